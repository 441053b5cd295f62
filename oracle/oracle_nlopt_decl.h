/* TEST INFRASTRUCTURE. Minimal C declarations for the part of the nlopt 2.4.2 C API that the
 * oracle calls, so that the oracle compiles without the reference tree being present
 * (the GPU box has no /root/reference; it only carries the prebuilt oracle/_ref/libnlopt_ref.a).
 * Each prototype restates the public API documented in
 * /root/reference/thirdLibs/nlopt-2.4.2/api/nlopt.h (create/destroy :194-195, optimize :198,
 * set_min_objective :201, bounds :214-221, xtol_rel :263, maxeval :270, result codes :160-170).
 * The reference reaches the same functions through the C++ wrapper nlopt.hpp
 * (src/hpmvs/PatchOptimizer.cpp:348-363). */
#ifndef ORACLE_NLOPT_DECL_H
#define ORACLE_NLOPT_DECL_H
#ifdef __cplusplus
extern "C" {
#endif

typedef double (*nlopt_func)(unsigned n, const double *x, double *gradient, void *func_data);
struct nlopt_opt_s;
typedef struct nlopt_opt_s *nlopt_opt;

/* nlopt_result values (nlopt.h:160-170) */
enum {
    ORC_NLOPT_FAILURE = -1, ORC_NLOPT_INVALID_ARGS = -2, ORC_NLOPT_OUT_OF_MEMORY = -3,
    ORC_NLOPT_ROUNDOFF_LIMITED = -4, ORC_NLOPT_FORCED_STOP = -5, ORC_NLOPT_SUCCESS = 1,
    ORC_NLOPT_STOPVAL_REACHED = 2, ORC_NLOPT_FTOL_REACHED = 3, ORC_NLOPT_XTOL_REACHED = 4,
    ORC_NLOPT_MAXEVAL_REACHED = 5, ORC_NLOPT_MAXTIME_REACHED = 6
};

/* algorithm ids are resolved at run time by name through nlopt_algorithm_name() */
const char *nlopt_algorithm_name(int a);
nlopt_opt nlopt_create(int algorithm, unsigned n);
void nlopt_destroy(nlopt_opt opt);
int nlopt_optimize(nlopt_opt opt, double *x, double *opt_f);
int nlopt_set_min_objective(nlopt_opt opt, nlopt_func f, void *f_data);
int nlopt_set_lower_bounds(nlopt_opt opt, const double *lb);
int nlopt_set_upper_bounds(nlopt_opt opt, const double *ub);
int nlopt_set_xtol_rel(nlopt_opt opt, double tol);
int nlopt_set_maxeval(nlopt_opt opt, int maxeval);

#ifdef __cplusplus
}
#endif
#endif
