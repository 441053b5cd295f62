// Test harness (tests/ only): runs the product's ask/tell BOBYQA (hpmvs_b200/csrc/bobyqa3.h, compiled for
// the host) on the oracle's analytic test objectives, so tests/test_bobyqa.py can compare the visited points
// bit-for-bit with the real nlopt (oracle/_ref) on the same objective code.
#include "../../hpmvs_b200/csrc/bobyqa3.h"
#include "../../oracle/hpmvs_oracle.h"

static int g_last_rescues = 0;

extern "C" int bq3_run_testfunc(int func_id, const double* x0, const double* lb, const double* ub, double xtol_rel,
                                int maxeval, double* xout, double* fout, double* trace_x, double* trace_f,
                                int trace_cap, int* nevals) {
    bq3::State S;
    double x[3];
    int n = 0;
    int act = bq3::start(S, x0, lb, ub, xtol_rel, maxeval, x);
    while (act == bq3::ASK) {
        const double f = orc_testfunc_eval(func_id, x);
        if (n < trace_cap) { trace_x[3 * n] = x[0]; trace_x[3 * n + 1] = x[1]; trace_x[3 * n + 2] = x[2]; trace_f[n] = f; }
        n++;
        act = bq3::advance(S, f, x);
    }
    g_last_rescues = S.n_rescue;
    bq3::result_x(S, xout);
    *fout = S.minf;
    *nevals = n;
    return S.rc;
}

extern "C" int bq3_last_rescues() { return g_last_rescues; }
extern "C" int bq3_state_bytes() { return (int)sizeof(bq3::State); }
