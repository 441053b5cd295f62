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

// The wavefront form of the same optimiser: the state is lane `lane`'s view into a tile of 32 interleaved states (bq3::StateTile) and
// advance() runs phase by phase - A (fresh objective value), T (trust-region step), B (shift / geometry step / Lagrange values) -
// handing the patch back (YIELD) in front of every heavy block that the current phase excludes, exactly as the per-phase kernels of
// hpmvs_b200/csrc/patch_kernels_wf.cuh call it.  Must visit the same points as the monolithic call above.
extern "C" int bq3_run_testfunc_tile(int func_id, const double* x0, const double* lb, const double* ub, double xtol_rel,
                                     int maxeval, double* xout, double* fout, double* trace_x, double* trace_f,
                                     int trace_cap, int* nevals, int lane, int split, int* nyields) {
    static_assert(sizeof(bq3::StateTile) % (8 * BQ_TILE_LANES) == 0, "tile = whole cells");
    unsigned char* tile = new unsigned char[sizeof(bq3::StateTile) + 8 * BQ_TILE_LANES];
    for (size_t i = 0; i < sizeof(bq3::StateTile) + 8 * BQ_TILE_LANES; i++) tile[i] = 0xA5;      // neighbours' lanes: must stay untouched
    bq3::StateTile& S = *reinterpret_cast<bq3::StateTile*>(tile + 8 * lane);
    double x[3];
    int n = 0, ny = 0;
    int act = bq3::start(S, x0, lb, ub, xtol_rel, maxeval, x);
    while (act == bq3::ASK) {
        const double f = orc_testfunc_eval(func_id, x);
        if (n < trace_cap) { trace_x[3 * n] = x[0]; trace_x[3 * n + 1] = x[1]; trace_x[3 * n + 2] = x[2]; trace_f[n] = f; }
        n++;
        if (!split) { act = bq3::advance<bq3::StateTile, bq3::PH_ALL, false>(S, f, x); continue; }
        act = bq3::advance<bq3::StateTile, bq3::PH_A, false>(S, f, x);
        int phase = 1;
        while (act == bq3::YIELD) {
            ny++;
            if (phase == 0) act = bq3::advance<bq3::StateTile, bq3::PH_A, false>(S, 0.0, x);
            else if (phase == 1) act = bq3::advance<bq3::StateTile, bq3::PH_T, false>(S, 0.0, x);
            else act = bq3::advance<bq3::StateTile, bq3::PH_B, false>(S, 0.0, x);
            phase = (phase + 1) % 3;
        }
    }
    g_last_rescues = S.n_rescue;
    bq3::result_x(S, xout);
    *fout = S.minf;
    *nevals = n;
    *nyields = ny;
    const int rc = S.rc;
    // every byte outside this lane's 8-byte column must be untouched
    int dirty = 0;
    for (size_t i = 0; i < sizeof(bq3::StateTile) + 8 * BQ_TILE_LANES; i++)
        if ((i % (8 * BQ_TILE_LANES)) / 8 != (size_t)lane && tile[i] != 0xA5) dirty++;
    delete[] tile;
    return dirty ? -1000 : rc;
}
extern "C" int bq3_tile_bytes() { return (int)sizeof(bq3::StateTile); }
