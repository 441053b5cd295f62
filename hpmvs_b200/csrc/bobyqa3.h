// BOBYQA (Powell 2009) specialised to n = 3 variables, npt = 2n+1 = 7 interpolation points, written as a
// resumable "ask / tell" state machine so that one GPU warp (or one lane) can own an optimisation and hand
// objective evaluations to warp-cooperative sampling code.
//
// Replaces, for the HPMVS hot path, the nlopt call chain
//   nlopt::opt(LN_BOBYQA,3).optimize()            /root/reference/src/hpmvs/PatchOptimizer.cpp:348-363
//   -> nlopt_optimize_ / default initial step      thirdLibs/nlopt-2.4.2/api/optimize.c:669-681, api/options.c:685-728
//   -> bobyqa() rescaling + bound preparation      thirdLibs/nlopt-2.4.2/bobyqa/bobyqa.c:3073-3268, util/rescale.c:29-82
//   -> bobyqb_/prelim_/trsbox_/altmov_/update_/rescue_   bobyqa.c:18-3055
// It is an independent implementation of the same algorithm with the same floating-point evaluation order,
// so that (with FMA contraction off) it visits bit-identical points; tests/test_bobyqa.py pins that against
// the real library on analytic objectives.  All state is FP64 and lives in the struct (about 1.6 KB).
//
// Conventions: zero-based indices everywhere; hq is the packed upper triangle (column by column);
// the scratch array w is partitioned exactly like the original workspace so that values which the algorithm
// reads back from scratch (e.g. the trust-region gradient in the bound test) are the same ones.
#pragma once

#if defined(__CUDACC__)
#define BQ_HD __host__ __device__ __forceinline__
#define BQ_HDN __host__ __device__ __noinline__
#else
#define BQ_HD inline
#define BQ_HDN inline
#endif

#include <math.h>
#include <type_traits>

// The loops below have tiny constant trip counts (3, 7, 10).  Fully unrolled, the state machine grows to >300 KB of
// SASS and the kernel becomes instruction-fetch bound (ncu: 65% "no_instructions" stalls); kept rolled it fits the
// SM's instruction cache.
#if defined(__CUDA_ARCH__)
#define BQ_NOUNROLL _Pragma("unroll 1")
// innermost 3-trip loops with one-statement bodies (and hess_mul) ARE unrolled: the per-label cycle profile
// (profiles/) showed the rolled multiply-accumulate loops costing ~10 instructions per MAC.
#define BQ_UNROLL _Pragma("unroll")
#define BQ_UNROLL4 _Pragma("unroll 4")
#else
#define BQ_NOUNROLL
#define BQ_UNROLL
#define BQ_UNROLL4
#endif

// Warp-converged dispatch of the label state machines below (device only).  Lanes of one warp run different
// optimisations and sit at different labels; executing `switch(label)` naively makes the warp run the UNION of
// all cases once per lane-group and per loop trip (ncu: ~42k warp instructions per round instead of ~5k).
// Instead every trip picks the smallest label (in flow order) present among the lanes and only those lanes
// execute it - lanes further along wait, so each expensive block runs about once per round for the whole warp.
#if defined(__CUDA_ARCH__)
#define BQ_SCHED_BEGIN(mask, done, key)                                   \
    {                                                                      \
        const int key__ = (done) ? 0x7fffffff : (key);                     \
        const int cur__ = __reduce_min_sync((mask), key__);                \
        if (cur__ == 0x7fffffff) break;                                    \
        if (key__ != cur__) continue;                                      \
    }
#define BQ_ACTIVE_MASK() __activemask()
#else
#define BQ_SCHED_BEGIN(mask, done, key) \
    if (done) break;
#define BQ_ACTIVE_MASK() 0xffffffffu
#endif

// The engine keeps every State in shared memory; telling the compiler lets it emit LDS/STS with immediate offsets
// instead of generic loads plus address arithmetic.
#if defined(__CUDA_ARCH__) && defined(BQ_STATE_IN_SHARED)
#define BQ_ASSUME_SHARED(S) do { if (!bq3::is_tile<typename std::remove_cv<typename std::remove_reference<decltype(S)>::type>::type>::value) __builtin_assume(__isShared(&(S))); else __builtin_assume(__isGlobal(&(S))); } while (0)
#else
#define BQ_ASSUME_SHARED(S)
#endif

namespace bq3 {

#if defined(__CUDACC__) && defined(HP_PROFILE)
__device__ unsigned long long g_prof_cycles[16];
__device__ unsigned long long g_prof_trips[16];
__device__ unsigned long long g_prof_lanes[16];
__device__ unsigned long long g_tprof_cycles[8];
__device__ unsigned long long g_tprof_trips[8];
__device__ unsigned long long g_tprof_lanes[8];
#endif

enum : int { N = 3, NPT = 7, NDIM = 10, NPTM = 3, NP = 4, NH = 6 };

// nlopt_result values the reference tests BQ_NOUNROLL for (api/nlopt.h:160-170)
enum Result : int {
    R_FAILURE = -1, R_INVALID_ARGS = -2, R_ROUNDOFF_LIMITED = -4,
    R_SUCCESS = 1, R_XTOL_REACHED = 4, R_MAXEVAL_REACHED = 5
};

enum Action : int { ASK = 0, DONE = 1, YIELD = 2 };   // YIELD (device, optional): call advance() again, no objective value needed

// BQ_DEFER_TRUST=<lanes>: a lane that needs a SECOND trust-region step inside one advance() call (after a rho reduction)
// gives the step back to the caller instead of making the whole warp wait for a trsbox trip with one or two lanes in it;
// it joins the first trip of the next round.  Only when at least <lanes> lanes entered together (busy rounds).
#ifndef BQ_DEFER_TRUST
#define BQ_DEFER_TRUST 0
#endif
#ifndef BQ_ALTSEARCH_UNROLL
#define BQ_ALTSEARCH_UNROLL 1
#endif
#define BQ_PRAGMA_(x) _Pragma(#x)
#define BQ_PRAGMA_UNROLL_N(n) BQ_PRAGMA_(unroll n)

BQ_HD double dmin(double a, double b) { return a <= b ? a : b; }
BQ_HD double dmax(double a, double b) { return a >= b ? a : b; }
BQ_HD int hidx(int i, int j) { return j * (j + 1) / 2 + i; }  // i <= j

// Strided cells for the wavefront kernels (patch_kernels_wf.cuh): the states of 32 patches are interleaved in one "tile" - every
// member of lane l's state is a cell of BQ_TILE_STRIDE bytes whose own value sits at byte offset 8*l, so that a warp whose lane l works
// on slot l of a tile reads / writes 32 consecutive 8-byte words per access (one fully coalesced 256-byte request) and every member keeps
// a compile-time offset from the lane's base pointer.  The cells behave like double / int in expressions; copies move the value only.
#ifndef BQ_TILE_LANES
#define BQ_TILE_LANES 32
#endif
struct alignas(8) SD {
    double v;
    double pad_[BQ_TILE_LANES - 1];
    BQ_HD operator double() const { return v; }
    BQ_HD SD& operator=(double x) { v = x; return *this; }
    BQ_HD SD& operator=(const SD& o) { v = o.v; return *this; }
    BQ_HD SD& operator+=(double x) { v += x; return *this; }
    BQ_HD SD& operator-=(double x) { v -= x; return *this; }
    BQ_HD SD& operator*=(double x) { v *= x; return *this; }
    BQ_HD SD& operator/=(double x) { v /= x; return *this; }
};
struct alignas(8) SI {
    int v;
    int pad_[2 * BQ_TILE_LANES - 1];
    BQ_HD operator int() const { return v; }
    BQ_HD SI& operator=(int x) { v = x; return *this; }
    BQ_HD SI& operator=(const SI& o) { v = o.v; return *this; }
    BQ_HD SI& operator+=(int x) { v += x; return *this; }
    BQ_HD SI& operator-=(int x) { v -= x; return *this; }
    BQ_HD SI& operator++() { ++v; return *this; }
    BQ_HD int operator++(int) { return v++; }
    BQ_HD SI& operator--() { --v; return *this; }
    BQ_HD int operator--(int) { return v--; }
};

template <class R, class I>
struct StateT {
    typedef R real;
    typedef I integer;
    // problem (scaled space)
    R xl[N], xu[N], scl[N];
    R rhobeg, rhoend;
    I maxeval, nevals;
    // model
    R xbase[N], xpt[NPT][N], fval[NPT], xopt[N], gopt[N], hq[NH], pq[NPT];
    R bmat[NDIM][N], zmat[NPT][NPTM], sl[N], su[N], xnew[N], xalt[N], d[N], vlag[NDIM];
    R w[3 * NDIM];
    R x[N];  // current point in scaled space (what prelim/bobyqb call X)
    // iteration scalars that live across evaluations
    R rho, delta, diffa, diffb, diffc, dsq, crvmin, dnorm, distsq, adelt, alpha, cauchy, beta, denom;
    R xoptsq, fsave, ratio, f, fbeg, stepa, stepb, vquad_r, fbase_r, minf;
    I ntrits, itest, nfsav, nresc, kopt, kbase, knew, nf, kpt, pc, rc, nrem_r;
    I in_rescue_from_main;
    I n_rescue;   // statistics only
};
typedef StateT<double, int> State;       // one patch, contiguous (host, shared-memory slots of the persistent kernel)
typedef StateT<SD, SI> StateTile;        // lane view into a tile of BQ_TILE_LANES interleaved states (wavefront kernels)
template <class ST> struct is_tile { static constexpr bool value = false; };
template <> struct is_tile<StateTile> { static constexpr bool value = true; };

// program counters (resume points)
enum : int {
    PC_PRELIM_EVAL = 1, PC_MAIN_EVAL = 2, PC_RESCUE_EVAL = 3, PC_FINISHED = 4, PC_YIELD_TRUST = 5,
    PC_YIELD_LABEL = 16        // + label: advance() stopped in front of a heavy label its PHASES mask excludes (wavefront kernels)
};

// labels of the main iteration, numbered in flow order (the warp scheduler runs the smallest one present)
enum : int {
    L_AFTER_EVAL = 0, L_RESCUE_LOOP = 1, L_RESCUE_DONE = 2, L_GOPT_FIX = 3, L_FARPOINT = 4, L_REDUCE_RHO = 5, L_TRUST = 6,
    L_SHIFT = 7, L_RESCUE = 8, L_ALTMOV = 9, L_VLAG = 10, L_EVAL = 11, L_EXIT = 12, L_NONE = 13
};
// The heavy label blocks.  advance<ST, PHASES> runs a heavy block only when its bit is set in PHASES and otherwise hands the patch
// back (YIELD, S.pc = PC_YIELD_LABEL + label); the light blocks (bookkeeping between them) run in every phase.  The wavefront engine
// launches one kernel per phase, so each kernel's instruction footprint is a fraction of the whole state machine.
enum : unsigned {
    PH_HEAVY = (1u << L_AFTER_EVAL) | (1u << L_RESCUE_LOOP) | (1u << L_TRUST) | (1u << L_SHIFT) | (1u << L_RESCUE) | (1u << L_ALTMOV) | (1u << L_VLAG),
    PH_A = (1u << L_AFTER_EVAL) | (1u << L_RESCUE_LOOP) | (1u << L_RESCUE),     // entered with a fresh objective value
    PH_T = (1u << L_TRUST),
    PH_B = (1u << L_SHIFT) | (1u << L_ALTMOV) | (1u << L_VLAG),
    PH_ALL = PH_A | PH_T | PH_B
};

// ---------------------------------------------------------------------------------------------------------
// default initial step (api/options.c:696-728) for one coordinate
// ---------------------------------------------------------------------------------------------------------
BQ_HD double default_step(double lb, double ub, double x) {
    double step = HUGE_VAL;
    const bool ubinf = isinf(ub), lbinf = isinf(lb);
    if (!ubinf && !lbinf && (ub - lb) * 0.25 < step && ub > lb) step = (ub - lb) * 0.25;
    if (!ubinf && ub - x < step && ub > x) step = (ub - x) * 0.75;
    if (!lbinf && x - lb < step && x > lb) step = (x - lb) * 0.75;
    if (isinf(step)) {
        if (!ubinf && fabs(ub - x) < fabs(step)) step = (ub - x) * 1.1;
        if (!lbinf && fabs(x - lb) < fabs(step)) step = (x - lb) * 1.1;
    }
    if (isinf(step) || step == 0) step = x;
    if (isinf(step) || step == 0) step = 1;
    return step;
}

// ---------------------------------------------------------------------------------------------------------
// H-matrix update when interpolation point `knew` moves (Powell's UPDATE)
// ---------------------------------------------------------------------------------------------------------
template <class ST>
BQ_HDN void update(ST& S, double beta, double denom, int knew, typename ST::real* w) {
    typedef typename ST::real R;
    BQ_ASSUME_SHARED(S);
    R (*zmat)[NPTM] = S.zmat;
    R (*bmat)[N] = S.bmat;
    R* vlag = S.vlag;
    double ztest = 0.0;
    BQ_NOUNROLL for (int k = 0; k < NPT; k++)
        BQ_UNROLL for (int j = 0; j < NPTM; j++) ztest = dmax(ztest, fabs(zmat[k][j]));
    ztest *= 1e-20;
    // rotations that zero the knew-th row of zmat beyond its first column
    BQ_NOUNROLL for (int j = 1; j < NPTM; j++) {
        if (fabs(zmat[knew][j]) > ztest) {
            const double a = zmat[knew][0], b = zmat[knew][j];
            double temp = sqrt(a * a + b * b);
            const double tempa = zmat[knew][0] / temp;
            const double tempb = zmat[knew][j] / temp;
            BQ_NOUNROLL for (int i = 0; i < NPT; i++) {
                temp = tempa * zmat[i][0] + tempb * zmat[i][j];
                zmat[i][j] = tempa * zmat[i][j] - tempb * zmat[i][0];
                zmat[i][0] = temp;
            }
        }
        zmat[knew][j] = 0.0;
    }
    BQ_NOUNROLL for (int i = 0; i < NPT; i++) w[i] = zmat[knew][0] * zmat[i][0];
    const double alpha = w[knew];
    const double tau = vlag[knew];
    vlag[knew] -= 1.0;
    {
        const double temp = sqrt(denom);
        const double tempb = zmat[knew][0] / temp;
        const double tempa = tau / temp;
        BQ_NOUNROLL for (int i = 0; i < NPT; i++) zmat[i][0] = tempa * zmat[i][0] - tempb * vlag[i];
    }
    BQ_NOUNROLL for (int j = 0; j < N; j++) {
        const int jp = NPT + j;
        w[jp] = bmat[knew][j];
        const double tempa = (alpha * vlag[jp] - tau * w[jp]) / denom;
        const double tempb = (-beta * w[jp] - tau * vlag[jp]) / denom;
        BQ_NOUNROLL for (int i = 0; i <= jp; i++) {
            bmat[i][j] = bmat[i][j] + tempa * vlag[i] + tempb * w[i];
            if (i >= NPT) bmat[jp][i - NPT] = bmat[i][j];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// ALTMOV: alternative positions for interpolation point knew (line search through xopt + Cauchy step)
// glag = w[0..2], hcol = w[3..9], wa = w[10..15]
// ---------------------------------------------------------------------------------------------------------
template <class ST>
BQ_HDN void altmov(ST& S) {
    typedef typename ST::real R;
    BQ_ASSUME_SHARED(S);
    R (*xpt)[N] = S.xpt;
    R (*zmat)[NPTM] = S.zmat;
    R (*bmat)[N] = S.bmat;
    R* xopt = S.xopt; R* sl = S.sl; R* su = S.su; R* xnew = S.xnew; R* xalt = S.xalt;
    R* glag = S.w; R* hcol = S.w + NP - 1; R* wa = S.w + NDIM;
    const int kopt = S.kopt, knew = S.knew;
    const double adelt = S.adelt;
    const double cnst = 1.0 + sqrt(2.0);
    BQ_NOUNROLL for (int k = 0; k < NPT; k++) hcol[k] = 0.0;
    BQ_NOUNROLL for (int j = 0; j < NPTM; j++) {
        const double temp = zmat[knew][j];
        BQ_NOUNROLL for (int k = 0; k < NPT; k++) hcol[k] += temp * zmat[k][j];
    }
    S.alpha = hcol[knew];
    const double ha = 0.5 * S.alpha;
    BQ_UNROLL for (int i = 0; i < N; i++) glag[i] = bmat[knew][i];
    BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
        double temp = 0.0;
        BQ_UNROLL for (int j = 0; j < N; j++) temp += xpt[k][j] * xopt[j];
        temp = hcol[k] * temp;
        BQ_UNROLL for (int i = 0; i < N; i++) glag[i] += temp * xpt[k][i];
    }
    // search along lines through xopt and the other points
    double presav = 0.0, stpsav = 0.0;
    int ksav = 0, ibdsav = 0;
    BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
        if (k == kopt) continue;
        double dderiv = 0.0, distsq = 0.0;
        BQ_NOUNROLL for (int i = 0; i < N; i++) {
            const double temp = xpt[k][i] - xopt[i];
            dderiv += glag[i] * temp;
            distsq += temp * temp;
        }
        double subd = adelt / sqrt(distsq);
        double slbd = -subd;
        int ilbd = 0, iubd = 0;
        const double sumin = dmin(1.0, subd);
        BQ_NOUNROLL for (int i = 0; i < N; i++) {
            const double temp = xpt[k][i] - xopt[i];
            if (temp > 0.0) {
                if (slbd * temp < sl[i] - xopt[i]) { slbd = (sl[i] - xopt[i]) / temp; ilbd = -(i + 1); }
                if (subd * temp > su[i] - xopt[i]) { subd = dmax(sumin, (su[i] - xopt[i]) / temp); iubd = i + 1; }
            } else if (temp < 0.0) {
                if (slbd * temp > su[i] - xopt[i]) { slbd = (su[i] - xopt[i]) / temp; ilbd = i + 1; }
                if (subd * temp < sl[i] - xopt[i]) { subd = dmax(sumin, (sl[i] - xopt[i]) / temp); iubd = -(i + 1); }
            }
        }
        double step, vlag;
        int isbd;
        if (k == knew) {
            const double diff = dderiv - 1.0;
            step = slbd;
            vlag = slbd * (dderiv - slbd * diff);
            isbd = ilbd;
            double temp = subd * (dderiv - subd * diff);
            if (fabs(temp) > fabs(vlag)) { step = subd; vlag = temp; isbd = iubd; }
            const double tempd = 0.5 * dderiv;
            const double tempa = tempd - diff * slbd;
            const double tempb = tempd - diff * subd;
            if (tempa * tempb < 0.0) {
                temp = tempd * tempd / diff;
                if (fabs(temp) > fabs(vlag)) { step = tempd / diff; vlag = temp; isbd = 0; }
            }
        } else {
            step = slbd;
            vlag = slbd * (1.0 - slbd);
            isbd = ilbd;
            const double temp = subd * (1.0 - subd);
            if (fabs(temp) > fabs(vlag)) { step = subd; vlag = temp; isbd = iubd; }
            if (subd > 0.5) {
                if (fabs(vlag) < 0.25) { step = 0.5; vlag = 0.25; isbd = 0; }
            }
            vlag *= dderiv;
        }
        const double temp = step * (1.0 - step) * distsq;
        const double predsq = vlag * vlag * (vlag * vlag + ha * temp * temp);
        if (predsq > presav) { presav = predsq; ksav = k; stpsav = step; ibdsav = isbd; }
    }
    BQ_NOUNROLL for (int i = 0; i < N; i++) {
        const double temp = xopt[i] + stpsav * (xpt[ksav][i] - xopt[i]);
        xnew[i] = dmax(sl[i], dmin(su[i], temp));
    }
    if (ibdsav < 0) xnew[-ibdsav - 1] = sl[-ibdsav - 1];
    if (ibdsav > 0) xnew[ibdsav - 1] = su[ibdsav - 1];

    // constrained Cauchy step, tried for both signs of glag
    const double bigstp = adelt + adelt;
    double csave = 0.0, step = 0.0;
    BQ_NOUNROLL for (int iflag = 0; iflag < 2; iflag++) {
        double wfixsq = 0.0, ggfree = 0.0;
        BQ_NOUNROLL for (int i = 0; i < N; i++) {
            wa[i] = 0.0;
            const double tempa = dmin(xopt[i] - sl[i], glag[i]);
            const double tempb = dmax(xopt[i] - su[i], glag[i]);
            if (tempa > 0.0 || tempb < 0.0) { wa[i] = bigstp; ggfree += glag[i] * glag[i]; }
        }
        if (ggfree == 0.0) { S.cauchy = 0.0; return; }
        for (;;) {
            const double temp = adelt * adelt - wfixsq;
            if (!(temp > 0.0)) break;
            const double wsqsav = wfixsq;
            step = sqrt(temp / ggfree);
            ggfree = 0.0;
            BQ_NOUNROLL for (int i = 0; i < N; i++) {
                if (wa[i] == bigstp) {
                    const double t = xopt[i] - step * glag[i];
                    if (t <= sl[i]) { wa[i] = sl[i] - xopt[i]; wfixsq += wa[i] * wa[i]; }
                    else if (t >= su[i]) { wa[i] = su[i] - xopt[i]; wfixsq += wa[i] * wa[i]; }
                    else ggfree += glag[i] * glag[i];
                }
            }
            if (!(wfixsq > wsqsav && ggfree > 0.0)) break;
        }
        double gw = 0.0;
        BQ_NOUNROLL for (int i = 0; i < N; i++) {
            if (wa[i] == bigstp) {
                wa[i] = -step * glag[i];
                xalt[i] = dmax(sl[i], dmin(su[i], xopt[i] + wa[i]));
            } else if (wa[i] == 0.0) xalt[i] = xopt[i];
            else if (glag[i] > 0.0) xalt[i] = sl[i];
            else xalt[i] = su[i];
            gw += glag[i] * wa[i];
        }
        double curv = 0.0;
        BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
            double temp = 0.0;
            BQ_UNROLL for (int j = 0; j < N; j++) temp += xpt[k][j] * wa[j];
            curv += hcol[k] * temp * temp;
        }
        if (iflag == 1) curv = -curv;
        if (curv > -gw && curv < -cnst * gw) {
            const double scale = -gw / curv;
            BQ_NOUNROLL for (int i = 0; i < N; i++) {
                const double temp = xopt[i] + scale * wa[i];
                xalt[i] = dmax(sl[i], dmin(su[i], temp));
            }
            const double t = 0.5 * gw * scale;
            S.cauchy = t * t;
        } else {
            const double t = gw + 0.5 * curv;
            S.cauchy = t * t;
        }
        if (iflag == 0) {
            BQ_UNROLL for (int i = 0; i < N; i++) { glag[i] = -glag[i]; wa[N + i] = xalt[i]; }
            csave = S.cauchy;
        }
    }
    if (csave > S.cauchy) {
        BQ_UNROLL for (int i = 0; i < N; i++) xalt[i] = wa[N + i];
        S.cauchy = csave;
    }
}

// ---------------------------------------------------------------------------------------------------------
// TRSBOX: truncated conjugate gradients inside the trust region with simple bounds, followed by the
// boundary (two-dimensional) refinements.  gnew=w[0..2] xbdi=w[3..5] s=w[6..8] hs=w[9..11] hred=w[12..14]
// ---------------------------------------------------------------------------------------------------------
template <class ST>
BQ_HD void hess_mul(const ST& S, const double* s, double* hs) {
    BQ_ASSUME_SHARED(S);
    int ih = 0;
    BQ_UNROLL for (int j = 0; j < N; j++) {
        hs[j] = 0.0;
        BQ_UNROLL for (int i = 0; i <= j; i++) {
            if (i < j) hs[j] += S.hq[ih] * s[i];
            hs[i] += S.hq[ih] * s[j];
            ih++;
        }
    }
    BQ_UNROLL for (int k = 0; k < NPT; k++) {
        if (S.pq[k] != 0.0) {
            double temp = 0.0;
            BQ_UNROLL for (int j = 0; j < N; j++) temp += S.xpt[k][j] * s[j];
            temp *= S.pq[k];
            BQ_UNROLL for (int i = 0; i < N; i++) hs[i] += temp * S.xpt[k][i];
        }
    }
}

template <class ST>
BQ_HDN void trsbox(ST& S, unsigned wmask) {
    BQ_ASSUME_SHARED(S);
    // All N-vectors of this routine live in registers (every loop over N below is fully unrolled so that the
    // indices are static); shared memory is only read (xpt, hq, pq) until the results are stored at T_FINISH.
    double xopt[N], gopt[N], sl[N], su[N], xnew[N], d[N], gnew[N], xbdi[N], s[N], hs[N], hred[N];
    BQ_UNROLL for (int i = 0; i < N; i++) { xopt[i] = S.xopt[i]; gopt[i] = S.gopt[i]; sl[i] = S.sl[i]; su[i] = S.su[i]; s[i] = 0.0; hs[i] = 0.0; hred[i] = 0.0; }
    const double delta = S.delta;
    int iterc = 0, nact = 0, itermax = 0, itcsav = 0, iact = 0;
    double gredsq = 0, ggsav = 0, dredsq = 0, dredg = 0, sredg = 0, angbd = 0, xsav = 0, beta = 0, stepsq = 0;
    BQ_UNROLL for (int i = 0; i < N; i++) {
        xbdi[i] = 0.0;
        if (xopt[i] <= sl[i]) { if (gopt[i] >= 0.0) xbdi[i] = -1.0; }
        else if (xopt[i] >= su[i]) { if (gopt[i] <= 0.0) xbdi[i] = 1.0; }
        if (xbdi[i] != 0.0) ++nact;
        d[i] = 0.0;
        gnew[i] = gopt[i];
    }
    double delsq = delta * delta;
    double qred = 0.0;
    double crvmin = -1.0;

    // labels are numbered in flow order: the scheduler always runs the smallest one present in the warp
    enum { T_RESTART, T_DIRECTION, T_CGSTEP, T_BOUNDARY, T_ALT_PREP, T_ALT_DIR, T_ALT_SEARCH, T_FINISH };
    int lbl = T_RESTART;
    bool fin = false;
    for (;;) {
        BQ_SCHED_BEGIN(wmask, fin, lbl)
#if defined(__CUDA_ARCH__) && defined(HP_PROFILE)
        const long long tp0 = clock64();
        const int tp_lbl = lbl;
#endif
        switch (lbl) {
        case T_RESTART:
            beta = 0.0;
            lbl = T_DIRECTION;
            break;
        case T_DIRECTION: {
            stepsq = 0.0;
            BQ_UNROLL for (int i = 0; i < N; i++) {
                if (xbdi[i] != 0.0) s[i] = 0.0;
                else if (beta == 0.0) s[i] = -gnew[i];
                else s[i] = beta * s[i] - gnew[i];
                stepsq += s[i] * s[i];
            }
            if (stepsq == 0.0) { lbl = T_FINISH; break; }
            if (beta == 0.0) { gredsq = stepsq; itermax = iterc + N - nact; }
            if (gredsq * delsq <= qred * 1e-4 * qred) { lbl = T_FINISH; break; }
            hess_mul(S, s, hs);
            lbl = T_CGSTEP;
            break;
        }
        case T_CGSTEP: {
            double resid = delsq, ds = 0.0, shs = 0.0;
            BQ_UNROLL for (int i = 0; i < N; i++)
                if (xbdi[i] == 0.0) { resid -= d[i] * d[i]; ds += s[i] * d[i]; shs += s[i] * hs[i]; }
            if (resid <= 0.0) { lbl = T_BOUNDARY; break; }
            double temp = sqrt(stepsq * resid + ds * ds);
            double blen;
            if (ds < 0.0) blen = (temp - ds) / stepsq;
            else blen = resid / (temp + ds);
            double stplen = blen;
            if (shs > 0.0) stplen = dmin(blen, gredsq / shs);
            iact = 0;
            BQ_UNROLL for (int i = 0; i < N; i++) {
                if (s[i] != 0.0) {
                    const double xsum = xopt[i] + d[i];
                    if (s[i] > 0.0) temp = (su[i] - xsum) / s[i];
                    else temp = (sl[i] - xsum) / s[i];
                    if (temp < stplen) { stplen = temp; iact = i + 1; }
                }
            }
            double sdec = 0.0;
            if (stplen > 0.0) {
                ++iterc;
                temp = shs / stepsq;
                if (iact == 0 && temp > 0.0) {
                    crvmin = dmin(crvmin, temp);
                    if (crvmin == -1.0) crvmin = temp;
                }
                ggsav = gredsq;
                gredsq = 0.0;
                BQ_UNROLL for (int i = 0; i < N; i++) {
                    gnew[i] += stplen * hs[i];
                    if (xbdi[i] == 0.0) gredsq += gnew[i] * gnew[i];
                    d[i] += stplen * s[i];
                }
                sdec = dmax(stplen * (ggsav - 0.5 * stplen * shs), 0.0);
                qred += sdec;
            }
            if (iact > 0) {
                ++nact;
                BQ_UNROLL for (int i = 0; i < N; i++)
                    if (i == iact - 1) {
                        xbdi[i] = 1.0;
                        if (s[i] < 0.0) xbdi[i] = -1.0;
                        delsq -= d[i] * d[i];
                    }
                if (delsq <= 0.0) { lbl = T_BOUNDARY; break; }
                lbl = T_RESTART;
                break;
            }
            if (stplen < blen) {
                if (iterc == itermax) { lbl = T_FINISH; break; }
                if (sdec <= qred * 0.01) { lbl = T_FINISH; break; }
                beta = gredsq / ggsav;
                lbl = T_DIRECTION;
                break;
            }
            lbl = T_BOUNDARY;
            break;
        }
        case T_BOUNDARY:
            crvmin = 0.0;
            lbl = T_ALT_PREP;
            break;
        case T_ALT_PREP: {
            if (nact >= N - 1) { lbl = T_FINISH; break; }
            dredsq = 0.0; dredg = 0.0; gredsq = 0.0;
            BQ_UNROLL for (int i = 0; i < N; i++) {
                if (xbdi[i] == 0.0) {
                    dredsq += d[i] * d[i];
                    dredg += d[i] * gnew[i];
                    gredsq += gnew[i] * gnew[i];
                    s[i] = d[i];
                } else s[i] = 0.0;
            }
            itcsav = iterc;
            hess_mul(S, s, hs);
            BQ_UNROLL for (int i = 0; i < N; i++) hred[i] = hs[i];
            lbl = T_ALT_DIR;
            break;
        }
        case T_ALT_DIR: {
            ++iterc;
            double temp = gredsq * dredsq - dredg * dredg;
            if (temp <= qred * 1e-4 * qred) { lbl = T_FINISH; break; }
            temp = sqrt(temp);
            BQ_UNROLL for (int i = 0; i < N; i++) {
                if (xbdi[i] == 0.0) s[i] = (dredg * d[i] - dredsq * gnew[i]) / temp;
                else s[i] = 0.0;
            }
            sredg = -temp;
            angbd = 1.0;
            iact = 0;
            bool refix = false;
            BQ_UNROLL for (int i = 0; i < N; i++) {
                if (xbdi[i] == 0.0) {
                    const double tempa = xopt[i] + d[i] - sl[i];
                    const double tempb = su[i] - xopt[i] - d[i];
                    if (tempa <= 0.0) { ++nact; xbdi[i] = -1.0; refix = true; break; }
                    else if (tempb <= 0.0) { ++nact; xbdi[i] = 1.0; refix = true; break; }
                    const double ssq = d[i] * d[i] + s[i] * s[i];
                    double t = xopt[i] - sl[i];
                    temp = ssq - t * t;
                    if (temp > 0.0) {
                        temp = sqrt(temp) - s[i];
                        if (angbd * temp > tempa) { angbd = tempa / temp; iact = i + 1; xsav = -1.0; }
                    }
                    t = su[i] - xopt[i];
                    temp = ssq - t * t;
                    if (temp > 0.0) {
                        temp = sqrt(temp) + s[i];
                        if (angbd * temp > tempb) { angbd = tempb / temp; iact = i + 1; xsav = 1.0; }
                    }
                }
            }
            if (refix) { lbl = T_ALT_PREP; break; }
            hess_mul(S, s, hs);
            lbl = T_ALT_SEARCH;
            break;
        }
        case T_ALT_SEARCH: {
            double shs = 0.0, dhs = 0.0, dhd = 0.0;
            BQ_UNROLL for (int i = 0; i < N; i++)
                if (xbdi[i] == 0.0) { shs += s[i] * hs[i]; dhs += d[i] * hs[i]; dhd += d[i] * hred[i]; }
            double redmax = 0.0, redsav = 0.0, rdprev = 0.0, rdnext = 0.0;
            int isav = 0;
            const int iu = (int)(angbd * 17. + 3.1);
#if defined(__CUDA_ARCH__)
            BQ_PRAGMA_UNROLL_N(BQ_ALTSEARCH_UNROLL)
#endif
            for (int i = 1; i <= iu; i++) {
                const double angt = angbd * (double)i / (double)iu;
                const double sth = (angt + angt) / (1.0 + angt * angt);
                const double temp = shs + angt * (angt * dhd - dhs - dhs);
                const double rednew = sth * (angt * dredg - sredg - 0.5 * sth * temp);
                if (rednew > redmax) { redmax = rednew; isav = i; rdprev = redsav; }
                else if (i == isav + 1) rdnext = rednew;
                redsav = rednew;
            }
            if (isav == 0) { lbl = T_FINISH; break; }
            double angt = angbd;  // value at i == iu
            if (isav < iu) {
                const double temp = (rdnext - rdprev) / (redmax + redmax - rdprev - rdnext);
                angt = angbd * ((double)isav + 0.5 * temp) / (double)iu;
            } else {
                angt = angbd * (double)iu / (double)iu;
            }
            const double cth = (1.0 - angt * angt) / (1.0 + angt * angt);
            const double sth = (angt + angt) / (1.0 + angt * angt);
            const double temp = shs + angt * (angt * dhd - dhs - dhs);
            const double sdec = sth * (angt * dredg - sredg - 0.5 * sth * temp);
            if (sdec <= 0.0) { lbl = T_FINISH; break; }
            dredg = 0.0; gredsq = 0.0;
            BQ_UNROLL for (int i = 0; i < N; i++) {
                gnew[i] = gnew[i] + (cth - 1.0) * hred[i] + sth * hs[i];
                if (xbdi[i] == 0.0) {
                    d[i] = cth * d[i] + sth * s[i];
                    dredg += d[i] * gnew[i];
                    gredsq += gnew[i] * gnew[i];
                }
                hred[i] = cth * hred[i] + sth * hs[i];
            }
            qred += sdec;
            if (iact > 0 && isav == iu) {
                ++nact;
                BQ_UNROLL for (int i = 0; i < N; i++) if (i == iact - 1) xbdi[i] = xsav;
                lbl = T_ALT_PREP;
                break;
            }
            if (sdec > qred * 0.01) { lbl = T_ALT_DIR; break; }
            lbl = T_FINISH;
            break;
        }
        case T_FINISH: {
            double dsq = 0.0;
            BQ_UNROLL for (int i = 0; i < N; i++) {
                xnew[i] = dmax(dmin(xopt[i] + d[i], su[i]), sl[i]);
                if (xbdi[i] == -1.0) xnew[i] = sl[i];
                if (xbdi[i] == 1.0) xnew[i] = su[i];
                d[i] = xnew[i] - xopt[i];
                dsq += d[i] * d[i];
            }
            S.dsq = dsq;
            S.crvmin = crvmin;
            BQ_UNROLL for (int i = 0; i < N; i++) {
                S.xnew[i] = xnew[i]; S.d[i] = d[i];
                S.w[i] = gnew[i]; S.w[3 + i] = xbdi[i]; S.w[6 + i] = s[i]; S.w[9 + i] = hs[i]; S.w[12 + i] = hred[i];
            }
            fin = true;
            break;
        }
        }
#if defined(__CUDA_ARCH__) && defined(HP_PROFILE)
        {
            const unsigned now = __activemask();
            unsigned lane_id; asm("mov.u32 %0, %%laneid;" : "=r"(lane_id));
            if (lane_id == (unsigned)(__ffs(now) - 1)) {
                atomicAdd(&g_tprof_cycles[tp_lbl], (unsigned long long)(clock64() - tp0));
                atomicAdd(&g_tprof_trips[tp_lbl], 1ull);
                atomicAdd(&g_tprof_lanes[tp_lbl], (unsigned long long)__popc(now));
            }
        }
#endif
    }
}

// point handed to the objective: x = clamp(xbase + p) with exact bounds where p sits on sl/su
template <class ST>
BQ_HD void point_from(ST& S, const typename ST::real* p) {
    BQ_ASSUME_SHARED(S);
    BQ_NOUNROLL for (int i = 0; i < N; i++) {
        S.x[i] = dmin(dmax(S.xl[i], S.xbase[i] + p[i]), S.xu[i]);
        if (p[i] == S.sl[i]) S.x[i] = S.xl[i];
        if (p[i] == S.su[i]) S.x[i] = S.xu[i];
    }
}

// ---------------------------------------------------------------------------------------------------------
// RESCUE, part 1: rebuild bmat/zmat around xopt with provisional points (no objective evaluations yet).
// ptsaux = w[0..5] (ptsaux[j][0|1] -> w[2j], w[2j+1]), ptsid = w[6..12], scratch wr = w[13..29]
// ---------------------------------------------------------------------------------------------------------
template <class ST>
BQ_HDN void rescue_setup(ST& S) {
    typedef typename ST::real R;
    BQ_ASSUME_SHARED(S);
    R (*xpt)[N] = S.xpt; R (*bmat)[N] = S.bmat; R (*zmat)[NPTM] = S.zmat;
    R* xopt = S.xopt; R* sl = S.sl; R* su = S.su; R* hq = S.hq; R* pq = S.pq;
    R* vlag = S.vlag;
    R* ptsaux = S.w; R* ptsid = S.w + 6; R* wr = S.w + 13;
    const double delta = S.delta;
    const double sfrac = 0.5 / (double)NP;
    double sumpq = 0.0, winc = 0.0;
    BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
        double distsq = 0.0;
        BQ_UNROLL for (int j = 0; j < N; j++) { xpt[k][j] -= xopt[j]; distsq += xpt[k][j] * xpt[k][j]; }
        sumpq += pq[k];
        wr[NDIM + k] = distsq;
        winc = dmax(winc, distsq);
        BQ_UNROLL for (int j = 0; j < NPTM; j++) zmat[k][j] = 0.0;
    }
    {
        int ih = 0;
        BQ_NOUNROLL for (int j = 0; j < N; j++) {
            wr[j] = 0.5 * sumpq * xopt[j];
            BQ_NOUNROLL for (int k = 0; k < NPT; k++) wr[j] += pq[k] * xpt[k][j];
            BQ_NOUNROLL for (int i = 0; i <= j; i++) { hq[ih] = hq[ih] + wr[i] * xopt[j] + wr[j] * xopt[i]; ih++; }
        }
    }
    BQ_NOUNROLL for (int j = 0; j < N; j++) {
        S.xbase[j] += xopt[j];
        sl[j] -= xopt[j];
        su[j] -= xopt[j];
        xopt[j] = 0.0;
        ptsaux[2 * j] = dmin(delta, su[j]);
        ptsaux[2 * j + 1] = dmax(-delta, sl[j]);
        if (ptsaux[2 * j] + ptsaux[2 * j + 1] < 0.0) {
            const double t = ptsaux[2 * j]; ptsaux[2 * j] = ptsaux[2 * j + 1]; ptsaux[2 * j + 1] = t;
        }
        if (fabs(ptsaux[2 * j + 1]) < 0.5 * fabs(ptsaux[2 * j])) ptsaux[2 * j + 1] = 0.5 * ptsaux[2 * j];
        BQ_NOUNROLL for (int i = 0; i < NDIM; i++) bmat[i][j] = 0.0;
    }
    S.fbase_r = S.fval[S.kopt];
    ptsid[0] = sfrac;
    BQ_NOUNROLL for (int j = 0; j < N; j++) {
        const int jp = j + 1, jpn = jp + N;   // zero-based rows of the +/- points along e_j
        ptsid[jp] = (double)(j + 1) + sfrac;
        // npt = 2n+1 so jpn < NPT always
        ptsid[jpn] = (double)(j + 1) / (double)NP + sfrac;
        const double temp = 1.0 / (ptsaux[2 * j] - ptsaux[2 * j + 1]);
        bmat[jp][j] = -temp + 1.0 / ptsaux[2 * j];
        bmat[jpn][j] = temp + 1.0 / ptsaux[2 * j + 1];
        bmat[0][j] = -bmat[jp][j] - bmat[jpn][j];
        zmat[0][j] = sqrt(2.) / fabs(ptsaux[2 * j] * ptsaux[2 * j + 1]);
        zmat[jp][j] = zmat[0][j] * ptsaux[2 * j + 1] * temp;
        zmat[jpn][j] = -zmat[0][j] * ptsaux[2 * j] * temp;
    }
    int nrem = NPT, kold = 0, knew = S.kopt;
    double beta = 0.0, denom = 0.0;
    bool swap_phase = true;
    for (;;) {
        if (swap_phase) {
            BQ_UNROLL for (int j = 0; j < N; j++) { const double t = bmat[kold][j]; bmat[kold][j] = bmat[knew][j]; bmat[knew][j] = t; }
            BQ_UNROLL for (int j = 0; j < NPTM; j++) { const double t = zmat[kold][j]; zmat[kold][j] = zmat[knew][j]; zmat[knew][j] = t; }
            ptsid[kold] = ptsid[knew];
            ptsid[knew] = 0.0;
            wr[NDIM + knew] = 0.0;
            --nrem;
            if (knew != S.kopt) {
                const double t = vlag[kold]; vlag[kold] = vlag[knew]; vlag[knew] = t;
                update(S, beta, denom, knew, wr);   // scratch = rescue's own workspace, as in the original call
                if (nrem == 0) { S.nrem_r = 0; return; }
                BQ_NOUNROLL for (int k = 0; k < NPT; k++) wr[NDIM + k] = fabs(wr[NDIM + k]);
            }
        }
        // pick the original point to reinstate next
        double dsqmin = 0.0;
        BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
            if (wr[NDIM + k] > 0.0) {
                if (dsqmin == 0.0 || wr[NDIM + k] < dsqmin) { knew = k; dsqmin = wr[NDIM + k]; }
            }
        }
        if (dsqmin == 0.0) break;
        BQ_UNROLL for (int j = 0; j < N; j++) wr[NPT + j] = xpt[knew][j];
        BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
            double sum = 0.0;
            if (k == S.kopt) {
            } else if (ptsid[k] == 0.0) {
                BQ_UNROLL for (int j = 0; j < N; j++) sum += wr[NPT + j] * xpt[k][j];
            } else {
                const int ip = (int)ptsid[k];
                if (ip > 0) sum = wr[NPT + ip - 1] * ptsaux[2 * (ip - 1)];
                const int iq = (int)((double)NP * ptsid[k] - (double)(ip * NP));
                if (iq > 0) {
                    int iw = 0;
                    if (ip == 0) iw = 1;
                    sum += wr[NPT + iq - 1] * ptsaux[2 * (iq - 1) + iw];
                }
            }
            wr[k] = 0.5 * sum * sum;
        }
        BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
            double sum = 0.0;
            BQ_UNROLL for (int j = 0; j < N; j++) sum += bmat[k][j] * wr[NPT + j];
            vlag[k] = sum;
        }
        beta = 0.0;
        BQ_NOUNROLL for (int j = 0; j < NPTM; j++) {
            double sum = 0.0;
            BQ_NOUNROLL for (int k = 0; k < NPT; k++) sum += zmat[k][j] * wr[k];
            beta -= sum * sum;
            BQ_NOUNROLL for (int k = 0; k < NPT; k++) vlag[k] += sum * zmat[k][j];
        }
        double bsum = 0.0, distsq = 0.0;
        BQ_NOUNROLL for (int j = 0; j < N; j++) {
            double sum = 0.0;
            BQ_NOUNROLL for (int k = 0; k < NPT; k++) sum += bmat[k][j] * wr[k];
            const int jp = j + NPT;
            bsum += sum * wr[jp];
            BQ_NOUNROLL for (int ip = NPT; ip < NDIM; ip++) sum += bmat[ip][j] * wr[ip];
            bsum += sum * wr[jp];
            vlag[jp] = sum;
            distsq += xpt[knew][j] * xpt[knew][j];
        }
        beta = 0.5 * distsq * distsq + beta - bsum;
        vlag[S.kopt] += 1.0;
        denom = 0.0;
        double vlmxsq = 0.0;
        BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
            if (ptsid[k] != 0.0) {
                double hdiag = 0.0;
                BQ_UNROLL for (int j = 0; j < NPTM; j++) hdiag += zmat[k][j] * zmat[k][j];
                const double den = beta * hdiag + vlag[k] * vlag[k];
                if (den > denom) { kold = k; denom = den; }
            }
            vlmxsq = dmax(vlmxsq, vlag[k] * vlag[k]);
        }
        if (denom <= vlmxsq * .01) {
            wr[NDIM + knew] = -wr[NDIM + knew] - winc;
            swap_phase = false;
            continue;
        }
        swap_phase = true;
    }
    S.nrem_r = nrem;
}

// RESCUE, part 2a: place provisional point kpt, predict the model there, emit the point to evaluate.
template <class ST>
BQ_HDN void rescue_place(ST& S, int kpt) {
    typedef typename ST::real R;
    BQ_ASSUME_SHARED(S);
    R (*xpt)[N] = S.xpt;
    R* hq = S.hq; R* pq = S.pq; R* gopt = S.gopt;
    R* ptsaux = S.w; R* ptsid = S.w + 6; R* wr = S.w + 13;
    int ih = 0;
    BQ_NOUNROLL for (int j = 0; j < N; j++) {
        wr[j] = xpt[kpt][j];
        xpt[kpt][j] = 0.0;
        const double temp = pq[kpt] * wr[j];
        BQ_NOUNROLL for (int i = 0; i <= j; i++) { hq[ih] += temp * wr[i]; ih++; }
    }
    pq[kpt] = 0.0;
    const int ip = (int)ptsid[kpt];
    const int iq = (int)((double)NP * ptsid[kpt] - (double)(ip * NP));
    double xp = 0.0, xq = 0.0;
    if (ip > 0) { xp = ptsaux[2 * (ip - 1)]; xpt[kpt][ip - 1] = xp; }
    if (iq > 0) {
        xq = ptsaux[2 * (iq - 1)];
        if (ip == 0) xq = ptsaux[2 * (iq - 1) + 1];
        xpt[kpt][iq - 1] = xq;
    }
    double vquad = S.fbase_r;
    int ihp = 0, ihq = 0;
    if (ip > 0) {
        ihp = (ip + ip * ip) / 2;
        vquad += xp * (gopt[ip - 1] + 0.5 * xp * hq[ihp - 1]);
    }
    if (iq > 0) {
        ihq = (iq + iq * iq) / 2;
        vquad += xq * (gopt[iq - 1] + 0.5 * xq * hq[ihq - 1]);
        if (ip > 0) {
            const int diff = ip - iq;
            const int iw = (ihp >= ihq ? ihp : ihq) - (diff < 0 ? -diff : diff);
            vquad += xp * xq * hq[iw - 1];
        }
    }
    BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
        double temp = 0.0;
        if (ip > 0) temp += xp * xpt[k][ip - 1];
        if (iq > 0) temp += xq * xpt[k][iq - 1];
        vquad += 0.5 * pq[k] * temp * temp;
    }
    S.vquad_r = vquad;
    // rescue evaluates at w[0..2] of ITS scratch; same clamp rule as elsewhere
    point_from(S, xpt[kpt]);
}

// RESCUE, part 2b: absorb f at provisional point kpt into the model.
template <class ST>
BQ_HDN void rescue_absorb(ST& S, int kpt, double f) {
    typedef typename ST::real R;
    BQ_ASSUME_SHARED(S);
    R (*bmat)[N] = S.bmat; R (*zmat)[NPTM] = S.zmat;
    R* hq = S.hq; R* pq = S.pq; R* gopt = S.gopt;
    R* ptsaux = S.w; R* ptsid = S.w + 6;
    const double diff = f - S.vquad_r;
    BQ_UNROLL for (int i = 0; i < N; i++) gopt[i] += diff * bmat[kpt][i];
    BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
        double sum = 0.0;
        BQ_UNROLL for (int j = 0; j < NPTM; j++) sum += zmat[k][j] * zmat[kpt][j];
        const double temp = diff * sum;
        if (ptsid[k] == 0.0) pq[k] += temp;
        else {
            const int ip = (int)ptsid[k];
            const int iq = (int)((double)NP * ptsid[k] - (double)(ip * NP));
            const int ihq = (iq * iq + iq) / 2;
            if (ip == 0) {
                const double t = ptsaux[2 * (iq - 1) + 1];
                hq[ihq - 1] += temp * (t * t);
            } else {
                const int ihp = (ip * ip + ip) / 2;
                const double t = ptsaux[2 * (ip - 1)];
                hq[ihp - 1] += temp * (t * t);
                if (iq > 0) {
                    const double u = ptsaux[2 * (iq - 1)];
                    hq[ihq - 1] += temp * (u * u);
                    const int dd = iq - ip;
                    const int iw = (ihp >= ihq ? ihp : ihq) - (dd < 0 ? -dd : dd);
                    hq[iw - 1] += temp * ptsaux[2 * (ip - 1)] * ptsaux[2 * (iq - 1)];
                }
            }
        }
    }
    ptsid[kpt] = 0.0;
}

// ---------------------------------------------------------------------------------------------------------
// start: nlopt_optimize_ glue + bobyqa() preparation.  x0/lb/ub are in the caller's (unscaled) space.
// Returns ASK with the first point in xs_out (unscaled), or DONE with S.rc set on invalid arguments.
// ---------------------------------------------------------------------------------------------------------

template <class ST>
BQ_HDN int start(ST& S, const double* x0, const double* lb, const double* ub, double xtol_rel, int maxeval,
                 double* xs_out) {
    BQ_ASSUME_SHARED(S);
    S.maxeval = maxeval;
    S.nevals = 0;
    S.rc = R_SUCCESS;
    S.minf = HUGE_VAL;
    S.in_rescue_from_main = false;
    S.n_rescue = 0;
    BQ_NOUNROLL for (int i = 0; i < N; i++) {
        if (lb[i] > ub[i] || x0[i] < lb[i] || x0[i] > ub[i]) { S.rc = R_INVALID_ARGS; S.pc = PC_FINISHED; return DONE; }
    }
    double dx[N];
    BQ_UNROLL for (int i = 0; i < N; i++) dx[i] = default_step(lb[i], ub[i], x0[i]);
    // rescale so that all initial steps equal dx[0]  (util/rescale.c:29-44)
    BQ_UNROLL for (int i = 0; i < N; i++) S.scl[i] = 1.0;
    {
        int i = 1;
        BQ_NOUNROLL for (; i < N && dx[i] == dx[i - 1]; ++i) ;
        if (i < N) BQ_NOUNROLL for (i = 1; i < N; ++i) S.scl[i] = dx[i] / dx[0];
    }
    BQ_NOUNROLL for (int i = 0; i < N; i++) {
        S.x[i] = x0[i] / S.scl[i];
        S.xl[i] = lb[i] / S.scl[i];
        S.xu[i] = ub[i] / S.scl[i];
        if (S.xl[i] > S.xu[i]) { const double t = S.xl[i]; S.xl[i] = S.xu[i]; S.xu[i] = t; }
    }
    S.rhobeg = fabs(dx[0] / S.scl[0]);
    S.rhoend = xtol_rel * S.rhobeg;   // xtol_abs is zero on this path
    const double rhobeg = S.rhobeg;
    BQ_NOUNROLL for (int j = 0; j < N; j++) {
        const double temp = S.xu[j] - S.xl[j];
        if (temp < rhobeg + rhobeg) { S.rc = R_INVALID_ARGS; S.pc = PC_FINISHED; return DONE; }
        S.sl[j] = S.xl[j] - S.x[j];
        S.su[j] = S.xu[j] - S.x[j];
        if (S.sl[j] >= -rhobeg) {
            if (S.sl[j] >= 0.0) { S.x[j] = S.xl[j]; S.sl[j] = 0.0; S.su[j] = temp; }
            else { S.x[j] = S.xl[j] + rhobeg; S.sl[j] = -rhobeg; S.su[j] = dmax(S.xu[j] - S.x[j], rhobeg); }
        } else if (S.su[j] <= rhobeg) {
            if (S.su[j] <= 0.0) { S.x[j] = S.xu[j]; S.sl[j] = -temp; S.su[j] = 0.0; }
            else { S.x[j] = S.xu[j] - rhobeg; S.sl[j] = dmin(S.xl[j] - S.x[j], -rhobeg); S.su[j] = rhobeg; }
        }
    }
    // PRELIM initialisation
    BQ_NOUNROLL for (int j = 0; j < N; j++) {
        S.xbase[j] = S.x[j];
        BQ_NOUNROLL for (int k = 0; k < NPT; k++) S.xpt[k][j] = 0.0;
        BQ_NOUNROLL for (int i = 0; i < NDIM; i++) S.bmat[i][j] = 0.0;
    }
    BQ_NOUNROLL for (int ih = 0; ih < NH; ih++) S.hq[ih] = 0.0;
    BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
        S.pq[k] = 0.0;
        BQ_UNROLL for (int j = 0; j < NPTM; j++) S.zmat[k][j] = 0.0;
    }
    S.nf = 0;
    S.pc = PC_PRELIM_EVAL;
    // emit the first point (the base point itself)
    S.nf = 1;
    point_from(S, S.xpt[0]);
    BQ_UNROLL for (int i = 0; i < N; i++) xs_out[i] = S.x[i] * S.scl[i];
    return ASK;
}

// final point in the caller's space (bobyqb exit block + unscale)
template <class ST>
BQ_HD void result_x(const ST& S, double* xs_out) {
    BQ_UNROLL for (int i = 0; i < N; i++) xs_out[i] = S.x[i] * S.scl[i];
}

template <class ST, unsigned PHASES = PH_ALL, bool DEFER = true>
BQ_HDN int advance(ST& S, double f_in, double* xs_out) {
    typedef typename ST::real R;
    BQ_ASSUME_SHARED(S);
    R (*xpt)[N] = S.xpt; R (*bmat)[N] = S.bmat; R (*zmat)[NPTM] = S.zmat;
    R* xopt = S.xopt; R* gopt = S.gopt; R* hq = S.hq; R* pq = S.pq; R* fval = S.fval;
    R* sl = S.sl; R* su = S.su; R* xnew = S.xnew; R* xalt = S.xalt; R* d = S.d;
    R* vlag = S.vlag; R* w = S.w;
    int lbl = L_NONE;
    int result = DONE;
    bool done = false;
    const unsigned wmask = BQ_ACTIVE_MASK();   // the lanes that entered together stay together until all are done

    if (S.pc == PC_FINISHED) { done = true; }
    else

    // ------------------------------------------------------------------ PRELIM (first 7 evaluations)
    if (S.pc == PC_PRELIM_EVAL) {
        const double rhobeg = S.rhobeg;
        const double rhosq = rhobeg * rhobeg;
        const int nf = S.nf;           // 1-based count of the value just received
        const int nfm = nf - 1, nfx = nfm - N;
        const double f = f_in;
        S.nevals++;
        fval[nf - 1] = f;
        if (nf == 1) { S.fbeg = f; S.kopt = 0; }
        else if (f < fval[S.kopt]) S.kopt = nf - 1;
        if (nf >= 2 && nf <= N + 1) {
            gopt[nfm - 1] = (f - S.fbeg) / S.stepa;
        } else if (nf >= N + 2) {
            const int c = nfx - 1;                 // coordinate index
            const int ih = hidx(c, c);
            const double stepa = S.stepa, stepb = S.stepb;
            const double temp = (f - S.fbeg) / stepb;
            const double diff = stepb - stepa;
            hq[ih] = 2.0 * (temp - gopt[c]) / diff;
            gopt[c] = (gopt[c] * stepb - temp * stepa) / diff;
            if (stepa * stepb < 0.0) {
                if (f < fval[nf - 1 - N]) {
                    fval[nf - 1] = fval[nf - 1 - N];
                    fval[nf - 1 - N] = f;
                    if (S.kopt == nf - 1) S.kopt = nf - 1 - N;
                    xpt[nf - 1 - N][c] = stepb;
                    xpt[nf - 1][c] = stepa;
                }
            }
            bmat[0][c] = -(stepa + stepb) / (stepa * stepb);
            bmat[nf - 1][c] = -0.5 / xpt[nf - 1 - N][c];
            bmat[nf - 1 - N][c] = -bmat[0][c] - bmat[nf - 1][c];
            zmat[0][c] = sqrt(2.0) / (stepa * stepb);
            zmat[nf - 1][c] = sqrt(0.5) / rhosq;
            zmat[nf - 1 - N][c] = -zmat[0][c] - zmat[nf - 1][c];
        }
        const bool stop_evals = (S.maxeval > 0 && S.nevals >= S.maxeval);
        if (!stop_evals && nf < NPT) {
            // next preliminary point
            const int nf2 = nf + 1, nfm2 = nf, nfx2 = nfm2 - N;
            if (nfm2 >= 1 && nfm2 <= N) {
                double stepa = rhobeg;
                if (su[nfm2 - 1] == 0.0) stepa = -stepa;
                xpt[nf2 - 1][nfm2 - 1] = stepa;
                S.stepa = stepa;
            } else {
                const double stepa = xpt[nf2 - 1 - N][nfx2 - 1];
                double stepb = -rhobeg;
                if (sl[nfx2 - 1] == 0.0) stepb = dmin(2.0 * rhobeg, su[nfx2 - 1]);
                if (su[nfx2 - 1] == 0.0) stepb = dmax(-2.0 * rhobeg, sl[nfx2 - 1]);
                xpt[nf2 - 1][nfx2 - 1] = stepb;
                S.stepa = stepa; S.stepb = stepb;
            }
            S.nf = nf2;
            point_from(S, xpt[nf2 - 1]);
            BQ_UNROLL for (int i = 0; i < N; i++) xs_out[i] = S.x[i] * S.scl[i];
            result = ASK; done = true;
        } else {
        // prelim finished (or ran out of evaluations): bobyqb start-up
        S.xoptsq = 0.0;
        BQ_UNROLL for (int i = 0; i < N; i++) { xopt[i] = xpt[S.kopt][i]; S.xoptsq += xopt[i] * xopt[i]; }
        S.fsave = fval[0];
        if (stop_evals) { S.rc = R_MAXEVAL_REACHED; lbl = L_EXIT; }
        else {
            S.kbase = 0;
            S.rho = rhobeg;
            S.delta = S.rho;
            S.nresc = S.nevals;
            S.ntrits = 0;
            S.diffa = 0.0; S.diffb = 0.0; S.diffc = 0.0;
            S.itest = 0;
            S.nfsav = S.nevals;
            S.ratio = 0.0;
            lbl = L_GOPT_FIX;
        }
        }
    } else if (S.pc == PC_YIELD_TRUST) {
        S.pc = PC_MAIN_EVAL;
        lbl = L_TRUST;
    } else if (S.pc >= PC_YIELD_LABEL) {
        lbl = S.pc - PC_YIELD_LABEL;
        S.pc = PC_MAIN_EVAL;
    } else if (S.pc == PC_MAIN_EVAL) {
        S.nevals++;
        S.f = f_in;
        lbl = L_AFTER_EVAL;
    } else {  // PC_RESCUE_EVAL
        S.nevals++;
        const double f = f_in;
        const int kpt = S.kpt;
        fval[kpt] = f;
        if (f < fval[S.kopt]) S.kopt = kpt;
        if (S.maxeval > 0 && S.nevals >= S.maxeval) {
            // rescue returns MAXEVAL_REACHED after recording f but before updating the model
            S.rc = R_MAXEVAL_REACHED;
            lbl = L_RESCUE_DONE;
        } else {
            rescue_absorb(S, kpt, f);
            S.kpt = kpt + 1;
            lbl = L_RESCUE_LOOP;
        }
    }

    int ntrust = 0;   // trust-region steps taken in this call (BQ_DEFER_TRUST)
    (void)ntrust;
    for (;;) {
        if (PHASES != PH_ALL && !done && lbl < 32 && ((PH_HEAVY >> lbl) & 1u) && !((PHASES >> lbl) & 1u)) {
            S.pc = PC_YIELD_LABEL + lbl;      // this phase does not run the block: the kernel of its phase picks the patch up
            result = YIELD; done = true;
        }
        BQ_SCHED_BEGIN(wmask, done, lbl)
#if defined(__CUDA_ARCH__)
        const unsigned sel = __activemask();
#else
        const unsigned sel = 0xffffffffu;
#endif
#if defined(__CUDA_ARCH__) && defined(HP_PROFILE)
        const long long prof_t0 = clock64();
        const int prof_lbl = lbl;
#endif
        switch (lbl) {
        // -------------------------------------------------------------- gopt correction when kopt moved
        case L_GOPT_FIX: {
            if (S.kopt != S.kbase) {
                int ih = 0;
                BQ_NOUNROLL for (int j = 0; j < N; j++)
                    BQ_NOUNROLL for (int i = 0; i <= j; i++) {
                        if (i < j) gopt[j] += hq[ih] * xopt[i];
                        gopt[i] += hq[ih] * xopt[j];
                        ih++;
                    }
                if (S.nevals > NPT) {
                    BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
                        double temp = 0.0;
                        BQ_UNROLL for (int j = 0; j < N; j++) temp += xpt[k][j] * xopt[j];
                        temp = pq[k] * temp;
                        BQ_UNROLL for (int i = 0; i < N; i++) gopt[i] += temp * xpt[k][i];
                    }
                }
            }
            lbl = L_TRUST;
            break;
        }
        // -------------------------------------------------------------- trust-region step
        case L_TRUST: {
#if defined(__CUDA_ARCH__) && BQ_DEFER_TRUST > 0
            if (DEFER && ntrust >= 1 && __popc(wmask) >= BQ_DEFER_TRUST) {
                S.pc = PC_YIELD_TRUST;
                result = YIELD; done = true;
                break;
            }
            ++ntrust;
#endif
            trsbox(S, sel);
            S.dnorm = dmin(S.delta, sqrt(S.dsq));
            if (S.dnorm < 0.5 * S.rho) {
                S.ntrits = -1;
                const double t = 10.0 * S.rho;
                S.distsq = t * t;
                if (S.nevals <= S.nfsav + 2) { lbl = L_FARPOINT; break; }
                const double errbig = dmax(dmax(S.diffa, S.diffb), S.diffc);
                const double frhosq = S.rho * .125 * S.rho;
                if (S.crvmin > 0.0 && errbig > frhosq * S.crvmin) { lbl = L_FARPOINT; break; }
                const double bdtol = errbig / S.rho;
                bool far = false;
                BQ_NOUNROLL for (int j = 0; j < N; j++) {
                    double bdtest = bdtol;
                    if (xnew[j] == sl[j]) bdtest = w[j];
                    if (xnew[j] == su[j]) bdtest = -w[j];
                    if (bdtest < bdtol) {
                        double curv = hq[hidx(j, j)];
                        BQ_NOUNROLL for (int k = 0; k < NPT; k++) curv += pq[k] * (xpt[k][j] * xpt[k][j]);
                        bdtest += 0.5 * curv * S.rho;
                        if (bdtest < bdtol) { far = true; break; }
                    }
                }
                lbl = far ? L_FARPOINT : L_REDUCE_RHO;
                break;
            }
            ++S.ntrits;
            lbl = L_SHIFT;
            break;
        }
        // -------------------------------------------------------------- shift xbase to xopt when far away
        case L_SHIFT: {
            if (S.dsq <= S.xoptsq * .001) {
                const double xoptsq = S.xoptsq;
                const double fracsq = xoptsq * .25;
                double sumpq = 0.0;
                BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
                    sumpq += pq[k];
                    double sum = -0.5 * xoptsq;
                    BQ_UNROLL for (int i = 0; i < N; i++) sum += xpt[k][i] * xopt[i];
                    w[NPT + k] = sum;
                    const double temp = fracsq - 0.5 * sum;
                    BQ_NOUNROLL for (int i = 0; i < N; i++) {
                        w[i] = bmat[k][i];
                        vlag[i] = sum * xpt[k][i] + temp * xopt[i];
                        const int ip = NPT + i;
                        BQ_NOUNROLL for (int j = 0; j <= i; j++) bmat[ip][j] = bmat[ip][j] + w[i] * vlag[j] + vlag[i] * w[j];
                    }
                }
                BQ_NOUNROLL for (int jj = 0; jj < NPTM; jj++) {
                    double sumz = 0.0, sumw = 0.0;
                    BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
                        sumz += zmat[k][jj];
                        vlag[k] = w[NPT + k] * zmat[k][jj];
                        sumw += vlag[k];
                    }
                    BQ_NOUNROLL for (int j = 0; j < N; j++) {
                        double sum = (fracsq * sumz - 0.5 * sumw) * xopt[j];
                        BQ_NOUNROLL for (int k = 0; k < NPT; k++) sum += vlag[k] * xpt[k][j];
                        w[j] = sum;
                        BQ_NOUNROLL for (int k = 0; k < NPT; k++) bmat[k][j] += sum * zmat[k][jj];
                    }
                    BQ_NOUNROLL for (int i = 0; i < N; i++) {
                        const int ip = i + NPT;
                        const double temp = w[i];
                        BQ_NOUNROLL for (int j = 0; j <= i; j++) bmat[ip][j] += temp * w[j];
                    }
                }
                int ih = 0;
                BQ_NOUNROLL for (int j = 0; j < N; j++) {
                    w[j] = -0.5 * sumpq * xopt[j];
                    BQ_NOUNROLL for (int k = 0; k < NPT; k++) { w[j] += pq[k] * xpt[k][j]; xpt[k][j] -= xopt[j]; }
                    BQ_NOUNROLL for (int i = 0; i <= j; i++) {
                        hq[ih] = hq[ih] + w[i] * xopt[j] + xopt[i] * w[j];
                        bmat[NPT + i][j] = bmat[NPT + j][i];
                        ih++;
                    }
                }
                BQ_NOUNROLL for (int i = 0; i < N; i++) {
                    S.xbase[i] += xopt[i];
                    xnew[i] -= xopt[i];
                    sl[i] -= xopt[i];
                    su[i] -= xopt[i];
                    xopt[i] = 0.0;
                }
                S.xoptsq = 0.0;
            }
            lbl = (S.ntrits == 0) ? L_ALTMOV : L_VLAG;
            break;
        }
        // -------------------------------------------------------------- RESCUE
        case L_RESCUE: {
            S.nfsav = S.nevals;
            S.kbase = S.kopt;
            S.n_rescue++;
            rescue_setup(S);
            if (S.nrem_r == 0) { lbl = L_RESCUE_DONE; break; }
            S.kpt = 0;
            lbl = L_RESCUE_LOOP;
            break;
        }
        case L_RESCUE_LOOP: {
            R* ptsid = S.w + 6;
            int kpt = S.kpt;
            while (kpt < NPT && ptsid[kpt] == 0.0) kpt++;
            if (kpt >= NPT) { lbl = L_RESCUE_DONE; break; }
            if (S.maxeval > 0 && S.nevals >= S.maxeval) { S.rc = R_MAXEVAL_REACHED; lbl = L_RESCUE_DONE; break; }
            S.kpt = kpt;
            rescue_place(S, kpt);
            S.pc = PC_RESCUE_EVAL;
            BQ_UNROLL for (int i = 0; i < N; i++) xs_out[i] = S.x[i] * S.scl[i];
            result = ASK; done = true;
            break;
        }
        case L_RESCUE_DONE: {
            S.xoptsq = 0.0;
            if (S.kopt != S.kbase) {
                BQ_UNROLL for (int i = 0; i < N; i++) { xopt[i] = xpt[S.kopt][i]; S.xoptsq += xopt[i] * xopt[i]; }
            }
            if (S.rc != R_SUCCESS) { lbl = L_EXIT; break; }
            S.nresc = S.nevals;
            if (S.nfsav < S.nevals) { S.nfsav = S.nevals; lbl = L_GOPT_FIX; break; }
            if (S.ntrits > 0) { lbl = L_TRUST; break; }
            lbl = L_ALTMOV;
            break;
        }
        // -------------------------------------------------------------- alternative (geometry) step
        case L_ALTMOV: {
            altmov(S);
            BQ_UNROLL for (int i = 0; i < N; i++) d[i] = xnew[i] - xopt[i];
            lbl = L_VLAG;
            break;
        }
        // -------------------------------------------------------------- Lagrange values, beta, denominator
        case L_VLAG: {
            BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
                double suma = 0.0, sumb = 0.0, sum = 0.0;
                BQ_NOUNROLL for (int j = 0; j < N; j++) {
                    suma += xpt[k][j] * d[j];
                    sumb += xpt[k][j] * xopt[j];
                    sum += bmat[k][j] * d[j];
                }
                w[k] = suma * (0.5 * suma + sumb);
                vlag[k] = sum;
                w[NPT + k] = suma;
            }
            double beta = 0.0;
            BQ_NOUNROLL for (int jj = 0; jj < NPTM; jj++) {
                double sum = 0.0;
                BQ_NOUNROLL for (int k = 0; k < NPT; k++) sum += zmat[k][jj] * w[k];
                beta -= sum * sum;
                BQ_NOUNROLL for (int k = 0; k < NPT; k++) vlag[k] += sum * zmat[k][jj];
            }
            double dsq = 0.0, bsum = 0.0, dx = 0.0;
            BQ_NOUNROLL for (int j = 0; j < N; j++) {
                dsq += d[j] * d[j];
                double sum = 0.0;
                BQ_NOUNROLL for (int k = 0; k < NPT; k++) sum += w[k] * bmat[k][j];
                bsum += sum * d[j];
                const int jp = NPT + j;
                BQ_UNROLL for (int i = 0; i < N; i++) sum += bmat[jp][i] * d[i];
                vlag[jp] = sum;
                bsum += sum * d[j];
                dx += d[j] * xopt[j];
            }
            S.dsq = dsq;
            beta = dx * dx + dsq * (S.xoptsq + dx + dx + 0.5 * dsq) + beta - bsum;
            S.beta = beta;
            vlag[S.kopt] += 1.0;
            if (S.ntrits == 0) {
                const double vk = vlag[S.knew];
                S.denom = vk * vk + S.alpha * beta;
                if (S.denom < S.cauchy && S.cauchy > 0.0) {
                    BQ_UNROLL for (int i = 0; i < N; i++) { xnew[i] = xalt[i]; d[i] = xnew[i] - xopt[i]; }
                    S.cauchy = 0.0;
                    lbl = L_VLAG;
                    break;
                }
                if (S.denom <= 0.5 * (vk * vk)) {
                    if (S.nevals > S.nresc) { lbl = L_RESCUE; break; }
                    S.rc = R_ROUNDOFF_LIMITED;
                    lbl = L_EXIT;
                    break;
                }
            } else {
                const double delsq = S.delta * S.delta;
                double scaden = 0.0, biglsq = 0.0;
                S.knew = -1;
                BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
                    if (k == S.kopt) continue;
                    double hdiag = 0.0;
                    BQ_UNROLL for (int jj = 0; jj < NPTM; jj++) hdiag += zmat[k][jj] * zmat[k][jj];
                    const double den = beta * hdiag + vlag[k] * vlag[k];
                    double distsq = 0.0;
                    BQ_UNROLL for (int j = 0; j < N; j++) { const double t = xpt[k][j] - xopt[j]; distsq += t * t; }
                    const double r = distsq / delsq;
                    const double temp = dmax(1.0, r * r);
                    if (temp * den > scaden) { scaden = temp * den; S.knew = k; S.denom = den; }
                    biglsq = dmax(biglsq, temp * (vlag[k] * vlag[k]));
                }
                if (scaden <= 0.5 * biglsq) {
                    if (S.nevals > S.nresc) { lbl = L_RESCUE; break; }
                    S.rc = R_ROUNDOFF_LIMITED;
                    lbl = L_EXIT;
                    break;
                }
            }
            lbl = L_EVAL;
            break;
        }
        // -------------------------------------------------------------- ask for f(xbase + xnew)
        case L_EVAL: {
            point_from(S, xnew);
            if (S.maxeval > 0 && S.nevals >= S.maxeval) { S.rc = R_MAXEVAL_REACHED; lbl = L_EXIT; break; }
            S.pc = PC_MAIN_EVAL;
            BQ_UNROLL for (int i = 0; i < N; i++) xs_out[i] = S.x[i] * S.scl[i];
            result = ASK; done = true;
            break;
        }
        case L_AFTER_EVAL: {
            const double f = S.f;
            if (S.ntrits == -1) {
                S.fsave = f;
                S.rc = R_XTOL_REACHED;
                if (S.fsave < fval[S.kopt]) { S.minf = f; S.pc = PC_FINISHED; result = DONE; done = true; break; }  // x stays at the new point
                lbl = L_EXIT;
                break;
            }
            const double fopt = fval[S.kopt];
            double vquad = 0.0;
            {
                int ih = 0;
                BQ_NOUNROLL for (int j = 0; j < N; j++) {
                    vquad += d[j] * gopt[j];
                    BQ_NOUNROLL for (int i = 0; i <= j; i++) {
                        double temp = d[i] * d[j];
                        if (i == j) temp = 0.5 * temp;
                        vquad += hq[ih] * temp;
                        ih++;
                    }
                }
            }
            BQ_NOUNROLL for (int k = 0; k < NPT; k++) vquad += 0.5 * pq[k] * (w[NPT + k] * w[NPT + k]);
            const double diff = f - fopt - vquad;
            S.diffc = S.diffb;
            S.diffb = S.diffa;
            S.diffa = fabs(diff);
            if (S.dnorm > S.rho) S.nfsav = S.nevals;
            if (S.ntrits > 0) {
                if (vquad >= 0.0) { S.rc = R_ROUNDOFF_LIMITED; lbl = L_EXIT; break; }
                S.ratio = (f - fopt) / vquad;
                if (S.ratio <= 0.1) S.delta = dmin(0.5 * S.delta, S.dnorm);
                else if (S.ratio <= .7) S.delta = dmax(0.5 * S.delta, S.dnorm);
                else S.delta = dmax(0.5 * S.delta, S.dnorm + S.dnorm);
                if (S.delta <= S.rho * 1.5) S.delta = S.rho;
                if (f < fopt) {
                    const int ksav = S.knew;
                    const double densav = S.denom;
                    const double delsq = S.delta * S.delta;
                    double scaden = 0.0, biglsq = 0.0;
                    S.knew = -1;
                    BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
                        double hdiag = 0.0;
                        BQ_UNROLL for (int jj = 0; jj < NPTM; jj++) hdiag += zmat[k][jj] * zmat[k][jj];
                        const double den = S.beta * hdiag + vlag[k] * vlag[k];
                        double distsq = 0.0;
                        BQ_UNROLL for (int j = 0; j < N; j++) { const double t = xpt[k][j] - xnew[j]; distsq += t * t; }
                        const double r = distsq / delsq;
                        const double temp = dmax(1.0, r * r);
                        if (temp * den > scaden) { scaden = temp * den; S.knew = k; S.denom = den; }
                        biglsq = dmax(biglsq, temp * (vlag[k] * vlag[k]));
                    }
                    if (scaden <= 0.5 * biglsq) { S.knew = ksav; S.denom = densav; }
                }
            }
            const int knew = S.knew;
            update(S, S.beta, S.denom, knew, w);
            {
                int ih = 0;
                const double pqold = pq[knew];
                pq[knew] = 0.0;
                BQ_NOUNROLL for (int i = 0; i < N; i++) {
                    const double temp = pqold * xpt[knew][i];
                    BQ_NOUNROLL for (int j = 0; j <= i; j++) { hq[ih] += temp * xpt[knew][j]; ih++; }
                }
            }
            BQ_NOUNROLL for (int jj = 0; jj < NPTM; jj++) {
                const double temp = diff * zmat[knew][jj];
                BQ_NOUNROLL for (int k = 0; k < NPT; k++) pq[k] += temp * zmat[k][jj];
            }
            fval[knew] = f;
            BQ_UNROLL for (int i = 0; i < N; i++) { xpt[knew][i] = xnew[i]; w[i] = bmat[knew][i]; }
            bool singular = false;
            BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
                double suma = 0.0;
                BQ_UNROLL for (int jj = 0; jj < NPTM; jj++) suma += zmat[knew][jj] * zmat[k][jj];
                if (isinf(suma)) { singular = true; break; }
                double sumb = 0.0;
                BQ_UNROLL for (int j = 0; j < N; j++) sumb += xpt[k][j] * xopt[j];
                const double temp = suma * sumb;
                BQ_UNROLL for (int i = 0; i < N; i++) w[i] += temp * xpt[k][i];
            }
            if (singular) { S.rc = R_ROUNDOFF_LIMITED; lbl = L_EXIT; break; }
            BQ_UNROLL for (int i = 0; i < N; i++) gopt[i] += diff * w[i];
            if (f < fopt) {
                S.kopt = knew;
                S.xoptsq = 0.0;
                int ih = 0;
                BQ_NOUNROLL for (int j = 0; j < N; j++) {
                    xopt[j] = xnew[j];
                    S.xoptsq += xopt[j] * xopt[j];
                    BQ_NOUNROLL for (int i = 0; i <= j; i++) {
                        if (i < j) gopt[j] += hq[ih] * d[i];
                        gopt[i] += hq[ih] * d[j];
                        ih++;
                    }
                }
                BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
                    double temp = 0.0;
                    BQ_UNROLL for (int j = 0; j < N; j++) temp += xpt[k][j] * d[j];
                    temp = pq[k] * temp;
                    BQ_UNROLL for (int i = 0; i < N; i++) gopt[i] += temp * xpt[k][i];
                }
                // nlopt_stop_ftol never fires here: ftol_rel = ftol_abs = 0 on this path (util/stop.c:28-34)
            }
            if (S.ntrits > 0) {
                // least Frobenius norm interpolant test
                BQ_NOUNROLL for (int k = 0; k < NPT; k++) { vlag[k] = fval[k] - fval[S.kopt]; w[k] = 0.0; }
                BQ_NOUNROLL for (int j = 0; j < NPTM; j++) {
                    double sum = 0.0;
                    BQ_NOUNROLL for (int k = 0; k < NPT; k++) sum += zmat[k][j] * vlag[k];
                    BQ_NOUNROLL for (int k = 0; k < NPT; k++) w[k] += sum * zmat[k][j];
                }
                BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
                    double sum = 0.0;
                    BQ_UNROLL for (int j = 0; j < N; j++) sum += xpt[k][j] * xopt[j];
                    w[k + NPT] = w[k];
                    w[k] = sum * w[k];
                }
                double gqsq = 0.0, gisq = 0.0;
                BQ_NOUNROLL for (int i = 0; i < N; i++) {
                    double sum = 0.0;
                    BQ_NOUNROLL for (int k = 0; k < NPT; k++) sum = sum + bmat[k][i] * vlag[k] + xpt[k][i] * w[k];
                    if (xopt[i] == sl[i]) {
                        const double a = dmin(0.0, gopt[i]); gqsq += a * a;
                        const double b = dmin(0.0, sum); gisq += b * b;
                    } else if (xopt[i] == su[i]) {
                        const double a = dmax(0.0, gopt[i]); gqsq += a * a;
                        const double b = dmax(0.0, sum); gisq += b * b;
                    } else {
                        gqsq += gopt[i] * gopt[i];
                        gisq += sum * sum;
                    }
                    vlag[NPT + i] = sum;
                }
                ++S.itest;
                if (gqsq < 10.0 * gisq) S.itest = 0;
                if (S.itest >= 3) {
                    BQ_UNROLL for (int i = 0; i < N; i++) gopt[i] = vlag[NPT + i];
                    BQ_NOUNROLL for (int k = 0; k < NPT; k++) pq[k] = w[NPT + k];
                    BQ_NOUNROLL for (int ih = 0; ih < NH; ih++) hq[ih] = 0.0;
                    S.itest = 0;
                }
            }
            if (S.ntrits == 0) { lbl = L_TRUST; break; }
            if (f <= fopt + 0.1 * vquad) { lbl = L_TRUST; break; }
            {
                const double a = 2.0 * S.delta, b = 10.0 * S.rho;
                S.distsq = dmax(a * a, b * b);
            }
            lbl = L_FARPOINT;
            break;
        }
        // -------------------------------------------------------------- is some point too far from xopt?
        case L_FARPOINT: {
            S.knew = -1;
            BQ_NOUNROLL for (int k = 0; k < NPT; k++) {
                double sum = 0.0;
                BQ_UNROLL for (int j = 0; j < N; j++) { const double t = xpt[k][j] - xopt[j]; sum += t * t; }
                if (sum > S.distsq) { S.knew = k; S.distsq = sum; }
            }
            if (S.knew >= 0) {
                const double dist = sqrt(S.distsq);
                if (S.ntrits == -1) {
                    S.delta = dmin(0.1 * S.delta, 0.5 * dist);
                    if (S.delta <= S.rho * 1.5) S.delta = S.rho;
                }
                S.ntrits = 0;
                S.adelt = dmax(dmin(0.1 * dist, S.delta), S.rho);
                S.dsq = S.adelt * S.adelt;
                lbl = L_SHIFT;
                break;
            }
            if (S.ntrits == -1) { lbl = L_REDUCE_RHO; break; }
            if (S.ratio > 0.0) { lbl = L_TRUST; break; }
            if (dmax(S.delta, S.dnorm) > S.rho) { lbl = L_TRUST; break; }
            lbl = L_REDUCE_RHO;
            break;
        }
        // -------------------------------------------------------------- next rho, or finish
        case L_REDUCE_RHO: {
            if (S.rho > S.rhoend) {
                S.delta = 0.5 * S.rho;
                S.ratio = S.rho / S.rhoend;
                if (S.ratio <= 16.) S.rho = S.rhoend;
                else if (S.ratio <= 250.) S.rho = sqrt(S.ratio) * S.rhoend;
                else S.rho = 0.1 * S.rho;
                S.delta = dmax(S.delta, S.rho);
                S.ntrits = 0;
                S.nfsav = S.nevals;
                lbl = L_TRUST;
                break;
            }
            if (S.ntrits == -1) { lbl = L_EVAL; break; }
            lbl = L_EXIT;
            break;
        }
        case L_EXIT: {
            point_from(S, xopt);
            S.minf = fval[S.kopt];
            S.pc = PC_FINISHED;
            result = DONE; done = true;
            break;
        }
        default:
            S.rc = R_FAILURE; S.pc = PC_FINISHED;
            result = DONE; done = true;
            break;
        }
#if defined(__CUDA_ARCH__) && defined(HP_PROFILE)
        {
            const unsigned now = __activemask();
            unsigned lane_id; asm("mov.u32 %0, %%laneid;" : "=r"(lane_id));
            if (lane_id == (unsigned)(__ffs(now) - 1)) {
                atomicAdd(&g_prof_cycles[prof_lbl], (unsigned long long)(clock64() - prof_t0));
                atomicAdd(&g_prof_trips[prof_lbl], 1ull);
                atomicAdd(&g_prof_lanes[prof_lbl], (unsigned long long)__popc(now));
            }
        }
#endif
    }
    return result;
}

}  // namespace bq3
