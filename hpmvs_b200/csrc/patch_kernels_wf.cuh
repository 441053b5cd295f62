// Wavefront form of the fused optimize path (sm_100a): the same per-patch work as hp::optimize_kernel (patch_kernels.cuh) - everything
// below PatchOptimizer::optimize(), /root/reference/src/hpmvs/PatchOptimizer.cpp:48-103 - but decomposed into one small kernel per
// PHASE of a refinement round instead of one persistent warp-specialised kernel:
//
//     fill   (once)      : every slot fetches a patch and runs the stages before the refinement              (warp = patch)
//     loop while patches are alive  (a CUDA-graph WHILE node, no host round trip per round):
//       advance A / T / B : BOBYQA (bobyqa3.h) absorbs the objective value (A), takes its trust-region step (T), shifts / picks
//                           the geometry step / forms the Lagrange values (B) and emits the next point       (lane = patch)
//       eval              : objective at the emitted points: 7x7 gather in every view + NCC                  (warp = patch)
//       post              : finished refinements: stages after the refinement, result record, slot refill    (warp = patch)
//                           + (last CTA) list bookkeeping and the loop condition
//
// Why: in the persistent kernel both roles are latency bound with 12 warps per SM (2 optimizer warps walking ~100 KB of FP64 code,
// 10 sampler warps; profiles/r1_cycle_breakdown.md, r2_cycle_dump_city100_r1kernel.txt) and every extra optimizer warp slowed the
// others down (instruction cache).  Here every phase is its own kernel with a small instruction footprint, all warps of an SM run
// the same code, and each phase gets the whole SM's occupancy to hide the dependent FP64 / shared-memory latency chains.
//
// State layout: the optimizer states live in HBM/L2 as TILES of 32 interleaved states (bq3::StateTile): member m of slot 32*t + l
// sits at tile t, byte 256*m + 8*l, so lane l of a warp that owns tile t issues one fully coalesced 256-byte request per member
// access and every member keeps an immediate offset.  Patch contexts (LaneCtx, 272 B) are an AoS array.  Nothing of the arithmetic
// changes: the device functions below are the ones of patch_kernels.cuh / bobyqa3.h, so results are bit-identical to the persistent
// kernel (tests/test_gpu_parity.py::test_wavefront_*).
#pragma once

#include "patch_kernels.cuh"

namespace hp {

enum : int {
    WS_EMPTY = 0, WS_NEW = 1, WS_NEED_EVAL = 2, WS_HAVE_F = 3, WS_POST = 4, WS_DEAD = 5,
    WS_YIELD = 32          // + label: the optimizer stopped in front of heavy block `label` (bq3::PC_YIELD_LABEL)
};

struct WfCtl {
    int work_counter;      // next input patch (shared by all slots of this launch)
    int dead;              // slots that will not receive another patch
    int eval_cnt;          // entries in eval_list (produced by advance, consumed by eval of the same round)
    int post_cnt;          // entries in post_list (produced by advance, consumed by post of the same round)
    int post_ticket;       // dynamic work distribution inside the post kernel
    int fill_ticket;       // ... and inside the fill kernel (runs once, before the loop)
    int round;
    int live;              // != 0 while another round is needed (host-loop mode reads it back)
    int overrun;           // set when max_rounds was hit with live slots (never expected; reported by the host)
    int ctas_done;         // post kernel: CTAs that have finished this round (the last one runs wf_sched)
    int run_post;          // sched's decision for the NEXT round: the post kernel runs (it is latency bound: ~100 us even for one patch)
};

struct WfParams {
    KParams K;
    int M;                             // slots in flight for this launch (multiple of 32, <= capacity)
    int max_rounds;
    LaneCtx* ctx;                      // [capacity]
    unsigned char* tiles;              // capacity / 32 tiles of sizeof(bq3::StateTile) bytes
    double* fval;                      // [capacity] objective values
    int* sstate;                       // [capacity]
    int* eval_list;                    // [capacity]
    int* post_list;                    // [capacity]
    WfCtl* ctl;
    cudaGraphConditionalHandle cond;   // WHILE node of the launch graph (graph mode)
    int use_cond;
    int kernels_per_round;             // 3 (advance, eval, post) or 5 (advance phases A / T / B)
    unsigned long long* round_log;     // optional (HPMVS_WF_LOG): per round {globaltimer ns, eval_cnt, post_cnt, dead}
    int round_log_cap;
};

constexpr int WF_SAMPLER_WARPS = 4;    // eval / post / fill kernels: 4 warps x 7 KB scratch per CTA (7 CTAs = 28 warps per SM)
constexpr int WF_ADV_THREADS = 128;    // advance kernels: 4 tiles per CTA
#ifndef WF_ADV_MIN_CTAS
#define WF_ADV_MIN_CTAS 4              // 128 registers: 61.1 vs 64.9 ms per city100 step against the uncapped 162 (profiles/r2_wavefront_experiments.md)
#endif
#ifndef WF_PREFETCH
#define WF_PREFETCH 0              // measured neutral (city100 61.6 vs 62.3 ms per step): the optimizer is bound by its own dependent chains
#endif

__device__ __forceinline__ bq3::StateTile& wf_state(const WfParams& P, int slot) {
    return *reinterpret_cast<bq3::StateTile*>(P.tiles + (size_t)(slot >> 5) * sizeof(bq3::StateTile) + (size_t)(slot & 31) * 8);
}

// warp-aggregated append of `slot` to a list for the lanes with `pred`
__device__ __forceinline__ void wf_push(int* list, int* cnt, bool pred, int slot) {
    const unsigned act = __activemask();
    const unsigned m = __ballot_sync(act, pred);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(cnt, __popc(m));
    base = __shfl_sync(act, base, leader);
    if (pred) list[base + __popc(m & ((1u << lane) - 1u))] = slot;
}

// ---- sched: end of a round (run by the last CTA of the post kernel) ----------------------------------------------------------------
__device__ __forceinline__ void wf_sched(const WfParams& P) {
    WfCtl& c = *P.ctl;
    if (P.round_log && c.round < P.round_log_cap) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        unsigned long long* r = P.round_log + 4 * (size_t)c.round;
        r[0] = t; r[1] = (unsigned long long)c.eval_cnt; r[2] = (unsigned long long)c.post_cnt; r[3] = (unsigned long long)c.dead;
    }
    c.eval_cnt = 0;
    if (c.run_post) { c.post_cnt = 0; c.post_ticket = 0; }     // the post pass of this round consumed the list
    atomicAdd(&P.K.counters[13], (unsigned long long)P.kernels_per_round);   // launches are counted where they happen: on the device
    c.round = c.round + 1;
    // Post passes are batched: a pass costs the latency of one post-stage (several scoring evaluations per patch, ~100-200 us) whether
    // it serves one patch or ten thousand, so finished patches are collected and served when nothing else is left to do, or - while
    // input patches are still waiting for a slot - when enough slots (1/16) have piled up to be worth a refill pass.
    const int pending = c.post_cnt;
    const int active = P.M - c.dead - pending;
    const bool more_input = c.work_counter < P.K.n;
    int thr = P.M / 16;
    if (thr < 64) thr = 64;
    c.run_post = (pending > 0 && (active <= 0 || (more_input && pending >= thr))) ? 1 : 0;
    int live = (c.dead < P.M) ? 1 : 0;
    if (live && c.round >= P.max_rounds) { live = 0; c.overrun = 1; }
    c.live = live;
    if (P.use_cond) cudaGraphSetConditional(P.cond, live ? 1u : 0u);
}

// ---- fill (round 0: all slots) and post (finished refinements + refill): warp = patch --------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(WF_SAMPLER_WARPS * 32) wf_post_kernel(const WfParams* __restrict__ Pp) {
    const WfParams& P = *Pp;
    const KParams& K = P.K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NccWarp& WS = reinterpret_cast<NccWarp*>(smem_raw)[threadIdx.x >> 5];
    Scratch& W = WS.S;
    LaneCtx& C = WS.P;
    const int lane = threadIdx.x & 31;
    // finished patches wait in post_list until sched asks for a post pass
    const int count = FILL ? P.M : (P.ctl->run_post ? P.ctl->post_cnt : 0);
    unsigned long long cnt[4] = {0, 0, 0, 0};
    int ndead = 0;
    while (count > 0) {                                  // (a round without a post pass must not draw tickets)
        int t = 0;
        if (lane == 0) t = atomicAdd(FILL ? &P.ctl->fill_ticket : &P.ctl->post_ticket, 1);
        t = __shfl_sync(FULL, t, 0);
        if (t >= count) break;
        const int slot = FILL ? t : P.post_list[t];
        if (!FILL) {
            copy_in<(int)sizeof(LaneCtx)>(&C, &P.ctx[slot], lane);
            __syncwarp();
            int st = C.status;
            if (st == HPMVS_OK) st = post_stage(W, C, K, lane);
            retire_patch(W, C, K, lane, st);
            if (lane == 0) { cnt[0]++; cnt[1] += (st == HPMVS_OK); cnt[2] += C.evals; cnt[3] += C.textures; }
            __syncwarp();
        }
        const int ns = serve_fill(C, W, K, lane, cnt);
        __syncwarp();
        if (ns == ST_NEW) copy_out<(int)sizeof(LaneCtx)>(&P.ctx[slot], &C, lane);
        if (lane == 0) { P.sstate[slot] = (ns == ST_NEW) ? WS_NEW : WS_DEAD; ndead += (ns != ST_NEW); }
        __syncwarp();
    }
    if (lane == 0) {
        if (ndead) atomicAdd(&P.ctl->dead, ndead);
        if (cnt[0]) {
            atomicAdd(&K.counters[0], cnt[0]); atomicAdd(&K.counters[1], cnt[1]);
            atomicAdd(&K.counters[2], cnt[2]); atomicAdd(&K.counters[3], cnt[3]);
        }
    }
    if (!FILL) {
        // the post kernel closes the round: its last CTA to finish does the bookkeeping and sets the loop condition
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(&P.ctl->ctas_done, 1) == (int)gridDim.x - 1) {
                P.ctl->ctas_done = 0;
                __threadfence();
                wf_sched(P);
            }
        }
    }
}

// ---- eval: objective_fn (PatchOptimizer.cpp:286-311) at the point the optimizer emitted: warp = patch ------------------------------
__global__ void __launch_bounds__(WF_SAMPLER_WARPS * 32) wf_eval_kernel(const WfParams* __restrict__ Pp) {
    const WfParams& P = *Pp;
    const KParams& K = P.K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NccWarp& WS = reinterpret_cast<NccWarp*>(smem_raw)[threadIdx.x >> 5];
    Scratch& W = WS.S;
    LaneCtx& C = WS.P;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int count = P.ctl->eval_cnt;
    for (int i = warp; i < count; i += nwarps) {
        const int slot = P.eval_list[i];
        copy_in<(int)sizeof(LaneCtx)>(&C, &P.ctx[slot], lane);
        __syncwarp();
        const int tex0 = C.textures;
        eval_dots(W, C, K, lane, 0, false, true);
        const double f = objective_value(W, C, K, lane);
        __syncwarp();
        if (lane == 0) {
            if (C.textures != tex0) P.ctx[slot].textures = C.textures;
            P.fval[slot] = f;
            P.sstate[slot] = WS_HAVE_F;
        }
        __syncwarp();
    }
}

// ---- advance: BOBYQA, lane = patch slot, one kernel per phase (bq3::PH_A / PH_T / PH_B, or PH_ALL in one) ----------------------------
template <unsigned PHASES>
__global__ void __launch_bounds__(WF_ADV_THREADS, WF_ADV_MIN_CTAS) wf_advance_kernel(const WfParams* __restrict__ Pp) {
    const WfParams& P = *Pp;
    const KParams& K = P.K;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if ((slot & ~31) >= P.M) return;                       // whole warps leave together
    const int st0 = (slot < P.M) ? P.sstate[slot] : WS_DEAD;
    bool take = false;
    if (st0 >= WS_YIELD) take = ((PHASES >> (st0 - WS_YIELD)) & 1u) != 0;
    else if (st0 == WS_NEW || st0 == WS_HAVE_F) take = (PHASES & bq3::PH_A) == bq3::PH_A;
    int st = 0;                                              // next slot state (0 = unchanged)
    if (take) {
        LaneCtx& mine = P.ctx[slot];
        bq3::StateTile& bq = wf_state(P, slot);
#if WF_PREFETCH
        // every kernel starts with a cold L1: without this each first touch of a state member is a dependent L2 round trip (~210 of
        // them per round); issued up front they overlap, and the optimizer's own accesses then hit L1
        if (st0 != WS_NEW) {
            const char* base = reinterpret_cast<const char*>(&bq);
#pragma unroll 1
            for (int c = 0; c < (int)(sizeof(bq3::StateTile) / (8 * BQ_TILE_LANES)); c++)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(base + (size_t)c * 8 * BQ_TILE_LANES));
        }
#endif
        double xcur[3] = {0.0, 0.0, 0.0};
        int act;
        if (st0 == WS_NEW) {
            const double lb[3] = {-HUGE_VAL, -23.99999, -23.99999};
            const double ub[3] = {HUGE_VAL, 23.99999, 23.99999};
            double x0[3];
            init_parameters(mine, K, lb, ub, x0);
            act = bq3::start(bq, x0, lb, ub, 1.e-7, 1000, xcur);
        } else {
            const double f = (st0 == WS_HAVE_F) ? P.fval[slot] : 0.0;
            act = bq3::advance<bq3::StateTile, PHASES, false>(bq, f, xcur);
        }
        if (act == bq3::ASK) { set_center_norm(mine, K, xcur); st = WS_NEED_EVAL; }
        else if (act == bq3::YIELD) {
            const int pc = bq.pc;
            st = WS_YIELD + (pc == bq3::PC_YIELD_TRUST ? (int)bq3::L_TRUST : pc - (int)bq3::PC_YIELD_LABEL);
        } else {
            st = WS_POST;
            const int rc = bq.rc;                            // optimizePatch's epilogue (:364-381)
            mine.nlopt_rc = rc; mine.evals = bq.nevals; mine.score = bq.minf;
            if (rc >= 1 && rc <= 4) {
                double xf[3];
                bq3::result_x(bq, xf);
                set_center_norm(mine, K, xf);
                mine.status = HPMVS_OK;
            } else {
                mine.status = rc == bq3::R_ROUNDOFF_LIMITED ? HPMVS_FAIL_OPT_ROUNDOFF
                              : rc == bq3::R_MAXEVAL_REACHED ? HPMVS_FAIL_OPT_MAXEVAL : HPMVS_FAIL_OPT_OTHER;
            }
        }
        P.sstate[slot] = st;
    }
    __syncwarp();
    wf_push(P.eval_list, &P.ctl->eval_cnt, st == WS_NEED_EVAL, slot);
    wf_push(P.post_list, &P.ctl->post_cnt, st == WS_POST, slot);
}


}  // namespace hp
