// C-ABI implementation of the hpmvs_b200 engine (see include/hpmvs_b200.h for what each entry point replaces
// in the reference).  Host side: owns the HBM-resident scene (camera table, RGBX pyramids, covisibility CSR),
// batch staging buffers and the stream; launches the kernels in patch_kernels.cuh.
// There is deliberately no CPU implementation behind this ABI.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "patch_kernels.cuh"
#include "patch_kernels_wf.cuh"

// NVTX ranges around the host-visible steps of a batch (tracing): free when no profiler is attached
struct HpNvtxRange {
    explicit HpNvtxRange(const char* name) { nvtxRangePushA(name); }
    ~HpNvtxRange() { nvtxRangePop(); }
};
#define HP_NVTX(name) HpNvtxRange hp_nvtx_range_##__LINE__(name)

#define HP_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t err__ = (call);                                                                     \
        if (err__ != cudaSuccess) {                                                                     \
            snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s at %s:%d: %s", #call, __FILE__, __LINE__, \
                     cudaGetErrorString(err__));                                                        \
            return HPMVS_E_CUDA;                                                                        \
        }                                                                                               \
    } while (0)

static thread_local char g_last_cuda_error[512] = "";

namespace {

struct LevelImage {
    uchar4* data = nullptr;
    int w = 0, h = 0, pitch = 0;   // pitch in pixels
};

inline int pitch_for(int w) { return (w + 3) & ~3; }   // 16-byte aligned rows (TMA-compatible box loads)

// host-side f32 helpers in the same evaluation order as the device ones
inline float h_dot3(const float* a, const float* b) {
    const float p0 = a[0] * b[0], p1 = a[1] * b[1], p2 = a[2] * b[2];
    return p0 + (p1 + p2);
}
inline void h_normalized3(const float* a, float* o) {
    const float z = h_dot3(a, a);
    if (z > 0.0f) { const float s = sqrtf(z); o[0] = a[0] / s; o[1] = a[1] / s; o[2] = a[2] / s; }
    else { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
}

}  // namespace

enum { HP_RING = 4, HP_WF_PARTS = 8, HP_WF_BATCHES = 16 };   // wavefront: up to 16 launches in flight, each with its own slots
#ifndef HP_WF_DEFAULT_MODE
#define HP_WF_DEFAULT_MODE (-1)     // auto: wavefront kernels for large batches, the persistent kernel below wf_min_batch patches
#endif

struct hpmvs_engine {
    int device = 0;
    int sm_count = 0;
    hpmvs_options_t opt{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_kernel_ms = 0.0f;
    // scene
    int ncams = 0;
    std::vector<hp::DevCamera> h_cams;
    std::vector<std::vector<LevelImage>> images;   // [cam][level]
    std::vector<std::vector<float*>> depths;        // [cam][level] Scene::m_depths
    std::vector<std::vector<size_t>> depth_cells;   // [cam][level] floats allocated (a camera table with other image sizes reallocates)
    int* d_accept = nullptr;
    size_t cap_accept = 0;
    int ncc_tma = 0;                    // HPMVS_NCC_TMA=1: the scoring kernel stages its image windows with the bulk-async copy engine (A/B experiment)
    unsigned long long* d_tma_fallbacks = nullptr;
    unsigned char* d_dedup = nullptr;   // hash table + per-record slots of hpmvs_dedup_border_device
    size_t cap_dedup = 0;
    hp::DevCamera* d_cams = nullptr;
    bool cams_dirty = true;
    int* d_covis_off = nullptr;
    int* d_covis_ids = nullptr;
    bool have_covis = false;
    // batch staging
    hpmvs_patch_t* d_in = nullptr;
    hpmvs_patch_t* d_out = nullptr;
    float* d_inccs = nullptr;
    size_t cap_patches = 0, cap_inccs = 0;
    // hpmvs_optimize_batch_submit: a second staging set so that two host-buffer batches can be in flight
    struct Stage { hpmvs_patch_t* d_in = nullptr; hpmvs_patch_t* d_out = nullptr; size_t cap = 0; double* d_start = nullptr;
                   double* h_start = nullptr; size_t cap_start = 0; cudaEvent_t done = nullptr; } stage2[16];
    unsigned long long submit_seq = 0;
    int start_mode = 0;              // hpmvs_engine_set_start_mode
    double* d_start = nullptr;       // host-evaluated start angles of the batch (start_mode 1)
    double* h_start = nullptr;       // pinned staging for them
    size_t cap_start = 0;
    const double* next_start = nullptr;   // device array handed to the next launch (consumed by launch_optimize)
    unsigned char* d_stage = nullptr;
    size_t cap_stage = 0;
    int* d_work = nullptr;           // HP_RING work counters: launches on different streams may be in flight together
    cudaEvent_t slot_done[4] = {nullptr, nullptr, nullptr, nullptr};   // slot k is reused only after its previous launch finished
    cudaStream_t slot_stream[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned long long launch_seq = 0;
    unsigned long long* d_counters = nullptr;
    unsigned long long launches = 0;
    size_t smem_bytes = 0, smem_opt_bytes = 0;
    int blocks_per_sm = 0;
    int force_lanes = 0;
    int variant = 0;
    int pvariant = 0;
    int parked_mode = -1;            // -1 auto, 0 never, 1 always (HPMVS_PARKED)
    hp::BqSlot* d_pool_bq2[2] = {nullptr, nullptr};     // parked variant: two pool sets, so that two launches can overlap
    hp::LaneCtx* d_pool_ctx2[2] = {nullptr, nullptr};
    int pool_ctas2[2] = {0, 0};
    cudaEvent_t pool_done[2] = {nullptr, nullptr};
    unsigned long long parked_seq = 0;
    // wavefront form of the fused path (patch_kernels_wf.cuh): 0 = persistent kernels, 1 = per-phase kernels in a CUDA-graph WHILE loop,
    // 2 = the same kernels launched round by round from the host (debugging; the call blocks), -1 = 1 for batches >= wf_min_batch else 0
    int wf_mode = 0;
    int wf_min_batch = 32000;        // synchronous call (one batch at a time; profiles/r2_sync_call_latency_by_batch_size.txt): 24 k patches 44.2 ms
                                     // persistent vs 46.0 ms wavefront, 40 k: 64.1 vs 59.6 ms, 96 k: 156.6 vs 127.5 ms
    int wf_min_batch_async = 4000;   // asynchronous / device-resident calls (the caller keeps several batches in flight): with 8 in flight the
                                     // wavefront kernels win from ~5 k patches on (5.5 k: 10.1 vs 15.9 ms, 11 k: 16.0 vs 22.4 ms per step)
    int wf_split = 0;                // 1: advance phases A / T / B as three kernels, 0: one kernel with every phase (default: 61.5 vs 71.7 ms per city100 step)
    int wf_capacity = 0;             // slots in flight per wavefront context
    struct WfContext {
        int capacity = 0;
        hp::LaneCtx* ctx = nullptr; unsigned char* tiles = nullptr; double* fval = nullptr; int* sstate = nullptr;
        int* eval_list = nullptr; int* post_list = nullptr; hp::WfCtl* ctl = nullptr;
        hp::WfParams* d_params = nullptr;
        hp::WfParams* h_params[2] = {nullptr, nullptr};      // pinned ring; a slot is rewritten only after the launch that used it finished
        cudaEvent_t h_params_free[2] = {nullptr, nullptr};
        hp::WfCtl* h_ctl = nullptr;                           // pinned: the control block of the last finished launch (overrun check)
        unsigned long long seq = 0;
        unsigned long long* round_log = nullptr;              // HPMVS_WF_LOG: per-round {ns, evals, posts, dead} of the last launch
        cudaGraphConditionalHandle cond = 0;
    };
    // one launch = one CUDA graph with up to HP_WF_PARTS independent branches (fill kernel -> WHILE loop), one per sub-batch
    struct WfBatch {
        WfContext part[HP_WF_PARTS];
        cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
        cudaEvent_t done = nullptr;
        cudaStream_t last_stream = nullptr;
    } wfb[HP_WF_BATCHES];
    int wf_parts = 1;                // a batch is cut into up to this many sub-batches: their round loops interleave on the GPU
    unsigned long long wf_seq = 0;
    int wf_last = 0;                 // batch context of the most recent wavefront launch
    int wf_sampler_ctas_per_sm = 6;
    unsigned long long opt_calls = 0, last_concurrent_call = 0;   // asynchronous calls seen / the last one that found other streams busy
    unsigned long long wf_overruns = 0;
    std::mutex mu;
};

// Instantiated CTA shapes of the fused kernel: (optimizer warps, sampler warps).  The default is picked from
// measurements; HPMVS_CONFIG="ow,sw" selects another one for experiments.
struct KernelVariant {
    int ow, sw, lpw;
    void (*fn)(const hp::KParams);
    size_t smem;
};
#define HP_VARIANT(OW, SW, LPW) {OW, SW, LPW, hp::optimize_kernel<OW, SW, LPW>, sizeof(hp::CtaSharedT<OW, SW, LPW>)}
static const KernelVariant g_variants[] = {HP_VARIANT(2, 10, 32), HP_VARIANT(2, 12, 32)};   // other shapes measured: profiles/r1_*
static const int g_default_variant = 0;

// parked-slot variant of the same kernel (patch state pools in HBM/L2): used when a batch has more patches than the
// resident variant can keep in flight (64 per SM)
struct ParkedVariant {
    int ow, sw, lpw;
    void (*fn)(const hp::KParams);
    size_t smem;
};
#define HP_PVARIANT(OW, SW, LPW) {OW, SW, LPW, hp::optimize_kernel_parked<OW, SW, LPW>, sizeof(hp::CtaSharedP<OW, SW, LPW>)}
static const ParkedVariant g_pvariants[] = {HP_PVARIANT(2, 10, 32)};
static const int g_default_pvariant = 0;

static int ensure_patch_capacity(hpmvs_engine* e, size_t n) {
    if (n <= e->cap_patches) return 0;
    const size_t cap = n + n / 4 + 1024;
    if (e->d_in) cudaFree(e->d_in);
    if (e->d_out) cudaFree(e->d_out);
    e->d_in = e->d_out = nullptr;
    e->cap_patches = 0;
    HP_CUDA(cudaMalloc(&e->d_in, cap * sizeof(hpmvs_patch_t)));
    HP_CUDA(cudaMalloc(&e->d_out, cap * sizeof(hpmvs_patch_t)));
    e->cap_patches = cap;
    return 0;
}

static int sync_cameras(hpmvs_engine* e) {
    if (!e->cams_dirty) return 0;
    for (int c = 0; c < e->ncams; c++)
        for (int l = 0; l < HPMVS_LEVELS; l++) {
            const LevelImage& li = e->images[c][l];
            e->h_cams[c].img[l] = li.data;
            e->h_cams[c].pitch[l] = li.pitch;
            e->h_cams[c].depth[l] = (c < (int)e->depths.size() && l < (int)e->depths[c].size()) ? e->depths[c][l] : nullptr;
            e->h_cams[c].drows[l] = (int)(e->h_cams[c].h[l] / 2.0);     // Scene.cpp:76-77 (DEPTH_SUBSAMPLE is a double)
            e->h_cams[c].dcols[l] = (int)(e->h_cams[c].w[l] / 2.0);
        }
    HP_CUDA(cudaMemcpyAsync(e->d_cams, e->h_cams.data(), sizeof(hp::DevCamera) * e->ncams, cudaMemcpyHostToDevice, e->stream));
    HP_CUDA(cudaStreamSynchronize(e->stream));
    e->cams_dirty = false;
    return 0;
}

static int check_ready(hpmvs_engine* e) {
    if (!e || e->ncams <= 0 || !e->have_covis) return HPMVS_E_STATE;
    const int nl = e->opt.maxlevel + 1;
    for (int c = 0; c < e->ncams; c++)
        for (int l = 0; l < nl && l < HPMVS_LEVELS; l++)
            if (!e->images[c][l].data) return HPMVS_E_STATE;
    return 0;
}

// view ids index the camera table on the device: every host-buffer entry point rejects out-of-range ids instead of faulting there
static bool valid_view_ids(const hpmvs_engine* e, int n, const hpmvs_patch_t* in) {
    for (int i = 0; i < n; i++) {
        const int k = in[i].nimages < HPMVS_MAX_VIEWS ? in[i].nimages : HPMVS_MAX_VIEWS;
        for (int j = 0; j < k; j++)
            if (in[i].images[j] < 0 || in[i].images[j] >= e->ncams) return false;
    }
    return true;
}

static hp::KParams make_params(hpmvs_engine* e, const hpmvs_patch_t* d_in, hpmvs_patch_t* d_out, int n) {
    hp::KParams K{};
    K.cams = e->d_cams;
    K.ncams = e->ncams;
    K.covis_off = e->d_covis_off;
    K.covis_ids = e->d_covis_ids;
    K.opt = e->opt;
    // constants the reference forms with libm on the host (PatchOptimizer.cpp:485,129,239,184,398)
    K.cos_max_d = cos((double)e->opt.max_angle);
    K.cos_max_f = cosf(e->opt.max_angle);
    K.sort_thr = (float)(1.0f - cos(10.0 * M_PI / 180.0));
    K.angle_scale = (float)(M_PI / 48.0f);
    K.in = d_in;
    K.out = d_out;
    K.n = n;
    K.work_counter = e->d_work;
    K.counters = e->d_counters;
    K.lanes_per_warp = 1;
    return K;
}

extern "C" {

int hpmvs_abi_version(void) { return 1; }

const char* hpmvs_error_string(int code) {
    switch (code) {
    case 0: return "ok";
    case HPMVS_E_ARG: return "invalid argument";
    case HPMVS_E_CUDA: return g_last_cuda_error[0] ? g_last_cuda_error : "CUDA error";
    case HPMVS_E_STATE: return "scene incomplete: cameras, all pyramid levels and covisibility must be uploaded first";
    case HPMVS_E_NODEVICE: return "no CUDA device (this engine has no CPU fallback)";
    default: return "unknown error";
    }
}

int hpmvs_engine_create(const hpmvs_options_t* opt, int device, hpmvs_engine_t** out) {
    if (!opt || !out) return HPMVS_E_ARG;
    *out = nullptr;
    if (opt->maxlevel < 1 || opt->maxlevel >= HPMVS_LEVELS || opt->min_images_per_patch < 1) return HPMVS_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return HPMVS_E_NODEVICE;
    if (device < 0 || device >= ndev) return HPMVS_E_ARG;
    HP_CUDA(cudaSetDevice(device));
    hpmvs_engine* e = new hpmvs_engine;
    e->device = device;
    e->opt = *opt;
    cudaDeviceProp prop;
    HP_CUDA(cudaGetDeviceProperties(&prop, device));
    e->sm_count = prop.multiProcessorCount;
    HP_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    HP_CUDA(cudaEventCreate(&e->ev0));
    HP_CUDA(cudaEventCreate(&e->ev1));
    HP_CUDA(cudaMalloc(&e->d_work, HP_RING * sizeof(int)));
    for (int i = 0; i < HP_RING; i++) HP_CUDA(cudaEventCreateWithFlags(&e->slot_done[i], cudaEventDisableTiming));
    HP_CUDA(cudaMalloc(&e->d_counters, 16 * sizeof(unsigned long long)));
    HP_CUDA(cudaMemset(e->d_counters, 0, 16 * sizeof(unsigned long long)));
    e->smem_bytes = sizeof(hp::NccWarp) * hp::WARPS_PER_BLOCK;             // ncc_kernel
    e->variant = g_default_variant;
    if (const char* cfg = getenv("HPMVS_CONFIG")) {
        int ow = 0, sw = 0, lpw = 0;
        if (sscanf(cfg, "%d,%d,%d", &ow, &sw, &lpw) == 3)
            for (size_t i = 0; i < sizeof(g_variants) / sizeof(g_variants[0]); i++)
                if (g_variants[i].ow == ow && g_variants[i].sw == sw && g_variants[i].lpw == lpw) e->variant = (int)i;
    }
    e->smem_opt_bytes = g_variants[e->variant].smem;                        // optimize_kernel: one CTA per SM
    HP_CUDA(cudaFuncSetAttribute(g_variants[e->variant].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_opt_bytes));
    HP_CUDA(cudaFuncSetAttribute(hp::ncc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
    HP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->blocks_per_sm, hp::ncc_kernel, hp::WARPS_PER_BLOCK * 32,
                                                          e->smem_bytes));
    if (e->blocks_per_sm < 1) e->blocks_per_sm = 1;
    if (const char* fl = getenv("HPMVS_FORCE_LANES")) e->force_lanes = atoi(fl);
    if (const char* tm = getenv("HPMVS_NCC_TMA")) e->ncc_tma = atoi(tm);
    if (e->ncc_tma) {
        HP_CUDA(cudaFuncSetAttribute(hp::ncc_kernel_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(hp::NccWarpTma) * hp::WARPS_PER_BLOCK)));
        HP_CUDA(cudaMalloc(&e->d_tma_fallbacks, sizeof(unsigned long long)));
        HP_CUDA(cudaMemset(e->d_tma_fallbacks, 0, sizeof(unsigned long long)));
    }
    if (const char* pk = getenv("HPMVS_PARKED")) e->parked_mode = atoi(pk);
    e->pvariant = g_default_pvariant;
    if (const char* cfg = getenv("HPMVS_PCONFIG")) {
        int ow = 0, sw = 0, lpw = 0;
        if (sscanf(cfg, "%d,%d,%d", &ow, &sw, &lpw) == 3)
            for (size_t i = 0; i < sizeof(g_pvariants) / sizeof(g_pvariants[0]); i++)
                if (g_pvariants[i].ow == ow && g_pvariants[i].sw == sw && g_pvariants[i].lpw == lpw) e->pvariant = (int)i;
    }
    HP_CUDA(cudaFuncSetAttribute(g_pvariants[e->pvariant].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_pvariants[e->pvariant].smem));
    e->wf_mode = HP_WF_DEFAULT_MODE;
    if (const char* wf = getenv("HPMVS_WF")) e->wf_mode = atoi(wf);
    if (const char* ws = getenv("HPMVS_WF_SPLIT")) e->wf_split = atoi(ws);
    if (const char* wm = getenv("HPMVS_WF_MIN_BATCH")) e->wf_min_batch = e->wf_min_batch_async = atoi(wm);
    {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hp::wf_eval_kernel, hp::WF_SAMPLER_WARPS * 32,
                                                          sizeof(hp::NccWarp) * hp::WF_SAMPLER_WARPS) == cudaSuccess && occ >= 1)
            e->wf_sampler_ctas_per_sm = occ > 8 ? 8 : occ;
        if (const char* sc = getenv("HPMVS_WF_SAMPLER_CTAS")) e->wf_sampler_ctas_per_sm = atoi(sc);
    }
    e->wf_capacity = e->sm_count * 12 * 32;                                 // one full wave of advance threads (12 warps per SM)
    if (const char* wp = getenv("HPMVS_WF_PARTS")) e->wf_parts = atoi(wp);
    if (e->wf_parts < 1) e->wf_parts = 1;
    if (e->wf_parts > HP_WF_PARTS) e->wf_parts = HP_WF_PARTS;
    e->wf_capacity = e->wf_capacity / e->wf_parts;
    if (const char* wc = getenv("HPMVS_WF_SLOTS")) e->wf_capacity = atoi(wc);
    e->wf_capacity = (e->wf_capacity + hp::WF_ADV_THREADS - 1) / hp::WF_ADV_THREADS * hp::WF_ADV_THREADS;
    if (e->wf_capacity < hp::WF_ADV_THREADS) e->wf_capacity = hp::WF_ADV_THREADS;
    *out = e;
    return 0;
}

void hpmvs_engine_destroy(hpmvs_engine_t* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    // batches submitted on caller streams (hpmvs_optimize_batch_submit / _device) may still be running on our buffers
    for (int i = 0; i < HP_RING; i++) if (e->slot_done[i]) cudaEventSynchronize(e->slot_done[i]);
    for (auto& st : e->stage2) if (st.done) cudaEventSynchronize(st.done);
    for (auto& cam : e->images)
        for (auto& li : cam)
            if (li.data) cudaFree(li.data);
    for (auto& cam : e->depths)
        for (auto* d : cam)
            if (d) cudaFree(d);
    cudaFree(e->d_accept); cudaFree(e->d_dedup); cudaFree(e->d_tma_fallbacks);
    for (int i = 0; i < 2; i++) if (e->pool_done[i]) cudaEventDestroy(e->pool_done[i]);
    cudaFree(e->d_cams); cudaFree(e->d_covis_off); cudaFree(e->d_covis_ids);
    cudaFree(e->d_in); cudaFree(e->d_out); cudaFree(e->d_inccs); cudaFree(e->d_stage);
    cudaFree(e->d_start); if (e->h_start) cudaFreeHost(e->h_start);
    for (auto& st : e->stage2) { cudaFree(st.d_in); cudaFree(st.d_out); cudaFree(st.d_start); if (st.h_start) cudaFreeHost(st.h_start); if (st.done) cudaEventDestroy(st.done); }
    cudaFree(e->d_work); cudaFree(e->d_counters);
    for (int i = 0; i < HP_RING; i++) if (e->slot_done[i]) cudaEventDestroy(e->slot_done[i]);
    for (int i = 0; i < 2; i++) { cudaFree(e->d_pool_bq2[i]); cudaFree(e->d_pool_ctx2[i]); }
    for (auto& b : e->wfb) {
        if (b.done) { cudaEventSynchronize(b.done); cudaEventDestroy(b.done); }
        if (b.exec) cudaGraphExecDestroy(b.exec);
        if (b.graph) cudaGraphDestroy(b.graph);
        for (auto& w : b.part) {
            cudaFree(w.ctx); cudaFree(w.tiles); cudaFree(w.fval); cudaFree(w.sstate); cudaFree(w.eval_list); cudaFree(w.post_list);
            cudaFree(w.ctl); cudaFree(w.d_params); cudaFree(w.round_log);
            for (int i = 0; i < 2; i++) { if (w.h_params[i]) cudaFreeHost(w.h_params[i]); if (w.h_params_free[i]) cudaEventDestroy(w.h_params_free[i]); }
            if (w.h_ctl) cudaFreeHost(w.h_ctl);
        }
    }
    cudaEventDestroy(e->ev0); cudaEventDestroy(e->ev1);
    cudaStreamDestroy(e->stream);
    delete e;
}

int hpmvs_engine_set_cameras(hpmvs_engine_t* e, int n, const hpmvs_camera_t* cams) {
    if (!e || n <= 0 || n > 65535 || !cams) return HPMVS_E_ARG;      // view ids are carried as 16 bits on the device
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    if (n != e->ncams) {
        for (auto& cam : e->images)
            for (auto& li : cam)
                if (li.data) cudaFree(li.data);
        e->images.assign(n, std::vector<LevelImage>(HPMVS_LEVELS));
        if (e->d_cams) cudaFree(e->d_cams);
        e->d_cams = nullptr;
        HP_CUDA(cudaMalloc(&e->d_cams, sizeof(hp::DevCamera) * n));
        e->have_covis = false;
    }
    e->ncams = n;
    e->h_cams.assign(n, hp::DevCamera{});
    for (int c = 0; c < n; c++) {
        hp::DevCamera& d = e->h_cams[c];
        const hpmvs_camera_t& s = cams[c];
        for (int l = 0; l < HPMVS_LEVELS; l++) {
            memcpy(d.P[l], s.P[l], sizeof(float) * 12);
            d.w[l] = s.width[l];
            d.h[l] = s.height[l];
        }
        for (int i = 0; i < 4; i++) d.center[i] = s.center[i];
        for (int i = 0; i < 3; i++) { d.xaxis[i] = s.xaxis[i]; d.yaxis[i] = s.yaxis[i]; d.zaxis[i] = s.zaxis[i]; }
        d.xaxis[3] = d.yaxis[3] = d.zaxis[3] = 0.0f;
        h_normalized3(s.xaxis, d.nx); h_normalized3(s.yaxis, d.ny); h_normalized3(s.zaxis, d.nz);
        d.nx[3] = d.ny[3] = d.nz[3] = 0.0f;
        d.ksum = s.k00 + s.k11;
        d.nlevels = e->opt.maxlevel + 1;
    }
    e->cams_dirty = true;
    return 0;
}

static int ensure_level(hpmvs_engine* e, int cam, int level, int w, int h) {
    LevelImage& li = e->images[cam][level];
    if (li.data && li.w == w && li.h == h) return 0;
    if (li.data) cudaFree(li.data);
    li = LevelImage{};
    const int pitch = pitch_for(w);
    // two spare rows: bilinear taps of the reference read row ly+1 / column lx+1 (guarded by the 3 px margin)
    HP_CUDA(cudaMalloc(&li.data, sizeof(uchar4) * (size_t)pitch * (h + 2)));
    HP_CUDA(cudaMemsetAsync(li.data, 0, sizeof(uchar4) * (size_t)pitch * (h + 2), e->stream));
    li.w = w; li.h = h; li.pitch = pitch;
    e->cams_dirty = true;
    return 0;
}

int hpmvs_engine_upload_image(hpmvs_engine_t* e, int cam, int level, const uint8_t* rgb, int w, int h, size_t pitch_bytes) {
    if (!e || !rgb || cam < 0 || cam >= e->ncams || level < 0 || level >= HPMVS_LEVELS || w <= 0 || h <= 0 ||
        pitch_bytes < (size_t)3 * w)
        return HPMVS_E_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    int rc = ensure_level(e, cam, level, w, h);
    if (rc) return rc;
    const size_t need = (size_t)3 * w * h;
    if (need > e->cap_stage) {
        if (e->d_stage) cudaFree(e->d_stage);
        e->d_stage = nullptr; e->cap_stage = 0;
        HP_CUDA(cudaMalloc(&e->d_stage, need));
        e->cap_stage = need;
    }
    HP_CUDA(cudaMemcpy2DAsync(e->d_stage, (size_t)3 * w, rgb, pitch_bytes, (size_t)3 * w, h, cudaMemcpyHostToDevice, e->stream));
    const LevelImage& li = e->images[cam][level];
    dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
    hp::rgb_to_rgbx_kernel<<<grd, blk, 0, e->stream>>>(e->d_stage, w, h, li.data, li.pitch);
    e->launches++;
    HP_CUDA(cudaGetLastError());
    HP_CUDA(cudaStreamSynchronize(e->stream));
    if (e->h_cams[cam].w[level] != w || e->h_cams[cam].h[level] != h) return HPMVS_E_ARG;
    return 0;
}

// Image::load's undistortion step (Image.cpp:51-53) on the device: uploads the DISTORTED level-0 image and writes the undistorted one
// as level 0 of view `cam`; r == 0 is a plain upload.
int hpmvs_engine_upload_image_undistort(hpmvs_engine_t* e, int cam, const uint8_t* rgb, int w, int h, size_t pitch_bytes, double f, double r) {
    if ((float)r == 0.0f) return hpmvs_engine_upload_image(e, cam, 0, rgb, w, h, pitch_bytes);
    if (!e || !rgb || cam < 0 || cam >= e->ncams || w <= 0 || h <= 0 || pitch_bytes < (size_t)3 * w) return HPMVS_E_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    int rc = ensure_level(e, cam, 0, w, h);
    if (rc) return rc;
    const size_t need = (size_t)3 * w * h;
    if (need > e->cap_stage) {
        if (e->d_stage) cudaFree(e->d_stage);
        e->d_stage = nullptr; e->cap_stage = 0;
        HP_CUDA(cudaMalloc(&e->d_stage, need));
        e->cap_stage = need;
    }
    HP_CUDA(cudaMemcpy2DAsync(e->d_stage, (size_t)3 * w, rgb, pitch_bytes, (size_t)3 * w, h, cudaMemcpyHostToDevice, e->stream));
    const LevelImage& li = e->images[cam][0];
    dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
    hp::undistort_kernel<<<grd, blk, 0, e->stream>>>(e->d_stage, w, h, (float)f, (float)r, li.data, li.pitch);
    e->launches++;
    HP_CUDA(cudaGetLastError());
    HP_CUDA(cudaStreamSynchronize(e->stream));
    if (e->h_cams[cam].w[0] != w || e->h_cams[cam].h[0] != h) return HPMVS_E_ARG;
    return 0;
}

int hpmvs_engine_build_pyramid(hpmvs_engine_t* e, int cam) {
    if (!e || cam < 0 || cam >= e->ncams) return HPMVS_E_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    if (!e->images[cam][0].data) return HPMVS_E_STATE;
    for (int l = 1; l <= e->opt.maxlevel; l++) {
        const LevelImage src = e->images[cam][l - 1];
        const int w2 = src.w / 2, h2 = src.h / 2;
        if (w2 <= 0 || h2 <= 0) return HPMVS_E_ARG;
        int rc = ensure_level(e, cam, l, w2, h2);
        if (rc) return rc;
        const LevelImage& dst = e->images[cam][l];
        dim3 blk(32, 8), grd((w2 + 31) / 32, (h2 + 7) / 8);
        hp::half_xy_kernel<<<grd, blk, 0, e->stream>>>(src.data, src.pitch, src.w, src.h, dst.data, dst.pitch, w2, h2);
        e->launches++;
        HP_CUDA(cudaGetLastError());
        if (e->h_cams[cam].w[l] != w2 || e->h_cams[cam].h[l] != h2) return HPMVS_E_ARG;
    }
    HP_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

int hpmvs_engine_download_image(hpmvs_engine_t* e, int cam, int level, uint8_t* rgb, size_t pitch_bytes) {
    if (!e || !rgb || cam < 0 || cam >= e->ncams || level < 0 || level >= HPMVS_LEVELS) return HPMVS_E_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    const LevelImage& li = e->images[cam][level];
    if (!li.data) return HPMVS_E_STATE;
    if (pitch_bytes < (size_t)3 * li.w) return HPMVS_E_ARG;
    const size_t need = (size_t)3 * li.w * li.h;
    if (need > e->cap_stage) {
        if (e->d_stage) cudaFree(e->d_stage);
        e->d_stage = nullptr; e->cap_stage = 0;
        HP_CUDA(cudaMalloc(&e->d_stage, need));
        e->cap_stage = need;
    }
    dim3 blk(32, 8), grd((li.w + 31) / 32, (li.h + 7) / 8);
    hp::rgbx_to_rgb_kernel<<<grd, blk, 0, e->stream>>>(li.data, li.pitch, li.w, li.h, e->d_stage);
    e->launches++;
    HP_CUDA(cudaGetLastError());
    HP_CUDA(cudaMemcpy2DAsync(rgb, pitch_bytes, e->d_stage, (size_t)3 * li.w, (size_t)3 * li.w, li.h, cudaMemcpyDeviceToHost, e->stream));
    HP_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

int hpmvs_engine_set_covis(hpmvs_engine_t* e, const int32_t* offsets, const int32_t* ids) {
    if (!e || !offsets || e->ncams <= 0) return HPMVS_E_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    const int n = e->ncams;
    const int total = offsets[n];
    if (total < 0 || (total > 0 && !ids)) return HPMVS_E_ARG;
    for (int i = 0; i < n; i++) if (offsets[i + 1] < offsets[i]) return HPMVS_E_ARG;
    for (int i = 0; i < total; i++) if (ids[i] < 0 || ids[i] >= n) return HPMVS_E_ARG;
    if (e->d_covis_off) cudaFree(e->d_covis_off);
    if (e->d_covis_ids) cudaFree(e->d_covis_ids);
    e->d_covis_off = e->d_covis_ids = nullptr;
    HP_CUDA(cudaMalloc(&e->d_covis_off, sizeof(int) * (n + 1)));
    HP_CUDA(cudaMalloc(&e->d_covis_ids, sizeof(int) * (total > 0 ? total : 1)));
    HP_CUDA(cudaMemcpy(e->d_covis_off, offsets, sizeof(int) * (n + 1), cudaMemcpyHostToDevice));
    if (total > 0) HP_CUDA(cudaMemcpy(e->d_covis_ids, ids, sizeof(int) * total, cudaMemcpyHostToDevice));
    e->have_covis = true;
    return 0;
}

// parametersFromCenterNorm's two angles (src/hpmvs/PatchOptimizer.cpp:421-443) for every input patch with the HOST's libm:
// std::asin(float), std::cos(double), std::acos(double) - the reference's own calls, so that the engine starts every
// optimisation from exactly the reference's x[1], x[2] whatever libm the reference was linked against.
static void host_start_parameters(hpmvs_engine* e, int n, const hpmvs_patch_t* in, double* out) {
    HP_NVTX("hpmvs:host_start_parameters (libm asin/acos per patch)");
    const double lb = -23.99999, ub = 23.99999;                 // optimizePatch's bounds (:333-340)
    const float angle_scale = (float)(M_PI / 48.0f);            // :398
    for (int i = 0; i < n; i++) {
        out[2 * i] = 0.0; out[2 * i + 1] = 0.0;
        if (in[i].nimages < 1 || in[i].images[0] < 0 || in[i].images[0] >= e->ncams) continue;
        const hp::DevCamera& cam = e->h_cams[in[i].images[0]];
        const float* nn = in[i].normal;
        const float fx = h_dot3(cam.nx, nn), fy = h_dot3(cam.ny, nn), fz = h_dot3(cam.nz, nn);
        double x1, x2 = std::asin(fy);                            // float overload
        const float cosb = std::cos(std::max(-1.0, std::min(1.0, x2)));
        if (cosb == 0.0) x1 = 0.0;
        else {
            const double sina = fx / cosb;
            const double cosa = -fz / cosb;
            x1 = std::acos(std::min(1.0, std::max(-1.0, cosa)));
            if (sina < 0.0) x1 = -x1;
        }
        x1 /= angle_scale; x2 /= angle_scale;
        out[2 * i] = std::min(ub, std::max(lb, x1));
        out[2 * i + 1] = std::min(ub, std::max(lb, x2));
    }
}

// ---- wavefront form: per-phase kernels in a CUDA-graph WHILE loop (patch_kernels_wf.cuh) ---------------------------------------------
static int wf_ensure_context(hpmvs_engine* e, hpmvs_engine::WfContext& w) {
    if (w.capacity) return 0;
    const size_t cap = (size_t)e->wf_capacity;
    HP_CUDA(cudaMalloc(&w.ctx, cap * sizeof(hp::LaneCtx)));
    HP_CUDA(cudaMalloc(&w.tiles, cap / 32 * sizeof(bq3::StateTile)));
    HP_CUDA(cudaMalloc(&w.fval, cap * sizeof(double)));
    HP_CUDA(cudaMalloc(&w.sstate, cap * sizeof(int)));
    HP_CUDA(cudaMalloc(&w.eval_list, cap * sizeof(int)));
    HP_CUDA(cudaMalloc(&w.post_list, cap * sizeof(int)));
    HP_CUDA(cudaMalloc(&w.ctl, sizeof(hp::WfCtl)));
    HP_CUDA(cudaMalloc(&w.d_params, sizeof(hp::WfParams)));
    for (int i = 0; i < 2; i++) {
        HP_CUDA(cudaMallocHost(&w.h_params[i], sizeof(hp::WfParams)));
        HP_CUDA(cudaEventCreateWithFlags(&w.h_params_free[i], cudaEventDisableTiming));
    }
    HP_CUDA(cudaMallocHost(&w.h_ctl, sizeof(hp::WfCtl)));
    memset(w.h_ctl, 0, sizeof(hp::WfCtl));
    if (getenv("HPMVS_WF_LOG")) HP_CUDA(cudaMalloc(&w.round_log, sizeof(unsigned long long) * 4 * 65536));
    w.capacity = (int)cap;
    return 0;
}

struct WfLaunchShape { dim3 grid_s, grid_p, grid_a; size_t smem_s; };
static WfLaunchShape wf_shape(const hpmvs_engine* e, const hpmvs_engine::WfContext& w) {
    WfLaunchShape sh;
    sh.smem_s = sizeof(hp::NccWarp) * hp::WF_SAMPLER_WARPS;
    // as many sampler CTAs as are RESIDENT per SM (4 warps x ~9 KB scratch each: 6 per SM): the scoring kernel strides statically over its
    // list, so a CTA that has to wait for a free slot would start its share only after the others have finished theirs
    int per_sm = e->wf_sampler_ctas_per_sm / e->wf_parts;
    if (per_sm < 2) per_sm = 2;
    sh.grid_s = dim3((unsigned)(e->sm_count * per_sm));
    sh.grid_p = dim3((unsigned)(e->sm_count * (per_sm > 4 ? 4 : per_sm)));   // post pass: dynamic tickets; most rounds it has nothing to do
    sh.grid_a = dim3((unsigned)(w.capacity / hp::WF_ADV_THREADS));
    return sh;
}

// one round of the loop, enqueued on `s` (host-loop mode) - the graph body holds the same nodes in the same order
static void wf_enqueue_round(const hpmvs_engine* e, const hpmvs_engine::WfContext& w, cudaStream_t s) {
    const WfLaunchShape sh = wf_shape(e, w);
    if (e->wf_split) {
        hp::wf_advance_kernel<bq3::PH_A><<<sh.grid_a, hp::WF_ADV_THREADS, 0, s>>>(w.d_params);
        hp::wf_advance_kernel<bq3::PH_T><<<sh.grid_a, hp::WF_ADV_THREADS, 0, s>>>(w.d_params);
        hp::wf_advance_kernel<bq3::PH_B><<<sh.grid_a, hp::WF_ADV_THREADS, 0, s>>>(w.d_params);
    } else {
        hp::wf_advance_kernel<bq3::PH_ALL><<<sh.grid_a, hp::WF_ADV_THREADS, 0, s>>>(w.d_params);
    }
    hp::wf_eval_kernel<<<sh.grid_s, hp::WF_SAMPLER_WARPS * 32, sh.smem_s, s>>>(w.d_params);
    hp::wf_post_kernel<false><<<sh.grid_p, hp::WF_SAMPLER_WARPS * 32, sh.smem_s, s>>>(w.d_params);
}

static int wf_build_graph(hpmvs_engine* e, hpmvs_engine::WfBatch& b) {
    if (b.exec) return 0;
    HP_CUDA(cudaGraphCreate(&b.graph, 0));
    for (int k = 0; k < e->wf_parts; k++) {
        hpmvs_engine::WfContext& w = b.part[k];
        const WfLaunchShape sh = wf_shape(e, w);
        void* args[1] = {(void*)&w.d_params};
        auto knode = [&](cudaGraph_t g, cudaGraphNode_t* node, const cudaGraphNode_t* dep, void* fn, dim3 grid, unsigned block, size_t smem) {
            cudaKernelNodeParams kp{};
            kp.func = fn; kp.gridDim = grid; kp.blockDim = dim3(block); kp.sharedMemBytes = (unsigned)smem; kp.kernelParams = args; kp.extra = nullptr;
            return cudaGraphAddKernelNode(node, g, dep, dep ? 1 : 0, &kp);
        };
        cudaGraphNode_t fill, loop;
        HP_CUDA(knode(b.graph, &fill, nullptr, (void*)hp::wf_post_kernel<true>, sh.grid_s, hp::WF_SAMPLER_WARPS * 32, sh.smem_s));
        HP_CUDA(cudaGraphConditionalHandleCreate(&w.cond, b.graph, 1, cudaGraphCondAssignDefault));
        cudaGraphNodeParams cp{};
        cp.type = cudaGraphNodeTypeConditional;
        cp.conditional.handle = w.cond;
        cp.conditional.type = cudaGraphCondTypeWhile;
        cp.conditional.size = 1;
        HP_CUDA(cudaGraphAddNode(&loop, b.graph, &fill, 1, &cp));
        cudaGraph_t body = cp.conditional.phGraph_out[0];
        cudaGraphNode_t prev, node;
        bool first = true;
        auto chain = [&](void* fn, dim3 grid, unsigned block, size_t smem) {
            const cudaError_t err = knode(body, &node, first ? nullptr : &prev, fn, grid, block, smem);
            prev = node; first = false;
            return err;
        };
        if (e->wf_split) {
            HP_CUDA(chain((void*)hp::wf_advance_kernel<bq3::PH_A>, sh.grid_a, hp::WF_ADV_THREADS, 0));
            HP_CUDA(chain((void*)hp::wf_advance_kernel<bq3::PH_T>, sh.grid_a, hp::WF_ADV_THREADS, 0));
            HP_CUDA(chain((void*)hp::wf_advance_kernel<bq3::PH_B>, sh.grid_a, hp::WF_ADV_THREADS, 0));
        } else {
            HP_CUDA(chain((void*)hp::wf_advance_kernel<bq3::PH_ALL>, sh.grid_a, hp::WF_ADV_THREADS, 0));
        }
        HP_CUDA(chain((void*)hp::wf_eval_kernel, sh.grid_s, hp::WF_SAMPLER_WARPS * 32, sh.smem_s));
        HP_CUDA(chain((void*)hp::wf_post_kernel<false>, sh.grid_p, hp::WF_SAMPLER_WARPS * 32, sh.smem_s));
    }
    HP_CUDA(cudaGraphInstantiate(&b.exec, b.graph, 0));
    return 0;
}

// parameters + control block of one sub-batch, enqueued on `s` ahead of the graph launch
static int wf_prepare_part(hpmvs_engine* e, hpmvs_engine::WfContext& w, int n, const hpmvs_patch_t* d_in, hpmvs_patch_t* d_out,
                           const double* d_start, cudaStream_t s) {
    const int ring = (int)(w.seq++ % 2);
    HP_CUDA(cudaEventSynchronize(w.h_params_free[ring]));           // the launch that read this pinned block has finished
    if (w.h_ctl->overrun) { e->wf_overruns++; w.h_ctl->overrun = 0; }
    hp::WfParams& P = *w.h_params[ring];
    memset(&P, 0, sizeof(P));
    P.K = make_params(e, d_in, d_out, n);
    P.K.work_counter = &w.ctl->work_counter;
    P.K.start = d_start;
    int M = (n + 31) / 32 * 32;
    if (M > w.capacity) M = w.capacity;
    P.M = M;
    P.max_rounds = 4096 + 8 * (int)(((long long)n + M) / (M + 1)) * 1024;    // safety net: a patch needs <= ~1000 evaluations
    P.ctx = w.ctx; P.tiles = w.tiles; P.fval = w.fval; P.sstate = w.sstate; P.eval_list = w.eval_list; P.post_list = w.post_list;
    P.ctl = w.ctl;
    P.cond = w.cond;
    P.use_cond = (e->wf_mode != 2) ? 1 : 0;
    P.kernels_per_round = e->wf_split ? 5 : 3;
    P.round_log = w.round_log; P.round_log_cap = 65536;
    HP_CUDA(cudaMemcpyAsync(w.d_params, &P, sizeof(P), cudaMemcpyHostToDevice, s));
    HP_CUDA(cudaMemsetAsync(w.ctl, 0, sizeof(hp::WfCtl), s));
    if (M > 0) HP_CUDA(cudaMemsetAsync(w.sstate, 0, sizeof(int) * (size_t)M, s));
    HP_CUDA(cudaEventRecord(w.h_params_free[ring], s));
    return 0;
}

// Choose an in-flight slot (wavefront batch context / host-buffer staging set): an idle one that already owns its buffers, else a fresh
// one (buffers + graph are created on first use: tens of milliseconds), else the oldest (the caller's stream then waits for it).  The
// number of slots that ever get created is thus the number of launches the caller really keeps in flight.
}  // extern "C"
template <class Slot>
static int pick_slot(Slot* a, int n, unsigned long long& rr) {
    int fresh = -1;
    for (int i = 0; i < n; i++) {
        if (!a[i].done) { if (fresh < 0) fresh = i; continue; }
        if (cudaEventQuery(a[i].done) == cudaSuccess) return i;
    }
    if (fresh >= 0) return fresh;
    return (int)(rr++ % (unsigned long long)n);
}
extern "C" {

// One launch = one CUDA graph: fill kernel -> WHILE loop over the rounds (wf_build_graph).  HPMVS_WF_PARTS > 1 cuts the batch into
// contiguous sub-batches with their own loops as parallel branches of that graph (measured: no gain over one loop, default 1;
// what pays is keeping whole launches in flight on different streams, each on its own WfBatch).
static int launch_wavefront(hpmvs_engine* e, int n, const hpmvs_patch_t* d_in, hpmvs_patch_t* d_out, cudaStream_t s) {
    const double* d_start = e->next_start;
    e->next_start = nullptr;
    e->wf_last = pick_slot(e->wfb, (int)HP_WF_BATCHES, e->wf_seq);
    hpmvs_engine::WfBatch& b = e->wfb[e->wf_last];
    int rc;
    for (int k = 0; k < e->wf_parts; k++) if ((rc = wf_ensure_context(e, b.part[k]))) return rc;
    if (!b.done) HP_CUDA(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
    const bool graph_mode = e->wf_mode != 2;
    if (graph_mode && (rc = wf_build_graph(e, b))) return rc;
    int parts = graph_mode ? e->wf_parts : 1;                        // host-loop mode (debugging) drives one loop
    int used = parts;
    while (used > 1 && n / used < 2048) used--;                      // small batches are not cut further
    HP_CUDA(cudaStreamWaitEvent(s, b.done, 0));                      // the previous launch on this batch context has left its buffers
    HP_CUDA(cudaEventRecord(e->ev0, s));
    for (int k = 0; k < parts; k++) {
        const int a = k < used ? (int)((long long)n * k / used) : n, z = k < used ? (int)((long long)n * (k + 1) / used) : n;
        if ((rc = wf_prepare_part(e, b.part[k], z - a, d_in + a, d_out + a, d_start ? d_start + 2 * (size_t)a : nullptr, s))) return rc;
    }
    if (graph_mode) {
        HP_CUDA(cudaGraphLaunch(b.exec, s));
        e->launches += (unsigned long long)parts;                    // the fill kernels; the kernels of the rounds are counted on the device (wf_sched)
    } else {
        hpmvs_engine::WfContext& w = b.part[0];
        const WfLaunchShape sh = wf_shape(e, w);
        hp::wf_post_kernel<true><<<sh.grid_s, hp::WF_SAMPLER_WARPS * 32, sh.smem_s, s>>>(w.d_params);
        e->launches++;
        const int max_rounds = w.h_params[(w.seq + 1) % 2]->max_rounds;
        for (int live = 1, guard = 0; live && guard < max_rounds; guard += 4) {
            for (int r = 0; r < 4; r++) wf_enqueue_round(e, w, s);
            HP_CUDA(cudaMemcpyAsync(w.h_ctl, w.ctl, sizeof(hp::WfCtl), cudaMemcpyDeviceToHost, s));
            HP_CUDA(cudaStreamSynchronize(s));
            live = w.h_ctl->live;
        }
    }
    for (int k = 0; k < parts; k++) HP_CUDA(cudaMemcpyAsync(b.part[k].h_ctl, b.part[k].ctl, sizeof(hp::WfCtl), cudaMemcpyDeviceToHost, s));
    HP_CUDA(cudaEventRecord(e->ev1, s));
    HP_CUDA(cudaEventRecord(b.done, s));
    b.last_stream = s;
    HP_CUDA(cudaGetLastError());
    return 0;
}

static int launch_optimize(hpmvs_engine* e, int n, const hpmvs_patch_t* d_in, hpmvs_patch_t* d_out, cudaStream_t s, bool async_call = true) {
    HP_NVTX("hpmvs:launch_optimize");
    int rc = check_ready(e);
    if (rc) return rc;
    rc = sync_cameras(e);
    if (rc) return rc;
    if (async_call && e->wf_mode < 0) {
        // An asynchronous caller that keeps several batches in flight ON DIFFERENT STREAMS gets the throughput-optimal choice (wavefront
        // kernels from 4 k patches on); one that runs a batch at a time (nothing in flight, or only earlier launches on this very
        // stream, which are serialised anyway) gets the latency-optimal one (persistent kernel below 32 k patches).
        bool concurrent = false;
        for (auto& b : e->wfb) if (b.done && b.last_stream != s && cudaEventQuery(b.done) != cudaSuccess) concurrent = true;
        for (int i = 0; i < HP_RING; i++)
            if (e->slot_done[i] && e->slot_stream[i] && e->slot_stream[i] != s && cudaEventQuery(e->slot_done[i]) != cudaSuccess) concurrent = true;
        (void)cudaGetLastError();
        // sticky for a while: the first launch after a drained pipeline (start of a timed region, a synchronisation point) belongs to
        // the same multi-stream caller as the 64 launches before it
        e->opt_calls++;
        if (concurrent) e->last_concurrent_call = e->opt_calls;
        if (!concurrent && (e->last_concurrent_call == 0 || e->opt_calls - e->last_concurrent_call > 64)) async_call = false;
    }
    if (e->wf_mode > 0 || (e->wf_mode < 0 && n >= (async_call ? e->wf_min_batch_async : e->wf_min_batch))) return launch_wavefront(e, n, d_in, d_out, s);
    // launches on different streams overlap (a CTA of the next launch starts on an SM as soon as the previous launch's CTA
    // there has drained its slots): every launch gets its own work counter from a small ring; a ring slot is reused only after
    // the launch that used it last has completed
    const int slot = (int)(e->launch_seq++ % HP_RING);
    HP_CUDA(cudaStreamWaitEvent(s, e->slot_done[slot], 0));
    HP_CUDA(cudaMemsetAsync(e->d_work + slot, 0, sizeof(int), s));
    hp::KParams K = make_params(e, d_in, d_out, n);
    K.work_counter = e->d_work + slot;
    // persistent grid: one warp-specialised CTA per SM.  Each optimizer warp keeps `lanes` patches in flight;
    // small batches are spread over all SMs first (lanes < 32) so that no SM idles.
    const KernelVariant& V = g_variants[e->variant];
    const int slots_per_cta = V.ow;                                // x lanes
    int grid = e->sm_count;
    if (n < grid * slots_per_cta) grid = (n + slots_per_cta - 1) / slots_per_cta;
    int lanes = (n + grid * slots_per_cta - 1) / (grid * slots_per_cta);
    if (lanes < 1) lanes = 1;
    if (lanes > V.lpw) lanes = V.lpw;
    if (e->force_lanes > 0) lanes = e->force_lanes;
    K.lanes_per_warp = lanes;
    K.start = e->next_start;
    e->next_start = nullptr;
    // more patches than resident slots -> parked variant (virtual slots per CTA, state pools in HBM/L2)
    const bool parked = e->parked_mode == 1 || (e->parked_mode != 0 && (long long)n > (long long)e->sm_count * V.ow * V.lpw * 5 / 4);
    HP_CUDA(cudaEventRecord(e->ev0, s));
    if (!parked) {
        V.fn<<<grid, (V.ow + V.sw) * 32, e->smem_opt_bytes, s>>>(K);
    } else {
        const ParkedVariant& PV = g_pvariants[e->pvariant];
        const int g_parked_ow = PV.ow;
        grid = e->sm_count;
        const int ps = (int)(e->parked_seq++ % 2);
        if (!e->pool_done[ps]) HP_CUDA(cudaEventCreateWithFlags(&e->pool_done[ps], cudaEventDisableTiming));
        HP_CUDA(cudaStreamWaitEvent(s, e->pool_done[ps], 0));
        if (e->pool_ctas2[ps] < grid) {
            cudaFree(e->d_pool_bq2[ps]); cudaFree(e->d_pool_ctx2[ps]);
            e->d_pool_bq2[ps] = nullptr; e->d_pool_ctx2[ps] = nullptr; e->pool_ctas2[ps] = 0;
            HP_CUDA(cudaMalloc(&e->d_pool_bq2[ps], sizeof(hp::BqSlotP) * (size_t)grid * hp::VMAX));
            HP_CUDA(cudaMalloc(&e->d_pool_ctx2[ps], sizeof(hp::LaneCtx) * (size_t)grid * hp::VMAX));
            e->pool_ctas2[ps] = grid;
        }
        int v = (n + grid - 1) / grid;                          // virtual slots per CTA: all patches in flight if they fit
        v = (v + g_parked_ow - 1) / g_parked_ow * g_parked_ow;
        if (v > hp::VMAX) v = hp::VMAX;
        if (v < g_parked_ow) v = g_parked_ow;
        if (const char* vs = getenv("HPMVS_VSLOTS")) { v = atoi(vs) / g_parked_ow * g_parked_ow; if (v < g_parked_ow) v = g_parked_ow; if (v > hp::VMAX) v = hp::VMAX; }
        K.vslots = v;
        K.pool_bq = e->d_pool_bq2[ps];
        K.pool_ctx = e->d_pool_ctx2[ps];
        PV.fn<<<grid, (PV.ow + PV.sw) * 32, PV.smem, s>>>(K);
        HP_CUDA(cudaEventRecord(e->pool_done[ps], s));
    }
    HP_CUDA(cudaEventRecord(e->ev1, s));
    HP_CUDA(cudaEventRecord(e->slot_done[slot], s));
    e->slot_stream[slot] = s;
    e->launches++;
    HP_CUDA(cudaGetLastError());
    return 0;
}

int hpmvs_optimize_batch_device(hpmvs_engine_t* e, int n, const hpmvs_patch_t* d_in, hpmvs_patch_t* d_out, void* stream) {
    if (!e || n < 0 || (n > 0 && (!d_in || !d_out))) return HPMVS_E_ARG;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    return launch_optimize(e, n, d_in, d_out, stream ? (cudaStream_t)stream : e->stream);
}

int hpmvs_start_parameters(hpmvs_engine_t* e, int n, const hpmvs_patch_t* in, double* out) {
    if (!e || n < 0 || (n > 0 && (!in || !out))) return HPMVS_E_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    if (e->ncams <= 0) return HPMVS_E_STATE;
    host_start_parameters(e, n, in, out);
    return 0;
}

int hpmvs_optimize_batch_device_start(hpmvs_engine_t* e, int n, const hpmvs_patch_t* d_in, hpmvs_patch_t* d_out, const double* d_start,
                                      void* stream) {
    if (!e || n < 0 || (n > 0 && (!d_in || !d_out))) return HPMVS_E_ARG;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    e->next_start = d_start;
    return launch_optimize(e, n, d_in, d_out, stream ? (cudaStream_t)stream : e->stream);
}

int hpmvs_engine_set_start_mode(hpmvs_engine_t* e, int mode) {
    if (!e || (mode != 0 && mode != 1)) return HPMVS_E_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    e->start_mode = mode;
    return 0;
}

int hpmvs_optimize_batch(hpmvs_engine_t* e, int n, const hpmvs_patch_t* in, hpmvs_patch_t* out, void* stream) {
    if (!e || n < 0 || (n > 0 && (!in || !out))) return HPMVS_E_ARG;
    if (n == 0) return 0;
    HP_NVTX("hpmvs_optimize_batch (H2D + kernel + D2H)");
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    int rc = check_ready(e);
    if (rc) return rc;
    rc = ensure_patch_capacity(e, (size_t)n);
    if (rc) return rc;
    if (!valid_view_ids(e, n, in)) return HPMVS_E_ARG;
    HP_CUDA(cudaMemcpyAsync(e->d_in, in, sizeof(hpmvs_patch_t) * n, cudaMemcpyHostToDevice, s));
    if (e->start_mode == 1) {
        // the one libm-dependent scalar step of the path, evaluated where the reference evaluates it: on the host
        if ((size_t)n > e->cap_start) {
            if (e->d_start) cudaFree(e->d_start);
            if (e->h_start) cudaFreeHost(e->h_start);
            e->d_start = nullptr; e->h_start = nullptr; e->cap_start = 0;
            const size_t cap = (size_t)n + (size_t)n / 4 + 1024;
            HP_CUDA(cudaMalloc(&e->d_start, cap * 2 * sizeof(double)));
            HP_CUDA(cudaMallocHost(&e->h_start, cap * 2 * sizeof(double)));
            e->cap_start = cap;
        }
        host_start_parameters(e, n, in, e->h_start);
        HP_CUDA(cudaMemcpyAsync(e->d_start, e->h_start, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, s));
        e->next_start = e->d_start;
    }
    rc = launch_optimize(e, n, e->d_in, e->d_out, s, /*async_call=*/false);
    if (rc) return rc;
    HP_CUDA(cudaMemcpyAsync(out, e->d_out, sizeof(hpmvs_patch_t) * n, cudaMemcpyDeviceToHost, s));
    HP_CUDA(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&e->last_kernel_ms, e->ev0, e->ev1);
    return 0;
}

// Asynchronous form of hpmvs_optimize_batch: H2D, kernel and D2H are only ENQUEUED on `stream` (which must not be NULL); the
// call returns at once and the caller synchronises its stream.  Two staging sets alternate, so two batches submitted on two
// streams overlap: the next batch's CTAs start on the SMs the previous batch has already drained.
int hpmvs_optimize_batch_submit(hpmvs_engine_t* e, int n, const hpmvs_patch_t* in, hpmvs_patch_t* out, void* stream) {
    if (!e || !stream || n < 0 || (n > 0 && (!in || !out))) return HPMVS_E_ARG;
    if (n == 0) return 0;
    HP_NVTX("hpmvs_optimize_batch_submit (enqueue H2D + kernel + D2H)");
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check_ready(e);
    if (rc) return rc;
    if (!valid_view_ids(e, n, in)) return HPMVS_E_ARG;
    hpmvs_engine::Stage& st = e->stage2[pick_slot(e->stage2, 16, e->submit_seq)];
    if (!st.done) HP_CUDA(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
    if ((size_t)n > st.cap) {
        HP_CUDA(cudaEventSynchronize(st.done));
        cudaFree(st.d_in); cudaFree(st.d_out); st.d_in = st.d_out = nullptr; st.cap = 0;
        const size_t cap = (size_t)n + (size_t)n / 4 + 1024;
        HP_CUDA(cudaMalloc(&st.d_in, cap * sizeof(hpmvs_patch_t)));
        HP_CUDA(cudaMalloc(&st.d_out, cap * sizeof(hpmvs_patch_t)));
        st.cap = cap;
    }
    if (e->start_mode == 1) {
        HP_CUDA(cudaEventSynchronize(st.done));          // the pinned angle buffer of this set is free again
        if ((size_t)n > st.cap_start) {
            cudaFree(st.d_start); if (st.h_start) cudaFreeHost(st.h_start);
            st.d_start = nullptr; st.h_start = nullptr; st.cap_start = 0;
            const size_t cap = (size_t)n + (size_t)n / 4 + 1024;
            HP_CUDA(cudaMalloc(&st.d_start, cap * 2 * sizeof(double)));
            HP_CUDA(cudaMallocHost(&st.h_start, cap * 2 * sizeof(double)));
            st.cap_start = cap;
        }
        host_start_parameters(e, n, in, st.h_start);
    }
    HP_CUDA(cudaStreamWaitEvent(s, st.done, 0));         // the previous batch that used this staging set has left it
    HP_CUDA(cudaMemcpyAsync(st.d_in, in, sizeof(hpmvs_patch_t) * n, cudaMemcpyHostToDevice, s));
    if (e->start_mode == 1) {
        HP_CUDA(cudaMemcpyAsync(st.d_start, st.h_start, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, s));
        e->next_start = st.d_start;
    }
    rc = launch_optimize(e, n, st.d_in, st.d_out, s);
    if (rc) return rc;
    HP_CUDA(cudaMemcpyAsync(out, st.d_out, sizeof(hpmvs_patch_t) * n, cudaMemcpyDeviceToHost, s));
    HP_CUDA(cudaEventRecord(st.done, s));
    return 0;
}

int hpmvs_ncc_batch(hpmvs_engine_t* e, int n, const hpmvs_patch_t* in, int ref_idx, int robust, float* inccs, void* stream) {
    if (!e || n < 0 || ref_idx < 0 || ref_idx >= HPMVS_MAX_VIEWS || (n > 0 && (!in || !inccs))) return HPMVS_E_ARG;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    int rc = check_ready(e);
    if (rc) return rc;
    if (!valid_view_ids(e, n, in)) return HPMVS_E_ARG;
    rc = sync_cameras(e);
    if (rc) return rc;
    rc = ensure_patch_capacity(e, (size_t)n);
    if (rc) return rc;
    const size_t ni = (size_t)n * HPMVS_MAX_VIEWS;
    if (ni > e->cap_inccs) {
        if (e->d_inccs) cudaFree(e->d_inccs);
        e->d_inccs = nullptr; e->cap_inccs = 0;
        HP_CUDA(cudaMalloc(&e->d_inccs, ni * sizeof(float)));
        e->cap_inccs = ni;
    }
    HP_CUDA(cudaMemcpyAsync(e->d_in, in, sizeof(hpmvs_patch_t) * n, cudaMemcpyHostToDevice, s));
    const hp::KParams K = make_params(e, e->d_in, e->d_out, n);
    int grid = e->sm_count * e->blocks_per_sm;
    const int need = (n + hp::WARPS_PER_BLOCK - 1) / hp::WARPS_PER_BLOCK;
    if (need < grid) grid = need;
    hp::ncc_kernel<<<grid, hp::WARPS_PER_BLOCK * 32, e->smem_bytes, s>>>(K, ref_idx, robust, e->d_inccs);
    e->launches++;
    HP_CUDA(cudaGetLastError());
    HP_CUDA(cudaMemcpyAsync(inccs, e->d_inccs, ni * sizeof(float), cudaMemcpyDeviceToHost, s));
    HP_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int hpmvs_ncc_batch_device(hpmvs_engine_t* e, int n, const hpmvs_patch_t* d_in, int ref_idx, int robust, float* d_inccs, void* stream) {
    if (!e || n < 0 || ref_idx < 0 || ref_idx >= HPMVS_MAX_VIEWS || (n > 0 && (!d_in || !d_inccs))) return HPMVS_E_ARG;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    int rc = check_ready(e);
    if (rc) return rc;
    rc = sync_cameras(e);
    if (rc) return rc;
    const hp::KParams K = make_params(e, d_in, e->d_out, n);
    int grid = e->sm_count * e->blocks_per_sm;
    const int need = (n + hp::WARPS_PER_BLOCK - 1) / hp::WARPS_PER_BLOCK;
    if (need < grid) grid = need;
    HP_CUDA(cudaEventRecord(e->ev0, s));
    if (e->ncc_tma) {
        int g2 = e->sm_count * 3;                              // ~70 KB per CTA: 3 CTAs (12 warps) per SM
        if (need < g2) g2 = need;
        hp::ncc_kernel_tma<<<g2, hp::WARPS_PER_BLOCK * 32, sizeof(hp::NccWarpTma) * hp::WARPS_PER_BLOCK, s>>>(K, ref_idx, robust, d_inccs, e->d_tma_fallbacks);
    } else {
        hp::ncc_kernel<<<grid, hp::WARPS_PER_BLOCK * 32, e->smem_bytes, s>>>(K, ref_idx, robust, d_inccs);
    }
    HP_CUDA(cudaEventRecord(e->ev1, s));
    e->launches++;
    HP_CUDA(cudaGetLastError());
    return 0;
}

// ---- "next" row f-2: depth maps + acceptance tests -----------------------------------------------------------
int hpmvs_engine_depth_reset(hpmvs_engine_t* e) {
    if (!e) return HPMVS_E_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    if (e->ncams <= 0) return HPMVS_E_STATE;
    const int nl = e->opt.maxlevel + 1;
    if ((int)e->depths.size() != e->ncams) {
        for (auto& cam : e->depths) for (auto* d : cam) if (d) cudaFree(d);
        e->depths.assign(e->ncams, std::vector<float*>(HPMVS_LEVELS, nullptr));
        e->depth_cells.assign(e->ncams, std::vector<size_t>(HPMVS_LEVELS, 0));
        e->cams_dirty = true;
    }
    if ((int)e->depth_cells.size() != e->ncams) e->depth_cells.assign(e->ncams, std::vector<size_t>(HPMVS_LEVELS, 0));
    for (int c = 0; c < e->ncams; c++)
        for (int l = 0; l < nl; l++) {
            const size_t rows = (size_t)(int)(e->h_cams[c].h[l] / 2.0), cols = (size_t)(int)(e->h_cams[c].w[l] / 2.0);
            const size_t cnt = rows * cols > 0 ? rows * cols : 1;
            if (e->depths[c][l] && e->depth_cells[c][l] != cnt) { cudaFree(e->depths[c][l]); e->depths[c][l] = nullptr; }
            if (!e->depths[c][l]) { HP_CUDA(cudaMalloc(&e->depths[c][l], cnt * sizeof(float))); e->depth_cells[c][l] = cnt; e->cams_dirty = true; }
            hp::depth_fill_kernel<<<(unsigned)((cnt + 255) / 256 > 1024 ? 1024 : (cnt + 255) / 256), 256, 0, e->stream>>>(e->depths[c][l], cnt);
            e->launches++;
        }
    HP_CUDA(cudaGetLastError());
    HP_CUDA(cudaStreamSynchronize(e->stream));
    return sync_cameras(e);
}

static int ensure_depths(hpmvs_engine* e) {
    if ((int)e->depths.size() == e->ncams && e->ncams > 0 && e->depths[0][0]) return 0;
    return HPMVS_E_STATE;
}

static int depth_set_impl(hpmvs_engine_t* e, int n, const hpmvs_patch_t* patches, void* stream, int subtract);
int hpmvs_depth_set_batch(hpmvs_engine_t* e, int n, const hpmvs_patch_t* patches, void* stream) { return depth_set_impl(e, n, patches, stream, 0); }
int hpmvs_depth_unset_batch(hpmvs_engine_t* e, int n, const hpmvs_patch_t* patches, void* stream) { return depth_set_impl(e, n, patches, stream, 1); }
static int depth_set_impl(hpmvs_engine_t* e, int n, const hpmvs_patch_t* patches, void* stream, int subtract) {
    if (!e || n < 0 || (n > 0 && !patches)) return HPMVS_E_ARG;
    HP_NVTX("hpmvs_depth_set/unset_batch");
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    int rc = ensure_depths(e); if (rc) return rc;
    if (!valid_view_ids(e, n, patches)) return HPMVS_E_ARG;
    rc = sync_cameras(e); if (rc) return rc;
    rc = ensure_patch_capacity(e, (size_t)n); if (rc) return rc;
    HP_CUDA(cudaMemcpyAsync(e->d_in, patches, sizeof(hpmvs_patch_t) * n, cudaMemcpyHostToDevice, s));
    const hp::KParams K = make_params(e, e->d_in, e->d_out, n);
    const int total = n * HPMVS_MAX_VIEWS;
    hp::depth_set_kernel<<<(total + 255) / 256, 256, 0, s>>>(K, e->d_in, n, subtract);
    e->launches++;
    HP_CUDA(cudaGetLastError());
    HP_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int hpmvs_accept_batch(hpmvs_engine_t* e, int n, const hpmvs_patch_t* patches, float margin, int32_t* out, void* stream) {
    if (!e || n < 0 || (n > 0 && (!patches || !out))) return HPMVS_E_ARG;
    HP_NVTX("hpmvs_accept_batch");
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    int rc = ensure_depths(e); if (rc) return rc;
    if (!valid_view_ids(e, n, patches)) return HPMVS_E_ARG;
    rc = sync_cameras(e); if (rc) return rc;
    rc = ensure_patch_capacity(e, (size_t)n); if (rc) return rc;
    if ((size_t)n * 3 > e->cap_accept) {
        if (e->d_accept) cudaFree(e->d_accept);
        e->d_accept = nullptr; e->cap_accept = 0;
        HP_CUDA(cudaMalloc(&e->d_accept, sizeof(int) * 3 * ((size_t)n + 1024)));
        e->cap_accept = 3 * ((size_t)n + 1024);
    }
    HP_CUDA(cudaMemcpyAsync(e->d_in, patches, sizeof(hpmvs_patch_t) * n, cudaMemcpyHostToDevice, s));
    const hp::KParams K = make_params(e, e->d_in, e->d_out, n);
    int grid = (n + 7) / 8;
    if (grid > e->sm_count * 8) grid = e->sm_count * 8;
    hp::accept_kernel<<<grid, 256, 0, s>>>(K, e->d_in, n, margin, e->d_accept);
    e->launches++;
    HP_CUDA(cudaGetLastError());
    HP_CUDA(cudaMemcpyAsync(out, e->d_accept, sizeof(int) * 3 * (size_t)n, cudaMemcpyDeviceToHost, s));
    HP_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// Border de-duplication on device-resident records (the gathered patch set on the root rank): keep[i] = 1 for the survivors, *nkeep
// their number (device int, may be NULL).  Asynchronous on `stream`.
int hpmvs_dedup_border_device(hpmvs_engine_t* e, int n, const hpmvs_patch_t* d_records, const int32_t* d_owner, const double origin[3],
                              double cell, uint8_t* d_keep, int32_t* d_nkeep, void* stream) {
    if (!e || n < 0 || (n > 0 && (!d_records || !d_owner || !d_keep)) || !(cell > 0.0)) return HPMVS_E_ARG;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    size_t cap = 1024;
    while (cap < (size_t)n * 2) cap <<= 1;
    const size_t need = cap * (8 + 4 + 8 + 8) + (size_t)n * 4 + 64;
    if (need > e->cap_dedup) {
        if (e->d_dedup) cudaFree(e->d_dedup);
        e->d_dedup = nullptr; e->cap_dedup = 0;
        HP_CUDA(cudaMalloc(&e->d_dedup, need));
        e->cap_dedup = need;
    }
    hp::DedupTable T;
    unsigned char* p = e->d_dedup;
    T.keys = (unsigned long long*)p; p += cap * 8;
    T.best_score = (unsigned long long*)p; p += cap * 8;
    T.best_who = (unsigned long long*)p; p += cap * 8;
    T.best_nimg = (int*)p; p += cap * 4;
    T.slot_of = (int*)p;
    T.cap_mask = (unsigned)(cap - 1);
    HP_CUDA(cudaMemsetAsync(T.keys, 0, cap * 8, s));
    HP_CUDA(cudaMemsetAsync(T.best_score, 0xff, cap * 16, s));           // best_score and best_who: all ones = "no candidate yet"
    HP_CUDA(cudaMemsetAsync(T.best_nimg, 0, cap * 4, s));
    if (d_nkeep) HP_CUDA(cudaMemsetAsync(d_nkeep, 0, sizeof(int), s));
    const double zero[3] = {0.0, 0.0, 0.0};
    if (!origin) origin = zero;
    int grid = (n + 255) / 256;
    if (grid > e->sm_count * 8) grid = e->sm_count * 8;
    hp::dedup_insert_kernel<<<grid, 256, 0, s>>>(d_records, n, origin[0], origin[1], origin[2], cell, T);
    hp::dedup_score_kernel<<<grid, 256, 0, s>>>(d_records, n, T);
    hp::dedup_who_kernel<<<grid, 256, 0, s>>>(d_records, d_owner, n, T);
    int* nk = d_nkeep;
    if (!nk) { nk = (int*)(e->d_dedup + need - 16); HP_CUDA(cudaMemsetAsync(nk, 0, sizeof(int), s)); }   // no counter wanted: count into a scratch word
    hp::dedup_verdict_kernel<<<grid, 256, 0, s>>>(d_owner, n, T, d_keep, nk);
    e->launches += 4;
    HP_CUDA(cudaGetLastError());
    return 0;
}

// ---- device-resident forms of the steps around optimize() in a level of the loop: no host staging, asynchronous on `stream` -------
int hpmvs_expand_candidates_device(hpmvs_engine_t* e, int n, const hpmvs_patch_t* d_parents, const float* d_widths, int mode,
                                   hpmvs_patch_t* d_out, void* stream) {
    if (!e || n < 0 || (n > 0 && (!d_parents || !d_widths || !d_out)) || (mode != 4 && mode != 6)) return HPMVS_E_ARG;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    if (e->ncams <= 0) return HPMVS_E_STATE;
    int rc = sync_cameras(e); if (rc) return rc;
    hp::ExpandDirs D{};
    for (int ii = 0; ii < mode; ii++) {                       // the host function's own angles and libm calls (host_scene.cpp)
        const float angle = (mode == 6) ? (float)(2.0 * M_PI / mode * ii) : (float)(2.0 * M_PI / mode * ii + M_PI / 4);
        D.dx[ii] = (float)cos((double)angle); D.dy[ii] = (float)sin((double)angle);
    }
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    const int total = n * mode;
    hp::expand_candidates_kernel<<<(total + 127) / 128, 128, 0, s>>>(e->d_cams, e->ncams, d_parents, d_widths, n, mode, D, d_out);
    e->launches++;
    HP_CUDA(cudaGetLastError());
    return 0;
}

int hpmvs_depth_set_batch_device(hpmvs_engine_t* e, int n, const hpmvs_patch_t* d_patches, int subtract, void* stream) {
    if (!e || n < 0 || (n > 0 && !d_patches)) return HPMVS_E_ARG;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    int rc = ensure_depths(e); if (rc) return rc;
    rc = sync_cameras(e); if (rc) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    const hp::KParams K = make_params(e, d_patches, e->d_out, n);
    const int total = n * HPMVS_MAX_VIEWS;
    hp::depth_set_kernel<<<(total + 255) / 256, 256, 0, s>>>(K, d_patches, n, subtract ? 1 : 0);
    e->launches++;
    HP_CUDA(cudaGetLastError());
    return 0;
}

int hpmvs_accept_batch_device(hpmvs_engine_t* e, int n, const hpmvs_patch_t* d_patches, float margin, int32_t* d_out, void* stream) {
    if (!e || n < 0 || (n > 0 && (!d_patches || !d_out))) return HPMVS_E_ARG;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    int rc = ensure_depths(e); if (rc) return rc;
    rc = sync_cameras(e); if (rc) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    const hp::KParams K = make_params(e, d_patches, e->d_out, n);
    int grid = (n + 7) / 8;
    if (grid > e->sm_count * 8) grid = e->sm_count * 8;
    hp::accept_kernel<<<grid, 256, 0, s>>>(K, d_patches, n, margin, d_out);
    e->launches++;
    HP_CUDA(cudaGetLastError());
    return 0;
}

int hpmvs_engine_download_depth(hpmvs_engine_t* e, int cam, int level, float* out, int* rows, int* cols) {
    if (!e || cam < 0 || cam >= e->ncams || level < 0 || level > e->opt.maxlevel || !rows || !cols) return HPMVS_E_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    int rc = ensure_depths(e); if (rc) return rc;
    *rows = (int)(e->h_cams[cam].h[level] / 2.0); *cols = (int)(e->h_cams[cam].w[level] / 2.0);
    if (out) HP_CUDA(cudaMemcpy(out, e->depths[cam][level], sizeof(float) * (size_t)(*rows) * (*cols), cudaMemcpyDeviceToHost));
    return 0;
}

// debugging aid (not in the header): textures of the TMA-staged scoring kernel whose footprint did not fit the staged window
long long hpmvs_engine_tma_fallbacks(hpmvs_engine_t* e) {
    if (!e || !e->d_tma_fallbacks) return -1;
    unsigned long long v = 0;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    cudaMemcpy(&v, e->d_tma_fallbacks, sizeof(v), cudaMemcpyDeviceToHost);
    return (long long)v;
}

// debugging aid (not in the header): writes the per-round log of the most recent wavefront launch as CSV (needs HPMVS_WF_LOG at create)
int hpmvs_engine_dump_round_log(hpmvs_engine_t* e, const char* path) {
    if (!e || !path) return HPMVS_E_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    HP_CUDA(cudaDeviceSynchronize());
    hpmvs_engine::WfContext& w = e->wfb[e->wf_last].part[0];
    if (!w.round_log) return HPMVS_E_STATE;
    hp::WfCtl c;
    HP_CUDA(cudaMemcpy(&c, w.ctl, sizeof(c), cudaMemcpyDeviceToHost));
    const int n = c.round < 65536 ? c.round : 65536;
    std::vector<unsigned long long> h((size_t)4 * n);
    HP_CUDA(cudaMemcpy(h.data(), w.round_log, sizeof(unsigned long long) * 4 * n, cudaMemcpyDeviceToHost));
    FILE* fh = fopen(path, "w");
    if (!fh) return HPMVS_E_ARG;
    fprintf(fh, "round,us,evals,posts,dead\n");
    for (int i = 0; i < n; i++)
        fprintf(fh, "%d,%.3f,%llu,%llu,%llu\n", i, i ? (h[4 * i] - h[0]) / 1e3 : 0.0, h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
    fclose(fh);
    return n;
}

int hpmvs_engine_counters(hpmvs_engine_t* e, hpmvs_counters_t* out, int reset) {
    if (!e || !out) return HPMVS_E_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    HP_CUDA(cudaSetDevice(e->device));
    unsigned long long c[16];
    HP_CUDA(cudaStreamSynchronize(e->stream));
    HP_CUDA(cudaMemcpy(c, e->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
    if (getenv("HPMVS_PROFILE_PRINT"))
    {
#ifdef HP_PROFILE
        unsigned long long pc[16], pt[16], pl[16];
        cudaMemcpyFromSymbol(pc, bq3::g_prof_cycles, sizeof(pc)); cudaMemcpyFromSymbol(pt, bq3::g_prof_trips, sizeof(pt));
        cudaMemcpyFromSymbol(pl, bq3::g_prof_lanes, sizeof(pl));
        unsigned long long ev[8];
        cudaMemcpyFromSymbol(ev, hp::g_eval_prof, sizeof(ev));
        fprintf(stderr, "[hpmvs profile] eval Gcyc: setup %.2f sample %.2f stats %.2f normref %.2f dots %.2f\n", ev[0] / 1e9, ev[1] / 1e9, ev[2] / 1e9, ev[3] / 1e9, ev[4] / 1e9);
        {
            unsigned long long tc[8], tt[8], tl[8];
            cudaMemcpyFromSymbol(tc, bq3::g_tprof_cycles, sizeof(tc)); cudaMemcpyFromSymbol(tt, bq3::g_tprof_trips, sizeof(tt)); cudaMemcpyFromSymbol(tl, bq3::g_tprof_lanes, sizeof(tl));
            const char* tn[8] = {"T_RESTART", "T_DIRECTION", "T_CGSTEP", "T_BOUNDARY", "T_ALT_PREP", "T_ALT_DIR", "T_ALT_SEARCH", "T_FINISH"};
            for (int i = 0; i < 8; i++) if (tt[i]) fprintf(stderr, "[hpmvs profile]     trsbox %-12s trips %10llu cycles/trip %7.0f lanes/trip %5.1f total Gcyc %7.2f\n", tn[i], tt[i], (double)tc[i] / tt[i], (double)tl[i] / tt[i], tc[i] / 1e9);
        }
        const char* names[13] = {"AFTER_EVAL", "RESCUE_LOOP", "RESCUE_DONE", "GOPT_FIX", "FARPOINT", "REDUCE_RHO", "TRUST", "SHIFT", "RESCUE", "ALTMOV", "VLAG", "EVAL", "EXIT"};
        for (int i = 0; i < 13; i++)
            if (pt[i]) fprintf(stderr, "[hpmvs profile]   %-12s trips %10llu  cycles/trip %8.0f  lanes/trip %5.1f  total Gcyc %7.2f\n", names[i], pt[i], (double)pc[i] / pt[i], (double)pl[i] / pt[i], pc[i] / 1e9);
        fprintf(stderr, "[hpmvs profile] opt: wait %llu adv %llu rounds %llu lanes %llu | sampler: idle %llu eval %llu n_eval %llu | parked: copy %llu sampler_other %llu\n", c[4], c[5], c[6], c[7], c[8], c[9], c[10], c[11], c[12]);
#endif
    }
    out->patches = c[0]; out->patches_ok = c[1]; out->evals = c[2]; out->textures = c[3];
    out->kernel_launches = e->launches + c[13];                  // host-side launches + the round kernels of the wavefront graphs
    if (reset) {
        HP_CUDA(cudaMemset(e->d_counters, 0, sizeof(c)));
        e->launches = 0;
    }
    return 0;
}

void* hpmvs_engine_stream(hpmvs_engine_t* e) { return e ? (void*)e->stream : nullptr; }

float hpmvs_engine_last_kernel_ms(hpmvs_engine_t* e) {
    if (!e) return 0.0f;
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) e->last_kernel_ms = ms;
    return e->last_kernel_ms;
}

}  // extern "C"
