// Device code of the hpmvs_b200 engine (sm_100a): everything below PatchOptimizer::optimize()
// (/root/reference/src/hpmvs/PatchOptimizer.cpp:48-103) for a batch of patches, in one persistent kernel.
//
// Work decomposition ("lane = patch for control, warp = patch for pixels"):
//   * every warp keeps up to 32 patches in flight, one per lane slot.  The FP64 BOBYQA state machine
//     (bobyqa3.h, ~1.6 KB of state per patch, thread-private => local memory) advances for all of a warp's
//     patches at once in SIMT fashion, so its long scalar instruction stream is amortised over the lanes;
//   * whenever patches need their photometric objective, the warp serves them one after the other with all
//     32 lanes co-operating: per-view projection set-up (lane = view), 7x7 bilinear RGB sampling
//     (lane = sample) and the mean / variance / correlation reductions, which are evaluated as SEQUENTIAL f32
//     chains (lane = texture) so that every rounding matches the reference's scalar loops
//     (Patch2d.hpp:37-84) - that is what makes the BOBYQA trajectory reproducible bit for bit;
//   * the short view-list stages before and after the refinement (addImages, filterImagesNCC, sortImages,
//     setRefImage, ...) also run warp-co-operatively on the patch's context in shared memory.
// Warps pull patches from a global atomic counter (the work per patch varies by 10x) and refill a lane slot
// as soon as its patch retires.
//
// Compile with -fmad=false: every f32/f64 expression below is written in the reference's evaluation order and
// must not be contracted into FMAs.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hpmvs_b200.h"
#define BQ_STATE_IN_SHARED 1
#include "bobyqa3.h"
#include "undistort_math.h"

namespace hp {

constexpr int MAXV = HPMVS_MAX_VIEWS;
constexpr int VC = 8;            // texture slots resident per warp (slot 0 = reference view)
constexpr int TEXN = 147;        // 7*7*3 floats per texture (Patch2d.hpp:31,88)
constexpr int TEXS = 148;        // padded stride: (148*slot) mod 32 distinct for 8 slots -> conflict-free chains
constexpr int QS = 52;
// The variance chain of PatchTex::normalize adds one term (f0^2 + f1^2) + f2^2 per sample; forming the terms inside the serial chain
// costs 12 instructions per sample with one lane per texture.  HP_VAR_PREPASS forms them in a flattened pass (lane = sample) and leaves
// a 2-instruction chain (load + add) - same terms, same order of additions, so the same bits.
#ifndef HP_VAR_PREPASS
#define HP_VAR_PREPASS 1
#endif
constexpr int WARPS_PER_BLOCK = 4;
constexpr unsigned FULL = 0xffffffffu;

#ifdef HP_PROFILE
__device__ unsigned long long g_eval_prof[8];
#define HP_EVT(i, t0) do { if (lane == 0) atomicAdd(&g_eval_prof[i], (unsigned long long)(clock64() - (t0))); } while (0)
#else
#define HP_EVT(i, t0) do { } while (0)
#endif

// -DHP_PROFILE adds HP_CLOCK() accounting of optimizer / sampler time to counters[4..10] (see engine.cu)
#ifdef HP_PROFILE
#define HP_CLOCK() clock64()
#else
#define HP_CLOCK() 0ll
#endif

struct DevCamera {
    float P[HPMVS_LEVELS][12];
    float center[4];
    float xaxis[4], yaxis[4], zaxis[4];   // Camera::xAxis_/yAxis_/zAxis_
    float nx[4], ny[4], nz[4];            // .normalized() of the above (setOptimizationFields, :388-390)
    float ksum;                           // kMat_[0](0,0) + kMat_[0](1,1) in f32
    int nlevels;
    int w[HPMVS_LEVELS], h[HPMVS_LEVELS];
    int pitch[HPMVS_LEVELS];              // row pitch in pixels (uchar4)
    const uchar4* img[HPMVS_LEVELS];
    float* depth[HPMVS_LEVELS];           // Scene::m_depths[cam][level] (Scene.h:75-76), rows x cols, row-major
    int drows[HPMVS_LEVELS], dcols[HPMVS_LEVELS];
};

struct KParams {
    const DevCamera* cams;
    int ncams;
    const int* covis_off;
    const int* covis_ids;
    hpmvs_options_t opt;
    double cos_max_d;     // cos((double)MAX_ANGLE): sampleTexture's gate compares float < double (:485)
    float cos_max_f;      // std::cos(MAX_ANGLE) in f32: addImages / filterImagesByAngle (:129,:239)
    float sort_thr;       // 1.0f - cos(10 deg) narrowed to f32 (:184)
    float angle_scale;    // (float)(M_PI / 48.0f) (:398)
    const hpmvs_patch_t* in;
    hpmvs_patch_t* out;
    int n;
    int* work_counter;
    unsigned long long* counters;   // patches, ok, evals, textures
    int lanes_per_warp;             // patches kept in flight per warp (1..32)
    // optional: parametersFromCenterNorm's angles (x[1], x[2], scaled and clamped) per input patch, evaluated by the caller with
    // the HOST's libm exactly where the reference evaluates them (PatchOptimizer.cpp:427-437); null = evaluate here
    // parked variant: per-CTA pools of virtual patch slots in HBM/L2 (state + context), `vslots` per CTA
    struct BqSlot* pool_bq;
    struct LaneCtx* pool_ctx;
    int vslots;
    const double* start;            // see above (2 doubles per input patch) or null
};

struct ViewSetup {
    float tlx, tly, dxx, dxy, dyx, dyy;
    int pitch;
    int pad;
    const uchar4* img;
};

// per-warp scratch for one photometric evaluation
struct __align__(16) Scratch {
    union {
        float tex[VC][TEXS];            // raw -> normalised -> product textures of the views being evaluated
        float rays[MAXV][4];            // view rays of the view-list stages (never live at the same time as tex)
    };
#if HP_VAR_PREPASS
    float sq[VC][QS];               // per-sample squared deviations (f0^2 + f1^2) + f2^2 of the textures being evaluated
#endif
    float mean[VC][4];
    float sigma[VC];
    int slot_view[VC];
    ViewSetup vs[MAXV];
    float dots[MAXV];
    float incc[MAXV];
    float tmp_f[MAXV];
    int tmp_i[MAXV];
    int vlist[MAXV];
    unsigned char vvalid[MAXV];
};

// per-lane-slot patch context (shared memory): what the co-operative stages read and write for one patch
struct __align__(16) LaneCtx {
    float center[4], normal[4];          // pCenter_, pNormal_
    float refCenter[4], refRay[4];       // refCenter_, refRay_
    float X0[4], Y0[4], Z0[4];           // imgX_[0], imgY_[0], imgZ_[0]
    float xa[4], ya[4], za[4];           // pXaxis_, pYaxis_, pZaxis_ for the current centre / normal (objective only)
    double score;
    float scale;                         // pScale_
    int nimg;                            // pImages_.size()
    int textures;
    int patch_index;
    int status, nlopt_rc, evals, pad;
    unsigned short images[MAXV];         // pImages_
};

// one warp of the stand-alone scoring kernel (K1): 7 KB, so that 28 warps are resident per SM
struct __align__(16) NccWarp {
    Scratch S;
    LaneCtx P;
};

// ----------------------------------------------------------------------------------------------------------
// f32 helpers in Eigen's evaluation order (see oracle/hpmvs_oracle.cpp header for the conventions)
// ----------------------------------------------------------------------------------------------------------
struct f4 { float x, y, z, w; };
struct f3 { float x, y, z; };

__device__ __forceinline__ f4 ld4(const float* p) { return f4{p[0], p[1], p[2], p[3]}; }
__device__ __forceinline__ f4 sub4(f4 a, f4 b) { return f4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
__device__ __forceinline__ f4 add4(f4 a, f4 b) { return f4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
__device__ __forceinline__ float dot4(f4 a, f4 b) {
    const float p0 = a.x * b.x, p1 = a.y * b.y, p2 = a.z * b.z, p3 = a.w * b.w;
    return (p0 + p2) + (p1 + p3);
}
__device__ __forceinline__ f4 normalized4(f4 a) {
    const float z = dot4(a, a);
    if (z > 0.0f) { const float s = sqrtf(z); return f4{a.x / s, a.y / s, a.z / s, a.w / s}; }
    return a;
}
__device__ __forceinline__ float dot3(f3 a, f3 b) {
    const float p0 = a.x * b.x, p1 = a.y * b.y, p2 = a.z * b.z;
    return p0 + (p1 + p2);
}
__device__ __forceinline__ f3 normalized3(f3 a) {
    const float z = dot3(a, a);
    if (z > 0.0f) { const float s = sqrtf(z); return f3{a.x / s, a.y / s, a.z / s}; }
    return a;
}
__device__ __forceinline__ f3 cross3(f3 a, f3 b) {
    return f3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Camera::project (Camera.h:45-62); returns x,y (z is not needed by the path)
__device__ __forceinline__ void project(const DevCamera& cam, f4 X, int level, float& u, float& v) {
    const float* p = cam.P[level];
    float r0 = (p[0] * X.x + p[1] * X.y) + (p[2] * X.z + p[3] * X.w);
    float r1 = (p[4] * X.x + p[5] * X.y) + (p[6] * X.z + p[7] * X.w);
    const float r2 = (p[8] * X.x + p[9] * X.y) + (p[10] * X.z + p[11] * X.w);
    if (r2 <= 0.0f) { u = -65535.0f; v = -65535.0f; return; }
    r0 = r0 / r2; r1 = r1 / r2;
    const float lo = -2147483648.0f, hi = 2147483648.0f;   // (float)(INT_MIN + 3.0f), (float)(INT_MAX - 3.0f)
    u = fmaxf(lo, fminf(hi, r0));
    v = fmaxf(lo, fminf(hi, r1));
}

// Camera::getLevel (Camera.cpp:92-95) given fz = |coord - center|
__device__ __forceinline__ float level_from(const DevCamera& cam, float fz, float scale) {
    return (float)log2((double)(scale * cam.ksum) / (2.0 * (double)fz));
}
__device__ __forceinline__ int leveli_from(const DevCamera& cam, float fz, float scale, int maxLevel) {
    const int l = (int)roundf(level_from(cam, fz, scale));
    return max(0, min(maxLevel, l));
}

// Image::getColor (Image.h:89-115) on the RGBX layout
__device__ __forceinline__ f3 get_color(const uchar4* img, int pitch, float x, float y) {
    const int lx = (int)x, ly = (int)y;
    const float dx1 = x - (float)lx, dx0 = 1.0f - dx1;
    const float dy1 = y - (float)ly, dy0 = 1.0f - dy1;
    const float f00 = dx0 * dy0, f01 = dx0 * dy1, f10 = dx1 * dy0, f11 = dx1 * dy1;
    const uchar4* p = img + (size_t)ly * pitch + lx;
    const uchar4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + pitch), d = __ldg(p + pitch + 1);
    f3 o;
    o.x = ((float)a.x * f00 + (float)c.x * f01) + ((float)b.x * f10 + (float)d.x * f11);
    o.y = ((float)a.y * f00 + (float)c.y * f01) + ((float)b.y * f10 + (float)d.y * f11);
    o.z = ((float)a.z * f00 + (float)c.z * f01) + ((float)b.z * f10 + (float)d.z * f11);
    return o;
}

// ----------------------------------------------------------------------------------------------------------
// calculatePatchAxis (:532-548): all lanes compute the same registers
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void patch_axes(const DevCamera& rc, f4 n, float scale, f4& xa, f4& ya, f4& za) {
    // n.head(3).normalized() (PatchOptimizer.cpp:536): dynamic-size block -> Eigen's scalar reduction loop (a0+a1)+a2
    f3 z = f3{n.x, n.y, n.z};
    {
        const float zz = (n.x * n.x + n.y * n.y) + n.z * n.z;
        if (zz > 0.0f) { const float s = sqrtf(zz); z = f3{n.x / s, n.y / s, n.z / s}; }
    }
    f3 y = normalized3(cross3(z, f3{rc.xaxis[0], rc.xaxis[1], rc.xaxis[2]}));
    f3 x = normalized3(cross3(y, z));
    x.x *= scale; x.y *= scale; x.z *= scale;
    y.x *= scale; y.y *= scale; y.z *= scale;
    const float s = dot3(normalized3(y), normalized3(f3{rc.yaxis[0], rc.yaxis[1], rc.yaxis[2]}));
    y.x = y.x * s; y.y = y.y * s; y.z = y.z * s;
    xa = f4{x.x, x.y, x.z, 0.0f};
    ya = f4{y.x, y.y, y.z, 0.0f};
    za = f4{z.x, z.y, z.z, 0.0f};
}

// ----------------------------------------------------------------------------------------------------------
// sampleTexture part 1 (:484-507): gate, level, three projections, border test.  lane = view.
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool view_setup(const KParams& K, const DevCamera& cam, f4 c, float scale, f4 xa, f4 ya,
                                           f4 zgate, ViewSetup& vs) {
    const f4 d = sub4(ld4(cam.center), c);
    const float z = dot4(d, d);
    const float fz = sqrtf(z);                      // == |c - center| as well (same squares)
    f4 nrm = d;
    if (z > 0.0f) nrm = f4{d.x / fz, d.y / fz, d.z / fz, d.w / fz};
    if ((double)dot4(nrm, zgate) < K.cos_max_d) return false;
    const int lvl = leveli_from(cam, fz, scale, K.opt.maxlevel - 1);
    float cu, cv, xu, xv, yu, yv;
    project(cam, c, lvl, cu, cv);
    project(cam, add4(c, xa), lvl, xu, xv);
    project(cam, add4(c, ya), lvl, yu, yv);
    const float dxx = xu - cu, dxy = xv - cv, dyx = yu - cu, dyy = yv - cv;
    const float hs = 3.5f;
    const float tlx = cu - hs * dxx - hs * dyx, tly = cv - hs * dxy - hs * dyy;
    const float trx = cu + hs * dxx - hs * dyx, try_ = cv + hs * dxy - hs * dyy;
    const float blx = cu - hs * dxx + hs * dyx, bly = cv - hs * dxy + hs * dyy;
    const float brx = cu + hs * dxx + hs * dyx, bry = cv + hs * dxy + hs * dyy;
    const float mnx = fminf(fminf(fminf(tlx, trx), blx), brx), mny = fminf(fminf(fminf(tly, try_), bly), bry);
    const float mxx = fmaxf(fmaxf(fmaxf(tlx, trx), blx), brx), mxy = fmaxf(fmaxf(fmaxf(tly, try_), bly), bry);
    const int m = 3;
    if (mnx < (float)m || mny < (float)m || mxx >= (float)(cam.w[lvl] - m) || mxy >= (float)(cam.h[lvl] - m)) return false;
    vs.tlx = tlx; vs.tly = tly; vs.dxx = dxx; vs.dxy = dxy; vs.dyx = dyx; vs.dyy = dyy;
    vs.pitch = cam.pitch[lvl];
    vs.pad = cam.h[lvl];            // rows of the level (the staged-window variant of the scoring kernel checks its window against them)
    vs.img = cam.img[lvl];
    return true;
}

// sampleTexture part 2 (:509-525) for `ns` textures at once: the (slot, sample) pairs are flattened over the lanes,
// two independent samples per lane and trip so that their loads and conversions overlap.  Raw RGB -> tex[slot].
__device__ __forceinline__ void sample_one(Scratch& W, int idx) {
    const int slot = idx / 49, s = idx - 49 * slot;
    const ViewSetup& v = W.vs[W.slot_view[slot]];
    const int yy = s / 7, xx = s - 7 * yy;
    const float dyx = v.dyx, dyy = v.dyy, dxx = v.dxx, dxy = v.dxy;
    float px = v.tlx, py = v.tly;
    // the reference walks the grid with repeated f32 additions (l += dy; c += dx)
#pragma unroll
    for (int i = 0; i < 6; i++) if (i < yy) { px += dyx; py += dyy; }
#pragma unroll
    for (int i = 0; i < 6; i++) if (i < xx) { px += dxx; py += dxy; }
    const f3 col = get_color(v.img, v.pitch, px, py);
    float* t = W.tex[slot] + 3 * s;
    t[0] = col.x; t[1] = col.y; t[2] = col.z;
}
__device__ __forceinline__ void sample_slots(Scratch& W, int first, int ns, int lane) {
    const int beg = first * 49, end = (first + ns) * 49;
    for (int base = beg + lane; base < end; base += 64) {
        sample_one(W, base);
        if (base + 32 < end) sample_one(W, base + 32);
    }
}

// PatchTex::normalize statistics (Patch2d.hpp:46-71) for slots [first, first+ns): per-channel means and the joint
// standard deviation, each as the reference's SEQUENTIAL f32 chain (lane = one chain).
__device__ __forceinline__ void stats_slots(Scratch& W, int first, int ns, int lane) {
    __syncwarp();
    if (lane < 3 * ns) {
        const int slot = first + lane / 3, ch = lane % 3;
        const float* t = W.tex[slot] + ch;
        float a = 0.0f;
#pragma unroll 7
        for (int i = 0; i < 49; i++) a += t[3 * i];
        W.mean[slot][ch] = a / 49.0f;
    }
    __syncwarp();
#if HP_VAR_PREPASS
    for (int idx = lane; idx < 49 * ns; idx += 32) {
        const int so = idx / 49, i = idx - 49 * so, slot = first + so;
        const float* t = W.tex[slot] + 3 * i;
        const float f0 = W.mean[slot][0] - t[0], f1 = W.mean[slot][1] - t[1], f2 = W.mean[slot][2] - t[2];
        W.sq[slot][i] = f0 * f0 + f1 * f1 + f2 * f2;
    }
    __syncwarp();
    if (lane < ns) {
        const int slot = first + lane;
        const float* q = W.sq[slot];
        float a = 0.0f;
#pragma unroll 7
        for (int i = 0; i < 49; i++) a += q[i];
        float sg = sqrtf(a / 147.0f);
        if (sg == 0.0f) sg = 1.0f;
        W.sigma[slot] = sg;
    }
#else
    if (lane < ns) {
        const int slot = first + lane;
        const float* t = W.tex[slot];
        const float m0 = W.mean[slot][0], m1 = W.mean[slot][1], m2 = W.mean[slot][2];
        float a = 0.0f;
#pragma unroll 7
        for (int i = 0; i < 49; i++) {
            const float f0 = m0 - t[3 * i], f1 = m1 - t[3 * i + 1], f2 = m2 - t[3 * i + 2];
            a += f0 * f0 + f1 * f1 + f2 * f2;
        }
        float sg = sqrtf(a / 147.0f);
        if (sg == 0.0f) sg = 1.0f;
        W.sigma[slot] = sg;
    }
#endif
    __syncwarp();
}

// normalise the reference texture in place (Patch2d.hpp:73-83)
// Element i = lane + 32k of a 147-float texture belongs to channel (lane + 32k) % 3 = (lane % 3 + 2k) % 3: the channel walks a fixed
// rotation as k advances, so the per-element "% 3" (and, in dot_slots, the "/ 147" of a flattened index) is not needed.
__device__ __forceinline__ float chan_mean(int c0, int k, float m0, float m1, float m2) {
    const int r = (2 * k) % 3;                                   // compile-time after unrolling
    // channel = (c0 + r) % 3
    if (r == 0) return c0 == 0 ? m0 : (c0 == 1 ? m1 : m2);
    if (r == 1) return c0 == 0 ? m1 : (c0 == 1 ? m2 : m0);
    return c0 == 0 ? m2 : (c0 == 1 ? m0 : m1);
}
__device__ __forceinline__ void normalize_ref(Scratch& W, int lane) {
    const float sg = W.sigma[0];
    const float m0 = W.mean[0][0], m1 = W.mean[0][1], m2 = W.mean[0][2];
    const int c0 = lane % 3;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int i = lane + 32 * k;
        if (i < TEXN) {
            float v = W.tex[0][i];
            v -= chan_mean(c0, k, m0, m1, m2);
            v /= sg;
            W.tex[0][i] = v;
        }
    }
    __syncwarp();
}

// normalise slots [1, 1+no) and multiply with the normalised reference in one pass, then PatchTex::dot's
// 147-term sequential chain (Patch2d.hpp:37-44), lane = slot
__device__ __forceinline__ void dot_slots(Scratch& W, int no, int lane) {
    const int c0 = lane % 3;
    float r[5];                                                   // this lane's elements of the normalised reference
#pragma unroll
    for (int k = 0; k < 5; k++) r[k] = (lane + 32 * k < TEXN) ? W.tex[0][lane + 32 * k] : 0.0f;
    for (int so = 0; so < no; so++) {
        const int slot = 1 + so;
        const float sg = W.sigma[slot];
        const float m0 = W.mean[slot][0], m1 = W.mean[slot][1], m2 = W.mean[slot][2];
        float* t = W.tex[slot];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const int i = lane + 32 * k;
            if (i < TEXN) {
                float v = t[i];
                v -= chan_mean(c0, k, m0, m1, m2);
                v /= sg;
                t[i] = r[k] * v;
            }
        }
    }
    __syncwarp();
    if (lane < no) {
        const int slot = 1 + lane;
        const float* t = W.tex[slot];
        float a = 0.0f;
#pragma unroll 7
        for (int i = 0; i < TEXN; i++) a += t[i];
        W.dots[W.slot_view[slot]] = a / 147.0f;
    }
    __syncwarp();
}

// ----------------------------------------------------------------------------------------------------------
// The photometric core shared by objective_fn (:286-311) and setINCCs (:448-474):
// for patch context P (centre / normal / scale / view list) with view `refIdx` as reference, fill
// W.vvalid[k] (sampleTexture succeeded) and W.dots[k] = refTex.dot(tex_k) for every valid k != refIdx.
// If the reference view itself fails nothing else is sampled (both callers return early).
// `axes_ready`: P.xa/ya/za already hold calculatePatchAxis for (P.images[0], P.normal) - the optimizer lane
// computes them when it moves the patch, so the objective's critical path skips five normalisations.
// ----------------------------------------------------------------------------------------------------------
__device__ __noinline__ void eval_dots(Scratch& W, LaneCtx& P, const KParams& K, int lane, int refIdx, bool z_is_normal,
                                       bool axes_ready) {
    const long long t0 = HP_CLOCK();
    const int nimg = P.nimg;
    const f4 c = ld4(P.center);
    const f4 n = ld4(P.normal);
    const float scale = P.scale;
    f4 xa, ya, za;
    if (axes_ready) { xa = ld4(P.xa); ya = ld4(P.ya); za = ld4(P.za); }
    else patch_axes(K.cams[P.images[refIdx]], n, scale, xa, ya, za);
    const f4 zgate = z_is_normal ? n : za;     // setINCCs passes pNormal_, objective_fn passes pZaxis_ (:456 vs :292)
    bool ok = false;
    if (lane < nimg) {
        ViewSetup vs;
        ok = view_setup(K, K.cams[P.images[lane]], c, scale, xa, ya, zgate, vs);
        if (ok) W.vs[lane] = vs;
        W.vvalid[lane] = ok ? 1 : 0;
    }
    const unsigned vmask = __ballot_sync(FULL, ok);
    __syncwarp();
    HP_EVT(0, t0);
    if (!((vmask >> refIdx) & 1u)) return;
    // ordered list of valid non-reference views
    const unsigned omask = vmask & ~(1u << refIdx);
    if (ok && lane != refIdx) W.vlist[__popc(omask & ((1u << lane) - 1u))] = lane;
    const int nother = __popc(omask);
    if (lane == 0) { P.textures += 1 + nother; W.slot_view[0] = refIdx; }
    __syncwarp();
    // reference texture -> slot 0, then the others in groups of VC-1
    int done = 0;
    bool first = true;
    do {
        const int no = min(VC - 1, nother - done);
        if (lane < no) W.slot_view[1 + lane] = W.vlist[done + lane];
        __syncwarp();
        const long long t1 = HP_CLOCK();
        if (first) {
            sample_slots(W, 0, 1 + no, lane);
            __syncwarp();
            HP_EVT(1, t1);
            const long long t2 = HP_CLOCK();
            stats_slots(W, 0, 1 + no, lane);
            HP_EVT(2, t2);
            const long long t3 = HP_CLOCK();
            normalize_ref(W, lane);
            HP_EVT(3, t3);
        } else {
            sample_slots(W, 1, no, lane);
            stats_slots(W, 1, no, lane);
        }
        const long long t4 = HP_CLOCK();
        if (no > 0) dot_slots(W, no, lane);
        HP_EVT(4, t4);
        done += no;
        first = false;
    } while (done < nother);
}

// objective_fn's reduction (:294-310): robust value per view with lane = view, then the reference's sequential
// f64 sum over the views in list order
__device__ __forceinline__ double objective_value(Scratch& W, const LaneCtx& P, const KParams& K, int lane) {
    if (!W.vvalid[0]) return 2.0;
    const int nimg = P.nimg;
    if (lane >= 1 && lane < nimg && W.vvalid[lane]) {
        const float r = (float)(1.0 - (double)W.dots[lane]);
        W.incc[lane] = r / (1.0f + 3.0f * r);
    }
    __syncwarp();
    double val = 0.0;
    int nImgs = 0;
    for (int ii = 1; ii < nimg; ii++) {
        if (!W.vvalid[ii]) continue;
        val += (double)W.incc[ii];
        nImgs++;
    }
    __syncwarp();
    if (nImgs < K.opt.min_images_per_patch - 1) return 2.0;
    return val / nImgs;
}

// setINCCs (:448-474): lane = view
__device__ __forceinline__ void set_inccs(Scratch& W, LaneCtx& P, const KParams& K, int lane, int refIdx, int robust) {
    eval_dots(W, P, K, lane, refIdx, true, false);
    if (lane < P.nimg) {
        float v;
        if (!W.vvalid[refIdx]) v = 2.0f;
        else if (lane == refIdx) v = 0.0f;
        else if (!W.vvalid[lane]) v = 2.0f;
        else {
            const float r = 1.0f - W.dots[lane];
            v = robust ? r / (1.0f + 3.0f * r) : r;
        }
        W.incc[lane] = v;
    }
    __syncwarp();
}

// ordered compaction of the view list: keep[lane] for lane < nimg
__device__ __forceinline__ void compact_images(LaneCtx& P, int lane, bool keep) {
    const int n = P.nimg;
    const int img = (lane < n) ? P.images[lane] : 0;
    const unsigned mask = __ballot_sync(FULL, keep && lane < n);
    __syncwarp();
    if (keep && lane < n) P.images[__popc(mask & ((1u << lane) - 1u))] = (unsigned short)img;
    if (lane == 0) P.nimg = __popc(mask);
    __syncwarp();
}

// filterImagesNCC (:138-152)
__device__ __forceinline__ bool filter_images_ncc(Scratch& W, LaneCtx& P, const KParams& K, int lane, float threshold) {
    set_inccs(W, P, K, lane, 0, 0);
    const bool keep = (lane == 0) || (lane < P.nimg && W.incc[lane] < 1.0f - threshold);
    __syncwarp();
    compact_images(P, lane, keep);
    return P.nimg >= K.opt.min_images_per_patch;
}

// addImages (:225-258).  Returns 1 ok, 0 fail, -1 view list overflow.
__device__ __forceinline__ int add_images(LaneCtx& P, const KParams& K, int lane) {
    if (P.nimg <= 0) return 0;
    const int ref = P.images[0];
    const int n0 = P.nimg;
    const int beg = K.covis_off[ref], end = K.covis_off[ref + 1];
    const f4 c = ld4(P.center), nrm = ld4(P.normal);
    const float scale = P.scale;
    int count = n0;
    for (int base = beg; base < end; base += 32) {
        const int j = base + lane;
        bool keep = false;
        int cand = 0;
        if (j < end) {
            cand = K.covis_ids[j];
            keep = true;
            for (int i = 0; i < n0; i++) if (P.images[i] == cand) keep = false;
            if (keep) {
                const DevCamera& cam = K.cams[cand];
                const f4 d = sub4(ld4(cam.center), c);
                const float z = dot4(d, d);
                const float fz = sqrtf(z);
                f4 r = d;
                if (z > 0.0f) r = f4{d.x / fz, d.y / fz, d.z / fz, d.w / fz};
                if (dot4(r, nrm) < K.cos_max_f) keep = false;
                if (keep) {
                    const int lvl = (int)roundf(level_from(cam, fz, scale));
                    if (lvl < K.opt.minlevel || lvl >= K.opt.maxlevel - 2) keep = false;
                    if (keep) {
                        float u, v;
                        project(cam, c, lvl, u, v);
                        if (u < 0.0f || (float)(cam.w[lvl] - 1) <= u || v < 0.0f || (float)(cam.h[lvl] - 1) <= v) keep = false;
                    }
                }
            }
        }
        const unsigned mask = __ballot_sync(FULL, keep);
        const int pos = count + __popc(mask & ((1u << lane) - 1u));
        if (keep && pos < MAXV) P.images[pos] = (unsigned short)cand;
        count += __popc(mask);
    }
    __syncwarp();
    if (count > MAXV) return -1;
    if (lane == 0) P.nimg = count;
    __syncwarp();
    return count >= K.opt.min_images_per_patch ? 1 : 0;
}

// rays[i] = (cam_i.center - center).normalized() for all views; lane = view
__device__ __forceinline__ f4 view_ray(const LaneCtx& P, const KParams& K, int lane, float* fz_out) {
    const DevCamera& cam = K.cams[P.images[lane]];
    const f4 d = sub4(ld4(cam.center), ld4(P.center));
    const float z = dot4(d, d);
    const float fz = sqrtf(z);
    if (fz_out) *fz_out = fz;
    if (z > 0.0f) return f4{d.x / fz, d.y / fz, d.z / fz, d.w / fz};
    return d;
}

// sortImages + getAngleWeightedScales (:183-223, :260-284).  Return value is ignored by the reference (:54).
__device__ __forceinline__ void sort_images(Scratch& W, LaneCtx& P, const KParams& K, int lane) {
    const int nimg = P.nimg;
    if (nimg == 0) return;
    float fz0;
    {
        const DevCamera& cam0 = K.cams[P.images[0]];
        const f4 d = sub4(ld4(P.center), ld4(cam0.center));
        fz0 = sqrtf(dot4(d, d));
    }
    const int refLevel = max(0, min(K.opt.maxlevel - 1, (int)roundf(level_from(K.cams[P.images[0]], fz0, P.scale))));
    bool keep = false;
    f4 ray = f4{0, 0, 0, 0};
    float ws = 0.0f;
    int img = 0;
    if (lane < nimg) {
        img = P.images[lane];
        float fz;
        ray = view_ray(P, K, lane, &fz);
        const float cosa = dot4(ray, normalized4(ld4(P.normal)));
        if (cosa > 0.0f) {
            keep = true;
            const DevCamera& cam = K.cams[img];
            float sc;
            if (cam.ksum == 0.0f) sc = 1.0f;
            else sc = (float)(2.0 * (double)fz * (double)(1 << refLevel) / (double)cam.ksum);   // Camera::getScale
            ws = sc / cosa;
        }
    }
    const unsigned mask = __ballot_sync(FULL, keep);
    __syncwarp();
    if (keep) {
        const int p = __popc(mask & ((1u << lane) - 1u));
        W.tmp_i[p] = img; W.tmp_f[p] = ws;
        W.rays[p][0] = ray.x; W.rays[p][1] = ray.y; W.rays[p][2] = ray.z; W.rays[p][3] = ray.w;
    }
    __syncwarp();
    if (lane == 0) {
        int m = __popc(mask);
        int out = 0;
        if (m >= 2) {
            const float thr = K.sort_thr;
            W.tmp_f[0] = 0.0f;   // keep the reference image
            while (m > 0) {
                int index = 0;
                float best = W.tmp_f[0];
                for (int j = 1; j < m; j++) if (W.tmp_f[j] < best) { best = W.tmp_f[j]; index = j; }
                P.images[out++] = (unsigned short)W.tmp_i[index];
                const f4 ri = ld4(W.rays[index]);
                int jj = 0;
                for (int j = 0; j < m; j++) {
                    if (j == index) continue;
                    const f4 rj = ld4(W.rays[j]);
                    const float ftmp = fminf(thr, fmaxf(thr / 2.0f, 1.0f - dot4(ri, rj)));
                    const float wj = W.tmp_f[j] * (thr / ftmp);
                    W.tmp_i[jj] = W.tmp_i[j];
                    W.tmp_f[jj] = wj;
                    W.rays[jj][0] = rj.x; W.rays[jj][1] = rj.y; W.rays[jj][2] = rj.z; W.rays[jj][3] = rj.w;
                    jj++;
                }
                m = jj;
            }
        }
        P.nimg = out;   // pImages_.clear() happens before the size test (:190-193)
    }
    __syncwarp();
}

// assureImageAngles (:105-123)
__device__ __forceinline__ bool assure_image_angles(Scratch& W, LaneCtx& P, const KParams& K, int lane) {
    const int nimg = P.nimg;
    if (lane < nimg) {
        const f4 r = view_ray(P, K, lane, nullptr);
        W.rays[lane][0] = r.x; W.rays[lane][1] = r.y; W.rays[lane][2] = r.z; W.rays[lane][3] = r.w;
    }
    __syncwarp();
    bool found = false;
    const int npairs = nimg * (nimg - 1) / 2;
    for (int p = lane; p < npairs; p += 32) {
        // unrank pair p -> (ii < jj)
        int ii = 0, rem = p;
        while (rem >= nimg - 1 - ii) { rem -= nimg - 1 - ii; ii++; }
        const int jj = ii + 1 + rem;
        const float a = (float)acos((double)dot4(ld4(W.rays[ii]), ld4(W.rays[jj])));
        if (a < K.opt.max_angle && a > K.opt.min_angle) found = true;
    }
    const bool any = __any_sync(FULL, found);
    __syncwarp();
    return any;
}

// filterImagesByAngle (:125-136)
__device__ __forceinline__ bool filter_images_by_angle(LaneCtx& P, const KParams& K, int lane) {
    bool keep = false;
    if (lane < P.nimg) {
        const f4 r = view_ray(P, K, lane, nullptr);
        keep = dot4(r, ld4(P.normal)) > K.cos_max_f;
    }
    __syncwarp();
    compact_images(P, lane, keep);
    return P.nimg >= K.opt.min_images_per_patch;
}

// setRefImage (:154-181)
__device__ __forceinline__ void set_ref_image(Scratch& W, LaneCtx& P, const KParams& K, int lane) {
    const int nimg = P.nimg;
    if (nimg <= 1) return;
    int refindex = -1;
    float refncc = 3.402823466e+38f;
    for (int ii = 0; ii < nimg; ii++) {
        set_inccs(W, P, K, lane, ii, 1);
        float sum = 0.0f;
        for (int k = 0; k < nimg; k++) sum = sum + W.incc[k];
        if (sum < refncc) { refncc = sum; refindex = ii; }
        __syncwarp();
    }
    if (lane == 0 && refindex > 0) {
        const unsigned short t = P.images[0];
        P.images[0] = P.images[refindex];
        P.images[refindex] = t;
    }
    __syncwarp();
}

// setCenterNorm (:401-414): the owning lane writes its patch's centre / normal for parameters x
__device__ __forceinline__ void set_center_norm(LaneCtx& P, const KParams& K, const double* x) {
    const float x0 = (float)x[0];
    for (int i = 0; i < 4; i++) P.center[i] = P.refCenter[i] + (x0 * P.refRay[i]) * 1.0f;
    const float angle1 = (float)(x[1] * (double)K.angle_scale);
    const float angle2 = (float)(x[2] * (double)K.angle_scale);
    double s1, c1, s2, c2;
    sincos((double)angle1, &s1, &c1);
    sincos((double)angle2, &s2, &c2);
    const float fx = (float)(s1 * c2);
    const float fy = (float)s2;
    const float fz = (float)(-c1 * c2);
    for (int i = 0; i < 3; i++) P.normal[i] = (P.X0[i] * fx + P.Y0[i] * fy) + P.Z0[i] * fz;
    P.normal[3] = 0.0f;
    // calculatePatchAxis for the objective (objective_fn :289), done here lane-parallel for all moving patches
    f4 xa, ya, za;
    patch_axes(K.cams[P.images[0]], ld4(P.normal), P.scale, xa, ya, za);
    P.xa[0] = xa.x; P.xa[1] = xa.y; P.xa[2] = xa.z; P.xa[3] = 0.0f;
    P.ya[0] = ya.x; P.ya[1] = ya.y; P.ya[2] = ya.z; P.ya[3] = 0.0f;
    P.za[0] = za.x; P.za[1] = za.y; P.za[2] = za.z; P.za[3] = 0.0f;
}

// setOptimizationFields + parametersFromCenterNorm (:384-399, :416-446): owning lane only
__device__ __forceinline__ void init_parameters(LaneCtx& P, const KParams& K, const double* lb, const double* ub, double* x) {
    const DevCamera& cam = K.cams[P.images[0]];
    for (int i = 0; i < 3; i++) { P.X0[i] = cam.nx[i]; P.Y0[i] = cam.ny[i]; P.Z0[i] = cam.nz[i]; }
    for (int i = 0; i < 4; i++) P.refCenter[i] = P.center[i];
    const f4 rr = normalized4(sub4(ld4(P.refCenter), ld4(cam.center)));
    P.refRay[0] = rr.x; P.refRay[1] = rr.y; P.refRay[2] = rr.z; P.refRay[3] = rr.w;
    // c == refCenter here, so x[0] = 0 . refRay / depthScale
    x[0] = (double)(dot4(sub4(ld4(P.center), ld4(P.refCenter)), rr) / 1.0f);
    const f3 n3 = f3{P.normal[0], P.normal[1], P.normal[2]};
    const float fx = dot3(f3{P.X0[0], P.X0[1], P.X0[2]}, n3);
    const float fy = dot3(f3{P.Y0[0], P.Y0[1], P.Y0[2]}, n3);
    const float fz = dot3(f3{P.Z0[0], P.Z0[1], P.Z0[2]}, n3);
    x[2] = (double)(float)asin((double)fy);                       // std::asin(float), correctly rounded
    const float cosb = (float)cos(fmax(-1.0, fmin(1.0, x[2])));
    if (cosb == 0.0f) x[1] = 0.0;
    else {
        const double sina = (double)(fx / cosb);
        const double cosa = (double)(-fz / cosb);
        x[1] = acos(fmin(1.0, fmax(-1.0, cosa)));
        if (sina < 0.0) x[1] = -x[1];
    }
    x[1] /= (double)K.angle_scale;
    x[2] /= (double)K.angle_scale;
    for (int i = 0; i < 3; i++) x[i] = fmin(ub[i], fmax(lb[i], x[i]));
    // host-evaluated angles were formed with the camera axes of the INPUT reference view; sortImages (:183-223) drops views with
    // cosa <= 0, so the reference view can change - then the device values above stand (the host cannot know the new view)
    if (K.start && (int)P.images[0] == K.in[P.patch_index].images[0]) { x[1] = K.start[2 * P.patch_index]; x[2] = K.start[2 * P.patch_index + 1]; }
}

// Scene::getColor(const Patch3d&) (Scene.cpp:300-327): lane = view; stable rank by colour norm
__device__ __forceinline__ f3 patch_color(Scratch& W, LaneCtx& P, const KParams& K, int lane) {
    const int nimg = P.nimg;
    f3 col = f3{0, 0, 0};
    float nrm = 0.0f;
    if (lane < nimg) {
        const DevCamera& cam = K.cams[P.images[lane]];
        const f4 c = ld4(P.center);
        const f4 d = sub4(c, ld4(cam.center));
        const float fz = sqrtf(dot4(d, d));
        const int lvl = leveli_from(cam, fz, P.scale, cam.nlevels - 1);
        float u, v;
        project(cam, c, lvl, u, v);
        col = get_color(cam.img[lvl], cam.pitch[lvl], u, v);
        nrm = sqrtf(dot3(col, col));
        W.tmp_f[lane] = nrm;
    }
    __syncwarp();
    int rank = 0;
    if (lane < nimg)
        for (int j = 0; j < nimg; j++) {
            const float nj = W.tmp_f[j];
            if (nj < nrm || (nj == nrm && j < lane)) rank++;
        }
    const int mid = nimg / 2;
    const unsigned mmid = __ballot_sync(FULL, lane < nimg && rank == mid);
    const unsigned mlow = __ballot_sync(FULL, lane < nimg && rank == 0);
    const int lmid = __ffs(mmid) - 1, llow = __ffs(mlow) - 1;
    const float nmid = __shfl_sync(FULL, nrm, lmid);
    const int src = ((double)nmid > 250.0) ? llow : lmid;
    f3 o;
    o.x = __shfl_sync(FULL, col.x, src);
    o.y = __shfl_sync(FULL, col.y, src);
    o.z = __shfl_sync(FULL, col.z, src);
    __syncwarp();
    return o;
}

__device__ __forceinline__ void load_patch(LaneCtx& P, const hpmvs_patch_t& p, int pi, int lane) {
    if (lane < 4) { P.center[lane] = p.center[lane]; P.normal[lane] = p.normal[lane]; }
    if (lane == 0) {
        P.scale = p.scale; P.nimg = min(max(p.nimages, 0), MAXV); P.textures = 0; P.patch_index = pi;
        P.status = HPMVS_OK; P.nlopt_rc = 0; P.evals = 0; P.score = 0.0;
    }
    P.images[lane] = (unsigned short)p.images[lane];
    __syncwarp();
}

// runOptimization (:48-57): the stages before the refinement.  Returns HPMVS_OK when the patch may be refined.
__device__ __noinline__ int pre_stage(Scratch& W, LaneCtx& P, const KParams& K, int lane) {
    const int r = add_images(P, K, lane);
    if (r < 0) return HPMVS_FAIL_TOO_MANY_VIEWS;
    if (r == 0) return HPMVS_FAIL_ADD_IMAGES;
    if (!filter_images_ncc(W, P, K, lane, K.opt.ncc_alpha_1)) return HPMVS_FAIL_NCC1;
    sort_images(W, P, K, lane);
    if (!assure_image_angles(W, P, K, lane)) return HPMVS_FAIL_ANGLES;
    if (P.nimg < K.opt.min_images_per_patch) return HPMVS_FAIL_OPT_MINIMAGES;   // optimizePatch (:323)
    return HPMVS_OK;
}

// runOptimization (:62-75): the stages after a successful refinement
__device__ __noinline__ int post_stage(Scratch& W, LaneCtx& P, const KParams& K, int lane) {
    const int r = add_images(P, K, lane);
    if (r < 0) return HPMVS_FAIL_TOO_MANY_VIEWS;
    if (r == 0) return HPMVS_FAIL_ADD_IMAGES2;
    if (!filter_images_ncc(W, P, K, lane, K.opt.ncc_alpha_2)) return HPMVS_FAIL_NCC2;
    if (!filter_images_by_angle(P, K, lane)) return HPMVS_FAIL_ANGLE_FILTER;
    if (!assure_image_angles(W, P, K, lane)) return HPMVS_FAIL_ANGLES2;
    set_ref_image(W, P, K, lane);
    if (!filter_images_ncc(W, P, K, lane, K.opt.ncc_alpha_2)) return HPMVS_FAIL_NCC3;
    return HPMVS_OK;
}

// write one result record (PatchOptimizer.cpp:86-100); all lanes of the warp participate
__device__ __forceinline__ void retire_patch(Scratch& W, LaneCtx& P, const KParams& K, int lane, int status) {
    const int pi = P.patch_index;
    const hpmvs_patch_t& pin = K.in[pi];
    hpmvs_patch_t& po = K.out[pi];
    if (status == HPMVS_OK) {
        const f3 col = patch_color(W, P, K, lane);
        if (lane < 4) { po.center[lane] = P.center[lane]; po.normal[lane] = P.normal[lane]; }
        po.images[lane] = (lane < P.nimg) ? (int)P.images[lane] : 0;
        if (lane == 0) {
            po.scale = P.scale; po.nimages = P.nimg;
            po.color[0] = col.x; po.color[1] = col.y; po.color[2] = col.z;
            po.ncc = 1.4f;
        }
    } else {
        // rejected: geometry and view list stay as given (PatchOptimizer.cpp:86-93 runs only on success)
        if (lane < 4) { po.center[lane] = pin.center[lane]; po.normal[lane] = pin.normal[lane]; }
        po.images[lane] = pin.images[lane];
        if (lane == 0) {
            po.scale = pin.scale; po.nimages = pin.nimages;
            po.color[0] = 0.0f; po.color[1] = 0.0f; po.color[2] = 0.0f;
            po.ncc = 0.0f;
        }
    }
    if (lane == 0) {
        po.status = status; po.nlopt_result = P.nlopt_rc; po.evals = P.evals; po.textures = P.textures; po.score = P.score;
    }
    __syncwarp();
}

// ----------------------------------------------------------------------------------------------------------
// K2: the fused optimize kernel.  One persistent, warp-specialised CTA per SM:
//   * OPT_WARPS "optimizer" warps: lane = patch slot.  Each lane owns one in-flight patch and its FP64 BOBYQA
//     state (thread-private, local memory) and advances it in SIMT fashion.  The optimiser's large, branchy
//     instruction stream is therefore fetched by very few warps per SM.
//   * SAMPLER_WARPS "sampler" warps: serve requests from a shared-memory ticket queue with all 32 lanes
//     co-operating on one patch at a time: EVAL (objective at the slot's current centre/normal), POST (stages
//     after the refinement + result record + refill of the slot) and FILL (fetch a patch, stages before the
//     refinement).  Their hot loop is small and stays resident in the instruction cache.
// Slots hand over through a per-slot state word in shared memory; patches come from a global atomic counter.
// ----------------------------------------------------------------------------------------------------------
#ifndef HP_SLEEP_Q
#define HP_SLEEP_Q 64        // ns between polls of an idle sampler warp
#endif
#ifndef HP_SLEEP_OPT
#define HP_SLEEP_OPT 200     // ns between polls of an optimizer warp waiting for its objectives
#endif
constexpr int QCAP = 1024;                // >= slots + sampler warps outstanding entries, power of two

enum : int {
    ST_EMPTY = 0,        // slot has no patch; a FILL request is (about to be) queued
    ST_FILLING = 1,      // a sampler is fetching a patch / running the pre-stage
    ST_NEW = 2,          // patch loaded, pre-stage passed: optimizer must start()
    ST_EVAL_PENDING = 3, // objective requested
    ST_EVAL_DONE = 4,    // objective value available in fval[slot]
    ST_POSTING = 5,      // refinement finished, sampler runs the post-stage and retires the patch
    ST_DEAD = 6          // no more work for this slot
};
enum : int { REQ_FILL = 1, REQ_EVAL = 2, REQ_POST = 3, REQ_EXIT = 4 };

struct QueueShared {
    int queue[QCAP];
    unsigned q_head, q_tail;
    int opt_alive;
    int pad;
};

// BOBYQA state of one slot, padded so that the 8-byte stride between lanes is odd (conflict-free 64-bit accesses
// when all lanes of an optimizer warp touch the same field)
struct __align__(8) BqSlot {
    bq3::State s;
    double pad[((sizeof(bq3::State) / 8) % 2 == 0) ? 1 : 2];
};
static_assert((sizeof(BqSlot) / 8) % 2 == 1, "BqSlot stride must be an odd number of 8-byte words");

template <int OPT_WARPS, int SAMPLER_WARPS, int LPW>
struct __align__(16) CtaSharedT {
    static constexpr int NSLOTS = OPT_WARPS * LPW;   // LPW = lane slots used per optimizer warp (<= 32)
    BqSlot bq[NSLOTS];
    LaneCtx ctx[NSLOTS];
    double fval[NSLOTS];
    int sstate[NSLOTS];
    QueueShared Q;
    Scratch scratch[SAMPLER_WARPS];
};

__device__ __forceinline__ int ld_state(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_state(int* p, int v) {
    __threadfence_block();
    *reinterpret_cast<volatile int*>(p) = v;
}
__device__ __forceinline__ void q_push(QueueShared& C, int kind, int slot) {
    const unsigned t = atomicAdd(&C.q_tail, 1u);
    __threadfence_block();
    *reinterpret_cast<volatile int*>(&C.queue[t & (QCAP - 1)]) = (kind << 16) | slot;
}
// whole warp: take the next ticket and wait for its entry
__device__ __forceinline__ int q_pop(QueueShared& C, int lane) {
    int e = 0;
    if (lane == 0) {
        const unsigned h = atomicAdd(&C.q_head, 1u);
        volatile int* q = &C.queue[h & (QCAP - 1)];
        while ((e = *q) == 0) __nanosleep(HP_SLEEP_Q);
        *q = 0;
        __threadfence_block();
    }
    return __shfl_sync(FULL, e, 0);
}

// FILL: fetch patches until one passes the pre-stage (-> ST_NEW) or the queue is empty (-> ST_DEAD)
// returns ST_NEW (P holds a patch that passed the pre-stage) or ST_DEAD (work queue empty)
__device__ __noinline__ int serve_fill(LaneCtx& P, Scratch& W, const KParams& K, int lane, unsigned long long* cnt) {
    for (;;) {
        int pi = 0;
        if (lane == 0) pi = atomicAdd(K.work_counter, 1);
        pi = __shfl_sync(FULL, pi, 0);
        if (pi >= K.n) return ST_DEAD;
        load_patch(P, K.in[pi], pi, lane);
        // a view list longer than the engine's capacity is rejected, never truncated
        const int st = (K.in[pi].nimages > MAXV) ? (int)HPMVS_FAIL_TOO_MANY_VIEWS : pre_stage(W, P, K, lane);
        if (st == HPMVS_OK) {
            __syncwarp();
            return ST_NEW;
        }
        retire_patch(W, P, K, lane, st);
        if (lane == 0) { cnt[0]++; cnt[3] += P.textures; }
    }
}

template <int OPT_WARPS, int SAMPLER_WARPS, int LPW>
__global__ void __launch_bounds__((OPT_WARPS + SAMPLER_WARPS) * 32, 1) optimize_kernel(const KParams K) {
    using CtaShared = CtaSharedT<OPT_WARPS, SAMPLER_WARPS, LPW>;
    constexpr int NSLOTS = CtaShared::NSLOTS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CtaShared& C = *reinterpret_cast<CtaShared*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    // ---- CTA set-up ------------------------------------------------------------------------------------------
    for (int i = threadIdx.x; i < QCAP; i += blockDim.x) C.Q.queue[i] = 0;
    for (int i = threadIdx.x; i < NSLOTS; i += blockDim.x) C.sstate[i] = ST_EMPTY;
    if (threadIdx.x == 0) { C.Q.q_head = 0; C.Q.q_tail = 0; C.Q.opt_alive = OPT_WARPS; }
    __syncthreads();

    if (warp < OPT_WARPS) {
        // =========================================== optimizer warp ===========================================
        const bool has_slot = lane < LPW;
        const int slot = warp * LPW + (has_slot ? lane : 0);     // lanes without a slot alias slot 0 but never touch it
        LaneCtx& mine = C.ctx[slot];
        int* my_state = &C.sstate[slot];
        bq3::State& bq = C.bq[slot].s;   // shared memory: guaranteed on-chip (local memory thrashed L1, see profiles/)
        double xcur[3] = {0.0, 0.0, 0.0};
        // slots beyond the per-warp quota never receive work (small batches are spread over all SMs instead)
        if (has_slot) {
            if (lane < K.lanes_per_warp) { st_state(my_state, ST_FILLING); q_push(C.Q, REQ_FILL, slot); }
            else st_state(my_state, ST_DEAD);
        }
        long long t_wait = 0, t_adv = 0, n_rounds = 0, n_lanes = 0;
        for (;;) {
            // wait until every requested objective of this warp has arrived and at least one lane can move
            int st;
            const long long tw0 = HP_CLOCK();
            for (;;) {
                st = has_slot ? ld_state(my_state) : ST_DEAD;
                const unsigned waiting = __ballot_sync(FULL, st == ST_EVAL_PENDING);
                const unsigned ready = __ballot_sync(FULL, st == ST_EVAL_DONE || st == ST_NEW);
                const unsigned dead = __ballot_sync(FULL, st == ST_DEAD);
                if (dead == FULL) { st = -1; break; }
                if (!waiting && ready) break;
                __nanosleep(HP_SLEEP_OPT);
            }
            if (st == -1) break;
            __threadfence_block();
            const long long ta0 = HP_CLOCK();
            t_wait += ta0 - tw0; n_rounds++; n_lanes += __popc(__ballot_sync(FULL, st == ST_EVAL_DONE || st == ST_NEW));
            if (st == ST_NEW) {
                const double lb[3] = {-HUGE_VAL, -23.99999, -23.99999};
                const double ub[3] = {HUGE_VAL, 23.99999, 23.99999};
                double x0[3];
                init_parameters(mine, K, lb, ub, x0);
                const int act = bq3::start(bq, x0, lb, ub, 1.e-7, 1000, xcur);
                if (act == bq3::ASK) { set_center_norm(mine, K, xcur); st = ST_EVAL_PENDING; }
                else st = ST_POSTING;
            } else if (st == ST_EVAL_DONE) {
                const double f = *reinterpret_cast<volatile double*>(&C.fval[slot]);
                const int act = bq3::advance(bq, f, xcur);
                if (act == bq3::ASK) { set_center_norm(mine, K, xcur); st = ST_EVAL_PENDING; }
                else if (act == bq3::YIELD) st = ST_EVAL_DONE;     // ready again next round, no objective needed
                else st = ST_POSTING;
            } else {
                st = 0;   // FILLING / POSTING / DEAD: nothing to do for this lane in this round
            }
            if (st == ST_POSTING) {
                // optimizePatch's epilogue (:364-381)
                const int rc = bq.rc;
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
            if (st == ST_EVAL_PENDING) { st_state(my_state, ST_EVAL_PENDING); q_push(C.Q, REQ_EVAL, slot); }
            else if (st == ST_POSTING) { st_state(my_state, ST_POSTING); q_push(C.Q, REQ_POST, slot); }
            // st == ST_EVAL_DONE (yield): the slot state is still ST_EVAL_DONE, it takes part in the next round as it is
            __syncwarp();
            t_adv += HP_CLOCK() - ta0;
        }
#ifdef HP_PROFILE
        if (lane == 0) {
            atomicAdd(&K.counters[4], (unsigned long long)t_wait); atomicAdd(&K.counters[5], (unsigned long long)t_adv);
            atomicAdd(&K.counters[6], (unsigned long long)n_rounds); atomicAdd(&K.counters[7], (unsigned long long)n_lanes);
        }
#endif
        // the last optimizer warp to finish releases the samplers
        __syncwarp();
        if (lane == 0) {
            if (atomicSub(&C.Q.opt_alive, 1) == 1)
                for (int i = 0; i < SAMPLER_WARPS; i++) q_push(C.Q, REQ_EXIT, 0);
        }
    } else {
        // ============================================ sampler warp ============================================
        Scratch& W = C.scratch[warp - OPT_WARPS];
        unsigned long long cnt[4] = {0, 0, 0, 0};   // patches, ok, evals, textures (lane 0 only)
        long long t_idle = 0, t_eval = 0, t_other = 0, n_eval = 0;
        for (;;) {
            const long long tq0 = HP_CLOCK();
            const int e = q_pop(C.Q, lane);
            const long long tq1 = HP_CLOCK();
            t_idle += tq1 - tq0;
            const int kind = e >> 16, slot = e & 0xffff;
            if (kind == REQ_EXIT) break;
            LaneCtx& P = C.ctx[slot];
            if (kind == REQ_EVAL) {
                eval_dots(W, P, K, lane, 0, false, true);
                const double f = objective_value(W, P, K, lane);
                __syncwarp();
                if (lane == 0) {
                    *reinterpret_cast<volatile double*>(&C.fval[slot]) = f;
                    st_state(&C.sstate[slot], ST_EVAL_DONE);
                }
                t_eval += HP_CLOCK() - tq1; n_eval++;
            } else if (kind == REQ_POST) {
                int st = P.status;
                if (st == HPMVS_OK) st = post_stage(W, P, K, lane);
                retire_patch(W, P, K, lane, st);
                if (lane == 0) { cnt[0]++; cnt[1] += (st == HPMVS_OK); cnt[2] += P.evals; cnt[3] += P.textures; }
                __syncwarp();
                const int ns = serve_fill(P, W, K, lane, cnt);
                if (lane == 0) st_state(&C.sstate[slot], ns);
                t_other += HP_CLOCK() - tq1;
            } else {   // REQ_FILL
                const int ns = serve_fill(P, W, K, lane, cnt);
                if (lane == 0) st_state(&C.sstate[slot], ns);
                t_other += HP_CLOCK() - tq1;
            }
            __syncwarp();
        }
        if (lane == 0 && cnt[0]) {
            atomicAdd(&K.counters[0], cnt[0]); atomicAdd(&K.counters[1], cnt[1]);
            atomicAdd(&K.counters[2], cnt[2]); atomicAdd(&K.counters[3], cnt[3]);
        }
#ifdef HP_PROFILE
        if (lane == 0) {
            atomicAdd(&K.counters[8], (unsigned long long)t_idle); atomicAdd(&K.counters[9], (unsigned long long)t_eval);
            atomicAdd(&K.counters[10], (unsigned long long)n_eval); atomicAdd(&K.counters[12], (unsigned long long)t_other);
        }
#endif
        (void)t_other;
    }
}

// ----------------------------------------------------------------------------------------------------------
// K2p: the same kernel with PARKED patch slots.  Shared memory caps the resident variant at 64 patches per SM
// (1.6 KB of FP64 optimizer state each) and throughput = patches in flight / round latency.  Here every CTA owns
// `vslots` (<= VMAX) virtual slots whose optimizer state and patch context live in an HBM/L2 pool; the shared-memory
// slots are only staging buffers: an optimizer warp claims up to LPW READY slots, copies their state in (16-byte
// vectors, L2-resident), advances them lane-parallel, copies them back and publishes the objective requests.  While
// those wait for the samplers the warp is already advancing another batch, so it never idles as long as enough
// patches are in flight.  Samplers work on a private shared-memory copy of the patch context.
// ----------------------------------------------------------------------------------------------------------
constexpr int VMAX = 256;
constexpr int MIN_BATCH = 12;            // lanes worth starting an optimizer round for while objectives are still pending

// Parked variant: the staging slots are filled and drained by the bulk-async copy engine (cp.async.bulk, SASS UBLKCP), which needs
// 16-byte aligned addresses and sizes: the slot is the bare State (1616 B = 16 * 101).  Its stride is an even number of 8-byte words,
// so lane-parallel FP64 accesses take 2-way bank conflicts here (the resident variant keeps the conflict-free odd stride) - the
// optimizer warps are latency bound, the copies were 13 % of their time.
#ifndef HP_PARKED_TMA
#define HP_PARKED_TMA 1
#endif
#if HP_PARKED_TMA
struct __align__(16) BqSlotP { bq3::State s; };
static_assert(sizeof(BqSlotP) % 16 == 0 && sizeof(LaneCtx) % 16 == 0, "bulk copies move multiples of 16 bytes");
#else
typedef BqSlot BqSlotP;
#endif

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk group
__device__ __forceinline__ void bulk_s2g(void* gmem, const void* smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem), "r"(smem_u32(smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_all() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int OPT_WARPS, int SAMPLER_WARPS, int LPW>
struct __align__(16) CtaSharedP {
    BqSlotP stage_bq[OPT_WARPS * LPW];
    LaneCtx stage_ctx[OPT_WARPS * LPW];
    int claim[OPT_WARPS][32];
    double fval[VMAX];
    int sstate[VMAX];
    QueueShared Q;
    struct Samp { Scratch S; LaneCtx P; } samp[SAMPLER_WARPS];
    unsigned long long mbar[OPT_WARPS];          // one transaction barrier per optimizer warp (stage-in of a round)
};

// warp-cooperative copies between the pool (global, read/written around L1) and shared memory, 16 bytes per lane
// (8-byte granules: the padded optimizer slots are only 8-byte aligned)
template <int BYTES>
__device__ __forceinline__ void copy_in(void* smem, const void* gmem, int lane) {
    static_assert(BYTES % 8 == 0, "8-byte granules");
    const uint2* g = reinterpret_cast<const uint2*>(gmem);
    uint2* s = reinterpret_cast<uint2*>(smem);
#pragma unroll
    for (int i = lane; i < BYTES / 8; i += 32) s[i] = __ldcg(g + i);
}
template <int BYTES>
__device__ __forceinline__ void copy_out(void* gmem, const void* smem, int lane) {
    static_assert(BYTES % 8 == 0, "8-byte granules");
    uint2* g = reinterpret_cast<uint2*>(gmem);
    const uint2* s = reinterpret_cast<const uint2*>(smem);
#pragma unroll
    for (int i = lane; i < BYTES / 8; i += 32) __stcg(g + i, s[i]);
}
constexpr int BQ_COPY_BYTES = (int)sizeof(bq3::State);
static_assert(sizeof(bq3::State) % 8 == 0 && sizeof(LaneCtx) % 8 == 0, "pool copies in 8-byte granules");

template <int OPT_WARPS, int SAMPLER_WARPS, int LPW>
__global__ void __launch_bounds__((OPT_WARPS + SAMPLER_WARPS) * 32, 1) optimize_kernel_parked(const KParams K) {
    using CtaShared = CtaSharedP<OPT_WARPS, SAMPLER_WARPS, LPW>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CtaShared& C = *reinterpret_cast<CtaShared*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int V = K.vslots;                       // multiple of OPT_WARPS
    BqSlotP* pool_bq = reinterpret_cast<BqSlotP*>(K.pool_bq) + (size_t)blockIdx.x * VMAX;
    LaneCtx* pool_ctx = K.pool_ctx + (size_t)blockIdx.x * VMAX;

    for (int i = threadIdx.x; i < QCAP; i += blockDim.x) C.Q.queue[i] = 0;
    for (int i = threadIdx.x; i < VMAX; i += blockDim.x) C.sstate[i] = ST_DEAD;
    if (threadIdx.x == 0) {
        C.Q.q_head = 0; C.Q.q_tail = 0; C.Q.opt_alive = OPT_WARPS;
#if HP_PARKED_TMA
        for (int i = 0; i < OPT_WARPS; i++) mbar_init(&C.mbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    }
    __syncthreads();

    if (warp < OPT_WARPS) {
        // =========================================== optimizer warp ===========================================
        const int VW = V / OPT_WARPS, v0 = warp * VW;
        for (int i = lane; i < VW; i += 32) { st_state(&C.sstate[v0 + i], ST_FILLING); q_push(C.Q, REQ_FILL, v0 + i); }
        __syncwarp();
        LaneCtx& mine = C.stage_ctx[warp * LPW + (lane < LPW ? lane : 0)];
        bq3::State& bq = C.stage_bq[warp * LPW + (lane < LPW ? lane : 0)].s;
        int tries = 0;
        unsigned mbar_phase = 0;
        (void)mbar_phase;
        long long t_wait = 0, t_adv = 0, t_copy = 0, n_rounds = 0, n_lanes = 0;
        long long tw0 = HP_CLOCK();
        for (;;) {
            // ---- claim up to LPW ready slots of this warp's range ---------------------------------------------
            int count = 0, npend = 0, nlive = 0;
            for (int base = 0; base < VW; base += 32) {
                const int i = base + lane;
                const int st = (i < VW) ? ld_state(&C.sstate[v0 + i]) : ST_DEAD;
                const bool ready = (st == ST_EVAL_DONE || st == ST_NEW);
                const unsigned m = __ballot_sync(FULL, ready);
                const int pos = count + __popc(m & ((1u << lane) - 1u));
                if (ready && pos < LPW) C.claim[warp][pos] = (v0 + i) | (st == ST_NEW ? 0x10000 : 0);
                count = min(LPW, count + __popc(m));
                npend += __popc(__ballot_sync(FULL, st == ST_EVAL_PENDING));
                nlive += __popc(__ballot_sync(FULL, st != ST_DEAD));
            }
            __syncwarp();
            if (count == 0) {
                if (nlive == 0) break;
                __nanosleep(200);
                continue;
            }
            if (count < MIN_BATCH && npend > 0 && tries < 40) { ++tries; __nanosleep(500); continue; }
            tries = 0;
            __threadfence_block();
            const long long tc0 = HP_CLOCK();
            t_wait += tc0 - tw0; n_rounds++; n_lanes += count;
            // ---- stage in: context always, optimizer state unless the patch is new -------------------------------
#if HP_PARKED_TMA
            {
                // lane j fetches slot j: two bulk-async copies (context, and the optimizer state unless the patch is new) that complete
                // on the warp's transaction barrier; one wait for the whole round instead of one L2 round trip per slot
                const int ej = (lane < count) ? C.claim[warp][lane] : 0;
                const unsigned mybytes = (lane < count) ? (unsigned)sizeof(LaneCtx) + ((ej >> 16) ? 0u : (unsigned)sizeof(BqSlotP)) : 0u;
                const unsigned total = __reduce_add_sync(FULL, mybytes);
                fence_proxy_async();      // the pools were last written through the generic proxy (samplers) or by our own bulk stores
                if (lane == 0) mbar_expect_tx(&C.mbar[warp], total);
                __syncwarp();
                if (lane < count) {
                    const int sl = ej & 0xffff;
                    bulk_g2s(&C.stage_ctx[warp * LPW + lane], &pool_ctx[sl], (unsigned)sizeof(LaneCtx), &C.mbar[warp]);
                    if (!(ej >> 16)) bulk_g2s(&C.stage_bq[warp * LPW + lane], &pool_bq[sl], (unsigned)sizeof(BqSlotP), &C.mbar[warp]);
                }
                mbar_wait(&C.mbar[warp], mbar_phase);
                mbar_phase ^= 1u;
            }
#else
            for (int j = 0; j < count; j++) {
                const int e = C.claim[warp][j], slot = e & 0xffff;
                copy_in<(int)sizeof(LaneCtx)>(&C.stage_ctx[warp * LPW + j], &pool_ctx[slot], lane);
                if (!(e >> 16)) copy_in<BQ_COPY_BYTES>(&C.stage_bq[warp * LPW + j], &pool_bq[slot], lane);
            }
#endif
            __syncwarp();
            const long long ta0 = HP_CLOCK();
            t_copy += ta0 - tc0;
            // ---- advance lane-parallel ------------------------------------------------------------------------------
            int st = 0, slot = 0;
            if (lane < count) {
                const int e = C.claim[warp][lane];
                slot = e & 0xffff;
                double xcur[3];
                int act;
                if (e >> 16) {
                    const double lb[3] = {-HUGE_VAL, -23.99999, -23.99999};
                    const double ub[3] = {HUGE_VAL, 23.99999, 23.99999};
                    double x0[3];
                    init_parameters(mine, K, lb, ub, x0);
                    act = bq3::start(bq, x0, lb, ub, 1.e-7, 1000, xcur);
                } else {
                    const double f = *reinterpret_cast<volatile double*>(&C.fval[slot]);
                    act = bq3::advance(bq, f, xcur);
                }
                if (act == bq3::ASK) { set_center_norm(mine, K, xcur); st = ST_EVAL_PENDING; }
                else if (act == bq3::YIELD) st = ST_EVAL_DONE;  // ready again next round, no objective needed
                else {
                    st = ST_POSTING;
                    const int rc = bq.rc;                      // optimizePatch's epilogue (:364-381)
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
            }
            __syncwarp();
            const long long to0 = HP_CLOCK();
            t_adv += to0 - ta0;
            // ---- stage out and publish ------------------------------------------------------------------------------
#if HP_PARKED_TMA
            // lane j wrote slot j's state and context (generic proxy); make that visible to the async proxy, then drain both with bulk
            // stores and wait for their completion before the slot is published
            fence_proxy_async();
            if (lane < count) {
                bulk_s2g(&pool_ctx[slot], &C.stage_ctx[warp * LPW + lane], (unsigned)sizeof(LaneCtx));
                bulk_s2g(&pool_bq[slot], &C.stage_bq[warp * LPW + lane], (unsigned)sizeof(BqSlotP));
            }
            bulk_commit_wait_all();
            fence_proxy_async();
#else
            for (int j = 0; j < count; j++) {
                const int sl = C.claim[warp][j] & 0xffff;
                copy_out<(int)sizeof(LaneCtx)>(&pool_ctx[sl], &C.stage_ctx[warp * LPW + j], lane);
                copy_out<BQ_COPY_BYTES>(&pool_bq[sl], &C.stage_bq[warp * LPW + j], lane);
            }
#endif
            __threadfence();
            __syncwarp();
            if (lane < count) {
                st_state(&C.sstate[slot], st);
                if (st != ST_EVAL_DONE) q_push(C.Q, st == ST_EVAL_PENDING ? REQ_EVAL : REQ_POST, slot);
            }
            __syncwarp();
            tw0 = HP_CLOCK();
            t_copy += tw0 - to0;
        }
#ifdef HP_PROFILE
        if (lane == 0) {
            atomicAdd(&K.counters[4], (unsigned long long)t_wait); atomicAdd(&K.counters[5], (unsigned long long)t_adv);
            atomicAdd(&K.counters[6], (unsigned long long)n_rounds); atomicAdd(&K.counters[7], (unsigned long long)n_lanes);
            atomicAdd(&K.counters[11], (unsigned long long)t_copy);
        }
#endif
        __syncwarp();
        if (lane == 0) {
            if (atomicSub(&C.Q.opt_alive, 1) == 1)
                for (int i = 0; i < SAMPLER_WARPS; i++) q_push(C.Q, REQ_EXIT, 0);
        }
    } else {
        // ============================================ sampler warp ============================================
        Scratch& W = C.samp[warp - OPT_WARPS].S;
        LaneCtx& P = C.samp[warp - OPT_WARPS].P;
        unsigned long long cnt[4] = {0, 0, 0, 0};
        long long t_idle = 0, t_eval = 0, t_other = 0, n_eval = 0;
        for (;;) {
            const long long tq0 = HP_CLOCK();
            const int e = q_pop(C.Q, lane);
            const long long tq1 = HP_CLOCK();
            t_idle += tq1 - tq0;
            const int kind = e >> 16, slot = e & 0xffff;
            if (kind == REQ_EXIT) break;
            int ns = -1;
            if (kind == REQ_EVAL) {
                copy_in<(int)sizeof(LaneCtx)>(&P, &pool_ctx[slot], lane);
                __syncwarp();
                const int tex0 = P.textures;
                eval_dots(W, P, K, lane, 0, false, true);
                const double f = objective_value(W, P, K, lane);
                __syncwarp();
                if (lane == 0) {
                    if (P.textures != tex0) __stcg(&pool_ctx[slot].textures, P.textures);
                    __threadfence();
                    *reinterpret_cast<volatile double*>(&C.fval[slot]) = f;
                    st_state(&C.sstate[slot], ST_EVAL_DONE);
                }
                t_eval += HP_CLOCK() - tq1; n_eval++;
            } else {
                if (kind == REQ_POST) {
                    copy_in<(int)sizeof(LaneCtx)>(&P, &pool_ctx[slot], lane);
                    __syncwarp();
                    int st = P.status;
                    if (st == HPMVS_OK) st = post_stage(W, P, K, lane);
                    retire_patch(W, P, K, lane, st);
                    if (lane == 0) { cnt[0]++; cnt[1] += (st == HPMVS_OK); cnt[2] += P.evals; cnt[3] += P.textures; }
                    __syncwarp();
                }
                ns = serve_fill(P, W, K, lane, cnt);
                __syncwarp();
                if (ns == ST_NEW) copy_out<(int)sizeof(LaneCtx)>(&pool_ctx[slot], &P, lane);
                __threadfence();
                __syncwarp();
                if (lane == 0) st_state(&C.sstate[slot], ns);
                t_other += HP_CLOCK() - tq1;
            }
            __syncwarp();
        }
        if (lane == 0 && cnt[0]) {
            atomicAdd(&K.counters[0], cnt[0]); atomicAdd(&K.counters[1], cnt[1]);
            atomicAdd(&K.counters[2], cnt[2]); atomicAdd(&K.counters[3], cnt[3]);
        }
#ifdef HP_PROFILE
        if (lane == 0) {
            atomicAdd(&K.counters[8], (unsigned long long)t_idle); atomicAdd(&K.counters[9], (unsigned long long)t_eval);
            atomicAdd(&K.counters[10], (unsigned long long)n_eval); atomicAdd(&K.counters[12], (unsigned long long)t_other);
        }
#endif
        (void)t_other; (void)t_eval; (void)n_eval; (void)t_idle;
    }
}

// ----------------------------------------------------------------------------------------------------------
// K1: setINCCs for a batch (parity vehicle for the photometric core)
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) ncc_kernel(const KParams K, int ref_idx, int robust, float* inccs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NccWarp& WS = reinterpret_cast<NccWarp*>(smem_raw)[threadIdx.x >> 5];
    Scratch& W = WS.S;
    LaneCtx& P = WS.P;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long c_tex = 0;
    for (int pi = warp; pi < K.n; pi += nwarps) {
        load_patch(P, K.in[pi], pi, lane);
        float v = 2.0f;
        if (ref_idx < P.nimg) {
            set_inccs(W, P, K, lane, ref_idx, robust);
            if (lane < P.nimg) v = W.incc[lane];
        }
        inccs[(size_t)pi * MAXV + lane] = (lane < P.nimg) ? v : 0.0f;
        if (lane == 0) c_tex += P.textures;
        __syncwarp();
    }
    if (lane == 0 && c_tex) atomicAdd(&K.counters[3], c_tex);
}

// ----------------------------------------------------------------------------------------------------------
// K1-staged: the same scoring (setINCCs for a batch) with the per-view image window of every texture STAGED IN SHARED MEMORY BY THE
// BULK-ASYNC COPY ENGINE (north_star: "TMA staging per-view tiles into shared memory") - the A/B that closes the question
// (profiles/r2_tma_ab.md).  For every texture of a group the TWR rows of a TWC-pixel window at the footprint's
// corner (start column rounded down to 16 bytes) are fetched with cp.async.bulk (SASS UBLKCP, the engine the parked kernel already
// uses) onto the warp's mbarrier - 4 row copies per lane; after the wait the 49 x 4 bilinear taps read the window with LDS instead
// of LDG.  A footprint that does not fit the window falls back to the global loads for that view.  Same arithmetic, same results.
// (A first version used one 2-D tensor map per view and level and cp.async.bulk.tensor.2d, SASS UTMALDG: it compiled, but every launch
// died with "illegal instruction" at the UTMALDG - with the descriptors in global memory and the boxes issued from a loop - and
// could not be debugged further inside this round's GPU budget; the row form moves the same bytes through the same engine.)
// ----------------------------------------------------------------------------------------------------------
constexpr int TWC = 20, TWR = 16;         // window: 20 pixels (80 bytes, 16-byte multiples) x 16 rows
struct __align__(128) NccWarpTma {
    uchar4 win[VC][TWR * TWC];            // 10 KB
    Scratch S;
    LaneCtx P;
    unsigned long long mbar;
    int wx0[VC], wy0[VC], wok[VC];
};
// Image::getColor on a staged window (same taps, same accumulation order as get_color)
__device__ __forceinline__ f3 get_color_win(const uchar4* win, int wx0, int wy0, float x, float y) {
    const int lx = (int)x, ly = (int)y;
    const float dx1 = x - (float)lx, dx0 = 1.0f - dx1;
    const float dy1 = y - (float)ly, dy0 = 1.0f - dy1;
    const float f00 = dx0 * dy0, f01 = dx0 * dy1, f10 = dx1 * dy0, f11 = dx1 * dy1;
    const uchar4* p = win + (ly - wy0) * TWC + (lx - wx0);
    const uchar4 a = p[0], b = p[1], c = p[TWC], d = p[TWC + 1];
    f3 o;
    o.x = ((float)a.x * f00 + (float)c.x * f01) + ((float)b.x * f10 + (float)d.x * f11);
    o.y = ((float)a.y * f00 + (float)c.y * f01) + ((float)b.y * f10 + (float)d.y * f11);
    o.z = ((float)a.z * f00 + (float)c.z * f01) + ((float)b.z * f10 + (float)d.z * f11);
    return o;
}
__device__ __forceinline__ void sample_one_win(NccWarpTma& Wt, int idx) {
    Scratch& W = Wt.S;
    const int slot = idx / 49, s = idx - 49 * slot;
    const ViewSetup& v = W.vs[W.slot_view[slot]];
    const int yy = s / 7, xx = s - 7 * yy;
    const float dyx = v.dyx, dyy = v.dyy, dxx = v.dxx, dxy = v.dxy;
    float px = v.tlx, py = v.tly;
#pragma unroll
    for (int i = 0; i < 6; i++) if (i < yy) { px += dyx; py += dyy; }
#pragma unroll
    for (int i = 0; i < 6; i++) if (i < xx) { px += dxx; py += dxy; }
    const f3 col = Wt.wok[slot] ? get_color_win(Wt.win[slot], Wt.wx0[slot], Wt.wy0[slot], px, py) : get_color(v.img, v.pitch, px, py);
    float* t = W.tex[slot] + 3 * s;
    t[0] = col.x; t[1] = col.y; t[2] = col.z;
}
// stage the windows of slots [first, first+ns)
__device__ __forceinline__ void stage_windows(NccWarpTma& Wt, const LaneCtx& P, const KParams& K, int first, int ns,
                                              int lane, unsigned& phase, unsigned long long* fallbacks) {
    Scratch& W = Wt.S;
    bool fits = false;
    if (lane < ns) {
        const int slot = first + lane;
        const ViewSetup& v = W.vs[W.slot_view[slot]];
        // the footprint's bounding box from its four corners (the grid is affine in the sample index)
        const float ex = 6.0f * v.dxx, ey = 6.0f * v.dxy, fx = 6.0f * v.dyx, fy = 6.0f * v.dyy;
        const float mnx = v.tlx + fminf(0.0f, ex) + fminf(0.0f, fx) - 0.01f, mxx = v.tlx + fmaxf(0.0f, ex) + fmaxf(0.0f, fx) + 0.01f;
        const float mny = v.tly + fminf(0.0f, ey) + fminf(0.0f, fy) - 0.01f, mxy = v.tly + fmaxf(0.0f, ey) + fmaxf(0.0f, fy) + 0.01f;
        const int x0 = ((int)mnx) & ~3, y0 = (int)mny;               // 16-byte aligned start column (rows are 16-byte aligned: pitch % 4 == 0)
        fits = ((int)mxx + 1 < x0 + TWC) && ((int)mxy + 1 < y0 + TWR) && mnx >= 0.0f && mny >= 0.0f &&
               x0 + TWC <= v.pitch && y0 + TWR <= v.pad + 2;      // the window stays inside the level's allocation (pitch x (rows + 2))
        Wt.wx0[slot] = x0; Wt.wy0[slot] = y0; Wt.wok[slot] = fits ? 1 : 0;
        if (!fits && fallbacks) atomicAdd(fallbacks, 1ull);
    }
    const unsigned m = __ballot_sync(FULL, fits);
    __syncwarp();
    if (m) {
        fence_proxy_async();       // the windows were last read through the generic proxy
        if (lane == 0) mbar_expect_tx(&Wt.mbar, (unsigned)__popc(m) * TWR * TWC * 4u);
        __syncwarp();
        // lane = (slot, quarter of the rows): up to 8 slots x 4 quarters, 4 row copies each
        const int sj = lane >> 2, q = lane & 3;
        if (sj < ns && ((m >> sj) & 1u)) {
            const int slot = first + sj;
            const ViewSetup& v = W.vs[W.slot_view[slot]];
            const int x0 = Wt.wx0[slot], y0 = Wt.wy0[slot];
#pragma unroll
            for (int r = 4 * q; r < 4 * q + 4; r++)
                bulk_g2s(&Wt.win[slot][r * TWC], v.img + (size_t)(y0 + r) * v.pitch + x0, TWC * 4u, &Wt.mbar);
        }
        // bounded wait: a faulting copy must end in a trap, never in a hang
        unsigned okw = 0;
        for (unsigned tries = 0; !okw; tries++) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(okw) : "r"(smem_u32(&Wt.mbar)), "r"(phase) : "memory");
            if (tries > (1u << 24)) __trap();
        }
        phase ^= 1u;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) ncc_kernel_tma(const KParams K, int ref_idx, int robust,
                                                                       float* inccs, unsigned long long* fallbacks) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    NccWarpTma& Wt = reinterpret_cast<NccWarpTma*>(smem_raw)[threadIdx.x >> 5];
    Scratch& W = Wt.S;
    LaneCtx& P = Wt.P;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    if (lane == 0) { mbar_init(&Wt.mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    unsigned phase = 0;
    unsigned long long c_tex = 0;
    for (int pi = warp; pi < K.n; pi += nwarps) {
        load_patch(P, K.in[pi], pi, lane);
        const int nimg = P.nimg;
        float v = 2.0f;
        if (ref_idx < nimg) {
            // set_inccs / eval_dots (above) with the sampling stage reading staged windows
            const f4 c = ld4(P.center), n = ld4(P.normal);
            f4 xa, ya, za;
            patch_axes(K.cams[P.images[ref_idx]], n, P.scale, xa, ya, za);
            bool ok = false;
            if (lane < nimg) {
                ViewSetup vs;
                ok = view_setup(K, K.cams[P.images[lane]], c, P.scale, xa, ya, n, vs);
                if (ok) W.vs[lane] = vs;
                W.vvalid[lane] = ok ? 1 : 0;
            }
            const unsigned vmask = __ballot_sync(FULL, ok);
            __syncwarp();
            if ((vmask >> ref_idx) & 1u) {
                const unsigned omask = vmask & ~(1u << ref_idx);
                if (ok && lane != ref_idx) W.vlist[__popc(omask & ((1u << lane) - 1u))] = lane;
                const int nother = __popc(omask);
                if (lane == 0) { P.textures += 1 + nother; W.slot_view[0] = ref_idx; }
                __syncwarp();
                int done = 0;
                bool first = true;
                do {
                    const int no = min(VC - 1, nother - done);
                    if (lane < no) W.slot_view[1 + lane] = W.vlist[done + lane];
                    __syncwarp();
                    const int s0 = first ? 0 : 1, ns = first ? 1 + no : no;
                    stage_windows(Wt, P, K, s0, ns, lane, phase, fallbacks);
                    for (int base = s0 * 49 + lane; base < (s0 + ns) * 49; base += 32) sample_one_win(Wt, base);
                    __syncwarp();
                    stats_slots(W, s0, ns, lane);
                    if (first) normalize_ref(W, lane);
                    if (no > 0) dot_slots(W, no, lane);
                    done += no;
                    first = false;
                } while (done < nother);
            }
            if (lane < nimg) {
                if (!W.vvalid[ref_idx]) v = 2.0f;
                else if (lane == ref_idx) v = 0.0f;
                else if (!W.vvalid[lane]) v = 2.0f;
                else { const float r = 1.0f - W.dots[lane]; v = robust ? r / (1.0f + 3.0f * r) : r; }
            }
        }
        inccs[(size_t)pi * MAXV + lane] = (lane < nimg) ? v : 0.0f;
        if (lane == 0) c_tex += P.textures;
        __syncwarp();
    }
    if (lane == 0 && c_tex) atomicAdd(&K.counters[3], c_tex);
}

// ----------------------------------------------------------------------------------------------------------
// K5 ("next" row f-2): depth-map bookkeeping and the acceptance tests that follow optimize() in
// CellProcessor::extend (src/hpmvs/CellProcessor.cpp:134-142, 197-201): Scene::setDepths (Scene.cpp:351-381),
// depthTests / depthTest (:518-585), pixelFreeTests (:587-611), viewBlockTest (:613-644).
// ----------------------------------------------------------------------------------------------------------
constexpr float MAX_DEPTH = 1000.0f;      // Scene.cpp:33

__global__ void depth_fill_kernel(float* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = MAX_DEPTH;
}

// Camera::mult (Camera.h:76-78)
__device__ __forceinline__ f3 mult_pt(const DevCamera& cam, f4 X, int level) {
    const float* p = cam.P[level];
    return f3{(p[0] * X.x + p[1] * X.y) + (p[2] * X.z + p[3] * X.w), (p[4] * X.x + p[5] * X.y) + (p[6] * X.z + p[7] * X.w),
              (p[8] * X.x + p[9] * X.y) + (p[10] * X.z + p[11] * X.w)};
}
// (int)(a / b + 0.5): f32 quotient, f64 addition, truncation
__device__ __forceinline__ int round_px(float a, float b) { return (int)((double)(a / b) + 0.5); }
// integer pixel / DEPTH_SUBSAMPLE (a double 2): f64 quotient, truncation
__device__ __forceinline__ int sub2(int v) { return (int)((double)v / 2.0); }

// setDepths(patch, subtract) for every (patch, view) pair; the reference's "d < old -> old = d" is an atomic float min, its
// "old == d -> MAX_DEPTH" (subtract, Scene.cpp:372-373) an atomic compare-and-swap on the same bits
__global__ void depth_set_kernel(const KParams K, const hpmvs_patch_t* __restrict__ patches, int n, int subtract) {
    const int total = n * MAXV;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const hpmvs_patch_t& p = patches[t / MAXV];
        const int k = t % MAXV;
        if (p.status != HPMVS_OK || k >= p.nimages) continue;
        const int idx = p.images[k];
        const DevCamera& cam = K.cams[idx];
        const f4 c = ld4(p.center);
        const f4 dd = sub4(c, ld4(cam.center));
        const float fz = sqrtf(dot4(dd, dd));
        const int level = leveli_from(cam, fz, p.scale, cam.nlevels - 1);
        const f3 imgC = mult_pt(cam, c, level);
        const int x = sub2(round_px(imgC.x, imgC.z)), y = sub2(round_px(imgC.y, imgC.z));
        const float d = imgC.z;
        if (!(d >= 0.0f)) continue;
        if (x < 0 || x >= cam.dcols[level] || y < 0 || y >= cam.drows[level]) continue;
        int* cell = reinterpret_cast<int*>(cam.depth[level] + (size_t)y * cam.dcols[level] + x);
        if (subtract) atomicCAS(cell, __float_as_int(d), __float_as_int(MAX_DEPTH));
        else atomicMin(cell, __float_as_int(d));
    }
}

__device__ __forceinline__ float full_depth(const DevCamera& cam, int xx, int yy) {
    float depth = MAX_DEPTH;
    int x = sub2(xx), y = sub2(yy);
    for (int level = 0; level < cam.nlevels; level++) {
        if (x < 0 || x >= cam.dcols[level] || y < 0 || y >= cam.drows[level]) return depth;
        depth = fminf(depth, cam.depth[level][(size_t)y * cam.dcols[level] + x]);
        x /= 2; y /= 2;
    }
    return depth;
}

// Scene::depthTest with neighbours = true (Scene.cpp:534-585); `abs(diff)` is float std::abs(float) with Eigen's include chain (see oracle/shim/Eigen/Dense)
__device__ __forceinline__ bool depth_test(const DevCamera& cam, f4 c, f4 nrm, float scale, float margin, bool viewBlock) {
    const f3 imgC = mult_pt(cam, c, 0);
    const int ix0 = round_px(imgC.x, imgC.z) - 1, iy0 = round_px(imgC.y, imgC.z) - 1;
    const float depth = imgC.z;
    const f4 ray = normalized4(sub4(c, ld4(cam.center)));
    const float factor = fminf(2.0f, 2.0f + dot4(ray, nrm));
    const double thr = (double)(scale * margin * factor) * 2.0;
    for (int yy = 0; yy < 3; yy++)
        for (int xx = 0; xx < 3; xx++) {
            const int ix = ix0 + xx, iy = iy0 + yy;
            if (depth < 0.0f || ix < 0 || ix >= cam.w[0] || iy < 0 || iy >= cam.h[0]) return false;
            const float imgDepth = full_depth(cam, ix, iy);
            if (imgDepth >= MAX_DEPTH) { if (viewBlock) return false; else continue; }
            const float diff = imgDepth - depth;
            if (!viewBlock) { if (!((double)fabsf(diff) < thr)) return false; }
            else { if (!((double)diff > thr)) return false; }
        }
    return true;
}

// one warp per patch: out[3*i] = depthTests, [3*i+1] = viewBlockTest, [3*i+2] = pixelFreeTests
__global__ void accept_kernel(const KParams K, const hpmvs_patch_t* __restrict__ patches, int n, float margin, int* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int pi = warp; pi < n; pi += nwarps) {
        const hpmvs_patch_t& p = patches[pi];
        const f4 c = ld4(p.center), nrm = ld4(p.normal);
        const float scale = p.scale;
        const int nimg = min(max(p.nimages, 0), MAXV);
        bool vis = false, fre = false;
        if (lane < nimg) {
            const DevCamera& cam = K.cams[p.images[lane]];
            vis = depth_test(cam, c, nrm, scale, margin, false);
            // pixelFreeTest (Scene.cpp:595-611)
            const f4 dd = sub4(c, ld4(cam.center));
            const int level = (int)roundf(level_from(cam, sqrtf(dot4(dd, dd)), scale));
            if (level >= 0 && level < cam.nlevels) {
                float u, v;
                project(cam, c, level, u, v);
                // project() already divided: imgC[2] is 1 (or -1 behind the camera)
                float w = 1.0f;
                const float* pr = cam.P[level];
                if ((pr[8] * c.x + pr[9] * c.y) + (pr[10] * c.z + pr[11] * c.w) <= 0.0f) w = -1.0f;
                const int ix = round_px(u, w), iy = round_px(v, w);
                if (ix >= 0 && ix < cam.w[level] && iy >= 0 && iy < cam.h[level]) {
                    const int x = sub2(ix), y = sub2(iy);
                    float dm = MAX_DEPTH;
                    if (x >= 0 && x < cam.dcols[level] && y >= 0 && y < cam.drows[level]) dm = cam.depth[level][(size_t)y * cam.dcols[level] + x];
                    fre = (dm == MAX_DEPTH);
                }
            }
        }
        const int nvis = __popc(__ballot_sync(FULL, vis)), nfree = __popc(__ballot_sync(FULL, fre));
        int nblock = 0;
        for (int base = 0; base < K.ncams; base += 32) {
            const int img = base + lane;
            bool blk = false;
            if (img < K.ncams) {
                const DevCamera& cam = K.cams[img];
                const f4 dd = sub4(c, ld4(cam.center));
                const int level = (int)roundf(level_from(cam, sqrtf(dot4(dd, dd)), scale));
                if (level >= 0 && level <= cam.nlevels - 1) {
                    float u, v;
                    project(cam, c, level, u, v);
                    if (!(u < 0.0f || u > (float)cam.w[level] || v < 0.0f || v > (float)cam.h[level]))
                        blk = depth_test(cam, c, nrm, scale, margin, true);
                }
            }
            nblock += __popc(__ballot_sync(FULL, blk));
        }
        if (lane == 0) { out[3 * pi] = nvis; out[3 * pi + 1] = nblock; out[3 * pi + 2] = nfree; }
    }
}

// ----------------------------------------------------------------------------------------------------------
// Candidate construction of CellProcessor::extend (mode 6, CellProcessor.cpp:98-119) and ::branch (mode 4, :227-249) on
// device-resident parents: one thread per (parent, direction).  The direction cosines are formed on the HOST (cos / sin of six or four
// constant angles with the caller's libm, exactly the values the host function hpmvs_expand_candidates uses) and passed in.
// ----------------------------------------------------------------------------------------------------------
struct ExpandDirs { float dx[6], dy[6]; };
__global__ void expand_candidates_kernel(const DevCamera* __restrict__ cams, int ncams, const hpmvs_patch_t* __restrict__ parents,
                                         const float* __restrict__ widths, int n, int mode, ExpandDirs D, hpmvs_patch_t* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * mode) return;
    const int i = t / mode, ii = t - mode * i;
    hpmvs_patch_t q = parents[i];
    const int ref = (q.nimages > 0 && q.images[0] >= 0 && q.images[0] < ncams) ? q.images[0] : 0;
    const DevCamera& rc = cams[ref];
    const f3 nrm = f3{q.normal[0], q.normal[1], q.normal[2]};
    const f3 ya = normalized3(cross3(nrm, f3{rc.xaxis[0], rc.xaxis[1], rc.xaxis[2]}));
    const f3 xa = cross3(ya, nrm);
    const float width = widths[i];
    const float extend = (mode == 6) ? width : (float)((double)width / 4.0);
    const float dx = D.dx[ii], dy = D.dy[ii];
    q.center[0] = q.center[0] + (dx * xa.x + dy * ya.x) * extend;
    q.center[1] = q.center[1] + (dx * xa.y + dy * ya.y) * extend;
    q.center[2] = q.center[2] + (dx * xa.z + dy * ya.z) * extend;
    q.scale = (mode == 6) ? (float)((double)width * 0.9 / 2.0) : (float)((double)width * 0.45 / 2.0);
    out[t] = q;
}

// ----------------------------------------------------------------------------------------------------------
// Border de-duplication after the final multi-GPU gather (device form of hpmvs_dedup_border, host_pipeline.cpp): patches of DIFFERENT
// ranks that fall into the same cubic cell are reduced to the best supported one - most views (CellProcessor::filter's spirit,
// CellProcessor.cpp:43-82), then the lower final score, then the lower rank, then the lower index; patches of the winner's own rank
// in that cell all stay.  Open-addressing hash table on the packed cell key, then three atomic reduction passes and a verdict pass.
// ----------------------------------------------------------------------------------------------------------
struct DedupTable {
    unsigned long long* keys;        // [cap] packed cell key + 1 (0 = empty)
    int* best_nimg;                  // [cap]
    unsigned long long* best_score;  // [cap] order-preserving bits of the score
    unsigned long long* best_who;    // [cap] (owner << 32) | index
    int* slot_of;                    // [n]
    unsigned cap_mask;
};
__device__ __forceinline__ unsigned long long dedup_key(const float* c, double ox, double oy, double oz, double cell) {
    const long long off = 1ll << 20, hi = (1ll << 21) - 1;
    long long k0 = (long long)floor(((double)c[0] - ox) / cell) + off, k1 = (long long)floor(((double)c[1] - oy) / cell) + off,
              k2 = (long long)floor(((double)c[2] - oz) / cell) + off;
    k0 = k0 < 0 ? 0 : (k0 > hi ? hi : k0); k1 = k1 < 0 ? 0 : (k1 > hi ? hi : k1); k2 = k2 < 0 ? 0 : (k2 > hi ? hi : k2);
    return ((unsigned long long)k0 << 42) | ((unsigned long long)k1 << 21) | (unsigned long long)k2;
}
__device__ __forceinline__ unsigned long long ordered_bits(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void dedup_insert_kernel(const hpmvs_patch_t* __restrict__ rec, int n, double ox, double oy, double oz, double cell, DedupTable T) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int slot = -1;
        if (rec[i].status == HPMVS_OK) {
            const unsigned long long key = dedup_key(rec[i].center, ox, oy, oz, cell) + 1ull;
            unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 32) & T.cap_mask;
            for (;;) {
                const unsigned long long prev = atomicCAS(&T.keys[h], 0ull, key);
                if (prev == 0ull || prev == key) { slot = (int)h; break; }
                h = (h + 1) & T.cap_mask;
            }
            atomicMax(&T.best_nimg[slot], rec[i].nimages);
        }
        T.slot_of[i] = slot;
    }
}
__global__ void dedup_score_kernel(const hpmvs_patch_t* __restrict__ rec, int n, DedupTable T) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int slot = T.slot_of[i];
        if (slot >= 0 && rec[i].nimages == T.best_nimg[slot]) atomicMin(&T.best_score[slot], ordered_bits(rec[i].score));
    }
}
__global__ void dedup_who_kernel(const hpmvs_patch_t* __restrict__ rec, const int* __restrict__ owner, int n, DedupTable T) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int slot = T.slot_of[i];
        if (slot >= 0 && rec[i].nimages == T.best_nimg[slot] && ordered_bits(rec[i].score) == T.best_score[slot])
            atomicMin(&T.best_who[slot], ((unsigned long long)(unsigned)owner[i] << 32) | (unsigned)i);
    }
}
__global__ void dedup_verdict_kernel(const int* __restrict__ owner, int n, DedupTable T, unsigned char* __restrict__ keep, int* __restrict__ nkeep) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int slot = T.slot_of[i];
        unsigned char k = 0;
        if (slot >= 0) {
            const unsigned long long w = T.best_who[slot];
            k = ((unsigned)(w & 0xffffffffull) == (unsigned)i || (int)(w >> 32) == owner[i]) ? 1 : 0;
        }
        keep[i] = k;
        if (k) atomicAdd(nkeep, 1);
    }
}

// ----------------------------------------------------------------------------------------------------------
// image layout kernels
// ----------------------------------------------------------------------------------------------------------
// interleaved RGB (tightly packed staging copy) -> pitched RGBX
__global__ void rgb_to_rgbx_kernel(const unsigned char* __restrict__ src, int w, int h, uchar4* __restrict__ dst, int pitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const unsigned char* s = src + 3 * ((size_t)y * w + x);
    dst[(size_t)y * pitch + x] = make_uchar4(s[0], s[1], s[2], 255);
}
__global__ void rgbx_to_rgb_kernel(const uchar4* __restrict__ src, int pitch, int w, int h, unsigned char* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uchar4 p = src[(size_t)y * pitch + x];
    unsigned char* d = dst + 3 * ((size_t)y * w + x);
    d[0] = p.x; d[1] = p.y; d[2] = p.z;
}

// Image::undistort (src/hpmvs/Image.cpp:68-149) on the device: one thread per TARGET pixel finds its source position in the distorted
// level-0 image in closed form (undistort_math.h), samples it with CImg's _linear_atXY (thirdLibs/cimg/CImg.h:12218-12235, f32) when
// it lies strictly inside the 1-pixel border, and truncates to u8; other target pixels are 0 (the reference leaves them
// uninitialised, Q13).  `src` is the tightly packed interleaved RGB staging copy, `dst` the pitched RGBX level 0.
__global__ void undistort_kernel(const unsigned char* __restrict__ src, int w, int h, float f, float k1, uchar4* __restrict__ dst, int pitch) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix >= w || iy >= h) return;
    float y = (float)((double)iy - (double)h / 2.0);
    float x = (float)((double)ix - (double)w / 2.0);
    x /= f;
    y /= f;
    if (y == 0.0f) y = (float)1e-3;
    const double kr = (double)k1 * ((double)(y * y) + (double)(x * x));
    const ud::Source s = (k1 > 0.0f) ? ud::source_positive_k1(x, y, kr) : ud::source_negative_k1_polar(x, y, kr);
    x = s.mx * f + (float)w / 2.0f;
    y = s.my * f + (float)h / 2.0f;
    uchar4 o = make_uchar4(0, 0, 0, 255);
    if (x > 1.0f && x < (float)(w - 1) && y > 1.0f && y < (float)(h - 1)) {
        const unsigned px = (unsigned)x, py = (unsigned)y;
        const float dx = x - (float)px, dy = y - (float)py;
        const unsigned nx = dx > 0.0f ? px + 1 : px, ny = dy > 0.0f ? py + 1 : py;
        const unsigned char* a = src + 3 * ((size_t)py * w + px);
        const unsigned char* b = src + 3 * ((size_t)py * w + nx);
        const unsigned char* c = src + 3 * ((size_t)ny * w + px);
        const unsigned char* d = src + 3 * ((size_t)ny * w + nx);
        unsigned char r[3];
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float Icc = (float)a[ch], Inc = (float)b[ch], Icn = (float)c[ch], Inn = (float)d[ch];
            const float v = Icc + dx * (Inc - Icc + dy * (Icc + Inn - Icn - Inc)) + dy * (Icn - Icc);
            r[ch] = (unsigned char)v;
        }
        o = make_uchar4(r[0], r[1], r[2], 255);
    }
    dst[(size_t)iy * pitch + ix] = o;
}

// CImg::get_resize_halfXY (thirdLibs/cimg/CImg.h:21189-21203): 3x3 mask at odd (x,y), Neumann border,
// f32 left-to-right accumulation, truncation to u8.  One thread per output pixel, all three channels.
__global__ void half_xy_kernel(const uchar4* __restrict__ src, int sp, int W, int H, uchar4* __restrict__ dst, int dp, int w2, int h2) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ox >= w2 || oy >= h2) return;
    const int x = 2 * ox + 1, y = 2 * oy + 1;
    const int xp = x - 1, xn = (x + 1 >= W) ? W - 1 : x + 1;
    const int yp = y - 1, yn = (y + 1 >= H) ? H - 1 : y + 1;
    const float m0 = 0.07842776544f, m1 = 0.1231940459f, m4 = 0.1935127547f;
    const uchar4 a0 = src[(size_t)yp * sp + xp], a1 = src[(size_t)yp * sp + x], a2 = src[(size_t)yp * sp + xn];
    const uchar4 b0 = src[(size_t)y * sp + xp], b1 = src[(size_t)y * sp + x], b2 = src[(size_t)y * sp + xn];
    const uchar4 c0 = src[(size_t)yn * sp + xp], c1 = src[(size_t)yn * sp + x], c2 = src[(size_t)yn * sp + xn];
#define HP_TAP(ch) ((float)a0.ch * m0 + (float)a1.ch * m1 + (float)a2.ch * m0 + (float)b0.ch * m1 + (float)b1.ch * m4 + \
                    (float)b2.ch * m1 + (float)c0.ch * m0 + (float)c1.ch * m1 + (float)c2.ch * m0)
    const float r = HP_TAP(x), g = HP_TAP(y), b = HP_TAP(z);
#undef HP_TAP
    dst[(size_t)oy * dp + ox] = make_uchar4((unsigned char)(int)r, (unsigned char)(int)g, (unsigned char)(int)b, 255);
}

}  // namespace hp
