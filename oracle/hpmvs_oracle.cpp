// TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT.  See hpmvs_oracle.h for scope and parity status.
//
// Every function cites the reference lines it restates (paths relative to /root/reference).
// Arithmetic conventions (the reference's third-party arithmetic is Eigen3, which is absent
// from /root/reference and unpinned by it - CMakeLists.txt:12):
//   * Eigen >= 3.3 semantics, x86-64 baseline (SSE2 on, no FMA, no AVX: CMakeLists.txt:4-7 sets
//     only -std=c++11 -O3):
//       - Vector4f reductions (dot/squaredNorm) are SSE2-vectorised: (a0+a2)+(a1+a3)
//       - Vector3f / Vector2f (fixed size, no packet access) reductions are unrolled binary trees: a0+(a1+a2)
//       - DYNAMIC-size blocks (n.head(3), row(2).head(3): Block<.,Dynamic,1>) are not unrolled and are below one
//         packet: Redux.h takes its scalar loop (a0+a1)+a2 - used at PatchOptimizer.cpp:536 and Camera.cpp:71
//       - fixed-size matrix*vector is coefficient based: inner 4 -> (p0+p1)+(p2+p3), inner 3 -> p0+(p1+p2)
//       - normalized(): z=squaredNorm; if (z>0) v / sqrt(z)  (true division per element)
//       - double scalars multiplying float vectors are narrowed to float first
//   * unqualified sin/cos on float arguments resolve to the double overloads (only <cmath> is
//     included, PatchOptimizer.cpp:21-33), std::cos/std::acos/std::asin/std::round on floats to
//     the float overloads.
//   * compiled with -ffp-contract=off (see Makefile).
#include "hpmvs_oracle.h"
#include "oracle_nlopt_decl.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// std::asin(float) at PatchOptimizer.cpp:427 is the only libm call on the path whose result feeds the optimiser
// directly.  glibc 2.39's asinf is not correctly rounded (differs from RN(asin(x)) for ~3.8% of inputs), newer
// glibc (CORE-MATH) is.  g_cr_asinf = 1 evaluates it as (float)asin((double)x) = correctly rounded, which is what
// the GPU engine does; 0 = this box's libm.  Tests run both and state the difference.
int g_cr_asinf = 0;

// ------------------------------------------------------------------------------------------
// small fixed-size vector helpers with Eigen's evaluation order
// ------------------------------------------------------------------------------------------
struct V4 { float v[4]; float& operator[](int i) { return v[i]; } float operator[](int i) const { return v[i]; } };
struct V3 { float v[3]; float& operator[](int i) { return v[i]; } float operator[](int i) const { return v[i]; } };
struct V2 { float v[2]; float& operator[](int i) { return v[i]; } float operator[](int i) const { return v[i]; } };

inline V4 sub4(const V4& a, const V4& b) { return V4{{a[0] - b[0], a[1] - b[1], a[2] - b[2], a[3] - b[3]}}; }
inline V4 add4(const V4& a, const V4& b) { return V4{{a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]}}; }
inline float dot4(const V4& a, const V4& b) {
    const float p0 = a[0] * b[0], p1 = a[1] * b[1], p2 = a[2] * b[2], p3 = a[3] * b[3];
    return (p0 + p2) + (p1 + p3);
}
inline float norm4(const V4& a) { return std::sqrt(dot4(a, a)); }
inline V4 normalized4(const V4& a) {
    const float z = dot4(a, a);
    if (z > 0.0f) { const float s = std::sqrt(z); return V4{{a[0] / s, a[1] / s, a[2] / s, a[3] / s}}; }
    return a;
}
inline float dot3(const V3& a, const V3& b) {
    const float p0 = a[0] * b[0], p1 = a[1] * b[1], p2 = a[2] * b[2];
    return p0 + (p1 + p2);
}
inline float norm3(const V3& a) { return std::sqrt(dot3(a, a)); }
inline V3 normalized3(const V3& a) {
    const float z = dot3(a, a);
    if (z > 0.0f) { const float s = std::sqrt(z); return V3{{a[0] / s, a[1] / s, a[2] / s}}; }
    return a;
}
inline V3 cross3(const V3& a, const V3& b) {
    return V3{{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}};
}
inline V3 head3(const V4& a) { return V3{{a[0], a[1], a[2]}}; }
// reductions over a dynamic-size head(3) block: scalar loop, (a0+a1)+a2
inline float sqnorm3_dyn(const V3& a) { return (a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]; }
inline V3 normalized3_dyn(const V3& a) {
    const float z = sqnorm3_dyn(a);
    if (z > 0.0f) { const float s = std::sqrt(z); return V3{{a[0] / s, a[1] / s, a[2] / s}}; }
    return a;
}

// ------------------------------------------------------------------------------------------
// Scene data
// ------------------------------------------------------------------------------------------
struct Camera {
    float P[ORC_LEVELS][3][4];
    float K0[3][3];
    V4 center;
    V3 xAxis, yAxis, zAxis;
    int nlevels;
};

struct Image {
    int w[ORC_LEVELS], h[ORC_LEVELS];
    std::vector<uint8_t> lvl[ORC_LEVELS];  // interleaved RGB, row stride 3*w (Image.cpp:62-63)
};

struct DepthMap { int rows = 0, cols = 0; std::vector<float> d; };   // Scene::m_depths[cam][level] (Scene.h:75-76)

struct Scene {
    orc_options_t opt;
    std::vector<Camera> cameras;
    std::vector<Image> images;
    std::vector<std::vector<int>> covis;
    std::vector<std::vector<DepthMap>> depths;
};

const float MAX_DEPTH = 1000.0f;        // Scene.cpp:33
const double DEPTH_SUBSAMPLE = 2;       // Scene.h:78 (a double: integer pixel coordinates are divided in f64, then truncated)

// Camera::init, src/hpmvs/Camera.cpp:34-81
void camera_init(Camera& cam, double f, const double q[4], const double c[3], int width, int height, int maxLevel) {
    cam.nlevels = maxLevel + 1;
    float K[3][3] = {{(float)f, 0.0f, (float)(width / 2.0)}, {0.0f, (float)f, (float)(height / 2.0)}, {0.0f, 0.0f, 1.0f}};
    std::memcpy(cam.K0, K, sizeof(K));
    // Eigen::Quaterniond::toRotationMatrix (double), then cast<float>()  (Camera.cpp:43-50)
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    double Rd[3][3];
    Rd[0][0] = 1.0 - (tyy + tzz); Rd[0][1] = txy - twz;         Rd[0][2] = txz + twy;
    Rd[1][0] = txy + twz;         Rd[1][1] = 1.0 - (txx + tzz); Rd[1][2] = tyz - twx;
    Rd[2][0] = txz - twy;         Rd[2][1] = tyz + twx;         Rd[2][2] = 1.0 - (txx + tyy);
    float R[3][3], cf[3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[i][j] = (float)Rd[i][j];
    for (int i = 0; i < 3; i++) cf[i] = (float)c[i];
    float M[3][4];
    for (int i = 0; i < 3; i++) {
        // col(3) = (-R) * c, coefficient based, inner size 3
        const float p0 = (-R[i][0]) * cf[0], p1 = (-R[i][1]) * cf[1], p2 = (-R[i][2]) * cf[2];
        M[i][3] = p0 + (p1 + p2);
        for (int j = 0; j < 3; j++) M[i][j] = R[i][j];
    }
    // projection_[0] = kMat_[0] * projection_[0]   (3x3 * 3x4, inner size 3)
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) {
            const float p0 = K[i][0] * M[0][j], p1 = K[i][1] * M[1][j], p2 = K[i][2] * M[2][j];
            cam.P[0][i][j] = p0 + (p1 + p2);
        }
    for (int l = 1; l < cam.nlevels; l++)
        for (int j = 0; j < 4; j++) {
            cam.P[l][0][j] = cam.P[l - 1][0][j] / 2.0f;
            cam.P[l][1][j] = cam.P[l - 1][1][j] / 2.0f;
            cam.P[l][2][j] = cam.P[l - 1][2][j];
        }
    cam.center = V4{{cf[0], cf[1], cf[2], 1.0f}};
    // oAxis_ = row(2) / row(2).head(3).norm()   (Camera.cpp:66-67)
    const V3 r2{{cam.P[0][2][0], cam.P[0][2][1], cam.P[0][2][2]}};
    const float n2 = std::sqrt(sqnorm3_dyn(r2));   // row(2).head(3).norm(): dynamic size
    cam.zAxis = V3{{r2[0] / n2, r2[1] / n2, r2[2] / n2}};
    const V3 r0{{cam.P[0][0][0], cam.P[0][0][1], cam.P[0][0][2]}};
    cam.yAxis = normalized3(cross3(cam.zAxis, r0));
    cam.xAxis = normalized3(cross3(cam.yAxis, cam.zAxis));
}

// Camera::project, include/hpmvs/Camera.h:45-62
inline V3 project(const Camera& cam, const V4& X, int level) {
    V3 r;
    for (int i = 0; i < 3; i++) {
        const float* p = cam.P[level][i];
        r[i] = (p[0] * X[0] + p[1] * X[1]) + (p[2] * X[2] + p[3] * X[3]);
    }
    if (r[2] <= 0.0f) {
        r = V3{{-(float)0xffff, -(float)0xffff, -1.0f}};
    } else {
        const float z = r[2];
        r[0] = r[0] / z; r[1] = r[1] / z; r[2] = r[2] / z;
        const float lo = (float)(INT_MIN + 3.0f), hi = (float)(INT_MAX - 3.0f);
        r[0] = std::max(lo, std::min(hi, r[0]));
        r[1] = std::max(lo, std::min(hi, r[1]));
    }
    return r;
}

// Camera::getScale, Camera.cpp:83-90
inline float get_scale(const Camera& cam, const V4& coord, int level) {
    const float fz = norm4(sub4(coord, cam.center));
    const float ftmp = cam.K0[0][0] + cam.K0[1][1];
    if (ftmp == 0.0) return 1.0;
    return (float)(2.0 * fz * (0x0001 << level) / ftmp);
}
// Camera::getLevel, Camera.cpp:92-95
inline float get_level(const Camera& cam, const V4& coord, float scale) {
    const float fz = norm4(sub4(coord, cam.center));
    return (float)std::log2(scale * (float)(cam.K0[0][0] + cam.K0[1][1]) / (2.0 * fz));
}
// Camera::getLeveli, Camera.cpp:97-99
inline int get_leveli(const Camera& cam, const V4& coord, float scale, int maxLevel) {
    return std::max(0, std::min(maxLevel, (int)std::round(get_level(cam, coord, scale))));
}

// Image::getColor, include/hpmvs/Image.h:89-115
inline V3 get_color(const Image& img, float x, float y, int level) {
    const int W = img.w[level];
    const int lx = static_cast<int>(x);
    const int ly = static_cast<int>(y);
    const int index = 3 * (ly * W + lx);
    const float dx1 = x - lx; const float dx0 = 1.0f - dx1;
    const float dy1 = y - ly; const float dy0 = 1.0f - dy1;
    const float f00 = dx0 * dy0; const float f01 = dx0 * dy1;
    const float f10 = dx1 * dy0; const float f11 = dx1 * dy1;
    const int index2 = index + 3 * W;
    const uint8_t* p0 = img.lvl[level].data() + index;
    const uint8_t* p1 = img.lvl[level].data() + index2;
    float r = 0.0f, g = 0.0f, b = 0.0f;
    r += p0[0] * f00 + p1[0] * f01;
    g += p0[1] * f00 + p1[1] * f01;
    b += p0[2] * f00 + p1[2] * f01;
    r += p0[3] * f10 + p1[3] * f11;
    g += p0[4] * f10 + p1[4] * f11;
    b += p0[5] * f10 + p1[5] * f11;
    return V3{{r, g, b}};
}

// CImg::get_resize_halfXY on one planar channel, thirdLibs/cimg/CImg.h:21189-21203
// (3x3 mask sampled at odd (x,y), Neumann border from cimg_for3x3, float -> u8 truncation)
void half_xy_interleaved(const std::vector<uint8_t>& src, int W, int H, std::vector<uint8_t>& dst, int& w2, int& h2) {
    static const float mask[9] = {0.07842776544f, 0.1231940459f, 0.07842776544f, 0.1231940459f, 0.1935127547f,
                                  0.1231940459f, 0.07842776544f, 0.1231940459f, 0.07842776544f};
    w2 = W / 2; h2 = H / 2;
    dst.assign((size_t)w2 * h2 * 3, 0);
    for (int c = 0; c < 3; c++)
        for (int y = 1, oy = 0; y < H && oy < h2; y += 2, oy++) {
            const int yp = y - 1, yn = (y + 1 >= H) ? H - 1 : y + 1;
            for (int x = 1, ox = 0; x < W && ox < w2; x += 2, ox++) {
                const int xp = x - 1, xn = (x + 1 >= W) ? W - 1 : x + 1;
                auto at = [&](int xx, int yy) -> float { return (float)src[3 * ((size_t)yy * W + xx) + c]; };
                const float s = at(xp, yp) * mask[0] + at(x, yp) * mask[1] + at(xn, yp) * mask[2] + at(xp, y) * mask[3] +
                                at(x, y) * mask[4] + at(xn, y) * mask[5] + at(xp, yn) * mask[6] + at(x, yn) * mask[7] +
                                at(xn, yn) * mask[8];
                dst[3 * ((size_t)oy * w2 + ox) + c] = (uint8_t)s;
            }
        }
}

// ------------------------------------------------------------------------------------------
// PatchTex (Patch2d<7,float>), include/hpmvs/Patch2d.hpp
// ------------------------------------------------------------------------------------------
struct PatchTex {
    float data[7 * 7 * 3];
    // Patch2d.hpp:37-44
    float dot(const PatchTex& o) const {
        const int size = 7 * 7 * 3;
        float ans = 0.0f;
        for (int i = 0; i < size; ++i) ans += data[i] * o.data[i];
        return ans / size;
    }
    // Patch2d.hpp:46-84
    void normalize() {
        const int size = 7 * 7 * 3, size3 = 7 * 7;
        float ave[3] = {0.0f, 0.0f, 0.0f};
        for (int i = 0; i < size3; ++i) { ave[0] += data[3 * i]; ave[1] += data[3 * i + 1]; ave[2] += data[3 * i + 2]; }
        ave[0] /= (float)size3; ave[1] /= (float)size3; ave[2] /= (float)size3;
        float ave2 = 0.0f;
        for (int i = 0; i < size3; ++i) {
            const float f0 = ave[0] - data[3 * i], f1 = ave[1] - data[3 * i + 1], f2 = ave[2] - data[3 * i + 2];
            ave2 += f0 * f0 + f1 * f1 + f2 * f2;
        }
        ave2 = std::sqrt(ave2 / size);
        if (ave2 == 0.0f) ave2 = 1.0f;
        for (int i = 0; i < size3; ++i)
            for (int c = 0; c < 3; c++) { data[3 * i + c] -= ave[c]; data[3 * i + c] /= ave2; }
    }
};

// ------------------------------------------------------------------------------------------
// PatchOptimizer, src/hpmvs/PatchOptimizer.cpp
// ------------------------------------------------------------------------------------------
struct PatchOptimizer {
    const Scene* scene;
    const orc_options_t* opt;
    V4 pCenter, pNormal;
    float pScale;
    V4 pXaxis, pYaxis, pZaxis;
    std::vector<int> pImages;
    V4 refCenter, refRay;
    std::vector<V3> imgX, imgY, imgZ;
    float depthScale, angleScale;
    PatchTex refTex, comTex;
    // instrumentation (not in the reference)
    int evals = 0, textures = 0, nlopt_result = 0, fail_stage = 0;
    double last_val = 0.0;

    explicit PatchOptimizer(const Scene* s) : scene(s), opt(&s->opt) {}

    static inline float robustincc(const float rhs) { return rhs / (1 + 3 * rhs); }  // PatchOptimizer.h:92-94

    // :532-548
    void calculatePatchAxis(int refIndex, const V4& n, float scale) {
        const Camera& rc = scene->cameras[refIndex];
        V3 z = normalized3_dyn(head3(n));   // n.head(3).normalized(): dynamic size
        V3 y = normalized3(cross3(z, rc.xAxis));
        V3 x = normalized3(cross3(y, z));
        for (int i = 0; i < 3; i++) { x[i] *= scale; y[i] *= scale; }
        const float s = dot3(normalized3(y), normalized3(rc.yAxis));
        for (int i = 0; i < 3; i++) y[i] = y[i] * s;
        pXaxis = V4{{x[0], x[1], x[2], 0.0f}};
        pYaxis = V4{{y[0], y[1], y[2], 0.0f}};
        pZaxis = V4{{z[0], z[1], z[2], 0.0f}};
    }

    // :476-529
    bool sampleTexture(const V4& c4, float scale, const V4& xax, const V4& yax, const V4& zax, int camIdx, PatchTex& tex) {
        const Image& image = scene->images[camIdx];
        const Camera& camera = scene->cameras[camIdx];
        if (dot4(normalized4(sub4(camera.center, c4)), zax) < ::cos((double)opt->max_angle)) return false;  // unqualified cos(float) -> double overload
        const int lvl = get_leveli(camera, c4, scale, opt->maxlevel - 1);
        const V3 pc = project(camera, c4, lvl);
        const V3 px = project(camera, add4(c4, xax), lvl);
        const V3 py = project(camera, add4(c4, yax), lvl);
        const V2 center{{pc[0], pc[1]}};
        const V2 dx{{px[0] - center[0], px[1] - center[1]}};
        const V2 dy{{py[0] - center[0], py[1] - center[1]}};
        const float hs = 7 / 2.0f;
        V2 tl, tr, bl, br;
        for (int i = 0; i < 2; i++) {
            tl[i] = center[i] - hs * dx[i] - hs * dy[i];
            tr[i] = center[i] + hs * dx[i] - hs * dy[i];
            bl[i] = center[i] - hs * dx[i] + hs * dy[i];
            br[i] = center[i] + hs * dx[i] + hs * dy[i];
        }
        V2 mn, mx;
        for (int i = 0; i < 2; i++) {
            mn[i] = std::min(std::min(std::min(tl[i], tr[i]), bl[i]), br[i]);
            mx[i] = std::max(std::max(std::max(tl[i], tr[i]), bl[i]), br[i]);
        }
        const int m = 3;
        if (mn[0] < m || mn[1] < m || mx[0] >= image.w[lvl] - m || mx[1] >= image.h[lvl] - m) return false;
        textures++;
        float* target = tex.data;
        V2 l = tl;
        for (int yy = 0; yy < 7; yy++) {
            V2 c = l;
            l[0] += dy[0]; l[1] += dy[1];
            for (int xx = 0; xx < 7; xx++) {
                const V3 col = get_color(image, c[0], c[1], lvl);
                *(target++) = col[0]; *(target++) = col[1]; *(target++) = col[2];
                c[0] += dx[0]; c[1] += dx[1];
            }
        }
        tex.normalize();
        return true;
    }

    // :448-474
    void setINCCs(std::vector<float>& inccs, const std::vector<int>& indexes, int refIdx, int robust) {
        inccs.resize(indexes.size());
        calculatePatchAxis(indexes[refIdx], pNormal, pScale);
        if (!sampleTexture(pCenter, pScale, pXaxis, pYaxis, pNormal, indexes[refIdx], refTex)) {
            std::fill(inccs.begin(), inccs.end(), 2.0f);
            return;
        }
        for (int ii = 0; ii < (int)indexes.size(); ii++) {
            if (ii == refIdx) inccs[ii] = 0.0f;
            else if (!sampleTexture(pCenter, pScale, pXaxis, pYaxis, pNormal, indexes[ii], comTex)) inccs[ii] = 2.0f;
            else if (robust) inccs[ii] = robustincc(1.0f - refTex.dot(comTex));
            else inccs[ii] = 1.0f - refTex.dot(comTex);
        }
    }

    // :225-258
    bool addImages() {
        if (pImages.size() <= 0) return false;
        const int refImg = pImages[0];
        std::set<int> existing(pImages.begin(), pImages.end());
        for (const int covisImg : scene->covis[refImg]) {
            if (existing.find(covisImg) != existing.end()) continue;
            const Camera& cam = scene->cameras[covisImg];
            if (dot4(normalized4(sub4(cam.center, pCenter)), pNormal) < std::cos(opt->max_angle)) continue;
            int imgLevel = (int)std::round(get_level(cam, pCenter, pScale));
            if (imgLevel < opt->minlevel || imgLevel >= opt->maxlevel - 2) continue;
            const V3 imgC = project(cam, pCenter, imgLevel);
            const Image& im = scene->images[covisImg];
            if (imgC[0] < 0.0f || im.w[imgLevel] - 1 <= imgC[0] || imgC[1] < 0.0f || im.h[imgLevel] - 1 <= imgC[1]) continue;
            pImages.push_back(covisImg);
        }
        return (int)pImages.size() >= opt->min_images_per_patch;
    }

    // :138-152
    bool filterImagesNCC(const float threshold) {
        std::vector<float> inccs;
        setINCCs(inccs, pImages, 0, 0);
        std::vector<int> newimages;
        newimages.push_back(pImages[0]);
        for (int i = 1; i < (int)pImages.size(); ++i)
            if (inccs[i] < 1.0f - threshold) newimages.push_back(pImages[i]);
        pImages.swap(newimages);
        return (int)pImages.size() >= opt->min_images_per_patch;
    }

    // :260-284
    void getAngleWeightedScales(std::vector<int>& indexes, std::vector<float>& scales, std::vector<V4>& rays) {
        if (pImages.empty()) return;
        const int refLevel = std::max(0, std::min(opt->maxlevel - 1, (int)std::round(get_level(scene->cameras[pImages[0]], pCenter, pScale))));
        indexes.clear(); scales.clear(); rays.clear();
        for (const int imgIdx : pImages) {
            const Camera& cam = scene->cameras[imgIdx];
            const V4 ray = normalized4(sub4(cam.center, pCenter));
            const float cosa = dot4(ray, normalized4(pNormal));
            if (cosa > 0) {
                indexes.push_back(imgIdx);
                rays.push_back(ray);
                const float scale = get_scale(cam, pCenter, refLevel);
                scales.push_back(scale / cosa);
            }
        }
    }

    // :183-223
    bool sortImages() {
        const float threshold = 1.0f - ::cos(10.0 * M_PI / 180.0);
        std::vector<int> indexes, indexes2;
        std::vector<float> wScales, wScales2;
        std::vector<V4> rays, rays2;
        getAngleWeightedScales(indexes, wScales, rays);
        pImages.clear();
        if (indexes.size() < 2) return false;
        wScales[0] = 0.0f;
        while (!indexes.empty()) {
            const int index = (int)(std::min_element(wScales.begin(), wScales.end()) - wScales.begin());
            pImages.push_back(indexes[index]);
            indexes2.clear(); wScales2.clear(); rays2.clear();
            for (int j = 0; j < (int)rays.size(); ++j) {
                if (j == index) continue;
                indexes2.push_back(indexes[j]);
                rays2.push_back(rays[j]);
                const float ftmp = std::min(threshold, std::max(threshold / 2.0f, 1.0f - dot4(rays[index], rays[j])));
                wScales2.push_back(wScales[j] * (threshold / ftmp));
            }
            indexes2.swap(indexes); wScales2.swap(wScales); rays2.swap(rays);
        }
        return (int)pImages.size() >= opt->min_images_per_patch;
    }

    // :105-123
    bool assureImageAngles() {
        std::vector<V4> rays;
        for (const int img : pImages) rays.push_back(normalized4(sub4(scene->cameras[img].center, pCenter)));
        const int nrImgs = (int)pImages.size();
        for (int ii = 0; ii < nrImgs - 1; ii++)
            for (int jj = ii + 1; jj < nrImgs; jj++) {
                const float a = std::acos(dot4(rays[ii], rays[jj]));
                if (a < opt->max_angle && a > opt->min_angle) return true;
            }
        return false;
    }

    // :125-136
    bool filterImagesByAngle() {
        std::vector<int> newImages;
        for (const int imgId : pImages)
            if (dot4(normalized4(sub4(scene->cameras[imgId].center, pCenter)), pNormal) > std::cos(opt->max_angle))
                newImages.push_back(imgId);
        pImages.swap(newImages);
        return (int)pImages.size() >= opt->min_images_per_patch;
    }

    // :154-181
    void setRefImage() {
        if (pImages.size() <= 1) return;
        std::vector<float> incc;
        int refindex = -1;
        float refncc = std::numeric_limits<float>::max();
        for (int ii = 0; ii < (int)pImages.size(); ii++) {
            setINCCs(incc, pImages, ii, 1);
            float sum = 0.0f;
            for (float v : incc) sum = sum + v;
            if (sum < refncc) { refncc = sum; refindex = ii; }
        }
        const int refIndex = pImages[refindex];
        for (int i = 0; i < (int)pImages.size(); ++i)
            if (pImages[i] == refIndex) { const int t = pImages[0]; pImages[0] = refIndex; pImages[i] = t; break; }
    }

    // :384-399
    void setOptimizationFields() {
        imgX.clear(); imgY.clear(); imgZ.clear();
        for (int img : pImages) {
            imgX.push_back(normalized3(scene->cameras[img].xAxis));
            imgY.push_back(normalized3(scene->cameras[img].yAxis));
            imgZ.push_back(normalized3(scene->cameras[img].zAxis));
        }
        refCenter = pCenter;
        refRay = normalized4(sub4(refCenter, scene->cameras[pImages[0]].center));
        depthScale = 1.0;
        angleScale = M_PI / 48.0f;
    }

    // :401-414
    void setCenterNorm(const double* x) {
        const float x0 = (float)x[0];  // Eigen narrows the double scalar before scaling the Vector4f
        for (int i = 0; i < 4; i++) pCenter[i] = refCenter[i] + (x0 * refRay[i]) * depthScale;
        const float angle1 = x[1] * angleScale;
        const float angle2 = x[2] * angleScale;
        const float fx = ::sin((double)angle1) * ::cos((double)angle2);  // double overloads, narrowed on assignment
        const float fy = ::sin((double)angle2);
        const float fz = -::cos((double)angle1) * ::cos((double)angle2);
        for (int i = 0; i < 3; i++) pNormal[i] = (imgX[0][i] * fx + imgY[0][i] * fy) + imgZ[0][i] * fz;
        pNormal[3] = 0.0f;
    }

    // :416-446
    void parametersFromCenterNorm(const V4& c, const V4 n, const double* lb, const double* ub, double* x) {
        x[0] = dot4(sub4(c, refCenter), refRay) / depthScale;
        const V3 n3 = head3(n);
        const float fx = dot3(imgX[0], n3);
        const float fy = dot3(imgY[0], n3);
        const float fz = dot3(imgZ[0], n3);
        x[2] = g_cr_asinf ? (double)(float)std::asin((double)fy) : (double)std::asin(fy);
        const float cosb = std::cos(std::max(-1.0, std::min(1.0, x[2])));
        if (cosb == 0.0) x[1] = 0.0;
        else {
            const double sina = fx / cosb;
            const double cosa = -fz / cosb;
            x[1] = std::acos(std::min(1.0, std::max(-1.0, cosa)));
            if (sina < 0.0) x[1] = -x[1];
        }
        x[1] /= angleScale;
        x[2] /= angleScale;
        for (int i = 0; i < 3; i++) x[i] = std::min(ub[i], std::max(lb[i], x[i]));
    }

    // :286-311
    double objective_fn() {
        evals++;
        calculatePatchAxis(pImages[0], pNormal, pScale);
        if (!sampleTexture(pCenter, pScale, pXaxis, pYaxis, pZaxis, pImages[0], refTex)) return 2.0;
        double val = 0.0;
        int nImgs = 0;
        for (int ii = 1; ii < (int)pImages.size(); ii++) {
            if (!sampleTexture(pCenter, pScale, pXaxis, pYaxis, pZaxis, pImages[ii], comTex)) continue;
            val += robustincc(1.0 - refTex.dot(comTex));
            nImgs++;
        }
        if (nImgs < opt->min_images_per_patch - 1) return 2.0;
        return val / nImgs;
    }

    // :313-320
    static double static_objective_fn(unsigned, const double* x, double*, void* data) {
        PatchOptimizer* o = static_cast<PatchOptimizer*>(data);
        o->setCenterNorm(x);
        return o->objective_fn();
    }

    static int bobyqa_id() {
        static int id = -1;
        if (id < 0)
            for (int a = 0; a < 64; a++) {
                const char* nm = nlopt_algorithm_name(a);
                if (nm && std::strstr(nm, "BOBYQA")) { id = a; break; }
            }
        return id;
    }

    // :322-382 (nlopt::opt wrapper calls restated through the C API they forward to)
    bool optimizePatch() {
        if ((int)pImages.size() < opt->min_images_per_patch) { fail_stage = ORC_FAIL_OPT_MINIMAGES; return false; }
        const double min_angle = -23.99999, max_angle = 23.99999;
        double lb[3] = {-HUGE_VAL, min_angle, min_angle};
        double ub[3] = {HUGE_VAL, max_angle, max_angle};
        double x[3] = {0, 0, 0};
        nlopt_opt o = nlopt_create(bobyqa_id(), 3);
        nlopt_set_min_objective(o, static_objective_fn, this);
        nlopt_set_xtol_rel(o, 1.e-7);
        nlopt_set_maxeval(o, 1000);
        nlopt_set_lower_bounds(o, lb);
        nlopt_set_upper_bounds(o, ub);
        setOptimizationFields();
        parametersFromCenterNorm(refCenter, pNormal, lb, ub, x);
        double minf = 0;
        const int result = nlopt_optimize(o, x, &minf);
        nlopt_destroy(o);
        nlopt_result = result;
        last_val = minf;
        // negative codes throw in nlopt.hpp:138-147 and are caught at PatchOptimizer.cpp:369-372
        const bool success = (result == ORC_NLOPT_SUCCESS || result == ORC_NLOPT_STOPVAL_REACHED ||
                              result == ORC_NLOPT_FTOL_REACHED || result == ORC_NLOPT_XTOL_REACHED);
        if (!success) {
            fail_stage = (result == ORC_NLOPT_ROUNDOFF_LIMITED) ? ORC_FAIL_OPT_ROUNDOFF
                         : (result == ORC_NLOPT_MAXEVAL_REACHED) ? ORC_FAIL_OPT_MAXEVAL : ORC_FAIL_OPT_OTHER;
            return false;
        }
        setCenterNorm(x);
        return true;
    }

    // :48-76
    bool runOptimization() {
        if (!addImages()) { fail_stage = ORC_FAIL_ADD_IMAGES; return false; }
        if (!filterImagesNCC(opt->ncc_alpha_1)) { fail_stage = ORC_FAIL_NCC1; return false; }
        sortImages();
        if (!assureImageAngles()) { fail_stage = ORC_FAIL_ANGLES; return false; }
        if (!optimizePatch()) return false;
        if (!addImages()) { fail_stage = ORC_FAIL_ADD_IMAGES2; return false; }
        if (!filterImagesNCC(opt->ncc_alpha_2)) { fail_stage = ORC_FAIL_NCC2; return false; }
        if (!filterImagesByAngle()) { fail_stage = ORC_FAIL_ANGLE_FILTER; return false; }
        if (!assureImageAngles()) { fail_stage = ORC_FAIL_ANGLES2; return false; }
        setRefImage();
        if (!filterImagesNCC(opt->ncc_alpha_2)) { fail_stage = ORC_FAIL_NCC3; return false; }
        return true;
    }

    void load(const orc_patch_t& p) {
        for (int i = 0; i < 4; i++) { pCenter[i] = p.center[i]; pNormal[i] = p.normal[i]; }
        pScale = p.scale;
        pImages.assign(p.images, p.images + p.nimages);
    }
};

// Scene::getColor(const Patch3d&), src/hpmvs/Scene.cpp:300-327
V3 patch_color(const Scene& s, const V4& center, float scale, const int* images, int n) {
    std::vector<V3> colors;
    for (int k = 0; k < n; k++) {
        const Camera& cam = s.cameras[images[k]];
        const int lvl = get_leveli(cam, center, scale, cam.nlevels - 1);
        const V3 c = project(cam, center, lvl);
        colors.push_back(get_color(s.images[images[k]], c[0], c[1], lvl));
    }
    std::sort(colors.begin(), colors.end(), [](const V3& a, const V3& b) { return norm3(a) < norm3(b); });
    if (norm3(colors[colors.size() / 2]) > 250.0) return colors.front();
    return colors[colors.size() / 2];
}

// Camera::mult (Camera.h:76-78): projection without the perspective divide
inline V3 mult(const Camera& cam, const V4& X, int level) {
    V3 r;
    for (int i = 0; i < 3; i++) {
        const float* p = cam.P[level][i];
        r[i] = (p[0] * X[0] + p[1] * X[1]) + (p[2] * X[2] + p[3] * X[3]);
    }
    return r;
}

// Scene::addCameras depth-map allocation (Scene.cpp:74-81)
void alloc_depths(Scene& s) {
    s.depths.assign(s.cameras.size(), {});
    for (size_t c = 0; c < s.cameras.size(); c++) {
        s.depths[c].resize(s.cameras[c].nlevels);
        for (int l = 0; l < s.cameras[c].nlevels; l++) {
            DepthMap& m = s.depths[c][l];
            m.rows = (int)(s.images[c].h[l] / DEPTH_SUBSAMPLE);
            m.cols = (int)(s.images[c].w[l] / DEPTH_SUBSAMPLE);
            m.d.assign((size_t)m.rows * m.cols, MAX_DEPTH);
        }
    }
}

// Scene::setDepths(patch, subtract) (Scene.cpp:351-381): subtract == false keeps the smaller depth; subtract == true resets the cell to
// MAX_DEPTH when it still holds exactly this patch's depth (a patch that is removed from the tree: CellProcessor.cpp:76, :273)
void set_depths(Scene& s, const orc_patch_t& p, bool subtract = false) {
    const V4 c{{p.center[0], p.center[1], p.center[2], p.center[3]}};
    for (int k = 0; k < p.nimages; k++) {
        const int idx = p.images[k];
        const Camera& cam = s.cameras[idx];
        const int level = get_leveli(cam, c, p.scale, cam.nlevels - 1);
        const V3 imgC = mult(cam, c, level);
        const int x = (int)((int)(imgC[0] / imgC[2] + 0.5) / DEPTH_SUBSAMPLE);
        const int y = (int)((int)(imgC[1] / imgC[2] + 0.5) / DEPTH_SUBSAMPLE);
        const float d = imgC[2];
        if (!(d >= 0)) continue;              // the reference CHECK-aborts here (Scene.cpp:363)
        DepthMap& m = s.depths[idx][level];
        if (x < 0 || x >= m.cols || y < 0 || y >= m.rows) continue;
        float& old = m.d[(size_t)y * m.cols + x];
        if (old == d && subtract) old = MAX_DEPTH;
        else if (!subtract && d < old) old = d;
    }
}

// Scene::getFullDepth (Scene.cpp:404-432)
float get_full_depth(const Scene& s, int img, int xx, int yy) {
    float depth = MAX_DEPTH;
    int x = (int)(xx / DEPTH_SUBSAMPLE), y = (int)(yy / DEPTH_SUBSAMPLE);
    const int levels = s.cameras[img].nlevels;
    for (int level = 0; level < levels; level++) {
        const DepthMap& m = s.depths[img][level];
        if (x < 0 || x >= m.cols || y < 0 || y >= m.rows) return depth;
        depth = std::min(depth, m.d[(size_t)y * m.cols + x]);
        x /= 2; y /= 2;
    }
    return depth;
}

// Scene::getDetphAtLevel (Scene.cpp:383-402)
float get_depth_at_level(const Scene& s, int img, int xx, int yy, int level) {
    const int x = (int)(xx / DEPTH_SUBSAMPLE), y = (int)(yy / DEPTH_SUBSAMPLE);
    const DepthMap& m = s.depths[img][level];
    if (x < 0 || x >= m.cols || y < 0 || y >= m.rows) return MAX_DEPTH;
    return m.d[(size_t)y * m.cols + x];
}

// Scene::depthTest(patch, ix, iy, depth, image, margin, viewBlock) (Scene.cpp:558-585).
// NOTE `abs(diff)` there is float std::abs(float): real Eigen/Core includes <emmintrin.h> -> <mm_malloc.h> -> <stdlib.h>, whose
// libstdc++ wrapper does `using std::abs` (checked with g++ 13; round 1 had restated the C int abs(int) - former quirk Q18, withdrawn).
bool depth_test_px(const Scene& s, const orc_patch_t& p, int ix, int iy, float depth, int image, float margin, bool viewBlock) {
    if (depth < 0 || ix < 0 || ix >= s.images[image].w[0] || iy < 0 || iy >= s.images[image].h[0]) return false;
    const float imgDepth = get_full_depth(s, image, ix, iy);
    if (imgDepth >= MAX_DEPTH) return viewBlock ? false : true;
    const V4 c{{p.center[0], p.center[1], p.center[2], p.center[3]}};
    const V4 n{{p.normal[0], p.normal[1], p.normal[2], p.normal[3]}};
    const V4 ray = normalized4(sub4(c, s.cameras[image].center));
    const float diff = imgDepth - depth;
    const float factor = std::min(2.0f, 2.0f + dot4(ray, n));
    if (!viewBlock) return std::fabs(diff) < p.scale * margin * factor * 2.0;
    return diff > p.scale * margin * factor * 2.0;
}

// Scene::depthTest(patch, image, margin, neighbours=true, viewBlock) (Scene.cpp:534-556)
bool depth_test(const Scene& s, const orc_patch_t& p, int image, float margin, bool viewBlock) {
    const V4 c{{p.center[0], p.center[1], p.center[2], p.center[3]}};
    const V3 imgC = mult(s.cameras[image], c, 0);
    int ix = (int)(imgC[0] / imgC[2] + 0.5);
    int iy = (int)(imgC[1] / imgC[2] + 0.5);
    ix--; iy--;
    for (int yy = 0; yy < 3; yy++)
        for (int xx = 0; xx < 3; xx++)
            if (!depth_test_px(s, p, ix + xx, iy + yy, imgC[2], image, margin, viewBlock)) return false;
    return true;
}

// Scene::pixelFreeTest (Scene.cpp:595-611)
bool pixel_free_test(const Scene& s, const orc_patch_t& p, int image) {
    const V4 c{{p.center[0], p.center[1], p.center[2], p.center[3]}};
    const Camera& cam = s.cameras[image];
    const int level = (int)std::round(get_level(cam, c, p.scale));
    if (level < 0 || level >= cam.nlevels) return false;
    const V3 imgC = project(cam, c, level);
    const int ix = (int)(imgC[0] / imgC[2] + 0.5), iy = (int)(imgC[1] / imgC[2] + 0.5);
    if (ix < 0 || ix >= s.images[image].w[level] || iy < 0 || iy >= s.images[image].h[level]) return false;
    return get_depth_at_level(s, image, ix, iy, level) == MAX_DEPTH;
}

// depthTests / viewBlockTest / pixelFreeTests (Scene.cpp:518-524, 613-644, 587-593)
void acceptance_counts(const Scene& s, const orc_patch_t& p, float margin, int32_t out[3]) {
    const V4 c{{p.center[0], p.center[1], p.center[2], p.center[3]}};
    int nvis = 0, nblock = 0, nfree = 0;
    for (int k = 0; k < p.nimages; k++) if (depth_test(s, p, p.images[k], margin, false)) ++nvis;
    for (int img = 0; img < (int)s.images.size(); img++) {
        const Camera& cam = s.cameras[img];
        const int level = (int)std::round(get_level(cam, c, p.scale));
        if (level < 0 || level > cam.nlevels - 1) continue;
        const V3 imgC = project(cam, c, level);
        if (imgC[0] < 0 || imgC[0] > s.images[img].w[level] || imgC[1] < 0 || imgC[1] > s.images[img].h[level]) continue;
        if (depth_test(s, p, img, margin, true)) nblock++;
    }
    for (int k = 0; k < p.nimages; k++) if (pixel_free_test(s, p, p.images[k])) ++nfree;
    out[0] = nvis; out[1] = nblock; out[2] = nfree;
}

// candidate construction of CellProcessor::extend (mode 6, CellProcessor.cpp:98-119) and ::branch (mode 4, :227-249)
void expand_candidates(const Scene& s, const orc_patch_t& p, float width, int mode, orc_patch_t* out) {
    const V3 n{{p.normal[0], p.normal[1], p.normal[2]}};
    const V3 imgX = s.cameras[p.images[0]].xAxis;
    const V3 yaxis = normalized3(cross3(n, imgX));
    const V3 xaxis = cross3(yaxis, n);
    const int N = mode;
    const float extend = (mode == 6) ? width : (float)(width / 4.0);
    for (int ii = 0; ii < N; ii++) {
        const float angle = (mode == 6) ? (float)(2.0 * M_PI / N * ii) : (float)(2.0 * M_PI / N * ii + M_PI / 4);
        const float dx = ::cos((double)angle);
        const float dy = ::sin((double)angle);
        orc_patch_t q = p;
        for (int k = 0; k < 3; k++) q.center[k] = p.center[k] + (dx * xaxis[k] + dy * yaxis[k]) * extend;
        q.scale = (mode == 6) ? (float)(width * 0.9 / 2.0) : (float)(width * 0.45 / 2.0);
        out[ii] = q;
    }
}

int optimize_one(const Scene& s, orc_patch_t& p) {
    PatchOptimizer po(&s);
    po.load(p);
    const bool ok = po.runOptimization();
    p.evals = po.evals; p.textures = po.textures; p.nlopt_result = po.nlopt_result; p.last_val = po.last_val;
    if (!ok) { p.status = po.fail_stage; return p.status; }
    if ((int)po.pImages.size() > ORC_MAX_VIEWS) { p.status = ORC_FAIL_TOO_MANY_VIEWS; return p.status; }
    // PatchOptimizer.cpp:86-100
    for (int i = 0; i < 4; i++) { p.center[i] = po.pCenter[i]; p.normal[i] = po.pNormal[i]; }
    p.scale = po.pScale;
    p.nimages = (int)po.pImages.size();
    for (int i = 0; i < p.nimages; i++) p.images[i] = po.pImages[i];
    p.ncc = 1.4f;
    const V3 col = patch_color(s, po.pCenter, po.pScale, p.images, p.nimages);
    p.color[0] = col[0]; p.color[1] = col[1]; p.color[2] = col[2];
    p.status = ORC_OK;
    return ORC_OK;
}

// ------------------------------------------------------------------------------------------
// analytic test objectives (n=3) used to pin BOBYQA implementations against the real nlopt
// ------------------------------------------------------------------------------------------
double testfunc(int id, const double* x) {
    switch (id) {
    case 0: {  // Box-Betts, thirdLibs/nlopt-2.4.2/test/testfuncs.c:65-89
        double f = 0;
        for (int i = 1; i <= 10; ++i) {
            const double e0 = std::exp(-0.1 * i * x[0]);
            const double e1 = std::exp(-0.1 * i * x[1]);
            const double e2 = std::exp(-0.1 * i) - std::exp((double)-i);
            const double g = e0 - e1 - e2 * x[2];
            f += g * g;
        }
        return f;
    }
    case 1: {  // 3-D Rosenbrock
        const double a = x[1] - x[0] * x[0], b = 1 - x[0], c = x[2] - x[1] * x[1], d = 1 - x[1];
        return 100 * a * a + b * b + 100 * c * c + d * d;
    }
    case 2: {  // smooth NCC-shaped basin r/(1+3r), r = 1-exp(-|A(x-m)|^2/2)
        const double u = (x[0] - 0.013) / 0.05, v = (x[1] - 3.7) / 9.0, w = (x[2] + 2.2) / 7.0;
        const double r = 1.0 - std::exp(-0.5 * (u * u + v * v + w * w + 0.3 * u * v));
        return r / (1.0 + 3.0 * r);
    }
    case 3: {  // same basin, value quantised to float (f32 noise floor like the real objective)
        const double u = (x[0] + 0.021) / 0.03, v = (x[1] + 5.1) / 10.0, w = (x[2] - 1.4) / 6.0;
        const float r = (float)(1.0 - std::exp(-0.5 * (u * u + v * v + w * w)));
        return (double)(r / (1 + 3 * r));
    }
    case 4: {  // non-smooth: |.| kinks + a plateau of 2.0 outside a slab (mimics invalid samples)
        if (std::fabs(x[0]) > 0.8) return 2.0;
        return std::fabs(x[0] - 0.1) + 0.05 * std::fabs(x[1] - 2.0) + 0.02 * std::fabs(x[2] + 3.0) +
               0.001 * std::floor(40.0 * x[0]) * 0.01;
    }
    case 5: {  // narrow depth basin (many evaluations) with coarse quantisation -> exercises rescue/roundoff exits
        const double u = (x[0] - 0.0004) / 0.001, v = (x[1] - 1.0) / 12.0, w = (x[2] - 0.5) / 12.0;
        const double r = 1.0 - std::exp(-0.5 * (u * u + v * v + w * w));
        return std::floor(1e6 * r / (1.0 + 3.0 * r)) * 1e-6;
    }
    case 6: {  // flat: constant (degenerate model)
        return 0.5;
    }
    case 7: {  // separable quadratic with minimum on a bound
        const double a = x[0] - 0.3, b = x[1] - 40.0, c = x[2] + 40.0;
        return a * a + 0.01 * b * b + 0.02 * c * c;
    }
    case 8: {  // ill-conditioned valley with a rotating axis (long runs, near-singular interpolation sets)
        const double a = x[0] * 1e3 - x[1] * x[1] * 1e-3, b = x[1] - 7.0 + 1e-2 * x[2] * x[2], c = x[2] + 11.0;
        return 1e6 * a * a + b * b + 1e-6 * c * c * c * c;
    }
    default: {  // id >= 100: seeded family of noisy / quantised / kinked basins (fuzzing of rare BOBYQA branches)
        unsigned h = (unsigned)id * 2654435761u;
        auto rnd = [&h]() { h ^= h << 13; h ^= h >> 17; h ^= h << 5; return (double)(h & 0xffffff) / 16777216.0; };
        const double c0 = (rnd() - 0.5) * 0.4, c1 = (rnd() - 0.5) * 40.0, c2 = (rnd() - 0.5) * 40.0;
        const double s0 = 0.002 + 0.3 * rnd() * rnd(), s1 = 2.0 + 20.0 * rnd(), s2 = 2.0 + 20.0 * rnd();
        const double q = std::pow(10.0, -2.0 - 6.0 * rnd());
        const double namp = std::pow(10.0, -1.0 - 7.0 * rnd());
        const int mode = (int)(rnd() * 4.0);
        const double u = (x[0] - c0) / s0, v = (x[1] - c1) / s1, w = (x[2] - c2) / s2;
        double r = 1.0 - std::exp(-0.5 * (u * u + v * v + w * w + 0.5 * u * w));
        if (mode == 1) r = std::min(1.0, 0.3 * (std::fabs(u) + std::fabs(v) + std::fabs(w)));
        double f = r / (1.0 + 3.0 * r);
        // deterministic pseudo-noise from the bit pattern of x
        unsigned long long b0, b1, b2;
        std::memcpy(&b0, &x[0], 8); std::memcpy(&b1, &x[1], 8); std::memcpy(&b2, &x[2], 8);
        unsigned long long hh = (b0 * 0x9E3779B97F4A7C15ull) ^ (b1 * 0xC2B2AE3D27D4EB4Full) ^ (b2 * 0x165667B19E3779F9ull);
        hh ^= hh >> 29; hh *= 0xBF58476D1CE4E5B9ull; hh ^= hh >> 32;
        if (mode >= 2) f += namp * ((double)(hh & 0xfffff) / 1048576.0 - 0.5);
        if (mode == 3 || mode == 0) f = std::floor(f / q) * q;
        if (std::fabs(u) > 60.0) f = 2.0;
        if (id >= 5000) {  // second family: badly scaled polynomial valleys (aims at the RESCUE branch)
            const double k0 = std::pow(10.0, 6.0 * rnd()), k1 = std::pow(10.0, -6.0 * rnd());
            const double a = u - 0.5 * v * v * rnd(), b = v + w * w * rnd(), c = w;
            f = k0 * a * a + b * b + k1 * c * c * c * c;
            if (mode >= 2) f *= 1.0 + 1e-13 * ((double)(hh & 0xff) - 128.0);
        }
        return f;
    }
    }
}

struct TraceData { int id; double* tx; double* tf; int cap; int n; };
double trace_fn(unsigned, const double* x, double*, void* d) {
    TraceData* t = static_cast<TraceData*>(d);
    const double f = testfunc(t->id, x);
    if (t->n < t->cap) { t->tx[3 * t->n] = x[0]; t->tx[3 * t->n + 1] = x[1]; t->tx[3 * t->n + 2] = x[2]; t->tf[t->n] = f; }
    t->n++;
    return f;
}

}  // namespace

// ==========================================================================================
// C API
// ==========================================================================================
extern "C" {

void* orc_scene_new(const orc_options_t* opt) { Scene* s = new Scene; s->opt = *opt; return s; }
void orc_scene_free(void* scene) { delete static_cast<Scene*>(scene); }

int orc_add_camera(void* scene, double f, const double q[4], const double c[3], int width, int height, const uint8_t* rgb) {
    Scene* s = static_cast<Scene*>(scene);
    s->cameras.emplace_back();
    s->images.emplace_back();
    Image& im = s->images.back();
    const int maxLevel = std::max(1, s->opt.maxlevel);  // Image.cpp:34-39
    im.w[0] = width; im.h[0] = height;
    im.lvl[0].assign(rgb, rgb + (size_t)3 * width * height);
    for (int l = 1; l <= maxLevel; l++) half_xy_interleaved(im.lvl[l - 1], im.w[l - 1], im.h[l - 1], im.lvl[l], im.w[l], im.h[l]);
    camera_init(s->cameras.back(), f, q, c, width, height, s->opt.maxlevel);
    s->covis.resize(s->cameras.size());
    return (int)s->cameras.size() - 1;
}

int orc_num_cameras(void* scene) { return (int)static_cast<Scene*>(scene)->cameras.size(); }

void orc_get_camera(void* scene, int idx, orc_camera_t* out) {
    Scene* s = static_cast<Scene*>(scene);
    const Camera& c = s->cameras[idx];
    std::memcpy(out->P, c.P, sizeof(c.P));
    for (int i = 0; i < 4; i++) out->center[i] = c.center[i];
    for (int i = 0; i < 3; i++) { out->xaxis[i] = c.xAxis[i]; out->yaxis[i] = c.yAxis[i]; out->zaxis[i] = c.zAxis[i]; }
    out->k00 = c.K0[0][0]; out->k11 = c.K0[1][1];
    for (int l = 0; l < ORC_LEVELS; l++) { out->width[l] = s->images[idx].w[l]; out->height[l] = s->images[idx].h[l]; }
}

const uint8_t* orc_get_image(void* scene, int cam, int level, int* w, int* h) {
    Scene* s = static_cast<Scene*>(scene);
    *w = s->images[cam].w[level]; *h = s->images[cam].h[level];
    return s->images[cam].lvl[level].data();
}

// Scene.cpp:241-298.  NOTE the counter is indexed by measurement POSITION (ii,jj), not by camera id
// (Scene.cpp:260-264) - reproduced on purpose.
void orc_extract_covis(void* scene, int npoints, const int32_t* off, const int32_t* cam) {
    Scene* s = static_cast<Scene*>(scene);
    const int n = (int)s->cameras.size();
    std::vector<int> vis((size_t)n * n, 0);
    for (int p = 0; p < npoints; p++) {
        const int m = off[p + 1] - off[p];
        (void)cam;
        for (int ii = 0; ii < m; ii++)
            for (int jj = 0; jj < m; jj++)
                if (ii != jj && ii < n && jj < n) vis[(size_t)ii * n + jj]++;
    }
    s->covis.assign(n, {});
    for (int ii = 0; ii < n; ii++)
        for (int jj = 0; jj < n; jj++)
            if (vis[(size_t)ii * n + jj] >= 50) s->covis[ii].push_back(jj);
}

void orc_set_covis(void* scene, const int32_t* offsets, const int32_t* ids) {
    Scene* s = static_cast<Scene*>(scene);
    const int n = (int)s->cameras.size();
    s->covis.assign(n, {});
    for (int i = 0; i < n; i++) s->covis[i].assign(ids + offsets[i], ids + offsets[i + 1]);
}

int orc_get_covis(void* scene, int cam, int32_t* out, int cap) {
    Scene* s = static_cast<Scene*>(scene);
    const auto& v = s->covis[cam];
    for (int i = 0; i < (int)v.size() && i < cap; i++) out[i] = v[i];
    return (int)v.size();
}

// Scene.cpp:116-165
void orc_seed_patches(void* scene, int npoints, const double* xyz, const int32_t* off, const int32_t* mcam, orc_patch_t* out, uint8_t* valid) {
    Scene* s = static_cast<Scene*>(scene);
    const orc_options_t& o = s->opt;
    const int cSize = 2;
    for (int ii = 0; ii < npoints; ii++) {
        orc_patch_t& p = out[ii];
        std::memset(&p, 0, sizeof(p));
        valid[ii] = 0;
        V4 center{{(float)xyz[3 * ii], (float)xyz[3 * ii + 1], (float)xyz[3 * ii + 2], 1.0f}};
        const int nm = off[ii + 1] - off[ii];
        if (nm < o.min_images_per_patch) continue;
        std::vector<int> imgs;
        for (int k = off[ii]; k < off[ii + 1]; k++) {
            const int idx = mcam[k];
            if (idx < 0) continue;
            const V3 pr = project(s->cameras[idx], center, o.start_level);
            const int margin = cSize;
            if (pr[0] < margin || pr[1] < margin || pr[0] >= s->images[idx].w[o.start_level] - margin ||
                pr[1] >= s->images[idx].h[o.start_level] - margin)
                continue;
            imgs.push_back(idx);
        }
        if (imgs.size() < 2 || (int)imgs.size() > ORC_MAX_VIEWS) continue;
        V4 nrm = sub4(s->cameras[imgs[0]].center, center);  // Scene.cpp:159 (first camera only)
        // Vector4f::normalize(): z = squaredNorm; if (z>0) v /= sqrt(z)
        nrm = normalized4(nrm);
        nrm[3] = 0.0f;
        for (int i = 0; i < 4; i++) { p.center[i] = center[i]; p.normal[i] = nrm[i]; }
        p.scale = get_scale(s->cameras[imgs[0]], center, o.start_level);
        p.nimages = (int)imgs.size();
        for (int i = 0; i < p.nimages; i++) p.images[i] = imgs[i];
        valid[ii] = 1;
    }
}

int orc_optimize(void* scene, orc_patch_t* patch) { return optimize_one(*static_cast<Scene*>(scene), *patch); }

// same parallel structure as Scene.cpp:94-114: one optimizer per thread, omp parallel for over patches
void orc_optimize_batch(void* scene, int n, orc_patch_t* patches, int nthreads) {
    const Scene& s = *static_cast<Scene*>(scene);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 8)
#endif
    for (int i = 0; i < n; i++) optimize_one(s, patches[i]);
}

void orc_set_inccs(void* scene, const orc_patch_t* patch, int ref_idx, int robust, float* inccs) {
    PatchOptimizer po(static_cast<Scene*>(scene));
    po.load(*patch);
    std::vector<float> v;
    po.setINCCs(v, po.pImages, ref_idx, robust);
    for (size_t i = 0; i < v.size(); i++) inccs[i] = v[i];
}

int orc_sample_texture(void* scene, const float center[4], float scale, const float xaxis[4], const float yaxis[4],
                       const float zaxis[4], int cam, float* tex147) {
    PatchOptimizer po(static_cast<Scene*>(scene));
    V4 c, x, y, z;
    for (int i = 0; i < 4; i++) { c[i] = center[i]; x[i] = xaxis[i]; y[i] = yaxis[i]; z[i] = zaxis[i]; }
    PatchTex t;
    if (!po.sampleTexture(c, scale, x, y, z, cam, t)) return 0;
    std::memcpy(tex147, t.data, sizeof(t.data));
    return 1;
}

double orc_objective(void* scene, const orc_patch_t* patch, const double x[3]) {
    PatchOptimizer po(static_cast<Scene*>(scene));
    po.load(*patch);
    po.setOptimizationFields();
    po.setCenterNorm(x);
    return po.objective_fn();
}

void orc_patch_color(void* scene, const orc_patch_t* patch, float rgb[3]) {
    V4 c{{patch->center[0], patch->center[1], patch->center[2], patch->center[3]}};
    const V3 col = patch_color(*static_cast<Scene*>(scene), c, patch->scale, patch->images, patch->nimages);
    rgb[0] = col[0]; rgb[1] = col[1]; rgb[2] = col[2];
}

void orc_set_cr_asinf(int on) { g_cr_asinf = on; }

void orc_depth_reset(void* scene) { alloc_depths(*static_cast<Scene*>(scene)); }
void orc_depth_set_batch(void* scene, int n, const orc_patch_t* patches) {
    Scene& s = *static_cast<Scene*>(scene);
    if (s.depths.size() != s.cameras.size()) alloc_depths(s);
    for (int i = 0; i < n; i++) if (patches[i].status == ORC_OK) set_depths(s, patches[i]);
}
void orc_depth_unset_batch(void* scene, int n, const orc_patch_t* patches) {
    Scene& s = *static_cast<Scene*>(scene);
    if (s.depths.size() != s.cameras.size()) alloc_depths(s);
    for (int i = 0; i < n; i++) if (patches[i].status == ORC_OK) set_depths(s, patches[i], true);
}
const float* orc_get_depth(void* scene, int cam, int level, int* rows, int* cols) {
    Scene& s = *static_cast<Scene*>(scene);
    if (s.depths.size() != s.cameras.size()) alloc_depths(s);
    *rows = s.depths[cam][level].rows; *cols = s.depths[cam][level].cols;
    return s.depths[cam][level].d.data();
}
void orc_accept_batch(void* scene, int n, const orc_patch_t* patches, float margin, int32_t* out) {
    Scene& s = *static_cast<Scene*>(scene);
    if (s.depths.size() != s.cameras.size()) alloc_depths(s);
    for (int i = 0; i < n; i++) acceptance_counts(s, patches[i], margin, out + 3 * i);
}
void orc_expand_candidates(void* scene, int n, const orc_patch_t* parents, const float* widths, int mode, orc_patch_t* out) {
    const Scene& s = *static_cast<Scene*>(scene);
    for (int i = 0; i < n; i++) expand_candidates(s, parents[i], widths[i], mode, out + (size_t)mode * i);
}

double orc_testfunc_eval(int func_id, const double x[3]) { return testfunc(func_id, x); }

int orc_bobyqa_testfunc(int func_id, const double x0[3], const double lb[3], const double ub[3], double xtol_rel,
                        int maxeval, double xout[3], double* fout, double* trace_x, double* trace_f, int trace_cap, int* nevals) {
    TraceData t{func_id, trace_x, trace_f, trace_cap, 0};
    nlopt_opt o = nlopt_create(PatchOptimizer::bobyqa_id(), 3);
    nlopt_set_min_objective(o, trace_fn, &t);
    nlopt_set_xtol_rel(o, xtol_rel);
    nlopt_set_maxeval(o, maxeval);
    nlopt_set_lower_bounds(o, lb);
    nlopt_set_upper_bounds(o, ub);
    double x[3] = {x0[0], x0[1], x0[2]};
    double minf = 0;
    const int r = nlopt_optimize(o, x, &minf);
    nlopt_destroy(o);
    xout[0] = x[0]; xout[1] = x[1]; xout[2] = x[2];
    *fout = minf;
    *nevals = t.n;
    return r;
}

}  // extern "C"
