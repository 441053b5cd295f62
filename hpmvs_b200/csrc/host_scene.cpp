// Host-side scene surface of hpmvs_b200: the small amount of C++ that sits between an NVM model and the GPU
// engine (camera tables, covisibility lists, seed-patch candidates).  It mirrors the reference's host code for
// these steps and keeps its f32 evaluation order (Eigen >= 3.3 on baseline x86-64: 4-vectors reduce as
// (a0+a2)+(a1+a3), 3-vectors as a0+(a1+a2), small matrix products are coefficient based), because the values
// produced here are inputs of a bit-reproducible optimisation.  Build with -ffp-contract=off.
#include <limits.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "../../include/hpmvs_b200.h"

namespace {

struct Vec3 { float x, y, z; };

inline float dot(const Vec3& a, const Vec3& b) { const float p0 = a.x * b.x, p1 = a.y * b.y, p2 = a.z * b.z; return p0 + (p1 + p2); }
inline Vec3 cross(const Vec3& a, const Vec3& b) { return Vec3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline Vec3 unit(const Vec3& a) {
    const float z = dot(a, a);
    if (!(z > 0.0f)) return a;
    const float s = sqrtf(z);
    return Vec3{a.x / s, a.y / s, a.z / s};
}
inline float sum4(float p0, float p1, float p2, float p3) { return (p0 + p2) + (p1 + p3); }

// Camera::project (Camera.h:45-62) on a camera record
inline void project_pt(const hpmvs_camera_t& cam, const float X[4], int level, float r[3]) {
    for (int i = 0; i < 3; i++) {
        const float* p = cam.P[level][i];
        r[i] = (p[0] * X[0] + p[1] * X[1]) + (p[2] * X[2] + p[3] * X[3]);
    }
    if (r[2] <= 0.0f) { r[0] = -65535.0f; r[1] = -65535.0f; r[2] = -1.0f; return; }
    const float z = r[2];
    r[0] /= z; r[1] /= z; r[2] /= z;
    const float lo = (float)(INT_MIN + 3.0f), hi = (float)(INT_MAX - 3.0f);
    r[0] = fmaxf(lo, fminf(hi, r[0]));
    r[1] = fmaxf(lo, fminf(hi, r[1]));
}

}  // namespace

extern "C" {

int hpmvs_camera_from_nvm(double f, const double q[4], const double c[3], int width, int height, int maxlevel,
                          hpmvs_camera_t* out) {
    if (!q || !c || !out || width <= 0 || height <= 0 || maxlevel < 1 || maxlevel >= HPMVS_LEVELS) return HPMVS_E_ARG;
    memset(out, 0, sizeof(*out));
    const float k[3][3] = {{(float)f, 0.0f, (float)(width / 2.0)}, {0.0f, (float)f, (float)(height / 2.0)}, {0.0f, 0.0f, 1.0f}};
    // unit quaternion (w,x,y,z) -> rotation, in double, then narrowed (Camera.cpp:43-50)
    const double qw = q[0], qx = q[1], qy = q[2], qz = q[3];
    const double x2 = 2.0 * qx, y2 = 2.0 * qy, z2 = 2.0 * qz;
    const double wx = x2 * qw, wy = y2 * qw, wz = z2 * qw;
    const double xx = x2 * qx, xy = y2 * qx, xz = z2 * qx;
    const double yy = y2 * qy, yz = z2 * qy, zz = z2 * qz;
    const double rd[3][3] = {{1.0 - (yy + zz), xy - wz, xz + wy}, {xy + wz, 1.0 - (xx + zz), yz - wx}, {xz - wy, yz + wx, 1.0 - (xx + yy)}};
    float rt[3][4];
    const float cf[3] = {(float)c[0], (float)c[1], (float)c[2]};
    for (int i = 0; i < 3; i++) {
        float ri[3];
        for (int j = 0; j < 3; j++) { ri[j] = (float)rd[i][j]; rt[i][j] = ri[j]; }
        rt[i][3] = (-ri[0]) * cf[0] + ((-ri[1]) * cf[1] + (-ri[2]) * cf[2]);
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) out->P[0][i][j] = k[i][0] * rt[0][j] + (k[i][1] * rt[1][j] + k[i][2] * rt[2][j]);
    for (int l = 1; l <= maxlevel; l++)
        for (int j = 0; j < 4; j++) {
            out->P[l][0][j] = out->P[l - 1][0][j] / 2.0f;
            out->P[l][1][j] = out->P[l - 1][1][j] / 2.0f;
            out->P[l][2][j] = out->P[l - 1][2][j];
        }
    out->center[0] = cf[0]; out->center[1] = cf[1]; out->center[2] = cf[2]; out->center[3] = 1.0f;
    const Vec3 row2{out->P[0][2][0], out->P[0][2][1], out->P[0][2][2]};
    // row(2).head(3).norm(): a dynamic-size block, reduced by Eigen's scalar loop (a0+a1)+a2 (Camera.cpp:71)
    const float n2 = sqrtf((row2.x * row2.x + row2.y * row2.y) + row2.z * row2.z);
    const Vec3 zax{row2.x / n2, row2.y / n2, row2.z / n2};
    const Vec3 row0{out->P[0][0][0], out->P[0][0][1], out->P[0][0][2]};
    const Vec3 yax = unit(cross(zax, row0));
    const Vec3 xax = unit(cross(yax, zax));
    out->xaxis[0] = xax.x; out->xaxis[1] = xax.y; out->xaxis[2] = xax.z;
    out->yaxis[0] = yax.x; out->yaxis[1] = yax.y; out->yaxis[2] = yax.z;
    out->zaxis[0] = zax.x; out->zaxis[1] = zax.y; out->zaxis[2] = zax.z;
    out->k00 = k[0][0]; out->k11 = k[1][1];
    int w = width, h = height;
    for (int l = 0; l <= maxlevel; l++) { out->width[l] = w; out->height[l] = h; w /= 2; h /= 2; }
    return 0;
}

int hpmvs_extract_covis(int ncams, int npoints, const int32_t* off, const int32_t* mcam, int compat, int32_t* out_offsets,
                        int32_t* out_ids, int ids_cap) {
    if (ncams <= 0 || npoints < 0 || !off || !out_offsets) return HPMVS_E_ARG;
    std::vector<int> shared((size_t)ncams * ncams, 0);
    for (int p = 0; p < npoints; p++) {
        const int a = off[p], b = off[p + 1];
        for (int i = a; i < b; i++)
            for (int j = a; j < b; j++) {
                if (i == j) continue;
                const int r = compat ? (i - a) : mcam[i], c = compat ? (j - a) : mcam[j];
                if (r >= 0 && r < ncams && c >= 0 && c < ncams) shared[(size_t)r * ncams + c]++;
            }
    }
    int total = 0;
    for (int i = 0; i < ncams; i++) {
        out_offsets[i] = total;
        for (int j = 0; j < ncams; j++)
            if (shared[(size_t)i * ncams + j] >= 50) {
                if (out_ids && total < ids_cap) out_ids[total] = j;
                total++;
            }
    }
    out_offsets[ncams] = total;
    if (total > ids_cap) return -total;
    return total;
}

int hpmvs_seed_patches(const hpmvs_options_t* opt, int ncams, const hpmvs_camera_t* cams, int npoints, const double* xyz,
                       const int32_t* off, const int32_t* mcam, hpmvs_patch_t* out, uint8_t* valid) {
    if (!opt || !cams || ncams <= 0 || npoints < 0 || !xyz || !off || !mcam || !out || !valid) return HPMVS_E_ARG;
    const int lvl = opt->start_level;
    if (lvl < 0 || lvl >= HPMVS_LEVELS) return HPMVS_E_ARG;
    const int margin = 2;
    for (int ii = 0; ii < npoints; ii++) {
        hpmvs_patch_t& p = out[ii];
        memset(&p, 0, sizeof(p));
        valid[ii] = 0;
        const float X[4] = {(float)xyz[3 * ii], (float)xyz[3 * ii + 1], (float)xyz[3 * ii + 2], 1.0f};
        if (off[ii + 1] - off[ii] < opt->min_images_per_patch) continue;
        int n = 0, overflow = 0;
        for (int k = off[ii]; k < off[ii + 1]; k++) {
            const int idx = mcam[k];
            if (idx < 0 || idx >= ncams) continue;
            float r[3];
            project_pt(cams[idx], X, lvl, r);
            if (r[0] < margin || r[1] < margin || r[0] >= cams[idx].width[lvl] - margin || r[1] >= cams[idx].height[lvl] - margin) continue;
            if (n < HPMVS_MAX_VIEWS) p.images[n++] = idx; else overflow = 1;
        }
        if (n < 2 || overflow) continue;
        const hpmvs_camera_t& c0 = cams[p.images[0]];
        const float d[4] = {c0.center[0] - X[0], c0.center[1] - X[1], c0.center[2] - X[2], c0.center[3] - X[3]};
        const float z = sum4(d[0] * d[0], d[1] * d[1], d[2] * d[2], d[3] * d[3]);
        const float s = sqrtf(z);
        for (int i = 0; i < 4; i++) { p.center[i] = X[i]; p.normal[i] = (z > 0.0f) ? d[i] / s : d[i]; }
        p.normal[3] = 0.0f;
        // Camera::getScale(centre, START_LEVEL) (Camera.cpp:83-90); |X - c| has the same squares as |c - X|
        const float ftmp = c0.k00 + c0.k11;
        p.scale = (ftmp == 0.0f) ? 1.0f : (float)(2.0 * s * (0x0001 << lvl) / ftmp);
        p.nimages = n;
        valid[ii] = 1;
    }
    return 0;
}

int hpmvs_expand_candidates(int ncams, const hpmvs_camera_t* cams, int n, const hpmvs_patch_t* parents, const float* widths,
                            int mode, hpmvs_patch_t* out) {
    if (!cams || ncams <= 0 || n < 0 || (n > 0 && (!parents || !widths || !out)) || (mode != 4 && mode != 6)) return HPMVS_E_ARG;
    for (int i = 0; i < n; i++) {
        const hpmvs_patch_t& p = parents[i];
        if (p.nimages <= 0 || p.images[0] < 0 || p.images[0] >= ncams) return HPMVS_E_ARG;
        const hpmvs_camera_t& rc = cams[p.images[0]];
        const Vec3 nrm{p.normal[0], p.normal[1], p.normal[2]};
        const Vec3 ya = unit(cross(nrm, Vec3{rc.xaxis[0], rc.xaxis[1], rc.xaxis[2]}));
        const Vec3 xa = cross(ya, nrm);
        const float width = widths[i];
        const float extend = (mode == 6) ? width : (float)(width / 4.0);
        for (int ii = 0; ii < mode; ii++) {
            const float angle = (mode == 6) ? (float)(2.0 * M_PI / mode * ii) : (float)(2.0 * M_PI / mode * ii + M_PI / 4);
            const float dx = (float)cos((double)angle), dy = (float)sin((double)angle);
            hpmvs_patch_t q = p;
            q.center[0] = p.center[0] + (dx * xa.x + dy * ya.x) * extend;
            q.center[1] = p.center[1] + (dx * xa.y + dy * ya.y) * extend;
            q.center[2] = p.center[2] + (dx * xa.z + dy * ya.z) * extend;
            q.scale = (mode == 6) ? (float)(width * 0.9 / 2.0) : (float)(width * 0.45 / 2.0);
            out[(size_t)mode * i + ii] = q;
        }
    }
    return 0;
}

}  // extern "C"
