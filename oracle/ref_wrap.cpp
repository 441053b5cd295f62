// TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT.
//
// C entry points around the REFERENCE'S OWN classes (mo3d::Scene, mo3d::PatchOptimizer, mo3d::NVMReader ...),
// compiled from /root/reference/src/hpmvs/*.cpp where they lie (oracle/Makefile, target `refhpmvs`) into
// oracle/_ref/libhpmvs_ref.so.  Nothing of the reference is copied here: this file only calls its public
// interface (include/hpmvs/Scene.h, PatchOptimizer.h, NVMReader.h, Patch3d.h) and converts to the oracle's PODs.
// Used by tests/ to pin the restatement in hpmvs_oracle.cpp against the real code path, and by bench.py's
// reference arm (cpu_baseline.kind = "reference").
// include order as src/hpmvs/PatchOptimizer.cpp:21-33 (HpmvsOptions.h is not self-contained)
#include <cmath>
#include <string>
#include <hpmvs/Scene.h>
#include <hpmvs/PatchOptimizer.h>

#include <cstring>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "hpmvs_oracle.h"

namespace {

struct RefScene {
    mo3d::HpmvsOptions options;
    mo3d::NVM_Model model;
    mo3d::Scene scene;
};

mo3d::HpmvsOptions to_options(const orc_options_t* o) {
    mo3d::HpmvsOptions r;   // defaults of include/hpmvs/HpmvsOptions.h:29-58
    if (o) {
        r.MAXLEVEL = o->maxlevel; r.MINLEVEL = o->minlevel; r.START_LEVEL = o->start_level;
        r.MAX_ANGLE = o->max_angle; r.MIN_ANGLE = o->min_angle;
        r.MAX_IMAGES_PER_PATCH = o->max_images_per_patch; r.MIN_IMAGES_PER_PATCH = o->min_images_per_patch;
        r.NCC_ALPHA_1 = o->ncc_alpha_1; r.NCC_ALPHA_2 = o->ncc_alpha_2;
    }
    return r;
}

void to_patch3d(const orc_patch_t& in, mo3d::Patch3d& p) {
    for (int i = 0; i < 4; i++) { p.center_[i] = in.center[i]; p.normal_[i] = in.normal[i]; }
    p.scale_3dx_ = in.scale;
    p.images_.assign(in.images, in.images + in.nimages);
}

void from_patch3d(const mo3d::Patch3d& p, orc_patch_t& out) {
    for (int i = 0; i < 4; i++) { out.center[i] = p.center_[i]; out.normal[i] = p.normal_[i]; }
    out.scale = p.scale_3dx_;
    out.nimages = (int32_t)std::min<size_t>(p.images_.size(), ORC_MAX_VIEWS);
    for (int i = 0; i < out.nimages; i++) out.images[i] = p.images_[i];
    for (int i = 0; i < 3; i++) out.color[i] = p.color_[i];
    out.ncc = p.ncc_;
}

}  // namespace

extern "C" {

// main.cpp:104-113: NVMReader::readFile(..., true) -> models[0]; Scene::addCameras; Scene::extractCoVisiblilty
void* refh_scene_load(const char* nvm_path, const orc_options_t* opt) {
    std::unique_ptr<RefScene> s(new RefScene);
    s->options = to_options(opt);
    std::vector<mo3d::NVM_Model> models;
    mo3d::NVMReader::readFile(nvm_path, models, true);
    if (models.empty()) return nullptr;
    s->model = models[0];
    if (!s->scene.addCameras(s->model, s->options)) return nullptr;
    if (!s->scene.extractCoVisiblilty(s->model, s->options)) return nullptr;
    return s.release();
}

void refh_scene_free(void* h) { delete static_cast<RefScene*>(h); }

// NVMReader::saveNVM (NVMReader.cpp:157-183) of the model that was loaded
void refh_save_nvm(void* h, const char* path) {
    std::vector<mo3d::NVM_Model> models(1, static_cast<RefScene*>(h)->model);
    mo3d::NVMReader::saveNVM(path, models);
}

int refh_num_cameras(void* h) { return (int)static_cast<RefScene*>(h)->scene.cameras_.size(); }
int refh_num_points(void* h) { return (int)static_cast<RefScene*>(h)->model.points.size(); }

void refh_get_camera(void* h, int idx, orc_camera_t* out) {
    const RefScene* s = static_cast<RefScene*>(h);
    const mo3d::Camera& c = s->scene.cameras_[idx];
    std::memset(out, 0, sizeof(*out));
    for (int l = 0; l < c.getLevels() && l < ORC_LEVELS; l++) {
        for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) out->P[l][i][j] = c.projection_[l](i, j);
        out->width[l] = s->scene.images_[idx].getWidth(l);
        out->height[l] = s->scene.images_[idx].getHeight(l);
    }
    for (int i = 0; i < 4; i++) out->center[i] = c.center_[i];
    for (int i = 0; i < 3; i++) { out->xaxis[i] = c.xAxis_[i]; out->yaxis[i] = c.yAxis_[i]; out->zaxis[i] = c.zAxis_[i]; }
    out->k00 = c.kMat_[0](0, 0); out->k11 = c.kMat_[0](1, 1);
}

// pyramid level as interleaved u8 RGB, row stride 3*w (Image::getImage un-permutes to planar x,y,1,c)
int refh_get_image(void* h, int cam, int level, uint8_t* out, int cap, int* w, int* h_out) {
    const RefScene* s = static_cast<RefScene*>(h);
    const cimg_library::CImg<unsigned char> img = s->scene.images_[cam].getImage(level);
    *w = img.width(); *h_out = img.height();
    const int need = img.width() * img.height() * 3;
    if (!out || cap < need) return need;
    for (int y = 0; y < img.height(); y++)
        for (int x = 0; x < img.width(); x++)
            for (int c = 0; c < 3; c++) out[3 * (y * img.width() + x) + c] = img(x, y, 0, c);
    return need;
}

int refh_get_covis(void* h, int cam, int32_t* out, int cap) {
    const RefScene* s = static_cast<RefScene*>(h);
    const std::vector<int>& v = s->scene.covis_[cam];
    for (int i = 0; i < (int)v.size() && i < cap; i++) out[i] = v[i];
    return (int)v.size();
}

// Scene::getColor(idx, x, y, level) -> Image::getColor (Scene.h:122-124, Image.h:89-115)
void refh_get_color(void* h, int cam, float x, float y, int level, float rgb[3]) {
    const RefScene* s = static_cast<RefScene*>(h);
    const Eigen::Vector3f c = s->scene.getColor(cam, x, y, level);
    rgb[0] = c[0]; rgb[1] = c[1]; rgb[2] = c[2];
}

// Camera::project / getScale / getLevel / getLeveli (Camera.h:45-62, Camera.cpp:83-99)
void refh_project(void* h, int cam, const float X[4], int level, float out[3]) {
    const RefScene* s = static_cast<RefScene*>(h);
    const Eigen::Vector3f r = s->scene.cameras_[cam].project(Eigen::Vector4f(X[0], X[1], X[2], X[3]), level);
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}
void refh_scale_level(void* h, int cam, const float X[4], float scale, int level, int maxLevel, float* getScale, float* getLevel, int* getLeveli) {
    const RefScene* s = static_cast<RefScene*>(h);
    const Eigen::Vector4f c(X[0], X[1], X[2], X[3]);
    *getScale = s->scene.cameras_[cam].getScale(c, level);
    *getLevel = s->scene.cameras_[cam].getLevel(c, scale);
    *getLeveli = s->scene.cameras_[cam].getLeveli(c, scale, maxLevel);
}

// PatchOptimizer::optimize(Patch3d&) (PatchOptimizer.cpp:78-103) per patch, one optimizer per thread as main.cpp:123-125.
// status: 0 = optimize() returned true, 100 = false (the reference does not say which stage failed)
void refh_optimize_batch(void* h, int n, orc_patch_t* patches, int nthreads) {
    RefScene* s = static_cast<RefScene*>(h);
    if (nthreads < 1) nthreads = 1;
    std::vector<mo3d::PatchOptimizer> optimizers;
    for (int i = 0; i < nthreads; i++) optimizers.emplace_back(s->options, &s->scene);
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads)
    for (int i = 0; i < n; i++) {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        mo3d::Patch3d p;
        to_patch3d(patches[i], p);
        const bool ok = optimizers[t].optimize(p);
        if (ok) from_patch3d(p, patches[i]);
        patches[i].status = ok ? 0 : 100;
    }
}

// Scene::initPatches up to the tree insertion is not separable; run it whole (Scene.cpp:90-208) and read the tree back.
// Returns the number of patches in patchTree_ (leaf-iterator order), writes up to cap of them.
int refh_init_patches(void* h, orc_patch_t* out, int cap) {
    RefScene* s = static_cast<RefScene*>(h);
    s->scene.initPatches(s->model, s->options);
    int n = 0;
    Leaf_iterator<mo3d::Ppatch3d> end = s->scene.patchTree_.end();
    for (Leaf_iterator<mo3d::Ppatch3d> it = s->scene.patchTree_.begin(); it != end; it++) {
        for (const mo3d::Ppatch3d& p : it->data) {
            if (n < cap) { std::memset(&out[n], 0, sizeof(orc_patch_t)); from_patch3d(*p, out[n]); }
            n++;
        }
    }
    return n;
}

// depth maps: Scene.cpp:74-81 (reset), :351-381 (setDepths)
void refh_depth_reset(void* h) {
    RefScene* s = static_cast<RefScene*>(h);
    for (auto& cam : s->scene.m_depths) for (auto& m : cam) *m = Eigen::MatrixXf::Ones(m->rows(), m->cols()) * mo3d::Scene::MAX_DEPTH;
}
void refh_depth_unset_batch(void* h, int n, const orc_patch_t* patches) {
    RefScene* s = static_cast<RefScene*>(h);
    for (int i = 0; i < n; i++) {
        if (patches[i].status != 0) continue;
        mo3d::Patch3d p;
        to_patch3d(patches[i], p);
        s->scene.setDepths(p, true);
    }
}
void refh_depth_set_batch(void* h, int n, const orc_patch_t* patches) {
    RefScene* s = static_cast<RefScene*>(h);
    for (int i = 0; i < n; i++) {
        if (patches[i].status != 0) continue;
        mo3d::Patch3d p;
        to_patch3d(patches[i], p);
        s->scene.setDepths(p, false);
    }
}
int refh_get_depth(void* h, int cam, int level, float* out, int cap, int* rows, int* cols) {
    RefScene* s = static_cast<RefScene*>(h);
    const Eigen::MatrixXf& m = *s->scene.m_depths[cam][level];
    *rows = (int)m.rows(); *cols = (int)m.cols();
    const int need = *rows * *cols;
    if (!out || cap < need) return need;
    for (int r = 0; r < *rows; r++) for (int c = 0; c < *cols; c++) out[r * *cols + c] = m(r, c);
    return need;
}
// out[3*i+0..2] = depthTests, viewBlockTest, pixelFreeTests (Scene.cpp:518-644)
void refh_accept_batch(void* h, int n, const orc_patch_t* patches, float margin, int32_t* out) {
    RefScene* s = static_cast<RefScene*>(h);
    for (int i = 0; i < n; i++) {
        mo3d::Patch3d p;
        to_patch3d(patches[i], p);
        out[3 * i + 0] = s->scene.depthTests(p, margin);
        out[3 * i + 1] = s->scene.viewBlockTest(p, margin);
        out[3 * i + 2] = s->scene.pixelFreeTests(p);
    }
}

}  // extern "C"
