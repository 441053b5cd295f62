// Header-only C++ adaptor over the C ABI (include/hpmvs_b200.h) with the reference's call shape:
//     mo3d::PatchOptimizer(const HpmvsOptions&, const Scene*) ; bool optimize(Patch3d&)
//     (/root/reference/include/hpmvs/PatchOptimizer.h:39-43, src/hpmvs/PatchOptimizer.cpp:38-45,78-103)
// It is a template over the patch / camera / options types so that it compiles against the reference's own
// Eigen-based mo3d::Patch3d (fields center_, normal_, scale_3dx_, images_, ncc_, color_; Patch3d.h:55-82) without
// this repository depending on Eigen.  See INTEGRATION.md for the three call sites it replaces.
#pragma once

#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "hpmvs_b200.h"

namespace hpmvs_b200 {

struct EngineDeleter { void operator()(hpmvs_engine_t* e) const { hpmvs_engine_destroy(e); } };

inline void check(int rc, const char* what) {
    if (rc < 0) throw std::runtime_error(std::string(what) + ": " + hpmvs_error_string(rc));
}

// Options: any struct with the HpmvsOptions field names (HpmvsOptions.h:31-52)
template <class Options>
hpmvs_options_t make_options(const Options& o) {
    hpmvs_options_t r;
    r.maxlevel = o.MAXLEVEL; r.minlevel = o.MINLEVEL; r.start_level = o.START_LEVEL;
    r.max_angle = o.MAX_ANGLE; r.min_angle = o.MIN_ANGLE;
    r.max_images_per_patch = o.MAX_IMAGES_PER_PATCH; r.min_images_per_patch = o.MIN_IMAGES_PER_PATCH;
    r.ncc_alpha_1 = o.NCC_ALPHA_1; r.ncc_alpha_2 = o.NCC_ALPHA_2;
    return r;
}

template <class Patch>
void to_record(const Patch& p, hpmvs_patch_t& r) {
    std::memset(&r, 0, sizeof(r));
    for (int i = 0; i < 4; i++) { r.center[i] = p.center_[i]; r.normal[i] = p.normal_[i]; }
    r.scale = p.scale_3dx_;
    r.nimages = (int32_t)p.images_.size();
    if (r.nimages > HPMVS_MAX_VIEWS) throw std::length_error("patch has more views than HPMVS_MAX_VIEWS");
    for (int i = 0; i < r.nimages; i++) r.images[i] = p.images_[i];
}

// mirrors PatchOptimizer.cpp:86-100: fields are written back only on success
template <class Patch>
bool from_record(const hpmvs_patch_t& r, Patch& p) {
    if (r.status != HPMVS_OK) return false;
    for (int i = 0; i < 4; i++) { p.center_[i] = r.center[i]; p.normal_[i] = r.normal[i]; }
    p.scale_3dx_ = r.scale;
    p.images_.assign(r.images, r.images + r.nimages);
    p.ncc_ = r.ncc;
    for (int i = 0; i < 3; i++) p.color_[i] = r.color[i];
    return true;
}

class PatchOptimizer {
public:
    // `cams` / pyramids / covisibility are uploaded by the caller through engine() (hpmvs_engine_set_cameras, ...)
    template <class Options>
    explicit PatchOptimizer(const Options& options, int device = 0) {
        const hpmvs_options_t o = make_options(options);
        hpmvs_engine_t* e = nullptr;
        check(hpmvs_engine_create(&o, device, &e), "hpmvs_engine_create");
        engine_.reset(e);
        // drop-in fidelity: the start angles are evaluated with the host's libm, as the reference does (PatchOptimizer.cpp:427-437)
        check(hpmvs_engine_set_start_mode(e, 1), "hpmvs_engine_set_start_mode");
    }
    hpmvs_engine_t* engine() const { return engine_.get(); }

    // drop-in for `bool PatchOptimizer::optimize(Patch3d&)` (a batch of one; prefer optimizeBatch)
    template <class Patch>
    bool optimize(Patch& patch) {
        hpmvs_patch_t rec;
        to_record(patch, rec);
        check(hpmvs_optimize_batch(engine_.get(), 1, &rec, &rec, nullptr), "hpmvs_optimize_batch");
        return from_record(rec, patch);
    }

    // the batched form the GPU wants: optimises every patch, returns per-patch success like n optimize() calls
    template <class PatchPtr>
    std::vector<char> optimizeBatch(std::vector<PatchPtr>& patches) {
        buf_.resize(patches.size());
        for (size_t i = 0; i < patches.size(); i++) to_record(*patches[i], buf_[i]);
        check(hpmvs_optimize_batch(engine_.get(), (int)buf_.size(), buf_.data(), buf_.data(), nullptr), "hpmvs_optimize_batch");
        std::vector<char> ok(patches.size());
        for (size_t i = 0; i < patches.size(); i++) ok[i] = from_record(buf_[i], *patches[i]) ? 1 : 0;
        return ok;
    }

private:
    std::unique_ptr<hpmvs_engine_t, EngineDeleter> engine_;
    std::vector<hpmvs_patch_t> buf_;
};

}  // namespace hpmvs_b200
