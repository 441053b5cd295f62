// Compile-and-link check of include/hpmvs_b200_adaptor.hpp against a stand-in for mo3d::Patch3d / HpmvsOptions
// (same field names as /root/reference/include/hpmvs/Patch3d.h:55-82 and HpmvsOptions.h:31-52, without Eigen).
#include <array>
#include <cstdio>
#include <memory>
#include <vector>

#include "../../include/hpmvs_b200_adaptor.hpp"

struct Patch3d {
    std::array<float, 4> center_{}, normal_{};
    std::vector<int> images_;
    float scale_3dx_ = 0, dscale_ = 0, ncc_ = 0;
    std::array<float, 3> color_{};
};
struct HpmvsOptions {
    int MAXLEVEL = 5, MINLEVEL = 0, START_LEVEL = 4;
    float MAX_ANGLE = 60.0f * 3.14159265358979323846 / 180.0f, MIN_ANGLE = 10.0f * 3.14159265358979323846 / 180.0f;
    int MAX_IMAGES_PER_PATCH = 6, MIN_IMAGES_PER_PATCH = 3;
    float NCC_ALPHA_1 = 0.4f, NCC_ALPHA_2 = 0.5f;
};

extern "C" int adaptor_selftest() {
    // record round trip (no GPU needed)
    Patch3d p; p.center_ = {1, 2, 3, 1}; p.normal_ = {0, 0, -1, 0}; p.scale_3dx_ = 0.1f; p.images_ = {3, 1, 2};
    hpmvs_patch_t r; hpmvs_b200::to_record(p, r);
    if (r.nimages != 3 || r.images[1] != 1 || r.center[2] != 3.0f) return 1;
    r.status = HPMVS_FAIL_NCC1;
    Patch3d q = p; q.scale_3dx_ = 7;
    if (hpmvs_b200::from_record(r, q) || q.scale_3dx_ != 7) return 2;      // untouched on failure
    r.status = HPMVS_OK; r.ncc = 1.4f; r.nimages = 2;
    if (!hpmvs_b200::from_record(r, q) || q.images_.size() != 2 || q.ncc_ != 1.4f) return 3;
    // constructing the optimizer needs a device; without one it must throw, not fall back
    try {
        hpmvs_b200::PatchOptimizer opt{HpmvsOptions{}};
        std::vector<std::shared_ptr<Patch3d>> v;
        opt.optimizeBatch(v);
        (void)opt;
        return 0;
    } catch (const std::exception& e) {
        return std::string(e.what()).find("no CUDA device") != std::string::npos ? 0 : 4;
    }
}
