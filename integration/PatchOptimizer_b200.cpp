// Drop-in replacement translation unit for the reference's src/hpmvs/PatchOptimizer.cpp.
//
// It implements the two members of mo3d::PatchOptimizer that the rest of HPMVS calls -
//     PatchOptimizer(const HpmvsOptions&, const Scene*)      (include/hpmvs/PatchOptimizer.h:41)
//     bool optimize(Patch3d&)                                (include/hpmvs/PatchOptimizer.h:43)
// - on top of the C ABI (include/hpmvs_b200.h) through the C++ shim include/hpmvs_b200_adaptor.hpp.  Nothing else of the
// reference changes: link this file INSTEAD of src/hpmvs/PatchOptimizer.cpp (plus libhpmvs_b200.so) and the reference's
// own CLI, Scene::initPatches and CellProcessor run on the B200 engine.  oracle/Makefile target `dropin` does exactly that
// with the reference's sources where they lie (-> oracle/_ref/hpmvs_ref_b200); tests/test_dropin.py runs both CLIs on the
// same NVM scene and compares the PLY files they write.
//
// The reference calls optimize() one patch at a time from its serial per-subtree queues.  With one host thread that is a batch
// of ONE per call (correct, and as slow as a GPU is on one patch).  With many host threads - the reference runs one queue per
// OpenMP thread over >= --subtrees sub-trees (src/main.cpp:128-155) and seeds in an OpenMP loop (src/hpmvs/Scene.cpp:114) - the
// calls that arrive while the GPU is busy are COALESCED into one hpmvs_optimize_batch(): run the unmodified scheduler with a few
// hundred (oversubscribed, mostly blocked) threads, e.g. OMP_NUM_THREADS=512 hpmvs --subtrees=2000, and the engine sees batches of
// a few hundred patches.  The batched call sites a maintainer would write instead are described in INTEGRATION.md section 3.
#include <cmath>
#include <string>
#include <hpmvs/Scene.h>
#include <hpmvs/PatchOptimizer.h>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "hpmvs_b200_adaptor.hpp"

namespace {

// Request coalescing: host threads park their record and wait; one worker thread sends everything that has arrived as ONE batch.
class Coalescer {
public:
    explicit Coalescer(hpmvs_engine_t* e) : engine_(e), worker_([this] { run(); }) {}
    ~Coalescer() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_submit_.notify_all();
        worker_.join();
        if (getenv("HPMVS_DROPIN_STATS"))
            fprintf(stderr, "[hpmvs_b200 drop-in] %llu optimize() calls in %llu batches (largest %d)\n", calls_, batches_, largest_);
    }
    void optimize(hpmvs_patch_t& rec) {
        Request r{&rec, false};
        std::unique_lock<std::mutex> lk(mu_);
        pending_.push_back(&r);
        cv_submit_.notify_one();
        cv_done_.wait(lk, [&] { return r.done; });
    }
private:
    struct Request { hpmvs_patch_t* rec; bool done; };
    void run() {
        std::vector<Request*> batch;
        std::vector<hpmvs_patch_t> buf;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_submit_.wait(lk, [&] { return stop_ || !pending_.empty(); });
                if (pending_.empty()) return;      // stop requested and nothing left to do
                batch.swap(pending_);
            }
            buf.resize(batch.size());
            for (size_t i = 0; i < batch.size(); i++) buf[i] = *batch[i]->rec;
            hpmvs_b200::check(hpmvs_optimize_batch(engine_, (int)buf.size(), buf.data(), buf.data(), nullptr), "hpmvs_optimize_batch");
            for (size_t i = 0; i < batch.size(); i++) *batch[i]->rec = buf[i];
            {
                std::lock_guard<std::mutex> lk(mu_);
                for (Request* r : batch) r->done = true;
                calls_ += batch.size(); batches_++; if ((int)batch.size() > largest_) largest_ = (int)batch.size();
            }
            cv_done_.notify_all();
            batch.clear();
        }
    }
    hpmvs_engine_t* engine_;
    std::mutex mu_;
    std::condition_variable cv_submit_, cv_done_;
    std::vector<Request*> pending_;
    bool stop_ = false;
    unsigned long long calls_ = 0, batches_ = 0;
    int largest_ = 0;
    std::thread worker_;
};

struct Gpu {
    std::shared_ptr<hpmvs_b200::PatchOptimizer> shim;
    std::unique_ptr<Coalescer> coalescer;     // destroyed first (declared last)
};

// One engine per Scene, shared by all PatchOptimizer instances (the reference creates one per OpenMP thread,
// src/main.cpp:123-125, src/hpmvs/Scene.cpp:94-97).
std::mutex g_mu;
std::map<const mo3d::Scene*, std::shared_ptr<Gpu>> g_engines;

std::shared_ptr<Gpu> engine_for(const mo3d::HpmvsOptions& options, const mo3d::Scene* scene) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_engines.find(scene);
    if (it != g_engines.end()) return it->second;
    auto holder = std::make_shared<Gpu>();
    holder->shim = std::make_shared<hpmvs_b200::PatchOptimizer>(options, /*device*/ 0);
    hpmvs_b200::PatchOptimizer* gpu = holder->shim.get();
    // ---- scene upload: what PatchOptimizer.cpp:38-41 borrows as raw pointers (cameras_, images_, covis_) -------------
    const int ncams = (int)scene->cameras_.size();
    std::vector<hpmvs_camera_t> cams(ncams);
    for (int i = 0; i < ncams; i++) {
        const mo3d::Camera& c = scene->cameras_[i];                       // include/hpmvs/Camera.h:87-106
        hpmvs_camera_t& o = cams[i];
        std::memset(&o, 0, sizeof(o));
        for (int l = 0; l < c.getLevels() && l < HPMVS_LEVELS; l++) {
            for (int r = 0; r < 3; r++) for (int k = 0; k < 4; k++) o.P[l][r][k] = c.projection_[l](r, k);
            o.width[l] = scene->images_[i].getWidth(l);                   // include/hpmvs/Image.h:62-63
            o.height[l] = scene->images_[i].getHeight(l);
        }
        for (int k = 0; k < 4; k++) o.center[k] = c.center_[k];
        for (int k = 0; k < 3; k++) { o.xaxis[k] = c.xAxis_[k]; o.yaxis[k] = c.yAxis_[k]; o.zaxis[k] = c.zAxis_[k]; }
        o.k00 = c.kMat_[0](0, 0); o.k11 = c.kMat_[0](1, 1);
    }
    hpmvs_b200::check(hpmvs_engine_set_cameras(gpu->engine(), ncams, cams.data()), "hpmvs_engine_set_cameras");
    std::vector<unsigned char> rgb;
    for (int i = 0; i < ncams; i++)
        for (int l = 0; l < scene->cameras_[i].getLevels() && l < HPMVS_LEVELS; l++) {
            // Image::getImage un-permutes the interleaved storage (Image.h:65, Image.cpp:62-63) to planar x,y,1,c
            const cimg_library::CImg<unsigned char> img = scene->images_[i].getImage(l);
            const int w = img.width(), h = img.height();
            rgb.resize((size_t)w * h * 3);
            for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) for (int ch = 0; ch < 3; ch++) rgb[3 * ((size_t)y * w + x) + ch] = img(x, y, 0, ch);
            hpmvs_b200::check(hpmvs_engine_upload_image(gpu->engine(), i, l, rgb.data(), w, h, (size_t)3 * w), "hpmvs_engine_upload_image");
        }
    std::vector<int32_t> off{0}, ids;
    for (const std::vector<int>& l : scene->covis_) { ids.insert(ids.end(), l.begin(), l.end()); off.push_back((int32_t)ids.size()); }
    if (ids.empty()) ids.push_back(0);
    hpmvs_b200::check(hpmvs_engine_set_covis(gpu->engine(), off.data(), ids.data()), "hpmvs_engine_set_covis");
    holder->coalescer.reset(new Coalescer(gpu->engine()));
    g_engines[scene] = holder;
    return holder;
}

}  // namespace

namespace mo3d {

PatchOptimizer::PatchOptimizer(const mo3d::HpmvsOptions& options, const mo3d::Scene* scene) : options_p(&options), scene_p(scene) {
    camera_p = scene->cameras_.data();
    images_p = scene->images_.data();
    covis_p = &scene->covis_;
    engine_for(options, scene);      // create + upload on first use
}

bool PatchOptimizer::optimize(mo3d::Patch3d& patch) {
    hpmvs_patch_t rec;
    hpmvs_b200::to_record(patch, rec);
    engine_for(*options_p, scene_p)->coalescer->optimize(rec);
    return hpmvs_b200::from_record(rec, patch);          // fields are written back only on success (PatchOptimizer.cpp:86-100)
}

}  // namespace mo3d
