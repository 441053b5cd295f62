// Level-synchronous host driver of the expand -> optimize -> filter loop around the engine, in C++ behind the C ABI
// (hpmvs_pipeline_run, include/hpmvs_b200.h).  It is the batching stand-in for the reference's scheduler
// (CellProcessor + DynOctTree, /root/reference/src/hpmvs/CellProcessor.cpp:369-420, src/main.cpp:145-155): per tree level it
// collects the candidates of ALL cells, optimises them in one engine batch, runs the acceptance tests against a snapshot of the
// depth maps and commits the survivors in a deterministic order.  Kept from the reference, per cell:
//   seeds   reject when the optimised centre moved more than 2*scale (Scene.cpp:171); one patch per cell, the better supported
//           patch wins (CellProcessor::filter, CellProcessor.cpp:43-82)
//   extend  6 candidates one cell width away, scale = width*0.9/2 (CellProcessor.cpp:98-119); accepted when width/2 < 2*scale < width,
//           drift < 1.5*width, depthTests >= MIN_IMAGES, viewBlockTest < MIN_IMAGES, pixelFreeTests >= MIN_IMAGES-1 and > 75 % of the
//           views (:129-142); a patch that leaves the root cube is dropped (:147-153, :533-540)
//   branch  4 candidates at width/4, scale = width*0.45/2, kept when they stay inside the parent's cell (:227-264); a cell that
//           yields no child keeps its patch from PATCH_FINAL_MINLEVEL on, below that it loses it (:266-283)
// hpmvs_b200/pipeline.py is the same algorithm in numpy against a backend protocol (it also runs on the CPU oracle);
// tests/test_pipeline.py requires identical patch arrays from both.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <chrono>
#include <algorithm>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/hpmvs_b200.h"

namespace {

typedef std::vector<hpmvs_patch_t> Recs;
const int MIN_IMAGES = 3;
const float DEPTH_TEST_FACTOR = 1.0f;

// f32 helpers in the evaluation order of the device code (patch_kernels.cuh): Vector4f reductions as (p0 + p2) + (p1 + p3)
inline float h_dot4(const float* a, const float* b) {
    const float p0 = a[0] * b[0], p1 = a[1] * b[1], p2 = a[2] * b[2], p3 = a[3] * b[3];
    return (p0 + p2) + (p1 + p3);
}
// std::round(Camera::getLevel(center, scale)) (src/hpmvs/Camera.cpp:92-95) on the host, as hp::level_from computes it on the device
inline int host_level(const hpmvs_camera_t& cam, const float* center, float scale) {
    float d[4];
    for (int i = 0; i < 4; i++) d[i] = center[i] - cam.center[i];
    const float fz = std::sqrt(h_dot4(d, d));
    const float ksum = cam.k00 + cam.k11;
    const float lvl = (float)std::log2((double)(scale * ksum) / (2.0 * (double)fz));
    return (int)std::round(lvl);
}

struct Driver {
    hpmvs_engine_t* e;
    const hpmvs_pipeline_params_t* p;
    hpmvs_pipeline_stats_t st;
    double origin[3];
    std::unordered_map<int64_t, int> sub_rank;      // (level, cell key) -> rank, from the sub-tree table
    std::vector<int> sub_levels;
    bool trace = getenv("HPMVS_PIPELINE_TRACE") != nullptr;     // one stderr line per engine batch (debugging / profiling)

    double width(int level) const { return p->root_width / (double)(1 << level); }
    bool exchanging() const { return p->exchange != nullptr && p->shard_count > 1; }

    int64_t key_of(const float* c, double w) const {
        int64_t k[3];
        for (int i = 0; i < 3; i++) k[i] = (int64_t)std::floor(((double)c[i] - origin[i]) / w) + ((int64_t)1 << 20);
        return (k[0] << 42) | (k[1] << 21) | k[2];
    }

    void init_shards() {
        for (int i = 0; i < p->nsub; i++) {
            const int L = p->sub_level[i];
            const int64_t off = (int64_t)1 << 20;
            const int64_t k = ((p->sub_key[3 * i] + off) << 42) | ((p->sub_key[3 * i + 1] + off) << 21) | (p->sub_key[3 * i + 2] + off);
            sub_rank[k * 32 + L] = p->sub_rank[i];
            if (std::find(sub_levels.begin(), sub_levels.end(), L) == sub_levels.end()) sub_levels.push_back(L);
        }
    }

    // multi-GPU: which rank grows the cell that holds `c`.  With a sub-tree table (hpmvs_shard_cells = the reference's getSubTrees split,
    // src/main.cpp:50-96) a centre belongs to the rank of the sub-tree that contains it and to NOBODY where the split left no sub-tree
    // (the reference drops a patch that no CellProcessor's tree contains, CellProcessor.cpp:533-540); without a table the cells of tree
    // level shard_level are dealt out by a hash (round-1 behaviour, kept for the single-process tests)
    int owner(const float* c) const {
        if (p->shard_count <= 1) return 0;
        if (p->nsub > 0) {
            for (int L : sub_levels) {
                auto it = sub_rank.find(key_of(c, width(L)) * 32 + L);
                if (it != sub_rank.end()) return it->second;
            }
            return -1;
        }
        const double w = width(p->shard_level);
        int64_t k[3];
        for (int i = 0; i < 3; i++) k[i] = (int64_t)std::floor(((double)c[i] - origin[i]) / w);
        const int64_t cell = (k[0] * 73856093ll) ^ (k[1] * 19349663ll) ^ (k[2] * 83492791ll);
        return (int)(((cell % p->shard_count) + p->shard_count) % p->shard_count);
    }
    bool mine(const float* c) const { return p->shard_count <= 1 || owner(c) == p->shard_rank; }

    int optimize(Recs& r) {
        if (r.empty()) return 0;
        const auto t0 = std::chrono::steady_clock::now();
        const int rc = hpmvs_optimize_batch(e, (int)r.size(), r.data(), r.data(), nullptr);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        st.seconds_optimize += dt;
        if (trace) fprintf(stderr, "[hpmvs pipeline] optimize batch of %7d patches: %8.2f ms\n", (int)r.size(), 1e3 * dt);
        st.optimize_calls += (int64_t)r.size();
        for (const hpmvs_patch_t& q : r) st.optimized_ok += (q.status == HPMVS_OK);
        return rc;
    }

    // all ranks' records of this step, rank by rank (the per-round border hand-off of the reference, CellProcessor.cpp:147-153 +
    // distributeBorderCell :487-540, as one variable-length all-gather); identity without an exchange callback
    int exchange(const Recs& mine_recs, Recs& all) {
        if (!exchanging()) { all = mine_recs; return 0; }
        hpmvs_patch_t* recv = nullptr;
        int nrecv = 0;
        const int rc = p->exchange(p->exchange_user, (int)mine_recs.size(), mine_recs.data(), &recv, &nrecv);
        if (rc < 0) return rc;
        all.assign(recv, recv + nrecv);
        st.exchanged += nrecv;
        return 0;
    }

    // CellProcessor::filter (CellProcessor.cpp:43-82): of the patches that share a cell keep the one with the smallest mean signed
    // distance of the OTHERS' centres along its own normal (first minimum in arrival order)
    static int filter_pick(const std::vector<const hpmvs_patch_t*>& m) {
        int best = 0;
        float bestd = 3.402823466e+38f;
        for (size_t a = 0; a < m.size(); a++) {
            float n[3] = {m[a]->normal[0], m[a]->normal[1], m[a]->normal[2]};
            const float z = (n[0] * n[0] + n[1] * n[1]) + n[2] * n[2];
            if (z > 0.0f) { const float s = std::sqrt(z); n[0] /= s; n[1] /= s; n[2] /= s; }
            float dist = 0.0f;
            for (size_t b = 0; b < m.size(); b++) {
                if (a == b) continue;
                const float d0 = m[b]->center[0] - m[a]->center[0], d1 = m[b]->center[1] - m[a]->center[1], d2 = m[b]->center[2] - m[a]->center[2];
                dist += (n[0] * d0 + n[1] * d1) + n[2] * d2;
            }
            dist /= (float)(m.size() - 1);
            if (dist < bestd) { bestd = dist; best = (int)a; }
        }
        return best;
    }

    // one patch per cell.  `cells` keeps insertion order; live[i] = rec[i] now lives in the grid; a patch that loses its cell to a new
    // one is appended to `removed` (its depths are subtracted by the caller, Scene::setDepths(p, true), CellProcessor.cpp:73-79)
    void insert(Recs& cells, std::unordered_map<int64_t, int>& index, const Recs& rec, double w, std::vector<char>* live, Recs* removed) {
        if (live) live->assign(rec.size(), 0);
        std::unordered_map<int64_t, std::vector<int>> groups;
        std::vector<int64_t> order;
        for (size_t i = 0; i < rec.size(); i++) {
            const int64_t k = key_of(rec[i].center, w);
            auto it = groups.find(k);
            if (it == groups.end()) { groups[k] = std::vector<int>(1, (int)i); order.push_back(k); }
            else it->second.push_back((int)i);
        }
        for (const int64_t k : order) {
            const std::vector<int>& g = groups[k];
            auto it = index.find(k);
            std::vector<const hpmvs_patch_t*> members;
            if (it != index.end()) members.push_back(&cells[it->second]);
            for (int i : g) members.push_back(&rec[i]);
            const int b = members.size() > 1 ? filter_pick(members) : 0;
            if (it != index.end()) {
                if (b == 0) continue;                                    // the resident patch keeps its cell
                if (removed) removed->push_back(cells[it->second]);
                cells[it->second] = *members[b];
                if (live) (*live)[g[b - 1]] = 1;
            } else {
                index[k] = (int)cells.size();
                cells.push_back(*members[b]);
                if (live) (*live)[g[b]] = 1;
            }
        }
    }

    static float norm3f(const float* a, const float* b) {
        const float d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
        return std::sqrt((d0 * d0 + d1 * d1) + d2 * d2);
    }

    // Scene::getLevelSupport(patch, MINLEVEL) (Scene.cpp:335-344)
    int level_support(const hpmvs_patch_t& q) const {
        int n = 0;
        for (int k = 0; k < q.nimages && k < HPMVS_MAX_VIEWS; k++)
            if (host_level(p->cams[q.images[k]], q.center, q.scale) > p->minlevel) n++;
        return n;
    }

    // the accepted candidates of one round thinned to the first one per image cell of their reference view (see pipeline.py)
    void first_per_ref_pixel(const Recs& rec, double w, std::vector<char>& keep) const {
        keep.assign(rec.size(), 1);
        if (!p->dedup_ref_pixel) return;
        std::unordered_set<int64_t> seen;
        for (size_t i = 0; i < rec.size(); i++) {
            const int ref = rec[i].images[0];
            const hpmvs_camera_t& cam = p->cams[ref];
            double r[3];
            for (int a = 0; a < 3; a++)
                r[a] = (((double)cam.P[0][a][0] * (double)rec[i].center[0] + (double)cam.P[0][a][1] * (double)rec[i].center[1]) +
                        (double)cam.P[0][a][2] * (double)rec[i].center[2]) + (double)cam.P[0][a][3] * (double)rec[i].center[3];
            const double z = r[2] > 1e-9 ? r[2] : 1e-9;
            double cell_px = w * (double)cam.k00 / z;
            if (!(cell_px > 1e-6)) cell_px = 1e-6;
            const int64_t ku = (int64_t)std::floor(r[0] / z / cell_px), kv = (int64_t)std::floor(r[1] / z / cell_px);
            const int64_t key = ((int64_t)ref << 44) | ((ku + ((int64_t)1 << 20)) << 22) | (kv + ((int64_t)1 << 20));
            if (!seen.insert(key).second) keep[i] = 0;
        }
    }

    // commit a step's accepted records: all ranks learn all of them (depth maps stay identical on every rank), each rank keeps the ones
    // in its own cells.  Returns the records that now live in this rank's grid.
    int commit(Recs& cells, std::unordered_map<int64_t, int>& index, const Recs& accepted, double w, Recs& fresh, bool* any) {
        Recs all, removed;
        int rc = exchange(accepted, all);
        if (rc < 0) return rc;
        if (any) *any = !all.empty();
        Recs own;
        for (const hpmvs_patch_t& q : all) if (mine(q.center)) own.push_back(q);
        std::vector<char> live;
        insert(cells, index, own, w, &live, &removed);
        fresh.clear();
        for (size_t i = 0; i < own.size(); i++) if (live[i]) fresh.push_back(own[i]);
        // depth bookkeeping on EVERY rank: the losers of a cell give their depths back, the new residents set theirs
        Recs all_removed, all_fresh;
        if ((rc = exchange(removed, all_removed)) < 0) return rc;
        if (exchanging()) { if ((rc = exchange(fresh, all_fresh)) < 0) return rc; }
        else all_fresh = fresh;
        if (!all_removed.empty() && (rc = hpmvs_depth_unset_batch(e, (int)all_removed.size(), all_removed.data(), nullptr)) < 0) return rc;
        if (!all_fresh.empty() && (rc = hpmvs_depth_set_batch(e, (int)all_fresh.size(), all_fresh.data(), nullptr)) < 0) return rc;
        return 0;
    }

    int run(const Recs& seeds, Recs& final_out) {
        int rc;
        init_shards();
        if ((rc = hpmvs_engine_depth_reset(e)) < 0) return rc;
        Recs out = seeds;
        if ((rc = optimize(out)) < 0) return rc;
        Recs first;
        for (size_t i = 0; i < out.size(); i++) {
            if (out[i].status != HPMVS_OK) continue;
            if (norm3f(out[i].center, seeds[i].center) > out[i].scale * 2) continue;             // Scene.cpp:171
            if (!exchanging() && !mine(out[i].center)) continue;
            first.push_back(out[i]);
        }
        Recs cells, fresh;
        std::unordered_map<int64_t, int> index;
        int level = p->start_level;
        if ((rc = commit(cells, index, first, width(level), fresh, nullptr)) < 0) return rc;
        std::vector<char> keep;
        std::vector<float> widths;
        std::vector<int32_t> counts;
        for (;;) {
            const double w = width(level);
            Recs frontier = cells;
            int64_t n_ext = 0;
            for (int round = 0; round < p->max_rounds && (exchanging() || !frontier.empty()); round++) {   // extend until the wavefront dies out
                Recs cand(frontier.size() * 6);
                widths.assign(frontier.size(), (float)w);
                if (!frontier.empty() &&
                    (rc = hpmvs_expand_candidates(p->ncams, p->cams, (int)frontier.size(), frontier.data(), widths.data(), 6, cand.data())) < 0) return rc;
                // one candidate per free cell and round (first parent wins), like a cell being filled once
                std::unordered_set<int64_t> seen;
                Recs sel; std::vector<int> parent;
                for (size_t i = 0; i < cand.size(); i++) {
                    const int64_t k = key_of(cand[i].center, w);
                    const bool uniq = seen.insert(k).second;
                    if (uniq && index.find(k) == index.end()) { sel.push_back(cand[i]); parent.push_back((int)(i / 6)); }
                }
                if (sel.empty() && !exchanging()) break;
                if ((rc = optimize(sel)) < 0) return rc;
                counts.assign(sel.size() * 3, 0);
                const auto t0 = std::chrono::steady_clock::now();
                if (!sel.empty() && (rc = hpmvs_accept_batch(e, (int)sel.size(), sel.data(), DEPTH_TEST_FACTOR, counts.data(), nullptr)) < 0) return rc;
                st.seconds_accept += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                Recs acc;
                for (size_t i = 0; i < sel.size(); i++) {
                    const hpmvs_patch_t& r = sel[i];
                    bool good = r.status == HPMVS_OK;
                    good = good && (r.scale * 2.0f < (float)w) && (r.scale * 2.0f > (float)(w / 2.0));
                    good = good && norm3f(r.center, frontier[parent[i]].center) < (float)(w * 1.5);
                    for (int a = 0; a < 3 && good; a++) {
                        const double rel = ((double)r.center[a] - origin[a]) / p->root_width;
                        good = rel >= 0.0 && rel < 1.0;
                    }
                    // without an exchange a patch that leaves this rank's cells is lost (nobody to hand it to); with one its owner gets it
                    good = good && (exchanging() ? owner(r.center) >= 0 : mine(r.center));
                    const int nimg = r.nimages > 1 ? r.nimages : 1;
                    good = good && counts[3 * i] >= MIN_IMAGES && counts[3 * i + 1] < MIN_IMAGES;
                    good = good && counts[3 * i + 2] >= MIN_IMAGES - 1 && (counts[3 * i + 2] * 1.0 / nimg > 0.75);
                    if (good) acc.push_back(r);
                }
                first_per_ref_pixel(acc, w, keep);
                Recs acc2;
                for (size_t i = 0; i < acc.size(); i++) if (keep[i]) acc2.push_back(acc[i]);
                if (acc2.empty() && !exchanging()) break;
                bool any = false;
                if ((rc = commit(cells, index, acc2, w, fresh, &any)) < 0) return rc;
                if (!any) break;                                                        // no rank accepted anything: the level is done
                n_ext += (int64_t)fresh.size();
                frontier.swap(fresh);
            }
            const Recs& patches = cells;
            if (st.nlevels < HPMVS_PIPELINE_MAX_LEVELS) { st.level[st.nlevels] = level; st.extended[st.nlevels] = n_ext; st.branched[st.nlevels] = 0; }
            if (level >= p->final_level || (patches.empty() && !exchanging())) {
                final_out.insert(final_out.end(), patches.begin(), patches.end());
                if (st.nlevels < HPMVS_PIPELINE_MAX_LEVELS) st.nlevels++;
                break;
            }
            // branch into the next level (CellProcessor::branch, CellProcessor.cpp:210-307).  A patch without level support (every view
            // already at the finest pyramid level, :222-225) is exhausted: its cell keeps it and it is a final result at whatever level.
            std::vector<char> exhausted(patches.size(), 0);
            Recs parents; std::vector<int> parent_of;
            for (size_t i = 0; i < patches.size(); i++) {
                if (level_support(patches[i]) < 1) { exhausted[i] = 1; final_out.push_back(patches[i]); }
                else { parents.push_back(patches[i]); parent_of.push_back((int)i); }
            }
            // 4 candidates per patch, kept when they stay inside the parent's cell
            Recs cand4(parents.size() * 4);
            widths.assign(parents.size(), (float)w);
            if (!parents.empty() &&
                (rc = hpmvs_expand_candidates(p->ncams, p->cams, (int)parents.size(), parents.data(), widths.data(), 4, cand4.data())) < 0) return rc;
            Recs cand; std::vector<int> parent; std::vector<int64_t> pkeys;
            for (size_t i = 0; i < cand4.size(); i++) {
                const int64_t pk = key_of(parents[i / 4].center, w);
                if (key_of(cand4[i].center, w) == pk) { cand.push_back(cand4[i]); parent.push_back(parent_of[i / 4]); pkeys.push_back(pk); }
            }
            if ((rc = optimize(cand)) < 0) return rc;
            Recs children;
            std::vector<char> branched(patches.size(), 0);
            for (size_t i = 0; i < cand.size(); i++)
                if (cand[i].status == HPMVS_OK && key_of(cand[i].center, w) == pkeys[i]) { children.push_back(cand[i]); branched[parent[i]] = 1; }
            // a cell that yields no child keeps its patch from PATCH_FINAL_MINLEVEL on (:266-269); every other cell is split and its
            // patch leaves the tree: Scene::setDepths(old, true) (:271-279)
            Recs gone;
            for (size_t i = 0; i < patches.size(); i++) {
                if (exhausted[i]) continue;
                if (!branched[i] && level >= p->final_min_level) final_out.push_back(patches[i]);
                else gone.push_back(patches[i]);
            }
            Recs all_gone;
            if ((rc = exchange(gone, all_gone)) < 0) return rc;
            if (!all_gone.empty() && (rc = hpmvs_depth_unset_batch(e, (int)all_gone.size(), all_gone.data(), nullptr)) < 0) return rc;
            if (st.nlevels < HPMVS_PIPELINE_MAX_LEVELS) { st.branched[st.nlevels] = (int64_t)children.size(); st.nlevels++; }
            Recs next; std::unordered_map<int64_t, int> nindex;
            level += 1;
            cells.swap(next); index.swap(nindex);
            if ((rc = commit(cells, index, children, width(level), fresh, nullptr)) < 0) return rc;
        }
        return 0;
    }
};

}  // namespace

extern "C" {

int hpmvs_pipeline_run(hpmvs_engine_t* e, const hpmvs_pipeline_params_t* params, int nseeds, const hpmvs_patch_t* seeds,
                       hpmvs_patch_t** out, int* nout, hpmvs_pipeline_stats_t* stats) {
    if (!e || !params || !out || !nout || nseeds < 0 || (nseeds > 0 && !seeds) || !params->cams || params->ncams <= 0 ||
        params->start_level < 0 || params->final_level < params->start_level || params->final_level > 20 || !(params->root_width > 0.0) ||
        (params->shard_count > 1 && (params->shard_rank < 0 || params->shard_rank >= params->shard_count || params->shard_level < 0 ||
                                     params->shard_level > params->start_level)))
        return HPMVS_E_ARG;
    Driver d;
    d.e = e; d.p = params;
    std::memset(&d.st, 0, sizeof(d.st));
    for (int i = 0; i < 3; i++) d.origin[i] = params->origin[i];
    Recs s(seeds, seeds + nseeds), fin;
    const int rc = d.run(s, fin);
    if (stats) *stats = d.st;
    if (rc < 0) return rc;
    *nout = (int)fin.size();
    *out = nullptr;
    if (!fin.empty()) {
        *out = (hpmvs_patch_t*)std::malloc(sizeof(hpmvs_patch_t) * fin.size());
        if (!*out) return HPMVS_E_ARG;
        std::memcpy(*out, fin.data(), sizeof(hpmvs_patch_t) * fin.size());
    }
    return 0;
}

void hpmvs_free(void* p) { std::free(p); }

// Root cube of the patch octree as Scene::initPatches forms it (src/hpmvs/Scene.cpp:186-193, getBoundingBox :329-349): f32 bounding box
// of the patch centres, edge = the largest extent, centred on the box.  origin = low corner.
int hpmvs_root_cube(int n, const hpmvs_patch_t* patches, double origin[3], double* width) {
    if (n <= 0 || !patches || !origin || !width) return HPMVS_E_ARG;
    float mn[3], mx[3];
    for (int a = 0; a < 3; a++) { mn[a] = patches[0].center[a]; mx[a] = patches[0].center[a]; }
    for (int i = 1; i < n; i++)
        for (int a = 0; a < 3; a++) { mn[a] = std::min(mn[a], patches[i].center[a]); mx[a] = std::max(mx[a], patches[i].center[a]); }
    const float w = std::max(mx[0] - mn[0], std::max(mx[1] - mn[1], mx[2] - mn[2]));
    for (int a = 0; a < 3; a++) origin[a] = (double)((mn[a] + mx[a]) / 2.0f) - (double)w / 2.0;
    *width = (double)w;
    return 0;
}

// Multi-GPU partition of a patch set by octree sub-tree, the reference's own split (getSubTrees, src/main.cpp:50-96, on top of
// DynOctTree::getSubTrees, include/hpmvs/doctree.h:513-523): the root cube is split into its (non-empty) children, then the sub-tree
// holding the most patches is split again until there are at least `min_subtrees` of them or the biggest holds fewer than 100
// (main.cpp:74).  The sub-trees are dealt to `nranks` ranks greedily: costliest first, each to the least loaded rank (the reference
// lets OpenMP's dynamic schedule do that, main.cpp:150).
struct SubTree { int level; uint32_t k[3]; std::vector<int> pts; int rank; };
static int build_subtrees(int n, const hpmvs_patch_t* patches, const double origin[3], double root_width, int min_subtrees, int nranks,
                          std::vector<SubTree>& subs) {
    const int MAXL = 20;
    std::vector<uint32_t> q((size_t)n * 3);        // integer cell coordinates at level MAXL
    SubTree root; root.level = 0; root.k[0] = root.k[1] = root.k[2] = 0; root.rank = 0;
    for (int i = 0; i < n; i++) {
        bool inside = true;
        for (int a = 0; a < 3; a++) {
            double rel = ((double)patches[i].center[a] - origin[a]) / root_width;
            // the box is closed at the top in the reference (a centre on the max face sits in the last cell); the cube is the f32
            // bounding box of these very centres (hpmvs_root_cube), so a centre ON a face may miss it by a rounding of origin / width
            if (rel < 0.0 && rel > -1e-6) rel = 0.0;
            if (rel > 1.0 && rel < 1.0 + 1e-6) rel = 1.0;
            if (!(rel >= 0.0 && rel <= 1.0)) { inside = false; break; }
            const double c = std::floor(rel * (double)(1u << MAXL));
            q[3 * (size_t)i + a] = (uint32_t)std::min(c, (double)((1u << MAXL) - 1));
        }
        if (inside) root.pts.push_back(i);
    }
    auto split = [&](const SubTree& s, std::vector<SubTree>& out) {
        SubTree ch[8];
        const int sh = MAXL - (s.level + 1);
        for (int c = 0; c < 8; c++) {
            ch[c].level = s.level + 1; ch[c].rank = 0;
            for (int a = 0; a < 3; a++) ch[c].k[a] = (s.k[a] << 1) | ((c >> a) & 1);
        }
        for (int i : s.pts) {
            int c = 0;
            for (int a = 0; a < 3; a++) c |= (int)((q[3 * (size_t)i + a] >> sh) & 1u) << a;
            ch[c].pts.push_back(i);
        }
        for (int c = 0; c < 8; c++) if (!ch[c].pts.empty()) out.push_back(std::move(ch[c]));
    };
    subs.clear();
    if (min_subtrees < 2) subs.push_back(std::move(root));
    else {
        split(root, subs);
        while ((int)subs.size() < min_subtrees && !subs.empty()) {
            size_t big = 0;
            for (size_t i = 1; i < subs.size(); i++) if (subs[i].pts.size() > subs[big].pts.size()) big = i;
            if (subs[big].pts.size() < 100 || subs[big].level >= MAXL) break;
            std::vector<SubTree> next;
            split(subs[big], next);
            for (size_t i = 0; i < subs.size(); i++) if (i != big) next.push_back(std::move(subs[i]));
            subs.swap(next);
        }
    }
    std::vector<int> order(subs.size());
    for (size_t i = 0; i < subs.size(); i++) order[i] = (int)i;
    // the cost of a patch grows with the number of views it is measured in (textures per evaluation): the sub-trees are dealt by the
    // sum of their patches' view counts, not by their patch count (configs[3] at 8 ranks: scoring work max/mean 1.18 by count, 1.10 by
    // views - profiles/r2_work_balance.txt; the measured step time per rank is latency bound and hardly moves: 10.9 -> 10.7 ms)
    std::vector<int64_t> weight(subs.size(), 0);
    for (size_t si = 0; si < subs.size(); si++)
        for (int i : subs[si].pts) weight[si] += (int64_t)std::max(1, patches[i].nimages);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return weight[a] > weight[b]; });
    std::vector<int64_t> load((size_t)nranks, 0);
    for (int si : order) {
        int r = 0;
        for (int j = 1; j < nranks; j++) if (load[j] < load[r]) r = j;
        load[r] += weight[si];
        subs[si].rank = r;
    }
    return (int)subs.size();
}

// cell_of[i] = sub-tree of patch i (-1: outside the root cube), rank_of[i] = its rank (-1 likewise).  Returns the number of sub-trees.
int hpmvs_shard_cells(int n, const hpmvs_patch_t* patches, const double origin[3], double root_width, int min_subtrees, int nranks,
                      int32_t* cell_of, int32_t* rank_of) {
    if (n < 0 || (n > 0 && (!patches || !cell_of || !rank_of)) || !origin || !(root_width > 0.0) || nranks < 1) return HPMVS_E_ARG;
    std::vector<SubTree> subs;
    build_subtrees(n, patches, origin, root_width, min_subtrees, nranks, subs);
    for (int i = 0; i < n; i++) { cell_of[i] = -1; rank_of[i] = -1; }
    for (size_t si = 0; si < subs.size(); si++)
        for (int i : subs[si].pts) { cell_of[i] = (int)si; rank_of[i] = subs[si].rank; }
    return (int)subs.size();
}

int hpmvs_shard_subtrees(int n, const hpmvs_patch_t* patches, const double origin[3], double root_width, int min_subtrees, int nranks,
                         int cap, int32_t* sub_level, int64_t* sub_key, int32_t* sub_rank) {
    if (n < 0 || (n > 0 && !patches) || !origin || !(root_width > 0.0) || nranks < 1 || cap < 0 || (cap > 0 && (!sub_level || !sub_key || !sub_rank)))
        return HPMVS_E_ARG;
    std::vector<SubTree> subs;
    build_subtrees(n, patches, origin, root_width, min_subtrees, nranks, subs);
    if ((int)subs.size() > cap) return -(int)subs.size() - 100;
    for (size_t i = 0; i < subs.size(); i++) {
        sub_level[i] = subs[i].level;
        for (int a = 0; a < 3; a++) sub_key[3 * i + a] = (int64_t)subs[i].k[a];
        sub_rank[i] = subs[i].rank;
    }
    return (int)subs.size();
}

// Border de-duplication after the final multi-GPU gather: patches of DIFFERENT ranks that fall into the same cubic cell of edge `cell`
// are reduced to the best-supported one - most views first (CellProcessor::filter, src/hpmvs/CellProcessor.cpp:43-82), then the lower
// final score, then the lower rank; patches of the winner's own rank in that cell all stay (merging inside a shard is the scheduler's
// job).  keep[] receives the surviving indices in ascending order; returns their number.
int hpmvs_dedup_border(int n, const hpmvs_patch_t* rec, const int32_t* owner, const double origin[3], double cell, int32_t* keep) {
    if (n < 0 || (n > 0 && (!rec || !owner || !keep)) || !(cell > 0.0)) return HPMVS_E_ARG;
    const double zero[3] = {0.0, 0.0, 0.0};
    if (!origin) origin = zero;
    struct Best { int idx; };
    struct Key { int64_t k[3]; bool operator==(const Key& o) const { return k[0] == o.k[0] && k[1] == o.k[1] && k[2] == o.k[2]; } };
    struct KeyHash { size_t operator()(const Key& a) const { return (size_t)(a.k[0] * 73856093ll ^ a.k[1] * 19349663ll ^ a.k[2] * 83492791ll); } };
    std::unordered_map<Key, int, KeyHash> best;
    std::vector<Key> keys((size_t)n);
    auto better = [&](int a, int b) {     // a beats b
        if (rec[a].nimages != rec[b].nimages) return rec[a].nimages > rec[b].nimages;
        if (rec[a].score != rec[b].score) return rec[a].score < rec[b].score;
        if (owner[a] != owner[b]) return owner[a] < owner[b];
        return a < b;
    };
    for (int i = 0; i < n; i++) {
        if (rec[i].status != HPMVS_OK) continue;
        for (int a = 0; a < 3; a++) keys[i].k[a] = (int64_t)std::floor(((double)rec[i].center[a] - origin[a]) / cell);
        auto it = best.find(keys[i]);
        if (it == best.end()) best[keys[i]] = i;
        else if (better(i, it->second)) it->second = i;
    }
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (rec[i].status != HPMVS_OK) continue;
        const int w = best[keys[i]];
        if (i == w || owner[i] == owner[w]) keep[m++] = i;
    }
    return m;
}

}  // extern "C"
