/* hpmvs_b200 - C ABI of the B200-native patch-optimisation engine.
 *
 * This is the drop-in boundary for ONE path of alexlocher/hpmvs: everything below
 *     bool mo3d::PatchOptimizer::optimize(mo3d::Patch3d&)
 *         (/root/reference/include/hpmvs/PatchOptimizer.h:39-43, src/hpmvs/PatchOptimizer.cpp:78-103)
 * i.e. view selection + 3-DoF BOBYQA refinement of the mean robust (1-NCC) photometric score + reference
 * view re-selection + patch colour, batched over many patches and executed by hand-written sm_100a kernels.
 * Callers in the reference: Scene::initPatches (src/hpmvs/Scene.cpp:167), CellProcessor::extend / branch
 * (src/hpmvs/CellProcessor.cpp:129,256).
 *
 * Plain C: pointers and sizes only.  All functions return 0 on success or a negative HPMVS_E_* code;
 * a patch that the reference would reject (optimize() == false) is NOT an error - it is reported per patch
 * in hpmvs_patch_t.status and its geometry fields are left as given (the reference mutates a Patch3d only
 * on success, PatchOptimizer.cpp:86-93).  There is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef HPMVS_B200_H
#define HPMVS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPMVS_LEVELS 6        /* pyramid levels 0..MAXLEVEL (Image.cpp:38,43; Camera.cpp:36) */
#define HPMVS_MAX_VIEWS 32    /* capacity of a patch's view list (the reference's std::vector<int> images_) */

/* error codes (function return values) */
#define HPMVS_E_ARG      (-1)
#define HPMVS_E_CUDA     (-2)
#define HPMVS_E_STATE    (-3)   /* cameras / images / covisibility not uploaded yet */
#define HPMVS_E_NODEVICE (-4)

/* Replaces mo3d::HpmvsOptions (include/hpmvs/HpmvsOptions.h:29-58); only the fields the path reads. */
typedef struct hpmvs_options {
    int32_t maxlevel;              /* MAXLEVEL 5 */
    int32_t minlevel;              /* MINLEVEL 0 */
    int32_t start_level;           /* START_LEVEL 4 (used by the seeding caller) */
    float   max_angle;             /* MAX_ANGLE 60 deg in radians, f32 */
    float   min_angle;             /* MIN_ANGLE 10 deg in radians, f32 */
    int32_t max_images_per_patch;  /* MAX_IMAGES_PER_PATCH 6 - carried but unused, as in PatchOptimizer.cpp:298 */
    int32_t min_images_per_patch;  /* MIN_IMAGES_PER_PATCH 3 */
    float   ncc_alpha_1;           /* NCC_ALPHA_1 0.4 */
    float   ncc_alpha_2;           /* NCC_ALPHA_2 0.5 */
} hpmvs_options_t;

/* Replaces the read-only part of mo3d::Camera the path touches (include/hpmvs/Camera.h:87-106):
 * projection_[level], center_, xAxis_/yAxis_/zAxis_, kMat_[0](0,0),(1,1); plus the pyramid sizes the path
 * asks mo3d::Image for (Image.h:62-63). */
typedef struct hpmvs_camera {
    float   P[HPMVS_LEVELS][3][4];
    float   center[4];
    float   xaxis[3], yaxis[3], zaxis[3];
    float   k00, k11;
    int32_t width[HPMVS_LEVELS], height[HPMVS_LEVELS];
} hpmvs_camera_t;

/* per-patch outcome = the stage at which PatchOptimizer::runOptimization (PatchOptimizer.cpp:48-76) bailed out */
enum {
    HPMVS_OK = 0,
    HPMVS_FAIL_ADD_IMAGES = 1,     /* :50  */
    HPMVS_FAIL_NCC1 = 2,           /* :52  */
    HPMVS_FAIL_ANGLES = 3,         /* :55  */
    HPMVS_FAIL_OPT_MINIMAGES = 4,  /* :323 */
    HPMVS_FAIL_OPT_ROUNDOFF = 5,   /* nlopt::roundoff_limited caught at :369 */
    HPMVS_FAIL_OPT_MAXEVAL = 6,    /* MAXEVAL_REACHED is not in the success list at :367 */
    HPMVS_FAIL_OPT_OTHER = 7,
    HPMVS_FAIL_ADD_IMAGES2 = 8,    /* :63  */
    HPMVS_FAIL_NCC2 = 9,           /* :65  */
    HPMVS_FAIL_ANGLE_FILTER = 10,  /* :67  */
    HPMVS_FAIL_ANGLES2 = 11,       /* :69  */
    HPMVS_FAIL_NCC3 = 12,          /* :72  */
    HPMVS_FAIL_TOO_MANY_VIEWS = 13 /* view list would exceed HPMVS_MAX_VIEWS (engine limit, not in the reference) */
};

/* Replaces the fields of mo3d::Patch3d that optimize() reads and writes (include/hpmvs/Patch3d.h:55-82).
 * One fixed-size record per patch, 208 bytes, used for input and output. */
typedef struct hpmvs_patch {
    /* in / out */
    float   center[4];             /* center_ (w = 1) */
    float   normal[4];             /* normal_ (w = 0) */
    float   scale;                 /* scale_3dx_ */
    int32_t nimages;               /* images_.size() */
    int32_t images[HPMVS_MAX_VIEWS]; /* images_, [0] = reference view */
    /* out */
    float   color[3];              /* color_ = Scene::getColor(patch), Scene.cpp:300-327 */
    float   ncc;                   /* ncc_ : the reference hard-codes 1.4 (PatchOptimizer.cpp:95) */
    int32_t status;                /* HPMVS_OK or HPMVS_FAIL_* */
    int32_t nlopt_result;          /* nlopt result code of the refinement (1,4 = ok; 5, -4 = rejected) */
    int32_t evals;                 /* objective evaluations spent by the refinement */
    int32_t textures;              /* 7x7x3 textures sampled for this patch (unit of the roofline model) */
    double  score;                 /* final mean robust (1-NCC) of the refinement (discarded by the reference) */
} hpmvs_patch_t;

typedef struct hpmvs_counters {
    uint64_t patches;              /* patches processed since creation / last reset */
    uint64_t patches_ok;
    uint64_t evals;                /* objective evaluations */
    uint64_t textures;             /* textures sampled (each = 49 bilinear RGB taps = 588 gathered bytes) */
    uint64_t kernel_launches;      /* engine kernels launched (all kinds) */
} hpmvs_counters_t;

typedef struct hpmvs_engine hpmvs_engine_t;

/* Replaces PatchOptimizer::PatchOptimizer(const HpmvsOptions&, const Scene*) (PatchOptimizer.cpp:38-45).
 * One engine per GPU; `device` is the CUDA ordinal. */
int  hpmvs_engine_create(const hpmvs_options_t *opt, int device, hpmvs_engine_t **out);
void hpmvs_engine_destroy(hpmvs_engine_t *e);

/* Replaces the borrowed pointer scene->cameras_.data() (PatchOptimizer.cpp:39). Copies n cameras to HBM. */
int  hpmvs_engine_set_cameras(hpmvs_engine_t *e, int n, const hpmvs_camera_t *cams);

/* Replaces the borrowed pointer scene->images_.data() (PatchOptimizer.cpp:40): uploads ONE pyramid level of
 * one view.  `rgb` is host memory, interleaved u8 RGB exactly as mo3d::Image stores it (Image.cpp:62-63),
 * `pitch_bytes` between rows (>= 3*w).  On the device the level lives as a pitched RGBX u8 array. */
int  hpmvs_engine_upload_image(hpmvs_engine_t *e, int cam, int level, const uint8_t *rgb, int w, int h,
                               size_t pitch_bytes);

/* Replaces Image::load's undistortion step (Image.cpp:51-53 -> Image::undistort, :68-149) on the GPU: `rgb` is the DISTORTED level-0
 * image of view `cam` (focal length f, VisualSFM radial parameter r from the NVM camera line); level 0 on the device becomes the
 * undistorted image (target pixels without a source are 0).  r == 0: plain upload.  The source positions are computed with CUDA's
 * double-precision libm: identical to the host function hpmvs_undistort_rgb() except for rare last-place roundings (<= 1 grey level on
 * < 0.01 % of the pixels, tests/test_next_rows.py). */
int  hpmvs_engine_upload_image_undistort(hpmvs_engine_t *e, int cam, const uint8_t *rgb, int w, int h, size_t pitch_bytes,
                                         double f, double r);

/* Optional replacement for Image::load's pyramid loop (Image.cpp:56-57, CImg get_resize_halfXY): builds
 * levels 1..maxlevel of view `cam` on the GPU from the already uploaded level 0 (bit-exact with CImg). */
int  hpmvs_engine_build_pyramid(hpmvs_engine_t *e, int cam);
/* Reads a device level back as interleaved RGB (tests, debugging). */
int  hpmvs_engine_download_image(hpmvs_engine_t *e, int cam, int level, uint8_t *rgb, size_t pitch_bytes);

/* Replaces the borrowed pointer &scene->covis_ (PatchOptimizer.cpp:41): CSR lists, offsets has ncams+1 entries. */
int  hpmvs_engine_set_covis(hpmvs_engine_t *e, const int32_t *offsets, const int32_t *ids);

/* Replaces n calls of PatchOptimizer::optimize(Patch3d&) (PatchOptimizer.cpp:78-103).
 * Host buffers (pinned recommended): copies `in` to the device, runs the fused kernel, copies results to `out`
 * (may alias `in`).  `stream` is a cudaStream_t (NULL = the engine's own stream); the call returns after the
 * stream has been synchronised. */
int  hpmvs_optimize_batch(hpmvs_engine_t *e, int n, const hpmvs_patch_t *in, hpmvs_patch_t *out, void *stream);

/* Asynchronous form: the copies and the kernel are only enqueued on `stream` (a cudaStream_t, not NULL; host buffers must be pinned
 * for the copies to be asynchronous) and the call returns at once; the caller synchronises the stream before reading `out`.
 * Batches submitted on different streams overlap - the persistent CTAs of the next batch start on the SMs the previous batch has
 * already drained, which hides the tail of long optimisations (about +25 % on 10 k-patch batches). */
int  hpmvs_optimize_batch_submit(hpmvs_engine_t *e, int n, const hpmvs_patch_t *in, hpmvs_patch_t *out, void *stream);

/* Where the one libm-dependent scalar step of the path is evaluated: parametersFromCenterNorm's starting angles
 * std::asin(float) / std::cos / std::acos (src/hpmvs/PatchOptimizer.cpp:427-437).
 *   mode 0 (default): on the device, asin correctly rounded (= a reference linked against a correctly rounded libm);
 *   mode 1: hpmvs_optimize_batch() evaluates them on the HOST with the caller's libm - the reference's own calls in the
 *           reference's own place - and hands them to the kernel, so that results are bit-identical to a reference built
 *           on the same machine (glibc 2.39's asinf is not correctly rounded).  Costs ~60 ns of host time per patch.
 * The device-resident call always uses mode 0. */
int  hpmvs_engine_set_start_mode(hpmvs_engine_t *e, int mode);

/* Same work on device-resident records (no copies, asynchronous on `stream`). */
int  hpmvs_optimize_batch_device(hpmvs_engine_t *e, int n, const hpmvs_patch_t *d_in, hpmvs_patch_t *d_out,
                                 void *stream);

/* Start mode 1 for device-resident records: hpmvs_start_parameters() evaluates parametersFromCenterNorm's two angles
 * (PatchOptimizer.cpp:427-443) for n HOST records with the caller's libm into out[2*n] (host); the caller copies them to the device
 * next to the records and passes them as d_start (NULL = mode 0). */
int  hpmvs_start_parameters(hpmvs_engine_t *e, int n, const hpmvs_patch_t *in, double *out);
int  hpmvs_optimize_batch_device_start(hpmvs_engine_t *e, int n, const hpmvs_patch_t *d_in, hpmvs_patch_t *d_out,
                                       const double *d_start, void *stream);

/* Replaces n calls of PatchOptimizer::setINCCs (PatchOptimizer.cpp:448-474) on patches as given:
 * inccs[i*HPMVS_MAX_VIEWS + k] = (robust ? r/(1+3r) : r), r = 1-NCC(view ref_idx, view k); 2.0 where invalid. */
int  hpmvs_ncc_batch(hpmvs_engine_t *e, int n, const hpmvs_patch_t *in, int ref_idx, int robust, float *inccs,
                     void *stream);
/* Same scoring on device-resident records and a device-resident result array (n * HPMVS_MAX_VIEWS floats): no copies,
 * asynchronous on `stream`; hpmvs_engine_last_kernel_ms() then reports this kernel's duration. */
int  hpmvs_ncc_batch_device(hpmvs_engine_t *e, int n, const hpmvs_patch_t *d_in, int ref_idx, int robust,
                            float *d_inccs, void *stream);

/* ---- "next" rows: what CellProcessor::extend / branch do right around optimize() ---------------------------------- */

/* Replaces the depth-map allocation in Scene::addCameras (src/hpmvs/Scene.cpp:74-81): (re)creates every
 * (view, level) depth map at MAX_DEPTH = 1000.  Call after the cameras are set. */
int  hpmvs_engine_depth_reset(hpmvs_engine_t *e);
/* Replaces n calls of Scene::setDepths(patch, false) (Scene.cpp:351-381) for the records with status == HPMVS_OK;
 * the "smaller depth wins" update is an atomic float min on the device. */
int  hpmvs_depth_set_batch(hpmvs_engine_t *e, int n, const hpmvs_patch_t *patches, void *stream);
/* Replaces n calls of Scene::setDepths(patch, true) (Scene.cpp:351-381, :372-373) - a patch leaves the tree (CellProcessor::filter :76,
 * ::branch :273): every depth cell that still holds exactly this patch's depth goes back to MAX_DEPTH (atomic compare-and-swap). */
int  hpmvs_depth_unset_batch(hpmvs_engine_t *e, int n, const hpmvs_patch_t *patches, void *stream);
/* Replaces, per patch, Scene::depthTests / viewBlockTest / pixelFreeTests (Scene.cpp:518-644) as used by
 * CellProcessor::extend (src/hpmvs/CellProcessor.cpp:134-142): out[3*i+0..2] = the three counts. */
int  hpmvs_accept_batch(hpmvs_engine_t *e, int n, const hpmvs_patch_t *patches, float margin, int32_t *out, void *stream);
/* Reads one depth map back (rows x cols f32, row-major); out may be NULL to query the size. */
int  hpmvs_engine_download_depth(hpmvs_engine_t *e, int cam, int level, float *out, int *rows, int *cols);
/* Replaces the candidate construction of CellProcessor::extend (mode 6, CellProcessor.cpp:98-119) and ::branch
 * (mode 4, :227-249) for n parent patches in cells of width widths[i]; out receives mode*n records. Host function. */
int  hpmvs_expand_candidates(int ncams, const hpmvs_camera_t *cams, int n, const hpmvs_patch_t *parents,
                             const float *widths, int mode, hpmvs_patch_t *out);

/* Device-resident forms of the three steps above (records, widths and results live in HBM; no host staging; asynchronous on `stream`):
 * a level of the loop can chain candidates -> hpmvs_optimize_batch_device -> acceptance counts -> depth updates without the patch
 * records crossing PCIe; only the per-patch verdicts (3 ints) and the cell bookkeeping need the host.  View ids in device-resident
 * records are the caller's responsibility (the host-buffer calls validate them). */
int  hpmvs_expand_candidates_device(hpmvs_engine_t *e, int n, const hpmvs_patch_t *d_parents, const float *d_widths, int mode,
                                    hpmvs_patch_t *d_out, void *stream);
int  hpmvs_depth_set_batch_device(hpmvs_engine_t *e, int n, const hpmvs_patch_t *d_patches, int subtract, void *stream);
int  hpmvs_accept_batch_device(hpmvs_engine_t *e, int n, const hpmvs_patch_t *d_patches, float margin, int32_t *d_out, void *stream);

/* ---- the loop around the path: level-synchronous expand -> optimize -> filter driver (host C++, hpmvs_b200/csrc/host_pipeline.cpp) ----
 * Batching stand-in for the reference's scheduler (CellProcessor::processQueue + DynOctTree, src/main.cpp:145-155): per octree level
 * all candidates of all cells go through ONE hpmvs_optimize_batch / hpmvs_accept_batch.  See the file header for what is kept. */
#define HPMVS_PIPELINE_MAX_LEVELS 24
typedef struct hpmvs_pipeline_params {
    double  origin[3];             /* low corner of the octree's root cube (Scene.cpp:186-193: centre - width/2) */
    double  root_width;            /* its edge length; cells of tree level L are root_width / 2^L wide */
    int32_t start_level;           /* level the accepted seeds are inserted at */
    int32_t final_level;           /* last level (inclusive) */
    int32_t final_min_level;       /* HpmvsOptions::PATCH_FINAL_MINLEVEL (9 from the CLI, src/main.cpp:44,234) */
    int32_t max_rounds;            /* extension rounds per level (64) */
    int32_t dedup_ref_pixel;       /* != 0: one accepted candidate per reference-view image cell and round */
    int32_t ncams;
    const hpmvs_camera_t *cams;    /* the cameras the engine was given (candidate construction runs on the host) */
    int32_t shard_count;           /* multi-GPU: > 1 = this call grows only the cells of tree level shard_level dealt to shard_rank; */
    int32_t shard_rank;            /*   0 or 1 = everything.  Merge the ranks' results with an NCCL gather + hpmvs_dedup_border. */
    int32_t shard_level;           /*   (<= start_level; used when there is no sub-tree table: cells of that level dealt out by a hash) */
    int32_t minlevel;              /* HpmvsOptions::MINLEVEL (0): a patch whose views are all at pyramid level <= minlevel is exhausted and
                                      never branches (Scene::getLevelSupport, CellProcessor.cpp:222-225) */
    /* sub-tree table from hpmvs_shard_subtrees() (the reference's getSubTrees split): sub-tree i is the cell sub_key[3i..3i+2] of tree
     * level sub_level[i] and belongs to rank sub_rank[i]; space outside every sub-tree belongs to nobody.  nsub == 0: hash by shard_level. */
    int32_t nsub;
    const int32_t *sub_level;
    const int64_t *sub_key;
    const int32_t *sub_rank;
    /* per-step border hand-off between ranks (CellProcessor.cpp:147-153, distributeBorderCell :487-540): called by EVERY rank the same
     * number of times with the records it accepted in this step; *recv must point at the concatenation of all ranks' records in rank
     * order (valid until the next call), *n_recv their number.  NULL: no hand-off (a patch that leaves the rank's cells is dropped). */
    int (*exchange)(void *user, int n_send, const hpmvs_patch_t *send, hpmvs_patch_t **recv, int *n_recv);
    void *exchange_user;
} hpmvs_pipeline_params_t;
typedef struct hpmvs_pipeline_stats {
    int64_t optimize_calls, optimized_ok;
    double  seconds_optimize, seconds_accept;
    int32_t nlevels;
    int32_t level[HPMVS_PIPELINE_MAX_LEVELS];
    int64_t extended[HPMVS_PIPELINE_MAX_LEVELS], branched[HPMVS_PIPELINE_MAX_LEVELS];
    int64_t exchanged;             /* records received through the exchange callback */
} hpmvs_pipeline_stats_t;
/* seeds: candidate patches as hpmvs_seed_patches() builds them.  *out receives a malloc'ed array of *nout final patches
 * (release with hpmvs_free).  Resets the engine's depth maps first. */
int  hpmvs_pipeline_run(hpmvs_engine_t *e, const hpmvs_pipeline_params_t *params, int nseeds, const hpmvs_patch_t *seeds,
                        hpmvs_patch_t **out, int *nout, hpmvs_pipeline_stats_t *stats);
void hpmvs_free(void *p);
/* Border de-duplication after the final multi-GPU gather of the patch records (host C++; the gather itself is an NCCL all_gather,
 * hpmvs_b200/gather.py): patches of different ranks (owner[i]) in the same cubic cell of edge `cell` are reduced to the best-supported
 * one (most views, CellProcessor::filter, CellProcessor.cpp:43-82; then lower score, then lower rank); cells are counted from `origin`
 * (the octree's low corner; NULL = world origin); keep[] receives the surviving
 * indices in ascending order (capacity n), the return value is their number. */
int  hpmvs_dedup_border(int n, const hpmvs_patch_t *records, const int32_t *owner, const double origin[3], double cell, int32_t *keep);
/* The same de-duplication on device-resident records (the gathered set on the root rank, straight out of the NCCL receive buffer):
 * d_keep[i] = 1 for the survivors, *d_nkeep (device int, may be NULL) their number.  Four small kernels (hash insert on the packed cell
 * key, then atomic reductions); asynchronous on `stream`. */
int  hpmvs_dedup_border_device(hpmvs_engine_t *e, int n, const hpmvs_patch_t *d_records, const int32_t *d_owner, const double origin[3],
                               double cell, uint8_t *d_keep, int32_t *d_nkeep, void *stream);
/* Root cube of the patch octree as Scene::initPatches forms it (src/hpmvs/Scene.cpp:186-193): f32 bounding box of the centres, edge =
 * largest extent, centred on the box; origin = its low corner. */
int  hpmvs_root_cube(int n, const hpmvs_patch_t *patches, double origin[3], double *width);
/* Replaces getSubTrees (src/main.cpp:50-96; DynOctTree::getSubTrees, include/hpmvs/doctree.h:513-523) as a partition of a patch set:
 * split the root cube into its non-empty children, keep splitting the sub-tree with the most patches until there are >= min_subtrees
 * (or the biggest holds < 100, main.cpp:74); deal the sub-trees to nranks ranks, costliest first (cost = sum of its patches' view
 * counts) to the least loaded rank.
 * cell_of[i] / rank_of[i] = sub-tree / rank of patch i (-1 outside the cube).  Returns the number of sub-trees. */
int  hpmvs_shard_cells(int n, const hpmvs_patch_t *patches, const double origin[3], double root_width, int min_subtrees, int nranks,
                       int32_t *cell_of, int32_t *rank_of);
/* The same split as a table for hpmvs_pipeline_params_t: sub-tree i = cell sub_key[3i..3i+2] of tree level sub_level[i], dealt to rank
 * sub_rank[i].  Arrays need room for `cap` sub-trees; returns their number (or -needed when cap is too small). */
int  hpmvs_shard_subtrees(int n, const hpmvs_patch_t *patches, const double origin[3], double root_width, int min_subtrees, int nranks,
                          int cap, int32_t *sub_level, int64_t *sub_key, int32_t *sub_rank);

/* ---- host-side scene surface (plain C++ on the host, no GPU needed): what feeds the engine ---- */

/* Replaces mo3d::Camera::init (src/hpmvs/Camera.cpp:34-81) for one NVM camera line
 * (NVMReader.cpp:63-74: focal f, quaternion wxyz, centre c; principal point = image centre) and fills the
 * pyramid sizes Image::load would produce (Image.cpp:56-57: each level is floor(w/2) x floor(h/2)). */
int  hpmvs_camera_from_nvm(double f, const double q_wxyz[4], const double c[3], int width, int height,
                           int maxlevel, hpmvs_camera_t *out);

/* Replaces Scene::extractCoVisiblilty (src/hpmvs/Scene.cpp:241-298).  Points are given as CSR measurement
 * lists (meas_offsets[npoints+1], meas_cam[]).  compat != 0 reproduces the reference's counter indexing by
 * measurement POSITION (Scene.cpp:260-264); compat == 0 counts by camera id.  Two cameras are covisible when
 * they share >= 50 points.  out_offsets has ncams+1 entries; returns the number of ids written, or the
 * required capacity (negated) if ids_cap is too small. */
int  hpmvs_extract_covis(int ncams, int npoints, const int32_t *meas_offsets, const int32_t *meas_cam,
                         int compat, int32_t *out_offsets, int32_t *out_ids, int ids_cap);

/* Replaces the candidate construction of Scene::initPatches (src/hpmvs/Scene.cpp:116-165), i.e. everything
 * before the optimize() call: visibility at START_LEVEL with a 2 px margin, normal towards the first view,
 * scale = getScale(centre, START_LEVEL).  valid[i] = 0 where the reference skips the point. */
int  hpmvs_seed_patches(const hpmvs_options_t *opt, int ncams, const hpmvs_camera_t *cams, int npoints,
                        const double *xyz, const int32_t *meas_offsets, const int32_t *meas_cam,
                        hpmvs_patch_t *out, uint8_t *valid);

/* ---- file formats on either side of the path (host) ---------------------------------------------------------------- */

typedef struct hpmvs_nvm hpmvs_nvm_t;
/* Replaces mo3d::NVMReader::readFile (src/hpmvs/NVMReader.cpp:115-155): NVM_V3 text; cameras
 * `file f qw qx qy qz cx cy cz r 0`, points `x y z r g b n (img feat u v)*n`; models are consumed until an empty
 * one, model 0 is kept (src/main.cpp:112-116).  fix_path != 0 resolves relative image names against the NVM folder. */
int  hpmvs_nvm_open(const char *path, int fix_path, hpmvs_nvm_t **out);
void hpmvs_nvm_close(hpmvs_nvm_t *m);
int  hpmvs_nvm_num_models(const hpmvs_nvm_t *m);
int  hpmvs_nvm_num_cameras(const hpmvs_nvm_t *m);
int  hpmvs_nvm_num_points(const hpmvs_nvm_t *m);
int  hpmvs_nvm_num_measurements(const hpmvs_nvm_t *m);
int  hpmvs_nvm_camera(const hpmvs_nvm_t *m, int i, char *filename, int cap, double *f, double q_wxyz[4], double c[3], double *r);
/* any output pointer may be NULL; offsets has num_points+1 entries (CSR over the measurement arrays) */
int  hpmvs_nvm_points(const hpmvs_nvm_t *m, double *xyz, double *rgb, int32_t *offsets, int32_t *meas_cam,
                      int32_t *meas_feat, double *meas_xy);
/* Replaces mo3d::NVMReader::saveNVM (src/hpmvs/NVMReader.cpp:157-183) for the model held by `m`: same text, byte for byte. */
int  hpmvs_nvm_write(const hpmvs_nvm_t *m, const char *path);
/* Replaces Image::undistort (src/hpmvs/Image.cpp:68-149): undoes VisualSFM's radial distortion `r` (NVM camera line, NVMReader.cpp:70)
 * on an interleaved u8 RGB level-0 image with focal length f, before the pyramid is built; r == 0 copies.  Host function (same libm
 * as the reference for its pow / complex pow), bit-exact against the reference build incl. the f32 -> u8 truncation.  Target pixels
 * whose source falls outside the image are left 0 (the reference leaves them UNINITIALISED, Image.cpp:79); `written` (width*height
 * bytes, may be NULL) marks the pixels that were set. */
int  hpmvs_undistort_rgb(const uint8_t *rgb, int width, int height, double f, double r, uint8_t *out, uint8_t *written);
/* Level-0 image reader (binary PPM "P6", maxval 255); call with rgb == NULL to get the size first. */
int  hpmvs_ppm_read(const char *path, int *width, int *height, uint8_t *rgb);
/* Replaces DynOctTree::toExtPly (include/hpmvs/doctree.h:525-622): vertex element {x y z [nx ny nz] red green blue
 * [scalar_scale]} + point_visibility element {list uint uint visible_cameras}; ascii or binary little endian. */
int  hpmvs_ply_write_ext(const char *path, int n, const hpmvs_patch_t *patches, int binary, int normal, int scale,
                         int visibility);

int  hpmvs_engine_counters(hpmvs_engine_t *e, hpmvs_counters_t *out, int reset);
/* The engine's own stream as a cudaStream_t, so callers can record events around asynchronous calls. */
void *hpmvs_engine_stream(hpmvs_engine_t *e);
/* Device-side duration of the most recent fused optimize kernel (or *_device scoring kernel) in milliseconds (CUDA events
 * on its stream). */
float hpmvs_engine_last_kernel_ms(hpmvs_engine_t *e);
const char *hpmvs_error_string(int code);
int  hpmvs_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HPMVS_B200_H */
