/* TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT.
 *
 * CPU oracle for the HPMVS hot path: a plain C++ (Eigen-free) restatement of
 *   /root/reference/src/hpmvs/PatchOptimizer.cpp      (optimize() and everything below it)
 *   /root/reference/include/hpmvs/Patch2d.hpp:37-84   (normalize / dot)
 *   /root/reference/include/hpmvs/Image.h:89-115      (bilinear getColor)
 *   /root/reference/include/hpmvs/Camera.h:45-62, src/hpmvs/Camera.cpp:34-99
 *   /root/reference/src/hpmvs/Scene.cpp:90-208,241-327 (seeding, covisibility, median colour)
 *   /root/reference/src/hpmvs/Image.cpp:41-66 + thirdLibs/cimg/CImg.h:21189-21203 (pyramid)
 * linked against the REAL vendored BOBYQA (oracle/_ref/libnlopt_ref.a, built by oracle/Makefile
 * from /root/reference/thirdLibs/nlopt-2.4.2 where it lies).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (hpmvs_b200/) never does.
 *
 * PARITY STATUS: pinned against the reference's own code.  HPMVS ships no golden vectors, but its sources compile
 * here where they lie (oracle/Makefile `refhpmvs` -> oracle/_ref/libhpmvs_ref.so + the hpmvs_ref CLI) once its
 * external dependencies that this image lacks are stood in for (oracle/shim: Eigen with its published evaluation
 * order restated, glog, gflags, jpeglib declarations).  tests/test_reference_golden.py checks this restatement
 * bit for bit against that build - Camera::init, the CImg pyramid, covisibility, seeds, optimize() (centre, normal,
 * views, colour, success), Scene::initPatches run whole, setDepths and the three acceptance tests - live and through
 * committed fixtures (tests/golden/ref_*.npz).  What stays unpinned is Eigen itself (absent, not vendored, no version
 * pinned by the reference): the shim's evaluation-order rules are documented in oracle/shim/Eigen/Dense.
 * The optimizer component is the reference's own BOBYQA object code; tests/test_oracle.py also checks it against
 * nlopt's known-answer functions (thirdLibs/nlopt-2.4.2/test/testfuncs.c:65-89,445-447).
 */
#ifndef HPMVS_ORACLE_H
#define HPMVS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_VIEWS 64
#define ORC_LEVELS 6

/* mirrors HpmvsOptions.h:29-58 (only the fields the hot path reads) */
typedef struct {
    int32_t maxlevel;            /* 5 */
    int32_t minlevel;            /* 0 */
    int32_t start_level;         /* 4 */
    float   max_angle;           /* 60 deg in rad (f32) */
    float   min_angle;           /* 10 deg in rad (f32) */
    int32_t max_images_per_patch;/* 6 (unused by the reference, PatchOptimizer.cpp:298) */
    int32_t min_images_per_patch;/* 3 */
    float   ncc_alpha_1;         /* 0.4 */
    float   ncc_alpha_2;         /* 0.5 */
} orc_options_t;

typedef struct {
    float P[ORC_LEVELS][3][4];   /* Camera::projection_ */
    float center[4];
    float xaxis[3], yaxis[3], zaxis[3];
    float k00, k11;              /* kMat_[0](0,0), (1,1) */
    int32_t width[ORC_LEVELS], height[ORC_LEVELS];
} orc_camera_t;

/* stage at which optimize() returned false (PatchOptimizer.cpp:48-76) */
enum {
    ORC_OK = 0, ORC_FAIL_ADD_IMAGES = 1, ORC_FAIL_NCC1 = 2, ORC_FAIL_ANGLES = 3,
    ORC_FAIL_OPT_MINIMAGES = 4, ORC_FAIL_OPT_ROUNDOFF = 5, ORC_FAIL_OPT_MAXEVAL = 6,
    ORC_FAIL_OPT_OTHER = 7, ORC_FAIL_ADD_IMAGES2 = 8, ORC_FAIL_NCC2 = 9,
    ORC_FAIL_ANGLE_FILTER = 10, ORC_FAIL_ANGLES2 = 11, ORC_FAIL_NCC3 = 12,
    ORC_FAIL_TOO_MANY_VIEWS = 13
};

typedef struct {
    /* in/out (Patch3d.h:55-82) */
    float center[4];
    float normal[4];
    float scale;
    int32_t nimages;
    int32_t images[ORC_MAX_VIEWS];
    /* out */
    float color[3];
    float ncc;                   /* reference hard-codes 1.4 (PatchOptimizer.cpp:95) */
    int32_t status;
    int32_t nlopt_result;
    int32_t evals;               /* objective evaluations */
    int32_t textures;            /* sampleTexture calls that reached the sampling loop */
    double  last_val;            /* BOBYQA's final objective (discarded by the reference) */
} orc_patch_t;

void *orc_scene_new(const orc_options_t *opt);
void  orc_scene_free(void *scene);
/* NVM camera (NVMReader.cpp:63-74) + interleaved u8 RGB level-0 image; builds the pyramid */
int   orc_add_camera(void *scene, double f, const double q_wxyz[4], const double c[3],
                     int width, int height, const uint8_t *rgb);
int   orc_num_cameras(void *scene);
void  orc_get_camera(void *scene, int idx, orc_camera_t *out);
const uint8_t *orc_get_image(void *scene, int cam, int level, int *w, int *h);
/* Scene::extractCoVisiblilty incl. its index bug (Scene.cpp:241-298) */
void  orc_extract_covis(void *scene, int npoints, const int32_t *meas_offsets, const int32_t *meas_cam);
void  orc_set_covis(void *scene, const int32_t *offsets, const int32_t *ids);
int   orc_get_covis(void *scene, int cam, int32_t *out, int cap);
/* Scene::initPatches up to (not including) optimize(): Scene.cpp:116-165. valid[i]=0 if skipped */
void  orc_seed_patches(void *scene, int npoints, const double *xyz, const int32_t *meas_offsets,
                       const int32_t *meas_cam, orc_patch_t *out, uint8_t *valid);
int   orc_optimize(void *scene, orc_patch_t *patch);                 /* returns status */
void  orc_optimize_batch(void *scene, int n, orc_patch_t *patches, int nthreads);
/* PatchOptimizer::setINCCs (:448-474) on the patch as given */
void  orc_set_inccs(void *scene, const orc_patch_t *patch, int ref_idx, int robust, float *inccs);
/* PatchOptimizer::sampleTexture (:476-529): returns 1 and 147 floats on success */
int   orc_sample_texture(void *scene, const float center[4], float scale, const float xaxis[4],
                         const float yaxis[4], const float zaxis[4], int cam, float *tex147);
/* objective_fn (:286-311) at parameters x for the patch (after setOptimizationFields) */
double orc_objective(void *scene, const orc_patch_t *patch, const double x[3]);
/* Scene::getColor(const Patch3d&) (Scene.cpp:300-327) */
void  orc_patch_color(void *scene, const orc_patch_t *patch, float rgb[3]);

/* --- "next" rows: the steps right after optimize() in CellProcessor::extend (CellProcessor.cpp:134-142,197-201) --- */
void  orc_depth_reset(void *scene);                                           /* Scene.cpp:74-81 */
void  orc_depth_set_batch(void *scene, int n, const orc_patch_t *patches);    /* setDepths(p,false) for status==OK */
void  orc_depth_unset_batch(void *scene, int n, const orc_patch_t *patches);  /* setDepths(p,true) for status==OK */
const float *orc_get_depth(void *scene, int cam, int level, int *rows, int *cols);
/* out[3*i+0..2] = depthTests, viewBlockTest, pixelFreeTests (Scene.cpp:518-644) */
void  orc_accept_batch(void *scene, int n, const orc_patch_t *patches, float margin, int32_t *out);
/* candidates of extend (mode 6) / branch (mode 4): CellProcessor.cpp:98-119, 227-249; out has mode*n records */
void  orc_expand_candidates(void *scene, int n, const orc_patch_t *parents, const float *widths, int mode, orc_patch_t *out);

/* 1: evaluate std::asin(float) of PatchOptimizer.cpp:427 correctly rounded instead of with this box's libm */
void  orc_set_cr_asinf(int on);

/* Real nlopt BOBYQA (n=3, default initial step, xtol_rel, maxeval) on a built-in analytic test
 * function; records every evaluated point.  Used to pin the product's own BOBYQA. */
int   orc_bobyqa_testfunc(int func_id, const double x0[3], const double lb[3], const double ub[3],
                          double xtol_rel, int maxeval, double xout[3], double *fout,
                          double *trace_x, double *trace_f, int trace_cap, int *nevals);
double orc_testfunc_eval(int func_id, const double x[3]);

#ifdef __cplusplus
}
#endif
#endif
