/*
 * icp_oracle.h -- CPU restatement of the ICP-Variants registration inner loop.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under icp_variants_b200/ or include/ may include, link or
 * call this.  Users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference.
 *
 * PARITY PINNED AGAINST THE REFERENCE'S OWN CODE, NOT AGAINST ITS THIRD-PARTY LIBRARIES.  The
 * reference (/root/reference/icp-variants) ships no tests or golden vectors, and Eigen 3.3, FLANN
 * 1.8.4, Ceres 2.x and PCL are neither vendored nor installed.  oracle/_ref/libicp_ref.so is the
 * reference's headers compiled where they lie against small stand-ins for those libraries
 * (oracle/ref_shim/, oracle/ref_driver.cpp); tests/test_oracle_vs_reference.py runs every function of
 * this file against it (bit-exact transforms / correspondences / weights / rejection / pyramid /
 * selection / LM path, 1e-5 rad / 1e-5 m for the fp32 linear solves), and
 * tests/golden/reference_outputs.npz stores its outputs for machines without the library.  What stays
 * unpinned is the last-bit behaviour of the absent libraries themselves (Eigen's reduction order and
 * SVD, FLANN's approximate search, Ceres' DENSE_QR): there the oracle fixes a contract (DESIGN.md
 * "Numerics contract") and says so at the function.  Each function cites the file:line it follows.
 *
 * Layouts are the reference's: points/normals packed float[3N] (std::vector<Vector3f>), colours
 * uint8[4N] (std::vector<Vector4uc>), poses float[16] column-major (Eigen Matrix4f),
 * Match = {int idx; float weight}.
 */
#ifndef ICP_ORACLE_H
#define ICP_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t idx; float weight; } orc_match;

enum { ORC_METRIC_P2P = 0, ORC_METRIC_P2PLANE = 1, ORC_METRIC_SYMMETRIC = 2 };
enum { ORC_MIN_LINEAR = 0, ORC_MIN_LM = 1 };
enum { ORC_MATCH_KNN = 0, ORC_MATCH_PROJECTIVE = 1 };
enum { ORC_SELECT_ALL = 0, ORC_SELECT_RANDOM = 1 };
enum { ORC_WEIGHT_CONSTANT = 0, ORC_WEIGHT_DISTANCES = 1, ORC_WEIGHT_NORMALS = 2, ORC_WEIGHT_COLORS = 3 };
enum { ORC_NN_BRUTE = 0, ORC_NN_KDTREE = 1 };

typedef struct {
    int32_t metric;          /* ICPOptimizer::setMetric            (ICPOptimizer.h:46)  */
    int32_t minimizer;       /* LinearICPOptimizer / CeresICPOptimizer                  */
    int32_t matching;        /* setMatchingMethod                  (ICPOptimizer.h:71)  */
    int32_t selection;       /* setSelectionMethod                 (ICPOptimizer.h:58)  */
    double  proba;
    uint32_t seed;           /* reference seeds from random_device (selection.h:76-79); explicit here */
    int32_t weighting;       /* setWeightingMethod                 (ICPOptimizer.h:67)  */
    int32_t rejection;       /* setRejectionMethod, default 1      (ICPOptimizer.h:30)  */
    float   max_distance_sq; /* the matcher's m_maxDistance, squared (NearestNeighbor.h:17-19); setMatchingMaxDistance (ICPOptimizer.h:41) sets it */
    int32_t color_icp;       /* enableColorICP                     (ICPOptimizer.h:54)  */
    int32_t multires;        /* enableMultiResolution              (ICPOptimizer.h:50)  */
    int32_t n_iterations;    /* setNbOfIterations                  (ICPOptimizer.h:84)  */
    float   fx, fy, cx, cy;  /* setCameraParamsMatchingMethod      (ICPOptimizer.h:80)  */
    uint32_t width, height;
    int32_t nn_mode;         /* ORC_NN_BRUTE (literal O(N*M)) or ORC_NN_KDTREE (same answers, fast) */
    int32_t lm_max_iterations; /* Ceres options.max_num_iterations = 10 (ICPOptimizer.h:358) */
    int32_t pyramid_mode;    /* 0 = the reference's stride pyramid (PointCloud.h:325-343); 1 = voxel levels (extension of
                                this repository, include/icp_gpu.h ICP_GPU_PYRAMID_VOXEL; not reference behaviour) */
    float   weight_max_distance_sq; /* ICPOptimizer::maxDistance as handed to WeightingMethod (ICPOptimizer.h:220,528): kept apart from the
                                matcher's threshold by the reference (setMatchingMethod resets only the matcher, :71-78); 0 = max_distance_sq */
} orc_config;

void orc_default_config(orc_config* c);

/* utils.h:106-118 / :122-133 */
void orc_transform_points(const float pose[16], const float* pts, int64_t n, float* out);
void orc_transform_normals(const float pose[16], const float* nrm, int64_t n, float* out);

/* NearestNeighbor.h:81-97 tie rule + :181-186 threshold; exact 1-NN, lowest index on ties. */
void orc_knn3_brute(const float* tgt, int64_t nt, const float* qry, int64_t nq, float max_d2, orc_match* out);
/* NearestNeighborSearchBruteForce as written: compared and thresholded on the rounded norm (NearestNeighbor.h:81-97) */
void orc_knn3_brute_norm(const float* tgt, int64_t nt, const float* qry, int64_t nq, float max_d, orc_match* out);
void orc_knn6_brute(const float* tgt, const uint8_t* tgt_rgba, int64_t nt,
                    const float* qry, const uint8_t* qry_rgba, int64_t nq, float max_d2, orc_match* out);
/* Same answers as the brute-force versions, via an exact kd-tree (dim 3 or 6). */
typedef struct orc_kdtree orc_kdtree;
orc_kdtree* orc_kdtree_build(const float* tgt, const uint8_t* tgt_rgba /*NULL => 3-D*/, int64_t nt);
void orc_kdtree_query(const orc_kdtree* t, const float* qry, const uint8_t* qry_rgba, int64_t nq,
                      float max_d2, orc_match* out);
void orc_kdtree_free(orc_kdtree* t);

/* NearestNeighbor.h:333-421 */
void orc_projective(const float* tgt, uint32_t width, uint32_t height, float fx, float fy, float cx, float cy,
                    const float* qry, int64_t nq, float max_d2, orc_match* out);

/* weighting.h:39-99 */
void orc_apply_weights(int method, float max_d2, const float* sp, const float* tp, const float* sn, const float* tn,
                       const uint8_t* sc, const uint8_t* tc, int64_t n, orc_match* m);
/* ICPOptimizer.h:157-174 */
void orc_prune(const float* sn, const float* tn, int64_t n, orc_match* m);

/* One iteration's stages 2-4 for a given pose (teacher-forced parity): transform source by pose,
 * match, weight, reject.  sel_idx (nullable) lists the selected source indices (ascending); if
 * NULL all n_src points are used.  out has n_sel entries. Optional outputs tp_out/tn_out are the
 * transformed points / normals (3*n_sel floats each, nullable). */
int orc_match_pipeline(const orc_config* cfg, const float pose[16],
                       const float* src, const float* src_n, const uint8_t* src_c, int64_t n_src,
                       const int32_t* sel_idx, int64_t n_sel,
                       const float* tgt, const float* tgt_n, const uint8_t* tgt_c, int64_t n_tgt,
                       const orc_kdtree* tree /*nullable*/, orc_match* out, float* tp_out, float* tn_out);

/* Linear solvers on already-matched (gathered) arrays, ICPOptimizer.h:666-898 + ProcrustesAligner.h.
 * Return 0 on success, -1 when there are no matches (the reference hangs in ASSERT there). */
int orc_solve_p2p(const float* s, const float* d, const float* w, int64_t m, float out_pose[16]);
int orc_solve_p2plane(const float* s, const float* d, const float* n, const float* w, int64_t m, float out_pose[16]);
int orc_solve_symmetric(const float* s, const float* d, const float* ns, const float* nt, const float* w, int64_t m,
                        float out_pose[16]);
/* Ceres LM restatement (ICPOptimizer.h:283-310,352-482 + constraints.h + utils.h:25-102).
 * Works on the per-source-point arrays + matches (not gathered), like prepareConstraints*. */
int orc_solve_lm(int metric, const float* sp, const float* sn, const float* tgt, const float* tgt_n,
                 const orc_match* m, int64_t n, int max_iterations, double x_out[6], float out_pose[16],
                 int* n_lm_iterations);

/* selection.h:88-104 with std::mt19937 + libstdc++ uniform_real_distribution<double>. */
typedef struct { uint32_t mt[624]; int idx; } orc_mt19937;
void orc_mt_seed(orc_mt19937* r, uint32_t seed);
uint32_t orc_mt_next(orc_mt19937* r);
double orc_mt_canonical(orc_mt19937* r);
/* PointCloud.h:325-343: indices i=0,f,2f,.. with finite point & normal. Returns count. */
int64_t orc_coarse_indices(const float* pts, const float* nrm, int64_t n, int stride, int32_t* out_idx);
/* ICPOptimizer.h:503-516 */
int orc_coarsest_stride(int64_t n);
/* Voxel pyramid level (extension, not in the reference): the valid point with the lowest index of every occupied cell of
 * the uniform source grid at depth min(T,22) - ceil(1.5 log2 stride).  Restates icp_variants_b200/csrc/grid.cu
 * (grid_params_kernel, cell_code, voxel_*_kernel) operation by operation in fp32.  Ascending indices; returns the count. */
int64_t orc_voxel_indices(const float* pts, const float* nrm, int64_t n, int stride, int32_t* out_idx);

/* Whole registration: LinearICPOptimizer::estimatePose (ICPOptimizer.h:493-663) or
 * CeresICPOptimizer::estimatePose (:185-349).  pose_history (nullable) receives 16 floats per
 * executed iteration; n_iters_out the number executed (max(nIter, levels) in multires).
 * Returns 0, or -1 if an iteration had no surviving matches (pose left at the last good one). */
int orc_estimate_pose(const orc_config* cfg,
                      const float* src, const float* src_n, const uint8_t* src_c, int64_t n_src,
                      const float* tgt, const float* tgt_n, const uint8_t* tgt_c, int64_t n_tgt,
                      float pose_inout[16], float* pose_history, int* n_iters_out, int64_t* n_queries_out);

/* ConvergenceMeasure.h:50-66 */
float orc_rmse(const float pose[16], const float* src, const float* ref, int64_t n);
/* PointCloud(pcl cloud) (PointCloud.h:41-76): k-NN PCA normals (PCL NormalEstimation restated), viewpoint vp. */
void orc_pca_normals(const float* pts, int64_t n, int k, const float vp[3], float* out_nrm, float* out_curv);
/* ConvergenceMeasure.h:104-151 (Fontana benchmark error) */
double orc_benchmark_error(const float pose[16], const float* src, const float* ref, int64_t n);

/* PointCloud.h:78-165: depth map (+ RGBX frame) -> points, central-difference normals, colours; returns the count.
 * K column-major 3x3, E (nullable = identity) column-major 4x4 depth extrinsics. */
int64_t orc_cloud_from_depth(const float* depth, const uint8_t* color, const float K[9], const float E[16],
                             uint32_t width, uint32_t height, int keep_original_size, uint32_t downsample, float max_distance,
                             float* pts_out, float* nrm_out, uint8_t* rgba_out);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
