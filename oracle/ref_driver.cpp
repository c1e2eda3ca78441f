// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Compiles the reference's OWN headers, from where they lie (-I /root/reference/icp-variants),
// into oracle/_ref/libicp_ref.so and exposes them through a flat C ABI for ctypes
// (oracle/ref.py).  The third-party libraries those headers include (Eigen, FLANN, Ceres, PCL)
// are not available in this image; oracle/ref_shim/ holds small stand-ins for the API slices the
// headers use (see the header comment of each).  Therefore this library pins the reference's own
// code -- every line of NearestNeighbor.h, weighting.h, selection.h, utils.h, constraints.h,
// ProcrustesAligner.h, PointCloud.h, ConvergenceMeasure.h and ICPOptimizer.h that the path
// executes -- but not the last-bit behaviour of Eigen / FLANN / Ceres themselves.
//
// No reference source is copied: the headers are #included by path at build time.
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <limits>
#include <memory>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <math.h>
#include <assert.h>

#include <Eigen/Dense>
#include <ceres/ceres.h>
#include <ceres/rotation.h>
#include <flann/flann.hpp>
#include <pcl/point_types.h>

// The reference's ASSERT spins forever (Eigen.h:9) but is #ifndef-guarded: make it throw, keeping
// the reference's (unparenthesised) expansion of the condition.
struct ref_assert_failure { int line; };
#define ASSERT(a) { if (!a) { throw ref_assert_failure{__LINE__}; } }

// selection.h:76-79 seeds std::mt19937 from std::random_device; make that seed explicit.
static unsigned g_ref_seed = 0;
namespace std { struct ref_fixed_random_device { unsigned operator()() { return g_ref_seed; } }; }
#define random_device ref_fixed_random_device

// In-memory stand-in for VirtualSensor.h (a FreeImage-based TUM file reader, out of scope): the
// same accessors over caller-provided frames, the same constants (VirtualSensor.h:38-46).
typedef unsigned char BYTE;
class VirtualSensor {
public:
    float* depth = nullptr; BYTE* color = nullptr; unsigned w = 640, h = 480; int frame = 0;
    Eigen::Matrix3f K; Eigen::Matrix4f E;
    VirtualSensor() { K.setZero(); K(0, 0) = 525.0f; K(1, 1) = 525.0f; K(0, 2) = 319.5f; K(1, 2) = 239.5f; K(2, 2) = 1.0f; E.setIdentity(); }
    float* getDepth() { return depth; }
    BYTE* getColorRGBX() { return color; }
    Eigen::Matrix3f getDepthIntrinsics() { return K; }
    Eigen::Matrix4f getDepthExtrinsics() { return E; }
    Eigen::Matrix3f getColorIntrinsics() { return K; }
    Eigen::Matrix4f getColorExtrinsics() { return E; }
    unsigned getDepthImageWidth() { return w; }
    unsigned getDepthImageHeight() { return h; }
    unsigned getColorImageWidth() { return w; }
    unsigned getColorImageHeight() { return h; }
    int getCurrentFrameCnt() { return frame; }
};

// The solvers and pruneCorrespondences are private/protected members; open them for the tests.
#define private public
#define protected public
#include "ICPOptimizer.h"
#undef private
#undef protected
#undef random_device

namespace {
struct Quiet { Quiet() { std::cout.rdbuf(nullptr); } } g_quiet;  // the reference logs every stage to stdout

std::vector<Vector3f> vec3(const float* p, int64_t n) {
    std::vector<Vector3f> v((size_t)n);
    for (int64_t i = 0; i < n; ++i) v[i] = Vector3f(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    return v;
}
std::vector<Vector4uc> vec4uc(const uint8_t* p, int64_t n) {
    std::vector<Vector4uc> v((size_t)n, Vector4uc::Zero());
    if (p) for (int64_t i = 0; i < n; ++i) v[i] = Vector4uc(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
    return v;
}
void put3(const std::vector<Vector3f>& v, float* out) { for (size_t i = 0; i < v.size(); ++i) { out[3 * i] = v[i][0]; out[3 * i + 1] = v[i][1]; out[3 * i + 2] = v[i][2]; } }
Matrix4f mat4(const float* p) { Matrix4f m; for (int i = 0; i < 16; ++i) m.data()[i] = p[i]; return m; }
void put4(const Matrix4f& m, float* out) { for (int i = 0; i < 16; ++i) out[i] = m.data()[i]; }
void put_matches(const std::vector<Match>& m, int32_t* idx, float* w) { for (size_t i = 0; i < m.size(); ++i) { idx[i] = m[i].idx; w[i] = m[i].weight; } }
std::vector<Match> get_matches(const int32_t* idx, const float* w, int64_t n) { std::vector<Match> m((size_t)n); for (int64_t i = 0; i < n; ++i) { m[i].idx = idx[i]; m[i].weight = w[i]; } return m; }

PointCloud make_cloud(const float* pts, const float* nrm, const uint8_t* rgba, int64_t n) {
    PointCloud c;
    c.getPoints() = vec3(pts, n);
    c.getNormals() = nrm ? vec3(nrm, n) : std::vector<Vector3f>((size_t)n, Vector3f::Zero());
    c.getColors() = vec4uc(rgba, n);
    return c;
}
}  // namespace

extern "C" {

typedef struct {
    int32_t minimizer;       // 0 LinearICPOptimizer, 1 CeresICPOptimizer
    int32_t metric, selection, weighting, rejection, matching, color_icp, multires, n_iterations;
    double proba; uint32_t seed;
    float max_distance_sq;
    float K[9];              // column-major 3x3 (Eigen Matrix3f storage), projective only
    uint32_t width, height;
    int32_t setter_order;    // 0: setMatchingMethod, then setMatchingMaxDistance (main.cpp's order: both distances = max_distance_sq);
                             // 1: setMatchingMaxDistance first -- setMatchingMethod then re-creates the matcher with MAX_DISTANCE and only
                             //    WeightingMethod keeps max_distance_sq; 2: setMatchingMaxDistance never called (0.005 / 0.0003)
} ref_config;

// utils.h:106-118 / :122-133
void ref_transform_points(const float* pose, const float* pts, int64_t n, float* out) { put3(transformPoints(vec3(pts, n), mat4(pose)), out); }
void ref_transform_normals(const float* pose, const float* nrm, int64_t n, float* out) { put3(transformNormals(vec3(nrm, n), mat4(pose)), out); }

// NearestNeighbor.h:104-314 (FLANN front-end over the exact stand-in), 3-D or 6-D
int ref_knn_flann(const float* tgt, const uint8_t* tgt_rgba, int64_t nt, const float* qry, const uint8_t* qry_rgba, int64_t nq,
                  float max_d2, int32_t* idx, float* w) {
    NearestNeighborSearchFlann nn; nn.setMatchingMaxDistance(max_d2);
    std::vector<Match> m;
    if (tgt_rgba) { nn.buildIndex(vec3(tgt, nt), vec4uc(tgt_rgba, nt)); m = nn.queryMatches(vec3(qry, nq), vec4uc(qry_rgba, nq)); }
    else { nn.buildIndex(vec3(tgt, nt)); m = nn.queryMatches(vec3(qry, nq)); }
    if ((int64_t)m.size() != nq) return -1;
    put_matches(m, idx, w); return 0;
}
// NearestNeighbor.h:42-98
int ref_knn_brute(const float* tgt, int64_t nt, const float* qry, int64_t nq, float max_d, int32_t* idx, float* w) {
    NearestNeighborSearchBruteForce nn; nn.setMatchingMaxDistance(max_d);
    nn.buildIndex(vec3(tgt, nt));
    std::vector<Match> m = nn.queryMatches(vec3(qry, nq));
    if ((int64_t)m.size() != nq) return -1;
    put_matches(m, idx, w); return 0;
}
// NearestNeighbor.h:317-444
int ref_projective(const float* tgt, uint32_t width, uint32_t height, const float* K9, const float* qry, int64_t nq, float max_d2,
                   int32_t* idx, float* w) {
    NearestNeighborSearchProjective nn; nn.setMatchingMaxDistance(max_d2);
    Matrix3f K; for (int i = 0; i < 9; ++i) K.data()[i] = K9[i];
    nn.setCameraParams(K, width, height);
    nn.buildIndex(vec3(tgt, (int64_t)width * height));
    std::vector<Match> m = nn.queryMatches(vec3(qry, nq));
    if ((int64_t)m.size() != nq) return -1;
    put_matches(m, idx, w); return 0;
}
// weighting.h:39-99
void ref_apply_weights(int method, float max_d2, const float* sp, const float* sn, const uint8_t* sc, int64_t n,
                       const float* tp, const float* tn, const uint8_t* tc, int64_t nt, int32_t* idx, float* w) {
    WeightingMethod wm(method, max_d2);
    std::vector<Match> m = get_matches(idx, w, n);
    wm.applyWeights(vec3(sp, n), vec3(tp, nt), vec3(sn, n), vec3(tn, nt), vec4uc(sc, n), vec4uc(tc, nt), m);
    put_matches(m, idx, w);
}
// ICPOptimizer.h:157-174
void ref_prune(const float* sn, int64_t n, const float* tn, int64_t nt, int32_t* idx, float* w) {
    LinearICPOptimizer opt;
    std::vector<Match> m = get_matches(idx, w, n);
    opt.pruneCorrespondences(vec3(sn, n), vec3(tn, nt), m);
    put_matches(m, idx, w);
}
// ICPOptimizer.h:666-898 + ProcrustesAligner.h on gathered pairs
int ref_solve_linear(int metric, const float* s, const float* d, const float* ns, const float* nt, const float* w, int64_t m, float* out16) {
    try {
        LinearICPOptimizer opt;
        std::vector<float> wv(w, w + m);
        Matrix4f P;
        if (metric == 0) P = opt.estimatePosePointToPoint(vec3(s, m), vec3(d, m), wv);
        else if (metric == 1) P = opt.estimatePosePointToPlane(vec3(s, m), vec3(d, m), vec3(nt, m), wv);
        else P = opt.estimatePoseSymmetricICP(vec3(s, m), vec3(d, m), vec3(ns, m), vec3(nt, m), wv);
        put4(P, out16); return 0;
    } catch (const ref_assert_failure&) { return -2; }
}
// constraints.h functors evaluated in double at the increment x (residual values only)
int ref_residuals(int kind /*0 p2p,1 plane,2 symmetric*/, const double* x, const float* s, const float* d, const float* ns, const float* nt, float w, double* out) {
    Vector3f S(s[0], s[1], s[2]), D(d[0], d[1], d[2]);
    if (kind == 0) { PointToPointConstraint c(S, D, w); c(x, out); return 3; }
    Vector3f NT(nt[0], nt[1], nt[2]);
    if (kind == 1) { PointToPlaneConstraint c(S, D, NT, w); c(x, out); return 1; }
    Vector3f NS(ns[0], ns[1], ns[2]);
    SymmetricConstraint c(S, D, NS, NT, w); c(x, out); return 1;
}
// utils.h:79-98
void ref_increment_to_matrix(const double* x6, float* out16) { double x[6]; std::memcpy(x, x6, sizeof(x)); put4(PoseIncrement<double>::convertToMatrix(PoseIncrement<double>(x)), out16); }

// PointCloud.h:325-343; returns the number of points kept
int64_t ref_coarse_resolution(const float* pts, const float* nrm, const uint8_t* rgba, int64_t n, int factor, float* pts_out, float* nrm_out, uint8_t* rgba_out) {
    PointCloud c = make_cloud(pts, nrm, rgba, n).getCoarseResolution(factor);
    put3(c.getPoints(), pts_out); put3(c.getNormals(), nrm_out);
    for (size_t i = 0; i < c.getColors().size(); ++i) for (int k = 0; k < 4; ++k) rgba_out[4 * i + k] = c.getColors()[i][k];
    return (int64_t)c.getPoints().size();
}
// selection.h: n_resamples calls of resample(); returns the size of the last sample
int64_t ref_selection(const float* pts, const float* nrm, const uint8_t* rgba, int64_t n, double proba, uint32_t seed, int n_resamples,
                      float* pts_out, float* nrm_out, int64_t* n_colors_out) {
    g_ref_seed = seed;
    PointSelection sel(make_cloud(pts, nrm, rgba, n), RANDOM_SAMPLING, (float)proba);
    for (int i = 0; i < n_resamples; ++i) sel.resample();
    put3(sel.getPoints(), pts_out); put3(sel.getNormals(), nrm_out);
    if (n_colors_out) *n_colors_out = (int64_t)sel.getColors().size();  // grows without bound: selection.h never clears m_colors
    return (int64_t)sel.getPoints().size();
}
// ConvergenceMeasure.h:50-66, :104-151
float ref_rmse(const float* pose, const float* src, const float* ref, int64_t n) { ConvergenceMeasure cm(vec3(src, n), vec3(ref, n)); return cm.rmseAlignmentError(mat4(pose)); }
double ref_benchmark_error(const float* pose, const float* src, const float* ref, int64_t n) { ConvergenceMeasure cm(vec3(src, n), vec3(ref, n), true); return cm.benchmarkError(mat4(pose)); }

// SimpleMesh::loadMesh (SimpleMesh.h:161-229) + PointCloud(const SimpleMesh&) (PointCloud.h:12-39)
int64_t ref_cloud_from_off(const char* path, int64_t cap, float* pts_out, float* nrm_out) {
    SimpleMesh mesh;
    if (!mesh.loadMesh(path)) return -1;
    PointCloud c{mesh};
    if ((int64_t)c.getPoints().size() > cap) return -(int64_t)c.getPoints().size();
    put3(c.getPoints(), pts_out); put3(c.getNormals(), nrm_out);
    return (int64_t)c.getPoints().size();
}
// PointCloud(float* depthMap, BYTE* colorFrame, ...) (PointCloud.h:78-165); returns the number of points
int64_t ref_cloud_from_depth(float* depth, uint8_t* color, const float* K9, const float* E16, uint32_t width, uint32_t height, int keep_original_size,
                             uint32_t downsample, float max_distance, float* pts_out, float* nrm_out, uint8_t* rgba_out) {
    Matrix3f K; for (int i = 0; i < 9; ++i) K.data()[i] = K9[i];
    PointCloud c(depth, color, K, mat4(E16), width, height, keep_original_size != 0, downsample, max_distance);
    put3(c.getPoints(), pts_out); put3(c.getNormals(), nrm_out);
    for (size_t i = 0; i < c.getColors().size(); ++i) for (int k = 0; k < 4; ++k) rgba_out[4 * i + k] = c.getColors()[i][k];
    return (int64_t)c.getPoints().size();
}
// PointCloud(pcl cloud) (PointCloud.h:41-76): k = 5 normals through the PCL stand-in
int64_t ref_cloud_from_xyz(const float* pts, int64_t n, float* nrm_out, uint8_t* rgba_out) {
    pcl::PointCloud<pcl::PointXYZ>::Ptr src(new pcl::PointCloud<pcl::PointXYZ>());
    src->points.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) { src->points[i].x = pts[3 * i]; src->points[i].y = pts[3 * i + 1]; src->points[i].z = pts[3 * i + 2]; }
    PointCloud c(src);
    put3(c.getNormals(), nrm_out);
    for (size_t i = 0; i < c.getColors().size(); ++i) for (int k = 0; k < 4; ++k) rgba_out[4 * i + k] = c.getColors()[i][k];
    return (int64_t)c.getPoints().size();
}

// LinearICPOptimizer::estimatePose (ICPOptimizer.h:493-663) / CeresICPOptimizer::estimatePose (:185-349).
// rmse_out receives one value per executed iteration (ConvergenceMeasure over gt_src/gt_ref);
// returns the number of iterations executed, -2 if the reference's ASSERT fired (no matches).
int ref_estimate_pose(const ref_config* cfg,
                      const float* src, const float* src_n, const uint8_t* src_c, int64_t n_src,
                      const float* tgt, const float* tgt_n, const uint8_t* tgt_c, int64_t n_tgt,
                      const float* gt_src, const float* gt_ref, int64_t n_gt,
                      float* pose_inout, float* rmse_out, int rmse_cap, double* stage_times_out /*6, nullable*/) {
    try {
        g_ref_seed = cfg->seed;
        std::unique_ptr<ICPOptimizer> opt;
        if (cfg->minimizer == 0) opt.reset(new LinearICPOptimizer()); else opt.reset(new CeresICPOptimizer());
        if (cfg->setter_order == 1) opt->setMatchingMaxDistance(cfg->max_distance_sq);
        opt->setMatchingMethod((unsigned)cfg->matching);   // replaces the matcher (ICPOptimizer.h:71-78)
        opt->setMetric((unsigned)cfg->metric);
        opt->setNbOfIterations((unsigned)cfg->n_iterations);
        if (cfg->setter_order == 0) opt->setMatchingMaxDistance(cfg->max_distance_sq);
        opt->setSelectionMethod((unsigned)cfg->selection, cfg->proba);
        opt->setRejectionMethod((unsigned)cfg->rejection);
        opt->setWeightingMethod((unsigned)cfg->weighting);
        opt->enableMultiResolution(cfg->multires != 0);
        opt->enableColorICP(cfg->color_icp != 0);
        if (cfg->matching == 1) { Matrix3f K; for (int i = 0; i < 9; ++i) K.data()[i] = cfg->K[i]; opt->setCameraParamsMatchingMethod(K, cfg->width, cfg->height); }
        PointCloud source = make_cloud(src, src_n, src_c, n_src), target = make_cloud(tgt, tgt_n, tgt_c, n_tgt);
        ConvergenceMeasure cm(vec3(gt_src, n_gt), vec3(gt_ref, n_gt));
        TimeMeasure tm;
        opt->setConvergenceMeasure(cm); opt->setTimeMeasure(tm);
        Matrix4f pose = mat4(pose_inout);
        opt->estimatePose(source, target, pose, true);
        put4(pose, pose_inout);
        const int n_it = (int)cm.iterationErrorsRMSE.size();
        for (int i = 0; i < n_it && i < rmse_cap; ++i) rmse_out[i] = cm.iterationErrorsRMSE[i];
        if (stage_times_out) { stage_times_out[0] = tm.selectionTime; stage_times_out[1] = tm.matchingTime; stage_times_out[2] = tm.weighingTime; stage_times_out[3] = tm.rejectionTime; stage_times_out[4] = tm.solverTime; stage_times_out[5] = tm.convergenceTime; }
        return n_it;
    } catch (const ref_assert_failure&) { return -2; }
}

// threads of the stand-in matcher's OpenMP loop (the reference's own loop is single-threaded apart from Ceres)
void ref_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}
int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// test hook: answer FLANN searches by the literal O(N*M) scan instead of the exact kd-tree (identical results)
void ref_set_flann_exhaustive(int on) { flann::exhaustive() = on != 0; }

const char* ref_describe(void) {
    return "reference headers from /root/reference/icp-variants compiled in place against oracle/ref_shim stand-ins "
           "(Eigen/FLANN/Ceres/PCL are not installed): pins the reference's own code, not the third-party numerics";
}
}  // extern "C"
