// value_classes_driver.cpp -- the reference API's value classes as a reference user holds them (selection.h, weighting.h, constraints.h,
// ProcrustesAligner.h, utils.h, PointCloud.h, ConvergenceMeasure.h), compiled against the drop-in headers and run on the device.
// Input: two raw clouds (int32 n, then n*3 float points, n*3 float normals).  Output: one line per class, compared by
// tests/test_cpp_dropin.py with the oracle / the Python binding.   usage: value_classes_driver <source.bin> <target.bin>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "icp_b200/ICPOptimizer.h"

static bool readCloud(const char* path, PointCloud& out) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    int32_t n = 0;
    if (std::fread(&n, 4, 1, f) != 1 || n < 0) { std::fclose(f); return false; }
    std::vector<Vector3f> p((size_t)n), m((size_t)n);
    bool ok = std::fread(p.data(), 12, (size_t)n, f) == (size_t)n && std::fread(m.data(), 12, (size_t)n, f) == (size_t)n;
    std::fclose(f);
    if (ok) out = PointCloud(p, m);
    return ok;
}
// order-independent fingerprint of the bit patterns
static unsigned long long bits(const std::vector<Vector3f>& v) {
    unsigned long long h = 0;
    for (size_t i = 0; i < v.size(); ++i) for (int k = 0; k < 3; ++k) { unsigned int u; float f = v[i][k]; std::memcpy(&u, &f, 4); h += (unsigned long long)u * (2 * (3 * i + k) + 1); }
    return h;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s source.bin target.bin\n", argv[0]); return 2; }
    PointCloud source, target;
    if (!readCloud(argv[1], source) || !readCloud(argv[2], target)) { std::fprintf(stderr, "cannot read clouds\n"); return 2; }
    // a small rigid motion as PoseIncrement builds it (utils.h:75-99)
    double x[6] = {0.01, -0.02, 0.03, 0.001, 0.002, -0.001};
    PoseIncrement<double> inc(x);
    const Matrix4f pose = PoseIncrement<double>::convertToMatrix(inc);
    std::printf("POSE"); for (int i = 0; i < 16; ++i) std::printf(" %.9g", pose.data()[i]); std::printf("\n");
    // transformPoints / transformNormals (utils.h:106-133)
    const std::vector<Vector3f> q = transformPoints(source.getPoints(), pose), qn = transformNormals(source.getNormals(), pose);
    std::printf("TRANSFORM %llu %llu\n", bits(q), bits(qn));
    // NearestNeighborSearchFlann + WeightingMethod (NearestNeighbor.h:104-207, weighting.h:39-99)
    NearestNeighborSearchFlann nn;
    nn.setMatchingMaxDistance(0.0003f);
    nn.buildIndex(target.getPoints());
    std::vector<Match> matches = nn.queryMatches(q);
    WeightingMethod weighting(DISTANCES_WEIGHTING, 0.0003f);
    weighting.applyWeights(q, target.getPoints(), qn, target.getNormals(), source.getColors(), target.getColors(), matches);
    long long nMatched = 0, idxSum = 0; double wSum = 0.0;
    std::vector<Vector3f> s, d; std::vector<float> w;
    for (size_t i = 0; i < matches.size(); ++i) if (matches[i].idx >= 0) {
        ++nMatched; idxSum += matches[i].idx; wSum += matches[i].weight;
        s.push_back(q[i]); d.push_back(target.getPoints()[(size_t)matches[i].idx]); w.push_back(matches[i].weight);
    }
    std::printf("MATCH %lld %lld %.9g\n", nMatched, idxSum, wSum);
    // ProcrustesAligner (ProcrustesAligner.h:6-29)
    ProcrustesAligner aligner;
    const Matrix4f p2p = aligner.estimatePose(s, d, w);
    std::printf("PROCRUSTES"); for (int i = 0; i < 16; ++i) std::printf(" %.9g", p2p.data()[i]); std::printf("\n");
    // PointSelection (selection.h:12-107) with an explicit seed
    PointSelection selection(source, RANDOM_SAMPLING, 0.25f, 42u);
    selection.resample();
    long long selSum = 0; for (int i : selection.getSelectedIndexes()) selSum += i;
    std::printf("SELECTION %zu %lld\n", selection.getPoints().size(), selSum);
    // the three functors (constraints.h:9-143) at the increment x
    double r[3], rp[1], rs[1];
    PointToPointConstraint(source.getPoints()[0], target.getPoints()[0], 0.7f)(x, r);
    PointToPlaneConstraint(source.getPoints()[0], target.getPoints()[0], target.getNormals()[0], 0.7f)(x, rp);
    SymmetricConstraint(source.getPoints()[0], target.getPoints()[0], source.getNormals()[0], target.getNormals()[0], 0.7f)(x, rs);
    std::printf("FUNCTORS %.17g %.17g %.17g %.17g %.17g\n", r[0], r[1], r[2], rp[0], rs[0]);
    // PointCloud::change_pose (PointCloud.h:277-283) and ConvergenceMeasure::rmseAlignmentError / benchmarkError
    PointCloud moved = source.copy_point_cloud();
    moved.change_pose(pose);
    std::printf("CHANGEPOSE %llu\n", bits(moved.getPoints()));
    std::vector<Vector3f> gs(source.getPoints().begin(), source.getPoints().begin() + 64), gt(target.getPoints().begin(), target.getPoints().begin() + 64);
    ConvergenceMeasure cm(gs, gt, true);
    std::printf("ERRORS %.9g %.17g\n", cm.rmseAlignmentError(pose), cm.benchmarkError(pose));
    return 0;
}
