// bunny_driver.cpp -- a driver written against the reference's API (compare alignBunnyWithICP,
// icp-variants/main.cpp:43-181) compiled against the drop-in headers.  Input: two raw clouds
// (int32 n, then n*3 float points, n*3 float normals), output: the 16 floats of the estimated pose
// (column-major) on stdout.   usage: bunny_driver <source.bin> <target.bin> <linear 0|1> <metric> <iterations>
#include <cstdio>
#include <cstdlib>
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

int main(int argc, char** argv) {
    if (argc < 6) { std::fprintf(stderr, "usage: %s source.bin target.bin linear metric iterations\n", argv[0]); return 2; }
    PointCloud source, target;
    if (!readCloud(argv[1], source) || !readCloud(argv[2], target)) { std::fprintf(stderr, "cannot read clouds\n"); return 2; }
    ICPOptimizer* optimizer = nullptr;
    if (std::atoi(argv[3])) optimizer = new LinearICPOptimizer(); else optimizer = new CeresICPOptimizer();
    optimizer->setMatchingMethod(0);                       // main.cpp:74
    optimizer->setMatchingMaxDistance(0.0003f);            // main.cpp:75
    optimizer->setMetric((unsigned)std::atoi(argv[4]));
    optimizer->setNbOfIterations((unsigned)std::atoi(argv[5]));
    optimizer->setSelectionMethod(SELECT_ALL);
    optimizer->setWeightingMethod(CONSTANT_WEIGHTING);
    // the hand-picked ground-truth correspondences of the bunny pair (main.cpp:105-120), when the clouds are large enough
    const int gtSource[4] = {215, 424, 640, 1023}, gtTarget[4] = {294, 258, 1238, 1310};
    std::vector<Vector3f> gs, gt;
    if (source.getPoints().size() > 1023 && target.getPoints().size() > 1310)
        for (int i = 0; i < 4; ++i) { gs.push_back(source.getPoints()[(size_t)gtSource[i]]); gt.push_back(target.getPoints()[(size_t)gtTarget[i]]); }
    TimeMeasure timeMeasure; ConvergenceMeasure convergenceMeasure(gs, gt, true);
    optimizer->setTimeMeasure(timeMeasure);
    optimizer->setConvergenceMeasure(convergenceMeasure);
    Matrix4f estimatedPose = Matrix4f::Identity();
    optimizer->estimatePose(source, target, estimatedPose);
    std::printf("POSE");
    for (int i = 0; i < 16; ++i) std::printf(" %.9g", estimatedPose.data()[i]);
    std::printf("\n");
    if (!convergenceMeasure.getRMSE().empty())
        std::printf("RMSE %.9g BENCHMARK %.9g ITERATIONS %d\n", convergenceMeasure.getFinalErrorRMSE(), convergenceMeasure.getFinalErrorBenchmark(),
                    (int)convergenceMeasure.getRMSE().size());
    delete optimizer;
    return 0;
}
