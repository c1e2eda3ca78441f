// ConvergenceMeasure.h -- RMSE / benchmark error over known correspondences (drop-in for icp-variants/ConvergenceMeasure.h:15-184;
// fed from the loop at ICPOptimizer.h:629-631).  Inside estimatePose the errors of ALL iterations are evaluated on the device from the
// loop's pose history (icp_gpu_set_correspondences / icp_gpu_convergence_errors) and stored here; rmseAlignmentError(pose) /
// benchmarkError(pose) / recordAlignmentError(pose) for a pose of the caller's own go through icp_gpu_alignment_error.
#pragma once
#include <cstdio>
#include <fstream>
#include <string>
#include "Eigen.h"
#include "detail.h"

class ConvergenceMeasure {
public:
    ConvergenceMeasure() {}
    ConvergenceMeasure(const std::vector<Vector3f>& sourcePoints, const std::vector<Vector3f>& unchangedPoints, const bool runBenchmark = false)
        : m_sourcePoints(sourcePoints), m_unchangedPoints(unchangedPoints), m_runBenchmark(runBenchmark) {
        if (sourcePoints.size() != unchangedPoints.size() || sourcePoints.empty())      // the reference ASSERTs (and hangs) here, :34-35
            std::cout << "ConvergenceMeasure: the number of points must be the same and > 0." << std::endl;
    }

    // used by ICPOptimizer::run to hand the correspondences to the device and to store what it computed
    const std::vector<Vector3f>& sourcePoints() const { return m_sourcePoints; }
    const std::vector<Vector3f>& unchangedPoints() const { return m_unchangedPoints; }
    bool runBenchmark() const { return m_runBenchmark; }
    void recordDeviceErrors(float rmse, double benchmark) {
        iterationErrorsRMSE.push_back(rmse);
        if (m_runBenchmark) iterationErrorsBenchmark.push_back((float)benchmark);
    }

    float rmseAlignmentError(const Matrix4f& pose) {                          // :50-66
        float rmse = 0.f; double bench = 0.0;
        errorsOf(pose, rmse, nullptr, bench);
        return rmse;
    }
    double benchmarkError(const Matrix4f& pose) {                             // :104-151
        float rmse = 0.f; double bench = 0.0;
        errorsOf(pose, rmse, &bench, bench);
        return bench;
    }
    void recordAlignmentError(const Matrix4f& pose) {                         // :69-78
        float rmse = 0.f; double bench = 0.0;
        errorsOf(pose, rmse, m_runBenchmark ? &bench : nullptr, bench);
        std::cout << "RMSE Alignment errors: " << rmse << "\n";
        iterationErrorsRMSE.push_back(rmse);
        if (m_runBenchmark) { std::cout << "Benchmark errors: " << (float)bench << "\n"; iterationErrorsBenchmark.push_back((float)bench); }
    }
    void outputAlignmentError() {                                             // :81-102
        if (iterationErrorsRMSE.empty()) { std::cout << "No recorded alignment error.\n"; return; }
        std::cout << "Recorded RMSE Alginment Error!\n\tIter \t RMSE Error\n";
        for (size_t i = 0; i < iterationErrorsRMSE.size(); i++) std::printf("\t%02d \t %01.6f\n", (int)i, iterationErrorsRMSE[i]);
        if (m_runBenchmark) {
            if (iterationErrorsBenchmark.empty()) { std::cout << "No recorded alignment error for benchmark.\n"; return; }
            std::cout << "Recorded benchmark Alginment Error!\n\tIter \t Benchmark Error\n";
            for (size_t i = 0; i < iterationErrorsBenchmark.size(); i++) std::printf("\t%02d \t %01.6f\n", (int)i, iterationErrorsBenchmark[i]);
        }
    }
    void writeRMSEToFile(std::string nameFile) { writeTo(nameFile, iterationErrorsRMSE); }                 // :153-163
    void writeBenchmarkToFile(std::string nameFile) { writeTo(nameFile, iterationErrorsBenchmark); }       // :165-175
    float getFinalErrorRMSE() const { return iterationErrorsRMSE.back(); }                                  // :177-179
    float getFinalErrorBenchmark() const { return iterationErrorsBenchmark.back(); }                        // :181-183
    const std::vector<float>& getRMSE() const { return iterationErrorsRMSE; }
    const std::vector<float>& getBenchmark() const { return iterationErrorsBenchmark; }

private:
    std::vector<Vector3f> m_sourcePoints, m_unchangedPoints;
    std::vector<float> iterationErrorsRMSE, iterationErrorsBenchmark;
    bool m_runBenchmark = false;

    void errorsOf(const Matrix4f& pose, float& rmse, double* benchOrNull, double& bench) {
        icp_gpu_ctx* ctx = icp_b200::sharedContext();
        if (!ctx || m_sourcePoints.empty() || m_sourcePoints.size() != m_unchangedPoints.size()) return;
        int rc = icp_gpu_set_correspondences(ctx, reinterpret_cast<const float*>(m_sourcePoints.data()), reinterpret_cast<const float*>(m_unchangedPoints.data()),
                                             (int64_t)m_sourcePoints.size());
        if (rc == ICP_GPU_OK) rc = icp_gpu_alignment_error(ctx, pose.data(), &rmse, benchOrNull ? &bench : nullptr);
        icp_b200::report(ctx, rc, "ConvergenceMeasure");
    }
    static void writeTo(const std::string& nameFile, const std::vector<float>& v) {
        std::ofstream f(nameFile);
        for (float x : v) f << x << std::endl;
    }
};
