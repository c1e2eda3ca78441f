// ConvergenceMeasure.h -- RMSE over known correspondences after every iteration
// (reference: icp-variants/ConvergenceMeasure.h:15-78; fed from the loop at ICPOptimizer.h:629-631).
// Inside estimatePose the errors of all iterations are evaluated on the device from the loop's pose history
// (icp_gpu_set_correspondences / icp_gpu_convergence_errors) and stored here; recordAlignmentError(pose) remains
// for callers that evaluate a pose of their own.
#pragma once
#include "Eigen.h"

class ConvergenceMeasure {
public:
    ConvergenceMeasure() {}
    ConvergenceMeasure(const std::vector<Vector3f>& sourceCorrespondences, const std::vector<Vector3f>& targetCorrespondences, const bool runBenchmark = false)
        : m_source(sourceCorrespondences), m_target(targetCorrespondences), m_runBenchmark(runBenchmark) {}

    // used by ICPOptimizer::run to hand the correspondences to the device and to store what it computed
    const std::vector<Vector3f>& sourcePoints() const { return m_source; }
    const std::vector<Vector3f>& unchangedPoints() const { return m_target; }
    bool runBenchmark() const { return m_runBenchmark; }
    void recordDeviceErrors(float rmse, double benchmark) { m_rmse.push_back(rmse); if (m_runBenchmark) m_benchmark.push_back((float)benchmark); }
    float getFinalErrorRMSE() const { return m_rmse.back(); }                 // ConvergenceMeasure.h:176-178
    float getFinalErrorBenchmark() const { return m_benchmark.back(); }       // ConvergenceMeasure.h:180-182
    const std::vector<float>& getBenchmark() const { return m_benchmark; }

    void recordAlignmentError(const Matrix4f& pose) {   // ConvergenceMeasure.h:50-78
        int counter = 0; float rmse = 0.f;
        for (size_t i = 0; i < m_source.size() && i < m_target.size(); ++i) {
            const Vector3f& s = m_source[i]; const Vector3f& t = m_target[i];
            const float x = ((pose(0, 0) * s[0] + pose(0, 1) * s[1]) + pose(0, 2) * s[2]) + pose(0, 3);
            const float y = ((pose(1, 0) * s[0] + pose(1, 1) * s[1]) + pose(1, 2) * s[2]) + pose(1, 3);
            const float z = ((pose(2, 0) * s[0] + pose(2, 1) * s[1]) + pose(2, 2) * s[2]) + pose(2, 3);
            if (std::isfinite(x) && std::isfinite(y) && std::isfinite(z) && std::isfinite(t[0]) && std::isfinite(t[1]) && std::isfinite(t[2])) {
                const float dx = x - t[0], dy = y - t[1], dz = z - t[2];
                rmse += (dx * dx + dy * dy) + dz * dz; ++counter;
            }
        }
        m_rmse.push_back(counter ? std::sqrt(rmse / counter) : 0.f);
    }
    const std::vector<float>& getRMSE() const { return m_rmse; }

private:
    std::vector<Vector3f> m_source, m_target;
    std::vector<float> m_rmse, m_benchmark;
    bool m_runBenchmark = false;
};
