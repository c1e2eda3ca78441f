// ProcrustesAligner.h -- drop-in for icp-variants/ProcrustesAligner.h:4-73: the rigid pose between matched point sets (unweighted means,
// rotation from the SVD of sum_i (t_i - t_mean) (w_i (s_i - s_mean))^T, det-corrected; translation R (t_mean - s_mean) - R t_mean + t_mean).
// One call into icp_gpu_solve_linear (metric 0): the moments are accumulated in fp64 on the device (the reference does it in fp32), the
// 3x3 SVD is a one-sided Jacobi in fp64.
#pragma once
#include "Eigen.h"
#include "detail.h"

class ProcrustesAligner {
public:
    Matrix4f estimatePose(const std::vector<Vector3f>& sourcePoints, const std::vector<Vector3f>& targetPoints, const std::vector<float>& weights) {
        Matrix4f pose = Matrix4f::Identity();
        if (sourcePoints.size() != targetPoints.size() || sourcePoints.empty() || (!weights.empty() && weights.size() != sourcePoints.size())) {
            std::cout << "ProcrustesAligner: the number of source points, target points and weights must agree and be positive." << std::endl;   // :8 ASSERT
            return pose;
        }
        icp_gpu_ctx* ctx = icp_b200::sharedContext();
        if (!ctx) return pose;
        icp_b200::report(ctx, icp_gpu_solve_linear(ctx, ICP_GPU_METRIC_P2P, reinterpret_cast<const float*>(sourcePoints.data()), nullptr,
                                                   reinterpret_cast<const float*>(targetPoints.data()), nullptr, weights.empty() ? nullptr : weights.data(),
                                                   (int64_t)sourcePoints.size(), pose.data()), "ProcrustesAligner::estimatePose");
        return pose;
    }
};
