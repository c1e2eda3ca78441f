// weighting.h -- drop-in for icp-variants/weighting.h:8-99.  Inside estimatePose the weights are computed by the matching / reduction
// kernels; WeightingMethod::applyWeights is the same device function behind icp_gpu_apply_weights for callers that hold the class:
//   CONSTANT   matches untouched (:44)                    DISTANCES  1 - |s - t|^2 / maxDistance, 0 for non-finite points (:16-20,:56-66)
//   NORMALS    n_s . n_t (unclamped), 0 if non-finite (:22-25,:70-79)     COLORS   the distance weight x (1 - |c_s - c_t|^2 / 195075) with the
//   uchar wrap-around of the colour difference (:27-30,:82-87)
#pragma once
#include <vector>
#include "NearestNeighbor.h"
#include "PointCloud.h"
#include "detail.h"

#define MAX_COLOR_DIFFERENCE 195075

enum weighting_methods { CONSTANT_WEIGHTING = 0, DISTANCES_WEIGHTING, NORMALS_WEIGHTING, COLORS_WEIGHTING };

class WeightingMethod {
public:
    WeightingMethod(int method = DISTANCES_WEIGHTING, float maxDistance = 0.0003f) : method(method), maxDistance(maxDistance) {}

    // sourcePoints / sourceNormals are the TRANSFORMED source (one entry per match); the target arrays are indexed by matches[i].idx
    void applyWeights(const std::vector<Vector3f>& sourcePoints, const std::vector<Vector3f>& targetPoints, const std::vector<Vector3f>& sourceNormals,
                      const std::vector<Vector3f>& targetNormals, const std::vector<Vector4uc>& sourceColors, const std::vector<Vector4uc>& targetColors,
                      std::vector<Match>& matches) {
        if (method == CONSTANT_WEIGHTING || matches.empty()) return;
        icp_gpu_ctx* ctx = icp_b200::sharedContext();
        if (!ctx) return;
        const size_t n = matches.size();
        std::vector<int32_t> idx(n); std::vector<float> w(n);
        for (size_t i = 0; i < n; ++i) { idx[i] = matches[i].idx; w[i] = matches[i].weight; }
        const bool sn = sourceNormals.size() >= n, sc = sourceColors.size() >= n;
        const bool tn = targetNormals.size() == targetPoints.size() && !targetNormals.empty(), tc = targetColors.size() == targetPoints.size() && !targetColors.empty();
        if (sourcePoints.size() < n) { std::cout << "WeightingMethod: fewer source points than matches." << std::endl; return; }
        if (!icp_b200::report(ctx, icp_gpu_apply_weights(ctx, method, maxDistance, reinterpret_cast<const float*>(sourcePoints.data()),
                                                         sn ? reinterpret_cast<const float*>(sourceNormals.data()) : nullptr,
                                                         sc ? reinterpret_cast<const uint8_t*>(sourceColors.data()) : nullptr, (int64_t)n,
                                                         targetPoints.empty() ? nullptr : reinterpret_cast<const float*>(targetPoints.data()),
                                                         tn ? reinterpret_cast<const float*>(targetNormals.data()) : nullptr,
                                                         tc ? reinterpret_cast<const uint8_t*>(targetColors.data()) : nullptr, (int64_t)targetPoints.size(),
                                                         idx.data(), w.data()), "WeightingMethod::applyWeights")) return;
        for (size_t i = 0; i < n; ++i) matches[i].weight = w[i];
    }

private:
    int method;
    float maxDistance;
};
