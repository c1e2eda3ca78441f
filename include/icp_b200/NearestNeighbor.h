// NearestNeighbor.h -- drop-in for the reference's matcher classes (icp-variants/NearestNeighbor.h):
// same class names, method signatures and return type, backed by the icp_gpu_* C ABI (icp_gpu.h).
//   Match                                NearestNeighbor.h:7-10
//   NearestNeighborSearch                NearestNeighbor.h:12-36
//   NearestNeighborSearchBruteForce      NearestNeighbor.h:42-98
//   NearestNeighborSearchFlann           NearestNeighbor.h:104-314   (exact here; FLANN's is approximate)
//   NearestNeighborSearchProjective      NearestNeighbor.h:317-444
// Error behaviour follows the reference: a message on stdout and an empty result.  No CPU fallback:
// without a usable B200-class device the constructor reports the failure and every query returns {}.
#pragma once
#include <memory>
#include "../icp_gpu.h"
#include "Eigen.h"

#define MAX_DISTANCE 0.005f

struct Match {
    int idx;
    float weight;
};

class NearestNeighborSearch {
public:
    virtual ~NearestNeighborSearch() { if (m_ctx) icp_gpu_destroy(m_ctx); }

    virtual void setMatchingMaxDistance(float maxDistance) { m_maxDistance = maxDistance; }   // squared
    float getMatchingMaxDistance(float) { return m_maxDistance; }

    virtual void buildIndex(const std::vector<Eigen::Vector3f>& targetPoints) { build(targetPoints, nullptr); }
    virtual void buildIndex(const std::vector<Eigen::Vector3f>& targetPoints, const std::vector<Vector4uc>& targetColors) { build(targetPoints, &targetColors); }
    virtual std::vector<Match> queryMatches(const std::vector<Vector3f>& transformedPoints) { return query(transformedPoints, nullptr); }
    virtual std::vector<Match> queryMatches(const std::vector<Vector3f>& transformedPoints, const std::vector<Vector4uc>& transformedColors) {
        return query(transformedPoints, &transformedColors);
    }
    virtual void setCameraParams(const Eigen::Matrix3f& depthIntrinsics, const unsigned width, const unsigned height) {
        if (m_ctx) icp_gpu_set_camera(m_ctx, depthIntrinsics.data(), width, height);
    }

protected:
    float m_maxDistance;
    icp_gpu_ctx* m_ctx;
    int m_matching, m_nn, m_built, m_colors;

    NearestNeighborSearch(int matching, int nn_algorithm, int device = 0)
        : m_maxDistance{MAX_DISTANCE}, m_ctx{nullptr}, m_matching{matching}, m_nn{nn_algorithm}, m_built{0}, m_colors{0} {
        if (icp_gpu_create(&m_ctx, device) != ICP_GPU_OK) { m_ctx = nullptr; std::cout << "icp_gpu: no usable CUDA device (there is no CPU fallback)." << std::endl; }
    }

    void build(const std::vector<Eigen::Vector3f>& pts, const std::vector<Vector4uc>* cols) {
        if (!m_ctx) return;
        const int rc = icp_gpu_set_target(m_ctx, pts.empty() ? nullptr : reinterpret_cast<const float*>(pts.data()), nullptr,
                                          (cols && !cols->empty()) ? reinterpret_cast<const uint8_t*>(cols->data()) : nullptr, (int64_t)pts.size());
        if (rc != ICP_GPU_OK) { std::cout << "icp_gpu: " << icp_gpu_last_error(m_ctx) << std::endl; return; }
        m_built = 1; m_colors = cols ? 1 : 0;
    }

    std::vector<Match> query(const std::vector<Vector3f>& pts, const std::vector<Vector4uc>* cols) {
        if (!m_ctx || !m_built) {   // NearestNeighbor.h:144-147
            std::cout << "FLANN index needs to be build before querying any matches." << std::endl;
            return {};
        }
        if ((cols != nullptr) != (m_colors != 0)) {   // NearestNeighbor.h:148-152
            std::cout << "Index and query dimensionality do not agree." << std::endl;
            return {};
        }
        icp_gpu_config cfg; icp_gpu_default_config(&cfg);
        cfg.matching = m_matching; cfg.nn_algorithm = m_nn; cfg.max_distance_sq = m_maxDistance;
        cfg.rejection = 0; cfg.weighting = ICP_GPU_WEIGHT_CONSTANT; cfg.color_icp = m_colors;
        const float eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        std::vector<int32_t> idx(pts.size()); std::vector<float> w(pts.size());
        int rc = icp_gpu_set_config(m_ctx, &cfg);
        if (rc == ICP_GPU_OK) rc = icp_gpu_set_source(m_ctx, pts.empty() ? nullptr : reinterpret_cast<const float*>(pts.data()), nullptr,
                                                      (cols && !cols->empty()) ? reinterpret_cast<const uint8_t*>(cols->data()) : nullptr, (int64_t)pts.size());
        if (rc == ICP_GPU_OK) rc = icp_gpu_query_matches(m_ctx, eye, nullptr, 0, idx.data(), w.data());
        if (rc != ICP_GPU_OK) { std::cout << "icp_gpu: " << icp_gpu_last_error(m_ctx) << std::endl; return {}; }
        std::vector<Match> matches(pts.size());
        for (size_t i = 0; i < pts.size(); ++i) matches[i] = Match{idx[i], w[i]};
        return matches;
    }
};

class NearestNeighborSearchBruteForce : public NearestNeighborSearch {
public:
    // NearestNeighbor.h:81-97 compares (p - m).norm() with m_maxDistance: a plain distance, not a squared one -- ICP_GPU_NN_BRUTE_NORM
    NearestNeighborSearchBruteForce() : NearestNeighborSearch(ICP_GPU_MATCH_KNN, ICP_GPU_NN_BRUTE_NORM) {}
};

class NearestNeighborSearchFlann : public NearestNeighborSearch {
public:
    NearestNeighborSearchFlann() : NearestNeighborSearch(ICP_GPU_MATCH_KNN, ICP_GPU_NN_AUTO) {}
};

class NearestNeighborSearchProjective : public NearestNeighborSearch {
public:
    NearestNeighborSearchProjective() : NearestNeighborSearch(ICP_GPU_MATCH_PROJECTIVE, ICP_GPU_NN_AUTO) {}
};
