// ICPOptimizer.h -- drop-in for the reference's optimizer classes (icp-variants/ICPOptimizer.h):
//   ICPOptimizer          options + setters + estimatePose      ICPOptimizer.h:27-175
//   CeresICPOptimizer     Levenberg-Marquardt minimiser          ICPOptimizer.h:181-483
//   LinearICPOptimizer    closed-form / linear-system minimiser  ICPOptimizer.h:489-899
// Same names, signatures, defaults and call order as the reference, so main.cpp / experiment.cpp style
// drivers compile against it unchanged.  The whole iteration loop runs on the device behind
// icp_gpu_estimate_pose (icp_gpu.h); the per-iteration poses come back for ConvergenceMeasure and the
// CUDA-event stage times for TimeMeasure.
// Deviations from the reference, all on error paths: where the reference hangs in ASSERT (`while(1);`,
// Eigen.h:9 -- e.g. no surviving correspondence, ICPOptimizer.h:668,680,788) the pose is left at the last
// good estimate and a message is printed; m_timeMeasure / m_convergenceMeasure may be left unset.
#pragma once
#include <memory>
#include "../icp_gpu.h"
#include "ConvergenceMeasure.h"
#include "Eigen.h"
#include "NearestNeighbor.h"
#include "PointCloud.h"
#include "ProcrustesAligner.h"
#include "TimeMeasure.h"
#include "constraints.h"
#include "selection.h"
#include "utils.h"
#include "weighting.h"

class ICPOptimizer {
public:
    ICPOptimizer()
        : metric{0}, colorICP{false}, multiResolutionICP{false}, selectionMethod{0}, proba{1.0}, rejectionMethod{1}, weightingMethod{0},
          matchingMethod{0}, m_nIterations{20}, m_timeMeasure{nullptr}, m_convergenceMeasure{nullptr}, maxDistance{0.0003f},
          m_matcherMaxDistance{MAX_DISTANCE},
          m_ctx{nullptr}, m_haveCamera{false}, m_seed{0}, m_selectionRng{ICP_GPU_RNG_MT19937}, m_width{0}, m_height{0}, m_pyramidMode{ICP_GPU_PYRAMID_STRIDE},
          m_stopRot{0.f}, m_stopTrans{0.f} {
        if (icp_gpu_create(&m_ctx, 0) != ICP_GPU_OK) { m_ctx = nullptr; std::cout << "icp_gpu: no usable CUDA device (there is no CPU fallback)." << std::endl; }
    }
    virtual ~ICPOptimizer() { if (m_ctx) icp_gpu_destroy(m_ctx); }

    // ICPOptimizer.h:41-44: the matcher's threshold AND the distance WeightingMethod divides by
    void setMatchingMaxDistance(float maxDistance) { m_matcherMaxDistance = maxDistance; this->maxDistance = maxDistance; }
    void setMetric(unsigned int metric) { this->metric = metric; }
    void enableMultiResolution(bool enableMultiResolution) { this->multiResolutionICP = enableMultiResolution; }
    // extension: ICP_GPU_PYRAMID_VOXEL builds the levels by voxel downsampling instead of the reference's index stride
    void setPyramidMode(int pyramidMode) { m_pyramidMode = pyramidMode; }
    // extension: stop once an applied increment is below both thresholds (radians, metres); 0 = the reference's behaviour (all iterations)
    void setEarlyStop(float rotation, float translation) { m_stopRot = rotation; m_stopTrans = translation; }
    void enableColorICP(bool colorICP) { this->colorICP = colorICP; }
    void setSelectionMethod(unsigned int selectionMethod, double proba = 1.0) { this->selectionMethod = selectionMethod; this->proba = proba; }
    void setRejectionMethod(unsigned int rejectionMethod) { this->rejectionMethod = rejectionMethod; }
    void setWeightingMethod(unsigned int weightingMethod) { this->weightingMethod = weightingMethod; }
    void setMatchingMethod(unsigned int matchingMethod) {
        this->matchingMethod = matchingMethod;
        m_matcherMaxDistance = MAX_DISTANCE;   // ICPOptimizer.h:71-78 re-creates the matcher: ITS distance is back at the default, maxDistance is not touched
    }
    void setCameraParamsMatchingMethod(const Eigen::Matrix3f& depthIntrinsics, const unsigned width, const unsigned height) {
        m_K = depthIntrinsics; m_width = width; m_height = height; m_haveCamera = true;
    }
    void setNbOfIterations(unsigned nIterations) { m_nIterations = nIterations; }
    void setTimeMeasure(TimeMeasure& timeMeasure) { timeMeasure.nIterations = &m_nIterations; m_timeMeasure = &timeMeasure; }
    void setConvergenceMeasure(ConvergenceMeasure& convergenMearsure) { m_convergenceMeasure = &convergenMearsure; }
    // The reference seeds its sampler from std::random_device (selection.h:76-79); here the seed is explicit.
    void setSelectionSeed(unsigned seed, bool deviceStream = false) { m_seed = seed; m_selectionRng = deviceStream ? ICP_GPU_RNG_DEVICE : ICP_GPU_RNG_MT19937; }

    // extension: the C-ABI handle behind this optimizer, for what the reference's interface has no name for (icp_gpu_set_stream,
    // icp_gpu_peer_export / icp_gpu_peer_attach: after attaching, estimatePose of every rank is one point-sharded registration --
    // the source handed to it is the rank's shard, the target the whole cloud)
    icp_gpu_ctx* context() const { return m_ctx; }

    void printICPConfiguration() {   // ICPOptimizer.h:97-138
        std::cout << "\n\n*-*-*-*-*-*-*-*-*-*-*-*-*-*-*-*-*\nStarting ICP with the following configuration:\n";
        if (colorICP) std::cout << "Color-ICP enabled\n";
        if (multiResolutionICP) std::cout << "Multi-Resolution ICP enabled\n";
        std::cout << (selectionMethod == SELECT_ALL ? "1. Selection: all\n" : "1. Selection: random\n");
        std::cout << (matchingMethod == 1 ? "2. Matching: projective (max distance " : "2. Matching: k-nn (max distance ") << maxDistance << " m)\n";
        const char* w[] = {"constant", "point distances", "normals", "colors"};
        std::cout << "3. Weighting: " << w[weightingMethod & 3] << "\n";
        std::cout << (rejectionMethod == 1 ? "4. Rejection: angle of normals\n" : "4. Rejection: keep all\n");
        const char* m[] = {"Point to Point", "Point to Plane", "Symmetric"};
        std::cout << "5. Metric: " << m[metric % 3] << "\n*-*-*-*-*-*-*-*-*-*-*-*-*-*-*-*-*\n\n";
    }

    virtual void estimatePose(const PointCloud& source, const PointCloud& target, Matrix4f& initialPose, bool calculateRMSE = true) = 0;

protected:
    unsigned int metric;
    bool colorICP;
    bool multiResolutionICP;
    unsigned int selectionMethod;
    double proba;
    unsigned int rejectionMethod;
    unsigned int weightingMethod;
    unsigned int matchingMethod;
    unsigned m_nIterations;
    TimeMeasure* m_timeMeasure;
    ConvergenceMeasure* m_convergenceMeasure;
    float maxDistance;   // squared distance: what WeightingMethod is constructed with (ICPOptimizer.h:220,528); default 0.0003f
    float m_matcherMaxDistance;   // NearestNeighborSearch::m_maxDistance of the matcher the reference owns (NearestNeighbor.h:17-19,35)
    icp_gpu_ctx* m_ctx;
    Eigen::Matrix3f m_K; bool m_haveCamera; unsigned m_seed; int m_selectionRng; unsigned m_width, m_height; int m_pyramidMode;
    float m_stopRot, m_stopTrans;

    // body of estimatePose shared by both minimisers (ICPOptimizer.h:185-349 / :493-663)
    void run(int minimizer, const PointCloud& source, const PointCloud& target, Matrix4f& initialPose, bool calculateRMSE) {
        if (!m_ctx) return;
        printICPConfiguration();
        icp_gpu_config cfg; icp_gpu_default_config(&cfg);
        cfg.metric = (int32_t)metric; cfg.minimizer = minimizer; cfg.matching = (int32_t)matchingMethod;
        cfg.selection = (int32_t)selectionMethod; cfg.proba = proba; cfg.seed = m_seed; cfg.selection_rng = m_selectionRng;
        cfg.weighting = (int32_t)weightingMethod; cfg.rejection = (int32_t)rejectionMethod;
        cfg.max_distance_sq = m_matcherMaxDistance; cfg.weight_max_distance_sq = maxDistance;
        cfg.color_icp = colorICP ? 1 : 0; cfg.multires = multiResolutionICP ? 1 : 0; cfg.n_iterations = (int32_t)m_nIterations; cfg.pyramid_mode = m_pyramidMode;
        cfg.early_stop_rotation = m_stopRot; cfg.early_stop_translation = m_stopTrans;
        int rc = icp_gpu_set_config(m_ctx, &cfg);
        if (rc == ICP_GPU_OK && m_haveCamera) rc = icp_gpu_set_camera(m_ctx, m_K.data(), m_width, m_height);
        const auto& tp = target.getPoints(); const auto& tn = target.getNormals(); const auto& tc = target.getColors();
        const auto& sp = source.getPoints(); const auto& sn = source.getNormals(); const auto& sc = source.getColors();
        // buildIndex (ICPOptimizer.h:532-535) and the source upload
        if (rc == ICP_GPU_OK) rc = icp_gpu_set_target(m_ctx, tp.empty() ? nullptr : reinterpret_cast<const float*>(tp.data()),
                                                      tn.size() == tp.size() && !tn.empty() ? reinterpret_cast<const float*>(tn.data()) : nullptr,
                                                      tc.size() == tp.size() && !tc.empty() ? reinterpret_cast<const uint8_t*>(tc.data()) : nullptr, (int64_t)tp.size());
        if (rc == ICP_GPU_OK) rc = icp_gpu_set_source(m_ctx, sp.empty() ? nullptr : reinterpret_cast<const float*>(sp.data()),
                                                      sn.size() == sp.size() && !sn.empty() ? reinterpret_cast<const float*>(sn.data()) : nullptr,
                                                      sc.size() == sp.size() && !sc.empty() ? reinterpret_cast<const uint8_t*>(sc.data()) : nullptr, (int64_t)sp.size());
        if (rc != ICP_GPU_OK) { std::cout << "icp_gpu: " << icp_gpu_last_error(m_ctx) << std::endl; return; }
        const int cap = icp_gpu_max_iterations(m_ctx);
        std::vector<float> history((size_t)(cap > 0 ? cap : 1) * 16);
        int32_t nIt = 0; icp_gpu_timings tm = {};
        rc = icp_gpu_estimate_pose(m_ctx, initialPose.data(), history.data(), &nIt, m_timeMeasure ? &tm : nullptr);
        if (rc != ICP_GPU_OK) std::cout << "icp_gpu: " << icp_gpu_last_error(m_ctx) << " -- pose left at the last good estimate" << std::endl;
        if (m_timeMeasure && rc == ICP_GPU_OK) {
            m_timeMeasure->matchingTime += tm.matching_ms * 1e-3; m_timeMeasure->solverTime += tm.solver_ms * 1e-3;
            m_timeMeasure->convergenceTime += tm.total_ms * 1e-3; m_timeMeasure->indexTime += tm.index_ms * 1e-3;
        }
        if (calculateRMSE && m_convergenceMeasure && nIt > 0) {
            // recordAlignmentError after every iteration (ICPOptimizer.h:629-631), evaluated on the device from the pose history
            const auto& gs = m_convergenceMeasure->sourcePoints(); const auto& gu = m_convergenceMeasure->unchangedPoints();
            std::vector<float> rmse((size_t)nIt); std::vector<double> bench((size_t)nIt); int32_t nErr = 0;
            int rc2 = gs.empty() || gs.size() != gu.size() ? ICP_GPU_E_ARG
                      : icp_gpu_set_correspondences(m_ctx, reinterpret_cast<const float*>(gs.data()), reinterpret_cast<const float*>(gu.data()), (int64_t)gs.size());
            if (rc2 == ICP_GPU_OK) rc2 = icp_gpu_convergence_errors(m_ctx, rmse.data(), m_convergenceMeasure->runBenchmark() ? bench.data() : nullptr, nIt, &nErr);
            if (rc2 == ICP_GPU_OK) for (int32_t i = 0; i < nErr; ++i) m_convergenceMeasure->recordDeviceErrors(rmse[(size_t)i], bench[(size_t)i]);
            else std::cout << "icp_gpu: convergence errors unavailable (" << (gs.empty() ? "no correspondences" : icp_gpu_last_error(m_ctx)) << ")" << std::endl;
        }
    }
};

class CeresICPOptimizer : public ICPOptimizer {
public:
    CeresICPOptimizer() {}
    virtual void estimatePose(const PointCloud& source, const PointCloud& target, Matrix4f& initialPose, bool calculateRMSE = true) override {
        run(ICP_GPU_MIN_LM, source, target, initialPose, calculateRMSE);
    }
};

class LinearICPOptimizer : public ICPOptimizer {
public:
    LinearICPOptimizer() {}
    virtual void estimatePose(const PointCloud& source, const PointCloud& target, Matrix4f& initialPose, bool calculateRMSE = true) override {
        run(ICP_GPU_MIN_LINEAR, source, target, initialPose, calculateRMSE);
    }
};
