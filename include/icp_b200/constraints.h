// constraints.h -- drop-in for icp-variants/constraints.h:9-143: the three residual functors of the non-linear minimiser,
//   PointToPointConstraint   r[0..2] = 0.1 w (T(x) s - d)                         LAMBDA 0.1  (:9-45)
//   PointToPlaneConstraint   r[0]    = 1.0 w n_t . (T(x) s - d)                   LAMBDA 1.0  (:47-89)
//   SymmetricConstraint      r[0]    = 1.0 w (n_t + n_s) . (T(x) s - R(x)^-1 d)   LAMBDA 1.0  (:91-143)
// with T(x) the PoseIncrement of utils.h.  Same constructors and the same templated operator()(pose, residuals), so they evaluate
// with doubles or with any automatic-differentiation scalar.  Inside CeresICPOptimizer::estimatePose the drop-in does not
// instantiate them: lm.cu evaluates the same residuals with forward-mode jets on the device.  `create` (the Ceres cost-function
// factory) exists only where <ceres/ceres.h> does.
#pragma once
#include "Eigen.h"
#include "utils.h"
#if defined(__has_include)
#if __has_include(<ceres/ceres.h>)
#include <ceres/ceres.h>
#define ICP_B200_HAVE_CERES 1
#endif
#endif

class PointToPointConstraint {
public:
    PointToPointConstraint(const Vector3f& sourcePoint, const Vector3f& targetPoint, const float weight)
        : m_sourcePoint{sourcePoint}, m_targetPoint{targetPoint}, m_weight{weight} {}
    template <typename T>
    bool operator()(const T* const pose, T* residuals) const {
        PoseIncrement<T> inc(const_cast<T*>(pose));
        T s[3] = {T(m_sourcePoint[0]), T(m_sourcePoint[1]), T(m_sourcePoint[2])}, y[3];
        inc.apply(s, y);
        for (int k = 0; k < 3; ++k) residuals[k] = T(LAMBDA) * T(m_weight) * (y[k] - T(m_targetPoint[k]));
        return true;
    }
#ifdef ICP_B200_HAVE_CERES
    static ceres::CostFunction* create(const Vector3f& sourcePoint, const Vector3f& targetPoint, const float weight) {
        return new ceres::AutoDiffCostFunction<PointToPointConstraint, 3, 6>(new PointToPointConstraint(sourcePoint, targetPoint, weight));
    }
#endif
protected:
    const Vector3f m_sourcePoint, m_targetPoint;
    const float m_weight;
    const float LAMBDA = 0.1f;
};

class PointToPlaneConstraint {
public:
    PointToPlaneConstraint(const Vector3f& sourcePoint, const Vector3f& targetPoint, const Vector3f& targetNormal, const float weight)
        : m_sourcePoint{sourcePoint}, m_targetPoint{targetPoint}, m_targetNormal{targetNormal}, m_weight{weight} {}
    template <typename T>
    bool operator()(const T* const pose, T* residuals) const {
        PoseIncrement<T> inc(const_cast<T*>(pose));
        T s[3] = {T(m_sourcePoint[0]), T(m_sourcePoint[1]), T(m_sourcePoint[2])}, y[3];
        inc.apply(s, y);
        const T x = T(m_targetNormal[0]) * (y[0] - T(m_targetPoint[0])), yy = T(m_targetNormal[1]) * (y[1] - T(m_targetPoint[1])),
                z = T(m_targetNormal[2]) * (y[2] - T(m_targetPoint[2]));
        residuals[0] = T(LAMBDA) * T(m_weight) * (x + yy + z);
        return true;
    }
#ifdef ICP_B200_HAVE_CERES
    static ceres::CostFunction* create(const Vector3f& sourcePoint, const Vector3f& targetPoint, const Vector3f& targetNormal, const float weight) {
        return new ceres::AutoDiffCostFunction<PointToPlaneConstraint, 1, 6>(new PointToPlaneConstraint(sourcePoint, targetPoint, targetNormal, weight));
    }
#endif
protected:
    const Vector3f m_sourcePoint, m_targetPoint, m_targetNormal;
    const float m_weight;
    const float LAMBDA = 1.0f;
};

class SymmetricConstraint {
public:
    SymmetricConstraint(const Vector3f& sourcePoint, const Vector3f& targetPoint, const Vector3f& sourceNormal, const Vector3f& targetNormal, const float weight)
        : m_sourcePoint{sourcePoint}, m_targetPoint{targetPoint}, m_sourceNormal{sourceNormal}, m_targetNormal{targetNormal}, m_weight{weight} {}
    template <typename T>
    bool operator()(const T* const pose, T* residuals) const {
        PoseIncrement<T> inc(const_cast<T*>(pose));
        T s[3] = {T(m_sourcePoint[0]), T(m_sourcePoint[1]), T(m_sourcePoint[2])}, d[3] = {T(m_targetPoint[0]), T(m_targetPoint[1]), T(m_targetPoint[2])}, y[3], z[3];
        inc.apply(s, y);                    // R s + t
        inc.apply_inv_rotation(d, z);       // R^-1 d
        const T a = (T(m_targetNormal[0]) + T(m_sourceNormal[0])) * (y[0] - z[0]), b = (T(m_targetNormal[1]) + T(m_sourceNormal[1])) * (y[1] - z[1]),
                c = (T(m_targetNormal[2]) + T(m_sourceNormal[2])) * (y[2] - z[2]);
        residuals[0] = T(LAMBDA) * T(m_weight) * (a + b + c);
        return true;
    }
#ifdef ICP_B200_HAVE_CERES
    static ceres::CostFunction* create(const Vector3f& sourcePoint, const Vector3f& targetPoint, const Vector3f& sourceNormal, const Vector3f& targetNormal, const float weight) {
        return new ceres::AutoDiffCostFunction<SymmetricConstraint, 1, 6>(new SymmetricConstraint(sourcePoint, targetPoint, sourceNormal, targetNormal, weight));
    }
#endif
protected:
    const Vector3f m_sourcePoint, m_targetPoint, m_sourceNormal, m_targetNormal;
    const float m_weight;
    const float LAMBDA = 1.0f;
};
