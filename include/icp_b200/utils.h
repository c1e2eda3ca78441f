// utils.h -- drop-in for icp-variants/utils.h:
//   fillVector                         utils.h:10-15
//   PoseIncrement<T>                   utils.h:25-102   (angle-axis + translation over a caller-owned array of 6; apply, apply_inv_rotation,
//                                                        convertToMatrix)
//   transformPoints / transformNormals utils.h:106-133  -> icp_gpu_transform_points / icp_gpu_transform_normals (device)
//   computeMean, gettranslationMatrix, crossProductMatrix, getRodriguesMatrix   utils.h:136-176 (small host algebra)
// The reference leans on ceres/rotation.h for the angle-axis algebra; here it is written out (same formulas, same small-angle
// branches: Rodrigues for theta^2 > epsilon, the first-order form below it), generic in T so that the functors of constraints.h
// evaluate with doubles, floats or automatic-differentiation jets alike.  saveRoomToFile (:178-197, mesh output) is not part of the
// registration path and stays with the caller.
#pragma once
#include <cmath>
#include <limits>
#include <vector>
#include "Eigen.h"
#include "detail.h"

template <typename T>
static inline void fillVector(const Vector3f& input, T* output) {
    output[0] = T(input[0]); output[1] = T(input[1]); output[2] = T(input[2]);
}

namespace icp_b200 {
// result = R(angle_axis) * pt   (ceres::AngleAxisRotatePoint)
template <typename T>
inline void angleAxisRotatePoint(const T* aa, const T* pt, T* result) {
    using std::sqrt; using std::cos; using std::sin;
    const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
    if (theta2 > T(std::numeric_limits<double>::epsilon())) {
        const T theta = sqrt(theta2), c = cos(theta), s = sin(theta), inv = T(1.0) / theta;
        const T w[3] = {aa[0] * inv, aa[1] * inv, aa[2] * inv};
        const T wxp[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
        const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - c);
        result[0] = pt[0] * c + wxp[0] * s + w[0] * tmp;
        result[1] = pt[1] * c + wxp[1] * s + w[1] * tmp;
        result[2] = pt[2] * c + wxp[2] * s + w[2] * tmp;
    } else {   // near zero: R ~ I + [aa]x, which keeps the derivative information
        const T wxp[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2], aa[0] * pt[1] - aa[1] * pt[0]};
        result[0] = pt[0] + wxp[0]; result[1] = pt[1] + wxp[1]; result[2] = pt[2] + wxp[2];
    }
}
// column-major 3x3 (ceres::AngleAxisToRotationMatrix)
inline void angleAxisToRotationMatrix(const double* aa, double* R) {
    const double theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
    if (theta2 > std::numeric_limits<double>::epsilon()) {
        const double theta = std::sqrt(theta2), wx = aa[0] / theta, wy = aa[1] / theta, wz = aa[2] / theta, c = std::cos(theta), s = std::sin(theta);
        R[0] = c + wx * wx * (1 - c);       R[1] = wz * s + wx * wy * (1 - c);  R[2] = -wy * s + wx * wz * (1 - c);
        R[3] = wx * wy * (1 - c) - wz * s;  R[4] = c + wy * wy * (1 - c);       R[5] = wx * s + wy * wz * (1 - c);
        R[6] = wy * s + wx * wz * (1 - c);  R[7] = -wx * s + wy * wz * (1 - c); R[8] = c + wz * wz * (1 - c);
    } else {
        R[0] = 1; R[1] = aa[2]; R[2] = -aa[1]; R[3] = -aa[2]; R[4] = 1; R[5] = aa[0]; R[6] = aa[1]; R[7] = -aa[0]; R[8] = 1;
    }
}
}  // namespace icp_b200

// Interface onto an array of 6 (no copy): [0..2] angle-axis rotation, [3..5] translation.
template <typename T>
class PoseIncrement {
public:
    explicit PoseIncrement(T* const array) : m_array{array} {}
    void setZero() { for (int i = 0; i < 6; ++i) m_array[i] = T(0); }
    T* getData() const { return m_array; }
    void apply(T* inputPoint, T* outputPoint) const {
        T temp[3];
        icp_b200::angleAxisRotatePoint(m_array, inputPoint, temp);
        outputPoint[0] = temp[0] + m_array[3]; outputPoint[1] = temp[1] + m_array[4]; outputPoint[2] = temp[2] + m_array[5];
    }
    void apply_inv_rotation(T* inputPoint, T* outputPoint) const {      // the inverse rotation only, no translation
        const T inv[3] = {-m_array[0], -m_array[1], -m_array[2]};
        icp_b200::angleAxisRotatePoint(inv, inputPoint, outputPoint);
    }
    static Matrix4f convertToMatrix(const PoseIncrement<double>& poseIncrement) {
        const double* pose = poseIncrement.getData();
        double R[9];
        icp_b200::angleAxisToRotationMatrix(pose, R);
        Matrix4f matrix = Matrix4f::Identity();
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) matrix(r, c) = float(R[c * 3 + r]); matrix(r, 3) = float(pose[3 + r]); }
        return matrix;
    }
private:
    T* m_array;
};

// utils.h:106-118 on the device: q = R p + t
inline std::vector<Vector3f> transformPoints(const std::vector<Vector3f>& sourcePoints, const Matrix4f& pose) {
    std::vector<Vector3f> out(sourcePoints.size());
    if (sourcePoints.empty()) return out;
    icp_gpu_ctx* ctx = icp_b200::sharedContext();
    if (!ctx || !icp_b200::report(ctx, icp_gpu_transform_points(ctx, pose.data(), reinterpret_cast<const float*>(sourcePoints.data()), (int64_t)sourcePoints.size(),
                                                                reinterpret_cast<float*>(out.data())), "transformPoints")) return std::vector<Vector3f>();
    return out;
}
// utils.h:122-133 on the device: n' = (R^-1)^T n
inline std::vector<Vector3f> transformNormals(const std::vector<Vector3f>& sourceNormals, const Matrix4f& pose) {
    std::vector<Vector3f> out(sourceNormals.size());
    if (sourceNormals.empty()) return out;
    icp_gpu_ctx* ctx = icp_b200::sharedContext();
    if (!ctx || !icp_b200::report(ctx, icp_gpu_transform_normals(ctx, pose.data(), reinterpret_cast<const float*>(sourceNormals.data()), (int64_t)sourceNormals.size(),
                                                                 reinterpret_cast<float*>(out.data())), "transformNormals")) return std::vector<Vector3f>();
    return out;
}

inline Vector3f computeMean(const std::vector<Vector3f>& points) {          // utils.h:136-146 (an empty input yields (0,0,0) instead of hanging)
    float m[3] = {0.f, 0.f, 0.f};
    for (const auto& p : points) { m[0] += p[0]; m[1] += p[1]; m[2] += p[2]; }
    const float n = points.empty() ? 1.f : (float)points.size();
    return Vector3f(m[0] / n, m[1] / n, m[2] / n);
}
inline Matrix4f gettranslationMatrix(const Vector3f& translation) {         // utils.h:150-157
    Matrix4f matrix = Matrix4f::Identity();
    matrix(0, 3) = translation[0]; matrix(1, 3) = translation[1]; matrix(2, 3) = translation[2];
    return matrix;
}
inline Matrix3f crossProductMatrix(const Vector3f& k) {                     // utils.h:161-168
    Matrix3f m; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m(r, c) = 0.f;
    m(0, 1) = -k[2]; m(0, 2) = k[1]; m(1, 0) = k[2]; m(1, 2) = -k[0]; m(2, 0) = -k[1]; m(2, 1) = k[0];
    return m;
}
inline Matrix3f getRodriguesMatrix(const Vector3f& axis, const float& sin_theta, const float& cos_theta) {   // utils.h:171-176
    const Matrix3f K = crossProductMatrix(axis);
    Matrix3f R;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
        float kk = 0.f; for (int j = 0; j < 3; ++j) kk += ((1 - cos_theta) * K(r, j)) * K(j, c);
        R(r, c) = (r == c ? 1.f : 0.f) + (sin_theta * K(r, c) + kk);
    }
    return R;
}

// utils.h:178-197: the current depth frame as a mesh next to a camera glyph, written as .off -- debugging output of reconstructRoom,
// available where the reference's SimpleMesh.h is on the include path (mesh output is not part of the registration path).
#if defined(__has_include)
#if __has_include("SimpleMesh.h") && __has_include("VirtualSensor.h")
#include <sstream>
#include <string>
#include "VirtualSensor.h"
#include "SimpleMesh.h"
inline int saveRoomToFile(VirtualSensor& sensor, const Matrix4f& currentCameraPose, const std::string& filenameBaseOut) {
    SimpleMesh currentDepthMesh{sensor, currentCameraPose, 0.1f};
    SimpleMesh currentCameraMesh = SimpleMesh::camera(currentCameraPose, 0.0015f);
    SimpleMesh resultingMesh = SimpleMesh::joinMeshes(currentDepthMesh, currentCameraMesh, Matrix4f::Identity());
    std::stringstream ss;
    ss << filenameBaseOut << sensor.getCurrentFrameCnt() << ".off";
    std::cout << ss.str() << std::endl;
    if (!resultingMesh.writeMesh(ss.str())) { std::cout << "Failed to write mesh!\nCheck file path!" << std::endl; return -1; }
    return 0;
}
#endif
#endif
