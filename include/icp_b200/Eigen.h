// Eigen.h -- the value types that cross the drop-in boundary (reference: icp-variants/Eigen.h:21,36).
//
// With real Eigen available, define ICP_B200_USE_EIGEN: the reference's own types are used and the
// drop-in classes pass their storage straight to the icp_gpu_* C ABI (std::vector<Eigen::Vector3f>
// is packed float[3N], Eigen::Matrix4f is float[16] column-major, Vector4uc is uchar[4]).
// Without Eigen (this container has none) a minimal layout-compatible stand-in is used so that the
// headers and the example driver compile and run.
#pragma once
#include <cmath>
#include <cstring>
#include <iostream>
#include <limits>
#include <vector>

#ifndef MINF
#define MINF -std::numeric_limits<float>::infinity()
#endif
#ifndef M_PI
#define M_PI 3.14159265359
#endif
#ifndef VERBOSE
#define VERBOSE(msg)
#endif
// The reference's ASSERT prints and then spins forever (`while(1);`, Eigen.h:9 -- and, unparenthesised, tests `!a` of the first operand
// only).  The drop-in's prints and aborts the process: nothing hangs.
#ifndef ASSERT
#include <cstdlib>
#define ASSERT(a) { if (!(a)) { std::cerr << "Error:\nFile: " << __FILE__ << "\nLine: " << __LINE__ << "\nFunction: " << __FUNCTION__ << std::endl; std::abort(); } }
#endif
#ifndef SAFE_DELETE
#define SAFE_DELETE(ptr) { if (ptr != nullptr) { delete ptr; ptr = nullptr; } }
#endif
#ifndef SAFE_DELETE_ARRAY
#define SAFE_DELETE_ARRAY(ptr) { if (ptr != nullptr) { delete[] ptr; ptr = nullptr; } }
#endif

#ifdef ICP_B200_USE_EIGEN
#include <Eigen/Dense>
#include <Eigen/StdVector>
typedef Eigen::Matrix<unsigned char, 4, 1> Vector4uc;
typedef Eigen::Matrix<unsigned char, 3, 1> Vector3uc;
using namespace Eigen;      // as the reference's Eigen.h:43 does: its drivers name Matrix4f, Vector3f, MatrixXf, JacobiSVD ... unqualified
// stream operators of the reference's Eigen.h:66-79 (VirtualSensor.h reads the trajectory quaternions with them)
template <typename T> std::istream& operator>>(std::istream& in, Eigen::Quaternion<T>& other) { in >> other.x() >> other.y() >> other.z() >> other.w(); return in; }
template <typename T> std::ostream& operator<<(std::ostream& out, const Eigen::Quaternion<T>& other) {
    std::fixed(out);
    out << other.x() << "\t" << other.y() << "\t" << other.z() << "\t" << other.w();
    return out;
}
#else
namespace Eigen {
struct Vector3f {
    float v[3];
    Vector3f() : v{0.f, 0.f, 0.f} {}
    Vector3f(float x, float y, float z) : v{x, y, z} {}
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    float& x() { return v[0]; } float& y() { return v[1]; } float& z() { return v[2]; }
    float x() const { return v[0]; } float y() const { return v[1]; } float z() const { return v[2]; }
    const float* data() const { return v; }
    bool allFinite() const { return std::isfinite(v[0]) && std::isfinite(v[1]) && std::isfinite(v[2]); }
    static Vector3f Zero() { return Vector3f(); }
};
struct Vector4uc {
    unsigned char v[4];
    Vector4uc() : v{0, 0, 0, 0} {}
    Vector4uc(unsigned char r, unsigned char g, unsigned char b, unsigned char a) : v{r, g, b, a} {}
    unsigned char& operator[](int i) { return v[i]; }
    unsigned char operator[](int i) const { return v[i]; }
    static Vector4uc Zero() { return Vector4uc(); }
};
// column-major like Eigen's default
template <int N>
struct MatrixNf {
    float m[N * N];
    MatrixNf() { std::memset(m, 0, sizeof(m)); }
    static MatrixNf Identity() { MatrixNf r; for (int i = 0; i < N; ++i) r.m[i * N + i] = 1.f; return r; }
    void setIdentity() { *this = Identity(); }
    float& operator()(int r, int c) { return m[c * N + r]; }
    float operator()(int r, int c) const { return m[c * N + r]; }
    float* data() { return m; }
    const float* data() const { return m; }
};
typedef MatrixNf<3> Matrix3f;
typedef MatrixNf<4> Matrix4f;
}  // namespace Eigen
typedef Eigen::Vector4uc Vector4uc;
using Eigen::Matrix3f;
using Eigen::Matrix4f;
using Eigen::Vector3f;
static_assert(sizeof(Eigen::Vector3f) == 12 && sizeof(Vector4uc) == 4 && sizeof(Eigen::Matrix4f) == 64, "packed layouts");
#endif
