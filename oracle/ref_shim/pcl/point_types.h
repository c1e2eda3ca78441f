// oracle/ref_shim/pcl/point_types.h -- TEST INFRASTRUCTURE ONLY.
//
// Stand-in for the slice of PCL the reference's headers touch (PointCloud.h:41-76,229-261,
// ConvergenceMeasure.h:104-151): point structs, pcl::PointCloud, NormalEstimation with
// setKSearch(k), compute3DCentroid, euclideanDistance.  NOT PCL.  NormalEstimation restates PCL's
// published algorithm (features/normal_3d.h): k nearest neighbours of every point (the point
// itself included), covariance of the neighbourhood, eigenvector of the smallest eigenvalue,
// flipped towards the viewpoint (0,0,0); curvature = l0 / (l0+l1+l2).  Neighbours are found
// exactly (exhaustive scan, ties to the lowest index); the covariance is accumulated in double.
#ifndef ICP_REF_SHIM_PCL
#define ICP_REF_SHIM_PCL
#include <Eigen/Dense>
#include <algorithm>
#include <cmath>
#include <memory>
#include <string>
#include <vector>
namespace pcl {
struct PointXYZ { float x, y, z; PointXYZ() : x(0), y(0), z(0) {} PointXYZ(float a, float b, float c) : x(a), y(b), z(c) {} };
struct Normal { float normal_x, normal_y, normal_z, curvature; Normal() : normal_x(0), normal_y(0), normal_z(0), curvature(0) {} };
struct PointXYZINormal { float x, y, z, intensity, normal_x, normal_y, normal_z, curvature; };
template <class P> class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<P> > Ptr;
    typedef std::shared_ptr<const PointCloud<P> > ConstPtr;
    std::vector<P> points; unsigned width = 0, height = 0;
    size_t size() const { return points.size(); }
    P& at(size_t i) { return points.at(i); }
    const P& at(size_t i) const { return points.at(i); }
    void push_back(const P& p) { points.push_back(p); }
    Ptr makeShared() const { return Ptr(new PointCloud<P>(*this)); }
};
namespace search { template <class P> class KdTree { public: typedef std::shared_ptr<KdTree<P> > Ptr; }; }
namespace io { template <class C> int savePLYFile(const std::string&, const C&) { return -1; /* file output is not part of the path */ } }

namespace shim {
// Jacobi eigen-decomposition of a symmetric 3x3 (row-major), eigenvalues ascending.
inline void eig3(const double A_[9], double evals[3], double evecs[9] /* column k = evecs[3*i+k] */) {
    double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; ++i) A[i] = A_[i];
    for (int sweep = 0; sweep < 64; ++sweep) {
        const double off = (std::fabs(A[1]) + std::fabs(A[2])) + std::fabs(A[5]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
            const double apq = A[p * 3 + q];
            if (apq == 0.0) continue;
            const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 3; ++k) { const double akp = A[k * 3 + p], akq = A[k * 3 + q]; A[k * 3 + p] = c * akp - s * akq; A[k * 3 + q] = s * akp + c * akq; }
            for (int k = 0; k < 3; ++k) { const double apk = A[p * 3 + k], aqk = A[q * 3 + k]; A[p * 3 + k] = c * apk - s * aqk; A[q * 3 + k] = s * apk + c * aqk; }
            for (int k = 0; k < 3; ++k) { const double vkp = V[k * 3 + p], vkq = V[k * 3 + q]; V[k * 3 + p] = c * vkp - s * vkq; V[k * 3 + q] = s * vkp + c * vkq; }
        }
    }
    int ord[3] = {0, 1, 2};   // stable ascending order of the diagonal
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2 - i; ++j) if (A[ord[j + 1] * 4] < A[ord[j] * 4]) std::swap(ord[j], ord[j + 1]);
    for (int k = 0; k < 3; ++k) { evals[k] = A[ord[k] * 3 + ord[k]]; for (int i = 0; i < 3; ++i) evecs[i * 3 + k] = V[i * 3 + ord[k]]; }
}
}  // namespace shim

template <class PIn, class POut> class NormalEstimation {
public:
    NormalEstimation() : k_(0), vpx_(0), vpy_(0), vpz_(0) {}
    void setInputCloud(const typename PointCloud<PIn>::Ptr& c) { in_ = c; }
    void setSearchMethod(const typename search::KdTree<PIn>::Ptr&) {}
    void setKSearch(int k) { k_ = k; }
    void setViewPoint(float x, float y, float z) { vpx_ = x; vpy_ = y; vpz_ = z; }
    void compute(PointCloud<POut>& out) {
        const long n = (long)in_->points.size();
        out.points.assign(n, POut()); out.width = (unsigned)n; out.height = 1;
        const std::vector<PIn>& P = in_->points;
#pragma omp parallel for schedule(dynamic, 64)
        for (long i = 0; i < n; ++i) {
            POut& o = out.points[i];
            const float nan = std::numeric_limits<float>::quiet_NaN();
            o.normal_x = o.normal_y = o.normal_z = o.curvature = nan;
            if (!std::isfinite(P[i].x) || !std::isfinite(P[i].y) || !std::isfinite(P[i].z)) continue;
            // k nearest (exact, includes i itself)
            std::vector<std::pair<float, long> > best; best.reserve(k_ + 1);
            for (long j = 0; j < n; ++j) {
                const float dx = P[j].x - P[i].x, dy = P[j].y - P[i].y, dz = P[j].z - P[i].z;
                const float d = dx * dx + dy * dy + dz * dz;
                if (!std::isfinite(d)) continue;
                if ((int)best.size() < k_ || d < best.back().first) {
                    std::pair<float, long> e(d, j);
                    best.insert(std::upper_bound(best.begin(), best.end(), e), e);
                    if ((int)best.size() > k_) best.pop_back();
                }
            }
            if ((int)best.size() < 3) continue;
            double m[3] = {0, 0, 0};
            for (auto& e : best) { m[0] += P[e.second].x; m[1] += P[e.second].y; m[2] += P[e.second].z; }
            for (int a = 0; a < 3; ++a) m[a] /= (double)best.size();
            double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (auto& e : best) {
                const double d[3] = {P[e.second].x - m[0], P[e.second].y - m[1], P[e.second].z - m[2]};
                for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[a * 3 + b] += d[a] * d[b];
            }
            for (int a = 0; a < 9; ++a) C[a] /= (double)best.size();
            double ev[3], V[9]; shim::eig3(C, ev, V);
            double nx = V[0], ny = V[3], nz = V[6];
            const double sum = (ev[0] + ev[1]) + ev[2];
            // flipNormalTowardsViewpoint
            const double vx = vpx_ - P[i].x, vy = vpy_ - P[i].y, vz = vpz_ - P[i].z;
            if ((vx * nx + vy * ny) + vz * nz < 0) { nx = -nx; ny = -ny; nz = -nz; }
            o.normal_x = (float)nx; o.normal_y = (float)ny; o.normal_z = (float)nz;
            o.curvature = sum != 0 ? (float)std::fabs(ev[0] / sum) : 0.f;
        }
    }
private:
    typename PointCloud<PIn>::Ptr in_; int k_; float vpx_, vpy_, vpz_;
};

template <class P, class S>
unsigned compute3DCentroid(const PointCloud<P>& c, Eigen::Matrix<S, 4, 1>& centroid) {
    centroid.setZero(); unsigned n = 0;
    for (const P& p : c.points) { if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue; centroid[0] += p.x; centroid[1] += p.y; centroid[2] += p.z; ++n; }
    if (n) { centroid /= (S)n; } centroid[3] = 1;
    return n;
}
template <class P1, class P2> float euclideanDistance(const P1& a, const P2& b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return std::sqrt(dx * dx + dy * dy + dz * dz);
}
}  // namespace pcl
#endif
