// PointCloud.h -- the container estimatePose receives (drop-in for icp-variants/PointCloud.h:8-349): points, normals, colours and
// the constructors / helpers the reference's drivers use.
//   PointCloud(mesh)                        :12-39    vertices + area-weighted vertex normals of any mesh type with getVertices() / getTriangles()
//                                                     (the reference's SimpleMesh) -- host, once per file
//   PointCloud(pcl::PointCloud<PointXYZ>::Ptr)  :41-76  points + k = 5 PCA normals (pcl::NormalEstimation) ON THE DEVICE (icp_gpu_target_normals),
//                                                     colours (255,255,255,1); any pointer-like with ->points[i].x/y/z
//   PointCloud(depthMap, colorFrame, ...)   :78-165   back-projection, central-difference normals, filter ON THE DEVICE (icp_gpu_cloud_from_depth)
//   readFromFile / writeToFile              :167-247  the reference's binary dump; PLY (ascii: x y z nx ny nz) without PCL
//   copy_point_cloud / change_pose          :263-283  change_pose = transformPoints on the device, normals by the rotation block (:280)
//   getClosestPoint / getCoarseResolution   :310-343
#pragma once
#include <cmath>
#include <fstream>
#include <string>
#include <utility>
#include "Eigen.h"
#include "detail.h"
#include "../icp_gpu.h"

typedef unsigned char BYTE;   // VirtualSensor.h:11

class PointCloud {
public:
    PointCloud() {}

    // PointCloud(const SimpleMesh&) (:12-39) for any mesh type of that shape: vertex.position.x()/y()/z(), triangle.idx0/idx1/idx2
    template <class Mesh, decltype(std::declval<const Mesh&>().getVertices(), 0) = 0>
    PointCloud(const Mesh& mesh) {
        const auto& vertices = mesh.getVertices();
        const auto& triangles = mesh.getTriangles();
        m_points.reserve(vertices.size());
        for (const auto& v : vertices) m_points.push_back(Vector3f(v.position.x(), v.position.y(), v.position.z()));
        std::vector<float> acc(3 * vertices.size(), 0.f);
        m_colors.assign(vertices.size(), Vector4uc(0, 0, 0, 0));
        for (const auto& t : triangles) {
            const Vector3f& a = m_points[t.idx0]; const Vector3f& b = m_points[t.idx1]; const Vector3f& c = m_points[t.idx2];
            const float u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, w[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
            const float f[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};      // face normal, not normalised (:30)
            for (unsigned k : {t.idx0, t.idx1, t.idx2}) { acc[3 * k] += f[0]; acc[3 * k + 1] += f[1]; acc[3 * k + 2] += f[2]; }
        }
        m_normals.resize(vertices.size());
        for (size_t i = 0; i < vertices.size(); ++i) {
            const float n = std::sqrt((acc[3 * i] * acc[3 * i] + acc[3 * i + 1] * acc[3 * i + 1]) + acc[3 * i + 2] * acc[3 * i + 2]);
            m_normals[i] = n > 0.f ? Vector3f(acc[3 * i] / n, acc[3 * i + 1] / n, acc[3 * i + 2] / n) : Vector3f(acc[3 * i], acc[3 * i + 1], acc[3 * i + 2]);
        }
    }

    // PointCloud(pcl::PointCloud<pcl::PointXYZ>::Ptr) (:41-76) for any pointer-like cloud with ->points[i].x / .y / .z
    template <class CloudPtr, decltype(std::declval<const CloudPtr&>()->points.size(), '0') = '0'>
    PointCloud(const CloudPtr src) {
        std::vector<Vector3f> p; p.reserve(src->points.size());
        for (const auto& q : src->points) p.push_back(Vector3f(q.x, q.y, q.z));
        *this = fromPoints(p, 5);
    }

    PointCloud(const std::vector<Vector3f>& points, const std::vector<Vector3f>& normals) : m_points(points), m_normals(normals) {
        m_colors.resize(points.size());   // PointCloud.h:26: colours zeroed when the input has none
    }
    PointCloud(const std::vector<Vector3f>& points, const std::vector<Vector3f>& normals, const std::vector<Vector4uc>& colors)
        : m_points(points), m_normals(normals), m_colors(colors) {}

    // PointCloud.h:78-165, same signature and defaults.  colorFrame is the RGBX frame (4*width*height bytes) or null.
    PointCloud(float* depthMap, BYTE* colorFrame, const Matrix3f& depthIntrinsics, const Matrix4f& depthExtrinsics, const unsigned width,
               const unsigned height, bool keepOriginalSize = false, unsigned downsampleFactor = 1, float maxDistance = 0.1f) {
        icp_gpu_ctx* ctx = nullptr;
        if (icp_gpu_create(&ctx, 0) != ICP_GPU_OK) { std::cout << "icp_gpu: no usable CUDA device (there is no CPU fallback)" << std::endl; return; }
        const size_t cap = downsampleFactor ? ((size_t)width * height + downsampleFactor - 1) / downsampleFactor : 0;
        m_points.resize(cap); m_normals.resize(cap); m_colors.resize(cap);
        int64_t n = 0;
        const int rc = icp_gpu_cloud_from_depth(ctx, depthMap, colorFrame, depthIntrinsics.data(), depthExtrinsics.data(), width, height,
                                                keepOriginalSize ? 1 : 0, downsampleFactor, maxDistance, ICP_GPU_CLOUD_ONLY,
                                                cap ? reinterpret_cast<float*>(m_points.data()) : nullptr, cap ? reinterpret_cast<float*>(m_normals.data()) : nullptr,
                                                cap ? reinterpret_cast<uint8_t*>(m_colors.data()) : nullptr, &n);
        if (rc != ICP_GPU_OK) { std::cout << "icp_gpu: " << icp_gpu_last_error(ctx) << std::endl; n = 0; }
        m_points.resize((size_t)n); m_normals.resize((size_t)n); m_colors.resize((size_t)n);
        icp_gpu_destroy(ctx);
    }

    // PointCloud(pcl::PointCloud<pcl::PointXYZ>::Ptr) (PointCloud.h:41-76) without the PCL types: points in, k = 5 PCA normals
    // (pcl::NormalEstimation, viewpoint at the origin) computed on the device, colours (255, 255, 255, 1) (:74).
    static PointCloud fromPoints(const std::vector<Vector3f>& points, int kSearch = 5) {
        PointCloud c;
        c.m_points = points;
        c.m_normals.resize(points.size());
        c.m_colors.assign(points.size(), Vector4uc(255, 255, 255, 1));
        icp_gpu_ctx* ctx = nullptr;
        if (icp_gpu_create(&ctx, 0) != ICP_GPU_OK) { std::cout << "icp_gpu: no usable CUDA device (there is no CPU fallback)" << std::endl; return c; }
        int rc = points.empty() ? ICP_GPU_OK : icp_gpu_set_target(ctx, reinterpret_cast<const float*>(points.data()), nullptr, nullptr, (int64_t)points.size());
        if (rc == ICP_GPU_OK && !points.empty()) rc = icp_gpu_target_normals(ctx, kSearch, nullptr, reinterpret_cast<float*>(c.m_normals.data()), nullptr);
        if (rc != ICP_GPU_OK) std::cout << "icp_gpu: " << icp_gpu_last_error(ctx) << std::endl;
        icp_gpu_destroy(ctx);
        return c;
    }

    std::vector<Vector3f>& getPoints() { return m_points; }
    const std::vector<Vector3f>& getPoints() const { return m_points; }
    std::vector<Vector3f>& getNormals() { return m_normals; }
    const std::vector<Vector3f>& getNormals() const { return m_normals; }
    std::vector<Vector4uc>& getColors() { return m_colors; }
    const std::vector<Vector4uc>& getColors() const { return m_colors; }

    bool readFromFile(const std::string& filename) {            // :167-217: char nBytes, uint n, then n points and n normals as float or double triples
        std::ifstream is(filename, std::ios::in | std::ios::binary);
        if (!is.is_open()) { std::cout << "ERROR: unable to read input file!" << std::endl; return false; }
        char nBytes = 0; unsigned int n = 0;
        is.read(&nBytes, sizeof(char)); is.read((char*)&n, sizeof(unsigned int));
        if (!is || (nBytes != (char)sizeof(float) && nBytes != (char)sizeof(double))) return false;
        for (int part = 0; part < 2; ++part) {
            std::vector<Vector3f>& dst = part == 0 ? m_points : m_normals;
            if (nBytes == (char)sizeof(float)) {
                std::vector<float> ps(3 * (size_t)n); is.read((char*)ps.data(), (std::streamsize)(ps.size() * sizeof(float)));
                for (unsigned int i = 0; i < n; i++) dst.push_back(Vector3f(ps[3 * i], ps[3 * i + 1], ps[3 * i + 2]));
            } else {
                std::vector<double> ps(3 * (size_t)n); is.read((char*)ps.data(), (std::streamsize)(ps.size() * sizeof(double)));
                for (unsigned int i = 0; i < n; i++) dst.push_back(Vector3f((float)ps[3 * i], (float)ps[3 * i + 1], (float)ps[3 * i + 2]));
            }
        }
        m_colors.resize(m_points.size());
        return (bool)is;
    }
    // :219-236 writes a PLY through pcl::io::savePLYFile; this one needs no PCL: ascii PLY with x y z nx ny nz
    bool writeToFile(const std::string& filename) {
        std::ofstream os(filename);
        if (!os.is_open()) return false;
        os << "ply\nformat ascii 1.0\nelement vertex " << m_points.size()
           << "\nproperty float x\nproperty float y\nproperty float z\nproperty float normal_x\nproperty float normal_y\nproperty float normal_z\nend_header\n";
        for (size_t i = 0; i < m_points.size(); ++i) {
            const Vector3f n = i < m_normals.size() ? m_normals[i] : Vector3f(0.f, 0.f, 0.f);
            os << m_points[i][0] << " " << m_points[i][1] << " " << m_points[i][2] << " " << n[0] << " " << n[1] << " " << n[2] << "\n";
        }
        return (bool)os;
    }
    PointCloud copy_point_cloud() { return *this; }             // :263-275

    // :277-283 on the device: points by the pose, normals by its rotation block (pose * (n, 0))
    void change_pose(const Matrix4f& pose) {
        if (m_points.empty()) return;
        icp_gpu_ctx* ctx = icp_b200::sharedContext();
        if (!ctx) return;
        Matrix4f rot = pose; rot(0, 3) = 0.f; rot(1, 3) = 0.f; rot(2, 3) = 0.f;
        std::vector<Vector3f> p(m_points.size()), n(m_normals.size());
        int rc = icp_gpu_transform_points(ctx, pose.data(), reinterpret_cast<const float*>(m_points.data()), (int64_t)m_points.size(), reinterpret_cast<float*>(p.data()));
        if (rc == ICP_GPU_OK && !m_normals.empty())
            rc = icp_gpu_transform_points(ctx, rot.data(), reinterpret_cast<const float*>(m_normals.data()), (int64_t)m_normals.size(), reinterpret_cast<float*>(n.data()));
        if (!icp_b200::report(ctx, rc, "PointCloud::change_pose")) return;
        m_points.swap(p); m_normals.swap(n);
    }

    unsigned int getClosestPoint(Vector3f& p) {                 // :310-322 (first minimum of the Euclidean norm)
        unsigned int idx = 0; float best = std::numeric_limits<float>::max();
        for (unsigned int i = 0; i < m_points.size(); ++i) {
            const float dx = p[0] - m_points[i][0], dy = p[1] - m_points[i][1], dz = p[2] - m_points[i][2];
            const float d = std::sqrt((dx * dx + dy * dy) + dz * dz);
            if (best > d) { idx = i; best = d; }
        }
        return idx;
    }

    // PointCloud.h:325-343: every `stride`-th point whose point and normal are finite.  (The device loop
    // builds its pyramid levels itself; this host version exists for callers that use it directly.)
    PointCloud getCoarseResolution(int stride) const {
        PointCloud c;
        for (size_t i = 0; i < m_points.size(); i += (size_t)stride) {
            const Vector3f& p = m_points[i]; const Vector3f& n = m_normals[i];
            if (std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]) && std::isfinite(n[0]) && std::isfinite(n[1]) && std::isfinite(n[2])) {
                c.m_points.push_back(p); c.m_normals.push_back(n);
                if (i < m_colors.size()) c.m_colors.push_back(m_colors[i]);
            }
        }
        return c;
    }

private:
    std::vector<Vector3f> m_points;
    std::vector<Vector3f> m_normals;
    std::vector<Vector4uc> m_colors;
};
