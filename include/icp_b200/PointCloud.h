// PointCloud.h -- the container estimatePose receives (reference: icp-variants/PointCloud.h).
// The storage and its accessors are the hot path's boundary.  Of the reference's constructors the depth-map
// one (PointCloud.h:78-165, the step right before the loop in reconstructRoom) is provided, computed on the
// device through icp_gpu_cloud_from_depth; construction from meshes and PCD files (PointCloud.h:12-76) stays
// with the caller (file I/O and PCL).
#pragma once
#include "Eigen.h"
#include "../icp_gpu.h"

typedef unsigned char BYTE;   // VirtualSensor.h:11

class PointCloud {
public:
    PointCloud() {}
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
