// PointCloud.h -- the container estimatePose receives (reference: icp-variants/PointCloud.h).
// Only the storage and its accessors are part of the hot path's boundary; construction from meshes,
// PCD files and depth maps (PointCloud.h:12-165) is input preparation and stays with the caller.
#pragma once
#include "Eigen.h"

class PointCloud {
public:
    PointCloud() {}
    PointCloud(const std::vector<Vector3f>& points, const std::vector<Vector3f>& normals) : m_points(points), m_normals(normals) {
        m_colors.resize(points.size());   // PointCloud.h:26: colours zeroed when the input has none
    }
    PointCloud(const std::vector<Vector3f>& points, const std::vector<Vector3f>& normals, const std::vector<Vector4uc>& colors)
        : m_points(points), m_normals(normals), m_colors(colors) {}

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
