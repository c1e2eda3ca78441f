#pragma once
// oracle/ref_shim: stand-in header (test infrastructure only); everything lives in pcl/point_types.h.
#include <pcl/point_types.h>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#ifndef PCL_ERROR
#define PCL_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
#endif
namespace pcl { namespace io {
// ascii .pcd with x y z as the first three fields (the ETH "Challenging data sets" exports); returns 0 / -1 like PCL
template <class P> int loadPCDFile(const std::string& path, pcl::PointCloud<P>& cloud) {
    std::ifstream is(path);
    if (!is.is_open()) return -1;
    std::string line; bool data = false; cloud.points.clear();
    while (std::getline(is, line)) {
        if (!data) { if (line.compare(0, 4, "DATA") == 0) { if (line.find("ascii") == std::string::npos) return -1; data = true; } continue; }
        std::istringstream ss(line); P p; if (ss >> p.x >> p.y >> p.z) cloud.points.push_back(p);
    }
    cloud.width = (unsigned)cloud.points.size(); cloud.height = 1;
    return data ? 0 : -1;
}
} }
