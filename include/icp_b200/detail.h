// detail.h -- one process-wide device context for the reference API's value classes (WeightingMethod, ProcrustesAligner,
// transformPoints, PointCloud::change_pose, ConvergenceMeasure): the reference's versions are free-standing CPU code, the
// drop-in's are thin calls into the icp_gpu_* C ABI and need a context to run on.  There is no CPU fallback: without a usable
// sm_100-class device the context is null and every caller reports that and returns its input unchanged / an empty result.
#pragma once
#include <iostream>
#include "../icp_gpu.h"

namespace icp_b200 {
struct SharedContext {
    icp_gpu_ctx* ctx;
    SharedContext() : ctx(nullptr) {
        if (icp_gpu_create(&ctx, 0) != ICP_GPU_OK) { ctx = nullptr; std::cout << "icp_gpu: no usable CUDA device (there is no CPU fallback)." << std::endl; }
    }
    ~SharedContext() { if (ctx) icp_gpu_destroy(ctx); }
    SharedContext(const SharedContext&) = delete;
    SharedContext& operator=(const SharedContext&) = delete;
};
// function-local static of an inline function: one instance per process, created on first use
inline icp_gpu_ctx* sharedContext() { static SharedContext holder; return holder.ctx; }
inline bool report(icp_gpu_ctx* ctx, int rc, const char* what) {
    if (rc == ICP_GPU_OK) return true;
    std::cout << "icp_gpu: " << what << ": " << (ctx ? icp_gpu_last_error(ctx) : "no device") << std::endl;
    return false;
}
}  // namespace icp_b200
