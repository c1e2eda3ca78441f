// peak.cu -- the FP32 roofline denominators of this device, measured: the k-NN kernels do non-tensor fp32 arithmetic, and
// MEASURED_PEAKS.json only holds the HBM copy bandwidth and the bf16 tensor throughput.  Two streams of independent
// register-resident chains per thread, all warp slots occupied:
//   mode 0  FFMA            2 flop per instruction (the usual "peak" figure)
//   mode 1  FMUL + FADD     1 flop per instruction -- what contract D1's un-fused squared distances can reach at best
// Reported as TFLOP/s from CUDA-event time on the context's stream (best of a few repetitions).
#include "icp_internal.cuh"

template <int MODE>
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* __restrict__ out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = (float)(threadIdx.x + k) * 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (MODE == 0) x[k] = __fmaf_rn(x[k], a, b);
            else { x[k] = __fmul_rn(x[k], a); x[k] = __fadd_rn(x[k], b); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += x[k];
    if (s == 123456.789f) out[0] = s;           // never true: keeps the chains alive
}

cudaError_t icp_measure_fp32_peak(int mode, int n_sms, cudaStream_t s, double* tflops) {
    float* out = nullptr;
    cudaError_t e = cudaMalloc(&out, 64);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, blocks = n_sms * 8;      // 8 x 256 threads = every warp slot of an SM
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, s);
        if (mode == 0) fp32_peak_kernel<0><<<blocks, 256, 0, s>>>(out, iters, 0.999f, 1e-4f);
        else fp32_peak_kernel<1><<<blocks, 256, 0, s>>>(out, iters, 0.999f, 1e-4f);
        cudaEventRecord(e1, s);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks;    // mode 0: 1 instruction = 2 flop; mode 1: 2 instructions = 2 flop
        if (rep > 0 && ms > 0.f) { const double t = flop / (ms * 1e-3) / 1e12; if (t > best) best = t; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (e == cudaSuccess) e = cudaGetLastError();
    *tflops = best;
    return e;
}
