// solve.cu -- stages 5-6 of one ICP iteration: residual / Jacobian assembly reduced straight to the
// normal equations (never materialising the reference's 4M x 6 matrix A), the 6x6 solve (or the 3x3
// Procrustes SVD), and the pose update, in ONE launch per iteration:
//   every block: per-thread fp64 accumulators -> recursive-halving warp reduce-scatter (62 shuffles
//   for 32 values) -> fixed-order block sum -> partial row in global memory;
//   last block (atomic ticket): fixed-order sum of all partial rows -> solve -> pose <- inc * pose.
// Deterministic: no floating-point atomics anywhere.
//
// Reference: gather (ICPOptimizer.h:583-610), estimatePosePointToPoint -> ProcrustesAligner
// (ICPOptimizer.h:666-674, ProcrustesAligner.h:6-70), estimatePosePointToPlane (:676-782),
// estimatePoseSymmetricICP (:784-898), pose accumulation (:614-620).
#include "icp_internal.cuh"
#include <stdlib.h>

struct PoseSmR { float P[16]; };

// One-sided Jacobi SVD of a 3x3 (row-major), singular values descending, A = U diag(S) V^T.
__device__ void svd3_dev(const double* A, double* U, double* S, double* V) {
    double W[9];
    for (int i = 0; i < 9; ++i) { W[i] = A[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
            double a = 0, b = 0, c = 0;
            for (int k = 0; k < 3; ++k) { a += W[k * 3 + p] * W[k * 3 + p]; b += W[k * 3 + q] * W[k * 3 + q]; c += W[k * 3 + p] * W[k * 3 + q]; }
            if (fabs(c) <= 1e-300 || fabs(c) <= 1e-17 * sqrt(a * b)) continue;
            off += fabs(c);
            const double zeta = (b - a) / (2.0 * c);
            const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
            for (int k = 0; k < 3; ++k) {
                const double wp = W[k * 3 + p], wq = W[k * 3 + q];
                W[k * 3 + p] = cs * wp - sn * wq; W[k * 3 + q] = sn * wp + cs * wq;
                const double vp = V[k * 3 + p], vq = V[k * 3 + q];
                V[k * 3 + p] = cs * vp - sn * vq; V[k * 3 + q] = sn * vp + cs * vq;
            }
        }
        if (off == 0.0) break;
    }
    for (int j = 0; j < 3; ++j) S[j] = sqrt(W[j] * W[j] + W[3 + j] * W[3 + j] + W[6 + j] * W[6 + j]);
    int ord[3] = {0, 1, 2};
    for (int i = 0; i < 2; ++i) for (int j = i + 1; j < 3; ++j) if (S[ord[j]] > S[ord[i]]) { const int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    double Ws[9], Vs[9], Ss[3];
    for (int j = 0; j < 3; ++j) { Ss[j] = S[ord[j]]; for (int k = 0; k < 3; ++k) { Ws[k * 3 + j] = W[k * 3 + ord[j]]; Vs[k * 3 + j] = V[k * 3 + ord[j]]; } }
    for (int j = 0; j < 3; ++j) S[j] = Ss[j];
    for (int i = 0; i < 9; ++i) V[i] = Vs[i];
    const double tiny = 1e-14 * (S[0] > 0 ? S[0] : 1.0);
    for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) U[k * 3 + j] = S[j] > tiny ? Ws[k * 3 + j] / S[j] : 0.0;
    if (!(S[0] > tiny)) { for (int i = 0; i < 9; ++i) U[i] = (i % 4 == 0) ? 1.0 : 0.0; return; }
    if (!(S[1] > tiny)) {   // complete a rank-1 U
        const double u0[3] = {U[0], U[3], U[6]};
        const int m = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2) : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
        double e[3] = {0, 0, 0}; e[m] = 1.0;
        const double u1[3] = {u0[1] * e[2] - u0[2] * e[1], u0[2] * e[0] - u0[0] * e[2], u0[0] * e[1] - u0[1] * e[0]};
        const double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
        for (int k = 0; k < 3; ++k) U[k * 3 + 1] = u1[k] / n1;
    }
    if (!(S[2] > tiny)) {
        const double u0[3] = {U[0], U[3], U[6]}, u1[3] = {U[1], U[4], U[7]};
        U[2] = u0[1] * u1[2] - u0[2] * u1[1]; U[5] = u0[2] * u1[0] - u0[0] * u1[2]; U[8] = u0[0] * u1[1] - u0[1] * u1[0];
    }
}

// ---------------------------------------------------------------------------- finishers: summed row -> increment
// Row layouts.  PLANE / SYMMETRIC: [0..20] A^T A upper triangle row-major, [21..26] A^T b, [27] count.
//               P2P: [0] count, [1..3] sum s, [4..6] sum d, [7] sum w, [8..10] sum w s, [11..13] sum w d, [14..22] sum w d s^T.
//               SUMS: [0] count, [1..3] sum s, [4..6] sum d.
// 6x6 Cholesky solve of the normal equations (symmetric positive definite by construction), every loop fully
// unrolled so that the factor lives in registers -- one thread runs this on the critical path of every iteration,
// and a local-memory (dynamically indexed) elimination costs ~10 us there.  A non-positive pivot (degenerate
// geometry) falls back to Gaussian elimination with partial pivoting, the solver the oracle uses.
__device__ __forceinline__ bool chol6_solve(const double* row, double lambda2, double* x) {
    double L[6][6];
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i; j < 6; ++j) { L[j][i] = row[k]; ++k; }        // lower triangle
#pragma unroll
    for (int i = 0; i < 6; ++i) L[i][i] += lambda2;
    bool ok = true;
    double Linv[6];          // 1 / L[j][j]: one rsqrt per column instead of a sqrt and 3 divisions (this is the serial tail
                             // of every iteration: one thread, every dependent fp64 division costs ~100 cycles)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = L[j][j];
#pragma unroll
        for (int m = 0; m < j; ++m) d -= L[j][m] * L[j][m];
        if (!(d > 0.0)) ok = false;
        const double inv = rsqrt(d);
        Linv[j] = inv;
        L[j][j] = d * inv;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double v = L[i][j];
#pragma unroll
            for (int m = 0; m < j; ++m) v -= L[i][m] * L[j][m];
            L[i][j] = v * inv;
        }
    }
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double v = row[21 + i];
#pragma unroll
        for (int m = 0; m < i; ++m) v -= L[i][m] * y[m];
        y[i] = v * Linv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double v = y[i];
#pragma unroll
        for (int m = i + 1; m < 6; ++m) v -= L[m][i] * x[m];
        x[i] = v * Linv[i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) if (!isfinite(x[i])) ok = false;
    return ok;
}

__device__ __noinline__ int expand_and_solve_pivoted(const double* row, double lambda2, double* x) {
    double A[36], b[6];
    int k = 0;
    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { A[i * 6 + j] = row[k]; A[j * 6 + i] = row[k]; ++k; }
    for (int i = 0; i < 6; ++i) { A[i * 6 + i] += lambda2; b[i] = row[21 + i]; }
    return solve6_dev(A, b, x);
}

__device__ int expand_and_solve(const double* row, double lambda2, double* x) {
    if (chol6_solve(row, lambda2, x)) return 0;
    return expand_and_solve_pivoted(row, lambda2, x);
}

__device__ int finish_p2plane(const double* row, float* inc) {
    mat4_identity_dev(inc);
    if (!(row[27] > 0.0)) return ICP_GPU_E_NO_MATCHES;
    double x[6];
    if (expand_and_solve(row, 0.0, x) != 0) return ICP_GPU_E_NUMERIC;
    // ICPOptimizer.h:768-779: R = Rx(alpha) * Ry(beta) * Rz(gamma) in fp32, t = x[3..5]
    const float al = (float)x[0], be = (float)x[1], ga = (float)x[2];
    double sd, cd;
    sincos((double)al, &sd, &cd); const float ca = (float)cd, sa = (float)sd;
    sincos((double)be, &sd, &cd); const float cb = (float)cd, sb = (float)sd;
    sincos((double)ga, &sd, &cd); const float cg = (float)cd, sg = (float)sd;
    float Rx[16], Ry[16], Rz[16], T[16];
    mat4_identity_dev(Rx); mat4_identity_dev(Ry); mat4_identity_dev(Rz);
    Rx[5] = ca; Rx[9] = -sa; Rx[6] = sa; Rx[10] = ca;
    Ry[0] = cb; Ry[8] = sb; Ry[2] = -sb; Ry[10] = cb;
    Rz[0] = cg; Rz[4] = -sg; Rz[1] = sg; Rz[5] = cg;
    mat4_mul_pinned(Rx, Ry, T); mat4_mul_pinned(T, Rz, inc);
    inc[12] = (float)x[3]; inc[13] = (float)x[4]; inc[14] = (float)x[5];
    return 0;
}

__device__ int finish_symmetric(const double* row, const float* meanS, const float* meanT, float* inc) {
    mat4_identity_dev(inc);
    if (!(row[27] > 0.0)) return ICP_GPU_E_NO_MATCHES;
    const float lambda = 0.0001f;                                   // ICPOptimizer.h:858-864
    double x[6];
    if (expand_and_solve(row, (double)pmul(lambda, lambda), x) != 0) return ICP_GPU_E_NUMERIC;
    // ICPOptimizer.h:876-895
    const float at0 = (float)x[0], at1 = (float)x[1], at2 = (float)x[2];
    const float tt[3] = {(float)x[3], (float)x[4], (float)x[5]};
    const float tan_theta = __fsqrt_rn(padd(padd(pmul(at0, at0), pmul(at1, at1)), pmul(at2, at2)));
    float R4[16]; mat4_identity_dev(R4);
    float cos_theta = 1.0f;
    if (tan_theta > 0.f) {       // the reference divides 0/0 here; guarded (documented deviation)
        const float a[3] = {pdiv(at0, tan_theta), pdiv(at1, tan_theta), pdiv(at2, tan_theta)};
        const float sin_theta = (float)((double)tan_theta / sqrt(1.0 + (double)pmul(tan_theta, tan_theta)));
        cos_theta = pdiv(sin_theta, tan_theta);
        const float K[9] = {0.f, -a[2], a[1], a[2], 0.f, -a[0], -a[1], a[0], 0.f};   // row-major [a]x
        const float omc = psub(1.0f, cos_theta);
        float K1[9], KK[9];
        for (int i = 0; i < 9; ++i) K1[i] = pmul(omc, K[i]);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
            KK[r * 3 + c] = padd(padd(pmul(K1[r * 3], K[c]), pmul(K1[r * 3 + 1], K[3 + c])), pmul(K1[r * 3 + 2], K[6 + c]));
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
            R4[r + 4 * c] = padd((r == c) ? 1.0f : 0.0f, padd(pmul(sin_theta, K[r * 3 + c]), KK[r * 3 + c]));   // getRodriguesMatrix, utils.h:171-176
    }
    float Td[16], Tt[16], Ts[16], M1[16], M2[16], M3[16];
    mat4_identity_dev(Td); mat4_identity_dev(Tt); mat4_identity_dev(Ts);
    for (int k = 0; k < 3; ++k) { Td[12 + k] = meanT[k]; Tt[12 + k] = pmul(tt[k], cos_theta); Ts[12 + k] = -meanS[k]; }
    mat4_mul_pinned(Td, R4, M1); mat4_mul_pinned(M1, Tt, M2); mat4_mul_pinned(M2, R4, M3); mat4_mul_pinned(M3, Ts, inc);
    return 0;
}

__device__ int finish_p2p(const double* row, float* inc) {
    mat4_identity_dev(inc);
    const double M = row[0];
    if (!(M > 0.0)) return ICP_GPU_E_NO_MATCHES;
    double sm[3], dm[3];
    for (int k = 0; k < 3; ++k) { sm[k] = row[1 + k] / M; dm[k] = row[4 + k] / M; }
    // ProcrustesAligner.h:50-54 with unweighted means: sum w (d - dm)(s - sm)^T expanded in raw moments
    double A[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
        A[r * 3 + c] = row[14 + r * 3 + c] - dm[r] * row[8 + c] - row[11 + r] * sm[c] + row[7] * dm[r] * sm[c];
    double U[9], S[3], V[9];
    svd3_dev(A, U, S, V);
    double UVt[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) UVt[r * 3 + c] = U[r * 3] * V[c * 3] + U[r * 3 + 1] * V[c * 3 + 1] + U[r * 3 + 2] * V[c * 3 + 2];
    const double dd = UVt[0] * (UVt[4] * UVt[8] - UVt[5] * UVt[7]) - UVt[1] * (UVt[3] * UVt[8] - UVt[5] * UVt[6]) + UVt[2] * (UVt[3] * UVt[7] - UVt[4] * UVt[6]);
    double R[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r * 3 + c] = U[r * 3] * V[c * 3] + U[r * 3 + 1] * V[c * 3 + 1] + dd * U[r * 3 + 2] * V[c * 3 + 2];
    for (int r = 0; r < 3; ++r) {
        double rt = 0, rd = 0;
        for (int c = 0; c < 3; ++c) { rt += R[r * 3 + c] * (dm[c] - sm[c]); rd += R[r * 3 + c] * dm[c]; }
        inc[r + 12] = (float)(rt - rd + dm[r]);                       // ProcrustesAligner.h:26
        for (int c = 0; c < 3; ++c) inc[r + 4 * c] = (float)R[r * 3 + c];
    }
    for (int i = 0; i < 16; ++i) if (!isfinite(inc[i])) return ICP_GPU_E_NUMERIC;
    return 0;
}

__device__ void finish_row(DevState* st, const double* row, int mode, float* history) {
    float inc[16]; int rc;
    if (mode == 0) rc = finish_p2p(row, inc);
    else if (mode == 1) rc = finish_p2plane(row, inc);
    else rc = finish_symmetric(row, st->mean_s, st->mean_d, inc);
    apply_increment(st, inc, rc, history);
}

// The same for the last block of reduce_kernel: the current pose and the means come from the block's shared-memory copies and
// the three loop counters from registers (loaded at kernel start), so the serial tail makes no dependent global-memory round
// trip; everything is written once at the end.
struct LoopRegs { int status, iters_done, iter; };

__device__ __forceinline__ void finish_row_local(DevState* st, const double* row, int mode, float* history, const float* P, const float* mS,
                                                 const float* mD, const LoopRegs lr) {
    float inc[16]; int rc;
    if (mode == 0) rc = finish_p2p(row, inc);
    else if (mode == 1) rc = finish_p2plane(row, inc);
    else rc = finish_symmetric(row, mS, mD, inc);
    if (lr.status == 0) {
        if (rc != 0) st->status = rc;
        else {
            float np[16], nm[9];
            mat4_mul_pinned(inc, P, np);                       // ICPOptimizer.h:614-620
            inv_transpose3_pinned(np, nm);
#pragma unroll
            for (int i = 0; i < 16; ++i) st->pose[i] = np[i];
#pragma unroll
            for (int i = 0; i < 9; ++i) st->nrm[i] = nm[i];
            if (history) {
#pragma unroll
                for (int i = 0; i < 16; ++i) history[16 * lr.iters_done + i] = np[i];
            }
            st->iters_done = lr.iters_done + 1;
            if (increment_is_small(inc, st->stop_rot, st->stop_trans)) st->converged = 1;
        }
    }
    st->iter = lr.iter + 1;
}

__device__ void finish_sums(DevState* st, const double* row) {
    // unweighted means of the kept matches, rounded to fp32 (ICPOptimizer.h:797-798)
    const double M = row[0];
    for (int k = 0; k < 3; ++k) {
        st->mean_s[k] = M > 0.0 ? (float)(row[1 + k] / M) : 0.f;
        st->mean_d[k] = M > 0.0 ? (float)(row[4 + k] / M) : 0.f;
    }
}

#ifdef ICP_TIMELINE   // diagnostic build (profiles/): see match.cu; per block of iteration 12 {start, loop end}, [511] = {pose written, 0}
__device__ unsigned long long g_timeline_r[512][2];
__device__ __forceinline__ unsigned long long tlr_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
extern "C" int icp_gpu_debug_timeline_reduce(unsigned long long* out, int reset) {
    if (reset) { void* p = nullptr; cudaGetSymbolAddress(&p, g_timeline_r); return (int)cudaMemset(p, 0, sizeof(g_timeline_r)); }
    return (int)cudaMemcpyFromSymbol(out, g_timeline_r, sizeof(g_timeline_r));
}
#endif

// ---------------------------------------------------------------------------- the reduction kernel
// MODE 0 p2p moments, 1 point-to-plane, 2 symmetric (needs means in state), 3 sums only
template <int MODE, bool FUSED>
__global__ void __launch_bounds__(ICP_REDUCE_THREADS) reduce_kernel(const ReduceArgs a) {
    __shared__ float P[16];
    __shared__ float mS[3], mD[3], Nm[9];
    __shared__ double red[ICP_REDUCE_THREADS / 32][32];
    __shared__ double fin[ICP_REDUCE_THREADS / 32][32];
    __shared__ bool is_last;
    unsigned long long t_start = 0, t_loop = 0;
    if (a.state->converged) return;                          // early stop reached (uniform across the grid)
    if (a.profile && threadIdx.x == 0) { t_start = global_timer_ns(); if (blockIdx.x == 0) a.state->prof[0] = t_start; }
    // the first point's independent loads are requested before the pose / descriptor loads and the barrier
    const int stride = gridDim.x * blockDim.x;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int pos_n = -1; float wf_n = 0.f;
    float4 sp_n = make_float4(0.f, 0.f, 0.f, 0.f), sn_n = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < a.n_src) {
        pos_n = FUSED ? a.nn_pos[i] : a.match_pos[i];
        sp_n = __ldg(&a.src_pts[i]);
        if (FUSED || MODE == 2) sn_n = __ldg(&a.src_nrm[i]);
        if (!FUSED) wf_n = a.match_w[i];
    }
    const int state_iter = a.state->iter;
#ifdef ICP_TIMELINE
    const bool tl_on = state_iter == 12 && threadIdx.x == 0 && blockIdx.x < 511;
    if (tl_on) g_timeline_r[blockIdx.x][0] = tlr_now();
#endif
    LoopRegs lr; lr.status = 0; lr.iters_done = 0; lr.iter = state_iter;
    if (threadIdx.x == 0) { lr.status = a.state->status; lr.iters_done = a.state->iters_done; }
    if (threadIdx.x < 16) P[threadIdx.x] = a.state->pose[threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 35) { mS[threadIdx.x - 32] = a.state->mean_s[threadIdx.x - 32]; mD[threadIdx.x - 32] = a.state->mean_d[threadIdx.x - 32]; }
    if (threadIdx.x >= 64 && threadIdx.x < 73) Nm[threadIdx.x - 64] = a.state->nrm[threadIdx.x - 64];
    __syncthreads();
    double v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = 0.0;
    const double LP = (double)0.1f, LQ = (double)1.0f;   // LAMBDA_POINT / LAMBDA_PLANE|SYMMETRIC (ICPOptimizer.h:737-738, :840-841)
    // queries are addressed by their position in the Morton-sorted source
    IterDesc d; d.stride = 1; d.filter_finite = 0; d.mask_word_offset = -1; d.rng_key = 0u; d.proba = -1.0f;
    if (FUSED) d = a.desc[a.desc_index >= 0 ? a.desc_index : state_iter];
    // Software pipeline over the thread's points: the loads that depend on nothing (search result / match record,
    // source point and normal) are issued one point ahead, so the only latency a point exposes is its gather of the
    // matched target point -- the kernel is latency-bound (4 warps per scheduler at 128 registers), not HBM-bound.
    for (; i < a.n_src; i += stride) {
        const int pos = pos_n; float wf = wf_n; const float4 sp = sp_n, sn4 = sn_n;
        // a query without a (surviving) match has pos -1
        const bool gather = pos >= 0 && (!FUSED || pos < a.n_tgt);
        float4 tp = make_float4(0.f, 0.f, 0.f, 0.f), tn = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gather) {
            tp = __ldg(&a.tgt_pts[pos]);
            if (FUSED || MODE == 1 || MODE == 2) tn = __ldg(&a.tgt_nrm[pos]);
        }
        {
            const int i2 = i + stride;
            if (i2 < a.n_src) {
                pos_n = FUSED ? a.nn_pos[i2] : a.match_pos[i2];
                sp_n = __ldg(&a.src_pts[i2]);
                if (FUSED || MODE == 2) sn_n = __ldg(&a.src_nrm[i2]);
                if (!FUSED) wf_n = a.match_w[i2];
            }
        }
        if (!gather) continue;
        float sxf, syf, szf;
        if (FUSED) {
            // stages 3-4 evaluated here from the search result (same code path as match_finish_kernel)
            if (!query_active(d, a.mask, sp, sn4)) continue;
            xform_point(P, sp.x, sp.y, sp.z, sxf, syf, szf);
            if (!finite3(sxf, syf, szf)) continue;
            float nx, ny, nz;
            xform_normal(Nm, sn4.x, sn4.y, sn4.z, nx, ny, nz);
            wf = 1.0f;
            if (!match_weight_and_reject(a.weighting, a.rejection, a.weight_max_d2, sxf, syf, szf, nx, ny, nz, __float_as_uint(sn4.w), tp, tn, wf)) continue;
        } else {
            xform_point(P, sp.x, sp.y, sp.z, sxf, syf, szf);
        }
        if (!finite3(sxf, syf, szf) || !finite3(tp.x, tp.y, tp.z)) continue;        // ICPOptimizer.h:590-592
        const double w = (double)wf;
        if (MODE == 3) {
            v[0] += 1.0; v[1] += sxf; v[2] += syf; v[3] += szf; v[4] += tp.x; v[5] += tp.y; v[6] += tp.z;
        } else if (MODE == 0) {
            const double s[3] = {sxf, syf, szf}, t[3] = {tp.x, tp.y, tp.z};
            v[0] += 1.0; v[7] += w;
#pragma unroll
            for (int k = 0; k < 3; ++k) { v[1 + k] += s[k]; v[4 + k] += t[k]; v[8 + k] += w * s[k]; v[11 + k] += w * t[k]; }
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) v[14 + r * 3 + c] += w * t[r] * s[c];
        } else {
            double s[3] = {sxf, syf, szf}, t[3] = {tp.x, tp.y, tp.z};
            double n[3] = {tn.x, tn.y, tn.z};
            bool use_row = finite3(tn.x, tn.y, tn.z);
            double u[3] = {s[0], s[1], s[2]};
            if (MODE == 2) {
                // the source normal is transformed by the current pose's inverse-transpose (ICPOptimizer.h:554)
                float nx, ny, nz;
                xform_normal(Nm, sn4.x, sn4.y, sn4.z, nx, ny, nz);
                use_row = use_row && finite3(nx, ny, nz);
#pragma unroll
                for (int k = 0; k < 3; ++k) { s[k] -= (double)mS[k]; t[k] -= (double)mD[k]; }   // :797-806
                n[0] += (double)nx; n[1] += (double)ny; n[2] += (double)nz;                      // n_t + n_s
#pragma unroll
                for (int k = 0; k < 3; ++k) u[k] = s[k] + t[k];
            }
            const double e[3] = {t[0] - s[0], t[1] - s[1], t[2] - s[2]};
            const double b2 = (LP * w) * (LP * w);
            // three point rows [-[s]x | I], rhs e  (:716-733 / :817-835):  P^T P = [[|s|^2 I - s s^T, [s]x], [-[s]x, I]]
            const double ss = s[0] * s[0] + s[1] * s[1] + s[2] * s[2];
            double c[6] = {0, 0, 0, 0, 0, 0}; double r = 0.0, a2 = 0.0;
            if (use_row) {
                // plane row [u x n | n], rhs n.e  (:698-710 ; symmetric :809-815 with u = s~ + d~, n = n_t + n_s)
                c[0] = u[1] * n[2] - u[2] * n[1]; c[1] = u[2] * n[0] - u[0] * n[2]; c[2] = u[0] * n[1] - u[1] * n[0];
                c[3] = n[0]; c[4] = n[1]; c[5] = n[2];
                r = n[0] * e[0] + n[1] * e[1] + n[2] * e[2];
                a2 = (LQ * w) * (LQ * w);
            }
            // upper triangle, row-major
            v[0] += a2 * c[0] * c[0] + b2 * (ss - s[0] * s[0]);
            v[1] += a2 * c[0] * c[1] - b2 * s[0] * s[1];
            v[2] += a2 * c[0] * c[2] - b2 * s[0] * s[2];
            v[3] += a2 * c[0] * c[3];
            v[4] += a2 * c[0] * c[4] - b2 * s[2];
            v[5] += a2 * c[0] * c[5] + b2 * s[1];
            v[6] += a2 * c[1] * c[1] + b2 * (ss - s[1] * s[1]);
            v[7] += a2 * c[1] * c[2] - b2 * s[1] * s[2];
            v[8] += a2 * c[1] * c[3] + b2 * s[2];
            v[9] += a2 * c[1] * c[4];
            v[10] += a2 * c[1] * c[5] - b2 * s[0];
            v[11] += a2 * c[2] * c[2] + b2 * (ss - s[2] * s[2]);
            v[12] += a2 * c[2] * c[3] - b2 * s[1];
            v[13] += a2 * c[2] * c[4] + b2 * s[0];
            v[14] += a2 * c[2] * c[5];
            v[15] += a2 * c[3] * c[3] + b2;
            v[16] += a2 * c[3] * c[4];
            v[17] += a2 * c[3] * c[5];
            v[18] += a2 * c[4] * c[4] + b2;
            v[19] += a2 * c[4] * c[5];
            v[20] += a2 * c[5] * c[5] + b2;
            // rhs: a^2 r c + b^2 [s x d ; e]   (s x e == s x d)
            v[21] += a2 * r * c[0] + b2 * (s[1] * e[2] - s[2] * e[1]);
            v[22] += a2 * r * c[1] + b2 * (s[2] * e[0] - s[0] * e[2]);
            v[23] += a2 * r * c[2] + b2 * (s[0] * e[1] - s[1] * e[0]);
            v[24] += a2 * r * c[3] + b2 * e[0];
            v[25] += a2 * r * c[4] + b2 * e[1];
            v[26] += a2 * r * c[5] + b2 * e[2];
            v[27] += 1.0;
        }
    }
    if (a.profile && threadIdx.x == 0) t_loop = global_timer_ns();
#ifdef ICP_TIMELINE
    if (tl_on) g_timeline_r[blockIdx.x][1] = tlr_now();
#endif
    if (!grid_reduce_row<ICP_REDUCE_THREADS>(v, a.partials, &a.state->ticket, red, fin, &is_last, a.solve ? &a.peer : nullptr, a.state)) return;
    if (threadIdx.x == 0) {
        if (a.profile) { a.state->prof[1] = t_start; a.state->prof[2] = t_loop; a.state->prof[3] = global_timer_ns(); }
        if (!a.solve) { for (int k = 0; k < ICP_NRED; ++k) a.state->shard_partials[k] = fin[0][k]; }
        else if (MODE == 3) finish_sums(a.state, fin[0]);
        else finish_row_local(a.state, fin[0], MODE, a.pose_history, P, mS, mD, lr);
        if (a.profile) { __threadfence(); a.state->prof[4] = global_timer_ns(); }
#ifdef ICP_TIMELINE
        if (state_iter == 12) g_timeline_r[511][0] = tlr_now();
#endif
    }
}

int icp_reduce_blocks(int n_src, int n_sms) {
    int nb = (n_src + ICP_REDUCE_THREADS - 1) / ICP_REDUCE_THREADS;
    int per_sm = 2;
    if (const char* e = getenv("ICP_GPU_REDUCE_BLOCKS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v <= 32) per_sm = v; }   // tuning knob
    if (nb > per_sm * n_sms) nb = per_sm * n_sms;
    if (nb < 1) nb = 1;
    return nb;
}

template <int MODE>
static void launch_reduce_mode(const ReduceArgs& a, int n_blocks, cudaStream_t s) {
    if (a.fused) reduce_kernel<MODE, true><<<n_blocks, ICP_REDUCE_THREADS, 0, s>>>(a);
    else reduce_kernel<MODE, false><<<n_blocks, ICP_REDUCE_THREADS, 0, s>>>(a);
}

cudaError_t icp_launch_reduce(const ReduceArgs& a, int n_blocks, cudaStream_t s, int* n_launches) {
    int launches = 0;
    if (a.metric == ICP_GPU_METRIC_P2P) { launch_reduce_mode<0>(a, n_blocks, s); ++launches; }
    else if (a.metric == ICP_GPU_METRIC_P2PLANE) { launch_reduce_mode<1>(a, n_blocks, s); ++launches; }
    else {
        launch_reduce_mode<3>(a, n_blocks, s); ++launches;
        launch_reduce_mode<2>(a, n_blocks, s); ++launches;
    }
    if (n_launches) *n_launches += launches;
    return cudaGetLastError();
}

cudaError_t icp_launch_reduce_phase(const ReduceArgs& a, int n_blocks, int phase, cudaStream_t s, int* n_launches) {
    if (a.metric == ICP_GPU_METRIC_P2P) launch_reduce_mode<0>(a, n_blocks, s);
    else if (a.metric == ICP_GPU_METRIC_P2PLANE) launch_reduce_mode<1>(a, n_blocks, s);
    else if (phase == 0) launch_reduce_mode<3>(a, n_blocks, s);
    else launch_reduce_mode<2>(a, n_blocks, s);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------- pose upload / shard apply
__global__ void pose_init_kernel(DevState* st, const float* pose16, float stop_rot, float stop_trans) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int i = 0; i < 16; ++i) st->pose[i] = pose16[i];
    inv_transpose3_pinned(st->pose, st->nrm);
    st->iter = 0; st->iters_done = 0; st->status = 0; st->ticket = 0; st->ticket2 = 0;
    st->n_queries = 0; st->n_matched = 0; st->n_evals = 0; st->n_nodes = 0;
    for (int k = 0; k < 3; ++k) { st->mean_s[k] = 0.f; st->mean_d[k] = 0.f; }
    st->lm_done = 0; st->lm_iter = 0;
    st->stop_rot = stop_rot; st->stop_trans = stop_trans; st->converged = 0;
}

cudaError_t icp_launch_pose_init(DevState* st, const float* pose_dev16, cudaStream_t s, float stop_rot, float stop_trans) {
    pose_init_kernel<<<1, 32, 0, s>>>(st, pose_dev16, stop_rot, stop_trans);
    return cudaGetLastError();
}

__global__ void shard_apply_kernel(DevState* st, int mode, float* history) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (mode == 3) finish_sums(st, st->shard_partials);
    else finish_row(st, st->shard_partials, mode, history);
}

cudaError_t icp_launch_shard_apply(DevState* st, int mode, float* history, cudaStream_t s) {
    shard_apply_kernel<<<1, 32, 0, s>>>(st, mode, history);
    return cudaGetLastError();
}
