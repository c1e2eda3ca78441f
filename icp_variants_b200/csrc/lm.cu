// lm.cu -- the Levenberg-Marquardt minimiser of CeresICPOptimizer on the device.
//
// Reference: CeresICPOptimizer::estimatePose inner solve (ICPOptimizer.h:283-310), configureSolver
// (:352-360: LEVENBERG_MARQUARDT, monotonic steps, DENSE_QR, max_num_iterations 10, Ceres defaults
// otherwise), prepareConstraints* (:362-482), the functors of constraints.h and PoseIncrement
// (utils.h:25-102).  Ceres itself is an un-vendored dependency; its trust-region loop is restated from
// the published algorithm exactly as oracle/icp_oracle.c:orc_solve_lm does (same order of tests).
//
// One launch = one evaluation of all residual blocks at a parameter vector x: cost, J^T J and J^T r
// (29 doubles) are reduced across the grid and the block that finishes last advances the
// trust-region state machine in DevState (accept / reject, radius update, next candidate, or
// convergence -> pose update).  An outer ICP iteration is 1 + max_num_iterations such launches,
// enqueued back to back; launches that arrive after convergence return immediately.  Jacobians are
// forward-mode dual numbers in fp64 (what Ceres' autodiff computes), with the rotation jets shared by
// all residuals of a thread.
#include "icp_internal.cuh"
#include <float.h>

namespace {

struct Jet { double v; double d[6]; };

__device__ __forceinline__ Jet jconst(double v) { Jet r; r.v = v; for (int i = 0; i < 6; ++i) r.d[i] = 0.0; return r; }
__device__ __forceinline__ Jet jadd(const Jet& a, const Jet& b) { Jet r; r.v = a.v + b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ __forceinline__ Jet jsub(const Jet& a, const Jet& b) { Jet r; r.v = a.v - b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
__device__ __forceinline__ Jet jmul(const Jet& a, const Jet& b) { Jet r; r.v = a.v * b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
__device__ __forceinline__ Jet jscale(const Jet& a, double c) { Jet r; r.v = a.v * c; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * c; return r; }
__device__ __forceinline__ Jet jaddc(const Jet& a, double c) { Jet r = a; r.v += c; return r; }
__device__ __forceinline__ Jet jneg(const Jet& a) { Jet r; r.v = -a.v; for (int i = 0; i < 6; ++i) r.d[i] = -a.d[i]; return r; }
__device__ __forceinline__ Jet jdiv(const Jet& a, const Jet& b) { Jet r; const double inv = 1.0 / b.v; r.v = a.v * inv; for (int i = 0; i < 6; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv; return r; }
__device__ __forceinline__ Jet jsqrt(const Jet& a) { Jet r; r.v = sqrt(a.v); const double k = 1.0 / (2.0 * r.v); for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * k; return r; }
__device__ __forceinline__ Jet jcos(const Jet& a) { Jet r; r.v = cos(a.v); const double k = -sin(a.v); for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * k; return r; }
__device__ __forceinline__ Jet jsin(const Jet& a) { Jet r; r.v = sin(a.v); const double k = cos(a.v); for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * k; return r; }

// ceres::AngleAxisRotatePoint for a fixed angle-axis jet triple, applied to constant points.
struct Rotator {
    bool big;          // theta^2 > DBL_EPSILON: Rodrigues; else first-order p + aa x p
    Jet aa[3], w[3], c, s, omc;
};

__device__ void rotator_init(Rotator& R, const Jet aa[3]) {
    for (int k = 0; k < 3; ++k) R.aa[k] = aa[k];
    const Jet theta2 = jadd(jadd(jmul(aa[0], aa[0]), jmul(aa[1], aa[1])), jmul(aa[2], aa[2]));
    R.big = theta2.v > DBL_EPSILON;
    if (R.big) {
        const Jet theta = jsqrt(theta2);
        R.c = jcos(theta); R.s = jsin(theta);
        const Jet ti = jdiv(jconst(1.0), theta);
        for (int k = 0; k < 3; ++k) R.w[k] = jmul(aa[k], ti);
        R.omc = jsub(jconst(1.0), R.c);
    }
}

__device__ void rotate_const(const Rotator& R, const double p[3], Jet out[3]) {
    if (R.big) {
        const Jet wxp[3] = {jsub(jscale(R.w[1], p[2]), jscale(R.w[2], p[1])), jsub(jscale(R.w[2], p[0]), jscale(R.w[0], p[2])),
                            jsub(jscale(R.w[0], p[1]), jscale(R.w[1], p[0]))};
        const Jet tmp = jmul(jadd(jadd(jscale(R.w[0], p[0]), jscale(R.w[1], p[1])), jscale(R.w[2], p[2])), R.omc);
        for (int i = 0; i < 3; ++i) out[i] = jadd(jadd(jscale(R.c, p[i]), jmul(wxp[i], R.s)), jmul(R.w[i], tmp));
    } else {
        const Jet wxp[3] = {jsub(jscale(R.aa[1], p[2]), jscale(R.aa[2], p[1])), jsub(jscale(R.aa[2], p[0]), jscale(R.aa[0], p[2])),
                            jsub(jscale(R.aa[0], p[1]), jscale(R.aa[1], p[0]))};
        for (int i = 0; i < 3; ++i) out[i] = jaddc(wxp[i], p[i]);
    }
}

// Row layout: [0..20] J^T J upper triangle row-major, [21..26] J^T r, [27] sum r^2, [28] residual count.
__device__ __forceinline__ void accumulate(double (&v)[32], const Jet& r) {
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b) v[k++] += r.d[a] * r.d[b];
#pragma unroll
    for (int a = 0; a < 6; ++a) v[21 + a] += r.d[a] * r.v;
    v[27] += r.v * r.v;
    v[28] += 1.0;
}

// PoseIncrement<double>::convertToMatrix (utils.h:79-98) via ceres::AngleAxisToRotationMatrix
__device__ void increment_matrix(const double* x, float* inc) {
    double R[9];
    const double theta2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    if (theta2 > DBL_EPSILON) {
        const double theta = sqrt(theta2);
        const double wx = x[0] / theta, wy = x[1] / theta, wz = x[2] / theta;
        const double ct = cos(theta), st = sin(theta);
        R[0] = ct + wx * wx * (1.0 - ct);      R[1] = wz * st + wx * wy * (1.0 - ct);  R[2] = -wy * st + wx * wz * (1.0 - ct);
        R[3] = wx * wy * (1.0 - ct) - wz * st; R[4] = ct + wy * wy * (1.0 - ct);       R[5] = wx * st + wy * wz * (1.0 - ct);
        R[6] = wy * st + wx * wz * (1.0 - ct); R[7] = -wx * st + wy * wz * (1.0 - ct); R[8] = ct + wz * wz * (1.0 - ct);
    } else {
        R[0] = 1; R[1] = x[2]; R[2] = -x[1]; R[3] = -x[2]; R[4] = 1; R[5] = x[0]; R[6] = x[1]; R[7] = -x[0]; R[8] = 1;
    }
    mat4_identity_dev(inc);
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) inc[r + 4 * c] = (float)R[r + 3 * c];   // column-major both sides
    inc[12] = (float)x[3]; inc[13] = (float)x[4]; inc[14] = (float)x[5];
}

__device__ void lm_finalize(DevState* st, int rc, float* history) {
    float inc[16];
    increment_matrix(st->lm_x, inc);
    apply_increment(st, inc, rc, history);
    st->lm_done = 1;
}

__device__ void unpack_system(const double* row, double* H, double* g, double* cost) {
    int k = 0;
    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { H[i * 6 + j] = row[k]; H[j * 6 + i] = row[k]; ++k; }
    for (int i = 0; i < 6; ++i) g[i] = row[21 + i];
    *cost = 0.5 * row[27];
}

// Top of TrustRegionMinimizer's loop: termination tests, then LevenbergMarquardtStrategy::ComputeStep
// in the Jacobi-scaled space; loops over invalid steps (they need no evaluation).  Leaves the next
// candidate in lm_cand, or finalises.
__device__ void lm_next_step(DevState* st, int max_iterations, float* history) {
    const double max_radius = 1e16, min_radius = 1e-32, min_diag = 1e-6, max_diag = 1e32, gradient_tolerance = 1e-10;
    (void)max_radius;
    for (;;) {
        double gmax = 0.0;
        for (int i = 0; i < 6; ++i) gmax = fmax(gmax, fabs(st->lm_g[i]));
        if (st->lm_iter >= max_iterations || (st->lm_step_ok && gmax <= gradient_tolerance) || st->lm_radius <= min_radius) {
            lm_finalize(st, 0, history);
            return;
        }
        st->lm_iter += 1;
        double Hs[36], gs[6], A[36], b[6], ds[6];
        for (int a = 0; a < 6; ++a) {
            gs[a] = st->lm_g[a] * st->lm_scale[a];
            for (int c = 0; c < 6; ++c) Hs[a * 6 + c] = st->lm_H[a * 6 + c] * st->lm_scale[a] * st->lm_scale[c];
        }
        if (!st->lm_reuse_diag)
            for (int a = 0; a < 6; ++a) { const double v = Hs[a * 6 + a]; st->lm_diag[a] = v < min_diag ? min_diag : (v > max_diag ? max_diag : v); }
        for (int i = 0; i < 36; ++i) A[i] = Hs[i];
        for (int a = 0; a < 6; ++a) { A[a * 6 + a] += st->lm_diag[a] / st->lm_radius; b[a] = -gs[a]; }
        const bool lin_ok = solve6_dev(A, b, ds) == 0;
        double model_cost_change = 0.0;
        if (lin_ok)
            for (int a = 0; a < 6; ++a) {
                double hd = 0.0;
                for (int c = 0; c < 6; ++c) hd += Hs[a * 6 + c] * ds[c];
                model_cost_change -= ds[a] * (gs[a] + 0.5 * hd);
            }
        if (!lin_ok || !(model_cost_change > 0.0)) {            // HandleInvalidStep
            st->lm_invalid += 1;
            if (st->lm_invalid >= 5) { lm_finalize(st, 0, history); return; }
            st->lm_radius *= 0.5; st->lm_reuse_diag = 1; st->lm_step_ok = 0;
            continue;
        }
        st->lm_invalid = 0;
        st->lm_model_change = model_cost_change;
        for (int a = 0; a < 6; ++a) st->lm_cand[a] = st->lm_x[a] + ds[a] * st->lm_scale[a];
        st->lm_have_cand = 1;
        return;
    }
}

__device__ void lm_advance(DevState* st, const double* row, int step_index, int max_iterations, float* history) {
    const double min_relative_decrease = 1e-3, function_tolerance = 1e-6, parameter_tolerance = 1e-8, max_radius = 1e16;
    if (step_index == 0) {
        // initial evaluation at x = 0 (poseIncrement.setZero(), ICPOptimizer.h:236,310)
        for (int i = 0; i < 6; ++i) st->lm_x[i] = 0.0;
        st->lm_iter = 0; st->lm_done = 0; st->lm_reuse_diag = 0; st->lm_invalid = 0; st->lm_step_ok = 1; st->lm_have_cand = 0;
        st->lm_radius = 1e4; st->lm_decrease = 2.0;
        if (!(row[28] > 0.0)) { lm_finalize(st, ICP_GPU_E_NO_MATCHES, history); return; }
        unpack_system(row, st->lm_H, st->lm_g, &st->lm_cost);
        for (int i = 0; i < 6; ++i) st->lm_scale[i] = 1.0 / (1.0 + sqrt(st->lm_H[i * 6 + i]));   // jacobi_scaling
        lm_next_step(st, max_iterations, history);
        return;
    }
    // evaluation at the candidate
    double Hc[36], gc[6], cand_cost;
    unpack_system(row, Hc, gc, &cand_cost);
    double step_norm = 0.0, x_norm = 0.0;
    for (int a = 0; a < 6; ++a) { const double d = st->lm_x[a] - st->lm_cand[a]; step_norm += d * d; x_norm += st->lm_x[a] * st->lm_x[a]; }
    step_norm = sqrt(step_norm); x_norm = sqrt(x_norm);
    if (step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) { lm_finalize(st, 0, history); return; }   // ParameterToleranceReached
    const double cost_change = st->lm_cost - cand_cost;
    if (fabs(cost_change) <= function_tolerance * st->lm_cost) { lm_finalize(st, 0, history); return; }              // FunctionToleranceReached
    const double rho = cost_change / st->lm_model_change;
    if (rho > min_relative_decrease) {                                                                               // HandleSuccessfulStep
        for (int a = 0; a < 6; ++a) { st->lm_x[a] = st->lm_cand[a]; st->lm_g[a] = gc[a]; }
        for (int i = 0; i < 36; ++i) st->lm_H[i] = Hc[i];
        st->lm_cost = cand_cost;
        const double t = 2.0 * rho - 1.0;
        double denom = 1.0 - t * t * t; if (denom < 1.0 / 3.0) denom = 1.0 / 3.0;
        st->lm_radius = st->lm_radius / denom; if (st->lm_radius > max_radius) st->lm_radius = max_radius;
        st->lm_decrease = 2.0; st->lm_reuse_diag = 0; st->lm_step_ok = 1;
    } else {                                                                                                         // HandleUnsuccessfulStep
        st->lm_radius = st->lm_radius / st->lm_decrease; st->lm_decrease *= 2.0; st->lm_reuse_diag = 1; st->lm_step_ok = 0;
    }
    lm_next_step(st, max_iterations, history);
}

template <int METRIC>
__global__ void __launch_bounds__(ICP_REDUCE_THREADS) lm_eval_kernel(const ReduceArgs a, int step_index, int max_iterations) {
    __shared__ float P[16];
    __shared__ float Nm[9];
    __shared__ double xs[6];
    __shared__ double red[ICP_REDUCE_THREADS / 32][32];
    __shared__ double fin[ICP_REDUCE_THREADS / 32][32];
    __shared__ bool is_last;
    if (a.state->converged) return;                          // early stop reached in an earlier outer iteration
    if (step_index > 0 && a.state->lm_done) return;          // converged earlier in this outer iteration (uniform across the grid)
    if (threadIdx.x < 16) P[threadIdx.x] = a.state->pose[threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 41) Nm[threadIdx.x - 32] = a.state->nrm[threadIdx.x - 32];
    if (threadIdx.x >= 64 && threadIdx.x < 70) xs[threadIdx.x - 64] = step_index == 0 ? 0.0 : a.state->lm_cand[threadIdx.x - 64];
    __syncthreads();
    double v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = 0.0;
    Jet pose[6];
    for (int i = 0; i < 6; ++i) { pose[i] = jconst(xs[i]); pose[i].d[i] = 1.0; }
    Rotator rot; rotator_init(rot, pose);                     // PoseIncrement::apply, utils.h:44-56
    Rotator rinv;
    if (METRIC == ICP_GPU_METRIC_SYMMETRIC) { const Jet ninv[3] = {jneg(pose[0]), jneg(pose[1]), jneg(pose[2])}; rotator_init(rinv, ninv); }   // utils.h:60-72
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_src; i += gridDim.x * blockDim.x) {
        const int pos = a.match_pos[i];
        if (pos < 0) continue;
        const float4 sp = __ldg(&a.src_pts[i]);
        float sxf, syf, szf;
        xform_point(P, sp.x, sp.y, sp.z, sxf, syf, szf);
        const float4 tp = __ldg(&a.tgt_pts[pos]);
        if (!finite3(sxf, syf, szf) || !finite3(tp.x, tp.y, tp.z)) continue;        // ICPOptimizer.h:378
        const double w = (double)a.match_w[i];
        const double s[3] = {sxf, syf, szf}, t[3] = {tp.x, tp.y, tp.z};
        Jet y[3];
        rotate_const(rot, s, y);
        for (int k = 0; k < 3; ++k) y[k] = jadd(y[k], pose[3 + k]);
        const double lw_pt = (double)0.1f * w;                                        // PointToPointConstraint LAMBDA (constraints.h)
        for (int k = 0; k < 3; ++k) accumulate(v, jscale(jaddc(y[k], -t[k]), lw_pt));
        if (METRIC == ICP_GPU_METRIC_P2PLANE) {
            const float4 tn = __ldg(&a.tgt_nrm[pos]);
            if (finite3(tn.x, tn.y, tn.z)) {                                          // ICPOptimizer.h:420-423
                const Jet acc = jadd(jadd(jscale(jaddc(y[0], -t[0]), (double)tn.x), jscale(jaddc(y[1], -t[1]), (double)tn.y)),
                                     jscale(jaddc(y[2], -t[2]), (double)tn.z));
                accumulate(v, jscale(acc, (double)1.0f * w));
            }
        } else if (METRIC == ICP_GPU_METRIC_SYMMETRIC) {
            const float4 tn = __ldg(&a.tgt_nrm[pos]);
            const float4 sn4 = __ldg(&a.src_nrm[i]);
            float nx, ny, nz;
            xform_normal(Nm, sn4.x, sn4.y, sn4.z, nx, ny, nz);
            if (finite3(tn.x, tn.y, tn.z) && finite3(nx, ny, nz)) {                   // ICPOptimizer.h:465-469
                Jet z[3];
                rotate_const(rinv, t, z);
                const double m[3] = {(double)tn.x + (double)nx, (double)tn.y + (double)ny, (double)tn.z + (double)nz};
                const Jet acc = jadd(jadd(jscale(jsub(y[0], z[0]), m[0]), jscale(jsub(y[1], z[1]), m[1])), jscale(jsub(y[2], z[2]), m[2]));
                accumulate(v, jscale(acc, (double)1.0f * w));
            }
        }
    }
    if (!grid_reduce_row<ICP_REDUCE_THREADS>(v, a.partials, &a.state->ticket, red, fin, &is_last, &a.peer, a.state)) return;
    if (threadIdx.x == 0) lm_advance(a.state, fin[0], step_index, max_iterations, a.pose_history);
}

}  // namespace

cudaError_t icp_launch_lm(const ReduceArgs& a, int n_blocks, int lm_max_iterations, cudaStream_t s, int* n_launches) {
    for (int step = 0; step <= lm_max_iterations; ++step) {
        if (a.metric == ICP_GPU_METRIC_P2P) lm_eval_kernel<0><<<n_blocks, ICP_REDUCE_THREADS, 0, s>>>(a, step, lm_max_iterations);
        else if (a.metric == ICP_GPU_METRIC_P2PLANE) lm_eval_kernel<1><<<n_blocks, ICP_REDUCE_THREADS, 0, s>>>(a, step, lm_max_iterations);
        else lm_eval_kernel<2><<<n_blocks, ICP_REDUCE_THREADS, 0, s>>>(a, step, lm_max_iterations);
        if (n_launches) *n_launches += 1;
    }
    return cudaGetLastError();
}
