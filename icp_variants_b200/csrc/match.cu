// match.cu -- stages 1-4 of one ICP iteration fused into one kernel per matching method:
//   selection slot -> source index            (selection.h:28-60, PointCloud.h:325-343)
//   transformPoints / transformNormals        (utils.h:106-133)
//   queryMatches: exact 1-NN (3-D / 6-D) or projective window search
//                                             (NearestNeighbor.h:143-207, :234-303, :333-421)
//   WeightingMethod::applyWeights             (weighting.h:39-99)
//   ICPOptimizer::pruneCorrespondences        (ICPOptimizer.h:157-174)
// The reference's FLANN search (1 randomized kd-tree, 16 checks) is approximate; these kernels return
// the exact nearest neighbour under contract D1-D3 with ties to the lowest target index, i.e. what
// NearestNeighborSearchBruteForce's scan order (NearestNeighbor.h:81-97) yields on squared distances.
#include "icp_internal.cuh"
#include <limits.h>

#define MINF_F (-INFINITY)

__device__ __forceinline__ int sel3i(int a, int x, int y, int z) { return a == 0 ? x : (a == 1 ? y : z); }
__device__ __forceinline__ float sel3f(int a, float x, float y, float z) { return a == 0 ? x : (a == 1 ? y : z); }

struct Query {
    float x, y, z;        // transformed source point
    float cr, cg, cb;     // colour features (6-D search only)
};

struct Best { float d; int idx; int pos; };

// (d, idx) lexicographic '<' : contract D2
__device__ __forceinline__ bool better(float d, int idx, const Best& b) { return d < b.d || (d == b.d && idx < b.idx); }

__device__ __forceinline__ float color_feature(unsigned int rgba, int k) {
    // NearestNeighbor.h:212-221,245-254: color_scale(1) * color_normalize(1/float(255)) * uchar
    return pmul(1.0f / 255.0f, (float)((rgba >> (8 * k)) & 0xFFu));
}

template <bool COLOR>
__device__ __forceinline__ float dist2(const Query& q, const float4 c, float best, const float4* __restrict__ nrm, unsigned int i) {
    const float dx = psub(q.x, c.x), dy = psub(q.y, c.y), dz = psub(q.z, c.z);
    float d = padd(padd(pmul(dx, dx), pmul(dy, dy)), pmul(dz, dz));            // D1
    if (COLOR) {
        if (d > best) return d;                                                // 3-D part already loses; adding squares cannot help
        const unsigned int rgba = __float_as_uint(__ldg(&nrm[i].w));
        const float dr = psub(q.cr, color_feature(rgba, 0)), dg = psub(q.cg, color_feature(rgba, 1)), db = psub(q.cb, color_feature(rgba, 2));
        d = padd(d, pmul(dr, dr)); d = padd(d, pmul(dg, dg)); d = padd(d, pmul(db, db));   // FLANN L2 order for 6 dims
    }
    return d;
}

template <bool COLOR>
__device__ __forceinline__ void scan_range(const float4* __restrict__ pts, const float4* __restrict__ nrm, unsigned int s, unsigned int e,
                                           const Query& q, Best& b, unsigned int& evals) {
    for (unsigned int i = s; i < e; ++i) {
        const float4 c = __ldg(&pts[i]);
        const float d = dist2<COLOR>(q, c, b.d, nrm, i);
        const int idx = __float_as_int(c.w);
        if (better(d, idx, b)) { b.d = d; b.idx = idx; b.pos = (int)i; }
    }
    evals += e - s;
}

// Squared gap between q and the slab of cells [lo, lo + 2^r) on axis a, a lower bound (under fp32
// rounding, by monotonicity) of the per-axis term of D1 for every point stored in those cells.
__device__ __forceinline__ float gap2(const GridParams& g, int a, float q, int lo, int r) {
    const float o = sel3f(a, g.o[0], g.o[1], g.o[2]), h = sel3f(a, g.h[0], g.h[1], g.h[2]), dl = sel3f(a, g.delta[0], g.delta[1], g.delta[2]);
    const float L = (o + (float)lo * h) - dl;
    const float U = (o + (float)(lo + (1 << r)) * h) + dl;
    const float t = q < L ? psub(L, q) : (q > U ? psub(q, U) : 0.0f);
    return pmul(t, t);
}

// Exact nearest neighbour by depth-first descent of the implicit tree (grid.cu): near child first,
// far child only if its box bound does not exceed the best distance so far ('>' keeps equal bounds
// alive so that an equally distant point with a lower index is still found).
template <bool COLOR>
__device__ void grid_search(const GridParams& g, const unsigned int* __restrict__ cs, const float4* __restrict__ pts,
                            const float4* __restrict__ nrm, const Query& q, Best& b, unsigned int& evals, unsigned int& nodes) {
    const int T = g.T;
    const unsigned int s0 = __ldg(&cs[0]), e0 = __ldg(&cs[1u << T]);
    if (e0 == s0) return;
    int lo0 = 0, lo1 = 0, lo2 = 0, r0 = g.bits[0], r1 = g.bits[1], r2 = g.bits[2];
    float t0 = gap2(g, 0, q.x, 0, r0), t1 = gap2(g, 1, q.y, 0, r1), t2 = gap2(g, 2, q.z, 0, r2);
    ++nodes;
    if (padd(padd(t0, t1), t2) > b.d) return;
    if (e0 - s0 <= ICP_LEAF_MAX || T == 0) { scan_range<COLOR>(pts, nrm, s0, e0, q, b, evals); return; }
    int d = 0, stage = 0, from = 0; unsigned int p = 0; bool asc = false;
    for (;;) {
        const int a = (int)((g.axis_seq >> (2 * d)) & 3ull);
        const int lo = sel3i(a, lo0, lo1, lo2), r = sel3i(a, r0, r1, r2) - 1;
        const float qa = sel3f(a, q.x, q.y, q.z);
        const float plane = sel3f(a, g.o[0], g.o[1], g.o[2]) + (float)(lo + (1 << r)) * sel3f(a, g.h[0], g.h[1], g.h[2]);
        const int nb = qa >= plane ? 1 : 0;                 // visiting order only; any choice is correct
        if (asc) { stage = (from == nb) ? 1 : 2; asc = false; }
        if (stage == 2) {                                   // both children done: ascend
            if (d == 0) break;
            from = (int)(p & 1u); --d; p >>= 1;
            const int a2 = (int)((g.axis_seq >> (2 * d)) & 3ull);
            const int rc = sel3i(a2, r0, r1, r2), lc = sel3i(a2, lo0, lo1, lo2);
            const int lp = lc - (from ? (1 << rc) : 0), rp = rc + 1;
            const float tp = gap2(g, a2, sel3f(a2, q.x, q.y, q.z), lp, rp);
            if (a2 == 0) { lo0 = lp; r0 = rp; t0 = tp; } else if (a2 == 1) { lo1 = lp; r1 = rp; t1 = tp; } else { lo2 = lp; r2 = rp; t2 = tp; }
            asc = true;
            continue;
        }
        const int bch = stage == 0 ? nb : (nb ^ 1);
        ++stage;
        const int lc = lo + (bch ? (1 << r) : 0);
        const float tc = gap2(g, a, qa, lc, r);
        const float lb = padd(padd(a == 0 ? tc : t0, a == 1 ? tc : t1), a == 2 ? tc : t2);
        ++nodes;
        if (lb > b.d) continue;
        const unsigned int c = 2u * p + (unsigned int)bch; const int sh = T - d - 1;
        const unsigned int ns = __ldg(&cs[c << sh]), ne = __ldg(&cs[(c + 1u) << sh]);
        if (ne == ns) continue;
        if (ne - ns <= ICP_LEAF_MAX || d + 1 == T) { scan_range<COLOR>(pts, nrm, ns, ne, q, b, evals); continue; }
        if (a == 0) { lo0 = lc; r0 = r; t0 = tc; } else if (a == 1) { lo1 = lc; r1 = r; t1 = tc; } else { lo2 = lc; r2 = r; t2 = tc; }
        ++d; p = c; stage = 0;
    }
}

// x86-64 gcc semantics of `unsigned = std::round(float)` (NearestNeighbor.h:378-379): cvttss2si to
// 64 bits, low 32 bits kept; NaN / out of range -> 0.
__device__ __forceinline__ unsigned int x86_float_to_u32(float t) {
    if (!(fabsf(t) < 9223372036854775808.0f)) return 0u;
    return (unsigned int)(long long)t;
}

// Shared per-block copy of the pose and the normal matrix.
struct PoseSm { float P[16]; float N[9]; };

__device__ __forceinline__ void load_pose(PoseSm& sm, const DevState* st) {
    if (threadIdx.x < 16) sm.P[threadIdx.x] = st->pose[threadIdx.x];
    else if (threadIdx.x < 25) sm.N[threadIdx.x - 16] = st->nrm[threadIdx.x - 16];
    __syncthreads();
}

__device__ __forceinline__ float unit_hash(unsigned int key, unsigned int k) {
    // counter-based stream for ICP_GPU_RNG_DEVICE (two rounds of a 32-bit mix)
    unsigned int x = k * 0x9E3779B9u + key;
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    x += key * 0x85EBCA6Bu; x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return (float)(x >> 8) * (1.0f / 16777216.0f);
}

// Stages 3-4 for one query given its match; writes the slot's outputs.
__device__ __forceinline__ void finish_match(const MatchArgs& a, int slot, bool matched, float w_const, int t_idx, int t_pos,
                                             float sx, float sy, float sz, float snx, float sny, float snz, unsigned int s_rgba,
                                             unsigned int& n_matched) {
    int out_idx = -1, out_pos = -1; float w = 0.0f;
    if (matched) {
        out_idx = t_idx; out_pos = t_pos; w = w_const;
        const float4 tn = __ldg(&a.tgt_nrm[t_pos]);
        if (a.weighting != ICP_GPU_WEIGHT_CONSTANT) {                         // weighting.h:44 early return
            const float4 tp = __ldg(&a.tgt_pts[t_pos]);
            w = 0.0f;
            if (a.weighting == ICP_GPU_WEIGHT_DISTANCES || a.weighting == ICP_GPU_WEIGHT_COLORS) {
                if (finite3(sx, sy, sz) && finite3(tp.x, tp.y, tp.z)) {        // weighting.h:58-59
                    const float d0 = psub(sx, tp.x), d1 = psub(sy, tp.y), d2 = psub(sz, tp.z);
                    const float q = pdiv(padd(padd(pmul(d0, d0), pmul(d1, d1)), pmul(d2, d2)), a.max_d2);
                    w = (float)(1.0 - (double)q);                              // weighting.h:16-20
                }
            }
            if (a.weighting == ICP_GPU_WEIGHT_NORMALS) {
                if (finite3(snx, sny, snz) && finite3(tn.x, tn.y, tn.z))       // weighting.h:72-73
                    w = padd(padd(pmul(snx, tn.x), pmul(sny, tn.y)), pmul(snz, tn.z));   // weighting.h:22-25 (unclamped)
            }
            if (a.weighting == ICP_GPU_WEIGHT_COLORS) {
                // weighting.h:27-30: Vector4uc difference wraps modulo 256 before squaring
                const unsigned int t_rgba = __float_as_uint(tn.w);
                int s = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int e = (int)((((s_rgba >> (8 * k)) & 0xFFu) - ((t_rgba >> (8 * k)) & 0xFFu)) & 0xFFu);
                    s += e * e;
                }
                const float cw = (float)(1.0 - (double)pdiv((float)s, 195075.0f));
                w = pmul(w, cw);
            }
        }
        if (a.rejection == 1) {                                                // ICPOptimizer.h:157-174
            const float dot = padd(padd(pmul(snx, tn.x), pmul(sny, tn.y)), pmul(snz, tn.z));
            const float na = __fsqrt_rn(padd(padd(pmul(snx, snx), pmul(sny, sny)), pmul(snz, snz)));
            const float nb = __fsqrt_rn(padd(padd(pmul(tn.x, tn.x), pmul(tn.y, tn.y)), pmul(tn.z, tn.z)));
            const float c = pdiv(dot, pmul(na, nb));
            if (c <= 0.5f && c >= -1.0f) { out_idx = -1; out_pos = -1; }       // D6; NaN and |c|>1 are kept like acos()'s NaN
        }
        if (out_pos >= 0) ++n_matched;
    }
    a.match_pos[slot] = out_pos;
    a.match_w[slot] = w;
    if (a.match_idx) a.match_idx[slot] = out_idx;
}

__device__ __forceinline__ void flush_stats(DevState* st, unsigned int nq, unsigned int nm, unsigned int ev, unsigned int nd) {
    nq = __reduce_add_sync(0xFFFFFFFFu, nq); nm = __reduce_add_sync(0xFFFFFFFFu, nm);
    // evals / nodes can exceed 32 bits only per launch, not per warp
    ev = __reduce_add_sync(0xFFFFFFFFu, ev); nd = __reduce_add_sync(0xFFFFFFFFu, nd);
    if ((threadIdx.x & 31) == 0) {
        if (nq) atomicAdd(&st->n_queries, (unsigned long long)nq);
        if (nm) atomicAdd(&st->n_matched, (unsigned long long)nm);
        if (ev) atomicAdd(&st->n_evals, (unsigned long long)ev);
        if (nd) atomicAdd(&st->n_nodes, (unsigned long long)nd);
    }
}

// Resolves the slot to a source point and transforms it. Returns false when the slot is not a query.
__device__ __forceinline__ bool prepare_query(const MatchArgs& a, const IterDesc& d, const PoseSm& sm, int slot, Query& q,
                                              float& snx, float& sny, float& snz, unsigned int& s_rgba) {
    const int i = slot_source_index(d, a.sel, slot, a.n_src);
    if (i < 0) return false;
    const float4 p = __ldg(&a.src_pts[i]);
    const float4 n = __ldg(&a.src_nrm[i]);
    if (d.filter_finite && !(finite3(p.x, p.y, p.z) && finite3(n.x, n.y, n.z))) return false;   // PointCloud.h:335
    if (d.proba >= 0.0f && !(unit_hash(d.rng_key, (unsigned int)slot) < d.proba)) return false;   // device selection stream
    xform_point(sm.P, p.x, p.y, p.z, q.x, q.y, q.z);
    xform_normal(sm.N, n.x, n.y, n.z, snx, sny, snz);
    s_rgba = __float_as_uint(p.w);
    q.cr = color_feature(s_rgba, 0); q.cg = color_feature(s_rgba, 1); q.cb = color_feature(s_rgba, 2);
    return true;
}

template <bool COLOR>
__global__ void __launch_bounds__(ICP_MATCH_THREADS) knn_grid_kernel(const MatchArgs a) {
    __shared__ PoseSm sm;
    load_pose(sm, a.state_ro);
    const IterDesc d = a.desc[a.desc_index >= 0 ? a.desc_index : a.state_ro->iter];
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int nq = 0, nm = 0, ev = 0, nd = 0;
    if (slot < d.n_queries) {
        Query q; float snx, sny, snz; unsigned int s_rgba;
        if (prepare_query(a, d, sm, slot, q, snx, sny, snz, s_rgba)) {
            ++nq;
            Best b; b.d = fminf(a.max_d2, 3.4028234e38f); b.idx = INT_MAX; b.pos = -1;
            if (finite3(q.x, q.y, q.z)) {
                const GridParams g = *a.grid;
                grid_search<COLOR>(g, a.cell_start, a.tgt_pts, a.tgt_nrm, q, b, ev, nd);
            }
            finish_match(a, slot, b.pos >= 0, 1.0f, b.idx, b.pos, q.x, q.y, q.z, snx, sny, snz, s_rgba, nm);
        } else {
            a.match_pos[slot] = -1; a.match_w[slot] = 0.0f; if (a.match_idx) a.match_idx[slot] = -1;
        }
    }
    flush_stats(a.state, nq, nm, ev, nd);
}

// Small targets: one warp per query, lanes stride over the target (original order, L1-resident),
// warp-shuffle arg-min on (d, idx).
template <bool COLOR>
__global__ void __launch_bounds__(ICP_MATCH_THREADS) knn_brute_kernel(const MatchArgs a) {
    __shared__ PoseSm sm;
    load_pose(sm, a.state_ro);
    const IterDesc d = a.desc[a.desc_index >= 0 ? a.desc_index : a.state_ro->iter];
    const int lane = threadIdx.x & 31;
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned int nq = 0, nm = 0, ev = 0, nd = 0;
    if (slot < d.n_queries) {
        Query q; float snx, sny, snz; unsigned int s_rgba;
        if (prepare_query(a, d, sm, slot, q, snx, sny, snz, s_rgba)) {     // warp-uniform
            Best b; b.d = fminf(a.max_d2, 3.4028234e38f); b.idx = INT_MAX; b.pos = -1;
            if (finite3(q.x, q.y, q.z)) {
                for (int j = lane; j < a.n_tgt; j += 32) {
                    const float4 c = __ldg(&a.tgt_pts[j]);
                    const float dd = dist2<COLOR>(q, c, b.d, a.tgt_nrm, (unsigned int)j);
                    if (better(dd, j, b)) { b.d = dd; b.idx = j; b.pos = j; }
                }
                ev += (unsigned int)((a.n_tgt - lane + 31) / 32);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    Best t; t.d = __shfl_xor_sync(0xFFFFFFFFu, b.d, o); t.idx = __shfl_xor_sync(0xFFFFFFFFu, b.idx, o); t.pos = t.idx;
                    if (better(t.d, t.idx, b)) b = t;
                }
                if (b.idx == INT_MAX) b.pos = -1; else b.pos = b.idx;
            }
            if (lane == 0) { ++nq; finish_match(a, slot, b.pos >= 0, 1.0f, b.idx, b.pos, q.x, q.y, q.z, snx, sny, snz, s_rgba, nm); }
        } else if (lane == 0) {
            a.match_pos[slot] = -1; a.match_w[slot] = 0.0f; if (a.match_idx) a.match_idx[slot] = -1;
        }
    }
    flush_stats(a.state, nq, nm, ev, nd);
}

// Projective matching (NearestNeighbor.h:333-421): literal restatement of the window scan, unsigned
// wrap-around included, one thread per query; neighbouring threads read neighbouring target pixels.
__global__ void __launch_bounds__(ICP_MATCH_THREADS) projective_kernel(const MatchArgs a) {
    __shared__ PoseSm sm;
    load_pose(sm, a.state_ro);
    const IterDesc d = a.desc[a.desc_index >= 0 ? a.desc_index : a.state_ro->iter];
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int nq = 0, nm = 0, ev = 0, nd = 0;
    if (slot < d.n_queries) {
        Query q; float snx, sny, snz; unsigned int s_rgba;
        if (prepare_query(a, d, sm, slot, q, snx, sny, snz, s_rgba)) {
            ++nq;
            if (q.x == MINF_F) {
                // :372-373 `continue` leaves the value-initialised Match{0, 0.f} (:353)
                finish_match(a, slot, true, 0.0f, 0, 0, q.x, q.y, q.z, snx, sny, snz, s_rgba, nm);
            } else {
                const unsigned int searchWindow = 12u;                         // NearestNeighbor.h:319
                const unsigned int uP = x86_float_to_u32(roundf(padd(pdiv(pmul(q.x, a.fx), q.z), a.cx)));
                const unsigned int vP = x86_float_to_u32(roundf(padd(pdiv(pmul(q.y, a.fy), q.z), a.cy)));
                float minDist = 3.4028234e38f; unsigned int idx = 0xFFFFFFFFu;
                for (unsigned int v = vP - searchWindow; (v < a.height && v <= vP + searchWindow); v++) {
                    for (unsigned int u = uP - searchWindow; (u < a.width && u <= uP + searchWindow); u++) {
                        const unsigned int ni = a.width * v + u;
                        const float4 t = __ldg(&a.tgt_pts[ni]);
                        if (t.x == MINF_F) continue;
                        const float dx = psub(q.x, t.x), dy = psub(q.y, t.y), dz = psub(q.z, t.z);
                        const float dist = padd(padd(pmul(dx, dx), pmul(dy, dy)), pmul(dz, dz));
                        ++ev;
                        if (minDist > dist) { idx = ni; minDist = dist; }
                    }
                }
                const bool ok = minDist <= a.max_d2 && idx != 0xFFFFFFFFu;
                finish_match(a, slot, ok, 1.0f, (int)idx, (int)idx, q.x, q.y, q.z, snx, sny, snz, s_rgba, nm);
            }
        } else {
            a.match_pos[slot] = -1; a.match_w[slot] = 0.0f; if (a.match_idx) a.match_idx[slot] = -1;
        }
    }
    flush_stats(a.state, nq, nm, ev, nd);
}

cudaError_t icp_launch_match(const MatchArgs& a, int algorithm, int max_queries, cudaStream_t s) {
    if (max_queries <= 0) return cudaSuccess;
    const int T = ICP_MATCH_THREADS;
    if (algorithm == 2) {
        projective_kernel<<<(max_queries + T - 1) / T, T, 0, s>>>(a);
    } else if (algorithm == 1) {
        const long long threads = (long long)max_queries * 32;
        const int nb = (int)((threads + T - 1) / T);
        if (a.color_icp) knn_brute_kernel<true><<<nb, T, 0, s>>>(a); else knn_brute_kernel<false><<<nb, T, 0, s>>>(a);
    } else {
        const int nb = (max_queries + T - 1) / T;
        if (a.color_icp) knn_grid_kernel<true><<<nb, T, 0, s>>>(a); else knn_grid_kernel<false><<<nb, T, 0, s>>>(a);
    }
    return cudaGetLastError();
}
