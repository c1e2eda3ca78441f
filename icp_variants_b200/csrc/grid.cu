// grid.cu -- buildIndex: the on-device search structure that replaces FLANN's kd-tree
// (reference: NearestNeighborSearchFlann::buildIndex, NearestNeighbor.h:122-141 / :209-232).
//
// Structure: a uniform grid of 2^T cells whose cell codes are Morton-style bit interleavings of
// the per-axis cell indices (the axis taken at each bit is chosen so that cells end up near-cubic).
// Points are counting-sorted by cell code (one radix pass, radix 2^T), and the exclusive prefix sum
// of the per-cell counts, cell_start[0..2^T], doubles as an implicit binary tree: the node at depth d
// with code prefix p owns points [cell_start[p << (T-d)], cell_start[(p+1) << (T-d)]).  No node
// storage, no pointers; empty subtrees are recognised by an empty range.
//
// Launches (all on one stream, no host synchronisation): pack -> bbox -> params -> memset ->
// keys+count -> scan (3) -> scatter.
#include "icp_internal.cuh"
#include <string.h>
#include <stdlib.h>

// ---------------------------------------------------------------------------- pack AoS3 -> float4
__global__ void pack_cloud_kernel(const float* __restrict__ xyz, const float* __restrict__ nrm, const uint8_t* __restrict__ rgba,
                                  int n, float4* __restrict__ pts, float4* __restrict__ nrmo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // Layout in HBM: points {x,y,z,original index bits}, normals {nx,ny,nz,rgba bits}.
    unsigned int c = 0;
    if (rgba) c = reinterpret_cast<const unsigned int*>(rgba)[i];
    float4 p; p.x = xyz[3 * (size_t)i]; p.y = xyz[3 * (size_t)i + 1]; p.z = xyz[3 * (size_t)i + 2]; p.w = __int_as_float(i);
    float4 m = make_float4(0.f, 0.f, 0.f, __uint_as_float(c));
    if (nrm) { m.x = nrm[3 * (size_t)i]; m.y = nrm[3 * (size_t)i + 1]; m.z = nrm[3 * (size_t)i + 2]; }
    pts[i] = p; nrmo[i] = m;
}

cudaError_t icp_launch_pack_cloud(const float* xyz, const float* nrm, const uint8_t* rgba, int n, float4* pts, float4* nrmo,
                                  cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    pack_cloud_kernel<<<(n + 255) / 256, 256, 0, s>>>(xyz, nrm, rgba, n, pts, nrmo);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------- bounding box
__device__ __forceinline__ unsigned int enc_f(float f) { unsigned int b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float dec_f(unsigned int e) { return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e); }

// bbox[0..2] = encoded min, bbox[3..5] = encoded max (initialised to 0xFFFFFFFF / 0)
__global__ void bbox_kernel(const float4* __restrict__ pts, int n, unsigned int* __restrict__ bbox) {
    unsigned int mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        if (!finite3(p.x, p.y, p.z)) continue;
        const unsigned int e[3] = {enc_f(p.x), enc_f(p.y), enc_f(p.z)};
#pragma unroll
        for (int a = 0; a < 3; ++a) { mn[a] = min(mn[a], e[a]); mx[a] = max(mx[a], e[a]); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        mn[a] = __reduce_min_sync(0xFFFFFFFFu, mn[a]);
        mx[a] = __reduce_max_sync(0xFFFFFFFFu, mx[a]);
    }
    // one atomic per block and value (thousands of same-address atomics from every warp cost tens of microseconds)
    __shared__ unsigned int s_mn[3][8], s_mx[3][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { s_mn[a][wid] = mn[a]; s_mx[a][wid] = mx[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int a = threadIdx.x % 3; const bool is_max = threadIdx.x >= 3;
        unsigned int v = is_max ? 0u : 0xFFFFFFFFu;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v = is_max ? max(v, s_mx[a][w]) : min(v, s_mn[a][w]);
        if (is_max) atomicMax(&bbox[3 + a], v); else atomicMin(&bbox[a], v);
    }
}

__global__ void grid_params_kernel(const unsigned int* __restrict__ bbox, int T, GridParams* __restrict__ g) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    GridParams P;
    float e[3], maxabs[3];
    const bool empty = bbox[0] == 0xFFFFFFFFu && bbox[3] == 0u;
    for (int a = 0; a < 3; ++a) {
        const float lo = empty ? 0.f : dec_f(bbox[a]), hi = empty ? 0.f : dec_f(bbox[3 + a]);
        P.o[a] = lo;
        maxabs[a] = fmaxf(fabsf(lo), fabsf(hi));
        e[a] = fmaxf(psub(hi, lo), padd(1e-20f, pmul(1e-6f, maxabs[a])));   // zero-extent axes stay well defined (no FMA: the oracle restates this)
        P.bits[a] = 0;
    }
    // Level k of the implicit tree halves the axis whose cells are currently the longest.
    float cur[3] = {e[0], e[1], e[2]};
    unsigned long long seq = 0ull;
    for (int k = 0; k < T; ++k) {
        int a = -1; float best = -1.f;
        for (int c = 0; c < 3; ++c) if (P.bits[c] < ICP_MAX_BITS_PER_AXIS && cur[c] > best) { best = cur[c]; a = c; }
        if (a < 0) a = 0;   // unreachable for T <= 3*ICP_MAX_BITS_PER_AXIS
        seq |= (unsigned long long)a << (2 * k);
        P.bits[a] += 1; cur[a] = pmul(cur[a], 0.5f);
    }
    for (int a = 0; a < 3; ++a) {
        P.h[a] = pdiv(pmul(e[a], 1.00001f), (float)(1 << P.bits[a]));
        P.inv_h[a] = pdiv(1.0f, P.h[a]);
        P.delta[a] = 1e-3f * P.h[a] + 1e-6f * maxabs[a];
    }
    P.T = T; P.axis_seq = seq; P.n_finite = 0; P.pad = 0;
    *g = P;
}

__device__ __forceinline__ int cell_index(const GridParams& g, int a, float x) {
    const float u = pmul(psub(x, g.o[a]), g.inv_h[a]);
    int i = (int)floorf(u);
    const int hi = (1 << g.bits[a]) - 1;
    return min(max(i, 0), hi);
}

__device__ __forceinline__ unsigned int cell_code(const GridParams& g, float x, float y, float z) {
    const int c0 = cell_index(g, 0, x), c1 = cell_index(g, 1, y), c2 = cell_index(g, 2, z);
    int r0 = g.bits[0], r1 = g.bits[1], r2 = g.bits[2];
    unsigned int code = 0;
    unsigned long long seq = g.axis_seq;
    for (int k = 0; k < g.T; ++k) {
        const int a = (int)(seq & 3ull); seq >>= 2;
        int bit;
        if (a == 0) { --r0; bit = (c0 >> r0) & 1; }
        else if (a == 1) { --r1; bit = (c1 >> r1) & 1; }
        else { --r2; bit = (c2 >> r2) & 1; }
        code = (code << 1) | (unsigned int)bit;
    }
    return code;
}

// key per point + rank within its cell (the count array becomes cell_start after the scan)
__global__ void keys_kernel(const float4* __restrict__ pts, int n, const GridParams* __restrict__ gp,
                            unsigned int* __restrict__ keys, unsigned int* __restrict__ ranks, unsigned int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const GridParams g = *gp;
    const float4 p = pts[i];
    if (!finite3(p.x, p.y, p.z)) { keys[i] = 0xFFFFFFFFu; return; }   // can never win the strict '>' scan of NearestNeighbor.h:87
    const unsigned int c = cell_code(g, p.x, p.y, p.z);
    keys[i] = c;
    ranks[i] = atomicAdd(&counts[c], 1u);
}

// ---------------------------------------------------------------------------- exclusive scan over 2^T+1 counters
#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ unsigned int block_exclusive_scan(unsigned int v, unsigned int* total) {
    __shared__ unsigned int warp_sums[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned int s = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0u, si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xFFFFFFFFu, si, o); if (lane >= o) si += t; }
        warp_sums[lane] = si - s;
        if (lane == 31 && total) *total = si;
    }
    __syncthreads();
    const unsigned int r = inc - v + warp_sums[w];
    __syncthreads();
    return r;
}

__global__ void scan_tile_sums_kernel(const unsigned int* __restrict__ data, int n, unsigned int* __restrict__ tile_sums) {
    const int base = blockIdx.x * SCAN_TILE;
    unsigned int s = 0;
    if (base + SCAN_TILE <= n) {                      // full tile: 16-byte loads (base is a multiple of 4096 entries)
        const uint4* d4 = reinterpret_cast<const uint4*>(data + base);
        for (int k = threadIdx.x; k < SCAN_TILE / 4; k += SCAN_THREADS) { const uint4 v = d4[k]; s += (v.x + v.y) + (v.z + v.w); }
    } else {
        for (int k = threadIdx.x; k < SCAN_TILE; k += SCAN_THREADS) { const int i = base + k; if (i < n) s += data[i]; }
    }
    s = __reduce_add_sync(0xFFFFFFFFu, s);
    __shared__ unsigned int ws[SCAN_THREADS / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned int t = 0; for (int w = 0; w < SCAN_THREADS / 32; ++w) t += ws[w]; tile_sums[blockIdx.x] = t; }
}

// single block: exclusive scan of up to 1024*8 tile sums in place
__global__ void scan_tile_offsets_kernel(unsigned int* __restrict__ tile_sums, int n_tiles) {
    __shared__ unsigned int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const unsigned int v = i < n_tiles ? tile_sums[i] : 0u;
        __shared__ unsigned int total;
        const unsigned int ex = block_exclusive_scan(v, &total);
        const unsigned int carry = carry_s;
        if (i < n_tiles) tile_sums[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
}

__global__ void scan_apply_kernel(unsigned int* __restrict__ data, int n, const unsigned int* __restrict__ tile_offsets) {
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned int v[SCAN_ITEMS]; unsigned int s = 0;
    const bool full = blockIdx.x * SCAN_TILE + SCAN_TILE <= n;      // 16-byte accesses on full tiles
    if (full) {
        const uint4* d4 = reinterpret_cast<const uint4*>(data + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) { const uint4 q = d4[k]; v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w; }
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) s += v[k];
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) { const int i = base + k; v[k] = i < n ? data[i] : 0u; s += v[k]; }
    }
    unsigned int ex = block_exclusive_scan(s, nullptr) + tile_offsets[blockIdx.x];
    if (full) {
        uint4* d4 = reinterpret_cast<uint4*>(data + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
            uint4 q;
            q.x = ex; ex += v[4 * k]; q.y = ex; ex += v[4 * k + 1]; q.z = ex; ex += v[4 * k + 2]; q.w = ex; ex += v[4 * k + 3];
            d4[k] = q;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) { const int i = base + k; if (i < n) data[i] = ex; ex += v[k]; }
    }
}

static cudaError_t launch_exclusive_scan(unsigned int* data, int n, unsigned int* block_sums, cudaStream_t s, int* launches) {
    const int n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_tile_sums_kernel<<<n_tiles, SCAN_THREADS, 0, s>>>(data, n, block_sums);
    scan_tile_offsets_kernel<<<1, 1024, 0, s>>>(block_sums, n_tiles);
    scan_apply_kernel<<<n_tiles, SCAN_THREADS, 0, s>>>(data, n, block_sums);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------- scatter into cell order
__global__ void scatter_kernel(const float4* __restrict__ pts, const float4* __restrict__ nrm, int n,
                               const unsigned int* __restrict__ keys, const unsigned int* __restrict__ ranks,
                               const unsigned int* __restrict__ cell_start, unsigned int n_cells, int keep_nonfinite,
                               unsigned int* __restrict__ nonfinite_counter, float4* __restrict__ pts_sorted,
                               float4* __restrict__ nrm_sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int k = keys[i];
    unsigned int pos;
    if (k == 0xFFFFFFFFu) {
        if (!keep_nonfinite) return;
        pos = cell_start[n_cells] + atomicAdd(nonfinite_counter, 1u);   // after the last cell; never part of any search
    } else {
        pos = cell_start[k] + ranks[i];
    }
    pts_sorted[pos] = pts[i];                        // .w already holds the original index: tie-break + API output
    nrm_sorted[pos] = nrm[i];
}

__global__ void extract_order_kernel(const float4* __restrict__ pts_sorted, int n, int* __restrict__ order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[i] = __float_as_int(pts_sorted[i].w);
}
cudaError_t icp_launch_extract_order(const float4* pts_sorted, int n, int* order, cudaStream_t s) {
    if (n > 0) extract_order_kernel<<<(n + 255) / 256, 256, 0, s>>>(pts_sorted, n, order);
    return cudaGetLastError();
}

__global__ void fill_int_kernel(int* __restrict__ p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// First-iteration seeds from the cell table: a query that remembers no neighbour yet (nn_pos < 0) gets a point of the
// smallest cell-tree node around its transformed position that holds any point (binary search over the depth: the count
// of the node around a position is monotone in the depth).  The search uses a seed only as its starting bound and first
// leaf, so any target point is a valid seed -- a near one lets the first iteration run like the later ones (fast path
// for most queries) instead of walking the tree from the root for every query.
__global__ void seed_from_grid_kernel(const float4* __restrict__ src_pts, int n_src, const DevState* __restrict__ st,
                                      const GridParams* __restrict__ gp, const unsigned int* __restrict__ cs, int T,
                                      const unsigned int* __restrict__ leaf_rank, int* __restrict__ nn_pos, int* __restrict__ nn_leaf) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_src) return;
    if (nn_pos[p] >= 0) return;
    const unsigned int n_finite = cs[(size_t)1 << T];
    if (n_finite == 0u) return;
    const float4 s = src_pts[p];
    float P[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) P[k] = st->pose[k];
    float x, y, z;
    xform_point(P, s.x, s.y, s.z, x, y, z);
    if (!finite3(x, y, z)) return;
    const GridParams g = *gp;
    const unsigned int c = cell_code(g, x, y, z);
    int lo = 0, hi = T;                                  // smallest shift whose node is not empty (shift T = the whole cloud)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const unsigned int pre = c >> mid;
        const unsigned int cnt = cs[((size_t)pre + 1) << mid] - cs[(size_t)pre << mid];
        if (cnt > 0u) hi = mid; else lo = mid + 1;
    }
    const unsigned int pre = c >> lo;
    const unsigned int s0 = cs[(size_t)pre << lo], e0 = cs[((size_t)pre + 1) << lo];
    if (e0 <= s0) return;
    const unsigned int pos = s0 + ((e0 - s0) >> 1);
    nn_pos[p] = (int)pos;
    nn_leaf[p] = (int)(leaf_rank[pos + 1] - 1u);
}

cudaError_t icp_launch_seed_from_grid(const float4* src_pts, int n_src, const DevState* st, const GridParams* grid, const unsigned int* cell_start,
                                      int T, const unsigned int* leaf_rank, int* nn_pos, int* nn_leaf, cudaStream_t s) {
    if (n_src <= 0) return cudaSuccess;
    seed_from_grid_kernel<<<(n_src + 255) / 256, 256, 0, s>>>(src_pts, n_src, st, grid, cell_start, T, leaf_rank, nn_pos, nn_leaf);
    return cudaGetLastError();
}

cudaError_t icp_launch_fill_int(int* p, int n, int v, cudaStream_t s) {
    if (n > 0) fill_int_kernel<<<(n + 255) / 256, 256, 0, s>>>(p, n, v);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------- refinement of over-full finest cells
// The dense table caps the sort key at T <= 24 bits; where the cloud is much denser than the grid (near the sensor) a
// finest cell still holds dozens of points in arbitrary order.  Those cells get a local counting sort by 6 more bits
// (2 per axis: the position inside the cell), so that the runs of 32 the leaves are cut from are compact.
#define REFINE_MAX 1024
// thread per sorted point: the first point of a finest cell with 33 .. REFINE_MAX points registers the cell
__global__ void find_overfull_cells_kernel(const float4* __restrict__ pts, int n, const GridParams* __restrict__ gp,
                                           const unsigned int* __restrict__ cs, int T, unsigned int* __restrict__ list,
                                           unsigned int* __restrict__ n_list, unsigned int capacity) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (unsigned int)i >= cs[(size_t)1 << T]) return;
    const GridParams g = *gp;
    const float4 p = pts[i];
    const unsigned int c = cell_code(g, p.x, p.y, p.z);
    const unsigned int s = cs[c], cnt = cs[(size_t)c + 1] - s;
    if ((unsigned int)i == s && cnt > 32u && cnt <= REFINE_MAX) { const unsigned int k = atomicAdd(n_list, 1u); if (k < capacity) list[k] = c; }
}

__global__ void __launch_bounds__(128) refine_cells_kernel(const unsigned int* __restrict__ cs, const GridParams* __restrict__ gp,
                                                           const unsigned int* __restrict__ list, const unsigned int* __restrict__ n_list,
                                                           unsigned int capacity, float4* __restrict__ pts, float4* __restrict__ nrm) {
    __shared__ float4 sp[REFINE_MAX];
    __shared__ float4 sn[REFINE_MAX];
    __shared__ unsigned short skey[REFINE_MAX];
    __shared__ unsigned int hist[64];
    const GridParams g = *gp;
    const unsigned int n = min(*n_list, capacity);
    for (unsigned int w = blockIdx.x; w < n; w += gridDim.x) {
        const unsigned int c = list[w], s = cs[c], cnt = cs[c + 1] - s;
        if (threadIdx.x < 64) hist[threadIdx.x] = 0u;
        __syncthreads();
        for (unsigned int k = threadIdx.x; k < cnt; k += blockDim.x) {
            const float4 p = pts[s + k];
            sp[k] = p; sn[k] = nrm[s + k];
            unsigned int key = 0;
            const float x[3] = {p.x, p.y, p.z};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float u = pmul(psub(x[a], g.o[a]), g.inv_h[a]);
                const float f = u - floorf(u);                      // position inside the cell (clamped cells: any value is fine)
                const int q = min(max((int)(f * 4.0f), 0), 3);
                key |= (unsigned int)(((q >> 1) & 1) << (5 - a)) | (unsigned int)((q & 1) << (2 - a));
            }
            skey[k] = (unsigned short)key;
            atomicAdd(&hist[key], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) { unsigned int acc = 0; for (int b = 0; b < 64; ++b) { const unsigned int v = hist[b]; hist[b] = acc; acc += v; } }
        __syncthreads();
        for (unsigned int k = threadIdx.x; k < cnt; k += blockDim.x) {
            const unsigned int pos = atomicAdd(&hist[skey[k]], 1u);
            pts[s + pos] = sp[k]; nrm[s + pos] = sn[k];
        }
        __syncthreads();
    }
}

cudaError_t icp_launch_refine_cells(const unsigned int* cell_start, int T, const GridParams* grid, unsigned int* list, unsigned int capacity,
                                    unsigned int* n_list, float4* pts_sorted, float4* nrm_sorted, int n, int n_sms, cudaStream_t s,
                                    int* n_launches) {
    cudaError_t e;
    if ((e = cudaMemsetAsync(n_list, 0, sizeof(unsigned int), s)) != cudaSuccess) return e;
    if (n <= 0) return cudaSuccess;
    find_overfull_cells_kernel<<<(n + 255) / 256, 256, 0, s>>>(pts_sorted, n, grid, cell_start, T, list, n_list, capacity);
    refine_cells_kernel<<<n_sms * 4, 128, 0, s>>>(cell_start, grid, list, n_list, capacity, pts_sorted, nrm_sorted);
    if (n_launches) *n_launches += 2;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------- BVH over the sorted cloud
// Leaves = nodes of the implicit cell tree with <= 32 points whose parent has more (an over-full finest cell is
// cut into runs of 32): cell-aligned, hence pairwise disjoint in space -- a search ball meets only the few
// leaves around it, unlike fixed runs of the Z-curve, whose boxes straddle the curve's jumps.
__global__ void mark_leaves_kernel(const float4* __restrict__ pts, int n, const GridParams* __restrict__ gp,
                                   const unsigned int* __restrict__ cs, int T, unsigned int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const unsigned int n_finite = cs[(size_t)1 << T];
    if ((unsigned int)i >= n_finite) { flags[i] = 0u; return; }
    const GridParams g = *gp;
    const float4 p = pts[i];
    const unsigned int c = cell_code(g, p.x, p.y, p.z);
    // smallest depth whose node holds <= 32 points (the count is monotone in the depth)
    const unsigned int sT = cs[c], eT = cs[(size_t)c + 1];
    if (eT - sT > 32u) { flags[i] = ((unsigned int)i - sT) % 32u == 0u ? 1u : 0u; return; }
    int lo = 0, hi = T;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1, sh = T - mid;
        const unsigned int pre = c >> sh;
        const unsigned int cnt = cs[((size_t)pre + 1) << sh] - cs[(size_t)pre << sh];
        if (cnt <= 32u) hi = mid; else lo = mid + 1;
    }
    const int sh = T - lo;
    flags[i] = (unsigned int)i == cs[(size_t)(c >> sh) << sh] ? 1u : 0u;
}

// flags have been exclusive-scanned in place: rank[i] = number of leaf starts before i
__global__ void leaf_starts_kernel(const unsigned int* __restrict__ rank, int n, const unsigned int* __restrict__ cs, int T,
                                   unsigned int* __restrict__ leaf_start, BvhDesc* __restrict__ bvh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int n_finite = cs[(size_t)1 << T];
    if (i < n && (unsigned int)i < n_finite && rank[i + 1] != rank[i]) leaf_start[rank[i]] = (unsigned int)i;
    if (i == 0) {
        const int L = (int)rank[n_finite];
        leaf_start[L] = n_finite;
        BvhDesc b;
        b.n_leaves = L; b.n_levels = 1;
        for (int k = 0; k < ICP_BVH_MAX_LEVELS; ++k) { b.count[k] = 0; b.offset[k] = 0; b.coffset[k] = 0; }
        b.count[0] = L;
        *bvh = b;
    }
}

// Number of level-`lvl` nodes that start before sorted point position s (s is a boundary of a cell-tree node at or
// above the level's nodes): leaves by leaf_rank, upper levels by the chain of per-level ranks.
__device__ __forceinline__ unsigned int nodes_before(const BvhDesc& b, const unsigned int* __restrict__ leaf_rank,
                                                     const unsigned int* __restrict__ node_rank, int lvl, unsigned int s) {
    unsigned int a = leaf_rank[s];
    for (int j = 1; j <= lvl; ++j) a = node_rank[b.coffset[j] + a];
    return a;
}

// Level `lvl` (>= 1) from level lvl-1, step 1: node k of level lvl-1 starts a level-lvl node iff it is the first node of
// the shallowest cell-tree node around it that holds <= 32 nodes of level lvl-1.  flags -> node_rank[coffset[lvl] + k].
__global__ void mark_level_kernel(const float4* __restrict__ pts, const GridParams* __restrict__ gp, const unsigned int* __restrict__ cs,
                                  int T, const unsigned int* __restrict__ leaf_start, const unsigned int* __restrict__ leaf_rank,
                                  const unsigned int* __restrict__ pstart, unsigned int* __restrict__ node_rank,
                                  const BvhDesc* __restrict__ bvh, int lvl) {
    const BvhDesc b = *bvh;
    if (b.n_levels != lvl) return;                         // the level below is not the current top, or the tree is complete
    const int prev = lvl - 1, n_prev = b.count[prev];
    if (n_prev <= 32) return;                              // the level below already is the top
    const GridParams g = *gp;
    unsigned int* flags = node_rank + b.coffset[lvl];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= n_prev; k += gridDim.x * blockDim.x) {
        if (k == n_prev) { flags[k] = 0u; continue; }
        const unsigned int pos = prev == 0 ? leaf_start[k] : pstart[b.offset[prev] + k];
        const float4 p = pts[pos];
        const unsigned int c = cell_code(g, p.x, p.y, p.z);
        const unsigned int sT = cs[c], eT = cs[(size_t)c + 1];
        const unsigned int aT = nodes_before(b, leaf_rank, node_rank, prev, sT), bT = nodes_before(b, leaf_rank, node_rank, prev, eT);
        if (bT - aT > 32u) { flags[k] = ((unsigned int)k - aT) % 32u == 0u ? 1u : 0u; continue; }   // over-full finest cell: runs of 32
        int lo = 0, hi = T;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1, sh = T - mid;
            const unsigned int pre = c >> sh;
            const unsigned int cnt = nodes_before(b, leaf_rank, node_rank, prev, cs[((size_t)pre + 1) << sh]) -
                                     nodes_before(b, leaf_rank, node_rank, prev, cs[(size_t)pre << sh]);
            if (cnt <= 32u) hi = mid; else lo = mid + 1;
        }
        const int sh = T - lo;
        flags[k] = (unsigned int)k == nodes_before(b, leaf_rank, node_rank, prev, cs[(size_t)(c >> sh) << sh]) ? 1u : 0u;
    }
}

// step 2 (one block): exclusive scan of the level's flags in place; the descriptor gains the level
__global__ void scan_level_kernel(unsigned int* __restrict__ node_rank, BvhDesc* __restrict__ bvh, int lvl) {
    __shared__ unsigned int carry_s;
    __shared__ unsigned int total;
    const BvhDesc b = *bvh;
    if (b.n_levels != lvl || b.count[lvl - 1] <= 32) return;
    const int n = b.count[lvl - 1] + 1;
    unsigned int* data = node_rank + b.coffset[lvl];
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const unsigned int v = i < n ? data[i] : 0u;
        const unsigned int ex = block_exclusive_scan(v, &total);
        const unsigned int carry = carry_s;
        if (i < n) data[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int cnt = (int)carry_s;
        bvh->count[lvl] = cnt;
        bvh->offset[lvl] = b.offset[lvl - 1] + b.count[lvl - 1];
        if (lvl + 1 < ICP_BVH_MAX_LEVELS) bvh->coffset[lvl + 1] = b.coffset[lvl] + n;     // this level used n rank entries and cnt+1 <= n child entries
        bvh->n_levels = lvl + 1;
    }
}

// step 3: children ranges and start positions of the level's nodes
__global__ void level_children_kernel(const unsigned int* __restrict__ leaf_start, const unsigned int* __restrict__ node_rank,
                                      unsigned int* __restrict__ child_start, unsigned int* __restrict__ pstart,
                                      const BvhDesc* __restrict__ bvh, int lvl) {
    const BvhDesc b = *bvh;
    if (b.n_levels != lvl + 1) return;                     // the level was not created
    const int prev = lvl - 1, n_prev = b.count[prev];
    const unsigned int* rank = node_rank + b.coffset[lvl];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= n_prev; k += gridDim.x * blockDim.x) {
        if (k == n_prev) { child_start[b.coffset[lvl] + b.count[lvl]] = (unsigned int)n_prev; continue; }
        if (rank[k + 1] != rank[k]) {
            child_start[b.coffset[lvl] + rank[k]] = (unsigned int)k;
            pstart[b.offset[lvl] + rank[k]] = prev == 0 ? leaf_start[k] : pstart[b.offset[prev] + k];
        }
    }
}

// one warp per node (grid-stride): level 0 reads the leaf's points, level l > 0 reads its (<= 32) child boxes
// The w components of the two box entries carry the node's COLOUR range (min / max of r, g, b as packed bytes), which the
// 6-D colour search adds to its lower bounds (match.cu: box_dist2c).
__global__ void bvh_level_kernel(const float4* __restrict__ pts, const float4* __restrict__ nrm, const unsigned int* __restrict__ leaf_start,
                                 const unsigned int* __restrict__ child_start, const BvhDesc* __restrict__ bvh, float4* __restrict__ box,
                                 int level) {
    const BvhDesc b = *bvh;
    if (level >= b.n_levels) return;
    const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
    const int count = b.count[level], offset = b.offset[level];
    for (int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; node < count; node += warps) {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        unsigned int clo[3] = {255u, 255u, 255u}, chi[3] = {0u, 0u, 0u};
        if (level == 0) {
            const unsigned int i = leaf_start[node] + lane;
            if (i < leaf_start[node + 1]) {
                const float4 p = pts[i]; lo[0] = hi[0] = p.x; lo[1] = hi[1] = p.y; lo[2] = hi[2] = p.z;
                const unsigned int c = __float_as_uint(nrm[i].w);
#pragma unroll
                for (int k = 0; k < 3; ++k) clo[k] = chi[k] = (c >> (8 * k)) & 0xFFu;
            }
        } else {
            const unsigned int c = child_start[b.coffset[level] + node] + lane;
            if (c < child_start[b.coffset[level] + node + 1]) {
                const float4 u = box[2 * (size_t)(b.offset[level - 1] + c)], v = box[2 * (size_t)(b.offset[level - 1] + c) + 1];
                lo[0] = u.x; lo[1] = u.y; lo[2] = u.z; hi[0] = v.x; hi[1] = v.y; hi[2] = v.z;
                const unsigned int cu = __float_as_uint(u.w), cv = __float_as_uint(v.w);
#pragma unroll
                for (int k = 0; k < 3; ++k) { clo[k] = (cu >> (8 * k)) & 0xFFu; chi[k] = (cv >> (8 * k)) & 0xFFu; }
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { clo[k] = __reduce_min_sync(0xFFFFFFFFu, clo[k]); chi[k] = __reduce_max_sync(0xFFFFFFFFu, chi[k]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], o)); }
        if (lane == 0) {
            box[2 * (size_t)(offset + node)] = make_float4(lo[0], lo[1], lo[2], __uint_as_float(clo[0] | (clo[1] << 8) | (clo[2] << 16)));
            box[2 * (size_t)(offset + node) + 1] = make_float4(hi[0], hi[1], hi[2], __uint_as_float(chi[0] | (chi[1] << 8) | (chi[2] << 16)));
        }
    }
}

// ---------------------------------------------------------------------------- voxel pyramid levels
// ICP_GPU_PYRAMID_VOXEL: a pyramid level keeps ONE point per occupied cell of the source grid at depth D -- the valid
// (finite point and normal) point with the lowest original index -- instead of every f-th point of the scan order
// (PointCloud.h:325-343).  The level is delivered as a selection mask (bit per original index).
__global__ void voxel_min_kernel(const float4* __restrict__ pts, const float4* __restrict__ nrm, int n, const GridParams* __restrict__ gp,
                                 int T, int D, unsigned int* __restrict__ table) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i], m = nrm[i];
    if (!finite3(p.x, p.y, p.z) || !finite3(m.x, m.y, m.z)) return;
    const GridParams g = *gp;
    atomicMin(&table[cell_code(g, p.x, p.y, p.z) >> (T - D)], (unsigned int)__float_as_int(p.w));
}
__global__ void voxel_mask_kernel(const float4* __restrict__ pts, const float4* __restrict__ nrm, int n, const GridParams* __restrict__ gp,
                                  int T, int D, const unsigned int* __restrict__ table, unsigned int* __restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i], m = nrm[i];
    if (!finite3(p.x, p.y, p.z) || !finite3(m.x, m.y, m.z)) return;
    const GridParams g = *gp;
    const unsigned int o = (unsigned int)__float_as_int(p.w);
    if (table[cell_code(g, p.x, p.y, p.z) >> (T - D)] == o) atomicOr(&mask[o >> 5], 1u << (o & 31u));
}

cudaError_t icp_launch_voxel_level(const float4* pts_sorted, const float4* nrm_sorted, int n, const GridParams* grid, int T, int D,
                                   unsigned int* table, unsigned int* mask, size_t mask_words, cudaStream_t s, int* n_launches) {
    cudaError_t e;
    if ((e = cudaMemsetAsync(table, 0xFF, sizeof(unsigned int) * ((size_t)1 << D), s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(mask, 0, sizeof(unsigned int) * mask_words, s)) != cudaSuccess) return e;
    if (n > 0) {
        voxel_min_kernel<<<(n + 255) / 256, 256, 0, s>>>(pts_sorted, nrm_sorted, n, grid, T, D, table);
        voxel_mask_kernel<<<(n + 255) / 256, 256, 0, s>>>(pts_sorted, nrm_sorted, n, grid, T, D, table, mask);
        if (n_launches) *n_launches += 2;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------- leaf adjacency
// One warp per leaf l: all other leaves whose box meets box(l) inflated by R, found by a box query on the BVH
// (lane = child).  R starts at twice the leaf's largest extent and is halved until the list fits in 32 entries.
#define ADJ_WARPS 4
__device__ __forceinline__ bool boxes_meet(const float* lo, const float* hi, const float4 blo, const float4 bhi) {
    return blo.x <= hi[0] && bhi.x >= lo[0] && blo.y <= hi[1] && bhi.y >= lo[1] && blo.z <= hi[2] && bhi.z >= lo[2];
}

__global__ void __launch_bounds__(ADJ_WARPS * 32) leaf_adjacency_kernel(const BvhDesc* __restrict__ bvh, const float4* __restrict__ box,
                                                                         const unsigned int* __restrict__ child_start,
                                                                         unsigned int* __restrict__ adj, float4* __restrict__ adj_box,
                                                                         int capacity, float r_factor, int lv) {
    __shared__ unsigned int s_node[ADJ_WARPS][32 * ICP_BVH_MAX_LEVELS];
    __shared__ unsigned int s_list[ADJ_WARPS][64];
    const BvhDesc b = *bvh;
    const unsigned int FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, warps = (gridDim.x * blockDim.x) >> 5;
    // lv = 0: lists of leaves (the node itself excluded: a search has scanned it already); lv = 1: lists of level-1 nodes
    // (the node itself included)
    if (lv >= b.n_levels) return;
    const int n_leaves = min(b.count[lv], capacity);
    const int top_level = b.n_levels - 1;
    unsigned int* st = s_node[wid]; unsigned int* list = s_list[wid];
    const unsigned int lt = (1u << lane) - 1u;
    for (int l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; l < n_leaves; l += warps) {
        const float4 mlo = box[2 * (size_t)(b.offset[lv] + l)], mhi = box[2 * (size_t)(b.offset[lv] + l) + 1];
        float R = r_factor * fmaxf(fmaxf(mhi.x - mlo.x, mhi.y - mlo.y), mhi.z - mlo.z);
        int count = 0; bool ok = false;
        for (int attempt = 0; attempt < 6 && !ok; ++attempt, R *= 0.5f) {
            const float lo[3] = {__fsub_rd(mlo.x, R), __fsub_rd(mlo.y, R), __fsub_rd(mlo.z, R)};
            const float hi[3] = {__fadd_ru(mhi.x, R), __fadd_ru(mhi.y, R), __fadd_ru(mhi.z, R)};
            count = 0; ok = true;
            int top = 0;
            // virtual root over the top level, then depth-first; leaves are appended to the list
            for (unsigned int base = 0; base < (unsigned int)b.count[top_level] && ok; base += 32) {
                int L = top_level; unsigned int first = base, last = min(base + 32u, (unsigned int)b.count[top_level]);
                for (;;) {
                    const unsigned int c = first + lane;
                    bool keep = false;
                    if (c < last) keep = boxes_meet(lo, hi, box[2 * (size_t)(b.offset[L] + c)], box[2 * (size_t)(b.offset[L] + c) + 1]);
                    if (L == lv && lv == 0) keep = keep && c != (unsigned int)l;
                    const unsigned int mk = __ballot_sync(FULL, keep);
                    if (L == lv) {
                        if (count + __popc(mk) > 32) { ok = false; break; }
                        if (keep) list[count + __popc(mk & lt)] = c;
                        count += __popc(mk);
                    } else {
                        if (keep) st[top + __popc(mk & lt)] = ((unsigned int)L << 27) | c;
                        top += __popc(mk);
                    }
                    __syncwarp();
                    if (top == 0) break;
                    --top;
                    const unsigned int id = st[top];
                    __syncwarp();
                    const int lvl = (int)(id >> 27); const unsigned int j = id & 0x7FFFFFFu;
                    first = child_start[b.coffset[lvl] + j]; last = child_start[b.coffset[lvl] + j + 1];
                    L = lvl - 1;
                }
            }
            if (ok) break;
        }
        __syncwarp();
        if (lane < count && ok) adj[(size_t)l * 32 + lane] = list[lane];
        if (lane == 0) {
            // the inflated box exactly as the query above used it (so that "ball inside this box" implies "every leaf the
            // ball meets is in the list"); inverted when there is no list
            if (ok) {
                adj_box[2 * (size_t)l] = make_float4(__fsub_rd(mlo.x, R), __fsub_rd(mlo.y, R), __fsub_rd(mlo.z, R), __int_as_float(count));
                adj_box[2 * (size_t)l + 1] = make_float4(__fadd_ru(mhi.x, R), __fadd_ru(mhi.y, R), __fadd_ru(mhi.z, R), 0.f);
            } else {
                adj_box[2 * (size_t)l] = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(0));
                adj_box[2 * (size_t)l + 1] = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
            }
        }
        __syncwarp();
    }
}

cudaError_t icp_launch_leaf_adjacency(const BvhDesc* bvh_dev, const float4* box, const unsigned int* child_start, unsigned int* adj,
                                      float4* adj_box, int capacity, int level, int n_sms, cudaStream_t s, int* n_launches) {
    long long nb = ((long long)capacity + ADJ_WARPS - 1) / ADJ_WARPS;
    if (nb > 16ll * n_sms) nb = 16ll * n_sms;
    if (nb < 1) nb = 1;
    float r_factor = 2.0f;
    if (const char* e = getenv("ICP_GPU_ADJ_FACTOR")) r_factor = (float)atof(e);   // tuning knob
    leaf_adjacency_kernel<<<(int)nb, ADJ_WARPS * 32, 0, s>>>(bvh_dev, box, child_start, adj, adj_box, capacity, r_factor, level);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

size_t icp_bvh_max_nodes(int n) {
    // worst case: every point its own leaf, and every upper level only halves the node count until the cap
    return (size_t)(n > 0 ? n : 1) * 2 + 64;
}

cudaError_t icp_launch_bvh_build(const float4* pts_sorted, const float4* nrm_sorted, int n, const GridParams* grid, const unsigned int* cell_start, int T,
                                 unsigned int* leaf_rank, unsigned int* block_sums, unsigned int* leaf_start, unsigned int* node_rank,
                                 unsigned int* child_start, unsigned int* pstart, BvhDesc* bvh_dev, float4* box, int n_sms,
                                 cudaStream_t s, int* n_launches) {
    int launches = 0;
    mark_leaves_kernel<<<(n + 1 + 255) / 256, 256, 0, s>>>(pts_sorted, n, grid, cell_start, T, leaf_rank); ++launches;
    cudaError_t e = launch_exclusive_scan(leaf_rank, n + 1, block_sums, s, &launches);
    if (e != cudaSuccess) return e;
    leaf_starts_kernel<<<(n + 1 + 255) / 256, 256, 0, s>>>(leaf_rank, n, cell_start, T, leaf_start, bvh_dev); ++launches;
    long long nb0 = ((long long)(n > 0 ? n : 1) + 7) / 8; if (nb0 > 8ll * n_sms) nb0 = 8ll * n_sms;
    bvh_level_kernel<<<(int)nb0, 256, 0, s>>>(pts_sorted, nrm_sorted, leaf_start, child_start, bvh_dev, box, 0); ++launches;
    // Upper levels: the node counts live on the device, so every possible level gets its (tiny) launches; the ones
    // past the top return at once.  An n-point cloud with healthy fill needs log_16(n / 16) levels; cap the launches there + 2.
    int max_levels = 2; { long long c = (n > 0 ? n : 1) / 16; while (c > 32 && max_levels < ICP_BVH_MAX_LEVELS) { c /= 8; ++max_levels; } }
    if (max_levels > ICP_BVH_MAX_LEVELS) max_levels = ICP_BVH_MAX_LEVELS;
    for (int l = 1; l < max_levels; ++l) {
        const int nb = l == 1 ? 2 * n_sms : n_sms / 2;
        mark_level_kernel<<<nb, 256, 0, s>>>(pts_sorted, grid, cell_start, T, leaf_start, leaf_rank, pstart, node_rank, bvh_dev, l); ++launches;
        scan_level_kernel<<<1, 1024, 0, s>>>(node_rank, bvh_dev, l); ++launches;
        level_children_kernel<<<nb, 256, 0, s>>>(leaf_start, node_rank, child_start, pstart, bvh_dev, l); ++launches;
        bvh_level_kernel<<<nb, 256, 0, s>>>(pts_sorted, nrm_sorted, leaf_start, child_start, bvh_dev, box, l); ++launches;
    }
    if (n_launches) *n_launches += launches;
    return cudaGetLastError();
}

cudaError_t icp_launch_grid_build(const float4* pts_in, const float4* nrm_in, int n, int T, GridParams* grid,
                                  unsigned int* bbox_scratch, unsigned int* keys, unsigned int* ranks, unsigned int* cell_start,
                                  unsigned int* block_sums, float4* pts_sorted, float4* nrm_sorted, int keep_nonfinite,
                                  cudaStream_t s, int* n_launches) {
    cudaError_t e;
    const int n_cells1 = (1 << T) + 1;
    if ((e = cudaMemsetAsync(bbox_scratch, 0xFF, 3 * sizeof(unsigned int), s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(bbox_scratch + 3, 0x00, 5 * sizeof(unsigned int), s)) != cudaSuccess) return e;   // max[3], spare, non-finite counter
    if ((e = cudaMemsetAsync(cell_start, 0, sizeof(unsigned int) * (size_t)n_cells1, s)) != cudaSuccess) return e;
    int launches = 0;
    if (n > 0) {
        const int nb = min((n + 255) / 256, 148 * 8);
        bbox_kernel<<<nb, 256, 0, s>>>(pts_in, n, bbox_scratch); ++launches;
    }
    grid_params_kernel<<<1, 32, 0, s>>>(bbox_scratch, T, grid); ++launches;
    if (n > 0) { keys_kernel<<<(n + 255) / 256, 256, 0, s>>>(pts_in, n, grid, keys, ranks, cell_start); ++launches; }
    const int n_tiles = (n_cells1 + SCAN_TILE - 1) / SCAN_TILE;
    scan_tile_sums_kernel<<<n_tiles, SCAN_THREADS, 0, s>>>(cell_start, n_cells1, block_sums); ++launches;
    scan_tile_offsets_kernel<<<1, 1024, 0, s>>>(block_sums, n_tiles); ++launches;
    scan_apply_kernel<<<n_tiles, SCAN_THREADS, 0, s>>>(cell_start, n_cells1, block_sums); ++launches;
    if (n > 0) {
        scatter_kernel<<<(n + 255) / 256, 256, 0, s>>>(pts_in, nrm_in, n, keys, ranks, cell_start, 1u << T, keep_nonfinite, bbox_scratch + 7,
                                                       pts_sorted, nrm_sorted);
        ++launches;
    }
    if (n_launches) *n_launches += launches;
    return cudaGetLastError();
}
