/*
 * icp_oracle.c -- CPU restatement of the ICP-Variants registration inner loop (see icp_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY.  Pinned against the reference's own headers compiled in place
 * (oracle/_ref, tests/test_oracle_vs_reference.py, tests/golden/reference_outputs.npz); the absent
 * third-party libraries' last bits stay a stated contract -- see the header.  Build: oracle/Makefile
 *   gcc -std=c11 -O2 -ffp-contract=off -fopenmp -shared -fPIC
 * -ffp-contract=off is part of the numerics contract: every fp32 expression below is evaluated
 * exactly as written (no FMA), left to right.
 *
 * Numerics contract: D1 squared distance fp32 ((dx*dx+dy*dy)+dz*dz) [+dr^2+dg^2+db^2];
 * D2 ties -> lowest target index; D3 valid iff d2 <= max (fp32); D4 transform
 * ((r0*p0+r1*p1)+r2*p2)+t in fp32; normals with R^-T built once from fp32 cofactors;
 * D5 normal equations / moments in fp64 from fp32 inputs promoted to double, solve in fp64,
 * increment rounded to fp32, pose product in fp32; D6 rejection cos <= 0.5f; D7 mt19937 masks.
 */
#include "icp_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MINF (-INFINITY)

static int g_threads = 0;
int orc_num_threads(void) {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
    g_threads = n;
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

void orc_default_config(orc_config* c) {
    /* ICPOptimizer.h:29-31 constructor defaults */
    memset(c, 0, sizeof(*c));
    c->metric = 0; c->minimizer = 0; c->matching = 0; c->selection = 0; c->proba = 1.0; c->seed = 0;
    c->weighting = 0; c->rejection = 1; c->max_distance_sq = 0.0003f; c->color_icp = 0; c->multires = 0;
    c->n_iterations = 20; c->nn_mode = ORC_NN_KDTREE; c->lm_max_iterations = 10; c->pyramid_mode = 0;
}

static inline int finite3(const float* p) { return isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]); }

/* ------------------------------------------------------------------ transforms (utils.h:106-133) */

void orc_transform_points(const float P[16], const float* pts, int64_t n, float* out) {
    /* utils.h:106-118: rotation * point + translation, fp32. Contract D4. */
    const float r00 = P[0], r10 = P[1], r20 = P[2], r01 = P[4], r11 = P[5], r21 = P[6], r02 = P[8], r12 = P[9], r22 = P[10];
    const float t0 = P[12], t1 = P[13], t2 = P[14];
    for (int64_t i = 0; i < n; ++i) {
        const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        out[3 * i + 0] = ((r00 * x + r01 * y) + r02 * z) + t0;
        out[3 * i + 1] = ((r10 * x + r11 * y) + r12 * z) + t1;
        out[3 * i + 2] = ((r20 * x + r21 * y) + r22 * z) + t2;
    }
}

static void inv_transpose3(const float P[16], float N[9] /* row-major N[i*3+j] */) {
    /* utils.h:129: rotation.inverse().transpose(); Eigen's 3x3 inverse is the cofactor formula.
     * (R^-1)^T_ij = cof_ij / det.  Computed once (the reference recomputes it per normal). */
    const float r00 = P[0], r10 = P[1], r20 = P[2], r01 = P[4], r11 = P[5], r21 = P[6], r02 = P[8], r12 = P[9], r22 = P[10];
    const float c00 = r11 * r22 - r12 * r21, c01 = r12 * r20 - r10 * r22, c02 = r10 * r21 - r11 * r20;
    const float c10 = r02 * r21 - r01 * r22, c11 = r00 * r22 - r02 * r20, c12 = r01 * r20 - r00 * r21;
    const float c20 = r01 * r12 - r02 * r11, c21 = r02 * r10 - r00 * r12, c22 = r00 * r11 - r01 * r10;
    const float det = (r00 * c00 + r01 * c01) + r02 * c02;
    const float id = 1.0f / det;
    N[0] = c00 * id; N[1] = c01 * id; N[2] = c02 * id;
    N[3] = c10 * id; N[4] = c11 * id; N[5] = c12 * id;
    N[6] = c20 * id; N[7] = c21 * id; N[8] = c22 * id;
}

void orc_transform_normals(const float P[16], const float* nrm, int64_t n, float* out) {
    float N[9];
    inv_transpose3(P, N);
    for (int64_t i = 0; i < n; ++i) {
        const float x = nrm[3 * i], y = nrm[3 * i + 1], z = nrm[3 * i + 2];
        out[3 * i + 0] = (N[0] * x + N[1] * y) + N[2] * z;
        out[3 * i + 1] = (N[3] * x + N[4] * y) + N[5] * z;
        out[3 * i + 2] = (N[6] * x + N[7] * y) + N[8] * z;
    }
}

static void mat4_mul(const float A[16], const float B[16], float C[16]) {
    /* Matrix4f * Matrix4f in fp32 (ICPOptimizer.h:614-620), sum left to right. */
    float T[16];
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
            T[i + 4 * j] = ((A[i] * B[4 * j] + A[i + 4] * B[1 + 4 * j]) + A[i + 8] * B[2 + 4 * j]) + A[i + 12] * B[3 + 4 * j];
    memcpy(C, T, sizeof(T));
}
static void mat4_identity(float M[16]) { memset(M, 0, 16 * sizeof(float)); M[0] = M[5] = M[10] = M[15] = 1.f; }

/* ------------------------------------------------------------------ exact 1-NN */

static inline float d2_3(const float* q, const float* p) {
    /* D1: FLANN L2<float> accumulation order for 3 dims; same as weighting.h:19 */
    const float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
    return (dx * dx + dy * dy) + dz * dz;
}
static inline void feat6(const float* p, const uint8_t* c, float f[6]) {
    /* NearestNeighbor.h:212-221,245-254: color_scale(1) * color_normalize(1/float(255)) * uchar */
    const float color_normalize = 1 / (float)255;
    const float color_scale = 1;
    f[0] = p[0]; f[1] = p[1]; f[2] = p[2];
    f[3] = color_scale * color_normalize * c[0];
    f[4] = color_scale * color_normalize * c[1];
    f[5] = color_scale * color_normalize * c[2];
}
static inline float d2_6(const float* a, const float* b) {
    float r = 0.f;
    /* FLANN L2: first four dims as one group ((d0^2+d1^2)+d2^2)+d3^2, then one add per dim */
    const float d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2], d3 = a[3] - b[3], d4 = a[4] - b[4], d5 = a[5] - b[5];
    r = ((d0 * d0 + d1 * d1) + d2 * d2) + d3 * d3;
    r = r + d4 * d4;
    r = r + d5 * d5;
    return r;
}

void orc_knn3_brute(const float* tgt, int64_t nt, const float* qry, int64_t nq, float max_d2, orc_match* out) {
    /* NearestNeighbor.h:81-97 (strict '>' => lowest index on ties) on squared distances,
     * thresholded as NearestNeighborSearchFlann does (:181-186). */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; ++i) {
        float best = FLT_MAX; int32_t bi = -1;
        for (int64_t j = 0; j < nt; ++j) {
            const float d = d2_3(qry + 3 * i, tgt + 3 * j);
            if (best > d) { best = d; bi = (int32_t)j; }
        }
        if (bi >= 0 && best <= max_d2) { out[i].idx = bi; out[i].weight = 1.f; }
        else { out[i].idx = -1; out[i].weight = 0.f; }
    }
}

/* NearestNeighborSearchBruteForce::getClosestPoint as written (NearestNeighbor.h:81-97): candidates compared on the rounded
 * Euclidean norm (strict '>' => lowest index among equal norms), the norm compared with m_maxDistance (:93). */
void orc_knn3_brute_norm(const float* tgt, int64_t nt, const float* qry, int64_t nq, float max_d, orc_match* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; ++i) {
        float best = FLT_MAX; int32_t bi = -1;
        for (int64_t j = 0; j < nt; ++j) {
            const float d = sqrtf(d2_3(qry + 3 * i, tgt + 3 * j));
            if (best > d) { best = d; bi = (int32_t)j; }
        }
        if (best <= max_d) { out[i].idx = bi; out[i].weight = 1.f; }
        else { out[i].idx = -1; out[i].weight = 0.f; }
    }
}

void orc_knn6_brute(const float* tgt, const uint8_t* tc, int64_t nt, const float* qry, const uint8_t* qc, int64_t nq,
                    float max_d2, orc_match* out) {
    float* tf = (float*)malloc(sizeof(float) * 6 * (size_t)(nt > 0 ? nt : 1));
    for (int64_t j = 0; j < nt; ++j) feat6(tgt + 3 * j, tc + 4 * j, tf + 6 * j);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; ++i) {
        float qf[6]; feat6(qry + 3 * i, qc + 4 * i, qf);
        float best = FLT_MAX; int32_t bi = -1;
        for (int64_t j = 0; j < nt; ++j) {
            const float d = d2_6(qf, tf + 6 * j);
            if (best > d) { best = d; bi = (int32_t)j; }
        }
        if (bi >= 0 && best <= max_d2) { out[i].idx = bi; out[i].weight = 1.f; }
        else { out[i].idx = -1; out[i].weight = 0.f; }
    }
    free(tf);
}

/* Exact kd-tree: returns exactly what the brute-force scan returns (argmin of (d2, idx)
 * lexicographically).  Pruning uses fl((q-split)^2) > best, which is safe because fp32 rounding is
 * monotone: any point beyond the split plane has fl(dx*dx) >= fl((q-split)^2) and adding
 * non-negative terms cannot decrease the rounded sum. */
struct orc_kdtree {
    int dim; int64_t n;       /* n = number of finite points inserted */
    float* pts;               /* [n*dim], reordered */
    int32_t* orig;            /* [n] original index */
    int32_t* node_lo; int32_t* node_hi; int32_t* node_left; int32_t* node_right; int8_t* node_dim; float* node_split;
    int32_t n_nodes, cap_nodes;
};
#define KD_LEAF 12

static void kd_swap(orc_kdtree* t, int64_t a, int64_t b) {
    if (a == b) return;
    float tmp[6];
    memcpy(tmp, t->pts + a * t->dim, sizeof(float) * t->dim);
    memcpy(t->pts + a * t->dim, t->pts + b * t->dim, sizeof(float) * t->dim);
    memcpy(t->pts + b * t->dim, tmp, sizeof(float) * t->dim);
    int32_t o = t->orig[a]; t->orig[a] = t->orig[b]; t->orig[b] = o;
}
static void kd_select(orc_kdtree* t, int64_t lo, int64_t hi, int64_t k, int d) {
    /* quickselect on coordinate d over [lo,hi) so that element k is in sorted position */
    while (hi - lo > 1) {
        int64_t mid = lo + (hi - lo) / 2;
        float a = t->pts[lo * t->dim + d], b = t->pts[mid * t->dim + d], c = t->pts[(hi - 1) * t->dim + d];
        float pv = (a < b) ? ((b < c) ? b : (a < c ? c : a)) : ((a < c) ? a : (b < c ? c : b));
        int64_t i = lo, j = hi - 1;
        while (i <= j) {
            while (t->pts[i * t->dim + d] < pv) ++i;
            while (t->pts[j * t->dim + d] > pv) --j;
            if (i <= j) { kd_swap(t, i, j); ++i; --j; }
        }
        if (k <= j) hi = j + 1; else if (k >= i) lo = i; else return;
    }
}
static int32_t kd_build_rec(orc_kdtree* t, int64_t lo, int64_t hi) {
    if (t->n_nodes == t->cap_nodes) {
        t->cap_nodes *= 2;
        t->node_lo = realloc(t->node_lo, sizeof(int32_t) * t->cap_nodes); t->node_hi = realloc(t->node_hi, sizeof(int32_t) * t->cap_nodes);
        t->node_left = realloc(t->node_left, sizeof(int32_t) * t->cap_nodes); t->node_right = realloc(t->node_right, sizeof(int32_t) * t->cap_nodes);
        t->node_dim = realloc(t->node_dim, sizeof(int8_t) * t->cap_nodes); t->node_split = realloc(t->node_split, sizeof(float) * t->cap_nodes);
    }
    int32_t id = t->n_nodes++;
    t->node_lo[id] = (int32_t)lo; t->node_hi[id] = (int32_t)hi; t->node_left[id] = t->node_right[id] = -1; t->node_dim[id] = -1; t->node_split[id] = 0.f;
    if (hi - lo <= KD_LEAF) return id;
    int bd = 0; float bext = -1.f;
    for (int d = 0; d < t->dim; ++d) {
        float mn = FLT_MAX, mx = -FLT_MAX;
        for (int64_t i = lo; i < hi; ++i) { float v = t->pts[i * t->dim + d]; if (v < mn) mn = v; if (v > mx) mx = v; }
        if (mx - mn > bext) { bext = mx - mn; bd = d; }
    }
    if (!(bext > 0.f)) return id; /* all points identical: keep as (large) leaf */
    int64_t mid = lo + (hi - lo) / 2;
    kd_select(t, lo, hi, mid, bd);
    float split = t->pts[mid * t->dim + bd];
    t->node_dim[id] = (int8_t)bd; t->node_split[id] = split;
    /* left: [lo,mid) coords <= split ; right: [mid,hi) coords >= split */
    int32_t l = kd_build_rec(t, lo, mid);
    int32_t r = kd_build_rec(t, mid, hi);
    t->node_left[id] = l; t->node_right[id] = r;
    return id;
}

orc_kdtree* orc_kdtree_build(const float* tgt, const uint8_t* tc, int64_t nt) {
    orc_kdtree* t = (orc_kdtree*)calloc(1, sizeof(*t));
    t->dim = tc ? 6 : 3;
    t->pts = (float*)malloc(sizeof(float) * t->dim * (size_t)(nt > 0 ? nt : 1));
    t->orig = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nt > 0 ? nt : 1));
    int64_t m = 0;
    for (int64_t j = 0; j < nt; ++j) {
        if (!finite3(tgt + 3 * j)) continue; /* a non-finite target can never win the strict '>' scan */
        if (tc) feat6(tgt + 3 * j, tc + 4 * j, t->pts + 6 * m);
        else memcpy(t->pts + 3 * m, tgt + 3 * j, 3 * sizeof(float));
        t->orig[m++] = (int32_t)j;
    }
    t->n = m;
    t->cap_nodes = 64;
    t->node_lo = malloc(sizeof(int32_t) * t->cap_nodes); t->node_hi = malloc(sizeof(int32_t) * t->cap_nodes);
    t->node_left = malloc(sizeof(int32_t) * t->cap_nodes); t->node_right = malloc(sizeof(int32_t) * t->cap_nodes);
    t->node_dim = malloc(sizeof(int8_t) * t->cap_nodes); t->node_split = malloc(sizeof(float) * t->cap_nodes);
    if (m > 0) kd_build_rec(t, 0, m);
    return t;
}
void orc_kdtree_free(orc_kdtree* t) {
    if (!t) return;
    free(t->pts); free(t->orig); free(t->node_lo); free(t->node_hi); free(t->node_left); free(t->node_right); free(t->node_dim); free(t->node_split);
    free(t);
}
static void kd_query_rec(const orc_kdtree* t, int32_t id, const float* q, float* best, int32_t* bi) {
    int d = t->node_dim[id];
    if (d < 0) {
        for (int32_t i = t->node_lo[id]; i < t->node_hi[id]; ++i) {
            const float dd = (t->dim == 3) ? d2_3(q, t->pts + 3 * (size_t)i) : d2_6(q, t->pts + 6 * (size_t)i);
            const int32_t o = t->orig[i];
            if (dd < *best || (dd == *best && o < *bi)) { *best = dd; *bi = o; }
        }
        return;
    }
    const float diff = q[d] - t->node_split[id];
    const float pd = diff * diff;
    int32_t nearc = (diff <= 0.f) ? t->node_left[id] : t->node_right[id];
    int32_t farc = (diff <= 0.f) ? t->node_right[id] : t->node_left[id];
    kd_query_rec(t, nearc, q, best, bi);
    if (!(pd > *best)) kd_query_rec(t, farc, q, best, bi);
}
void orc_kdtree_query(const orc_kdtree* t, const float* qry, const uint8_t* qc, int64_t nq, float max_d2, orc_match* out) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < nq; ++i) {
        float qf[6];
        if (t->dim == 6) feat6(qry + 3 * i, qc + 4 * i, qf); else memcpy(qf, qry + 3 * i, 3 * sizeof(float));
        /* start at (FLT_MAX, INT_MAX): brute force starts at FLT_MAX with strict '>' */
        float best = FLT_MAX; int32_t bi = INT32_MAX;
        int ok = 1;
        for (int d = 0; d < 3; ++d) if (!isfinite(qf[d])) ok = 0;
        if (ok && t->n > 0) kd_query_rec(t, 0, qf, &best, &bi);
        if (bi != INT32_MAX && best < FLT_MAX && best <= max_d2) { out[i].idx = bi; out[i].weight = 1.f; }
        else { out[i].idx = -1; out[i].weight = 0.f; }
    }
}

/* ------------------------------------------------------------------ projective (NearestNeighbor.h:333-421) */

static inline uint32_t x86_float_to_u32(float t) {
    /* `unsigned = std::round(float)` (NearestNeighbor.h:378-379) is UB for negative / huge values;
     * x86-64 gcc emits cvttss2si r64 and keeps the low 32 bits: NaN / |t| >= 2^63 -> 0,
     * negatives wrap modulo 2^32. */
    if (!(fabsf(t) < 9223372036854775808.0f)) return 0u;
    return (uint32_t)(int64_t)t;
}

void orc_projective(const float* tgt, uint32_t width, uint32_t height, float fx, float fy, float mx, float my,
                    const float* qry, int64_t nq, float max_d2, orc_match* out) {
    const uint32_t searchWindow = 12; /* NearestNeighbor.h:319 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; ++i) {
        const float* p = qry + 3 * i;
        out[i].idx = 0; out[i].weight = 0.f; /* std::vector<Match> matches(nMatches) value-initialises (:353) */
        if (p[0] == MINF) continue;           /* :372-373 leaves {0, 0.f} */
        uint32_t uPoint = x86_float_to_u32(roundf(((p[0] * fx) / p[2]) + mx));
        uint32_t vPoint = x86_float_to_u32(roundf(((p[1] * fy) / p[2]) + my));
        float minDist = FLT_MAX;
        uint32_t idx = (uint32_t)-1;
        for (uint32_t v = vPoint - searchWindow; (v < height && v <= vPoint + searchWindow); v++) {
            for (uint32_t u = uPoint - searchWindow; (u < width && u <= uPoint + searchWindow); u++) {
                uint32_t neighborIndex = width * v + u;
                const float* t = tgt + 3 * (size_t)neighborIndex;
                if (t[0] == MINF) continue;
                /* squaredNorm of the difference; contract D1 association */
                float dist = d2_3(p, t);
                if (minDist > dist) { idx = neighborIndex; minDist = dist; }
            }
        }
        if (minDist <= max_d2) { out[i].idx = (int32_t)idx; out[i].weight = 1.f; }
        else { out[i].idx = -1; out[i].weight = 0.f; }
    }
}

/* ------------------------------------------------------------------ weighting (weighting.h:39-99) */

void orc_apply_weights(int method, float max_d2, const float* sp, const float* tp, const float* sn, const float* tn,
                       const uint8_t* sc, const uint8_t* tc, int64_t n, orc_match* m) {
    if (method == ORC_WEIGHT_CONSTANT) return; /* :44 */
    for (int64_t i = 0; i < n; ++i) {
        if (m[i].idx < 0) continue;
        const int64_t j = m[i].idx;
        float w = 0.0f;
        if (method == ORC_WEIGHT_DISTANCES || method == ORC_WEIGHT_COLORS) {
            if (!finite3(sp + 3 * i) || !finite3(tp + 3 * j)) w += 0.0f;
            else {
                /* :16-20: 1.0 - ((d0*d0 + d1*d1 + d2*d2) / maxDistance): float quotient, double subtraction, float result */
                const float d0 = sp[3 * i] - tp[3 * j], d1 = sp[3 * i + 1] - tp[3 * j + 1], d2 = sp[3 * i + 2] - tp[3 * j + 2];
                const float q = ((d0 * d0 + d1 * d1) + d2 * d2) / max_d2;
                w += (float)(1.0 - (double)q);
            }
        }
        if (method == ORC_WEIGHT_NORMALS) {
            if (!finite3(sn + 3 * i) || !finite3(tn + 3 * j)) w += 0.0f;
            else w += (sn[3 * i] * tn[3 * j] + sn[3 * i + 1] * tn[3 * j + 1]) + sn[3 * i + 2] * tn[3 * j + 2]; /* :22-25 */
        }
        if (method == ORC_WEIGHT_COLORS) {
            /* :27-30: Vector4uc difference wraps modulo 256 before squaring (ints), / 195075 in float */
            const uint8_t e0 = (uint8_t)(sc[4 * i] - tc[4 * j]), e1 = (uint8_t)(sc[4 * i + 1] - tc[4 * j + 1]), e2 = (uint8_t)(sc[4 * i + 2] - tc[4 * j + 2]);
            const int s = (int)e0 * e0 + (int)e1 * e1 + (int)e2 * e2;
            const float cw = (float)(1.0 - (double)((float)s / (float)195075));
            w *= cw;
        }
        m[i].weight = w;
    }
}

/* ------------------------------------------------------------------ rejection (ICPOptimizer.h:157-174) */

void orc_prune(const float* sn, const float* tn, int64_t n, orc_match* m) {
    const double threshold = 60 * 3.141592653589793238462643383279502884 / 180.0; /* EIGEN_PI */
    for (int64_t i = 0; i < n; ++i) {
        if (m[i].idx < 0) continue;
        const float* a = sn + 3 * i; const float* b = tn + 3 * (int64_t)m[i].idx;
        const float dot = (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
        const float na = sqrtf((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]);
        const float nb = sqrtf((b[0] * b[0] + b[1] * b[1]) + b[2] * b[2]);
        /* acos(float) resolves to the float overload; the comparison is in double. NaN compares false => kept. */
        if ((double)acosf(dot / (na * nb)) > threshold) m[i].idx = -1;
    }
}

/* ------------------------------------------------------------------ small dense linear algebra (fp64) */

static int solve6(double A[36] /* row-major, destroyed */, double b[6], double x[6]) {
    /* Gaussian elimination with partial pivoting */
    for (int k = 0; k < 6; ++k) {
        int p = k; double mx = fabs(A[k * 6 + k]);
        for (int i = k + 1; i < 6; ++i) if (fabs(A[i * 6 + k]) > mx) { mx = fabs(A[i * 6 + k]); p = i; }
        if (!(mx > 0.0)) return -1;
        if (p != k) { for (int j = 0; j < 6; ++j) { double t = A[k * 6 + j]; A[k * 6 + j] = A[p * 6 + j]; A[p * 6 + j] = t; } double t = b[k]; b[k] = b[p]; b[p] = t; }
        for (int i = k + 1; i < 6; ++i) {
            double f = A[i * 6 + k] / A[k * 6 + k];
            for (int j = k; j < 6; ++j) A[i * 6 + j] -= f * A[k * 6 + j];
            b[i] -= f * b[k];
        }
    }
    for (int i = 5; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < 6; ++j) s -= A[i * 6 + j] * x[j];
        x[i] = s / A[i * 6 + i];
    }
    return 0;
}

static void svd3(const double A[9] /* row-major */, double U[9], double S[3], double V[9]) {
    /* One-sided Jacobi (Hestenes) SVD, singular values sorted descending, A = U diag(S) V^T. */
    double W[9]; memcpy(W, A, sizeof(W));
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
            double a = 0, b = 0, c = 0;
            for (int k = 0; k < 3; ++k) { a += W[k * 3 + p] * W[k * 3 + p]; b += W[k * 3 + q] * W[k * 3 + q]; c += W[k * 3 + p] * W[k * 3 + q]; }
            if (fabs(c) <= 1e-300 || fabs(c) <= 1e-17 * sqrt(a * b)) continue;
            off += fabs(c);
            double zeta = (b - a) / (2.0 * c);
            double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
            for (int k = 0; k < 3; ++k) {
                double wp = W[k * 3 + p], wq = W[k * 3 + q];
                W[k * 3 + p] = cs * wp - sn * wq; W[k * 3 + q] = sn * wp + cs * wq;
                double vp = V[k * 3 + p], vq = V[k * 3 + q];
                V[k * 3 + p] = cs * vp - sn * vq; V[k * 3 + q] = sn * vp + cs * vq;
            }
        }
        if (off == 0.0) break;
    }
    for (int j = 0; j < 3; ++j) S[j] = sqrt(W[j] * W[j] + W[3 + j] * W[3 + j] + W[6 + j] * W[6 + j]);
    /* sort descending */
    int ord[3] = {0, 1, 2};
    for (int i = 0; i < 2; ++i) for (int j = i + 1; j < 3; ++j) if (S[ord[j]] > S[ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    double Ws[9], Vs[9], Ss[3];
    for (int j = 0; j < 3; ++j) { Ss[j] = S[ord[j]]; for (int k = 0; k < 3; ++k) { Ws[k * 3 + j] = W[k * 3 + ord[j]]; Vs[k * 3 + j] = V[k * 3 + ord[j]]; } }
    memcpy(S, Ss, sizeof(Ss)); memcpy(V, Vs, sizeof(Vs));
    const double tiny = 1e-14 * (S[0] > 0 ? S[0] : 1.0);
    for (int j = 0; j < 3; ++j) {
        if (S[j] > tiny) for (int k = 0; k < 3; ++k) U[k * 3 + j] = Ws[k * 3 + j] / S[j];
        else for (int k = 0; k < 3; ++k) U[k * 3 + j] = 0.0;
    }
    /* complete rank-deficient U to an orthonormal basis */
    if (!(S[0] > tiny)) { for (int i = 0; i < 9; ++i) U[i] = (i % 4 == 0) ? 1.0 : 0.0; return; }
    if (!(S[1] > tiny)) {
        double u0[3] = {U[0], U[3], U[6]};
        int m = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2) : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
        double e[3] = {0, 0, 0}; e[m] = 1.0;
        double u1[3] = {u0[1] * e[2] - u0[2] * e[1], u0[2] * e[0] - u0[0] * e[2], u0[0] * e[1] - u0[1] * e[0]};
        double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
        for (int k = 0; k < 3; ++k) U[k * 3 + 1] = u1[k] / n1;
    }
    if (!(S[2] > tiny)) {
        double u0[3] = {U[0], U[3], U[6]}, u1[3] = {U[1], U[4], U[7]};
        U[2] = u0[1] * u1[2] - u0[2] * u1[1]; U[5] = u0[2] * u1[0] - u0[0] * u1[2]; U[8] = u0[0] * u1[1] - u0[1] * u1[0];
    }
}

static double det3(const double M[9]) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

/* ------------------------------------------------------------------ linear solvers */

int orc_solve_p2p(const float* s, const float* d, const float* w, int64_t m, float out[16]) {
    /* ICPOptimizer.h:666-674 -> ProcrustesAligner::estimatePose (ProcrustesAligner.h:6-70).
     * Unweighted means; A = sum (d - dbar) (w (s - sbar))^T; R = U diag(1,1,det(U V^T)) V^T;
     * translation column = R*(dbar - sbar) - R*dbar + dbar.  Contract D5: fp64 throughout. */
    mat4_identity(out);
    if (m <= 0) return -1;
    double sm[3] = {0, 0, 0}, dm[3] = {0, 0, 0};
    for (int64_t i = 0; i < m; ++i) for (int k = 0; k < 3; ++k) { sm[k] += s[3 * i + k]; dm[k] += d[3 * i + k]; }
    for (int k = 0; k < 3; ++k) { sm[k] /= (double)m; dm[k] /= (double)m; }
    double A[9] = {0};
    for (int64_t i = 0; i < m; ++i) {
        double sv[3], dv[3];
        for (int k = 0; k < 3; ++k) { sv[k] = (double)w[i] * ((double)s[3 * i + k] - sm[k]); dv[k] = (double)d[3 * i + k] - dm[k]; }
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) A[r * 3 + c] += dv[r] * sv[c];
    }
    double U[9], S[3], V[9];
    svd3(A, U, S, V);
    double UVt[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) UVt[r * 3 + c] = U[r * 3] * V[c * 3] + U[r * 3 + 1] * V[c * 3 + 1] + U[r * 3 + 2] * V[c * 3 + 2];
    const double dd = det3(UVt);
    double R[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r * 3 + c] = U[r * 3] * V[c * 3] + U[r * 3 + 1] * V[c * 3 + 1] + dd * U[r * 3 + 2] * V[c * 3 + 2];
    for (int r = 0; r < 3; ++r) {
        double rt = 0, rd = 0;
        for (int c = 0; c < 3; ++c) { rt += R[r * 3 + c] * (dm[c] - sm[c]); rd += R[r * 3 + c] * dm[c]; }
        out[r + 12] = (float)(rt - rd + dm[r]);
        for (int c = 0; c < 3; ++c) out[r + 4 * c] = (float)R[r * 3 + c];
    }
    return 0;
}

static void add_row(double AtA[36], double Atb[6], const double row[6], double rhs, double scale) {
    double r[6];
    for (int i = 0; i < 6; ++i) r[i] = row[i] * scale;
    rhs *= scale;
    for (int i = 0; i < 6; ++i) { Atb[i] += r[i] * rhs; for (int j = 0; j < 6; ++j) AtA[i * 6 + j] += r[i] * r[j]; }
}

int orc_solve_p2plane(const float* s_, const float* d_, const float* n_, const float* w, int64_t m, float out[16]) {
    /* ICPOptimizer.h:676-782. Rows exactly as the reference builds them (plane row, three point rows,
     * scaled by LAMBDA*weight); the least-squares solution of A x = b is obtained from the fp64
     * normal equations (== JacobiSVD::solve for full column rank).  A non-finite target normal
     * drops only the plane row (the rule CeresICPOptimizer applies, ICPOptimizer.h:420-423);
     * the linear reference would produce a NaN pose there -- documented deviation. */
    mat4_identity(out);
    if (m <= 0) return -1;
    const double LAMBDA_POINT = (double)0.1f, LAMBDA_PLANE = (double)1.0f;
    double AtA[36] = {0}, Atb[6] = {0};
    for (int64_t i = 0; i < m; ++i) {
        const double s[3] = {s_[3 * i], s_[3 * i + 1], s_[3 * i + 2]}, d[3] = {d_[3 * i], d_[3 * i + 1], d_[3 * i + 2]};
        const double wt = (double)w[i];
        if (finite3(n_ + 3 * i)) {
            const double n[3] = {n_[3 * i], n_[3 * i + 1], n_[3 * i + 2]};
            const double row[6] = {n[2] * s[1] - n[1] * s[2], n[0] * s[2] - n[2] * s[0], n[1] * s[0] - n[0] * s[1], n[0], n[1], n[2]};
            const double rhs = (n[0] * d[0] + n[1] * d[1] + n[2] * d[2]) - (n[0] * s[0] + n[1] * s[1] + n[2] * s[2]);
            add_row(AtA, Atb, row, rhs, LAMBDA_PLANE * wt);
        }
        const double r1[6] = {0, s[2], -s[1], 1, 0, 0}, r2[6] = {-s[2], 0, s[0], 0, 1, 0}, r3[6] = {s[1], -s[0], 0, 0, 0, 1};
        add_row(AtA, Atb, r1, d[0] - s[0], LAMBDA_POINT * wt);
        add_row(AtA, Atb, r2, d[1] - s[1], LAMBDA_POINT * wt);
        add_row(AtA, Atb, r3, d[2] - s[2], LAMBDA_POINT * wt);
    }
    double x[6];
    if (solve6(AtA, Atb, x) != 0) return -2;
    /* :768-779: R = Rx(alpha) * Ry(beta) * Rz(gamma), fp32 */
    const float al = (float)x[0], be = (float)x[1], ga = (float)x[2];
    const float ca = (float)cos((double)al), sa = (float)sin((double)al), cb = (float)cos((double)be), sb = (float)sin((double)be), cg = (float)cos((double)ga), sg = (float)sin((double)ga);
    float Rx[16], Ry[16], Rz[16], T[16];
    mat4_identity(Rx); mat4_identity(Ry); mat4_identity(Rz);
    Rx[5] = ca; Rx[9] = -sa; Rx[6] = sa; Rx[10] = ca;
    Ry[0] = cb; Ry[8] = sb; Ry[2] = -sb; Ry[10] = cb;
    Rz[0] = cg; Rz[4] = -sg; Rz[1] = sg; Rz[5] = cg;
    mat4_mul(Rx, Ry, T); mat4_mul(T, Rz, out);
    out[12] = (float)x[3]; out[13] = (float)x[4]; out[14] = (float)x[5];
    return 0;
}

static void translation4(const float t[3], float M[16]) { mat4_identity(M); M[12] = t[0]; M[13] = t[1]; M[14] = t[2]; }

int orc_solve_symmetric(const float* s_, const float* d_, const float* ns_, const float* nt_, const float* w, int64_t m, float out[16]) {
    /* ICPOptimizer.h:784-898. */
    mat4_identity(out);
    if (m <= 0) return -1;
    const double LAMBDA_POINT = (double)0.1f, LAMBDA_SYMMETRIC = (double)1.0f;
    double sm[3] = {0, 0, 0}, dm[3] = {0, 0, 0};
    for (int64_t i = 0; i < m; ++i) for (int k = 0; k < 3; ++k) { sm[k] += s_[3 * i + k]; dm[k] += d_[3 * i + k]; }
    float meanS[3], meanT[3];
    for (int k = 0; k < 3; ++k) { meanS[k] = (float)(sm[k] / (double)m); meanT[k] = (float)(dm[k] / (double)m); }
    double AtA[36] = {0}, Atb[6] = {0};
    for (int64_t i = 0; i < m; ++i) {
        double s[3], d[3];
        for (int k = 0; k < 3; ++k) { s[k] = (double)s_[3 * i + k] - (double)meanS[k]; d[k] = (double)d_[3 * i + k] - (double)meanT[k]; }
        const double wt = (double)w[i];
        if (finite3(nt_ + 3 * i) && finite3(ns_ + 3 * i)) {
            double nsum[3], u[3];
            for (int k = 0; k < 3; ++k) { nsum[k] = (double)nt_[3 * i + k] + (double)ns_[3 * i + k]; u[k] = s[k] + d[k]; }
            const double row[6] = {u[1] * nsum[2] - u[2] * nsum[1], u[2] * nsum[0] - u[0] * nsum[2], u[0] * nsum[1] - u[1] * nsum[0], nsum[0], nsum[1], nsum[2]};
            const double rhs = (d[0] - s[0]) * nsum[0] + (d[1] - s[1]) * nsum[1] + (d[2] - s[2]) * nsum[2];
            add_row(AtA, Atb, row, rhs, LAMBDA_SYMMETRIC * wt);
        }
        const double r1[6] = {0, s[2], -s[1], 1, 0, 0}, r2[6] = {-s[2], 0, s[0], 0, 1, 0}, r3[6] = {s[1], -s[0], 0, 0, 0, 1};
        add_row(AtA, Atb, r1, d[0] - s[0], LAMBDA_POINT * wt);
        add_row(AtA, Atb, r2, d[1] - s[1], LAMBDA_POINT * wt);
        add_row(AtA, Atb, r3, d[2] - s[2], LAMBDA_POINT * wt);
    }
    const float lambda = 0.0001f;
    for (int i = 0; i < 6; ++i) AtA[i * 6 + i] += (double)(lambda * lambda); /* :863-864 */
    double x[6];
    if (solve6(AtA, Atb, x) != 0) return -2;
    /* :876-895 (fp32 scalars as in the reference; trig-free) */
    const float a_t[3] = {(float)x[0], (float)x[1], (float)x[2]}, t_t[3] = {(float)x[3], (float)x[4], (float)x[5]};
    const float tan_theta = sqrtf((a_t[0] * a_t[0] + a_t[1] * a_t[1]) + a_t[2] * a_t[2]);
    float R4[16]; mat4_identity(R4);
    float cos_theta = 1.0f;
    if (tan_theta > 0.f) { /* reference divides 0/0 here -> NaN pose; guarded (documented deviation) */
        const float a[3] = {a_t[0] / tan_theta, a_t[1] / tan_theta, a_t[2] / tan_theta};
        const float sin_theta = (float)((double)tan_theta / sqrt(1.0 + (double)(tan_theta * tan_theta)));
        cos_theta = sin_theta / tan_theta;
        /* getRodriguesMatrix (utils.h:171-176): I + sin*K + (1-cos)*K*K */
        const float K[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0}; /* row-major */
        const float omc = 1 - cos_theta;
        float K1[9], KK[9];
        for (int i = 0; i < 9; ++i) K1[i] = omc * K[i];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) KK[r * 3 + c] = (K1[r * 3] * K[c] + K1[r * 3 + 1] * K[3 + c]) + K1[r * 3 + 2] * K[6 + c];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R4[r + 4 * c] = ((r == c) ? 1.0f : 0.0f) + (sin_theta * K[r * 3 + c] + KK[r * 3 + c]);
    }
    const float t[3] = {t_t[0] * cos_theta, t_t[1] * cos_theta, t_t[2] * cos_theta};
    const float negS[3] = {-meanS[0], -meanS[1], -meanS[2]};
    float Td[16], Tt[16], Ts[16], M1[16], M2[16], M3[16];
    translation4(meanT, Td); translation4(t, Tt); translation4(negS, Ts);
    mat4_mul(Td, R4, M1); mat4_mul(M1, Tt, M2); mat4_mul(M2, R4, M3); mat4_mul(M3, Ts, out);
    return 0;
}

/* ------------------------------------------------------------------ Ceres restatement */

typedef struct { double v; double d[6]; } jet;
static inline jet jc(double v) { jet r; r.v = v; memset(r.d, 0, sizeof(r.d)); return r; }
static inline jet jadd(jet a, jet b) { jet r; r.v = a.v + b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
static inline jet jsub(jet a, jet b) { jet r; r.v = a.v - b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
static inline jet jmul(jet a, jet b) { jet r; r.v = a.v * b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
static inline jet jdiv(jet a, jet b) { jet r; const double inv = 1.0 / b.v; r.v = a.v * inv; for (int i = 0; i < 6; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv; return r; }
static inline jet jsqrt(jet a) { jet r; r.v = sqrt(a.v); const double k = 1.0 / (2.0 * r.v); for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * k; return r; }
static inline jet jcos(jet a) { jet r; r.v = cos(a.v); const double k = -sin(a.v); for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * k; return r; }
static inline jet jsin(jet a) { jet r; r.v = sin(a.v); const double k = cos(a.v); for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * k; return r; }
static inline jet jneg(jet a) { jet r; r.v = -a.v; for (int i = 0; i < 6; ++i) r.d[i] = -a.d[i]; return r; }

static void jet_angle_axis_rotate_point(const jet aa[3], const jet pt[3], jet result[3]) {
    /* ceres/rotation.h AngleAxisRotatePoint (Ceres 2.x, restated from its published algorithm) */
    const jet theta2 = jadd(jadd(jmul(aa[0], aa[0]), jmul(aa[1], aa[1])), jmul(aa[2], aa[2]));
    if (theta2.v > DBL_EPSILON) {
        const jet theta = jsqrt(theta2);
        const jet costheta = jcos(theta), sintheta = jsin(theta);
        const jet theta_inverse = jdiv(jc(1.0), theta);
        const jet w[3] = {jmul(aa[0], theta_inverse), jmul(aa[1], theta_inverse), jmul(aa[2], theta_inverse)};
        const jet wxp[3] = {jsub(jmul(w[1], pt[2]), jmul(w[2], pt[1])), jsub(jmul(w[2], pt[0]), jmul(w[0], pt[2])), jsub(jmul(w[0], pt[1]), jmul(w[1], pt[0]))};
        const jet tmp = jmul(jadd(jadd(jmul(w[0], pt[0]), jmul(w[1], pt[1])), jmul(w[2], pt[2])), jsub(jc(1.0), costheta));
        for (int i = 0; i < 3; ++i) result[i] = jadd(jadd(jmul(pt[i], costheta), jmul(wxp[i], sintheta)), jmul(w[i], tmp));
    } else {
        const jet wxp[3] = {jsub(jmul(aa[1], pt[2]), jmul(aa[2], pt[1])), jsub(jmul(aa[2], pt[0]), jmul(aa[0], pt[2])), jsub(jmul(aa[0], pt[1]), jmul(aa[1], pt[0]))};
        for (int i = 0; i < 3; ++i) result[i] = jadd(pt[i], wxp[i]);
    }
}

typedef struct {
    int metric; const float* sp; const float* sn; const float* tgt; const float* tgt_n; const orc_match* m; int64_t n;
} lm_problem;

/* Evaluate cost = 1/2 sum r^2 and (optionally) J^T J (row-major 6x6), J^T r at x.
 * Residual blocks as ICPOptimizer.h:362-482 adds them; functors from constraints.h. */
static int64_t lm_eval(const lm_problem* P, const double x[6], double* cost, double* JtJ, double* Jtr) {
    double c = 0.0; int64_t nres = 0;
    if (JtJ) { memset(JtJ, 0, 36 * sizeof(double)); memset(Jtr, 0, 6 * sizeof(double)); }
    jet pose[6];
    for (int i = 0; i < 6; ++i) { pose[i] = jc(x[i]); pose[i].d[i] = 1.0; }
    jet ninv[3] = {jneg(pose[0]), jneg(pose[1]), jneg(pose[2])};
    for (int64_t i = 0; i < P->n; ++i) {
        if (P->m[i].idx < 0) continue;
        const float* s = P->sp + 3 * i; const float* d = P->tgt + 3 * (int64_t)P->m[i].idx;
        if (!finite3(s) || !finite3(d)) continue;
        const float wgt = P->m[i].weight;
        jet sj[3] = {jc(s[0]), jc(s[1]), jc(s[2])}, y[3], rot[3];
        jet_angle_axis_rotate_point(pose, sj, rot);           /* PoseIncrement::apply, utils.h:44-56 */
        for (int k = 0; k < 3; ++k) y[k] = jadd(rot[k], pose[3 + k]);
        jet r[4]; int nr = 0;
        const jet lw_pt = jmul(jc((double)0.1f), jc((double)wgt)); /* PointToPointConstraint LAMBDA = 0.1f */
        for (int k = 0; k < 3; ++k) r[nr++] = jmul(lw_pt, jsub(y[k], jc(d[k])));
        if (P->metric == ORC_METRIC_P2PLANE) {
            const float* n = P->tgt_n + 3 * (int64_t)P->m[i].idx;
            if (finite3(n)) {
                jet acc = jadd(jadd(jmul(jc(n[0]), jsub(y[0], jc(d[0]))), jmul(jc(n[1]), jsub(y[1], jc(d[1])))), jmul(jc(n[2]), jsub(y[2], jc(d[2]))));
                r[nr++] = jmul(jmul(jc((double)1.0f), jc((double)wgt)), acc);
            }
        } else if (P->metric == ORC_METRIC_SYMMETRIC) {
            const float* n = P->tgt_n + 3 * (int64_t)P->m[i].idx; const float* ns = P->sn + 3 * i;
            if (finite3(n) && finite3(ns)) {
                jet dj[3] = {jc(d[0]), jc(d[1]), jc(d[2])}, z[3];
                jet_angle_axis_rotate_point(ninv, dj, z);      /* apply_inv_rotation, utils.h:60-72 */
                jet acc = jc(0.0);
                jet comp[3];
                for (int k = 0; k < 3; ++k) comp[k] = jmul(jc((double)n[k] + (double)ns[k]), jsub(y[k], z[k]));
                acc = jadd(jadd(comp[0], comp[1]), comp[2]);
                r[nr++] = jmul(jmul(jc((double)1.0f), jc((double)wgt)), acc);
            }
        }
        for (int k = 0; k < nr; ++k) {
            c += r[k].v * r[k].v;
            if (JtJ) for (int a = 0; a < 6; ++a) { Jtr[a] += r[k].d[a] * r[k].v; for (int b = 0; b < 6; ++b) JtJ[a * 6 + b] += r[k].d[a] * r[k].d[b]; }
        }
        nres += nr;
    }
    *cost = 0.5 * c;
    return nres;
}

static void angle_axis_to_rotation(const double aa[3], double R[9] /* column-major like Ceres */) {
    /* ceres/rotation.h AngleAxisToRotationMatrix */
    const double theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
    if (theta2 > DBL_EPSILON) {
        const double theta = sqrt(theta2);
        const double wx = aa[0] / theta, wy = aa[1] / theta, wz = aa[2] / theta;
        const double ct = cos(theta), st = sin(theta);
        R[0] = ct + wx * wx * (1.0 - ct);      R[1] = wz * st + wx * wy * (1.0 - ct);  R[2] = -wy * st + wx * wz * (1.0 - ct);
        R[3] = wx * wy * (1.0 - ct) - wz * st; R[4] = ct + wy * wy * (1.0 - ct);       R[5] = wx * st + wy * wz * (1.0 - ct);
        R[6] = wy * st + wx * wz * (1.0 - ct); R[7] = -wx * st + wy * wz * (1.0 - ct); R[8] = ct + wz * wz * (1.0 - ct);
    } else {
        R[0] = 1; R[1] = aa[2]; R[2] = -aa[1]; R[3] = -aa[2]; R[4] = 1; R[5] = aa[0]; R[6] = aa[1]; R[7] = -aa[0]; R[8] = 1;
    }
}

int orc_solve_lm(int metric, const float* sp, const float* sn, const float* tgt, const float* tgt_n,
                 const orc_match* m, int64_t n, int max_iterations, double x_out[6], float out_pose[16], int* n_lm) {
    /* ceres::Solve with the options of ICPOptimizer.h:352-360 (LEVENBERG_MARQUARDT, monotonic,
     * DENSE_QR, max_num_iterations 10, defaults otherwise), restated from Ceres 2.x's published
     * TrustRegionMinimizer / LevenbergMarquardtStrategy (Ceres is an un-vendored dependency). */
    lm_problem P = {metric, sp, sn, tgt, tgt_n, m, n};
    double x[6] = {0, 0, 0, 0, 0, 0}; /* poseIncrement.setZero(), ICPOptimizer.h:236,310 */
    mat4_identity(out_pose);
    if (n_lm) *n_lm = 0;
    double cost, H[36], g[6];
    if (lm_eval(&P, x, &cost, H, g) == 0) return -1;
    double scale[6];
    for (int i = 0; i < 6; ++i) scale[i] = 1.0 / (1.0 + sqrt(H[i * 6 + i])); /* jacobi_scaling */
    double radius = 1e4, decrease_factor = 2.0;                               /* initial_trust_region_radius */
    const double max_radius = 1e16, min_radius = 1e-32, min_diag = 1e-6, max_diag = 1e32;
    const double min_relative_decrease = 1e-3, function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    double diag[6]; int reuse_diagonal = 0; int invalid_steps = 0;
    double gmax = 0; for (int i = 0; i < 6; ++i) if (fabs(g[i]) > gmax) gmax = fabs(g[i]);
    int iter = 0; int step_successful = 1;
    for (;;) {
        /* FinalizeIterationAndCheckIfMinimizerCanContinue */
        if (iter >= max_iterations) break;
        if (step_successful && gmax <= gradient_tolerance) break;
        if (radius <= min_radius) break;
        ++iter;
        /* LevenbergMarquardtStrategy::ComputeStep in the Jacobi-scaled space */
        double Hs[36], gs[6];
        for (int a = 0; a < 6; ++a) { gs[a] = g[a] * scale[a]; for (int b = 0; b < 6; ++b) Hs[a * 6 + b] = H[a * 6 + b] * scale[a] * scale[b]; }
        if (!reuse_diagonal) for (int a = 0; a < 6; ++a) { double v = Hs[a * 6 + a]; diag[a] = v < min_diag ? min_diag : (v > max_diag ? max_diag : v); }
        double A[36], b[6], ds[6];
        memcpy(A, Hs, sizeof(A));
        for (int a = 0; a < 6; ++a) { A[a * 6 + a] += diag[a] / radius; b[a] = -gs[a]; }
        int lin_ok = solve6(A, b, ds) == 0;
        double model_cost_change = 0.0;
        if (lin_ok) {
            for (int a = 0; a < 6; ++a) { double hd = 0; for (int c = 0; c < 6; ++c) hd += Hs[a * 6 + c] * ds[c]; model_cost_change -= ds[a] * (gs[a] + 0.5 * hd); }
        }
        if (!lin_ok || !(model_cost_change > 0.0)) {
            /* HandleInvalidStep / StepIsInvalid */
            if (++invalid_steps >= 5) break;
            radius *= 0.5; reuse_diagonal = 1; step_successful = 0;
            continue;
        }
        invalid_steps = 0;
        double delta[6], cand[6], cand_cost;
        for (int a = 0; a < 6; ++a) { delta[a] = ds[a] * scale[a]; cand[a] = x[a] + delta[a]; }
        lm_eval(&P, cand, &cand_cost, NULL, NULL);
        /* ParameterToleranceReached */
        double step_norm = 0, x_norm = 0;
        for (int a = 0; a < 6; ++a) { step_norm += (x[a] - cand[a]) * (x[a] - cand[a]); x_norm += x[a] * x[a]; }
        step_norm = sqrt(step_norm); x_norm = sqrt(x_norm);
        if (step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) break;
        /* FunctionToleranceReached */
        const double cost_change = cost - cand_cost;
        if (fabs(cost_change) <= function_tolerance * cost) break;
        const double rho = cost_change / model_cost_change;
        if (rho > min_relative_decrease) {
            memcpy(x, cand, sizeof(x));
            lm_eval(&P, x, &cost, H, g);
            gmax = 0; for (int i = 0; i < 6; ++i) if (fabs(g[i]) > gmax) gmax = fabs(g[i]);
            const double t = 2.0 * rho - 1.0;
            double denom = 1.0 - t * t * t; if (denom < 1.0 / 3.0) denom = 1.0 / 3.0;
            radius = radius / denom; if (radius > max_radius) radius = max_radius;
            decrease_factor = 2.0; reuse_diagonal = 0; step_successful = 1;
        } else {
            radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1; step_successful = 0;
        }
    }
    if (n_lm) *n_lm = iter;
    memcpy(x_out, x, sizeof(x));
    /* PoseIncrement<double>::convertToMatrix (utils.h:79-98) */
    double R[9]; angle_axis_to_rotation(x, R);
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) out_pose[r + 4 * c] = (float)R[r + 3 * c];
    out_pose[12] = (float)x[3]; out_pose[13] = (float)x[4]; out_pose[14] = (float)x[5];
    return 0;
}

/* ------------------------------------------------------------------ selection / pyramid */

void orc_mt_seed(orc_mt19937* r, uint32_t seed) {
    r->mt[0] = seed;
    for (int i = 1; i < 624; ++i) r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
    r->idx = 624;
}
uint32_t orc_mt_next(orc_mt19937* r) {
    if (r->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (r->mt[i] & 0x80000000u) | (r->mt[(i + 1) % 624] & 0x7fffffffu);
            r->mt[i] = r->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        r->idx = 0;
    }
    uint32_t y = r->mt[r->idx++];
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
    return y;
}
double orc_mt_canonical(orc_mt19937* r) {
    /* libstdc++ std::generate_canonical<double, 53>(mt19937): two 32-bit draws, (lo + hi*2^32) / 2^64,
     * clamped below 1 -- what uniform_real_distribution<double>(0,1) evaluates (selection.h:89,96). */
    const double lo = (double)orc_mt_next(r);
    const double hi = (double)orc_mt_next(r);
    double v = (lo + hi * 4294967296.0) / 18446744073709551616.0;
    if (v >= 1.0) v = nextafter(1.0, 0.0);
    return v;
}

int64_t orc_coarse_indices(const float* pts, const float* nrm, int64_t n, int stride, int32_t* out) {
    /* PointCloud.h:325-343 */
    int64_t c = 0;
    for (int64_t i = 0; i < n; i += stride)
        if (finite3(pts + 3 * i) && finite3(nrm + 3 * i)) out[c++] = (int32_t)i;
    return c;
}

int orc_coarsest_stride(int64_t n) {
    /* ICPOptimizer.h:503-516 */
    float currentResolution = 1.0f;
    int originalSize = (int)n;
    for (;;) {
        originalSize = (int)(originalSize / 2.0);
        if (originalSize < 100) break; /* MULTI_RESOLUTION_MINIMUM_POINTS */
        currentResolution *= 2.0f;
    }
    return (int)currentResolution;
}

/* ------------------------------------------------------------------ voxel pyramid levels (extension; restates csrc/grid.cu) */

typedef struct { float o[3], h[3], inv_h[3]; int bits[3]; int T; int axis[64]; } orc_grid;

static int orc_pick_T(int64_t n) {
    int T = 3;
    while (T < 24 && ((int64_t)1 << T) < 4 * (n > 0 ? n : 1)) ++T;   /* ICP_SOURCE_CELLS_PER_POINT */
    return T;
}

static void orc_grid_params(const float* pts, int64_t n, int T, orc_grid* g) {
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}; int any = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!finite3(pts + 3 * i)) continue;
        for (int a = 0; a < 3; ++a) {
            const float v = pts[3 * i + a];
            if (!any || v < lo[a]) lo[a] = v;
            if (!any || v > hi[a]) hi[a] = v;
        }
        any = 1;
    }
    float e[3], cur[3];
    for (int a = 0; a < 3; ++a) {
        g->o[a] = lo[a];
        const float maxabs = fmaxf(fabsf(lo[a]), fabsf(hi[a]));
        const float floor_ = 1e-20f + 1e-6f * maxabs;
        const float ext = hi[a] - lo[a];
        e[a] = ext > floor_ ? ext : floor_;
        cur[a] = e[a]; g->bits[a] = 0;
    }
    for (int k = 0; k < T; ++k) {
        int ax = -1; float best = -1.f;
        for (int c = 0; c < 3; ++c) if (g->bits[c] < 10 && cur[c] > best) { best = cur[c]; ax = c; }
        if (ax < 0) ax = 0;
        g->axis[k] = ax; g->bits[ax] += 1; cur[ax] = cur[ax] * 0.5f;
    }
    for (int a = 0; a < 3; ++a) {
        g->h[a] = (e[a] * 1.00001f) / (float)(1 << g->bits[a]);
        g->inv_h[a] = 1.0f / g->h[a];
    }
    g->T = T;
}

static uint32_t orc_cell_code(const orc_grid* g, const float* p) {
    int c[3], r[3];
    for (int a = 0; a < 3; ++a) {
        const float u = (p[a] - g->o[a]) * g->inv_h[a];
        int i = (int)floorf(u);
        const int hi = (1 << g->bits[a]) - 1;
        c[a] = i < 0 ? 0 : (i > hi ? hi : i);
        r[a] = g->bits[a];
    }
    uint32_t code = 0;
    for (int k = 0; k < g->T; ++k) {
        const int a = g->axis[k];
        r[a] -= 1;
        code = (code << 1) | (uint32_t)((c[a] >> r[a]) & 1);
    }
    return code;
}

typedef struct { uint32_t prefix; int32_t idx; } orc_vox;
static int orc_vox_cmp(const void* a, const void* b) {
    const orc_vox* x = (const orc_vox*)a; const orc_vox* y = (const orc_vox*)b;
    if (x->prefix != y->prefix) return x->prefix < y->prefix ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}
static int orc_i32_cmp(const void* a, const void* b) { const int32_t x = *(const int32_t*)a, y = *(const int32_t*)b; return x < y ? -1 : (x > y ? 1 : 0); }

int64_t orc_voxel_indices(const float* pts, const float* nrm, int64_t n, int stride, int32_t* out) {
    if (stride <= 1) return orc_coarse_indices(pts, nrm, n, 1, out);
    const int T = orc_pick_T(n);
    orc_grid g; orc_grid_params(pts, n, T, &g);
    int k = 0; while ((1 << k) < stride) ++k;
    int D = (T < 22 ? T : 22) - (3 * k + 1) / 2; if (D < 0) D = 0;
    orc_vox* v = (orc_vox*)malloc(sizeof(orc_vox) * (size_t)(n > 0 ? n : 1));
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!finite3(pts + 3 * i) || !finite3(nrm + 3 * i)) continue;
        v[m].prefix = orc_cell_code(&g, pts + 3 * i) >> (T - D); v[m].idx = (int32_t)i; ++m;
    }
    qsort(v, (size_t)m, sizeof(orc_vox), orc_vox_cmp);
    int64_t c = 0;
    for (int64_t j = 0; j < m; ++j) if (j == 0 || v[j].prefix != v[j - 1].prefix) out[c++] = v[j].idx;   /* lowest index of the cell */
    free(v);
    qsort(out, (size_t)c, sizeof(int32_t), orc_i32_cmp);
    return c;
}

/* ------------------------------------------------------------------ pipeline */

int orc_match_pipeline(const orc_config* cfg, const float pose[16],
                       const float* src, const float* src_n, const uint8_t* src_c, int64_t n_src,
                       const int32_t* sel, int64_t n_sel,
                       const float* tgt, const float* tgt_n, const uint8_t* tgt_c, int64_t n_tgt,
                       const orc_kdtree* tree, orc_match* out, float* tp_out, float* tn_out) {
    if (!sel) n_sel = n_src;
    float* sp = (float*)malloc(sizeof(float) * 3 * (size_t)(n_sel > 0 ? n_sel : 1));
    float* sn = (float*)malloc(sizeof(float) * 3 * (size_t)(n_sel > 0 ? n_sel : 1));
    uint8_t* sc = (uint8_t*)malloc(4 * (size_t)(n_sel > 0 ? n_sel : 1));
    for (int64_t k = 0; k < n_sel; ++k) {
        const int64_t i = sel ? sel[k] : k;
        memcpy(sp + 3 * k, src + 3 * i, 12);
        if (src_n) memcpy(sn + 3 * k, src_n + 3 * i, 12); else { sn[3 * k] = sn[3 * k + 1] = sn[3 * k + 2] = 0.f; }
        if (src_c) memcpy(sc + 4 * k, src_c + 4 * i, 4); else memset(sc + 4 * k, 0, 4);
    }
    float* tp = (float*)malloc(sizeof(float) * 3 * (size_t)(n_sel > 0 ? n_sel : 1));
    float* tn = (float*)malloc(sizeof(float) * 3 * (size_t)(n_sel > 0 ? n_sel : 1));
    orc_transform_points(pose, sp, n_sel, tp);   /* ICPOptimizer.h:553 */
    orc_transform_normals(pose, sn, n_sel, tn);  /* :554 */
    int rc = 0;
    if (cfg->matching == ORC_MATCH_PROJECTIVE) {
        if ((int64_t)cfg->width * cfg->height != n_tgt || cfg->height == 0) rc = -3; /* NearestNeighbor.h:341-349 */
        else orc_projective(tgt, cfg->width, cfg->height, cfg->fx, cfg->fy, cfg->cx, cfg->cy, tp, n_sel, cfg->max_distance_sq, out);
    } else if (cfg->color_icp) {
        if (tree) orc_kdtree_query(tree, tp, sc, n_sel, cfg->max_distance_sq, out);
        else orc_knn6_brute(tgt, tgt_c, n_tgt, tp, sc, n_sel, cfg->max_distance_sq, out);
    } else {
        if (tree) orc_kdtree_query(tree, tp, NULL, n_sel, cfg->max_distance_sq, out);
        else orc_knn3_brute(tgt, n_tgt, tp, n_sel, cfg->max_distance_sq, out);
    }
    if (rc == 0) {
        orc_apply_weights(cfg->weighting, cfg->weight_max_distance_sq > 0.f ? cfg->weight_max_distance_sq : cfg->max_distance_sq,
                          tp, tgt, tn, tgt_n, sc, tgt_c, n_sel, out); /* :571-572; WeightingMethod(weightingMethod, maxDistance) :528 */
        if (cfg->rejection == 1) orc_prune(tn, tgt_n, n_sel, out);                                        /* :578-579 */
    }
    if (tp_out) memcpy(tp_out, tp, sizeof(float) * 3 * (size_t)n_sel);
    if (tn_out) memcpy(tn_out, tn, sizeof(float) * 3 * (size_t)n_sel);
    free(sp); free(sn); free(sc); free(tp); free(tn);
    return rc;
}

int orc_estimate_pose(const orc_config* cfg,
                      const float* src, const float* src_n, const uint8_t* src_c, int64_t n_src,
                      const float* tgt, const float* tgt_n, const uint8_t* tgt_c, int64_t n_tgt,
                      float pose[16], float* hist, int* n_iters_out, int64_t* n_queries_out) {
    int rc = 0;
    int stride = 1;
    if (cfg->multires) stride = orc_coarsest_stride(n_src);
    int32_t* level = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_src > 0 ? n_src : 1));
    int32_t* sel = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_src > 0 ? n_src : 1));
    int64_t n_level;
    if (cfg->multires) n_level = cfg->pyramid_mode == 1 ? orc_voxel_indices(src, src_n, n_src, stride, level) : orc_coarse_indices(src, src_n, n_src, stride, level);
    else { n_level = n_src; for (int64_t i = 0; i < n_src; ++i) level[i] = (int32_t)i; }
    orc_mt19937 rng; uint32_t n_selections = 0;
    orc_mt_seed(&rng, cfg->seed + n_selections++); /* PointSelection ctor -> initSampler */
    orc_kdtree* tree = NULL;
    if (cfg->matching == ORC_MATCH_KNN && cfg->nn_mode == ORC_NN_KDTREE)
        tree = orc_kdtree_build(tgt, cfg->color_icp ? tgt_c : NULL, n_tgt); /* buildIndex, ICPOptimizer.h:532-535 */
    orc_match* m = (orc_match*)malloc(sizeof(orc_match) * (size_t)(n_src > 0 ? n_src : 1));
    float* tp = (float*)malloc(sizeof(float) * 3 * (size_t)(n_src > 0 ? n_src : 1));
    float* tn = (float*)malloc(sizeof(float) * 3 * (size_t)(n_src > 0 ? n_src : 1));
    float* gs = (float*)malloc(sizeof(float) * 3 * (size_t)(n_src > 0 ? n_src : 1));
    float* gd = (float*)malloc(sizeof(float) * 3 * (size_t)(n_src > 0 ? n_src : 1));
    float* gns = (float*)malloc(sizeof(float) * 3 * (size_t)(n_src > 0 ? n_src : 1));
    float* gnt = (float*)malloc(sizeof(float) * 3 * (size_t)(n_src > 0 ? n_src : 1));
    float* gw = (float*)malloc(sizeof(float) * (size_t)(n_src > 0 ? n_src : 1));
    int iters = 0; int64_t nq = 0;
    for (int i = 0; i < cfg->n_iterations || cfg->multires; ++i) {
        const int32_t* cur = level; int64_t n_cur = n_level;
        if (cfg->selection == ORC_SELECT_RANDOM) { /* resample(), selection.h:57-60,88-104 */
            int64_t c = 0;
            for (int64_t k = 0; k < n_level; ++k) if (orc_mt_canonical(&rng) < cfg->proba) sel[c++] = level[k];
            cur = sel; n_cur = c;
        }
        nq += n_cur;
        rc = orc_match_pipeline(cfg, pose, src, src_n, src_c, n_src, cur, n_cur, tgt, tgt_n, tgt_c, n_tgt, tree, m, tp, tn);
        if (rc != 0) break;
        float inc[16];
        if (cfg->minimizer == ORC_MIN_LM) {
            double x[6]; int nlm;
            rc = orc_solve_lm(cfg->metric, tp, tn, tgt, tgt_n, m, n_cur, cfg->lm_max_iterations, x, inc, &nlm);
        } else {
            /* gather, ICPOptimizer.h:583-610 */
            int64_t g = 0;
            for (int64_t k = 0; k < n_cur; ++k) {
                if (m[k].idx < 0) continue;
                const float* d = tgt + 3 * (int64_t)m[k].idx;
                if (!finite3(tp + 3 * k) || !finite3(d)) continue;
                memcpy(gs + 3 * g, tp + 3 * k, 12); memcpy(gd + 3 * g, d, 12); gw[g] = m[k].weight;
                if (cfg->metric != 0) memcpy(gnt + 3 * g, tgt_n + 3 * (int64_t)m[k].idx, 12);
                if (cfg->metric == 2) memcpy(gns + 3 * g, tn + 3 * k, 12);
                ++g;
            }
            if (cfg->metric == 1) rc = orc_solve_p2plane(gs, gd, gnt, gw, g, inc);
            else if (cfg->metric == 0) rc = orc_solve_p2p(gs, gd, gw, g, inc);
            else rc = orc_solve_symmetric(gs, gd, gns, gnt, gw, g, inc);
        }
        if (rc != 0) break;
        mat4_mul(inc, pose, pose); /* estimatedPose = increment * estimatedPose */
        if (hist) memcpy(hist + 16 * (size_t)iters, pose, 16 * sizeof(float));
        ++iters;
        if (cfg->multires) { /* ICPOptimizer.h:634-655 */
            if (stride == 1 && i >= cfg->n_iterations - 1) break;
            if (stride == 1) continue;
            stride /= 2; if (stride < 1) stride = 1;
            n_level = cfg->pyramid_mode == 1 ? orc_voxel_indices(src, src_n, n_src, stride, level) : orc_coarse_indices(src, src_n, n_src, stride, level);
            orc_mt_seed(&rng, cfg->seed + n_selections++);
        }
    }
    if (n_iters_out) *n_iters_out = iters;
    if (n_queries_out) *n_queries_out = nq;
    orc_kdtree_free(tree);
    free(level); free(sel); free(m); free(tp); free(tn); free(gs); free(gd); free(gns); free(gnt); free(gw);
    return rc;
}

float orc_rmse(const float pose[16], const float* src, const float* ref, int64_t n) {
    /* ConvergenceMeasure.h:50-66 */
    int counter = 0; float rmse = 0.0f;
    for (int64_t i = 0; i < n; ++i) {
        float t[3]; orc_transform_points(pose, src + 3 * i, 1, t);
        if (finite3(t) && finite3(ref + 3 * i)) {
            const float dx = t[0] - ref[3 * i], dy = t[1] - ref[3 * i + 1], dz = t[2] - ref[3 * i + 2];
            rmse += (dx * dx + dy * dy) + dz * dz; counter++;
        }
    }
    rmse /= counter;
    return sqrtf(rmse);
}

/* ------------------------------------------------------------------ input preparation */

/* Inverse of a rigid/affine 4x4 (column-major) in double, rounded to fp32.  The reference calls Eigen's
 * Matrix4f::inverse() (PointCloud.h:88); its drivers only ever pass the identity (VirtualSensor.h:52), for which
 * every implementation returns the identity exactly.  For other extrinsics the contract is "inverse in fp64". */
static void inv4_f64_to_f32(const float M[16], float out[16]) {
    double a[4][8];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { a[r][c] = M[r + 4 * c]; a[r][4 + c] = (r == c) ? 1.0 : 0.0; }
    for (int k = 0; k < 4; ++k) {
        int p = k; for (int i = k + 1; i < 4; ++i) if (fabs(a[i][k]) > fabs(a[p][k])) p = i;
        if (p != k) for (int j = 0; j < 8; ++j) { const double t = a[k][j]; a[k][j] = a[p][j]; a[p][j] = t; }
        const double d = a[k][k];
        for (int j = 0; j < 8; ++j) a[k][j] /= d;
        for (int i = 0; i < 4; ++i) if (i != k) { const double f = a[i][k]; if (f != 0.0) for (int j = 0; j < 8; ++j) a[i][j] -= f * a[k][j]; }
    }
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[r + 4 * c] = (float)a[r][4 + c];
}

/* PointCloud(float* depthMap, BYTE* colorFrame, intrinsics, extrinsics, width, height, keepOriginalSize,
 * downsampleFactor, maxDistance) -- PointCloud.h:78-165.  Returns the number of points written.
 * colour of kept pixel i = bytes colorFrame[i .. i+3] (the reference indexes the RGBX frame with the PIXEL index,
 * PointCloud.h:151-152), so color must hold width*height + 3 bytes at least. */
int64_t orc_cloud_from_depth(const float* depth, const uint8_t* color, const float K[9] /* column-major */, const float E[16] /* nullable */,
                             uint32_t width, uint32_t height, int keep_original_size, uint32_t downsample, float max_distance,
                             float* pts_out, float* nrm_out, uint8_t* rgba_out) {
    const float fovX = K[0], fovY = K[4], cX = K[6], cY = K[7];      /* depthIntrinsics(0,0),(1,1),(0,2),(1,2) */
    const float half = max_distance / 2.f;
    float Einv[16];
    if (E) inv4_f64_to_f32(E, Einv); else mat4_identity(Einv);
    const int64_t n = (int64_t)width * height;
    const float minf = -INFINITY;
    int64_t m = 0;
    if (downsample == 0) return -1;
    for (int64_t i = 0; i < n; i += downsample) {
        const int u = (int)(i % width), v = (int)(i / width);
        float p[3], nr[3] = {minf, minf, minf};
        const float d = depth[i];
        if (d == minf) { p[0] = p[1] = p[2] = minf; }
        else {
            const float c[3] = {(u - cX) / fovX * d, (v - cY) / fovY * d, d};
            for (int r = 0; r < 3; ++r) p[r] = ((Einv[r] * c[0] + Einv[r + 4] * c[1]) + Einv[r + 8] * c[2]) + Einv[r + 12];
        }
        if (u >= 1 && v >= 1 && u < (int)width - 1 && v < (int)height - 1) {
            const float du = 0.5f * (depth[i + 1] - depth[i - 1]);
            const float dv = 0.5f * (depth[i + width] - depth[i - width]);
            if (isfinite(du) && isfinite(dv) && !(fabsf(du) > half) && !(fabsf(dv) > half)) {
                const float a = -du, b = -dv, c1 = 1.0f;
                const float nn = sqrtf((a * a + b * b) + c1 * c1);
                nr[0] = a / nn; nr[1] = b / nn; nr[2] = c1 / nn;
            }
        }
        if (keep_original_size || (finite3(p) && finite3(nr))) {
            for (int r = 0; r < 3; ++r) { pts_out[3 * m + r] = p[r]; nrm_out[3 * m + r] = nr[r]; }
            if (rgba_out) for (int r = 0; r < 4; ++r) rgba_out[4 * m + r] = color ? color[i + r] : 0;
            ++m;
        }
    }
    return m;
}

/* ConvergenceMeasure::benchmarkError / calculate_error (ConvergenceMeasure.h:104-151): mean over i of
 * |T s_i - u_i| / |T s_i - centroid(T s)|, pcl::euclideanDistance in fp32, centroid accumulated in double
 * (pcl::compute3DCentroid into Eigen::Vector4d) and narrowed to float (pcl::PointXYZ), the sum in double. */
double orc_benchmark_error(const float pose[16], const float* src, const float* ref, int64_t n) {
    float* t = (float*)malloc((size_t)(n > 0 ? n : 1) * 3 * sizeof(float));
    orc_transform_points(pose, src, n, t);
    double c[3] = {0, 0, 0}; int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) if (finite3(t + 3 * i)) { c[0] += t[3 * i]; c[1] += t[3 * i + 1]; c[2] += t[3 * i + 2]; ++cnt; }
    if (cnt) { c[0] /= (double)cnt; c[1] /= (double)cnt; c[2] /= (double)cnt; }
    const float cf[3] = {(float)c[0], (float)c[1], (float)c[2]};
    double err = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float* a = t + 3 * i; const float* b = ref + 3 * i;
        const float ex = a[0] - cf[0], ey = a[1] - cf[1], ez = a[2] - cf[2];
        const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        const double centroid_distance = sqrtf((ex * ex + ey * ey) + ez * ez);
        err += sqrtf((dx * dx + dy * dy) + dz * dz) / centroid_distance;
    }
    free(t);
    return n ? err / (double)n : 0.0;
}

/* ------------------------------------------------------------------ k-NN PCA normals (second half) */

/* Jacobi eigen-decomposition of a symmetric 3x3 (row-major) in double, eigenvalues ascending; column k of V
 * (V[3*i+k]) is the k-th eigenvector.  Cyclic sweeps (0,1), (0,2), (1,2) until the off-diagonal sum is below 1e-300 or 64
 * sweeps.  The device kernel (csrc/normals.cu) executes exactly these operations in this order, without FMA contraction. */
static void orc_eig3(const double A_[9], double evals[3], double V[9]) {
    double A[9];
    for (int i = 0; i < 9; ++i) { A[i] = A_[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 64; ++sweep) {
        const double off = (fabs(A[1]) + fabs(A[2])) + fabs(A[5]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
            const double apq = A[p * 3 + q];
            if (apq == 0.0) continue;
            const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 3; ++k) { const double akp = A[k * 3 + p], akq = A[k * 3 + q]; A[k * 3 + p] = c * akp - s * akq; A[k * 3 + q] = s * akp + c * akq; }
            for (int k = 0; k < 3; ++k) { const double apk = A[p * 3 + k], aqk = A[q * 3 + k]; A[p * 3 + k] = c * apk - s * aqk; A[q * 3 + k] = s * apk + c * aqk; }
            for (int k = 0; k < 3; ++k) { const double vkp = V[k * 3 + p], vkq = V[k * 3 + q]; V[k * 3 + p] = c * vkp - s * vkq; V[k * 3 + q] = s * vkp + c * vkq; }
        }
    }
    int ord[3] = {0, 1, 2};   /* stable ascending order of the diagonal */
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2 - i; ++j) if (A[ord[j + 1] * 4] < A[ord[j] * 4]) { const int t = ord[j]; ord[j] = ord[j + 1]; ord[j + 1] = t; }
    double Vs[9];
    for (int k = 0; k < 3; ++k) { evals[k] = A[ord[k] * 4]; for (int i = 0; i < 3; ++i) Vs[i * 3 + k] = V[i * 3 + ord[k]]; }
    for (int i = 0; i < 9; ++i) V[i] = Vs[i];
}

/* PointCloud(pcl::PointCloud<PointXYZ>::Ptr) (PointCloud.h:41-76): pcl::NormalEstimation with setKSearch(k = 5) and the
 * default viewpoint (0,0,0).  PCL is an un-vendored dependency; restated from its published algorithm
 * (features/normal_3d.h): the k nearest neighbours of every point (itself included; exact, (d2, index) order under contract
 * D1), the covariance of the neighbourhood about its mean (double), the eigenvector of the smallest eigenvalue, flipped
 * towards the viewpoint; curvature = |l0 / (l0 + l1 + l2)|.  Non-finite points, and points with fewer than 3 finite
 * neighbours, get NaN normals.  out_nrm: 3n floats, out_curv (nullable): n floats. */
void orc_pca_normals(const float* pts, int64_t n, int k, const float vp[3], float* out_nrm, float* out_curv) {
    if (k > 16) k = 16;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        const float nanf_ = NAN;
        float* o = out_nrm + 3 * i;
        o[0] = o[1] = o[2] = nanf_; if (out_curv) out_curv[i] = nanf_;
        const float* q = pts + 3 * i;
        if (!finite3(q)) continue;
        float bd[16]; int64_t bi[16]; int m = 0;
        for (int64_t j = 0; j < n; ++j) {
            const float* p = pts + 3 * j;
            const float d = d2_3(q, p);
            if (!isfinite(d)) continue;
            /* insert (d, j) into the ascending list; equal distances keep index order because j ascends */
            if (m < k || d < bd[m - 1]) {
                int pos = m < k ? m : k - 1;
                while (pos > 0 && d < bd[pos - 1]) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
                bd[pos] = d; bi[pos] = j;
                if (m < k) ++m;
            }
        }
        if (m < 3) continue;
        double mean[3] = {0, 0, 0};
        for (int a = 0; a < m; ++a) { mean[0] += pts[3 * bi[a]]; mean[1] += pts[3 * bi[a] + 1]; mean[2] += pts[3 * bi[a] + 2]; }
        for (int a = 0; a < 3; ++a) mean[a] /= (double)m;
        double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int a = 0; a < m; ++a) {
            const double d[3] = {pts[3 * bi[a]] - mean[0], pts[3 * bi[a] + 1] - mean[1], pts[3 * bi[a] + 2] - mean[2]};
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C[r * 3 + c] += d[r] * d[c];
        }
        for (int a = 0; a < 9; ++a) C[a] /= (double)m;
        double ev[3], V[9];
        orc_eig3(C, ev, V);
        double nx = V[0], ny = V[3], nz = V[6];
        const double sum = (ev[0] + ev[1]) + ev[2];
        const double vx = (double)vp[0] - q[0], vy = (double)vp[1] - q[1], vz = (double)vp[2] - q[2];
        if ((vx * nx + vy * ny) + vz * nz < 0) { nx = -nx; ny = -ny; nz = -nz; }          /* flipNormalTowardsViewpoint */
        o[0] = (float)nx; o[1] = (float)ny; o[2] = (float)nz;
        if (out_curv) out_curv[i] = sum != 0 ? (float)fabs(ev[0] / sum) : 0.f;
    }
}
