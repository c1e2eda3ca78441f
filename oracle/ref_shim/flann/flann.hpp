// oracle/ref_shim/flann/flann.hpp -- TEST INFRASTRUCTURE ONLY.
//
// Stand-in for the slice of the FLANN 1.8.4 API that NearestNeighbor.h:104-314 calls, so that the
// reference's own NearestNeighborSearchFlann compiles and runs where it lies.  It is NOT FLANN:
// FLANN's KDTreeIndex with checks=16 is an approximate search; this stand-in answers every query
// EXACTLY (a median-split kd-tree whose pruning is safe under fp32 rounding, OpenMP over the
// queries) with FLANN's L2<float> functor semantics -- the squared distance accumulated feature by
// feature, result += diff*diff in fp32 -- and the lowest index on ties (NearestNeighbor.h:81-97 is
// the reference's own statement of the tie rule).  That is the comparand BASELINE.json names ("the
// reference's exact brute-force k-NN").  Define ICP_REF_FLANN_EXHAUSTIVE for the literal O(N*M) scan
// (tests/test_oracle_vs_reference.py checks both give identical answers).
#ifndef ICP_REF_SHIM_FLANN
#define ICP_REF_SHIM_FLANN
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>
namespace flann {
template <class T> class Matrix {
public:
    size_t rows, cols;
    Matrix() : rows(0), cols(0), data_(nullptr) {}
    Matrix(T* d, size_t r, size_t c) : rows(r), cols(c), data_(d) {}
    T* operator[](size_t i) const { return data_ + i * cols; }
    T* ptr() const { return data_; }
private:
    T* data_;
};
template <class T> struct L2 { typedef T ElementType; typedef T ResultType; };
struct KDTreeIndexParams { int trees; explicit KDTreeIndexParams(int t = 4) : trees(t) {} };
struct SearchParams { int checks; float eps; bool sorted; int cores; SearchParams(int c = 32, float e = 0.f, bool s = true) : checks(c), eps(e), sorted(s), cores(1) {} };

inline bool& exhaustive() { static bool e = false; return e; }   // test hook: literal O(N*M) scan

template <class Distance> class Index {
public:
    typedef typename Distance::ElementType E;
    Index(const Matrix<E>& data, const KDTreeIndexParams&) : data_(data), dim_(data.cols) {}
    void buildIndex() {
        order_.clear(); nodes_.clear();
        for (size_t j = 0; j < data_.rows; ++j) {
            bool fin = true; for (size_t k = 0; k < dim_; ++k) fin = fin && std::isfinite(data_[j][k]);
            if (fin) order_.push_back((int)j);      // a non-finite point can never win the strict '<' scan
        }
        if (!order_.empty()) build(0, (int)order_.size());
    }
    int knnSearch(const Matrix<E>& q, Matrix<int>& indices, Matrix<E>& dists, size_t knn, const SearchParams&) const {
        (void)knn;  // the reference only asks for k = 1
        const long nq = (long)q.rows;
        const bool brute = exhaustive();
#pragma omp parallel for schedule(dynamic, 1024)
        for (long i = 0; i < nq; ++i) {
            const E* a = q[i];
            E best = std::numeric_limits<E>::max(); int bi = -1;
            if (brute) {
                for (long j = 0; j < (long)data_.rows; ++j) { const E r = dist(a, data_[j]); if (r < best) { best = r; bi = (int)j; } }
            } else {
                bool fin = true; for (size_t k = 0; k < dim_; ++k) fin = fin && std::isfinite(a[k]);
                if (fin && !nodes_.empty()) query(0, a, best, bi);
            }
            *indices[i] = bi; *dists[i] = best;
        }
        return (int)nq;
    }
private:
    struct Node { int lo, hi, left, right, dim; E split; };
    Matrix<E> data_; size_t dim_;
    std::vector<int> order_; std::vector<Node> nodes_;
    E dist(const E* a, const E* b) const { E r = E(); for (size_t k = 0; k < dim_; ++k) { const E d = a[k] - b[k]; r += d * d; } return r; }
    int build(int lo, int hi) {
        const int id = (int)nodes_.size();
        nodes_.push_back(Node{lo, hi, -1, -1, -1, E()});
        if (hi - lo <= 12) return id;
        int bd = 0; E bext = E(-1);
        for (size_t d = 0; d < dim_; ++d) {
            E mn = std::numeric_limits<E>::max(), mx = -std::numeric_limits<E>::max();
            for (int i = lo; i < hi; ++i) { const E v = data_[order_[i]][d]; mn = std::min(mn, v); mx = std::max(mx, v); }
            if (mx - mn > bext) { bext = mx - mn; bd = (int)d; }
        }
        if (!(bext > E(0))) return id;                  // identical points: one (large) leaf
        const int mid = lo + (hi - lo) / 2;
        std::nth_element(order_.begin() + lo, order_.begin() + mid, order_.begin() + hi,
                         [&](int x, int y) { return data_[x][bd] < data_[y][bd]; });
        const E split = data_[order_[mid]][bd];
        nodes_[id].dim = bd; nodes_[id].split = split;
        const int l = build(lo, mid); const int r = build(mid, hi);     // left: coords <= split, right: >= split
        nodes_[id].left = l; nodes_[id].right = r;
        return id;
    }
    void query(int id, const E* q, E& best, int& bi) const {
        const Node& n = nodes_[id];
        if (n.dim < 0) {
            for (int i = n.lo; i < n.hi; ++i) {
                const int o = order_[i];
                const E r = dist(q, data_[o]);
                if (r < best || (r == best && (bi < 0 || o < bi))) { best = r; bi = o; }   // lowest index on ties, like the in-order scan
            }
            return;
        }
        const E diff = q[n.dim] - n.split;
        const E pd = diff * diff;   // fl((q-split)^2) <= fl(dx*dx) of any point beyond the plane: rounding is monotone
        query(diff <= E(0) ? n.left : n.right, q, best, bi);
        if (!(pd > best)) query(diff <= E(0) ? n.right : n.left, q, best, bi);
    }
};
}  // namespace flann
#endif
