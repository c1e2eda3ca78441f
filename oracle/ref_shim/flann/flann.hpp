// oracle/ref_shim/flann/flann.hpp -- TEST INFRASTRUCTURE ONLY.
//
// Stand-in for the slice of the FLANN 1.8.4 API that NearestNeighbor.h:104-314 calls, so that the
// reference's own NearestNeighborSearchFlann compiles and runs where it lies.  It is NOT FLANN:
// FLANN's KDTreeIndex with checks=16 is an approximate search; this stand-in answers every query
// EXACTLY (exhaustive scan, OpenMP over queries) with FLANN's L2<float> functor semantics -- the
// squared distance accumulated feature by feature, result += diff*diff in fp32 -- and the lowest
// index on ties (NearestNeighbor.h:81-97 is the reference's own statement of the tie rule).
// That is the comparand BASELINE.json names ("the reference's exact brute-force k-NN").
#ifndef ICP_REF_SHIM_FLANN
#define ICP_REF_SHIM_FLANN
#include <cstddef>
#include <limits>
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
template <class Distance> class Index {
public:
    typedef typename Distance::ElementType E;
    Index(const Matrix<E>& data, const KDTreeIndexParams&) : data_(data) {}
    void buildIndex() {}
    int knnSearch(const Matrix<E>& q, Matrix<int>& indices, Matrix<E>& dists, size_t knn, const SearchParams&) const {
        (void)knn;  // the reference only asks for k = 1
        const long nq = (long)q.rows, nt = (long)data_.rows; const size_t dim = data_.cols;
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < nq; ++i) {
            const E* a = q[i];
            E best = std::numeric_limits<E>::max(); int bi = -1;
            for (long j = 0; j < nt; ++j) {
                const E* b = data_[j];
                E r = E();
                for (size_t k = 0; k < dim; ++k) { const E d = a[k] - b[k]; r += d * d; }
                if (r < best) { best = r; bi = (int)j; }
            }
            *indices[i] = bi; *dists[i] = best;
        }
        return (int)nq;
    }
private:
    Matrix<E> data_;
};
}  // namespace flann
#endif
