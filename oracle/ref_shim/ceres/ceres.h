// oracle/ref_shim/ceres/ceres.h -- TEST INFRASTRUCTURE ONLY.
//
// Stand-in for the slice of the Ceres 2.x API that ICPOptimizer.h:283-310,352-482 and
// constraints.h use: Jet-based AutoDiffCostFunction, Problem::AddResidualBlock with ONE parameter
// block and no loss, Solver::Options / Summary, Solve().  Solve() restates Ceres'
// TrustRegionMinimizer + LevenbergMarquardtStrategy (defaults of Ceres 2.x: initial radius 1e4,
// max 1e16, min 1e-32, min_relative_decrease 1e-3, function/gradient/parameter tolerances
// 1e-6/1e-10/1e-8, min/max LM diagonal 1e-6/1e32, Jacobi scaling) on the dense normal equations.
// It is NOT Ceres (DENSE_QR is replaced by a 6x6 elimination).  What the library built with it
// pins is the reference's own problem construction -- which residual blocks are added, with which
// functor, weight and lambda -- and its outer loop; the LM restatement itself is the same
// algorithm as oracle/icp_oracle.c:orc_solve_lm, written independently against the Jet functors.
#ifndef ICP_REF_SHIM_CERES
#define ICP_REF_SHIM_CERES
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>
namespace ceres {

template <typename T, int N>
struct Jet {
    T a; T v[N];
    Jet() : a() { for (int i = 0; i < N; ++i) v[i] = T(); }
    Jet(const T& s) : a(s) { for (int i = 0; i < N; ++i) v[i] = T(); }  // implicit, like Ceres
    template <class U, class = typename std::enable_if<std::is_arithmetic<U>::value && !std::is_same<U, T>::value>::type>
    explicit Jet(const U& s) : a(static_cast<T>(s)) { for (int i = 0; i < N; ++i) v[i] = T(); }
};
#define ICP_JET_BIN(op, expr_a, expr_v)                                                                 \
    template <typename T, int N> inline Jet<T, N> operator op(const Jet<T, N>& f, const Jet<T, N>& g) { \
        Jet<T, N> h; h.a = expr_a; for (int i = 0; i < N; ++i) h.v[i] = expr_v; return h; }
ICP_JET_BIN(+, f.a + g.a, f.v[i] + g.v[i])
ICP_JET_BIN(-, f.a - g.a, f.v[i] - g.v[i])
ICP_JET_BIN(*, f.a * g.a, f.a * g.v[i] + f.v[i] * g.a)
#undef ICP_JET_BIN
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
    Jet<T, N> h; const T g_a_inverse = T(1.0) / g.a; h.a = f.a * g_a_inverse; const T f_a_by_g_a = h.a;
    for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse; return h; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f) { Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <typename T, int N> inline bool operator>(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a > g.a; }
template <typename T, int N> inline bool operator<(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a < g.a; }
template <typename T, int N> inline Jet<T, N> sqrt(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::sqrt(f.a); const T k = T(1.0) / (T(2.0) * h.a); for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * k; return h; }
template <typename T, int N> inline Jet<T, N> cos(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::cos(f.a); const T k = -std::sin(f.a); for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * k; return h; }
template <typename T, int N> inline Jet<T, N> sin(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::sin(f.a); const T k = std::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * k; return h; }

class CostFunction {
public:
    virtual ~CostFunction() {}
    virtual int num_residuals() const = 0;
    // residuals[num_residuals]; jacobian (nullable) row-major num_residuals x 6
    virtual bool Evaluate(const double* x, double* residuals, double* jacobian) const = 0;
};
class LossFunction;

template <class Functor, int kNumResiduals, int N0>
class AutoDiffCostFunction : public CostFunction {
public:
    explicit AutoDiffCostFunction(Functor* f) : f_(f) {}
    int num_residuals() const override { return kNumResiduals; }
    bool Evaluate(const double* x, double* residuals, double* jacobian) const override {
        if (!jacobian) return (*f_)(x, residuals);
        typedef Jet<double, N0> J;
        J xj[N0], r[kNumResiduals];
        for (int i = 0; i < N0; ++i) { xj[i] = J(x[i]); xj[i].v[i] = 1.0; }
        if (!(*f_)(xj, r)) return false;
        for (int k = 0; k < kNumResiduals; ++k) { residuals[k] = r[k].a; for (int i = 0; i < N0; ++i) jacobian[k * N0 + i] = r[k].v[i]; }
        return true;
    }
private:
    std::unique_ptr<Functor> f_;
};

enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };

class Problem {
public:
    Problem() : params_(nullptr) {}
    void AddResidualBlock(CostFunction* c, LossFunction*, double* params) { blocks_.emplace_back(c); params_ = params; }
    int NumResidualBlocks() const { return (int)blocks_.size(); }
    std::vector<std::unique_ptr<CostFunction> > blocks_;
    double* params_;
};

struct Solver {
    struct Options {
        TrustRegionStrategyType trust_region_strategy_type = LEVENBERG_MARQUARDT;
        bool use_nonmonotonic_steps = false;
        LinearSolverType linear_solver_type = DENSE_QR;
        bool minimizer_progress_to_stdout = false;
        int max_num_iterations = 50;
        int num_threads = 1;
        double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
        double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
        int max_num_consecutive_invalid_steps = 5;
        double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
        bool jacobi_scaling = true;
    };
    struct Summary {
        double initial_cost = 0, final_cost = 0; int num_iterations = 0;
        std::string BriefReport() const { return "shim: iterations " + std::to_string(num_iterations) + ", cost " + std::to_string(initial_cost) + " -> " + std::to_string(final_cost); }
        std::string FullReport() const { return BriefReport(); }
    };
};

namespace shim {
// last Solve()'s iteration count, readable by the test driver
inline int& last_num_iterations() { static thread_local int n = 0; return n; }
inline bool eval(const Problem& p, const double x[6], double* cost, double* H /*36 row-major, nullable*/, double* g) {
    double c = 0; if (H) { std::memset(H, 0, 36 * sizeof(double)); std::memset(g, 0, 6 * sizeof(double)); }
    double r[8], J[8 * 6];
    for (const auto& b : p.blocks_) {
        const int nr = b->num_residuals();
        if (!b->Evaluate(x, r, H ? J : nullptr)) return false;
        for (int k = 0; k < nr; ++k) {
            c += r[k] * r[k];
            if (H) for (int a = 0; a < 6; ++a) { g[a] += J[k * 6 + a] * r[k]; for (int q = 0; q < 6; ++q) H[a * 6 + q] += J[k * 6 + a] * J[k * 6 + q]; }
        }
    }
    *cost = 0.5 * c; return true;
}
inline bool solve6(double A[36], double b[6], double x[6]) {  // partial-pivot elimination
    for (int k = 0; k < 6; ++k) {
        int p = k; for (int i = k + 1; i < 6; ++i) if (std::fabs(A[i * 6 + k]) > std::fabs(A[p * 6 + k])) p = i;
        if (!(std::fabs(A[p * 6 + k]) > 0.0)) return false;
        if (p != k) { for (int j = 0; j < 6; ++j) std::swap(A[k * 6 + j], A[p * 6 + j]); std::swap(b[k], b[p]); }
        for (int i = k + 1; i < 6; ++i) { const double f = A[i * 6 + k] / A[k * 6 + k]; for (int j = k; j < 6; ++j) A[i * 6 + j] -= f * A[k * 6 + j]; b[i] -= f * b[k]; }
    }
    for (int i = 5; i >= 0; --i) { double acc = b[i]; for (int j = i + 1; j < 6; ++j) acc -= A[i * 6 + j] * x[j]; x[i] = acc / A[i * 6 + i]; }
    for (int i = 0; i < 6; ++i) if (!std::isfinite(x[i])) return false;
    return true;
}
}  // namespace shim

inline void Solve(const Solver::Options& o, Problem* problem, Solver::Summary* summary) {
    double* x = problem->params_;
    shim::last_num_iterations() = 0;
    if (!x || problem->blocks_.empty()) return;
    double cost, H[36], g[6];
    if (!shim::eval(*problem, x, &cost, H, g)) return;
    summary->initial_cost = cost;
    double scale[6];
    for (int i = 0; i < 6; ++i) scale[i] = o.jacobi_scaling ? 1.0 / (1.0 + std::sqrt(H[i * 6 + i])) : 1.0;
    double radius = o.initial_trust_region_radius, decrease_factor = 2.0, diag[6];
    bool reuse_diagonal = false, step_successful = true; int invalid_steps = 0, iter = 0;
    auto max_abs = [](const double* v) { double m = 0; for (int i = 0; i < 6; ++i) m = std::max(m, std::fabs(v[i])); return m; };
    double gmax = max_abs(g);
    for (;;) {
        if (iter >= o.max_num_iterations) break;
        if (step_successful && gmax <= o.gradient_tolerance) break;
        if (radius <= o.min_trust_region_radius) break;
        ++iter;
        double Hs[36], gs[6];
        for (int a = 0; a < 6; ++a) { gs[a] = g[a] * scale[a]; for (int b = 0; b < 6; ++b) Hs[a * 6 + b] = H[a * 6 + b] * scale[a] * scale[b]; }
        if (!reuse_diagonal) for (int a = 0; a < 6; ++a) diag[a] = std::min(std::max(Hs[a * 6 + a], o.min_lm_diagonal), o.max_lm_diagonal);
        double A[36], b[6], ds[6];
        std::memcpy(A, Hs, sizeof(A));
        for (int a = 0; a < 6; ++a) { A[a * 6 + a] += diag[a] / radius; b[a] = -gs[a]; }
        const bool lin_ok = shim::solve6(A, b, ds);
        double model_cost_change = 0.0;
        if (lin_ok) for (int a = 0; a < 6; ++a) { double hd = 0; for (int c = 0; c < 6; ++c) hd += Hs[a * 6 + c] * ds[c]; model_cost_change -= ds[a] * (gs[a] + 0.5 * hd); }
        if (!lin_ok || !(model_cost_change > 0.0)) {
            if (++invalid_steps >= o.max_num_consecutive_invalid_steps) break;
            radius *= 0.5; reuse_diagonal = true; step_successful = false; continue;
        }
        invalid_steps = 0;
        double cand[6], cand_cost, step_norm = 0, x_norm = 0;
        for (int a = 0; a < 6; ++a) { cand[a] = x[a] + ds[a] * scale[a]; step_norm += (x[a] - cand[a]) * (x[a] - cand[a]); x_norm += x[a] * x[a]; }
        shim::eval(*problem, cand, &cand_cost, nullptr, nullptr);
        if (std::sqrt(step_norm) <= o.parameter_tolerance * (std::sqrt(x_norm) + o.parameter_tolerance)) break;
        const double cost_change = cost - cand_cost;
        if (std::fabs(cost_change) <= o.function_tolerance * cost) break;
        const double rho = cost_change / model_cost_change;
        if (rho > o.min_relative_decrease) {
            std::memcpy(x, cand, 6 * sizeof(double));
            shim::eval(*problem, x, &cost, H, g); gmax = max_abs(g);
            const double t = 2.0 * rho - 1.0;
            radius = std::min(o.max_trust_region_radius, radius / std::max(1.0 / 3.0, 1.0 - t * t * t));
            decrease_factor = 2.0; reuse_diagonal = false; step_successful = true;
        } else {
            radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true; step_successful = false;
        }
    }
    summary->final_cost = cost; summary->num_iterations = iter;
    shim::last_num_iterations() = iter;
}
}  // namespace ceres
#endif
