// Host-side mirror of the reference's QP solver interface, backed by the B200 CUDA library.
//
//   qp_solver::QuadraticProblem<Scalar>   reference include/solvers/qp.hpp:19-34
//   qp_solver::QPSolverSettings<Scalar>   reference include/solvers/qp.hpp:36-68
//   qp_solver::QPSolverStatus             reference include/solvers/qp.hpp:70
//   qp_solver::QPSolverInfo<Scalar>       reference include/solvers/qp.hpp:72-108
//   qp_solver::QPSolver<Scalar>           reference include/solvers/qp.hpp:113-250, src/qp.cpp
//   qp_solver::BatchQPSolver              NEW: B independent QPSolver<double> instances in one object
//                                         (the data-parallel axis; reference call site src/sqp.cpp:221-222)
//
// Same names, fields, defaults, enum values and call order as the reference: this directory (host/overlay) holds ONLY
// solvers/qp.hpp (+ its dense.hpp), so putting it BEFORE the reference's include/ on the include path replaces exactly that
// header -- the reference's own src/sqp.cpp, include/solvers/sqp.hpp and bfgs.hpp then compile unmodified against it
// (tests/test_dropin.py builds the reference's src/sqp.cpp and its gtest files that way). Everything numeric happens on the GPU
// through the C-ABI of include/sqp_b200_qp.h; there is no CPU fallback (constructing a solver without a B200 throws).
// QPSolver<float> is the fp32 instantiation on the device too (sqpb200_qp_batch_set_precision); the ABI's arrays stay double.
#pragma once
#include <cstdio>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../../include/sqp_b200_qp.h"
#include "dense.hpp"

#define QP_SOLVER_PRINTING

namespace qp_solver {

template <typename Scalar = double>
struct QuadraticProblem {
    using Vector = sqpb200_dense::Vector<Scalar>;
    using Matrix = sqpb200_dense::Matrix<Scalar>;
    const Matrix *P;
    const Vector *q;
    const Matrix *A;
    const Vector *l;
    const Vector *u;
};

template <typename Scalar>
struct QPSolverSettings {
    Scalar rho = 1e-1;
    Scalar sigma = 1e-6;
    Scalar alpha = 1.0;
    Scalar eps_rel = 1e-3;
    Scalar eps_abs = 1e-3;
    int max_iter = 1000;
    int check_termination = 25;
    bool warm_start = false;
    bool adaptive_rho = false;
    Scalar adaptive_rho_tolerance = 5;
    int adaptive_rho_interval = 25;
    bool verbose = false;

    void print() const {
        printf("ADMM settings:\n");
        printf("  sigma %.2e\n", (double)sigma);
        printf("  rho %.2e\n", (double)rho);
        printf("  alpha %.2f\n", (double)alpha);
        printf("  eps_rel %.1e\n", (double)eps_rel);
        printf("  eps_abs %.1e\n", (double)eps_abs);
        printf("  max_iter %d\n", max_iter);
        printf("  adaptive_rho %d\n", adaptive_rho);
        printf("  warm_start %d\n", warm_start);
    }
    sqpb200_qp_settings to_c() const {
        sqpb200_qp_settings s;
        s.rho = rho; s.sigma = sigma; s.alpha = alpha; s.eps_rel = eps_rel; s.eps_abs = eps_abs;
        s.max_iter = max_iter; s.check_termination = check_termination; s.warm_start = warm_start;
        s.adaptive_rho = adaptive_rho; s.adaptive_rho_tolerance = adaptive_rho_tolerance;
        s.adaptive_rho_interval = adaptive_rho_interval; s.verbose = verbose;
        return s;
    }
};

typedef enum { SOLVED, MAX_ITER_EXCEEDED, UNSOLVED, NUMERICAL_ISSUES, UNINITIALIZED } QPSolverStatus;

template <typename Scalar>
struct QPSolverInfo {
    QPSolverStatus status = UNINITIALIZED;
    int iter = 0;
    int rho_updates = 0;
    Scalar rho_estimate = 0;
    Scalar res_prim = 0;
    Scalar res_dual = 0;

    void print() const {
        static const char *names[] = {"SOLVED", "MAX_ITER_EXCEEDED", "UNSOLVED", "NUMERICAL_ISSUES", "UNINITIALIZED"};
        printf("ADMM info:\n");
        printf("  status %s\n", names[status <= UNINITIALIZED ? status : UNINITIALIZED]);
        printf("  iter %d\n", iter);
        printf("  rho_updates %d\n", rho_updates);
        printf("  rho_estimate %f\n", (double)rho_estimate);
        printf("  res_prim %f\n", (double)res_prim);
        printf("  res_dual %f\n", (double)res_dual);
    }
};

// One CUDA context per (thread, device), shared by every solver object created on that thread.
class Device {
   public:
    static std::shared_ptr<Device> get(int device = 0) {
        static thread_local std::vector<std::weak_ptr<Device>> cache;
        if ((int)cache.size() <= device) cache.resize(device + 1);
        if (auto sp = cache[device].lock()) return sp;
        auto sp = std::shared_ptr<Device>(new Device(device));
        cache[device] = sp;
        return sp;
    }
    ~Device() { sqpb200_ctx_destroy(ctx_); }
    sqpb200_ctx *ctx() const { return ctx_; }
    void check(int rc, const char *what) const {
        if (rc) throw std::runtime_error(std::string(what) + ": " + sqpb200_last_error(ctx_));
    }

   private:
    explicit Device(int device) {
        if (int rc = sqpb200_ctx_create(device, &ctx_))
            throw std::runtime_error(std::string("sqpb200_ctx_create failed (") + std::to_string(rc) + "): " + sqpb200_last_error(nullptr));
    }
    sqpb200_ctx *ctx_ = nullptr;
};

/** B independent QP solver instances of identical size, solved concurrently on one B200.
 *  Arrays are batch-major, each matrix column-major: P[B][n*n] q[B][n] A[B][m*n] l[B][m] u[B][m]. */
class BatchQPSolver {
   public:
    using Settings = QPSolverSettings<double>;
    BatchQPSolver(int batch, int n, int m, int device = 0) : dev_(Device::get(device)), batch_(batch), n_(n), m_(m) {
        dev_->check(sqpb200_qp_batch_create(dev_->ctx(), batch, n, m, &h_), "sqpb200_qp_batch_create");
        // results land in page-locked host memory (asynchronous device-to-host copies, no driver-side staging)
        const size_t B = (size_t)batch;
        x_ = pinned<double>(B * n);
        y_ = pinned<double>(B * (m > 0 ? m : 1));
        rho_estimate_ = pinned<double>(B);
        res_prim_ = pinned<double>(B);
        res_dual_ = pinned<double>(B);
        status_ = pinned<int>(B);
        iter_ = pinned<int>(B);
        rho_updates_ = pinned<int>(B);
        for (size_t i = 0; i < B; ++i) {
            status_[i] = UNINITIALIZED;
            iter_[i] = rho_updates_[i] = 0;
            rho_estimate_[i] = res_prim_[i] = res_dual_[i] = 0.0;
        }
    }
    ~BatchQPSolver() {
        if (stream_) {
            sqpb200_stream_sync(dev_->ctx(), stream_);
            sqpb200_stream_destroy(dev_->ctx(), stream_);
        }
        sqpb200_qp_batch_destroy(h_);
        for (void *p : pinned_) sqpb200_host_free(dev_->ctx(), p);
    }
    BatchQPSolver(const BatchQPSolver &) = delete;
    BatchQPSolver &operator=(const BatchQPSolver &) = delete;

    // compute in fp32 (the reference's QPSolver<float>, qp.cpp:386) instead of fp64; the arrays of the interface stay double
    void set_precision_fp32(bool fp32) { dev_->check(sqpb200_qp_batch_set_precision(h_, fp32 ? 1 : 0), "set_precision"); }

    int batch() const { return batch_; }
    int num_var() const { return n_; }
    int num_constr() const { return m_; }
    Settings &settings() { return settings_; }
    const Settings &settings() const { return settings_; }
    // read back rho_updates / rho_estimate / res_prim / res_dual after every call (default) or only x, y, status, iter
    void fetch_full_info(bool on) { full_info_ = on; }

    // QPSolver::setup / update_qp / solve over the first `count` instances (host pointers)
    void setup(const double *P, const double *q, const double *A, const double *l, const double *u, int count = -1) {
        call(sqpb200_qp_batch_setup, "setup", P, q, A, l, u, count);
    }
    void update_qp(const double *P, const double *q, const double *A, const double *l, const double *u, int count = -1) {
        call(sqpb200_qp_batch_update_qp, "update_qp", P, q, A, l, u, count);
    }
    void solve(const double *P, const double *q, const double *A, const double *l, const double *u, int count = -1) {
        call(sqpb200_qp_batch_solve, "solve", P, q, A, l, u, count);
    }
    // setup immediately followed by solve in one launch: the pattern of SQP<T>::run_solve_qp (sqp.cpp:221-222)
    // opts: SQPB200_KEEP_FACTOR / SQPB200_REUSE_FACTOR for re-solves of the same P, A with new q, l, u
    void setup_solve(const double *P, const double *q, const double *A, const double *l, const double *u, int count = -1,
                     unsigned opts = 0) {
        auto fn = [opts](sqpb200_qp_batch *h, const sqpb200_qp_settings *s, int c, const double *P_, const double *q_, const double *A_,
                         const double *l_, const double *u_, unsigned flags, void *stream) {
            return sqpb200_qp_batch_setup_solve_opts(h, s, c, P_, q_, A_, l_, u_, flags, stream, opts);
        };
        call(fn, "setup_solve", P, q, A, l, u, count);
    }

    // Page-locked input buffers owned by this object, sized for the whole batch: callers that re-solve every few hundred
    // microseconds (the SQP outer loop) pack the problem data straight into them and call setup_solve_staged().
    double *staged_P() { return staged(0, (size_t)n_ * n_); }
    double *staged_q() { return staged(1, (size_t)n_); }
    double *staged_A() { return staged(2, (size_t)(m_ > 0 ? m_ : 1) * n_); }
    double *staged_l() { return staged(3, (size_t)(m_ > 0 ? m_ : 1)); }
    double *staged_u() { return staged(4, (size_t)(m_ > 0 ? m_ : 1)); }
    void setup_solve_staged(int count = -1, unsigned opts = 0) {
        setup_solve(staged_P(), staged_q(), staged_A(), staged_l(), staged_u(), count, opts);
    }
    // The same, asynchronously on this object's own stream: copies of the staged inputs, the launch and the read-back of x, y, status
    // and iter are enqueued and the call returns at once; wait() makes the results (primal_solution(i), info(i), ...) valid. The staged
    // buffers must not be touched in between. Lets a caller pipeline host work against the GPU (sqp::BatchSQP alternates two groups).
    void setup_solve_staged_async(int count = -1, unsigned opts = 0) {
        if (count < 0) count = batch_;
        if (!stream_) dev_->check(sqpb200_stream_create(dev_->ctx(), &stream_), "sqpb200_stream_create");
        sqpb200_qp_settings s = settings_.to_c();
        dev_->check(sqpb200_qp_batch_setup_solve_opts(h_, &s, count, staged_P(), staged_q(), staged_A(), staged_l(), staged_u(), SQPB200_HOST_ASYNC,
                                                      stream_, opts),
                    "setup_solve (async)");
        dev_->check(sqpb200_qp_batch_get(h_, count, x_, y_, nullptr, status_, iter_, full_info_ ? rho_updates_ : nullptr,
                                         full_info_ ? rho_estimate_ : nullptr, full_info_ ? res_prim_ : nullptr,
                                         full_info_ ? res_dual_ : nullptr, SQPB200_HOST_ASYNC, stream_),
                    "get (async)");
    }
    void wait() {
        if (stream_) dev_->check(sqpb200_stream_sync(dev_->ctx(), stream_), "sqpb200_stream_sync");
    }

    // The reference's intended sparse variant (Eigen::SparseMatrix A: include/solvers/qp.hpp:22-25,
    // include/unsupported/qp_solver.hpp:363-394; tests/qp_solver_sparse_test.cpp): A in compressed column storage
    // (layout SQPB200_SPARSE_CSC, Eigen's outerIndexPtr / innerIndexPtr / valuePtr) or compressed row storage, ONE pattern for
    // the batch, values[batch][nnz]. setup + solve in one launch.
    void setup_solve_sparse(const double *P, const double *q, const double *A_values, const int *A_outer, const int *A_inner, int nnz,
                            int layout, const double *l, const double *u, int count = -1) {
        call_sparse(sqpb200_qp_batch_setup_solve_sparse, "setup_solve_sparse", P, q, A_values, A_outer, A_inner, nnz, layout, l, u, count);
    }
    // the separate calls of the object API with a sparse A (tests/qp_solver_sparse_test.cpp:68-98: setup, solve, solve, update_qp, solve)
    void setup_sparse(const double *P, const double *q, const double *A_values, const int *A_outer, const int *A_inner, int nnz, int layout,
                      const double *l, const double *u, int count = -1) {
        call_sparse(sqpb200_qp_batch_setup_sparse, "setup_sparse", P, q, A_values, A_outer, A_inner, nnz, layout, l, u, count);
    }
    void update_qp_sparse(const double *P, const double *q, const double *A_values, const int *A_outer, const int *A_inner, int nnz,
                          int layout, const double *l, const double *u, int count = -1) {
        call_sparse(sqpb200_qp_batch_update_qp_sparse, "update_qp_sparse", P, q, A_values, A_outer, A_inner, nnz, layout, l, u, count);
    }
    void solve_sparse(const double *P, const double *q, const double *A_values, const int *A_outer, const int *A_inner, int nnz, int layout,
                      const double *l, const double *u, int count = -1) {
        call_sparse(sqpb200_qp_batch_solve_sparse, "solve_sparse", P, q, A_values, A_outer, A_inner, nnz, layout, l, u, count);
    }

    const double *primal_solution(int i = 0) const { return x_ + (size_t)i * n_; }
    const double *dual_solution(int i = 0) const { return y_ + (size_t)i * m_; }
    QPSolverInfo<double> info(int i) const {
        QPSolverInfo<double> r;
        r.status = (QPSolverStatus)status_[i];
        r.iter = iter_[i];
        r.rho_updates = rho_updates_[i];
        r.rho_estimate = rho_estimate_[i];
        r.res_prim = res_prim_[i];
        r.res_dual = res_dual_[i];
        return r;
    }
    long long total_iterations() {
        long long t = 0;
        dev_->check(sqpb200_qp_batch_total_iters(h_, &t, nullptr), "total_iters");
        return t;
    }

   private:
    template <typename T>
    T *pinned(size_t count) {
        void *p = nullptr;
        dev_->check(sqpb200_host_alloc(dev_->ctx(), sizeof(T) * (count ? count : 1), &p), "sqpb200_host_alloc");
        pinned_.push_back(p);
        return static_cast<T *>(p);
    }
    double *staged(int k, size_t per_instance) {
        if (!in_[k]) in_[k] = pinned<double>((size_t)batch_ * per_instance);
        return in_[k];
    }
    void fetch(int count) {
        dev_->check(sqpb200_qp_batch_get(h_, count, x_, y_, nullptr, status_, iter_, full_info_ ? rho_updates_ : nullptr,
                                         full_info_ ? rho_estimate_ : nullptr, full_info_ ? res_prim_ : nullptr,
                                         full_info_ ? res_dual_ : nullptr, SQPB200_HOST_PTRS, nullptr),
                    "get");
    }
    template <typename F>
    void call_sparse(F fn, const char *what, const double *P, const double *q, const double *A_values, const int *A_outer, const int *A_inner,
                     int nnz, int layout, const double *l, const double *u, int count) {
        if (count < 0) count = batch_;
        sqpb200_qp_settings s = settings_.to_c();
        dev_->check(fn(h_, &s, count, P, q, A_values, A_outer, A_inner, nnz, layout, l, u, SQPB200_HOST_PTRS, nullptr), what);
        fetch(count);
    }
    template <typename F>
    void call(F fn, const char *what, const double *P, const double *q, const double *A, const double *l, const double *u, int count) {
        if (count < 0) count = batch_;
        sqpb200_qp_settings s = settings_.to_c();
        dev_->check(fn(h_, &s, count, P, q, A, l, u, SQPB200_HOST_PTRS, nullptr), what);
        fetch(count);
    }
    std::shared_ptr<Device> dev_;
    sqpb200_qp_batch *h_ = nullptr;
    int batch_, n_, m_;
    Settings settings_;
    bool full_info_ = true;
    double *x_ = nullptr, *y_ = nullptr, *rho_estimate_ = nullptr, *res_prim_ = nullptr, *res_dual_ = nullptr;
    int *status_ = nullptr, *iter_ = nullptr, *rho_updates_ = nullptr;
    double *in_[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    std::vector<void *> pinned_;
    void *stream_ = nullptr;  // created on first asynchronous call
};

/**
 *  minimize        0.5 x' P x + q' x
 *  subject to      l <= A x <= u
 *
 *  Drop-in for the reference's QPSolver<SCALAR> (a batch of one on the GPU). For throughput use
 *  BatchQPSolver: a single small QP cannot fill a B200.
 */
template <typename SCALAR>
class QPSolver {
   public:
    using Scalar = SCALAR;
    using QP = QuadraticProblem<Scalar>;
    using Vector = sqpb200_dense::Vector<Scalar>;
    using Matrix = sqpb200_dense::Matrix<Scalar>;
    using Settings = QPSolverSettings<Scalar>;
    using Info = QPSolverInfo<Scalar>;

    enum { INEQUALITY_CONSTRAINT, EQUALITY_CONSTRAINT, LOOSE_BOUNDS } ConstraintType;

    static constexpr Scalar RHO_MIN = 1e-6;
    static constexpr Scalar RHO_MAX = 1e+6;
    static constexpr Scalar RHO_TOL = 1e-4;
    static constexpr Scalar RHO_EQ_FACTOR = 1e+3;
    static constexpr Scalar LOOSE_BOUNDS_THRESH = 1e+16;
    static constexpr Scalar DIV_BY_ZERO_REGUL = std::numeric_limits<Scalar>::epsilon();

    QPSolver() = default;

    /** Setup solver for QP. (reference src/qp.cpp:11-44) */
    void setup(const QP &qp) {
        n = (size_t)qp.P->rows();
        m = (size_t)qp.A->rows();
        if (!batch_ || batch_->num_var() != (int)n || batch_->num_constr() != (int)m) {
            // a new batch object is a default-constructed solver; carry the cumulative counter over (qp.cpp:313)
            int carried = info_.rho_updates;
            batch_.reset(new BatchQPSolver(1, (int)n, (int)m));
            batch_->set_precision_fp32(sizeof(Scalar) == 4);  // QPSolver<float> computes in fp32 on the device too
            carry_rho_updates_ = carried;
        }
        x.resize(n);
        y.resize(m);
        run(qp, &BatchQPSolver::setup);
    }
    /** Update solver for QP of same size as initial setup. (reference src/qp.cpp:46-62) */
    void update_qp(const QP &qp) {
        require_setup("update_qp");
        run(qp, &BatchQPSolver::update_qp);
    }
    /** Solve the QP. (reference src/qp.cpp:64-157) */
    void solve(const QP &qp) {
        if (info_.status == UNINITIALIZED || info_.status == NUMERICAL_ISSUES) return;  // qp.cpp:68-71
        if (settings_.verbose) settings_.print();
        run(qp, &BatchQPSolver::solve);
        if (settings_.verbose) info_.print();
    }

    inline const Vector &primal_solution() const { return x; }
    inline Vector &primal_solution() { return x; }
    inline const Vector &dual_solution() const { return y; }
    inline Vector &dual_solution() { return y; }
    inline const Settings &settings() const { return settings_; }
    inline Settings &settings() { return settings_; }
    inline const Info &info() const { return info_; }
    inline Info &info() { return info_; }

    /* Public function for unit testing (reference qp.hpp:173, src/qp.cpp:283-294) */
    static void constr_type_init(const Vector &l, const Vector &u, sqpb200_dense::VectorXi &constr_type) {
        std::vector<double> ld((size_t)l.rows()), ud((size_t)u.rows());
        std::vector<int> out((size_t)l.rows());
        for (size_t i = 0; i < ld.size(); ++i) { ld[i] = l(i); ud[i] = u(i); }
        sqpb200_constr_type_init(ld.data(), ud.data(), (int)ld.size(), out.data());
        for (size_t i = 0; i < out.size(); ++i) constr_type(i) = out[i];
    }

   private:
    void require_setup(const char *what) const {
        if (!batch_) throw std::logic_error(std::string("QPSolver::") + what + " called before setup()");
    }
    typedef void (BatchQPSolver::*BatchFn)(const double *, const double *, const double *, const double *, const double *, int);
    void run(const QP &qp, BatchFn fn) {
        // QuadraticProblem holds non-owning pointers to column-major storage: pass them straight through for
        // double; convert at the boundary for float.
        const size_t nn = n * n, mn = m * n;
        stage(P_, qp.P->data(), nn);
        stage(q_, qp.q->data(), n);
        stage(A_, qp.A->data(), mn);
        stage(l_, qp.l->data(), m);
        stage(u_, qp.u->data(), m);
        BatchQPSolver::Settings &bs = batch_->settings();
        bs.rho = settings_.rho; bs.sigma = settings_.sigma; bs.alpha = settings_.alpha;
        bs.eps_rel = settings_.eps_rel; bs.eps_abs = settings_.eps_abs; bs.max_iter = settings_.max_iter;
        bs.check_termination = settings_.check_termination; bs.warm_start = settings_.warm_start;
        bs.adaptive_rho = settings_.adaptive_rho; bs.adaptive_rho_tolerance = settings_.adaptive_rho_tolerance;
        bs.adaptive_rho_interval = settings_.adaptive_rho_interval; bs.verbose = settings_.verbose;
        (batch_.get()->*fn)(ptr(P_, qp.P->data()), ptr(q_, qp.q->data()), ptr(A_, qp.A->data()), ptr(l_, qp.l->data()),
                            ptr(u_, qp.u->data()), 1);
        const double *xs = batch_->primal_solution(0), *ys = batch_->dual_solution(0);
        for (size_t i = 0; i < n; ++i) x(i) = (Scalar)xs[i];
        for (size_t i = 0; i < m; ++i) y(i) = (Scalar)ys[i];
        QPSolverInfo<double> bi = batch_->info(0);
        info_.status = bi.status;
        info_.iter = bi.iter;
        info_.rho_updates = bi.rho_updates + carry_rho_updates_;
        info_.rho_estimate = (Scalar)bi.rho_estimate;
        info_.res_prim = (Scalar)bi.res_prim;
        info_.res_dual = (Scalar)bi.res_dual;
    }
    static void stage(std::vector<double> &, const double *, size_t) {}
    static void stage(std::vector<double> &buf, const float *src, size_t len) { buf.assign(src, src + len); }
    static const double *ptr(const std::vector<double> &, const double *src) { return src; }
    static const double *ptr(const std::vector<double> &buf, const float *) { return buf.data(); }

    size_t n = 0;  //< number of variables
    size_t m = 0;  //< number of constraints
    Vector x;      //< primal variable, size n
    Vector y;      //< dual variable, size m
    Settings settings_;
    Info info_;
    int carry_rho_updates_ = 0;
    std::unique_ptr<BatchQPSolver> batch_;
    std::vector<double> P_, q_, A_, l_, u_;  // float -> double staging
};

}  // namespace qp_solver
