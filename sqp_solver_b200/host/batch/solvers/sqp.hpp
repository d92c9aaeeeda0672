// Host-side SQP outer loop on top of the GPU QP solver -- the caller of the hot path.
//
//   sqp::sqp_settings_t / Status / Info / NonLinearProblem / SQP   reference include/solvers/sqp.hpp:13-165, src/sqp.cpp
//   sqp::BatchSQP   NEW: B problem instances advanced in lock-step; every outer iteration forms B QP
//                   subproblems on the host and solves them in ONE batched GPU launch
//                   (BASELINE.json config 4).
//
// Algorithm: N&W Alg. 18.3 as the reference implements it -- linearise, damped-BFGS Hessian of the
// Lagrangian, positive-definiteness repair, QP subproblem, optional second-order correction,
// l1-merit backtracking line search, step, termination on step norms + constraint violation.
// The outer loop and BFGS stay on the host (user callbacks are host virtuals); only the QP is on the GPU.
#pragma once
#include <chrono>
#include <cmath>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <vector>

#include "../../overlay/solvers/qp.hpp"
#include "bfgs.hpp"

namespace sqp {

template <typename T>
class SQP;

template <typename Scalar>
struct sqp_settings_t {
    Scalar tau = 0.5;       /**< line search iteration decrease, 0 < tau < 1 */
    Scalar eta = 0.25;      /**< line search parameter, 0 < eta < 1 */
    Scalar rho = 0.5;       /**< line search parameter, 0 < rho < 1 */
    Scalar eps_prim = 1e-4; /**< primal step termination threshold, eps_prim > 0 */
    Scalar eps_dual = 1e-4; /**< dual step termination threshold, eps_dual > 0 */
    int max_iter = 100;
    int line_search_max_iter = 20;
    bool second_order_correction = false;
    std::function<void(SQP<Scalar> &)> iteration_callback;

    // The reference's validate() demands eps_prim < 0 (sqp.hpp:28) and is never called; this one checks
    // the documented ranges instead.
    bool validate() const {
        return 0.0 < tau && tau < 1.0 && 0.0 < eta && eta < 1.0 && 0.0 < rho && rho < 1.0 && eps_prim > 0.0 && eps_dual > 0.0 &&
               max_iter > 0 && line_search_max_iter > 0;
    }
};

typedef enum { SOLVED, MAX_ITER_EXCEEDED, INVALID_SETTINGS } Status;

struct Info {
    int iter = 0;
    int qp_solver_iter = 0;
    Status status = MAX_ITER_EXCEEDED;

    void print() const {
        static const char *names[] = {"SOLVED", "MAX_ITER_EXCEEDED", "INVALID_SETTINGS"};
        printf("SQP info:\n  iter: %d\n  qp_solver_iter: %d\n  status: %s\n", iter, qp_solver_iter,
               status <= INVALID_SETTINGS ? names[status] : "UNKNOWN");
    }
};

template <typename Scalar_ = double>
struct NonLinearProblem {
    using Scalar = Scalar_;
    using Matrix = sqpb200_dense::Matrix<Scalar>;
    using Vector = sqpb200_dense::Vector<Scalar>;

    int num_var;
    int num_constr;

    virtual ~NonLinearProblem() = default;
    virtual void objective(const Vector &x, Scalar &obj) = 0;
    virtual void objective_linearized(const Vector &x, Vector &grad, Scalar &obj) = 0;
    virtual void constraint(const Vector &x, Vector &c, Vector &l, Vector &u) = 0;
    virtual void constraint_linearized(const Vector &x, Matrix &Jc, Vector &c, Vector &l, Vector &u) = 0;
};

namespace detail {

// The QP settings the reference's SQP constructor installs (src/sqp.cpp:13-24).
template <typename S>
void install_qp_settings(qp_solver::QPSolverSettings<S> &s) {
    s.warm_start = true;
    s.check_termination = 10;
    s.eps_abs = 1e-4;
    s.eps_rel = 1e-4;
    s.max_iter = 100;
    s.adaptive_rho = true;
    s.adaptive_rho_interval = 50;
    s.alpha = 1.6;
}

// Cholesky success test: what Eigen::LLT::info() != NumericalIssue means (src/sqp.cpp:115-122).
template <typename Matrix>
bool is_posdef(const Matrix &H) {
    using S = typename Matrix::Scalar;
    const std::ptrdiff_t n = H.rows();
    std::vector<S> L((size_t)(n * n));
    for (std::ptrdiff_t j = 0; j < n; ++j)
        for (std::ptrdiff_t i = 0; i < n; ++i) L[(size_t)(i + n * j)] = H(i, j);
    for (std::ptrdiff_t k = 0; k < n; ++k) {
        S x = L[(size_t)(k + n * k)];
        for (std::ptrdiff_t j = 0; j < k; ++j) x -= L[(size_t)(k + n * j)] * L[(size_t)(k + n * j)];
        if (x <= S(0)) return false;
        x = std::sqrt(x);
        L[(size_t)(k + n * k)] = x;
        for (std::ptrdiff_t i = k + 1; i < n; ++i) {
            S v = L[(size_t)(i + n * k)];
            for (std::ptrdiff_t j = 0; j < k; ++j) v -= L[(size_t)(i + n * j)] * L[(size_t)(k + n * j)];
            L[(size_t)(i + n * k)] = v / x;
        }
    }
    return true;
}

// One problem instance's outer-loop state and the host-side steps of an SQP iteration. SQP<T> drives one
// of these with a QPSolver<T>; BatchSQP drives B of them with one BatchQPSolver.
template <typename Scalar>
struct Instance {
    using Problem = NonLinearProblem<Scalar>;
    using Matrix = typename Problem::Matrix;
    using Vector = typename Problem::Vector;
    static constexpr Scalar DIV_BY_ZERO_REGUL = std::numeric_limits<Scalar>::epsilon();

    Vector x_, lambda_, step_prev_, grad_L_, delta_grad_L_;
    Matrix Hess_;
    Vector grad_obj_;
    Scalar obj_ = 0;
    Matrix Jac_constr_;
    Vector constr_, l_, u_;
    Vector p, p_lambda;  // search directions (run_solve locals in the reference, src/sqp.cpp:45-46)
    Vector ql, qu;       // bounds of the current QP subproblem
    Scalar dual_step_norm_ = 0, primal_step_norm_ = 0;
    Info info_;
    bool active = true;
    bool verbose = true;      // print "Hessian not positive definite" like the reference (sqp.cpp:172); BatchSQP counts instead
    int hessian_repairs = 0;  // how many times the PD repair of sqp.cpp:171-180 ran
    int last_qp_iter = 0;  // info().iter of this instance's own QPSolver: a setup that fails leaves it stale (qp.cpp:68-71, sqp.cpp:224)

    void init(Problem &prob) {  // src/sqp.cpp:48-66
        const int nx = prob.num_var, nc = prob.num_constr;
        p.resize(nx); p.setZero();
        p_lambda.resize(nc); p_lambda.setZero();
        step_prev_.resize(nx); step_prev_.setZero();
        grad_L_.resize(nx); grad_L_.setZero();
        delta_grad_L_.resize(nx); delta_grad_L_.setZero();
        Hess_.resize(nx, nx);
        grad_obj_.resize(nx);
        Jac_constr_.resize(nc, nx);
        constr_.resize(nc);
        l_.resize(nc);
        u_.resize(nc);
        ql.resize(nc);
        qu.resize(nc);
        info_.qp_solver_iter = 0;
        info_.iter = 0;
        info_.status = MAX_ITER_EXCEEDED;
        active = true;
        last_qp_iter = 0;
    }

    // Linearise and build the QP  min 0.5 p'Hp + g'p  s.t.  l - c <= J p <= u - c   (src/sqp.cpp:139-197)
    void form_qp(Problem &prob) {
        const int nx = prob.num_var, nc = prob.num_constr;
        prob.objective_linearized(x_, grad_obj_, obj_);
        prob.constraint_linearized(x_, Jac_constr_, constr_, l_, u_);
        for (int i = 0; i < nx; ++i) delta_grad_L_(i) = -grad_L_(i);
        for (int j = 0; j < nx; ++j) {
            Scalar acc = 0;
            for (int i = 0; i < nc; ++i) acc += Jac_constr_(i, j) * lambda_(i);
            grad_L_(j) = grad_obj_(j) + acc;
        }
        if (info_.iter == 1) {
            Hess_.setIdentity();
        } else {
            for (int i = 0; i < nx; ++i) delta_grad_L_(i) += grad_L_(i);
            BFGS_update(Hess_, step_prev_, delta_grad_L_);
        }
        if (!is_posdef(Hess_)) {
            if (verbose) std::cout << "Hessian not positive definite\n";
            ++hessian_repairs;
            Scalar tau = 1e-3;
            while (!is_posdef(Hess_)) {
                for (int i = 0; i < nx; ++i) Hess_(i, i) += tau;
                tau *= 10;
            }
        }
        for (int i = 0; i < nc; ++i) {
            ql(i) = l_(i) - constr_(i);
            qu(i) = u_(i) - constr_(i);
        }
    }

    // Bounds of the second-order-correction QP: same P, q, A, corrected l and u (src/sqp.cpp:244-276)
    void form_soc_bounds(Problem &prob) {
        const int nx = prob.num_var, nc = prob.num_constr;
        Vector x_step(nx), constr_step(nc);
        for (int i = 0; i < nx; ++i) x_step(i) = x_(i) + p(i);
        prob.constraint(x_step, constr_step, l_, u_);
        for (int i = 0; i < nc; ++i) {
            Scalar Ap = 0;
            for (int j = 0; j < nx; ++j) Ap += Jac_constr_(i, j) * p(j);
            const Scalar d = constr_step(i) - Ap;
            ql(i) = l_(i) - d;
            qu(i) = u_(i) - d;
        }
    }

    Scalar constraint_norm(const Vector &constr, const Vector &l, const Vector &u) const {  // src/sqp.cpp:310-318
        Scalar c_l1 = DIV_BY_ZERO_REGUL, a = 0, b = 0;
        for (std::ptrdiff_t i = 0; i < constr.rows(); ++i) a += std::max<Scalar>(l(i) - constr(i), 0);
        for (std::ptrdiff_t i = 0; i < constr.rows(); ++i) b += std::max<Scalar>(constr(i) - u(i), 0);
        c_l1 += a;
        c_l1 += b;
        return c_l1;
    }
    Scalar constraint_norm(const Vector &x, Problem &prob) {  // src/sqp.cpp:320-327
        prob.constraint(x, constr_, l_, u_);
        return constraint_norm(constr_, l_, u_);
    }
    Scalar max_constraint_violation(const Vector &x, Problem &prob) {  // src/sqp.cpp:329-344
        Scalar c_max = 0;
        prob.constraint(x, constr_, l_, u_);
        if (prob.num_constr > 0) {
            Scalar a = -std::numeric_limits<Scalar>::infinity(), b = a;
            for (int i = 0; i < prob.num_constr; ++i) a = std::max(a, l_(i) - constr_(i));
            for (int i = 0; i < prob.num_constr; ++i) b = std::max(b, constr_(i) - u_(i));
            c_max = std::fmax(c_max, a);
            c_max = std::fmax(c_max, b);
        }
        return c_max;
    }

    // l1-merit backtracking line search along p (src/sqp.cpp:278-308)
    Scalar line_search(Problem &prob, const sqp_settings_t<Scalar> &settings) {
        const int nx = prob.num_var;
        const Scalar constr_l1 = constraint_norm(constr_, l_, u_);
        Scalar gp = 0, pHp = 0;
        for (int i = 0; i < nx; ++i) gp += grad_obj_(i) * p(i);
        for (int i = 0; i < nx; ++i) {
            Scalar acc = 0;
            for (int j = 0; j < nx; ++j) acc += Hess_(i, j) * p(j);
            pHp += p(i) * acc;
        }
        const Scalar mu = (gp + Scalar(0.5) * pHp) / ((1 - settings.rho) * constr_l1);
        const Scalar phi_l1 = obj_ + mu * constr_l1;
        const Scalar Dp_phi_l1 = gp - mu * constr_l1;
        Scalar alpha = 1.0;
        Vector x_step(nx);
        for (int i = 1; i < settings.line_search_max_iter; i++) {
            Scalar obj_step;
            for (int k = 0; k < nx; ++k) x_step(k) = x_(k) + alpha * p(k);
            prob.objective(x_step, obj_step);
            const Scalar phi_l1_step = obj_step + mu * constraint_norm(x_step, prob);
            if (phi_l1_step <= phi_l1 + alpha * settings.eta * Dp_phi_l1) break;
            alpha = settings.tau * alpha;
        }
        return alpha;
    }

    // p_lambda -= lambda; line search; step; step norms; termination (src/sqp.cpp:76-96). Returns true when done.
    bool finish_iteration(Problem &prob, const sqp_settings_t<Scalar> &settings) {
        const int nx = prob.num_var, nc = prob.num_constr;
        for (int i = 0; i < nc; ++i) p_lambda(i) -= lambda_(i);
        const Scalar alpha = line_search(prob, settings);
        Scalar pn = 0, dn = 0;
        for (int i = 0; i < nx; ++i) {
            x_(i) = x_(i) + alpha * p(i);
            step_prev_(i) = alpha * p(i);
            pn = std::max(pn, std::abs(p(i)));
        }
        for (int i = 0; i < nc; ++i) {
            lambda_(i) = lambda_(i) + alpha * p_lambda(i);
            dn = std::max(dn, std::abs(p_lambda(i)));
        }
        primal_step_norm_ = alpha * pn;
        dual_step_norm_ = alpha * dn;
        return primal_step_norm_ <= settings.eps_prim && dual_step_norm_ <= settings.eps_dual &&
               max_constraint_violation(x_, prob) <= settings.eps_prim;  // src/sqp.cpp:124-131
    }
};

}  // namespace detail

/*
 * minimize     f(x)
 * subject to   l <= c(x) <= u
 */
template <typename Scalar_>
class SQP {
   public:
    using Scalar = Scalar_;
    using Matrix = sqpb200_dense::Matrix<Scalar>;
    using Vector = sqpb200_dense::Vector<Scalar>;
    using Problem = NonLinearProblem<Scalar>;
    using Settings = sqp_settings_t<Scalar>;

    static constexpr Scalar DIV_BY_ZERO_REGUL = std::numeric_limits<Scalar>::epsilon();

    SQP() { detail::install_qp_settings(qp_solver_.settings()); }
    ~SQP() = default;

    void solve(Problem &prob, const Vector &x0, const Vector &lambda0) {
        st_.x_ = x0;
        st_.lambda_ = lambda0;
        run_solve(prob);
    }
    void solve(Problem &prob) {
        st_.x_.resize(prob.num_var); st_.x_.setZero();
        st_.lambda_.resize(prob.num_constr); st_.lambda_.setZero();
        run_solve(prob);
    }

    inline const Vector &primal_solution() const { return st_.x_; }
    inline Vector &primal_solution() { return st_.x_; }
    inline const Vector &dual_solution() const { return st_.lambda_; }
    inline Vector &dual_solution() { return st_.lambda_; }
    inline const Settings &settings() const { return settings_; }
    inline Settings &settings() { return settings_; }
    inline const Info &info() const { return st_.info_; }
    inline Info &info() { return st_.info_; }
    qp_solver::QPSolver<Scalar> &qp_solver() { return qp_solver_; }

    void run_solve(Problem &prob) {  // src/sqp.cpp:43-101
        st_.init(prob);
        if (settings_.iteration_callback) settings_.iteration_callback(*this);
        int &iter = st_.info_.iter;
        for (iter = 1; iter <= settings_.max_iter; iter++) {
            st_.form_qp(prob);
            run_solve_qp(st_.Hess_, st_.grad_obj_, st_.Jac_constr_, st_.ql, st_.qu, st_.p, st_.p_lambda);
            if (settings_.second_order_correction) {
                st_.form_soc_bounds(prob);
                run_solve_qp(st_.Hess_, st_.grad_obj_, st_.Jac_constr_, st_.ql, st_.qu, st_.p, st_.p_lambda);
            }
            const bool done = st_.finish_iteration(prob, settings_);
            if (settings_.iteration_callback) settings_.iteration_callback(*this);
            if (done) {
                st_.info_.status = SOLVED;
                break;
            }
        }
        if (iter > settings_.max_iter) st_.info_.status = MAX_ITER_EXCEEDED;
    }

    bool run_solve_qp(const Matrix &P, const Vector &q, const Matrix &A, const Vector &l, const Vector &u, Vector &prim,
                      Vector &dual) {  // src/sqp.cpp:210-242
        qp_solver::QuadraticProblem<Scalar> qp_;
        qp_.P = &P;
        qp_.q = &q;
        qp_.A = &A;
        qp_.l = &l;
        qp_.u = &u;
        qp_solver_.setup(qp_);
        qp_solver_.solve(qp_);
        st_.info_.qp_solver_iter += qp_solver_.info().iter;
        if (qp_solver_.info().status == qp_solver::NUMERICAL_ISSUES) {
            std::cout << "QPSolver NUMERICAL_ISSUES\n";
            return false;
        }
        prim = qp_solver_.primal_solution();
        dual = qp_solver_.dual_solution();
        return true;
    }

    // Solver state (public like the reference's, which comments out "private:" at sqp.hpp:116)
    detail::Instance<Scalar> st_;
    Settings settings_;
    qp_solver::QPSolver<Scalar> qp_solver_;
};

/** B SQP instances in lock-step; the QP subproblems of one outer iteration are batched GPU solves.
 *  The instances are split into (up to) two groups, each with its own BatchQPSolver and stream: while the GPU solves one group's QPs
 *  the host forms the other group's QPs or runs its line searches, so host and device work overlap (each instance's trajectory is
 *  independent of the grouping). */
class BatchSQP {
   public:
    using Scalar = double;
    using Problem = NonLinearProblem<double>;
    using Vector = Problem::Vector;
    using Settings = sqp_settings_t<double>;

    /** All problems must share num_var and num_constr. The problems are not owned. `groups`: 0 = automatic (two groups from 512
     *  instances on), 1 = a single group (one batched solve per outer iteration). */
    explicit BatchSQP(const std::vector<Problem *> &problems, int device = 0, int groups = 0)
        : probs_(problems), nx_(problems.at(0)->num_var), nc_(problems.at(0)->num_constr), inst_(problems.size()) {
        for (auto *p : probs_)
            if (p->num_var != nx_ || p->num_constr != nc_) throw std::invalid_argument("BatchSQP: problems differ in size");
        const size_t B = probs_.size();
        const int ng = groups > 0 ? (groups > 2 ? 2 : groups) : (B >= 512 ? 2 : 1);
        for (int g = 0; g < ng; ++g) {
            Group G;
            G.lo = B * g / ng;
            G.hi = B * (g + 1) / ng;
            G.qp.reset(new qp_solver::BatchQPSolver((int)(G.hi - G.lo), nx_, nc_, device));
            detail::install_qp_settings(G.qp->settings());
            G.qp->fetch_full_info(false);  // the outer loop reads x, y, status and iter only (sqp.cpp:224-239)
            // the packed QP arrays ARE the solver's page-locked staging buffers: one asynchronous copy per array per outer iteration
            G.P = G.qp->staged_P(); G.q = G.qp->staged_q(); G.A = G.qp->staged_A(); G.l = G.qp->staged_l(); G.u = G.qp->staged_u();
            G.slot.resize(G.hi - G.lo);
            groups_.push_back(std::move(G));
        }
    }

    Settings &settings() { return settings_; }
    qp_solver::BatchQPSolver &qp_solver(int group = 0) { return *groups_[group].qp; }
    size_t size() const { return probs_.size(); }
    const Vector &primal_solution(size_t i) const { return inst_[i].x_; }
    const Vector &dual_solution(size_t i) const { return inst_[i].lambda_; }
    const Info &info(size_t i) const { return inst_[i].info_; }
    int qp_launches() const { return launches_; }
    // wall-clock seconds of the last solve() spent in: [0] forming and packing the QPs (host), [1] WAITING for the batched GPU QP calls
    // (whatever of staging + launch + read-back did not hide behind host work), [2] unpacking the results (host), [3] line search /
    // step / termination (host)
    const double *phase_seconds() const { return phase_; }
    long long hessian_repairs() const {
        long long t = 0;
        for (const auto &I : inst_) t += I.hessian_repairs;
        return t;
    }

    void solve(const std::vector<Vector> &x0, const std::vector<Vector> &lambda0) {
        const size_t B = probs_.size();
        for (size_t i = 0; i < B; ++i) {
            inst_[i].x_ = x0[i];
            inst_[i].lambda_ = lambda0[i];
            inst_[i].init(*probs_[i]);
            inst_[i].verbose = false;  // thousands of instances: the repair message of sqp.cpp:172 is counted, not printed
            inst_[i].hessian_repairs = 0;
        }
        launches_ = 0;
        for (double &t : phase_) t = 0;
        // Each group is its own little state machine (form + launch -> wait -> line search -> form + launch ...) and the host
        // alternates between them: while the GPU solves one group's QPs the host runs the other group's line searches and forms
        // its next QPs, so with two groups the GPU latency of an outer iteration hides behind host work.
        for (auto &G : groups_) {
            G.iter = 1;
            G.in_flight = start(G);
        }
        for (bool any = true; any;) {
            any = false;
            for (auto &G : groups_) {
                if (G.in_flight) {
                    finish(G);
                    G.in_flight = (++G.iter <= settings_.max_iter) && start(G);
                }
                any = any || G.in_flight;
            }
        }
        for (auto &I : inst_)
            if (I.active) {
                I.info_.iter = settings_.max_iter + 1;
                I.info_.status = MAX_ITER_EXCEEDED;
                I.active = false;
            }
    }

   private:
    struct Group {
        size_t lo = 0, hi = 0;  // instances [lo, hi)
        std::unique_ptr<qp_solver::BatchQPSolver> qp;
        double *P = nullptr, *q = nullptr, *A = nullptr, *l = nullptr, *u = nullptr;  // packed QP arrays, owned by qp
        std::vector<int> slot;
        int na = 0, iter = 0;
        bool in_flight = false;
    };
    typedef std::chrono::steady_clock::time_point tp;
    static tp now() { return std::chrono::steady_clock::now(); }
    static double secs(tp a, tp b) { return std::chrono::duration<double>(b - a).count(); }

    // form the QPs of the still-active instances of a group and hand them to the GPU (asynchronous); false when none is left
    bool start(Group &G) {
        const auto t0 = now();
        G.na = 0;
        for (size_t i = G.lo; i < G.hi; ++i)
            if (inst_[i].active) G.slot[G.na++] = (int)i;  // compact into the leading slots of the packed QP arrays
        if (G.na == 0) return false;
        const int iter = G.iter;
        // host side of the outer iteration: independent per instance (user callbacks must be re-entrant across
        // DIFFERENT problem objects when built with OpenMP)
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
        for (int k = 0; k < G.na; ++k) {
            auto &I = inst_[G.slot[k]];
            I.info_.iter = iter;
            I.form_qp(*probs_[G.slot[k]]);
            pack(G, k, I, true);
        }
        phase_[0] += secs(t0, now());
        G.qp->setup_solve_staged_async(G.na, settings_.second_order_correction ? SQPB200_KEEP_FACTOR : 0u);
        ++launches_;
        return true;
    }
    // collect the steps of a group, optionally re-solve with the second-order-corrected bounds, line search, step, termination
    void finish(Group &G) {
        auto t0 = now();
        G.qp->wait();
        auto t1 = now();
        phase_[1] += secs(t0, t1);
        unpack(G);
        phase_[2] += secs(t1, now());
        if (settings_.second_order_correction) {
            t0 = now();
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
            for (int k = 0; k < G.na; ++k) {
                auto &I = inst_[G.slot[k]];
                I.form_soc_bounds(*probs_[G.slot[k]]);
                pack(G, k, I, false);  // only l and u change (the TODO at src/sqp.cpp:273)
            }
            t1 = now();
            phase_[0] += secs(t0, t1);
            // same P, A: instances with unchanged constraint classes skip the factorisation
            G.qp->setup_solve_staged_async(G.na, SQPB200_REUSE_FACTOR);
            ++launches_;
            G.qp->wait();
            const auto t2 = now();
            phase_[1] += secs(t1, t2);
            unpack(G);
            phase_[2] += secs(t2, now());
        }
        t0 = now();
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
        for (int k = 0; k < G.na; ++k) {
            auto &I = inst_[G.slot[k]];
            if (I.finish_iteration(*probs_[G.slot[k]], settings_)) {
                I.info_.status = SOLVED;
                I.active = false;
            }
        }
        phase_[3] += secs(t0, now());
    }

    void pack(Group &G, int k, const detail::Instance<double> &I, bool all) {
        const size_t nn = (size_t)nx_ * nx_, mn = (size_t)nc_ * nx_;
        if (all) {
            for (int j = 0; j < nx_; ++j)
                for (int i = 0; i < nx_; ++i) G.P[k * nn + i + (size_t)nx_ * j] = I.Hess_(i, j);
            for (int j = 0; j < nx_; ++j)
                for (int i = 0; i < nc_; ++i) G.A[k * mn + i + (size_t)nc_ * j] = I.Jac_constr_(i, j);
            for (int i = 0; i < nx_; ++i) G.q[(size_t)k * nx_ + i] = I.grad_obj_(i);
        }
        for (int i = 0; i < nc_; ++i) {
            G.l[(size_t)k * nc_ + i] = I.ql(i);
            G.u[(size_t)k * nc_ + i] = I.qu(i);
        }
    }
    void unpack(Group &G) {  // the tail of run_solve_qp for every active instance (src/sqp.cpp:224-239)
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
        for (int k = 0; k < G.na; ++k) {
            auto &I = inst_[G.slot[k]];
            const auto qi = G.qp->info(k);
            // A QP whose setup fails is not solved (qp.cpp:68-71), so the reference adds its solver's STALE info().iter (sqp.cpp:224):
            // the count of this instance's previous subproblem -- not whatever instance last occupied batch slot k.
            if (qi.status != qp_solver::NUMERICAL_ISSUES) I.last_qp_iter = qi.iter;
            I.info_.qp_solver_iter += I.last_qp_iter;
            if (qi.status == qp_solver::NUMERICAL_ISSUES) continue;  // keep the stale step, like the reference
            const double *xs = G.qp->primal_solution(k), *ys = G.qp->dual_solution(k);
            for (int i = 0; i < nx_; ++i) I.p(i) = xs[i];
            for (int i = 0; i < nc_; ++i) I.p_lambda(i) = ys[i];
        }
    }

    std::vector<Problem *> probs_;
    int nx_, nc_;
    std::vector<Group> groups_;
    std::vector<detail::Instance<double>> inst_;
    Settings settings_;
    int launches_ = 0;
    double phase_[4] = {0, 0, 0, 0};
};

}  // namespace sqp
