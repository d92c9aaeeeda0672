// ORACLE/_ref (test infrastructure, NOT product code): C entry points around the REFERENCE'S OWN solver classes.
//
// This file is compiled together with /root/reference/src/qp.cpp and /root/reference/src/sqp.cpp (unmodified, from where they
// lie) against oracle/eigen_lite (the stand-in for the absent Eigen dependency) into oracle/_ref/libsqp_ref.so -- see
// oracle/Makefile, target `ref`. Every line of the ADMM loop (src/qp.cpp:64-157) and of the SQP outer loop (src/sqp.cpp:43-308)
// that runs behind these entry points is the reference's; only what Eigen would have supplied (LDLT, LLT, products, reductions)
// comes from eigen_lite. tests/test_reference_build.py holds the C restatement (oracle/qp_oracle_impl.h, oracle/sqp_oracle.c) to
// bit identity with this build, which is what pins the oracle; bench.py may time it as the CPU baseline.
//
// The struct layouts are those of oracle/qp_oracle_impl.h and oracle/sqp_oracle.c so that the same ctypes definitions serve both.
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <solvers/qp.hpp>
#include <solvers/sqp.hpp>

extern "C" {
struct ref_settings_f64 {
    double rho, sigma, alpha, eps_rel, eps_abs;
    int max_iter, check_termination, warm_start, adaptive_rho;
    double adaptive_rho_tolerance;
    int adaptive_rho_interval, verbose;
};
struct ref_settings_f32 {
    float rho, sigma, alpha, eps_rel, eps_abs;
    int max_iter, check_termination, warm_start, adaptive_rho;
    float adaptive_rho_tolerance;
    int adaptive_rho_interval, verbose;
};
struct ref_info_f64 {
    int status, iter, rho_updates;
    double rho_estimate, res_prim, res_dual;
};
struct ref_info_f32 {
    int status, iter, rho_updates;
    float rho_estimate, res_prim, res_dual;
};
// oracle/sqp_oracle.c
struct sqp_problem {
    int num_var, num_constr;
    void (*objective)(const sqp_problem *, const double *x, double *obj);
    void (*objective_linearized)(const sqp_problem *, const double *x, double *grad, double *obj);
    void (*constraint)(const sqp_problem *, const double *x, double *c, double *l, double *u);
    void (*constraint_linearized)(const sqp_problem *, const double *x, double *Jc, double *c, double *l, double *u);
    double par[4];
};
int oracle_sqp_make_problem(int id, int n, sqp_problem *p);
struct sqp_settings {
    double tau, eta, rho, eps_prim, eps_dual;
    int max_iter, line_search_max_iter, second_order_correction;
};
struct sqp_info {
    int iter, qp_solver_iter, status;
};
// per OUTER iteration (the reference exposes its state to settings.iteration_callback): iterate, cumulative ADMM iterations,
// status and iteration count of the iteration's last QP
struct ref_sqp_trace {
    int cap, count, nx, nc;
    double *x, *lambda, *qp_x;
    int *qp_solver_iter, *qp_status, *qp_iter;
};
}

namespace {

template <typename S, typename Settings>
void install(qp_solver::QPSolver<S> &solver, const Settings &s) {
    auto &t = solver.settings();
    t.rho = s.rho; t.sigma = s.sigma; t.alpha = s.alpha; t.eps_rel = s.eps_rel; t.eps_abs = s.eps_abs;
    t.max_iter = s.max_iter; t.check_termination = s.check_termination; t.warm_start = s.warm_start != 0;
    t.adaptive_rho = s.adaptive_rho != 0; t.adaptive_rho_tolerance = s.adaptive_rho_tolerance;
    t.adaptive_rho_interval = s.adaptive_rho_interval; t.verbose = s.verbose != 0;
}

template <typename S>
struct Problem {
    using Matrix = Eigen::Matrix<S, Eigen::Dynamic, Eigen::Dynamic>;
    using Vector = Eigen::Matrix<S, Eigen::Dynamic, 1>;
    Matrix P, A;
    Vector q, l, u;
    qp_solver::QuadraticProblem<S> qp;
    void load(int n, int m, const S *P_, const S *q_, const S *A_, const S *l_, const S *u_) {
        P.resize(n, n); A.resize(m, n); q.resize(n); l.resize(m); u.resize(m);
        std::memcpy(P.data(), P_, sizeof(S) * (size_t)n * n);
        if (m * n) std::memcpy(A.data(), A_, sizeof(S) * (size_t)m * n);
        std::memcpy(q.data(), q_, sizeof(S) * (size_t)n);
        if (m) { std::memcpy(l.data(), l_, sizeof(S) * (size_t)m); std::memcpy(u.data(), u_, sizeof(S) * (size_t)m); }
        qp.P = &P; qp.q = &q; qp.A = &A; qp.l = &l; qp.u = &u;
    }
};

// one reference solver object + the problem it currently looks at (QuadraticProblem holds non-owning pointers, qp.hpp:29-33)
template <typename S>
struct Handle {
    qp_solver::QPSolver<S> solver;
    Problem<S> prob;
};

template <typename S, typename Settings, typename Info>
int solve_batch(const Settings *settings, int batch, int n, int m, const S *P, const S *q, const S *A, const S *l, const S *u, S *x,
                S *y, int *status, int *iter, S *res_prim, S *res_dual, int *rho_updates, S *rho_estimate, int nthreads) {
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#pragma omp parallel num_threads(nthreads)
#endif
    {
        Problem<S> pr;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int b = 0; b < batch; ++b) {
            qp_solver::QPSolver<S> solver;  // a fresh solver per QP: the pattern of sqp.cpp:210-242 with default-constructed state
            install(solver, *settings);
            pr.load(n, m, P + (size_t)b * n * n, q + (size_t)b * n, A + (size_t)b * m * n, l + (size_t)b * m, u + (size_t)b * m);
            solver.setup(pr.qp);
            solver.solve(pr.qp);
            const auto &info = solver.info();
            if (x) for (int i = 0; i < n; ++i) x[(size_t)b * n + i] = info.status == qp_solver::NUMERICAL_ISSUES ? S(0) : solver.primal_solution()(i);
            if (y) for (int i = 0; i < m; ++i) y[(size_t)b * m + i] = info.status == qp_solver::NUMERICAL_ISSUES ? S(0) : solver.dual_solution()(i);
            if (status) status[b] = info.status;
            if (iter) iter[b] = info.iter;
            if (res_prim) res_prim[b] = info.res_prim;
            if (res_dual) res_dual[b] = info.res_dual;
            if (rho_updates) rho_updates[b] = info.rho_updates;
            if (rho_estimate) rho_estimate[b] = info.rho_estimate;
        }
    }
    return used;
}

// adapter: the oracle's C problem definitions (hand-derived gradients of the reference's test problems) as a NonLinearProblem
struct CProblem : sqp::NonLinearProblem<double> {
    sqp_problem p;
    explicit CProblem(const sqp_problem &p_) : p(p_) {
        num_var = p.num_var;
        num_constr = p.num_constr;
    }
    void objective(const Vector &x, Scalar &obj) override { p.objective(&p, x.data(), &obj); }
    void objective_linearized(const Vector &x, Vector &grad, Scalar &obj) override {
        grad.resize(num_var);
        p.objective_linearized(&p, x.data(), grad.data(), &obj);
    }
    void constraint(const Vector &x, Vector &c, Vector &l, Vector &u) override { p.constraint(&p, x.data(), c.data(), l.data(), u.data()); }
    void constraint_linearized(const Vector &x, Matrix &Jc, Vector &c, Vector &l, Vector &u) override {
        Jc.resize(num_constr, num_var);
        p.constraint_linearized(&p, x.data(), Jc.data(), c.data(), l.data(), u.data());
    }
};

}  // namespace

extern "C" {

const char *ref_describe(void) {
    return "reference src/qp.cpp + src/sqp.cpp (unmodified) compiled against oracle/eigen_lite (stand-in for Eigen: LDLT/LLT/products restated)";
}

#define REF_QP_API(SUF, S, SETTINGS, INFO)                                                                                          \
    void *ref_qp_new##SUF(void) { return new Handle<S>(); }                                                                         \
    void ref_qp_free##SUF(void *h) { delete static_cast<Handle<S> *>(h); }                                                          \
    void ref_qp_set_settings##SUF(void *h, const SETTINGS *s) { install(static_cast<Handle<S> *>(h)->solver, *s); }                 \
    void ref_qp_setup##SUF(void *h_, int n, int m, const S *P, const S *q, const S *A, const S *l, const S *u) {                    \
        auto *h = static_cast<Handle<S> *>(h_);                                                                                     \
        h->prob.load(n, m, P, q, A, l, u);                                                                                          \
        h->solver.setup(h->prob.qp);                                                                                                \
    }                                                                                                                               \
    void ref_qp_update_qp##SUF(void *h_, int n, int m, const S *P, const S *q, const S *A, const S *l, const S *u) {                \
        auto *h = static_cast<Handle<S> *>(h_);                                                                                     \
        h->prob.load(n, m, P, q, A, l, u);                                                                                          \
        h->solver.update_qp(h->prob.qp);                                                                                            \
    }                                                                                                                               \
    void ref_qp_solve##SUF(void *h_, int n, int m, const S *P, const S *q, const S *A, const S *l, const S *u) {                    \
        auto *h = static_cast<Handle<S> *>(h_);                                                                                     \
        h->prob.load(n, m, P, q, A, l, u);                                                                                          \
        h->solver.solve(h->prob.qp);                                                                                                \
    }                                                                                                                               \
    void ref_qp_get##SUF(void *h_, S *x, S *y, INFO *info) {                                                                        \
        auto *h = static_cast<Handle<S> *>(h_);                                                                                     \
        const auto &xs = h->solver.primal_solution();                                                                               \
        const auto &ys = h->solver.dual_solution();                                                                                 \
        if (x) for (Eigen::Index i = 0; i < xs.rows(); ++i) x[i] = xs(i);                                                           \
        if (y) for (Eigen::Index i = 0; i < ys.rows(); ++i) y[i] = ys(i);                                                           \
        if (info) {                                                                                                                 \
            const auto &fi = h->solver.info();                                                                                      \
            info->status = fi.status; info->iter = fi.iter; info->rho_updates = fi.rho_updates;                                     \
            info->rho_estimate = fi.rho_estimate; info->res_prim = fi.res_prim; info->res_dual = fi.res_dual;                       \
        }                                                                                                                           \
    }                                                                                                                               \
    int ref_qp_solve_batch##SUF(const SETTINGS *settings, int batch, int n, int m, const S *P, const S *q, const S *A, const S *l,  \
                                const S *u, S *x, S *y, int *status, int *iter, S *res_prim, S *res_dual, int *rho_updates,         \
                                S *rho_estimate, int nthreads) {                                                                    \
        return solve_batch<S, SETTINGS, INFO>(settings, batch, n, m, P, q, A, l, u, x, y, status, iter, res_prim, res_dual,         \
                                              rho_updates, rho_estimate, nthreads);                                                 \
    }

REF_QP_API(_f64, double, ref_settings_f64, ref_info_f64)
REF_QP_API(_f32, float, ref_settings_f32, ref_info_f32)

void ref_constr_type_init_f64(const double *l, const double *u, int m, int *constr_type) {  // static QPSolver::constr_type_init, qp.cpp:283-294
    Eigen::VectorXd lv(m), uv(m);
    Eigen::VectorXi t(m);
    for (int i = 0; i < m; ++i) { lv(i) = l[i]; uv(i) = u[i]; }
    qp_solver::QPSolver<double>::constr_type_init(lv, uv, t);
    for (int i = 0; i < m; ++i) constr_type[i] = t(i);
}

// SQP<double>::solve(prob, x0, lambda0) on a built-in test problem (problem callbacks shared with oracle/sqp_oracle.c)
int ref_sqp_solve_builtin(int prob_id, int n, const sqp_settings *settings, const double *x0, const double *lambda0, double *x_out,
                          double *lambda_out, sqp_info *info_out, ref_sqp_trace *trace) {
    sqp_problem cp;
    if (oracle_sqp_make_problem(prob_id, n, &cp)) return 1;
    CProblem prob(cp);
    sqp::SQP<double> solver;
    auto &st = solver.settings();
    st.tau = settings->tau; st.eta = settings->eta; st.rho = settings->rho; st.eps_prim = settings->eps_prim;
    st.eps_dual = settings->eps_dual; st.max_iter = settings->max_iter; st.line_search_max_iter = settings->line_search_max_iter;
    st.second_order_correction = settings->second_order_correction != 0;
    const int nx = cp.num_var, nc = cp.num_constr;
    if (trace) {
        trace->count = 0; trace->nx = nx; trace->nc = nc;
        bool first = true;
        st.iteration_callback = [trace, nx, nc, first](sqp::SQP<double> &s) mutable {
            if (first) { first = false; return; }  // the call before the loop (sqp.cpp:68-70)
            if (trace->count >= trace->cap) return;
            const int k = trace->count++;
            for (int i = 0; i < nx; ++i) trace->x[(size_t)k * nx + i] = s.x_(i);
            for (int i = 0; i < nc; ++i) trace->lambda[(size_t)k * nc + i] = s.lambda_(i);
            for (int i = 0; i < nx; ++i) trace->qp_x[(size_t)k * nx + i] = s.qp_solver_.primal_solution()(i);
            trace->qp_solver_iter[k] = s.info_.qp_solver_iter;
            trace->qp_status[k] = s.qp_solver_.info().status;
            trace->qp_iter[k] = s.qp_solver_.info().iter;
        };
    }
    Eigen::VectorXd xv(nx), lv(nc);
    for (int i = 0; i < nx; ++i) xv(i) = x0[i];
    for (int i = 0; i < nc; ++i) lv(i) = lambda0[i];
    solver.solve(prob, xv, lv);
    for (int i = 0; i < nx; ++i) x_out[i] = solver.primal_solution()(i);
    for (int i = 0; i < nc; ++i) lambda_out[i] = solver.dual_solution()(i);
    info_out->iter = solver.info().iter;
    info_out->qp_solver_iter = solver.info().qp_solver_iter;
    info_out->status = solver.info().status;
    return 0;
}

int ref_num_procs(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

}  // extern "C"
