/*
 * ORACLE (test infrastructure, NOT product code) -- CPU restatement of the reference SQP outer loop,
 * the CALLER of the QP hot path.
 *
 * Follows /root/reference/src/sqp.cpp (SQP<T>::SQP :13-24, run_solve :43-101, is_posdef :115-122,
 * termination_criteria :124-131, solve_qp :139-208, run_solve_qp :210-242, second_order_correction
 * :244-276, line_search :278-308, constraint_norm :310-327, max_constraint_violation :329-344) and
 * /root/reference/include/solvers/bfgs.hpp:15-41 (damped BFGS). Every QP is solved by the QP oracle
 * of qp_oracle.c. Eigen::LLT (is_posdef) is restated as an unblocked Cholesky that fails on a
 * pivot <= 0 (Eigen/src/Cholesky/LLT.h, llt_inplace<Lower>::unblocked).
 *
 * The reference's NonLinearProblem is a class with four virtuals (sqp.hpp:62-76); here a problem is a
 * struct of four function pointers. The test problems of tests/sqp_test.cpp and
 * tests/sqp_test_autodiff.cpp are built in with hand-derived gradients (the reference uses
 * Eigen::AutoDiffScalar, which gives the same derivatives up to rounding).
 *
 * PARITY PINNING: the reference's own assertions for these problems (solution within isApprox 1e-2, iter < max_iter), and BIT
 * IDENTITY of whole trajectories (outer and inner iteration counts, iterates, per-iteration ADMM totals) with oracle/_ref -- the
 * reference's own src/sqp.cpp + src/qp.cpp compiled unmodified against oracle/eigen_lite (tests/test_reference_build.py).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- QP oracle (qp_oracle.c) ---------------------------------------------------------------- */
typedef struct {
    double rho, sigma, alpha, eps_rel, eps_abs;
    int max_iter, check_termination, warm_start, adaptive_rho;
    double adaptive_rho_tolerance;
    int adaptive_rho_interval, verbose;
} qp_settings_t;
typedef struct {
    int status, iter, rho_updates;
    double rho_estimate, res_prim, res_dual;
} qp_info_t;
void *oracle_qp_new_f64(void);
void oracle_qp_free_f64(void *);
qp_settings_t *oracle_qp_settings_f64(void *);
qp_info_t *oracle_qp_info_f64(void *);
double *oracle_qp_primal_f64(void *);
double *oracle_qp_dual_f64(void *);
void oracle_qp_setup_f64(void *, int, int, const double *, const double *, const double *, const double *, const double *);
void oracle_qp_solve_f64(void *, const double *, const double *, const double *, const double *, const double *);

enum { SQP_SOLVED = 0, SQP_MAX_ITER_EXCEEDED = 1, SQP_INVALID_SETTINGS = 2 }; /* sqp.hpp:33 */
#define QP_NUMERICAL_ISSUES 3

/* ---- problem interface (sqp.hpp:62-76) ------------------------------------------------------- */
typedef struct sqp_problem {
    int num_var, num_constr;
    void (*objective)(const struct sqp_problem *, const double *x, double *obj);
    void (*objective_linearized)(const struct sqp_problem *, const double *x, double *grad, double *obj);
    void (*constraint)(const struct sqp_problem *, const double *x, double *c, double *l, double *u);
    /* Jc is num_constr x num_var, column-major */
    void (*constraint_linearized)(const struct sqp_problem *, const double *x, double *Jc, double *c, double *l, double *u);
    double par[4];
} sqp_problem;

/* ---- built-in test problems -------------------------------------------------------------------- */
enum { PROB_CONSTRAINED_ROSENBROCK_2D = 0, PROB_SIMPLE_NLP = 1, PROB_SIMPLE_QP = 2, PROB_SIMPLE_NLP2 = 3, PROB_ROSENBROCK_BOX = 4 };

static void rosen_obj(const sqp_problem *p, const double *x, double *obj) { /* tests/sqp_test_autodiff.cpp:61-71 */
    double z = 0;
    for (int i = 0; i < p->num_var - 1; i++) {
        double a = 1.0 - x[i], b = x[i + 1] - x[i] * x[i];
        z += a * a + 100.0 * b * b;
    }
    *obj = z;
}
static void rosen_lin(const sqp_problem *p, const double *x, double *g, double *obj) {
    int n = p->num_var;
    rosen_obj(p, x, obj);
    for (int i = 0; i < n; i++) g[i] = 0;
    for (int i = 0; i < n - 1; i++) {
        double b = x[i + 1] - x[i] * x[i];
        g[i] += -2.0 * (1.0 - x[i]) - 400.0 * x[i] * b;
        g[i + 1] += 200.0 * b;
    }
}
/* ConstrainedRosenbrock2D, tests/sqp_test_autodiff.cpp:73-99: c = [x0 - x1, x0^2 + x1^2], l = [-inf, 1], u = [0, 1] */
static void crosen_con(const sqp_problem *p, const double *x, double *c, double *l, double *u) {
    (void)p;
    c[0] = x[0] - x[1];
    c[1] = x[0] * x[0] + x[1] * x[1];
    u[0] = 0; u[1] = 1;
    l[0] = -INFINITY; l[1] = 1;
}
static void crosen_con_lin(const sqp_problem *p, const double *x, double *J, double *c, double *l, double *u) {
    crosen_con(p, x, c, l, u);
    J[0] = 1; J[1] = 2 * x[0];  /* column 0 */
    J[2] = -1; J[3] = 2 * x[1]; /* column 1 */
}
/* SimpleNLP, tests/sqp_test.cpp:8-44 */
static void snlp_obj(const sqp_problem *p, const double *x, double *obj) { (void)p; *obj = -(x[0] + x[1]); }
static void snlp_lin(const sqp_problem *p, const double *x, double *g, double *obj) { snlp_obj(p, x, obj); g[0] = -1; g[1] = -1; }
static void snlp_con(const sqp_problem *p, const double *x, double *c, double *l, double *u) {
    (void)p;
    c[0] = x[0] * x[0] + x[1] * x[1]; c[1] = x[0]; c[2] = x[1];
    l[0] = 1; l[1] = 0; l[2] = 0;
    u[0] = 2; u[1] = INFINITY; u[2] = INFINITY;
}
static void snlp_con_lin(const sqp_problem *p, const double *x, double *J, double *c, double *l, double *u) {
    snlp_con(p, x, c, l, u);
    J[0] = 2 * x[0]; J[1] = 1; J[2] = 0;
    J[3] = 2 * x[1]; J[4] = 0; J[5] = 1;
}
/* SimpleQP posed as an NLP, tests/sqp_test.cpp:92-124 */
static void sqpn_obj(const sqp_problem *p, const double *x, double *obj) {
    (void)p;
    double Px0 = 4 * x[0] + 1 * x[1], Px1 = 1 * x[0] + 2 * x[1];
    *obj = 0.5 * (x[0] * Px0 + x[1] * Px1) + (x[0] + x[1]);
}
static void sqpn_lin(const sqp_problem *p, const double *x, double *g, double *obj) {
    sqpn_obj(p, x, obj);
    g[0] = 4 * x[0] + 1 * x[1] + 1;
    g[1] = 1 * x[0] + 2 * x[1] + 1;
}
static void sqpn_con(const sqp_problem *p, const double *x, double *c, double *l, double *u) {
    (void)p;
    c[0] = x[0] + x[1]; c[1] = x[0]; c[2] = x[1];
    l[0] = 1; l[1] = 0; l[2] = 0;
    u[0] = 1; u[1] = 0.7; u[2] = 0.7;
}
static void sqpn_con_lin(const sqp_problem *p, const double *x, double *J, double *c, double *l, double *u) {
    sqpn_con(p, x, c, l, u);
    J[0] = 1; J[1] = 1; J[2] = 0;
    J[3] = 1; J[4] = 0; J[5] = 1;
}
/* SimpleNLP2 (N&W example 12.1), tests/sqp_test_autodiff.cpp:245-265 */
static void snlp2_obj(const sqp_problem *p, const double *x, double *obj) { (void)p; *obj = x[0] + x[1]; }
static void snlp2_lin(const sqp_problem *p, const double *x, double *g, double *obj) { snlp2_obj(p, x, obj); g[0] = 1; g[1] = 1; }
static void snlp2_con(const sqp_problem *p, const double *x, double *c, double *l, double *u) {
    (void)p;
    c[0] = x[0] * x[0] + x[1] * x[1];
    l[0] = 2; u[0] = 2;
}
static void snlp2_con_lin(const sqp_problem *p, const double *x, double *J, double *c, double *l, double *u) {
    snlp2_con(p, x, c, l, u);
    J[0] = 2 * x[0]; J[1] = 2 * x[1];
}
/* Rosenbrock(n) with box constraints, tests/sqp_test_autodiff.cpp:122-145 */
static void rbox_con(const sqp_problem *p, const double *x, double *c, double *l, double *u) {
    for (int i = 0; i < p->num_var; i++) { c[i] = x[i]; u[i] = 1; l[i] = 0; }
}
static void rbox_con_lin(const sqp_problem *p, const double *x, double *J, double *c, double *l, double *u) {
    int n = p->num_var;
    rbox_con(p, x, c, l, u);
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) J[i + n * j] = (i == j) ? 1.0 : 0.0;
}

int oracle_sqp_make_problem(int id, int n, sqp_problem *p) {
    memset(p, 0, sizeof *p);
    switch (id) {
        case PROB_CONSTRAINED_ROSENBROCK_2D:
            p->num_var = 2; p->num_constr = 2;
            p->objective = rosen_obj; p->objective_linearized = rosen_lin;
            p->constraint = crosen_con; p->constraint_linearized = crosen_con_lin;
            return 0;
        case PROB_SIMPLE_NLP:
            p->num_var = 2; p->num_constr = 3;
            p->objective = snlp_obj; p->objective_linearized = snlp_lin;
            p->constraint = snlp_con; p->constraint_linearized = snlp_con_lin;
            return 0;
        case PROB_SIMPLE_QP:
            p->num_var = 2; p->num_constr = 3;
            p->objective = sqpn_obj; p->objective_linearized = sqpn_lin;
            p->constraint = sqpn_con; p->constraint_linearized = sqpn_con_lin;
            return 0;
        case PROB_SIMPLE_NLP2:
            p->num_var = 2; p->num_constr = 1;
            p->objective = snlp2_obj; p->objective_linearized = snlp2_lin;
            p->constraint = snlp2_con; p->constraint_linearized = snlp2_con_lin;
            return 0;
        case PROB_ROSENBROCK_BOX:
            p->num_var = n; p->num_constr = n;
            p->objective = rosen_obj; p->objective_linearized = rosen_lin;
            p->constraint = rbox_con; p->constraint_linearized = rbox_con_lin;
            return 0;
    }
    return 1;
}

#ifndef SQP_ORACLE_PROBLEMS_ONLY /* oracle/_ref links only the problem definitions above (oracle/ref_shim.cpp) */
/* ---- SQP --------------------------------------------------------------------------------------- */
typedef struct { /* sqp_settings_t, sqp.hpp:13-31 */
    double tau, eta, rho, eps_prim, eps_dual;
    int max_iter, line_search_max_iter, second_order_correction;
} sqp_settings;
typedef struct { /* sqp::Info, sqp.hpp:35-60 */
    int iter, qp_solver_iter, status;
} sqp_info;

void oracle_sqp_default_settings(sqp_settings *s) {
    s->tau = 0.5; s->eta = 0.25; s->rho = 0.5; s->eps_prim = 1e-4; s->eps_dual = 1e-4;
    s->max_iter = 100; s->line_search_max_iter = 20; s->second_order_correction = 0;
}

/* optional per-QP trace for subproblem-level parity tests */
typedef struct {
    int cap, count; /* QP solves recorded */
    int nx, nc;
    double *P, *q, *A, *l, *u, *x, *y; /* [cap][...] */
    int *status, *iter;
} sqp_trace;

typedef struct {
    const sqp_problem *prob;
    int nx, nc;
    double *x, *lambda, *step_prev, *grad_L, *delta_grad_L, *Hess, *grad_obj, obj, *Jac, *constr, *l, *u;
    double dual_step_norm, primal_step_norm;
    sqp_settings settings;
    sqp_info info;
    void *qp;
    sqp_trace *trace;
    double *tmp_n, *tmp_c, *ql, *qu;
} sqp_state;

/* Eigen::LLT success test, sqp.cpp:115-122 */
static int is_posdef(const double *H, int n, double *work) {
    memcpy(work, H, sizeof(double) * n * n);
    for (int k = 0; k < n; k++) {
        double x = work[k + n * k];
        for (int j = 0; j < k; j++) x -= work[k + n * j] * work[k + n * j];
        if (x <= 0.0) return 0;
        x = sqrt(x);
        work[k + n * k] = x;
        for (int i = k + 1; i < n; i++) {
            double v = work[i + n * k];
            for (int j = 0; j < k; j++) v -= work[i + n * j] * work[k + n * j];
            work[i + n * k] = v / x;
        }
    }
    return 1;
}

/* Damped BFGS, include/solvers/bfgs.hpp:15-41 */
static void bfgs_update(double *B, int n, const double *s, const double *y, double *Bs, double *r) {
    double sBs = 0, sy = 0, sr;
    for (int i = 0; i < n; i++) {
        double acc = 0;
        for (int j = 0; j < n; j++) acc += B[i + n * j] * s[j];
        Bs[i] = acc;
    }
    for (int i = 0; i < n; i++) { sBs += s[i] * Bs[i]; sy += s[i] * y[i]; }
    if (sy < 0.2 * sBs) {
        double theta = 0.8 * sBs / (sBs - sy);
        for (int i = 0; i < n; i++) r[i] = theta * y[i] + (1 - theta) * Bs[i];
        sr = theta * sy + (1 - theta) * sBs;
    } else {
        for (int i = 0; i < n; i++) r[i] = y[i];
        sr = sy;
    }
    if (sr < DBL_EPSILON) return;
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) B[i + n * j] += -Bs[i] * Bs[j] / sBs + r[i] * r[j] / sr;
}

/* sqp.cpp:310-318 */
static double constraint_norm_v(const double *c, const double *l, const double *u, int nc) {
    double c_l1 = DBL_EPSILON, a = 0, b = 0;
    for (int i = 0; i < nc; i++) { double v = l[i] - c[i]; a += v > 0.0 ? v : 0.0; }
    for (int i = 0; i < nc; i++) { double v = c[i] - u[i]; b += v > 0.0 ? v : 0.0; }
    c_l1 += a;
    c_l1 += b;
    return c_l1;
}
/* sqp.cpp:320-327: uses members constr_, l_, u_ as temporaries */
static double constraint_norm_x(sqp_state *s, const double *x) {
    s->prob->constraint(s->prob, x, s->constr, s->l, s->u);
    return constraint_norm_v(s->constr, s->l, s->u, s->nc);
}
/* sqp.cpp:329-344 */
static double max_constraint_violation(sqp_state *s, const double *x) {
    double c_max = 0;
    s->prob->constraint(s->prob, x, s->constr, s->l, s->u);
    if (s->nc > 0) {
        double a = -INFINITY, b = -INFINITY;
        for (int i = 0; i < s->nc; i++) { double v = s->l[i] - s->constr[i]; if (v > a) a = v; }
        for (int i = 0; i < s->nc; i++) { double v = s->constr[i] - s->u[i]; if (v > b) b = v; }
        c_max = fmax(c_max, a);
        c_max = fmax(c_max, b);
    }
    return c_max;
}

/* sqp.cpp:210-242 */
static int run_solve_qp(sqp_state *s, const double *P, const double *q, const double *A, const double *l, const double *u,
                        double *prim, double *dual) {
    int nx = s->nx, nc = s->nc;
    oracle_qp_setup_f64(s->qp, nx, nc, P, q, A, l, u);
    oracle_qp_solve_f64(s->qp, P, q, A, l, u);
    qp_info_t *qi = oracle_qp_info_f64(s->qp);
    s->info.qp_solver_iter += qi->iter;
    if (s->trace && s->trace->count < s->trace->cap) {
        sqp_trace *t = s->trace;
        int k = t->count++;
        memcpy(t->P + (size_t)k * nx * nx, P, sizeof(double) * nx * nx);
        memcpy(t->q + (size_t)k * nx, q, sizeof(double) * nx);
        memcpy(t->A + (size_t)k * nc * nx, A, sizeof(double) * nc * nx);
        memcpy(t->l + (size_t)k * nc, l, sizeof(double) * nc);
        memcpy(t->u + (size_t)k * nc, u, sizeof(double) * nc);
        memcpy(t->x + (size_t)k * nx, oracle_qp_primal_f64(s->qp), sizeof(double) * nx);
        memcpy(t->y + (size_t)k * nc, oracle_qp_dual_f64(s->qp), sizeof(double) * nc);
        t->status[k] = qi->status;
        t->iter[k] = qi->iter;
    }
    if (qi->status == QP_NUMERICAL_ISSUES) return 0;
    memcpy(prim, oracle_qp_primal_f64(s->qp), sizeof(double) * nx);
    memcpy(dual, oracle_qp_dual_f64(s->qp), sizeof(double) * nc);
    return 1;
}

/* sqp.cpp:244-276 */
static void second_order_correction(sqp_state *s, double *p, double *lambda) {
    int nx = s->nx, nc = s->nc;
    double *x_step = s->tmp_n, *constr_step = s->tmp_c;
    for (int i = 0; i < nx; i++) x_step[i] = s->x[i] + p[i];
    s->prob->constraint(s->prob, x_step, constr_step, s->l, s->u);
    for (int i = 0; i < nc; i++) {
        double Ap = 0;
        for (int j = 0; j < nx; j++) Ap += s->Jac[i + nc * j] * p[j];
        double d = constr_step[i] - Ap;
        s->ql[i] = s->l[i] - d;
        s->qu[i] = s->u[i] - d;
    }
    run_solve_qp(s, s->Hess, s->grad_obj, s->Jac, s->ql, s->qu, p, lambda);
}

/* sqp.cpp:139-208 */
static void solve_qp(sqp_state *s, double *step, double *lambda, double *work) {
    int nx = s->nx, nc = s->nc;
    s->prob->objective_linearized(s->prob, s->x, s->grad_obj, &s->obj);
    s->prob->constraint_linearized(s->prob, s->x, s->Jac, s->constr, s->l, s->u);
    for (int i = 0; i < nx; i++) s->delta_grad_L[i] = -s->grad_L[i];
    for (int j = 0; j < nx; j++) {
        double acc = 0;
        for (int i = 0; i < nc; i++) acc += s->Jac[i + nc * j] * s->lambda[i];
        s->grad_L[j] = s->grad_obj[j] + acc;
    }
    if (s->info.iter == 1) {
        for (int j = 0; j < nx; j++) for (int i = 0; i < nx; i++) s->Hess[i + nx * j] = (i == j) ? 1.0 : 0.0;
    } else {
        for (int i = 0; i < nx; i++) s->delta_grad_L[i] += s->grad_L[i];
        bfgs_update(s->Hess, nx, s->step_prev, s->delta_grad_L, work + nx * nx, work + nx * nx + nx);
    }
    if (!is_posdef(s->Hess, nx, work)) {
        double tau = 1e-3;
        while (!is_posdef(s->Hess, nx, work)) {
            for (int i = 0; i < nx; i++) s->Hess[i + nx * i] += tau;
            tau *= 10;
        }
    }
    for (int i = 0; i < nc; i++) { s->ql[i] = s->l[i] - s->constr[i]; s->qu[i] = s->u[i] - s->constr[i]; }
    run_solve_qp(s, s->Hess, s->grad_obj, s->Jac, s->ql, s->qu, step, lambda);
    if (s->settings.second_order_correction) second_order_correction(s, step, lambda);
}

/* sqp.cpp:278-308 */
static double line_search(sqp_state *s, const double *p) {
    int nx = s->nx;
    const double tau = s->settings.tau;
    double constr_l1 = constraint_norm_v(s->constr, s->l, s->u, s->nc);
    double gp = 0, pHp = 0;
    for (int i = 0; i < nx; i++) gp += s->grad_obj[i] * p[i];
    for (int i = 0; i < nx; i++) {
        double acc = 0;
        for (int j = 0; j < nx; j++) acc += s->Hess[i + nx * j] * p[j];
        pHp += p[i] * acc;
    }
    double mu = (gp + 0.5 * pHp) / ((1 - s->settings.rho) * constr_l1);
    double phi_l1 = s->obj + mu * constr_l1;
    double Dp_phi_l1 = gp - mu * constr_l1;
    double alpha = 1.0;
    double *x_step = s->tmp_n;
    for (int i = 1; i < s->settings.line_search_max_iter; i++) {
        double obj_step;
        for (int k = 0; k < nx; k++) x_step[k] = s->x[k] + alpha * p[k];
        s->prob->objective(s->prob, x_step, &obj_step);
        double phi_l1_step = obj_step + mu * constraint_norm_x(s, x_step);
        if (phi_l1_step <= phi_l1 + alpha * s->settings.eta * Dp_phi_l1) break;
        else alpha = tau * alpha;
    }
    return alpha;
}

/* SQP<T>::solve(prob, x0, lambda0) -> run_solve, sqp.cpp:26-101. Returns final x, lambda, info. */
int oracle_sqp_solve(const sqp_problem *prob, const sqp_settings *settings, const double *x0, const double *lambda0,
                     double *x_out, double *lambda_out, sqp_info *info_out, sqp_trace *trace) {
    sqp_state S;
    memset(&S, 0, sizeof S);
    int nx = prob->num_var, nc = prob->num_constr;
    S.prob = prob; S.nx = nx; S.nc = nc; S.settings = *settings; S.trace = trace;
    size_t nn = (size_t)nx * nx;
    S.x = calloc(nx + 1, 8); S.lambda = calloc(nc + 1, 8); S.step_prev = calloc(nx + 1, 8);
    S.grad_L = calloc(nx + 1, 8); S.delta_grad_L = calloc(nx + 1, 8); S.Hess = calloc(nn + 1, 8);
    S.grad_obj = calloc(nx + 1, 8); S.Jac = calloc((size_t)nc * nx + 1, 8); S.constr = calloc(nc + 1, 8);
    S.l = calloc(nc + 1, 8); S.u = calloc(nc + 1, 8); S.tmp_n = calloc(nx + 1, 8); S.tmp_c = calloc(nc + 1, 8);
    S.ql = calloc(nc + 1, 8); S.qu = calloc(nc + 1, 8);
    double *p = calloc(nx + 1, 8), *p_lambda = calloc(nc + 1, 8), *work = calloc(nn + 2 * nx + 2, 8);
    if (trace) { trace->count = 0; trace->nx = nx; trace->nc = nc; }
    memcpy(S.x, x0, sizeof(double) * nx);
    memcpy(S.lambda, lambda0, sizeof(double) * nc);

    /* SQP<T>::SQP(), sqp.cpp:13-24 */
    S.qp = oracle_qp_new_f64();
    qp_settings_t *qs = oracle_qp_settings_f64(S.qp);
    qs->warm_start = 1; qs->check_termination = 10; qs->eps_abs = 1e-4; qs->eps_rel = 1e-4; qs->max_iter = 100;
    qs->adaptive_rho = 1; qs->adaptive_rho_interval = 50; qs->alpha = 1.6;

    S.info.qp_solver_iter = 0;
    S.info.status = SQP_MAX_ITER_EXCEEDED;
    int iter;
    for (iter = 1; iter <= S.settings.max_iter; iter++) {
        S.info.iter = iter;
        solve_qp(&S, p, p_lambda, work);
        for (int i = 0; i < nc; i++) p_lambda[i] -= S.lambda[i];
        double alpha = line_search(&S, p);
        for (int i = 0; i < nx; i++) S.x[i] = S.x[i] + alpha * p[i];
        for (int i = 0; i < nc; i++) S.lambda[i] = S.lambda[i] + alpha * p_lambda[i];
        double pn = 0, dn = 0;
        for (int i = 0; i < nx; i++) { S.step_prev[i] = alpha * p[i]; if (fabs(p[i]) > pn) pn = fabs(p[i]); }
        for (int i = 0; i < nc; i++) if (fabs(p_lambda[i]) > dn) dn = fabs(p_lambda[i]);
        S.primal_step_norm = alpha * pn;
        S.dual_step_norm = alpha * dn;
        /* termination_criteria, sqp.cpp:124-131 */
        if (S.primal_step_norm <= S.settings.eps_prim && S.dual_step_norm <= S.settings.eps_dual &&
            max_constraint_violation(&S, S.x) <= S.settings.eps_prim) {
            S.info.status = SQP_SOLVED;
            break;
        }
    }
    S.info.iter = iter;
    if (iter > S.settings.max_iter) S.info.status = SQP_MAX_ITER_EXCEEDED;
    memcpy(x_out, S.x, sizeof(double) * nx);
    memcpy(lambda_out, S.lambda, sizeof(double) * nc);
    *info_out = S.info;
    oracle_qp_free_f64(S.qp);
    free(S.x); free(S.lambda); free(S.step_prev); free(S.grad_L); free(S.delta_grad_L); free(S.Hess); free(S.grad_obj);
    free(S.Jac); free(S.constr); free(S.l); free(S.u); free(S.tmp_n); free(S.tmp_c); free(S.ql); free(S.qu);
    free(p); free(p_lambda); free(work);
    return 0;
}

/* convenience for ctypes: built-in problem by id */
int oracle_sqp_solve_builtin(int prob_id, int n, const sqp_settings *settings, const double *x0, const double *lambda0,
                             double *x_out, double *lambda_out, sqp_info *info_out, sqp_trace *trace) {
    sqp_problem p;
    if (oracle_sqp_make_problem(prob_id, n, &p)) return 1;
    return oracle_sqp_solve(&p, settings, x0, lambda0, x_out, lambda_out, info_out, trace);
}

/* B independent SQP solves of one built-in problem from B starting points, OpenMP dynamic schedule over the host cores: the CPU
 * side of BASELINE config 4 (bench.py's cpu_baseline for the batched-SQP record). */
int oracle_sqp_solve_batch(int prob_id, int n, const sqp_settings *settings, int batch, const double *x0, const double *lambda0,
                           double *x_out, double *lambda_out, sqp_info *info_out, int nthreads) {
    sqp_problem p;
    if (oracle_sqp_make_problem(prob_id, n, &p)) return -1;
    const int nx = p.num_var, nc = p.num_constr;
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads)
#endif
    for (int b = 0; b < batch; ++b)
        oracle_sqp_solve(&p, settings, x0 + (size_t)b * nx, lambda0 + (size_t)b * nc, x_out + (size_t)b * nx, lambda_out + (size_t)b * nc,
                         info_out + b, NULL);
    return used;
}

/* standalone BFGS entry for tests/bfgs_test.cpp-style checks */
void oracle_bfgs_update(double *B, int n, const double *s, const double *y) {
    double *w = malloc(sizeof(double) * 2 * (n + 1));
    bfgs_update(B, n, s, y, w, w + n);
    free(w);
}
int oracle_is_posdef(const double *H, int n) {
    double *w = malloc(sizeof(double) * (n * n + 1));
    int r = is_posdef(H, n, w);
    free(w);
    return r;
}
#endif /* SQP_ORACLE_PROBLEMS_ONLY */
