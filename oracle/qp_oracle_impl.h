/*
 * ORACLE (test infrastructure, NOT product code) -- CPU restatement of the reference QP solver.
 *
 * Follows /root/reference/src/qp.cpp line by line (dense branch only; the
 * QP_SOLVER_USE_SPARSE branches do not compile in the reference).
 * The arithmetic the reference delegates to Eigen (Eigen::LDLT<MatrixXd, Lower>,
 * include/solvers/qp.hpp:129; call sites src/qp.cpp:90, :242, :253) is restated from
 * the published Eigen 3.3/3.4 algorithm (Eigen/src/Cholesky/LDLT.h: ldlt_inplace<Lower>::unblocked
 * and LDLT::_solve_impl). Eigen is an un-vendored, unpinned dependency
 * (CMakeLists.txt:12 "find_package(Eigen3 3.3 REQUIRED NO_MODULE)") and is absent from this
 * image, so the reference itself cannot be compiled here.
 *
 * PARITY PINNING: (1) every known-answer assertion the reference's tests hold for this path
 * (tests/qp_solver_test.cpp:43-156, see tests/test_oracle_reference_kats.py); (2) BIT IDENTITY with oracle/_ref --
 * the reference's own src/qp.cpp compiled unmodified against oracle/eigen_lite (a stand-in for the absent Eigen) --
 * over random shapes and settings, the object API, the float instantiation and NaN inputs
 * (tests/test_reference_build.py), plus the committed outputs of that build (tests/golden/reference_outputs.json).
 * Every line of the ADMM loop is thereby pinned to the reference's source; what stays unpinned against a real Eigen
 * build is only the arithmetic inside Eigen's own kernels (LDLT pivot ties, reduction order), restated here and in
 * eigen_lite from the published algorithm.
 *
 * This header is included twice by qp_oracle.c with SCALAR = double and float
 * (the reference instantiates both, src/qp.cpp:385-386).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use this code.
 */

#ifndef SCALAR
#error "define SCALAR, SUFFIX before including"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* include/solvers/qp.hpp:36-53 */
typedef struct {
    SCALAR rho;
    SCALAR sigma;
    SCALAR alpha;
    SCALAR eps_rel;
    SCALAR eps_abs;
    int max_iter;
    int check_termination;
    int warm_start;
    int adaptive_rho;
    SCALAR adaptive_rho_tolerance;
    int adaptive_rho_interval;
    int verbose;
} FN(oracle_settings);

/* include/solvers/qp.hpp:72-79 */
typedef struct {
    int status;
    int iter;
    int rho_updates;
    SCALAR rho_estimate;
    SCALAR res_prim;
    SCALAR res_dual;
} FN(oracle_info);

/* members of QPSolver, include/solvers/qp.hpp:213-250 */
typedef struct {
    int n, m;
    int iter;
    SCALAR *x, *z, *y;
    SCALAR *x_tilde, *z_tilde, *z_prev;
    SCALAR *rho_vec, *rho_inv_vec;
    SCALAR rho;
    SCALAR *rhs, *x_tilde_nu;
    SCALAR max_Ax_z_norm, max_Px_ATy_q_norm;
    int *constr_type;
    FN(oracle_settings) settings;
    FN(oracle_info) info;
    SCALAR *kkt_mat; /* (n+m)^2, column-major like Eigen's default */
    /* Eigen::LDLT state */
    SCALAR *ldlt_mat;
    int *ldlt_transp;
    SCALAR *ldlt_tmp;
    int ldlt_ok;
    /* scratch for mat-vecs */
    SCALAR *tmp_n, *tmp_n2, *tmp_m;
    /* test diagnostics (not in the reference): smallest normalised residuals ever fed to rho_estimate. When one of them is
     * at rounding level the adaptive-rho decision is decided by rounding noise and is not reproducible across implementations. */
    double diag_min_rp_norm, diag_min_rd_norm;
} FN(oracle_solver);

/* include/solvers/qp.hpp:136-141 */
#define ORC_RHO_MIN ((SCALAR)1e-6)
#define ORC_RHO_MAX ((SCALAR)1e+6)
#define ORC_RHO_TOL ((SCALAR)1e-4)
#define ORC_RHO_EQ_FACTOR ((SCALAR)1e+3)
#define ORC_LOOSE_BOUNDS_THRESH ((SCALAR)1e+16)

static void FN(orc_default_settings)(FN(oracle_settings) * s) {
    s->rho = (SCALAR)1e-1;
    s->sigma = (SCALAR)1e-6;
    s->alpha = (SCALAR)1.0;
    s->eps_rel = (SCALAR)1e-3;
    s->eps_abs = (SCALAR)1e-3;
    s->max_iter = 1000;
    s->check_termination = 25;
    s->warm_start = 0;
    s->adaptive_rho = 0;
    s->adaptive_rho_tolerance = (SCALAR)5;
    s->adaptive_rho_interval = 25;
    s->verbose = 0;
}

FN(oracle_solver) * FN(oracle_qp_new)(void) {
    FN(oracle_solver) *s = (FN(oracle_solver) *)calloc(1, sizeof(FN(oracle_solver)));
    FN(orc_default_settings)(&s->settings);
    s->info.status = ORC_UNINITIALIZED; /* qp.hpp:74 */
    s->info.iter = 0;
    s->info.rho_updates = 0;
    s->info.rho_estimate = 0;
    s->info.res_prim = 0;
    s->info.res_dual = 0;
    return s;
}

static void FN(orc_free_bufs)(FN(oracle_solver) * s) {
    free(s->x); free(s->z); free(s->y);
    free(s->x_tilde); free(s->z_tilde); free(s->z_prev);
    free(s->rho_vec); free(s->rho_inv_vec);
    free(s->rhs); free(s->x_tilde_nu);
    free(s->constr_type); free(s->kkt_mat);
    free(s->ldlt_mat); free(s->ldlt_transp); free(s->ldlt_tmp);
    free(s->tmp_n); free(s->tmp_n2); free(s->tmp_m);
}

void FN(oracle_qp_free)(FN(oracle_solver) * s) {
    if (!s) return;
    FN(orc_free_bufs)(s);
    free(s);
}

FN(oracle_settings) * FN(oracle_qp_settings)(FN(oracle_solver) * s) { return &s->settings; }
FN(oracle_info) * FN(oracle_qp_info)(FN(oracle_solver) * s) { return &s->info; }
SCALAR *FN(oracle_qp_primal)(FN(oracle_solver) * s) { return s->x; }
SCALAR *FN(oracle_qp_dual)(FN(oracle_solver) * s) { return s->y; }
SCALAR *FN(oracle_qp_z)(FN(oracle_solver) * s) { return s->z; }
SCALAR FN(oracle_qp_rho)(FN(oracle_solver) * s) { return s->rho; }

/* ---- Eigen::LDLT<Matrix, Lower> restatement ------------------------------------------- */

#define M_(i, j) mat[(size_t)(i) + (size_t)(j) * (size_t)size]

/* Eigen/src/Cholesky/LDLT.h, ldlt_inplace<Lower>::unblocked: diagonal pivoting on the
 * largest |diagonal| entry of the trailing block (first occurrence wins), symmetric swap
 * inside the lower triangle, then the unblocked column update. Returns Eigen's "ret"
 * (Success / NumericalIssue). */
static int FN(orc_ldlt_compute)(SCALAR *mat, int *transp, SCALAR *temp, int size) {
    int found_zero_pivot = 0;
    int ret = 1;
    if (size <= 1) {
        for (int i = 0; i < size; i++) transp[i] = i;
        return 1;
    }
    for (int k = 0; k < size; ++k) {
        /* mat.diagonal().tail(size-k).cwiseAbs().maxCoeff(&idx): strict '>' scan, so NaN
         * never becomes the max unless it is first. */
        int big = k;
        SCALAR best = ORC_FABS(M_(k, k));
        for (int i = k + 1; i < size; ++i) {
            SCALAR v = ORC_FABS(M_(i, i));
            if (v > best) { best = v; big = i; }
        }
        transp[k] = big;
        if (k != big) {
            int s = size - big - 1;
            for (int j = 0; j < k; ++j) { SCALAR t = M_(k, j); M_(k, j) = M_(big, j); M_(big, j) = t; }
            for (int i = 0; i < s; ++i) {
                SCALAR t = M_(big + 1 + i, k);
                M_(big + 1 + i, k) = M_(big + 1 + i, big);
                M_(big + 1 + i, big) = t;
            }
            { SCALAR t = M_(k, k); M_(k, k) = M_(big, big); M_(big, big) = t; }
            for (int i = k + 1; i < big; ++i) {
                SCALAR t = M_(i, k);
                M_(i, k) = M_(big, i);
                M_(big, i) = t;
            }
        }
        int rs = size - k - 1;
        if (k > 0) {
            /* temp.head(k) = D[0:k] .* A10^T ; A11 -= A10*temp ; A21 -= A20*temp */
            for (int j = 0; j < k; ++j) temp[j] = M_(j, j) * M_(k, j);
            SCALAR acc = 0;
            for (int j = 0; j < k; ++j) acc += M_(k, j) * temp[j];
            M_(k, k) -= acc;
            if (rs > 0) {
                /* column-major GEMV: accumulate column by column like Eigen's general
                 * matrix-vector kernel */
                for (int j = 0; j < k; ++j) {
                    SCALAR tj = temp[j];
                    const SCALAR *col = &M_(k + 1, j);
                    SCALAR *dst = &M_(k + 1, k);
                    for (int i = 0; i < rs; ++i) dst[i] -= col[i] * tj;
                }
            }
        }
        SCALAR akk = M_(k, k);
        int pivot_is_valid = (ORC_FABS(akk) > (SCALAR)0);
        if (k == 0 && !pivot_is_valid) {
            for (int j = 0; j < size; ++j) {
                transp[j] = j;
                for (int i = j + 1; i < size; ++i) ret = ret && (M_(i, j) == (SCALAR)0);
            }
            return ret;
        }
        if (rs > 0 && pivot_is_valid) {
            for (int i = 0; i < rs; ++i) M_(k + 1 + i, k) /= akk;
        } else if (rs > 0) {
            for (int i = 0; i < rs; ++i) ret = ret && (M_(k + 1 + i, k) == (SCALAR)0);
        }
        if (found_zero_pivot && pivot_is_valid) ret = 0;
        else if (!pivot_is_valid) found_zero_pivot = 1;
    }
    return ret;
}

/* LDLT::_solve_impl: dst = P^T L^-T D^+ L^-1 P rhs, with D^+ zeroing entries whose
 * |D_i| <= numeric_limits<Scalar>::min(). */
static void FN(orc_ldlt_solve)(const SCALAR *mat, const int *transp, int size, const SCALAR *rhs, SCALAR *dst) {
    for (int i = 0; i < size; ++i) dst[i] = rhs[i];
    for (int k = 0; k < size; ++k) {
        int j = transp[k];
        if (j != k) { SCALAR t = dst[k]; dst[k] = dst[j]; dst[j] = t; }
    }
    /* unit-lower forward substitution, column oriented (Eigen's col-major triangular solver) */
    for (int k = 0; k < size; ++k) {
        SCALAR v = dst[k];
        if (v != (SCALAR)0) {
            const SCALAR *col = &M_(0, k);
            for (int i = k + 1; i < size; ++i) dst[i] -= col[i] * v;
        }
    }
    for (int i = 0; i < size; ++i) {
        SCALAR d = M_(i, i);
        if (ORC_FABS(d) > ORC_MINPOS) dst[i] /= d;
        else dst[i] = 0;
    }
    /* unit-upper (L^T) backward substitution: row i of L^T is column i of L */
    for (int i = size - 1; i >= 0; --i) {
        const SCALAR *col = &M_(0, i);
        SCALAR acc = dst[i];
        for (int r = i + 1; r < size; ++r) acc -= col[r] * dst[r];
        dst[i] = acc;
    }
    for (int k = size - 1; k >= 0; --k) {
        int j = transp[k];
        if (j != k) { SCALAR t = dst[k]; dst[k] = dst[j]; dst[j] = t; }
    }
}
#undef M_

/* ---- helpers ----------------------------------------------------------------------------- */

static SCALAR FN(orc_inf_norm)(const SCALAR *v, int len) {
    SCALAR r = 0;
    for (int i = 0; i < len; ++i) { SCALAR a = ORC_FABS(v[i]); if (a > r) r = a; }
    return r;
}
/* out = A*x, A is rows x cols column-major */
static void FN(orc_gemv)(const SCALAR *A, int rows, int cols, const SCALAR *x, SCALAR *out) {
    for (int i = 0; i < rows; ++i) out[i] = 0;
    for (int j = 0; j < cols; ++j) {
        SCALAR xj = x[j];
        const SCALAR *col = A + (size_t)j * rows;
        for (int i = 0; i < rows; ++i) out[i] += col[i] * xj;
    }
}
/* out = A^T*y */
static void FN(orc_gemv_t)(const SCALAR *A, int rows, int cols, const SCALAR *y, SCALAR *out) {
    for (int j = 0; j < cols; ++j) {
        const SCALAR *col = A + (size_t)j * rows;
        SCALAR acc = 0;
        for (int i = 0; i < rows; ++i) acc += col[i] * y[i];
        out[j] = acc;
    }
}

/* src/qp.cpp:283-294 (static, public in the reference for unit testing) */
void FN(oracle_constr_type_init)(const SCALAR *l, const SCALAR *u, int m, int *constr_type) {
    for (int i = 0; i < m; i++) {
        if (l[i] < -ORC_LOOSE_BOUNDS_THRESH && u[i] > ORC_LOOSE_BOUNDS_THRESH) {
            constr_type[i] = ORC_LOOSE_BOUNDS;
        } else if (u[i] - l[i] < ORC_RHO_TOL) {
            constr_type[i] = ORC_EQUALITY_CONSTRAINT;
        } else {
            constr_type[i] = ORC_INEQUALITY_CONSTRAINT;
        }
    }
}

/* src/qp.cpp:296-314 */
static void FN(orc_rho_vec_update)(FN(oracle_solver) * s, SCALAR rho0) {
    for (int i = 0; i < s->m; i++) {
        switch (s->constr_type[i]) {
            case ORC_LOOSE_BOUNDS: s->rho_vec[i] = ORC_RHO_MIN; break;
            case ORC_EQUALITY_CONSTRAINT: s->rho_vec[i] = ORC_RHO_EQ_FACTOR * rho0; break;
            default: s->rho_vec[i] = rho0;
        }
    }
    for (int i = 0; i < s->m; i++) s->rho_inv_vec[i] = (SCALAR)1 / s->rho_vec[i];
    s->rho = rho0;
    s->info.rho_updates += 1;
}

/* src/qp.cpp:185-187: only the lower-triangle blocks are written; the upper-right block is
 * left as it is (LDLT<Lower> never reads it). The top-left block receives all of P. */
static void FN(orc_construct_KKT)(FN(oracle_solver) * s, const SCALAR *P, const SCALAR *A) {
    int n = s->n, m = s->m, N = n + m;
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < n; ++i)
            s->kkt_mat[(size_t)i + (size_t)j * N] = P[(size_t)i + (size_t)j * n] + (i == j ? s->settings.sigma : (SCALAR)0);
        for (int i = 0; i < m; ++i)
            s->kkt_mat[(size_t)(n + i) + (size_t)j * N] = A[(size_t)i + (size_t)j * m];
    }
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i)
            s->kkt_mat[(size_t)(n + i) + (size_t)(n + j) * N] = (i == j) ? (SCALAR)-1.0 * s->rho_inv_vec[i] : (SCALAR)0;
}

/* src/qp.cpp:225-235 (dense) */
static void FN(orc_update_KKT_rho)(FN(oracle_solver) * s) {
    int n = s->n, m = s->m, N = n + m;
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i)
            s->kkt_mat[(size_t)(n + i) + (size_t)(n + j) * N] = (i == j) ? (SCALAR)-1.0 * s->rho_inv_vec[i] : (SCALAR)0;
}

/* src/qp.cpp:237-259: compute_KKT / factorize_KKT are the same call in the dense build */
static int FN(orc_factorize_KKT)(FN(oracle_solver) * s) {
    int N = s->n + s->m;
    memcpy(s->ldlt_mat, s->kkt_mat, sizeof(SCALAR) * (size_t)N * N);
    s->ldlt_ok = FN(orc_ldlt_compute)(s->ldlt_mat, s->ldlt_transp, s->ldlt_tmp, N);
    return s->ldlt_ok;
}

/* src/qp.cpp:11-44 */
void FN(oracle_qp_setup)(FN(oracle_solver) * s, int n, int m, const SCALAR *P, const SCALAR *q,
                         const SCALAR *A, const SCALAR *l, const SCALAR *u) {
    (void)q;
    int N = n + m;
    if (s->n != n || s->m != m || !s->x) {
        FN(orc_free_bufs)(s);
        s->n = n; s->m = m;
        s->x = (SCALAR *)malloc(sizeof(SCALAR) * (n + 1));
        s->z = (SCALAR *)malloc(sizeof(SCALAR) * (m + 1));
        s->y = (SCALAR *)malloc(sizeof(SCALAR) * (m + 1));
        s->x_tilde = (SCALAR *)malloc(sizeof(SCALAR) * (n + 1));
        s->z_tilde = (SCALAR *)malloc(sizeof(SCALAR) * (m + 1));
        s->z_prev = (SCALAR *)malloc(sizeof(SCALAR) * (m + 1));
        s->rho_vec = (SCALAR *)malloc(sizeof(SCALAR) * (m + 1));
        s->rho_inv_vec = (SCALAR *)malloc(sizeof(SCALAR) * (m + 1));
        s->rhs = (SCALAR *)malloc(sizeof(SCALAR) * (N + 1));
        s->x_tilde_nu = (SCALAR *)malloc(sizeof(SCALAR) * (N + 1));
        s->constr_type = (int *)malloc(sizeof(int) * (m + 1));
        /* dense resize() does not zero; the upper-right block is never read. calloc keeps
         * valgrind quiet without changing any result. */
        s->kkt_mat = (SCALAR *)calloc((size_t)N * N + 1, sizeof(SCALAR));
        s->ldlt_mat = (SCALAR *)malloc(sizeof(SCALAR) * ((size_t)N * N + 1));
        s->ldlt_transp = (int *)malloc(sizeof(int) * (N + 1));
        s->ldlt_tmp = (SCALAR *)malloc(sizeof(SCALAR) * (N + 1));
        s->tmp_n = (SCALAR *)malloc(sizeof(SCALAR) * (n + 1));
        s->tmp_n2 = (SCALAR *)malloc(sizeof(SCALAR) * (n + 1));
        s->tmp_m = (SCALAR *)malloc(sizeof(SCALAR) * (m + 1));
    }
    s->diag_min_rp_norm = s->diag_min_rd_norm = 1e300;
    for (int i = 0; i < n; ++i) s->x[i] = 0; /* qp.cpp:16-18, the only real cold start */
    for (int i = 0; i < m; ++i) { s->z[i] = 0; s->y[i] = 0; }

    FN(oracle_constr_type_init)(l, u, m, s->constr_type);
    FN(orc_rho_vec_update)(s, s->settings.rho);
    FN(orc_construct_KKT)(s, P, A);
    if (FN(orc_factorize_KKT)(s)) s->info.status = ORC_UNSOLVED;
    else s->info.status = ORC_NUMERICAL_ISSUES;
}

/* src/qp.cpp:46-62 */
void FN(oracle_qp_update_qp)(FN(oracle_solver) * s, const SCALAR *P, const SCALAR *q, const SCALAR *A,
                             const SCALAR *l, const SCALAR *u) {
    (void)q;
    FN(oracle_constr_type_init)(l, u, s->m, s->constr_type);
    FN(orc_rho_vec_update)(s, s->settings.rho);
    FN(orc_construct_KKT)(s, P, A); /* update_KKT_mat == construct_KKT_mat in the dense build, qp.cpp:220 */
    if (FN(orc_factorize_KKT)(s)) s->info.status = ORC_UNSOLVED;
    else s->info.status = ORC_NUMERICAL_ISSUES;
}

/* src/qp.cpp:353-361 */
static SCALAR FN(orc_residual_prim)(FN(oracle_solver) * s, const SCALAR *A) {
    FN(orc_gemv)(A, s->m, s->n, s->x, s->tmp_m);
    for (int i = 0; i < s->m; ++i) s->tmp_m[i] -= s->z[i];
    return FN(orc_inf_norm)(s->tmp_m, s->m);
}
static SCALAR FN(orc_residual_dual)(FN(oracle_solver) * s, const SCALAR *P, const SCALAR *q, const SCALAR *A) {
    FN(orc_gemv)(P, s->n, s->n, s->x, s->tmp_n);
    FN(orc_gemv_t)(A, s->m, s->n, s->y, s->tmp_n2);
    for (int i = 0; i < s->n; ++i) s->tmp_n[i] = s->tmp_n[i] + q[i] + s->tmp_n2[i];
    return FN(orc_inf_norm)(s->tmp_n, s->n);
}

/* src/qp.cpp:316-331 */
static void FN(orc_update_state)(FN(oracle_solver) * s, const SCALAR *P, const SCALAR *q, const SCALAR *A) {
    SCALAR norm_Ax, norm_z, norm_Px, norm_ATy, norm_q;
    FN(orc_gemv)(A, s->m, s->n, s->x, s->tmp_m);
    norm_Ax = FN(orc_inf_norm)(s->tmp_m, s->m);
    norm_z = FN(orc_inf_norm)(s->z, s->m);
    s->max_Ax_z_norm = ORC_FMAX(norm_Ax, norm_z);

    FN(orc_gemv)(P, s->n, s->n, s->x, s->tmp_n);
    norm_Px = FN(orc_inf_norm)(s->tmp_n, s->n);
    FN(orc_gemv_t)(A, s->m, s->n, s->y, s->tmp_n2);
    norm_ATy = FN(orc_inf_norm)(s->tmp_n2, s->n);
    norm_q = FN(orc_inf_norm)(q, s->n);
    s->max_Px_ATy_q_norm = ORC_FMAX(norm_Px, ORC_FMAX(norm_ATy, norm_q));

    s->info.res_prim = FN(orc_residual_prim)(s, A);
    s->info.res_dual = FN(orc_residual_dual)(s, P, q, A);
}

/* src/qp.cpp:333-341 */
static SCALAR FN(orc_rho_estimate)(FN(oracle_solver) * s, SCALAR rho0) {
    SCALAR rp_norm = s->info.res_prim / (s->max_Ax_z_norm + ORC_EPS);
    SCALAR rd_norm = s->info.res_dual / (s->max_Px_ATy_q_norm + ORC_EPS);
    if ((double)rp_norm < s->diag_min_rp_norm) s->diag_min_rp_norm = (double)rp_norm;
    if ((double)rd_norm < s->diag_min_rd_norm) s->diag_min_rd_norm = (double)rd_norm;
    return ORC_RHO_EST(rho0, rp_norm / (rd_norm + ORC_EPS));
}

/* src/qp.cpp:363-371 with eps_prim / eps_dual :343-351 */
static int FN(orc_termination)(FN(oracle_solver) * s) {
    SCALAR ep = s->settings.eps_abs + s->settings.eps_rel * s->max_Ax_z_norm;
    SCALAR ed = s->settings.eps_abs + s->settings.eps_rel * s->max_Px_ATy_q_norm;
    return (s->info.res_prim <= ep && s->info.res_dual <= ed);
}

/* src/qp.cpp:64-157 */
void FN(oracle_qp_solve)(FN(oracle_solver) * s, const SCALAR *P, const SCALAR *q, const SCALAR *A,
                         const SCALAR *l, const SCALAR *u) {
    int check_termination = 0;
    int n = s->n, m = s->m, N = n + m;
    if (s->info.status == ORC_UNINITIALIZED || s->info.status == ORC_NUMERICAL_ISSUES) return;

    /* qp.cpp:78-82: "x.Zero(n)" builds and discards a temporary -- x, z, y are NOT reset
     * (SURVEY.md section 0 fact 4). Restated as the no-op it is. */
    if (!s->settings.warm_start) { /* no-op */ }

    int iter;
    for (iter = 1; iter <= s->settings.max_iter; iter++) {
        const SCALAR alpha = s->settings.alpha;
        memcpy(s->z_prev, s->z, sizeof(SCALAR) * m);

        /* form_KKT_rhs, qp.cpp:272-276 */
        for (int i = 0; i < n; ++i) s->rhs[i] = s->settings.sigma * s->x[i] - q[i];
        for (int i = 0; i < m; ++i) s->rhs[n + i] = s->z[i] - s->rho_inv_vec[i] * s->y[i];
        FN(orc_ldlt_solve)(s->ldlt_mat, s->ldlt_transp, N, s->rhs, s->x_tilde_nu);

        for (int i = 0; i < n; ++i) s->x_tilde[i] = s->x_tilde_nu[i];
        for (int i = 0; i < m; ++i)
            s->z_tilde[i] = s->z_prev[i] + s->rho_inv_vec[i] * (s->x_tilde_nu[n + i] - s->y[i]);

        for (int i = 0; i < n; ++i) s->x[i] = alpha * s->x_tilde[i] + ((SCALAR)1 - alpha) * s->x[i];

        for (int i = 0; i < m; ++i) {
            SCALAR zi = alpha * s->z_tilde[i] + ((SCALAR)1 - alpha) * s->z_prev[i] + s->rho_inv_vec[i] * s->y[i];
            /* box_projection qp.cpp:278-281: cwiseMax(l) then cwiseMin(u) */
            zi = (zi < l[i]) ? l[i] : zi;
            zi = (u[i] < zi) ? u[i] : zi;
            s->z[i] = zi;
        }
        for (int i = 0; i < m; ++i)
            s->y[i] = s->y[i] + s->rho_vec[i] * (alpha * s->z_tilde[i] + ((SCALAR)1 - alpha) * s->z_prev[i] - s->z[i]);

        if (s->settings.check_termination != 0 && iter % s->settings.check_termination == 0) check_termination = 1;
        else check_termination = 0;

        if (check_termination) {
            FN(orc_update_state)(s, P, q, A);
            if (FN(orc_termination)(s)) { s->info.status = ORC_SOLVED; break; }
        }

        if (s->settings.adaptive_rho && iter % s->settings.adaptive_rho_interval == 0) {
            if (!check_termination) FN(orc_update_state)(s, P, q, A);
            SCALAR new_rho = FN(orc_rho_estimate)(s, s->rho);
            new_rho = ORC_FMAX(ORC_RHO_MIN, ORC_FMIN(new_rho, ORC_RHO_MAX));
            s->info.rho_estimate = new_rho;
            if (new_rho < s->rho / s->settings.adaptive_rho_tolerance ||
                new_rho > s->rho * s->settings.adaptive_rho_tolerance) {
                FN(orc_rho_vec_update)(s, new_rho);
                FN(orc_update_KKT_rho)(s);
                if (!FN(orc_factorize_KKT)(s)) { s->info.status = ORC_NUMERICAL_ISSUES; break; }
            }
        }
    }
    if (iter > s->settings.max_iter) s->info.status = ORC_MAX_ITER_EXCEEDED;
    s->info.iter = iter; /* == max_iter + 1 on MAX_ITER_EXCEEDED, SURVEY.md section 0 fact 5 */
    s->iter = iter;
}

/* Expose the factor for test cross-checks against numpy (not part of the reference API). */
void FN(oracle_qp_kkt_solve)(FN(oracle_solver) * s, const SCALAR *rhs, SCALAR *out) {
    FN(orc_ldlt_solve)(s->ldlt_mat, s->ldlt_transp, s->n + s->m, rhs, out);
}
void FN(oracle_qp_ldlt_dump)(FN(oracle_solver) * s, SCALAR *D, int *transp) {
    int N = s->n + s->m;
    for (int i = 0; i < N; ++i) { D[i] = s->ldlt_mat[(size_t)i + (size_t)i * N]; transp[i] = s->ldlt_transp[i]; }
}

/* Batched convenience used by the parity tests and by bench.py's cpu_baseline leg:
 * the reference's calling pattern at src/sqp.cpp:221-222 (fresh setup + solve per QP),
 * one solver object per OpenMP thread, dynamic schedule over the batch. */
int FN(oracle_qp_solve_batch)(const FN(oracle_settings) * settings, int batch, int n, int m,
                              const SCALAR *P, const SCALAR *q, const SCALAR *A, const SCALAR *l, const SCALAR *u,
                              SCALAR *x, SCALAR *y, SCALAR *z, int *status, int *iter, SCALAR *res_prim,
                              SCALAR *res_dual, int *rho_updates, SCALAR *rho_estimate, int nthreads, double *diag_min_norms) {
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#pragma omp parallel num_threads(nthreads)
#endif
    {
        FN(oracle_solver) *s = FN(oracle_qp_new)();
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int b = 0; b < batch; ++b) {
            const SCALAR *Pb = P + (size_t)b * n * n, *qb = q + (size_t)b * n, *Ab = A + (size_t)b * m * n;
            const SCALAR *lb = l + (size_t)b * m, *ub = u + (size_t)b * m;
            s->settings = *settings;
            s->info.status = ORC_UNINITIALIZED;
            s->info.iter = 0; s->info.rho_updates = 0; s->info.rho_estimate = 0;
            s->info.res_prim = 0; s->info.res_dual = 0;
            FN(oracle_qp_setup)(s, n, m, Pb, qb, Ab, lb, ub);
            FN(oracle_qp_solve)(s, Pb, qb, Ab, lb, ub);
            if (x) memcpy(x + (size_t)b * n, s->x, sizeof(SCALAR) * n);
            if (y) memcpy(y + (size_t)b * m, s->y, sizeof(SCALAR) * m);
            if (z) memcpy(z + (size_t)b * m, s->z, sizeof(SCALAR) * m);
            if (status) status[b] = s->info.status;
            if (iter) iter[b] = s->info.iter;
            if (res_prim) res_prim[b] = s->info.res_prim;
            if (res_dual) res_dual[b] = s->info.res_dual;
            if (rho_updates) rho_updates[b] = s->info.rho_updates;
            if (rho_estimate) rho_estimate[b] = s->info.rho_estimate;
            if (diag_min_norms) { diag_min_norms[2 * b] = s->diag_min_rp_norm; diag_min_norms[2 * b + 1] = s->diag_min_rd_norm; }
        }
        FN(oracle_qp_free)(s);
    }
    return used;
}

#undef FN
#undef CAT
#undef CAT_
#undef ORC_RHO_MIN
#undef ORC_RHO_MAX
#undef ORC_RHO_TOL
#undef ORC_RHO_EQ_FACTOR
#undef ORC_LOOSE_BOUNDS_THRESH
