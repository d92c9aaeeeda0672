// GPU: mirrors the reference's tests/qp_solver_test.cpp (all six TESTs, same bodies) against the
// drop-in qp_solver::QPSolver<Scalar> of sqp_solver_b200/host/overlay/solvers/qp.hpp, plus the batched sibling.
#include "mini_test.hpp"
#include "solvers/qp.hpp"

using namespace qp_solver;

// The fixture of tests/qp_solver_test.cpp:6-41 (P = [[4,1],[1,2]], q = [1,1], A = [[1,1],[1,0],[0,1]],
// l = [1,0,0], u = [1,0.7,0.7], solution [0.3,0.7]); the problem owns its storage and exposes the
// reference's non-owning QuadraticProblem view.
template <typename Scalar>
struct SimpleQP : QuadraticProblem<Scalar> {
    using Matrix = sqpb200_dense::Matrix<Scalar>;
    using Vector = sqpb200_dense::Vector<Scalar>;
    Matrix Pm{2, 2}, Am{3, 2};
    Vector qv{Scalar(1), Scalar(1)}, lv{Scalar(1), Scalar(0), Scalar(0)}, uv{Scalar(1), Scalar(0.7), Scalar(0.7)};
    Vector SOLUTION{Scalar(0.3), Scalar(0.7)};
    SimpleQP() {
        const Scalar Pvals[2][2] = {{4, 1}, {1, 2}}, Avals[3][2] = {{1, 1}, {1, 0}, {0, 1}};
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) Pm(i, j) = Pvals[i][j];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 2; ++j) Am(i, j) = Avals[i][j];
        this->P = &Pm;
        this->q = &qv;
        this->A = &Am;
        this->l = &lv;
        this->u = &uv;
    }
    SimpleQP(const SimpleQP &) = delete;
};

// setup + solve of the fixture with caller-chosen settings; returns the solver for inspection
template <typename Scalar, typename Configure>
static QPSolver<Scalar> solved(const SimpleQP<Scalar> &qp, Configure configure) {
    QPSolver<Scalar> solver;
    configure(solver.settings());
    solver.setup(qp);
    solver.solve(qp);
    return solver;
}

TEST(QPSolverTest, testSimpleQP) {  // tests/qp_solver_test.cpp:43-56
    SimpleQP<double> qp;
    auto solver = solved<double>(qp, [](QPSolverSettings<double> &s) { s.max_iter = 1000; });
    EXPECT_TRUE(solver.primal_solution().isApprox(qp.SOLUTION, 1e-2));
    EXPECT_LT(solver.info().iter, solver.settings().max_iter);
    EXPECT_EQ(solver.info().status, SOLVED);
    EXPECT_EQ(solver.info().iter, 125);  // oracle regression value
}

TEST(QPSolverTest, testSinglePrecisionFloat) {  // tests/qp_solver_test.cpp:58-69
    SimpleQP<float> qp;
    auto solver = solved<float>(qp, [](QPSolverSettings<float> &) {});
    EXPECT_TRUE(solver.primal_solution().isApprox(qp.SOLUTION, 1e-2f));
    EXPECT_LT(solver.info().iter, solver.settings().max_iter);
    EXPECT_EQ(solver.info().status, SOLVED);
}

TEST(QPSolverTest, testConstraintViolation) {  // tests/qp_solver_test.cpp:71-87
    SimpleQP<double> qp;
    auto solver = solved<double>(qp, [](QPSolverSettings<double> &s) { s.eps_rel = s.eps_abs = 1e-4f; });
    const auto &sol = solver.primal_solution();
    // feasibility with the reference's epsilon margin: l - 1e-3 <= A x <= u + 1e-3
    for (int i = 0; i < 3; ++i) {
        const double Ax = qp.Am(i, 0) * sol(0) + qp.Am(i, 1) * sol(1);
        EXPECT_GE(Ax - qp.lv(i), -1e-3);
        EXPECT_LE(Ax - qp.uv(i), 1e-3);
    }
}

TEST(QPSolverTest, testAdaptiveRho) {  // tests/qp_solver_test.cpp:89-100
    SimpleQP<double> qp;
    auto solver = solved<double>(qp, [](QPSolverSettings<double> &s) {
        s.adaptive_rho = true;
        s.adaptive_rho_interval = 10;
    });
    EXPECT_EQ(solver.info().status, SOLVED);
    EXPECT_EQ(solver.info().rho_updates, 2);  // oracle regression value
}

TEST(QPSolverTest, testAdaptiveRhoImprovesConvergence) {  // tests/qp_solver_test.cpp:102-125
    SimpleQP<double> qp;
    auto solver = solved<double>(qp, [](QPSolverSettings<double> &s) {
        s.warm_start = false;
        s.max_iter = 1000;
        s.rho = 0.1;
        s.adaptive_rho = false;
    });
    const int iters_without = solver.info().iter;
    // second solve on the same solver (a warm start in the reference) with adaptive rho
    solver.settings().adaptive_rho = true;
    solver.settings().adaptive_rho_interval = 10;
    solver.solve(qp);
    EXPECT_LT(solver.info().iter, solver.settings().max_iter);
    EXPECT_LT(solver.info().iter, iters_without);
    EXPECT_EQ(solver.info().status, SOLVED);
}

TEST(QPSolverTest, TestConstraint) {
    using Solver = QPSolver<double>;
    sqpb200_dense::Vector<double> l(5), u(5);
    int type_expect[5];
    l(0) = -10 * Solver::LOOSE_BOUNDS_THRESH; u(0) = 10 * Solver::LOOSE_BOUNDS_THRESH; type_expect[0] = Solver::LOOSE_BOUNDS;
    l(1) = -1; u(1) = 10 * Solver::LOOSE_BOUNDS_THRESH; type_expect[1] = Solver::INEQUALITY_CONSTRAINT;
    l(2) = -10 * Solver::LOOSE_BOUNDS_THRESH; u(2) = 2; type_expect[2] = Solver::INEQUALITY_CONSTRAINT;
    l(3) = -3; u(3) = 4; type_expect[3] = Solver::INEQUALITY_CONSTRAINT;
    l(4) = 42; u(4) = 42; type_expect[4] = Solver::EQUALITY_CONSTRAINT;
    sqpb200_dense::VectorXi constr_type(5);
    Solver::constr_type_init(l, u, constr_type);
    for (int i = 0; i < l.rows(); i++) EXPECT_EQ(constr_type[i], type_expect[i]);
}

TEST(QPSolverTest, solveBeforeSetupIsNoop) {  // src/qp.cpp:68-71
    SimpleQP<double> qp;
    QPSolver<double> solver;
    solver.solve(qp);
    EXPECT_EQ(solver.info().status, UNINITIALIZED);
    EXPECT_EQ(solver.info().iter, 0);
}

TEST(BatchQPSolverTest, manyCopiesOfSimpleQP) {
    const int B = 64;
    std::vector<double> P, q, A, l, u;
    for (int b = 0; b < B; ++b) {
        const double Pm[] = {4, 1, 1, 2}, qm[] = {1, 1}, Am[] = {1, 1, 0, 1, 0, 1}, lm[] = {1, 0, 0}, um[] = {1, 0.7, 0.7};
        P.insert(P.end(), Pm, Pm + 4); q.insert(q.end(), qm, qm + 2); A.insert(A.end(), Am, Am + 6);
        l.insert(l.end(), lm, lm + 3); u.insert(u.end(), um, um + 3);
    }
    BatchQPSolver solver(B, 2, 3);
    solver.setup_solve(P.data(), q.data(), A.data(), l.data(), u.data());
    for (int b = 0; b < B; ++b) {
        EXPECT_EQ(solver.info(b).status, SOLVED);
        EXPECT_EQ(solver.info(b).iter, 125);
        EXPECT_TRUE(std::abs(solver.primal_solution(b)[0] - 0.3) < 3e-3 && std::abs(solver.primal_solution(b)[1] - 0.7) < 3e-3);
    }
    EXPECT_EQ(solver.total_iterations(), 125LL * B);
}

// tests/qp_solver_sparse_test.cpp:36-49 (dead code in the reference: QP_SOLVER_USE_SPARSE is never defined): SimpleQP with
// A = [[1,1],[1,0],[0,1]] stored as an Eigen::SparseMatrix (compressed columns), adaptive rho -> [0.3, 0.7]
TEST(BatchQPSolverTest, sparseSimpleQP) {
    const int B = 8;
    const int outer[] = {0, 2, 4}, inner[] = {0, 1, 0, 2};
    std::vector<double> P, q, vals, l, u;
    for (int b = 0; b < B; ++b) {
        const double Pm[] = {4, 1, 1, 2}, qm[] = {1, 1}, vm[] = {1, 1, 1, 1}, lm[] = {1, 0, 0}, um[] = {1, 0.7, 0.7};
        P.insert(P.end(), Pm, Pm + 4); q.insert(q.end(), qm, qm + 2); vals.insert(vals.end(), vm, vm + 4);
        l.insert(l.end(), lm, lm + 3); u.insert(u.end(), um, um + 3);
    }
    BatchQPSolver solver(B, 2, 3);
    solver.settings().max_iter = 1000;
    solver.settings().adaptive_rho = true;
    solver.setup_solve_sparse(P.data(), q.data(), vals.data(), outer, inner, 4, SQPB200_SPARSE_CSC, l.data(), u.data());
    for (int b = 0; b < B; ++b) {
        EXPECT_EQ(solver.info(b).status, SOLVED);
        EXPECT_TRUE(std::abs(solver.primal_solution(b)[0] - 0.3) < 3e-3 && std::abs(solver.primal_solution(b)[1] - 0.7) < 3e-3);
    }
}

// A banded sparse problem large enough for the thread-block-cluster kernel (n = 96 > 64): min 1/2 |x|^2 - sum x subject to
// 0 <= x_i + x_{i+1} <= 1 (bidiagonal A, CSR) and x_0 = 0.25 as an equality row. Checked against the same problem through the
// dense entry point (blocked kernel): same status and iteration count, solutions to 1e-6.
TEST(BatchQPSolverTest, sparseBandedMatchesDense) {
    const int B = 5, n = 96, m = 96;
    std::vector<int> outer(m + 1), inner;
    std::vector<double> v1, Ad((size_t)m * n, 0.0);
    for (int i = 0; i < m; ++i) {
        outer[i] = (int)inner.size();
        if (i == m - 1) {
            inner.push_back(0); v1.push_back(1.0); Ad[i + (size_t)m * 0] = 1.0;
        } else {
            inner.push_back(i); v1.push_back(1.0); Ad[i + (size_t)m * i] = 1.0;
            inner.push_back(i + 1); v1.push_back(1.0 + 0.01 * i); Ad[i + (size_t)m * (i + 1)] = 1.0 + 0.01 * i;
        }
    }
    outer[m] = (int)inner.size();
    const int nnz = (int)inner.size();
    std::vector<double> P, q, vals, A, l, u;
    for (int b = 0; b < B; ++b) {
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) P.push_back(i == j ? 1.0 + 0.1 * b : 0.0);
        for (int j = 0; j < n; ++j) q.push_back(-1.0);
        vals.insert(vals.end(), v1.begin(), v1.end());
        A.insert(A.end(), Ad.begin(), Ad.end());
        for (int i = 0; i < m; ++i) { l.push_back(i == m - 1 ? 0.25 : 0.0); u.push_back(i == m - 1 ? 0.25 : 1.0); }
    }
    BatchQPSolver sparse(B, n, m), dense(B, n, m);
    for (BatchQPSolver *s : {&sparse, &dense}) { s->settings().alpha = 1.6; s->settings().adaptive_rho = true; }
    sparse.setup_solve_sparse(P.data(), q.data(), vals.data(), outer.data(), inner.data(), nnz, SQPB200_SPARSE_CSR, l.data(), u.data());
    dense.setup_solve(P.data(), q.data(), A.data(), l.data(), u.data());
    for (int b = 0; b < B; ++b) {
        EXPECT_EQ(sparse.info(b).status, SOLVED);
        EXPECT_EQ(sparse.info(b).status, dense.info(b).status);
        EXPECT_EQ(sparse.info(b).iter, dense.info(b).iter);
        double worst = 0;
        for (int j = 0; j < n; ++j) worst = std::max(worst, std::abs(sparse.primal_solution(b)[j] - dense.primal_solution(b)[j]));
        EXPECT_TRUE(worst < 1e-6);
        EXPECT_TRUE(std::abs(sparse.primal_solution(b)[0] - 0.25) < 1e-2);
    }
}

MINI_TEST_MAIN()
