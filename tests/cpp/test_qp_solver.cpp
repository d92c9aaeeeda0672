// GPU: mirrors the reference's tests/qp_solver_test.cpp (all six TESTs, same bodies) against the
// drop-in qp_solver::QPSolver<Scalar> of sqp_solver_b200/host/solvers/qp.hpp, plus the batched sibling.
#include "mini_test.hpp"
#include "solvers/qp.hpp"

using namespace qp_solver;

template <typename Scalar>
class SimpleQP : public QuadraticProblem<Scalar> {  // tests/qp_solver_test.cpp:6-41
   public:
    using BASE = QuadraticProblem<Scalar>;
    using Matrix = sqpb200_dense::Matrix<Scalar>;
    using Vector = sqpb200_dense::Vector<Scalar>;
    SimpleQP() : SOLUTION(2) {
        Matrix *P = new Matrix(2, 2);
        Vector *q = new Vector(2);
        Matrix *A = new Matrix(3, 2);
        Vector *l = new Vector(3);
        Vector *u = new Vector(3);
        (*P)(0, 0) = 4; (*P)(0, 1) = 1; (*P)(1, 0) = 1; (*P)(1, 1) = 2;
        (*q)(0) = 1; (*q)(1) = 1;
        (*A)(0, 0) = 1; (*A)(0, 1) = 1; (*A)(1, 0) = 1; (*A)(1, 1) = 0; (*A)(2, 0) = 0; (*A)(2, 1) = 1;
        (*l)(0) = 1; (*l)(1) = 0; (*l)(2) = 0;
        (*u)(0) = 1; (*u)(1) = 0.7; (*u)(2) = 0.7;
        BASE::P = P; BASE::q = q; BASE::A = A; BASE::l = l; BASE::u = u;
        SOLUTION(0) = 0.3; SOLUTION(1) = 0.7;
    }
    ~SimpleQP() { delete BASE::P; delete BASE::q; delete BASE::A; delete BASE::l; delete BASE::u; }
    Vector SOLUTION;
};

TEST(QPSolverTest, testSimpleQP) {
    SimpleQP<double> qp;
    QPSolver<double> solver;
    solver.settings().max_iter = 1000;
    solver.setup(qp);
    solver.solve(qp);
    auto sol = solver.primal_solution();
    EXPECT_TRUE(sol.isApprox(qp.SOLUTION, 1e-2));
    EXPECT_LT(solver.info().iter, solver.settings().max_iter);
    EXPECT_EQ(solver.info().status, SOLVED);
    EXPECT_EQ(solver.info().iter, 125);  // oracle regression value
}

TEST(QPSolverTest, testSinglePrecisionFloat) {
    SimpleQP<float> qp;
    QPSolver<float> solver;
    solver.setup(qp);
    solver.solve(qp);
    auto sol = solver.primal_solution();
    EXPECT_TRUE(sol.isApprox(qp.SOLUTION, 1e-2f));
    EXPECT_LT(solver.info().iter, solver.settings().max_iter);
    EXPECT_EQ(solver.info().status, SOLVED);
}

TEST(QPSolverTest, testConstraintViolation) {
    SimpleQP<double> qp;
    QPSolver<double> solver;
    solver.settings().eps_rel = 1e-4f;
    solver.settings().eps_abs = 1e-4f;
    solver.setup(qp);
    solver.solve(qp);
    auto sol = solver.primal_solution();
    double lower = 1e300, upper = -1e300;
    for (int i = 0; i < 3; ++i) {
        double Ax = (*qp.A)(i, 0) * sol(0) + (*qp.A)(i, 1) * sol(1);
        lower = std::min(lower, Ax - (*qp.l)(i));
        upper = std::max(upper, Ax - (*qp.u)(i));
    }
    EXPECT_GE(lower, -1e-3);
    EXPECT_LE(upper, 1e-3);
}

TEST(QPSolverTest, testAdaptiveRho) {
    SimpleQP<double> qp;
    QPSolver<double> solver;
    solver.settings().adaptive_rho = true;
    solver.settings().adaptive_rho_interval = 10;
    solver.setup(qp);
    solver.solve(qp);
    EXPECT_EQ(solver.info().status, SOLVED);
    EXPECT_EQ(solver.info().rho_updates, 2);
}

TEST(QPSolverTest, testAdaptiveRhoImprovesConvergence) {
    SimpleQP<double> qp;
    QPSolver<double> solver;
    solver.settings().warm_start = false;
    solver.settings().max_iter = 1000;
    solver.settings().rho = 0.1;
    solver.settings().adaptive_rho = false;
    solver.setup(qp);
    solver.solve(qp);
    int prev_iter = solver.info().iter;
    solver.settings().adaptive_rho = true;
    solver.settings().adaptive_rho_interval = 10;
    solver.solve(qp);
    auto info = solver.info();
    EXPECT_LT(info.iter, solver.settings().max_iter);
    EXPECT_LT(info.iter, prev_iter);
    EXPECT_EQ(info.status, SOLVED);
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

MINI_TEST_MAIN()
