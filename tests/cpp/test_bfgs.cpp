// Host-only: mirrors the reference's tests/bfgs_test.cpp:21-65 against sqp_solver_b200/host/batch/solvers/bfgs.hpp.
#include <cmath>

#include "mini_test.hpp"
#include "solvers/sqp.hpp"

using Mat = sqpb200_dense::Matrix<double>;
using Vec = sqpb200_dense::Vector<double>;

static bool approx(const Mat &A, const Mat &B, double prec) {
    double d = 0, a = 0, b = 0;
    for (int j = 0; j < A.cols(); ++j)
        for (int i = 0; i < A.rows(); ++i) {
            d += (A(i, j) - B(i, j)) * (A(i, j) - B(i, j));
            a += A(i, j) * A(i, j);
            b += B(i, j) * B(i, j);
        }
    return d <= prec * prec * std::min(a, b);
}

static void run(double h11, bool expect_converge) {
    Mat H(2, 2), B(2, 2);
    H.setZero();
    H(0, 0) = 2;
    H(1, 1) = h11;
    B.setIdentity();
    Vec step(2), delta_grad(2);
    for (int i = 0; i < 10; i++) {
        step(0) = std::sin(i);
        step(1) = std::cos(i);
        delta_grad(0) = H(0, 0) * step(0) + H(0, 1) * step(1);
        delta_grad(1) = H(1, 0) * step(0) + H(1, 1) * step(1);
        BFGS_update(B, step, delta_grad);
        EXPECT_TRUE(sqp::detail::is_posdef(B));
    }
    if (expect_converge) EXPECT_TRUE(approx(B, H, 1e-3));
}

TEST(BFGSTestCase, Test2D_posdef) { run(1.0, true); }      // bfgs_test.cpp:21-43
TEST(BFGSTestCase, Test2D_indefinite) { run(-1.0, false); }  // bfgs_test.cpp:45-65

TEST(DenseShim, ColumnMajorLikeEigen) {
    Mat M(3, 2);
    for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 3; ++i) M(i, j) = 10 * i + j;
    EXPECT_EQ(M.data()[1], 10.0);  // (1,0) follows (0,0): column-major
    EXPECT_EQ(M.data()[3], 1.0);   // (0,1) starts the second column
    Vec a = {3.0, 4.0};
    EXPECT_EQ(a.norm(), 5.0);
    Vec b = {3.0, 4.01};
    EXPECT_TRUE(a.isApprox(b, 1e-2));
    EXPECT_TRUE(!a.isApprox(b, 1e-4));
}

TEST(SQPSettings, ValidateAcceptsDefaults) {
    sqp::sqp_settings_t<double> s;
    EXPECT_TRUE(s.validate());
    s.tau = 1.5;
    EXPECT_TRUE(!s.validate());
}

MINI_TEST_MAIN()
