// NLP test problems of the reference's tests/sqp_test.cpp and tests/sqp_test_autodiff.cpp with hand-written
// derivatives (Eigen's AutoDiff module is not available), shared by test_sqp.cpp and sqp_cli.cpp.
#pragma once
#include <limits>

#include "solvers/sqp.hpp"

using namespace sqp;
using Vec = NonLinearProblem<double>::Vector;
using Mat = NonLinearProblem<double>::Matrix;
static const double inf = std::numeric_limits<double>::infinity();

struct SimpleNLP : public NonLinearProblem<double> {  // tests/sqp_test.cpp:8-44
    SimpleNLP() { num_var = 2; num_constr = 3; }
    void objective(const Vec &x, double &obj) override { obj = -(x(0) + x(1)); }
    void objective_linearized(const Vec &x, Vec &grad, double &obj) override {
        grad.resize(num_var);
        objective(x, obj);
        grad(0) = -1; grad(1) = -1;
    }
    void constraint(const Vec &x, Vec &c, Vec &l, Vec &u) override {
        c(0) = x(0) * x(0) + x(1) * x(1); c(1) = x(0); c(2) = x(1);
        l(0) = 1; l(1) = 0; l(2) = 0;
        u(0) = 2; u(1) = inf; u(2) = inf;
    }
    void constraint_linearized(const Vec &x, Mat &Jc, Vec &c, Vec &l, Vec &u) override {
        Jc.resize(3, 2);
        constraint(x, c, l, u);
        Jc(0, 0) = 2 * x(0); Jc(0, 1) = 2 * x(1);
        Jc(1, 0) = 1; Jc(1, 1) = 0;
        Jc(2, 0) = 0; Jc(2, 1) = 1;
    }
};

struct SimpleQPasNLP : public NonLinearProblem<double> {  // tests/sqp_test.cpp:92-124
    SimpleQPasNLP() { num_var = 2; num_constr = 3; }
    void objective(const Vec &x, double &obj) override {
        obj = 0.5 * (x(0) * (4 * x(0) + x(1)) + x(1) * (x(0) + 2 * x(1))) + x(0) + x(1);
    }
    void objective_linearized(const Vec &x, Vec &grad, double &obj) override {
        objective(x, obj);
        grad(0) = 4 * x(0) + x(1) + 1;
        grad(1) = x(0) + 2 * x(1) + 1;
    }
    void constraint(const Vec &x, Vec &c, Vec &l, Vec &u) override {
        c(0) = x(0) + x(1); c(1) = x(0); c(2) = x(1);
        l(0) = 1; l(1) = 0; l(2) = 0;
        u(0) = 1; u(1) = 0.7; u(2) = 0.7;
    }
    void constraint_linearized(const Vec &x, Mat &Jc, Vec &c, Vec &l, Vec &u) override {
        constraint(x, c, l, u);
        Jc(0, 0) = 1; Jc(0, 1) = 1; Jc(1, 0) = 1; Jc(1, 1) = 0; Jc(2, 0) = 0; Jc(2, 1) = 1;
    }
};

struct ConstrainedRosenbrock2D : public NonLinearProblem<double> {  // tests/sqp_test_autodiff.cpp:73-99
    ConstrainedRosenbrock2D() { num_var = 2; num_constr = 2; }
    void objective(const Vec &x, double &obj) override {
        const double a = 1 - x(0), b = x(1) - x(0) * x(0);
        obj = a * a + 100 * b * b;
    }
    void objective_linearized(const Vec &x, Vec &grad, double &obj) override {
        objective(x, obj);
        const double b = x(1) - x(0) * x(0);
        grad(0) = -2 * (1 - x(0)) - 400 * x(0) * b;
        grad(1) = 200 * b;
    }
    void constraint(const Vec &x, Vec &c, Vec &l, Vec &u) override {
        c(0) = x(0) - x(1); c(1) = x(0) * x(0) + x(1) * x(1);
        u(0) = 0; u(1) = 1;
        l(0) = -inf; l(1) = 1;
    }
    void constraint_linearized(const Vec &x, Mat &Jc, Vec &c, Vec &l, Vec &u) override {
        constraint(x, c, l, u);
        Jc(0, 0) = 1; Jc(0, 1) = -1; Jc(1, 0) = 2 * x(0); Jc(1, 1) = 2 * x(1);
    }
};

struct SimpleNLP2 : public NonLinearProblem<double> {  // N&W example 12.1, tests/sqp_test_autodiff.cpp:245-265
    SimpleNLP2() { num_var = 2; num_constr = 1; }
    void objective(const Vec &x, double &obj) override { obj = x(0) + x(1); }
    void objective_linearized(const Vec &x, Vec &grad, double &obj) override { objective(x, obj); grad(0) = 1; grad(1) = 1; }
    void constraint(const Vec &x, Vec &c, Vec &l, Vec &u) override { c(0) = x(0) * x(0) + x(1) * x(1); l(0) = 2; u(0) = 2; }
    void constraint_linearized(const Vec &x, Mat &Jc, Vec &c, Vec &l, Vec &u) override {
        constraint(x, c, l, u);
        Jc(0, 0) = 2 * x(0); Jc(0, 1) = 2 * x(1);
    }
};

struct RosenbrockBox : public NonLinearProblem<double> {  // `Rosenbrock`, tests/sqp_test_autodiff.cpp:61-71 and :122-144: 0 <= x <= 1
    explicit RosenbrockBox(int n) { num_var = n; num_constr = n; }
    void objective(const Vec &x, double &obj) override {
        obj = 0;
        for (int i = 0; i < num_var - 1; ++i) {
            const double a = 1 - x(i), b = x(i + 1) - x(i) * x(i);
            obj += a * a + 100 * b * b;
        }
    }
    void objective_linearized(const Vec &x, Vec &grad, double &obj) override {
        objective(x, obj);
        for (int i = 0; i < num_var; ++i) grad(i) = 0;
        for (int i = 0; i < num_var - 1; ++i) {
            const double b = x(i + 1) - x(i) * x(i);
            grad(i) += -2 * (1 - x(i)) - 400 * x(i) * b;
            grad(i + 1) += 200 * b;
        }
    }
    void constraint(const Vec &x, Vec &c, Vec &l, Vec &u) override {
        for (int i = 0; i < num_var; ++i) { c(i) = x(i); l(i) = 0; u(i) = 1; }
    }
    void constraint_linearized(const Vec &x, Mat &Jc, Vec &c, Vec &l, Vec &u) override {
        constraint(x, c, l, u);
        for (int j = 0; j < num_var; ++j)
            for (int i = 0; i < num_var; ++i) Jc(i, j) = i == j ? 1.0 : 0.0;
    }
};

static Vec v2(double a, double b) { Vec v(2); v(0) = a; v(1) = b; return v; }
static Vec zeros(int n) { Vec v(n); v.setZero(); return v; }

