// BASELINE.json config 4 -- "full SQP: batch=4096 constrained-Rosenbrock (tests/sqp_test_autodiff.cpp:73-99) with BFGS Hessian, host
// outer loop + GPU QP": B instances of the reference's ConstrainedRosenbrock2D test problem from B starting points, advanced in
// lock-step by sqp::BatchSQP (host C++: linearisation, damped BFGS, PD repair, l1-merit line search, termination -- the reference's
// src/sqp.cpp logic) with ONE batched GPU QP solve per outer iteration. Prints one JSON line; bench.py records it.
//
//   batch_sqp_bench [batch=4096] [runs=3] [device=0]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>

#include "../batch/solvers/sqp.hpp"

using namespace sqp;
using Vec = NonLinearProblem<double>::Vector;
using Mat = NonLinearProblem<double>::Matrix;

struct ConstrainedRosenbrock2D : public NonLinearProblem<double> {  // hand-derived gradients (the reference uses Eigen's AutoDiff)
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
        l(0) = -std::numeric_limits<double>::infinity(); l(1) = 1;
    }
    void constraint_linearized(const Vec &x, Mat &Jc, Vec &c, Vec &l, Vec &u) override {
        constraint(x, c, l, u);
        Jc(0, 0) = 1; Jc(0, 1) = -1; Jc(1, 0) = 2 * x(0); Jc(1, 1) = 2 * x(1);
    }
};

int main(int argc, char **argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 4096, runs = argc > 2 ? atoi(argv[2]) : 3, device = argc > 3 ? atoi(argv[3]) : 0;
    std::vector<ConstrainedRosenbrock2D> probs(B);
    std::vector<NonLinearProblem<double> *> ptrs;
    std::vector<Vec> x0, l0;
    for (int i = 0; i < B; ++i) {
        ptrs.push_back(&probs[i]);
        // deterministic starts on a 64 x 64 grid over [-0.6, 0.6]^2 (the same as tests/cpp/sqp_cli.cpp and oracle-side bench.py)
        Vec x(2), l(2);
        x(0) = -0.6 + 1.2 * (i % 64) / 63.0 + 1e-3 * (i / 4096);
        x(1) = -0.6 + 1.2 * ((i / 64) % 64) / 63.0;
        l.setZero();
        x0.push_back(x);
        l0.push_back(l);
    }
    try {
        BatchSQP batch(ptrs, device);
        batch.settings().max_iter = 100;
        double best = 1e30;
        for (int r = 0; r < runs + 1; ++r) {  // first run warms up (context, staging buffers)
            const auto t0 = std::chrono::steady_clock::now();
            batch.solve(x0, l0);
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (r > 0 && sec < best) best = sec;
        }
        long long qp_iters = 0;
        int solved = 0;
        double checksum = 0;
        for (int i = 0; i < B; ++i) {
            qp_iters += batch.info(i).qp_solver_iter;
            solved += batch.info(i).status == SOLVED;
            if (batch.info(i).status == SOLVED) checksum += batch.primal_solution(i)(0) + batch.primal_solution(i)(1);
        }
        const double *ph = batch.phase_seconds();  // of the last run
        printf("{\"batch\": %d, \"runs\": %d, \"seconds\": %.6f, \"sqp_per_s\": %.1f, \"qp_launches\": %d, \"solved\": %d, "
               "\"qp_solver_iter_total\": %lld, \"checksum_solved_x\": %.12g, \"hessian_repairs\": %lld, "
               "\"phase_seconds\": {\"form_and_pack_qps_host\": %.6f, \"gpu_qp_calls\": %.6f, \"unpack_host\": %.6f, \"line_search_step_host\": %.6f}}\n",
               B, runs, best, B / best, batch.qp_launches(), solved, qp_iters, checksum, batch.hessian_repairs(), ph[0], ph[1], ph[2], ph[3]);
    } catch (const std::exception &e) {
        fprintf(stderr, "batch_sqp_bench: %s\n", e.what());
        return 1;
    }
    return 0;
}
