// Prints one JSON line per SQP test problem solved by sqp::SQP<double> (QP subproblems on the GPU), for the
// Python parity test that compares the trajectory with the CPU oracle (oracle/sqp_oracle.c).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "sqp_problems.hpp"

static void report(const char *name, SQP<double> &s) {
    printf("{\"name\": \"%s\", \"iter\": %d, \"qp_solver_iter\": %d, \"status\": %d, \"x\": [", name, s.info().iter,
           s.info().qp_solver_iter, (int)s.info().status);
    for (int i = 0; i < s.primal_solution().rows(); ++i) printf("%s%.17g", i ? ", " : "", s.primal_solution()(i));
    printf("], \"lambda\": [");
    for (int i = 0; i < s.dual_solution().rows(); ++i) printf("%s%.17g", i ? ", " : "", s.dual_solution()(i));
    printf("]}\n");
}

// BASELINE.json config 4: a batch of constrained-Rosenbrock instances (BFGS Hessian, host outer loop, one batched GPU QP
// solve per outer iteration). Prints one JSON line with timing and per-instance results for `sample` instances.
static int run_batch(int B, int sample) {
    std::vector<ConstrainedRosenbrock2D> probs(B);
    std::vector<NonLinearProblem<double> *> ptrs;
    std::vector<Vec> x0, l0;
    for (int i = 0; i < B; ++i) {
        ptrs.push_back(&probs[i]);
        // deterministic starts on a 64 x 64 grid over [-0.6, 0.6]^2, nudged off the feasible circle
        const double gx = -0.6 + 1.2 * (i % 64) / 63.0, gy = -0.6 + 1.2 * ((i / 64) % 64) / 63.0;
        x0.push_back(v2(gx + 1e-3 * (i / 4096), gy));
        l0.push_back(zeros(2));
    }
    BatchSQP batch(ptrs);
    batch.settings().max_iter = 100;
    const auto t0 = std::chrono::steady_clock::now();
    batch.solve(x0, l0);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long long qp_iters = 0;
    int solved = 0, feasible = 0;
    for (int i = 0; i < B; ++i) {
        qp_iters += batch.info(i).qp_solver_iter;
        if (batch.info(i).status == SOLVED) {
            ++solved;
            const auto &x = batch.primal_solution(i);
            if (x(0) - x(1) <= 1e-3 && std::abs(x(0) * x(0) + x(1) * x(1) - 1.0) <= 1e-3) ++feasible;
        }
    }
    printf("{\"name\": \"batch\", \"batch\": %d, \"seconds\": %.6f, \"sqp_per_s\": %.1f, \"qp_launches\": %d, \"solved\": %d, "
           "\"solved_feasible\": %d, \"qp_solver_iter_total\": %lld, \"instances\": [",
           B, sec, B / sec, batch.qp_launches(), solved, feasible, qp_iters);
    for (int k = 0; k < sample; ++k) {
        const int i = (int)((long long)k * B / sample);
        printf("%s{\"i\": %d, \"x0\": [%.17g, %.17g], \"iter\": %d, \"qp_solver_iter\": %d, \"status\": %d, \"x\": [%.17g, %.17g]}", k ? ", " : "",
               i, x0[i](0), x0[i](1), batch.info(i).iter, batch.info(i).qp_solver_iter, (int)batch.info(i).status,
               batch.primal_solution(i)(0), batch.primal_solution(i)(1));
    }
    printf("]}\n");
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 3 && std::string(argv[1]) == "--batch") return run_batch(atoi(argv[2]), argc >= 4 ? atoi(argv[3]) : 32);
    {
        ConstrainedRosenbrock2D p; SQP<double> s; s.settings().max_iter = 100;
        s.solve(p, v2(0, 0), zeros(2)); report("ConstrainedRosenbrock2D", s);
    }
    {
        SimpleNLP p; SQP<double> s; s.settings().max_iter = 100; s.settings().second_order_correction = true;
        s.solve(p, v2(1.2, 0.1), zeros(3)); report("SimpleNLP_feasible_SOC", s);
    }
    {
        SimpleNLP p; SQP<double> s; s.settings().max_iter = 100; s.settings().second_order_correction = true;
        Vec y0(3); y0.setConstant(1);
        s.solve(p, v2(2, -1), y0); report("SimpleNLP_infeasible_SOC", s);
    }
    {
        SimpleQPasNLP p; SQP<double> s; s.settings().second_order_correction = true;
        s.solve(p, v2(0, 0), zeros(3)); report("SimpleQP_as_NLP_SOC", s);
    }
    {
        SimpleNLP2 p; SQP<double> s;
        s.solve(p, v2(1.2, 0.1), zeros(1)); report("SimpleNLP2", s);
    }
    {
        RosenbrockBox p(2); SQP<double> s; s.settings().max_iter = 100;
        s.solve(p, zeros(2), zeros(2)); report("RosenbrockBox2", s);
    }
    return 0;
}
