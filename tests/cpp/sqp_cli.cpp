// Prints one JSON line per SQP test problem solved by sqp::SQP<double> (QP subproblems on the GPU), for the
// Python parity test that compares the trajectory with the CPU oracle (oracle/sqp_oracle.c).
#include <cstdio>

#include "sqp_problems.hpp"

static void report(const char *name, SQP<double> &s) {
    printf("{\"name\": \"%s\", \"iter\": %d, \"qp_solver_iter\": %d, \"status\": %d, \"x\": [", name, s.info().iter,
           s.info().qp_solver_iter, (int)s.info().status);
    for (int i = 0; i < s.primal_solution().rows(); ++i) printf("%s%.17g", i ? ", " : "", s.primal_solution()(i));
    printf("], \"lambda\": [");
    for (int i = 0; i < s.dual_solution().rows(); ++i) printf("%s%.17g", i ? ", " : "", s.dual_solution()(i));
    printf("]}\n");
}

int main() {
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
    return 0;
}
