// Damped BFGS update for SQP -- host side (the north star keeps BFGS and the SQP outer loop on the host).
// Same signature and semantics as the reference's include/solvers/bfgs.hpp:15-41
// (Procedure 18.2 "Damped BFGS updating for SQP", Nocedal & Wright): B is updated in place from the
// step s = x - x_prev and the Lagrangian-gradient change y = grad - grad_prev.
#pragma once
#include <limits>
#include <vector>

template <typename Mat, typename Vec>
void BFGS_update(Mat &B, const Vec &s, const Vec &y) {
    using Scalar = typename Mat::Scalar;
    const std::ptrdiff_t n = s.rows();
    std::vector<Scalar> Bs((size_t)n), r((size_t)n);
    Scalar sBs = 0, sy = 0, sr;
    for (std::ptrdiff_t i = 0; i < n; ++i) {
        Scalar acc = 0;
        for (std::ptrdiff_t j = 0; j < n; ++j) acc += B(i, j) * s(j);
        Bs[(size_t)i] = acc;
    }
    for (std::ptrdiff_t i = 0; i < n; ++i) {
        sBs += s(i) * Bs[(size_t)i];
        sy += s(i) * y(i);
    }
    if (sy < Scalar(0.2) * sBs) {
        // damped update keeps B positive definite: r = theta y + (1 - theta) B s
        const Scalar theta = Scalar(0.8) * sBs / (sBs - sy);
        for (std::ptrdiff_t i = 0; i < n; ++i) r[(size_t)i] = theta * y(i) + (1 - theta) * Bs[(size_t)i];
        sr = theta * sy + (1 - theta) * sBs;
    } else {
        for (std::ptrdiff_t i = 0; i < n; ++i) r[(size_t)i] = y(i);
        sr = sy;
    }
    if (sr < std::numeric_limits<Scalar>::epsilon()) return;
    for (std::ptrdiff_t j = 0; j < n; ++j)
        for (std::ptrdiff_t i = 0; i < n; ++i) B(i, j) += -Bs[(size_t)i] * Bs[(size_t)j] / sBs + r[(size_t)i] * r[(size_t)j] / sr;
}
