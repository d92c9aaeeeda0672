"""Synthetic batched dense QPs: the exact recipe of SURVEY.md section 8(d) / BASELINE.md section 4.

Per QP i, rng = numpy.random.default_rng(seed0 + i) (PCG64), draws in this order:
  M = N(0,1)[n,n]/sqrt(n);  P = M M^T + 1e-2 I   (symmetric positive definite)
  q = N(0,1)[n];            A = N(0,1)[m,n]/sqrt(n)  (dense)
  x0 = N(0,1)[n]; c = A x0; l = c - U(0,1)[m]; u = c + U(0,1)[m]; k = U(0,1)[m]
  rows with k < 0.1 become equalities l = u = c; rows with k > 0.9 become loose (-1e20, +1e20).
Always feasible (x0). Arrays are batch-major; each matrix is stored COLUMN-major, i.e.
bit-compatible with Eigen::MatrixXd::data() (the reference's QuadraticProblem, qp.hpp:19-34).
"""
import numpy as np


def make_qp(n, m, seed):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n)) / np.sqrt(n)
    P = M @ M.T + 1e-2 * np.eye(n)
    q = rng.standard_normal(n)
    A = rng.standard_normal((m, n)) / np.sqrt(n)
    x0 = rng.standard_normal(n)
    c = A @ x0
    l = c - rng.uniform(0.0, 1.0, m)
    u = c + rng.uniform(0.0, 1.0, m)
    k = rng.uniform(0.0, 1.0, m)
    eq = k < 0.1
    l[eq] = c[eq]
    u[eq] = c[eq]
    loose = k > 0.9
    l[loose] = -1e20
    u[loose] = 1e20
    return P, q, A, l, u


def make_batch(batch, n, m, seed0=0):
    """Returns dict of float64 arrays P[B,n*n] q[B,n] A[B,m*n] l[B,m] u[B,m] (matrices column-major)."""
    P = np.empty((batch, n * n))
    q = np.empty((batch, n))
    A = np.empty((batch, m * n))
    l = np.empty((batch, m))
    u = np.empty((batch, m))
    for i in range(batch):
        Pi, qi, Ai, li, ui = make_qp(n, m, seed0 + i)
        P[i] = Pi.reshape(-1, order="F")
        q[i] = qi
        A[i] = Ai.reshape(-1, order="F")
        l[i] = li
        u[i] = ui
    return dict(P=P, q=q, A=A, l=l, u=u, n=n, m=m, batch=batch, seed0=seed0)
