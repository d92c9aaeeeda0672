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


def make_sparse_batch(batch, n, m, density=0.03, seed0=0, pattern_seed=1):
    """BASELINE.json config 5: QPs whose constraint matrices share ONE sparsity pattern (the Jacobian pattern of a batch of
    same-structure NLPs). Pattern: each entry stored with probability `density`, plus one forced entry per row (no empty rows);
    values and P, q as in make_qp; bounds rebuilt around c = A_sparse x0 so that every instance stays feasible.
    Returns the dense-free dict: P[B,n*n] q[B,n] l[B,m] u[B,m], CSR pattern outer[m+1], inner[nnz] (int32) and vals[B,nnz]."""
    prng = np.random.default_rng(pattern_seed)
    mask = prng.uniform(size=(m, n)) < density
    mask[np.arange(m), prng.integers(0, n, m)] = True
    rows, cols = np.nonzero(mask)
    outer = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int32)
    inner = np.ascontiguousarray(cols.astype(np.int32))
    nnz = int(inner.shape[0])
    P = np.empty((batch, n * n))
    q = np.empty((batch, n))
    l = np.empty((batch, m))
    u = np.empty((batch, m))
    vals = np.empty((batch, nnz))
    for i in range(batch):
        rng = np.random.default_rng(seed0 + i)
        M = rng.standard_normal((n, n)) / np.sqrt(n)
        P[i] = (M @ M.T + 1e-2 * np.eye(n)).reshape(-1, order="F")
        q[i] = rng.standard_normal(n)
        v = rng.standard_normal(nnz) / np.sqrt(n)
        x0 = rng.standard_normal(n)
        c = np.bincount(rows, weights=v * x0[cols], minlength=m)
        li = c - rng.uniform(0.0, 1.0, m)
        ui = c + rng.uniform(0.0, 1.0, m)
        k = rng.uniform(0.0, 1.0, m)
        eq = k < 0.1
        li[eq] = c[eq]
        ui[eq] = c[eq]
        loose = k > 0.9
        li[loose] = -1e20
        ui[loose] = 1e20
        l[i], u[i], vals[i] = li, ui, v
    return dict(P=P, q=q, l=l, u=u, vals=vals, outer=outer, inner=inner, rows=rows, cols=cols, nnz=nnz, n=n, m=m, batch=batch,
                seed0=seed0, density=density)


def densify(d, first=0, count=None):
    """Dense column-major A[B, m*n] of instances [first, first+count) of a make_sparse_batch dict (for the CPU oracle)."""
    count = d["batch"] - first if count is None else count
    n, m = d["n"], d["m"]
    A = np.zeros((count, n, m))
    A[:, d["cols"], d["rows"]] = d["vals"][first:first + count]
    return A.reshape(count, n * m)
