"""The object API with a sparse A: setup / solve / solve / update_qp / solve as separate launches on persistent state.
Mirrors the reference's (dead: QP_SOLVER_USE_SPARSE is never defined) tests/qp_solver_sparse_test.cpp:68-98 testCanMultipleSolve and
testCanUpdateQP exactly, then holds the same call sequence to oracle parity on the shapes of every sparse code path: densified into
the thread-per-QP / register-tiled kernels, the 4-CTA and 8-CTA cluster kernels (the factor is rebuilt by solve), the blocked kernel
(factor in the slab)."""
import numpy as np
import pytest

from helpers import assert_parity, is_approx, oracle_settings_from

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from sqp_solver_b200 import api

    return api


@pytest.fixture(scope="module")
def ctx(api):
    c = api.Context(0)
    yield c
    c.close()


@pytest.fixture(autouse=True)
def _reset(api, ctx):
    yield
    ctx.set_option(api.OPT_KERNEL, api.KERNEL_AUTO)


def simple_sparse_qp(golden):
    g = golden["simple_qp"]
    A = np.array(g["A"], dtype=float)
    cols, rows = np.nonzero(A.T)  # CSC = Eigen::SparseMatrix: A.sparseView()
    outer = np.concatenate([[0], np.cumsum((A != 0).sum(axis=0))]).astype(np.int32)
    vals = np.ascontiguousarray(A[rows, cols][None, :])
    P = np.array(g["P"], dtype=float).reshape(1, 4)
    q, l, u = (np.array(g[k], dtype=float).reshape(1, -1) for k in ("q", "l", "u"))
    return P, q, vals, outer, rows.astype(np.int32), l, u, g["solution"]


def test_reference_sparse_testCanMultipleSolve(api, ctx, golden):
    """tests/qp_solver_sparse_test.cpp:68-78: setup; solve -> SOLVED; solve -> SOLVED."""
    P, q, vals, outer, inner, l, u, _ = simple_sparse_qp(golden)
    b = api.QPBatch(ctx, 1, 2, 3)
    b.setup_sparse(P, q, vals, outer, inner, l, u)
    b.solve_sparse(P, q, vals, outer, inner, l, u)
    assert b.info()["status"][0] == api.SOLVED
    b.solve_sparse(P, q, vals, outer, inner, l, u)
    assert b.info()["status"][0] == api.SOLVED
    b.close()


def test_reference_sparse_testCanUpdateQP(api, ctx, golden):
    """tests/qp_solver_sparse_test.cpp:80-98: setup; solve -> [0.3, 0.7]; P = I, q = 0; update_qp; solve -> [0.5, 0.5]."""
    P, q, vals, outer, inner, l, u, sol = simple_sparse_qp(golden)
    b = api.QPBatch(ctx, 1, 2, 3)
    b.setup_sparse(P, q, vals, outer, inner, l, u)
    b.solve_sparse(P, q, vals, outer, inner, l, u)
    out = b.get()
    assert is_approx(out["x"][0], sol, 1e-2) and out["status"][0] == api.SOLVED
    P2, q2 = np.eye(2).reshape(1, 4), np.zeros((1, 2))
    b.update_qp_sparse(P2, q2, vals, outer, inner, l, u)
    b.solve_sparse(P2, q2, vals, outer, inner, l, u)
    out = b.get()
    assert is_approx(out["x"][0], [0.5, 0.5], 1e-2) and out["status"][0] == api.SOLVED
    b.close()


@pytest.mark.parametrize("n,m,batch,density,layout,expect", [
    (2, 3, 5, 0.7, "csc", "small<"), (40, 60, 6, 0.15, "csr", "tile<"), (100, 150, 5, 0.08, "csr", "cluster<4>"),
    (256, 512, 3, 0.03, "csr", "cluster<4>"), (200, 1100, 2, 0.006, "csc", "cluster<8>"), (130, 200, 3, 0.6, "csr", "block")])
def test_sparse_object_api_against_the_oracle(api, ctx, oracle, n, m, batch, density, layout, expect):
    """setup(); solve() [stops early]; solve() [warm, adaptive rho on]; update_qp() [new P, q, values, bounds: no reset of x, z, y];
    solve() -- every step against the oracle's object API on the densified problem."""
    from sqp_solver_b200.synth import densify, make_sparse_batch

    d, d2 = (make_sparse_batch(batch, n, m, density=density, seed0=s0, pattern_seed=3) for s0 in (71000, 72000))
    if layout == "csc":  # same pattern, column-compressed: reorder the stored entries
        order = np.lexsort((d["rows"], d["cols"]))
        for dd in (d, d2):
            dd["vals"] = np.ascontiguousarray(dd["vals"][:, order])
        outer = np.concatenate([[0], np.cumsum(np.bincount(d["cols"], minlength=n))]).astype(np.int32)
        inner = np.ascontiguousarray(d["rows"][order].astype(np.int32))
        rows_c, cols_c = d["rows"][order], d["cols"][order]
        for dd in (d, d2):
            dd["rows"], dd["cols"] = rows_c, cols_c
    else:
        outer, inner = d["outer"], d["inner"]
    lay = api.SPARSE_CSC if layout == "csc" else api.SPARSE_CSR
    A1, A2 = densify(d), densify(d2)
    sols = [oracle.QPSolver() for _ in range(batch)]
    mk = lambda dd, A, i: oracle.QuadraticProblem(dd["P"][i].reshape(n, n, order="F"), dd["q"][i], A[i].reshape(m, n, order="F"), dd["l"][i], dd["u"][i])
    qps, qps2 = [mk(d, A1, i) for i in range(batch)], [mk(d2, A2, i) for i in range(batch)]
    state = lambda: dict(x=np.array([s.primal_solution() for s in sols]), y=np.array([s.dual_solution() for s in sols]),
                         status=np.array([s.info().status for s in sols]), iter=np.array([s.info().iter for s in sols]),
                         rho_updates=np.array([s.info().rho_updates for s in sols]))
    b = api.QPBatch(ctx, batch, n, m)

    def both(fn, dd, q_list, **kw):
        for k, v in kw.items():
            setattr(b.settings, k, v)
        getattr(b, fn + "_sparse")(dd["P"], dd["q"], dd["vals"], outer, inner, dd["l"], dd["u"], layout=lay)
        for s, qp in zip(sols, q_list):
            for k, v in kw.items():
                setattr(s.settings(), k, v)
            getattr(s, fn)(qp)

    both("setup", d, qps, max_iter=40)
    assert ctx.last_kernel.startswith(expect), ctx.last_kernel
    assert (b.info()["status"] == api.UNSOLVED).all()
    both("solve", d, qps)
    assert ctx.last_kernel.startswith(expect), ctx.last_kernel
    assert_parity(b.get(), state(), what="first solve (%s)" % ctx.last_kernel)
    both("solve", d, qps, max_iter=1000, adaptive_rho=1, adaptive_rho_interval=25, alpha=1.6)
    assert_parity(b.get(), state(), what="second (warm) solve")
    both("update_qp", d2, qps2)
    both("solve", d2, qps2)
    assert_parity(b.get(), state(), what="update_qp + solve")
    b.close()
