"""GPU tests of the thread-per-QP literal KKT kernel (csrc/qp_small.cu; n + m <= 16, the SQP regime).

The kernel restates the reference's formulation operation by operation -- full pivoted KKT LDL^T (Eigen::LDLT<MatrixXd, Lower>,
qp.hpp:129) and a substitution per iteration (qp.cpp:90), no FMA contraction -- so the bar here is BIT identity with the CPU
oracle on everything the solver returns, including ill-conditioned BFGS-like Hessians and rounding-noise-decided adaptive-rho steps
where the Schur-complement kernels can only agree to the conditioning of the problem."""
import numpy as np
import pytest

from helpers import oracle_settings_from

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


def assert_bit_identical(got, ref, what, fields=("status", "iter", "rho_updates", "x", "y", "z", "res_prim", "res_dual", "rho_estimate")):
    for k in fields:
        a, b = np.asarray(got[k]), np.asarray(ref[k])
        same = (a == b) | (np.isnan(a) & np.isnan(b)) if a.dtype.kind == "f" else (a == b)
        assert same.all(), "%s: %s differs on instances %s (max abs diff %.3e)" % (
            what, k, np.unique(np.nonzero(~same)[0])[:10], np.nanmax(np.abs(a.astype(float) - b.astype(float))))


def ill_conditioned_batch(batch, n, m, seed, cond=1e12):
    """BFGS-like Hessians (cond up to `cond`), a mix of equality / inequality / one-sided / loose rows"""
    rng = np.random.default_rng(seed)
    P = np.zeros((batch, n * n))
    A = rng.standard_normal((batch, m * n))
    q = 60.0 * rng.standard_normal((batch, n))
    l = np.zeros((batch, m))
    u = np.zeros((batch, m))
    for i in range(batch):
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        ev = np.logspace(0, -np.log10(cond) * rng.uniform(0.3, 1.0), n) * 10.0 ** rng.uniform(-2, 3)
        Pi = (Q * ev) @ Q.T
        P[i] = ((Pi + Pi.T) / 2).reshape(-1, order="F")
        c = A[i].reshape(m, n, order="F") @ rng.standard_normal(n)
        kind = rng.integers(0, 4, m)
        l[i] = np.where(kind == 0, c, np.where(kind == 1, c - rng.uniform(0, 1, m), np.where(kind == 2, -np.inf, -1e20)))
        u[i] = np.where(kind == 0, c, np.where(kind == 1, c + rng.uniform(0, 1, m), np.where(kind == 2, c, 1e20)))
    return dict(P=P, q=q, A=np.ascontiguousarray(A), l=l, u=u, n=n, m=m, batch=batch)


SMALL_SHAPES = [(2, 2), (2, 3), (3, 3), (2, 1), (1, 1), (4, 4), (5, 7), (8, 8), (1, 15), (15, 1), (16, 0), (3, 0), (6, 10)]


@pytest.mark.parametrize("n,m", SMALL_SHAPES)
def test_small_kernel_is_bit_identical_to_the_oracle(api, ctx, oracle, n, m):
    from sqp_solver_b200.synth import make_batch

    cases = [("S1", api.default_settings()), ("S2", api.default_settings(alpha=1.6, adaptive_rho=1)), ("sqp_ctor", api.sqp_ctor_settings()),
             ("odd", api.default_settings(alpha=1.8, adaptive_rho=1, adaptive_rho_interval=7, check_termination=3, max_iter=300,
                                          adaptive_rho_tolerance=2.0, eps_abs=1e-5, eps_rel=1e-5, sigma=1e-4, rho=1.0))]
    for name, s in cases:
        for d in (make_batch(67, n, m, seed0=52000 + 100 * n + m), ill_conditioned_batch(45, n, m, seed=53000 + 100 * n + m)):
            b = api.QPBatch(ctx, d["batch"], n, m)
            b.settings = s
            b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
            got = b.get()
            assert ctx.last_kernel.startswith("small<"), ctx.last_kernel
            ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
            assert_bit_identical(got, ref, "%s n=%d m=%d" % (name, n, m))
            assert b.total_iters() == int(np.minimum(ref["iter"], s.max_iter)[ref["status"] != api.NUMERICAL_ISSUES].sum())
            b.close()


def test_small_kernel_object_api_and_nan(api, ctx, oracle):
    """setup(); solve(); solve() (warm); update_qp(); solve() through separate launches (the factor is rebuilt from the
    stored classes and rho) -- bit identical to the oracle's object API; NaN instances report NUMERICAL_ISSUES and stay untouched."""
    from sqp_solver_b200.synth import make_batch

    B, n, m = 40, 3, 4
    d, d2 = make_batch(B, n, m, seed0=54000), make_batch(B, n, m, seed0=55000)
    d["P"][7, n * n - 1] = np.nan
    sols = [oracle.QPSolver() for _ in range(B)]
    mk = lambda dd, i: oracle.QuadraticProblem(dd["P"][i].reshape(n, n, order="F"), dd["q"][i], dd["A"][i].reshape(m, n, order="F"), dd["l"][i], dd["u"][i])
    qps, qps2 = [mk(d, i) for i in range(B)], [mk(d2, i) for i in range(B)]
    state = lambda: dict(x=np.array([s.primal_solution() for s in sols]), y=np.array([s.dual_solution() for s in sols]),
                         z=np.array([s.z() for s in sols]), status=np.array([s.info().status for s in sols]),
                         iter=np.array([s.info().iter for s in sols]), rho_updates=np.array([s.info().rho_updates for s in sols]),
                         res_prim=np.array([s.info().res_prim for s in sols]), res_dual=np.array([s.info().res_dual for s in sols]),
                         rho_estimate=np.array([s.info().rho_estimate for s in sols]))
    args, args2 = (d["P"], d["q"], d["A"], d["l"], d["u"]), (d2["P"], d2["q"], d2["A"], d2["l"], d2["u"])
    b = api.QPBatch(ctx, B, n, m)
    b.solve(*args)  # before setup: silent no-op (qp.cpp:68-71)
    assert (b.info()["status"] == api.UNINITIALIZED).all()

    def both(settings_kw, fn_name, a, q_list):
        for k, v in settings_kw.items():
            setattr(b.settings, k, v)
        getattr(b, fn_name)(*a)
        for s, qp in zip(sols, q_list):
            for k, v in settings_kw.items():
                setattr(s.settings(), k, v)
            getattr(s, fn_name)(qp)

    both(dict(max_iter=30), "setup", args, qps)
    assert ctx.last_kernel.startswith("small<")
    assert b.info()["status"][7] == api.NUMERICAL_ISSUES
    both({}, "solve", args, qps)
    assert_bit_identical(b.get(), state(), "first solve")
    both(dict(max_iter=1000, adaptive_rho=1, adaptive_rho_interval=10, alpha=1.6), "solve", args, qps)
    assert_bit_identical(b.get(), state(), "second (warm) solve")
    both({}, "solve", args, qps)
    assert_bit_identical(b.get(), state(), "third solve")
    both({}, "update_qp", args2, qps2)
    both({}, "solve", args2, qps2)
    assert_bit_identical(b.get(), state(), "update_qp + solve")
    b.close()


def test_small_kernel_fp32_is_bit_identical_to_the_float_oracle(api, ctx, oracle):
    """QPSolver<float> (qp.cpp:386) at SQP sizes: the float instantiation of the same literal arithmetic."""
    import ctypes as C

    from sqp_solver_b200.synth import make_batch

    L = oracle.lib()
    for n, m in ((2, 3), (5, 7)):
        d = make_batch(33, n, m, seed0=56000 + n)
        for kw in ({}, dict(alpha=1.6, adaptive_rho=1)):
            b = api.QPBatch(ctx, d["batch"], n, m)
            b.settings = api.default_settings(**kw)
            b.set_precision(True)
            b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
            got = b.get()
            assert ctx.last_kernel.startswith("small<") and ",f32>" in ctx.last_kernel, ctx.last_kernel
            f = {k: np.ascontiguousarray(d[k], dtype=np.float32) for k in ("P", "q", "A", "l", "u")}
            B = d["batch"]
            out = dict(x=np.zeros((B, n), np.float32), y=np.zeros((B, m), np.float32), z=np.zeros((B, m), np.float32),
                       status=np.zeros(B, np.int32), iter=np.zeros(B, np.int32), res_prim=np.zeros(B, np.float32),
                       res_dual=np.zeros(B, np.float32), rho_updates=np.zeros(B, np.int32), rho_estimate=np.zeros(B, np.float32))
            s = oracle.default_settings(dtype=np.float32, **kw)
            fp, ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_float)), lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
            L.oracle_qp_solve_batch_f32(C.byref(s), B, n, m, fp(f["P"]), fp(f["q"]), fp(f["A"]), fp(f["l"]), fp(f["u"]), fp(out["x"]),
                                        fp(out["y"]), fp(out["z"]), ip(out["status"]), ip(out["iter"]), fp(out["res_prim"]),
                                        fp(out["res_dual"]), ip(out["rho_updates"]), fp(out["rho_estimate"]), 1, None)
            ref = {k: (v.astype(np.float64) if v.dtype == np.float32 else v) for k, v in out.items()}
            assert_bit_identical(got, ref, "fp32 n=%d m=%d %s" % (n, m, kw))
            b.close()


def test_small_kernel_solves_sqp_generated_subproblems_exactly(api, ctx, oracle):
    """Every QP an SQP run of the reference's test problems generates (BFGS Hessians up to cond(P) ~ 1e14, steps 1e-5 next to
    |q| ~ 60, infinite bounds, an infeasible subproblem whose duals diverge): bit identical, no tolerance at all."""
    from oracle import sqp_oracle as S

    for pid, x0, l0, soc in ((S.CONSTRAINED_ROSENBROCK_2D, [0, 0], [0, 0], 0), (S.SIMPLE_NLP, [2, -1], [1, 1, 1], 1),
                             (S.SIMPLE_NLP, [1.2, 0.1], [0, 0, 0], 0), (S.SIMPLE_QP, [0, 0], [0, 0, 0], 1),
                             (S.ROSENBROCK_BOX, [0, 0], [0, 0], 0)):
        tr = S.solve(pid, x0, l0, S.default_settings(second_order_correction=soc), n=len(x0), trace_cap=512)["qps"]
        k, nx, nc = tr["count"], tr["q"].shape[1], tr["l"].shape[1]
        b = api.QPBatch(ctx, k, nx, nc)
        b.settings = api.sqp_ctor_settings()
        b.setup_solve(tr["P"], tr["q"], tr["A"], tr["l"], tr["u"])
        assert ctx.last_kernel.startswith("small<")
        assert_bit_identical(b.get(), tr, "SQP-generated QPs of problem %d" % pid, fields=("status", "iter", "x", "y"))
        b.close()


def test_small_kernel_device_pointers_streams_and_host_staging(api, ctx, oracle):
    """device-pointer call on a non-default torch stream == host-pointer call (chunked staging with the ready flag); get() with no
    stream argument is ordered behind the launch (it used to read on the legacy stream)."""
    import torch

    from sqp_solver_b200.synth import make_batch

    d = make_batch(5000, 2, 2, seed0=57000)
    b = api.QPBatch(ctx, 5000, 2, 2)
    b.settings = api.sqp_ctor_settings()
    b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
    host = b.get()
    dev = {k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")}
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    b2 = api.QPBatch(ctx, 5000, 2, 2)
    b2.settings = api.sqp_ctor_settings()
    with torch.cuda.stream(st):
        b2.setup_solve(dev["P"], dev["q"], dev["A"], dev["l"], dev["u"], stream=st.cuda_stream)
    got = b2.get()  # no explicit synchronisation
    for k in ("x", "y", "status", "iter"):
        np.testing.assert_array_equal(got[k], host[k], err_msg=k)
    ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, b.settings))
    assert_bit_identical(got, ref, "5000 x (2,2)")
    b.close()
    b2.close()
