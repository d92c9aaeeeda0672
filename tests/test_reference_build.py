"""The oracle pinned to THE REFERENCE'S OWN CODE (CPU tests, no GPU).

oracle/_ref is /root/reference/src/qp.cpp + src/sqp.cpp compiled unmodified against oracle/eigen_lite (stand-in for the absent
Eigen; `make -C oracle ref`). Here:
  * the reference's own unit tests (tests/qp_solver_test.cpp, sqp_test.cpp, bfgs_test.cpp; built against gtest_lite) pass,
  * the C restatement (oracle/qp_oracle_impl.h, sqp_oracle.c) reproduces the reference build BIT FOR BIT: fused solves over
    random shapes and settings, the object API (setup / solve / warm solve / update_qp), the float instantiation,
    constraint classification, NaN inputs, and whole SQP trajectories,
  * the committed outputs of the reference build (tests/golden/reference_outputs.json) are reproduced by the oracle (this part
    also runs where /root/reference does not exist).
"""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref(oracle):
    from oracle import ref_build

    ref_build.build()
    if not ref_build.available():
        pytest.skip("oracle/_ref is not built and /root/reference is not present")
    return ref_build


@pytest.fixture(scope="module")
def ref_golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_outputs.json")) as f:
        return json.load(f)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(((a == b) | ((a != a) & (b != b))).all()) if a.dtype.kind == "f" else bool((a == b).all())


FIELDS = ("status", "iter", "rho_updates", "x", "y", "res_prim", "res_dual", "rho_estimate")


def test_reference_unit_tests_pass(ref):
    """/root/reference/tests/{qp_solver,sqp,bfgs}_test.cpp, unmodified, on the reference's own sources."""
    r = ref.run_reference_tests()
    assert r.returncode == 0, r.stdout[-3000:]
    assert "11 tests ran, 0 failed" in r.stdout
    for name in ("QPSolverTest.testSimpleQP", "QPSolverTest.testSinglePrecisionFloat", "QPSolverTest.testConstraintViolation",
                 "QPSolverTest.testAdaptiveRho", "QPSolverTest.testAdaptiveRhoImprovesConvergence", "QPSolverTest.TestConstraint",
                 "SQPTestCase.TestSimpleNLP", "SQPTestCase.SimpleNLP_InfeasibleStart", "SQPTestCase.TestSimpleQP",
                 "BFGSTestCase.Test2D_posdef", "BFGSTestCase.Test2D_indefinite"):
        assert "[       OK ] " + name in r.stdout, name


def test_oracle_is_bit_identical_to_the_reference_build(ref, oracle):
    from sqp_solver_b200.synth import make_batch

    rng = np.random.default_rng(7)
    shapes = [(64, 128), (32, 64), (2, 3), (2, 2), (1, 0), (7, 0), (1, 9), (40, 13)] + [(int(rng.integers(1, 50)), int(rng.integers(0, 90))) for _ in range(8)]
    for case, (n, m) in enumerate(shapes):
        B = 6 if n * m > 2000 else 12
        d = make_batch(B, n, m, seed0=8000 + 37 * case)
        if case % 3 == 0:
            d["P"][1, 0] = np.nan  # LDLT failure -> NUMERICAL_ISSUES (qp.cpp:39-43)
            d["l"][2, :1] = -np.inf
        kws = [{}, dict(alpha=1.6, adaptive_rho=1),
               dict(alpha=float(rng.choice([1.0, 1.6, 1.8])), adaptive_rho=1, adaptive_rho_interval=int(rng.choice([7, 25, 50])),
                    check_termination=int(rng.choice([0, 1, 10, 25])), max_iter=int(rng.choice([1, 60, 300])), rho=float(rng.choice([0.05, 1.0])),
                    sigma=float(rng.choice([1e-6, 1e-4])), eps_abs=float(rng.choice([1e-3, 1e-5])), eps_rel=float(rng.choice([1e-3, 1e-5])),
                    adaptive_rho_tolerance=float(rng.choice([2.0, 5.0])))]
        for kw in kws:
            s = oracle.default_settings(**kw)
            a = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], s)
            b = ref.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], s)
            for k in FIELDS:
                assert same(a[k], b[k]), "n=%d m=%d %s: %s differs (max |diff| %.3e)" % (n, m, kw, k, np.nanmax(np.abs(np.asarray(a[k], float) - b[k])))


def test_oracle_float_instantiation_is_bit_identical(ref, oracle):
    import ctypes as C

    from sqp_solver_b200.synth import make_batch

    L = oracle.lib()
    for n, m in ((2, 3), (10, 14), (32, 64)):
        d = make_batch(8, n, m, seed0=8800 + n)
        for kw in ({}, dict(alpha=1.6, adaptive_rho=1)):
            s = oracle.default_settings(dtype=np.float32, **kw)
            b = ref.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], s, dtype=np.float32)
            f = {k: np.ascontiguousarray(d[k], dtype=np.float32) for k in ("P", "q", "A", "l", "u")}
            B = 8
            a = dict(x=np.zeros((B, n), np.float32), y=np.zeros((B, m), np.float32), z=np.zeros((B, m), np.float32),
                     status=np.zeros(B, np.int32), iter=np.zeros(B, np.int32), res_prim=np.zeros(B, np.float32),
                     res_dual=np.zeros(B, np.float32), rho_updates=np.zeros(B, np.int32), rho_estimate=np.zeros(B, np.float32))
            fp, ip = lambda v: v.ctypes.data_as(C.POINTER(C.c_float)), lambda v: v.ctypes.data_as(C.POINTER(C.c_int))
            L.oracle_qp_solve_batch_f32(C.byref(s), B, n, m, fp(f["P"]), fp(f["q"]), fp(f["A"]), fp(f["l"]), fp(f["u"]), fp(a["x"]), fp(a["y"]),
                                        fp(a["z"]), ip(a["status"]), ip(a["iter"]), fp(a["res_prim"]), fp(a["res_dual"]), ip(a["rho_updates"]),
                                        fp(a["rho_estimate"]), 1, None)
            for k in FIELDS:
                assert same(a[k], b[k]), "float n=%d m=%d %s: %s differs" % (n, m, kw, k)


def test_oracle_object_api_is_bit_identical(ref, oracle):
    """setup(); solve(); solve() (always warm: qp.cpp:78-82 discards a temporary); update_qp(); solve(); solve() before any
    setup() is a no-op (qp.cpp:68-71); rho_updates accumulates (qp.cpp:313)."""
    from sqp_solver_b200.synth import make_batch

    n, m = 12, 20
    d, d2 = make_batch(5, n, m, seed0=8900), make_batch(5, n, m, seed0=8950)
    for i in range(5):
        mk = lambda dd: oracle.QuadraticProblem(dd["P"][i].reshape(n, n, order="F"), dd["q"][i], dd["A"][i].reshape(m, n, order="F"), dd["l"][i], dd["u"][i])
        qp, qp2 = mk(d), mk(d2)
        o, r = oracle.QPSolver(), ref.QPSolver()

        def step(fn, q, **kw):
            for k, v in kw.items():
                setattr(o.settings(), k, v)
                setattr(r.settings(), k, v)
            getattr(o, fn)(q)
            getattr(r, fn)(q)
            x, y, info = r.get()
            oi = o.info()
            assert same(o.primal_solution(), x) and same(o.dual_solution(), y), fn
            for k in ("status", "iter", "rho_updates", "rho_estimate", "res_prim", "res_dual"):
                assert getattr(oi, k) == getattr(info, k), (fn, k, getattr(oi, k), getattr(info, k))

        r.solve(qp)  # before setup: a silent no-op, status stays UNINITIALIZED (qp.cpp:68-71)
        assert r.get()[2].status == oracle.UNINITIALIZED and o.info().status == oracle.UNINITIALIZED
        step("setup", qp, max_iter=40)
        step("solve", qp)
        step("solve", qp, max_iter=1000, adaptive_rho=1, adaptive_rho_interval=10, alpha=1.6)
        step("solve", qp)
        step("update_qp", qp2)
        step("solve", qp2)


def test_constraint_classification_matches(ref, oracle):
    l = np.array([-1e17, -1.0, -1e17, -3.0, 42.0, 0.0, -np.inf, 1.0])
    u = np.array([1e17, 1e17, 2.0, 4.0, 42.0, 5e-5, np.inf, 1.0 + 1e-4])
    np.testing.assert_array_equal(ref.constr_type_init(l, u), oracle.QPSolver.constr_type_init(l, u))
    np.testing.assert_array_equal(ref.constr_type_init(l, u)[:5], [2, 0, 0, 0, 1])  # tests/qp_solver_test.cpp:127-156


def test_sqp_oracle_is_bit_identical_to_the_reference_build(ref, oracle):
    """Whole SQP trajectories (src/sqp.cpp outer loop, bfgs.hpp, line search, SOC) on the reference's test problems and on a grid
    of constrained-Rosenbrock starts (BASELINE config 4's instances): same outer/inner iteration counts, status and iterate,
    and the same cumulative ADMM count after every outer iteration."""
    from oracle import sqp_oracle as S

    cases = [(S.CONSTRAINED_ROSENBROCK_2D, [0, 0], [0, 0], 0), (S.SIMPLE_NLP, [1.2, 0.1], [0, 0, 0], 1), (S.SIMPLE_NLP, [2, -1], [1, 1, 1], 1),
             (S.SIMPLE_NLP, [1.2, 0.1], [0, 0, 0], 0), (S.SIMPLE_QP, [0, 0], [0, 0, 0], 1), (S.SIMPLE_NLP2, [1.2, 0.1], [0], 0),
             (S.ROSENBROCK_BOX, [0, 0], [0, 0], 0), (S.ROSENBROCK_BOX, [0, 0, 0], [0, 0, 0], 0)]
    for gx in np.linspace(-0.6, 0.6, 5):
        for gy in np.linspace(-0.6, 0.6, 5):
            cases.append((S.CONSTRAINED_ROSENBROCK_2D, [gx, gy], [0, 0], 0))
    for pid, x0, l0, soc in cases:
        st = S.default_settings(second_order_correction=soc)
        a = S.solve(pid, x0, l0, st, n=len(x0), trace_cap=512)
        b = ref.sqp_solve(pid, x0, l0, st, n=len(x0), trace_cap=128)
        assert (a["iter"], a["qp_solver_iter"], a["status"]) == (b["iter"], b["qp_solver_iter"], b["status"]), (pid, x0)
        assert same(a["x"], b["x"]) and same(a["lam"], b["lam"]), (pid, x0, a["x"], b["x"])
        per_outer = np.cumsum(a["qps"]["iter"][:a["qps"]["count"]])[(1 if soc else 0)::(2 if soc else 1)]
        np.testing.assert_array_equal(per_outer, b["trace"]["qp_solver_iter"][:len(per_outer)])


def test_oracle_reproduces_committed_reference_outputs(oracle, ref_golden):
    """tests/golden/reference_outputs.json holds outputs of the reference build (tests/golden/make_reference_golden.py); the oracle
    must reproduce them exactly. Runs anywhere (no /root/reference needed)."""
    from oracle import sqp_oracle as S
    from sqp_solver_b200.synth import make_batch

    assert "reference src/qp.cpp" in ref_golden["how"]
    for c in ref_golden["qp"]:
        d = make_batch(c["batch"], c["n"], c["m"], seed0=c["seed0"])
        a = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle.default_settings(**c["settings"]), nthreads=1)
        for k in FIELDS:
            assert same(a[k], np.array(c[k])), (c["name"], k)
    for c in ref_golden["sqp"]:
        a = S.solve(c["pid"], c["x0"], c["l0"], S.default_settings(second_order_correction=c["soc"]), n=len(c["x0"]))
        assert (a["iter"], a["qp_solver_iter"], a["status"]) == (c["iter"], c["qp_solver_iter"], c["status"]), c["name"]
        assert same(a["x"], np.array(c["x"])) and same(a["lam"], np.array(c["lam"])), c["name"]
