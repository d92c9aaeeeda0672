"""Pins the C oracle (oracle/qp_oracle.c) against every known-answer assertion the reference's
own tests hold for the QP hot path: /root/reference/tests/qp_solver_test.cpp:43-156.
Each test mirrors one TEST(QPSolverTest, ...) of that file. CPU only."""
import numpy as np


def is_approx(a, b, prec):
    """Eigen's isApprox: ||a-b||^2 <= prec^2 * min(||a||^2, ||b||^2)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.sum((a - b) ** 2) <= prec * prec * min(np.sum(a * a), np.sum(b * b))


def simple_qp(oracle, golden, dtype=np.float64):
    g = golden["simple_qp"]
    return oracle.QuadraticProblem(g["P"], g["q"], g["A"], g["l"], g["u"], dtype=dtype), np.array(g["solution"])


def test_simple_qp(oracle, golden):  # tests/qp_solver_test.cpp:43-56
    qp, sol = simple_qp(oracle, golden)
    solver = oracle.QPSolver()
    solver.settings().max_iter = 1000
    solver.setup(qp)
    solver.solve(qp)
    assert is_approx(solver.primal_solution(), sol, 1e-2)
    assert solver.info().iter < solver.settings().max_iter
    assert solver.info().status == oracle.SOLVED


def test_single_precision_float(oracle, golden):  # tests/qp_solver_test.cpp:58-69
    qp, sol = simple_qp(oracle, golden, np.float32)
    solver = oracle.QPSolver(np.float32)
    solver.setup(qp)
    solver.solve(qp)
    assert is_approx(solver.primal_solution(), sol, 1e-2)
    assert solver.info().iter < solver.settings().max_iter
    assert solver.info().status == oracle.SOLVED


def test_constraint_violation(oracle, golden):  # tests/qp_solver_test.cpp:71-87
    qp, _ = simple_qp(oracle, golden)
    solver = oracle.QPSolver()
    solver.settings().eps_rel = float(np.float32(1e-4))  # "1e-4f" in the reference
    solver.settings().eps_abs = float(np.float32(1e-4))
    solver.setup(qp)
    solver.solve(qp)
    sol = solver.primal_solution()
    assert (qp.A @ sol - qp.l).min() >= -1e-3
    assert (qp.A @ sol - qp.u).max() <= 1e-3


def test_adaptive_rho(oracle, golden):  # tests/qp_solver_test.cpp:89-100
    qp, _ = simple_qp(oracle, golden)
    solver = oracle.QPSolver()
    solver.settings().adaptive_rho = 1
    solver.settings().adaptive_rho_interval = 10
    solver.setup(qp)
    solver.solve(qp)
    assert solver.info().status == oracle.SOLVED


def test_adaptive_rho_improves_convergence(oracle, golden):  # tests/qp_solver_test.cpp:102-125
    qp, _ = simple_qp(oracle, golden)
    solver = oracle.QPSolver()
    solver.settings().warm_start = 0
    solver.settings().max_iter = 1000
    solver.settings().rho = 0.1
    solver.settings().adaptive_rho = 0
    solver.setup(qp)
    solver.solve(qp)
    prev_iter = solver.info().iter
    solver.settings().adaptive_rho = 1
    solver.settings().adaptive_rho_interval = 10
    solver.solve(qp)
    info = solver.info()
    assert info.iter < solver.settings().max_iter
    assert info.iter < prev_iter
    assert info.status == oracle.SOLVED
    g = golden["simple_qp"]["oracle_regression"]["improves_convergence"]
    assert (prev_iter, info.iter) == (g["first_iter"], g["second_iter"])
    np.testing.assert_allclose(solver.primal_solution(), g["second_x"], rtol=0, atol=1e-11)


def test_constraint_classification(oracle, golden):  # tests/qp_solver_test.cpp:127-156
    g = golden["test_constraint"]
    T = 1e16  # Solver::LOOSE_BOUNDS_THRESH
    l = [-10 * T, -1, -10 * T, -3, 42]
    u = [10 * T, 10 * T, 2, 4, 42]
    assert l == g["l"] and u == g["u"]
    got = oracle.QPSolver.constr_type_init(l, u)
    assert got.tolist() == g["type_expect"]
    assert (oracle.LOOSE_BOUNDS, oracle.INEQUALITY_CONSTRAINT, oracle.EQUALITY_CONSTRAINT) == (2, 0, 1)


def test_oracle_regression_values(oracle, golden):
    """Values recorded in SURVEY.md Appendix B.1 by an independent numpy restatement."""
    qp, _ = simple_qp(oracle, golden)
    reg = golden["simple_qp"]["oracle_regression"]

    def run(**kw):
        s = oracle.QPSolver()
        for k, v in kw.items():
            setattr(s.settings(), k, v)
        s.setup(qp)
        if "D" not in run.__dict__:
            run.D, run.T = s.ldlt_dump()
        s.solve(qp)
        return s

    e4 = float(np.float32(1e-4))
    cases = {
        "defaults": {},
        "eps_1e-4f": dict(eps_rel=e4, eps_abs=e4),
        "adaptive_interval_10": dict(adaptive_rho=1, adaptive_rho_interval=10),
        "sqp_ctor_settings": dict(warm_start=1, check_termination=10, eps_abs=1e-4, eps_rel=1e-4, max_iter=100,
                                  adaptive_rho=1, adaptive_rho_interval=50, alpha=1.6),  # src/sqp.cpp:16-23
    }
    for name, kw in cases.items():
        s = run(**kw)
        r = reg[name]
        i = s.info()
        assert (i.status, i.iter, i.rho_updates) == (r["status"], r["iter"], r["rho_updates"]), name
        np.testing.assert_allclose(s.primal_solution(), r["x"], rtol=0, atol=2e-12, err_msg=name)
        np.testing.assert_allclose(s.dual_solution(), r["y"], rtol=0, atol=2e-9, err_msg=name)
        if "rho" in r:
            assert abs(s.rho() - r["rho"]) < 1e-5
    np.testing.assert_allclose(run.D, reg["ldlt_D"], rtol=1e-11)
    assert run.T.tolist() == reg["ldlt_transpositions"]


def test_iter_is_max_iter_plus_one_on_exceed(oracle, golden):
    """src/qp.cpp:84,150."""
    qp, _ = simple_qp(oracle, golden)
    s = oracle.QPSolver()
    s.settings().max_iter = 30
    s.setup(qp)
    s.solve(qp)
    assert s.info().status == oracle.MAX_ITER_EXCEEDED
    assert s.info().iter == 31


def test_solve_without_setup_is_silent_noop(oracle, golden):
    """src/qp.cpp:68-71: UNINITIALIZED -> solve returns without touching anything."""
    qp, _ = simple_qp(oracle, golden)
    s = oracle.QPSolver()
    assert s.info().status == oracle.UNINITIALIZED
    # solve() would dereference unsized members in the reference only after the early return
    s.n, s.m = qp.n, qp.m
    s._f("oracle_qp_solve")(s._h, *s._args(qp))
    assert s.info().status == oracle.UNINITIALIZED and s.info().iter == 0


def test_nan_input_gives_numerical_issues(oracle, golden):
    """NaN in K => LDLT info()!=Success => NUMERICAL_ISSUES (qp.cpp:39-43)."""
    g = golden["simple_qp"]
    P = np.array(g["P"], dtype=float)
    P[1, 1] = np.nan
    qp = oracle.QuadraticProblem(P, g["q"], g["A"], g["l"], g["u"])
    s = oracle.QPSolver()
    s.setup(qp)
    assert s.info().status == oracle.NUMERICAL_ISSUES
    s.solve(qp)
    assert s.info().status == oracle.NUMERICAL_ISSUES and s.info().iter == 0


def test_ldlt_solve_matches_numpy(oracle):
    """Independent check of the restated Eigen::LDLT: K x = rhs against numpy.linalg.solve."""
    from sqp_solver_b200.synth import make_qp

    for n, m, seed in ((2, 3, 0), (8, 5, 1), (32, 64, 2), (64, 128, 3)):
        P, q, A, l, u = make_qp(n, m, seed)
        qp = oracle.QuadraticProblem(P, q, A, l, u)
        s = oracle.QPSolver()
        s.setup(qp)
        assert s.info().status == oracle.UNSOLVED
        typ = oracle.QPSolver.constr_type_init(l, u)
        rho = np.where(typ == 2, 1e-6, np.where(typ == 1, 100.0, 0.1))
        K = np.block([[P + 1e-6 * np.eye(n), A.T], [A, -np.diag(1.0 / rho)]])
        rhs = np.random.default_rng(seed).standard_normal(n + m)
        ref = np.linalg.solve(K, rhs)
        got = s.kkt_solve(rhs)
        assert np.linalg.norm(got - ref) <= 1e-9 * np.linalg.norm(ref)


def test_oracle_solution_is_the_qp_optimum(oracle):
    """ADMM fixed point check against an independent active-set-free KKT test: at tight
    tolerance the oracle's (x, y) satisfies stationarity, feasibility and complementarity."""
    from sqp_solver_b200.synth import make_qp

    n, m = 16, 24
    P, q, A, l, u = make_qp(n, m, 7)
    qp = oracle.QuadraticProblem(P, q, A, l, u)
    s = oracle.QPSolver()
    st = s.settings()
    st.eps_abs = st.eps_rel = 1e-9
    st.max_iter = 20000
    st.adaptive_rho = 1
    s.setup(qp)
    s.solve(qp)
    assert s.info().status == oracle.SOLVED
    x, y = s.primal_solution(), s.dual_solution()
    Ax = A @ x
    assert np.abs(P @ x + q + A.T @ y).max() < 1e-6
    assert (Ax >= l - 1e-6).all() and (Ax <= u + 1e-6).all()
    fin_l, fin_u = l > -1e16, u < 1e16
    assert (np.minimum(y, 0) * np.where(fin_l, Ax - l, 0.0)).__abs__().max() < 1e-5
    assert (np.maximum(y, 0) * np.where(fin_u, u - Ax, 0.0)).__abs__().max() < 1e-5


def test_batch_helper_matches_object_api(oracle):
    from sqp_solver_b200.synth import make_batch

    d = make_batch(6, 8, 12, seed0=100)
    out = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], nthreads=2)
    for i in range(6):
        qp = oracle.QuadraticProblem(d["P"][i].reshape(8, 8, order="F"), d["q"][i], d["A"][i].reshape(12, 8, order="F"),
                                     d["l"][i], d["u"][i])
        s = oracle.QPSolver()
        s.setup(qp)
        s.solve(qp)
        assert s.info().iter == out["iter"][i] and s.info().status == out["status"][i]
        np.testing.assert_array_equal(s.primal_solution(), out["x"][i])
        np.testing.assert_array_equal(s.dual_solution(), out["y"][i])


def test_oracle_reproduces_committed_golden_outputs(oracle):
    """tests/golden/oracle_synthetic.json (made by tests/golden/make_oracle_golden.py): the oracle built on THIS machine
    reproduces the committed outputs -- same status / iteration counts, x and y to 1e-9."""
    import json
    import os

    from sqp_solver_b200.synth import make_batch

    with open(os.path.join(os.path.dirname(__file__), "golden", "oracle_synthetic.json")) as f:
        gold = json.load(f)
    for c in gold["cases"]:
        d = make_batch(c["batch"], c["n"], c["m"], seed0=c["seed0"])
        r = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle.default_settings(**c["settings"]), nthreads=2)
        assert r["status"].tolist() == c["status"] and r["iter"].tolist() == c["iter"], c["name"]
        assert r["rho_updates"].tolist() == c["rho_updates"], c["name"]
        np.testing.assert_allclose(r["x"], np.array(c["x"]), rtol=0, atol=1e-9, err_msg=c["name"])
        np.testing.assert_allclose(r["y"], np.array(c["y"]), rtol=0, atol=1e-8, err_msg=c["name"])


def test_oracle_reproduces_committed_sparse_and_f32_golden(oracle):
    """tests/golden/oracle_sparse_f32.json: the densified sparse-A cases (BASELINE config 5's path) and the float instantiation."""
    import json
    import os

    from sqp_solver_b200.synth import densify, make_batch, make_sparse_batch

    with open(os.path.join(os.path.dirname(__file__), "golden", "oracle_sparse_f32.json")) as f:
        gold = json.load(f)
    for c in gold["sparse"][:2]:  # (the third case repeats the 256x512 shape with more iterations: left to the GPU test)
        d = make_sparse_batch(c["batch"], c["n"], c["m"], density=c["density"], seed0=c["seed0"])
        assert d["nnz"] == c["nnz"]
        r = oracle.solve_batch(d["P"], d["q"], densify(d), d["l"], d["u"], oracle.default_settings(**c["settings"]), nthreads=2)
        assert r["status"].tolist() == c["status"] and r["iter"].tolist() == c["iter"], c["name"]
        np.testing.assert_allclose(r["x"], np.array(c["x"]), rtol=0, atol=1e-9, err_msg=c["name"])
    for c in gold["f32"]:
        d = make_batch(c["batch"], c["n"], c["m"], seed0=c["seed0"])
        n, m = c["n"], c["m"]
        for i in range(c["batch"]):
            qp = oracle.QuadraticProblem(d["P"][i].reshape(n, n, order="F"), d["q"][i], d["A"][i].reshape(m, n, order="F"), d["l"][i],
                                         d["u"][i], dtype=np.float32)
            s = oracle.QPSolver(dtype=np.float32)
            s.setup(qp)
            s.solve(qp)
            assert int(s.info().status) == c["status"][i] and int(s.info().iter) == c["iter"][i], c["name"]
            np.testing.assert_allclose(s.primal_solution(), np.array(c["x"][i]), rtol=0, atol=2e-6, err_msg=c["name"])
