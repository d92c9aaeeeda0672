"""GPU parity tests: the CUDA path, called through the C-ABI (ctypes), against the CPU oracle on
the same seeded inputs. Bar (BASELINE.json north_star): identical termination status (and, because
termination is only tested every check_termination iterations, identical iteration count) and
primal solution within 1e-6 relative. Every kernel variant is exercised."""
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


KERNELS = ["generic", "tile", "tile8"]  # tile8 = the non-default warps-per-QP variants (8 for the 64x128 class, 2 for 32x64)


def select_kernel(api, ctx, kernel, n, m):
    if kernel == "tile8" and not (n > 16 or m > 32):
        pytest.skip("alternative warps-per-QP variants only exist for the 32x64 and 64x128 classes")
    ctx.set_option(api.OPT_KERNEL, {"generic": api.KERNEL_GENERIC, "tile": api.KERNEL_TILE, "tile8": api.KERNEL_TILE,
                                    "auto": api.KERNEL_AUTO, "block": api.KERNEL_BLOCK}[kernel])
    ctx.set_option(api.OPT_TILE_WARPS, (8 if (n > 32 or m > 64) else 2) if kernel == "tile8" else 0)


@pytest.fixture(autouse=True)
def _reset_kernel_option(api, ctx):
    yield
    ctx.set_option(api.OPT_KERNEL, api.KERNEL_AUTO)
    ctx.set_option(api.OPT_TILE_WARPS, 0)


def run_fused(api, ctx, d, settings, kernel):
    select_kernel(api, ctx, kernel, d["n"], d["m"])
    b = api.QPBatch(ctx, d["batch"], d["n"], d["m"])
    b.settings = settings
    try:
        b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
    except api.SolverError as e:
        if kernel.startswith("tile") and "outside its range" in str(e):
            pytest.skip("shape not covered by the register-tiled kernel")
        raise
    out = b.get()
    out["kernel"] = ctx.last_kernel
    out["total_iters"] = b.total_iters()
    b.close()
    return out


def simple_qp_batch(golden, copies=1):
    g = golden["simple_qp"]
    P = np.array(g["P"], dtype=np.float64).reshape(-1, order="F")
    A = np.array(g["A"], dtype=np.float64).reshape(-1, order="F")
    rep = lambda v: np.ascontiguousarray(np.tile(np.asarray(v, dtype=np.float64), (copies, 1)))
    return dict(P=rep(P), q=rep(g["q"]), A=rep(A), l=rep(g["l"]), u=rep(g["u"]), n=2, m=3, batch=copies)


@pytest.mark.parametrize("kernel", KERNELS)
def test_simple_qp_reference_cases(api, ctx, oracle, golden, kernel):
    """tests/qp_solver_test.cpp:43-100 through the CUDA path, plus exact agreement with the oracle."""
    d = simple_qp_batch(golden, copies=3)
    e4 = float(np.float32(1e-4))
    cases = {
        "testSimpleQP": api.default_settings(max_iter=1000),
        "testConstraintViolation": api.default_settings(eps_rel=e4, eps_abs=e4),
        "testAdaptiveRho": api.default_settings(adaptive_rho=1, adaptive_rho_interval=10),
        "sqp_ctor": api.sqp_ctor_settings(),
    }
    for name, s in cases.items():
        out = run_fused(api, ctx, d, s, kernel)
        ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
        assert_parity(out, ref, what=name)
        for i in range(3):
            assert out["status"][i] == api.SOLVED
            assert out["iter"][i] < s.max_iter
            assert is_approx(out["x"][i], golden["simple_qp"]["solution"], 1e-2)
        if name == "testConstraintViolation":
            A = np.array(golden["simple_qp"]["A"], dtype=float)
            assert (A @ out["x"][0] - np.array(golden["simple_qp"]["l"])).min() >= -1e-3
            assert (A @ out["x"][0] - np.array(golden["simple_qp"]["u"])).max() <= 1e-3
        # residuals are differences of O(1) terms: compare on the scale of those terms
        np.testing.assert_allclose(out["res_prim"], ref["res_prim"], rtol=1e-4, atol=1e-8)
        np.testing.assert_allclose(out["res_dual"], ref["res_dual"], rtol=1e-4, atol=1e-8)
        np.testing.assert_allclose(out["rho_estimate"], ref["rho_estimate"], rtol=1e-4)


SHAPES = [(32, 64, 96), (64, 128, 48), (5, 7, 16), (17, 3, 8), (1, 1, 4), (40, 100, 8), (64, 20, 8), (3, 128, 8)]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("n,m,batch", SHAPES)
def test_synthetic_defaults_S1(api, ctx, oracle, kernel, n, m, batch):
    """Reference default settings (S1 of SURVEY.md section 8d), configs 2 and 3 shapes plus ragged ones."""
    from sqp_solver_b200.synth import make_batch

    d = make_batch(batch, n, m, seed0=1000)
    s = api.default_settings()
    out = run_fused(api, ctx, d, s, kernel)
    ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
    worst = assert_parity(out, ref, what="S1 n=%d m=%d %s" % (n, m, out["kernel"]))
    executed = np.minimum(ref["iter"], s.max_iter).sum()
    assert out["total_iters"] == executed
    print("S1", n, m, out["kernel"], "worst x rel err %.2e" % worst)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("n,m,batch", [(32, 64, 96), (64, 128, 48), (5, 7, 16)])
def test_synthetic_adaptive_S2(api, ctx, oracle, kernel, n, m, batch):
    """alpha = 1.6 + adaptive rho every 25 iterations (S2): exercises in-kernel refactorisation."""
    from sqp_solver_b200.synth import make_batch

    d = make_batch(batch, n, m, seed0=2000)
    s = api.default_settings(alpha=1.6, adaptive_rho=1)
    out = run_fused(api, ctx, d, s, kernel)
    ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
    assert_parity(out, ref, what="S2 n=%d m=%d" % (n, m))
    assert (ref["rho_updates"] > 1).any()
    np.testing.assert_allclose(out["rho_estimate"], ref["rho_estimate"], rtol=1e-4)


@pytest.mark.parametrize("kernel", KERNELS)
def test_sqp_ctor_settings_interval_mismatch(api, ctx, oracle, kernel):
    """check_termination=10 with adaptive_rho_interval=50 and max_iter=100 (src/sqp.cpp:16-23):
    the adaptive block runs on iterations where a check also ran, and MAX_ITER gives iter=101."""
    from sqp_solver_b200.synth import make_batch

    d = make_batch(32, 16, 24, seed0=3000)
    s = api.sqp_ctor_settings()
    out = run_fused(api, ctx, d, s, kernel)
    ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
    assert_parity(out, ref, what="sqp ctor settings")
    s2 = api.default_settings(adaptive_rho=1, adaptive_rho_interval=7, check_termination=5, max_iter=60)
    out = run_fused(api, ctx, d, s2, kernel)
    ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s2))
    assert_parity(out, ref, what="interval 7 / check 5")
    assert (ref["status"] == api.MAX_ITER_EXCEEDED).any() and (ref["iter"][ref["status"] == 1] == 61).all()


@pytest.mark.parametrize("kernel", KERNELS)
def test_object_api_setup_solve_solve_update(api, ctx, oracle, golden, kernel):
    """setup(); solve(); solve() (warm, fact 0.4); update_qp(); solve() -- against the oracle object API,
    mirroring tests/qp_solver_test.cpp:102-125 and the dead update_qp case of qp_solver_sparse_test.cpp."""
    from sqp_solver_b200.synth import make_batch

    B, n, m = (6, 12, 20) if kernel != "tile8" else (4, 40, 70)
    select_kernel(api, ctx, kernel, n, m)
    d = make_batch(B, n, m, seed0=4000)
    d2 = make_batch(B, n, m, seed0=5000)
    b = api.QPBatch(ctx, B, n, m)
    sols = [oracle.QPSolver() for _ in range(B)]
    qps = [oracle.QuadraticProblem(d["P"][i].reshape(n, n, order="F"), d["q"][i], d["A"][i].reshape(m, n, order="F"),
                                   d["l"][i], d["u"][i]) for i in range(B)]
    qps2 = [oracle.QuadraticProblem(d2["P"][i].reshape(n, n, order="F"), d2["q"][i], d2["A"][i].reshape(m, n, order="F"),
                                    d2["l"][i], d2["u"][i]) for i in range(B)]

    def ref_state():
        return dict(x=np.array([s.primal_solution() for s in sols]), y=np.array([s.dual_solution() for s in sols]),
                    status=np.array([s.info().status for s in sols]), iter=np.array([s.info().iter for s in sols]),
                    rho_updates=np.array([s.info().rho_updates for s in sols]))

    args = (d["P"], d["q"], d["A"], d["l"], d["u"])
    args2 = (d2["P"], d2["q"], d2["A"], d2["l"], d2["u"])
    # solve() before setup(): silent no-op, status stays UNINITIALIZED (qp.cpp:68-71)
    try:
        b.solve(*args)
    except api.SolverError as e:
        if kernel.startswith("tile") and "outside its range" in str(e):
            pytest.skip("shape not covered by the register-tiled kernel")
        raise
    info = b.info()
    assert (info["status"] == api.UNINITIALIZED).all() and (info["iter"] == 0).all()

    b.settings.max_iter = 40  # first solve stops early: MAX_ITER_EXCEEDED with iter 41
    b.setup(*args)
    assert (b.info()["status"] == api.UNSOLVED).all()
    for s, qp in zip(sols, qps):
        s.settings().max_iter = 40
        s.setup(qp)
    b.solve(*args)
    for s, qp in zip(sols, qps):
        s.solve(qp)
    assert_parity(b.get(), ref_state(), what="first solve")
    # second solve warm-starts from the first (reference fact 0.4) with adaptive rho switched on
    b.settings.max_iter = 1000
    b.settings.adaptive_rho = 1
    b.settings.adaptive_rho_interval = 10
    b.solve(*args)
    for s, qp in zip(sols, qps):
        s.settings().max_iter = 1000
        s.settings().adaptive_rho = 1
        s.settings().adaptive_rho_interval = 10
        s.solve(qp)
    assert_parity(b.get(), ref_state(), what="second (warm) solve")
    # third solve continues with the adapted rho and factor kept from the second
    b.solve(*args)
    for s, qp in zip(sols, qps):
        s.solve(qp)
    assert_parity(b.get(), ref_state(), what="third solve")
    # update_qp with a different problem: no reset of x,z,y; rho back to settings.rho; rho_updates accumulates
    b.update_qp(*args2)
    assert (b.info()["status"] == api.UNSOLVED).all()
    b.solve(*args2)
    for s, qp in zip(sols, qps2):
        s.update_qp(qp)
        s.solve(qp)
    assert_parity(b.get(), ref_state(), what="update_qp + solve")
    b.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_numerical_issues_and_nan_inputs(api, ctx, oracle, kernel):
    from sqp_solver_b200.synth import make_batch

    d = make_batch(8, 6, 9, seed0=6000)
    d["P"][2, 6 * 6 - 1] = np.nan  # NaN on the diagonal -> LDLT failure -> NUMERICAL_ISSUES (qp.cpp:39-43)
    d["P"][5, 1] = np.nan  # NaN in the lower triangle
    d["l"][3, 0] = -np.inf  # infinities pass through unchanged (tests/sqp_test_autodiff.cpp:97)
    d["u"][3, 0] = 0.0
    d["l"][4, 1] = -np.inf
    d["u"][4, 1] = np.inf
    s = api.default_settings()
    out = run_fused(api, ctx, d, s, kernel)
    ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
    assert ref["status"][2] == api.NUMERICAL_ISSUES and ref["status"][5] == api.NUMERICAL_ISSUES
    np.testing.assert_array_equal(out["status"], ref["status"])
    np.testing.assert_array_equal(out["iter"], ref["iter"])
    ok = ref["status"] != api.NUMERICAL_ISSUES
    sub = lambda o: {k: (v[ok] if isinstance(v, np.ndarray) and v.shape[:1] == (8,) else v) for k, v in o.items()}
    assert_parity(sub(out), sub(ref), what="finite instances next to NaN ones")
    assert (out["x"][~ok] == 0).all()  # untouched cold-start iterates


@pytest.mark.parametrize("kernel", KERNELS)
def test_max_iter_edge_cases(api, ctx, oracle, kernel):
    from sqp_solver_b200.synth import make_batch

    d = make_batch(4, 8, 10, seed0=7000)
    for kw in (dict(max_iter=0), dict(max_iter=1), dict(check_termination=0, max_iter=30), dict(max_iter=25)):
        s = api.default_settings(**kw)
        out = run_fused(api, ctx, d, s, kernel)
        ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
        assert_parity(out, ref, what=str(kw))


def test_device_pointers_match_host_pointers(api, ctx):
    import torch
    from sqp_solver_b200.synth import make_batch

    d = make_batch(40, 32, 64, seed0=8000)
    s = api.default_settings()
    host = run_fused(api, ctx, d, s, "auto")
    dev = {k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")}
    b = api.QPBatch(ctx, 40, 32, 64)
    b.setup_solve(dev["P"], dev["q"], dev["A"], dev["l"], dev["u"])
    x = torch.empty(40, 32, dtype=torch.float64, device="cuda")
    st = torch.empty(40, dtype=torch.int32, device="cuda")
    it = torch.empty(40, dtype=torch.int32, device="cuda")
    b.get_into(x=x, status=st, iter=it)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(x.cpu().numpy(), host["x"])
    np.testing.assert_array_equal(st.cpu().numpy(), host["status"])
    np.testing.assert_array_equal(it.cpu().numpy(), host["iter"])
    # count < batch only touches the leading instances
    b2 = api.QPBatch(ctx, 40, 32, 64)
    b2.setup_solve(dev["P"], dev["q"], dev["A"], dev["l"], dev["u"], count=10)
    o2 = b2.get()
    np.testing.assert_array_equal(o2["x"][:10], host["x"][:10])
    assert (o2["status"][10:] == api.UNINITIALIZED).all() and (o2["x"][10:] == 0).all()
    b.close()
    b2.close()


def test_fused_then_solve_is_rejected(api, ctx):
    from sqp_solver_b200.synth import make_batch

    d = make_batch(2, 4, 4, seed0=1)
    b = api.QPBatch(ctx, 2, 4, 4)
    b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
    with pytest.raises(api.SolverError, match="setup\\(\\) first"):
        b.solve(d["P"], d["q"], d["A"], d["l"], d["u"])
    b.close()


def test_full_size_properties_config3(api, ctx, oracle):
    """BASELINE config 3 at full size (batch 8192, n=64, m=128): size-independent properties on every
    instance plus oracle parity on EVERY instance (the CPU oracle needs ~10 s for the batch on the box's cores)."""
    import torch
    from sqp_solver_b200.synth import make_batch

    B, n, m = 8192, 64, 128
    d = make_batch(B, n, m, seed0=0)
    s = api.default_settings()
    out = run_fused(api, ctx, d, s, "auto")
    assert set(np.unique(out["status"])) <= {api.SOLVED, api.MAX_ITER_EXCEEDED}
    assert ((out["iter"] % 25 == 0) | (out["iter"] == 1001)).all()
    assert (out["iter"][out["status"] == api.MAX_ITER_EXCEEDED] == 1001).all()
    # recompute the termination test independently (torch fp64 on the GPU) from the returned x, z, y
    P = torch.from_numpy(d["P"]).cuda().view(B, n, n).transpose(1, 2)  # column-major -> [B, row, col]
    A = torch.from_numpy(d["A"]).cuda().view(B, n, m).transpose(1, 2)
    q = torch.from_numpy(d["q"]).cuda()
    x, y, z = (torch.from_numpy(out[k]).cuda() for k in ("x", "y", "z"))
    Ax = torch.bmm(A, x.unsqueeze(2)).squeeze(2)
    Px = torch.bmm(P, x.unsqueeze(2)).squeeze(2)
    Aty = torch.bmm(A.transpose(1, 2), y.unsqueeze(2)).squeeze(2)
    rp = (Ax - z).abs().amax(1)
    rd = (Px + q + Aty).abs().amax(1)
    ep = s.eps_abs + s.eps_rel * torch.maximum(Ax.abs().amax(1), z.abs().amax(1))
    ed = s.eps_abs + s.eps_rel * torch.maximum(torch.maximum(Px.abs().amax(1), Aty.abs().amax(1)), q.abs().amax(1))
    solved = torch.from_numpy(out["status"] == api.SOLVED).cuda()
    ok = (rp <= ep * (1 + 1e-9)) & (rd <= ed * (1 + 1e-9))
    assert bool((ok == solved).all()), "termination test disagrees on %d instances" % int((ok != solved).sum())
    np.testing.assert_allclose(out["res_prim"], rp.cpu().numpy(), rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(out["res_dual"], rd.cpu().numpy(), rtol=1e-6, atol=1e-10)
    # z is inside the box, y obeys the sign pattern of an active-set multiplier
    l, u = torch.from_numpy(d["l"]).cuda(), torch.from_numpy(d["u"]).cuda()
    assert bool(((z >= l) & (z <= u)).all())
    ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
    worst = assert_parity(out, ref, what="config 3, all 8192 instances, S1")
    print("config 3 S1: all %d instances agree with the oracle (status, iter, rho_updates); worst x rel err %.2e" % (B, worst))


@pytest.mark.parametrize("config,B,n,m", [("config2", 1024, 32, 64), ("config3", 8192, 64, 128)])
@pytest.mark.parametrize("settings_name", ["S1", "S2"])
def test_full_batch_oracle_parity(api, ctx, oracle, config, B, n, m, settings_name):
    """BASELINE configs 2 and 3 at their full batch sizes, reference defaults (S1) and the SQP constructor's regime
    (S2: alpha 1.6 + adaptive rho): EVERY instance against the CPU oracle -- status, iteration count, rho updates, x within 1e-6
    relative, y within 1e-5. Failures are listed one by one (helpers.assert_parity), never sampled or averaged."""
    from sqp_solver_b200.synth import make_batch

    if config == "config3" and settings_name == "S1":
        pytest.skip("covered by test_full_size_properties_config3")
    d = make_batch(B, n, m, seed0=0)
    s = api.default_settings(**({} if settings_name == "S1" else dict(alpha=1.6, adaptive_rho=1)))
    out = run_fused(api, ctx, d, s, "auto")
    ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
    worst = assert_parity(out, ref, what="%s, all %d instances, %s" % (config, B, settings_name))
    assert out["total_iters"] == int(np.minimum(ref["iter"], s.max_iter).sum())
    print("%s %s (%s): all %d instances agree with the oracle; worst x rel err %.2e" % (config, settings_name, out["kernel"], B, worst))


@pytest.mark.parametrize("kernel", ["auto", "generic"])
def test_kernels_beyond_tile_range(api, ctx, oracle, kernel):
    """Shapes outside the register-tiled kernel (n > 64 or m > 128) dispatch to the blocked kernel (n <= 256, m <= 1024)
    and beyond that to the generic one; both are also checked when forced."""
    from sqp_solver_b200.synth import make_batch

    for n, m, batch in ((80, 150, 6), (100, 40, 4), (20, 300, 4), (129, 257, 3), (256, 512, 3), (300, 100, 2)):
        d = make_batch(batch, n, m, seed0=9000)
        for s in (api.default_settings(max_iter=300), api.default_settings(max_iter=200, alpha=1.6, adaptive_rho=1)):
            out = run_fused(api, ctx, d, s, kernel)
            expect = "generic" if (kernel == "generic" or n > 256 or m > 1024) else "block"
            assert out["kernel"].startswith(expect), out["kernel"]
            ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
            assert_parity(out, ref, what="%s n=%d m=%d" % (out["kernel"], n, m))


def test_blocked_kernel_object_api_and_numerical_issues(api, ctx, oracle):
    """setup / solve / update_qp as separate launches and a NaN instance through the blocked kernel (n = 96, m = 160)."""
    from sqp_solver_b200.synth import make_batch

    B, n, m = 4, 96, 160
    d = make_batch(B, n, m, seed0=9500)
    d["P"][1, 5 * n + 5] = np.nan
    args = (d["P"], d["q"], d["A"], d["l"], d["u"])
    b = api.QPBatch(ctx, B, n, m)
    b.settings = api.default_settings(adaptive_rho=1, alpha=1.6)
    b.setup(*args)
    assert ctx.last_kernel.startswith("block")
    st = b.info()["status"]
    assert st[1] == api.NUMERICAL_ISSUES and (np.delete(st, 1) == api.UNSOLVED).all()
    b.solve(*args)
    got = b.get()
    ref = oracle.solve_batch(*args, oracle_settings_from(oracle, b.settings))
    np.testing.assert_array_equal(got["status"], ref["status"])
    ok = ref["status"] != api.NUMERICAL_ISSUES
    sub = lambda o: {k: v[ok] for k, v in o.items() if isinstance(v, np.ndarray) and v.shape[:1] == (B,)}
    assert_parity(sub(got), sub(ref), what="blocked kernel, separate setup/solve")
    b.close()


def test_api_argument_errors(api, ctx):
    """API-level failures come back as error codes with text (SURVEY.md 8b 'Errors'), never as a crash."""
    from sqp_solver_b200.synth import make_batch

    with pytest.raises(api.SolverError, match="batch >= 1"):
        api.QPBatch(ctx, 0, 4, 4)
    with pytest.raises(api.SolverError, match="too large"):
        api.QPBatch(ctx, 1, 20000, 20000)
    d = make_batch(3, 4, 5, seed0=2)
    b = api.QPBatch(ctx, 3, 4, 5)
    with pytest.raises(api.SolverError, match="count outside"):
        b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"], count=4)
    b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"], count=0)  # no-op
    assert (b.info()["status"] == api.UNINITIALIZED).all()
    with pytest.raises(api.SolverError, match="host arrays or all"):
        import torch

        b.setup_solve(torch.from_numpy(d["P"]).cuda(), d["q"], d["A"], d["l"], d["u"])
    with pytest.raises(api.SolverError):
        ctx.set_option(api.OPT_KERNEL, 7)
    b.close()


def test_two_batches_and_streams_are_independent(api, ctx, oracle):
    """Two batch objects driven on two CUDA streams do not disturb each other's state."""
    import torch
    from sqp_solver_b200.synth import make_batch

    d1, d2 = make_batch(24, 32, 64, seed0=11000), make_batch(16, 16, 24, seed0=12000)
    dev1 = {k: torch.from_numpy(d1[k]).cuda() for k in ("P", "q", "A", "l", "u")}
    dev2 = {k: torch.from_numpy(d2[k]).cuda() for k in ("P", "q", "A", "l", "u")}
    torch.cuda.synchronize()
    b1, b2 = api.QPBatch(ctx, 24, 32, 64), api.QPBatch(ctx, 16, 16, 24)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        b1.setup_solve(dev1["P"], dev1["q"], dev1["A"], dev1["l"], dev1["u"], stream=s1.cuda_stream)
        b2.setup_solve(dev2["P"], dev2["q"], dev2["A"], dev2["l"], dev2["u"], stream=s2.cuda_stream)
    torch.cuda.synchronize()
    s = api.default_settings()
    for b, d in ((b1, d1), (b2, d2)):
        ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
        got = b.get()
        assert (got.pop("rho_updates") == 3).all()  # cumulative over the three setups, like the reference (qp.cpp:313)
        assert_parity(got, ref, what="concurrent batches")
    b1.close()
    b2.close()


def test_linearity_property_full_size_config2(api, ctx):
    """Size-independent property at BASELINE config 2 (batch 1024, n=32, m=64): scaling the objective (P, q) and
    rho by c > 0 leaves the primal ADMM trajectory unchanged and scales the duals by c."""
    from sqp_solver_b200.synth import make_batch

    d = make_batch(1024, 32, 64, seed0=5)
    # loose rows keep rho = RHO_MIN whatever rho is, which breaks the scaling symmetry: make them wide finite boxes
    d["l"] = np.maximum(d["l"], -50.0)
    d["u"] = np.minimum(d["u"], 50.0)
    c = 4.0  # a power of two: the scaled run is the same floating-point computation up to exact scaling
    s = api.default_settings(eps_abs=0.0, eps_rel=1e-3, sigma=1e-6)
    out1 = run_fused(api, ctx, d, s, "auto")
    d2 = dict(d, P=d["P"] * c, q=d["q"] * c)
    s2 = api.default_settings(eps_abs=0.0, eps_rel=1e-3, sigma=1e-6 * c, rho=0.1 * c)
    out2 = run_fused(api, ctx, d2, s2, "auto")
    np.testing.assert_array_equal(out1["iter"], out2["iter"])
    np.testing.assert_array_equal(out1["status"], out2["status"])
    np.testing.assert_allclose(out2["x"], out1["x"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out2["y"], c * out1["y"], rtol=1e-9, atol=1e-11)
    assert (out1["status"] == api.SOLVED).sum() > 512


@pytest.mark.parametrize("kernel", ["generic", "tile"])
@pytest.mark.parametrize("n,m", [(64, 128), (12, 20)])
def test_keep_and_reuse_factor_is_bit_identical(api, ctx, oracle, kernel, n, m):
    """Re-solving the same P, A with new q, l, u (the second-order-correction QP, sqp.cpp:244-276) with
    KEEP_FACTOR / REUSE_FACTOR gives exactly the plain setup_solve results, whether or not an instance's
    constraint classes changed (those refactor), and with adaptive rho moving rho in the first solve."""
    from sqp_solver_b200.synth import make_batch

    B = 24
    d = make_batch(B, n, m, seed0=13000)
    s = api.default_settings(alpha=1.6, adaptive_rho=1)
    select_kernel(api, ctx, kernel, n, m)
    b = api.QPBatch(ctx, B, n, m)
    b.settings = s
    b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"], opts=api.KEEP_FACTOR)
    first = b.get()
    plain = run_fused(api, ctx, d, s, kernel)
    for k in ("x", "y", "iter", "status"):
        np.testing.assert_array_equal(first[k], plain[k])
    # second problem: same P, A; shifted bounds and new q; instances 0..5 also change a constraint class
    rng = np.random.default_rng(1)
    d2 = dict(d, q=d["q"] + 0.1 * rng.standard_normal(d["q"].shape), l=d["l"].copy(), u=d["u"].copy())
    shift = 0.05 * rng.standard_normal(d["l"].shape)
    finite = np.abs(d["l"]) < 1e19
    d2["l"][finite] += shift[finite]
    d2["u"][finite] += shift[finite]
    for i in range(6):
        row = int(np.argmax((d2["u"][i] - d2["l"][i] > 1e-3) & finite[i]))
        d2["u"][i, row] = d2["l"][i, row]  # inequality -> equality
    b.setup_solve(d2["P"], d2["q"], d2["A"], d2["l"], d2["u"], opts=api.REUSE_FACTOR)
    reused = b.get()
    plain2 = run_fused(api, ctx, d2, s, kernel)
    for k in ("x", "y", "z", "iter", "status", "res_prim", "res_dual"):
        np.testing.assert_array_equal(reused[k], plain2[k], err_msg=k)
    ref = oracle.solve_batch(d2["P"], d2["q"], d2["A"], d2["l"], d2["u"], oracle_settings_from(oracle, s))
    reused.pop("rho_updates")
    assert_parity(reused, ref, what="reuse-factor re-solve")
    b.close()


def test_set_iterates_warm_start_from_checkpoint(api, ctx, oracle):
    """Checkpointed iterates restored with sqpb200_qp_batch_set_iterates continue exactly like the oracle whose
    x, y (qp.hpp:160,163 expose them by reference) and z are overwritten with the same values."""
    from sqp_solver_b200.synth import make_batch

    B, n, m = 12, 24, 40
    d = make_batch(B, n, m, seed0=14000)
    args = (d["P"], d["q"], d["A"], d["l"], d["u"])
    b = api.QPBatch(ctx, B, n, m)
    b.settings.max_iter = 60
    b.setup(*args)
    b.solve(*args)
    ck = b.get()  # checkpoint after 60 iterations
    b2 = api.QPBatch(ctx, B, n, m)
    b2.setup(*args)
    b2.set_iterates(x=ck["x"], y=ck["y"], z=ck["z"])
    b2.settings.max_iter = 1000
    b2.solve(*args)
    got = b2.get()
    refs = []
    for i in range(B):
        qp = oracle.QuadraticProblem(d["P"][i].reshape(n, n, order="F"), d["q"][i], d["A"][i].reshape(m, n, order="F"), d["l"][i], d["u"][i])
        s = oracle.QPSolver()
        s.setup(qp)
        s.set_iterates(x=ck["x"][i], y=ck["y"][i], z=ck["z"][i])
        s.solve(qp)
        refs.append(s)
    ref = dict(x=np.array([s.primal_solution() for s in refs]), y=np.array([s.dual_solution() for s in refs]),
               status=np.array([s.info().status for s in refs]), iter=np.array([s.info().iter for s in refs]),
               rho_updates=np.array([s.info().rho_updates for s in refs]))
    assert_parity(got, ref, what="warm start from checkpoint")
    assert (got["iter"] < 1000).any()
    b.close()
    b2.close()


def _sparse_batch(batch, n, m, density, seed0, layout):
    """Synthetic QPs whose A shares one sparsity pattern (density of nonzeros); returns the dense dict and the
    compressed arrays in the requested layout ('csc' like Eigen::SparseMatrix, or 'csr')."""
    from sqp_solver_b200.synth import make_batch

    d = make_batch(batch, n, m, seed0=seed0)
    rng = np.random.default_rng(seed0 + 12345)
    mask = rng.uniform(size=(m, n)) < density
    mask[np.arange(m), rng.integers(0, n, m)] = True  # no empty rows
    A3 = d["A"].reshape(batch, n, m).transpose(0, 2, 1) * mask  # [B, m, n]
    # keep the problems feasible: rebuild the bounds around c = A x0 with the same widths
    x0 = rng.standard_normal((batch, n))
    c = np.einsum("bij,bj->bi", A3, x0)
    width_l, width_u = rng.uniform(0, 1, (batch, m)), rng.uniform(0, 1, (batch, m))
    l, u = c - width_l, c + width_u
    eq = rng.uniform(size=(batch, m)) < 0.1
    l[eq] = c[eq]
    u[eq] = c[eq]
    d["l"], d["u"] = np.ascontiguousarray(l), np.ascontiguousarray(u)
    d["A"] = np.ascontiguousarray(A3.transpose(0, 2, 1).reshape(batch, n * m))  # column-major per instance
    if layout == "csc":
        cols, rows = np.nonzero(mask.T)  # column-major order of the stored entries
        outer = np.concatenate([[0], np.cumsum(mask.sum(axis=0))]).astype(np.int32)
        inner = rows.astype(np.int32)
        vals = np.ascontiguousarray(A3[:, rows, cols])
    else:
        rows, cols = np.nonzero(mask)
        outer = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int32)
        inner = cols.astype(np.int32)
        vals = np.ascontiguousarray(A3[:, rows, cols])
    return d, vals, outer, np.ascontiguousarray(inner)


@pytest.mark.parametrize("layout", ["csc", "csr"])
@pytest.mark.parametrize("n,m,batch,density", [(40, 60, 12, 0.15), (256, 512, 3, 0.03)])
def test_sparse_A_entry_point(api, ctx, oracle, layout, n, m, batch, density):
    """SURVEY.md 8f row 3 / BASELINE config 5 shape: A in compressed column (Eigen) or compressed row storage with a
    shared pattern. Parity against the oracle on the densified problem, and bit-identity with the dense entry point."""
    d, vals, outer, inner = _sparse_batch(batch, n, m, density, 15000, layout)
    s = api.default_settings(alpha=1.6, adaptive_rho=1, max_iter=400)
    b = api.QPBatch(ctx, batch, n, m)
    b.settings = s
    b.setup_solve_sparse(d["P"], d["q"], vals, outer, inner, d["l"], d["u"],
                         layout=api.SPARSE_CSC if layout == "csc" else api.SPARSE_CSR)
    got = b.get()
    kernel = ctx.last_kernel
    dense = run_fused(api, ctx, d, s, "auto")
    if n <= 64:  # register-tile shapes densify on the device: the very same kernel and arithmetic
        assert "sparse" not in kernel
        for k in ("x", "y", "iter", "status"):
            np.testing.assert_array_equal(got[k], dense[k], err_msg=k)
    else:  # blocked kernel walks the compressed pattern: different summation order, same algorithm
        assert "/sparse" in kernel, kernel
        np.testing.assert_array_equal(got["status"], dense["status"])
        np.testing.assert_allclose(got["x"], dense["x"], rtol=1e-6, atol=1e-9)
    ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
    assert_parity(got, ref, what="sparse %s n=%d m=%d" % (layout, n, m))
    b.close()


def test_sparse_simple_qp_reference_fixture(api, ctx, golden):
    """The reference's (dead) tests/qp_solver_sparse_test.cpp:36-49 and :85-98: SimpleQP with sparse A, adaptive rho ->
    [0.3, 0.7]; then P = I, q = 0 -> [0.5, 0.5]. A = [[1,1],[1,0],[0,1]] in CSC has 4 stored entries."""
    g = golden["simple_qp"]
    A = np.array(g["A"], dtype=float)
    cols, rows = np.nonzero(A.T)
    outer = np.concatenate([[0], np.cumsum((A != 0).sum(axis=0))]).astype(np.int32)
    vals = np.ascontiguousarray(A[rows, cols][None, :])
    P = np.array(g["P"], dtype=float).reshape(1, 4)
    q, l, u = (np.array(g[k], dtype=float).reshape(1, -1) for k in ("q", "l", "u"))
    b = api.QPBatch(ctx, 1, 2, 3)
    b.settings = api.default_settings(max_iter=1000, adaptive_rho=1)
    b.setup_solve_sparse(P, q, vals, outer, rows.astype(np.int32), l, u)
    out = b.get()
    assert out["status"][0] == api.SOLVED and is_approx(out["x"][0], g["solution"], 1e-2)
    b.setup_solve_sparse(np.eye(2).reshape(1, 4), np.zeros((1, 2)), vals, outer, rows.astype(np.int32), l, u)
    out = b.get()
    assert out["status"][0] == api.SOLVED and is_approx(out["x"][0], [0.5, 0.5], 1e-2)
    b.close()


@pytest.mark.parametrize("kernel", ["auto", "generic"])
def test_randomised_shapes_and_settings(api, ctx, oracle, kernel):
    """Fuzz-style parity sweep: random (n, m) across every tile configuration and the padding paths, random solver
    settings (relaxation, adaptive rho with odd intervals, check cadence, tolerances, rho, sigma), including m = 0."""
    from sqp_solver_b200.synth import make_batch

    rng = np.random.default_rng(2024)
    shapes = [(1, 0), (3, 0), (64, 128), (63, 127), (33, 65), (17, 33), (9, 17), (8, 16), (64, 1), (1, 128), (2, 3)]
    shapes += [(int(rng.integers(1, 65)), int(rng.integers(0, 129))) for _ in range(14)]
    n_total = n_stable = 0
    for case, (n, m) in enumerate(shapes):
        batch = int(rng.integers(2, 7))
        d = make_batch(batch, n, m, seed0=16000 + 10 * case)
        kw = dict(alpha=float(rng.choice([1.0, 1.6, 1.8])), adaptive_rho=int(rng.integers(0, 2)),
                  adaptive_rho_interval=int(rng.choice([7, 25, 50])), check_termination=int(rng.choice([1, 10, 25])),
                  max_iter=int(rng.choice([60, 300])), rho=float(rng.choice([0.05, 0.1, 1.0])),
                  sigma=float(rng.choice([1e-6, 1e-4])), eps_abs=float(rng.choice([1e-3, 1e-5])),
                  eps_rel=float(rng.choice([1e-3, 1e-5])), adaptive_rho_tolerance=float(rng.choice([2.0, 5.0])))
        s = api.default_settings(**kw)
        out = run_fused(api, ctx, d, s, kernel)
        os_ = oracle_settings_from(oracle, s)
        ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], os_)
        # Adaptive rho takes rho * sqrt(res_prim / res_dual) of NORMALISED residuals: when one of them sits at rounding level
        # (e.g. no active constraint: A x - z is pure solve noise, ~1e-13 through the KKT substitution and ~1e-16 through
        # z~ = A x~) the new rho -- and everything after it -- is decided by that noise and only the SAME arithmetic reproduces it
        # (found by this sweep: n=2, m=3, rho 5566 -> 1.5e-4 in the oracle vs the 1e-6 clamp in the Schur-complement kernels).
        # The thread-per-QP kernel (n + m <= 16 under "auto") is that arithmetic and is compared on every instance; for the other
        # kernels the oracle records the smallest normalised residual it ever fed to rho_estimate and instances below 1e-9 are excluded.
        stable = ref["diag_min_norms"].min(axis=1) > 1e-9
        if out["kernel"].startswith("small<"):
            stable[:] = True
        n_total += batch
        n_stable += int(stable.sum())
        if not stable.any():
            continue
        sub = lambda o: {k: v[stable] for k, v in o.items() if isinstance(v, np.ndarray) and v.shape[:1] == (batch,)}
        # strict bar everywhere: converged or not, 1e-6 relative on x with no floor
        assert_parity(sub(out), sub(ref), what="fuzz case %d n=%d m=%d %s %s" % (case, n, m, out["kernel"], kw))
    assert n_stable >= 0.9 * n_total, "only %d of %d fuzz instances are numerically stable in the oracle" % (n_stable, n_total)


@pytest.mark.parametrize("kernel", KERNELS)
def test_against_committed_golden_outputs(api, ctx, kernel):
    """The CUDA path against the COMMITTED oracle outputs of tests/golden/oracle_synthetic.json (no oracle run involved)."""
    import json
    import os

    from sqp_solver_b200.synth import make_batch

    with open(os.path.join(os.path.dirname(__file__), "golden", "oracle_synthetic.json")) as f:
        gold = json.load(f)
    for c in gold["cases"]:
        d = make_batch(c["batch"], c["n"], c["m"], seed0=c["seed0"])
        out = run_fused(api, ctx, d, api.default_settings(**c["settings"]), kernel)
        ref = dict(status=np.array(c["status"]), iter=np.array(c["iter"]), rho_updates=np.array(c["rho_updates"]),
                   x=np.array(c["x"]), y=np.array(c["y"]))
        assert_parity(out, ref, what="golden " + c["name"])


@pytest.mark.parametrize("settings_name", ["S1", "S2"])
@pytest.mark.parametrize("n,m,batch,density,layout", [(100, 150, 6, 0.08, "csr"), (128, 257, 5, 0.05, "csc"),
                                                      (200, 701, 5, 0.02, "csr"), (256, 512, 5, 0.03, "csc"),
                                                      (256, 512, 4, 0.065, "csr"), (250, 300, 4, 0.09, "csc"),
                                                      (65, 3, 3, 0.3, "csr"), (129, 5, 3, 0.2, "csc")])  # fewer rows than CTAs
def test_cluster_kernel_parity(api, ctx, oracle, n, m, batch, density, layout, settings_name):
    """The thread-block-cluster kernel (sparse A, 64 < n <= 256: H^-1 distributed over the shared memory of 4 CTAs) against the
    oracle on the densified problem: ragged n (padded to 128 / 256), m not divisible by the cluster size, both layouts, the
    reference defaults and the SQP regime, and one NaN instance (NUMERICAL_ISSUES, iter untouched)."""
    d, vals, outer, inner = _sparse_batch(batch, n, m, density, 16000 + n, layout)
    d["P"][1, 0] = np.nan  # instance 1: NaN pivot -> Eigen::LDLT::info() != Success -> NUMERICAL_ISSUES (qp.cpp:39-43)
    s = api.default_settings(max_iter=300) if settings_name == "S1" else api.default_settings(alpha=1.6, adaptive_rho=1, max_iter=300)
    ctx.set_option(api.OPT_KERNEL, api.KERNEL_CLUSTER)
    b = api.QPBatch(ctx, batch, n, m)
    b.settings = s
    b.setup_solve_sparse(d["P"], d["q"], vals, outer, inner, d["l"], d["u"],
                         layout=api.SPARSE_CSC if layout == "csc" else api.SPARSE_CSR)
    got = b.get()
    assert ctx.last_kernel.startswith("cluster"), ctx.last_kernel
    if density > 0.06 and n > 128 and m >= 256:  # too dense for four CTAs' shared memory at n > 128: eight CTAs per QP
        assert ctx.last_kernel.startswith("cluster<8>"), ctx.last_kernel
    assert got["status"][1] == api.NUMERICAL_ISSUES
    ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
    keep = np.array([i for i in range(batch) if i != 1])
    sub = lambda o: {k: np.asarray(v)[keep] for k, v in o.items() if isinstance(v, np.ndarray) and v.shape[:1] == (batch,)}
    assert ref["status"][1] == api.NUMERICAL_ISSUES
    assert_parity(sub(got), sub(ref), what="cluster kernel %s n=%d m=%d %s" % (layout, n, m, settings_name))
    # second call on the same batch object: rho_updates is cumulative (qp.cpp:313), everything else identical
    b.setup_solve_sparse(d["P"], d["q"], vals, outer, inner, d["l"], d["u"],
                         layout=api.SPARSE_CSC if layout == "csc" else api.SPARSE_CSR)
    again = b.get()
    np.testing.assert_array_equal(again["x"][keep], got["x"][keep])
    np.testing.assert_array_equal(again["iter"][keep], got["iter"][keep])
    b.close()


def test_cluster_kernel_many_waves_matches_blocked_kernel(api, ctx):
    """More QPs than resident clusters (the atomic work queue hands every cluster several QPs) and agreement with the blocked
    kernel on the same sparse inputs (same algorithm, different factorisation order: status identical, x to 1e-6)."""
    n, m, batch = 160, 300, 150
    d, vals, outer, inner = _sparse_batch(batch, n, m, 0.04, 17000, "csr")
    s = api.default_settings(alpha=1.6, adaptive_rho=1, max_iter=200)
    outs = {}
    for name, opt in (("cluster", api.KERNEL_CLUSTER), ("block", api.KERNEL_BLOCK)):
        ctx.set_option(api.OPT_KERNEL, opt)
        b = api.QPBatch(ctx, batch, n, m)
        b.settings = s
        b.setup_solve_sparse(d["P"], d["q"], vals, outer, inner, d["l"], d["u"], layout=api.SPARSE_CSR)
        outs[name] = b.get()
        assert ctx.last_kernel.startswith(name), ctx.last_kernel
        b.close()
    np.testing.assert_array_equal(outs["cluster"]["status"], outs["block"]["status"])
    np.testing.assert_array_equal(outs["cluster"]["iter"], outs["block"]["iter"])
    np.testing.assert_allclose(outs["cluster"]["x"], outs["block"]["x"], rtol=1e-6, atol=1e-9)


def _oracle_f32_batch(oracle, d, kw):
    """The oracle's QPSolver<float> (the reference's second instantiation, qp.cpp:386) instance by instance."""
    n, m, B = d["n"], d["m"], d["batch"]
    out = dict(x=np.zeros((B, n)), status=np.zeros(B, dtype=np.int32), iter=np.zeros(B, dtype=np.int32))
    for i in range(B):
        qp = oracle.QuadraticProblem(d["P"][i].reshape(n, n, order="F"), d["q"][i], d["A"][i].reshape(m, n, order="F"), d["l"][i],
                                     d["u"][i], dtype=np.float32)
        s = oracle.QPSolver(dtype=np.float32)
        for k, v in kw.items():
            setattr(s.settings(), k, v)
        s.setup(qp)
        s.solve(qp)
        out["x"][i] = s.primal_solution()
        out["status"][i], out["iter"][i] = s.info().status, s.info().iter
    return out


@pytest.mark.parametrize("n,m,batch", [(64, 128, 10), (32, 64, 10), (10, 14, 8), (2, 3, 4)])
def test_fp32_instantiation_against_float_oracle(api, ctx, oracle, n, m, batch):
    """QPSolver<float>. The register-tiled kernel instantiated for fp32 (registers, shared memory, arithmetic;
    the interface arrays stay float64 and carry float values) against the oracle's float instantiation.
    Bar (stated here, looser than the fp64 path's 1e-6): reference defaults -> identical status and iteration count, x within 1e-4
    relative; alpha 1.6 + adaptive rho -> identical status, x within 1e-2 relative (the reference's own float assertion,
    tests/qp_solver_test.cpp:58-69): the explicit fp32 inverse of an H with cond ~1e5 slows ADMM by a few checks there."""
    from sqp_solver_b200.synth import make_batch

    d = make_batch(batch, n, m, seed0=31000 + n)
    for kw, it_exact, tol in (({}, True, 1e-4), (dict(alpha=1.6, adaptive_rho=1), False, 1e-2)):
        b = api.QPBatch(ctx, batch, n, m)
        b.settings = api.default_settings(**kw)
        b.set_precision(True)
        b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
        got = b.get()
        assert ",f32>" in ctx.last_kernel, ctx.last_kernel
        ref = _oracle_f32_batch(oracle, d, kw)
        np.testing.assert_array_equal(got["status"], ref["status"])
        if it_exact:
            np.testing.assert_array_equal(got["iter"], ref["iter"])
        rel = np.linalg.norm(got["x"] - ref["x"], axis=1) / np.linalg.norm(ref["x"], axis=1)
        assert rel.max() < tol, (kw, rel)
        # every returned value is a float widened to double
        for k in ("x", "y", "z"):
            np.testing.assert_array_equal(got[k], got[k].astype(np.float32).astype(np.float64))
        # separate setup + solve launches (H^-1, x, z, y round-trip through the float64 arrays losslessly) == fused launch
        b2 = api.QPBatch(ctx, batch, n, m)
        b2.settings = api.default_settings(**kw)
        b2.set_precision(True)
        b2.setup(d["P"], d["q"], d["A"], d["l"], d["u"])
        b2.solve(d["P"], d["q"], d["A"], d["l"], d["u"])
        sep = b2.get()
        np.testing.assert_array_equal(sep["iter"], got["iter"])
        np.testing.assert_array_equal(sep["x"], got["x"])
        b.close()
        b2.close()


def test_fp32_reference_float_test_and_fallback(api, ctx, oracle, golden):
    """tests/qp_solver_test.cpp:58-69 (SimpleQP in float: SOLVED, x ~ [0.3, 0.7] to 1e-2) through the fp32 kernel; shapes beyond
    the register-tiled kernel keep computing in fp64 with the flag set."""
    d = simple_qp_batch(golden, copies=2)
    b = api.QPBatch(ctx, 2, 2, 3)
    b.set_precision(True)
    b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
    out = b.get()
    assert ",f32>" in ctx.last_kernel
    assert (out["status"] == api.SOLVED).all() and (out["iter"] < 1000).all()
    assert is_approx(out["x"][0], golden["simple_qp"]["solution"], 1e-2)
    b.close()
    from sqp_solver_b200.synth import make_batch

    # beyond the register-tiled kernel's shapes QPSolver<float> runs the generic kernel's float instantiation (float at every size,
    # like `template class QPSolver<float>`, qp.cpp:386): identical status and iteration count as the oracle's float instantiation at
    # the reference defaults, x within 1e-4 relative
    d = make_batch(3, 80, 100, seed0=3)
    b = api.QPBatch(ctx, 3, 80, 100)
    b.set_precision(True)
    b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
    assert ctx.last_kernel == "generic<f32>", ctx.last_kernel
    got = b.get()
    ref = _oracle_f32_batch(oracle, d, {})
    np.testing.assert_array_equal(got["status"], ref["status"])
    np.testing.assert_array_equal(got["iter"], ref["iter"])
    rel = np.linalg.norm(got["x"] - ref["x"], axis=1) / np.linalg.norm(ref["x"], axis=1)
    assert rel.max() < 1e-4, rel
    np.testing.assert_array_equal(got["x"], got["x"].astype(np.float32).astype(np.float64))
    # object API in float: setup + solve as separate launches (the packed float factor in the slab) == the fused launch
    b2 = api.QPBatch(ctx, 3, 80, 100)
    b2.set_precision(True)
    b2.setup(d["P"], d["q"], d["A"], d["l"], d["u"])
    b2.solve(d["P"], d["q"], d["A"], d["l"], d["u"])
    sep = b2.get()
    np.testing.assert_array_equal(sep["iter"], got["iter"])
    np.testing.assert_array_equal(sep["x"], got["x"])
    b.close()
    b2.close()


def test_cluster_kernel_concurrent_batches_on_streams(api, ctx):
    """Two batch objects solved by the cluster kernel at the same time on two CUDA streams (device pointers, asynchronous calls):
    each owns its exchange scratch, so the overlapped launches give the same results as running them one after the other."""
    import torch

    n, m, batch = 130, 260, 40
    ds = [_sparse_batch(batch, n, m, 0.04, 18000 + 100 * k, "csr") for k in range(2)]
    s = api.default_settings(alpha=1.6, adaptive_rho=1, max_iter=200)
    ctx.set_option(api.OPT_KERNEL, api.KERNEL_CLUSTER)
    dev = [{k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in
            dict(P=d["P"], q=d["q"], l=d["l"], u=d["u"], vals=vals, outer=outer, inner=inner).items()} for d, vals, outer, inner in ds]
    serial = []
    for k in range(2):
        b = api.QPBatch(ctx, batch, n, m)
        b.settings = s
        b.setup_solve_sparse(dev[k]["P"], dev[k]["q"], dev[k]["vals"], dev[k]["outer"], dev[k]["inner"], dev[k]["l"], dev[k]["u"], layout=api.SPARSE_CSR)
        torch.cuda.synchronize()
        serial.append(b.get())
        b.close()
    streams = [torch.cuda.Stream() for _ in range(2)]
    bs = [api.QPBatch(ctx, batch, n, m) for _ in range(2)]
    for rep in range(3):
        for k in range(2):
            bs[k].settings = s
            bs[k].setup_solve_sparse(dev[k]["P"], dev[k]["q"], dev[k]["vals"], dev[k]["outer"], dev[k]["inner"], dev[k]["l"], dev[k]["u"],
                                     layout=api.SPARSE_CSR, stream=streams[k].cuda_stream)
    torch.cuda.synchronize()
    for k in range(2):
        got = bs[k].get()
        np.testing.assert_array_equal(got["iter"], serial[k]["iter"])
        np.testing.assert_array_equal(got["x"], serial[k]["x"])
        bs[k].close()


def test_cluster_and_fp32_against_committed_golden_outputs(api, ctx):
    """The cluster kernel (sparse A) and the fp32 instantiation against the COMMITTED oracle outputs of
    tests/golden/oracle_sparse_f32.json (no oracle run involved)."""
    import json
    import os

    from sqp_solver_b200.synth import make_batch, make_sparse_batch

    with open(os.path.join(os.path.dirname(__file__), "golden", "oracle_sparse_f32.json")) as f:
        gold = json.load(f)
    for c in gold["sparse"]:
        d = make_sparse_batch(c["batch"], c["n"], c["m"], density=c["density"], seed0=c["seed0"])
        b = api.QPBatch(ctx, c["batch"], c["n"], c["m"])
        b.settings = api.default_settings(**c["settings"])
        b.setup_solve_sparse(d["P"], d["q"], d["vals"], d["outer"], d["inner"], d["l"], d["u"], layout=api.SPARSE_CSR)
        assert ctx.last_kernel.startswith("cluster"), ctx.last_kernel
        out = b.get()
        ref = dict(status=np.array(c["status"]), iter=np.array(c["iter"]), x=np.array(c["x"]), y=np.array(c["y"]))
        assert_parity({k: out[k] for k in ("status", "iter", "x", "y")}, ref, what="golden " + c["name"])
        b.close()
    for c in gold["f32"]:
        d = make_batch(c["batch"], c["n"], c["m"], seed0=c["seed0"])
        b = api.QPBatch(ctx, c["batch"], c["n"], c["m"])
        b.settings = api.default_settings(**c["settings"])
        b.set_precision(True)
        b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
        out = b.get()
        assert out["status"].tolist() == c["status"] and out["iter"].tolist() == c["iter"], c["name"]
        rel = np.linalg.norm(out["x"] - np.array(c["x"]), axis=1) / np.linalg.norm(np.array(c["x"]), axis=1)
        assert rel.max() < 1e-4, (c["name"], rel)
        b.close()


def test_verbose_prints_the_reference_status_lines():
    """settings.verbose: the per-check status line of the reference (print_status, qp.cpp:375-382: a header at the first check, then
    `iter obj rp rd` with "%4d  %.2e  %.2e  %.2e") comes from the device; SimpleQP with default settings checks at 25, 50, ..., 125."""
    import os
    import subprocess
    import sys

    code = (
        "import numpy as np\n"
        "from sqp_solver_b200 import api\n"
        "ctx = api.Context(0)\n"
        "b = api.QPBatch(ctx, 1, 2, 3)\n"
        "b.settings = api.default_settings(verbose=1)\n"
        "P = np.array([[4., 1., 1., 2.]]); q = np.array([[1., 1.]]); A = np.array([[1., 1., 0., 1., 0., 1.]])\n"
        "b.setup_solve(P, q, A, np.array([[1., 0., 0.]]), np.array([[1., .7, .7]]))\n"
        "o = b.get(); print('KERNEL', ctx.last_kernel, int(o['iter'][0]), int(o['status'][0]))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "iter   obj       rp        rd"
    rows = [ln for ln in lines[1:] if not ln.startswith("KERNEL")]
    assert [int(ln.split()[0]) for ln in rows] == [25, 50, 75, 100, 125]
    assert all(len(ln.split()) == 4 and "e" in ln.split()[1] for ln in rows)
    assert abs(float(rows[-1].split()[1]) - 1.88) < 0.02  # objective of SimpleQP at [0.3, 0.7]: 0.5 x'Px + q'x = 1.88
    assert "KERNEL generic 125 0" in lines[-1]


@pytest.mark.parametrize("kernel,n", [("tile", 5), ("tile", 64), ("generic", 12), ("block", 80), ("auto", 130)])
def test_unconstrained_qp_m_zero(api, ctx, oracle, kernel, n):
    """m = 0 (no constraint rows: Eigen handles the 0 x n matrices of the reference transparently): every kernel returns the
    oracle's result -- SOLVED at the first check with x = -P^-1 q."""
    rng = np.random.default_rng(50 + n)
    B = 3
    M = rng.standard_normal((B, n, n)) / np.sqrt(n)
    P = np.einsum("bij,bkj->bik", M, M) + 0.1 * np.eye(n)
    d = dict(P=np.ascontiguousarray(P.transpose(0, 2, 1).reshape(B, n * n)), q=rng.standard_normal((B, n)), A=np.zeros((B, 0)),
             l=np.zeros((B, 0)), u=np.zeros((B, 0)), n=n, m=0, batch=B)
    out = run_fused(api, ctx, d, api.default_settings(), kernel)
    ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"])
    assert (ref["status"] == api.SOLVED).all() and (ref["iter"] == 25).all()
    np.testing.assert_array_equal(out["status"], ref["status"])
    np.testing.assert_array_equal(out["iter"], ref["iter"])
    xs = np.stack([-np.linalg.solve(P[i], d["q"][i]) for i in range(B)])
    assert np.abs(out["x"] - ref["x"]).max() <= 1e-9 * np.abs(ref["x"]).max()
    assert np.abs(out["x"] - xs).max() <= 1e-4 * np.abs(xs).max()


def test_full_size_properties_config5(api, ctx, oracle):
    """BASELINE config 5 at full size (batch 2048 sparse-A QPs, n=256, m=512, one CSR pattern, SQP settings) through the cluster
    kernel: size-independent properties on every instance (the termination test recomputed independently in torch fp64 from the
    returned x, z, y; z inside the box; iteration counts on the check grid) plus oracle parity on a seeded sample of 192 instances
    (the densified CPU oracle runs ~100 of these per second on 16 cores)."""
    import torch
    from sqp_solver_b200.synth import densify, make_sparse_batch

    B, n, m = 2048, 256, 512
    d = make_sparse_batch(B, n, m, density=0.03, seed0=0)
    s = api.default_settings(alpha=1.6, adaptive_rho=1)
    b = api.QPBatch(ctx, B, n, m)
    b.settings = s
    b.setup_solve_sparse(d["P"], d["q"], d["vals"], d["outer"], d["inner"], d["l"], d["u"], layout=api.SPARSE_CSR)
    out = b.get()
    assert ctx.last_kernel.startswith("cluster<4>"), ctx.last_kernel
    assert (out["status"] == api.SOLVED).all()
    assert (out["iter"] % 25 == 0).all() and out["iter"].max() <= 1000
    P = torch.from_numpy(d["P"]).cuda().view(B, n, n).transpose(1, 2)
    A = torch.zeros(B, m, n, dtype=torch.float64, device="cuda")
    A[:, torch.from_numpy(d["rows"]).cuda(), torch.from_numpy(d["cols"]).cuda()] = torch.from_numpy(d["vals"]).cuda()
    q = torch.from_numpy(d["q"]).cuda()
    x, y, z = (torch.from_numpy(out[k]).cuda() for k in ("x", "y", "z"))
    Ax = torch.bmm(A, x.unsqueeze(2)).squeeze(2)
    Px = torch.bmm(P, x.unsqueeze(2)).squeeze(2)
    Aty = torch.bmm(A.transpose(1, 2), y.unsqueeze(2)).squeeze(2)
    rp, rd = (Ax - z).abs().amax(1), (Px + q + Aty).abs().amax(1)
    ep = s.eps_abs + s.eps_rel * torch.maximum(Ax.abs().amax(1), z.abs().amax(1))
    ed = s.eps_abs + s.eps_rel * torch.maximum(torch.maximum(Px.abs().amax(1), Aty.abs().amax(1)), q.abs().amax(1))
    assert bool(((rp <= ep * (1 + 1e-9)) & (rd <= ed * (1 + 1e-9))).all())
    np.testing.assert_allclose(out["res_prim"], rp.cpu().numpy(), rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(out["res_dual"], rd.cpu().numpy(), rtol=1e-6, atol=1e-10)
    l, u = torch.from_numpy(d["l"]).cuda(), torch.from_numpy(d["u"]).cuda()
    assert bool(((z >= l) & (z <= u)).all())
    idx = np.sort(np.random.default_rng(321).choice(B, 192, replace=False))
    sub_d = dict(d, vals=d["vals"][idx], batch=len(idx))
    ref = oracle.solve_batch(d["P"][idx], d["q"][idx], densify(sub_d), d["l"][idx], d["u"][idx], oracle_settings_from(oracle, s))
    sub = {k: v[idx] for k, v in out.items() if isinstance(v, np.ndarray)}
    assert_parity({k: sub[k] for k in ("status", "iter", "x", "y")}, {k: ref[k] for k in ("status", "iter", "x", "y")}, what="config 5 sample")
    b.close()
