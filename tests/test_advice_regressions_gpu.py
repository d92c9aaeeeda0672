"""Regression tests for the round-1 advisor findings: each one reproduces the failing sequence it described."""
import numpy as np
import pytest

from helpers import assert_parity, oracle_settings_from

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


def test_cluster_kernel_rows_beyond_1024(api, ctx, oracle):
    """qp_cluster.cu packed (row index | value position): the row field was 10 bits wide while the 8-CTA plan accepts m up to 2048,
    so rows >= 1024 ran into the value position. The field is 11 bits now (PACK_BITS); parity at m in (1024, 2048]."""
    from sqp_solver_b200.synth import densify, make_sparse_batch

    n, m, B = 200, 1500, 3
    d = make_sparse_batch(B, n, m, density=0.004, seed0=61000)
    s = api.default_settings(alpha=1.6, adaptive_rho=1)
    b = api.QPBatch(ctx, B, n, m)
    b.settings = s
    b.setup_solve_sparse(d["P"], d["q"], d["vals"], d["outer"], d["inner"], d["l"], d["u"], layout=api.SPARSE_CSR)
    got = b.get()
    assert ctx.last_kernel.startswith("cluster<8>"), ctx.last_kernel
    ref = oracle.solve_batch(d["P"], d["q"], densify(d), d["l"], d["u"], oracle_settings_from(oracle, s))
    assert_parity({k: got[k] for k in ("status", "iter", "x", "y")}, {k: ref[k] for k in ("status", "iter", "x", "y")}, what="cluster<8>, m = 1500")
    b.close()


def test_generic_kernel_scratch_is_per_batch_object(api, ctx, oracle):
    """Two batch objects on two streams that both take the GENERIC kernel: its n*n factorisation workspace used to be one buffer per
    context, shared by concurrent launches."""
    import torch

    from sqp_solver_b200.synth import make_batch

    ctx.set_option(api.OPT_KERNEL, api.KERNEL_GENERIC)
    d1, d2 = make_batch(700, 48, 60, seed0=62000), make_batch(650, 40, 90, seed0=63000)
    dev = [{k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")} for d in (d1, d2)]
    torch.cuda.synchronize()
    bs = [api.QPBatch(ctx, d["batch"], d["n"], d["m"]) for d in (d1, d2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    s = api.default_settings(alpha=1.6, adaptive_rho=1, max_iter=200)
    for _ in range(2):
        for b, dv, st in zip(bs, dev, streams):
            b.settings = s
            b.setup_solve(dv["P"], dv["q"], dv["A"], dv["l"], dv["u"], stream=st.cuda_stream)
            assert ctx.last_kernel == "generic"
    for b, d in zip(bs, (d1, d2)):
        got = b.get()  # ordered behind the object's last launch, no explicit synchronisation
        got.pop("rho_updates")
        ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
        ref.pop("rho_updates")
        assert_parity(got, ref, what="concurrent generic launches")
        b.close()


@pytest.mark.parametrize("kernel", ["tile", "generic", "block"])
def test_plain_factor_launch_forgets_the_kept_factor(api, ctx, oracle, kernel):
    """KEEP (classes C1) -> plain setup_solve with new bounds (classes C2) -> REUSE with C2: the kept factor belongs to C1's rho
    vector and must not be reused (the plain launch overwrote the stored classes but used to leave the slab's tag alone)."""
    from sqp_solver_b200.synth import make_batch

    n, m = (20, 30) if kernel != "block" else (80, 100)
    B = 12
    ctx.set_option(api.OPT_KERNEL, {"tile": api.KERNEL_TILE, "generic": api.KERNEL_GENERIC, "block": api.KERNEL_BLOCK}[kernel])
    d = make_batch(B, n, m, seed0=64000)
    s = api.default_settings()
    d2 = dict(d, l=d["l"].copy(), u=d["u"].copy())
    for i in range(B):  # turn two inequality rows per instance into equalities: different classes, different rho vector
        rows = np.nonzero((d["u"][i] - d["l"][i] > 1e-3) & (np.abs(d["l"][i]) < 1e19))[0][:2]
        d2["u"][i, rows] = d2["l"][i, rows]
    b = api.QPBatch(ctx, B, n, m)
    b.settings = s
    b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"], opts=api.KEEP_FACTOR)
    b.setup_solve(d2["P"], d2["q"], d2["A"], d2["l"], d2["u"])
    b.setup_solve(d2["P"], d2["q"], d2["A"], d2["l"], d2["u"], opts=api.REUSE_FACTOR)
    got = b.get()
    got.pop("rho_updates")
    ref = oracle.solve_batch(d2["P"], d2["q"], d2["A"], d2["l"], d2["u"], oracle_settings_from(oracle, s))
    ref.pop("rho_updates")
    assert_parity(got, ref, what="REUSE after a plain launch changed the classes (%s)" % ctx.last_kernel)
    b.close()


def test_solve_keeps_the_kernel_that_stored_the_factor(api, ctx, oracle):
    """setup() under one kernel option, solve() under another (64 < n <= 256: blocked LDL^T panels vs the generic kernel's dense
    inverse): solve() must run the kernel whose layout is in the slab."""
    from sqp_solver_b200.synth import make_batch

    B, n, m = 6, 80, 100
    d = make_batch(B, n, m, seed0=65000)
    s = api.default_settings()
    ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
    args = (d["P"], d["q"], d["A"], d["l"], d["u"])
    for first, second in ((api.KERNEL_AUTO, api.KERNEL_GENERIC), (api.KERNEL_GENERIC, api.KERNEL_AUTO)):
        b = api.QPBatch(ctx, B, n, m)
        b.settings = s
        ctx.set_option(api.OPT_KERNEL, first)
        b.setup(*args)
        stored_by = ctx.last_kernel
        ctx.set_option(api.OPT_KERNEL, second)
        b.solve(*args)
        assert ctx.last_kernel.split("<")[0] == stored_by.split("<")[0], (stored_by, ctx.last_kernel)
        assert_parity(b.get(), ref, what="setup under %s, solve under option %d" % (stored_by, second))
        b.close()


def test_fp32_reuse_factor_hits(api, ctx):
    """fp32: the slab's tag is compared in the compute scalar, so REUSE finds the kept factor (it never matched before:
    (double)(float)0.1 != 0.1) -- results equal the plain path either way; this checks them after a KEEP / REUSE pair."""
    from sqp_solver_b200.synth import make_batch

    B, n, m = 16, 24, 40
    d = make_batch(B, n, m, seed0=66000)
    b = api.QPBatch(ctx, B, n, m)
    b.set_precision(True)
    b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"], opts=api.KEEP_FACTOR)
    d2 = dict(d, q=d["q"] * 1.01)
    b.setup_solve(d2["P"], d2["q"], d2["A"], d2["l"], d2["u"], opts=api.REUSE_FACTOR)
    reused = b.get()
    p = api.QPBatch(ctx, B, n, m)
    p.set_precision(True)
    p.setup_solve(d2["P"], d2["q"], d2["A"], d2["l"], d2["u"])
    plain = p.get()
    for k in ("x", "y", "iter", "status"):
        np.testing.assert_array_equal(reused[k], plain[k], err_msg=k)
    b.close()
    p.close()


def test_empty_and_maximum_sizes(api, ctx, oracle):
    """Edge cases of the boundary: count = 0 is a no-op; the largest shapes of the blocked kernel (256, 1024) and a shape only the generic
    kernel takes (300, 700); a batch of one; counts smaller than the capacity leave the other instances untouched."""
    from sqp_solver_b200.synth import make_batch

    d = make_batch(4, 8, 10, seed0=67000)
    b = api.QPBatch(ctx, 4, 8, 10)
    b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"], count=0)  # nothing to do, nothing launched
    assert (b.info()["status"] == api.UNINITIALIZED).all()
    b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"], count=2)
    info = b.info()
    assert (info["status"][:2] != api.UNINITIALIZED).all() and (info["status"][2:] == api.UNINITIALIZED).all()
    b.close()
    s = api.default_settings(alpha=1.6, adaptive_rho=1, max_iter=200)
    for (n, m, batch, expect) in ((256, 1024, 2, "block"), (300, 700, 2, "generic"), (64, 128, 1, "tile"), (1, 0, 1, "small")):
        d = make_batch(batch, n, m, seed0=68000 + n)
        b = api.QPBatch(ctx, batch, n, m)
        b.settings = s
        b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
        assert ctx.last_kernel.startswith(expect), ctx.last_kernel
        got = b.get()
        ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle_settings_from(oracle, s))
        assert_parity(got, ref, what="n=%d m=%d (%s)" % (n, m, ctx.last_kernel))
        b.close()


@pytest.mark.parametrize("n,m,expect_direct", [(64, 128, True), (20, 30, True), (2, 3, False), (80, 100, False)])
def test_setup_solve_to_writes_caller_arrays(api, ctx, oracle, n, m, expect_direct):
    """sqpb200_qp_batch_setup_solve_to: fused setup + solve of FRESH instances whose results land in caller device arrays -- written by
    the register-tiled kernel's epilogue, or copied out behind the launch for the other kernels. Repeated calls return per-call info
    (rho_updates = 1 under the reference defaults), unlike the object API whose counter accumulates (qp.cpp:313)."""
    import torch

    from sqp_solver_b200.synth import make_batch

    B = 40
    d = make_batch(B, n, m, seed0=69000 + n)
    dev = {k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")}
    out = dict(x=torch.full((B, n), 7.0, dtype=torch.float64, device="cuda"), y=torch.full((B, m), 7.0, dtype=torch.float64, device="cuda"),
               z=torch.full((B, m), 7.0, dtype=torch.float64, device="cuda"), status=torch.full((B,), 77, dtype=torch.int32, device="cuda"),
               iter=torch.full((B,), 77, dtype=torch.int32, device="cuda"), rho_updates=torch.full((B,), 77, dtype=torch.int32, device="cuda"),
               rho_estimate=torch.full((B,), 7.0, dtype=torch.float64, device="cuda"), res_prim=torch.full((B,), 7.0, dtype=torch.float64, device="cuda"),
               res_dual=torch.full((B,), 7.0, dtype=torch.float64, device="cuda"))
    b = api.QPBatch(ctx, B, n, m)
    ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"])
    for call in range(2):
        b.setup_solve_to(dev["P"], dev["q"], dev["A"], dev["l"], dev["u"], out)
        torch.cuda.synchronize()
        got = {k: v.cpu().numpy() for k, v in out.items()}
        assert_parity({k: got[k] for k in ("status", "iter", "rho_updates", "x", "y")}, ref, what="setup_solve_to call %d (%s)" % (call, ctx.last_kernel))
        np.testing.assert_allclose(got["res_prim"], ref["res_prim"], rtol=1e-4, atol=1e-8)
    assert ctx.last_kernel.startswith("tile") == expect_direct, ctx.last_kernel
    b.close()
