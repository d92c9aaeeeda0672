"""Time slicing of the register-tiled kernel (SQPB200_OPT_SLICE_ITERS): a QP is suspended after a slice of iterations -- iterates, info and
H^-1 parked in the object's arrays -- and re-queued; a resumed QP carries its iteration counter on. The arithmetic sequence is the one of an
unsliced solve, so every returned value must be BIT-identical, whatever the slice length (also lengths that are no multiple of the check or
adaptive-rho interval), with adaptive-rho refactorisations inside slices, with NaN instances, and when the results go to caller arrays."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
FIELDS = ("x", "y", "z", "status", "iter", "rho_updates", "rho_estimate", "res_prim", "res_dual")


@pytest.fixture(scope="module")
def api():
    from sqp_solver_b200 import api

    return api


@pytest.fixture(scope="module")
def ctx(api):
    c = api.Context(0)
    yield c
    c.set_option(api.OPT_SLICE_ITERS, -1)
    c.close()


def solve(api, ctx, d, settings, slice_iters, device=False):
    import torch

    ctx.set_option(api.OPT_SLICE_ITERS, slice_iters)
    b = api.QPBatch(ctx, d["batch"], d["n"], d["m"])
    b.settings = settings
    if device:
        dev = {k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")}
        b.setup_solve(dev["P"], dev["q"], dev["A"], dev["l"], dev["u"])
    else:
        b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
    out = b.get()
    out["kernel"] = ctx.last_kernel
    out["total_iters"] = b.total_iters()
    b.close()
    return out


@pytest.mark.parametrize("n,m,batch", [(64, 128, 700), (40, 100, 300), (64, 65, 150)])
@pytest.mark.parametrize("settings_name", ["S1", "S2", "odd"])
def test_sliced_solve_is_bit_identical(api, ctx, n, m, batch, settings_name):
    from sqp_solver_b200.synth import make_batch

    d = make_batch(batch, n, m, seed0=81000 + n + m)
    d["P"][3, 0] = np.nan  # a NUMERICAL_ISSUES instance among the others
    s = {"S1": api.default_settings(), "S2": api.default_settings(alpha=1.6, adaptive_rho=1),
         "odd": api.default_settings(alpha=1.8, adaptive_rho=1, adaptive_rho_interval=7, check_termination=3, max_iter=300,
                                     adaptive_rho_tolerance=2.0)}[settings_name]
    ref = solve(api, ctx, d, s, 0, device=True)
    assert "/sliced" not in ref["kernel"]
    for sl in (250, 100, 37, 1, 150 * 65536 + 40):  # the last: a first slice of 150 iterations, then slices of 40
        if sl == 1 and batch > 200:
            continue  # one iteration per slice: a stress test of the queue, kept to the small batch
        got = solve(api, ctx, d, s, sl, device=True)
        if (sl & 0xffff) < s.max_iter:
            assert got["kernel"].endswith("/sliced"), got["kernel"]
        for k in FIELDS:
            a, b_ = got[k], ref[k]
            same = (a == b_) | ((a != a) & (b_ != b_)) if a.dtype.kind == "f" else (a == b_)
            assert same.all(), "%s slice %d: %s differs on instances %s" % (settings_name, sl, k, np.unique(np.nonzero(~same)[0])[:8])
        assert got["total_iters"] == ref["total_iters"]


def test_automatic_slicing_and_host_pointers(api, ctx):
    """-1 = automatic: on for a batch of a few QPs per CTA slot (device pointers, or host pointers staged in one piece), off for a large
    batch, a tiny one, or a host-pointer call whose chunked staging runs behind the kernel."""
    from sqp_solver_b200.synth import make_batch

    s = api.default_settings()
    d = make_batch(600, 64, 128, seed0=82000)
    ref = solve(api, ctx, d, s, 0)
    auto = solve(api, ctx, d, s, -1, device=True)
    assert auto["kernel"].endswith("/sliced"), auto["kernel"]
    host = solve(api, ctx, d, s, 250)  # host pointers, chunked staging: unsliced
    assert "/sliced" not in host["kernel"]
    for k in FIELDS:
        np.testing.assert_array_equal(auto[k], ref[k], err_msg=k)
        np.testing.assert_array_equal(host[k], ref[k], err_msg=k)
    small = solve(api, ctx, make_batch(40, 64, 128, seed0=82100), s, -1, device=True)
    assert "/sliced" not in small["kernel"]


def test_sliced_solve_into_caller_arrays(api, ctx):
    """setup_solve_to (the multi-GPU path: inputs possibly in peer memory -> local copies of P and A for the resumes, results written to
    caller arrays by the final slice)."""
    import torch

    from sqp_solver_b200.synth import make_batch

    B, n, m = 500, 64, 128
    d = make_batch(B, n, m, seed0=83000)
    s = api.default_settings(alpha=1.6, adaptive_rho=1)
    ref = solve(api, ctx, d, s, 0, device=True)
    dev = {k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")}
    shapes = dict(x=(B, n), y=(B, m), z=(B, m))
    out = {k: torch.full(shapes.get(k, (B,)), 7, dtype=torch.int32 if k in ("status", "iter", "rho_updates") else torch.float64, device="cuda") for k in FIELDS}
    ctx.set_option(api.OPT_SLICE_ITERS, 60)
    b = api.QPBatch(ctx, B, n, m)
    b.settings = s
    for _ in range(2):
        b.setup_solve_to(dev["P"], dev["q"], dev["A"], dev["l"], dev["u"], out)
        torch.cuda.synchronize()
        assert ctx.last_kernel.endswith("/sliced")
        for k in FIELDS:
            np.testing.assert_array_equal(out[k].cpu().numpy(), ref[k], err_msg=k)
    b.close()
