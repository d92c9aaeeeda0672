"""Small driver for compute-sanitizer runs (memcheck / racecheck / synccheck / initcheck) over every kernel
configuration, with adaptive rho on so the in-kernel refactorisation path runs too:
    compute-sanitizer --tool racecheck python tests/sanitizer_smoke.py
Not collected by pytest (no test_ prefix)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from sqp_solver_b200 import api
from sqp_solver_b200.synth import make_batch

ctx = api.Context(0)
for kernel, warps, n, m, batch in () if os.environ.get("SANITIZER_ONLY") else ((api.KERNEL_TILE, 0, 64, 128, 6), (api.KERNEL_TILE, 8, 64, 128, 4), (api.KERNEL_TILE, 0, 32, 64, 6),
                                   (api.KERNEL_TILE, 2, 32, 64, 4), (api.KERNEL_TILE, 0, 16, 32, 4), (api.KERNEL_TILE, 0, 5, 7, 4),
                                   (api.KERNEL_TILE, 0, 50, 100, 4), (api.KERNEL_GENERIC, 0, 20, 30, 4)):
    ctx.set_option(api.OPT_KERNEL, kernel)
    ctx.set_option(api.OPT_TILE_WARPS, warps)
    d = make_batch(batch, n, m, seed0=99)
    b = api.QPBatch(ctx, batch, n, m)
    b.settings = api.default_settings(alpha=1.6, adaptive_rho=1, max_iter=120)
    b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
    out = b.get()
    b.setup(d["P"], d["q"], d["A"], d["l"], d["u"])
    b.solve(d["P"], d["q"], d["A"], d["l"], d["u"])
    out2 = b.get()
    assert np.array_equal(out["iter"], out2["iter"]) and np.allclose(out["x"], out2["x"], rtol=0, atol=0), ctx.last_kernel
    print(ctx.last_kernel, n, m, out["iter"].tolist())
    b.close()
# fp32 instantiation of the register-tiled kernel (QPSolver<float>)
if not os.environ.get("SANITIZER_ONLY") or os.environ.get("SANITIZER_ONLY") == "fp32":
    ctx.set_option(api.OPT_KERNEL, api.KERNEL_TILE)
    ctx.set_option(api.OPT_TILE_WARPS, 0)
    for n, m, batch in ((64, 128, 4), (32, 64, 4), (50, 100, 3), (5, 7, 3)):
        d = make_batch(batch, n, m, seed0=99)
        b = api.QPBatch(ctx, batch, n, m)
        b.set_precision(True)
        b.settings = api.default_settings(alpha=1.6, adaptive_rho=1, max_iter=120)
        b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
        out = b.get()
        b.setup(d["P"], d["q"], d["A"], d["l"], d["u"])
        b.solve(d["P"], d["q"], d["A"], d["l"], d["u"])
        out2 = b.get()
        assert np.array_equal(out["iter"], out2["iter"]) and np.array_equal(out["x"], out2["x"]), ctx.last_kernel
        print(ctx.last_kernel, n, m, out["iter"].tolist())
        b.close()

# blocked kernel (dense and sparse A) and the thread-block-cluster kernel (sparse A): fused launches with adaptive rho
from sqp_solver_b200.synth import make_sparse_batch

only = os.environ.get("SANITIZER_ONLY", "")
for kernel, n, m, batch, dens in ((api.KERNEL_BLOCK, 96, 160, 3, 0.0), (api.KERNEL_BLOCK, 100, 150, 3, 0.08),
                                  (api.KERNEL_CLUSTER, 100, 150, 3, 0.08), (api.KERNEL_CLUSTER, 200, 301, 2, 0.03)):
    if only and only != {api.KERNEL_BLOCK: "block", api.KERNEL_CLUSTER: "cluster"}[kernel]:  # (SANITIZER_ONLY=fp32 skips them all)
        continue
    ctx.set_option(api.OPT_KERNEL, kernel)
    st = api.default_settings(alpha=1.6, adaptive_rho=1, max_iter=60)
    b = api.QPBatch(ctx, batch, n, m)
    b.settings = st
    if dens == 0.0:
        d = make_batch(batch, n, m, seed0=99)
        b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
    else:
        d = make_sparse_batch(batch, n, m, density=dens, seed0=99)
        b.setup_solve_sparse(d["P"], d["q"], d["vals"], d["outer"], d["inner"], d["l"], d["u"], layout=api.SPARSE_CSR)
    out = b.get()
    print(ctx.last_kernel, n, m, out["iter"].tolist(), out["status"].tolist())
    b.close()

# round 2: thread-per-QP kernels (register and shared-memory variants, double and float), time-sliced register-tiled kernel (suspend /
# re-queue / resume, also with results to caller arrays and local copies of A, P), the cluster kernel's object-API launch modes, the
# generic kernel's float instantiation, setup_solve_to
if not only or only == "r2":
    import torch

    ctx.set_option(api.OPT_KERNEL, api.KERNEL_AUTO)
    ctx.set_option(api.OPT_TILE_WARPS, 0)
    for n, m, batch, f32 in ((2, 2, 70, False), (2, 3, 40, True), (4, 4, 33, False), (5, 7, 40, False), (1, 15, 20, True), (3, 0, 9, False)):
        d = make_batch(batch, n, m, seed0=98)
        b = api.QPBatch(ctx, batch, n, m)
        b.set_precision(f32)
        b.settings = api.default_settings(alpha=1.6, adaptive_rho=1, adaptive_rho_interval=10, check_termination=5, max_iter=80)
        b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
        out = b.get()
        b.setup(d["P"], d["q"], d["A"], d["l"], d["u"])
        b.solve(d["P"], d["q"], d["A"], d["l"], d["u"])
        out2 = b.get()
        assert np.array_equal(out["iter"], out2["iter"]) and np.array_equal(out["x"], out2["x"]), ctx.last_kernel
        print(ctx.last_kernel, n, m, out["iter"][:6].tolist())
        b.close()
    for sl, n, m, batch in ((13, 64, 128, 20), (40, 50, 100, 12)):
        d = make_batch(batch, n, m, seed0=97)
        dev = {k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")}
        outs = []
        for s_ in (0, sl):
            ctx.set_option(api.OPT_SLICE_ITERS, s_)
            b = api.QPBatch(ctx, batch, n, m)
            b.settings = api.default_settings(alpha=1.6, adaptive_rho=1, max_iter=150)
            b.setup_solve(dev["P"], dev["q"], dev["A"], dev["l"], dev["u"])
            outs.append(b.get())
            print(ctx.last_kernel, n, m, outs[-1]["iter"][:6].tolist())
            if s_:
                tgt = dict(x=torch.zeros(batch, n, dtype=torch.float64, device="cuda"), y=torch.zeros(batch, m, dtype=torch.float64, device="cuda"),
                           z=torch.zeros(batch, m, dtype=torch.float64, device="cuda"), status=torch.zeros(batch, dtype=torch.int32, device="cuda"),
                           iter=torch.zeros(batch, dtype=torch.int32, device="cuda"), rho_updates=torch.zeros(batch, dtype=torch.int32, device="cuda"),
                           rho_estimate=torch.zeros(batch, dtype=torch.float64, device="cuda"), res_prim=torch.zeros(batch, dtype=torch.float64, device="cuda"),
                           res_dual=torch.zeros(batch, dtype=torch.float64, device="cuda"))
                b.setup_solve_to(dev["P"], dev["q"], dev["A"], dev["l"], dev["u"], tgt)
                torch.cuda.synchronize()
                assert np.array_equal(tgt["x"].cpu().numpy(), outs[0]["x"]) and np.array_equal(tgt["iter"].cpu().numpy(), outs[0]["iter"])
            b.close()
        ctx.set_option(api.OPT_SLICE_ITERS, -1)
        assert np.array_equal(outs[0]["x"], outs[1]["x"]) and np.array_equal(outs[0]["iter"], outs[1]["iter"])
    d = make_sparse_batch(3, 100, 150, density=0.08, seed0=96)
    b = api.QPBatch(ctx, 3, 100, 150)
    b.settings = api.default_settings(alpha=1.6, adaptive_rho=1, max_iter=40)
    sp = (d["P"], d["q"], d["vals"], d["outer"], d["inner"], d["l"], d["u"])
    b.setup_sparse(*sp, layout=api.SPARSE_CSR)
    b.solve_sparse(*sp, layout=api.SPARSE_CSR)
    b.solve_sparse(*sp, layout=api.SPARSE_CSR)
    b.update_qp_sparse(*sp, layout=api.SPARSE_CSR)
    b.solve_sparse(*sp, layout=api.SPARSE_CSR)
    print(ctx.last_kernel, "object API", b.get()["iter"].tolist())
    b.close()
    d = make_batch(3, 70, 90, seed0=95)
    b = api.QPBatch(ctx, 3, 70, 90)
    b.set_precision(True)
    b.settings = api.default_settings(max_iter=60)
    b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"])
    print(ctx.last_kernel, b.get()["iter"].tolist())
    b.close()
ctx.close()
print("sanitizer smoke done")
