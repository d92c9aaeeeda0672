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
ctx.close()
print("sanitizer smoke done")
