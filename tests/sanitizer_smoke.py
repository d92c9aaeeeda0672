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
for kernel, warps, n, m, batch in ((api.KERNEL_TILE, 0, 64, 128, 6), (api.KERNEL_TILE, 8, 64, 128, 4), (api.KERNEL_TILE, 0, 32, 64, 6),
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
ctx.close()
print("sanitizer smoke done")
