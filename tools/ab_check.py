"""A/B helper: the same batch through two OPT_TILE_WARPS variants; prints times and whether the results are bit-identical."""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from sqp_solver_b200 import api
from sqp_solver_b200.synth import make_batch
va, vb = int(sys.argv[1]), int(sys.argv[2])
ctx = api.Context(0)
B, n, m = 8192, 64, 128
d = make_batch(B, n, m, seed0=0)
dev = [torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")]
for sname, st in (("S1", api.default_settings()), ("S2", api.default_settings(alpha=1.6, adaptive_rho=1))):
    outs = []
    for v in (va, vb):
        ctx.set_option(api.OPT_TILE_WARPS, v)
        b = api.QPBatch(ctx, B, n, m); b.settings = st
        ts = []
        for rep in range(6):
            torch.cuda.synchronize(); t0 = time.perf_counter(); b.setup_solve(*dev); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        o = b.get(); outs.append(o)
        print(sname, "variant", v, ctx.last_kernel, "ms %.3f" % (1e3 * min(ts)), "iters", int(o["iter"].sum()), flush=True)
        b.close()
    same = all(np.array_equal(outs[0][k], outs[1][k], equal_nan=True) for k in ("x", "y", "z", "iter", "status", "rho_updates"))
    print(sname, "bit-identical:", same)
