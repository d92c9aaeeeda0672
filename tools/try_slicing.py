"""Timing of the time-sliced register-tiled kernel on one GPU at the per-GPU batch sizes of the strong-scaling flow (1024 = 8192 / 8, 2048 = 8192 / 4).
Usage: python tools/try_slicing.py [slice_iters ...]"""
import sys

import torch

sys.path.insert(0, ".")
from sqp_solver_b200 import api  # noqa: E402
from sqp_solver_b200.synth import make_batch  # noqa: E402

ctx = api.Context(0)
data = {B: make_batch(B, 64, 128, seed0=0) for B in (1024, 2048, 4096)}
for sl in [int(a) for a in sys.argv[1:]] or [0, 250, 125]:
    ctx.set_option(api.OPT_SLICE_ITERS, sl)
    for B, d in data.items():
        dev = {k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")}
        b = api.QPBatch(ctx, B, 64, 128)
        for _ in range(3):
            b.setup_solve(dev["P"], dev["q"], dev["A"], dev["l"], dev["u"])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            b.setup_solve(dev["P"], dev["q"], dev["A"], dev["l"], dev["u"])
        e1.record()
        torch.cuda.synchronize()
        print("slice", sl, "batch", B, ctx.last_kernel, "%.3f ms" % (e0.elapsed_time(e1) / 10), flush=True)
        b.close()
