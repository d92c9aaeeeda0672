"""Cost of the termination checks in the headline kernel: ns per QP-iteration with check_termination = 25 (default), 100, 1000."""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from sqp_solver_b200 import api
from sqp_solver_b200.synth import make_batch
ctx = api.Context(0)
B, n, m = 4096, 64, 128
d = make_batch(B, n, m, seed0=0)
dev = {k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")}
for ct, eps in ((25, 1e-3), (25, 1e-12), (100, 1e-12), (1000, 1e-12)):
    b = api.QPBatch(ctx, B, n, m)
    b.settings = api.default_settings(check_termination=ct, eps_abs=eps, eps_rel=eps)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        ev0.record(); b.setup_solve(*[dev[k] for k in ("P", "q", "A", "l", "u")]); ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1); its = b.total_iters()
    print("check_termination %4d eps %.0e: %.2f ms, %d iterations, %.3f ns per QP-iteration (whole GPU)" % (ct, eps, ms, its, 1e6 * ms / its))
    b.close()
