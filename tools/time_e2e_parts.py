"""Where the end-to-end step of the headline workload spends its time beyond the device-resident solve: wall clock of the staged
setup_solve(HOST_PTRS) call, of get_into, and of the device-pointer call, at several chunk counts.
Usage: python tools/time_e2e_parts.py"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from sqp_solver_b200 import api
from sqp_solver_b200.synth import make_batch

ctx = api.Context(0)
B, n, m = 8192, 64, 128
d = make_batch(B, n, m, seed0=0)
pin = {k: torch.from_numpy(d[k]).pin_memory() for k in ("P", "q", "A", "l", "u")}
hp = {k: v.numpy() for k, v in pin.items()}
dev = [torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")]
ox = torch.empty(B, n, dtype=torch.float64).pin_memory(); oy = torch.empty(B, m, dtype=torch.float64).pin_memory()
ost = torch.empty(B, dtype=torch.int32).pin_memory(); oit = torch.empty(B, dtype=torch.int32).pin_memory()
qb = api.QPBatch(ctx, B, n, m)
def t(f, reps=6):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return 1e3 * min(ts), 1e3 * sorted(ts)[len(ts) // 2]
print("device-pointer setup_solve  min %.3f median %.3f ms" % t(lambda: qb.setup_solve(*dev)))
for ch in (16, 32, 8):
    ctx.set_option(api.OPT_H2D_CHUNKS, ch)
    print("chunks %2d: host-pointer setup_solve  min %.3f median %.3f ms" % ((ch,) + t(lambda: qb.setup_solve(hp["P"], hp["q"], hp["A"], hp["l"], hp["u"], count=B))))
ctx.set_option(api.OPT_H2D_CHUNKS, 16)
print("get_into (x, y, status, iter)  min %.3f median %.3f ms" % t(lambda: qb.get_into(count=B, x=ox.numpy(), y=oy.numpy(), status=ost.numpy(), iter=oit.numpy())))
def both():
    qb.setup_solve(hp["P"], hp["q"], hp["A"], hp["l"], hp["u"], count=B)
    qb.get_into(count=B, x=ox.numpy(), y=oy.numpy(), status=ost.numpy(), iter=oit.numpy())
print("step (both)  min %.3f median %.3f ms" % t(both))
# pure H2D of the same bytes for reference
big = torch.empty(B * (n * n + n + m * n + 2 * m), dtype=torch.float64).pin_memory(); dbig = torch.empty_like(big, device="cuda")
print("plain H2D of the step's 826 MB  min %.3f median %.3f ms" % t(lambda: dbig.copy_(big, non_blocking=True)))
