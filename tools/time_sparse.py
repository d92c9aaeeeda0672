"""Config 5 timing: sparse-A QPs (shared pattern) through the sparse entry point. Usage: time_sparse.py [n m batch density S1|S2]"""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from sqp_solver_b200 import api
from sqp_solver_b200.synth import make_batch

n, m, B, dens = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])) if len(sys.argv) > 4 else (256, 512, 2048, 0.03)
sett = sys.argv[5] if len(sys.argv) > 5 else "S2"
kern = int(sys.argv[6]) if len(sys.argv) > 6 else 0
ctx = api.Context(0)
ctx.set_option(api.OPT_KERNEL, kern)
d = make_batch(B, n, m, seed0=0)
rng = np.random.default_rng(1); mask = rng.uniform(size=(m, n)) < dens; mask[np.arange(m), rng.integers(0, n, m)] = True
A3 = d["A"].reshape(B, n, m).transpose(0, 2, 1) * mask
x0 = rng.standard_normal((B, n)); c = np.einsum("bij,bj->bi", A3, x0)
d["l"] = c - rng.uniform(0, 1, (B, m)); d["u"] = c + rng.uniform(0, 1, (B, m))
rows, cols = np.nonzero(mask); outer = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int32)
vals = np.ascontiguousarray(A3[:, rows, cols]); inner = np.ascontiguousarray(cols.astype(np.int32))
dev = {k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "l", "u")}
dv, do, di = torch.from_numpy(vals).cuda(), torch.from_numpy(outer).cuda(), torch.from_numpy(inner).cuda()
b = api.QPBatch(ctx, B, n, m)
b.settings = api.default_settings(alpha=1.6, adaptive_rho=1) if sett == "S2" else api.default_settings()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b.setup_solve_sparse(dev["P"], dev["q"], dv, do, di, dev["l"], dev["u"], layout=api.SPARSE_CSR); torch.cuda.synchronize(); t1 = time.perf_counter()
info = b.info(); its = int(np.minimum(info["iter"], b.settings.max_iter).sum())
print("%dx%d batch %d nnz %d %s [%s]: %.2f ms, %d iterations, %d factorisations, status %s" % (
    n, m, B, len(inner), sett, ctx.last_kernel, 1e3 * (t1 - t0), its, int(info["rho_updates"].sum()), np.bincount(info["status"]).tolist()))
