import sys; sys.path.insert(0, '/root/repo')
import numpy as np
from sqp_solver_b200 import api
from sqp_solver_b200.synth import make_batch
from oracle import qp_oracle as O
ctx = api.Context(0)
for n, m, B in ((64, 128, 12), (32, 64, 12), (10, 14, 8)):
    d = make_batch(B, n, m, seed0=5)
    for name, kw in (("S1", {}), ("S2", dict(alpha=1.6, adaptive_rho=1))):
        b = api.QPBatch(ctx, B, n, m); b.settings = api.default_settings(**kw)
        b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"]); g64 = b.get()
        b.set_precision(True)
        b.setup_solve(d["P"], d["q"], d["A"], d["l"], d["u"]); g32 = b.get(); k = ctx.last_kernel
        rows = []
        for i in range(B):
            qp = O.QuadraticProblem(d["P"][i].reshape(n, n, order="F"), d["q"][i], d["A"][i].reshape(m, n, order="F"), d["l"][i], d["u"][i], dtype=np.float32)
            s = O.QPSolver(dtype=np.float32)
            st = s.settings()
            for kk, v in kw.items(): setattr(st, kk, v)
            s.setup(qp); s.solve(qp)
            x = s.primal_solution().astype(np.float64)
            inf = s.info()
            rel_o = np.linalg.norm(g32["x"][i] - x) / np.linalg.norm(x)
            rel_64 = np.linalg.norm(g32["x"][i] - g64["x"][i]) / np.linalg.norm(g64["x"][i])
            rel_o64 = np.linalg.norm(x - g64["x"][i]) / np.linalg.norm(g64["x"][i])
            rows.append((int(g32["status"][i]), int(inf.status), int(g64["status"][i]), int(g32["iter"][i]), int(inf.iter), int(g64["iter"][i]), rel_o, rel_64, rel_o64))
        print(k, n, m, name)
        for r in rows: print("   st gpu32/or32/gpu64 %d/%d/%d it %d/%d/%d  |gpu32-or32| %.1e |gpu32-gpu64| %.1e |or32-gpu64| %.1e" % r)
        b.close()
