import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from sqp_solver_b200 import api
from sqp_solver_b200.synth import make_batch
ctx=api.Context(0)
for n,m,B in ((256,512,296),(128,256,592)):
    d=make_batch(B,n,m,seed0=0)
    dev={k: torch.from_numpy(d[k]).cuda() for k in ("P","q","A","l","u")}
    b=api.QPBatch(ctx,B,n,m); b.settings=api.default_settings(alpha=1.6,adaptive_rho=1)
    args=[dev[k] for k in ("P","q","A","l","u")]
    for rep in range(2):
        torch.cuda.synchronize(); t0=time.perf_counter(); b.setup(*args); torch.cuda.synchronize(); t1=time.perf_counter(); b.solve(*args); torch.cuda.synchronize(); t2=time.perf_counter()
    info=b.info(); its=int(info["iter"].sum()); ru=info["rho_updates"].sum()
    print(n,m,B,"setup %.2f ms  solve %.2f ms  iters %d  rho_updates %d  -> per factorization %.2f ms/wave, per iteration %.1f us (per QP)"%(1e3*(t1-t0),1e3*(t2-t1),its,ru,1e3*(t1-t0), 1e6*(t2-t1)/(its/B)))
