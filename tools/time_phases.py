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

# sparse A (shared pattern, 3 % dense) through the blocked kernel: fused setup+solve, against the dense call on the same data
for n,m,B,dens in ((256,512,296,0.03),(128,256,592,0.06)):
    d=make_batch(B,n,m,seed0=0)
    rng=np.random.default_rng(1); mask=rng.uniform(size=(m,n))<dens; mask[np.arange(m),rng.integers(0,n,m)]=True
    A3=d["A"].reshape(B,n,m).transpose(0,2,1)*mask
    x0=rng.standard_normal((B,n)); c=np.einsum("bij,bj->bi",A3,x0)
    d["l"]=c-rng.uniform(0,1,(B,m)); d["u"]=c+rng.uniform(0,1,(B,m))
    d["A"]=np.ascontiguousarray(A3.transpose(0,2,1).reshape(B,n*m))
    rows,cols=np.nonzero(mask); outer=np.concatenate([[0],np.cumsum(mask.sum(axis=1))]).astype(np.int32)
    vals=np.ascontiguousarray(A3[:,rows,cols]); inner=np.ascontiguousarray(cols.astype(np.int32))
    dev={k: torch.from_numpy(d[k]).cuda() for k in ("P","q","A","l","u")}
    dv,do,di=torch.from_numpy(vals).cuda(),torch.from_numpy(outer).cuda(),torch.from_numpy(inner).cuda()
    b=api.QPBatch(ctx,B,n,m); b.settings=api.default_settings(alpha=1.6,adaptive_rho=1)
    for rep in range(2):
        torch.cuda.synchronize(); t0=time.perf_counter()
        b.setup_solve_sparse(dev["P"],dev["q"],dv,do,di,dev["l"],dev["u"],layout=api.SPARSE_CSR); torch.cuda.synchronize(); t1=time.perf_counter()
    ks=ctx.last_kernel; its=int(b.info()["iter"].sum()); xs=b.get()["x"]
    for rep in range(2):
        torch.cuda.synchronize(); t2=time.perf_counter()
        b.setup_solve(*[dev[k] for k in ("P","q","A","l","u")]); torch.cuda.synchronize(); t3=time.perf_counter()
    kd=ctx.last_kernel; itd=int(b.info()["iter"].sum()); xd=b.get()["x"]
    print(n,m,B,"nnz %d  sparse[%s] %.2f ms (%d it)   dense[%s] %.2f ms (%d it)  max|dx| %.2e"%(len(inner),ks,1e3*(t1-t0),its,kd,1e3*(t3-t2),itd,np.abs(xs-xd).max()))
