"""Peer-memory strong scaling (sharding.PeerBatch) with two PROCESSES sharing cuda:0: CUDA IPC export/import of the owner's
arrays, the solve kernel reading its slice through the imported mapping, results written into the owner's arrays. (On one GPU the
"peer" is the same device; the 2- and 8-GPU runs of `bench.py --scaling strong --transport p2p` exercise NVLink.) gloo carries the handle
exchange and the barrier -- NCCL refuses two ranks on one device."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, n, m, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sqp_solver_b200 import api, sharding
        from sqp_solver_b200.synth import make_batch

        torch.cuda.set_device(0)
        ctx = api.Context(0)
        lo, hi = sharding.shard_range(batch, rank, world)
        qb = api.QPBatch(ctx, max(hi - lo, 1), n, m)
        pb = sharding.PeerBatch(ctx, n, m, batch, root=0)
        prob = None
        if rank == 0:
            d = make_batch(batch, n, m, seed0=123)
            prob = {k: torch.from_numpy(d[k]).cuda() for k in sharding.PROBLEM_KEYS}
        pb.load(prob)
        torch.cuda.synchronize()
        dist.barrier()
        pb.solve(qb)
        out = pb.results()
        if rank == 0:
            whole = api.QPBatch(ctx, batch, n, m)
            whole.setup_solve(*[prob[k] for k in sharding.PROBLEM_KEYS])
            ref = whole.get()
            ok = all(np.array_equal(out[k].cpu().numpy(), ref[k]) for k in ("x", "y", "z", "status", "iter", "rho_updates", "res_prim", "res_dual"))
            q.put((ok, int(ref["iter"].sum())))
            whole.close()
        pb.close()
        qb.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch,n,m", [(11, 64, 128), (6, 20, 30)])
def test_peer_batch_two_processes_one_gpu(batch, n, m):
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, batch, n, m, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    ok, iters = q.get(timeout=5)
    assert ok and iters > 0
