"""N>1 path on CPU: world_size-2 (and 3, ragged) gloo runs of the split / solve / gather plumbing in
sqp_solver_b200/sharding.py. The per-rank solver is the CPU oracle here (the GPU kernel is covered by the
-m gpu tests); what is checked is that sharding reproduces the unsharded result exactly, in order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


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
        from oracle import qp_oracle as O
        from sqp_solver_b200 import sharding
        from sqp_solver_b200.synth import make_batch

        prob = None
        if rank == 0:
            d = make_batch(batch, n, m, seed0=77)
            prob = {k: torch.from_numpy(d[k]) for k in sharding.PROBLEM_KEYS}

        def solve_local(mine):
            cnt = mine["q"].shape[0]
            if cnt == 0:
                return {k: torch.empty((0,) + ((n,) if k == "x" else (m,) if k in ("y", "z") else ()), dtype=dt)
                        for k, dt in sharding.RESULT_DTYPES.items()}
            o = O.solve_batch(*(mine[k].numpy() for k in sharding.PROBLEM_KEYS), nthreads=1)
            return {k: torch.from_numpy(np.ascontiguousarray(o[k])).to(sharding.RESULT_DTYPES[k]) for k in sharding.RESULT_DTYPES}

        out = sharding.solve_sharded(prob, n, m, batch, solve_local, root=0)
        lo, hi = sharding.shard_range(batch, rank, world)
        if rank == 0:
            ref = O.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], nthreads=1)
            ok = all(np.array_equal(out[k].numpy(), ref[k].astype(out[k].numpy().dtype)) for k in sharding.RESULT_DTYPES)
            q.put(("result", ok, int(ref["iter"].sum())))
        else:
            assert out is None
        q.put(("range", rank, lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,batch", [(2, 10), (3, 7), (2, 1)])
def test_split_solve_gather_matches_unsharded(world, batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, 6, 9, q)) for r in range(world)]
    for p in procs:
        p.start()
    msgs = [q.get(timeout=120) for _ in range(world + 1)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = [m for m in msgs if m[0] == "result"][0]
    assert res[1], "sharded result differs from the unsharded one"
    ranges = sorted((m[1], m[2], m[3]) for m in msgs if m[0] == "range")
    assert ranges[0][1] == 0 and ranges[-1][2] == batch
    for a, b in zip(ranges, ranges[1:]):
        assert a[2] == b[1]  # contiguous, no gaps or overlap


def test_shard_range_is_balanced():
    from sqp_solver_b200.sharding import shard_range

    for batch in (0, 1, 7, 8192, 8193):
        for world in (1, 2, 3, 8):
            sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
            assert sum(sizes) == batch and max(sizes) - min(sizes) <= 1
            assert shard_range(batch, 0, world)[0] == 0 and shard_range(batch, world - 1, world)[1] == batch
