"""Batch-axis sharding of a QP batch over the GPUs of one node (one process per GPU, torch.distributed).

The QPs of a batch are independent, so the solve itself needs no collective.
NCCL (or gloo in the CPU tests) is used only where the north star names it: to SPLIT a batch that lives
on one rank across the ranks and to GATHER the results back.

    shard_range(batch, rank, world)        contiguous slice [lo, hi) owned by `rank`
    scatter_batch(problem, src)            rank `src` holds P,q,A,l,u for the whole batch -> every rank gets its slice
    gather_results(local, batch, dst)      per-rank x,y,z,status,iter,... -> whole-batch arrays on rank `dst`
    solve_sharded(problem, solve_local)    scatter -> solve_local(slice) on each rank -> gather

`solve_local` is the per-GPU solver (api.QPBatch.setup_solve + get in production; the tests inject the CPU
oracle to exercise the plumbing without a GPU).
"""
import torch
import torch.distributed as dist

PROBLEM_KEYS = ("P", "q", "A", "l", "u")
RESULT_DTYPES = dict(x=torch.float64, y=torch.float64, z=torch.float64, status=torch.int32, iter=torch.int32,
                     rho_updates=torch.int32, rho_estimate=torch.float64, res_prim=torch.float64, res_dual=torch.float64)


def shard_range(batch, rank, world):
    """Contiguous, balanced slices: the first batch % world ranks get one extra QP."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def scatter_batch(problem, n, m, batch, src=0):
    """problem: dict of [batch, width] float64 tensors on rank `src` (ignored elsewhere).
    Returns this rank's slice as a dict of tensors on the backend's device."""
    rank, world = dist.get_rank(), dist.get_world_size()
    widths = dict(P=n * n, q=n, A=m * n, l=m, u=m)
    dev = _device()
    lo, hi = shard_range(batch, rank, world)
    out = {}
    for k in PROBLEM_KEYS:
        mine = torch.empty(hi - lo, widths[k], dtype=torch.float64, device=dev)
        if world == 1:
            mine.copy_(problem[k].reshape(batch, widths[k]))
        else:
            # uneven slices: grouped point-to-point sends instead of dist.scatter (which needs equal sizes)
            ops = []
            if rank == src:
                full = problem[k].reshape(batch, widths[k]).to(dev)
                for r in range(world):
                    rlo, rhi = shard_range(batch, r, world)
                    if r == src:
                        mine.copy_(full[rlo:rhi])
                    elif rhi > rlo:
                        ops.append(dist.P2POp(dist.isend, full[rlo:rhi].contiguous(), r))
            elif hi > lo:
                ops.append(dist.P2POp(dist.irecv, mine, src))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
        out[k] = mine
    return out


def gather_results(local, batch, dst=0):
    """local: dict of per-rank result tensors (leading dim = slice length). Returns whole-batch tensors on
    rank `dst` (None elsewhere), in the original batch order."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _device()
    out = {} if rank == dst else None
    for k, v in local.items():
        v = v.to(dev).contiguous()
        if world == 1:
            out[k] = v.clone()
            continue
        ops = []
        if rank == dst:
            full = torch.empty((batch,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
            for r in range(world):
                rlo, rhi = shard_range(batch, r, world)
                if r == dst:
                    full[rlo:rhi].copy_(v)
                elif rhi > rlo:
                    ops.append(dist.P2POp(dist.irecv, full[rlo:rhi], r))
            out[k] = full
        elif v.shape[0] > 0:
            ops.append(dist.P2POp(dist.isend, v, dst))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
    return out


def solve_sharded(problem, n, m, batch, solve_local, root=0):
    """Strong-scaling path: a batch owned by `root` is split over the ranks, solved, and gathered back.
    `solve_local(slice_dict) -> result dict of tensors`."""
    mine = scatter_batch(problem, n, m, batch, src=root)
    local = solve_local(mine)
    return gather_results(local, batch, dst=root)


def gpu_solve_local(qbatch):
    """Per-rank solver for solve_sharded backed by an api.QPBatch (capacity >= slice length)."""
    def run(mine):
        cnt = mine["q"].shape[0]
        dev = mine["q"].device
        res = {k: torch.empty((cnt,) + ((qbatch.n,) if k == "x" else (qbatch.m,) if k in ("y", "z") else ()), dtype=dt, device=dev)
               for k, dt in RESULT_DTYPES.items()}
        if cnt:
            qbatch.setup_solve(mine["P"], mine["q"], mine["A"], mine["l"], mine["u"], count=cnt)
            qbatch.get_into(count=cnt, **res)
        return res
    return run
