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


# ---- strong scaling without a split/gather step: peer memory over NVLink -------------------------------------------------
class PeerBatch:
    """One batch owned by rank `root`, solved by all the GPUs of the node with NO data-path collective and no staging copy:
    the owner's input arrays and result arrays are shared with the other processes as CUDA IPC mappings, every rank's solve kernel
    reads its slice of P, q, A, l, u straight out of the owner's HBM over NVLink (the kernel's TMA bulk copies take the peer
    address, so the transfer overlaps the solve QP by QP) and its results go straight into the owner's arrays (device-to-peer
    copies issued by sqpb200_qp_batch_get). torch.distributed is used for the handle exchange and the closing barrier only.

        pb = PeerBatch(ctx, n, m, batch, root=0)      # collective: allocates on root, exchanges handles
        pb.load(problem)                              # root: dict of CUDA tensors P,q,A,l,u -> the shared input arrays
        pb.solve(qbatch, stream)                      # every rank: solve its slice (asynchronous on `stream`)
        out = pb.results()                            # root: dict of CUDA tensors (after a barrier)
    """

    IN_WIDTH = lambda self, k: dict(P=self.n * self.n, q=self.n, A=self.m * self.n, l=self.m, u=self.m)[k]

    def __init__(self, ctx, n, m, batch, root=0):
        import numpy as np

        from . import api

        self.ctx, self.n, self.m, self.batch, self.root = ctx, n, m, batch, root
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.lo, self.hi = shard_range(batch, self.rank, self.world)
        self._np = {torch.float64: np.float64, torch.int32: np.int32}
        self.out_width = dict(x=n, y=m, z=m, status=1, iter=1, rho_updates=1, rho_estimate=1, res_prim=1, res_dual=1)
        names = [("in", k) for k in PROBLEM_KEYS] + [("out", k) for k in RESULT_DTYPES]
        self.buf = {}
        handles = [None]
        if self.rank == root:
            hs = {}
            for kind, k in names:
                dt = np.float64 if kind == "in" else self._np[RESULT_DTYPES[k]]
                width = self.IN_WIDTH(k) if kind == "in" else self.out_width[k]
                self.buf[(kind, k)] = ctx.dev_alloc(batch * width * np.dtype(dt).itemsize, dt)
                hs[(kind, k)] = ctx.ipc_export(self.buf[(kind, k)])
            handles = [hs]
        dist.broadcast_object_list(handles, src=root)
        if self.rank != root:
            for kind, k in names:
                dt = np.float64 if kind == "in" else self._np[RESULT_DTYPES[k]]
                self.buf[(kind, k)] = ctx.ipc_import(handles[0][(kind, k)], dt)
        self._api = api

    def load(self, problem, stream=None):
        """Root copies the problem into the shared input arrays; returns on every rank once the data has landed (stream
        synchronisation on the root + barrier), so a following solve() on any rank reads complete inputs."""
        if self.rank == self.root:
            for k in PROBLEM_KEYS:
                t = problem[k].contiguous()
                self.ctx.dev_copy(self.buf[("in", k)], t, t.numel() * 8, stream)
            torch.cuda.synchronize()
        dist.barrier()

    def solve(self, qbatch, stream=None):
        cnt = self.hi - self.lo
        if cnt <= 0:
            return
        ins = [self.buf[("in", k)].offset(self.lo * self.IN_WIDTH(k)) for k in PROBLEM_KEYS]
        outs = {k: self.buf[("out", k)].offset(self.lo * self.out_width[k]) for k in RESULT_DTYPES}
        # ONE kernel per rank and nothing else: inputs pulled from the owner's HBM by the kernel's TMA loads, results written into
        # the owner's arrays by the kernel's epilogue
        qbatch.setup_solve_to(*ins, outs, count=cnt, stream=stream)

    def results(self):
        """Barrier, then (root) the whole-batch results as CUDA tensors."""
        torch.cuda.synchronize()
        dist.barrier()
        if self.rank != self.root:
            return None
        out = {}
        for k, dt in RESULT_DTYPES.items():
            w = self.out_width[k]
            t = torch.empty((self.batch, w) if k in ("x", "y", "z") else (self.batch,), dtype=dt, device="cuda")
            self.ctx.dev_copy(t, self.buf[("out", k)], t.numel() * t.element_size())
            out[k] = t
        torch.cuda.synchronize()
        return out

    def close(self):
        torch.cuda.synchronize()
        dist.barrier()
        for key, dp in self.buf.items():
            if self.rank == self.root:
                self.ctx.dev_free(dp)
            else:
                self.ctx.ipc_release(dp)
        self.buf = {}
