#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched QP-subproblem hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores (oracle)

A "step" is one pass of the hot path (fresh setup + ADMM solve, reference src/sqp.cpp:221-222)
over one batch of synthetic QPs: BASELINE.json configs[2], batch=8192 dense QPs n=64 m=128 fp64,
reference default settings. One process per GPU (torchrun for N>1); the batch axis is sharded
with no data-path collective (weak scaling: every rank solves its own 8192 QPs).

Prints ONE JSON line (rank 0). Keys follow the driver's contract; see DESIGN.md section
"Measurement" for how `roofline.achieved` is formed from the algorithmic bytes of SURVEY.md 8(d).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "QP-subproblems/sec (batch=8192, n=64, m=128)"
UNIT = "QP/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config3", choices=["config3", "config2", "config5"],
                    help="config3 (default, the headline: batch 8192 dense 64x128), config2 (batch 1024 dense 32x64, one warp per QP), "
                         "config5 (batch 2048 sparse-A 256x512, shared CSR pattern, cluster kernel). Only config3 is the driver's bench line")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--m", type=int, default=0)
    ap.add_argument("--density", type=float, default=0.03, help="config5: probability that an entry of A is stored")
    ap.add_argument("--settings", default="S1", choices=["S1", "S2"],
                    help="S1 = reference defaults (headline); S2 = alpha 1.6 + adaptive rho (SURVEY.md 8d)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "tile"])
    ap.add_argument("--tile-warps", type=int, default=0, choices=[0, 1, 2, 4, 8])
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): every rank solves its own batch. strong: ONE batch on rank 0 "
                         "is split over the ranks with NCCL send/recv, solved, and gathered back (north star's split/gather)")
    ap.add_argument("--transport", default="nccl", choices=["nccl", "p2p"],
                    help="--scaling strong only. nccl: split with NCCL send/recv, solve, gather with NCCL. p2p: no split/gather step at all -- "
                         "every rank's solve kernel reads its slice out of rank 0's HBM over NVLink (CUDA IPC mapping, TMA from peer "
                         "memory) and writes its results into rank 0's arrays")
    ap.add_argument("--cpu-sample", type=int, default=0, help="QPs in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--fp32", action="store_true", help="compute in fp32 (the reference's QPSolver<float>; NOT the headline: the parity bar "
                                                       "of the metric is fp64). dtype reads f32 and the CPU baseline is skipped")
    return ap.parse_args()


def algorithmic_bytes(n, m, iters_executed, checks, factorizations, count):
    """SURVEY.md 8(d) canonical accounting (fp64): per ADMM iteration A twice + packed chol(H) twice,
    per termination check P and A once, per QP compulsory I/O, per (re)factorisation P and A once."""
    b_iter = 16 * m * n + 8 * n * (n + 1)
    b_check = 8 * (n * n + m * n)
    b_io = 8 * (n * n + n + m * n + 2 * m) + 8 * (n + m) + 16
    b_fact = 8 * (n * n + m * n)
    return count * b_io + iters_executed * b_iter + checks * b_check + factorizations * b_fact


def settings_kwargs(name):
    return {} if name == "S1" else dict(alpha=1.6, adaptive_rho=1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(sm)[len(sm) // 2:]  # upper half: samples taken under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(d, settings_name, sample, steps=1):
    """The reference algorithm (oracle restatement -- Eigen is absent, so kind='port') on this box's host cores:
    headline = gcc -O2, OpenMP dynamic schedule over a bounded sample of the same workload on all cores; BASELINE.md
    section 4's other rows (one core; -O3 -march=native built on this box) ride along in `extra`."""
    from oracle import qp_oracle as O

    O.build()
    cores = O.num_procs()
    if sample <= 0:
        sample = min(d["batch"], 64 * cores)
    st = O.default_settings(**settings_kwargs(settings_name))

    def timed(count, nthreads, native=False):
        sl = slice(0, count)
        best, out = None, None
        for _ in range(steps):
            t0 = time.perf_counter()
            out = O.solve_batch(d["P"][sl], d["q"][sl], d["A"][sl], d["l"][sl], d["u"][sl], st, nthreads=nthreads, native=native)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        its = int(np.minimum(out["iter"], st.max_iter).sum())
        return count / best, its / best, best, out["threads"]

    qps, itps, sec, threads = timed(sample, cores)
    extra = {}
    try:
        one = max(4, min(sample, 24))
        q1, i1, s1, _ = timed(one, 1)
        extra["one_core_O2"] = {"value": q1, "admm_iters_per_s": i1, "sample": "%d QPs, %.2f s" % (one, s1)}
        O.lib(native=True)  # compiles on this box (outside the timed region)
        qn, inn, sn, _ = timed(sample, cores, native=True)
        extra["all_cores_O3_march_native"] = {"value": qn, "admm_iters_per_s": inn, "sample": "%d QPs, %.2f s" % (sample, sn)}
    except Exception as e:  # the extra rows are informative only
        extra["error"] = repr(e)
    return {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "first %d QPs of the same batch, %s, gcc -O2 oracle restatement of src/qp.cpp (Eigen absent), "
                      "OpenMP schedule(dynamic) over %d threads, %.2f s" % (sample, settings_name, threads, sec),
            "admm_iters_per_s": itps, "seconds": sec, "extra": extra}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from sqp_solver_b200.synth import make_batch
    from oracle import qp_oracle as O

    O.build()
    cores = O.num_procs()
    metric, wl = METRIC, "configs[%d]: batch=%d dense QPs n=%d m=%d fp64" % (1 if args.workload == "config2" else 2, args.batch, args.n, args.m)
    if args.workload == "config5":  # sparse A, densified for the CPU path (the reference's sparse variant is dead code)
        from sqp_solver_b200.synth import densify, make_sparse_batch

        sample = args.cpu_sample or min(args.batch, 2 * cores)
        d = make_sparse_batch(sample, args.n, args.m, density=args.density, seed0=0)
        d["A"] = densify(d)
        metric = "QP-subproblems/sec (batch=%d, n=%d, m=%d, sparse A nnz=%d)" % (args.batch, args.n, args.m, d["nnz"])
        wl = "configs[4]: batch=%d sparse-A QPs n=%d m=%d fp64 (densified for the CPU path)" % (args.batch, args.n, args.m)
    else:
        sample = args.cpu_sample or min(args.batch, 64 * cores)
        d = make_batch(sample, args.n, args.m, seed0=0)
        if (args.batch, args.n, args.m) != WORKLOADS["config3"]:
            metric = "QP-subproblems/sec (batch=%d, n=%d, m=%d)" % (args.batch, args.n, args.m)
    st = O.default_settings(**settings_kwargs(args.settings))
    for _ in range(max(1, min(args.warmup, 1))):
        O.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], st, nthreads=cores)
    t0 = time.perf_counter()
    its = 0
    for _ in range(args.steps):
        out = O.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], st, nthreads=cores)
        its += int(np.minimum(out["iter"], st.max_iter).sum())
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {"impl": "reference", "metric": metric, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s, settings %s; each step is a bounded sample of %d QPs of that batch on host cores" % (wl, args.settings, sample)},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d QPs per step, gcc -O2 oracle restatement of src/qp.cpp (Eigen absent from the image, "
                                       "so the reference cannot be compiled), OpenMP over %d threads" % (sample, out["threads"])},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "admm_iters_per_s": its / dt, "gpu_launches": 0}
    print(json.dumps(line))


def run_strong(args, rank, world, ctx, api, d, settings):
    """Strong scaling: rank 0 owns the whole batch; NCCL send/recv splits it, each rank solves its slice,
    NCCL gathers x, y, z and the info arrays back to rank 0. The split and gather are inside the timed region."""
    import torch
    import torch.distributed as dist

    from sqp_solver_b200 import sharding

    B, n, m = args.batch, args.n, args.m
    lo, hi = sharding.shard_range(B, rank, world)
    qb = api.QPBatch(ctx, max(hi - lo, 1), n, m)
    qb.settings = settings
    prob = {k: torch.from_numpy(d[k]).cuda() for k in sharding.PROBLEM_KEYS} if rank == 0 else None
    solve_local = sharding.gpu_solve_local(qb)
    stream = torch.cuda.current_stream()
    pb = None
    if args.transport == "p2p":
        pb = sharding.PeerBatch(ctx, n, m, B, root=0)
        pb.load(prob, stream=stream.cuda_stream)  # the batch is resident in rank 0's HBM before the timed region, as in the nccl flow
        torch.cuda.synchronize()
        dist.barrier()

    def step():
        if pb is None:
            return sharding.solve_sharded(prob, n, m, B, solve_local, root=0)
        pb.solve(qb, stream=stream.cuda_stream)
        return None

    for _ in range(args.warmup):
        out = step()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        out = step()
    ev1.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    if pb is not None:
        out = pb.results()
        if rank == 0:  # same answers as a plain single-GPU solve of the first shard
            chk = api.QPBatch(ctx, hi - lo, n, m)
            chk.settings = settings
            chk.setup_solve(*[prob[k][lo:hi].contiguous() for k in sharding.PROBLEM_KEYS])
            ref = chk.get(fields=("x", "iter"))
            assert (out["iter"][lo:hi].cpu().numpy() == ref["iter"]).all() and (out["x"][lo:hi].cpu().numpy() == ref["x"]).all()
            chk.close()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        its = int(torch.clamp(out["iter"], max=settings.max_iter).sum().item())
        print(json.dumps({"metric": METRIC, "value": B * args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": ("configs[2]: ONE batch=%d n=%d m=%d on rank 0, NCCL split -> solve -> NCCL gather (both inside the "
                                                  "timed region)" if pb is None else "configs[2]: ONE batch=%d n=%d m=%d in rank 0's HBM; every rank's solve "
                                                  "kernel reads its slice over NVLink (CUDA IPC peer mapping, TMA from peer memory) and writes its "
                                                  "results into rank 0's arrays: no split/gather step, no collective") % (B, n, m),
                                     "kernel": ctx.last_kernel, "transport": args.transport},
                          "admm_iters_per_s": its / (ms / 1e3 / args.steps)}))
    if pb is not None:
        pb.close()


WORKLOADS = {"config3": (8192, 64, 128), "config2": (1024, 32, 64), "config5": (2048, 256, 512)}


def sparse_algorithmic_bytes(n, m, nnz, iters_executed, checks, factorizations, count):
    """SURVEY.md 8(d) accounting carried over to compressed A (8-byte value + 4-byte index per stored entry):
    per iteration A twice + packed chol(H) twice, per check P and A once, per QP compulsory I/O, per factorisation P and A once."""
    a = 12 * nnz
    b_iter = 2 * a + 8 * n * (n + 1)
    b_check = 8 * n * n + a
    b_io = 8 * (n * n + n + 2 * m) + a + 8 * (n + m) + 16
    b_fact = 8 * n * n + a
    return count * b_io + iters_executed * b_iter + checks * b_check + factorizations * b_fact


def run_config5(args, rank, world, local_rank):
    """BASELINE.json configs[4]: sparse-A QPs (one CSR pattern for the batch) through sqpb200_qp_batch_setup_solve_sparse.
    Same JSON keys as the headline line; weak scaling (every rank its own batch)."""
    import torch
    import torch.distributed as dist

    from sqp_solver_b200 import api
    from sqp_solver_b200.synth import densify, make_sparse_batch

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = api.Context(local_rank)
    B, n, m = args.batch, args.n, args.m
    d = make_sparse_batch(B, n, m, density=args.density, seed0=rank * B)
    nnz = d["nnz"]
    settings = api.default_settings(**settings_kwargs(args.settings))
    dev = {k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "vals", "outer", "inner", "l", "u")}
    qb = api.QPBatch(ctx, B, n, m)
    qb.settings = settings
    stream = torch.cuda.current_stream()

    def step_device():
        qb.setup_solve_sparse(dev["P"], dev["q"], dev["vals"], dev["outer"], dev["inner"], dev["l"], dev["u"], layout=api.SPARSE_CSR,
                              stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    total_iters = qb.total_iters()
    info0 = qb.get(fields=("status", "iter", "rho_updates"))
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    agg = torch.tensor([float(total_iters), float(B)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    ms_max, all_iters, all_qps = float(t.item()), float(agg[0].item()), float(agg[1].item())

    e2e = None
    if not args.no_e2e:
        pin = {k: torch.from_numpy(d[k]).pin_memory() for k in ("P", "q", "vals", "l", "u")}
        hp = {k: v.numpy() for k, v in pin.items()}
        ox = torch.empty(B, n, dtype=torch.float64).pin_memory()
        oy = torch.empty(B, m, dtype=torch.float64).pin_memory()
        ost = torch.empty(B, dtype=torch.int32).pin_memory()
        oit = torch.empty(B, dtype=torch.int32).pin_memory()

        def step_host():
            qb.setup_solve_sparse(hp["P"], hp["q"], hp["vals"], d["outer"], d["inner"], hp["l"], hp["u"], layout=api.SPARSE_CSR)
            qb.get_into(x=ox.numpy(), y=oy.numpy(), status=ost.numpy(), iter=oit.numpy())

        for _ in range(min(args.warmup, 2)):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        assert (ost.numpy() == info0["status"]).all() and (oit.numpy() == info0["iter"]).all()
        e2e = {"value": all_qps * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": B * (8 * (n * n + n + 2 * m + nnz)) + 4 * (m + 1 + nnz),
               "d2h_bytes_per_step": B * (8 * (n + m) + 8), "ms_per_step": 1e3 * dt / args.steps,
               "how": "pinned host buffers -> sqpb200_qp_batch_setup_solve_sparse(HOST_PTRS): one persistent launch whose work queue is gated "
                      "on the chunk-by-chunk H2D staging -> sqpb200_qp_batch_get to pinned host; wall clock, max over ranks"}
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    executed = np.minimum(info0["iter"], settings.max_iter).astype(np.int64)
    executed[info0["status"] == api.NUMERICAL_ISSUES] = 0
    ct = settings.check_termination
    checks = int((executed // ct).sum()) if ct > 0 else 0
    # rho_updates is cumulative over the calls on this batch object (qp.cpp:313): per launch = total / launches so far
    calls = args.warmup + args.steps + (0 if args.no_e2e else min(args.warmup, 2) + args.steps)
    facts = int(info0["rho_updates"].sum()) // max(calls, 1)
    bytes_per_launch = sparse_algorithmic_bytes(n, m, nnz, int(executed.sum()), checks, facts, B)
    sec = ms / 1e3 / args.steps
    peak, peak_src = peaks()
    achieved = bytes_per_launch / sec / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("%s_%dx%d_b%d_%s" % (ctx.last_kernel, n, m, B, args.settings))
    except Exception:
        traffic = None
    line = {
        "metric": "QP-subproblems/sec (batch=%d, n=%d, m=%d, sparse A nnz=%d)" % (B, n, m, nnz), "value": all_qps * args.steps / (ms_max / 1e3),
        "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[4]: batch=%d sparse-A QPs n=%d m=%d fp64 per GPU, one CSR pattern for the batch (density %.3g + one "
                               "entry per row: nnz=%d), settings %s, fresh setup+solve per step" % (B, n, m, args.density, nnz, args.settings),
                   "batch_per_gpu": B, "n": n, "m": m, "nnz": nnz, "kernel": ctx.last_kernel,
                   "parallelism": "batch-sharded x%d, no collective on the data path" % world,
                   "l2": "inputs are %.0f MB per step, larger than the 126 MB L2" % (8 * B * (n * n + n + 2 * m + nnz) / 1e6)},
        "admm_iters_per_s": all_iters / (ms_max / 1e3 / args.steps), "admm_iters_per_step": all_iters,
        "factorisations_per_step": facts,
        "status_histogram": {api.STATUS_NAMES[k]: int((info0["status"] == k).sum()) for k in np.unique(info0["status"])},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch, "kernel_ms": 1e3 * sec,
                     "note": "SURVEY.md 8(d) accounting with A compressed (12 B per stored entry). H^-1 lives in the shared memory of a "
                             "4-CTA cluster, so DRAM traffic is the compulsory I/O only; the kernel is bound by cluster barriers and the "
                             "serial pivot chain of the factorisation (DESIGN.md 4.4), not by HBM"},
        "clocks": clocks,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu_baseline:
        from oracle import qp_oracle as O

        O.build()
        cores = O.num_procs()
        sample = args.cpu_sample or min(B, 2 * cores)
        A = densify(d, 0, sample)
        st = O.default_settings(**settings_kwargs(args.settings))
        t0 = time.perf_counter()
        out = O.solve_batch(d["P"][:sample], d["q"][:sample], A, d["l"][:sample], d["u"][:sample], st, nthreads=cores)
        dt = time.perf_counter() - t0
        assert (out["status"] == info0["status"][:sample]).all(), "GPU and oracle disagree on the status of the sampled QPs"
        line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "first %d QPs of the same batch (densified: the reference's sparse variant is dead code), %s, gcc -O2 "
                                          "oracle restatement of src/qp.cpp, OpenMP over %d threads, %.2f s" % (sample, args.settings, out["threads"], dt),
                                "admm_iters_per_s": int(np.minimum(out["iter"], st.max_iter).sum()) / dt}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    wb, wn, wm = WORKLOADS[args.workload]
    args.batch, args.n, args.m = args.batch or wb, args.n or wn, args.m or wm
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "config5":
        run_config5(args, rank, world, local_rank)
        return

    import torch
    import torch.distributed as dist

    from sqp_solver_b200 import api
    from sqp_solver_b200.synth import make_batch

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = api.Context(local_rank)
    ctx.set_option(api.OPT_KERNEL, {"auto": 0, "generic": 1, "tile": 2}[args.kernel])
    ctx.set_option(api.OPT_TILE_WARPS, args.tile_warps)
    ctx.set_option(api.OPT_CTAS_PER_SM, args.ctas_per_sm)
    B, n, m = args.batch, args.n, args.m
    strong = args.scaling == "strong" and world > 1
    # weak: every rank owns a disjoint shard of the seed sequence; strong: only rank 0 builds the (single) batch
    d = make_batch(B if (not strong or rank == 0) else 1, n, m, seed0=0 if strong else rank * B)
    settings = api.default_settings(**settings_kwargs(args.settings))
    if strong:
        run_strong(args, rank, world, ctx, api, d, settings)
        dist.destroy_process_group()
        return
    dev = {k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")}
    qb = api.QPBatch(ctx, B, n, m)
    qb.settings = settings
    if args.fp32:
        qb.set_precision(True)
        args.no_cpu_baseline = True
    stream = torch.cuda.current_stream()

    def step_device():
        qb.setup_solve(dev["P"], dev["q"], dev["A"], dev["l"], dev["u"], stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    total_iters = qb.total_iters()
    info = qb.get(fields=("status", "iter", "rho_updates"))
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    agg = torch.tensor([float(total_iters), float(B)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    all_iters, all_qps = float(agg[0].item()), float(agg[1].item())

    # ---- end to end: host (pinned) buffers through the C-ABI, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        pin = {k: torch.from_numpy(d[k]).pin_memory() for k in ("P", "q", "A", "l", "u")}
        hp = {k: v.numpy() for k, v in pin.items()}
        ox = torch.empty(B, n, dtype=torch.float64).pin_memory()
        oy = torch.empty(B, m, dtype=torch.float64).pin_memory()
        ost = torch.empty(B, dtype=torch.int32).pin_memory()
        oit = torch.empty(B, dtype=torch.int32).pin_memory()

        def step_host():
            qb.setup_solve(hp["P"], hp["q"], hp["A"], hp["l"], hp["u"])
            qb.get_into(x=ox.numpy(), y=oy.numpy(), status=ost.numpy(), iter=oit.numpy())

        for _ in range(min(args.warmup, 2)):
            step_host()
        barrier()
        l0 = ctx.launch_count
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e_launches = ctx.launch_count - l0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        assert (ost.numpy() == info["status"]).all() and (oit.numpy() == info["iter"]).all()
        e2e = {"value": all_qps * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": 8 * B * (n * n + n + m * n + 2 * m), "d2h_bytes_per_step": B * (8 * (n + m) + 8),
               "ms_per_step": 1e3 * dt / args.steps, "launches": e2e_launches,
               "how": "pinned host buffers -> sqpb200_qp_batch_setup_solve(HOST_PTRS): one persistent launch whose work queue is "
                      "gated on the chunk-by-chunk H2D staging -> sqpb200_qp_batch_get to pinned host; wall clock between device "
                      "synchronisations, max over ranks"}
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    executed = np.minimum(info["iter"], settings.max_iter).astype(np.int64)
    executed[info["status"] == api.NUMERICAL_ISSUES] = 0
    ct = settings.check_termination
    checks = int((executed // ct).sum()) if ct > 0 else 0
    facts = int(info["rho_updates"].sum())
    bytes_per_launch = algorithmic_bytes(n, m, int(executed.sum()), checks, facts, B)
    sec_per_launch = ms / 1e3 / args.steps  # rank 0's own kernel time
    peak, peak_src = peaks()
    achieved = bytes_per_launch / sec_per_launch / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("%s_%dx%d_b%d_%s" % (ctx.last_kernel, n, m, B, args.settings))
        except Exception:
            traffic = None
    line = {
        "metric": METRIC if (B, n, m) == WORKLOADS["config3"] else "QP-subproblems/sec (batch=%d, n=%d, m=%d)" % (B, n, m),
        "value": all_qps * args.steps / (ms_max / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.fp32 else "f64", "data": "synthetic",
        "config": {"workload": "configs[%d]: batch=%%d dense QPs n=%%d m=%%d fp64 per GPU, settings %%s (%%s), fresh setup+solve per step" % (1 if args.workload == "config2" else 2)
                               % (B, n, m, args.settings, "reference defaults qp.hpp:38-53" if args.settings == "S1" else
                                  "alpha=1.6 adaptive_rho"),
                   "batch_per_gpu": B, "n": n, "m": m, "kernel": ctx.last_kernel, "parallelism": "batch-sharded x%d, no collective on the data path" % world,
                   "l2": "inputs are %.0f MB per step, larger than the 126 MB L2" % (8 * B * (n * n + n + m * n + 2 * m) / 1e6)},
        "admm_iters_per_s": all_iters / (ms_max / 1e3 / args.steps),
        "admm_iters_per_step": all_iters,
        "status_histogram": {api.STATUS_NAMES[k]: int((info["status"] == k).sum()) for k in np.unique(info["status"])},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                     "kernel_ms": 1e3 * sec_per_launch,
                     "note": "algorithmic bytes of SURVEY.md 8(d); the working set is register/shared-memory resident, so measured "
                             "DRAM traffic is far below this and frac can exceed 1 (see DESIGN.md)"},
        # the honest binding resource: fp64 FMA work against the chip's nominal fp64 vector rate
        "compute": {"flops_per_iteration": 4 * m * n + 2 * n * n, "achieved_tflops": (4 * m * n + 2 * n * n) * total_iters / sec_per_launch / 1e12,
                    "peak_tflops": 148 * 58.98 * 2 * 1.965e9 / 1e12, "unit": "TFLOP/s fp64",
                    "peak_source": "measured: 58.98 DFMA/clk/SM x 148 SMs x 2 x 1.965 GHz (tools/proto/proto_lat.cu on this pool's B200; "
                                   "nominal 64/clk; MEASURED_PEAKS.json has no fp64 entry)",
                    "frac": (4 * m * n + 2 * n * n) * total_iters / sec_per_launch / (148 * 58.98 * 2 * 1.965e9)},
        "clocks": clocks,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(d, args.settings, args.cpu_sample)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
