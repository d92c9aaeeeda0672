#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched QP-subproblem hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores (oracle)

A "step" is one pass of the hot path (fresh setup + ADMM solve, reference src/sqp.cpp:221-222)
over one batch of synthetic QPs: BASELINE.json configs[2], batch=8192 dense QPs n=64 m=128 fp64,
reference default settings. One process per GPU (torchrun for N>1).
  N = 1: the batch is resident in HBM; `extra` carries the other regimes and configs (S2, config 2, config 5) measured in the same run.
  N > 1: the north star's flow -- ONE batch of 8192 owned by rank 0 is solved by all N GPUs (strong scaling): every rank's solve
         kernel pulls its slice out of rank 0's HBM over NVLink and writes its results into rank 0's arrays (no split/gather step;
         `--transport nccl` runs the literal NCCL split -> solve -> NCCL gather instead). The weak-scaling figure (every rank its
         own 8192 QPs, no data-path traffic at all) rides along under `extra.weak`.

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
    ap.add_argument("--slice-iters", type=int, default=-1, help="time slicing of the register-tiled kernel: -1 automatic, 0 off, else iterations per slice")
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="auto (default): strong when N > 1 -- ONE batch on rank 0 solved by all ranks, the north star's split/gather "
                         "flow -- with the weak figure under extra.weak. weak: every rank solves its own batch and nothing else")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra records (S2, config 2, config 5, weak) of the default line")
    ap.add_argument("--transport", default="p2p", choices=["nccl", "p2p"],
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


WORKLOADS = {"config3": (8192, 64, 128), "config2": (1024, 32, 64), "config5": (2048, 256, 512)}


KERNEL_SOURCES = {"tile": "qp_tile.cu", "cluster": "qp_cluster.cu", "small": "qp_small.cu", "block": "qp_block.cu", "generic": "qp_generic.cu"}


def csrc_hash(kernel):
    """Short hash of the sources of one kernel (its .cu + the shared headers): ncu-derived constants under profiles/ (DRAM traffic per
    launch) are only quoted while the sources they were captured from are unchanged. `kernel`: a name as reported by
    sqpb200_last_kernel ("tile<64,128,4>x2", "cluster<4>/sparse x33", ...)."""
    import hashlib

    family = next((f for f in KERNEL_SOURCES if kernel.startswith(f)), None)
    files = ["qp_common.cuh", "qp_tile.cuh"] + ([KERNEL_SOURCES[family]] if family else sorted(KERNEL_SOURCES.values()))
    h = hashlib.sha256()
    for f in files:
        h.update(f.encode())
        h.update(open(os.path.join(ROOT, "sqp_solver_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def traffic_for(key):
    """(bytes per launch or None, note). profiles/traffic.json maps a workload key to {"bytes": dram bytes of one launch from an
    `ncu --set full` capture, "csrc": hash of the sources it was captured from}."""
    try:
        e = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(key)
    except Exception:
        e = None
    if not isinstance(e, dict):
        return None, "no ncu capture recorded for this workload"
    if e.get("csrc") != csrc_hash(key):
        return None, "stale: the ncu capture (%s) predates the current kernel sources" % e.get("capture", "?")
    return float(e["bytes"]), "ncu --set full capture %s (dram__bytes_read.sum + dram__bytes_write.sum of the solve kernel, one launch)" % e.get("capture", "?")


def algorithmic_flops(n, m, iters_executed, checks, factorizations):
    """Least-work exact formulation ('flops per iteration: canonical 4mn + 2n^2'): per ADMM iteration A twice and the
    n x n solve; per check A x, A^T y, P x; per (re)factorisation the SYRK A^T diag(rho) A (n^2 m) and chol(H) (n^3 / 3)."""
    return (iters_executed + checks) * (4 * m * n + 2 * n * n) + factorizations * (n * n * m + n * n * n // 3)


def sparse_algorithmic_bytes(n, m, nnz, iters_executed, checks, factorizations, count):
    """SURVEY.md 8(d) accounting carried over to compressed A (8-byte value + 4-byte index per stored entry):
    per iteration A twice + packed chol(H) twice, per check P and A once, per QP compulsory I/O, per factorisation P and A once."""
    a = 12 * nnz
    b_iter = 2 * a + 8 * n * (n + 1)
    b_check = 8 * n * n + a
    b_io = 8 * (n * n + n + 2 * m) + a + 8 * (n + m) + 16
    b_fact = 8 * n * n + a
    return count * b_io + iters_executed * b_iter + checks * b_check + factorizations * b_fact


def fp64_peak(ctx):
    """fp64 FMA peak of this GPU measured now (MEASURED_PEAKS.json carries HBM and bf16 only), with the SM clock it ran at."""
    sampler = ClockSampler(ctx.device)
    tf, sec, cnt = 0.0, 0.0, 0.0
    for _ in range(6):  # ~60 ms of load so that nvidia-smi sees the clock
        tf, sec, cnt = ctx.measure_fp64_peak()
    clk = sampler.stop()
    sms = ctx.device_query()["sm_count"]
    out = {"peak_tflops": tf, "unit": "TFLOP/s fp64", "how": "sqpb200_measure_fp64_peak: saturating DFMA kernel (8 independent chains per "
           "thread, 8 warps per scheduler), best of 4 launches, CUDA events", "sm_count": sms, "clocks": clk}
    if clk.get("sm_mhz"):
        out["dfma_per_clk_per_sm"] = cnt / sec / (clk["sm_mhz"] * 1e6) / sms
    return out


def counts_from_info(info, settings, api):
    executed = np.minimum(info["iter"], settings.max_iter).astype(np.int64)
    executed[info["status"] == api.NUMERICAL_ISSUES] = 0
    ct = settings.check_termination
    checks = int((executed // ct).sum()) if ct > 0 else 0
    return int(executed.sum()), checks


def rooflines(n, m, B, iters, checks, facts, sec, peak64, kernel, settings_name):
    """(roofline bound by what binds -- the fp64 pipe --, the SURVEY 8(d) algorithmic-HBM figure beside it)."""
    hbm_peak, hbm_src = peaks()
    bytes_per_launch = algorithmic_bytes(n, m, iters, checks, facts, B)
    flops = algorithmic_flops(n, m, iters, checks, facts)
    traffic, traffic_note = traffic_for("%s_%dx%d_b%d_%s" % (kernel, n, m, B, settings_name))
    ach = flops / sec / 1e12
    roof = {"bound": "fp64", "achieved": ach, "peak": peak64["peak_tflops"], "unit": "TFLOP/s", "frac": ach / peak64["peak_tflops"],
            "traffic": traffic, "traffic_note": traffic_note, "kernel_ms": 1e3 * sec, "algorithmic_flops_per_launch": flops,
            "units_per_launch": {"qps": B, "admm_iterations": iters, "checks": checks, "factorisations": facts},
            "peak_source": "measured in this run: " + peak64["how"],
            "note": "the working set (A, H^-1, P) is register/shared-memory resident for the whole solve, so HBM carries the compulsory "
                    "I/O only; the kernel is bound by the fp64 pipe and the shuffle/barrier latencies around it (DESIGN.md 4.1)"}
    if traffic:
        roof["dram_gbs"] = traffic / sec / 1e9
        roof["dram_frac_of_hbm_peak"] = traffic / sec / 1e9 / hbm_peak
        roof["traffic_over_compulsory_io"] = traffic / (B * (8 * (n * n + n + m * n + 2 * m) + 8 * (n + 2 * m) + 40))
    hbm = {"bound": "hbm", "achieved": bytes_per_launch / sec / 1e9, "peak": hbm_peak, "unit": "GB/s",
           "frac": bytes_per_launch / sec / 1e9 / hbm_peak, "peak_source": hbm_src, "algorithmic_bytes_per_launch": bytes_per_launch,
           "note": "SURVEY.md 8(d) canonical bytes (B_iter = 16mn + 8n(n+1), B_check = 8(n^2+mn), B_io, 8(n^2+mn) per factorisation) x the "
                   "units of one launch / kernel time: what an implementation streaming A and the factor from HBM every iteration would "
                   "need; > 1 because nothing is streamed"}
    return roof, hbm


def time_device(qb, dev, stream, steps, warmup, barrier, torch):
    """W untimed + K timed device-resident fused setup+solve launches. Returns (ms of the K steps on this rank, launches, last info,
    factorisations per launch, ADMM iterations of the last launch)."""
    def step():
        qb.setup_solve(dev["P"], dev["q"], dev["A"], dev["l"], dev["u"], stream=stream.cuda_stream)

    for _ in range(warmup):
        step()
    barrier()
    ru0 = int(qb.get(fields=("rho_updates",))["rho_updates"].astype(np.int64).sum())
    l0 = qb.ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = qb.ctx.launch_count - l0
    total_iters = qb.total_iters()
    info = qb.get(fields=("status", "iter", "rho_updates"))
    # rho_updates is cumulative over the calls on a batch object (qp.cpp:313): per launch = the timed region's increase / K
    facts = (int(info["rho_updates"].astype(np.int64).sum()) - ru0) // max(steps, 1)
    return ms, launches, info, facts, total_iters


def time_e2e(qb, d, B, n, m, steps, warmup, barrier, torch, info=None):
    """The same step through the reference-facing C-ABI call with HOST buffers: pinned host arrays -> sqpb200_qp_batch_setup_solve
    (HOST_PTRS: one persistent launch gated on the chunk-by-chunk H2D staging) -> sqpb200_qp_batch_get to pinned host. Wall clock."""
    pin = {k: torch.from_numpy(d[k][:B]).pin_memory() for k in ("P", "q", "A", "l", "u")}
    hp = {k: v.numpy() for k, v in pin.items()}
    ox = torch.empty(B, n, dtype=torch.float64).pin_memory()
    oy = torch.empty(B, max(m, 1), dtype=torch.float64).pin_memory()[:, :m]
    ost = torch.empty(B, dtype=torch.int32).pin_memory()
    oit = torch.empty(B, dtype=torch.int32).pin_memory()
    oy_np = np.ascontiguousarray(oy.numpy()) if m == 0 else oy.numpy()

    def step():
        qb.setup_solve(hp["P"], hp["q"], hp["A"], hp["l"], hp["u"], count=B)
        qb.get_into(count=B, x=ox.numpy(), y=oy_np, status=ost.numpy(), iter=oit.numpy())

    for _ in range(min(warmup, 2)):
        step()
    barrier()
    l0 = qb.ctx.launch_count
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if info is not None:
        assert (ost.numpy() == info["status"][:B]).all() and (oit.numpy() == info["iter"][:B]).all()
    h2d = 8 * B * (n * n + n + m * n + 2 * m)
    return dt, {"h2d_bytes_per_step": h2d, "d2h_bytes_per_step": B * (8 * (n + m) + 8), "launches": qb.ctx.launch_count - l0,
                "h2d_gbs_this_rank": h2d * steps / dt / 1e9}


E2E_HOW = ("pinned host buffers -> sqpb200_qp_batch_setup_solve(HOST_PTRS): one persistent launch whose work queue is gated on the "
           "chunk-by-chunk H2D staging -> sqpb200_qp_batch_get to pinned host; wall clock between device synchronisations, max over ranks")


def dense_record(ctx, api, torch, B, n, m, settings_name, steps, warmup, peak64, seed0=0, d=None, cpu=True, cpu_sample=0, e2e=True):
    """One single-GPU record of a dense workload (used for the headline at N = 1 and for the `extra` records)."""
    from sqp_solver_b200.synth import make_batch

    d = d if d is not None else make_batch(B, n, m, seed0=seed0)
    settings = api.default_settings(**settings_kwargs(settings_name))
    dev = {k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")}
    qb = api.QPBatch(ctx, B, n, m)
    qb.settings = settings
    stream = torch.cuda.current_stream()
    barrier = torch.cuda.synchronize
    sampler = ClockSampler(ctx.device)
    ms, launches, info, facts, total_iters = time_device(qb, dev, stream, steps, warmup, barrier, torch)
    kernel = ctx.last_kernel
    sec = ms / 1e3 / steps
    iters, checks = counts_from_info(info, settings, api)
    roof, hbm = rooflines(n, m, B, iters, checks, facts, sec, peak64, kernel, settings_name)
    rec = {"value": B / sec, "unit": UNIT, "ms_per_step": 1e3 * sec, "steps": steps, "warmup": warmup, "kernel": kernel,
           "workload": "batch=%d dense QPs n=%d m=%d fp64, settings %s, fresh setup+solve per step, inputs resident in HBM (%.0f MB per step: %s)"
                       % (B, n, m, settings_name, 8 * B * (n * n + n + m * n + 2 * m) / 1e6,
                          "larger than the 126 MB L2" if 8 * B * (n * n + n + m * n + 2 * m) > 126e6 else "L2-resident across steps"),
           "admm_iters_per_s": total_iters / sec, "admm_iters_per_step": total_iters, "factorisations_per_step": facts,
           "status_histogram": {api.STATUS_NAMES[k]: int((info["status"] == k).sum()) for k in np.unique(info["status"])},
           "gpu_launches": int(launches), "roofline": roof, "roofline_hbm_algorithmic": hbm}
    if e2e:
        dt, ex = time_e2e(qb, d, B, n, m, steps, warmup, barrier, torch, info)
        rec["e2e"] = dict({"value": B * steps / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / steps, "how": E2E_HOW}, **ex)
    rec["clocks"] = sampler.stop()
    if cpu:
        rec["cpu_baseline"] = cpu_baseline(d, settings_name, cpu_sample)
    qb.close()
    return rec, d


def run_strong(args, rank, world, local_rank, ctx, api, torch, dist):
    """N > 1, the north star's flow: ONE batch (8192 QPs) owned by rank 0 is solved by all N GPUs.
    p2p (default): every rank's solve kernel reads its slice from rank 0's HBM over NVLink (CUDA IPC peer mapping, TMA from peer
    memory) and its epilogue writes the results into rank 0's arrays -- ONE kernel per rank per step, no split/gather step.
    nccl: NCCL send/recv split -> local solve -> NCCL gather, all inside the timed region."""
    from sqp_solver_b200 import sharding
    from sqp_solver_b200.synth import make_batch

    B, n, m = args.batch, args.n, args.m
    settings = api.default_settings(**settings_kwargs(args.settings))
    lo, hi = sharding.shard_range(B, rank, world)
    # every rank generates its own full batch (weak extra); rank 0's (seed0 = 0) is THE batch of the strong flow
    want_weak = not args.no_extras
    d = make_batch(B, n, m, seed0=rank * B) if (rank == 0 or want_weak) else None
    qb = api.QPBatch(ctx, max(hi - lo, 1) if not want_weak else B, n, m)
    qb.settings = settings
    prob = {k: torch.from_numpy(d[k]).cuda() for k in sharding.PROBLEM_KEYS} if rank == 0 else None
    solve_local = sharding.gpu_solve_local(qb)
    stream = torch.cuda.current_stream()
    pb = None
    if args.transport == "p2p":
        pb = sharding.PeerBatch(ctx, n, m, B, root=0)
        pb.load(prob, stream=stream.cuda_stream)  # the batch is resident in rank 0's HBM before the timed region, as in the nccl flow

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def step():
        if pb is None:
            return sharding.solve_sharded(prob, n, m, B, solve_local, root=0)
        pb.solve(qb, stream=stream.cuda_stream)
        return None

    for _ in range(args.warmup):
        out = step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        out = step()
    ev1.record(stream)
    barrier()
    launches = ctx.launch_count - l0
    kernel = ctx.last_kernel
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
    ms = float(t.item())
    nl = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
    dist.all_reduce(nl, op=dist.ReduceOp.SUM)

    # ---- end to end: the batch arrives SHARDED in host memory (each rank holds its slice in its own pinned buffers, the way a
    # data-parallel loader delivers it) -> every rank's C-ABI call with HOST pointers -> results back in pinned host memory
    e2e = None
    if not args.no_e2e:
        cnt = hi - lo
        dsl = make_batch(cnt, n, m, seed0=lo) if rank != 0 else {k: d[k][lo:hi] for k in ("P", "q", "A", "l", "u")}
        dt, ex = time_e2e(qb, dsl, cnt, n, m, args.steps, args.warmup, barrier, torch)
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        agg = torch.tensor([float(ex["h2d_bytes_per_step"]), float(ex["d2h_bytes_per_step"]), float(ex["launches"])], dtype=torch.float64, device="cuda")
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        dt = float(tt.item())
        e2e = {"value": B * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(agg[0].item()), "d2h_bytes_per_step": int(agg[1].item()),
               "ms_per_step": 1e3 * dt / args.steps, "launches": int(agg[2].item()), "h2d_gbs_rank0": ex["h2d_gbs_this_rank"],
               "how": "the 8192-QP batch sharded over the ranks' own pinned host buffers (%d QPs each) -> " % cnt + E2E_HOW}

    # ---- weak scaling beside it: every rank its own 8192 QPs, nothing on the data path
    weak = None
    if want_weak:
        dev = {k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")} if rank != 0 else prob
        wms, wl, winfo, wfacts, wit = time_device(qb, dev, stream, args.steps, args.warmup, barrier, torch)
        wt = torch.tensor([wms], dtype=torch.float64, device="cuda")
        dist.all_reduce(wt, op=dist.ReduceOp.MAX)
        wagg = torch.tensor([float(wit)], dtype=torch.float64, device="cuda")
        dist.all_reduce(wagg, op=dist.ReduceOp.SUM)
        wms = float(wt.item())
        weak = {"scaling": "weak", "value": world * B * args.steps / (wms / 1e3), "unit": UNIT, "ms_per_step": wms / args.steps,
                "admm_iters_per_s": float(wagg.item()) / (wms / 1e3 / args.steps), "batch_per_gpu": B,
                "workload": "every rank solves its own batch of %d QPs (seed0 = rank x %d), no data-path traffic; device time, max over ranks" % (B, B)}
        if not args.no_e2e:
            dt, ex = time_e2e(qb, d, B, n, m, args.steps, args.warmup, barrier, torch, winfo)
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            weak["e2e"] = {"value": world * B * args.steps / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / args.steps,
                           "h2d_bytes_per_step": world * ex["h2d_bytes_per_step"], "d2h_bytes_per_step": world * ex["d2h_bytes_per_step"],
                           "h2d_gbs_rank0": ex["h2d_gbs_this_rank"]}
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        its = int(torch.clamp(out["iter"], max=settings.max_iter).sum().item())
        info = {"iter": out["iter"].cpu().numpy(), "status": out["status"].cpu().numpy()}
        iters, checks = counts_from_info(info, settings, api)
        # p2p: fresh instances every step (setup_solve_to), so rho_updates is per launch; nccl: cumulative over the calls on the object
        facts = int(out["rho_updates"].sum().item()) // (1 if pb is not None else args.warmup + args.steps)
        peak64 = fp64_peak(ctx)
        sec = ms / 1e3 / args.steps
        # the whole job's work over the whole job's time against N GPUs' peak
        roof, hbm = rooflines(n, m, B, iters, checks, facts, sec, dict(peak64, peak_tflops=world * peak64["peak_tflops"]), kernel, args.settings)
        roof["peak_source"] += " x %d GPUs" % world
        hbm["peak"] *= world
        hbm["frac"] /= world
        line = {"metric": METRIC, "value": B * args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": ("configs[2]: ONE batch=%d n=%d m=%d fp64 (settings %s) on rank 0, NCCL split -> solve -> NCCL gather (both inside "
                                        "the timed region)" if pb is None else "configs[2]: ONE batch=%d n=%d m=%d fp64 (settings %s) resident in rank "
                                        "0's HBM (826 MB, larger than L2); every rank's solve kernel reads its slice over NVLink (CUDA IPC peer mapping, "
                                        "TMA from peer memory) and its epilogue writes the results into rank 0's arrays: one kernel per rank per "
                                        "step, no split/gather step, no collective") % (B, n, m, args.settings),
                           "kernel": kernel, "transport": args.transport, "n": n, "m": m, "batch": B,
                           "parallelism": "batch-sharded x%d (strong: %d QPs per GPU)" % (world, hi - lo)},
                "admm_iters_per_s": its / (ms / 1e3 / args.steps), "admm_iters_per_step": its,
                "gpu_launches": int(nl.item()), "roofline": roof, "roofline_hbm_algorithmic": hbm, "compute_peak": peak64, "clocks": clocks}
        if e2e is not None:
            line["e2e"] = e2e
        extra = {}
        if weak is not None:
            extra["weak"] = weak
        if extra:
            line["extra"] = extra
        print(json.dumps(line))
    if pb is not None:
        pb.close()
    qb.close()


def config5_record(args, rank, world, local_rank, ctx, api, torch, dist, B, n, m, settings_name, steps, warmup, no_e2e, no_cpu, cpu_sample):
    """BASELINE.json configs[4]: sparse-A QPs (one CSR pattern for the batch) through sqpb200_qp_batch_setup_solve_sparse.
    Same JSON keys as the headline line; weak scaling (every rank its own batch). Returns the record on rank 0 (None elsewhere)."""
    from sqp_solver_b200.synth import densify, make_sparse_batch

    d = make_sparse_batch(B, n, m, density=args.density, seed0=rank * B)
    nnz = d["nnz"]
    settings = api.default_settings(**settings_kwargs(settings_name))
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

    for _ in range(warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ru0 = int(qb.get(fields=("rho_updates",))["rho_updates"].astype(np.int64).sum())
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step_device()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    total_iters = qb.total_iters()
    info0 = qb.get(fields=("status", "iter", "rho_updates"))
    facts = (int(info0["rho_updates"].astype(np.int64).sum()) - ru0) // max(steps, 1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    agg = torch.tensor([float(total_iters), float(B)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    ms_max, all_iters, all_qps = float(t.item()), float(agg[0].item()), float(agg[1].item())

    e2e = None
    if not no_e2e:
        pin = {k: torch.from_numpy(d[k]).pin_memory() for k in ("P", "q", "vals", "l", "u")}
        hp = {k: v.numpy() for k, v in pin.items()}
        ox = torch.empty(B, n, dtype=torch.float64).pin_memory()
        oy = torch.empty(B, m, dtype=torch.float64).pin_memory()
        ost = torch.empty(B, dtype=torch.int32).pin_memory()
        oit = torch.empty(B, dtype=torch.int32).pin_memory()

        def step_host():
            qb.setup_solve_sparse(hp["P"], hp["q"], hp["vals"], d["outer"], d["inner"], hp["l"], hp["u"], layout=api.SPARSE_CSR)
            qb.get_into(x=ox.numpy(), y=oy.numpy(), status=ost.numpy(), iter=oit.numpy())

        for _ in range(min(warmup, 2)):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_host()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        assert (ost.numpy() == info0["status"]).all() and (oit.numpy() == info0["iter"]).all()
        e2e = {"value": all_qps * steps / dt, "unit": UNIT, "h2d_bytes_per_step": B * (8 * (n * n + n + 2 * m + nnz)) + 4 * (m + 1 + nnz),
               "d2h_bytes_per_step": B * (8 * (n + m) + 8), "ms_per_step": 1e3 * dt / steps,
               "how": "pinned host buffers -> sqpb200_qp_batch_setup_solve_sparse(HOST_PTRS): one persistent launch whose work queue is gated "
                      "on the chunk-by-chunk H2D staging -> sqpb200_qp_batch_get to pinned host; wall clock, max over ranks"}
    clocks = sampler.stop() if sampler else None
    kernel = ctx.last_kernel
    qb.close()
    if rank != 0:
        return None
    iters, checks = counts_from_info(info0, settings, api)
    bytes_per_launch = sparse_algorithmic_bytes(n, m, nnz, iters, checks, facts, B)
    sec = ms / 1e3 / steps
    peak, peak_src = peaks()
    achieved = bytes_per_launch / sec / 1e9
    traffic, traffic_note = traffic_for("%s_%dx%d_b%d_%s" % (kernel, n, m, B, settings_name))
    line = {
        "metric": "QP-subproblems/sec (batch=%d, n=%d, m=%d, sparse A nnz=%d)" % (B, n, m, nnz), "value": all_qps * steps / (ms_max / 1e3),
        "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[4]: batch=%d sparse-A QPs n=%d m=%d fp64 per GPU, one CSR pattern for the batch (density %.3g + one "
                               "entry per row: nnz=%d), settings %s, fresh setup+solve per step" % (B, n, m, args.density, nnz, settings_name),
                   "batch_per_gpu": B, "n": n, "m": m, "nnz": nnz, "kernel": kernel,
                   "parallelism": "batch-sharded x%d, no collective on the data path" % world,
                   "l2": "inputs are %.0f MB per step, larger than the 126 MB L2" % (8 * B * (n * n + n + 2 * m + nnz) / 1e6)},
        "admm_iters_per_s": all_iters / (ms_max / 1e3 / steps), "admm_iters_per_step": all_iters,
        "factorisations_per_step": facts,
        "status_histogram": {api.STATUS_NAMES[k]: int((info0["status"] == k).sum()) for k in np.unique(info0["status"])},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_note": traffic_note, "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch, "kernel_ms": 1e3 * sec,
                     "note": "SURVEY.md 8(d) accounting with A compressed (12 B per stored entry). H^-1 lives in the shared memory of a "
                             "4-CTA cluster, so DRAM traffic is the compulsory I/O only; the kernel is bound by cluster barriers and the "
                             "serial pivot chain of the factorisation (DESIGN.md 4.4), not by HBM"},
        "clocks": clocks,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if not no_cpu:
        from oracle import qp_oracle as O

        O.build()
        cores = O.num_procs()
        sample = cpu_sample or min(B, 2 * cores)
        A = densify(d, 0, sample)
        st = O.default_settings(**settings_kwargs(settings_name))
        t0 = time.perf_counter()
        out = O.solve_batch(d["P"][:sample], d["q"][:sample], A, d["l"][:sample], d["u"][:sample], st, nthreads=cores)
        dt = time.perf_counter() - t0
        assert (out["status"] == info0["status"][:sample]).all(), "GPU and oracle disagree on the status of the sampled QPs"
        line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "first %d QPs of the same batch (densified: the reference's sparse variant is dead code), %s, gcc -O2 "
                                          "oracle restatement of src/qp.cpp, OpenMP over %d threads, %.2f s" % (sample, settings_name, out["threads"], dt),
                                "admm_iters_per_s": int(np.minimum(out["iter"], st.max_iter).sum()) / dt}
    return line


def config4_record(no_cpu, batch=4096, runs=3, device=0):
    """BASELINE.json configs[3]: batch=4096 constrained-Rosenbrock SQPs (BFGS Hessian, host outer loop, one batched GPU QP solve per
    outer iteration) through sqp::BatchSQP (sqp_solver_b200/host/tools/batch_sqp_bench.cpp), next to the CPU oracle's SQP
    (oracle/sqp_oracle.c: the restated src/sqp.cpp + src/qp.cpp) over all host cores on the same starting points."""
    from sqp_solver_b200 import build as B_

    tool = B_.TOOL
    if not os.path.exists(tool):
        tool = B_.build_tools()
    r = subprocess.run([tool, str(batch), str(runs), str(device)], capture_output=True, text=True, timeout=600)
    if r.returncode:
        return {"error": r.stderr[-500:]}
    g = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    rec = {"metric": "SQP solves/sec (batch=%d constrained-Rosenbrock, n=2, m=2, BFGS, max_iter 100)" % batch, "value": g["sqp_per_s"],
           "unit": "SQP/s", "seconds_per_batch": g["seconds"], "qp_launches": g["qp_launches"], "solved": g["solved"],
           "admm_iters_per_batch": g["qp_solver_iter_total"], "admm_iters_per_s": g["qp_solver_iter_total"] / g["seconds"],
           "kernel": "small<8> (thread-per-QP literal KKT kernel), one launch per outer iteration over the still-active instances",
           "e2e": {"value": g["sqp_per_s"], "unit": "SQP/s", "how": "the number above IS end to end: host buffers in, host results out, every outer "
                   "iteration (pinned staging, %d QP launches, host line search / BFGS between them); best of %d runs, wall clock" % (g["qp_launches"], runs)},
           "note": "host-bound by design: the north star keeps the SQP outer loop and BFGS on the host, so every outer iteration is a host "
                   "round trip (pack -> H2D -> kernel -> D2H -> line search)"}
    if not no_cpu:
        from oracle import qp_oracle as O, sqp_oracle as S

        O.build()
        i = np.arange(batch)
        x0 = np.stack([-0.6 + 1.2 * (i % 64) / 63.0 + 1e-3 * (i // 4096), -0.6 + 1.2 * ((i // 64) % 64) / 63.0], 1)
        cores = O.num_procs()
        t0 = time.perf_counter()
        c = S.solve_batch(S.CONSTRAINED_ROSENBROCK_2D, x0, np.zeros((batch, 2)), S.default_settings(), nthreads=cores)
        dt = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": batch / dt, "unit": "SQP/s", "cores": cores, "kind": "port",
                               "sample": "all %d instances, same starting points, oracle restatement of src/sqp.cpp + src/qp.cpp (gcc -O2), OpenMP "
                                         "schedule(dynamic) over %d threads, %.2f s" % (batch, c["threads"], dt),
                               "solved": int((c["status"] == S.SOLVED).sum()), "admm_iters_per_batch": int(c["qp_solver_iter"].sum())}
        # the QP subproblems are solved with the reference's own arithmetic (bit-identical kernel): the two runs must tell the same story
        rec["same_as_cpu"] = {"solved": rec["solved"] == rec["cpu_baseline"]["solved"],
                              "admm_iterations": rec["admm_iters_per_batch"] == rec["cpu_baseline"]["admm_iters_per_batch"]}
    return rec


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

    import torch
    import torch.distributed as dist

    from sqp_solver_b200 import api

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = api.Context(local_rank)  # raises when the CUDA library or the device is missing: no fallback
    try:
        if args.workload == "config5":
            line = config5_record(args, rank, world, local_rank, ctx, api, torch, dist, args.batch, args.n, args.m, args.settings,
                                  args.steps, args.warmup, args.no_e2e, args.no_cpu_baseline, args.cpu_sample)
            if rank == 0:
                print(json.dumps(line))
            return
        ctx.set_option(api.OPT_KERNEL, {"auto": 0, "generic": 1, "tile": 2}[args.kernel])
        ctx.set_option(api.OPT_TILE_WARPS, args.tile_warps)
        ctx.set_option(api.OPT_CTAS_PER_SM, args.ctas_per_sm)
        ctx.set_option(api.OPT_SLICE_ITERS, args.slice_iters)
        if world > 1 and args.scaling in ("auto", "strong"):
            run_strong(args, rank, world, local_rank, ctx, api, torch, dist)
            return
        run_weak(args, rank, world, local_rank, ctx, api, torch, dist)
    finally:
        if world > 1:
            dist.destroy_process_group()


def run_weak(args, rank, world, local_rank, ctx, api, torch, dist):
    """N = 1 (the driver's headline line) or --scaling weak: every rank solves its own batch, no data-path traffic."""
    from sqp_solver_b200.synth import make_batch

    B, n, m = args.batch, args.n, args.m
    d = make_batch(B, n, m, seed0=rank * B)
    settings = api.default_settings(**settings_kwargs(args.settings))
    dev = {k: torch.from_numpy(d[k]).cuda() for k in ("P", "q", "A", "l", "u")}
    qb = api.QPBatch(ctx, B, n, m)
    qb.settings = settings
    if args.fp32:
        qb.set_precision(True)
        args.no_cpu_baseline = True
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak64 = fp64_peak(ctx) if rank == 0 else None
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, launches, info, facts, total_iters = time_device(qb, dev, stream, args.steps, args.warmup, barrier, torch)
    kernel = ctx.last_kernel
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    agg = torch.tensor([float(total_iters), float(B)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    all_iters, all_qps = float(agg[0].item()), float(agg[1].item())

    e2e = None
    if not args.no_e2e:
        dt, ex = time_e2e(qb, d, B, n, m, args.steps, args.warmup, barrier, torch, info)
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": all_qps * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": world * ex["h2d_bytes_per_step"],
               "d2h_bytes_per_step": world * ex["d2h_bytes_per_step"], "ms_per_step": 1e3 * dt / args.steps, "launches": ex["launches"],
               "h2d_gbs_rank0": ex["h2d_gbs_this_rank"], "how": E2E_HOW}
    clocks = sampler.stop() if sampler else None
    qb.close()
    if rank != 0:
        return

    iters, checks = counts_from_info(info, settings, api)
    sec_per_launch = ms / 1e3 / args.steps  # rank 0's own kernel time
    roof, hbm = rooflines(n, m, B, iters, checks, facts, sec_per_launch, peak64, kernel, args.settings)
    in_bytes = 8 * B * (n * n + n + m * n + 2 * m)
    line = {
        "metric": METRIC if (B, n, m) == WORKLOADS["config3"] else "QP-subproblems/sec (batch=%d, n=%d, m=%d)" % (B, n, m),
        "value": all_qps * args.steps / (ms_max / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.fp32 else "f64", "data": "synthetic",
        "config": {"workload": "configs[%d]: batch=%%d dense QPs n=%%d m=%%d fp64 per GPU, settings %%s (%%s), fresh setup+solve per step" % (1 if args.workload == "config2" else 2)
                               % (B, n, m, args.settings, "reference defaults qp.hpp:38-53" if args.settings == "S1" else
                                  "alpha=1.6 adaptive_rho"),
                   "batch_per_gpu": B, "n": n, "m": m, "kernel": kernel, "parallelism": "batch-sharded x%d, no collective on the data path" % world,
                   "l2": "inputs are %.0f MB per step, %s" % (in_bytes / 1e6, "larger than the 126 MB L2" if in_bytes > 126e6 else
                                                             "SMALLER than the 126 MB L2: they stay L2-resident across steps (no flush between steps)")},
        "admm_iters_per_s": all_iters / (ms_max / 1e3 / args.steps),
        "admm_iters_per_step": all_iters, "factorisations_per_step": facts,
        "status_histogram": {api.STATUS_NAMES[k]: int((info["status"] == k).sum()) for k in np.unique(info["status"])},
        "gpu_launches": int(launches),
        "roofline": roof, "roofline_hbm_algorithmic": hbm, "compute_peak": peak64,
        "clocks": clocks,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(d, args.settings, args.cpu_sample)
    # ---- the other regimes and configs, measured in the same run (N = 1 only; each a few seconds)
    if world == 1 and not args.no_extras and not args.fp32 and args.workload == "config3" and args.settings == "S1" and args.kernel == "auto":
        extra = {}
        xs, xw = min(args.steps, 10), 3
        try:
            extra["config3_S2"], _ = dense_record(ctx, api, torch, B, n, m, "S2", xs, xw, peak64, d=d, cpu=not args.no_cpu_baseline)
            extra["config3_S2"]["note"] = "the regime the reference's own caller uses (SQP constructor: alpha 1.6 + adaptive rho)"
            b2, n2, m2 = WORKLOADS["config2"]
            extra["config2_S1"], _ = dense_record(ctx, api, torch, b2, n2, m2, "S1", xs, xw, peak64, cpu=not args.no_cpu_baseline)
            extra["config2_S1"]["note"] = "BASELINE configs[1]: batch=1024 dense QPs n=32 m=64, one warp per QP"
            b5, n5, m5 = WORKLOADS["config5"]
            extra["config5_S2"] = config5_record(args, 0, 1, local_rank, ctx, api, torch, dist, b5, n5, m5, "S2", min(xs, 5), 2,
                                                 args.no_e2e, args.no_cpu_baseline, 0)
            extra["config4_sqp"] = config4_record(args.no_cpu_baseline, device=local_rank)
        except Exception as e:  # the extras never take the headline line down
            extra["error"] = repr(e)
        line["extra"] = extra
    print(json.dumps(line))


if __name__ == "__main__":
    main()
