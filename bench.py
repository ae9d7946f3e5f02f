#!/usr/bin/env python
"""bench.py — headline benchmark of the batched LQ hot path.

Metric (BASELINE.json): Riccati + rollout LQ solves/s on the legged-robot shape (nx = 24, nu = 24, N = 100, ILQR, LINE_SEARCH,
reduced Riccati form, DIAGONAL_SHIFT 1e-5). One "step" = one pass of the hot path (backward sweep + controller + one alpha = 1
rollout) over the whole batch of seeded synthetic problems.

Scaling: STRONG. The global batch (BASELINE.json config 5: 16384 legged problems) is split over the ranks by problem index
(`shard_bounds`, contiguous blocks, no collective on the data path): at N = 8 every GPU owns 2048 problems. `value` = global batch /
max-over-ranks device time. The weak-scaling figure (the full 16384 problems on EVERY GPU) rides along as `weak` when N > 1, and the
other BASELINE configs (ballbot 65536, quadrotor SLQ 32768, manipulator 16384; 3 steps each, sharded the same way) as `workloads`.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA, through the C ABI)
  python bench.py --impl reference [...]                       the reference's CPU path (oracle port: Eigen/Boost are absent, the
                                                               reference itself cannot be compiled here) on all host cores
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nx, nu, nc, algorithm, eps, BASELINE global batch)
    "legged": (24, 24, 0, 0, 1e-5, 16384),
    "ballbot": (10, 3, 0, 0, 1e-3, 65536),
    "quadrotor_slq": (12, 4, 0, 1, 1e-3, 32768),
    "manipulator": (9, 9, 3, 0, 1e-3, 16384),
    "cartpole": (4, 1, 0, 0, 1e-6, 1),
    # not BASELINE configs: what the reference's own legged example runs on top of config 5 (contact constraints,
    # LeggedRobotInterface.cpp:186-190; SLQ, ocs2_legged_robot/config/mpc/task.info:85)
    "legged_constrained": (24, 24, 12, 0, 1e-5, 16384),
    "legged_slq": (24, 24, 0, 1, 1e-5, 8192),
}
TABLE_WORKLOADS = ("ballbot", "quadrotor_slq", "manipulator", "legged_constrained", "legged_slq")
N_STAGES = 100
DT = 0.01
METRIC = "LQ solves/s (Riccati backward sweep + LQ rollout)"
FP64_PEAK_TFLOPS_MEASURED = 37.1  # profiles/r01_fp64_peak_microbench.log (DMMA m8n8k4, this pool's B200)
FP64_PEAK_SOURCE = ("profiles/r01_fp64_peak_microbench.log: FP64 DMMA (mma.sync m8n8k4.f64, the tensor pipe's FP64 sub-pipe) measured on this pool's "
                    "B200 = DFMA peak; MEASURED_PEAKS.json holds no FP64 entry, its bf16 figure does not apply to an f64 path")


def workload_label(name):
    """The workload string both arms print (config.workload): what is solved, not how."""
    n, m, nc, alg, eps, _ = WORKLOADS[name]
    if alg == 0:
        return f"{name} ILQR nx={n} nu={m} nc={nc} N={N_STAGES} LINE_SEARCH reduced DIAGONAL_SHIFT {eps}"
    return f"{name} SLQ-RK4 nx={n} nu={m} N={N_STAGES} timeStep={DT}"


def config_of(workload, global_batch, shards):
    """The `config` object of a bench line: the workload, its global batch, how it is split and why no L2 flush is needed. Both arms
    (ours and --impl reference) print exactly this object for the same command line."""
    n, m, nc, alg, _, _ = WORKLOADS[workload]
    bytes_solve, _ = algorithmic_per_solve(n, m, nc, N_STAGES, alg)
    per_shard = (global_batch + shards - 1) // shards
    return {"workload": workload_label(workload), "global_batch": global_batch,
            "sharding": "by problem index (contiguous blocks), no collective on the data path",
            "l2": "inputs larger than L2 (per-GPU LQ data %.1f GB >> 126 MB), no flush needed" % (bytes_solve * per_shard / 1e9)}


def algorithmic_per_solve(n, m, nc, N, alg=0):
    """compulsory bytes (each input read once, each output written once) and flops of one solve
    (ILQR sweep + discrete rollout, or SLQ-RK4 flow map with one step per interval + continuous rollout)."""
    bytes_in = 8 * (2 * n * n + 2 * n * m + m * m + 2 * n + m + 1 + nc * (n + m + 1))
    bytes_out = 8 * (n * m + m + n * n + n + 1 + n + m)
    if alg == 1:
        p = m - nc
        flops = 4 * (2 * n**3 + 4 * n * n * p + 7 * n * n + 4 * n * p + 3 * (3 * n * n + 2 * n * p + 3 * n + 2 * p + 1)) + 10 * (n * (n + 1) / 2 + n + 1)
        flops += (11.0 / 3.0) * m**3 + 3 * n * m * m + 4 * n * n * m  # per-node projection + controller
        flops += 4 * (2 * n * n + 4 * n * m)                          # continuous rollout, 4 RK4 stages per interval
        return N * (bytes_in + bytes_out), N * flops
    flops = 4 * n**3 + 6 * n * n * m + 4 * n * m * m + (2.0 / 3.0) * m**3 + 4 * n * n + 4 * n * m + 2 * m * m + 2 * n * n + 4 * n * m
    if nc:
        flops += 6 * m * m * nc + 2 * nc * nc * (m - nc / 3.0) + 2 * nc * n * m + 2 * nc * m
    return N * (bytes_in + bytes_out), N * flops


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def dram_traffic_entry(workload, kernel):
    """DRAM bytes per solve of the dominant kernel from the committed ncu capture (profiles/dram_traffic.json), with the capture's
    summary file and its hash so that the number can be traced; None when no capture of this kernel is committed."""
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if not os.path.exists(tpath):
        return None
    with open(tpath) as f:
        tj = json.load(f)
    ent = tj.get(f"{workload}:{kernel}")
    if not ent:
        return None
    out = {"dram_bytes_per_solve": ent["dram_bytes_per_solve"], "source": ent.get("source")}
    src = ent.get("file")
    if src and os.path.exists(os.path.join(ROOT, src)):
        with open(os.path.join(ROOT, src), "rb") as f:
            out["file"], out["sha256_16"] = src, hashlib.sha256(f.read()).hexdigest()[:16]
    return out


def roofline_of(workload, kernel, batch_local, kernel_s, peaks, which):
    """Roofline of one launch of the dominant kernel on one GPU: algorithmic flops against the measured FP64 peak, algorithmic bytes and
    (when a capture is committed) the DRAM bytes ncu counted against the measured HBM peak. `bound` names the binding roofline of the
    ALGORITHMIC work; `vs_traffic_roofline` is the fraction of max(t_fp64, t_dram_traffic)."""
    n, m, nc, alg, _, _ = WORKLOADS[workload]
    bytes_solve, flops_solve = algorithmic_per_solve(n, m, nc, N_STAGES, alg)
    fp64 = {"achieved": flops_solve * batch_local / kernel_s / 1e12, "peak": FP64_PEAK_TFLOPS_MEASURED, "unit": "TFLOP/s",
            "algorithmic_flops_per_solve": flops_solve}
    fp64["frac"] = fp64["achieved"] / fp64["peak"]
    hbm = {"achieved": bytes_solve * batch_local / kernel_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "algorithmic_bytes_per_solve": bytes_solve,
           "peak_source": f"MEASURED_PEAKS.json ({which})"}
    hbm["frac"] = hbm["achieved"] / hbm["peak"]
    t_hbm, t_fp64 = bytes_solve / (peaks["hbm_gbs"] * 1e9), flops_solve / (FP64_PEAK_TFLOPS_MEASURED * 1e12)
    tr = dram_traffic_entry(workload, kernel)
    traffic = None
    extra = {}
    if tr:
        traffic = tr["dram_bytes_per_solve"] * batch_local
        t_dram = tr["dram_bytes_per_solve"] / (peaks["hbm_gbs"] * 1e9)
        extra = {"traffic_source": tr,
                 "hbm_on_dram_traffic": {"achieved": traffic / kernel_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                         "frac": traffic / kernel_s / 1e9 / peaks["hbm_gbs"]},
                 "vs_traffic_roofline": {"roofline_solves_per_s": 1.0 / max(t_fp64, t_dram), "frac": (batch_local / kernel_s) * max(t_fp64, t_dram),
                                         "bound": "tensor" if t_fp64 >= t_dram else "hbm"}}
    if t_fp64 >= t_hbm:
        return {"bound": "tensor", **fp64, "peak_source": FP64_PEAK_SOURCE, "traffic": traffic, "kernel": kernel, "kernel_ms": kernel_s * 1e3,
                "units_per_launch": batch_local, "hbm": hbm, **extra}
    return {"bound": "hbm", **hbm, "traffic": traffic, "kernel": kernel, "kernel_ms": kernel_s * 1e3, "units_per_launch": batch_local,
            "fp64": {**fp64, "peak_source": FP64_PEAK_SOURCE}, **extra}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power), "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """The reference's own CPU implementation of the path (restated: oracle port), all host threads, bounded sample per step."""
    from oracle import oracle as orc

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, m, nc, alg, eps, default_batch = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    threads = args.cpu_threads or cores
    st = orc.make_settings(algorithm=alg, reduced_form=True, hessian_multiple=eps, time_step=DT)
    sample = args.cpu_sample or max(threads * 64, 256)  # ~0.4 s of CPU work per step on 16 threads: long enough to keep every thread busy
    for w in range(args.warmup):
        orc.baseline_run(st, 1, w * sample, min(sample, threads), n, m, nc, N_STAGES, DT, threads)
    total_s = 0.0
    for k in range(args.steps):
        secs, _ = orc.baseline_run(st, 1, k * sample, sample, n, m, nc, N_STAGES, DT, threads)
        total_s += secs
    value = args.steps * sample / total_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args.workload, args.batch or default_batch, max(1, args.gpus)),
        "config_detail": {"sample_problems_per_step": sample,
                          "note": "each step solves a bounded sample of the workload's seeded problem family on the host cores; solves/s does not depend on the sample size"},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} problems per step x {args.steps} steps of the same seeded family (oracle/lq_oracle.cpp, one problem per task)"},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(device_index):
    """One process per GPU: run on the CPUs of the NUMA node the GPU hangs off, so that the pinned host buffers of the end-to-end path
    (first-touch allocation) and the copy engines' reads stay on that socket. Returns the node, or None when it cannot be determined."""
    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:  # noqa: BLE001 - placement is an optimisation, never a requirement
        pass
    return None


class Ctx:
    """rank / device / collective plumbing shared by the measurements (torch.distributed only for the barrier and the max over ranks)."""

    def __init__(self, torch):
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None

    def init(self):
        torch = self.torch
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            import torch.distributed as dist_mod

            self.dist = dist_mod
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            self.dist.init_process_group("nccl", rank=self.rank, world_size=self.world, device_id=torch.device("cuda", self.local_rank))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def timed_steps(ctx, solver, steps, warmup, sampler=None):
    """`warmup` untimed steps, then exactly `steps` steps between barrier + synchronize on both sides, CUDA events on the library's compute
    stream; returns (max-over-ranks total ms, mean ms of the dominant kernel per step on this rank, launches in the region, clocks)."""
    torch = ctx.torch
    stream = torch.cuda.ExternalStream(solver.compute_stream, device=torch.device("cuda", ctx.local_rank))
    for _ in range(warmup):
        solver.solve(1.0)
    solver.sync()
    split = solver.kernel_variant not in ("ilqr_wpp_kernel", "ilqr_rpl_kernel")  # backward and rollout are separate launches there
    if sampler is not None:
        sampler.start()
        time.sleep(0.3)
    ctx.barrier()
    launches0 = solver.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
    t_wall0 = time.time()
    ev[0].record(stream)
    for k in range(steps):
        if split:
            solver.solveSequentialRiccatiEquations()
            ev[2 * k + 1].record(stream)
            solver.rolloutTrajectory((1.0,))
        else:
            solver.solve(1.0)
            ev[2 * k + 1].record(stream)
        ev[2 * k + 2].record(stream)
    solver.sync()
    ctx.barrier()
    t_wall1 = time.time()
    launches = solver.launch_count - launches0
    total_ms = ctx.max_over_ranks(ev[0].elapsed_time(ev[2 * steps]))
    sweep_ms = sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(steps)) / steps
    step_ms = ev[0].elapsed_time(ev[2 * steps]) / steps
    clocks = sampler.stop(t_wall0, t_wall1) if sampler is not None else None
    return total_ms, (step_ms if split else sweep_ms), launches, clocks


def measure_workload(ctx, o2, np, name, global_batch, steps, warmup, peaks, which, strong=True, sampler=None, keep=False):
    """One workload at its global batch: every rank owns its shard_bounds block (strong) or the whole batch (weak)."""
    from ocs2_b200.sharding import shard_bounds

    n, m, nc, alg, eps, _ = WORKLOADS[name]
    if strong:
        first, count = shard_bounds(global_batch, ctx.world, ctx.rank)
    else:
        first, count = ctx.rank * global_batch, global_batch
    st = o2.Settings(algorithm=alg, hessianCorrectionMultiple=eps, timeStep=DT)
    solver = o2.BatchedLqSolver(st, n, m, N_STAGES, count, nc_max=nc, device=ctx.local_rank)
    solver.generate_synthetic(seed=1, first_problem_index=first, dt=DT)
    solver.sync()
    total_ms, kernel_ms, launches, clocks = timed_steps(ctx, solver, steps, warmup, sampler)
    ms_per_step = total_ms / steps
    processed = global_batch if strong else ctx.world * global_batch
    # status check outside the timed region: the problems of the last step finished clean
    sol = solver.download(problem_begin=0, problem_count=min(count, 64))
    assert (sol.status == 0).all() and np.isfinite(sol.x).all(), "solver reported a failure status"
    kernel = solver.kernel_variant
    res = {"value": processed / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "kernel": kernel, "launches": launches, "clocks": clocks,
           "batch_local": count, "global_batch": processed, "first": first,
           "roofline": roofline_of(name, kernel, count, kernel_ms * 1e-3, peaks, which), "settings": st}
    if keep:
        res["solver"] = solver
    else:
        solver.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="legged", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="GLOBAL batch, split over the GPUs (default: the workload's BASELINE batch)")
    ap.add_argument("--weak", action="store_true", help="weak scaling as the headline: --batch problems on EVERY GPU")
    ap.add_argument("--e2e-batch", type=int, default=4096, help="GLOBAL problems per end-to-end step (host buffers, H2D/D2H timed)")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the table of the other BASELINE configs")
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaling sub-measurement at N > 1")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch

    import ocs2_b200 as o2
    from ocs2_b200.sharding import shard_bounds

    ctx = Ctx(torch)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (ocs2_b200 has no CPU fallback; use --impl reference for the CPU arm)")
    numa = bind_to_gpu_numa_node(ctx.local_rank)  # before any pinned allocation: first touch places the host buffers next to the GPU
    ctx.init()
    peaks, which = measured_peaks()
    n, m, nc, alg, eps, default_batch = WORKLOADS[args.workload]
    global_batch = args.batch or default_batch

    # ---- headline: the workload at its global batch, strong scaling ----
    head = measure_workload(ctx, o2, np, args.workload, global_batch, args.steps, args.warmup, peaks, which, strong=not args.weak,
                            sampler=ClockSampler(ctx.local_rank), keep=True)
    solver = head["solver"]

    # ---- end to end through the C ABI with host buffers (H2D + D2H inside the timed region), same problem family ----
    e2e = None
    if not args.no_e2e:
        ge = min(args.e2e_batch, global_batch)
        _, eb = shard_bounds(ge, ctx.world, ctx.rank)
        eb = min(eb, head["batch_local"])
        e2e = run_e2e(ctx, o2, np, solver, head["settings"], args.workload, eb, ge, args)
    solver.close()

    # ---- weak scaling as a sub-field (N > 1): the whole global batch on every GPU ----
    weak = None
    if ctx.world > 1 and not args.weak and not args.no_weak:
        w = measure_workload(ctx, o2, np, args.workload, global_batch, min(args.steps, 5), 3, peaks, which, strong=False)
        weak = {"value": w["value"], "unit": "solves/s", "ms_per_step": w["ms_per_step"], "batch_per_gpu": w["batch_local"], "global_batch": w["global_batch"],
                "steps": min(args.steps, 5), "roofline_frac": w["roofline"]["frac"]}

    # ---- the other BASELINE configs, 3 steps each, sharded the same way ----
    table = None
    if not args.no_workloads and args.workload == "legged":
        table = {}
        for name in TABLE_WORKLOADS:
            r = measure_workload(ctx, o2, np, name, WORKLOADS[name][5], 3, 3, peaks, which, strong=True)
            rf = r["roofline"]
            table[name] = {"workload": workload_label(name), "global_batch": r["global_batch"], "batch_per_gpu": r["batch_local"], "value": r["value"],
                           "unit": "solves/s", "ms_per_step": r["ms_per_step"], "steps": 3, "kernel": r["kernel"], "bound": rf["bound"],
                           "frac": rf["frac"], "achieved": rf["achieved"], "peak": rf["peak"], "roofline_unit": rf["unit"], "traffic": rf["traffic"],
                           "vs_traffic_roofline": rf.get("vs_traffic_roofline")}

    cpu_baseline = None
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc

        cores = os.cpu_count() or 1
        threads = args.cpu_threads or cores
        ost = orc.make_settings(algorithm=alg, reduced_form=True, hessian_multiple=eps, time_step=DT)
        orc.baseline_run(ost, 1, 0, threads, n, m, nc, N_STAGES, DT, threads)  # warm-up
        sample = args.cpu_sample or max(threads * 8, 64)
        secs, _ = orc.baseline_run(ost, 1, 0, sample, n, m, nc, N_STAGES, DT, threads)
        if secs < 5.0:  # grow to a ~10 s sample
            sample = int(sample * min(64.0, 10.0 / max(secs, 1e-3)))
            secs, _ = orc.baseline_run(ost, 1, 0, sample, n, m, nc, N_STAGES, DT, threads)
        cpu_baseline = {"value": sample / secs, "unit": "solves/s", "cores": threads, "kind": "port",
                        "sample": f"{sample} problems of the same seeded family, one problem per task on {threads} threads (oracle/lq_oracle.cpp; "
                                  "the reference itself needs Eigen3/Boost which are absent)"}

    if ctx.rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": "solves/s", "n_gpus": ctx.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak" if args.weak else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # `config` is the same object in both arms (what is solved); how this arm solves it goes to `config_detail`
            "config": config_of(args.workload, head["global_batch"], ctx.world),
            "config_detail": {"batch_per_gpu": head["batch_local"], "host_numa_node": numa, "kernel": head["kernel"]},
            "clocks": head["clocks"], "e2e": e2e, "gpu_launches": head["launches"], "roofline": head["roofline"], "cpu_baseline": cpu_baseline,
            "weak": weak, "workloads": table,
        }
        print(json.dumps(line), flush=True)
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()


class _DevArray:
    """Device memory of the library as a __cuda_array_interface__ object (so torch can read the resident LQ records back to the host)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


def run_e2e(ctx, o2, np, solver, st, workload, eb, global_eb, args):
    """Same metric through o2c_solve_host: pinned HOST SoA buffers in, HOST buffers out, chunked H2D / compute / D2H pipeline. The host
    data are the SAME seeded problems the device-resident measurement ran on (the first `eb` problems of this rank's shard, read back
    from the library's records before the timed region)."""
    import ctypes as C

    from ocs2_b200 import lib as _l

    torch = ctx.torch
    n, m, nc, alg, _, _ = WORKLOADS[workload]
    nodes = N_STAGES + 1 if alg == 1 else N_STAGES
    N = N_STAGES

    def pinned(shape):
        return torch.empty(shape, dtype=torch.float64).pin_memory()

    # resident records -> pinned host SoA arrays (column-major blocks, [problem][node][block] per field)
    dv = solver.device_lq_view()
    solver.sync()
    host = {}

    def pull(name, fld, block, nn):
        ps, ns = int(fld.problem_stride), int(fld.node_stride) if nn > 1 else block
        length = (eb - 1) * ps + (nn - 1) * ns + block
        flat = torch.as_tensor(_DevArray(int(fld.ptr), length), device=torch.device("cuda", ctx.local_rank))
        dst = pinned((eb, nn, block))
        dst.copy_(torch.as_strided(flat, (eb, nn, block), (ps, ns, 1)))
        host[name] = dst.numpy()

    pull("A", dv.A, n * n, nodes), pull("B", dv.B, n * m, nodes), pull("Q", dv.Q, n * n, nodes), pull("P", dv.P, m * n, nodes)
    pull("R", dv.R, m * m, nodes), pull("Hv", dv.Hv, n, nodes), pull("q", dv.q, n, nodes), pull("r", dv.r, m, nodes), pull("c", dv.c, 1, nodes)
    if nc:
        pull("C", dv.C, nc * n, nodes), pull("D", dv.D, nc * m, nodes), pull("e", dv.e, nc, nodes)
    pull("Qf", dv.Qf, n * n, 1), pull("qf", dv.qf, n, 1), pull("cf", dv.cf, 1, 1), pull("x0", dv.x0, n, 1)
    torch.cuda.synchronize()
    # packed upper triangles of the symmetric cost Hessians (O2C_LQ_SYMMETRIC_PACKED), produced outside the timed region like the
    # dense blocks are: a caller that assembles its LQ data writes whichever form the view declares
    packed = {}
    for name, k in (("Q", n), ("R", m), ("Qf", n)):
        full = host[name].reshape(host[name].shape[:-1] + (k, k))
        buf = pinned(full.shape[:-2] + (k * (k + 1) // 2,))
        buf.numpy()[:] = o2.pack_upper(full)  # blocks are symmetric: the row- / column-major reading of the block does not matter
        packed[name] = buf.numpy()

    def fld(a, block, nn):
        return _l.Field(a.ctypes.data, nn * block, block)

    def lq_view(sym):
        lv = _l.LqView()
        src = packed if sym else host
        lv.A, lv.B, lv.P = fld(host["A"], n * n, nodes), fld(host["B"], n * m, nodes), fld(host["P"], m * n, nodes)
        lv.Q = fld(src["Q"], n * (n + 1) // 2 if sym else n * n, nodes)
        lv.R = fld(src["R"], m * (m + 1) // 2 if sym else m * m, nodes)
        lv.Hv, lv.q, lv.r, lv.c = fld(host["Hv"], n, nodes), fld(host["q"], n, nodes), fld(host["r"], m, nodes), fld(host["c"], 1, nodes)
        if nc:
            lv.C, lv.D, lv.e = fld(host["C"], nc * n, nodes), fld(host["D"], nc * m, nodes), fld(host["e"], nc, nodes)
        lv.Qf = fld(src["Qf"], n * (n + 1) // 2 if sym else n * n, 1)
        lv.qf, lv.cf, lv.x0 = fld(host["qf"], n, 1), fld(host["cf"], 1, 1), fld(host["x0"], n, 1)
        lv.flags = _l.LQ_SYMMETRIC_PACKED if sym else 0
        nbytes = sum(v.nbytes for k_, v in host.items() if not (sym and k_ in packed)) + (sum(v.nbytes for v in packed.values()) if sym else 0)
        return lv, nbytes

    on = solver.rollout_num_nodes
    out = {"K": pinned((eb, N + 1, n, m)).numpy(), "dbias": pinned((eb, N + 1, m)).numpy(), "bias": pinned((eb, N + 1, m)).numpy(),
           "Sm": pinned((eb, N + 1, n, n)).numpy(), "Sv": pinned((eb, N + 1, n)).numpy(), "s": pinned((eb, N + 1)).numpy(),
           "x": pinned((eb, on, n)).numpy(), "u": pinned((eb, on, m)).numpy()}
    status = np.zeros(eb, dtype=np.int32)

    def sol_view(fields):
        sv = _l.SolutionView()
        blocks = {"K": (m * n, N + 1), "dbias": (m, N + 1), "bias": (m, N + 1), "Sm": (n * n, N + 1), "Sv": (n, N + 1), "s": (1, N + 1), "x": (n, on), "u": (m, on)}
        for name in fields:
            setattr(sv, name, fld(out[name], *blocks[name]))
        sv.status = status.ctypes.data
        return sv, sum(out[name].nbytes for name in fields) + status.nbytes

    lib = solver._lib
    ALL = ("K", "dbias", "bias", "Sm", "Sv", "s", "x", "u")
    POLICY = ("K", "dbias", "bias", "x", "u")

    def leg(sym, fields, steps):
        lv, h2d = lq_view(sym)
        sv, d2h = sol_view(fields)

        def step():
            _l.check(lib.o2c_solve_host(solver.handle, C.byref(lv), C.byref(sv), 1.0, eb, 0))  # returns after the last D2H completed

        for _ in range(2):
            step()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        secs = ctx.max_over_ranks(time.perf_counter() - t0)
        assert (status == 0).all() and np.isfinite(out["x"]).all()
        return {"value": global_eb * steps / secs, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "problems_per_step": global_eb, "problems_per_step_per_gpu": eb, "ms_per_step": 1e3 * secs / steps,
                "h2d_gb_per_s_per_gpu": h2d * steps / secs / 1e9, "d2h_gb_per_s_per_gpu": d2h * steps / secs / 1e9}

    main_leg = leg(True, ALL, args.steps)
    dense = leg(False, ALL, min(args.steps, 3))
    slim = leg(True, POLICY, min(args.steps, 3))

    # what the host link gives with every rank copying at once: pinned H2D alone and with D2H running against it (1 GiB buffers)
    nb = 1 << 27
    h_in, h_out = pinned((nb,)), pinned((nb,))
    d_in = torch.empty(nb, dtype=torch.float64, device="cuda")
    d_out = torch.ones(nb, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def link(h2d, d2h, reps=3):
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return nb * 8 * reps / ctx.max_over_ranks(time.perf_counter() - t0) / 1e9

    link(True, True, 1)
    h2d_alone, h2d_duplex = link(True, False), link(True, True)
    ceiling = h2d_duplex * 1e9 / (main_leg["h2d_bytes_per_step"] / eb) * ctx.world  # solves/s if the H2D ran at the duplex link rate
    main_leg.update({
        "input_format": "O2C_LQ_SYMMETRIC_PACKED (Q, R, Qf as packed upper triangles), every output field requested",
        "same_problems_as_value": True,
        "dense_input": {k: dense[k] for k in ("value", "h2d_bytes_per_step", "d2h_bytes_per_step", "ms_per_step", "h2d_gb_per_s_per_gpu")},
        "policy_only_output": {k: slim[k] for k in ("value", "h2d_bytes_per_step", "d2h_bytes_per_step", "ms_per_step", "h2d_gb_per_s_per_gpu")},
        "host_link": {"h2d_alone_gb_per_s_per_gpu": h2d_alone, "h2d_with_d2h_gb_per_s_per_gpu": h2d_duplex, "ranks_copying_at_once": ctx.world,
                      "how": "1 GiB pinned buffers, all ranks at once after a barrier, max-over-ranks time"},
        "h2d_ceiling_solves_per_s": ceiling, "frac_of_h2d_ceiling": main_leg["value"] / ceiling,
        "bound": "PCIe host->device: the pipeline moves the LQ data at the link rate measured in this run (host_link); the kernels are hidden behind the copies",
        "how": "o2c_solve_host: pinned host SoA buffers -> chunked H2D, pack, sweep+rollout, unpack, D2H on 3 overlapping stream lanes; host wall "
               "clock around the blocking call (it returns after the last D2H), max over ranks",
    })
    return main_leg


if __name__ == "__main__":
    main()
