#!/usr/bin/env python
"""bench.py -- throughput of the batched FGD hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--workload c5|c1|c2|c3|c4] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input: every
trajectory of the batch is optimised to completion (all outer / inner / line
search iterations) by the persistent sm_100a kernel.  Default workload = config 5
of BASELINE.json, the one north_star's target is quoted on: the 1 M-trajectory
random-restart sweep (4096 problems x 256 restarts, BLS), strong scaling - the
restart axis is sharded over the ranks, and the local argmin kernel plus the
sweep's single NCCL collective (all-gather of the per-problem order keys) are
INSIDE the timed step.  The single-GPU default run also measures configs 2, 3
and 4 (`secondary`), each with its own roofline object.  Prints ONE JSON line
(rank 0).  See DESIGN.md section 6 for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "optimized trajectories/sec"
UNIT = "trajectories/s"
DEFAULT_WORKLOAD = "c5"


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--batch", type=int, default=0, help="whole-job batch override (0 = the configuration's size)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="budget of the bounded CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run parity check of a subset against the oracle")
    ap.add_argument("--strict-math", action="store_true")
    ap.add_argument("--whole-arm", action="store_true", help="obstacle cost over all joint positions (SURVEY 8f-3) instead of the end effector")
    ap.add_argument("--presoak-seconds", type=float, default=1.0,
                    help="untimed repetitions of the step before the W warm-up steps, so that the GPU has left its idle clocks")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads of the default single-GPU run")
    ap.add_argument("--secondary", default="c2,c2sat,c3,c4,c1,c5_strict", help="secondary workloads of the default single-GPU run")
    ap.add_argument("--c4-relaunch", action="store_true", help="c4 through one launch per 8 iterations (round-1 path) instead of the live kernel")
    ap.add_argument("--c4-max-sets", type=int, default=-1,
                    help="c4 live: publish at most N obstacle sets per launch (-1 = until the launch ends; 0 = a deterministic launch for ncu "
                         "captures: under kernel replay the publisher's timing differs from pass to pass)")
    return ap.parse_args(argv)


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------

def hp_view(args_ns, T):
    hp = type("HP", (), dict(vars(args_ns)))()
    hp.n_timesteps = T
    return hp


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under load (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 <= ts <= t1 + 0.06):
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """all host threads this process may use (torchrun pins OMP_NUM_THREADS=1, so ask the OS)"""
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_baseline(wl, traj, alpha0, start, goal, seconds, threads=0):
    """The C mirror oracle (a port of the reference algorithm) on the host cores, on a bounded
    sample of the same workload."""
    from oracle import mirror as M
    m = M.Mirror(hp_view(wl.args, traj.N_timesteps), traj.km, traj.dkm, traj.jac, wl.obstacles, wl.mode)
    if threads <= 0:
        threads = host_threads()
    cores = threads
    n0 = min(len(alpha0), max(cores * 2, 16))
    t = time.perf_counter()
    m.optimize(alpha0[:n0], start[:n0], goal[:n0], nthreads=threads)
    rate = n0 / max(time.perf_counter() - t, 1e-6)
    n = int(min(len(alpha0), max(n0, rate * seconds)))
    t = time.perf_counter()                                # calibrate on one full pass (the first sample pays thread start-up)
    m.optimize(alpha0[:n], start[:n], goal[:n], nthreads=threads)
    rate = max(rate, n / max(time.perf_counter() - t, 1e-6))
    reps = max(1, int(round(rate * seconds / n)))          # small batches: repeat the pass to fill the budget
    t = time.perf_counter()
    iters = 0
    for _ in range(reps):
        _, fs, is_ = m.optimize(alpha0[:n], start[:n], goal[:n], nthreads=threads)
        iters += int(is_[:, M.I_INNER_TOTAL].sum())
    dt = time.perf_counter() - t
    return {"value": n * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n} trajectories of the {wl.name} batch x {reps} passes, C mirror oracle (oracle/fgd_mirror.c), "
                      f"OpenMP over trajectories, {dt:.1f} s",
            "fgd_iters_per_s": float(iters / dt)}, n * reps, dt


def config_dict(wl, T, world):
    """Identical in both arms (the driver compares them)."""
    return {"workload": f"{wl.name}: {wl.description}", "optimizer": wl.mode,
            "trajectories_total": int(wl.B if wl.n_problems else wl.B * world),
            "trajectories_per_gpu": int(wl.B // world) if wl.n_problems else int(wl.B), "n_timesteps": int(T),
            "n_obstacles": int(len(wl.obstacles)), "l2_flush_between_steps": True,
            "obstacle_cost": "whole arm (3 joint positions)" if getattr(wl.args, "whole_arm_cost", False) else "end effector"}


# --------------------------------------------------------------------------
# reference arm: the reference algorithm on the host cores
# --------------------------------------------------------------------------

def run_reference(a, rank, world):
    if rank != 0:
        return
    from irm_motion_planning_b200.trajectory import Trajectory
    from irm_motion_planning_b200.workloads import initial_alpha, make_workload
    wl = make_workload(a.workload, B=a.batch or None, seed=a.seed)
    wl.args.whole_arm_cost = bool(a.whole_arm)
    traj = Trajectory(wl.args, create_handle=False)
    # the sample is a prefix of the batch: generate only as many problems of the sweep as a step can consume
    if wl.n_problems:
        alpha0, start, goal = initial_alpha(wl, traj, a.seed, problems=(0, min(wl.n_problems, 512)))
    else:
        alpha0, start, goal = initial_alpha(wl, traj, a.seed)
    # each step: a bounded sample, sized so the whole run stays within a few minutes
    per_step = max(2.0, min(a.cpu_seconds, 120.0 / max(1, a.steps + a.warmup)))
    times, n_used, info = [], 0, None
    for i in range(a.warmup + a.steps):
        info, n, dt = cpu_baseline(wl, traj, alpha0, start, goal, per_step)
        if i >= a.warmup:
            times.append(dt); n_used = n
    value = n_used * len(times) / sum(times)
    info["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
            "scaling": "strong" if wl.n_problems else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(wl, traj.N_timesteps, max(1, a.gpus)),
            "note": f"each step = {n_used} trajectories on {info['cores']} host threads",
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------

class Ctx:
    """per-process state shared by the measurements of one bench run"""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.a = a
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the FGD hot path has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.dev = torch.device("cuda", self.local)
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)   # > 126 MB L2
        self.peak_fp32 = None
        self.peak_mufu = None
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            self.peaks = {}
        self.executed = _load_json(os.path.join(ROOT, "profiles", "executed.json"))
        self.traffic = _load_json(os.path.join(ROOT, "profiles", "traffic.json"))

    def barrier(self):
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce(self, vals, op="sum"):
        import torch
        import torch.distributed as dist
        t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]


def _load_json(path):
    try:
        return json.load(open(path))
    except Exception:
        return {}


def roofline_object(ctx, key, flops_per_gpu, kern_s, hbm_bytes):
    """FP32 roofline of one workload: algorithmic FLOPs (the reference algorithm's) per second against the
    FFMA peak measured in this run; beside it the EXECUTED FP32 fraction and, for the obstacle-bound shapes, the MUFU
    fraction - both from the ncu instruction counters of the same kernel committed under profiles/ (executed.json)."""
    peak, achieved = ctx.peak_fp32, flops_per_gpu / kern_s * 1e-12
    ex = ctx.executed.get(key) or {}
    r = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
         "traffic": ctx.traffic.get(key),
         "peak_source": "FFMA probe kernel measured in this run (fgd_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 CUDA-core figure; "
                        "nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
         "flops": "algorithmic FLOPs of the reference algorithm per launch (SURVEY 8d: F_grad, F_cost per consumed evaluation) / CUDA-event time",
         "hbm": {"algorithmic_bytes_per_launch": int(hbm_bytes), "achieved_gbs": hbm_bytes / kern_s * 1e-9,
                 "peak_gbs": ctx.peaks.get("hbm_gbs"), "note": "HBM is not the bound: ~25 B per trajectory-iteration"}}
    if ex.get("fp32_flop_per_alg_flop"):
        r["executed_frac"] = r["frac"] * ex["fp32_flop_per_alg_flop"] if r["frac"] is not None else None
        r["executed"] = {"fp32_flop_per_algorithmic_flop": ex["fp32_flop_per_alg_flop"], "source": ex.get("source"),
                         "note": "executed FP32 flops (SASS opcode histogram of the ncu --set full source page, warp level: FFMA2 = 128, "
                                 "FADD2 / FMUL2 / FFMA = 64, FADD / FMUL = 32 per warp instruction) per algorithmic FLOP of the same launch; "
                                 "< 1 because q, v and the loss of an accepted candidate are reused and the sparse dK product skips zeros"}
    if ex.get("mufu_per_alg_flop") and ctx.peak_mufu:
        r["mufu_frac"] = ex["mufu_per_alg_flop"] * flops_per_gpu / kern_s * 1e-12 / ctx.peak_mufu
        r["mufu_peak_trcp"] = ctx.peak_mufu
    return r


def measure(ctx, name, steps, warmup, batch=None, e2e=True, parity=True, presoak=None, clocks=False, cpu=False, key=None, strict=None):
    """One workload on this process's GPU (all ranks call it together).  Returns the fields of a bench line."""
    import torch
    from irm_motion_planning_b200 import backend
    from irm_motion_planning_b200.batch import BatchedFGD, decode_keys, gather_best_keys, restart_shard
    from irm_motion_planning_b200.trajectory import Trajectory
    from irm_motion_planning_b200.workloads import flops_total, initial_alpha, make_workload, obstacle_swap

    a, rank, world, dev = ctx.a, ctx.rank, ctx.world, ctx.dev
    key = key or name
    wl = make_workload(name, B=batch, seed=a.seed + (0 if name == "c5" else rank))
    strong = bool(wl.n_problems)
    wl.args.whole_arm_cost = bool(a.whole_arm)
    joints = 3 if a.whole_arm else 1
    strict = a.strict_math if strict is None else strict
    traj = Trajectory(wl.args, strict_math=strict)
    traj.set_obstacles(wl.obstacles)
    r_lo, r_hi = 0, wl.n_restarts
    if strong:                                     # every rank: all problems, its block of the restart axis
        r_lo, r_hi = restart_shard(wl.n_restarts, rank, world)
        alpha0, start, goal = initial_alpha(wl, traj, a.seed, restarts=(r_lo, r_hi))
    else:
        alpha0, start, goal = initial_alpha(wl, traj, a.seed + rank)
    P, R_local = wl.n_problems, r_hi - r_lo
    B, T = len(alpha0), traj.N_timesteps
    eng = BatchedFGD(traj, wl.mode)
    h = traj.handle

    n_total = warmup + steps
    a0_dev = torch.as_tensor(alpha0, device=dev)
    s_dev = torch.as_tensor(start, device=dev).contiguous()
    g_dev = torch.as_tensor(goal, device=dev).contiguous()
    # inputs of every step are resident in HBM before the timed region (one alpha buffer per step, recycled for very large batches)
    n_buf = n_total if B * T * 12 * n_total < 8e9 else 2
    bufs = [a0_dev.clone() for _ in range(n_buf)]
    live = name == "c4" and not a.c4_relaunch and hasattr(eng, "optimize_live")
    swaps = [torch.as_tensor(wl.obstacles).pin_memory()] + [torch.as_tensor(obstacle_swap(k, a.seed)).pin_memory() for k in range(1, 64)] \
        if name == "c4" else None
    info = {"swaps": 0}

    def one_step(buf):
        """Launch the hot path for one batch; returns (fstate, istate, winning keys of the sweep or None)."""
        fs, is_ = eng.new_state(B)
        keys = None
        if name == "c4" and live:
            info["swaps"] += eng.optimize_live(buf, s_dev, g_dev, fs, is_, swaps, poll_every=8,
                                               max_sets=None if a.c4_max_sets < 0 else a.c4_max_sets)
        elif name == "c4":
            for k in range(10000):
                h.set_obstacles(swaps[k % len(swaps)])
                eng.optimize_device(buf, s_dev, g_dev, fs, is_, max_launch_iters=8)
                if bool((is_[:, backend.I_STATUS] == backend.ST_DONE).all()):
                    break
        else:
            eng.optimize_device(buf, s_dev, g_dev, fs, is_)
        if strong:
            keys = eng.best_keys(fs, is_, P, R_local, index_offset=r_lo, problem_stride=wl.n_restarts)
            keys = gather_best_keys(keys)          # the single NCCL collective of the sweep
        return fs, is_, keys

    sampler = ClockSampler(ctx.local) if clocks and rank == 0 else None
    if sampler:
        sampler.start()
    # pre-soak (untimed): the host-side input generation above leaves the GPU idle for seconds; the first steps after that
    # run at ramping clocks, so the step is repeated for --presoak-seconds before the W warm-up steps
    # (the repetition count is agreed across ranks: a step may contain a collective)
    presoak = a.presoak_seconds if presoak is None else presoak
    if presoak > 0:
        torch.cuda.synchronize()
        t_one = time.perf_counter()
        one_step(bufs[0])
        torch.cuda.synchronize()
        dt_one = ctx.allreduce([time.perf_counter() - t_one], "max")[0]
        for _ in range(int(min(5000, presoak / max(dt_one, 1e-6)))):
            bufs[0].copy_(a0_dev)
            one_step(bufs[0])
        torch.cuda.synchronize()
    bufs[0].copy_(a0_dev)
    for i in range(warmup):
        if n_buf < n_total:
            bufs[i % n_buf].copy_(a0_dev)
        one_step(bufs[i % n_buf])
        ctx.flush.zero_()
    ctx.barrier()
    launches0 = h.kernel_launches()
    t_load0 = time.perf_counter()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    last, last_buf, info["swaps"] = None, None, 0
    wall0 = time.perf_counter()
    for i in range(steps):
        buf = bufs[(warmup + i) % n_buf]
        if n_buf < n_total:
            buf.copy_(a0_dev)
        ctx.flush.zero_()                               # evict the previous step's lines from L2 (outside the event pair)
        ev[i][0].record()
        last, last_buf = one_step(buf), buf
        ev[i][1].record()
    ctx.barrier()
    wall1 = time.perf_counter()
    launches = h.kernel_launches() - launches0
    step_ms = [s.elapsed_time(e) for s, e in ev]
    total_own = sum(step_ms) * 1e-3
    if world > 1:
        print(f"[bench] rank {rank}: {1e3 * total_own / steps:.3f} ms per step on its own device "
              f"(steps: {', '.join('%.2f' % m for m in step_ms)})", file=sys.stderr, flush=True)
    total_s = ctx.allreduce([total_own], "max")[0]
    total_min = -ctx.allreduce([-total_own], "max")[0]

    fs, is_ = last[0].cpu().numpy(), last[1].cpu().numpy()
    inner, cand = is_[:, backend.I_INNER_TOTAL], is_[:, backend.I_CAND_EVALS]
    outer = np.maximum(1, is_[:, backend.I_OUTER] + is_[:, backend.I_FULFILLED])
    assert (is_[:, backend.I_STATUS] == backend.ST_DONE).all(), "a timed step left unfinished trajectories"
    O_eff = len(wl.obstacles)
    flops_step = flops_total(wl.mode, T, O_eff, inner, cand, outer, joints)
    n_traj_all, n_iter_all, flops_all = ctx.allreduce([float(B), float(inner.sum()), flops_step], "sum")
    value = n_traj_all * steps / total_s

    # cost of the sweep's reduction (local argmin kernel + all-gather + min), timed apart after the timed region
    sweep = None
    if strong:
        fs_d, is_d = last[0], last[1]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        e0.record()
        for _ in range(10):
            gather_best_keys(eng.best_keys(fs_d, is_d, P, R_local, index_offset=r_lo, problem_stride=wl.n_restarts))
        e1.record()
        torch.cuda.synchronize()
        cost_w, idx_w, ful_w = decode_keys(last[2])
        sweep = {"problems": int(P), "restarts": int(wl.n_restarts), "restarts_per_gpu": int(R_local),
                 "sharding": "restart axis: every rank optimises restarts [lo, hi) of every problem",
                 "collective": "one all-gather of the per-problem order keys (8 B x problems per rank), reduced by an elementwise min",
                 "argmin_plus_gather_ms": e0.elapsed_time(e1) / 10,
                 "rank_ms_per_step_min": 1e3 * total_min / steps, "rank_ms_per_step_max": 1e3 * total_s / steps,
                 "problems_with_fulfilled_winner": float(ful_w.float().mean().item()),
                 "mean_winner_cost": float(cost_w.mean().item())}

    # soak: the same step repeated (untimed) so nvidia-smi sees the clocks this kernel runs at
    clocks_out = None
    if clocks:
        t_soak = time.perf_counter()
        n_soak = 0
        while True:
            bufs[0].copy_(a0_dev)
            one_step(bufs[0])
            torch.cuda.synchronize()
            n_soak += 1
            go = ctx.allreduce([1.0 if time.perf_counter() - t_soak < 1.0 else 0.0], "max")[0] if world > 1 else \
                (1.0 if time.perf_counter() - t_soak < 1.0 else 0.0)
            if not go:
                break
        t_load1 = time.perf_counter()
        clocks_out = sampler.stop(t_load0, t_load1) if sampler else None

    # end to end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timing)
    e2e_out = None
    if e2e and name != "c4":
        pin = lambda x: torch.as_tensor(x).pin_memory()
        a_pin, s_pin, g_pin = pin(alpha0), pin(start), pin(goal)
        out_a = torch.empty_like(a_pin).pin_memory()
        out_f = torch.empty(B, backend.FSTATE, dtype=torch.float32).pin_memory()
        out_i = torch.empty(B, backend.ISTATE, dtype=torch.int32).pin_memory()
        keys_pin = torch.empty(max(P, 1), dtype=torch.int64).pin_memory()

        def e2e_step():
            eng.optimize_pinned(a_pin, s_pin, g_pin, out_a, out_f, out_i)
            if strong:      # the argmin kernel reads the state the optimiser just wrote to the pinned buffers, the winners return to the host
                k = gather_best_keys(eng.best_keys(out_f, out_i, P, R_local, index_offset=r_lo, problem_stride=wl.n_restarts))
                keys_pin.copy_(k, non_blocking=True)
                torch.cuda.synchronize()

        for _ in range(max(1, min(warmup, 2))):
            e2e_step()
        ctx.barrier()
        zc0 = h.zero_copy_calls()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        ctx.barrier()
        zero_copy = (h.zero_copy_calls() - zc0) == steps
        e2e_s = ctx.allreduce([time.perf_counter() - t0], "max")[0]
        h2d = B * T * 12 + 2 * B * 12
        d2h = B * T * 12 + B * (backend.FSTATE + backend.ISTATE) * 4 + (P * 8 if strong else 0)
        e2e_out = {"value": n_traj_all * steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(d2h), "api": "fgd_optimize_host_io via BatchedFGD.optimize_pinned"
                   + (" + fgd_argmin_per_problem + all-gather, winners copied to the host" if strong else ""),
                   "transfer": "zero-copy: the kernel reads inputs from / writes results to the pinned host buffers over PCIe, per trajectory"
                               if zero_copy else "staged: cudaMemcpyAsync H2D, launch, cudaMemcpyAsync D2H"}
        assert np.array_equal(out_i.numpy()[:, backend.I_STATUS], np.full(B, backend.ST_DONE)), "e2e left unfinished trajectories"
        if strong:
            assert torch.equal(keys_pin, last[2].cpu()), "e2e sweep winners differ from the device-resident step"
        del a_pin, out_a, out_f, out_i

    out = {"rank": rank}
    if rank == 0:
        if ctx.peak_fp32 is None:
            ctx.peak_fp32 = h.measure_fp32_peak()
            ctx.peak_mufu = h.measure_mufu_peak()
        kern_s = total_s / steps
        hbm_bytes = B * (2 * T * 12 + 2 * 12 + 2 * (backend.FSTATE + backend.ISTATE) * 4)
        out.update({
            "value": value, "ms_per_step": 1e3 * total_s / steps, "scaling": "strong" if strong else "weak",
            "config": config_dict(wl, T, world), "launch": h.launch_geometry(B),
            "math": "strict (IEEE reciprocal, division, square root: bit-exact against the oracle)" if strict else "fast (rcp.approx)",
            "fgd_iters_per_s": n_iter_all * steps / total_s, "mean_inner_iters": float(inner.mean()),
            "fulfilled_frac": float(is_[:, backend.I_FULFILLED].mean()),
            "wall_ms_per_step": 1e3 * (wall1 - wall0) / steps, "step_ms": [round(m, 4) for m in step_ms],
            "e2e": e2e_out, "gpu_launches": int(launches),
            "roofline": roofline_object(ctx, key, flops_all / max(world, 1), kern_s, hbm_bytes)})
        if clocks:
            out["clocks"] = clocks_out
        if sweep:
            out["sweep"] = sweep
        if name == "c4":
            out["dynamic_obstacles"] = {"mode": "live: one persistent launch, the host publishes obstacle sets with fgd_set_obstacles_async "
                                                "while it runs, every team polls the generation counter every 8 inner iterations" if live
                                        else "one launch per 8 inner iterations", "sets_published_per_step": info["swaps"] / max(steps, 1)}
        if parity and not strict:
            out["parity"] = parity_check(ctx, wl, traj, alpha0, start, goal, last, last_buf, P, R_local, r_lo)
        if cpu:
            out["cpu_baseline"], _, _ = cpu_baseline(wl, traj, alpha0, start, goal, a.cpu_seconds)
    del bufs, a0_dev
    torch.cuda.empty_cache()
    return out


def parity_check(ctx, wl, traj, alpha0, start, goal, last, last_alpha, P, R_local, r_lo, n_sub=256):
    """After the timed region: a subset of the timed batch against the CPU mirror oracle.
    (1) the same subset re-run on the GPU in strict-math mode must equal the oracle bit for bit;
    (2) the timed (fast-math) results of the subset are compared trajectory by trajectory: identical decision traces,
        joint-angle and cost errors against the stated tolerance (5e-2 rad, 1e-2 relative: SURVEY 8d-ii);
    (3) sweeps: the GPU argmin over the subset's whole problems equals the oracle's argmin."""
    import torch
    from irm_motion_planning_b200 import backend
    from irm_motion_planning_b200.batch import BatchedFGD
    from irm_motion_planning_b200.trajectory import Trajectory
    from oracle import mirror as M
    B = len(alpha0)
    if P:                                           # whole problems: the first restarts block of a few problems
        n_prob = max(1, n_sub // R_local)
        sub = np.arange(n_prob * R_local)
    else:
        sub = np.linspace(0, B - 1, min(n_sub, B)).astype(np.int64)
    m = M.Mirror(hp_view(wl.args, traj.N_timesteps), traj.km, traj.dkm, traj.jac, wl.obstacles, wl.mode)
    ca, cfs, cis = m.optimize(alpha0[sub], start[sub], goal[sub], nthreads=host_threads())
    ts = Trajectory(wl.args, strict_math=True)
    ts.set_obstacles(wl.obstacles)
    eng = BatchedFGD(ts, wl.mode)
    a_s = torch.as_tensor(alpha0[sub], device=ctx.dev).clone()
    res = eng.optimize_device(a_s, torch.as_tensor(start[sub], device=ctx.dev).contiguous(), torch.as_tensor(goal[sub], device=ctx.dev).contiguous())
    torch.cuda.synchronize()
    strict_ok = bool(np.array_equal(res.alpha.cpu().numpy(), ca) and np.array_equal(res.istate.cpu().numpy(), cis))
    out = {"subset": int(len(sub)), "oracle": "oracle/fgd_mirror.c", "strict_math_bit_exact": strict_ok}
    if P:
        keys = eng.best_keys(res.fstate, res.istate, n_prob, R_local, index_offset=r_lo, problem_stride=wl.n_restarts).cpu().numpy()
        toc, ful = cfs[:, M.F_TOC].reshape(n_prob, R_local), cis[:, M.I_FULFILLED].reshape(n_prob, R_local).astype(bool)
        c = np.where(ful, toc, np.inf)
        c = np.where(np.isinf(c).all(1, keepdims=True), toc, c)
        r = c.argmin(1)
        out["argmin_matches_oracle"] = bool(np.array_equal(keys & 0x7FFFFFFF, r_lo + np.arange(n_prob) * wl.n_restarts + r))
    # the timed fast-math results of the same trajectories
    sub_d = torch.as_tensor(sub, device=ctx.dev)
    fs_t, is_t, a_t = last[0][sub_d].cpu().numpy(), last[1][sub_d].cpu().numpy(), last_alpha[sub_d].cpu().numpy()
    dq = np.abs(np.einsum("ij,bjk->bik", traj.km, a_t - ca) @ traj.jac).reshape(len(sub), -1).max(1)      # joint angles: q = K alpha J
    same = is_t[:, backend.I_HASH] == cis[:, M.I_HASH]
    both = (is_t[:, backend.I_FULFILLED] == 1) & (cis[:, M.I_FULFILLED] == 1)
    rel = np.abs(fs_t[:, backend.F_TOC] - cfs[:, M.F_TOC]) / np.maximum(cfs[:, M.F_TOC], 1e-6)
    out.update({"fast_math_identical_traces": float(same.mean()),
                "fast_math_fulfilled_agree": float((is_t[:, backend.I_FULFILLED] == cis[:, M.I_FULFILLED]).mean()),
                "fast_math_max_dq_same_trace": float(dq[same].max()) if same.any() else None,
                "fast_math_dq_p99": float(np.quantile(dq, 0.99)), "fast_math_dq_max": float(dq.max()),
                "fast_math_dq_outside_5e-2": float((dq > 5e-2).mean()),
                "fast_math_rel_cost_err_max_same_trace": float(rel[same].max()) if same.any() else None,
                "fast_math_rel_cost_err_p99": float(np.quantile(rel[both], 0.99)) if both.any() else None,
                "fast_math_rel_cost_outside_1e-2": float((rel[both] > 1e-2).mean()) if both.any() else None,
                "tolerance": "stated: max-abs joint angle 5e-2 rad, relative obstacle cost 1e-2 (SURVEY 8d-ii); trajectories whose decision "
                             "trace differs from the oracle's diverge chaotically (SURVEY 0.3-3) and are reported as a fraction"})
    return out


def main(argv=None):
    a = parse(argv)
    if a.impl == "reference":
        run_reference(a, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return
    import torch.distributed as dist
    ctx = Ctx(a)
    primary = measure(ctx, a.workload, a.steps, a.warmup, batch=a.batch or None, e2e=not a.no_e2e, parity=not a.no_parity,
                      clocks=True, cpu=(ctx.world == 1 and not a.no_cpu_baseline))
    secondary = {}
    if ctx.world == 1 and not a.no_secondary and not a.batch and a.workload == DEFAULT_WORKLOAD:
        # the other BASELINE configurations on the same GPU, same kernels: fewer steps, each with its own roofline object
        plan = {"c2": dict(name="c2", steps=10, warmup=3, e2e=True, presoak=0.2),
                "c2sat": dict(name="c2", steps=3, warmup=1, batch=65536, e2e=False, presoak=0.0, parity=False, key="c2_b65536"),
                "c3": dict(name="c3", steps=2, warmup=1, e2e=False, presoak=0.0),
                "c4": dict(name="c4", steps=2, warmup=1, e2e=False, presoak=0.0, parity=False),
                # the reference's own problem (one trajectory): ms_per_step = latency of one optimize() call
                "c1": dict(name="c1", steps=20, warmup=5, e2e=False, presoak=0.0, parity=False),
                # the primary workload in strict-math mode: what the bit-exact kernels cost
                "c5_strict": dict(name="c5", steps=3, warmup=1, e2e=False, presoak=0.0, parity=False, strict=True, key="c5_strict")}
        for k in [s for s in a.secondary.split(",") if s]:
            r = measure(ctx, **plan[k])
            r.pop("rank", None)
            secondary[k] = r
    if ctx.rank == 0:
        line = {"metric": METRIC, "value": primary["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": primary["ms_per_step"], "higher_is_better": True, "scaling": primary["scaling"],
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "presoak_s": a.presoak_seconds}
        line.update({k: v for k, v in primary.items() if k not in ("rank", "value", "ms_per_step", "scaling")})
        if secondary:
            line["secondary"] = secondary
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
