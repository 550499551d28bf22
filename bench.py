#!/usr/bin/env python
"""bench.py -- throughput of the batched FGD hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--workload c2|c1|c3|c4|c5] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input: every
trajectory of the batch is optimised to completion (all outer / inner / line
search iterations) by the persistent sm_100a kernel.  Default workload = config 2
of BASELINE.json (fixed-step GD, default scene, 4096 random-init trajectories per
GPU).  Prints ONE JSON line (rank 0).  See DESIGN.md section 6 for every field.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (0 = the configuration's size)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="budget of the bounded CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--strict-math", action="store_true")
    ap.add_argument("--whole-arm", action="store_true", help="obstacle cost over all joint positions (SURVEY 8f-3) instead of the end effector")
    ap.add_argument("--presoak-seconds", type=float, default=1.0,
                    help="untimed repetitions of the step before the W warm-up steps, so that the GPU has left its idle clocks")
    ap.add_argument("--no-saturated", action="store_true", help="skip the secondary large-batch measurement of the default run")
    ap.add_argument("--saturated-batch", type=int, default=65536)
    return ap.parse_args()


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


def cpu_baseline(wl, traj, alpha0, start, goal, seconds, threads=0):
    """The C mirror oracle (a port of the reference algorithm) on the host cores, on a bounded
    sample of the same workload."""
    from oracle import mirror as M
    m = M.Mirror(hp_view(wl.args, traj.N_timesteps), traj.km, traj.dkm, traj.jac, wl.obstacles, wl.mode)
    if threads <= 0:      # all host threads this process may use (torchrun pins OMP_NUM_THREADS=1, so ask the OS)
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
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
            "warmup": a.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(wl, traj, len(alpha0), note=f"each step = {n_used} trajectories on {info['cores']} host threads"),
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def config_dict(wl, traj, B, **extra):
    d = {"workload": f"{wl.name}: {wl.description}", "optimizer": wl.mode, "trajectories_per_gpu": int(B),
         "n_timesteps": traj.N_timesteps, "n_obstacles": int(len(wl.obstacles)), "l2_flush_between_steps": True,
         "obstacle_cost": "whole arm (3 joint positions)" if getattr(wl.args, "whole_arm_cost", False) else "end effector"}
    d.update(extra)
    return d


# --------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------

def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return

    import torch
    import torch.distributed as dist
    from irm_motion_planning_b200 import backend
    from irm_motion_planning_b200.batch import BatchedFGD, gather_best, shard_range
    from irm_motion_planning_b200.trajectory import Trajectory
    from irm_motion_planning_b200.workloads import flops_total, initial_alpha, make_workload, obstacle_swap

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the FGD hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    strong = a.workload == "c5"
    wl = make_workload(a.workload, B=a.batch or None, seed=a.seed + (0 if strong else rank))
    wl.args.whole_arm_cost = bool(a.whole_arm)
    joints = 3 if a.whole_arm else 1
    traj = Trajectory(wl.args, strict_math=a.strict_math)
    traj.set_obstacles(wl.obstacles)
    alpha0, start, goal = initial_alpha(wl, traj, a.seed + (0 if strong else rank))
    prob_lo = 0
    if strong and world > 1:                       # shard whole problems (all restarts of a problem on one rank)
        prob_lo, prob_hi = shard_range(wl.n_problems, rank, world)
        sl = slice(prob_lo * wl.n_restarts, prob_hi * wl.n_restarts)
        alpha0, start, goal = alpha0[sl], start[sl], goal[sl]
    n_prob_local = (len(alpha0) // wl.n_restarts) if wl.n_problems else 0
    B, T = len(alpha0), traj.N_timesteps
    eng = BatchedFGD(traj, wl.mode)
    h = traj.handle

    n_total = a.warmup + a.steps
    a0_dev = torch.as_tensor(alpha0, device=dev)
    s_dev = torch.as_tensor(start, device=dev).contiguous()
    g_dev = torch.as_tensor(goal, device=dev).contiguous()
    # inputs of every step are resident in HBM before the timed region (one alpha buffer per step,
    # recycled for very large batches)
    n_buf = n_total if B * T * 12 * n_total < 8e9 else 2
    bufs = [a0_dev.clone() for _ in range(n_buf)]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2
    swaps = [torch.as_tensor(wl.obstacles).pin_memory()] + [torch.as_tensor(obstacle_swap(k, a.seed)).pin_memory() for k in range(1, 64)] \
        if a.workload == "c4" else None

    def one_step(buf):
        """Launch the hot path for one batch; returns (fstate, istate)."""
        fs, is_ = eng.new_state(B)
        if a.workload == "c4":
            for k in range(10000):
                h.set_obstacles(swaps[k % len(swaps)])
                eng.optimize_device(buf, s_dev, g_dev, fs, is_, max_launch_iters=8)
                if bool((is_[:, backend.I_STATUS] == backend.ST_DONE).all()):
                    break
        else:
            eng.optimize_device(buf, s_dev, g_dev, fs, is_)
        if wl.n_problems:
            cost, idx = eng.best_per_problem(type("R", (), {"fstate": fs, "istate": is_})(), n_prob_local, wl.n_restarts,
                                             index_offset=prob_lo * wl.n_restarts)
            gather_best(cost, idx)                 # the single NCCL collective of the sweep
        return fs, is_

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # pre-soak (untimed): the host-side input generation above leaves the GPU idle for seconds; the first steps after that
    # run at ramping clocks (measured on c5 at N = 2: 487 / 378 / 359 ms for three consecutive steps), so the step is
    # repeated for --presoak-seconds before the W warm-up steps
    # (the repetition count is agreed across ranks: a step may contain a collective)
    if a.presoak_seconds > 0:
        torch.cuda.synchronize()
        t_one = time.perf_counter()
        one_step(bufs[0])
        torch.cuda.synchronize()
        dt_one = torch.tensor([time.perf_counter() - t_one], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt_one, op=dist.ReduceOp.MAX)
        for _ in range(int(min(5000, a.presoak_seconds / max(float(dt_one.item()), 1e-6)))):
            bufs[0].copy_(a0_dev)
            one_step(bufs[0])
        torch.cuda.synchronize()
    bufs[0].copy_(a0_dev)
    for i in range(a.warmup):
        if n_buf < n_total:
            bufs[i % n_buf].copy_(a0_dev)
        one_step(bufs[i % n_buf])
        flush.zero_()
    barrier()
    launches0 = h.kernel_launches()
    t_load0 = time.perf_counter()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    last = None
    wall0 = time.perf_counter()
    for i in range(a.steps):
        buf = bufs[(a.warmup + i) % n_buf]
        if n_buf < n_total:
            buf.copy_(a0_dev)
        flush.zero_()                               # evict the previous step's lines from L2 (outside the event pair)
        ev[i][0].record()
        last = one_step(buf)
        ev[i][1].record()
    barrier()
    wall1 = time.perf_counter()
    launches = h.kernel_launches() - launches0
    step_ms = [s.elapsed_time(e) for s, e in ev]
    total_s = sum(step_ms) * 1e-3
    t_max = torch.tensor([total_s], dtype=torch.float64, device=dev)
    if world > 1:
        print(f"[bench] rank {rank}: {1e3 * total_s / a.steps:.3f} ms per step on its own device "
              f"(steps: {', '.join('%.2f' % m for m in step_ms)})", file=sys.stderr, flush=True)
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    total_s = float(t_max.item())

    fs, is_ = last[0].cpu().numpy(), last[1].cpu().numpy()
    inner, cand = is_[:, backend.I_INNER_TOTAL], is_[:, backend.I_CAND_EVALS]
    outer = np.maximum(1, is_[:, backend.I_OUTER] + is_[:, backend.I_FULFILLED])
    assert (is_[:, backend.I_STATUS] == backend.ST_DONE).all(), "a timed step left unfinished trajectories"
    O_eff = len(wl.obstacles)
    flops_step = flops_total(wl.mode, T, O_eff, inner, cand, outer, joints)
    counts = torch.tensor([float(B), float(inner.sum()), flops_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    n_traj_all, n_iter_all, flops_all = [float(x) for x in counts.tolist()]
    value = n_traj_all * a.steps / total_s

    # soak: the same step repeated (untimed) so nvidia-smi sees the clocks this kernel runs at
    t_soak = time.perf_counter()
    while time.perf_counter() - t_soak < 1.0:
        bufs[0].copy_(a0_dev)
        one_step(bufs[0])
        torch.cuda.synchronize()
    t_load1 = time.perf_counter()
    clocks = sampler.stop(t_load0, t_load1) if rank == 0 else None

    # end to end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timing)
    e2e = None
    if not a.no_e2e and a.workload != "c4":
        pin = lambda x: torch.as_tensor(x).pin_memory()
        a_pin, s_pin, g_pin = pin(alpha0), pin(start), pin(goal)
        out_a = torch.empty_like(a_pin).pin_memory()
        out_f = torch.empty(B, backend.FSTATE, dtype=torch.float32).pin_memory()
        out_i = torch.empty(B, backend.ISTATE, dtype=torch.int32).pin_memory()
        for _ in range(max(1, min(a.warmup, 2))):
            eng.optimize_pinned(a_pin, s_pin, g_pin, out_a, out_f, out_i)
        barrier()
        zc0 = h.zero_copy_calls()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            eng.optimize_pinned(a_pin, s_pin, g_pin, out_a, out_f, out_i)
        barrier()
        zero_copy = (h.zero_copy_calls() - zc0) == a.steps
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        h2d = B * T * 12 + 2 * B * 12
        d2h = B * T * 12 + B * (backend.FSTATE + backend.ISTATE) * 4
        e2e = {"value": n_traj_all * a.steps / float(e2e_s.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "api": "fgd_optimize_host_io via BatchedFGD.optimize_pinned",
               "transfer": "zero-copy: the kernel reads inputs from / writes results to the pinned host buffers over PCIe, per trajectory"
                           if zero_copy else "staged: cudaMemcpyAsync H2D, launch, cudaMemcpyAsync D2H"}
        assert np.array_equal(out_i.numpy()[:, backend.I_STATUS], np.full(B, backend.ST_DONE)), "e2e left unfinished trajectories"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak = h.measure_fp32_peak()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    kern_s = total_s / a.steps
    achieved = flops_all / max(world, 1) / kern_s * 1e-12      # per GPU
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(a.workload)
    except Exception:
        pass
    hbm_bytes = B * (2 * T * 12 + 2 * 12 + 2 * (backend.FSTATE + backend.ISTATE) * 4)
    roofline = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "traffic": traffic,
                "peak_source": "FFMA probe kernel measured in this run (fgd_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 CUDA-core figure; "
                               "nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
                "flops": "algorithmic FLOPs of the reference algorithm per launch (SURVEY 8d: F_grad, F_cost per consumed evaluation) / CUDA-event time",
                "hbm": {"algorithmic_bytes_per_launch": int(hbm_bytes), "achieved_gbs": hbm_bytes / kern_s * 1e-9,
                        "peak_gbs": peaks.get("hbm_gbs"), "note": "HBM is not the bound: ~25 B per trajectory-iteration"}}

    # secondary measurement (default single-GPU run only): the same workload at a batch that fills the GPU many
    # times over, so that the kernel's throughput is visible next to the tail-limited headline batch
    saturated = None
    if world == 1 and not a.no_saturated and not a.batch and a.workload == "c2":
        wl2 = make_workload(a.workload, B=a.saturated_batch, seed=a.seed)
        wl2.args.whole_arm_cost = bool(a.whole_arm)
        al2, st2, go2 = initial_alpha(wl2, traj, a.seed)
        B2 = len(al2)
        a2 = torch.as_tensor(al2, device=dev); s2 = torch.as_tensor(st2, device=dev).contiguous(); g2 = torch.as_tensor(go2, device=dev).contiguous()
        ms2, st_last = [], None
        for i in range(1 + 3):
            buf2 = a2.clone(); fs2, is2 = eng.new_state(B2)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.optimize_device(buf2, s2, g2, fs2, is2); e1.record()
            torch.cuda.synchronize()
            if i >= 1:
                ms2.append(e0.elapsed_time(e1)); st_last = is2.cpu().numpy()
        t2 = sum(ms2) / len(ms2) * 1e-3
        fl2 = flops_total(wl2.mode, T, len(wl2.obstacles), st_last[:, backend.I_INNER_TOTAL], st_last[:, backend.I_CAND_EVALS],
                          np.maximum(1, st_last[:, backend.I_OUTER] + st_last[:, backend.I_FULFILLED]), joints)
        saturated = {"trajectories_per_gpu": int(B2), "steps": len(ms2), "ms_per_step": 1e3 * t2, "value": B2 / t2, "unit": UNIT,
                     "fgd_iters_per_s": float(st_last[:, backend.I_INNER_TOTAL].sum()) / t2,
                     "roofline_frac": fl2 / t2 * 1e-12 / peak if peak else None,
                     "note": "same workload and kernel, batch large enough to hide the ragged-convergence tail"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * total_s / a.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(wl, traj, B, launch=h.launch_geometry(B),
                                  math="strict" if a.strict_math else "fast (rcp.approx)"),
            "fgd_iters_per_s": n_iter_all * a.steps / total_s,
            "mean_inner_iters": float(inner.mean()), "fulfilled_frac": float(is_[:, backend.I_FULFILLED].mean()),
            "wall_ms_per_step": 1e3 * (wall1 - wall0) / a.steps, "step_ms": [round(m, 4) for m in step_ms],
            "presoak_s": a.presoak_seconds,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}
    if saturated is not None:
        line["saturated"] = saturated

    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"], _, _ = cpu_baseline(wl, traj, alpha0, start, goal, a.cpu_seconds)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
