#!/usr/bin/env python
"""bench.py — trajectory-steps/s of the hybrid ODE-NN rollout on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload hybrid_fwd|mech_rk4]
    python bench.py --impl reference ...      # the CPU restatement of the reference path

A "step" is one pass of the hot path (one libhode rollout launch) over one batch of
synthetic 4GI-shaped trajectories.  One JSON line is printed by rank 0.  Definitions of
every reported quantity are in DESIGN.md §Measurement.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work per unit (BASELINE.md §4 / SURVEY.md §8d)
FLOP_MLP_EVAL = 26496.0          # 13 248 MAC
FLOP_ATTEMPT_HYBRID = 6 * FLOP_MLP_EVAL + 700.0   # DP5(4) attempt, FSAL: 158 976 + ~700
FLOP_STEP_RK4_MECH = 310.0
BYTES_PER_TRAJ = lambda T, n_in, per_row_t: 24 + 4 * T * n_in + (4 * T if per_row_t else 0) + 24 * T


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hybrid_fwd", choices=["hybrid_fwd", "hybrid_fwdbwd", "mech_rk4"])
    ap.add_argument("--traj-per-gpu", type=int, default=0)
    ap.add_argument("--precision", default="tf32x3", choices=["fp32", "tf32x3", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    return ap.parse_args()


WORKLOADS = {
    # configs/default.yaml-shaped: hybrid 64x4 net with a non-zero head, dopri5 1e-6/1e-8,
    # 262 144 trajectories (the config's whole cohort on ONE GPU; weak scaling over N)
    "hybrid_fwd": dict(B=262144, T=61, solver="dopri5", rtol=1e-6, atol=1e-8, kinks="clip",
                       nn=True, name="default.yaml hybrid 64x4 dopri5 rtol1e-6 fwd, 262144 traj/GPU, T=61"),
    # the same with the discrete adjoint (hode_rollout_bwd), same cohort per GPU.  (--traj-per-gpu 32768 is
    # configs/default.yaml's 262 144 trajectories strong-scaled over 8 GPUs: both kernels are then bound by
    # the latency of the longest trajectories — 221 trajectories per SM — rather than by throughput.)
    "hybrid_fwdbwd": dict(B=262144, T=61, solver="dopri5", rtol=1e-6, atol=1e-8, kinks="clip", nn=True, bwd=True,
                          name="default.yaml hybrid 64x4 dopri5 rtol1e-6 fwd+adjoint, 262144 traj/GPU, T=61"),
    # configs/ablation_no_nn.yaml-shaped: mechanistic only, RK4, 4 substeps per 5-min interval
    "mech_rk4": dict(B=1048576, T=61, solver="rk4", n_substeps=4, nn=False,
                     name="ablation_no_nn mechanistic rk4 4 substeps, 1048576 traj/GPU, T=61"),
}


def make_workload(kind: str, B: int, seed: int):
    from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import THETA_DEFAULT, cohort, random_mlp
    w = dict(WORKLOADS[kind])
    if B:
        w["B"] = B
    y0, t, ins = cohort(w["B"], w["T"], seed=seed, tvns=(kind != "mech_rk4"))
    W = random_mlp(64, 4, seed=1234, out_std=0.05) if w["nn"] else None
    w.update(y0=y0, t=t, ins=ins, theta=THETA_DEFAULT.copy(), W=W)
    return w


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "power_w_max": float(np.max(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------- CPU arm
def cpu_oracle_run(w, n_traj: int, threads: int):
    """Time the C restatement of the reference path on `n_traj` trajectories of workload w."""
    from oracle import cpu_oracle as o
    sl = slice(0, n_traj)
    ins = {k: v[sl] for k, v in w["ins"].items()}
    t0 = time.perf_counter()
    if w["solver"] == "rk4":
        _, st, cn, nfev = o.rollout(w["y0"][sl], w["t"], ins, w["theta"], w["W"], solver="rk4",
                                    n_substeps=w["n_substeps"], n_threads=threads)
    else:
        _, st, cn, nfev = o.rollout(w["y0"][sl], w["t"], ins, w["theta"], w["W"], solver="dopri5",
                                    rtol=w["rtol"], atol=w["atol"], kinks=w["kinks"], n_threads=threads)
    dt = time.perf_counter() - t0
    return float(cn[0].sum() + cn[1].sum()), dt


def cpu_sample_size(w, threads: int, target_s: float) -> int:
    probe = min(w["B"], 32 * threads)
    cpu_oracle_run(w, min(probe, 2 * threads), threads)      # load/compile, warm caches
    _, dt = cpu_oracle_run(w, probe, threads)
    n = int(probe * target_s / max(dt, 1e-3))
    return max(probe, min(w["B"], (n // 64) * 64 or 64))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    w = make_workload(args.workload, args.traj_per_gpu, seed=1000)
    from oracle import cpu_oracle as o
    o.build()
    n = cpu_sample_size(w, threads, target_s=6.0)
    for _ in range(args.warmup):
        cpu_oracle_run(w, min(n, 8 * threads), threads)
    steps_total, t_total = 0.0, 0.0
    for _ in range(args.steps):
        s, dt = cpu_oracle_run(w, n, threads)
        steps_total += s
        t_total += dt
    value = steps_total / t_total
    sample = (f"{n} of {w['B']} trajectories per step, C restatement of the reference path "
              f"(oracle/hode_oracle.c: float32 RHS, float64 SciPy-RK45 stepping), {threads} pthreads")
    line = {
        "impl": "reference", "metric": "trajectory_steps_per_sec", "value": value,
        "unit": "trajectory-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": w["name"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "trajectory-steps/s", "cores": threads,
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local: int):
    """Pin this rank's host threads (and with them the first-touch placement of its pinned buffers) to the
    NUMA node its GPU hangs off.  torchrun leaves ranks unbound; with 8 ranks moving 0.5 GB per step each
    through host memory the e2e leg is otherwise bound by cross-socket traffic.  Best effort: any failure
    leaves the affinity as it was.  Returns the node or None."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if not out:
            return None
        bus = out if out.count(":") == 2 and len(out.split(":")[0]) == 4 else out[-12:]   # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


# ---------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from hybrid_ode_for_glp_1_and_glucose_b200 import _lib, build, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libhode has no CPU fallback")
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    build.build()
    L = _lib.lib()

    w = make_workload(args.workload, args.traj_per_gpu, seed=1000 + rank)
    B, T = w["B"], w["T"]
    # host copies in pinned memory (the e2e path starts from these)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_y0, h_t, h_theta = pin(w["y0"]), pin(w["t"]), pin(w["theta"])
    h_ins = {k: pin(v) for k, v in w["ins"].items()}
    h_W = pin(w["W"]) if w["W"] is not None else None
    d_y0, d_t, d_theta = h_y0.to(dev), h_t.to(dev), h_theta.to(dev)
    d_ins = {k: v.to(dev) for k, v in h_ins.items()}
    d_W = h_W.to(dev) if h_W is not None else None
    kw = dict(solver=w["solver"], device=dev)
    if w["solver"] == "rk4":
        kw.update(n_substeps=w["n_substeps"])
    else:
        kw.update(rtol=w["rtol"], atol=w["atol"], kinks=w["kinks"], precision=args.precision)

    bwd = bool(w.get("bwd"))
    d_g = torch.full((B, T, 6), 1.0 / (B * T * 6), dtype=torch.float32, device=dev) if bwd else None
    grads = {}

    def step():
        if not bwd:
            return ops.rollout(d_y0, d_t, d_ins, d_theta, d_W, **kw)
        traj, info, tape = ops.rollout(d_y0, d_t, d_ins, d_theta, d_W, save_steps=True, **kw)
        g_y0, g_theta, g_W = ops.rollout_bwd(tape, d_g)      # grad of mean(traj) w.r.t. y0, theta, W
        if world > 1:
            # the path's one exchange step: gradients of the shared parameters + a loss slot,
            # one packed float32 buffer (54 KB), one NCCL all-reduce
            packed = torch.cat([g_theta.reshape(-1), g_W.reshape(-1), traj.new_zeros(1)])
            dist.all_reduce(packed, op=dist.ReduceOp.SUM)
            grads["packed"] = packed
        grads["out"] = (g_y0, g_theta, g_W)
        return traj, info

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        traj, info = step()
    barrier()
    attempts_per_step = float((info.n_accept.sum() + info.n_reject.sum()).item())
    assert bool((info.status == 0).all()), "synthetic workload must integrate without failures"

    # ---- kernel-resident timing: inputs already in HBM, K launches, CUDA events ------------
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        traj, info = step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    # kernels of ours per step: rollout (+ weight-image prep on the tensor-core path); the adjoint adds
    # two image preps, the sort-key and schedule kernels, the sweep and the partial-gradient reduction
    # (cub's radix-sort kernels are library code and are not counted)
    BWD_LAUNCHES = 6
    per_step = (2 if (w["nn"] and args.precision != "fp32") else 1) + (BWD_LAUNCHES if bwd else 0)
    launches = args.steps * per_step

    # ---- e2e: host buffers in, host result out, through the C ABI host entry --------------
    cfg, _ = ops.prepare(h_y0, h_t, h_ins, h_theta, h_W, 64, 4, torch.device("cpu"))
    cfg.solver = ops.SOLVERS[w["solver"]]
    cfg.n_substeps = w.get("n_substeps", 1)
    cfg.rtol, cfg.atol = w.get("rtol", 1e-6), w.get("atol", 1e-8)
    cfg.kink_mode = ops.KINKS[w.get("kinks", "clip")]
    if cfg.mlp != _lib.MLP_NONE:
        cfg.mlp = ops.PRECISIONS[args.precision]
    h_traj = torch.empty((B, T, 6), dtype=torch.float32).pin_memory()
    h_status = torch.empty(B, dtype=torch.int32).pin_memory()
    h_cnt = torch.empty((2, B), dtype=torch.int32).pin_memory()
    vp = lambda x: None if x is None else ctypes.c_void_p(x.data_ptr())
    stream = torch.cuda.current_stream(dev)

    def e2e_step():
        rc = L.hode_rollout_fwd_host(ctypes.byref(cfg), vp(h_y0), vp(h_t), vp(h_ins.get("meal")),
                                     vp(h_ins.get("tVNS")), vp(h_ins.get("GD")), vp(h_theta), vp(h_W),
                                     vp(h_traj), vp(h_status), vp(h_cnt),
                                     ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "hode_rollout_fwd_host")

    e2e_warm = max(args.warmup, 3)
    for _ in range(e2e_warm):
        e2e_step()          # warm-up: stream-ordered pool growth, first touch of the pinned buffers
    barrier()
    e2e_steps = max(3, min(args.steps, 5))
    e2e_each = []
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        e2e_step()          # synchronises its stream before returning
        e2e_each.append(1e3 * (time.perf_counter() - t1))
    barrier()
    e2e_s = time.perf_counter() - t0
    launches += (e2e_steps + e2e_warm) * (2 if (w["nn"] and args.precision != "fp32") else 1)
    e2e_attempts = float(h_cnt.sum().item())
    h2d = sum(x.numel() * 4 for x in [h_y0, h_t, h_theta] + list(h_ins.values()) + ([h_W] if h_W is not None else []))
    d2h = h_traj.numel() * 4 + h_status.numel() * 4 + h_cnt.numel() * 4

    # ---- the other legs of the metric, per GPU, short (skipped with --no-extra) -------------
    also = {}
    if not args.no_extra and w["nn"] and not bwd:
        def timed(fn, n):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                out = fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n, out
        Bx = min(B, 262144)
        sl = slice(0, Bx)
        x_ins = {k: v[sl].contiguous() for k, v in d_ins.items()}
        x_g = torch.full((Bx, T, 6), 1.0 / (Bx * T * 6), dtype=torch.float32, device=dev)

        def fwdbwd():
            _, info_x, tape = ops.rollout(d_y0[sl], d_t, x_ins, d_theta, d_W, save_steps=True, **kw)
            g = ops.rollout_bwd(tape, x_g)
            return info_x, g
        ms, (info_x, _) = timed(fwdbwd, 2)
        att = float((info_x.n_accept.sum() + info_x.n_reject.sum()).item())
        also["fwd_bwd"] = {"value": att / (ms * 1e-3), "unit": "trajectory-steps/s", "per": "GPU",
                           "trajectories": Bx, "ms_per_step": ms,
                           "what": "hode_rollout_fwd (3xTF32, steps recorded) + hode_rollout_bwd (tcgen05 discrete "
                                   "adjoint: 3xTF32 recomputation and delta chain, BF16x3 weight gradients; "
                                   "grad y0, theta[17], W[13510])",
                           "algorithmic_tflops": att * 3 * FLOP_ATTEMPT_HYBRID / (ms * 1e-3) / 1e12}
        launches += 3 * (2 + BWD_LAUNCHES)
        Sx = 8
        Bv = min(B, 131072)
        slv = slice(0, Bv)
        v_ins = {k: v[slv].contiguous() for k, v in d_ins.items()}
        rng = np.random.default_rng(7)
        thS = torch.from_numpy((w["theta"][None, :] * (1 + 0.02 * rng.normal(0, 1, (Sx, 17)))).astype(np.float32)).to(dev)
        WS = torch.from_numpy((w["W"][None, :] + 0.01 * rng.normal(0, 1, (Sx, w["W"].size))).astype(np.float32)).to(dev)

        def vi():
            return ops.vi_predictive(d_y0[slv], d_t, v_ins, thS, WS, **kw)
        ms, (_, _, info_v) = timed(vi, 1)
        att = float((info_v.n_accept.sum() + info_v.n_reject.sum()).item())
        also["vi_predictive"] = {"value": att / (ms * 1e-3), "unit": "trajectory-steps/s", "per": "GPU",
                                 "samples": Sx, "trajectories": Bv, "ms_per_step": ms,
                                 "what": "hode_vi_predictive: S parameter sets x B trajectories, mean/std "
                                         "reduced in-kernel (3xTF32)"}
        launches += 2 * 2

    # ---- reduce over ranks: MAX time, SUM work ----------------------------------------------
    stats = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
    work = torch.tensor([attempts_per_step, e2e_attempts], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    ms_total, e2e_s = stats.tolist()
    attempts_all, e2e_attempts_all = work.tolist()
    value = attempts_all * args.steps / (ms_total * 1e-3)
    e2e_value = e2e_attempts_all * e2e_steps / e2e_s

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_max = clocks.get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        ms_kernel = ms_total / args.steps          # one launch per step, nothing else on the stream
        if w["nn"]:
            flops = attempts_per_step * FLOP_ATTEMPT_HYBRID * (3.0 if bwd else 1.0)
            achieved = flops / (ms_kernel * 1e-3) / 1e12
            if args.precision == "fp32":
                peak = sm_count * 128 * 2 * sm_max * 1e6 / 1e12
                peak_src = (f"FP32 FMA pipe: {sm_count} SM x 128 lanes x 2 flop x {sm_max:.0f} MHz "
                            "(derived; MEASURED_PEAKS.json has no FP32 CUDA-core figure)")
                bound = "fp32"
            else:
                peak = float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2.0
                peak_src = ("TF32 dense = measured sustained cuBLAS bf16 / 2 (MEASURED_PEAKS.json "
                            "bf16_tflops_sustained; the kernel is timed inside a long step). achieved "
                            "counts ALGORITHMIC flops (158 976 + 700 per attempt); the 3xTF32 "
                            "emulation issues 3 tensor passes per algorithmic pass"
                            if peaks else "fallback 1400/2 TFLOP/s")
                bound = "tensor"
            # dram__bytes_read + dram__bytes_write of the dominant kernel from the committed ncu
            # --set full capture of this exact configuration (profiles/r01_rollout_tc_tf32x3_ncu_full.txt)
            traffic = 532.13e6 if (args.workload == "hybrid_fwd" and B == 262144 and args.precision == "tf32x3") else None
            roof = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic,
                    "algorithmic_bytes_per_launch": B * BYTES_PER_TRAJ(T, len(w["ins"]), False),
                    "peak_source": peak_src,
                    "kernel": ("rollout_simt_kernel<2>" if args.precision == "fp32" else "rollout_tc_kernel<x3,dopri5>")
                              + (" + rollout_bwd_tc_kernel" if bwd else ""),
                    "tensor_passes_per_algorithmic_pass": 3 if args.precision == "tf32x3" else 1,
                    "kernel_ms": ms_kernel,
                    "algorithmic_flop_per_launch": flops}
        else:
            nbytes = B * BYTES_PER_TRAJ(T, len(w["ins"]), False)
            achieved = nbytes / (ms_kernel * 1e-3) / 1e9
            peak = float(peaks.get("hbm_gbs", 6650.0))
            flops = attempts_per_step * FLOP_STEP_RK4_MECH
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                    "kernel": "rollout_simt_kernel<0>", "kernel_ms": ms_kernel,
                    "algorithmic_bytes_per_launch": nbytes,
                    "fp32_tflops": flops / (ms_kernel * 1e-3) / 1e12}
        line = {
            "metric": "trajectory_steps_per_sec", "value": value, "unit": "trajectory-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "mlp_arithmetic": args.precision if w["nn"] else "none",
                       "trajectories_total": B * world, "attempts_per_trajectory": attempts_per_step / B,
                       "l2_policy": "inputs+outputs per step exceed L2 (no flush needed): "
                                    f"{(B * BYTES_PER_TRAJ(T, len(w['ins']), False)) / 2**20:.0f} MiB"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "ms_each_rank0": [round(x, 2) for x in e2e_each],
                    "api": "hode_rollout_fwd_host (pinned host buffers in, host trajectories out)",
                    "host_numa_node_rank0": numa_node},
            "gpu_launches": launches,
            "roofline": roof,
            "trajectories_per_sec": B * world * args.steps / (ms_total * 1e-3),
        }
        if also:
            line["also"] = also
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n = cpu_sample_size(w, threads, target_s=12.0)
            s, dt = cpu_oracle_run(w, n, threads)
            line["cpu_baseline"] = {
                "value": s / dt, "unit": "trajectory-steps/s", "cores": threads, "kind": "port",
                "sample": f"{n} of {B} trajectories of the same workload, oracle/hode_oracle.c "
                          f"(float32 RHS + float64 SciPy-RK45 stepping), {threads} pthreads, {dt:.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)
