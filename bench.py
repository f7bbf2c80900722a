#!/usr/bin/env python
"""bench.py — trajectory-steps/s of the hybrid ODE-NN rollout and gradient path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload hybrid_fwd|hybrid_fwdbwd|vi_predictive|mech_rk4]
    python bench.py --impl reference ...      # the CPU restatement of the reference path

A "step" is one pass of the hot path over one batch of synthetic 4GI-shaped trajectories.  One JSON line is
printed by rank 0.  The headline (`value`, `e2e`, `roofline`) is the workload named in config.workload; the
default run (hybrid_fwd) also times the other legs of the metric and keeps them under `roofline.legs`:
forward + discrete adjoint (+ the packed NCCL gradient all-reduce when N > 1, inside the timed region), the
same on config 3's real per-GPU shard, and the 64-sample x 1 048 576-trajectory posterior-predictive sweep of
config 4.  Definitions of every reported quantity are in DESIGN.md §Measurement.
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
VI_SAMPLES, VI_TRAJ_PER_BOX = 64, 1048576     # config 4 (configs/4gi_vi.yaml-shaped posterior predictive)
REC_CAPACITY = 128                            # accepted-step records per trajectory (bench cohort: max ~100)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hybrid_fwd",
                    choices=["hybrid_fwd", "hybrid_fwdbwd", "vi_predictive", "mech_rk4"])
    ap.add_argument("--traj-per-gpu", type=int, default=0)
    ap.add_argument("--precision", default="f16bf16x2", choices=["fp32", "tf32x3", "tf32", "tf32bf16", "tf32x2bf16", "f16bf16x2"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs of the default run")
    ap.add_argument("--no-e2e", action="store_true",
                    help="skip the host-buffer e2e leg (PROFILING RUNS ONLY: under ncu's serialisation the streamed host entry "
                         "waits seconds per launch for its own copies); the printed line then carries e2e = null")
    ap.add_argument("--arrival-order", action="store_true",
                    help="time every pass in arrival order (no launch-order hint from the previous pass)")
    return ap.parse_args()


WORKLOADS = {
    # configs/default.yaml-shaped: hybrid 64x4 net with a non-zero head, dopri5 1e-6/1e-8,
    # 262 144 trajectories (the config's whole cohort on ONE GPU; weak scaling over N)
    "hybrid_fwd": dict(B=262144, T=61, solver="dopri5", rtol=1e-6, atol=1e-8, kinks="clip", nn=True,
                       name="default.yaml hybrid 64x4 dopri5 rtol1e-6 fwd, {B} traj/GPU, T=61"),
    # the same with the discrete adjoint (hode_rollout_bwd) and, when N > 1, the packed gradient all-reduce
    "hybrid_fwdbwd": dict(B=262144, T=61, solver="dopri5", rtol=1e-6, atol=1e-8, kinks="clip", nn=True, bwd=True,
                          name="default.yaml hybrid 64x4 dopri5 rtol1e-6 fwd+adjoint(+allreduce), {B} traj/GPU, T=61"),
    # configs/4gi_vi.yaml-shaped posterior predictive: 64 parameter samples x 1 048 576 trajectories PER BOX
    # (B = 1 048 576 / N per GPU: the sweep is partitioned by trajectory, all samples of a trajectory on one GPU)
    "vi_predictive": dict(B=VI_TRAJ_PER_BOX, T=61, solver="dopri5", rtol=1e-6, atol=1e-8, kinks="clip", nn=True, vi=True,
                          name="4gi_vi.yaml posterior predictive 64 samples x {B} traj/GPU, hybrid 64x4 dopri5 rtol1e-6, T=61"),
    # configs/ablation_no_nn.yaml-shaped: mechanistic only, RK4, 4 substeps per 5-min interval
    "mech_rk4": dict(B=1048576, T=61, solver="rk4", n_substeps=4, nn=False,
                     name="ablation_no_nn mechanistic rk4 4 substeps, {B} traj/GPU, T=61"),
}


def make_workload(kind: str, B: int, seed: int, world: int = 1):
    from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import THETA_DEFAULT, cohort, random_mlp
    w = dict(WORKLOADS[kind])
    if B:
        w["B"] = B
    elif w.get("vi"):
        w["B"] = VI_TRAJ_PER_BOX // world
    y0, t, ins = cohort(w["B"], w["T"], seed=seed, tvns=(kind != "mech_rk4"))
    W = random_mlp(64, 4, seed=1234, out_std=0.05) if w["nn"] else None
    w.update(y0=y0, t=t, ins=ins, theta=THETA_DEFAULT.copy(), W=W, name=w["name"].format(B=w["B"]))
    return w


def vi_posterior_samples(theta, W, S, seed=7):
    """S draws from a diagonal Gaussian posterior shaped like the reference's (models/bayes.py:99-127): mean = the
    point parameters, std = 0.1 x prior std; ODE priors of configs/4gi_vi.yaml:26-33 (EC_50: the default 1.0),
    the other ODE parameters fixed; network prior std 0.1."""
    rng = np.random.default_rng(seed)
    prior = {0: 0.002, 1: 0.005, 2: 0.001, 5: 0.02, 6: 1.0, 8: 2.0, 9: 1.5, 10: 0.005}   # index into theta (include/hode.h)
    th = np.repeat(theta[None, :], S, 0).astype(np.float64)
    for i, sd in prior.items():
        th[:, i] += 0.1 * sd * rng.normal(0, 1, S)
    Ws = W[None, :] + 0.1 * 0.1 * rng.normal(0, 1, (S, W.size))
    return th.astype(np.float32), Ws.astype(np.float32)


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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = make_workload(args.workload, args.traj_per_gpu, seed=1000, world=world)
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
    what = "one parameter set of the sweep (the reference loops over the samples)" if w.get("vi") else "forward rollout"
    sample = (f"{n} of {w['B']} trajectories per step, {what}, C restatement of the reference path "
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
    NUMA node its GPU hangs off.  Best effort: any failure leaves the affinity as it was.  Returns the node or None."""
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


def load_json(path):
    try:
        return json.load(open(path))
    except Exception:
        return {}


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
    L.hode_launch_count.restype = ctypes.c_int64

    w = make_workload(args.workload, args.traj_per_gpu, seed=1000 + rank, world=world)
    B, T = w["B"], w["T"]
    pin = lambda a: torch.from_numpy(a).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max_sum(ms, work):
        """MAX time over ranks, SUM of work."""
        st = torch.tensor([ms], dtype=torch.float64, device=dev)
        wk = torch.tensor([work], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(st, op=dist.ReduceOp.MAX)
            dist.all_reduce(wk, op=dist.ReduceOp.SUM)
        return float(st.item()), float(wk.item())

    def timed(fn, steps, warmup):
        """W untimed + K timed calls of fn on torch's current stream, CUDA events, barrier + synchronize on both sides.
        Returns (ms for the K steps on this rank, last result, kernels of ours launched in the timed region)."""
        out = None
        for _ in range(warmup):
            out = fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = int(L.hode_launch_count())
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1), out, int(L.hode_launch_count()) - n0

    class Cohort:
        """One rank's device-resident inputs of a workload (+ pinned host copies for the e2e leg)."""
        def __init__(self, wl, keep_host=False):
            self.w = wl
            h = dict(y0=pin(wl["y0"]), t=pin(wl["t"]), theta=pin(wl["theta"]), ins={k: pin(v) for k, v in wl["ins"].items()},
                     W=pin(wl["W"]) if wl["W"] is not None else None)
            self.y0, self.t, self.theta = h["y0"].to(dev), h["t"].to(dev), h["theta"].to(dev)
            self.ins = {k: v.to(dev) for k, v in h["ins"].items()}
            self.W = h["W"].to(dev) if h["W"] is not None else None
            self.host = h if keep_host else None
            self.kw = dict(solver=wl["solver"], device=dev)
            if wl["solver"] == "rk4":
                self.kw.update(n_substeps=wl["n_substeps"])
            else:
                self.kw.update(rtol=wl["rtol"], atol=wl["atol"], kinks=wl["kinks"], precision=args.precision)

    def attempts_of(info):
        return float((info.n_accept.sum() + info.n_reject.sum()).item())

    def fwd_step(c, order=None):
        return lambda: ops.rollout(c.y0, c.t, c.ins, c.theta, c.W, order=order, **c.kw)

    def fwdbwd_step(c, d_g, order=None):
        def step():
            traj, info, tape = ops.rollout(c.y0, c.t, c.ins, c.theta, c.W, save_steps=True, max_saved_steps=REC_CAPACITY,
                                           order=order, **c.kw)
            g_y0, g_theta, g_W = ops.rollout_bwd(tape, d_g)      # grad of mean(traj) w.r.t. y0, theta, W
            if world > 1:
                # the path's one exchange step: gradients of the shared parameters + a loss slot,
                # one packed float32 buffer (54 KB), one NCCL all-reduce — INSIDE the timed region
                packed = torch.cat([g_theta.reshape(-1), g_W.reshape(-1), traj.new_zeros(1)])
                dist.all_reduce(packed, op=dist.ReduceOp.SUM)
            return traj, info
        return step

    def adaptive(make_step, c, *a):
        """A pass over a cohort that was integrated before hands its trajectories to the kernel longest first, in the
        order of the PREVIOUS pass's attempt counters (ops.launch_order: a device-side argsort, inside the timed step).
        Training and posterior sweeps re-integrate the same cohort every epoch / sample (reference
        train/train_hybrid.py:225-275); at 4.6 trajectories per lane the arrival-order schedule of the persistent
        kernel ends 30 % above its ideal makespan, the longest-first one 6 %.  The first pass of a
        cohort has no counters and runs in arrival order (leg fwd_first_pass)."""
        state = {"order": None}

        def step():
            out = make_step(c, *a, order=state["order"])()
            state["order"] = ops.launch_order(out[1])
            return out
        return step

    def vi_step(c, thS, WS):
        return lambda: ops.vi_predictive(c.y0, c.t, c.ins, thS, WS, **c.kw)

    KMODE = {"tf32x3": "x3", "tf32x2bf16": "mix3", "f16bf16x2": "h16", "tf32bf16": "mixed", "tf32": "tf32", "fp32": "fp32"}[args.precision]
    LAUNCH_ORDER_TEXT = ("longest first by the PREVIOUS pass's attempt counters (ops.launch_order: device-side argsort inside the "
                         "timed step); the first pass of a cohort runs in arrival order: roofline.legs.fwd_first_pass")
    peaks_box = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"))
    peaks_r2 = load_json(os.path.join(ROOT, "profiles", "r02_measured_peaks.json"))
    traffic_db = load_json(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"))
    tf32_peak = (peaks_r2.get("tf32") or {}).get("sustained_tflops")
    if tf32_peak:
        tensor_peak, tensor_src = float(tf32_peak), (
            "MEASURED cuBLAS TF32 dense, sustained (torch.matmul 8192^3 float32 with allow_tf32, back to back for 4 s; "
            f"burst {(peaks_r2['tf32']).get('burst_tflops', 0):.0f}): profiles/r02_measured_peaks.json, made by tools/measure_peaks.py "
            "on this pool's B200 the way MEASURED_PEAKS.json was made.  achieved counts ALGORITHMIC flops (158 976 + 700 per "
            "attempt); the split-precision MLP issues tensor_passes_per_algorithmic_pass TF32-equivalent passes per algorithmic pass")
    else:
        tensor_peak = float(peaks_box.get("bf16_tflops_sustained", 1400.0)) / 2.0
        tensor_src = "fallback: MEASURED_PEAKS.json bf16_tflops_sustained / 2 (no measured TF32 figure found)"
    fp32_peak = (peaks_r2.get("fp32_simt") or {}).get("sustained_tflops")

    def tensor_roof(attempts, ms_kernel, factor, kernel, passes=3):
        flops = attempts * FLOP_ATTEMPT_HYBRID * factor
        ach = flops / (ms_kernel * 1e-3) / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak,
                "kernel": kernel, "kernel_ms": ms_kernel, "algorithmic_flop_per_launch": flops,
                "tensor_passes_per_algorithmic_pass": passes}

    c = Cohort(w, keep_host=True)
    bwd, vi = bool(w.get("bwd")), bool(w.get("vi"))
    # adaptive launch order (see adaptive()): the tensor-core DP5(4) rollout only — fixed-step and FP32 launches are
    # one-thread-per-trajectory grids without a queue
    ADAPTIVE = bool(w["nn"]) and args.precision != "fp32" and w["solver"] != "rk4" and not vi and not args.arrival_order
    warm = max(args.warmup, 3)
    steps = args.steps
    if vi:
        thS_np, WS_np = vi_posterior_samples(w["theta"], w["W"], VI_SAMPLES)
        thS, WS = torch.from_numpy(thS_np).to(dev), torch.from_numpy(WS_np).to(dev)
        step = vi_step(c, thS, WS)
        warm, steps = max(1, min(args.warmup, 1)), max(1, min(args.steps, 3))   # one step = 64 x 1 M rollouts (~seconds)
    elif bwd:
        d_g = torch.full((B, T, 6), 1.0 / (B * T * 6), dtype=torch.float32, device=dev)
        step = adaptive(fwdbwd_step, c, d_g) if ADAPTIVE else fwdbwd_step(c, d_g)
    else:
        step = adaptive(fwd_step, c) if ADAPTIVE else fwd_step(c)

    # ---- headline: kernel-resident timing, inputs already in HBM, K steps, CUDA events ------------
    sampler = ClockSampler(local)
    sampler.start()
    ms_total, out, launches = timed(step, steps, warm)
    clocks = sampler.stop()
    info = out[-1]
    attempts_per_step = attempts_of(info)
    if not vi:
        assert bool((info.status == 0).all()), "synthetic workload must integrate without failures"
    ms_total, attempts_all = reduce_max_sum(ms_total, attempts_per_step)
    value = attempts_all * steps / (ms_total * 1e-3)

    # ---- e2e: host buffers in, host result out, through the C ABI host entry (forward workloads) ----
    e2e = None
    if args.no_e2e:
        pass
    elif not bwd and not vi:
        h = c.host
        cfg, _ = ops.prepare(h["y0"], h["t"], h["ins"], h["theta"], h["W"], 64, 4, torch.device("cpu"))
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

        # A cohort this entry integrated before is launched longest first (within 32 768-trajectory blocks) from the
        # previous call's counters — the host-side twin of the device leg's launch order: opts.prev_counters = the
        # counters buffer the previous call filled.  The first call over a cohort has no hint: e2e.first_call.
        hint = _lib.new_fwd_opts(prev_counters_ptr=h_cnt.data_ptr()) if (w["nn"] and args.precision != "fp32") else None

        def e2e_step(opts=hint):
            rc = L.hode_rollout_fwd_host_ex(ctypes.byref(cfg), None if opts is None else ctypes.byref(opts), vp(h["y0"]), vp(h["t"]),
                                            vp(h["ins"].get("meal")), vp(h["ins"].get("tVNS")), vp(h["ins"].get("GD")),
                                            vp(h["theta"]), vp(h["W"]), vp(h_traj), vp(h_status), vp(h_cnt),
                                            ctypes.c_void_p(stream.cuda_stream))
            _lib.check(rc, "hode_rollout_fwd_host_ex")

        first_each = []
        for _ in range(3):      # warm-up (stream-ordered pool growth, first touch of the pinned buffers) = calls WITHOUT a hint
            t1 = time.perf_counter()
            e2e_step(None)
            first_each.append(1e3 * (time.perf_counter() - t1))
        barrier()
        e2e_steps = max(3, min(args.steps, 5))
        e2e_each = []
        n0 = int(L.hode_launch_count())
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            t1 = time.perf_counter()
            e2e_step()          # synchronises its stream before returning
            e2e_each.append(1e3 * (time.perf_counter() - t1))
        barrier()
        e2e_s = time.perf_counter() - t0
        launches += int(L.hode_launch_count()) - n0
        e2e_s, e2e_attempts_all = reduce_max_sum(e2e_s, float(h_cnt.sum().item()))
        h2d = sum(x.numel() * 4 for x in [h["y0"], h["t"], h["theta"]] + list(h["ins"].values()) + ([h["W"]] if h["W"] is not None else []))
        d2h = h_traj.numel() * 4 + h_status.numel() * 4 + h_cnt.numel() * 4
        e2e = {"value": e2e_attempts_all * e2e_steps / e2e_s, "unit": "trajectory-steps/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_each_rank0": [round(x, 2) for x in e2e_each],
               "api": "hode_rollout_fwd_host_ex (pinned host buffers in, host trajectories out; inputs and results stream "
                      "behind / ahead of the one persistent launch)" + ("; opts.prev_counters = the previous call's counters" if hint else ""),
               "first_call": {"ms_rank0": round(first_each[-1], 2), "value": float(h_cnt.sum().item()) / (first_each[-1] * 1e-3),
                              "unit": "trajectory-steps/s", "what": "the same call without a launch-order hint (last warm-up call)"},
               "host_numa_node_rank0": numa_node}
        del h_traj
        # the same call with an output-state mask: only the glucose column travels back (what the reference's figures
        # and glucose metrics read, plots/plot_all.py:183; 1/6 of the D2H bytes)
        if w["nn"] and args.precision != "fp32":
            h_g = torch.empty((B, T, 1), dtype=torch.float32).pin_memory()
            opts = _lib.new_fwd_opts(out_state_mask=0b000001, prev_counters_ptr=h_cnt.data_ptr())

            def e2e_masked():
                rc = L.hode_rollout_fwd_host_ex(ctypes.byref(cfg), ctypes.byref(opts), vp(h["y0"]), vp(h["t"]), vp(h["ins"].get("meal")),
                                                vp(h["ins"].get("tVNS")), vp(h["ins"].get("GD")), vp(h["theta"]), vp(h["W"]),
                                                vp(h_g), vp(h_status), vp(h_cnt), ctypes.c_void_p(stream.cuda_stream))
                _lib.check(rc, "hode_rollout_fwd_host_ex")
            for _ in range(2):
                e2e_masked()
            barrier()
            n0 = int(L.hode_launch_count())
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_masked()
            barrier()
            m_s = time.perf_counter() - t0
            launches += int(L.hode_launch_count()) - n0
            m_s, m_att = reduce_max_sum(m_s, float(h_cnt.sum().item()))
            e2e["glucose_column_only"] = {"value": m_att * e2e_steps / m_s, "unit": "trajectory-steps/s", "h2d_bytes_per_step": h2d,
                                          "d2h_bytes_per_step": h_g.numel() * 4 + h_status.numel() * 4 + h_cnt.numel() * 4,
                                          "api": "hode_rollout_fwd_host_ex, out_state_mask = 0b000001, prev_counters"}
            del h_g
    elif bwd:
        # the gradient path's e2e: host inputs in (pinned), loss + gradients out to the host — what a training step moves
        h = c.host
        h_out = torch.empty(17 + w["W"].size + 1, dtype=torch.float32).pin_memory()
        obs = torch.zeros((B, T, 6), dtype=torch.float32, device=dev)

        def e2e_step():
            y0 = h["y0"].to(dev, non_blocking=True)
            ins = {k: v.to(dev, non_blocking=True) for k, v in h["ins"].items()}
            loss, _, g_theta, g_W, _, info_e = ops.data_loss_step(y0, h["t"].to(dev, non_blocking=True), ins,
                                                                  h["theta"].to(dev, non_blocking=True),
                                                                  h["W"].to(dev, non_blocking=True), obs,
                                                                  max_saved_steps=REC_CAPACITY, **c.kw)
            packed = torch.cat([g_theta.reshape(-1), g_W.reshape(-1), loss.reshape(-1)])
            if world > 1:
                dist.all_reduce(packed, op=dist.ReduceOp.SUM)
            h_out.copy_(packed, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return info_e
        for _ in range(2):
            info_e = e2e_step()
        barrier()
        e2e_steps = max(3, min(args.steps, 5))
        n0 = int(L.hode_launch_count())
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            info_e = e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
        launches += int(L.hode_launch_count()) - n0
        e2e_s, e2e_attempts_all = reduce_max_sum(e2e_s, attempts_of(info_e))
        h2d = sum(x.numel() * 4 for x in [h["y0"], h["t"], h["theta"], h["W"]] + list(h["ins"].values()))
        e2e = {"value": e2e_attempts_all * e2e_steps / e2e_s, "unit": "trajectory-steps/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": h_out.numel() * 4, "steps": e2e_steps,
               "api": "ops.data_loss_step = hode_loss_fused_fwd_bwd (pinned host inputs copied in, loss + packed gradients copied out)"}
        del obs
    else:
        # posterior-predictive sweep end to end: host inputs and the S parameter sets in, mean / std out to the host
        h = c.host
        h_mean = torch.empty((B, T, 6), dtype=torch.float32).pin_memory()
        h_std = torch.empty((B, T, 6), dtype=torch.float32).pin_memory()
        h_thS, h_WS = pin(thS_np), pin(WS_np)

        def e2e_step():
            y0 = h["y0"].to(dev, non_blocking=True)
            ins = {k: v.to(dev, non_blocking=True) for k, v in h["ins"].items()}
            mean, std, info_e = ops.vi_predictive(y0, h["t"].to(dev, non_blocking=True), ins, h_thS.to(dev, non_blocking=True),
                                                  h_WS.to(dev, non_blocking=True), **c.kw)
            h_mean.copy_(mean, non_blocking=True)
            h_std.copy_(std, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return info_e
        info_e = e2e_step()
        barrier()
        n0 = int(L.hode_launch_count())
        t0 = time.perf_counter()
        info_e = e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
        launches += int(L.hode_launch_count()) - n0
        e2e_s, e2e_attempts_all = reduce_max_sum(e2e_s, attempts_of(info_e))
        h2d = sum(x.numel() * 4 for x in [h["y0"], h["t"], h_thS, h_WS] + list(h["ins"].values()))
        e2e = {"value": e2e_attempts_all / e2e_s, "unit": "trajectory-steps/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 2 * h_mean.numel() * 4, "steps": 1,
               "api": "ops.vi_predictive = hode_vi_predictive (pinned host inputs + 64 parameter sets copied in, mean / std copied out)"}
        del h_mean, h_std

    # ---- the other legs of the metric (default run only), each timed like the headline ----------------
    legs = {}
    if not args.no_extra and args.workload == "hybrid_fwd" and not args.traj_per_gpu:
        # (1) forward + discrete adjoint on the same cohort, with the packed gradient all-reduce when N > 1
        d_g = torch.full((B, T, 6), 1.0 / (B * T * 6), dtype=torch.float32, device=dev)
        ms, o2, n_l = timed(adaptive(fwdbwd_step, c, d_g) if ADAPTIVE else fwdbwd_step(c, d_g), min(args.steps, 5), 2)
        launches += n_l
        ms, att = reduce_max_sum(ms, attempts_of(o2[-1]))
        k2 = min(args.steps, 5)
        leg = tensor_roof(att / world, ms / k2, 3.0, "rollout_tc_kernel<" + KMODE + ",dopri5> (steps recorded) + rollout_bwd_tc_kernel")
        legs["fwd_bwd"] = {"value": att * k2 / (ms * 1e-3), "unit": "trajectory-steps/s", "n_gpus": world, "scaling": "weak",
                           "trajectories_per_gpu": B, "ms_per_step": ms / k2, "steps": k2,
                           "launch_order": LAUNCH_ORDER_TEXT if ADAPTIVE else "arrival",
                           "collective": "packed NCCL all-reduce of grad theta[17] + grad W[13510] + loss slot inside the timed region"
                                         if world > 1 else "none (single GPU)",
                           "what": "hode_rollout_fwd (" + args.precision + ", accepted steps + stage derivatives recorded) + hode_rollout_bwd "
                                   "(tcgen05 discrete adjoint: grad y0, theta[17], W[13510])",
                           "roofline": {k: leg[k] for k in ("achieved", "peak", "frac", "unit")}}
        # (1b) the same passes WITHOUT the launch-order hint: what the first pass over a cohort costs, and (1c) with a
        #      hint that is one optimiser step old: the counters come from a pass whose network weights differ by 1 %
        if ADAPTIVE:
            for name, fn, fac in (("fwd_first_pass", fwd_step(c), 1.0), ("fwd_bwd_first_pass", fwdbwd_step(c, d_g), 3.0)):
                ms, o5, n_l = timed(fn, k2, 2)
                launches += n_l
                ms, att = reduce_max_sum(ms, attempts_of(o5[-1]))
                legs[name] = {"value": att * k2 / (ms * 1e-3), "unit": "trajectory-steps/s", "n_gpus": world, "scaling": "weak",
                              "trajectories_per_gpu": B, "ms_per_step": ms / k2, "steps": k2, "launch_order": "arrival (no counters yet)",
                              "frac": att / world * FLOP_ATTEMPT_HYBRID * fac / (ms / k2 * 1e-3) / 1e12 / tensor_peak}
            g = torch.Generator(device=dev); g.manual_seed(7)
            W_prev = c.W * (1.0 + 0.01 * torch.randn(c.W.shape, device=dev, generator=g))
            stale = ops.launch_order(ops.rollout(c.y0, c.t, c.ins, c.theta, W_prev, **c.kw)[1])
            ms, o6, n_l = timed(fwd_step(c, stale), k2, 2)
            launches += n_l
            ms, att = reduce_max_sum(ms, attempts_of(o6[-1]))
            legs["fwd_order_one_update_old"] = {
                "value": att * k2 / (ms * 1e-3), "unit": "trajectory-steps/s", "n_gpus": world, "scaling": "weak",
                "trajectories_per_gpu": B, "ms_per_step": ms / k2, "steps": k2,
                "launch_order": "longest first by the counters of a pass with every network weight changed by 1 % (N(0, 0.01) relative): "
                                "the hint a training epoch inherits from the previous one",
                "frac": att / world * FLOP_ATTEMPT_HYBRID / (ms / k2 * 1e-3) / 1e12 / tensor_peak}
            del W_prev, stale
        del d_g
        # (2) config 3 as BASELINE.json states it: 262 144 trajectories in total over the N GPUs (strong scaling);
        #     at N = 1 the real per-GPU shard of the 8-GPU job (32 768) is timed instead
        Bs = 262144 // world if world > 1 else 32768
        ws = make_workload("hybrid_fwdbwd", Bs, seed=2000 + rank, world=world)
        cs = Cohort(ws)
        d_gs = torch.full((Bs, T, 6), 1.0 / (Bs * T * 6), dtype=torch.float32, device=dev)
        # a small cohort is bound by its longest trajectories: hand them out first, in the order of the previous pass's
        # attempt counters (training re-integrates the same cohort every epoch)
        order = ops.launch_order(fwd_step(cs)()[1])
        for name, fn, fac in (("fwd", fwd_step(cs, order), 1.0), ("fwd_bwd", fwdbwd_step(cs, d_gs, order), 3.0),
                              ("fwd_unordered", fwd_step(cs), 1.0)):
            ms, o3, n_l = timed(fn, 5, 2)
            launches += n_l
            ms, att = reduce_max_sum(ms, attempts_of(o3[-1]))
            legs[f"config3_shard_{name}"] = {
                "value": att * 5 / (ms * 1e-3), "unit": "trajectory-steps/s", "n_gpus": world,
                "launch_order": "arrival" if name.endswith("unordered") else "longest first (previous pass's attempt counters)",
                "scaling": "strong (262 144 trajectories in total)" if world > 1 else "n/a (one 32 768-trajectory shard of the 8-GPU job)",
                "trajectories_per_gpu": Bs, "ms_per_step": ms / 5, "steps": 5,
                "frac": att / world * FLOP_ATTEMPT_HYBRID * fac / (ms / 5 * 1e-3) / 1e12 / tensor_peak}
        del cs, d_gs
        # (3) config 4: 64 posterior samples x 1 048 576 trajectories per box, mean / std reduced in the kernel
        wv = make_workload("vi_predictive", 0, seed=3000 + rank, world=world)
        cv = Cohort(wv)
        th_np, W_np = vi_posterior_samples(wv["theta"], wv["W"], VI_SAMPLES)
        thS, WS = torch.from_numpy(th_np).to(dev), torch.from_numpy(W_np).to(dev)
        ms, o4, n_l = timed(vi_step(cv, thS, WS), 1, 1)
        launches += n_l
        info_v = o4[-1]
        ms, att = reduce_max_sum(ms, attempts_of(info_v))
        n_fail = torch.tensor([float((info_v.status != 0).sum().item())], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(n_fail)
        leg = tensor_roof(att / world, ms, 1.0, "rollout_tc_kernel<" + KMODE + ",dopri5> (fused Welford mean/std)")
        legs["vi_predictive_64x1M"] = {
            "value": att / (ms * 1e-3), "unit": "trajectory-steps/s", "n_gpus": world, "scaling": "strong (1 048 576 trajectories per box)",
            "samples": VI_SAMPLES, "trajectories_per_gpu": wv["B"], "trajectories_total": wv["B"] * world, "ms_per_step": ms, "steps": 1,
            "failed_units": float(n_fail.item()),
            "what": "hode_vi_predictive: 64 parameter sets x B trajectories, unbiased mean / std over the samples reduced in the kernel; "
                    "posterior = point parameters with std 0.1 x prior (configs/4gi_vi.yaml priors, network prior std 0.1)",
            "roofline": {k: leg[k] for k in ("achieved", "peak", "frac", "unit")}}
        del cv, thS, WS
        # (5 below) ...
        # (4) the reference's Sobol sweep (plots/plot_all.py:124-224): 16 384 parameter sets x 1 trajectory, per-trajectory theta
        if rank == 0:
            from hybrid_ode_for_glp_1_and_glucose_b200 import HybridODENN, sensitivity
            mdl = HybridODENN(device=dev)
            with torch.no_grad():
                off = 0
                Wt = torch.from_numpy(w["W"])
                for _, p_ in mdl.nn_residual.named_parameters():
                    p_.copy_(Wt[off: off + p_.numel()].reshape(p_.shape)); off += p_.numel()
            names = ["a_GI", "k_I", "rho", "E_max", "V_max", "K_m", "k_L"]
            lo = np.array([0.008, 0.02, 0.002, 0.08, 7.0, 5.5, 0.015]); hi = np.array([0.012, 0.03, 0.004, 0.12, 11.0, 8.5, 0.025])
            smp = torch.from_numpy((lo + np.random.default_rng(5).uniform(0, 1, (16384, 7)) * (hi - lo)).astype(np.float32))
            sob = lambda: sensitivity.sobol_outputs(mdl, names, smp)
            sob(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = int(L.hode_launch_count())
            e0.record(); sob(); e1.record(); torch.cuda.synchronize()
            launches += int(L.hode_launch_count()) - n0
            att = attempts_of(mdl.last_info)
            # (5) one whole training step (reference train_epoch batch body) as ONE call: data + physics (20 stacked
            #     re-solves) + L2 loss, adjoint of the data term, clip, Adam — eager and as a CUDA-graph replay
            from hybrid_ode_for_glp_1_and_glucose_b200.training import FusedTrainer
            Bt = 32768
            wt_ = make_workload("hybrid_fwdbwd", Bt, seed=4000, world=1)
            tb = lambda a: torch.from_numpy(a).to(dev)
            batch = {"initial_state": tb(wt_["y0"]), "observations": tb(np.repeat(wt_["y0"][:, None, :], T, 1).copy()),
                     "time_points": tb(wt_["t"]), "external_inputs": {k: tb(v) for k, v in wt_["ins"].items()}}
            trn = FusedTrainer(mdl, lr=1e-4, data_gradient=True)
            e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for mode in ("eager", "cuda_graph"):
                fn = (lambda: trn.step(batch, 1.0, 1.0, True, 1.0, True)) if mode == "eager" else (lambda r=trn.capture(batch, 1.0, 1.0, True, 1.0): r(batch))
                fn(); torch.cuda.synchronize()
                n0 = int(L.hode_launch_count())
                t0_ = time.perf_counter()
                e2.record()
                for _ in range(3):
                    fn()
                e3.record(); torch.cuda.synchronize()
                wall = (time.perf_counter() - t0_) / 3
                launches += int(L.hode_launch_count()) - n0
                legs[f"train_step_{mode}"] = {"value": None, "unit": "ms", "ms_per_step": e2.elapsed_time(e3) / 3, "host_ms_per_step": 1e3 * wall,
                                              "trajectories": Bt, "physics_rows": 20 * Bt, "n_gpus": 1,
                                              "what": "hode_train_step: rollout + adjoint of the data term + 20 stacked physics re-solves + RHS + "
                                                      "RHS-VJP + L2 + clip + Adam on 32 768 trajectories, rank 0 only"}
            legs["sobol_16384_sets"] = {"value": att / (e0.elapsed_time(e1) * 1e-3), "unit": "trajectory-steps/s", "n_gpus": 1,
                                        "ms_per_step": e0.elapsed_time(e1), "parameter_sets": 16384,
                                        "what": "sensitivity.sobol_outputs: per-trajectory theta on the tensor-core rollout, rank 0 only"}

    if rank == 0:
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_max = clocks.get("sm_max_mhz") or peaks_box.get("sm_max_mhz") or 1965.0
        ms_step = ms_total / steps
        n_in = len(w["ins"])
        alg_bytes = B * BYTES_PER_TRAJ(T, n_in, False)
        if w["nn"] and args.precision != "fp32":
            kname = ("rollout_tc_kernel<" + KMODE + ",dopri5> (fused Welford mean/std)" if vi else
                     "rollout_tc_kernel<" + KMODE + ",dopri5>" + (" + rollout_bwd_tc_kernel" if bwd else ""))
            passes = {"tf32x3": 3, "tf32bf16": 2, "tf32x2bf16": 2.5, "f16bf16x2": 1.5, "tf32": 1}[args.precision]
            roof = tensor_roof(attempts_per_step, ms_step, 3.0 if bwd else 1.0, kname, passes)
            roof["peak_source"] = tensor_src
            bf16_peak = peaks_box.get("bf16_tflops_sustained") or (peaks_r2.get("bf16") or {}).get("sustained_tflops")
            if bf16_peak:
                # the same algorithmic rate against the box's dense BF16 figure (MEASURED_PEAKS.json): the f16bf16x2 mode
                # issues kind::f16 MMAs (3 FP16/BF16 passes per float32 product), so both denominators are shown
                roof["frac_of_bf16_peak"] = roof["achieved"] / float(bf16_peak)
                roof["bf16_peak"] = float(bf16_peak)
                roof["issued_frac_of_bf16_peak"] = roof["achieved"] * passes * 2.0 / float(bf16_peak) if args.precision == "f16bf16x2" else None
            key = f"{args.workload}:{B}:{args.precision}"
            roof["traffic"] = traffic_db.get(key, {}).get("dram_bytes")
            roof["traffic_source"] = traffic_db.get(key, {}).get("source")
            roof["algorithmic_bytes_per_launch"] = alg_bytes if not vi else B * (24 + 4 * T * n_in + 2 * 24 * T)
        elif w["nn"]:
            peak = float(fp32_peak) if fp32_peak else sm_count * 128 * 2 * sm_max * 1e6 / 1e12
            flops = attempts_per_step * FLOP_ATTEMPT_HYBRID
            ach = flops / (ms_step * 1e-3) / 1e12
            roof = {"bound": "fp32", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "kernel": "rollout_simt_kernel<2>", "kernel_ms": ms_step, "traffic": None,
                    "peak_source": "measured cuBLAS SGEMM without TF32 (profiles/r02_measured_peaks.json fp32_simt)" if fp32_peak
                                   else f"derived: {sm_count} SM x 128 lanes x 2 flop x {sm_max:.0f} MHz"}
        else:
            # mechanistic RK4: FP32-issue-bound (3 divisions and up to 3 pow per RHS), outputs only at the observation times
            peak = float(fp32_peak) if fp32_peak else sm_count * 128 * 2 * sm_max * 1e6 / 1e12
            flops = attempts_per_step * FLOP_STEP_RK4_MECH
            ach = flops / (ms_step * 1e-3) / 1e12
            roof = {"bound": "fp32", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "kernel": "rollout_simt_kernel<0>", "kernel_ms": ms_step,
                    "traffic": traffic_db.get(f"{args.workload}:{B}:fp32", {}).get("dram_bytes"),
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "hbm_gbs": alg_bytes / (ms_step * 1e-3) / 1e9, "hbm_peak_gbs": float(peaks_box.get("hbm_gbs", 6650.0)),
                    "peak_source": ("FP32 pipe: measured cuBLAS SGEMM without TF32 (profiles/r02_measured_peaks.json fp32_simt); the kernel is "
                                    "FP32-issue-bound (ncu: profiles/r02_rollout_simt_mech_ncu_full.txt), its HBM traffic is "
                                    "inputs + observation-time outputs only (hbm_gbs / hbm_peak_gbs)") if fp32_peak
                                   else f"derived: {sm_count} SM x 128 lanes x 2 flop x {sm_max:.0f} MHz"}
        if legs:
            roof["legs"] = legs
        line = {
            "metric": "trajectory_steps_per_sec", "value": value, "unit": "trajectory-steps/s",
            "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if vi else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "mlp_arithmetic": args.precision if w["nn"] else "none",
                       "trajectories_per_gpu": B, "trajectories_total": B * world,
                       "attempts_per_trajectory": attempts_per_step / B / (VI_SAMPLES if vi else 1),
                       "launch_order": LAUNCH_ORDER_TEXT if ADAPTIVE else "arrival (one launch per pass, no queue hint)",
                       "l2_policy": "inputs+outputs per step exceed L2 (no flush needed): "
                                    f"{alg_bytes / 2**20:.0f} MiB"},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches,
            "gpu_launches_how": "hode_launch_count(): every kernel libhode.so launched inside the timed regions (headline, e2e, legs)",
            "roofline": roof,
            "trajectories_per_sec": B * world * steps / (ms_total * 1e-3) * (VI_SAMPLES if vi else 1),
        }
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n = cpu_sample_size(w, threads, target_s=12.0)
            s, dt = cpu_oracle_run(w, n, threads)
            line["cpu_baseline"] = {
                "value": s / dt, "unit": "trajectory-steps/s", "cores": threads, "kind": "port",
                "sample": f"{n} of {B} trajectories of the same workload (forward rollout, one parameter set), oracle/hode_oracle.c "
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
