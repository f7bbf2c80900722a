#!/usr/bin/env python
"""Roofline denominators measured on the box (GPU): cuBLAS TF32 and BF16 dense matmul throughput, burst
(best of 10) and sustained (back to back for ~4 s) the way MEASURED_PEAKS.json was made, plus pinned
host<->device copy bandwidth of this rank (plain cudaMemcpyAsync of one large buffer per direction).

    python tools/measure_peaks.py [out.json]          # one GPU
    torchrun --nproc-per-node 8 tools/measure_peaks.py out.json --copies-only   # all ranks copy concurrently

bench.py reads profiles/r02_measured_peaks.json (a committed copy of this output) for the TF32 denominator.
"""
import json
import os
import subprocess
import sys
import threading
import time

import torch


def clocks_thread(stop, rows, idx):
    while not stop.is_set():
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                  "-i", str(idx)], capture_output=True, text=True, timeout=5).stdout.strip()
            f = [x.strip() for x in out.split(",")]
            rows.append((float(f[0]), float(f[1])))
        except Exception:
            pass
        time.sleep(0.1)


def matmul_tflops(dtype, tf32, n=8192, sustain_s=4.0, idx=0):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    flop = 2.0 * n ** 3
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flop / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        time.sleep(0.05)
    rows, stop = [], threading.Event()
    th = threading.Thread(target=clocks_thread, args=(stop, rows, idx), daemon=True)
    th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 0
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < sustain_s:
        for _ in range(20):
            torch.matmul(a, b, out=c)
        iters += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    th.join(timeout=2)
    sus = flop * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
    sm = sorted(r[0] for r in rows)
    return {"burst_tflops": best, "sustained_tflops": sus, "sm_mhz_median_sustained": sm[len(sm) // 2] if sm else None,
            "power_w_max": max((r[1] for r in rows), default=None)}


def copy_bandwidth(nbytes=1 << 30):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = {}
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        out[name + "_gbs"] = best

    # both directions at once on two streams
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    out["bidir_gbs_sum"] = 2 * nbytes / (time.perf_counter() - t0) / 1e9
    return out


def main():
    copies_only = "--copies-only" in sys.argv
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    res = {"gpu_name": torch.cuda.get_device_name(local), "torch": torch.__version__, "world": world,
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "host_cpus": os.cpu_count()}
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    cp = copy_bandwidth()
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([cp["h2d_gbs"], cp["d2h_gbs"], cp["bidir_gbs_sum"]], device="cuda", dtype=torch.float64)
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        # second pass: every rank copies at the same time (the e2e leg's situation)
        dist.barrier()
        cp2 = copy_bandwidth()
        t2 = torch.tensor([cp2["h2d_gbs"], cp2["d2h_gbs"], cp2["bidir_gbs_sum"]], device="cuda", dtype=torch.float64)
        allv2 = [torch.zeros_like(t2) for _ in range(world)]
        dist.all_gather(allv2, t2)
        res["pinned_copy_per_rank_concurrent"] = [dict(zip(("h2d_gbs", "d2h_gbs", "bidir_gbs_sum"), v.tolist())) for v in allv2]
        res["pinned_copy_sum_concurrent"] = {k: sum(r[k] for r in res["pinned_copy_per_rank_concurrent"])
                                             for k in ("h2d_gbs", "d2h_gbs", "bidir_gbs_sum")}
    else:
        res["pinned_copy"] = cp
    if not copies_only and rank == 0:
        res["tf32"] = matmul_tflops(torch.float32, True, idx=local)
        res["bf16"] = matmul_tflops(torch.bfloat16, False, idx=local)
        res["fp32_simt"] = matmul_tflops(torch.float32, False, n=4096, sustain_s=2.0, idx=local)
        res["how"] = ("torch.matmul 8192^3 (2 N^3 flop): best of 10 (burst) and back to back for 4 s (sustained); tf32 = float32 "
                      "inputs with torch.backends.cuda.matmul.allow_tf32; fp32_simt = float32 without TF32 (4096^3)")
    if rank == 0:
        s = json.dumps(res, indent=1)
        print(s)
        if args:
            open(args[0], "w").write(s + "\n")
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
