#!/usr/bin/env python
"""bench.py -- ms per r2c+c2r 3D FFT pair (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 1024] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one forward `fft` + one backward `tff` transform of a synthetic uniform random
nx*ny*nz double field through the C ABI (libp3dfft.so).  `value` is measured with the arrays
resident in HBM (device pointers, CUDA events on the library's stream, max over ranks);
`e2e` is the same pair through the same entry points with HOST (pinned) buffers, i.e. with the
PCIe copies inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GRID_FOR = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}
METRIC = "ms_per_r2c_c2r_pair"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.lines = []
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mxc = float(f[1]), float(f[2])
            except ValueError:
                continue
            mx = mxc
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            sm = [float(self.lines[-1][1].split(",")[1])] if self.lines else [0.0]
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def stage_bytes(info, r=8):
    """ALGORITHMIC HBM bytes per rank per transform, split per stage."""
    c = 2 * r
    x = info.nx * info.jisize * info.kjsize * r + info.nxhpc * info.jisize * info.kjsize * c
    y = info.iisize * info.ny * info.kjsize * c + info.iisize * info.nyc * info.kjsize * c
    z = info.iisize * info.jjsize * info.nz * c + info.iisize * info.jjsize * info.nzc * c
    return {"x": x, "y": y, "z": z}


def run_reference(args, rank):
    """Reference arm: the reference's own CPU stage sequence (restated; FFTW/MPI/Fortran are
    not in the image) on the host cores, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import p3dfft_oracle as po
    n = args.size
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    frac = max(8, n // 16)        # 1024 -> 1/64 of the lines of every stage per step
    vals = []
    desc = ""
    for i in range(args.warmup + args.steps):
        full, meas, desc = po.cpu_pair_sampled(n, n, n, frac, workers=cores)
        if i >= args.warmup:
            vals.append(full * 1e3)
    ms = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{n}^3 double r2c+c2r pair (forward fft + backward tff)", "grid": [1, 1]},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=str, default="")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cufft", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world > 1:
        args.gpus = world
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import p3dfft_b200 as pb

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = pb.load(False)
    L.p3dfft_clean()
    dims = tuple(int(x) for x in args.grid.split("x")) if args.grid else GRID_FOR.get(world, (1, world))
    assert dims[0] * dims[1] == world
    comm = 0
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(L.get_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        comm = L.comm_create(rank, world, bytes(uid.cpu().numpy().tobytes()), local)
    n = args.size
    L.p3dfft_setup(dims, n, n, n, comm)
    _, info = L.plan_steps(dims, n, n, n, rank, False, "fft")
    nreal = n * info.jisize * info.kjsize
    ncplx = info.iisize * info.jjsize * n
    g = torch.Generator(device="cuda").manual_seed(20240229 + rank)
    A = torch.rand(nreal, dtype=torch.float64, device="cuda", generator=g)
    F = torch.empty(2 * ncplx, dtype=torch.float64, device="cuda")
    B = torch.empty(nreal, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()
    L.set_stream(stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ---------------------------------------------------------------------------------
    L.set_async(False)
    for _ in range(args.warmup):
        L.p3dfft_ftran_r2c(A, F, "fft")
        L.p3dfft_btran_c2r(F, B, "tff")
    err = float((B / float(n) ** 3 - A).abs().max())

    # ---- timed region: K pairs, arrays resident in HBM --------------------------------------------
    # The calls are the reference's synchronous entry points; the library brackets every stage kernel
    # and every exchange with CUDA events on the stream it launches on, so the per-stage times used
    # for the roofline below come from THIS region (timers(12) of the reference, module.F90:106).
    L.set_timers()
    L.launch_count(True)
    L.fast_launch_count(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    t0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        L.p3dfft_ftran_r2c(A, F, "fft")
        L.p3dfft_btran_c2r(F, B, "tff")
    e1.record(stream)
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    tm = [t / args.steps for t in L.get_timers()]
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    launches = torch.tensor([L.launch_count(), L.fast_launch_count()], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    ms = float(ms_total) / args.steps

    # ---- e2e: same pair with host (pinned) buffers through the same C ABI -------------------
    e2e = None
    if not args.no_e2e:
        hA = torch.empty(nreal, dtype=torch.float64).pin_memory()
        hF = torch.empty(2 * ncplx, dtype=torch.float64).pin_memory()
        hB = torch.empty(nreal, dtype=torch.float64).pin_memory()
        hA.copy_(A)
        for _ in range(2):
            L.p3dfft_ftran_r2c(hA, hF, "fft")
            L.p3dfft_btran_c2r(hF, hB, "tff")
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            L.p3dfft_ftran_r2c(hA, hF, "fft")      # synchronous: returns with hF complete on the host
            L.p3dfft_btran_c2r(hF, hB, "tff")
        torch.cuda.synchronize()
        w = torch.tensor([(time.perf_counter() - w0) * 1e3 / args.e2e_steps], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        err_e2e = float((hB / float(n) ** 3 - hA).abs().max())
        e2e = {"value": float(w), "unit": "ms", "h2d_bytes_per_step": int((nreal + 2 * ncplx) * 8),
               "d2h_bytes_per_step": int((2 * ncplx + nreal) * 8), "steps": args.e2e_steps,
               "roundtrip_max_err": err_e2e,
               "note": "host pinned buffers through p3dfft_ftran_r2c/p3dfft_btran_c2r; wall clock, max over ranks"}
        del hA, hF, hB

    # ---- cuFFT, comparison only (north_star: "cuFFT is reported only as a comparison"); never on the product path
    cufft = None
    if world == 1 and not args.no_cufft:
        try:
            x = A.view(n, n, n)                      # same bytes, C order: a [z][y][x] array with x fastest
            for _ in range(2):
                y = torch.fft.rfftn(x)
                z = torch.fft.irfftn(y, s=(n, n, n))
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            for _ in range(3):
                y = torch.fft.rfftn(x)
                z = torch.fft.irfftn(y, s=(n, n, n))
            c1.record(stream)
            torch.cuda.synchronize()
            cufft = {"ms_per_pair": c0.elapsed_time(c1) / 3, "what": "torch.fft.rfftn + irfftn (cuFFT D2Z/Z2D, out of place, includes its 1/N scaling)"}
            del y, z
        except Exception as e:      # noqa: BLE001 - a comparison line must never fail the bench
            cufft = {"unavailable": repr(e)[:200]}
        torch.cuda.empty_cache()

    p2p_on = L.p2p_active()
    L.p3dfft_clean()
    L.reset_stream()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant stage kernel ------------------------------------------------
    peak, peak_src = load_peaks()
    sb = stage_bytes(info)
    # timers (1-based): 5 X r2c, 7 Y fwd, 8 Z fwd, 9 Z bwd, 10 Y bwd, 12 X c2r
    stage_t = {"x_r2c": (tm[4], sb["x"]), "y_fwd": (tm[6], sb["y"]), "z_fwd": (tm[7], sb["z"]),
               "z_bwd": (tm[8], sb["z"]), "y_bwd": (tm[9], sb["y"]), "x_c2r": (tm[11], sb["x"])}
    dom = max(stage_t, key=lambda k: stage_t[k][0])
    dt, db = stage_t[dom]
    # (the Z-backward stage of a 1024-point transform runs the split variant of the c2c kernel, fft_fast.cu)
    kname = {"x_r2c": "xr2c_kernel", "x_c2r": "xc2r_kernel", "z_bwd": "cstage_split_kernel" if n == 1024 else "cstage_kernel"}.get(
        dom, "cstage_kernel") + f" ({dom})"
    traffic = None      # DRAM bytes per launch of that kernel from the committed ncu --set full capture (same workload only)
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if world == 1 and n == 1024:
            traffic = tj["per_launch"][dom]["dram_bytes"]
    except Exception:
        traffic = None
    ach = db / dt / 1e9 if dt > 0 else 0.0
    hbm_pair = 2 * (sb["x"] + sb["y"] + sb["z"])
    M1, M2 = dims
    c = 16
    nvl_pair = 2 * (info.nxhpc * info.jisize * info.kjsize * c * (M1 - 1) / M1 + info.iisize * info.nyc * info.kjsize * c * (M2 - 1) / M2)
    roof_ms = max(hbm_pair / (peak * 1e9), nvl_pair / 900e9) * 1e3
    ntot = float(n) ** 3
    line = {
        "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{n}^3 double r2c+c2r pair (forward fft + backward tff), {M1}x{M2} pencil grid",
                   "grid": [M1, M2], "l2": f"per-rank arrays of {nreal * 8 / 2**30:.2f} GiB exceed the 126 MB L2 (no flush needed)"},
        "gflops_5NlogN": 2 * 5 * ntot * math.log2(ntot) / (ms * 1e-3) / 1e9,
        "roofline_pair_ms": roof_ms, "roofline_pair_frac": roof_ms / ms,
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": db, "avg_launch_ms": dt * 1e3,
                     "stages_ms": {k: v[0] * 1e3 for k, v in stage_t.items()},
                     "exchange_ms": {"T1": tm[0] * 1e3, "T2": tm[1] * 1e3, "T3": tm[2] * 1e3, "T4": tm[3] * 1e3}},
        "gpu_launches": int(launches[0]), "gpu_launches_specialised_kernels": int(launches[1]),
        "transpose": ("none" if world == 1 else ("nvlink peer stores from the stage kernels + barrier" if p2p_on
                      else "grouped ncclSend/ncclRecv")),
        "clocks": clocks, "roundtrip_max_err": err,
    }
    if cufft:
        line["cufft_comparison"] = cufft
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu:
        from oracle import p3dfft_oracle as po     # CPU baseline leg (checker code, never on the GPU path)
        cores = len(os.sched_getaffinity(0))
        frac = max(8, n // 16)
        full, meas, desc = po.cpu_pair_sampled(n, n, n, frac, workers=cores)
        if meas < 5.0:       # aim for ~10-30 s of CPU work in total
            reps = min(6, int(10.0 / max(meas, 0.1)))
            vals = [po.cpu_pair_sampled(n, n, n, frac, workers=cores)[0] for _ in range(reps)]
            full = sum(vals) / len(vals)
        line["cpu_baseline"] = {"value": full * 1e3, "unit": "ms", "cores": cores, "kind": "port", "sample": desc}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
