#!/usr/bin/env python
"""bench.py -- ms per r2c+c2r 3D FFT pair (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload (the metric's): 1024^3 double, forward `fft` + backward `tff`, grid 1x1 / 1x2 / 2x2 / 2x4.
Other BASELINE configurations (parity + measurement cases, not the headline line):
  --size 512                                   config 2
  --nx 2048 --ny 2048 --nz 2048 --dtype f32 --grid 1x8|2x4          config 4
  --nx 2048 --ny 512 --nz 513 --op cheby       config 5a (p3dfft_cheby + btran 'cff')
  --nx 2048 --ny 512 --nz 512 --op pruned      config 5b (2/3-rule pruned fft/tff + E(k) epilogue)

A "step" is one forward + one backward transform of a synthetic uniform random field through the C ABI
(libp3dfft[_single].so).  `value` is measured with the arrays resident in HBM (device pointers, CUDA events on
the library's stream, max over ranks); `e2e` is the same pair through the same entry points with HOST buffers,
i.e. with the PCIe copies inside the timed region.  Before the timed region the run checks itself against the
oracle (BASELINE config 1, 128^3, on this run's grid) and against a direct DFT of the timed field on a handful
of modes (`parity`); a failure exits non-zero.  Prints ONE JSON line on rank 0.
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


def workload_string(a):
    """The same string in both arms (the driver compares them)."""
    prec = "double" if a.dtype == "f64" else "single"
    size = f"{a.nx}^3" if a.nx == a.ny == a.nz else f"{a.nx}x{a.ny}x{a.nz}"
    what = {"fft": "r2c+c2r pair (forward fft + backward tff)",
            "cheby": "Chebyshev pair (p3dfft_cheby + backward cff)",
            "pruned": "pruned r2c+c2r pair (2/3 rule, forward fft + backward tff)"}[a.op]
    return f"{size} {prec} {what}"


def pruned_cut(a):
    return (2 * (a.nx // 3), 2 * (a.ny // 3), 2 * (a.nz // 3)) if a.op == "pruned" else (a.nx, a.ny, a.nz)


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


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU stage sequence, restated (oracle/), on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def run_reference(args, rank):
    """Reference arm: whole forward+backward pairs of the reference's P=1 stage sequence (restated with
    pocketfft; FFTW/MPI/Fortran are not in the image) on the host cores, thread count pinned explicitly.
    Every executed step is a complete pair on a complete field; when the requested size does not fit the
    time window, the largest power-of-two cube that does is timed whole and the line says so."""
    if rank != 0:
        return
    from oracle import p3dfft_oracle as po
    cores = cpu_cores()
    res = po.cpu_pair_measured(args.nx, args.ny, args.nz, dtype=args.dtype, op=args.op, workers=cores,
                               budget_s=args.ref_budget, max_steps=args.steps, cut=pruned_cut(args))
    ms = res["ms_per_pair"]
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus,
        "steps": res["steps"], "warmup": res["warmup"], "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": res["ms_per_executed_step"], "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload_string(args), "grid": [1, 1], "executed": res["executed"]},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port", "sample": res["sample"]},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "stage_s": res["stage_s"],
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# parity (checker code; outside every timed region)
# ------------------------------------------------------------------------------------------------
def parity_config1(pb, comm, dims, rank, torch, dist):
    """BASELINE config 1 (sample/C/driver_inverse.c size: 128^3 double) on THIS run's grid: every rank's slab of the
    forward and of the backward transform of one global Philox field against the oracle's slice."""
    import numpy as np
    from oracle import p3dfft_oracle as po      # checker only
    L = pb.load(False)
    n = 128
    L.p3dfft_setup(dims, n, n, n, comm)
    d = po.Decomp(n, n, n, dims, rank)
    G = po.philox_field(n, n, n)
    A = np.asfortranarray(G[po.local_in_slice(d)])
    tA = torch.from_numpy(A.ravel(order="F").copy()).cuda()
    ncplx = d.iisize * d.jjsize * d.nzc
    tF = torch.zeros(2 * ncplx, dtype=torch.float64, device="cuda")
    L.p3dfft_ftran_r2c(tA, tF, "fft")
    F = tF.cpu().numpy().view(np.complex128)
    e_f = po.rel_l2(F, np.asfortranarray(po.local_forward(G, d, "fft")).ravel(order="F"))
    Fg = po.global_forward(G, d, "fft")
    tFi = torch.from_numpy(np.asfortranarray(Fg[po.local_out_slice(d)]).ravel(order="F").view(np.float64).copy()).cuda()
    tB = torch.zeros(A.size, dtype=torch.float64, device="cuda")
    L.p3dfft_btran_c2r(tFi, tB, "tff")
    e_b = po.rel_l2(tB.cpu().numpy(), np.asfortranarray(po.local_backward(Fg, d, "tff")).ravel(order="F"))
    L.p3dfft_clean()
    e = torch.tensor([e_f, e_b], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
    return float(e[0]), float(e[1])


def spot_check_forward(A, F, info, n3, rank, torch, dist, nmodes=12):
    """Direct DFT of the TIMED field on a few modes: F[k] = sum_xyz A e^{-2 pi i (kx x/nx + ky y/ny + kz z/nz)},
    every rank summing over its own X pencil (fp64), partial sums added across ranks, compared with the value
    the rank that owns mode k holds in its Z pencil.  Independent of the transform code path (a matrix product and
    two weighted sums); catches any mis-routed block of the transposes at the full size.
    Returns max |F - F_direct| / ||A||_1   (the natural scale of the forward error of a DFT sum)."""
    nx, ny, nz = n3
    dev = A.device
    gen = torch.Generator().manual_seed(4242)
    kx = torch.randint(0, nx // 2 + 1, (nmodes,), generator=gen)
    ky = torch.randint(0, ny, (nmodes,), generator=gen)
    kz = torch.randint(0, nz, (nmodes,), generator=gen)
    kx[0] = ky[0] = kz[0] = 0                                   # DC
    kx[1], ky[1], kz[1] = nx // 2, ny // 2, nz // 2            # Nyquist corner
    kx[2], ky[2], kz[2] = 1, ny - 1, nz - 1                    # negative-frequency corner
    kx, ky, kz = kx.to(dev), ky.to(dev), kz.to(dev)
    f64 = torch.float64
    twopi = 2.0 * math.pi
    ji, kj = info.jisize, info.kjsize
    xs = torch.arange(nx, device=dev, dtype=torch.int64)
    ys = torch.arange(info.jistart - 1, info.jistart - 1 + ji, device=dev, dtype=torch.int64)
    zs = torch.arange(info.kjstart - 1, info.kjstart - 1 + kj, device=dev, dtype=torch.int64)

    def phase(idx, k, n):       # exp(-2 pi i idx k / n), reduced mod n in integers first
        m = (idx[:, None] * k[None, :]) % n
        ang = -twopi * m.to(f64) / n
        return torch.cos(ang), torch.sin(ang)
    wxr, wxi = phase(xs, kx, nx)
    A2 = A.view(kj * ji, nx).to(f64)
    tr = (A2 @ wxr).view(kj, ji, nmodes)
    ti = (A2 @ wxi).view(kj, ji, nmodes)
    wyr, wyi = phase(ys, ky, ny)
    ur = (tr * wyr[None] - ti * wyi[None]).sum(1)
    ui = (tr * wyi[None] + ti * wyr[None]).sum(1)
    wzr, wzi = phase(zs, kz, nz)
    direct = torch.stack([(ur * wzr - ui * wzi).sum(0), (ur * wzi + ui * wzr).sum(0)])       # [2][nmodes]
    l1 = A2.abs().sum().reshape(1)
    Fv = F.view(info.nzc, info.jjsize, info.iisize, 2)
    mine = torch.zeros(2, nmodes, dtype=f64, device=dev)
    for m in range(nmodes):
        ix, iy = int(kx[m]) - (info.iistart - 1), int(ky[m]) - (info.jjstart - 1)
        if 0 <= ix < info.iisize and 0 <= iy < info.jjsize:
            mine[:, m] = Fv[int(kz[m]), iy, ix].to(f64)
    if dist is not None:
        dist.all_reduce(direct)
        dist.all_reduce(mine)
        dist.all_reduce(l1)
    err = ((direct - mine) ** 2).sum(0).sqrt().max()
    return float(err / l1[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--nx", type=int, default=0)
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--nz", type=int, default=0)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--op", default="fft", choices=["fft", "cheby", "pruned"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=str, default="")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cufft", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-budget", type=float, default=100.0, help="seconds of CPU work the reference arm may spend")
    ap.add_argument("--cpu-budget", type=float, default=25.0, help="seconds of CPU work of the cpu_baseline leg")
    args = ap.parse_args()
    args.nx, args.ny, args.nz = args.nx or args.size, args.ny or args.size, args.nz or args.size
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
    single = args.dtype == "f32"
    rbytes = 4 if single else 8
    tdt = torch.float32 if single else torch.float64
    dims = tuple(int(x) for x in args.grid.split("x")) if args.grid else GRID_FOR.get(world, (1, world))
    assert dims[0] * dims[1] == world
    comms = {}
    for sp in sorted({False, single}):
        Lx = pb.load(sp)
        Lx.p3dfft_clean()
        comms[sp] = 0
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid = torch.frombuffer(bytearray(Lx.get_unique_id()), dtype=torch.uint8).cuda()
            dist.broadcast(uid, 0)
            comms[sp] = Lx.comm_create(rank, world, bytes(uid.cpu().numpy().tobytes()), local)
    L = pb.load(single)
    comm = comms[single]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity, part 1: BASELINE config 1 on this grid against the oracle (checker; untimed) ----------------
    parity = None
    if not args.no_parity:
        e_f, e_b = parity_config1(pb, comms[False], dims, rank, torch, dist)
        parity = {"config1_128cubed_fwd_rel_l2": e_f, "config1_128cubed_bwd_rel_l2": e_b, "grid": list(dims), "tol": 1e-12}

    nx, ny, nz = args.nx, args.ny, args.nz
    cut = pruned_cut(args)
    L.p3dfft_setup(dims, nx, ny, nz, comm, *cut)
    _, info = L.plan_steps(dims, nx, ny, nz, rank, False, "fft", nxc=cut[0], nyc=cut[1], nzc=cut[2])
    nreal = nx * info.jisize * info.kjsize
    ncplx = info.iisize * info.jjsize * info.nzc
    g = torch.Generator(device="cuda").manual_seed(20240229 + rank)
    A = torch.rand(nreal, dtype=tdt, device="cuda", generator=g)
    F = torch.empty(2 * ncplx, dtype=tdt, device="cuda")
    B = torch.empty(nreal, dtype=tdt, device="cuda")
    stream = torch.cuda.current_stream()
    L.set_stream(stream.cuda_stream)
    Lz = 2.0
    opb = "cff" if args.op == "cheby" else "tff"

    def forward(a, f):
        if args.op == "cheby":
            L.p3dfft_cheby(a, f, Lz)
        else:
            L.p3dfft_ftran_r2c(a, f, "fft")

    def backward(f, b):
        L.p3dfft_btran_c2r(f, b, opb)

    # ---- warm-up ---------------------------------------------------------------------------------
    L.set_async(False)
    for _ in range(args.warmup):
        forward(A, F)
        backward(F, B)
    ntot = float(nx) * ny * nz
    err = None
    if args.op == "fft":
        err = float((B / ntot - A).abs().max())

    # ---- parity, part 2: the timed field at full size ------------------------------------------------------------
    if parity is not None:
        tol = 1e-12 if not single else 1e-5
        if args.op == "fft":
            parity["timed_field_fwd_direct_dft_err_over_l1"] = spot_check_forward(A, F, info, (nx, ny, nz), rank, torch, dist)
            parity["timed_field_roundtrip_max_err"] = err
            parity["modes_checked"] = 12
        ok = parity["config1_128cubed_fwd_rel_l2"] <= 1e-12 and parity["config1_128cubed_bwd_rel_l2"] <= 1e-12
        if args.op == "fft":
            ok = ok and parity["timed_field_fwd_direct_dft_err_over_l1"] <= tol and err <= (1e-12 if not single else 1e-4)
        parity["pass"] = bool(ok)

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
        forward(A, F)
        backward(F, B)
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

    # ---- config 5b: the E(k) epilogue of driver_spec.c on the pruned spectrum (timed on its own) ----------------
    spectrum = None
    if args.op == "pruned":
        from math import sqrt
        kmax = int(sqrt(float(nx * nx + ny * ny + nz * nz)) * 0.5 + 0.5)
        E = torch.zeros(kmax + 1, dtype=torch.float64, device="cuda")
        forward(A, F)
        L.spectrum(F, kmax, 1.0 / ntot, out=E)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(5):
            L.spectrum(F, kmax, 1.0 / ntot, out=E)
        s1.record(stream)
        barrier()
        spectrum = {"ms": s0.elapsed_time(s1) / 5, "bins": kmax + 1, "sum": float(E.sum())}
        backward(F, B)

    # ---- e2e: same pair with host buffers through the same C ABI -------------------
    e2e = None
    if not args.no_e2e:
        hA = torch.empty(nreal, dtype=tdt).pin_memory()
        hF = torch.empty(2 * ncplx, dtype=tdt).pin_memory()
        hB = torch.empty(nreal, dtype=tdt).pin_memory()
        hA.copy_(A)
        for _ in range(2):
            forward(hA, hF)
            backward(hF, hB)
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            forward(hA, hF)      # synchronous: returns with hF complete on the host
            backward(hF, hB)
        torch.cuda.synchronize()
        w = torch.tensor([(time.perf_counter() - w0) * 1e3 / args.e2e_steps], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        err_e2e = float((hB / ntot - hA).abs().max()) if args.op == "fft" else None
        e2e = {"value": float(w), "unit": "ms", "h2d_bytes_per_step": int((nreal + 2 * ncplx) * rbytes),
               "d2h_bytes_per_step": int((2 * ncplx + nreal) * rbytes), "steps": args.e2e_steps,
               "roundtrip_max_err": err_e2e,
               "note": "host pinned buffers through the same entry points; wall clock, max over ranks"}
        del hA, hF, hB
        # pageable host arrays (what the reference's drivers pass: malloc, driver_sine.c:144-146)
        try:
            pA = np.empty(nreal, dtype=np.float32 if single else np.float64)
            pF = np.empty(2 * ncplx, dtype=pA.dtype)
            pB = np.empty(nreal, dtype=pA.dtype)
            pA[:] = 0.25
            forward(pA, pF)
            backward(pF, pB)
            barrier()
            w0 = time.perf_counter()
            forward(pA, pF)
            backward(pF, pB)
            torch.cuda.synchronize()
            w = torch.tensor([(time.perf_counter() - w0) * 1e3], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(w, op=dist.ReduceOp.MAX)
            e2e["pageable_ms"] = float(w)
            del pA, pF, pB
        except Exception as ex:      # noqa: BLE001 - a side line must never fail the bench
            e2e["pageable_ms"] = None
            e2e["pageable_error"] = repr(ex)[:200]

    # ---- cuFFT, comparison only (north_star: "cuFFT is reported only as a comparison"); never on the product path
    cufft = None
    if world == 1 and not args.no_cufft and args.op == "fft":
        try:
            x = A.view(nz, ny, nx)                      # same bytes, C order: a [z][y][x] array with x fastest
            for _ in range(2):
                y = torch.fft.rfftn(x)
                z = torch.fft.irfftn(y, s=(nz, ny, nx))
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            for _ in range(3):
                y = torch.fft.rfftn(x)
                z = torch.fft.irfftn(y, s=(nz, ny, nx))
            c1.record(stream)
            torch.cuda.synchronize()
            cufft = {"ms_per_pair": c0.elapsed_time(c1) / 3, "what": "torch.fft.rfftn + irfftn (cuFFT, out of place, includes its 1/N scaling)"}
            del y, z
        except Exception as e:      # noqa: BLE001 - a comparison line must never fail the bench
            cufft = {"unavailable": repr(e)[:200]}
        torch.cuda.empty_cache()

    p2p_on = L.p2p_active()
    variant = L.variant_string() if hasattr(L, "variant_string") else ""
    L.p3dfft_clean()
    L.reset_stream()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant stage kernel ------------------------------------------------
    peak, peak_src = load_peaks()
    sb = stage_bytes(info, rbytes)
    # timers (1-based): 5 X r2c, 7 Y fwd, 8 Z fwd, 9 Z bwd, 10 Y bwd, 12 X c2r
    stage_t = {"x_r2c": (tm[4], sb["x"]), "y_fwd": (tm[6], sb["y"]), "z_fwd": (tm[7], sb["z"]),
               "z_bwd": (tm[8], sb["z"]), "y_bwd": (tm[9], sb["y"]), "x_c2r": (tm[11], sb["x"])}
    dom = max(stage_t, key=lambda k: stage_t[k][0])
    dt, db = stage_t[dom]
    kname = {"x_r2c": "xr2c_kernel", "x_c2r": "xc2r_kernel"}.get(dom, "cstage kernel") + f" ({dom})"
    traffic = None      # DRAM bytes per launch of that kernel from the committed ncu --set full capture (same workload only)
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if world == 1 and (nx, ny, nz) == (1024, 1024, 1024) and not single and args.op == "fft":
            traffic = tj["per_launch"][dom]["dram_bytes"]
            kname = tj["per_launch"][dom]["kernel"] + f" ({dom})"      # the instantiation ncu saw for this stage
    except Exception:
        traffic = None
    ach = db / dt / 1e9 if dt > 0 else 0.0
    hbm_pair = 2 * (sb["x"] + sb["y"] + sb["z"])
    M1, M2 = dims
    c = 2 * rbytes
    nvl_pair = 2 * (info.nxhpc * info.jisize * info.kjsize * c * (M1 - 1) / M1 + info.iisize * info.nyc * info.kjsize * c * (M2 - 1) / M2)
    roof_ms = max(hbm_pair / (peak * 1e9), nvl_pair / 900e9) * 1e3
    line = {
        "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic",
        "config": {"workload": workload_string(args), "grid": [M1, M2], "cut": list(cut) if args.op == "pruned" else None,
                   "l2": f"per-rank arrays of {nreal * rbytes / 2**30:.2f} GiB exceed the 126 MB L2 (no flush needed)"},
        "gflops_5NlogN": 2 * 5 * ntot * math.log2(ntot) / (ms * 1e-3) / 1e9,
        "roofline_pair_ms": roof_ms, "roofline_pair_frac": roof_ms / ms,
        "roofline_pair_bound": "nvlink (900 GB/s/dir nominal)" if nvl_pair / 900e9 > hbm_pair / (peak * 1e9) else "hbm",
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
    if world > 1:
        line["roofline"]["note"] = ("N > 1: the stages that feed a transpose are bound by their NVLink peer stores, not by HBM (see "
                                    "roofline_pair_*); consumer chunks of the pipelined group run on a side stream and are not in stages_ms")
    if variant:
        line["library_variant"] = variant
    if parity is not None:
        line["parity"] = parity
    if spectrum:
        line["spectrum_epilogue"] = spectrum
    if cufft:
        line["cufft_comparison"] = cufft
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu and world == 1:             # (rank 0 at N = 1 only: the reference arm carries the CPU number at every N)
        from oracle import p3dfft_oracle as po     # CPU baseline leg (checker code, never on the GPU path)
        cores = cpu_cores()
        res = po.cpu_pair_measured(nx, ny, nz, dtype=args.dtype, op=args.op, workers=cores, budget_s=args.cpu_budget,
                                   max_steps=1, cut=cut, warmup=0)
        line["cpu_baseline"] = {"value": res["ms_per_pair"], "unit": "ms", "cores": cores, "kind": "port", "sample": res["sample"]}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    if parity is not None and not parity["pass"]:
        sys.exit(3)


if __name__ == "__main__":
    main()
