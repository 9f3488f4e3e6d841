#!/usr/bin/env python
"""Minimal driver for profilers: N forward+backward pairs on device-resident arrays, 1x1 grid.

  ncu --set full ... python tools/prof_pair.py --size 1024 --pairs 2 [--single] [--op fft]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import p3dfft_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, nargs="+", default=[1024])
ap.add_argument("--pairs", type=int, default=2)
ap.add_argument("--warm", type=int, default=0)
ap.add_argument("--single", action="store_true")
ap.add_argument("--opf", default="fft")
ap.add_argument("--opb", default="tff")
ap.add_argument("--cut", type=int, nargs=3, default=None)
a = ap.parse_args()
n = a.size * 3 if len(a.size) == 1 else a.size
nx, ny, nz = n
L = pb.load(a.single)
L.p3dfft_clean()
c = a.cut or (None, None, None)
L.p3dfft_setup((1, 1), nx, ny, nz, 0, *c)
_, _, isz = L.p3dfft_get_dims(1)
_, _, fsz = L.p3dfft_get_dims(2)
dt = torch.float32 if a.single else torch.float64
A = torch.rand(isz[0] * isz[1] * isz[2], dtype=dt, device="cuda")
F = torch.empty(2 * fsz[0] * fsz[1] * fsz[2], dtype=dt, device="cuda")
B = torch.empty_like(A)
import time
torch.cuda.synchronize()
for it in range(a.pairs + a.warm):
    if it == a.warm:
        L.set_timers()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
    L.p3dfft_ftran_r2c(A, F, a.opf)
    L.p3dfft_btran_c2r(F, B, a.opb)
torch.cuda.synchronize()
wall = (time.perf_counter() - w0) * 1e3 / a.pairs      # the calls are synchronous: includes what runs on the side stream
N = float(nx) * ny * nz
t = [x * 1e3 / a.pairs for x in L.get_timers()]
names = {4: "x_r2c", 6: "y_fwd", 7: "z_fwd", 8: "z_bwd", 9: "y_bwd", 11: "x_c2r"}
print(os.environ.get("TAG", ""), "roundtrip max err %.2e" % float((B / N - A).abs().max()), "pair ms %.3f (wall %.3f)" % (sum(t), wall),
      " ".join(f"{names[i]}={t[i]:.3f}" for i in sorted(names)))
L.p3dfft_clean()
