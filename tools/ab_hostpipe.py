#!/usr/bin/env python
"""A/B of the pageable-host-array path (ring of page-locked chunks filled by copy threads, api.cpp HostPipe) on one GPU:
one 1024^3 double pair per variant from ordinary numpy arrays, against the page-locked time of the same process.

  python tools/ab_hostpipe.py [--size 1024] [--variants "8:32768;16:32768;32:65536"]      (copy threads : chunk KB)
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import p3dfft_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--variants", default="8:32768;16:32768;24:32768;32:32768;16:65536;16:16384")
a = ap.parse_args()
n = a.size
L = pb.load(False)
L.p3dfft_clean()
print("host cores", len(os.sched_getaffinity(0)), flush=True)


def pair(A, F, B, reps):
    L.p3dfft_ftran_r2c(A, F, "fft")
    L.p3dfft_btran_c2r(F, B, "tff")
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for _ in range(reps):
        L.p3dfft_ftran_r2c(A, F, "fft")
        L.p3dfft_btran_c2r(F, B, "tff")
    torch.cuda.synchronize()
    return (time.perf_counter() - w0) * 1e3 / reps


L.p3dfft_setup((1, 1), n, n, n, 0)
_, _, isz = L.p3dfft_get_dims(1)
_, _, fsz = L.p3dfft_get_dims(2)
nreal, ncplx = int(np.prod(isz)), int(np.prod(fsz))
hA = torch.empty(nreal, dtype=torch.float64).pin_memory()
hF = torch.empty(2 * ncplx, dtype=torch.float64).pin_memory()
hB = torch.empty(nreal, dtype=torch.float64).pin_memory()
hA.fill_(0.25)
print("page-locked pair ms %.1f" % pair(hA, hF, hB, 2), flush=True)
del hA, hF, hB
L.p3dfft_clean()
pA = np.full(nreal, 0.25)
pF = np.zeros(2 * ncplx)
pB = np.zeros(nreal)
for var in a.variants.split(";"):
    th, kb = var.split(":")
    os.environ["P3DFFT_B200_COPY_THREADS"] = th
    os.environ["P3DFFT_B200_COPY_CHUNK_KB"] = kb
    L.p3dfft_setup((1, 1), n, n, n, 0)
    ms = pair(pA, pF, pB, 2)
    err = float(np.abs(pB[:1 << 20] / float(n) ** 3 - 0.25).max())
    L.p3dfft_clean()          # releases the ring and its threads: the next setup reads the environment again
    print("threads %3s chunk %6s KB  pageable pair ms %.1f  (round trip %.1e)" % (th, kb, ms, err), flush=True)
