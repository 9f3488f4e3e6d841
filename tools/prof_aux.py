#!/usr/bin/env python
"""One launch each of the kernels beside the transform path, for an ncu capture:
   ncu --set full --clock-control none -k regex:'rcopy_kernel|spectrum_kernel' -o gpurun_out/aux python tools/prof_aux.py 1024"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import p3dfft_b200 as pb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = pb.load(False)
L.p3dfft_clean()
L.p3dfft_setup((1, 1), n, n, n, 0)
A = torch.rand(n * n * n, dtype=torch.float64, device="cuda")
F = torch.empty(2 * (n // 2 + 1) * n * n, dtype=torch.float64, device="cuda")
B = torch.empty_like(A)
L.p3dfft_ftran_r2c(A, F, "fft")
kmax = int((3 * n * n) ** 0.5 * 0.5 + 0.5)
E = torch.zeros(kmax + 1, dtype=torch.float64, device="cuda")
L.spectrum(F, kmax, 1.0 / n ** 3, out=E)
L.rtran("x2y", A, B)
L.p3dfft_clean()
print("done")
