#!/usr/bin/env python
"""Times the routines beside the transform path on one GPU (1024^3 double by default): the fused normalisation,
the power-spectrum epilogue, the single-rank real-data transpose (a pure box copy) and p3dfft_ftran_r2c_1d.
Wall clock around the library's synchronous calls on device-resident arrays, best of `reps`.  One JSON line."""
import json
import sys
import time

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import p3dfft_b200 as pb


def best(f, reps=5):
    f()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    L = pb.load(False)
    L.p3dfft_clean()
    L.p3dfft_setup((1, 1), n, n, n, 0)
    nxhp = n // 2 + 1
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.rand(n * n * n, dtype=torch.float64, device="cuda", generator=g)
    F = torch.empty(2 * nxhp * n * n, dtype=torch.float64, device="cuda")
    B = torch.empty_like(A)
    out = {"n": n, "dtype": "f64"}
    out["fwd_ms"] = best(lambda: L.p3dfft_ftran_r2c(A, F, "fft"))
    out["bwd_ms"] = best(lambda: L.p3dfft_btran_c2r(F, B, "tff"))
    L.set_scale(1.0 / n ** 3, 1.0 / n ** 3)
    out["fwd_scaled_ms"] = best(lambda: L.p3dfft_ftran_r2c(A, F, "fft"))
    out["bwd_scaled_ms"] = best(lambda: L.p3dfft_btran_c2r(F, B, "tff"))
    L.set_scale(1.0, 1.0)
    kmax = int((3 * n * n) ** 0.5 * 0.5 + 0.5)
    E = torch.zeros(kmax + 1, dtype=torch.float64, device="cuda")
    out["spectrum_ms"] = best(lambda: L.spectrum(F, kmax, 1.0 / n ** 3, out=E))
    out["spectrum_GBs"] = 16.0 * nxhp * n * n / out["spectrum_ms"] / 1e6
    out["rtran_x2y_ms"] = best(lambda: L.rtran("x2y", A, B))
    out["rtran_GBs"] = 2 * 8.0 * n ** 3 / out["rtran_x2y_ms"] / 1e6
    out["r2c_1d_ms"] = best(lambda: L.p3dfft_ftran_r2c_1d(A, F))
    out["r2c_1d_GBs"] = (8.0 * n ** 3 + 16.0 * nxhp * n * n) / out["r2c_1d_ms"] / 1e6
    L.p3dfft_clean()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
