#!/usr/bin/env python
"""One forward + backward transform through the emulated library (CPU) under the stream policy and the library switches of
the environment (P3D_EMU_STREAMS, P3DFFT_B200_XYPIPE ...): prints the errors against the oracle, exit code 0 iff both are
below 1e-12.  Run once per policy by tests/test_emulated_library.py::test_two_stream_executor_under_every_stream_order --
the mock runtime reads its policy once per process (tests/emu/emu_streams.inc)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import p3dfft_b200 as pb  # noqa: E402
from oracle import p3dfft_oracle as po  # noqa: E402


def main():
    n = tuple(int(x) for x in sys.argv[1:4])
    path = os.environ.get("P3D_EMU_LIB") or os.path.join(ROOT, "tests", "emu", "lib", "libp3dfft_emu.so")
    lib = pb.P3DFFT(False, path=path)
    d = po.Decomp(*n, (1, 1), 0)
    A = np.asfortranarray(np.random.default_rng(5).random(n))
    exp = po.local_forward(A, d, "fft")
    lib.p3dfft_setup((1, 1), *n, 0)
    worst = 0.0
    for _ in range(2):                # twice: the second call reuses streams, events and ring slots
        F = np.zeros(exp.shape, dtype=np.complex128, order="F")
        lib.p3dfft_ftran_r2c(A, F, "fft")
        e1 = po.rel_l2(F, exp)
        B = np.zeros(n, order="F")
        lib.p3dfft_btran_c2r(F, B, "tff")
        e2 = float(np.max(np.abs(B / A.size - A)))
        worst = max(worst, e1, e2)
    lib.p3dfft_clean()
    print(f"worst error {worst:.2e}", "ok" if worst <= 1e-12 else "WRONG")
    sys.exit(0 if worst <= 1e-12 else 1)


if __name__ == "__main__":
    main()
