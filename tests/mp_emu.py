#!/usr/bin/env python
"""One rank of a multi-rank run of the EMULATED library (CPU only; launched P times by tests/test_emulated_multirank.py).

The product's api.cpp -- process bootstrap, communicator split, peer mapping of the work buffers, the executor with its
exchange steps, barriers and write-after-read rule -- runs unchanged on the mock CUDA runtime and mock NCCL of tests/emu
(emu_api.cpp, emu_mp.inc): "device" buffers live in POSIX shared memory so the stage kernels' peer-to-peer stores really
land in another process, NCCL messages are files.  The cases and the checking code are those of the GPU driver
tests/mp_parity.py (array adapter EmuArrays), plus call sequences that exercise the executor's hazard rule.

  RANK, WORLD_SIZE, P3D_EMU_UID (hex unique id made by the launcher), P3D_EMU_SHM=1 in the environment;
  the library's own switches (P3DFFT_B200_P2P, _FLAGBAR, _OVERLAP, _PLAIN ...) are read at p3dfft_setup as usual.
Exit code 0 iff every case passes on this rank.
"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import p3dfft_b200 as pb  # noqa: E402
from oracle import p3dfft_oracle as po  # noqa: E402
from tests import mp_parity as M  # noqa: E402

SUITES = {
    # (n, cut, opf, opb, stride1, nv, single)
    "fast": [((64, 64, 64), None, "fft", "tff", False, 1, False),            # specialised kernels, blocked buffers
             ((64, 64, 64), (32, 32, 32), "fft", "tff", False, 1, False),
             ((64, 64, 64), None, "fft", "tff", False, 2, False),            # _many: the work buffers grow, peers are re-mapped
             ((128, 64, 64), None, "fft", "tff", False, 1, True)],
    # the headline lengths (1024-point X and Z stages: split kernel, L2 prefetch, blocked tile order) at a thin y extent
    "long": [((1024, 32, 1024), None, "fft", "tff", False, 1, False),
             ((2048, 16, 512), None, "fft", "tff", False, 1, True)],
    "long-light": [((1024, 16, 64), None, "fft", "tff", False, 1, False),
                   ((64, 16, 1024), None, "fft", "tff", False, 1, False),
                   ((64, 1024, 16), None, "fft", "tff", False, 1, False)],
    "mixed": [((32, 32, 32), None, "fft", "tff", False, 1, False),
              ((14, 26, 38), None, "fft", "tff", False, 1, False),           # the reference's uneven case
              ((32, 32, 33), None, "ffc", "cff", False, 1, False),
              ((32, 32, 32), None, "ffn", "nff", False, 1, False),
              ((64, 32, 48), None, "fft", "tff", True, 1, False),            # STRIDE1
              ((64, 64, 64), (42, 42, 42), "fft", "tff", False, 1, True)],
}


def fuzz_cases(seed, count, dims):
    """`count` random cases, the same on every rank: sizes that take the specialised kernels (64, 128) mixed with sizes that
    take the any-length kernel, pruned grids, every third-dimension variant, STRIDE1, two variables, both precisions"""
    import random
    rng = random.Random(seed)
    cases = []
    while len(cases) < count:
        nx = rng.choice([64, 128, 32, 14, 30, 36])
        ny, nz = rng.choice([64, 32, 26, 18, 12, 21, 48]), rng.choice([64, 32, 38, 20, 12, 35, 128])
        if (nx // 2 + 1) < dims[0] or ny < max(dims) or nz < dims[1]:
            continue
        opf, opb = rng.choice([("fft", "tff")] * 4 + [("ffc", "cff"), ("ffs", "sff"), ("ffn", "nff")])
        if opf == "ffc" and nz % 2 == 0:
            nz += 1
        cut = None
        if rng.random() < 0.3:
            cut = tuple(min(v, max(2 * max(dims), (v * 2 // 3) // 2 * 2)) for v in (nx, ny, nz))
            if opf != "fft":
                cut = (cut[0], cut[1], nz)
        cases.append(((nx, ny, nz), cut, opf, opb, rng.random() < 0.25, rng.choice([1, 1, 1, 2]), rng.random() < 0.25))
    return cases


def emulib(single):
    path = os.path.join(ROOT, "tests", "emu", "lib", "libp3dfft_emu_single.so" if single else "libp3dfft_emu.so")
    if not single and os.environ.get("P3D_EMU_LIB"):      # mutation checks of the tests themselves (a deliberately broken build)
        path = os.environ["P3D_EMU_LIB"]
    return pb.P3DFFT(single, path=path)


def nccl_counts(L):
    out = (ctypes.c_longlong * 4)()
    L.lib.emu_nccl_counts(out)
    return list(out)      # sends, receives, all-reduces, all-gathers


def run_repeat(L, comm, dims, rank, n, X, asynchronous=False):
    """forward, forward, backward, backward on the same plan: on one-dimensional grids the second call of a pair stores
    into a receive buffer a slower peer may still be reading -- the executor's extra barrier (api.cpp run_plan) is what
    keeps the results right.  Ranks are desynchronised on purpose."""
    import time
    nx, ny, nz = n
    L.set_layout(False, False)
    L.p3dfft_setup(dims, nx, ny, nz, comm)
    d = po.Decomp(nx, ny, nz, dims, rank)
    _, _, isz = L.p3dfft_get_dims(1)
    _, _, fsz = L.p3dfft_get_dims(2)
    fields = [po.philox_field(nx, ny, nz, seed=500 + i) for i in range(3)]
    ins = [X.dev(np.asfortranarray(f[po.local_in_slice(d)]).ravel(order="F")) for f in fields]
    outs = [X.full(2 * int(np.prod(fsz)), 0.0, np.float64) for _ in fields]
    worst = 0.0
    L.set_async(asynchronous)       # asynchronous: the three calls are only enqueued (device arrays), one sync at the end
    for i in range(3):
        if rank == i % (dims[0] * dims[1]):
            time.sleep(0.05)
        L.p3dfft_ftran_r2c(ins[i], outs[i], "fft")
    L.sync()
    for i in range(3):
        exp = np.asfortranarray(po.local_forward(fields[i], d, "fft")).ravel(order="F")
        worst = max(worst, po.rel_l2(X.host(outs[i]).view(np.complex128), exp))
    backs = [X.full(int(np.prod(isz)), 0.0, np.float64) for _ in fields]
    for i in range(3):
        if rank == (i + 1) % (dims[0] * dims[1]):
            time.sleep(0.05)
        L.p3dfft_btran_c2r(outs[i], backs[i], "tff")
    L.sync()
    L.set_async(False)
    for i in range(3):
        worst = max(worst, float(np.abs(X.host(backs[i]) / (nx * ny * nz) - X.host(ins[i])).max()))
    L.p3dfft_clean()
    return worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", required=True)
    ap.add_argument("--suite", default="fast")
    ap.add_argument("--aux", action="store_true")
    ap.add_argument("--repeat", action="store_true")
    ap.add_argument("--async", dest="asynchronous", action="store_true")
    ap.add_argument("--expect-p2p", type=int, default=-1)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dims = tuple(int(x) for x in a.grid.split("x"))
    assert dims[0] * dims[1] == world
    uid = bytes.fromhex(os.environ["P3D_EMU_UID"])
    ok = True
    libs, comms = {}, {}
    for single in (False, True):
        libs[single] = emulib(single)
        libs[single].p3dfft_clean()
        # one mailbox directory per library: the two libraries are separate NCCL worlds
        u = bytearray(uid)
        u[uid.index(b"\0")] = ord("s" if single else "d")
        comms[single] = libs[single].comm_create(rank, world, bytes(u), -1)
    cases = SUITES[a.suite] if a.suite in SUITES else []
    if a.suite.startswith("fuzz:"):            # fuzz:<seed>:<count>
        _, seed, count = a.suite.split(":")
        cases = fuzz_cases(int(seed), int(count), dims)
    for case in cases:
        single = case[6]
        L = libs[single]
        X = M.EmuArrays(L)
        tol = 1e-5 if single else 1e-12
        before = nccl_counts(L)
        try:
            err = M.run_case(L, comms[single], dims, rank, case, X)
        except Exception as e:     # noqa: BLE001 - report and fail
            err = float("nan")
            print(f"rank {rank} grid {dims} case {case}: EXCEPTION {e!r}", flush=True)
            L.p3dfft_clean()
        after = nccl_counts(L)
        sent = after[0] - before[0]
        good = err <= tol
        if a.expect_p2p == 1 and sent != 0:
            good = False        # peer-to-peer plans never call ncclSend
        if a.expect_p2p == 0 and world > 1 and sent == 0:
            good = False
        print(f"rank {rank} grid {a.grid} n={case[0]} cut={case[1]} op={case[2]}/{case[3]} stride1={case[4]} nv={case[5]} "
              f"{'sp' if single else 'dp'}: rel-L2 {err:.2e} sends {sent} barriers {after[2] - before[2]} {'ok' if good else 'FAIL'}", flush=True)
        ok = ok and good
    if a.aux:
        for n, single in (((32, 24, 20), False), ((14, 26, 38), False)):
            L = libs[single]
            try:
                err = M.run_aux(L, comms[single], dims, rank, n, single, M.EmuArrays(L))
            except Exception as e:     # noqa: BLE001
                err = float("nan")
                print(f"rank {rank} grid {dims} aux n={n}: EXCEPTION {e!r}", flush=True)
                L.p3dfft_clean()
            good = err <= 1e-12
            print(f"rank {rank} grid {a.grid} n={n} rtran + r2c_1d + proc queries: {'ok' if good else 'FAIL'} ({err:.2e})", flush=True)
            ok = ok and good
    if a.repeat:
        L = libs[False]
        for n in ((64, 64, 64), (32, 32, 32)):
            try:
                err = run_repeat(L, comms[False], dims, rank, n, M.EmuArrays(L), a.asynchronous)
            except Exception as e:     # noqa: BLE001
                err = float("nan")
                print(f"rank {rank} grid {dims} repeat n={n}: EXCEPTION {e!r}", flush=True)
                L.p3dfft_clean()
            good = err <= 1e-12
            print(f"rank {rank} grid {a.grid} n={n} fwd x3, bwd x3{' (asynchronous)' if a.asynchronous else ''}: {err:.2e} {'ok' if good else 'FAIL'}", flush=True)
            ok = ok and good
    for single in (False, True):
        libs[single].comm_destroy(comms[single])
        # every p3dfft_clean above must have released its work buffers AND closed its mappings of the peers' buffers
        libs[single].lib.emu_live_allocations.restype = ctypes.c_longlong
        live = libs[single].lib.emu_live_allocations(None)
        if live != 0:
            print(f"rank {rank}: {live} device allocations / peer mappings still live after p3dfft_clean: FAIL", flush=True)
            ok = False
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
