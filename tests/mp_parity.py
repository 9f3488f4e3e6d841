#!/usr/bin/env python
"""Multi-GPU parity driver (one process per GPU, launched by torchrun; NCCL transposes).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
      --master-port 29511 tests/mp_parity.py [--grids 2x2,1x4,4x1]

Every rank transforms its pencil of ONE global Philox field through the C ABI and compares
with the oracle's slice of the global transform (relative L2 <= 1e-12 double / 1e-5 single).
Covers the reference's own matrix (extra/makejob.py:122-152): even 32^3 and 128^3, uneven
14x26x38, pruned 64^3 -> 32^3, Chebyshev 32x32x33, the *_many calls and the STRIDE1 layout.
Exit code 0 iff every case passes on every rank.
"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import p3dfft_b200 as pb
from oracle import p3dfft_oracle as po

CASES = [
    # (n, cut, opf, opb, stride1, nv, single)
    ((128, 128, 128), None, "fft", "tff", False, 1, False),        # BASELINE config 1 (driver_inverse/driver_sine size)
    ((32, 32, 32), None, "fft", "tff", False, 1, False),
    ((14, 26, 38), None, "fft", "tff", False, 1, False),
    ((64, 64, 64), (32, 32, 32), "fft", "tff", False, 1, False),
    ((32, 32, 33), None, "ffc", "cff", False, 1, False),
    ((32, 32, 32), None, "ffn", "nff", False, 1, False),
    ((32, 32, 32), None, "fft", "tff", False, 2, False),
    ((64, 32, 48), None, "fft", "tff", True, 1, False),
    ((256, 128, 64), None, "fft", "tff", False, 1, False),
    ((128, 128, 128), None, "fft", "tff", False, 1, True),
    ((64, 64, 64), (42, 42, 42), "fft", "tff", False, 1, True),
]


class TorchArrays:
    """device arrays of the GPU run: torch CUDA tensors"""

    def dev(self, a):
        import torch
        return torch.from_numpy(np.ascontiguousarray(a)).cuda()

    def full(self, n, value, np_dtype):
        import torch
        return torch.full((n,), value, dtype=torch.float32 if np_dtype == np.float32 else torch.float64, device="cuda")

    def host(self, t):
        return t.cpu().numpy()

    def sync(self):
        import torch
        torch.cuda.synchronize()


class EmuArrays:
    """'device' arrays of the CPU emulation (tests/mp_emu.py): numpy arrays registered with the mock runtime, which the
    library then uses in place like device memory"""

    def __init__(self, L):
        self.L, self.keep = L, []
        L.lib.emu_register_device_range.argtypes = [ctypes.c_void_p, ctypes.c_size_t]

    def dev(self, a):
        a = np.ascontiguousarray(a).copy()
        self.L.lib.emu_register_device_range(a.ctypes.data, a.nbytes)
        self.keep.append(a)
        return a

    def full(self, n, value, np_dtype):
        return self.dev(np.full(n, value, dtype=np_dtype))

    def host(self, t):
        return t

    def sync(self):
        pass


def run_case(L, comm, dims, rank, case, X=None):
    X = X or TorchArrays()
    n, cut, opf, opb, stride1, nv, single = case
    nx, ny, nz = n
    c = cut or (None, None, None)
    rt, ct = (np.float32, np.complex64) if single else (np.float64, np.complex128)
    L.set_layout(stride1, False)
    L.p3dfft_setup(dims, nx, ny, nz, comm, *c)
    d = po.Decomp(nx, ny, nz, dims, rank, *c, stride1=stride1, elem=4 if single else 8)
    _, _, isz = L.p3dfft_get_dims(1)
    _, _, fsz = L.p3dfft_get_dims(2)
    assert list(isz) == d.get_dims(1)[2] and list(fsz) == d.get_dims(2)[2]
    nreal, ncplx = int(np.prod(isz)), int(np.prod(fsz))
    fields = [po.philox_field(nx, ny, nz, seed=20240229 + v) for v in range(nv)]
    loc = [np.asfortranarray(f[po.local_in_slice(d)]).astype(rt) for f in fields]
    tA = X.dev(np.concatenate([a.ravel(order="F") for a in loc]))
    tF = X.full(2 * ncplx * nv, 0.0, rt)
    if nv == 1:
        L.p3dfft_ftran_r2c(tA, tF, opf)
    else:
        L.p3dfft_ftran_r2c_many(tA, nreal, tF, ncplx, nv, opf)
    F = X.host(tF).view(ct).reshape(nv, ncplx)
    errs = []
    for v in range(nv):
        exp = po.local_forward(fields[v].astype(rt).astype(np.float64), d, opf)
        errs.append(po.rel_l2(F[v], np.asfortranarray(exp).ravel(order="F")))
    # backward from the oracle's global spectrum
    Fg = [po.global_forward(f.astype(rt).astype(np.float64), d, opf) for f in fields]
    parts = []
    for f in Fg:
        l = f[po.local_out_slice(d)]
        if stride1:
            l = l.transpose(2, 1, 0)
        parts.append(np.asfortranarray(l).astype(ct).ravel(order="F"))
    tFi = X.dev(np.concatenate(parts).view(rt))
    tB = X.full(nreal * nv, 0.0, rt)
    if nv == 1:
        L.p3dfft_btran_c2r(tFi, tB, opb)
    else:
        L.p3dfft_btran_c2r_many(tFi, ncplx, tB, nreal, nv, opb)
    B = X.host(tB).reshape(nv, nreal)
    for v in range(nv):
        exp = po.local_backward(Fg[v], d, opb)
        errs.append(po.rel_l2(B[v], np.asfortranarray(exp).ravel(order="F")))
    L.p3dfft_clean()
    return max(errs)


def run_aux(L, comm, dims, rank, n, single, X=None):
    """Remaining module routines on this grid: the four real-data transposes (bit-exact: data movement only),
    p3dfft_ftran_r2c_1d and the process-map queries.  Returns the worst error (0.0 = exact)."""
    X = X or TorchArrays()
    nx, ny, nz = n
    rt, ct = (np.float32, np.complex64) if single else (np.float64, np.complex128)
    L.set_layout(False, False)
    L.p3dfft_setup(dims, nx, ny, nz, comm)
    d = po.Decomp(nx, ny, nz, dims, rank, elem=4 if single else 8)
    G = po.philox_field(nx, ny, nz, seed=77).astype(rt)
    worst = 0.0
    t_acc = 0.0
    for which in pb.RTRAN_NAMES * 2:            # twice: the second round reuses the receive buffers (hazard rule)
        src_sl, dst_sl = po.rtran_slices(d, which)
        src = X.dev(np.asfortranarray(G[src_sl]).ravel(order="F").copy())
        exp = po.rtran_local(G, d, which)
        dst = X.full(exp.size, float("nan"), rt)
        dstart, dend, dsize, t_acc = L.rtran(which, src, dst, t_acc)
        if [list(dstart), list(dend), list(dsize)] != [list(x) for x in po.rtran_dims(d, which)]:
            worst = max(worst, 2.0)       # no assert inside the collective sequence: a rank that raised would hang the others
        if not np.array_equal(X.host(dst), exp.ravel(order="F")):
            worst = max(worst, 1.0)
        # host arrays take the staged path
        dsth = np.full(exp.size, np.nan, dtype=rt)
        L.rtran(which, np.asfortranarray(G[src_sl]).ravel(order="F").copy(), dsth)
        if not np.array_equal(dsth, exp.ravel(order="F")):
            worst = max(worst, 1.0)
    A = np.asfortranarray(G[po.local_in_slice(d)])
    tA = X.dev(A.ravel(order="F").copy())
    tC = X.full(2 * d.nxhp * d.jisize * d.kjsize, 0.0, rt)
    L.p3dfft_ftran_r2c_1d(tA, tC)
    e = po.rel_l2(X.host(tC).view(ct), po.forward_r2c_1d(A.astype(np.float64)).ravel(order="F"))
    worst = max(worst, e / (1e-5 if single else 1e-12) * 1e-12)
    g = po.ProcGrid(nx, ny, nz, dims)
    P = dims[0] * dims[1]
    assert L.p3dfft_get_mpi_info()[:2] == (rank, P)
    for r in range(P):
        assert L.proc_id2coords(r) == (g.proc_id2coords[2 * r], g.proc_id2coords[2 * r + 1])
        assert L.proc_coords2id(*L.proc_id2coords(r)) == r
        for conf in (1, 2):
            assert L.proc_dims(conf, r) == [g.proc_dims[(conf, k, r)] for k in range(1, 10)]
        for o in (-1, 1):
            for di in (1, 2):
                assert L.proc_neighb(r, o, di) == g.proc_neighb(r, o, di)
    for base, size, conf in (((1, 1, 1), (nx, ny, nz), 1), ((2, 3, 1), (3, 4, nz), 2), ((1, 2, 2), (nx, 3, 2), 1)):
        exp = g.get_proc_parts(*base, *size, conf)
        assert L.get_proc_parts(base, size, conf, P) == (exp[0][:P], exp[1], exp[2])
    L.p3dfft_clean()
    return worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grids", default="")
    a = ap.parse_args()
    import torch                      # (imported here: tests/mp_emu.py reuses this module's cases and checks without torch)
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grids = [tuple(int(x) for x in g.split("x")) for g in a.grids.split(",") if g] or \
        [(m1, world // m1) for m1 in range(1, world + 1) if world % m1 == 0]
    ok = True
    comms = {}
    for single in (False, True):
        L = pb.load(single)
        L.p3dfft_clean()
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(L.get_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        comms[single] = L.comm_create(rank, world, bytes(uid.cpu().numpy().tobytes()), local)
    for dims in grids:
        for case in CASES:
            single = case[6]
            L = pb.load(single)
            tol = 1e-5 if single else 1e-12
            try:
                err = run_case(L, comms[single], dims, rank, case)
                good = err <= tol
            except Exception as e:     # noqa: BLE001 - report and fail
                err, good = float("nan"), False
                print(f"rank {rank} grid {dims} case {case}: EXCEPTION {e!r}", flush=True)
                L.p3dfft_clean()
            flag = torch.tensor([0 if good else 1], device="cuda")
            worst = torch.tensor([err if err == err else 1e9], dtype=torch.float64, device="cuda")
            dist.all_reduce(flag)
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"grid {dims[0]}x{dims[1]} n={case[0]} cut={case[1]} op={case[2]}/{case[3]} stride1={case[4]} "
                      f"nv={case[5]} {'sp' if single else 'dp'}: max rel-L2 {float(worst):.2e} "
                      f"{'ok' if int(flag) == 0 else 'FAIL'}", flush=True)
            ok = ok and int(flag) == 0
        for n, single in (((32, 24, 20), False), ((14, 26, 38), False), ((64, 64, 64), True)):
            L = pb.load(single)
            try:
                err = run_aux(L, comms[single], dims, rank, n, single)
                good = err <= 1e-12
            except Exception as e:     # noqa: BLE001 - report and fail
                err, good = float("nan"), False
                print(f"rank {rank} grid {dims} aux n={n}: EXCEPTION {e!r}", flush=True)
                L.p3dfft_clean()
            flag = torch.tensor([0 if good else 1], device="cuda")
            dist.all_reduce(flag)
            if rank == 0:
                print(f"grid {dims[0]}x{dims[1]} n={n} rtran x2y/y2x/x2z/z2x + r2c_1d + proc queries "
                      f"{'sp' if single else 'dp'}: {'ok' if int(flag) == 0 else 'FAIL'}", flush=True)
            ok = ok and int(flag) == 0
    for single, h in comms.items():
        pb.load(single).comm_destroy(h)
    dist.destroy_process_group()
    if rank == 0:
        print("MP PARITY", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
