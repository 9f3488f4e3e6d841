#!/usr/bin/env python
"""A/B of the library's run-time switches on N GPUs inside ONE torchrun job (no per-variant start-up cost).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/ab_multi.py [--size 1024] [--grid 2x4] [--pairs 10] [--variants "A=1 B=2;C=3"] [--dtype f64]

Every variant is a set of environment variables applied before p3dfft_setup (the library reads its switches there).
Per variant: 3 warm-up pairs, `--pairs` timed pairs (CUDA events, max over ranks), the per-stage timers, the round-trip
error and the direct-DFT spot check of bench.py on the forward result.  One line per variant on rank 0.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import bench
import p3dfft_b200 as pb

DEFAULT_VARIANTS = (";P3DFFT_B200_BULK=0;P3DFFT_B200_R32=0;P3DFFT_B200_FLAGBAR=1;"
                    "P3DFFT_B200_FLAGBAR=1 P3DFFT_B200_OVERLAP=2;P3DFFT_B200_FLAGBAR=1 P3DFFT_B200_OVERLAP=4;"
                    "P3DFFT_B200_FLAGBAR=1 P3DFFT_B200_OVERLAP=4 P3DFFT_B200_OVERLAP_SMS=40;"
                    "P3DFFT_B200_FLAGBAR=1 P3DFFT_B200_OVERLAP=8 P3DFFT_B200_OVERLAP_SMS=72;"
                    "P3DFFT_B200_FLAGBAR=1 P3DFFT_B200_BULK=0;P3DFFT_B200_P2P=0")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, nargs="+", default=[1024])
    ap.add_argument("--grid", default="")
    ap.add_argument("--pairs", type=int, default=10)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--variants", default=DEFAULT_VARIANTS)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    single = a.dtype == "f32"
    tdt = torch.float32 if single else torch.float64
    L = pb.load(single)
    L.p3dfft_clean()
    dims = tuple(int(x) for x in a.grid.split("x")) if a.grid else bench.GRID_FOR.get(world, (1, world))
    comm = 0
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(L.get_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        comm = L.comm_create(rank, world, bytes(uid.cpu().numpy().tobytes()), local)
    n = a.size * 3 if len(a.size) == 1 else a.size
    nx, ny, nz = n
    stream = torch.cuda.current_stream()
    L.set_stream(stream.cuda_stream)
    d = dist if world > 1 else None
    names = {4: "x_r2c", 6: "y_fwd", 7: "z_fwd", 8: "z_bwd", 9: "y_bwd", 11: "x_c2r", 0: "T1", 1: "T2", 2: "T3", 3: "T4"}
    for var in a.variants.split(";"):
        env = dict(kv.split("=") for kv in var.split())
        saved = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            L.p3dfft_setup(dims, nx, ny, nz, comm)
            _, info = L.plan_steps(dims, nx, ny, nz, rank, False, "fft")
            nreal, ncplx = nx * info.jisize * info.kjsize, info.iisize * info.jjsize * info.nzc
            g = torch.Generator(device="cuda").manual_seed(20240229 + rank)
            A = torch.rand(nreal, dtype=tdt, device="cuda", generator=g)
            F = torch.empty(2 * ncplx, dtype=tdt, device="cuda")
            B = torch.empty(nreal, dtype=tdt, device="cuda")
            for _ in range(3):
                L.p3dfft_ftran_r2c(A, F, "fft")
                L.p3dfft_btran_c2r(F, B, "tff")
            err = torch.tensor([float((B / (float(nx) * ny * nz) - A).abs().max())], dtype=torch.float64, device="cuda")
            spot = bench.spot_check_forward(A, F, info, (nx, ny, nz), rank, torch, d)
            L.set_timers()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if d:
                d.barrier()
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(a.pairs):
                L.p3dfft_ftran_r2c(A, F, "fft")
                L.p3dfft_btran_c2r(F, B, "tff")
            e1.record(stream)
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / a.pairs], dtype=torch.float64, device="cuda")
            tm = torch.tensor([t * 1e3 / a.pairs for t in L.get_timers()], dtype=torch.float64, device="cuda")
            if d:
                d.all_reduce(ms, op=d.ReduceOp.MAX)
                d.all_reduce(err, op=d.ReduceOp.MAX)
                d.all_reduce(tm, op=d.ReduceOp.MAX)
            p2p = L.p2p_active()
            L.p3dfft_clean()
            del A, F, B
            if rank == 0:
                st = " ".join(f"{names[i]}={float(tm[i]):.3f}" for i in (4, 0, 6, 1, 7, 8, 2, 9, 3, 11))
                print(f"[{var or 'default':<72}] {nx}x{ny}x{nz} {a.dtype} {dims[0]}x{dims[1]} p2p={int(p2p)} pair {float(ms):7.3f} ms | {st} | "
                      f"roundtrip {float(err):.1e} spot {spot:.1e}", flush=True)
        except Exception as ex:      # noqa: BLE001 - report and go on with the next variant
            print(f"rank {rank} [{var}] EXCEPTION {ex!r}", flush=True)
            L.p3dfft_clean()
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    L.reset_stream()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
