#!/usr/bin/env python
"""Multi-GPU stress of the peer-to-peer transposes and their barrier (one process per GPU, torchrun).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tests/mp_stress.py [--iters 1000] [--grid 1x2]

`--iters` transforms in a seeded pseudo-random order (forward / backward, so that repeated directions exercise the
write-after-read rule of api.cpp as well as the alternating pattern of the drivers), while one rank per iteration is
deliberately delayed on the host before it enters the call.  The library is deterministic: every forward result must be
BITWISE equal to the first one, every backward result likewise -- a barrier that lets a peer's stores race a reader, or a
flag that is seen before the data it guards, shows up as a mismatch.  Sizes are small so that the kernels are short
against the delays.  Exit code 0 iff no mismatch on any rank.
"""
import argparse
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import p3dfft_b200 as pb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--grid", default="")
    ap.add_argument("--sizes", default="128x128x128,256x64x128")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = pb.load(False)
    L.p3dfft_clean()
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(L.get_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    comm = L.comm_create(rank, world, bytes(uid.cpu().numpy().tobytes()), local)
    grids = [tuple(int(x) for x in a.grid.split("x"))] if a.grid else \
        [(m1, world // m1) for m1 in range(1, world + 1) if world % m1 == 0]
    bad = 0
    total = 0
    for dims in grids:
        for sz in a.sizes.split(","):
            nx, ny, nz = (int(x) for x in sz.split("x"))
            L.p3dfft_setup(dims, nx, ny, nz, comm)
            _, info = L.plan_steps(dims, nx, ny, nz, rank, False, "fft")
            nreal, ncplx = nx * info.jisize * info.kjsize, info.iisize * info.jjsize * info.nzc
            g = torch.Generator(device="cuda").manual_seed(99 + rank)
            A = torch.rand(nreal, dtype=torch.float64, device="cuda", generator=g)
            F0 = torch.empty(2 * ncplx, dtype=torch.float64, device="cuda")
            B0 = torch.empty(nreal, dtype=torch.float64, device="cuda")
            L.p3dfft_ftran_r2c(A, F0, "fft")
            L.p3dfft_btran_c2r(F0, B0, "tff")
            F, B = torch.empty_like(F0), torch.empty_like(B0)
            rng = random.Random(1234)            # the same sequence on every rank: the calls are collective
            per = max(1, a.iters // (len(grids) * len(a.sizes.split(","))))
            for it in range(per):
                fwd = rng.random() < 0.5
                slow, delay = rng.randrange(world), rng.random() * 2e-3
                if rank == slow:
                    time.sleep(delay)
                if fwd:
                    F.fill_(-1.0)
                    L.p3dfft_ftran_r2c(A, F, "fft")
                    ok = torch.equal(F, F0)
                else:
                    B.fill_(-1.0)
                    L.p3dfft_btran_c2r(F0, B, "tff")
                    ok = torch.equal(B, B0)
                bad += 0 if ok else 1
                total += 1
            L.p3dfft_clean()
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    flagbar = os.environ.get("P3DFFT_B200_FLAGBAR", "default")
    if rank == 0:
        print(f"MP STRESS {'PASS' if int(t) == 0 else 'FAIL'}: {total} transforms per rank on grids {grids}, "
              f"{int(t)} mismatching results (all ranks), FLAGBAR={flagbar} OVERLAP={os.environ.get('P3DFFT_B200_OVERLAP', 'default')}", flush=True)
    L.comm_destroy(comm)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 0 else 1)


if __name__ == "__main__":
    main()
