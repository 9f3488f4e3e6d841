#!/usr/bin/env python
"""p3drun -- start N ranks of a program on one node, one GPU per rank (the `mpirun -np N` of this build).

  python tools/p3drun.py -n 4 [--port 29600] ./driver_sine [args...]

Sets RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR=127.0.0.1 and MASTER_PORT for every rank; programs built
against include/mpi_shim/mpi.h (or using torch.distributed) pick them up.  Exit code = first non-zero
exit code of a rank; the other ranks are terminated when one fails.
"""
import argparse
import os
import subprocess
import sys
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-n", "--np", type=int, default=1)
    ap.add_argument("--port", type=int, default=29600)
    ap.add_argument("--timeout", type=float, default=600.0)
    ap.add_argument("cmd", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    if not a.cmd:
        ap.error("no program given")
    procs = []
    for r in range(a.np):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(a.np), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(a.port))
        procs.append(subprocess.Popen(a.cmd, env=env))
    t0, rc = time.time(), 0
    live = list(procs)
    while live:
        for p in list(live):
            c = p.poll()
            if c is None:
                continue
            live.remove(p)
            if c != 0 and rc == 0:
                rc = c
        if rc != 0 or time.time() - t0 > a.timeout:
            for p in live:
                p.terminate()
            if not rc:
                rc = 124
            break
        time.sleep(0.05)
    for p in procs:
        try:
            p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            p.kill()
    sys.exit(rc)


if __name__ == "__main__":
    main()
