"""Compiles the reference's own sample drivers -- UNCHANGED, from where they lie under /root/reference/sample/C -- against this
repository's headers (include/p3dfft.h, include/mpi_shim/mpi.h) and links them with the PRODUCT library libp3dfft[_single].so.

Output: oracle/_ref/drivers/<driver>[_sp]   (git-ignored; travels to the GPU box with the snapshot like the built libraries).
No reference source is copied into the repository; /root/reference only exists in the build container, so this step is skipped
elsewhere.  The binaries are the reference's acceptance tests at the drop-in boundary: tests/test_zzzz_reference_binaries.py runs
them on the GPU (they print their own verdict); tests/test_reference_drivers_emulated.py runs the same sources on the CPU emulation.
"""
from __future__ import annotations

import glob
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/sample/C"
OUT = os.path.join(ROOT, "oracle", "_ref", "drivers")
LIB = os.path.join(ROOT, "p3dfft_b200", "lib")


def build_ref_drivers(verbose: bool = False) -> list[str]:
    if not os.path.isdir(REF):
        return []
    os.makedirs(OUT, exist_ok=True)
    built = []
    for src in sorted(glob.glob(os.path.join(REF, "driver_*.c"))):
        name = os.path.basename(src)[:-2]
        for defs, lib, tag in (([], "p3dfft", ""), (["-DSINGLE_PREC"], "p3dfft_single", "_sp")):
            exe = os.path.join(OUT, name + tag)
            deps = [src, os.path.join(LIB, f"lib{lib}.so"), os.path.join(ROOT, "include", "p3dfft.h"),
                    os.path.join(ROOT, "include", "mpi_shim", "mpi.h"), os.path.abspath(__file__)]
            if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
                cmd = ["gcc", "-O1", "-w", *defs, f"-I{ROOT}/include/mpi_shim", f"-I{ROOT}/include", src, f"-L{LIB}", f"-l{lib}", "-lm",
                       "-Wl,-rpath,$ORIGIN/../../../p3dfft_b200/lib", "-o", exe]
                if verbose:
                    print(" ".join(cmd), flush=True)
                subprocess.check_call(cmd)
            built.append(exe)
    return built


if __name__ == "__main__":
    print("\n".join(build_ref_drivers(verbose=True)))
