"""Builds the C-ABI shared libraries in-tree with nvcc for sm_100a.

  p3dfft_b200/lib/libp3dfft.so          double precision (the reference's default build)
  p3dfft_b200/lib/libp3dfft_single.so   -DSINGLE_PREC     (configure --enable-single)

The STRIDE1 / DIMS_C switches of the reference's configure are run-time options of the same
binaries (p3dfft_b200_set_layout); `--variants` additionally emits libraries with those
defaults baked in (libp3dfft_stride1.so, ...), matching configure.ac:171-329.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-gnu-unique", "--expt-relaxed-constexpr"]
# -fno-gnu-unique: static locals of template functions (launch_spectrum<T>::configured ...) must stay private to each library;
# the double and single libraries are routinely loaded into one process (tests/mp_parity.py, bench.py)
SOURCES = ["fft_kernels.cu", "fft_fast.cu", "api.cpp"]
HEADERS = ["stage.h", "plan.h", "kernels.h", "fast.h", "fft_fast.cuh", "rcopy.h", "procmap.h"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_one(name: str, defines: list[str], verbose: bool = False, force: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    target = os.path.join(LIBDIR, f"lib{name}.so")
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [
        os.path.join(HERE, "..", "include", "p3dfft_b200.h"), os.path.abspath(__file__)]
    objs, cmds = [], []
    for src in SOURCES:
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        if force or _newer(obj, deps):
            cmd = [NVCC, *ARCH, *COMMON, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
            if src.endswith(".cu") and verbose:
                cmd += ["-Xptxas", "-v"]
            cmds.append(cmd)
    if cmds:
        def run(cmd):
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
        with ThreadPoolExecutor(max_workers=len(cmds)) as ex:
            list(ex.map(run, cmds))
    if force or _newer(target, objs):
        # -Bsymbolic: references inside a library bind to its own definitions even when the double and the single library
        # (same symbol names, different real type) sit in one process loaded RTLD_GLOBAL
        cmd = [NVCC, *ARCH, "-shared", "-Xlinker", "-Bsymbolic", "-o", target, *objs, "-ldl", "-lpthread"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return target


def build_all(verbose: bool = False, variants: bool = False, force: bool = False) -> list[str]:
    jobs = [("p3dfft", []), ("p3dfft_single", ["-DSINGLE_PREC"])]
    if variants:
        jobs += [("p3dfft_stride1", ["-DSTRIDE1"]), ("p3dfft_single_stride1", ["-DSINGLE_PREC", "-DSTRIDE1"])]
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        return list(ex.map(lambda j: build_one(j[0], j[1], verbose, force), jobs))


def build_c_drivers(verbose: bool = False) -> list[str]:
    """C acceptance drivers (tests/c/*.c) built with gcc against include/p3dfft.h and the MPI stand-in
    include/mpi_shim/mpi.h, next to the libraries (rpath $ORIGIN) so that they travel with them."""
    root = os.path.join(HERE, "..")
    inc = ["-I" + os.path.join(root, "include", "mpi_shim"), "-I" + os.path.join(root, "include")]
    jobs = [("wave_roundtrip", "wave_roundtrip.c", [], "p3dfft"),
            ("wave_roundtrip_single", "wave_roundtrip.c", ["-DSINGLE_PREC"], "p3dfft_single"),
            ("shim_selftest", "shim_selftest.c", [], "p3dfft"),
            ("spec_epilogue", "spec_epilogue.c", [], "p3dfft"),
            ("spec_epilogue_single", "spec_epilogue.c", ["-DSINGLE_PREC"], "p3dfft_single")]
    out = []
    for exe, src, defs, lib in jobs:
        target = os.path.join(LIBDIR, exe)
        srcp = os.path.join(root, "tests", "c", src)
        deps = [srcp, os.path.join(root, "include", "mpi_shim", "mpi.h"), os.path.join(root, "include", "p3dfft.h"),
                os.path.join(root, "include", "p3dfft_b200.h"), os.path.join(LIBDIR, f"lib{lib}.so")]
        if _newer(target, deps):
            cmd = ["gcc", "-O2", "-Wall", *defs, *inc, srcp, "-L" + LIBDIR, "-l" + lib, "-lm", "-Wl,-rpath,$ORIGIN", "-o", target]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
        out.append(target)
    return out


if __name__ == "__main__":
    libs = build_all(verbose="-v" in sys.argv, variants="--variants" in sys.argv, force="-f" in sys.argv)
    print("\n".join(libs + build_c_drivers(verbose="-v" in sys.argv)))
    # (the test-only host builds -- kernel emulation, mock runtime -- live in tests/emu/build.py)
