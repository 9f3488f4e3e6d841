"""Builds the TEST-ONLY host libraries into tests/emu/lib/ (never into the product's p3dfft_b200/lib/):

  librcopy_check.so              host harness of the real-copy address arithmetic (tests/c/rcopy_host.cpp)
  libemu_fast[_single].so        the specialised stage kernels' source compiled for the host (emu_fast.cpp)
  libp3dfft_emu[_single].so      the whole C ABI -- api.cpp, planner, every kernel file -- on the mock CUDA runtime and
                                 mock NCCL of this directory (emu_api.cpp, emu_kernels.cpp, emu_fast.cpp)

g++ only; nothing here is linked into, loaded by or shipped with libp3dfft.so.  See README.md.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

EMU = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(EMU))
CSRC = os.path.join(ROOT, "p3dfft_b200", "csrc")
LIBDIR = os.path.join(EMU, "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CUDA_INC = "-I" + os.path.join(os.path.dirname(os.path.dirname(NVCC)), "include")      # host-side declarations only (cuda_runtime.h, nccl.h)
SOURCES = ["fft_kernels.cu", "fft_fast.cu", "api.cpp"]
HEADERS = ["stage.h", "plan.h", "kernels.h", "fast.h", "fft_fast.cuh", "rcopy.h", "procmap.h"]
# -fno-gnu-unique / -Bsymbolic: these libraries share symbol names (p3d::launch_fast ...) with the product library, which the
# tests load into the same process; every library must bind to its own definitions
CXX = ["g++", "-O1", "-fno-gnu-unique", "-std=c++17", "-x", "c++", "-w", "-fPIC", "-pthread"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_emulation(verbose: bool = False) -> list[str]:
    os.makedirs(LIBDIR, exist_ok=True)
    out = []

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    target = os.path.join(LIBDIR, "librcopy_check.so")
    srcp = os.path.join(ROOT, "tests", "c", "rcopy_host.cpp")
    if _newer(target, [srcp, os.path.join(CSRC, "rcopy.h"), os.path.join(CSRC, "stage.h")]):
        run(["g++", "-O2", "-std=c++17", "-Wall", "-shared", "-fPIC", srcp, "-o", target])
    out.append(target)

    for name, defs in (("emu_fast", []), ("emu_fast_single", ["-DSINGLE_PREC"])):
        target = os.path.join(LIBDIR, f"lib{name}.so")
        deps = [os.path.join(EMU, f) for f in ("emu_fast.cpp", "cuda_emu.h", "emu_runtime.inc")] + [os.path.abspath(__file__)] + \
            [os.path.join(CSRC, f) for f in ("fft_fast.cu", "fft_fast.cuh", "fast.h", "stage.h")]
        if _newer(target, deps):
            run([*CXX, "-shared", "-Wl,-Bsymbolic", *defs, CUDA_INC, os.path.join(EMU, "emu_fast.cpp"), "-o", target])
        out.append(target)

    for name, defs in (("p3dfft_emu", []), ("p3dfft_emu_single", ["-DSINGLE_PREC"])):
        target = os.path.join(LIBDIR, f"lib{name}.so")
        objdir = os.path.join(LIBDIR, "obj_" + name)
        os.makedirs(objdir, exist_ok=True)
        deps = [os.path.join(EMU, f) for f in ("emu_api.cpp", "emu_kernels.cpp", "emu_fast.cpp", "cuda_emu.h", "emu_runtime.inc",
                                               "emu_streams.inc", "emu_mp.inc")] + \
            [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
        if _newer(target, deps):
            objs, cmds = [], []
            for src, extra in (("emu_api.cpp", []), ("emu_kernels.cpp", []), ("emu_fast.cpp", ["-DEMU_NO_RUNTIME"])):
                obj = os.path.join(objdir, src[:-4] + ".o")
                objs.append(obj)
                cmds.append([*CXX, "-c", *defs, *extra, CUDA_INC, os.path.join(EMU, src), "-o", obj])
            with ThreadPoolExecutor(max_workers=len(cmds)) as ex:
                list(ex.map(run, cmds))
            run(["g++", "-shared", "-pthread", "-Wl,-Bsymbolic", *objs, "-o", target, "-ldl", "-lrt"])
        out.append(target)
    return out


if __name__ == "__main__":
    print("\n".join(build_emulation(verbose="-v" in sys.argv)))
