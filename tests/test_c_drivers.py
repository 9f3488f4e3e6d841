"""The C side of the drop-in boundary: include/p3dfft.h + the MPI stand-in include/mpi_shim/mpi.h.

CPU: the stand-in's collectives (1, 3 and 5 ranks under tools/p3drun.py), and -- where the reference tree is
present (this container only, never on the GPU box) -- that every one of its sample/C drivers compiles and
links UNCHANGED against our headers and library.
GPU: tests/c/wave_roundtrip.c, our own acceptance driver with the reference drivers' checks (four forward
spikes of modulus N/8, round trip within 1e-14*N/4), through Cp3dfft_* with host arrays, 1 rank and -- when the
box has them -- 2 and 4 ranks on real GPUs.
"""
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "p3dfft_b200", "lib")
REF = "/root/reference/sample/C"
_port = [29750 + (os.getpid() % 97) * 2]


def _built():
    sys.path.insert(0, ROOT)
    from p3dfft_b200 import build as b
    b.build_all()
    return b.build_c_drivers()


def _run(n, exe, *args, timeout=300):
    _port[0] += 2
    cmd = [sys.executable, os.path.join(ROOT, "tools", "p3drun.py"), "-n", str(n), "--port", str(_port[0]), exe, *map(str, args)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("n", [1, 3, 5])
def test_mpi_standin_collectives(n):
    _built()
    r = _run(n, os.path.join(LIB, "shim_selftest"))
    assert r.returncode == 0 and "passed" in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_reference_c_drivers_link_unchanged(tmp_path):
    _built()
    drivers = sorted(glob.glob(os.path.join(REF, "*.c")))
    assert len(drivers) == 9
    for src in drivers:
        for defs, lib in (([], "p3dfft"), (["-DSINGLE_PREC"], "p3dfft_single")):
            exe = tmp_path / (os.path.basename(src)[:-2] + ("_sp" if defs else ""))
            cmd = ["gcc", "-O1", "-w", *defs, f"-I{ROOT}/include/mpi_shim", f"-I{ROOT}/include", src, f"-L{LIB}", f"-l{lib}", "-lm",
                   "-o", str(exe)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            assert r.returncode == 0, f"{os.path.basename(src)} {defs}: {r.stderr}"


@pytest.mark.gpu
@pytest.mark.parametrize("exe,n", [("wave_roundtrip", (64, 64, 64)), ("wave_roundtrip", (128, 32, 48)), ("wave_roundtrip", (20, 12, 36)),
                                   ("wave_roundtrip_single", (64, 64, 64))])
def test_c_driver_single_rank(exe, n):
    r = _run(1, os.path.join(LIB, exe), *n)
    assert r.returncode == 0 and "Results are correct" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("ranks,grid", [(2, (1, 2)), (2, (2, 1)), (4, (2, 2))])
def test_c_driver_multi_rank(ranks, grid):
    import torch
    if torch.cuda.device_count() < ranks:
        pytest.skip(f"needs {ranks} GPUs")
    r = _run(ranks, os.path.join(LIB, "wave_roundtrip"), 64, 48, 80, *grid)
    assert r.returncode == 0 and "Results are correct" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("exe,n", [("spec_epilogue", (64, 48, 40)), ("spec_epilogue", (128, 64, 64)), ("spec_epilogue", (30, 18, 14)),
                                   ("spec_epilogue_single", (64, 64, 64))])
def test_c_epilogue_driver_single_rank(exe, n):
    """tests/c/spec_epilogue.c: fused normalisation, device power spectrum, rtran_* and r2c_1d from C"""
    r = _run(1, os.path.join(LIB, exe), *n)
    assert r.returncode == 0 and "Results are correct" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("ranks,grid", [(2, (1, 2)), (2, (2, 1)), (4, (2, 2))])
def test_c_epilogue_driver_multi_rank(ranks, grid):
    import torch
    if torch.cuda.device_count() < ranks:
        pytest.skip(f"needs {ranks} GPUs")
    r = _run(ranks, os.path.join(LIB, "spec_epilogue"), 64, 48, 40, *grid)
    assert r.returncode == 0 and "Results are correct" in r.stdout, r.stdout + r.stderr


# ---- the same acceptance drivers on the CPU-emulated library (tests/emu), 1 / 2 / 4 ranks -----------------------------------
EMULIB = os.path.join(ROOT, "tests", "emu", "lib")
EMU_OK = os.path.exists(os.path.join(EMULIB, "libp3dfft_emu.so"))


@pytest.fixture(scope="module")
def emu_exes(tmp_path_factory):
    out = tmp_path_factory.mktemp("cdrivers_emu")
    built = {}
    for exe, src, defs, lib in (("wave_roundtrip", "wave_roundtrip.c", [], "libp3dfft_emu.so"),
                                ("wave_roundtrip_single", "wave_roundtrip.c", ["-DSINGLE_PREC"], "libp3dfft_emu_single.so"),
                                ("spec_epilogue", "spec_epilogue.c", [], "libp3dfft_emu.so"),
                                ("spec_epilogue_single", "spec_epilogue.c", ["-DSINGLE_PREC"], "libp3dfft_emu_single.so")):
        target = out / exe
        cmd = ["gcc", "-O2", "-Wall", *defs, f"-I{ROOT}/include/mpi_shim", f"-I{ROOT}/include", os.path.join(ROOT, "tests", "c", src),
               f"-L{EMULIB}", f"-l:{lib}", "-lm", f"-Wl,-rpath,{EMULIB}", "-o", str(target)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        built[exe] = str(target)
    return built


def _run_emu(n, exe, *args, env=None):
    _port[0] += 2
    e = {k: v for k, v in os.environ.items() if not k.startswith("P3DFFT_B200_")}
    e.update({"P3D_EMU_SHM": "1", "P3D_EMU_TIMEOUT": "60"})
    e.update(env or {})
    cmd = [sys.executable, os.path.join(ROOT, "tools", "p3drun.py"), "-n", str(n), "--port", str(_port[0]), "--timeout", "240", exe, *map(str, args)]
    try:
        return subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=e)
    finally:
        from tests.test_reference_drivers_emulated import sweep_shm
        sweep_shm()


@pytest.mark.skipif(not EMU_OK, reason="emulated library not built")
@pytest.mark.parametrize("exe,ranks,args", [("wave_roundtrip", 1, (64, 64, 64)), ("wave_roundtrip", 2, (64, 48, 80, 1, 2)),
                                            ("wave_roundtrip", 4, (64, 48, 80, 2, 2)), ("wave_roundtrip_single", 2, (64, 64, 64, 2, 1)),
                                            ("wave_roundtrip", 2, (20, 12, 36, 1, 2))])
def test_c_driver_on_emulated_ranks(emu_exes, exe, ranks, args):
    r = _run_emu(ranks, emu_exes[exe], *args)
    assert r.returncode == 0 and "Results are correct" in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(not EMU_OK, reason="emulated library not built")
@pytest.mark.parametrize("exe,ranks,args", [("spec_epilogue", 1, (64, 24, 20)), ("spec_epilogue", 2, (32, 24, 20, 1, 2)),
                                            ("spec_epilogue", 2, (64, 16, 12, 2, 1)), ("spec_epilogue_single", 4, (64, 32, 16, 2, 2)),
                                            ("spec_epilogue", 4, (30, 18, 14, 2, 2))])
def test_c_epilogue_driver_on_emulated_ranks(emu_exes, exe, ranks, args):
    """fused normalisation, device power spectrum (NCCL all-reduce over the ranks), rtran_* and r2c_1d from C, several ranks"""
    r = _run_emu(ranks, emu_exes[exe], *args)
    assert r.returncode == 0 and "Results are correct" in r.stdout, r.stdout + r.stderr
