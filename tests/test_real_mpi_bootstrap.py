"""Unmodified callers under an MPI of their own (reference: build/setup.F90:178-261, driver_sine.c:126).

tests/c/mpi_stub.c builds a stand-alone MPI shared library (MPICH-family handle values) that knows nothing about P3DFFT.
Drivers are compiled against ITS header (tests/c/mpi_stub/mpi.h: prototypes only) and linked with it -- the position of an
application built with a real MPI -- and hand MPI_Comm_c2f(MPI_COMM_WORLD) = 0x44000000 to p3dfft_setup.  The library finds
MPI_Comm_f2c / MPI_Comm_rank / MPI_Comm_size / MPI_Bcast in the process (dlsym), learns rank and size, distributes the NCCL id
and creates its own communicator.  Run here on the CPU-emulated library, 1, 2 and 4 ranks; the reference's own sample drivers
where the reference tree exists, this repository's wave_roundtrip.c everywhere.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "emu", "lib")
REF = "/root/reference/sample/C"
_port = [33750 + (os.getpid() % 83) * 2]

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(LIB, "libp3dfft_emu.so")), reason="emulated library not built")


@pytest.fixture(scope="module")
def stub(tmp_path_factory):
    out = tmp_path_factory.mktemp("mpistub")
    so = out / "libmpi_stub.so"
    r = subprocess.run(["gcc", "-O1", "-w", "-fPIC", "-shared", os.path.join(ROOT, "tests", "c", "mpi_stub.c"), "-o", str(so)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sym = subprocess.run(["nm", "-D", str(so)], capture_output=True, text=True).stdout
    assert " T MPI_Bcast" in sym and "p3dfft" not in sym          # an MPI that has never heard of the library
    return out


def build(stub, src, name, defs=(), lib="libp3dfft_emu.so"):
    exe = stub / name
    cmd = ["gcc", "-O1", "-w", *defs, f"-I{ROOT}/tests/c/mpi_stub", f"-I{ROOT}/include", src, f"-L{stub}", "-lmpi_stub", f"-L{LIB}",
           f"-l:{lib}", "-lm", f"-Wl,-rpath,{LIB}", f"-Wl,-rpath,{stub}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return str(exe)


def run(exe, ranks, args=(), cwd=None, stdin_fields=None):
    _port[0] += 2
    e = {k: v for k, v in os.environ.items() if not k.startswith("P3DFFT_B200_")}
    e.update({"P3D_EMU_SHM": "1", "P3D_EMU_TIMEOUT": "60"})
    cmd = [sys.executable, os.path.join(ROOT, "tools", "p3drun.py"), "-n", str(ranks), "--port", str(_port[0]), "--timeout", "240", exe,
           *map(str, args)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=cwd, env=e)


@pytest.mark.parametrize("ranks,grid", [(1, (1, 1)), (2, (1, 2)), (4, (2, 2))])
def test_own_driver_under_a_separate_mpi(stub, ranks, grid):
    exe = build(stub, os.path.join(ROOT, "tests", "c", "wave_roundtrip.c"), "wave_roundtrip_mpi")
    r = run(exe, ranks, (64, 48, 80, *grid))
    assert r.returncode == 0 and "Results are correct" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name,ranks,grid", [("driver_sine", 2, (2, 1)), ("driver_inverse", 4, (2, 2)), ("driver_rand", 1, (1, 1))])
def test_reference_driver_unchanged_under_a_separate_mpi(stub, tmp_path, name, ranks, grid):
    exe = build(stub, os.path.join(REF, name + ".c"), name + "_mpi")
    (tmp_path / "stdin").write_text("64 64 64 2 1\n")
    (tmp_path / "dims").write_text(f"{grid[0]} {grid[1]}\n")
    r = run(exe, ranks, cwd=str(tmp_path))
    assert r.returncode == 0 and "Results are correct" in r.stdout and "incorrect" not in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert f"Using processor grid {grid[0]} x {grid[1]}" in r.stdout


def test_unknown_handle_without_an_mpi_is_an_error():
    """no MPI in the process: a non-zero value that is not a library handle must not silently become a one-rank run"""
    sys.path.insert(0, ROOT)
    import p3dfft_b200 as pb
    lib = pb.P3DFFT(False, path=os.path.join(LIB, "libp3dfft_emu.so"))
    with pytest.raises(RuntimeError, match="neither a handle"):
        lib.p3dfft_setup((1, 1), 16, 16, 16, 0x44000000)
    lib.p3dfft_setup((1, 1), 16, 16, 16, 0)          # 0 stays the implicit one-rank communicator
    lib.p3dfft_clean()
