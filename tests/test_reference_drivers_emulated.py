"""The reference's OWN sample drivers, unchanged, run against this library on the CPU.

Where the reference tree is present (this container; never the GPU box) every sample/C/driver_*.c is compiled where it lies
-- no copy, no edit -- against include/p3dfft.h and the MPI stand-in include/mpi_shim/mpi.h and linked with the EMULATED
library (libp3dfft_emu.so: the product's api.cpp, planner and kernel sources on the mock CUDA runtime / mock NCCL of
tests/emu).  The drivers then run on 1, 2 and 4 ranks under tools/p3drun.py and must print their own verdict
("Results are correct": driver_sine.c:239-247 and siblings; driver_inverse.c:222-258 checks the four spikes of the known
answer) -- the drop-in boundary exercised by the code that defines it.  driver_spec prints a
spectrum instead of a verdict: E(2) = 3/16 for its sine field, everything else zero.
"""
import glob
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "emu", "lib")
REF = "/root/reference/sample/C"
_port = [31750 + (os.getpid() % 89) * 2]

pytestmark = [pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)"),
              pytest.mark.skipif(not os.path.exists(os.path.join(LIB, "libp3dfft_emu.so")), reason="emulated library not built")]


@pytest.fixture(scope="module")
def exes(tmp_path_factory):
    out = tmp_path_factory.mktemp("refdrivers")
    built = {}
    for src in sorted(glob.glob(os.path.join(REF, "driver_*.c"))):
        name = os.path.basename(src)[:-2]
        for defs, lib, tag in (([], "libp3dfft_emu.so", ""), (["-DSINGLE_PREC"], "libp3dfft_emu_single.so", "_sp")):
            exe = out / (name + tag)
            cmd = ["gcc", "-O1", "-w", *defs, f"-I{ROOT}/include/mpi_shim", f"-I{ROOT}/include", src, f"-L{LIB}", f"-l:{lib}", "-lm",
                   f"-Wl,-rpath,{LIB}", "-o", str(exe)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            assert r.returncode == 0, f"{name} {defs}: {r.stderr}"
            built[name + tag] = str(exe)
    return built


def run_driver(exe, tmp_path, ranks, grid, size, nv=None, reps=1, env=None):
    wd = tmp_path / f"run_{os.path.basename(exe)}_{ranks}"
    wd.mkdir()
    ndim = 2
    fields = [*size, ndim] + ([nv] if nv else []) + [reps]
    (wd / "stdin").write_text(" ".join(map(str, fields)) + "\n")
    (wd / "dims").write_text(f"{grid[0]} {grid[1]}\n")
    _port[0] += 2
    e = {k: v for k, v in os.environ.items() if not k.startswith("P3DFFT_B200_")}
    e.update({"P3D_EMU_SHM": "1", "P3D_EMU_TIMEOUT": "60"})
    e.update(env or {})
    cmd = [sys.executable, os.path.join(ROOT, "tools", "p3drun.py"), "-n", str(ranks), "--port", str(_port[0]), "--timeout", "240", exe]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=str(wd), env=e)
    finally:
        sweep_shm()
    return r


def sweep_shm():
    """removes what ranks that no longer exist left in /dev/shm (a driver that aborts does not run its atexit clean-up)"""
    for path in glob.glob("/dev/shm/p3demu_mem_*") + glob.glob("/dev/shm/p3demu_nccl_*"):
        m = re.match(r"p3demu_(?:mem|nccl)_(\d+)_", os.path.basename(path))
        if m and not os.path.exists(f"/proc/{m.group(1)}"):
            if os.path.isdir(path):
                shutil.rmtree(path, ignore_errors=True)
            else:
                try:
                    os.unlink(path)
                except OSError:
                    pass


VERDICT = ["driver_sine", "driver_sine_inplace", "driver_rand", "driver_noop", "driver_inverse"]
MANY = ["driver_sine_many", "driver_sine_inplace_many", "driver_rand_many"]


@pytest.mark.parametrize("name,ranks,grid", [(n, r, g) for n in VERDICT for r, g in ((1, (1, 1)), (4, (2, 2)))] +
                         [("driver_sine", 2, (1, 2)), ("driver_inverse", 2, (2, 1)), ("driver_sine_inplace", 4, (1, 4))])
def test_reference_driver_verdict(exes, tmp_path, name, ranks, grid):
    r = run_driver(exes[name], tmp_path, ranks, grid, (64, 64, 64))
    assert r.returncode == 0 and "Results are correct" in r.stdout and "incorrect" not in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"Using processor grid {grid[0]} x {grid[1]}" in r.stdout


@pytest.mark.parametrize("name", MANY)
@pytest.mark.parametrize("ranks,grid", [(1, (1, 1)), (4, (2, 2))])
def test_reference_many_driver_verdict(exes, tmp_path, name, ranks, grid):
    r = run_driver(exes[name], tmp_path, ranks, grid, (64, 32, 48), nv=3)
    assert r.returncode == 0 and "Results are correct" in r.stdout and "incorrect" not in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("name,size", [("driver_sine_sp", (64, 64, 64)), ("driver_rand_sp", (32, 48, 40)), ("driver_sine", (14, 26, 38)),
                                       ("driver_sine_inplace", (32, 20, 24))])
def test_reference_driver_other_sizes_and_precision(exes, tmp_path, name, size):
    """single precision builds (-DSINGLE_PREC, tolerance 1e-5 in the driver) and the uneven grid of the reference's test matrix"""
    r = run_driver(exes[name], tmp_path, 2, (2, 1), size)
    assert r.returncode == 0 and "Results are correct" in r.stdout and "incorrect" not in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("name,ranks,grid,tol", [("driver_spec", 1, (1, 1), 1e-12), ("driver_spec_sp", 1, (1, 1), 1e-6),
                                                 ("driver_spec_sp", 4, (2, 2), 1e-6), ("driver_spec_sp", 2, (1, 2), 1e-6)])
def test_reference_driver_spec(exes, tmp_path, name, ranks, grid, tol):
    """driver_spec.c:223-250: forward transform, normalisation and shell-summed power spectrum of sin x sin y sin z.
    Several ranks only in single precision: the driver reduces its partial spectra with MPI_FLOAT in both builds
    (driver_spec.c:384), which is only right when E is float."""
    r = run_driver(exes[name], tmp_path, ranks, grid, (32, 32, 32))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    spec = {int(k): float(v) for k, v in re.findall(r"\((\d+)\.0 ([-+0-9.eEinfa]+)\)", r.stdout)}
    assert spec, r.stdout[-3000:]
    assert abs(spec[2] - 3.0 / 16.0) <= tol, spec
    assert all(abs(v) <= tol for k, v in spec.items() if k != 2), spec


def test_baseline_config1_driver_inverse_128_cubed_2x2(exes, tmp_path):
    """BASELINE.json configs[0]: sample/C driver_inverse, 128^3 double, 4 ranks on a 2x2 grid -- the reference's own
    correctness case, its own binary logic and its own check (four spikes -+N/4 at x = nx, driver_inverse.c:222-258)"""
    r = run_driver(exes["driver_inverse"], tmp_path, 4, (2, 2), (128, 128, 128))
    assert r.returncode == 0 and "Results are correct" in r.stdout and "incorrect" not in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    spikes = re.findall(r"\((\d+),(\d+),(\d+)\) (-?[0-9.]+)", r.stdout)
    got = sorted((int(a), int(b), int(c), float(v)) for a, b, c, v in spikes)
    n4 = 128 ** 3 / 4
    assert got == sorted([(128, 3, 4, -n4), (128, 3, 126, n4), (128, 127, 4, n4), (128, 127, 126, -n4)]), got


def test_shipped_reference_binaries_run_on_the_emulation(tmp_path):
    """oracle/_ref/drivers/* are the binaries that travel to the GPU box (linked with the PRODUCT library, run there by
    tests/test_zzzz_reference_binaries.py).  Here the very same files run on the CPU: LD_LIBRARY_PATH, which the loader
    searches before their RUNPATH, presents the emulated library under the product's name."""
    drv = os.path.join(ROOT, "oracle", "_ref", "drivers")
    if not os.path.exists(os.path.join(drv, "driver_inverse")):
        pytest.skip("oracle/_ref/drivers not built (python oracle/build_ref_drivers.py)")
    link = tmp_path / "lib"
    link.mkdir()
    os.symlink(os.path.join(LIB, "libp3dfft_emu.so"), link / "libp3dfft.so")
    os.symlink(os.path.join(LIB, "libp3dfft_emu_single.so"), link / "libp3dfft_single.so")
    for name, ranks, grid in (("driver_inverse", 4, (2, 2)), ("driver_sine_sp", 1, (1, 1)), ("driver_rand_many", 2, (1, 2))):
        r = run_driver(os.path.join(drv, name), tmp_path, ranks, grid, (64, 64, 64), nv=2 if "many" in name else None,
                       env={"LD_LIBRARY_PATH": str(link)})
        assert r.returncode == 0 and "Results are correct" in r.stdout and "incorrect" not in r.stdout, name + r.stdout[-2000:] + r.stderr[-2000:]
