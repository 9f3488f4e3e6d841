"""The reference's own sample drivers on the B200, through this library.

oracle/_ref/drivers/ holds the binaries that oracle/build_ref_drivers.py compiled in the build container from the UNCHANGED
sources of /root/reference/sample/C (which does not exist on the GPU box) against include/p3dfft.h + include/mpi_shim/mpi.h,
linked with the product library.  Here they run on the GPU under tools/p3drun.py -- one rank, and 2 / 4 ranks where the box
has the GPUs -- and must print their own verdict ("Results are correct", e.g. driver_sine.c:239-247; driver_inverse.c:222-258
checks the four spikes of its known answer).  Host arrays, i.e. the staged path of the C ABI, exactly as the drivers are written.
(Sorted last on purpose: this file depends on build artefacts of another container.)
"""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "oracle", "_ref", "drivers")
_port = [33750 + (os.getpid() % 83) * 2]

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(os.path.join(DRV, "driver_sine")),
                                                   reason="reference driver binaries not built (oracle/build_ref_drivers.py, build container only)")]


def run(name, tmp_path, ranks, grid, size, nv=None, env=None):
    (tmp_path / "stdin").write_text(" ".join(map(str, [*size, 2] + ([nv] if nv else []) + [1])) + "\n")
    (tmp_path / "dims").write_text(f"{grid[0]} {grid[1]}\n")
    _port[0] += 2
    cmd = [sys.executable, os.path.join(ROOT, "tools", "p3drun.py"), "-n", str(ranks), "--port", str(_port[0]), "--timeout", "240",
           os.path.join(DRV, name)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=str(tmp_path), env=dict(os.environ, **(env or {})))


def gpus():
    import torch
    return torch.cuda.device_count()


VERDICT = ["driver_sine", "driver_sine_inplace", "driver_rand", "driver_noop", "driver_inverse"]
MANY = ["driver_sine_many", "driver_sine_inplace_many", "driver_rand_many"]


@pytest.mark.parametrize("name", VERDICT + [n + "_sp" for n in VERDICT])
def test_reference_driver_on_gpu(tmp_path, name):
    r = run(name, tmp_path, 1, (1, 1), (128, 128, 128))        # the drivers' default size
    assert r.returncode == 0 and "Results are correct" in r.stdout and "incorrect" not in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("name", MANY)
def test_reference_many_driver_on_gpu(tmp_path, name):
    r = run(name, tmp_path, 1, (1, 1), (64, 32, 48), nv=3)
    assert r.returncode == 0 and "Results are correct" in r.stdout and "incorrect" not in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("name,tol", [("driver_spec", 1e-12), ("driver_spec_sp", 1e-6)])
def test_reference_driver_spec_on_gpu(tmp_path, name, tol):
    r = run(name, tmp_path, 1, (1, 1), (64, 64, 64))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    spec = {int(k): float(v) for k, v in re.findall(r"\((\d+)\.0 ([-+0-9.eEinfa]+)\)", r.stdout)}
    assert spec and abs(spec[2] - 3.0 / 16.0) <= tol and all(abs(v) <= tol for k, v in spec.items() if k != 2), spec


@pytest.mark.parametrize("name,ranks,grid", [("driver_inverse", 4, (2, 2)), ("driver_sine", 2, (1, 2)), ("driver_rand", 2, (2, 1)),
                                             ("driver_sine_inplace", 4, (1, 4)), ("driver_sine", 8, (2, 4))])
def test_reference_driver_on_several_gpus(tmp_path, name, ranks, grid):
    """BASELINE configs[0] is the first case: driver_inverse, 128^3, 4 ranks on a 2x2 grid"""
    if gpus() < ranks:
        pytest.skip(f"needs {ranks} GPUs")
    r = run(name, tmp_path, ranks, grid, (128, 128, 128))
    assert r.returncode == 0 and "Results are correct" in r.stdout and "incorrect" not in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
