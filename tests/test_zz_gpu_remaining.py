"""GPU tests of the routines added after the transform path: p3dfft_ftran_r2c_1d, the
real-data transposes, the process-map queries on a live plan, the fused output scaling and the power-spectrum
epilogue.  Single GPU here; the multi-rank paths run under torchrun (tests/mp_parity.py), launched below when the
box has at least two GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

import p3dfft_b200 as pb
from oracle import p3dfft_oracle as po

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture
def lib():
    L = pb.load(False)
    L.p3dfft_clean()
    L.set_layout(False, False)
    L.set_async(False)
    yield L
    L.p3dfft_clean()
    L.set_layout(False, False)


@pytest.mark.parametrize("n", [(64, 8, 6), (128, 16, 4), (30, 6, 5), (14, 26, 38), (1024, 4, 4)])
@pytest.mark.parametrize("device", [False, True])
def test_ftran_r2c_1d(lib, n, device):
    import torch
    nx, ny, nz = n
    lib.p3dfft_setup((1, 1), nx, ny, nz, 0)
    A = np.asfortranarray(np.random.default_rng(1).random(n))
    exp = po.forward_r2c_1d(A)
    if device:
        tA = torch.from_numpy(A.ravel(order="F").copy()).cuda()
        tC = torch.zeros(2 * exp.size, dtype=torch.float64, device="cuda")
        lib.p3dfft_ftran_r2c_1d(tA, tC)
        C = tC.cpu().numpy().view(np.complex128)
        assert np.array_equal(tA.cpu().numpy(), A.ravel(order="F"))
    else:
        C = np.zeros(exp.size, dtype=np.complex128)
        lib.p3dfft_ftran_r2c_1d(A, C)
    assert po.rel_l2(C, exp.ravel(order="F")) <= 1e-12


@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("n", [(16, 12, 10), (256, 64, 32), (9, 7, 5)])
def test_rtran_single_rank_is_a_copy(n, single):
    """one rank: both communicators have one member, every transpose is one direct copy (bit-exact)"""
    import torch
    L = pb.load(single)
    L.p3dfft_clean()
    L.set_layout(False, False)
    rt = np.float32 if single else np.float64
    nx, ny, nz = n
    L.p3dfft_setup((1, 1), nx, ny, nz, 0)
    try:
        A = np.asfortranarray(np.random.default_rng(3).random(n).astype(rt))
        d = po.Decomp(nx, ny, nz, (1, 1), 0)
        t = 0.0
        for which in pb.RTRAN_NAMES:
            src = torch.from_numpy(A.ravel(order="F").copy()).cuda()
            dst = torch.full((A.size,), float("nan"), dtype=src.dtype, device="cuda")
            L.launch_count(True)
            dstart, dend, dsize, t = L.rtran(which, src, dst, t)
            assert L.launch_count() == 1
            assert [list(dstart), list(dend), list(dsize)] == [list(x) for x in po.rtran_dims(d, which)]
            assert np.array_equal(dst.cpu().numpy(), A.ravel(order="F"))
            dsth = np.full(A.size, np.nan, dtype=rt)
            L.rtran(which, A, dsth)                      # host arrays: staged inside the call
            assert np.array_equal(dsth, A.ravel(order="F"))
        assert t == 0.0                                  # no exchange on one rank
    finally:
        L.p3dfft_clean()


def test_proc_queries_on_live_plan(lib):
    lib.p3dfft_setup((1, 1), 16, 12, 10, 0)
    g = po.ProcGrid(16, 12, 10, (1, 1))
    assert lib.p3dfft_get_mpi_info()[:2] == (0, 1)
    assert lib.proc_id2coords(0) == (0, 0) and lib.proc_coords2id(0, 0) == 0
    assert lib.proc_id2coords(1) is None and lib.proc_neighb(0, 1, 1) == -1
    for conf in (1, 2):
        assert lib.proc_dims(conf, 0) == [g.proc_dims[(conf, k, 0)] for k in range(1, 10)]
    exp = g.get_proc_parts(2, 3, 4, 5, 6, 7, 1)
    assert lib.get_proc_parts((2, 3, 4), (5, 6, 7), 1, 1) == (exp[0][:1], exp[1], exp[2])
    assert lib.get_proc_parts((1, 1, 1), (1, 1, 1), 3, 1)[2] == 1


@pytest.mark.parametrize("ngpu", [2, 4])
def test_multi_rank_parity_under_torchrun(ngpu):
    """tests/mp_parity.py: the reference's test matrix plus the transposes on every grid of `ngpu` ranks"""
    import torch
    if torch.cuda.device_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    port = 29600 + (os.getpid() % 50) * 2 + ngpu
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ngpu}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mp_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "MP PARITY PASS" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]


@pytest.mark.parametrize("n,cut,stride1", [((64, 64, 64), None, False), ((128, 64, 256), None, False), ((64, 64, 64), None, True),
                                           ((30, 18, 50), None, False), ((128, 128, 128), (64, 64, 64), False)])
def test_fused_scale(lib, n, cut, stride1):
    """p3dfft_b200_set_scale: the drivers' normalisation pass fused into the last stage's stores (specialised
    SCALED kernels for power-of-two lengths, the any-length kernel otherwise): forward * 1/N, backward * 2."""
    import torch
    nx, ny, nz = n
    c = cut or (None, None, None)
    lib.set_layout(stride1, False)
    lib.p3dfft_setup((1, 1), nx, ny, nz, 0, *c)
    d = po.Decomp(nx, ny, nz, (1, 1), 0, *c, stride1=stride1)
    A = np.asfortranarray(np.random.default_rng(8).random(n))
    N = float(nx * ny * nz)
    try:
        lib.set_scale(1.0 / N, 2.0)
        exp = po.local_forward(A, d, "fft")
        tA = torch.from_numpy(A.ravel(order="F").copy()).cuda()
        tF = torch.zeros(2 * exp.size, dtype=torch.float64, device="cuda")
        lib.p3dfft_ftran_r2c(tA, tF, "fft")
        F = tF.cpu().numpy().view(np.complex128)
        assert po.rel_l2(F, exp.ravel(order="F") / N) <= 1e-12
        Fg = po.global_forward(A, d, "fft")
        tB = torch.zeros(A.size, dtype=torch.float64, device="cuda")
        tFi = torch.from_numpy(np.asfortranarray(exp).ravel(order="F").view(np.float64).copy()).cuda()
        lib.p3dfft_btran_c2r(tFi, tB, "tff")
        assert po.rel_l2(tB.cpu().numpy(), 2.0 * po.local_backward(Fg, d, "tff").ravel(order="F")) <= 1e-12
    finally:
        lib.set_scale(1.0, 1.0)


@pytest.mark.parametrize("n,cut,stride1", [((32, 32, 32), None, False), ((64, 48, 40), None, False), ((32, 32, 32), None, True),
                                           ((64, 64, 64), (42, 42, 42), False), ((256, 128, 64), None, False)])
@pytest.mark.parametrize("device", [False, True])
def test_power_spectrum(lib, n, cut, stride1, device):
    """driver_spec.c: forward transform, normalise by 1/N, shell-sum k^2 |B|^2 -- here the last two on the device"""
    import torch
    nx, ny, nz = n
    c = cut or (None, None, None)
    lib.set_layout(stride1, False)
    lib.p3dfft_setup((1, 1), nx, ny, nz, 0, *c)
    d = po.Decomp(nx, ny, nz, (1, 1), 0, *c, stride1=stride1)
    x = np.sin(2 * np.pi * np.arange(nx) / nx)
    A = np.asfortranarray(x[:, None, None] * np.sin(2 * np.pi * np.arange(ny) / ny)[None, :, None]
                          * np.sin(2 * np.pi * np.arange(nz) / nz)[None, None, :]) + \
        0.01 * np.random.default_rng(4).random(n)
    A = np.asfortranarray(A)
    kmax = po.spectrum_kmax(nx, ny, nz)
    factor = 1.0 / (nx * ny * nz)
    exp_F = po.local_forward(A, d, "fft")
    expE = po.power_spectrum(exp_F, d, kmax, factor)
    if device:
        tA = torch.from_numpy(A.ravel(order="F").copy()).cuda()
        tF = torch.zeros(2 * exp_F.size, dtype=torch.float64, device="cuda")
        lib.p3dfft_ftran_r2c(tA, tF, "fft")
        tE = torch.zeros(kmax + 1, dtype=torch.float64, device="cuda")
        lib.spectrum(tF, kmax, factor, out=tE)
        E = tE.cpu().numpy()
    else:
        F = np.zeros(exp_F.shape, dtype=np.complex128, order="F")
        lib.p3dfft_ftran_r2c(A, F, "fft")
        E = lib.spectrum(F, kmax, factor)
    assert np.max(np.abs(E - expE)) <= 1e-12 * np.max(np.abs(expE))
    if cut is None:
        # the sine wave's four spikes (driver_sine.c:203) sit in shell ik = round(sqrt(3)) = 2
        assert np.argmax(E) == 2
