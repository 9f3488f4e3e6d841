"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Tolerances (BASELINE.json north_star): relative L2 <= 1e-12 in double, <= 1e-5 in single,
plus the sample drivers' own criterion max|err| <= prec*N/4 (driver_sine.c:239-247)."""
import numpy as np
import pytest

import p3dfft_b200 as pb
from oracle import p3dfft_oracle as po

pytestmark = pytest.mark.gpu
TOL = {False: 1e-12, True: 1e-5}


@pytest.fixture
def lib():
    L = pb.load(False)
    L.p3dfft_clean()
    L.set_layout(False, False)
    L.set_async(False)
    yield L
    L.p3dfft_clean()
    L.set_layout(False, False)


@pytest.fixture
def libf():
    L = pb.load(True)
    L.p3dfft_clean()
    L.set_layout(False, False)
    yield L
    L.p3dfft_clean()


def _rand(n, dtype=np.float64, seed=3):
    rng = np.random.default_rng(seed)
    return np.asfortranarray(rng.random(n).astype(dtype))


def _fwd_bwd(L, n, cut=None, opf="fft", opb="tff", stride1=False, single=False, device=False):
    import torch
    nx, ny, nz = n
    c = cut or (None, None, None)
    rt = np.float32 if single else np.float64
    ct = np.complex64 if single else np.complex128
    L.set_layout(stride1, False)
    L.p3dfft_setup((1, 1), nx, ny, nz, 0, *c)
    d = po.Decomp(nx, ny, nz, (1, 1), 0, *c, stride1=stride1, elem=4 if single else 8)
    ist, ien, isz = L.p3dfft_get_dims(1)
    fst, fen, fsz = L.p3dfft_get_dims(2)
    assert [list(ist), list(ien), list(isz)] == [list(v) for v in d.get_dims(1)]
    assert [list(fst), list(fen), list(fsz)] == [list(v) for v in d.get_dims(2)]
    assert list(isz) == d.get_dims(1)[2] and list(fsz) == d.get_dims(2)[2]
    A = _rand(n, rt)
    F = np.zeros(fsz, dtype=ct, order="F")
    exp_F = po.local_forward(A.astype(np.float64), d, opf)
    if device:
        tA = torch.from_numpy(A.ravel(order="F").copy()).cuda()
        tF = torch.zeros(int(np.prod(fsz)) * 2, dtype=torch.float32 if single else torch.float64, device="cuda")
        L.p3dfft_ftran_r2c(tA, tF, opf)
        F = tF.cpu().numpy().view(ct).reshape(fsz, order="F")
        assert np.array_equal(tA.cpu().numpy(), A.ravel(order="F")), "forward must not modify its input"
    else:
        L.p3dfft_ftran_r2c(A, F, opf)
    assert po.rel_l2(F, exp_F) <= TOL[single], (n, cut, opf)
    # backward from the oracle's spectrum
    Fg = np.asfortranarray(exp_F.astype(ct))
    B = np.zeros(isz, dtype=rt, order="F")
    Fglob = po.global_forward(A.astype(np.float64), d, opf)
    exp_B = po.local_backward(Fglob, d, opb)
    if device:
        tF = torch.from_numpy(Fg.ravel(order="F").view(rt).copy()).cuda()
        keep = tF.clone()
        tB = torch.zeros(B.size, dtype=tF.dtype, device="cuda")
        L.p3dfft_btran_c2r(tF, tB, opb)
        B = tB.cpu().numpy().reshape(isz, order="F")
        assert torch.equal(tF, keep), "backward never overwrites its input in this build"
    else:
        L.p3dfft_btran_c2r(Fg, B, opb)
    assert po.rel_l2(B, exp_B) <= TOL[single], (n, cut, opb)
    L.p3dfft_clean()


@pytest.mark.parametrize("n,cut", [((32, 32, 32), None), ((14, 26, 38), None), ((64, 64, 64), (32, 32, 32)),
                                   ((128, 128, 128), None), ((40, 24, 20), (20, 12, 10)), ((30, 18, 50), None),
                                   ((256, 64, 32), None), ((24, 256, 16), None), ((16, 24, 512), None)])
@pytest.mark.parametrize("device", [False, True])
def test_forward_backward_double(lib, n, cut, device):
    _fwd_bwd(lib, n, cut, device=device)


@pytest.mark.parametrize("ops", [("ffc", "cff"), ("ffs", "sff"), ("ffn", "nff")])
@pytest.mark.parametrize("n,cut", [((32, 32, 33), None), ((16, 12, 9), None), ((32, 16, 24), (16, 8, 12))])
def test_third_dimension_variants(lib, ops, n, cut):
    _fwd_bwd(lib, n, cut, *ops)


@pytest.mark.parametrize("n,cut", [((32, 32, 32), None), ((14, 26, 38), None), ((64, 32, 48), (32, 16, 24))])
def test_stride1_layout(lib, n, cut):
    _fwd_bwd(lib, n, cut, stride1=True)
    _fwd_bwd(lib, n, cut, "ffc", "cff", stride1=True, device=True)


@pytest.mark.parametrize("n,cut", [((32, 32, 32), None), ((14, 26, 38), None), ((128, 64, 32), (64, 32, 16))])
def test_single_precision(libf, n, cut):
    _fwd_bwd(libf, n, cut, single=True)
    _fwd_bwd(libf, n, cut, single=True, device=True)


FAST_CASES = [((64, 64, 64), None), ((128, 256, 64), None), ((256, 64, 128), None), ((512, 128, 64), None),
              ((1024, 64, 256), None), ((2048, 64, 64), None), ((64, 1024, 64), None), ((64, 64, 2048), None),
              ((64, 512, 512), None), ((128, 128, 128), (64, 64, 64)), ((256, 256, 64), (170, 170, 42)),
              ((512, 64, 64), (340, 42, 42))]


@pytest.mark.parametrize("n,cut", FAST_CASES)
def test_fast_kernels_double(lib, n, cut):
    """Power-of-two lengths run on the specialised kernels (fft_fast.cuh): all six stages."""
    lib.fast_launch_count(True)
    _fwd_bwd(lib, n, cut, device=True)
    assert lib.fast_launch_count() == 6
    lib.force_generic(True)           # A/B: the any-length kernel on the same input
    try:
        lib.fast_launch_count(True)
        _fwd_bwd(lib, n, cut, device=True)
        assert lib.fast_launch_count() == 0
    finally:
        lib.force_generic(False)


NONPOW2_CASES = [((384, 64, 64), None), ((768, 64, 128), None), ((1536, 64, 64), None), ((640, 128, 64), None), ((1280, 64, 64), None),
                 ((64, 384, 64), None), ((64, 64, 768), None), ((64, 1536, 64), None), ((128, 640, 64), None), ((64, 64, 1280), None),
                 ((384, 384, 384), None), ((768, 384, 640), (512, 256, 426)), ((64, 64, 385), None)]


@pytest.mark.parametrize("n,cut", NONPOW2_CASES)
def test_fast_kernels_non_power_of_two(lib, libf, n, cut):
    """3 * 2^k and 5 * 2^k lengths (the DNS-typical 384, 768, 1536, 640, 1280) on the specialised kernels, every stage, double
    and single precision; the last case is the DCT-I of length 385 (a 768-point transform)"""
    ops = ("ffc", "cff") if n[2] % 2 else ("fft", "tff")
    for L, single in ((lib, False), (libf, True)):
        L.fast_launch_count(True)
        _fwd_bwd(L, n, cut, single=single, device=True, opf=ops[0], opb=ops[1])
        assert L.fast_launch_count() == 6


@pytest.mark.parametrize("n,cut", [FAST_CASES[0], FAST_CASES[1], FAST_CASES[8], FAST_CASES[10], ((64, 1024, 1024), None)])
def test_fast_kernels_row_bytes(lib, n, cut):
    """Both tile row widths of the internal layouts (64- and 128-byte rows) give the same transform."""
    for rb in (64, 128):
        lib.row_bytes(rb)
        try:
            lib.fast_launch_count(True)
            _fwd_bwd(lib, n, cut, device=True)
            assert lib.fast_launch_count() == 6
        finally:
            lib.row_bytes(0)


@pytest.mark.parametrize("n,cut", FAST_CASES[:9] + FAST_CASES[10:11])
def test_fast_kernels_single(libf, n, cut):
    libf.fast_launch_count(True)
    _fwd_bwd(libf, n, cut, single=True, device=True)
    assert libf.fast_launch_count() == 6


@pytest.mark.parametrize("stride1", [False, True])
@pytest.mark.parametrize("n,cut", [((64, 64, 33), None), ((128, 64, 65), None), ((64, 128, 129), None),
                                   ((64, 64, 513), None), ((128, 128, 65), (64, 64, 33))])
def test_fast_kernels_dct(lib, n, cut, stride1):
    """Chebyshev third dimension: DCT-I of odd nz as an even-extended FFT of length 2(nz-1)."""
    lib.fast_launch_count(True)
    _fwd_bwd(lib, n, cut, "ffc", "cff", stride1=stride1, device=True)
    assert lib.fast_launch_count() == 6


@pytest.mark.parametrize("stride1", [False, True])
@pytest.mark.parametrize("n,cut", [((64, 64, 31), None), ((128, 64, 63), None), ((64, 128, 127), (32, 64, 64)), ((64, 64, 511), None),
                                   ((64, 64, 1023), None), ((64, 64, 383), None)])
def test_fast_kernels_dst(lib, n, cut, stride1):
    """Sine third dimension: DST-I of nz = 2^k - 1 (and 3 * 2^k - 1) points as an odd-extended FFT of length 2 (nz + 1) on the
    specialised c2c kernel (fft_exec.F90:866-921); nz = 1023 takes 64-byte rows (2048-point tile)."""
    lib.fast_launch_count(True)
    _fwd_bwd(lib, n, cut, "ffs", "sff", stride1=stride1, device=True)
    assert lib.fast_launch_count() == 6


def test_fast_kernels_dst_single(libf):
    libf.fast_launch_count(True)
    _fwd_bwd(libf, (64, 128, 255), None, "ffs", "sff", single=True, device=True)
    assert libf.fast_launch_count() == 6


@pytest.mark.parametrize("n", [(64, 64, 64), (128, 64, 256)])
def test_fast_kernels_stride1(lib, n):
    lib.fast_launch_count(True)
    _fwd_bwd(lib, n, None, stride1=True, device=True)
    assert lib.fast_launch_count() == 6


def test_driver_sine_known_answer_and_roundtrip(lib):
    """driver_sine.c: spikes of modulus N/8 at 1-based (2,{2,ny},{2,nz}); round trip <= 1e-14*N/4."""
    nx = ny = nz = 64
    x = np.sin(2 * np.pi * np.arange(nx) / nx)
    A = np.asfortranarray(x[:, None, None] * np.sin(2 * np.pi * np.arange(ny) / ny)[None, :, None]
                          * np.sin(2 * np.pi * np.arange(nz) / nz)[None, None, :])
    lib.p3dfft_setup((1, 1), nx, ny, nz, 0, nx, ny, nz, True)
    _, _, fsz = lib.p3dfft_get_dims(2)
    F = np.zeros(fsz, dtype=np.complex128, order="F")
    lib.p3dfft_ftran_r2c(A, F, "fft")
    N = nx * ny * nz
    big = sorted(map(tuple, np.argwhere(np.abs(F) > N * 1.25e-4).tolist()))
    assert big == sorted([(1, 1, 1), (1, 1, nz - 1), (1, ny - 1, 1), (1, ny - 1, nz - 1)])
    assert all(abs(abs(F[i]) - N / 8) < 1e-14 * N for i in big)
    F *= 1.0 / N
    C = np.zeros((nx, ny, nz), order="F")
    lib.p3dfft_btran_c2r(F, C, "tff")
    assert np.max(np.abs(C - A)) <= 1e-14 * N * 0.25


def test_driver_inverse_known_answer(lib):
    """driver_inverse.c:222-240: c2r of e^{ix} sin2y sin3z -> four spikes -+N/4 at x=nx."""
    nx = ny = nz = 64
    lib.p3dfft_setup((1, 1), nx, ny, nz, 0, nx, ny, nz, True)
    xs = np.arange(nx // 2 + 1)
    Fg = (np.cos(xs * 2 * np.pi / nx) + 1j * np.sin(xs * 2 * np.pi / nx))[:, None, None] * \
        np.sin(2.0 * np.arange(ny) * 2 * np.pi / ny)[None, :, None] * \
        np.sin(3.0 * np.arange(nz) * 2 * np.pi / nz)[None, None, :]
    Fg = np.asfortranarray(Fg)
    B = np.zeros((nx, ny, nz), order="F")
    lib.p3dfft_btran_c2r(Fg, B, "tff")
    N = nx * ny * nz
    exp = np.zeros_like(B)
    exp[nx - 1, 2, 3] = -N / 4
    exp[nx - 1, 2, nz - 3] = N / 4
    exp[nx - 1, ny - 2, 3] = N / 4
    exp[nx - 1, ny - 2, nz - 3] = -N / 4
    assert np.max(np.abs(B - exp)) <= 1e-14 * N * 0.25


def _inplace_sine(L, n, stride1, arrays=None, iterations=2):
    """driver_sine_inplace.c:183-232: ONE array of max(real, 2 x complex) elements is input and output of both transforms
    (Cp3dfft_ftran_r2c(A,A,op_f) ... mult_array ... Cp3dfft_btran_c2r(A,A,op_b)), `iterations` times; the result must be the
    initial sine field to 1e-14*N/4.  arrays=None: a host array (staged); otherwise the adapter that makes device arrays."""
    nx, ny, nz = n
    L.set_layout(stride1, False)
    L.p3dfft_setup((1, 1), nx, ny, nz, 0, nx, ny, nz, True)
    _, _, isz = L.p3dfft_get_dims(1)
    _, _, fsz = L.p3dfft_get_dims(2)
    nm = max(int(np.prod(isz)), 2 * int(np.prod(fsz)))
    field = (np.sin(2 * np.pi * np.arange(nx) / nx)[:, None, None] * np.sin(2 * np.pi * np.arange(ny) / ny)[None, :, None]
             * np.sin(2 * np.pi * np.arange(nz) / nz)[None, None, :])
    host = np.zeros(nm)
    host[: nx * ny * nz] = field.ravel(order="F")
    A = arrays.dev(host) if arrays else host
    N = nx * ny * nz
    for _ in range(iterations):
        L.p3dfft_ftran_r2c(A, A, "fft")
        if arrays:
            A *= 1.0 / N                      # mult_array on the device array
            arrays.sync()
        else:
            A[: 2 * int(np.prod(fsz))] *= 1.0 / N
        L.p3dfft_btran_c2r(A, A, "tff")
    out = np.asarray(arrays.host(A) if arrays else A)[: nx * ny * nz].reshape((nx, ny, nz), order="F")
    L.p3dfft_clean()
    assert np.max(np.abs(out - field)) <= 1e-14 * N * 0.25


@pytest.mark.parametrize("stride1", [False, True])
@pytest.mark.parametrize("device", [False, True])
def test_driver_sine_inplace(lib, stride1, device):
    from tests import mp_parity as M
    _inplace_sine(lib, (64, 64, 64), stride1, M.TorchArrays() if device else None)
    _inplace_sine(lib, (30, 18, 50), stride1, M.TorchArrays() if device else None)


@pytest.mark.parametrize("stride1", [False, True])
def test_driver_cheby_sin_to_cos(lib, stride1):
    """driver_cheby.F90:218-285 with its in-place call (mem aliased for in and out)."""
    nx, ny, nz, Lz = 32, 32, 33, 2.0
    lib.set_layout(stride1, False)
    mem = lib.p3dfft_setup((1, 1), nx, ny, nz, 0)
    coordZ = np.cos(np.pi * np.arange(nz) / (nz - 1)) * 2.0 / Lz
    buf = np.zeros(int(np.prod(mem)), dtype=np.float64)
    buf[: nx * ny * nz] = np.broadcast_to(np.sin(coordZ)[None, None, :], (nx, ny, nz)).ravel(order="F")
    lib.p3dfft_cheby(buf, buf, Lz)            # in place: real in, complex out at the same address
    _, _, fsz = lib.p3dfft_get_dims(2)
    C = buf[: 2 * int(np.prod(fsz))].view(np.complex128).reshape(fsz, order="F")
    d = po.Decomp(nx, ny, nz, (1, 1), 0, stride1=stride1)
    A = np.asfortranarray(np.broadcast_to(np.sin(coordZ)[None, None, :], (nx, ny, nz)).copy())
    expC = po.global_cheby(A, d, Lz)
    if stride1:
        expC = expC.transpose(2, 1, 0)
    assert po.rel_l2(C, expC) <= 1e-12
    if stride1:
        C[0] *= 2.0
        C[nz - 1] *= 2.0
    else:
        C[:, :, 0] *= 2.0
        C[:, :, nz - 1] *= 2.0
    C *= 0.5
    lib.p3dfft_btran_c2r(buf, buf, "cff")
    R = buf[: nx * ny * nz].reshape((nx, ny, nz), order="F")
    assert np.max(np.abs(R - np.cos(coordZ)[None, None, :])) <= 1e-14 * nx * ny * nz * 0.25


@pytest.mark.parametrize("stride1", [False, True])
def test_many_variables(lib, stride1):
    """driver_sine_many.c / driver_rand_many.c: nv variables, dim_in/dim_out strides."""
    nx, ny, nz, nv = 32, 24, 16, 3
    lib.set_layout(stride1, False)
    lib.p3dfft_setup((1, 1), nx, ny, nz, 0)
    d = po.Decomp(nx, ny, nz, (1, 1), 0, stride1=stride1)
    _, _, fsz = lib.p3dfft_get_dims(2)
    dim_in = nx * ny * nz + 7
    dim_out = int(np.prod(fsz)) + 5
    A = np.zeros((dim_in, nv), order="F")
    fields = [_rand((nx, ny, nz), seed=20 + v) for v in range(nv)]
    for v in range(nv):
        A[: nx * ny * nz, v] = fields[v].ravel(order="F")
    F = np.zeros((dim_out, nv), dtype=np.complex128, order="F")
    lib.p3dfft_ftran_r2c_many(A, dim_in, F, dim_out, nv, "fft")
    for v in range(nv):
        exp = po.local_forward(fields[v], d, "fft")
        assert po.rel_l2(F[: exp.size, v], exp.ravel(order="F")) <= 1e-12
    B = np.zeros((dim_in, nv), order="F")
    lib.p3dfft_btran_c2r_many(F, dim_out, B, dim_in, nv, "tff")
    N = nx * ny * nz
    for v in range(nv):
        assert np.max(np.abs(B[:N, v] / N - fields[v].ravel(order="F"))) <= 1e-13
    # cheby_many needs odd nz >= 3 along z: reuse with nz=17
    lib.p3dfft_clean()
    nz = 17
    lib.p3dfft_setup((1, 1), nx, ny, nz, 0)
    d = po.Decomp(nx, ny, nz, (1, 1), 0, stride1=stride1)
    _, _, fsz = lib.p3dfft_get_dims(2)
    dim_in, dim_out = nx * ny * nz, int(np.prod(fsz))
    fields = [_rand((nx, ny, nz), seed=40 + v) for v in range(2)]
    A = np.asfortranarray(np.stack([f.ravel(order="F") for f in fields], axis=1))
    F = np.zeros((dim_out, 2), dtype=np.complex128, order="F")
    lib.p3dfft_cheby_many(A, dim_in, F, dim_out, 2, 3.0)
    for v in range(2):
        exp = po.global_cheby(fields[v], d, 3.0)
        if stride1:
            exp = exp.transpose(2, 1, 0)
        assert po.rel_l2(F[:, v], np.asfortranarray(exp).ravel(order="F")) <= 1e-12


def test_error_behaviour(lib):
    """call before setup -> message and return (ftran.F90:506-509); second setup without
    clean -> error (setup.F90:130-135); unknown op letter (ftran.F90:640-643)."""
    A = np.zeros(8)
    with pytest.raises(RuntimeError, match="call setup before other routines"):
        lib.p3dfft_ftran_r2c(A, A, "fft")
    lib.p3dfft_setup((1, 1), 8, 8, 8, 0)
    with pytest.raises(RuntimeError, match="already initialized"):
        lib.p3dfft_setup((1, 1), 8, 8, 8, 0)
    F = np.zeros((5, 8, 8), dtype=np.complex128, order="F")
    A = np.zeros((8, 8, 8), order="F")
    with pytest.raises(RuntimeError, match="Unknown transform type"):
        lib.p3dfft_ftran_r2c(A, F, "ffx")
    with pytest.raises(RuntimeError, match="Invalid processor geometry"):
        lib.p3dfft_clean()
        lib.p3dfft_setup((2, 2), 8, 8, 8, 0)
    lib.p3dfft_clean()
    lib.p3dfft_setup((1, 1), 8, 8, 8, 0)       # setup works again after clean (module.F90:309)
    t0 = lib.get_timers()
    lib.p3dfft_ftran_r2c(A, F, "fft")
    t1 = lib.get_timers()
    assert t1[4] > t0[4] and t1[6] > t0[6] and t1[7] > t0[7]   # timers 5, 7, 8
    lib.set_timers()
    assert lib.get_timers() == [0.0] * 12


def test_large_roundtrip_properties(lib):
    """BASELINE config 2 size (512^3 double, 1x1): round trip and Parseval, device arrays."""
    import torch
    n = 512
    lib.p3dfft_setup((1, 1), n, n, n, 0)
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.rand(n * n * n, dtype=torch.float64, device="cuda", generator=g)
    F = torch.empty((n // 2 + 1) * n * n * 2, dtype=torch.float64, device="cuda")
    B = torch.empty_like(A)
    lib.p3dfft_ftran_r2c(A, F, "fft")
    lib.p3dfft_btran_c2r(F, B, "tff")
    N = float(n) ** 3
    err = (B / N - A).abs().max().item()
    assert err <= 1e-14 * N * 0.25 and err < 1e-12
    # Parseval on the half spectrum: sum |A|^2 * N == sum w_k |F_k|^2, w = 1 for kx in {0, n/2}, else 2
    Fc = torch.view_as_complex(F.view(-1, 2)).view(n, n, n // 2 + 1)   # C-order view: (z, y, x)
    p = (Fc.real ** 2 + Fc.imag ** 2).sum(dim=(0, 1))
    w = torch.full((n // 2 + 1,), 2.0, dtype=torch.float64, device="cuda")
    w[0] = 1.0
    w[-1] = 1.0
    lhs = (A ** 2).sum().item() * N
    rhs = (p * w).sum().item()
    assert abs(lhs - rhs) / lhs < 1e-12
    # spot-check a few x-lines of the spectrum against numpy on the host
    A3 = A.view(n, n, n)          # (z, y, x)
    F0 = torch.fft.rfft(A3[0:1, :, :].cpu(), dim=2)     # partial transform only in x: consistency of DC line
    dc = A.sum().item()
    assert abs(Fc[0, 0, 0].real.item() - dc) / abs(dc) < 1e-12
