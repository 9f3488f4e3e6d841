"""Pin the oracle on the known-answer checks of the reference's own sample drivers
(SURVEY.md section 8(c)).  Every check below cites the driver lines it restates; the pass
criterion is the drivers' own:  max|err| <= prec * Nglob * 0.25, prec = 1e-14 (double)
(sample/C/driver_sine.c:239-247)."""
import itertools

import numpy as np
import pytest

from oracle import p3dfft_oracle as po

GRIDS = [(1, 1), (2, 2), (1, 4), (4, 1), (2, 3)]


def _sine(nx, ny, nz, fx=1, fy=1, fz=1):
    x = np.sin(fx * 2 * np.pi * np.arange(nx) / nx)
    y = np.sin(fy * 2 * np.pi * np.arange(ny) / ny)
    z = np.sin(fz * 2 * np.pi * np.arange(nz) / nz)
    return np.asfortranarray(x[:, None, None] * y[None, :, None] * z[None, None, :])


@pytest.mark.parametrize("dims", GRIDS)
@pytest.mark.parametrize("n", [(32, 32, 32), (14, 26, 38)])
def test_map_data_and_tables_consistent(dims, n):
    """Send counts of rank a to b equal receive counts of b from a (setup.F90:481-518)."""
    nx, ny, nz = n
    P = dims[0] * dims[1]
    ds = [po.Decomp(nx, ny, nz, dims, r) for r in range(P)]
    for d in ds:
        row, col = d.row_ranks(), d.col_ranks()
        me_r, me_c = row.index(d.rank), col.index(d.rank)
        for p, peer in enumerate(row):
            assert d.IfSndCnts[p] == ds[peer].IfRcvCnts[me_r]
            assert d.KrSndCnts[p] == ds[peer].KrRcvCnts[me_r]
        for p, peer in enumerate(col):
            assert d.KfSndCnts[p] == ds[peer].KfRcvCnts[me_c]
            assert d.JrSndCnts[p] == ds[peer].JrRcvCnts[me_c]
        assert sum(d.iisz) == d.nxhpc and sum(d.jisz) == ny
        assert sum(d.jjsz) == d.nyc and sum(d.kjsz) == nz
        # the last (data mod proc) ranks are the bigger ones (setup.F90:608-635)
        assert d.jisz == sorted(d.jisz)


def test_c1_numbers_from_survey():
    """SURVEY section 8 anchor C1: 128^3 on 2x2: iisz={32,33}, padi_work=1 on ipid=1, memsize=(128,64,66)."""
    d = po.Decomp(128, 128, 128, (2, 2), 1)
    assert d.nxhp == 65 and d.iisz == [32, 33] and d.jisz == [64, 64]
    assert (d.ipid, d.jpid) == (1, 0)
    assert d.padi_work == 1 and d.nm == 65 * 64 * 65
    assert d.memsize == (128, 64, 66)
    d3 = po.Decomp(128, 128, 128, (2, 2), 2)
    assert (d3.ipid, d3.jpid) == (0, 1)


@pytest.mark.parametrize("dims", GRIDS)
def test_driver_sine_forward_spikes_and_roundtrip(dims):
    """driver_sine.c:168-181 (input), :203 + :311-321 (spikes of modulus N/8 at 1-based
    (2,{2,ny},{2,nz})), :218-228 (round trip after 1/N normalisation)."""
    nx = ny = nz = 32
    A = _sine(nx, ny, nz)
    w = po.SimWorld(nx, ny, nz, dims)
    F = w.gather_wave(w.forward(w.scatter_real(A), "fft"))
    N = nx * ny * nz
    big = np.argwhere(np.abs(F) > N * 1.25e-4)
    assert sorted(map(tuple, big.tolist())) == sorted(
        [(1, 1, 1), (1, 1, nz - 1), (1, ny - 1, 1), (1, ny - 1, nz - 1)])
    for idx in big:
        assert abs(abs(F[tuple(idx)]) - N / 8) < 1e-14 * N
    B = w.gather_real(w.backward(w.scatter_wave(F / N), "tff"))
    assert np.max(np.abs(B - A)) <= 1e-14 * N * 0.25


@pytest.mark.parametrize("dims", GRIDS)
def test_driver_inverse_known_answer(dims):
    """driver_inverse.c:326-360 (init_wave2: e^{ix} sin2y sin3z in Fourier space) and
    :222-240: c2r gives four spikes -+N/4 at 1-based x=nx, y in {3,ny-1}, z in {4,nz-2}."""
    nx = ny = nz = 32
    d0 = po.Decomp(nx, ny, nz, dims, 0)
    x = np.arange(d0.nxhp)
    y = np.sin(2.0 * np.arange(ny) * 2 * np.pi / ny)
    z = np.sin(3.0 * np.arange(nz) * 2 * np.pi / nz)
    Fg = (np.cos(x * 2 * np.pi / nx) + 1j * np.sin(x * 2 * np.pi / nx))[:, None, None] * \
        y[None, :, None] * z[None, None, :]
    w = po.SimWorld(nx, ny, nz, dims)
    B = w.gather_real(w.backward(w.scatter_wave(np.asfortranarray(Fg)), "tff"))
    N = nx * ny * nz
    exp = np.zeros_like(B)
    X = nx - 1
    exp[X, 2, 3] = -N * 0.25
    exp[X, 2, nz - 3] = +N * 0.25
    exp[X, ny - 2, 3] = +N * 0.25
    exp[X, ny - 2, nz - 3] = -N * 0.25
    assert np.max(np.abs(B - exp)) <= 1e-14 * N * 0.25
    # and the global definition agrees with the structural one
    Bg = po.global_backward(np.asfortranarray(Fg), d0, "tff")
    assert np.max(np.abs(Bg - exp)) <= 1e-14 * N * 0.25


@pytest.mark.parametrize("ops", [("ffn", "nff"), ("ffc", "cff"), ("ffs", "sff")])
@pytest.mark.parametrize("dims", [(1, 1), (2, 2)])
def test_driver_noop_and_r2r_roundtrip(dims, ops):
    """driver_noop.c:77,151 ('ffn' then 'nff'); DCT-I/DST-I round trips with their own
    normalisation 2(N-1) / 2(N+1) (FFTW REDFT00/RODFT00 definitions)."""
    nx, ny, nz = 16, 12, 9
    rng = np.random.default_rng(1)
    A = np.asfortranarray(rng.random((nx, ny, nz)))
    w = po.SimWorld(nx, ny, nz, dims)
    F = w.forward(w.scatter_real(A), ops[0])
    B = w.gather_real(w.backward(F, ops[1]))
    znorm = {"n": 1, "c": 2 * (nz - 1), "s": 2 * (nz + 1)}[ops[0][2]]
    assert np.max(np.abs(B / (nx * ny * znorm) - A)) <= 1e-13


@pytest.mark.parametrize("dims", [(1, 1), (2, 2), (1, 2)])
def test_driver_cheby_sin_to_cos(dims):
    """driver_cheby.F90:100-105 (Chebyshev nodes), :218-229 (sin(z) in, p3dfft_cheby),
    :236-254 (double first/last coefficient, halve), :256 btran 'cff', :258-285 (== cos)."""
    nx, ny, nz, Lz = 16, 16, 33, 2.0
    coordZ = np.cos(np.pi * np.arange(nz) / (nz - 1)) * 2.0 / Lz
    A = np.asfortranarray(np.broadcast_to(np.sin(coordZ)[None, None, :], (nx, ny, nz)).copy())
    w = po.SimWorld(nx, ny, nz, dims)
    C = w.cheby(w.scatter_real(A), Lz)
    for c in C:
        c[:, :, 0] *= 2.0
        c[:, :, nz - 1] *= 2.0
        c *= 0.5
    B = w.gather_real(w.backward(C, "cff"))
    N = nx * ny * nz
    assert np.max(np.abs(B - np.cos(coordZ)[None, None, :])) <= 1e-14 * N * 0.25


@pytest.mark.parametrize("dims", [(1, 1), (2, 2), (2, 1), (1, 2)])
def test_driver_sine_pruned_roundtrip(dims):
    """driver_sine_pruned.F90:96,153,192-229 and extra/makejob.py:130 ('64 64 64 32 32 32'):
    a low-mode sine survives forward(pruned) -> backward(zero-padded)."""
    nx = ny = nz = 32
    nc = 16
    A = _sine(nx, ny, nz)
    w = po.SimWorld(nx, ny, nz, dims, nxc=nc, nyc=nc, nzc=nc)
    F = w.forward(w.scatter_real(A), "fft")
    assert F[0].shape == (w.d[0].iisize, w.d[0].jjsize, nc)
    B = w.gather_real(w.backward([f / (nx * ny * nz) for f in F], "tff"))
    assert np.max(np.abs(B - A)) <= 1e-14 * nx * ny * nz * 0.25


@pytest.mark.parametrize("dims", GRIDS)
@pytest.mark.parametrize("n,cut", [((14, 26, 38), None), ((32, 20, 12), (16, 10, 8))])
def test_structural_equals_global_definition(dims, n, cut):
    """The simulated-rank stage sequence equals the mathematical definition on the uneven
    matrix of extra/makejob.py:146-152 (14x26x38) and a pruned case."""
    nx, ny, nz = n
    c = cut or (None, None, None)
    rng = np.random.default_rng(7)
    A = np.asfortranarray(rng.random((nx, ny, nz)))
    w = po.SimWorld(nx, ny, nz, dims, *c)
    d0 = w.d[0]
    for opf, opb in (("fft", "tff"), ("ffc", "cff"), ("ffs", "sff"), ("ffn", "nff")):
        Fs = w.gather_wave(w.forward(w.scatter_real(A), opf))
        Fg = po.global_forward(A, d0, opf)
        assert po.rel_l2(Fs, Fg) < 1e-14
        Bs = w.gather_real(w.backward(w.scatter_wave(Fg), opb))
        Bg = po.global_backward(Fg, d0, opb)
        assert po.rel_l2(Bs, Bg) < 1e-14


def test_single_precision_tolerance():
    nx = ny = nz = 32
    A = _sine(nx, ny, nz).astype(np.float32)
    w = po.SimWorld(nx, ny, nz, (2, 2), dtype=np.float32)
    F = w.forward(w.scatter_real(A), "fft")
    assert F[0].dtype == np.complex64
    N = nx * ny * nz
    B = w.gather_real(w.backward([f / N for f in F], "tff"))
    assert B.dtype == np.float32
    assert np.max(np.abs(B - A)) <= 1e-5 * N * 0.25


def test_philox_field_is_rank_count_independent():
    a = po.philox_field(8, 6, 5)
    b = po.philox_field(8, 6, 5, sl=(slice(2, 5), slice(1, 4)))
    assert np.array_equal(a[:, 2:5, 1:4], b)
    assert a.min() >= 0 and a.max() < 1


@pytest.mark.parametrize("dims", GRIDS)
def test_driver_spec_power_spectrum_known_answer(dims):
    """driver_spec.c:196-203 (sine-product input), :223 (1/N), :298-384 (shell sums, MPI_Reduce): the four spikes of
    modulus 1/8 stored in the half spectrum (kx = 1, ky = +-1, kz = +-1) have k^2 = 3, i.e. shell int(sqrt(3)+.5) = 2,
    so E(2) = 4 * 3 / 64 and every other shell is empty.  Summed over the ranks of any grid."""
    nx, ny, nz = 32, 24, 20
    A = _sine(nx, ny, nz)
    kmax = po.spectrum_kmax(nx, ny, nz)
    assert kmax == int(np.sqrt(nx * nx + ny * ny + nz * nz) * 0.5 + 0.5)
    E = np.zeros(kmax + 1)
    for r in range(dims[0] * dims[1]):
        d = po.Decomp(nx, ny, nz, dims, r)
        E += po.power_spectrum(po.local_forward(A, d, "fft"), d, kmax, 1.0 / (nx * ny * nz))
    assert abs(E[2] - 0.1875) < 1e-14
    E[2] = 0.0
    assert np.max(np.abs(E)) < 1e-25


def test_rtran_restatement_is_a_permutation():
    """module.F90:1061-1361: the transposes only move data -- the structural restatement (pack, alltoallv with the
    Ii/Ji/Ij/Kj tables of setup.F90:531-549, unpack) reproduces the global slices and x2y/y2x, x2z/z2x are inverses."""
    for dims, n in (((2, 3), (14, 26, 38)), ((3, 2), (9, 7, 5)), ((1, 4), (16, 12, 10)), ((4, 1), (16, 12, 10))):
        w = po.SimWorld(*n, dims)
        A = np.asfortranarray(np.arange(np.prod(n), dtype=np.float64).reshape(n, order="F"))
        parts = w.scatter_real(A)
        for fwd, bwd in (("x2y", "y2x"), ("x2z", "z2x")):
            o = w.rtran(fwd, parts)
            for d, x in zip(w.d, o):
                assert np.array_equal(x, po.rtran_local(A, d, fwd))
                assert list(x.shape) == po.rtran_dims(d, fwd)[2]
            for x, y in zip(w.rtran(bwd, o), parts):
                assert np.array_equal(x, y)
        # the destination pencils tile the global array exactly once
        for which in ("x2y", "x2z"):
            cover = np.zeros(n, dtype=int)
            for d in w.d:
                cover[po.rtran_slices(d, which)[1]] += 1
            assert np.all(cover == 1)
