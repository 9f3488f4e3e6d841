"""Golden known-answer fixtures (tests/golden/known_answers.npz, written by tests/golden/make_golden.py
from the closed-form answers of the reference's sample drivers) against the oracle (CPU) and against the
CUDA path through the C ABI (GPU).  Criterion: the drivers' own max|err| <= 1e-14 * N / 4
(sample/C/driver_sine.c:239-247)."""
import os

import numpy as np
import pytest

import p3dfft_b200 as pb
from oracle import p3dfft_oracle as po

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "known_answers.npz"))
GRIDS = [(1, 1), (2, 2), (1, 4), (4, 1)]


@pytest.mark.parametrize("n", [16, 32])
@pytest.mark.parametrize("dims", GRIDS)
def test_oracle_matches_golden(n, dims):
    N = n ** 3
    tol = 1e-14 * N * 0.25
    d0 = po.Decomp(n, n, n, dims, 0)
    assert np.max(np.abs(po.global_backward(G[f"inverse_in_{n}"], d0, "tff") - G[f"inverse_out_{n}"])) <= tol
    assert np.max(np.abs(po.global_forward(G[f"sine_in_{n}"], d0, "fft") - G[f"sine_out_{n}"])) <= tol
    # the structural (pack / alltoallv / unpack) restatement on simulated ranks gives the same answers
    W = po.SimWorld(n, n, n, dims)
    out = W.gather_wave(W.forward(W.scatter_real(np.asfortranarray(G[f"sine_in_{n}"])), "fft"))
    assert np.max(np.abs(out - G[f"sine_out_{n}"])) <= tol
    back = W.gather_real(W.backward(W.scatter_wave(np.asfortranarray(G[f"inverse_in_{n}"])), "tff"))
    assert np.max(np.abs(back - G[f"inverse_out_{n}"])) <= tol


def test_oracle_cheby_matches_golden():
    A, D, Lz = G["cheby_in"], G["cheby_deriv"], float(G["cheby_Lz"])
    nx, ny, nz = A.shape
    d = po.Decomp(nx, ny, nz, (1, 1), 0)
    C = po.global_cheby(np.asfortranarray(A), d, Lz)            # Chebyshev coefficients of df/dz, Fourier in x,y
    # driver_cheby.F90:258-285: halve the end coefficients, backward 'cff', compare with cos(z)
    C = C.copy()
    C[:, :, 0] *= 2.0
    C[:, :, nz - 1] *= 2.0
    C *= 0.5
    R = po.global_backward(C, d, "cff")
    assert np.max(np.abs(R - D)) <= 1e-14 * nx * ny * nz * 0.25


@pytest.mark.gpu
@pytest.mark.parametrize("n", [16, 32])
def test_cuda_matches_golden(n):
    L = pb.load(False)
    L.p3dfft_clean()
    L.set_layout(False, False)
    N = n ** 3
    tol = 1e-14 * N * 0.25
    L.p3dfft_setup((1, 1), n, n, n, 0)
    F = np.zeros((n // 2 + 1, n, n), dtype=np.complex128, order="F")
    L.p3dfft_ftran_r2c(np.asfortranarray(G[f"sine_in_{n}"]), F, "fft")
    assert np.max(np.abs(F - G[f"sine_out_{n}"])) <= tol
    B = np.zeros((n, n, n), order="F")
    L.p3dfft_btran_c2r(np.asfortranarray(G[f"inverse_in_{n}"]), B, "tff")
    assert np.max(np.abs(B - G[f"inverse_out_{n}"])) <= tol
    L.p3dfft_clean()


@pytest.mark.gpu
def test_cuda_cheby_matches_golden():
    A, D, Lz = G["cheby_in"], G["cheby_deriv"], float(G["cheby_Lz"])
    nx, ny, nz = A.shape
    L = pb.load(False)
    L.p3dfft_clean()
    L.set_layout(False, False)
    L.p3dfft_setup((1, 1), nx, ny, nz, 0)
    C = np.zeros((nx // 2 + 1, ny, nz), dtype=np.complex128, order="F")
    L.p3dfft_cheby(np.asfortranarray(A), C, Lz)
    C[:, :, 0] *= 2.0
    C[:, :, nz - 1] *= 2.0
    C *= 0.5
    R = np.zeros((nx, ny, nz), order="F")
    L.p3dfft_btran_c2r(C, R, "cff")
    L.p3dfft_clean()
    assert np.max(np.abs(R - D)) <= 1e-14 * nx * ny * nz * 0.25
