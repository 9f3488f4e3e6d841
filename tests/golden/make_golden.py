#!/usr/bin/env python
"""Writes tests/golden/known_answers.npz -- ANALYTIC known answers of the reference's sample drivers.

The reference cannot be built in this image (no Fortran compiler, MPI or FFTW) and
ships no golden files, so these vectors are not outputs of the reference binary: they are the closed-form
answers its own drivers test against, evaluated here with numpy from the drivers' formulas only (no FFT is
computed by this script; neither the oracle nor the CUDA library is imported).

  inverse_*   sample/C/driver_inverse.c:209,222-240,326-360   c2r of e^{ikx} sin(2y) sin(3z): four spikes -+N/4
  sine_*      sample/C/driver_sine.c:168-181,203,311-321      r2c of sin x sin y sin z: four spikes of modulus N/8
  cheby_*     sample/FORTRAN/driver_cheby.F90:218-285          Chebyshev derivative of sin(z) is cos(z)
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def inverse_case(n):
    nx = ny = nz = n
    N = nx * ny * nz
    xs = np.arange(nx // 2 + 1)
    # driver_inverse.c:222-240: F(x,y,z) = (cos(x dx) + i sin(x dx)) * sin(2 y dy) * sin(3 z dz) on the half spectrum
    F = (np.cos(xs * 2 * np.pi / nx) + 1j * np.sin(xs * 2 * np.pi / nx))[:, None, None] * \
        np.sin(2.0 * np.arange(ny) * 2 * np.pi / ny)[None, :, None] * \
        np.sin(3.0 * np.arange(nz) * 2 * np.pi / nz)[None, None, :]
    # driver_inverse.c:326-360: the only non-zero outputs, at 1-based x = nx
    B = np.zeros((nx, ny, nz))
    B[nx - 1, 2, 3] = -N / 4
    B[nx - 1, 2, nz - 3] = N / 4
    B[nx - 1, ny - 2, 3] = N / 4
    B[nx - 1, ny - 2, nz - 3] = -N / 4
    return F, B


def sine_case(n):
    nx = ny = nz = n
    N = nx * ny * nz
    s = lambda m: np.sin(2 * np.pi * np.arange(m) / m)
    A = s(nx)[:, None, None] * s(ny)[None, :, None] * s(nz)[None, None, :]
    # sin x sin y sin z = product of (e^{i.} - e^{-i.})/(2i): modes (+-1,+-1,+-1), coefficient N * (1/(2i))^3 * signs
    F = np.zeros((nx // 2 + 1, ny, nz), dtype=np.complex128)
    for sy, ky in ((1, 1), (-1, ny - 1)):
        for sz, kz in ((1, 1), (-1, nz - 1)):
            F[1, ky, kz] = N * (1 / 2j) ** 3 * sy * sz
    return A, F


def cheby_case(nx, ny, nz, Lz):
    # driver_cheby.F90:218-230: f = sin(z) on the Chebyshev nodes z_k = cos(pi k/(nz-1)) * 2/Lz ... derivative is cos(z)
    z = np.cos(np.pi * np.arange(nz) / (nz - 1)) * 2.0 / Lz
    A = np.broadcast_to(np.sin(z)[None, None, :], (nx, ny, nz)).copy()
    D = np.broadcast_to(np.cos(z)[None, None, :], (nx, ny, nz)).copy()
    return A, D


if __name__ == "__main__":
    out = {}
    for n in (16, 32):
        F, B = inverse_case(n)
        out[f"inverse_in_{n}"], out[f"inverse_out_{n}"] = F, B
        A, S = sine_case(n)
        out[f"sine_in_{n}"], out[f"sine_out_{n}"] = A, S
    A, D = cheby_case(8, 8, 33, 2.0)
    out["cheby_in"], out["cheby_deriv"], out["cheby_Lz"] = A, D, np.float64(2.0)
    np.savez_compressed(os.path.join(HERE, "known_answers.npz"), **out)
    print({k: v.shape for k, v in out.items()})
