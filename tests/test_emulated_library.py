"""The whole C-ABI library on the CPU.

libp3dfft_emu[_single].so (tests/emu) is the product's api.cpp, planner and every kernel file compiled by g++ against a
mock CUDA runtime and the execution-model emulation of tests/emu/cuda_emu.h: "device memory" is host memory, streams run
at enqueue time, kernels run as OS threads.  These tests drive it through the same Python binding and the same helper
functions as the GPU parity suite (tests/test_gpu_parity.py), so the executor (run_plan), the staging of host arrays,
the zero-copy path for "device" arrays, scaling, epilogues, the auxiliary routines and the reference's error behaviour
are exercised in the CPU test-suite against the oracle -- with small sizes, since every CUDA thread is an OS thread.
Single rank only; nothing here says anything about speed."""
import ctypes as C
import os

import numpy as np
import pytest

import p3dfft_b200 as pb
from oracle import p3dfft_oracle as po
from tests import test_gpu_parity as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_cache = {}


def emulib(single=False):
    if single not in _cache:
        path = os.path.join(ROOT, "tests", "emu", "lib", "libp3dfft_emu_single.so" if single else "libp3dfft_emu.so")
        L = pb.P3DFFT(single, path=path)
        L.lib.emu_register_device_range.argtypes = [C.c_void_p, C.c_size_t]
        L.lib.emu_unregister_device_range.argtypes = [C.c_void_p]
        _cache[single] = L
    return _cache[single]


@pytest.fixture
def lib():
    L = emulib(False)
    L.p3dfft_clean()
    L.set_layout(False, False)
    L.set_async(False)
    L.set_scale(1.0, 1.0)
    yield L
    L.p3dfft_clean()
    L.set_layout(False, False)
    L.set_scale(1.0, 1.0)


class device_arrays:
    """numpy arrays registered as device memory with the mock runtime: the library uses them in place (no staging)"""

    def __init__(self, L, *arrays):
        self.L, self.arrays = L, arrays

    def __enter__(self):
        for a in self.arrays:
            self.L.lib.emu_register_device_range(a.ctypes.data, a.nbytes)
        return self.arrays

    def __exit__(self, *exc):
        for a in self.arrays:
            self.L.lib.emu_unregister_device_range(a.ctypes.data)


@pytest.mark.parametrize("n,cut", [((32, 32, 32), None), ((14, 26, 38), None), ((64, 64, 64), (32, 32, 32)), ((40, 24, 20), (20, 12, 10)),
                                   ((128, 64, 64), None)])
def test_forward_backward_host_arrays(lib, n, cut):
    G._fwd_bwd(lib, n, cut, device=False)


@pytest.mark.parametrize("ops", [("ffc", "cff"), ("ffs", "sff"), ("ffn", "nff")])
@pytest.mark.parametrize("n,cut", [((16, 12, 9), None), ((32, 16, 24), (16, 8, 12))])
def test_third_dimension_variants(lib, ops, n, cut):
    G._fwd_bwd(lib, n, cut, *ops)


@pytest.mark.parametrize("n,cut", [((32, 32, 32), None), ((64, 32, 48), (32, 16, 24))])
def test_stride1_layout(lib, n, cut):
    G._fwd_bwd(lib, n, cut, stride1=True)
    G._fwd_bwd(lib, n, cut, "ffc", "cff", stride1=True)


def test_single_precision():
    Lf = emulib(True)
    Lf.p3dfft_clean()
    Lf.set_layout(False, False)
    try:
        G._fwd_bwd(Lf, (64, 64, 64), None, single=True)
        G._fwd_bwd(Lf, (14, 26, 38), None, single=True)
    finally:
        Lf.p3dfft_clean()


def test_device_arrays_are_used_in_place(lib):
    """the zero-copy path: no staging buffers, the input is left untouched, the specialised kernels take every stage"""
    n = (64, 64, 64)
    lib.p3dfft_setup((1, 1), *n, 0)
    d = po.Decomp(*n, (1, 1), 0)
    A = np.asfortranarray(np.random.default_rng(3).random(n))
    keep = A.copy()
    F = np.zeros((d.nxhp, n[1], n[2]), dtype=np.complex128, order="F")
    B = np.zeros(n, order="F")
    with device_arrays(lib, A, F, B):
        lib.fast_launch_count(True)
        lib.p3dfft_ftran_r2c(A, F, "fft")
        lib.p3dfft_btran_c2r(F, B, "tff")
        assert lib.fast_launch_count() == 6
    assert np.array_equal(A, keep)
    assert po.rel_l2(F, po.local_forward(A, d, "fft")) <= 1e-13
    assert np.max(np.abs(B / A.size - A)) <= 1e-13


test_driver_sine_known_answer_and_roundtrip = G.test_driver_sine_known_answer_and_roundtrip
test_driver_inverse_known_answer = G.test_driver_inverse_known_answer
test_error_behaviour = G.test_error_behaviour


@pytest.mark.parametrize("stride1", [False, True])
@pytest.mark.parametrize("device", [False, True])
def test_driver_sine_inplace(lib, stride1, device):
    """driver_sine_inplace.c on the emulated library: host array (staged) and "device" array (used in place)"""
    from tests import mp_parity as M
    G._inplace_sine(lib, (64, 64, 64), stride1, M.EmuArrays(lib) if device else None)
    G._inplace_sine(lib, (30, 18, 50), stride1, M.EmuArrays(lib) if device else None)


@pytest.mark.parametrize("stride1", [False, True])
def test_driver_cheby_in_place(lib, stride1):
    G.test_driver_cheby_sin_to_cos(lib, stride1)


@pytest.mark.parametrize("stride1", [False, True])
def test_many_variables(lib, stride1):
    G.test_many_variables(lib, stride1)


@pytest.mark.parametrize("n,cut,stride1", [((64, 64, 64), None, False), ((30, 18, 14), None, False), ((64, 64, 64), (42, 42, 42), True)])
def test_fused_scale(lib, n, cut, stride1):
    """p3dfft_b200_set_scale through the executor: the scale lands on the stage that writes the user array, and only there"""
    nx, ny, nz = n
    c = cut or (None, None, None)
    lib.set_layout(stride1, False)
    lib.p3dfft_setup((1, 1), nx, ny, nz, 0, *c)
    d = po.Decomp(nx, ny, nz, (1, 1), 0, *c, stride1=stride1)
    A = np.asfortranarray(np.random.default_rng(8).random(n))
    N = float(nx * ny * nz)
    lib.set_scale(1.0 / N, 2.0)
    exp = po.local_forward(A, d, "fft")
    F = np.zeros(exp.shape, dtype=np.complex128, order="F")
    lib.p3dfft_ftran_r2c(A, F, "fft")
    assert po.rel_l2(F, exp / N) <= 1e-13
    B = np.zeros(n, order="F")
    lib.p3dfft_btran_c2r(np.asfortranarray(exp), B, "tff")
    assert po.rel_l2(B, 2.0 * po.local_backward(po.global_forward(A, d, "fft"), d, "tff")) <= 1e-13


@pytest.mark.parametrize("n,cut,stride1", [((16, 16, 16), None, False), ((16, 12, 10), None, True), ((32, 32, 32), (20, 20, 20), False)])
def test_power_spectrum(lib, n, cut, stride1):
    """the spectrum kernel (warp shuffles, ballots, shared and global atomics emulated) against the oracle"""
    nx, ny, nz = n
    c = cut or (None, None, None)
    lib.set_layout(stride1, False)
    lib.p3dfft_setup((1, 1), nx, ny, nz, 0, *c)
    d = po.Decomp(nx, ny, nz, (1, 1), 0, *c, stride1=stride1)
    A = np.asfortranarray(np.random.default_rng(4).random(n))
    F = np.asfortranarray(po.local_forward(A, d, "fft"))
    kmax = po.spectrum_kmax(nx, ny, nz)
    factor = 1.0 / (nx * ny * nz)
    E = lib.spectrum(F, kmax, factor)
    expE = po.power_spectrum(F, d, kmax, factor)
    assert np.max(np.abs(E - expE)) <= 1e-12 * np.max(np.abs(expE))


def test_r2c_1d_rtran_and_queries(lib):
    nx, ny, nz = 64, 12, 10
    lib.p3dfft_setup((1, 1), nx, ny, nz, 0)
    d = po.Decomp(nx, ny, nz, (1, 1), 0)
    A = np.asfortranarray(np.random.default_rng(1).random((nx, ny, nz)))
    exp = po.forward_r2c_1d(A)
    Cc = np.zeros(exp.size, dtype=np.complex128)
    lib.p3dfft_ftran_r2c_1d(A, Cc)
    assert po.rel_l2(Cc, exp.ravel(order="F")) <= 1e-13
    for which in pb.RTRAN_NAMES:
        dst = np.full(A.size, np.nan)
        dstart, dend, dsize, t = lib.rtran(which, A, dst)
        assert [list(dstart), list(dend), list(dsize)] == [list(x) for x in po.rtran_dims(d, which)]
        assert np.array_equal(dst, A.ravel(order="F")) and t == 0.0
    g = po.ProcGrid(nx, ny, nz, (1, 1))
    assert lib.p3dfft_get_mpi_info()[:2] == (0, 1)
    assert lib.proc_dims(2, 0) == [g.proc_dims[(2, k, 0)] for k in range(1, 10)]
    exp_parts = g.get_proc_parts(2, 3, 4, 5, 6, 3, 1)
    assert lib.get_proc_parts((2, 3, 4), (5, 6, 3), 1, 1) == (exp_parts[0][:1], exp_parts[1], exp_parts[2])


def test_opt_in_variants_through_the_executor(lib, monkeypatch):
    """the switchable kernel variants end to end through api.cpp (sizes whose six stages all run on the specialised
    kernels: the emulation of the any-length kernel's many small CTAs is slow)"""
    cases = [({"P3DFFT_B200_R32": "1"}, (64, 512, 64), False), ({"P3DFFT_B200_R32": "0"}, (64, 1024, 64), False),
             ({"P3DFFT_B200_BULK": "1"}, (64, 512, 64), True), ({"P3DFFT_B200_BULK": "1", "P3DFFT_B200_R32": "1"}, (64, 512, 64), True)]
    for env, n, both in cases:
        d = po.Decomp(*n, (1, 1), 0)
        A = np.asfortranarray(np.random.default_rng(5).random(n))
        exp = po.local_forward(A, d, "fft")
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        try:
            lib.p3dfft_setup((1, 1), *n, 0)
            F = np.zeros(exp.shape, dtype=np.complex128, order="F")
            lib.launch_count(True)
            lib.fast_launch_count(True)
            lib.p3dfft_ftran_r2c(A, F, "fft")
            nl = lib.launch_count()
            assert lib.fast_launch_count() == nl
            assert po.rel_l2(F, exp) <= 1e-13, env
            if both:
                B = np.zeros(n, order="F")
                lib.p3dfft_btran_c2r(F, B, "tff")
                assert np.max(np.abs(B / A.size - A)) <= 1e-13, env
                assert lib.launch_count() == 2 * nl
            assert nl == 3
        finally:
            lib.p3dfft_clean()
            for k in env:
                monkeypatch.delenv(k)


def test_guard_pages_catch_an_overrun():
    """the mock cudaMalloc end-aligns every allocation against an inaccessible page (tests/emu/emu_mp.inc): writing 16 bytes
    past the end of a "device" buffer kills the process, the last 16 bytes inside it are fine"""
    import subprocess
    import sys
    code = ("import ctypes as C, sys\n"
            f"L = C.CDLL({os.path.join(ROOT, 'tests', 'emu', 'lib', 'libp3dfft_emu.so')!r})\n"
            "p = C.c_void_p()\n"
            "assert L.cudaMalloc(C.byref(p), C.c_size_t(1000 * 16)) == 0\n"
            "C.memset(p.value + 999 * 16, 1, 16)\n"
            "print('inside ok', flush=True)\n"
            "if sys.argv[1] == 'over':\n"
            "    C.memset(p.value + 1000 * 16, 1, 16)\n"
            "print('done', flush=True)\n")
    ok = subprocess.run([sys.executable, "-c", code, "in"], capture_output=True, text=True)
    assert ok.returncode == 0 and "done" in ok.stdout, ok.stderr
    bad = subprocess.run([sys.executable, "-c", code, "over"], capture_output=True, text=True)
    assert bad.returncode < 0 and "inside ok" in bad.stdout and "done" not in bad.stdout, (bad.returncode, bad.stdout, bad.stderr)


def test_async_calls_on_a_user_stream(lib):
    """p3dfft_b200_set_stream + set_async: with device arrays the transform is only ENQUEUED on the caller's stream (the mock
    runtime, lazy policy, runs nothing before a synchronisation), p3dfft_b200_sync completes it; timers are not touched"""
    n = (64, 64, 64)
    st = C.c_void_p()
    assert lib.lib.cudaStreamCreateWithFlags(C.byref(st), 1) == 0         # a non-blocking stream of the caller
    lib.set_stream(st.value)
    lib.p3dfft_setup((1, 1), *n, 0)
    d = po.Decomp(*n, (1, 1), 0)
    A = np.asfortranarray(np.random.default_rng(9).random(n))
    F = np.zeros((d.nxhp, n[1], n[2]), dtype=np.complex128, order="F")
    B = np.zeros(n, order="F")
    try:
        with device_arrays(lib, A, F, B):
            lib.set_async(True)
            lib.set_timers()
            lib.p3dfft_ftran_r2c(A, F, "fft")
            lib.p3dfft_btran_c2r(F, B, "tff")
            if os.environ.get("P3D_EMU_STREAMS", "lazy") == "lazy":
                assert not F.any() and not B.any(), "asynchronous calls must return before the work has run"
            lib.sync()
            assert po.rel_l2(F, po.local_forward(A, d, "fft")) <= 1e-13
            assert np.max(np.abs(B / A.size - A)) <= 1e-13
            assert not any(lib.get_timers()), "asynchronous calls do not time their stages"
            # host arrays fall back to a synchronous call even in asynchronous mode (the copy back must have landed)
            F2 = np.zeros_like(F)
            lib.p3dfft_ftran_r2c(A.copy(order="F"), F2, "fft")
            assert po.rel_l2(F2, F) <= 1e-15
    finally:
        lib.set_async(False)
        lib.p3dfft_clean()
        lib.reset_stream()
        lib.lib.cudaStreamDestroy(st)


def test_clean_releases_every_device_allocation(lib):
    """p3dfft_clean (module.F90:309) frees work buffers, staging buffers, twiddle tables, the spectrum accumulator ...: the mock
    allocator counts live blocks, none may survive; a second setup / clean cycle must not accumulate any either"""
    nb = C.c_longlong()
    lib.lib.emu_live_allocations.restype = C.c_longlong
    base = lib.lib.emu_live_allocations(C.byref(nb))
    for cycle in range(2):
        n = (64, 16, 17)
        lib.p3dfft_setup((1, 1), *n, 0)
        d = po.Decomp(*n, (1, 1), 0)
        A = np.asfortranarray(np.random.default_rng(1).random(n))
        F = np.zeros((d.nxhp, n[1], n[2]), dtype=np.complex128, order="F")
        lib.p3dfft_ftran_r2c(A, F, "fft")
        lib.p3dfft_btran_c2r(F, A.copy(order="F"), "tff")
        lib.p3dfft_ftran_r2c(A, F, "ffc")
        A2 = np.asfortranarray(np.random.default_rng(2).random((2,) + n)).ravel()
        F2 = np.zeros(2 * F.size, dtype=np.complex128)
        lib.p3dfft_ftran_r2c_many(A2, A.size, F2, F.size, 2, "fft")          # the work buffers grow
        lib.spectrum(F, 12)
        dst = np.zeros(A.size)
        lib.rtran("x2y", A.ravel(order="F").copy(), dst)
        assert lib.lib.emu_live_allocations(C.byref(nb)) > base
        lib.p3dfft_clean()
        assert lib.lib.emu_live_allocations(C.byref(nb)) == base, (cycle, nb.value)


@pytest.mark.parametrize("threads", ["1", "3", "6"])
def test_pageable_host_arrays_take_the_chunk_ring(monkeypatch, threads):
    """host arrays that are not page-locked go through a ring of page-locked chunks filled / drained by copy threads
    (api.cpp copy_h2d / copy_d2h): forward, backward and in-place calls on arrays of many chunks, sizes that are no multiple
    of the chunk -- in a fresh process, since the ring is sized once"""
    import subprocess
    import sys
    code = ("import os, sys, numpy as np\n"
            f"sys.path.insert(0, {ROOT!r})\n"
            "import p3dfft_b200 as pb\n"
            "from oracle import p3dfft_oracle as po\n"
            f"lib = pb.P3DFFT(False, path={os.path.join(ROOT, 'tests', 'emu', 'lib', 'libp3dfft_emu.so')!r})\n"
            "n = (64, 48, 40)\n"
            "lib.p3dfft_setup((1, 1), *n, 0)\n"
            "d = po.Decomp(*n, (1, 1), 0)\n"
            "A = np.asfortranarray(np.random.default_rng(3).random(n))\n"
            "F = np.zeros((d.nxhp, n[1], n[2]), dtype=np.complex128, order='F')\n"
            "lib.p3dfft_ftran_r2c(A, F, 'fft')\n"
            "e1 = po.rel_l2(F, po.local_forward(A, d, 'fft'))\n"
            "B = np.zeros(n, order='F')\n"
            "lib.p3dfft_btran_c2r(F, B, 'tff')\n"
            "e2 = float(np.max(np.abs(B / A.size - A)))\n"
            "mem = lib.p3dfft_setup and None\n"
            "W = np.zeros(2 * d.nxhp * n[1] * n[2]); W[:A.size] = A.ravel(order='F')\n"
            "lib.p3dfft_ftran_r2c(W, W, 'fft')\n"
            "e3 = po.rel_l2(W.view(np.complex128), F.ravel(order='F'))\n"
            "lib.p3dfft_clean()\n"
            "Wkeep = W.copy()\n"
            "lib.p3dfft_setup((1, 1), *n, 0)\n"      # a second life of the ring: its copy threads start afresh and must
            "F2 = np.zeros_like(F)\n"                 # not replay the last job of the first life (its target was W)
            "lib.p3dfft_ftran_r2c(A, F2, 'fft')\n"
            "e4 = po.rel_l2(F2, F) + float(np.max(np.abs(W - Wkeep)))\n"
            "lib.p3dfft_clean()\n"
            "print('errors', e1, e2, e3, e4)\n"
            "sys.exit(0 if max(e1, e3, e4) < 1e-13 and e2 < 1e-13 else 1)\n")
    env = dict(os.environ, P3DFFT_B200_COPY_CHUNK_KB="100", P3DFFT_B200_COPY_THREADS=threads)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
