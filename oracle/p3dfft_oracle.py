"""CPU oracle for the P3DFFT r2c/c2r hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy restatement of the reference algorithm (sdsc/p3dfft 2.7.x,
Fortran + MPI + FFTW).  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing
under ``p3dfft_b200/`` (the product) imports or calls it.

Why numpy and not the reference itself: the reference cannot be built in this image
(no Fortran compiler, no MPI, no FFTW -- SURVEY.md section 0), and its 1D transforms live
in an un-vendored, un-pinned third-party dependency (FFTW 3.x, ``build/fft_spec.F90:88``,
``configure.ac:330-353``).  The published FFTW definitions restated here are
  r2c / c2c forward : Y_k = sum_j X_j exp(-2 pi i jk/N)   (unnormalised)
  c2r / c2c backward: Y_k = sum_j X_j exp(+2 pi i jk/N)   (unnormalised)
  REDFT00 (DCT-I)   : Y_k = X_0 + (-1)^k X_{N-1} + 2 sum_{j=1}^{N-2} X_j cos(pi jk/(N-1))
  RODFT00 (DST-I)   : Y_k = 2 sum_{j=0}^{N-1} X_j sin(pi (j+1)(k+1)/(N+1))
computed with scipy's pocketfft.  PARITY PINNING: the reference ships no golden vectors;
the oracle is pinned on the known-answer checks of the reference's own sample drivers
(``sample/C/driver_inverse.c:222-240``, ``driver_sine.c:168-228``, ``driver_noop.c``,
``sample/FORTRAN/driver_cheby.F90:258-285``, ``driver_sine_pruned.F90``) -- see
``tests/test_oracle_known_answers.py``.  Bit-level parity with "the FFTW build" is
therefore unpinned; the parity bar is relative L2 <= 1e-12 (double) / 1e-5 (single).

Two restatements are provided:
  * ``global_*``      -- the mathematical definition on the global array, sliced per rank.
  * ``SimWorld``      -- P simulated ranks running the reference's actual stage sequence
                         (FFT -> pack -> alltoallv -> unpack ...) with the reference's
                         buffer layouts and byte counts, so that the exchange tables are
                         exercised.  This is also what the CPU baseline times.

All arrays are Fortran-ordered (column-major, x fastest) like the reference.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np
import scipy.fft as sfft

__all__ = [
    "map_data_to_proc", "Decomp", "global_forward", "global_backward", "global_cheby",
    "local_forward", "local_backward", "SimWorld", "philox_field", "rel_l2",
    "rtran_slices", "rtran_local", "rtran_dims", "forward_r2c_1d", "ProcGrid", "power_spectrum", "spectrum_kmax",
]


# --------------------------------------------------------------------------------------
# decomposition arithmetic
# --------------------------------------------------------------------------------------
def map_data_to_proc(data: int, proc: int):
    """``MapDataToProc`` (build/setup.F90:608-635): the LAST ``data mod proc`` ranks get
    one extra element.  Returns 1-based (st, en, sz) lists."""
    size = data // proc
    nu = data - size * proc
    nl = proc - nu
    st, en, sz = [0] * proc, [0] * proc, [0] * proc
    st[0], sz[0], en[0] = 1, size, size
    for i in range(1, nl):
        st[i] = st[i - 1] + size
        sz[i] = size
        en[i] = en[i - 1] + size
    size1 = size + 1
    for i in range(max(nl, 1), proc):
        st[i] = en[i - 1] + 1
        sz[i] = size1
        en[i] = en[i - 1] + size1
    en[proc - 1] = data
    sz[proc - 1] = data - st[proc - 1] + 1
    return st, en, sz


@dataclass
class Decomp:
    """Per-rank plan integers.  Follows build/setup.F90:148-176 (derived extents),
    :192-219 (rank -> (ipid,jpid)), :279-312 (block maps), :382-398 (padi, nm),
    :481-518 (alltoallv tables), :580-603 (memsize); build/module.F90:225-273 (get_dims)."""
    nx: int
    ny: int
    nz: int
    dims: tuple
    rank: int = 0
    nxc: int | None = None
    nyc: int | None = None
    nzc: int | None = None
    dims_c: bool = False          # -DDIMS_C
    stride1: bool = False         # -DSTRIDE1 (changes conf-2 layout only)
    elem: int = 8                 # bytes per real (8 double, 4 single)

    def __post_init__(self):
        nx, ny, nz = self.nx, self.ny, self.nz
        if nx <= 0 or ny <= 0 or nz <= 0:
            raise ValueError(f"Invalid dimensions : {nx} {ny} {nz}")
        self.nxc = nx if self.nxc is None else self.nxc
        self.nyc = ny if self.nyc is None else self.nyc
        self.nzc = nz if self.nzc is None else self.nzc
        self.iproc, self.jproc = int(self.dims[0]), int(self.dims[1])
        if self.iproc <= 0 or self.jproc <= 0:
            raise ValueError("Invalid processor geometry")
        self.numtasks = self.iproc * self.jproc
        if not (0 <= self.rank < self.numtasks):
            raise ValueError("rank outside processor grid")
        self.nxh = nx // 2
        self.nxhp = self.nxh + 1
        self.nxhc = self.nxc // 2
        self.nxhpc = self.nxhc + 1
        self.nyh, self.nzh = ny // 2, nz // 2
        self.nyhc, self.nzhc = self.nyc // 2, self.nzc // 2
        self.nycph, self.nzcph = (self.nyc + 1) // 2, (self.nzc + 1) // 2
        # rank -> grid coordinates (setup.F90:195-219; row-major MPI_Cart_create)
        if self.dims_c:
            self.ipid, self.jpid = self.rank // self.jproc, self.rank % self.jproc
        else:
            self.ipid, self.jpid = self.rank % self.iproc, self.rank // self.iproc
        self.iist, self.iien, self.iisz = map_data_to_proc(self.nxhpc, self.iproc)
        self.jist, self.jien, self.jisz = map_data_to_proc(ny, self.iproc)
        self.jjst, self.jjen, self.jjsz = map_data_to_proc(self.nyc, self.jproc)
        self.kjst, self.kjen, self.kjsz = map_data_to_proc(nz, self.jproc)
        i, j = self.ipid, self.jpid
        self.iistart, self.iiend, self.iisize = self.iist[i], self.iien[i], self.iisz[i]
        self.jistart, self.jiend, self.jisize = self.jist[i], self.jien[i], self.jisz[i]
        self.jjstart, self.jjend, self.jjsize = self.jjst[j], self.jjen[j], self.jjsz[j]
        self.kjstart, self.kjend, self.kjsize = self.kjst[j], self.kjen[j], self.kjsz[j]
        # real-space x blocks of the rtran_* transposes (setup.F90:299-312)
        self.iiist, self.iiien, self.iiisz = map_data_to_proc(nx, self.iproc)
        self.ijst, self.ijen, self.ijsz = map_data_to_proc(nx, self.jproc)
        self.iiistart, self.iiiend, self.iiisize = self.iiist[i], self.iiien[i], self.iiisz[i]
        self.ijstart, self.ijend, self.ijsize = self.ijst[j], self.ijen[j], self.ijsz[j]
        # work-buffer padding (setup.F90:382-398)
        padd = max(self.iisize * self.jjsize * nz, self.iisize * ny * self.kjsize) \
            - self.nxhp * self.jisize * self.kjsize
        if padd <= 0:
            padi = 0
        else:
            d = self.nxhp * self.jisize
            padi = padd // d if padd % d == 0 else padd // d + 1
        self.padi_work = padi
        self.nm = self.nxhp * self.jisize * (self.kjsize + padi)
        # alltoallv tables in BYTES (setup.F90:481-518); c = bytes per complex
        c = 2 * self.elem
        ii, ji, jj, kj = self.iisize, self.jisize, self.jjsize, self.kjsize
        self.IfSndStrt = [(self.iist[p] - 1) * ji * kj * c for p in range(self.iproc)]
        self.IfSndCnts = [self.iisz[p] * ji * kj * c for p in range(self.iproc)]
        self.IfRcvStrt = [(self.jist[p] - 1) * ii * kj * c for p in range(self.iproc)]
        self.IfRcvCnts = [self.jisz[p] * ii * kj * c for p in range(self.iproc)]
        self.KfSndStrt = [(self.jjst[p] - 1) * ii * kj * c for p in range(self.jproc)]
        self.KfSndCnts = [ii * kj * self.jjsz[p] * c for p in range(self.jproc)]
        self.KfRcvStrt = [(self.kjst[p] - 1) * ii * jj * c for p in range(self.jproc)]
        self.KfRcvCnts = [ii * jj * self.kjsz[p] * c for p in range(self.jproc)]
        self.JrSndStrt, self.JrSndCnts = self.KfRcvStrt, self.KfRcvCnts
        self.JrRcvStrt, self.JrRcvCnts = self.KfSndStrt, self.KfSndCnts
        self.KrSndStrt, self.KrSndCnts = self.IfRcvStrt, self.IfRcvCnts
        self.KrRcvStrt, self.KrRcvCnts = self.IfSndStrt, self.IfSndCnts
        # rtran tables in BYTES (setup.F90:531-549); r = bytes per real
        r = self.elem
        iii, ijs = self.iiisize, self.ijsize
        self.IiStrt = [(self.iiist[p] - 1) * ji * kj * r for p in range(self.iproc)]
        self.IiCnts = [self.iiisz[p] * ji * kj * r for p in range(self.iproc)]
        self.JiStrt = [(self.jist[p] - 1) * iii * kj * r for p in range(self.iproc)]
        self.JiCnts = [self.jisz[p] * iii * kj * r for p in range(self.iproc)]
        self.IjStrt = [(self.ijst[p] - 1) * ji * kj * r for p in range(self.jproc)]
        self.IjCnts = [self.ijsz[p] * ji * kj * r for p in range(self.jproc)]
        self.KjStrt = [(self.kjst[p] - 1) * ji * ijs * r for p in range(self.jproc)]
        self.KjCnts = [self.kjsz[p] * ji * ijs * r for p in range(self.jproc)]
        # memsize (setup.F90:580-603): real elements needed for an in-place array
        pad1 = 2 * max(nz * jj * ii, ny * kj * ii) - nx * ji * kj
        if pad1 <= 0:
            pad1 = 0
        d = nx * ji
        padm = pad1 // d + (1 if pad1 % d else 0) if d > 0 else 0
        self.padi = padm                      # public module variable after setup
        self.memsize = (nx, ji, kj + padm)

    # ---- communicators -------------------------------------------------------------
    def rank_of(self, ipid: int, jpid: int) -> int:
        return ipid * self.jproc + jpid if self.dims_c else jpid * self.iproc + ipid

    def row_ranks(self):
        """mpi_comm_row: same jpid, ordered by ipid (setup.F90:245-261)."""
        return [self.rank_of(i, self.jpid) for i in range(self.iproc)]

    def col_ranks(self):
        """mpi_comm_col: same ipid, ordered by jpid."""
        return [self.rank_of(self.ipid, j) for j in range(self.jproc)]

    # ---- p3dfft_get_dims (module.F90:225-273) --------------------------------------
    def get_dims(self, conf: int):
        if conf == 1:
            return ([1, self.jistart, self.kjstart], [self.nx, self.jiend, self.kjend],
                    [self.nx, self.jisize, self.kjsize])
        if conf == 2:
            if self.stride1:
                return ([1, self.jjstart, self.iistart], [self.nzc, self.jjend, self.iiend],
                        [self.nzc, self.jjsize, self.iisize])
            return ([self.iistart, self.jjstart, 1], [self.iiend, self.jjend, self.nzc],
                    [self.iisize, self.jjsize, self.nzc])
        if conf == 3:
            m = list(self.memsize)
            return ([0, 0, 0], m, list(m))
        raise ValueError("conf must be 1, 2 or 3")

    # ---- pruning index maps (0-based kept -> full index) -----------------------------
    def kept_y(self):
        """Kept Y modes: first nycph, last nyc-nycph (seg_copy_y calls, ftran.F90:756-757).
        NOTE: for odd nyc and jproc>1 the reference's pack_fcomm2 splits at nyhc instead
        (fcomm2.F90:339-381) and loses an element; even nyc (all reference tests) agree."""
        h1 = self.nycph
        return np.concatenate([np.arange(h1), np.arange(h1, self.nyc) + (self.ny - self.nyc)])

    def kept_z(self):
        """Kept Z modes: first nzcph, last nzc-nzcph (seg_copy_z, ftran.F90:645-646)."""
        h1 = self.nzcph
        return np.concatenate([np.arange(h1), np.arange(h1, self.nzc) + (self.nz - self.nzc)])


# --------------------------------------------------------------------------------------
# 1D transforms (FFTW definitions, unnormalised)
# --------------------------------------------------------------------------------------
_WORKERS = int(os.environ.get("P3DFFT_ORACLE_WORKERS", "0")) or None


def _fft(a, axis, inverse=False):
    if inverse:
        return sfft.ifft(a, axis=axis, norm="forward", workers=_WORKERS)
    return sfft.fft(a, axis=axis, workers=_WORKERS)


def _ztrans(a, axis, ch, inverse):
    """Third-dimension transform selected by the op letter (ftran.F90:613-643,
    btran.F90:440-466): t/f = c2c FFT, c = DCT-I, s = DST-I on re and im separately
    (fft_exec.F90:670-671, 890-891), n/0 = nothing."""
    if ch in ("t", "f"):
        return _fft(a, axis, inverse)
    if ch == "c":
        return sfft.dct(a.real, type=1, axis=axis, workers=_WORKERS) + \
            1j * sfft.dct(a.imag, type=1, axis=axis, workers=_WORKERS)
    if ch == "s":
        return sfft.dst(a.real, type=1, axis=axis, workers=_WORKERS) + \
            1j * sfft.dst(a.imag, type=1, axis=axis, workers=_WORKERS)
    if ch in ("n", "0"):
        return a
    raise ValueError(f"Unknown transform type: {ch}")


def _ctype(real_dtype):
    return np.complex64 if np.dtype(real_dtype) == np.float32 else np.complex128


# --------------------------------------------------------------------------------------
# global definition
# --------------------------------------------------------------------------------------
def global_forward(A, d: Decomp, op="fft"):
    """Whole-array forward transform: real A[nx,ny,nz] -> complex [nxhpc, nyc, nzc]
    (x-pruned to the first nxhpc, y/z pruned to the outer modes).  ftran.F90:489-780."""
    A = np.asfortranarray(A)
    F = sfft.rfft(A, axis=0, workers=_WORKERS)[: d.nxhpc]
    F = _fft(F, 1)[:, d.kept_y(), :]
    F = _ztrans(F, 2, op[2], False)[:, :, d.kept_z()]
    return np.asfortranarray(F.astype(_ctype(A.dtype), copy=False))


def global_backward(F, d: Decomp, op="tff"):
    """Whole-array backward transform: complex [nxhpc,nyc,nzc] -> real [nx,ny,nz];
    pruned modes are zero-filled first (btran.F90:471-509, bcomm1.F90:367-374,
    bcomm2.F90:307-313).  Unnormalised."""
    ct = F.dtype
    Z = np.zeros((d.nxhpc, d.nyc, d.nz), dtype=ct, order="F")
    Z[:, :, d.kept_z()] = F
    Z = _ztrans(Z, 2, op[0], True)
    Y = np.zeros((d.nxhp, d.ny, d.nz), dtype=ct, order="F")
    Y[: d.nxhpc, d.kept_y(), :] = Z
    Y = _fft(Y, 1, inverse=True)
    R = sfft.irfft(Y, n=d.nx, axis=0, norm="forward", workers=_WORKERS)
    rt = np.float32 if ct == np.complex64 else np.float64
    return np.asfortranarray(R.astype(rt, copy=False))


def cheby_epilogue(out, d: Decomp, Lz):
    """Scaling + Chebyshev derivative recurrence along z (ftran.F90:408-451); ``out`` is
    the 'ffc' transform with z as LAST axis; returns a new array."""
    nzc = d.nzc
    a = out * (1.0 / (float(d.nx * d.ny) * float(nzc - 1)))
    res = np.array(a, copy=True)
    L = 4.0 / float(Lz)
    # 1-based k in the reference; here z index k-1
    res[..., nzc - 1] = 0
    res[..., nzc - 2] = L * (nzc - 1) * a[..., nzc - 1] * 0.5
    for k in range(nzc - 2, 0, -1):          # k = nzc-2 .. 1 (1-based)
        res[..., k - 1] = L * k * a[..., k] + res[..., k + 1]
    res[..., 0] = res[..., 0] * 0.5
    return res


def global_cheby(A, d: Decomp, Lz):
    return np.asfortranarray(cheby_epilogue(global_forward(A, d, "ffc"), d, Lz))


def local_in_slice(d: Decomp):
    return (slice(None), slice(d.jistart - 1, d.jiend), slice(d.kjstart - 1, d.kjend))


def local_out_slice(d: Decomp):
    return (slice(d.iistart - 1, d.iiend), slice(d.jjstart - 1, d.jjend), slice(None))


def local_forward(Aglobal, d: Decomp, op="fft"):
    """Rank-local wavenumber block in the layout get_dims(conf=2) describes."""
    F = global_forward(Aglobal, d, op)[local_out_slice(d)]
    return np.asfortranarray(F.transpose(2, 1, 0)) if d.stride1 else np.asfortranarray(F)


def local_backward(Fglobal, d: Decomp, op="tff"):
    return np.asfortranarray(global_backward(Fglobal, d, op)[local_in_slice(d)])


# --------------------------------------------------------------------------------------
# structural restatement: P simulated ranks, the reference's stage sequence
# --------------------------------------------------------------------------------------
class SimWorld:
    """All ranks of an iproc x jproc grid simulated in one process.  Each stage follows
    the non-STRIDE1 single-variable code path: ftran.F90:489-780 / btran.F90:396-679 with
    fcomm1.F90:209-327, fcomm2.F90:250-386, bcomm1.F90:250-380, bcomm2.F90:213-318.
    The alltoallv is a python loop moving the byte ranges given by the If/Kf/Jr/Kr tables."""

    def __init__(self, nx, ny, nz, dims, nxc=None, nyc=None, nzc=None, dtype=np.float64,
                 dims_c=False):
        self.rt = np.dtype(dtype)
        self.ct = np.dtype(_ctype(dtype))
        elem = self.rt.itemsize
        P = dims[0] * dims[1]
        self.d = [Decomp(nx, ny, nz, tuple(dims), r, nxc, nyc, nzc, dims_c=dims_c, elem=elem)
                  for r in range(P)]
        self.P = P

    # ---- helpers -----------------------------------------------------------------
    def scatter_real(self, Aglobal):
        return [np.asfortranarray(Aglobal[local_in_slice(d)].astype(self.rt)) for d in self.d]

    def gather_real(self, parts):
        d0 = self.d[0]
        G = np.zeros((d0.nx, d0.ny, d0.nz), dtype=self.rt, order="F")
        for d, p in zip(self.d, parts):
            G[local_in_slice(d)] = p
        return G

    def scatter_wave(self, Fglobal):
        return [np.asfortranarray(Fglobal[local_out_slice(d)].astype(self.ct)) for d in self.d]

    def gather_wave(self, parts):
        d0 = self.d[0]
        G = np.zeros((d0.nxhpc, d0.nyc, d0.nzc), dtype=self.ct, order="F")
        for d, p in zip(self.d, parts):
            G[local_out_slice(d)] = p
        return G

    def _alltoallv(self, sendbufs, group_of, snd_strt, snd_cnts, rcv_strt, rcv_cnts, recv_len):
        """MPI_Alltoallv with counts in bytes over flat complex buffers."""
        c = self.ct.itemsize
        recv = [np.zeros(recv_len(d), dtype=self.ct) for d in self.d]
        for d in self.d:
            grp = group_of(d)
            me = grp.index(d.rank)
            for p, peer in enumerate(grp):
                dp = self.d[peer]
                s0, n = snd_strt(d)[p] // c, snd_cnts(d)[p] // c
                r0, m = rcv_strt(dp)[me] // c, rcv_cnts(dp)[me] // c
                assert n == m, "alltoallv count mismatch"
                recv[peer][r0:r0 + n] = sendbufs[d.rank][s0:s0 + n]
        return recv

    # ---- forward -----------------------------------------------------------------
    def forward(self, parts, op="fft"):
        D = self.d
        # 1. X r2c: (nx,ji,kj) -> (nxhp,ji,kj)              exec_f_r2c, ftran.F90:530
        b2 = [np.asfortranarray(sfft.rfft(a, axis=0, workers=_WORKERS).astype(self.ct)) for a in parts]
        # 2. row transpose                                   fcomm1 / seg_copy_x
        ybuf = []
        if D[0].iproc > 1:
            send = []
            for d, s in zip(D, b2):
                buf1 = np.zeros(d.nxhpc * d.jisize * d.kjsize, dtype=self.ct)
                for p in range(d.iproc):                     # fcomm1.F90:239-253
                    pos = d.IfSndStrt[p] // self.ct.itemsize
                    blk = s[d.iist[p] - 1:d.iien[p], :, :]
                    buf1[pos:pos + blk.size] = blk.ravel(order="F")
                send.append(buf1)
            recv = self._alltoallv(send, lambda d: d.row_ranks(), lambda d: d.IfSndStrt,
                                   lambda d: d.IfSndCnts, lambda d: d.IfRcvStrt,
                                   lambda d: d.IfRcvCnts, lambda d: d.iisize * d.ny * d.kjsize)
            for d, r in zip(D, recv):                        # fcomm1.F90:284-320
                dest = np.zeros((d.iisize, d.ny, d.kjsize), dtype=self.ct, order="F")
                for p in range(d.iproc):
                    pos = d.IfRcvStrt[p] // self.ct.itemsize
                    n = d.iisize * d.jisz[p] * d.kjsize
                    dest[:, d.jist[p] - 1:d.jien[p], :] = \
                        r[pos:pos + n].reshape((d.iisize, d.jisz[p], d.kjsize), order="F")
                ybuf.append(dest)
        else:
            ybuf = [np.asfortranarray(s[: d.nxhpc]) for d, s in zip(D, b2)]   # ftran.F90:554
        # 3. Y c2c per z-plane                               ftran.F90:581-583
        ybuf = [_fft(b, 1) for b in ybuf]
        # 4. column transpose with Y pruning                 fcomm2 / seg_copy_y
        zbuf = []
        if D[0].jproc > 1:
            send = []
            for d, s in zip(D, ybuf):
                sp = s[:, d.kept_y(), :]                      # pack_fcomm2, fcomm2.F90:321-386
                buf1 = np.zeros(d.iisize * d.nyc * d.kjsize, dtype=self.ct)
                for p in range(d.jproc):
                    pos = d.KfSndStrt[p] // self.ct.itemsize
                    blk = sp[:, d.jjst[p] - 1:d.jjen[p], :]
                    buf1[pos:pos + blk.size] = blk.ravel(order="F")
                send.append(buf1)
            recv = self._alltoallv(send, lambda d: d.col_ranks(), lambda d: d.KfSndStrt,
                                   lambda d: d.KfSndCnts, lambda d: d.KfRcvStrt,
                                   lambda d: d.KfRcvCnts, lambda d: d.iisize * d.jjsize * d.nz)
            # lands directly as (iisize,jjsize,nz)            fcomm2.F90:313
            zbuf = [r.reshape((d.iisize, d.jjsize, d.nz), order="F") for d, r in zip(D, recv)]
        else:
            zbuf = [np.asfortranarray(s[:, d.kept_y(), :]) for d, s in zip(D, ybuf)]
        # 5. Z transform + Z pruning                         ftran.F90:605-683
        out = []
        for d, b in zip(D, zbuf):
            t = _ztrans(b, 2, op[2], False)
            out.append(np.asfortranarray(t[:, :, d.kept_z()].astype(self.ct)))
        return out

    # ---- backward ----------------------------------------------------------------
    def backward(self, parts, op="tff"):
        D = self.d
        # 1. zero-pad in z + Z inverse                        btran.F90:437-509
        zb = []
        for d, f in zip(D, parts):
            b = np.zeros((d.iisize, d.jjsize, d.nz), dtype=self.ct, order="F")
            b[:, :, d.kept_z()] = f
            zb.append(np.asfortranarray(_ztrans(b, 2, op[0], True).astype(self.ct)))
        # 2. column transpose + Y zero fill                   bcomm1 / seg_copy_y+seg_zero_y
        yb = []
        if D[0].jproc > 1:
            send = [b.ravel(order="F") for b in zb]           # sent from the array itself, bcomm1.F90:295
            recv = self._alltoallv(send, lambda d: d.col_ranks(), lambda d: d.JrSndStrt,
                                   lambda d: d.JrSndCnts, lambda d: d.JrRcvStrt,
                                   lambda d: d.JrRcvCnts, lambda d: d.iisize * d.nyc * d.kjsize)
            for d, r in zip(D, recv):                         # unpack_bcomm1, bcomm1.F90:309-380
                dest = np.zeros((d.iisize, d.ny, d.kjsize), dtype=self.ct, order="F")
                ky = d.kept_y()
                for p in range(d.jproc):
                    pos = d.JrRcvStrt[p] // self.ct.itemsize
                    n = d.iisize * d.jjsz[p] * d.kjsize
                    dest[:, ky[d.jjst[p] - 1:d.jjen[p]], :] = \
                        r[pos:pos + n].reshape((d.iisize, d.jjsz[p], d.kjsize), order="F")
                yb.append(dest)
        else:
            for d, b in zip(D, zb):
                dest = np.zeros((d.iisize, d.ny, d.nz), dtype=self.ct, order="F")
                dest[:, d.kept_y(), :] = b
                yb.append(dest)
        # 3. Y inverse                                       btran.F90:618-623
        yb = [_fft(b, 1, inverse=True) for b in yb]
        # 4. row transpose + X zero fill                      bcomm2 / seg_copy_x+seg_zero_x
        xb = []
        if D[0].iproc > 1:
            send = []
            for d, s in zip(D, yb):                           # bcomm2.F90:233-273
                buf1 = np.zeros(d.iisize * d.ny * d.kjsize, dtype=self.ct)
                for p in range(d.iproc):
                    pos = d.KrSndStrt[p] // self.ct.itemsize
                    blk = s[:, d.jist[p] - 1:d.jien[p], :]
                    buf1[pos:pos + blk.size] = blk.ravel(order="F")
                send.append(buf1)
            recv = self._alltoallv(send, lambda d: d.row_ranks(), lambda d: d.KrSndStrt,
                                   lambda d: d.KrSndCnts, lambda d: d.KrRcvStrt,
                                   lambda d: d.KrRcvCnts, lambda d: d.nxhpc * d.jisize * d.kjsize)
            for d, r in zip(D, recv):                         # bcomm2.F90:290-313
                dest = np.zeros((d.nxhp, d.jisize, d.kjsize), dtype=self.ct, order="F")
                for p in range(d.iproc):
                    pos = d.KrRcvStrt[p] // self.ct.itemsize
                    n = d.iisz[p] * d.jisize * d.kjsize
                    dest[d.iist[p] - 1:d.iien[p], :, :] = \
                        r[pos:pos + n].reshape((d.iisz[p], d.jisize, d.kjsize), order="F")
                xb.append(dest)
        else:
            for d, s in zip(D, yb):
                dest = np.zeros((d.nxhp, d.jisize, d.kjsize), dtype=self.ct, order="F")
                dest[: d.nxhpc] = s
                xb.append(dest)
        # 5. X c2r                                           btran.F90:655
        return [np.asfortranarray(
            sfft.irfft(b, n=d.nx, axis=0, norm="forward", workers=_WORKERS).astype(self.rt))
            for d, b in zip(D, xb)]

    # ---- real-data transposes (module.F90:1061-1361) -----------------------------------
    def _alltoallv_real(self, sendbufs, group_of, snd_strt, snd_cnts, rcv_strt, rcv_cnts, recv_len):
        r = self.rt.itemsize
        recv = [np.zeros(recv_len(d), dtype=self.rt) for d in self.d]
        for d in self.d:
            grp = group_of(d)
            me = grp.index(d.rank)
            for p, peer in enumerate(grp):
                dp = self.d[peer]
                s0, n = snd_strt(d)[p] // r, snd_cnts(d)[p] // r
                r0, m = rcv_strt(dp)[me] // r, rcv_cnts(dp)[me] // r
                assert n == m, "alltoallv count mismatch"
                recv[peer][r0:r0 + n] = sendbufs[d.rank][s0:s0 + n]
        return recv

    def rtran(self, which, parts):
        """``which`` in {"x2y", "y2x", "x2z", "z2x"}; parts[r] = rank r's source array (Fortran order).
        Pack loops, MPI_Alltoallv tables and unpack loops as module.F90:1075-1116 (x2y), 1151-1193 (y2x),
        1228-1271 (x2z), 1306-1347 (z2x)."""
        D = self.d
        out = []
        if which == "x2y":
            send = []
            for d, s in zip(D, parts):
                buf = np.zeros(d.nx * d.jisize * d.kjsize, dtype=self.rt)
                pos = 0
                for i in range(d.iproc):
                    blk = s[d.iiist[i] - 1:d.iiien[i], :, :]
                    buf[pos:pos + blk.size] = blk.ravel(order="F")
                    pos += blk.size
                send.append(buf)
            recv = self._alltoallv_real(send, lambda d: d.row_ranks(), lambda d: d.IiStrt, lambda d: d.IiCnts,
                                        lambda d: d.JiStrt, lambda d: d.JiCnts, lambda d: d.iiisize * d.ny * d.kjsize)
            for d, r in zip(D, recv):
                dest = np.zeros((d.iiisize, d.ny, d.kjsize), dtype=self.rt, order="F")
                pos = 0
                for i in range(d.iproc):
                    n = d.iiisize * d.jisz[i] * d.kjsize
                    dest[:, d.jist[i] - 1:d.jien[i], :] = r[pos:pos + n].reshape((d.iiisize, d.jisz[i], d.kjsize), order="F")
                    pos += n
                out.append(dest)
        elif which == "y2x":
            send = []
            for d, s in zip(D, parts):
                buf = np.zeros(d.iiisize * d.ny * d.kjsize, dtype=self.rt)
                pos = 0
                for i in range(d.iproc):
                    blk = s[:, d.jist[i] - 1:d.jien[i], :]
                    buf[pos:pos + blk.size] = blk.ravel(order="F")
                    pos += blk.size
                send.append(buf)
            recv = self._alltoallv_real(send, lambda d: d.row_ranks(), lambda d: d.JiStrt, lambda d: d.JiCnts,
                                        lambda d: d.IiStrt, lambda d: d.IiCnts, lambda d: d.nx * d.jisize * d.kjsize)
            for d, r in zip(D, recv):
                dest = np.zeros((d.nx, d.jisize, d.kjsize), dtype=self.rt, order="F")
                pos = 0
                for i in range(d.iproc):
                    n = d.iiisz[i] * d.jisize * d.kjsize
                    dest[d.iiist[i] - 1:d.iiien[i], :, :] = r[pos:pos + n].reshape((d.iiisz[i], d.jisize, d.kjsize), order="F")
                    pos += n
                out.append(dest)
        elif which == "x2z":
            send = []
            for d, s in zip(D, parts):
                buf = np.zeros(d.nx * d.jisize * d.kjsize, dtype=self.rt)
                pos = 0
                for i in range(d.jproc):
                    blk = s[d.ijst[i] - 1:d.ijen[i], :, :]
                    buf[pos:pos + blk.size] = blk.ravel(order="F")
                    pos += blk.size
                send.append(buf)
            recv = self._alltoallv_real(send, lambda d: d.col_ranks(), lambda d: d.IjStrt, lambda d: d.IjCnts,
                                        lambda d: d.KjStrt, lambda d: d.KjCnts, lambda d: d.ijsize * d.jisize * d.nz)
            for d, r in zip(D, recv):
                dest = np.zeros((d.ijsize, d.jisize, d.nz), dtype=self.rt, order="F")
                pos = 0
                for i in range(d.jproc):
                    n = d.ijsize * d.jisize * d.kjsz[i]
                    dest[:, :, d.kjst[i] - 1:d.kjen[i]] = r[pos:pos + n].reshape((d.ijsize, d.jisize, d.kjsz[i]), order="F")
                    pos += n
                out.append(dest)
        elif which == "z2x":
            send = []
            for d, s in zip(D, parts):
                buf = np.zeros(d.ijsize * d.jisize * d.nz, dtype=self.rt)
                pos = 0
                for i in range(d.jproc):
                    blk = s[:, :, d.kjst[i] - 1:d.kjen[i]]
                    buf[pos:pos + blk.size] = blk.ravel(order="F")
                    pos += blk.size
                send.append(buf)
            recv = self._alltoallv_real(send, lambda d: d.col_ranks(), lambda d: d.KjStrt, lambda d: d.KjCnts,
                                        lambda d: d.IjStrt, lambda d: d.IjCnts, lambda d: d.nx * d.jisize * d.kjsize)
            for d, r in zip(D, recv):
                dest = np.zeros((d.nx, d.jisize, d.kjsize), dtype=self.rt, order="F")
                pos = 0
                for i in range(d.jproc):
                    n = d.ijsz[i] * d.jisize * d.kjsize
                    dest[d.ijst[i] - 1:d.ijen[i], :, :] = r[pos:pos + n].reshape((d.ijsz[i], d.jisize, d.kjsize), order="F")
                    pos += n
                out.append(dest)
        else:
            raise ValueError(which)
        return out

    def cheby(self, parts, Lz):
        out = self.forward(parts, "ffc")
        return [np.asfortranarray(cheby_epilogue(o, d, Lz).astype(self.ct)) for d, o in zip(self.d, out)]


# --------------------------------------------------------------------------------------
# real-data transposes, X-only transform, process-map queries, power spectrum
# --------------------------------------------------------------------------------------
def rtran_slices(d: Decomp, which: str):
    """(source slice, destination slice) of the GLOBAL real array for one rank: the pencil each side of
    rtran_x2y / y2x / x2z / z2x holds (module.F90:1064-1065, 1140-1141, 1217-1218, 1295-1296)."""
    X = (slice(None), slice(d.jistart - 1, d.jiend), slice(d.kjstart - 1, d.kjend))
    Y = (slice(d.iiistart - 1, d.iiiend), slice(None), slice(d.kjstart - 1, d.kjend))
    Z = (slice(d.ijstart - 1, d.ijend), slice(d.jistart - 1, d.jiend), slice(None))
    return {"x2y": (X, Y), "y2x": (Y, X), "x2z": (X, Z), "z2x": (Z, X)}[which]


def rtran_local(Aglobal, d: Decomp, which: str):
    """Destination array of one rank given the global field (the transposes move data, nothing else)."""
    return np.asfortranarray(Aglobal[rtran_slices(d, which)[1]])


def rtran_dims(d: Decomp, which: str):
    """dstart, dend, dsize returned by the transposes (module.F90:1118-1127, 1195-1204, 1273-1282, 1349-1358)."""
    if which == "x2y":
        return ([d.iiistart, 1, d.kjstart], [d.iiiend, d.ny, d.kjend], [d.iiisize, d.ny, d.kjsize])
    if which == "x2z":
        return ([d.ijstart, d.jistart, 1], [d.ijend, d.jiend, d.nz], [d.ijsize, d.jisize, d.nz])
    return ([1, d.jistart, d.kjstart], [d.nx, d.jiend, d.kjend], [d.nx, d.jisize, d.kjsize])


def forward_r2c_1d(Alocal):
    """p3dfft_ftran_r2c_1d (ftran.F90:787-814): r2c along x only, (nx, ji, kj) -> (nxhp, ji, kj)."""
    A = np.asfortranarray(Alocal)
    return np.asfortranarray(sfft.rfft(A, axis=0, workers=_WORKERS).astype(_ctype(A.dtype), copy=False))


class ProcGrid:
    """The reference's "trans2proc" tables and queries (module.F90:168-176, 788-1054; setup.F90:224-230, 551-577),
    restated literally with 1-based proc_parts rows.  Where the reference reads outside its arrays
    (proc_neighb accepts coord+orient == iproc, module.F90:813-822; the final loop of get_proc_parts starts at
    row 0, module.F90:1025) the restatement stops at the array bounds instead."""

    def __init__(self, nx, ny, nz, dims, nxc=None, nyc=None, nzc=None, dims_c=False, stride1=False):
        self.iproc, self.jproc = dims
        P = self.iproc * self.jproc
        self.d = [Decomp(nx, ny, nz, tuple(dims), r, nxc, nyc, nzc, dims_c=dims_c, stride1=stride1) for r in range(P)]
        self.proc_id2coords = []
        for d in self.d:
            self.proc_id2coords += [d.ipid, d.jpid]
        self.proc_coords2id = {(d.ipid, d.jpid): d.rank for d in self.d}
        # proc_dims(conf, 1..9, id) = start(3), end(3), size(3)
        self.proc_dims = {}
        for d in self.d:
            for conf in (1, 2):
                st, en, sz = d.get_dims(conf)
                for k, v in enumerate(list(st) + list(en) + list(sz), start=1):
                    self.proc_dims[(conf, k, d.rank)] = v

    def proc_neighb(self, base, orient, direction):
        if base < 0 or base >= self.iproc * self.jproc or orient not in (1, -1) or direction not in (1, 2):
            return -1
        ci, cj = self.proc_id2coords[2 * base], self.proc_id2coords[2 * base + 1]
        if direction == 1:
            return self.proc_coords2id.get((ci + orient, cj), -1)
        return self.proc_coords2id.get((ci, cj + orient), -1)

    def search_proc(self, point_i, point_j, di, dj, conf):
        if not (1 <= di <= 3 and 1 <= dj <= 3 and conf in (1, 2)):
            return -1
        pd = self.proc_dims
        pid = self.proc_coords2id[(0, 0)]
        while not point_i < pd[(conf, di, pid)] + pd[(conf, di + 6, pid)]:
            pid = self.proc_neighb(pid, 1, 1)
            if pid < 0:
                return -1
        while not point_j < pd[(conf, dj, pid)] + pd[(conf, dj + 6, pid)]:
            pid = self.proc_neighb(pid, 1, 2)
            if pid < 0:
                return -1
        return pid

    def get_proc_parts(self, base_x, base_y, base_z, size_x, size_y, size_z, conf):
        """-> (proc_parts as a list of P rows of 7 ints, no_parts, ierr).  module.F90:888-1054."""
        P = self.iproc * self.jproc
        parts = [[-1] * 7 for _ in range(P + 1)]       # row 0 unused (1-based like the reference)
        pd = self.proc_dims
        if conf == 1:
            base_i, base_j, base_k, size_i, size_j, size_k, di, dj = base_y, base_z, base_x, size_y, size_z, size_x, 2, 3
        elif conf == 2:
            base_i, base_j, base_k, size_i, size_j, size_k, di, dj = base_x, base_y, base_z, size_x, size_y, size_z, 1, 2
        else:
            return parts[1:], 0, 1
        found = self.search_proc(base_i, base_j, di, dj, conf)
        if found < 0:
            return parts[1:], 0, -1
        hi = lambda off, pid: pd[(conf, off, pid)] + pd[(conf, off + 6, pid)]
        n = 0
        end_i = 0
        while end_i == 0:
            n += 1
            parts[n][0] = found
            start_id_j, start_base_j, start_size_j = found, base_j, size_j
            if base_i + size_i <= hi(di, found):
                parts[n][1], parts[n][4] = base_i, size_i
                end_i = 1
            else:
                parts[n][1], parts[n][4] = base_i, hi(di, found) - base_i
                size_i = base_i + size_i - hi(di, found)
                base_i = hi(di, found)
            parts[n][3], parts[n][6] = base_k, size_k
            end_j, size_i_exist = 0, 1
            base_j, size_j = start_base_j, start_size_j
            while end_j == 0:
                if size_i_exist == 1:
                    size_i_exist = 0
                else:
                    n += 1
                    parts[n][0] = found
                    parts[n][3], parts[n][4], parts[n][6] = parts[n - 1][3], parts[n - 1][4], parts[n - 1][6]
                if base_j + size_j <= hi(dj, found):
                    parts[n][2], parts[n][5] = base_j, size_j
                    end_j = 1
                else:
                    parts[n][2], parts[n][5] = base_j, hi(dj, found) - base_j
                    size_j = base_j + size_j - hi(dj, found)
                    base_j = hi(dj, found)
                if end_j == 0:
                    found = self.proc_neighb(found, 1, 2)
                    if found < 0:
                        break
            if end_i == 0:
                found = self.proc_neighb(start_id_j, 1, 1)
                if found < 0:
                    break
        if conf == 1:
            for p in range(1, P + 1):
                bi, bj, bk, si, sj, sk = parts[p][1:7]
                parts[p][1:7] = [bk, bi, bj, sk, si, sj]
        return parts[1:], n, 0


def spectrum_kmax(nx, ny, nz):
    """driver_spec.c:228: kmax = sqrt(nx^2 + ny^2 + nz^2) * 0.5 + 0.5 truncated."""
    return int(np.sqrt(float(nx * nx + ny * ny + nz * nz)) * 0.5 + 0.5)


def power_spectrum(Blocal, d: Decomp, kmax=None, factor=1.0):
    """compute_spectrum of sample/C/driver_spec.c:298-384 on one rank's wavenumber array (get_dims(2) layout,
    (nzc, jjsize, iisize) with STRIDE1): el[ik] += k2 * (re^2 + im^2) over the block, with kx = global x index
    (never folded: only kx <= nx/2 is stored), ky and kz the global indices folded about n/2 (:352-358), and
    ik = int(sqrt(k2) + 0.5) (:364-368).  ``factor`` multiplies B first (the driver's mult_array, :223).
    For a pruned transform the stored y/z indices are first mapped to the modes they hold (kept_y / kept_z);
    the reference driver does not prune, where the two agree.  The per-rank results add up to the global
    spectrum (MPI_Reduce, :381)."""
    B = np.asarray(Blocal)
    if d.stride1:
        B = B.transpose(2, 1, 0)
    kmax = spectrum_kmax(d.nx, d.ny, d.nz) if kmax is None else kmax
    kx = np.arange(d.iistart - 1, d.iiend)
    ky = d.kept_y()[d.jjstart - 1:d.jjend]
    ky = np.where(ky > d.ny // 2, d.ny - ky, ky)
    kz = d.kept_z()
    kz = np.where(kz > d.nz // 2, d.nz - kz, kz)
    k2 = kx[:, None, None] ** 2 + ky[None, :, None] ** 2 + kz[None, None, :] ** 2
    ik = (np.sqrt(k2.astype(np.float64)) + 0.5).astype(np.int64)
    w = k2 * (B.real.astype(np.float64) ** 2 + B.imag.astype(np.float64) ** 2) * float(factor) ** 2
    keep = ik <= kmax
    return np.bincount(ik[keep].ravel(), weights=w[keep].ravel(), minlength=kmax + 1)[: kmax + 1]


# --------------------------------------------------------------------------------------
# synthetic inputs and metrics
# --------------------------------------------------------------------------------------
def philox_field(nx, ny, nz, seed=20240229, dtype=np.float64, sl=None):
    """Uniform [0,1) field keyed by GLOBAL index, like
    driver_rand.c:194's rand()/RAND_MAX but reproducible for any rank count.
    ``sl`` = (yslice, zslice) restricts generation to an X-pencil."""
    ys = range(ny)[sl[0]] if sl else range(ny)
    zs = range(nz)[sl[1]] if sl else range(nz)
    out = np.empty((nx, len(ys), len(zs)), dtype=dtype, order="F")
    for kk, z in enumerate(zs):
        # one Philox stream per z-plane: value = f(seed, z)[x + nx*y]
        g = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, z]))
        plane = g.random(nx * ny).reshape((nx, ny), order="F")
        out[:, :, kk] = plane[:, ys.start:ys.stop] if isinstance(ys, range) else plane[:, ys]
    return out


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.linalg.norm(b.ravel().astype(np.complex128 if np.iscomplexobj(b) else np.float64))
    num = np.linalg.norm((a.ravel() - b.ravel()).astype(
        np.complex128 if np.iscomplexobj(b) else np.float64))
    return float(num / den) if den > 0 else float(num)


# --------------------------------------------------------------------------------------
# CPU baseline: the reference's single-rank stage sequence on a bounded sample
# --------------------------------------------------------------------------------------
def cpu_pair_sampled(nx, ny, nz, frac_planes, dtype=np.float64, workers=None, seed=1):
    """Times the reference's P=1 forward+backward stage sequence (ftran.F90:530,554,581-583,
    756-757,660-683; btran.F90:437-466,618-623,663-671,655: FFT executes plus the seg_copy /
    ar_copy passes) on 1/frac of the 1D lines of every stage at FULL line length, and
    extrapolates linearly.  X and Y stages run on nz/frac z-planes; the Z stage runs on an
    x-slab of nxhp/frac columns.  Returns (seconds_for_full_pair_extrapolated, seconds_measured,
    description)."""
    import time
    w = workers or os.cpu_count()
    rt = np.dtype(dtype)
    ct = np.dtype(_ctype(dtype))
    nxhp = nx // 2 + 1
    nzs = max(1, nz // frac_planes)
    nxs = max(1, nxhp // frac_planes)
    rng = np.random.default_rng(seed)
    A = np.asfortranarray(rng.random((nx, ny, nzs)).astype(rt))
    Zin = np.asfortranarray((rng.random((nxs, ny, nz)) + 1j * rng.random((nxs, ny, nz))).astype(ct))
    t = {}
    t0 = time.perf_counter()
    # forward
    b2 = sfft.rfft(A, axis=0, workers=w)                       # exec_f_r2c
    buf = np.array(b2[:nxhp], order="F", copy=True)            # seg_copy_x
    buf = sfft.fft(buf, axis=1, workers=w, overwrite_x=True)   # Y, per z-plane
    out = np.array(buf, order="F", copy=True)                  # seg_copy_y x2 into XYZg
    t1 = time.perf_counter()
    zf = sfft.fft(Zin, axis=2, workers=w)                      # exec_f_c2_same on the user array
    t2 = time.perf_counter()
    # backward
    zb = sfft.ifft(zf, axis=2, norm="forward", workers=w)      # exec_b_c2_same
    zc = np.array(zb, order="F", copy=True)                    # ar_copy into buf
    t3 = time.perf_counter()
    yb = sfft.ifft(out, axis=1, norm="forward", workers=w)     # Y inverse
    xb = np.array(yb, order="F", copy=True)                    # seg_copy_x + seg_zero_x into buf1
    R = sfft.irfft(xb, n=nx, axis=0, norm="forward", workers=w)   # exec_b_c2r
    t4 = time.perf_counter()
    meas_xy = (t1 - t0) + (t4 - t3)
    meas_z = (t3 - t1)
    full = meas_xy * (nz / nzs) + meas_z * (nxhp / nxs)
    desc = (f"reference P=1 stage sequence restated with scipy/pocketfft ({w} threads): X,Y stages on {nzs}/{nz} "
            f"z-planes, Z stage on {nxs}/{nxhp} x-columns, all 1D lines at full length; linearly extrapolated")
    del R, zc
    return full, meas_xy + meas_z, desc


def _kept(n, nc):
    """stored index -> logical mode index of a pruned axis (lower half, then the upper modes)"""
    h = (nc + 1) // 2
    return np.concatenate([np.arange(h), np.arange(n - (nc - h), n)]).astype(np.int64)


def cpu_pair_measured(nx, ny, nz, dtype="f64", op="fft", workers=None, budget_s=150.0, max_steps=20, cut=None,
                      warmup=0, min_fraction=0.125, seed=1):
    """The reference's P=1 forward+backward stage sequence (ftran.F90:530,554,581-583,756-757,605-683;
    btran.F90:437-466,471-509,618-623,655-671: the FFT executes plus the seg_copy / ar_copy passes), restated with
    pocketfft and run on COMPLETE fields: the X and Y stages over z-slabs, the Z stage over x-slabs, `workers`
    slabs in flight (one thread each -- the parallelism the reference gets from its MPI ranks), thread count fixed
    here and independent of OMP_NUM_THREADS.

    op: 'fft' (forward fft, backward tff), 'cheby' (forward ffc + the Chebyshev epilogue ftran.F90:408-451, backward
    cff) or 'pruned' (fft/tff with `cut`).  Steps are whole pairs; as many as fit `budget_s` (at least one, at most
    `max_steps`).  When one pair is predicted (from a probe of 1/32 of the slabs) not to fit the budget, only every
    k-th slab of each stage is executed (fraction >= `min_fraction`) and the pair time is extrapolated linearly --
    the returned description says which.  Returns a dict."""
    import time
    from concurrent.futures import ThreadPoolExecutor
    w = int(workers or os.cpu_count() or 1)
    rt = np.dtype(np.float32 if dtype in ("f32", np.float32) else np.float64)
    ct = np.dtype(_ctype(rt))
    nxc, nyc, nzc = cut if (cut and op == "pruned") else (nx, ny, nz)
    nxhp, nxhpc = nx // 2 + 1, nxc // 2 + 1
    ky, kz = _kept(ny, nyc), _kept(nz, nzc)
    zch = "c" if op == "cheby" else "f"
    SZ = max(1, min(8, nz // (2 * w) or 1))            # z-planes per slab of the X/Y stages
    SX = max(1, min(8, nxhpc // (2 * w) or 1))         # x-columns per slab of the Z stage
    zslabs = [(z, min(z + SZ, nz)) for z in range(0, nz, SZ)]
    xslabs = [(x, min(x + SX, nxhpc)) for x in range(0, nxhpc, SX)]
    pool = ThreadPoolExecutor(max_workers=w)
    A = np.empty((nx, ny, nz), dtype=rt, order="F")

    def fill(sl):
        A[:, :, sl[0]:sl[1]] = np.random.default_rng(seed + sl[0]).random((nx, ny, sl[1] - sl[0]), dtype=np.float64).astype(rt, copy=False)
    list(pool.map(fill, zslabs))
    Y = np.empty((nxhpc, nyc, nz), dtype=ct, order="F")       # buf after the Y stage (y-pruned), then XYZg (z-pruned view)
    F = Y if nzc == nz else np.empty((nxhpc, nyc, nzc), dtype=ct, order="F")
    B = np.empty((nx, ny, nz), dtype=rt, order="F")
    Lz = 2.0

    def fwd_xy(sl):
        a = A[:, :, sl[0]:sl[1]]
        b = sfft.rfft(a, axis=0, workers=1)                    # exec_f_r2c (ftran.F90:530)
        b = np.array(b[:nxhpc], order="F", copy=True)          # seg_copy_x (ftran.F90:554)
        b = sfft.fft(b, axis=1, workers=1, overwrite_x=True)   # Y stage, one z-plane at a time (ftran.F90:581-583)
        Y[:, :, sl[0]:sl[1]] = b if nyc == ny else b[:, ky, :]   # seg_copy_y (ftran.F90:756-757)

    def fwd_z(sl):
        v = Y[sl[0]:sl[1]]
        if zch == "c":                                         # exec_ctrans_r2_complex_same: re and im apart (fft_exec.F90:646-701)
            r = sfft.dct(v.real, type=1, axis=2, workers=1) + 1j * sfft.dct(v.imag, type=1, axis=2, workers=1)
        else:
            r = sfft.fft(v, axis=2, workers=1)                 # exec_f_c2_same on the user array (ftran.F90:605-683)
        if op == "cheby":
            r = cheby_epilogue(r, _ChebyDims(nx, ny, nzc), Lz)
        F[sl[0]:sl[1]] = r if nzc == nz else r[:, :, kz]        # seg_copy_z (ftran.F90:645-646)

    def bwd_z(sl):
        if nzc == nz:
            v = F[sl[0]:sl[1]]
        else:                                                  # zero-fill of the pruned band (btran.F90:471-509)
            v = np.zeros((sl[1] - sl[0], nyc, nz), dtype=ct, order="F")
            v[:, :, kz] = F[sl[0]:sl[1]]
        if zch == "c":
            r = sfft.dct(v.real, type=1, axis=2, workers=1) + 1j * sfft.dct(v.imag, type=1, axis=2, workers=1)
        else:
            r = sfft.ifft(v, axis=2, norm="forward", workers=1)     # exec_b_c2_same (btran.F90:437-466)
        Y[sl[0]:sl[1]] = r                                     # ar_copy into buf

    def bwd_yx(sl):
        if nyc == ny and nxhpc == nxhp:
            v = Y[:, :, sl[0]:sl[1]]
        else:                                                  # seg_copy_x / seg_zero_x, zero-fill of the pruned Y band (bcomm1.F90:367-374)
            v = np.zeros((nxhp, ny, sl[1] - sl[0]), dtype=ct, order="F")
            v[:nxhpc, ky, :] = Y[:, :, sl[0]:sl[1]]
        yb = sfft.ifft(v, axis=1, norm="forward", workers=1)    # Y inverse per z-plane (btran.F90:618-623)
        xb = np.array(yb, order="F", copy=True)                # unpack copy into buf1 (btran.F90:663-671)
        B[:, :, sl[0]:sl[1]] = sfft.irfft(xb, n=nx, axis=0, norm="forward", workers=1)   # exec_b_c2r (btran.F90:655)

    stages = (("fwd_xy", fwd_xy, zslabs), ("fwd_z", fwd_z, xslabs), ("bwd_z", bwd_z, xslabs), ("bwd_yx", bwd_yx, zslabs))

    def run_pair(stride):
        ts = {}
        for name, fn, slabs in stages:
            t0 = time.perf_counter()
            sub = slabs[::stride]
            list(pool.map(fn, sub))
            ts[name] = (time.perf_counter() - t0) * (len(slabs) / len(sub))
        return ts

    # probe: 1/32 of the slabs of every stage -> estimated seconds per complete pair (also warms the thread pool)
    probe = max(1, min(32, len(zslabs) // w, len(xslabs) // w))
    est = sum(run_pair(probe).values()) if probe > 1 else None
    stride = 1
    if est is not None and est > budget_s:
        stride = 2
        while est / stride > budget_s and 1.0 / (2 * stride) >= min_fraction:
            stride *= 2
    per = (est / stride) if est is not None else None
    steps = 1 if per is None else int(max(1, min(max_steps, budget_s // max(per, 1e-9))))
    for _ in range(warmup):
        run_pair(stride)
    t_steps, acc = [], {}
    wall0 = time.perf_counter()
    for _ in range(steps):
        ts = run_pair(stride)
        t_steps.append(sum(ts.values()))
        for k, v in ts.items():
            acc[k] = acc.get(k, 0.0) + v / steps
    wall = time.perf_counter() - wall0
    pool.shutdown()
    size = f"{nx}^3" if nx == ny == nz else f"{nx}x{ny}x{nz}"
    whole = stride == 1
    # self-check on the first slab (always executed): backward(forward(A)) = N * A for the band the transform keeps
    rt_err = None
    if op == "fft" and whole:
        sl = zslabs[0]
        rt_err = float(np.abs(B[:, :, sl[0]:sl[1]] / (float(nx) * ny * nz) - A[:, :, sl[0]:sl[1]]).max())
    sample = (f"reference P=1 stage sequence restated with scipy/pocketfft, {w} threads (one slab per thread): "
              + (f"{steps} complete {size} pair(s) on a complete field" if whole else
                 f"every {stride}-th slab of every stage of a {size} pair (fraction 1/{stride}), linearly extrapolated; {steps} step(s)"))
    return {"ms_per_pair": 1e3 * sum(t_steps) / steps, "ms_per_executed_step": 1e3 * wall / steps, "steps": steps,
            "warmup": warmup, "fraction": 1.0 / stride, "executed": "complete pairs" if whole else f"1/{stride} of the slabs per step",
            "sample": sample, "stage_s": acc, "cores": w, "roundtrip_max_err": rt_err}


class _ChebyDims:
    """the three integers cheby_epilogue reads"""
    def __init__(self, nx, ny, nzc):
        self.nx, self.ny, self.nzc = nx, ny, nzc
