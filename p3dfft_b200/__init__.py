"""p3dfft_b200 -- host-side mirror of the P3DFFT interface over the B200 C-ABI library.

The product is the shared library ``p3dfft_b200/lib/libp3dfft[_single].so`` (C++/CUDA,
``p3dfft_b200/csrc``); its entry points are declared in ``include/p3dfft.h`` and
``include/p3dfft_b200.h``.  This module is a thin ctypes binding with the reference's names and
argument meaning (``build/module.F90:178-186`` public list), used by the tests, ``bench.py``
and Python callers.  It contains no transform arithmetic and no CPU fallback: if the
library is missing, importing :func:`load` raises.

Arrays are passed as raw addresses: numpy arrays (host memory, staged over PCIe by the
library) or torch CUDA tensors (device memory, used in place).  All arrays are Fortran
ordered (x fastest) exactly as the reference expects.
"""
from __future__ import annotations

import ctypes as C
import os

__all__ = ["load", "P3DFFT", "LibraryMissing", "Step", "Stage", "Exchange", "Seg", "Side", "DecompInfo"]

_HERE = os.path.dirname(os.path.abspath(__file__))
MAXSEG, MAXFAC = 16, 24


class LibraryMissing(RuntimeError):
    pass


# ---- ctypes mirrors of csrc/stage.h ------------------------------------------------------
class Seg(C.Structure):
    _fields_ = [("base", C.c_void_p), ("buf", C.c_int32), ("peer", C.c_int32), ("off", C.c_int64),
                ("start", C.c_int32), ("len", C.c_int32), ("ps", C.c_int64), ("sa", C.c_int64),
                ("sb", C.c_int64), ("sc", C.c_int64), ("kw", C.c_int32), ("aw", C.c_int32),
                ("psh", C.c_int64), ("sah", C.c_int64), ("bw", C.c_int32), ("pad_", C.c_int32), ("sbh", C.c_int64)]


class Side(C.Structure):
    _fields_ = [("nseg", C.c_int32), ("cnt", C.c_int32), ("h1", C.c_int32), ("logical", C.c_int32),
                ("seg", Seg * MAXSEG)]


class Stage(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n", C.c_int32), ("nfft", C.c_int32), ("na", C.c_int32),
                ("nb", C.c_int32), ("nc", C.c_int32), ("tile", C.c_int32), ("layx", C.c_int32),
                ("need_zero", C.c_int32), ("nfac", C.c_int32), ("fac", C.c_int32 * MAXFAC),
                ("timer", C.c_int32), ("bord", C.c_int32), ("tw", C.c_void_p), ("scale", C.c_double),
                ("inp", Side), ("out", Side)]


class Exchange(C.Structure):
    _fields_ = [("comm", C.c_int32), ("npeer", C.c_int32), ("self", C.c_int32), ("sendbuf", C.c_int32),
                ("recvbuf", C.c_int32), ("timer", C.c_int32), ("p2p", C.c_int32), ("ebytes", C.c_int32),
                ("sndoff", C.c_int64 * MAXSEG), ("sndcnt", C.c_int64 * MAXSEG),
                ("rcvoff", C.c_int64 * MAXSEG), ("rcvcnt", C.c_int64 * MAXSEG)]


class Step(C.Structure):
    _fields_ = [("is_exchange", C.c_int32), ("pad_", C.c_int32), ("st", Stage), ("ex", Exchange)]


class DecompInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "nx", "ny", "nz", "nxc", "nyc", "nzc", "nxhp", "nxhpc", "nycph", "nzcph", "iproc", "jproc", "ipid", "jpid",
        "iistart", "iiend", "iisize", "jistart", "jiend", "jisize", "jjstart", "jjend", "jjsize",
        "kjstart", "kjend", "kjsize", "padi_work", "padi")] + [
        ("memsize", C.c_int32 * 3), ("nm", C.c_int64), ("work_elems", C.c_int64)]


RTRAN_NAMES = ("x2y", "y2x", "x2z", "z2x")      # index = `which` of the planner (csrc/plan.h RtranKind)
KIND_NAMES = {0: "c2c_fwd", 1: "c2c_bwd", 2: "r2c", 3: "c2r", 4: "dct1", 5: "dst1", 6: "noop", 7: "rcopy"}
BUF_USER_IN, BUF_USER_OUT, BUF_A, BUF_B, BUF_C = 0, 1, 2, 3, 4


def lib_path(single: bool = False) -> str:
    # P3DFFT_B200_LIB_SUFFIX=_x selects lib/libp3dfft_x.so: experimental builds side by side with the product (tools/)
    suf = os.environ.get("P3DFFT_B200_LIB_SUFFIX", "")
    return os.path.join(_HERE, "lib", f"libp3dfft{'_single' if single else ''}{suf}.so")


def _addr(x) -> int:
    """Raw address of a numpy array or torch tensor (no copies are ever made here)."""
    if x is None:
        return 0
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    if hasattr(x, "ctypes"):
        return int(x.ctypes.data)
    if isinstance(x, int):
        return x
    raise TypeError(f"cannot take the address of {type(x)}")


class P3DFFT:
    """One loaded library = one P3DFFT "module" (one plan at a time, module.F90:103-176)."""

    def __init__(self, single: bool = False, path: str | None = None):
        path = path or lib_path(single)
        if not os.path.exists(path):
            raise LibraryMissing(f"{path} not built; run `python p3dfft_b200/build.py` (no CPU fallback exists)")
        self.single = single
        self.lib = lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        if bool(lib.p3dfft_b200_build_flags() & 1) != bool(single):
            raise LibraryMissing(f"{path} has the wrong precision")
        self.real = C.c_float if single else C.c_double
        ip, vp = C.POINTER(C.c_int), C.c_void_p
        lib.p3dfft_setup.argtypes = [ip] * 10
        lib.p3dfft_setup.restype = None
        lib.p3dfft_get_dims.argtypes = [ip] * 4
        lib.p3dfft_ftran_r2c.argtypes = [vp, vp, C.c_char_p]
        lib.p3dfft_btran_c2r.argtypes = [vp, vp, C.c_char_p]
        lib.p3dfft_ftran_r2c_many.argtypes = [vp, ip, vp, ip, ip, C.c_char_p]
        lib.p3dfft_btran_c2r_many.argtypes = [vp, ip, vp, ip, ip, C.c_char_p]
        lib.p3dfft_cheby.argtypes = [vp, vp, C.POINTER(self.real)]
        lib.p3dfft_cheby_many.argtypes = [vp, ip, vp, ip, ip, C.POINTER(self.real)]
        lib.get_timers.argtypes = [C.POINTER(C.c_double)]
        lib.p3dfft_b200_get_unique_id.argtypes = [vp]
        lib.p3dfft_b200_comm_create.argtypes = [C.c_int, C.c_int, vp, C.c_int]
        lib.p3dfft_b200_last_error.argtypes = [C.c_char_p, C.c_int]
        lib.p3dfft_b200_set_stream.argtypes = [vp]
        lib.p3dfft_b200_launch_count.restype = C.c_longlong
        lib.p3dfft_b200_launch_count.argtypes = [C.c_int]
        lib.p3dfft_b200_fast_launch_count.restype = C.c_longlong
        lib.p3dfft_b200_fast_launch_count.argtypes = [C.c_int]
        lib.p3dfft_b200_plan_decomp.argtypes = [ip] + [C.c_int] * 8 + [C.POINTER(DecompInfo)]
        lib.p3dfft_b200_plan_steps.argtypes = [ip] + [C.c_int] * 9 + [C.c_char_p, C.c_int, C.c_int64, C.c_int64,
                                                                     C.c_int, vp, C.c_int]
        lib.p3dfft_ftran_r2c_1d.argtypes = [vp, vp]
        lib.p3dfft_ftran_r2c_1d.restype = None
        for w in RTRAN_NAMES:
            f = getattr(lib, "p3dfft_b200_rtran_" + w)
            f.argtypes = [vp, vp, ip, ip, ip, C.POINTER(C.c_double)]
            f.restype = None
        lib.p3dfft_get_mpi_info.argtypes = [ip, ip, ip]
        lib.p3dfft_b200_set_scale.argtypes = [C.c_double, C.c_double]
        lib.p3dfft_b200_set_scale.restype = None
        lib.p3dfft_b200_spectrum.argtypes = [vp, C.c_double, vp, C.c_int]
        lib.p3dfft_b200_spectrum.restype = None
        lib.p3dfft_b200_proc_id2coords.argtypes = [C.c_int, ip, ip]
        lib.p3dfft_b200_proc_coords2id.argtypes = [C.c_int, C.c_int]
        lib.p3dfft_b200_proc_dims.argtypes = [C.c_int, C.c_int, ip]
        lib.p3dfft_b200_proc_neighb.argtypes = [C.c_int] * 3
        lib.p3dfft_b200_get_proc_parts.argtypes = [C.c_int] * 7 + [ip, ip]
        lib.p3dfft_b200_plan_aux_steps.argtypes = [ip] + [C.c_int] * 10 + [vp, C.c_int]
        lib.p3dfft_b200_plan_rtran_info.argtypes = [ip] + [C.c_int] * 6 + [ip]
        lib.p3dfft_b200_plan_rtran_info.restype = C.c_longlong
        lib.p3dfft_b200_plan_proc_parts.argtypes = [ip] + [C.c_int] * 14 + [ip, ip]
        lib.p3dfft_b200_plan_proc_neighb.argtypes = [ip] + [C.c_int] * 4
        if lib.p3dfft_b200_sizeof_step() != C.sizeof(Step):
            raise LibraryMissing("ctypes mirror of P3dStep is out of date")
        lib.p3dfft_b200_set_error_mode(1)     # Python callers get exceptions instead of abort()

    # ---- errors ------------------------------------------------------------------------
    def _check(self):
        buf = C.create_string_buffer(2048)
        if self.lib.p3dfft_b200_last_error(buf, 2048) > 0:
            raise RuntimeError(buf.value.decode(errors="replace"))

    # ---- reference API -----------------------------------------------------------------
    def p3dfft_setup(self, dims, nx, ny, nz, comm=0, nxcut=None, nycut=None, nzcut=None, overwrite=True):
        """``p3dfft_setup`` (build/setup.F90:107); returns ``memsize``."""
        d = (C.c_int * 2)(*dims)
        mem = (C.c_int * 3)()
        i = lambda v: C.byref(C.c_int(int(v)))
        self.lib.p3dfft_setup(d, i(nx), i(ny), i(nz), i(comm), i(nx if nxcut is None else nxcut),
                              i(ny if nycut is None else nycut), i(nz if nzcut is None else nzcut),
                              i(1 if overwrite else 0), mem)
        self._check()
        return tuple(mem)

    def p3dfft_get_dims(self, conf):
        """``p3dfft_get_dims`` (build/module.F90:225): (istart, iend, isize), 1-based."""
        a, b, c = (C.c_int * 3)(), (C.c_int * 3)(), (C.c_int * 3)()
        self.lib.p3dfft_get_dims(a, b, c, C.byref(C.c_int(conf)))
        self._check()
        return tuple(a), tuple(b), tuple(c)

    def p3dfft_ftran_r2c(self, A, B, op="fft"):
        self.lib.p3dfft_ftran_r2c(_addr(A), _addr(B), op.encode() + b"\0")
        self._check()

    def p3dfft_btran_c2r(self, A, B, op="tff"):
        self.lib.p3dfft_btran_c2r(_addr(A), _addr(B), op.encode() + b"\0")
        self._check()

    def p3dfft_ftran_r2c_many(self, A, dim_in, B, dim_out, nv, op="fft"):
        i = lambda v: C.byref(C.c_int(int(v)))
        self.lib.p3dfft_ftran_r2c_many(_addr(A), i(dim_in), _addr(B), i(dim_out), i(nv), op.encode() + b"\0")
        self._check()

    def p3dfft_btran_c2r_many(self, A, dim_in, B, dim_out, nv, op="tff"):
        i = lambda v: C.byref(C.c_int(int(v)))
        self.lib.p3dfft_btran_c2r_many(_addr(A), i(dim_in), _addr(B), i(dim_out), i(nv), op.encode() + b"\0")
        self._check()

    def p3dfft_cheby(self, A, B, Lz):
        self.lib.p3dfft_cheby(_addr(A), _addr(B), C.byref(self.real(Lz)))
        self._check()

    def p3dfft_cheby_many(self, A, dim_in, B, dim_out, nv, Lz):
        i = lambda v: C.byref(C.c_int(int(v)))
        self.lib.p3dfft_cheby_many(_addr(A), i(dim_in), _addr(B), i(dim_out), i(nv), C.byref(self.real(Lz)))
        self._check()

    def p3dfft_clean(self):
        self.lib.p3dfft_clean()

    def get_timers(self):
        t = (C.c_double * 12)()
        self.lib.get_timers(t)
        return list(t)

    def set_timers(self):
        self.lib.set_timers()

    # ---- remaining module routines (module.F90:178-186) ----------------------------------------
    def p3dfft_ftran_r2c_1d(self, A, B):
        """``p3dfft_ftran_r2c_1d`` (build/ftran.F90:787): X transform only."""
        self.lib.p3dfft_ftran_r2c_1d(_addr(A), _addr(B))
        self._check()

    def rtran(self, which, source, dest, t=0.0):
        """``rtran_x2y`` / ``rtran_y2x`` / ``rtran_x2z`` / ``rtran_z2x`` (build/module.F90:1061-1361).
        Returns (dstart, dend, dsize, t)."""
        a, b, c = (C.c_int * 3)(), (C.c_int * 3)(), (C.c_int * 3)()
        tt = C.c_double(t)
        getattr(self.lib, "p3dfft_b200_rtran_" + which)(_addr(source), _addr(dest), a, b, c, C.byref(tt))
        self._check()
        return tuple(a), tuple(b), tuple(c), tt.value

    def p3dfft_get_mpi_info(self):
        a, b, c = C.c_int(-1), C.c_int(-1), C.c_int(-1)
        self.lib.p3dfft_get_mpi_info(C.byref(a), C.byref(b), C.byref(c))
        self._check()
        return a.value, b.value, c.value

    def proc_id2coords(self, pid):
        a, b = C.c_int(-1), C.c_int(-1)
        rc = self.lib.p3dfft_b200_proc_id2coords(pid, C.byref(a), C.byref(b))
        self._check()
        return (a.value, b.value) if rc == 0 else None

    def proc_coords2id(self, ipid, jpid):
        return int(self.lib.p3dfft_b200_proc_coords2id(ipid, jpid))

    def proc_dims(self, conf, pid):
        o = (C.c_int * 9)()
        rc = self.lib.p3dfft_b200_proc_dims(conf, pid, o)
        self._check()
        return list(o) if rc == 0 else None

    def proc_neighb(self, base, orient, direction):
        return int(self.lib.p3dfft_b200_proc_neighb(base, orient, direction))

    def get_proc_parts(self, base, size, conf, nproc):
        """``get_proc_parts`` (build/module.F90:888): (rows of 7 ints, number of parts, ierr)."""
        parts = (C.c_int * (7 * nproc))()
        ierr = C.c_int(0)
        n = self.lib.p3dfft_b200_get_proc_parts(*base, *size, conf, parts, C.byref(ierr))
        self._check()
        return [list(parts[7 * i:7 * i + 7]) for i in range(nproc)], n, ierr.value

    def set_scale(self, forward=1.0, backward=1.0):
        """Fused normalisation of the transforms' outputs (the drivers' ``mult_array`` pass)."""
        self.lib.p3dfft_b200_set_scale(float(forward), float(backward))

    def spectrum(self, B, kmax, factor=1.0, out=None):
        """``compute_spectrum`` of sample/C/driver_spec.c:298-384 on the device; returns kmax+1 doubles."""
        import numpy as np
        E = np.zeros(kmax + 1) if out is None else out
        self.lib.p3dfft_b200_spectrum(_addr(B), float(factor), _addr(E), int(kmax))
        self._check()
        return E

    # ---- extensions ----------------------------------------------------------------------
    def set_layout(self, stride1=False, dims_c=False):
        self.lib.p3dfft_b200_set_layout(int(stride1), int(dims_c))

    def get_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        if self.lib.p3dfft_b200_get_unique_id(buf) != 0:
            self._check()
            raise RuntimeError("ncclGetUniqueId failed")
        return buf.raw

    def comm_create(self, rank, size, unique_id: bytes | None, device=-1) -> int:
        buf = C.create_string_buffer(unique_id or b"\0" * 128, 128)
        h = self.lib.p3dfft_b200_comm_create(rank, size, buf, device)
        if h < 0:
            self._check()
            raise RuntimeError(f"p3dfft_b200_comm_create failed ({h})")
        return h

    def comm_destroy(self, handle):
        self.lib.p3dfft_b200_comm_destroy(handle)

    def set_stream(self, cuda_stream_ptr):
        self.lib.p3dfft_b200_set_stream(cuda_stream_ptr)

    def reset_stream(self):
        self.lib.p3dfft_b200_reset_stream()

    def force_generic(self, on=True):
        self.lib.p3dfft_b200_force_generic(int(on))

    def fast_launch_count(self, reset=False) -> int:
        return int(self.lib.p3dfft_b200_fast_launch_count(int(reset)))

    def set_async(self, flag):
        self.lib.p3dfft_b200_set_async(int(flag))

    def sync(self):
        self.lib.p3dfft_b200_sync()

    def launch_count(self, reset=False) -> int:
        return int(self.lib.p3dfft_b200_launch_count(int(reset)))

    # ---- host-only planner ---------------------------------------------------------------
    def set_p2p(self, on=True):
        self.lib.p3dfft_b200_set_p2p(int(on))

    def p2p_active(self) -> bool:
        return bool(self.lib.p3dfft_b200_p2p_active())

    def plain_layout(self, on=True):
        self.lib.p3dfft_b200_plain_layout(int(on))

    def row_bytes(self, rb=0):
        """Tile row width of the internal layouts: 0 = planner's rule, 64 or 128 forced (next setup)."""
        self.lib.p3dfft_b200_row_bytes(int(rb))

    def plan_decomp(self, dims, nx, ny, nz, rank=0, nxc=None, nyc=None, nzc=None, stride1=False, dims_c=False,
                    plain=False, row_bytes=0):
        info = DecompInfo()
        d = (C.c_int * 2)(*dims)
        flags = (1 if self.single else 0) | (2 if stride1 else 0) | (4 if dims_c else 0) | (8 if plain else 0)
        flags |= 32 if row_bytes == 64 else 64 if row_bytes == 128 else 0
        rc = self.lib.p3dfft_b200_plan_decomp(d, nx, ny, nz, rank, nxc or nx, nyc or ny, nzc or nz, flags,
                                              C.byref(info))
        if rc != 0:
            self._check()
            raise RuntimeError("plan_decomp failed")
        return info

    def plan_steps(self, dims, nx, ny, nz, rank, backward, op, nv=1, nxc=None, nyc=None, nzc=None, stride1=False,
                   dims_c=False, dim_real=None, dim_cplx=None, plain=False, p2p=False, row_bytes=0, overlap=0):
        info = self.plan_decomp(dims, nx, ny, nz, rank, nxc, nyc, nzc, stride1, dims_c, plain, row_bytes)
        if dim_real is None:
            dim_real = info.nx * info.jisize * info.kjsize
        if dim_cplx is None:
            dim_cplx = info.iisize * info.jjsize * info.nzc
        arr = (Step * 512)()
        d = (C.c_int * 2)(*dims)
        flags = (2 if stride1 else 0) | (4 if dims_c else 0) | (8 if plain else 0) | (16 if p2p else 0)
        flags |= 32 if row_bytes == 64 else 64 if row_bytes == 128 else 0
        flags |= (int(overlap) & 0xff) << 8
        n = self.lib.p3dfft_b200_plan_steps(d, nx, ny, nz, rank, nxc or nx, nyc or ny, nzc or nz, flags,
                                            1 if backward else 0, op.encode() + b"\0", nv, dim_real, dim_cplx,
                                            4 if self.single else 8, arr, 512)
        if n < 0:
            self._check()
            raise RuntimeError("plan_steps failed")
        return [arr[i] for i in range(n)], info


    def plan_aux_steps(self, dims, nx, ny, nz, rank, which, p2p=False, dims_c=False, nxc=None, nyc=None, nzc=None):
        """Step list of ``p3dfft_ftran_r2c_1d`` (which = "r2c_1d") or of a real-data transpose ("x2y", ...)."""
        w = 100 if which == "r2c_1d" else RTRAN_NAMES.index(which)
        arr = (Step * 8)()
        d = (C.c_int * 2)(*dims)
        flags = (4 if dims_c else 0) | (16 if p2p else 0)
        n = self.lib.p3dfft_b200_plan_aux_steps(d, nx, ny, nz, rank, nxc or nx, nyc or ny, nzc or nz, flags, w,
                                                4 if self.single else 8, arr, 8)
        if n < 0:
            self._check()
            raise RuntimeError("plan_aux_steps failed")
        return [arr[i] for i in range(n)]

    def plan_rtran_info(self, dims, nx, ny, nz, rank, which, dims_c=False):
        """(dstart, dend, dsize) of the destination array and the work-buffer bound (complex elements)."""
        d = (C.c_int * 2)(*dims)
        o = (C.c_int * 9)()
        w = self.lib.p3dfft_b200_plan_rtran_info(d, nx, ny, nz, rank, RTRAN_NAMES.index(which), 4 if dims_c else 0, o)
        if w < 0:
            self._check()
            raise RuntimeError("plan_rtran_info failed")
        return list(o[0:3]), list(o[3:6]), list(o[6:9]), int(w)

    def plan_proc_parts(self, dims, nx, ny, nz, base, size, conf, nxc=None, nyc=None, nzc=None, stride1=False,
                        dims_c=False):
        d = (C.c_int * 2)(*dims)
        P = dims[0] * dims[1]
        parts = (C.c_int * (7 * P))()
        ierr = C.c_int(0)
        flags = (2 if stride1 else 0) | (4 if dims_c else 0)
        n = self.lib.p3dfft_b200_plan_proc_parts(d, nx, ny, nz, nxc or nx, nyc or ny, nzc or nz, flags, *base, *size,
                                                 conf, parts, C.byref(ierr))
        if n < 0:
            self._check()
            raise RuntimeError("plan_proc_parts failed")
        return [list(parts[7 * i:7 * i + 7]) for i in range(P)], n, ierr.value

    def plan_proc_neighb(self, dims, base, orient, direction, dims_c=False):
        d = (C.c_int * 2)(*dims)
        return int(self.lib.p3dfft_b200_plan_proc_neighb(d, 4 if dims_c else 0, base, orient, direction))


_cache: dict = {}


def load(single: bool = False) -> P3DFFT:
    """Load (once) the double or single precision library."""
    if single not in _cache:
        _cache[single] = P3DFFT(single)
    return _cache[single]
