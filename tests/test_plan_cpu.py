"""CPU tests of the C-ABI library's host side: symbols, decomposition arithmetic and the
stage/exchange plans (interpreted with numpy) against the oracle.  No GPU needed."""
import ctypes
import itertools
import os
import re

import numpy as np
import pytest

import p3dfft_b200 as pb
from oracle import p3dfft_oracle as po
from tests import plan_interp as pi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return pb.load(False)


@pytest.mark.parametrize("single", [False, True])
def test_library_exports_every_declared_symbol(single):
    """Every function declared in include/*.h is exported by the built library."""
    L = pb.P3DFFT(single)
    names = set()
    for hdr in ("p3dfft.h", "p3dfft_b200.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"^\s*(?:void|int|long long)\s+(\w+)\s*\(", text, flags=re.M):
            names.add(m.group(1))
    names = {n for n in names if not n.startswith("Cp3dfft") and not n.startswith("Cget") and not n.startswith("Cset")}
    assert {"p3dfft_setup", "p3dfft_get_dims", "p3dfft_ftran_r2c", "p3dfft_btran_c2r", "p3dfft_ftran_r2c_many",
            "p3dfft_btran_c2r_many", "p3dfft_cheby", "p3dfft_cheby_many", "p3dfft_clean", "get_timers",
            "set_timers"} <= names
    for n in sorted(names):
        assert hasattr(L.lib, n), f"{n} not exported"
    assert bool(L.lib.p3dfft_b200_build_flags() & 1) == single


def _c_prototypes():
    """name -> number of parameters, for every function declared in include/*.h"""
    protos = {}
    for hdr in ("p3dfft.h", "p3dfft_b200.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"^\s*(?:void|int|long long)\s+(\w+)\s*\(([^)]*)\)\s*;", text, flags=re.M | re.S):
            args = m.group(2).strip()
            protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def test_fortran_module_binds_the_exported_c_abi():
    """fortran/p3dfft.F90 cannot be compiled here (no Fortran compiler in the image), so its ISO_C_BINDING interfaces are
    checked as text: every bind(C,name=...) is a symbol the library exports, with as many dummy arguments as the C prototype
    in include/*.h has parameters; and the module makes public every name of the reference's public list (module.F90:178-186)."""
    text = open(os.path.join(ROOT, "fortran", "p3dfft.F90")).read()
    code = "\n".join(ln.split("!")[0].rstrip() for ln in text.splitlines() if not ln.lstrip().startswith("!"))
    code = re.sub(r"&\s*\n\s*&?", " ", code)
    L = pb.P3DFFT(False)
    protos = _c_prototypes()
    binds = re.findall(r"(?:subroutine|function)\s+(\w+)\s*\(([^)]*)\)\s*(?:result\s*\(\w+\)\s*)?bind\s*\(\s*C\s*,\s*name\s*=\s*'(\w+)'\s*\)", code, flags=re.I)
    assert len(binds) >= 19, len(binds)
    for fname, fargs, cname in binds:
        assert hasattr(L.lib, cname), f"{fname}: {cname} is not exported by the library"
        assert cname in protos, f"{cname} is not declared in include/*.h"
        nf = 0 if not fargs.strip() else fargs.count(",") + 1
        assert nf == protos[cname], f"{fname} -> {cname}: {nf} Fortran dummies, {protos[cname]} C parameters"
    bound = {c for _, _, c in binds}
    assert {"p3dfft_setup", "p3dfft_get_dims", "p3dfft_ftran_r2c", "p3dfft_btran_c2r", "p3dfft_ftran_r2c_many", "p3dfft_btran_c2r_many",
            "p3dfft_cheby", "p3dfft_cheby_many", "p3dfft_clean", "get_timers", "set_timers"} <= bound
    public = set()
    for m in re.finditer(r"public\b[^:\n]*::\s*([^\n]*)", code, flags=re.I):
        public |= {w.split("(")[0].split("=")[0].strip().lower() for w in re.sub(r"\([^)]*\)", "", m.group(1)).split(",")}
    for name in ("p3dfft_type", "r8", "i8", "num_thr", "padi", "timers", "real_size", "complex_size", "p3dfft_setup", "p3dfft_get_dims",
                 "p3dfft_get_mpi_info", "p3dfft_ftran_r2c", "p3dfft_btran_c2r", "p3dfft_ftran_r2c_many", "p3dfft_btran_c2r_many",
                 "p3dfft_cheby", "p3dfft_cheby_many", "get_timers", "set_timers", "p3dfft_clean", "print_buf", "print_buf_real",
                 "rtran_x2y", "rtran_y2x", "rtran_x2z", "rtran_z2x", "get_proc_parts", "p3dfft_ftran_r2c_1d", "proc_id2coords",
                 "proc_coords2id", "proc_dims", "proc_parts"):
        assert name in public, f"module p3dfft does not make `{name}` public"


GRIDS = [(1, 1), (2, 2), (1, 4), (4, 1), (2, 3), (3, 2), (2, 4), (1, 8)]
SIZES = [((32, 32, 32), None), ((14, 26, 38), None), ((64, 64, 64), (32, 32, 32)), ((128, 128, 128), None),
         ((32, 20, 12), (16, 10, 8)), ((16, 16, 33), None)]


@pytest.mark.parametrize("dims", GRIDS)
@pytest.mark.parametrize("n,cut", SIZES)
@pytest.mark.parametrize("dims_c", [False, True])
def test_decomp_matches_oracle(lib, dims, n, cut, dims_c):
    """C++ Decomp == Python restatement of setup.F90 over the makejob.py matrix."""
    nx, ny, nz = n
    c = cut or (None, None, None)
    for r in range(dims[0] * dims[1]):
        o = po.Decomp(nx, ny, nz, dims, r, *c, dims_c=dims_c)
        i = lib.plan_decomp(dims, nx, ny, nz, r, *c, dims_c=dims_c)
        for f in ("nxhp", "nxhpc", "nycph", "nzcph", "ipid", "jpid", "iistart", "iiend", "iisize", "jistart", "jiend",
                  "jisize", "jjstart", "jjend", "jjsize", "kjstart", "kjend", "kjsize", "padi_work", "padi", "nm"):
            assert getattr(i, f) == getattr(o, f), f
        assert tuple(i.memsize) == o.memsize


def test_decomp_errors(lib):
    with pytest.raises(RuntimeError, match="Invalid dimensions"):
        lib.plan_decomp((1, 1), 0, 4, 4)
    with pytest.raises(RuntimeError, match="transform length|prime"):
        lib.plan_steps((1, 1), 8, 8, 2 * 4099, 0, False, "fft")     # prime factor 4099 > 4096
    steps, _ = lib.plan_steps((1, 1), 8, 8, 74, 0, False, "fft")    # 74 = 2*37: the O(r^2) pass covers it
    assert list(steps[-1].st.fac)[: steps[-1].st.nfac] == [2, 37]
    with pytest.raises(RuntimeError, match="Unknown transform type"):
        lib.plan_steps((1, 1), 8, 8, 8, 0, False, "ffx")


def _run(lib, n, dims, cut, opf, opb, stride1=False, nv=1, dims_c=False):
    for plain, p2p, rb in ((False, False, 0), (True, False, 0), (False, True, 0), (False, True, 64)):
        _run_layout(lib, n, dims, cut, opf, opb, stride1, nv, dims_c, plain, p2p, rb)


def _run_layout(lib, n, dims, cut, opf, opb, stride1, nv, dims_c, plain, p2p=False, rb=0):
    nx, ny, nz = n
    c = cut or (None, None, None)
    P = dims[0] * dims[1]
    rng = np.random.default_rng(11)
    A = [np.asfortranarray(rng.random((nx, ny, nz))) for _ in range(nv)]
    ds = [po.Decomp(nx, ny, nz, dims, r, *c, stride1=stride1, dims_c=dims_c) for r in range(P)]
    plans, infos = [], []
    for r in range(P):
        s, inf = lib.plan_steps(dims, nx, ny, nz, r, False, opf, nv, *c, stride1=stride1, dims_c=dims_c, plain=plain, p2p=p2p, row_bytes=rb)
        plans.append(s)
        infos.append(inf)
    ins = [np.concatenate([a[po.local_in_slice(d)].ravel(order="F") for a in A]) for d in ds]
    outs = pi.run_world(plans, infos, ins, False, nv, allow_padding=not plain)
    Fg = [po.global_forward(a, ds[0], opf) for a in A]
    for d, o in zip(ds, outs):
        exp = np.concatenate([po.local_forward(a, d, opf).ravel(order="F") for a in A])
        assert po.rel_l2(o, exp) < 1e-13
    # backward from the oracle's spectrum
    plans, infos = [], []
    for r in range(P):
        s, inf = lib.plan_steps(dims, nx, ny, nz, r, True, opb, nv, *c, stride1=stride1, dims_c=dims_c, plain=plain, p2p=p2p, row_bytes=rb)
        plans.append(s)
        infos.append(inf)
    ins = []
    for d in ds:
        parts = []
        for f in Fg:
            loc = f[po.local_out_slice(d)]
            if stride1:
                loc = loc.transpose(2, 1, 0)
            parts.append(np.asfortranarray(loc).ravel(order="F"))
        ins.append(np.concatenate(parts))
    outs = pi.run_world(plans, infos, ins, True, nv, allow_padding=not plain)
    for d, o in zip(ds, outs):
        exp = np.concatenate([po.local_backward(f, d, opb).ravel(order="F") for f in Fg])
        assert po.rel_l2(o, exp) < 1e-13


@pytest.mark.parametrize("dims", [(1, 1), (2, 2), (1, 4), (4, 1), (2, 3)])
@pytest.mark.parametrize("n,cut", [((14, 26, 38), None), ((16, 12, 10), (8, 6, 6)), ((16, 16, 16), None)])
def test_plan_interpreted_matches_oracle(lib, dims, n, cut):
    _run(lib, n, dims, cut, "fft", "tff")


@pytest.mark.parametrize("ops", [("ffc", "cff"), ("ffs", "sff"), ("ffn", "nff")])
@pytest.mark.parametrize("dims", [(1, 1), (2, 2)])
def test_plan_third_dimension_variants(lib, dims, ops):
    _run(lib, (12, 10, 9), dims, None, *ops)
    _run(lib, (12, 10, 9), dims, (8, 6, 6), *ops)


@pytest.mark.parametrize("dims", [(1, 1), (2, 2), (1, 2), (2, 1)])
def test_plan_stride1_and_many_and_dims_c(lib, dims):
    _run(lib, (12, 10, 14), dims, None, "fft", "tff", stride1=True)
    _run(lib, (12, 10, 14), dims, (8, 6, 8), "fft", "tff", stride1=True, nv=2)
    _run(lib, (12, 10, 14), dims, None, "fft", "tff", nv=3)
    _run(lib, (12, 10, 14), dims, None, "fft", "tff", dims_c=True)


def test_exchange_tables_equal_reference_byte_counts(lib):
    """Per-peer counts equal the If/Kf/Jr/Kr tables of setup.F90:481-518 (bytes / 16)."""
    n, dims = (128, 128, 128), (2, 2)
    for r in range(4):
        d = po.Decomp(*n, dims, r)
        steps, _ = lib.plan_steps(dims, *n, r, False, "fft", plain=True)
        ex = [s.ex for s in steps if s.is_exchange]
        assert [e.comm for e in ex] == [0, 1]
        assert list(ex[0].sndcnt[:2]) == [c // 16 for c in d.IfSndCnts]
        assert list(ex[0].sndoff[:2]) == [c // 16 for c in d.IfSndStrt]
        assert list(ex[0].rcvcnt[:2]) == [c // 16 for c in d.IfRcvCnts]
        assert list(ex[0].rcvoff[:2]) == [c // 16 for c in d.IfRcvStrt]
        assert list(ex[1].sndcnt[:2]) == [c // 16 for c in d.KfSndCnts]
        assert list(ex[1].rcvoff[:2]) == [c // 16 for c in d.KfRcvStrt]
        steps, _ = lib.plan_steps(dims, *n, r, True, "tff", plain=True)
        ex = [s.ex for s in steps if s.is_exchange]
        assert [e.comm for e in ex] == [1, 0]
        assert list(ex[0].sndcnt[:2]) == [c // 16 for c in d.JrSndCnts]
        assert list(ex[0].rcvoff[:2]) == [c // 16 for c in d.JrRcvStrt]
        assert list(ex[1].sndcnt[:2]) == [c // 16 for c in d.KrSndCnts]
        assert list(ex[1].rcvoff[:2]) == [c // 16 for c in d.KrRcvStrt]


@pytest.mark.parametrize("single", [False, True])
def test_tile_row_width_rule(single):
    """pick_W (plan.h): 128-byte rows wherever a kernel exists for the 128-byte tile of the longest Y/Z transform -- any length
    up to 1024, 1280, 1536 and, through the split kernel, 2048 (nz = 1025: the Chebyshev transform's 2048-point even
    extension; nz = 1023: the sine transform's odd extension) -- and 64-byte rows otherwise."""
    L = pb.load(single)
    full = 16 if single else 8            # lines per 128-byte row

    def width(ny, nz, op="fft"):
        steps, _ = L.plan_steps((1, 1), 64, ny, nz, 0, False, op)
        z = [s.st for s in steps if not s.is_exchange][-1]
        return z.inp.seg[0].aw

    assert width(64, 1024) == full and width(1024, 64) == full
    assert width(1280, 64) == full and width(64, 1536) == full
    assert width(64, 2048) == full and width(2048, 2048) == full
    assert width(64, 1025, "ffc") == full
    assert width(64, 1023, "ffs") == full
    assert width(64, 4096) == full // 2 and width(1792, 64) == full // 2
