"""CPU tests of the remaining module routines: the real-data pencil transposes
rtran_x2y / y2x / x2z / z2x (build/module.F90:1061-1361), p3dfft_ftran_r2c_1d (build/ftran.F90:787) and the
process-map queries (build/module.F90:788-1054), all through the C-ABI library's host side.

The transposes run as P3D_RCOPY stages (csrc/rcopy.h).  Their address arithmetic is the same host/device
code the CUDA kernel executes; here it is driven over numpy buffers by the test-only harness
tests/c/rcopy_host.cpp for P simulated ranks and compared with the oracle's restatement of the reference's
pack -> alltoallv -> unpack loops."""
import ctypes as C
import os

import numpy as np
import pytest

import p3dfft_b200 as pb
from oracle import p3dfft_oracle as po
from tests import plan_interp as pi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return pb.load(False)


@pytest.fixture(scope="module")
def harness():
    h = C.CDLL(os.path.join(ROOT, "tests", "emu", "lib", "librcopy_check.so"))
    h.rcopy_host_run.argtypes = [C.POINTER(pb.Stage), C.c_int]
    h.rcopy_host_contiguous_boxes.argtypes = [C.POINTER(pb.Stage), C.c_int]
    return h


def _resolve(st, bufs, world, esz):
    """what api.cpp does before a launch: seg.base = buffer address + element offset."""
    for side in (st.inp, st.out):
        for g in range(side.nseg):
            sg = side.seg[g]
            tgt = bufs if sg.peer < 0 else world[sg.peer]
            sg.base = tgt[sg.buf].ctypes.data + sg.off * esz


def _run_rtran_world(L, harness, n, dims, which, p2p, dtype=np.float64, dims_c=False):
    nx, ny, nz = n
    P = dims[0] * dims[1]
    esz = np.dtype(dtype).itemsize
    rng = np.random.default_rng(11)
    G = np.asfortranarray(rng.random(n).astype(dtype))
    D = [po.Decomp(nx, ny, nz, dims, r, dims_c=dims_c, elem=esz) for r in range(P)]
    plans = [L.plan_aux_steps(dims, nx, ny, nz, r, which, p2p=p2p, dims_c=dims_c) for r in range(P)]
    world = []
    for r in range(P):
        src_sl, dst_sl = po.rtran_slices(D[r], which)
        _, _, dsize, welems = L.plan_rtran_info(dims, nx, ny, nz, r, which, dims_c=dims_c)
        assert dsize == po.rtran_dims(D[r], which)[2]
        b = {pb.BUF_USER_IN: np.ascontiguousarray(G[src_sl].ravel(order="F")),
             pb.BUF_USER_OUT: np.full(int(np.prod(dsize)), np.nan, dtype=dtype),
             # sized by the library's own bound, in reals (the bound is in complex elements)
             pb.BUF_A: np.full(2 * welems, np.nan, dtype=dtype), pb.BUF_B: np.full(2 * welems, np.nan, dtype=dtype)}
        world.append(b)
    nsteps = len(plans[0])
    assert all(sum(s.is_exchange for s in p) == sum(s.is_exchange for s in plans[0]) for p in plans)
    # ranks whose stage is empty have shorter lists: walk stage-by-stage between exchanges
    pos = [0] * P
    while any(pos[r] < len(plans[r]) for r in range(P)):
        for r in range(P):          # all stages up to the next exchange
            while pos[r] < len(plans[r]) and not plans[r][pos[r]].is_exchange:
                st = plans[r][pos[r]].st
                assert st.kind == 7
                ks, addrs, _ = pi._side_indices(st.inp, st.na, st.nb, st.nc)      # every stored point covered once
                ks, addrs, _ = pi._side_indices(st.out, st.na, st.nb, st.nc)
                _resolve(st, world[r], world, esz)
                nb = harness.rcopy_host_run(C.byref(st), esz)
                assert nb >= 1
                assert harness.rcopy_host_contiguous_boxes(C.byref(st), esz) == nb, "rows must be contiguous on both sides"
                pos[r] += 1
        exs = [plans[r][pos[r]].ex if pos[r] < len(plans[r]) else None for r in range(P)]
        if all(e is None for e in exs):
            break
        assert all(e is not None for e in exs), "every rank reaches every exchange"
        for r in range(P):
            ex = exs[r]
            assert ex.ebytes == esz and bool(ex.p2p) == p2p
            if ex.p2p:
                continue
            me = D[r]
            for p in range(ex.npeer):
                if p == ex.self:
                    continue
                peer = me.rank_of(p, me.jpid) if ex.comm == 0 else me.rank_of(me.ipid, p)
                my_idx = me.ipid if ex.comm == 0 else me.jpid
                pex = exs[peer]
                cnt = ex.sndcnt[p]
                assert cnt == pex.rcvcnt[my_idx]
                src = world[r][ex.sendbuf][ex.sndoff[p]:ex.sndoff[p] + cnt]
                assert not np.any(np.isnan(src)), "exchange sends unwritten data"
                world[peer][pex.recvbuf][pex.rcvoff[my_idx]:pex.rcvoff[my_idx] + cnt] = src
        # the reference's byte tables (setup.F90:531-549) are what the exchange moves
        for r in range(P):
            ex, me = exs[r], D[r]
            row = ex.comm == 0
            fromx = which in ("x2y", "x2z")
            xs, xc = (me.IiStrt, me.IiCnts) if row else (me.IjStrt, me.IjCnts)
            fs, fc = (me.JiStrt, me.JiCnts) if row else (me.KjStrt, me.KjCnts)
            ss, sc, rs, rc = (xs, xc, fs, fc) if fromx else (fs, fc, xs, xc)
            for p in range(ex.npeer):
                assert (ex.sndoff[p] * esz, ex.sndcnt[p] * esz) == (ss[p], sc[p])
                assert (ex.rcvoff[p] * esz, ex.rcvcnt[p] * esz) == (rs[p], rc[p])
        for r in range(P):
            pos[r] += 1
    del nsteps
    # against the global definition and the structural restatement of the reference's loops
    sim = po.SimWorld(nx, ny, nz, dims, dtype=dtype, dims_c=dims_c)
    ref = sim.rtran(which, [np.asfortranarray(G[po.rtran_slices(d, which)[0]]) for d in D])
    for r in range(P):
        out = world[r][pb.BUF_USER_OUT]
        assert not np.any(np.isnan(out)), "destination not fully written"
        exp = po.rtran_local(G, D[r], which)
        assert np.array_equal(out, exp.ravel(order="F"))           # bit-exact: data movement only
        assert np.array_equal(ref[r], exp)


GRIDS = [(1, 1), (1, 2), (2, 1), (2, 2), (1, 4), (4, 1), (2, 3), (3, 2), (2, 4), (1, 8)]
SIZES = [(16, 12, 10), (14, 26, 38), (32, 32, 32), (9, 7, 5)]


@pytest.mark.parametrize("which", pb.RTRAN_NAMES)
@pytest.mark.parametrize("dims", GRIDS)
@pytest.mark.parametrize("n", SIZES)
def test_rtran_plans_move_the_reference_blocks(lib, harness, which, dims, n):
    if min(n[0], n[1]) < dims[0] or min(n[0], n[2]) < dims[1]:
        pytest.skip("grid larger than the array")
    for p2p in (False, True):
        _run_rtran_world(lib, harness, n, dims, which, p2p)


@pytest.mark.parametrize("which", pb.RTRAN_NAMES)
def test_rtran_plans_single_precision_and_dims_c(harness, which):
    Lf = pb.load(True)
    _run_rtran_world(Lf, harness, (16, 12, 10), (2, 2), which, True, dtype=np.float32)
    _run_rtran_world(Lf, harness, (16, 12, 10), (2, 3), which, False, dtype=np.float32, dims_c=True)


def test_rtran_work_bound_covers_every_rank(lib):
    """the lazily grown work buffers (api.cpp alloc_work) use one bound for all ranks"""
    for dims, n in (((2, 3), (14, 26, 38)), ((3, 2), (9, 7, 5)), ((1, 8), (32, 32, 32))):
        P = dims[0] * dims[1]
        bounds = set()
        for r in range(P):
            d = po.Decomp(*n, dims, r)
            w = lib.plan_rtran_info(dims, *n, r, "x2y")[3]
            bounds.add(w)
            need = max(d.nx * d.jisize * d.kjsize, d.iiisize * d.ny * d.kjsize, d.ijsize * d.jisize * d.nz)
            assert 2 * w >= need
        assert len(bounds) == 1


@pytest.mark.parametrize("n", [(16, 12, 10), (64, 8, 4), (14, 6, 5), (30, 4, 4)])
@pytest.mark.parametrize("dims", [(1, 1), (2, 2), (2, 3)])
def test_r2c_1d_plan(lib, n, dims):
    """p3dfft_ftran_r2c_1d = the X stage alone into the user array, nx+2 reals per line (ftran.F90:790-791)."""
    nx, ny, nz = n
    rng = np.random.default_rng(5)
    G = np.asfortranarray(rng.random(n))
    for r in range(dims[0] * dims[1]):
        d = po.Decomp(nx, ny, nz, dims, r)
        steps = lib.plan_aux_steps(dims, nx, ny, nz, r, "r2c_1d")
        assert len(steps) == 1 and not steps[0].is_exchange
        st = steps[0].st
        assert (st.kind, st.n, st.na, st.nb, st.timer) == (2, nx, d.jisize, d.kjsize, 5)
        A = np.asfortranarray(G[po.local_in_slice(d)])
        bufs = {pb.BUF_USER_IN: A.ravel(order="F").astype(np.complex128),
                pb.BUF_USER_OUT: np.full(d.nxhp * d.jisize * d.kjsize, np.nan + 0j)}
        pi.run_stage(st, bufs)
        exp = po.forward_r2c_1d(A)
        assert po.rel_l2(bufs[pb.BUF_USER_OUT], exp.ravel(order="F")) < 1e-14


# ---- process-map queries ---------------------------------------------------------------------------
@pytest.mark.parametrize("dims", [(1, 1), (2, 2), (2, 3), (3, 2), (1, 4), (2, 4)])
@pytest.mark.parametrize("dims_c", [False, True])
def test_proc_neighb(lib, dims, dims_c):
    g = po.ProcGrid(16, 12, 10, dims, dims_c=dims_c)
    P = dims[0] * dims[1]
    for base in range(-1, P + 1):
        for orient in (-1, 1, 2):
            for direction in (0, 1, 2):
                assert lib.plan_proc_neighb(dims, base, orient, direction, dims_c=dims_c) == \
                    g.proc_neighb(base, orient, direction)


@pytest.mark.parametrize("dims", [(1, 1), (2, 2), (2, 3), (3, 2), (2, 4)])
@pytest.mark.parametrize("stride1", [False, True])
def test_get_proc_parts_matches_restatement(lib, dims, stride1):
    """every box of a small sweep, both decompositions; the reference's quirks included (procmap.h)."""
    n, cut = (16, 12, 10), (12, 8, 6)
    g = po.ProcGrid(*n, dims, *cut, stride1=stride1)
    P = dims[0] * dims[1]
    rng = np.random.default_rng(2)
    cases = [((1, 1, 1), n, 1), ((1, 1, 1), (cut[0] // 2 + 1, cut[1], cut[2]), 2), ((1, 1, 1), (1, 1, 1), 3)]
    for _ in range(60):
        conf = int(rng.integers(1, 3))
        ext = n if conf == 1 else (cut[0] // 2 + 1, cut[1], cut[2])
        if conf == 2 and stride1:
            ext = ext[::-1]
        base = [int(rng.integers(1, e + 1)) for e in ext]
        size = [int(rng.integers(1, e - b + 2)) for e, b in zip(ext, base)]
        cases.append((tuple(base), tuple(size), conf))
    cases.append(((1, 40, 1), (1, 1, 1), 1))           # base point outside every block -> ierr -1
    for base, size, conf in cases:
        exp = g.get_proc_parts(*base, *size, conf)
        got = lib.plan_proc_parts(dims, *n, base, size, conf, *cut, stride1=stride1)
        assert got == (exp[0][:P], exp[1], exp[2]), (base, size, conf)


def test_get_proc_parts_tiles_single_direction_boxes(lib):
    """boxes that span ranks in ONE grid direction are where the reference routine is sound: the parts must
    tile the box exactly (conf 1: y over iproc, z over jproc)."""
    dims, n = (2, 3), (16, 12, 10)
    for base, size in (((1, 1, 2), (16, 12, 2)), ((3, 2, 1), (5, 3, 10)), ((1, 7, 4), (16, 6, 7))):
        parts, cnt, ierr = lib.plan_proc_parts(dims, *n, base, size, 1)
        assert ierr == 0
        live = parts[:cnt]
        if any(p[2] < 0 for p in live):        # j-continuation rows lack the y base (reference omission)
            live = [[p[0], p[1], base[1], p[3], p[4], p[5], p[6]] for p in live]
        vol = sum(p[4] * p[5] * p[6] for p in live)
        assert vol == size[0] * size[1] * size[2]
        for p in live:
            d = po.Decomp(*n, dims, p[0])
            assert d.jistart <= p[2] and p[2] + p[5] - 1 <= d.jiend
            assert d.kjstart <= p[3] and p[3] + p[6] - 1 <= d.kjend
