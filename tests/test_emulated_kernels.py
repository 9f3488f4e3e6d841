"""The specialised CUDA stage kernels, executed on the CPU.

tests/emu compiles the product's own kernel source (p3dfft_b200/csrc/fft_fast.cuh) and its host-side dispatch
(fft_fast.cu) with g++ against a small emulation of the CUDA execution model: one OS thread per CUDA thread, a
barrier for __syncthreads(), CTAs one after the other.  These tests run whole transforms stage by stage through the
emulated kernels (plans from the C-ABI planner, buffers in numpy) and compare with the oracle -- the kernels'
butterflies, twiddle tables, digit reversal, shared-memory swizzles, row tables, pruning maps, the r2c/c2r pair
passes and the tile loops are all exercised without a GPU.  What the emulation cannot show: performance, and
races that need real warp scheduling to appear."""
import ctypes as C
import os

import numpy as np
import pytest

import p3dfft_b200 as pb
from oracle import p3dfft_oracle as po
from tests import plan_interp as pi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_libs = {}


def emu(single=False):
    if single not in _libs:
        h = C.CDLL(os.path.join(ROOT, "tests", "emu", "lib", "libemu_fast_single.so" if single else "libemu_fast.so"))
        h.emu_run_fast.argtypes = [C.POINTER(pb.Stage)]
        _libs[single] = h
    return _libs[single]


def _esz(kind, side, r):
    if kind == 7 or (kind == 2 and side == 0) or (kind == 3 and side == 1):
        return r
    return 2 * r


def run_steps_emulated(steps, bufs, world, single, counts):
    """one rank's stage steps up to (not including) the next exchange; returns the number of steps consumed"""
    r = 4 if single else 8
    n = 0
    for s in steps:
        if s.is_exchange:
            break
        st = s.st
        for si, side in enumerate((st.inp, st.out)):
            for g in range(side.nseg):
                sg = side.seg[g]
                tgt = bufs if sg.peer < 0 else world[sg.peer]
                sg.base = tgt[sg.buf].ctypes.data + sg.off * _esz(st.kind, si, r)
        rc = emu(single).emu_run_fast(C.byref(st))
        assert rc in (0, 1), rc
        if rc == 1:                       # no specialised kernel for this stage: the any-length kernel's job
            assert not single, "the numpy interpreter of the generic path works in double"
            pi.run_stage(st, bufs, world)
        counts[rc] += 1
        n += 1
    return n


def _eager_order(steps):
    """The producer-runs-ahead schedule the events of an X<->Y pipeline allow: every producer chunk as early as legal --
    without a ring all producers first, with the two-slot ring the producer of chunk c+2 right after the consumer of c."""
    out, i = [], 0
    while i < len(steps):
        if not (steps[i].pad_ >> 1) & 3:
            out.append(steps[i])
            i += 1
            continue
        j = i
        while j < len(steps) and (steps[j].pad_ >> 1) & 3:
            j += 1
        group = steps[i:j]
        prod, cons = group[0::2], group[1::2]
        assert all(not (p.pad_ & 1) for p in prod) and all(c.pad_ & 1 for c in cons)
        ring = ((group[0].pad_ >> 1) & 3) == 2
        if not ring:
            out += prod + cons
        else:
            out += prod[:2]
            for c in range(len(cons)):
                out.append(cons[c])
                if c + 2 < len(prod):
                    out.append(prod[c + 2])
        i = j
    return out


def transform_world(n, dims, cut, opf, opb, stride1=False, single=False, p2p=False, row_bytes=0, nv=1, overlap=0,
                    eager=False):
    """forward and backward on P simulated ranks; returns (fast stage count, generic stage count)"""
    nx, ny, nz = n
    c = cut or (None, None, None)
    P = dims[0] * dims[1]
    rt, ct = (np.float32, np.complex64) if single else (np.float64, np.complex128)
    L = pb.load(single)
    D = [po.Decomp(nx, ny, nz, dims, r, *c, stride1=stride1, elem=4 if single else 8) for r in range(P)]
    A = [np.asfortranarray(np.random.default_rng(7 + v).random(n).astype(rt)) for v in range(nv)]
    Fg = [po.global_forward(a.astype(np.float64), D[0], opf) for a in A]
    counts = [0, 0]
    tol = 2e-6 if single else 1e-13
    for backward, op in ((False, opf), (True, opb)):
        plans, world = [], []
        for r, d in enumerate(D):
            steps, inf = L.plan_steps(dims, nx, ny, nz, r, backward, op, nv, *c, stride1=stride1, p2p=p2p, row_bytes=row_bytes,
                                       overlap=overlap)
            if eager:
                steps = _eager_order(steps)
            plans.append(steps)
            w = int(inf.work_elems) * nv
            if backward:
                parts = []
                for f in Fg:
                    loc = f[po.local_out_slice(d)]
                    parts.append(np.asfortranarray(loc.transpose(2, 1, 0) if stride1 else loc).astype(ct).ravel(order="F"))
                inp = np.concatenate(parts)
                out = np.full(nx * d.jisize * d.kjsize * nv, np.nan, dtype=rt)
            else:
                inp = np.concatenate([np.asfortranarray(a[po.local_in_slice(d)]).ravel(order="F") for a in A])
                out = np.full(d.iisize * d.jjsize * d.nzc * nv, np.nan, dtype=ct)
            world.append({pb.BUF_A: np.zeros(w, dtype=ct), pb.BUF_B: np.zeros(w, dtype=ct), pb.BUF_C: np.zeros(w, dtype=ct),
                          pb.BUF_USER_IN: inp, pb.BUF_USER_OUT: out})
        pos = [0] * P
        while any(pos[r] < len(plans[r]) for r in range(P)):
            for r in range(P):
                pos[r] += run_steps_emulated(plans[r][pos[r]:], world[r], world, single, counts)
            exs = [plans[r][pos[r]].ex if pos[r] < len(plans[r]) else None for r in range(P)]
            if all(e is None for e in exs):
                break
            for r in range(P):
                ex, me = exs[r], D[r]
                if not ex.p2p:
                    for p in range(ex.npeer):
                        if p == ex.self:
                            continue
                        peer = me.rank_of(p, me.jpid) if ex.comm == 0 else me.rank_of(me.ipid, p)
                        my_idx = me.ipid if ex.comm == 0 else me.jpid
                        pex = exs[peer]
                        cnt = ex.sndcnt[p]
                        assert cnt == pex.rcvcnt[my_idx]
                        world[peer][pex.recvbuf][pex.rcvoff[my_idx]:pex.rcvoff[my_idx] + cnt] = \
                            world[r][ex.sendbuf][ex.sndoff[p]:ex.sndoff[p] + cnt]
                pos[r] += 1
        for r, d in enumerate(D):
            if backward:
                exp = np.concatenate([po.local_backward(f, d, op).ravel(order="F") for f in Fg])
            else:
                exp = np.concatenate([np.asfortranarray(po.local_forward(a.astype(np.float64), d, op)).ravel(order="F") for a in A])
            got = world[r][pb.BUF_USER_OUT]
            assert not np.any(np.isnan(got)), "output not fully written"
            assert po.rel_l2(got.astype(np.complex128 if not backward else np.float64), exp) <= tol, (n, dims, cut, op, r)
    return counts


# every specialised length: X stage H = 32 ... 1024 (nx = 64 ... 2048), Y/Z stages 64 ... 2048.  The long axis is paired with
# short ones to keep the emulation quick; stages shorter than 64 points belong to the any-length kernel (numpy here).
# (n, cut, number of stages the specialised kernels must take)
LENGTHS = [((64, 64, 64), None, 6), ((128, 256, 64), None, 6), ((256, 64, 128), None, 6), ((512, 16, 16), None, 2),
           ((1024, 16, 16), None, 2), ((2048, 16, 16), None, 2), ((16, 512, 16), None, 2), ((16, 16, 1024), None, 2),
           ((16, 1024, 16), None, 2), ((16, 16, 2048), None, 2), ((16, 2048, 16), None, 2),
           ((128, 128, 128), (64, 64, 64), 6), ((256, 128, 64), (170, 84, 42), 6), ((16, 1024, 16), (16, 680, 16), 2)]


@pytest.mark.parametrize("n,cut,nfast", LENGTHS)
def test_emulated_kernels_double_all_lengths(n, cut, nfast):
    fast, generic = transform_world(n, (1, 1), cut, "fft", "tff")
    assert (fast, generic) == (nfast, 6 - nfast)


@pytest.mark.parametrize("n,cut,nfast", [LENGTHS[0], LENGTHS[1], LENGTHS[3]])
def test_emulated_kernels_single_precision(n, cut, nfast):
    # (short stages would go to the generic kernel, whose numpy stand-in works in double: all-specialised cases and X only)
    if nfast < 6:
        n = (n[0], 64, 64)
    fast, generic = transform_world(n, (1, 1), cut, "fft", "tff", single=True)
    assert (fast, generic) == (6, 0)


@pytest.mark.parametrize("rb", [64, 128])
@pytest.mark.parametrize("n,cut", [((64, 64, 64), None), ((64, 256, 64), None), ((128, 64, 64), (84, 42, 42))])
def test_emulated_kernels_row_widths(n, cut, rb):
    """both tile row widths of the internal layouts: 64-byte rows use the swizzled shared-memory tile"""
    fast, generic = transform_world(n, (1, 1), cut, "fft", "tff", row_bytes=rb)
    assert (fast, generic) == (6, 0)


@pytest.mark.parametrize("n,cut", [((64, 64, 33), None), ((64, 64, 65), None), ((128, 64, 129), (64, 32, 65))])
@pytest.mark.parametrize("stride1", [False, True])
def test_emulated_kernels_chebyshev_and_stride1(n, cut, stride1):
    """DCT-I as an even-extended FFT (mirror rows) through the c2c kernel; STRIDE1 output layout"""
    fast, generic = transform_world(n, (1, 1), cut, "ffc", "cff", stride1=stride1)
    assert (fast, generic) == (6, 0)


@pytest.mark.parametrize("n,cut,single,stride1", [((64, 64, 31), None, False, False), ((64, 64, 63), None, False, True),
                                                  ((128, 64, 127), (64, 32, 64), False, False), ((64, 64, 63), None, True, False),
                                                  ((64, 64, 191), None, False, False), ((64, 64, 1023), None, False, False)])
def test_emulated_kernels_sine_transform(n, cut, single, stride1):
    """DST-I (op letter 's', exec_strans_r2_complex_same, fft_exec.F90:866-921) as a 2 (nz + 1)-point FFT of the odd extension
    through the c2c kernel's DST instantiation: nfft = 64 ... 1024 and 384 = 3.128, pruned z, STRIDE1, single precision;
    nz = 1023 (nfft = 2048) through the DST instantiation of the split kernel (128-byte rows)."""
    fast, generic = transform_world(n, (1, 1), cut, "ffs", "sff", single=single, stride1=stride1)
    assert (fast, generic) == (6, 0)


def test_emulated_kernels_sine_transform_multi_rank():
    fast, generic = transform_world((64, 64, 63), (2, 2), None, "ffs", "sff", p2p=True)
    assert (fast, generic) == (24, 0)


@pytest.mark.parametrize("n,cut,single,stride1,ops", [((16, 16, 2048), None, False, False, ("fft", "tff")),
                                                      ((64, 2048, 64), (64, 1364, 64), True, False, ("fft", "tff")),
                                                      ((16, 16, 2048), None, False, True, ("fft", "tff")),
                                                      ((16, 16, 1025), None, False, False, ("ffc", "cff"))])
def test_emulated_2048_points_with_128_byte_rows(n, cut, single, stride1, ops):
    """2048-point Y/Z stages take 128-byte rows through the split kernel (half of the 256 KB tile waits in registers): the
    planner's rule (pick_W), double and single, pruned, STRIDE1, and the DCT-I of 1025 points (2048-point even extension)"""
    L = pb.load(single)
    _, inf = L.plan_steps((1, 1), *n, 0, False, ops[0], 1, *(cut or n), stride1=stride1)
    st = [s.st for s in L.plan_steps((1, 1), *n, 0, False, ops[0], 1, *(cut or n), stride1=stride1)[0] if not s.is_exchange]
    assert st[-1].inp.seg[0].aw == (16 if single else 8)       # the Z stage reads 128-byte rows
    fast, generic = transform_world(n, (1, 1), cut, *ops, single=single, stride1=stride1)
    long_stages = sum(1 for v in n if v >= 64 and v != 1025) + (1 if n[2] == 1025 else 0)
    assert fast == 2 * long_stages and fast + generic == 6


def test_emulated_2048_points_multi_rank_peer_stores():
    fast, generic = transform_world((64, 2048, 16), (2, 2), None, "fft", "tff", p2p=True)
    assert (fast, generic) == (16, 8)       # per rank: X and Y stages specialised, the 16-point Z stages on the any-length kernel


def test_emulated_split_kernel(monkeypatch):
    """the two-half-tiles variant of the 1024-point c2c kernel (taken on the GPU for far-pitch inputs)"""
    monkeypatch.setenv("P3DFFT_B200_SPLIT", "1")
    # the environment is read once per process by the dispatch: use a fresh copy of the emulator library
    import shutil
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        dst = os.path.join(tmp, "libemu_split.so")
        shutil.copy(os.path.join(ROOT, "tests", "emu", "lib", "libemu_fast.so"), dst)
        h = C.CDLL(dst)
        h.emu_run_fast.argtypes = [C.POINTER(pb.Stage)]
        old = _libs.get(False)
        _libs[False] = h
        try:
            fast, generic = transform_world((16, 1024, 16), (1, 1), None, "fft", "tff")
            assert (fast, generic) == (2, 4)
            fast, generic = transform_world((16, 16, 1024), (1, 1), (16, 16, 512), "fft", "tff")
            assert (fast, generic) == (2, 4)
        finally:
            if old is not None:
                _libs[False] = old
            else:
                _libs.pop(False, None)


@pytest.mark.parametrize("dims", [(1, 2), (2, 1), (2, 2), (2, 4)])
@pytest.mark.parametrize("p2p", [False, True])
def test_emulated_kernels_multi_rank(dims, p2p):
    """P simulated ranks: per-peer blocks, and with p2p the stage kernels store into the peers' buffers"""
    fast, generic = transform_world((64, 64, 64), dims, None, "fft", "tff", p2p=p2p)
    assert generic == 0 and fast == 6 * dims[0] * dims[1]


def test_emulated_kernels_uneven_multi_rank_and_many():
    fast, generic = transform_world((64, 64, 64), (2, 3), (42, 42, 42), "fft", "tff", p2p=True)
    assert generic == 0
    fast, generic = transform_world((64, 64, 64), (2, 2), None, "fft", "tff", nv=2)
    assert generic == 0 and fast == 24


@pytest.mark.parametrize("single", [False, True])
def test_emulated_two_pass_variant(monkeypatch, single):
    """opt-in 1024 = 32 x 32 and 512 = 16 x 32 schedules (P3DFFT_B200_R32=1): radix-32 butterfly, their own twiddle blocks"""
    monkeypatch.setenv("P3DFFT_B200_R32", "1")
    h = emu(single)
    nx = 64 if single else 16                 # (single precision: every stage must be a specialised one, see above)
    for n, cut in (((nx, 1024, 64 if single else 16), None), ((nx, 64 if single else 16, 512), (nx, 64 if single else 16, 340))):
        fast, generic = transform_world(n, (1, 1), cut, "fft", "tff", single=single)
        assert fast == (6 if single else 2)
    # the variant really ran for the (512-point) Y stage of a transform; at 512 points it is off without the switch
    monkeypatch.delenv("P3DFFT_B200_R32")
    steps, _ = pb.load(single).plan_steps((1, 1), 64, 512, 64, 0, False, "fft")
    st = steps[1].st
    w = int(pb.load(single).plan_decomp((1, 1), 64, 512, 64).work_elems)
    ct = np.complex64 if single else np.complex128
    bufs = {b: np.zeros(w, dtype=ct) for b in (pb.BUF_A, pb.BUF_B, pb.BUF_C)}
    for si, side in enumerate((st.inp, st.out)):
        for g in range(side.nseg):
            sg = side.seg[g]
            sg.base = bufs[sg.buf].ctypes.data + sg.off * (8 if single else 16)
    assert h.emu_run_fast(C.byref(st)) == 0 and h.emu_last_variant() == 0
    monkeypatch.setenv("P3DFFT_B200_R32", "1")
    assert h.emu_run_fast(C.byref(st)) == 0 and h.emu_last_variant() == 1


@pytest.mark.parametrize("dims,n,cut", [((1, 2), (64, 64, 64), None), ((2, 2), (64, 64, 64), (42, 42, 42)), ((2, 4), (64, 64, 64), None),
                                        ((2, 3), (16, 12, 10), None), ((2, 2), (20, 12, 33), None)])
@pytest.mark.parametrize("chunks", [2, 3, 5])
def test_pipelined_tail_plans(dims, n, cut, chunks):
    """opt-in chunked tail of the peer-to-peer plans (plan.h split_for_overlap): P_c, barrier_c, Q_c ... must give the
    same transform when every consumer chunk runs right after its own barrier (buffers start zeroed: a consumer that
    needed a later producer chunk would read zeros).  Power-of-two cases run the emulated CUDA kernels, the others numpy."""
    L = pb.load(False)
    opf, opb = ("ffc", "cff") if n[2] % 2 else ("fft", "tff")
    for backward, op in ((False, opf), (True, opb)):
        steps, _ = L.plan_steps(dims, *n, 0, backward, op, 1, *(cut or (None, None, None)), p2p=True, overlap=chunks)
        base, _ = L.plan_steps(dims, *n, 0, backward, op, 1, *(cut or (None, None, None)), p2p=True)
        # the LAST  stage, exchange, stage  triple of the plan is pipelined (backward on a 1 x N grid: Z T3 Y, in front of X)
        ex = [s.is_exchange for s in base]
        at = max((i for i in range(len(base) - 2) if ex[i:i + 3] == [0, 1, 0]), default=None)
        if at is None:
            assert len(steps) == len(base)
            continue
        grp = steps[at:at + 3 * chunks]
        assert [s.is_exchange for s in grp] == [0, 1, 0] * chunks
        assert [s.pad_ & 1 for s in grp] == [0, 0, 1] * chunks                       # consumers run on the side stream
        assert [(s.pad_ >> 8) - 1 for s in grp] == [c for c in range(chunks) for _ in range(3)]
        assert len(steps) == len(base) - 3 + 3 * chunks
        assert [(s.pad_ >> 8) for s in steps[at + 3 * chunks:]] == [0] * (len(base) - at - 3)   # what follows is not chunked
        prod, cons = base[at].st, base[at + 2].st                                    # the chunks partition the batch axis
        if cons.kind == 3:        # X c2r consumer: z planes
            assert sum(s.st.nb for s in grp[0::3]) == prod.nb and sum(s.st.nb for s in grp[2::3]) == cons.nb
        else:
            assert sum(s.st.na for s in grp[0::3]) == prod.na and sum(s.st.na for s in grp[2::3]) == cons.na
    transform_world(n, dims, cut, opf, opb, p2p=True, overlap=chunks)


def test_pipelined_tail_needs_a_peer_to_peer_exchange():
    L = pb.load(False)
    a, _ = L.plan_steps((1, 1), 64, 64, 64, 0, False, "fft", overlap=4)
    b, _ = L.plan_steps((2, 2), 64, 64, 64, 0, False, "fft", overlap=4)          # not a p2p plan
    c, _ = L.plan_steps((2, 1), 64, 64, 64, 0, False, "fft", p2p=True, overlap=4)   # forward on M2 = 1: X T1 Y is pipelined (z planes), Z follows
    assert len(a) == 3 and len(b) == 5 and len(c) == 3 * 4 + 1
    assert [s.is_exchange for s in c] == [0, 1, 0] * 4 + [0] and [(s.pad_ >> 8) for s in c] == [1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 0]


@pytest.mark.parametrize("single", [False, True])
def test_emulated_bulk_store_variant(monkeypatch, single):
    """opt-in bulk-copy stores (P3DFFT_B200_BULK=1): the last pass re-forms the tile in shared memory in natural row order and
    one cp.async.bulk per output run moves it out (memcpy in the emulation) -- Y forward / Z and Y backward, pruned runs,
    2 x 2 with the runs landing in the peers' buffers"""
    monkeypatch.setenv("P3DFFT_B200_BULK", "1")
    h = emu(single)
    if single:
        transform_world((64, 1024, 64), (1, 1), (64, 680, 42), "fft", "tff", single=True)
    else:
        fast, generic = transform_world((16, 1024, 16), (1, 1), None, "fft", "tff")
        assert (fast, generic) == (2, 4)
        transform_world((24, 512, 16), (1, 1), (24, 340, 16), "fft", "tff")
        transform_world((16, 16, 1024), (1, 1), (16, 16, 680), "fft", "tff")
        transform_world((32, 1024, 16), (2, 2), (32, 680, 16), "fft", "tff", p2p=True)
        transform_world((32, 16, 512), (2, 2), None, "fft", "tff", p2p=True)
    # the forward Y stage (contiguous rows per run) takes the variant -- together with the two-pass schedule, which a 1024-point
    # stage with contiguous output rows runs by default (bits 0 and 2) --, the forward Z stage (user layout) takes neither
    steps, _ = pb.load(single).plan_steps((1, 1), 64, 1024, 64, 0, False, "fft")
    w = int(pb.load(single).plan_decomp((1, 1), 64, 1024, 64).work_elems)
    ct = np.complex64 if single else np.complex128
    bufs = {b: np.zeros(w, dtype=ct) for b in (pb.BUF_A, pb.BUF_B, pb.BUF_C)}
    bufs[pb.BUF_USER_OUT] = np.zeros(33 * 1024 * 64, dtype=ct)
    for idx, want in ((1, 5), (2, 0)):
        st = steps[idx].st
        for si, side in enumerate((st.inp, st.out)):
            for g in range(side.nseg):
                sg = side.seg[g]
                sg.base = bufs[sg.buf].ctypes.data + sg.off * (8 if single else 16)
        assert h.emu_run_fast(C.byref(st)) == 0 and h.emu_last_variant() == want


@pytest.mark.parametrize("single", [False, True])
def test_emulated_async_staged_kernel(monkeypatch, single):
    """asynchronously staged 1024-point kernel (P3DFFT_B200_ASYNC=1): cp.async input staging through six 256-row units,
    decimation in time by 4 on the outside, 16 x 16 sub-transforms -- Y and Z stages, forward and backward, pruned rows
    (zero-filled copies), the DCT-I mirror rows, STRIDE1 output, 2 x 2 with peer stores, CTAs that walk 1, 2, 3 and more tiles"""
    monkeypatch.setenv("P3DFFT_B200_ASYNC", "1")
    h = emu(single)
    if single:
        transform_world((64, 1024, 64), (1, 1), (64, 680, 42), "fft", "tff", single=True)
        transform_world((64, 64, 1024), (1, 1), None, "fft", "tff", single=True, stride1=True)
    else:
        fast, generic = transform_world((16, 1024, 16), (1, 1), None, "fft", "tff")
        assert (fast, generic) == (2, 4)
        transform_world((24, 1024, 16), (1, 1), (24, 680, 16), "fft", "tff")
        transform_world((16, 16, 1024), (1, 1), (16, 16, 680), "fft", "tff")
        transform_world((16, 16, 513), (1, 1), None, "ffc", "cff")
        transform_world((16, 40, 1024), (1, 1), None, "fft", "tff", stride1=True)
        transform_world((32, 1024, 16), (2, 2), (32, 680, 16), "fft", "tff", p2p=True)
        transform_world((48, 8, 1024), (1, 2), None, "fft", "tff", p2p=True, overlap=3)
    steps, _ = pb.load(single).plan_steps((1, 1), 64, 1024, 64, 0, False, "fft")
    w = int(pb.load(single).plan_decomp((1, 1), 64, 1024, 64).work_elems)
    ct = np.complex64 if single else np.complex128
    bufs = {b: np.zeros(w, dtype=ct) for b in (pb.BUF_A, pb.BUF_B, pb.BUF_C)}
    st = steps[1].st
    for si, side in enumerate((st.inp, st.out)):
        for g in range(side.nseg):
            sg = side.seg[g]
            sg.base = bufs[sg.buf].ctypes.data + sg.off * (8 if single else 16)
    assert h.emu_run_fast(C.byref(st)) == 0 and h.emu_last_variant() == 8


@pytest.mark.parametrize("single", [False, True])
def test_emulated_staged_x_stores(monkeypatch, single):
    """X r2c with whole-row stores through a staging buffer (the rule for stages that store into a peer's memory;
    P3DFFT_B200_XSTAGE=1 forces it everywhere): every specialised X length, pruned in x, partial tiles, 2 x 2 with peer stores"""
    monkeypatch.setenv("P3DFFT_B200_XSTAGE", "1")
    h = emu(single)
    for nx in (64, 128, 256, 512, 1024, 2048):
        fast, generic = transform_world((nx, 64, 64) if single else (nx, 16, 16), (1, 1), None, "fft", "tff", single=single)
        assert fast >= 2
    transform_world((256, 64, 64), (1, 1), (170, 64, 64), "fft", "tff", single=single)
    transform_world((128, 22, 18) if not single else (128, 64, 64), (1, 1), None, "fft", "tff", single=single)      # partial X tiles
    transform_world((256, 64, 64), (2, 2), (170, 42, 64), "fft", "tff", single=single, p2p=True)
    monkeypatch.delenv("P3DFFT_B200_XSTAGE")
    transform_world((128, 64, 64), (2, 2), None, "fft", "tff", single=single, p2p=True)      # by rule: peer outputs
    steps, _ = pb.load(single).plan_steps((2, 1), 128, 64, 64, 0, False, "fft", p2p=True)
    assert steps[0].st.kind == 2 and any(steps[0].st.out.seg[g].peer >= 0 for g in range(steps[0].st.out.nseg))


@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("n", [384, 768, 1536, 640, 1280])
def test_emulated_non_power_of_two_c2c(n, single):
    """3 * 2^k and 5 * 2^k lengths on the specialised c2c kernels (odd radix first: 384 = 3.16.8, 768 = 3.16.16,
    1536 = 6.16.16, 640 = 5.16.8, 1280 = 5.16.16): Y and Z stages, forward and backward, pruned, DCT-I of the matching odd
    length, 2 x 2 with peer stores"""
    nx = 64 if single else 16
    fast, generic = transform_world((nx, n, nx), (1, 1), None, "fft", "tff", single=single)
    assert fast >= 2
    fast, generic = transform_world((nx, nx, n), (1, 1), (nx, nx, 2 * (n // 3)), "fft", "tff", single=single)
    assert fast >= 2
    if n <= 768 and not single:
        transform_world((16, 16, n // 2 + 1), (1, 1), None, "ffc", "cff")          # DCT-I: nfft = 2 (nz - 1) = n
        transform_world((32, n, 16), (2, 2), None, "fft", "tff", p2p=True)


@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("nx", [384, 768, 1536, 640, 1280])
def test_emulated_non_power_of_two_x_stage(nx, single):
    """nx = 3 * 2^k, 5 * 2^k on the specialised X kernels (even radices in the pair passes, the odd factor in the middle:
    192 = 4.6.8, 384 = 8.6.8, 768 = 8.3.4.8, 320 = 8.5.8, 640 = 8.5.2.8): r2c and c2r, pruned in x, partial tiles, staged
    whole-row stores towards a peer"""
    ny = 64 if single else 16
    fast, generic = transform_world((nx, ny, ny), (1, 1), None, "fft", "tff", single=single)
    assert fast >= 2 and (not single or generic == 0)
    transform_world((nx, ny, ny), (1, 1), (2 * (nx // 3), ny, ny), "fft", "tff", single=single)
    if not single:
        transform_world((nx, 22, 18), (1, 1), None, "fft", "tff")                    # partial X tiles
        if nx <= 768:
            transform_world((nx, 32, 16), (2, 2), None, "fft", "tff", p2p=True)      # staged stores into the row peer
