"""numpy interpreter of the C++ planner's step list (TEST CODE, CPU only).

The library exports the stage / exchange descriptors it would hand to the CUDA kernels and
to NCCL (p3dfft_b200_plan_steps).  This module executes such a step list with numpy for P
simulated ranks: gathers through the input segment lists, applies the 1D transform with the
oracle's FFT definitions, scatters through the output segment lists, and moves the byte
ranges of every exchange.  It lets the host-side logic (segment addressing, pruning maps,
buffer rotation, alltoallv tables) be verified on a machine without a GPU.
"""
import numpy as np
import scipy.fft as sfft

from p3dfft_b200 import BUF_A, BUF_B, BUF_C, BUF_USER_IN, BUF_USER_OUT


def _line_transform(X, kind, n):
    if kind == 0:
        return sfft.fft(X, axis=0)
    if kind == 1:
        return sfft.ifft(X, axis=0, norm="forward")
    if kind == 2:
        return sfft.fft(X.real, axis=0)
    if kind == 3:
        nxhp = n // 2 + 1
        return sfft.irfft(X[:nxhp], n=n, axis=0, norm="forward")
    if kind == 4:
        return sfft.dct(X.real, type=1, axis=0) + 1j * sfft.dct(X.imag, type=1, axis=0)
    if kind == 5:
        return sfft.dst(X.real, type=1, axis=0) + 1j * sfft.dst(X.imag, type=1, axis=0)
    if kind == 6:
        return X
    raise ValueError(kind)


def _side_indices(side, na, nb, nc):
    """-> (k_logical [cnt], addr [cnt,na,nb,nc], bufid [cnt]) concatenated over segments."""
    shift = side.logical - side.cnt
    ks, addrs, bufs = [], [], []
    a = np.arange(na, dtype=np.int64)[None, :, None, None]
    b = np.arange(nb, dtype=np.int64)[None, None, :, None]
    c = np.arange(nc, dtype=np.int64)[None, None, None, :]
    covered = np.zeros(side.cnt, dtype=int)
    for g in range(side.nseg):
        sg = side.seg[g]
        i = np.arange(sg.len, dtype=np.int64)
        s = sg.start + i
        covered[s] += 1
        k = np.where(s < side.h1, s, s + shift)
        ro = (i // sg.kw) * sg.psh + (i % sg.kw) * sg.ps if sg.kw > 1 else i * sg.ps
        ao = (a // sg.aw) * sg.sah + (a % sg.aw) * sg.sa if sg.aw > 1 else a * sg.sa
        bo = (b // sg.bw) * sg.sbh + (b % sg.bw) * sg.sb if sg.bw > 1 else b * sg.sb
        addr = sg.off + ro[:, None, None, None] + ao + bo + c * sg.sc
        ks.append(k)
        addrs.append(addr)
        bufs.append((sg.buf, sg.peer))
    assert np.all(covered == 1), "segments must cover every stored point exactly once"
    return ks, addrs, bufs


def run_stage(st, bufs, world=None):
    """bufs = this rank's buffers; world[r] = rank r's buffers (peer-to-peer plans store into them)."""
    na, nb, nc = st.na, st.nb, st.nc
    ks, addrs, binfo = _side_indices(st.inp, na, nb, nc)
    X = np.zeros((st.inp.logical, na, nb, nc), dtype=np.complex128)
    for k, addr, (bid, peer) in zip(ks, addrs, binfo):
        assert peer < 0, "stages read local memory only"
        X[k] = bufs[bid][addr]
    Y = _line_transform(X, st.kind, st.n) * st.scale
    ks, addrs, binfo = _side_indices(st.out, na, nb, nc)
    for k, addr, (bid, peer) in zip(ks, addrs, binfo):
        tgt = bufs[bid] if peer < 0 else world[peer][bid]
        flat = addr.ravel()
        assert len(np.unique(flat)) == flat.size, "output scatter writes an address twice"
        tgt[addr] = Y[k].real if st.kind == 3 else Y[k]


def run_world(plans, infos, inputs, backward, nv=1, check_hazards=True, allow_padding=False):
    """plans[r] = step list of rank r; inputs[r] = flat user input.  Returns flat outputs."""
    P = len(plans)
    out = []
    bufs = []
    for r in range(P):
        inf = infos[r]
        real_n = inf.nx * inf.jisize * inf.kjsize * nv
        cplx_n = inf.iisize * inf.jjsize * inf.nzc * nv
        w = int(inf.work_elems) * nv
        b = {BUF_A: np.full(w, np.nan + 0j), BUF_B: np.full(w, np.nan + 0j), BUF_C: np.full(w, np.nan + 0j)}
        b[BUF_USER_IN] = inputs[r]
        b[BUF_USER_OUT] = np.full(real_n, np.nan) if backward else np.full(cplx_n, np.nan + 0j)
        bufs.append(b)
    nsteps = len(plans[0])
    assert all(len(p) == nsteps for p in plans)
    for i in range(nsteps):
        if plans[0][i].is_exchange:
            # gather who is who: comm 0 = row (same jpid, ordered by ipid), comm 1 = column
            for r in range(P):
                ex = plans[r][i].ex
                me = infos[r]
                if ex.p2p:          # blocks were stored at their destination by the producing stage
                    continue
                for p in range(ex.npeer):
                    if p == ex.self:
                        continue
                    if ex.comm == 0:
                        peer = [q for q in range(P) if infos[q].jpid == me.jpid and infos[q].ipid == p][0]
                        my_idx = me.ipid
                    else:
                        peer = [q for q in range(P) if infos[q].ipid == me.ipid and infos[q].jpid == p][0]
                        my_idx = me.jpid
                    pex = plans[peer][i].ex
                    n = ex.sndcnt[p]
                    assert n == pex.rcvcnt[my_idx]
                    src = bufs[r][ex.sendbuf][ex.sndoff[p]:ex.sndoff[p] + n]
                    # blocked layouts pad every peer block in x to a multiple of W: those lanes are
                    # never written and never read as live lines (NaN would reach the user output)
                    assert allow_padding or not np.any(np.isnan(src)), "exchange sends unwritten data"
                    bufs[peer][pex.recvbuf][pex.rcvoff[my_idx]:pex.rcvoff[my_idx] + n] = src
        else:
            for r in range(P):
                st = plans[r][i].st
                if check_hazards:
                    ins = {st.inp.seg[g].buf for g in range(st.inp.nseg)}
                    outs = {st.out.seg[g].buf for g in range(st.out.nseg)}
                    assert not (ins & outs), "a stage must not write a buffer it reads"
                run_stage(st, bufs[r], bufs)
    for r in range(P):
        o = bufs[r][BUF_USER_OUT]
        assert not np.any(np.isnan(o)), "user output not fully written"
        out.append(o)
    return out
