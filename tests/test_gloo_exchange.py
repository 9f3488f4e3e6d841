"""world_size-2 test of the N>1 host path on CPU: two real processes (torch.distributed, gloo) each take
THEIR rank's step list from the C-ABI planner, run the stage steps with the numpy interpreter and carry
out every exchange step with point-to-point messages sized by the plan's per-peer offsets and counts --
the same tables the library hands to ncclSend/ncclRecv.  Checks each rank's pencil against the oracle."""
import os
import socket

import numpy as np
import pytest

import p3dfft_b200 as pb
from oracle import p3dfft_oracle as po
from tests import plan_interp as pi


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _exchange(dist, torch, ex, bufs, my_idx, group_ranks):
    """alltoallv of one P3dExchange over gloo; self block was written in place by the producing stage."""
    reqs, landing = [], []
    for p in range(ex.npeer):
        if p == ex.self:
            continue
        peer = group_ranks[p]
        n = ex.sndcnt[p]
        src = np.ascontiguousarray(bufs[ex.sendbuf][ex.sndoff[p]:ex.sndoff[p] + n])
        src = np.nan_to_num(src, nan=0.0)                      # padding lanes of the blocked layouts
        reqs.append(dist.isend(torch.from_numpy(src.view(np.float64)), peer))
        r = torch.empty(2 * ex.rcvcnt[p], dtype=torch.float64)
        reqs.append(dist.irecv(r, peer))
        landing.append((p, r))
    for q in reqs:
        q.wait()
    for p, r in landing:
        bufs[ex.recvbuf][ex.rcvoff[p]:ex.rcvoff[p] + ex.rcvcnt[p]] = r.numpy().view(np.complex128)


def _worker(rank, world, port, dims, n, cut, plain, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = pb.load(False)
        nx, ny, nz = n
        c = cut or (None, None, None)
        d = po.Decomp(nx, ny, nz, dims, rank, *c)
        A = po.philox_field(nx, ny, nz)
        Fg = po.global_forward(A, d, "fft")
        errs = []
        for backward, op, inp in ((False, "fft", np.asfortranarray(A[po.local_in_slice(d)]).ravel(order="F")),
                                  (True, "tff", np.asfortranarray(Fg[po.local_out_slice(d)]).ravel(order="F"))):
            steps, inf = lib.plan_steps(dims, nx, ny, nz, rank, backward, op, 1, *c, plain=plain)
            w = int(inf.work_elems)
            bufs = {pb.BUF_A: np.full(w, np.nan + 0j), pb.BUF_B: np.full(w, np.nan + 0j), pb.BUF_C: np.full(w, np.nan + 0j),
                    pb.BUF_USER_IN: inp,
                    pb.BUF_USER_OUT: np.full(nx * inf.jisize * inf.kjsize, np.nan) if backward
                    else np.full(inf.iisize * inf.jjsize * inf.nzc, np.nan + 0j)}
            for s in steps:
                if s.is_exchange:
                    ex = s.ex
                    if ex.comm == 0:     # row communicator: same jpid, ordered by ipid (setup.F90:245-261)
                        group = [d.rank_of(ip, d.jpid) for ip in range(d.iproc)]
                    else:
                        group = [d.rank_of(d.ipid, jp) for jp in range(d.jproc)]
                    _exchange(dist, torch, ex, bufs, ex.self, group)
                else:
                    pi.run_stage(s.st, bufs)
            out = bufs[pb.BUF_USER_OUT]
            exp = po.local_backward(Fg, d, op) if backward else po.local_forward(A, d, op)
            errs.append(po.rel_l2(out, np.asfortranarray(exp).ravel(order="F")))
        q.put((rank, max(errs)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dims,n,cut,plain", [((1, 2), (12, 10, 14), None, False), ((2, 1), (14, 26, 38), None, False),
                                              ((1, 2), (16, 12, 10), (8, 6, 6), False), ((2, 1), (16, 16, 16), None, True)])
def test_two_rank_transform_over_gloo(dims, n, cut, plain):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dims, n, cut, plain, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err < 1e-13, (rank, err)


def _rtran_worker(rank, world, port, dims, n, q):
    """real-data transposes on two real processes: RCOPY stages through the test harness (tests/c/rcopy_host.cpp, the
    address arithmetic the CUDA kernel runs), exchanges over gloo with the plan's Ii/Ji/Ij/Kj counts in REAL elements."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = pb.load(False)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        h = C.CDLL(os.path.join(root, "tests", "emu", "lib", "librcopy_check.so"))
        h.rcopy_host_run.argtypes = [C.POINTER(pb.Stage), C.c_int]
        nx, ny, nz = n
        d = po.Decomp(nx, ny, nz, dims, rank)
        G = po.philox_field(nx, ny, nz, seed=5)
        bad = 0
        for which in pb.RTRAN_NAMES:
            src_sl, dst_sl = po.rtran_slices(d, which)
            _, _, dsize, welems = lib.plan_rtran_info(dims, nx, ny, nz, rank, which)
            bufs = {pb.BUF_USER_IN: np.ascontiguousarray(G[src_sl].ravel(order="F")),
                    pb.BUF_USER_OUT: np.full(int(np.prod(dsize)), np.nan),
                    pb.BUF_A: np.full(2 * welems, np.nan), pb.BUF_B: np.full(2 * welems, np.nan)}
            for s in lib.plan_aux_steps(dims, nx, ny, nz, rank, which):
                if s.is_exchange:
                    ex = s.ex
                    assert ex.ebytes == 8 and not ex.p2p
                    group = [d.rank_of(ip, d.jpid) for ip in range(d.iproc)] if ex.comm == 0 else \
                        [d.rank_of(d.ipid, jp) for jp in range(d.jproc)]
                    reqs, landing = [], []
                    for p in range(ex.npeer):
                        if p == ex.self:
                            continue
                        snd = np.ascontiguousarray(bufs[ex.sendbuf][ex.sndoff[p]:ex.sndoff[p] + ex.sndcnt[p]])
                        assert not np.any(np.isnan(snd))
                        reqs.append(dist.isend(torch.from_numpy(snd), group[p]))
                        r = torch.empty(ex.rcvcnt[p], dtype=torch.float64)
                        reqs.append(dist.irecv(r, group[p]))
                        landing.append((p, r))
                    for rq in reqs:
                        rq.wait()
                    for p, r in landing:
                        bufs[ex.recvbuf][ex.rcvoff[p]:ex.rcvoff[p] + ex.rcvcnt[p]] = r.numpy()
                else:
                    st = s.st
                    for side in (st.inp, st.out):
                        for g in range(side.nseg):
                            sg = side.seg[g]
                            assert sg.peer < 0
                            sg.base = bufs[sg.buf].ctypes.data + sg.off * 8
                    assert h.rcopy_host_run(C.byref(st), 8) >= 1
            exp = po.rtran_local(G, d, which).ravel(order="F")
            bad += int(not np.array_equal(bufs[pb.BUF_USER_OUT], exp))
        q.put((rank, bad))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dims,n", [((1, 2), (12, 10, 14)), ((2, 1), (14, 26, 38)), ((2, 1), (9, 7, 5))])
def test_two_rank_rtran_over_gloo(dims, n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rtran_worker, args=(r, 2, port, dims, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, bad in res:
        assert bad == 0, (rank, bad)
