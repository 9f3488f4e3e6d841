"""Several ranks of the whole C-ABI library on the CPU: P processes of tests/mp_emu.py, each loading libp3dfft_emu.so
(api.cpp + planner + every kernel on the mock CUDA runtime) with shared-memory "device" buffers, mock cudaIpc peer
mappings and a mock NCCL (tests/emu/emu_mp.inc).  What runs is the product's own multi-GPU host path: communicator
split, peer mapping, stage kernels storing into the peers' buffers, exchange steps as barriers or grouped send/recv,
the executor's write-after-read rule, buffer regrowth with re-mapping -- checked against the oracle with the GPU
driver's own checking code (tests/mp_parity.py).  A mismatch of collectives shows up as a time-out of the mock NCCL
(and of the launcher), not as a hang of the test-suite."""
import glob
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "lib", "libp3dfft_emu.so")

pytestmark = pytest.mark.skipif(not os.path.exists(EMU), reason="emulated library not built (python tests/emu/build.py)")


def launch(grid, args=(), env=None, timeout=600):
    m1, m2 = (int(x) for x in grid.split("x"))
    world = m1 * m2
    uid = (b"p3demu_nccl_%d_%s" % (os.getpid(), os.urandom(6).hex().encode())).ljust(128, b"\0")
    base = dict(os.environ)
    for k in list(base):
        if k.startswith("P3DFFT_B200_"):
            del base[k]
    base.update({"WORLD_SIZE": str(world), "P3D_EMU_UID": uid.hex(), "P3D_EMU_SHM": "1", "P3D_EMU_TIMEOUT": "90",
                 "OMP_NUM_THREADS": "1", "OPENBLAS_NUM_THREADS": "1", "MKL_NUM_THREADS": "1"})
    base.update(env or {})
    procs = []
    try:
        for r in range(world):
            e = dict(base, RANK=str(r))
            procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mp_emu.py"), "--grid", grid, *args], env=e,
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = []
        for p in procs:
            try:
                outs.append(p.communicate(timeout=timeout)[0])
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                outs.append(p.communicate()[0] + "\n[launcher: time-out]")
        return [p.returncode for p in procs], outs
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
            for f in glob.glob(f"/dev/shm/p3demu_mem_{p.pid}_*"):
                os.unlink(f)
        for d in glob.glob("/dev/shm/" + uid.rstrip(b"\0").decode() + "*"):
            shutil.rmtree(d, ignore_errors=True)


def check(grid, args=(), env=None):
    codes, outs = launch(grid, args, env)
    text = "\n".join(outs)
    assert all(c == 0 for c in codes), f"grid {grid} {args} {env}: exit codes {codes}\n{text[-6000:]}"
    assert "FAIL" not in text and "EXCEPTION" not in text, text[-6000:]
    return text


@pytest.mark.parametrize("grid", ["1x2", "2x1", "2x2"])
def test_peer_to_peer_transposes(grid):
    """default multi-GPU path: every transpose is stores into the peers' mapped buffers + a barrier; no ncclSend at all"""
    text = check(grid, ["--suite", "fast", "--expect-p2p", "1"])
    assert text.count(" ok") >= 4 * int(grid[0]) * int(grid[2])


@pytest.mark.parametrize("grid", ["1x2", "2x2"])
def test_nccl_send_recv_transposes(grid):
    """P3DFFT_B200_P2P=0: grouped ncclSend/ncclRecv exchanges sized by the plan"""
    check(grid, ["--suite", "fast", "--expect-p2p", "0"], {"P3DFFT_B200_P2P": "0"})


def test_reference_matrix_2x2():
    """uneven sizes, Chebyshev, no-op, STRIDE1, pruned single precision (the any-length kernels) on a 2x2 grid"""
    check("2x2", ["--suite", "mixed"])


def test_plain_layout_exchanges():
    """the reference's own pack-buffer layouts and alltoallv tables (P3DFFT_B200_PLAIN=1) through ncclSend/ncclRecv"""
    check("2x2", ["--suite", "mixed", "--expect-p2p", "0"], {"P3DFFT_B200_PLAIN": "1"})


@pytest.mark.parametrize("grid", ["1x2", "2x2"])
def test_real_transposes_and_queries(grid):
    check(grid, ["--suite", "none", "--aux"])


@pytest.mark.parametrize("grid,env", [("1x2", {"P3D_EMU_DELAY": "1:3:2:250"}), ("1x4", {"P3D_EMU_DELAY": "2:3:2:250"}),
                                      ("2x2", {"P3D_EMU_DELAY": "3:3:2:250"}), ("2x1", {"P3D_EMU_DELAY": "1:3:2:250", "P3DFFT_B200_OVERLAP": "3"}),
                                      ("1x2", {"P3DFFT_B200_SCOPED": "0", "P3D_EMU_DELAY": "1:3:2:250"}),
                                      ("2x2", {"P3DFFT_B200_SCOPED": "0"}),
                                      ("1x2", {"P3DFFT_B200_FLAGBAR": "0", "P3D_EMU_DELAY": "1:3:2:250"}),
                                      ("1x4", {"P3DFFT_B200_FLAGBAR": "0", "P3D_EMU_DELAY": "2:3:2:250"})])
def test_repeated_direction_hazard_rule(grid, env):
    """forward x3 then backward x3: on a one-dimensional grid the producing stage of call k+1 stores into the receive buffer a
    slower peer is still reading in the last stage of call k.  P3D_EMU_DELAY makes one rank's last stage slow (emu_runtime.inc),
    so the executor's write-after-read rule is what keeps the results right -- in each of its three forms: the scoped flag
    synchronisation (default: wait for the target peers' release epoch of the buffer, api.cpp `hazard_wait`), world flag
    barriers (P3DFFT_B200_SCOPED=0) and NCCL barriers (P3DFFT_B200_FLAGBAR=0; both `L.dirty[nex->recvbuf]`).  Checked by hand
    for each form: the same run against a build with the rule removed fails with errors of order 1."""
    check(grid, ["--suite", "none", "--repeat"], env)


@pytest.mark.parametrize("grid,chunks,policy", [("1x2", 2, "lazy"), ("2x2", 3, "lazy"), ("2x2", 3, "eager"), ("1x2", 3, "random:5"),
                                               ("2x2", 2, "random:11")])
def test_pipelined_tail_executor(grid, chunks, policy):
    """opt-in P3DFFT_B200_OVERLAP=C: chunked producer / barrier / consumer lists through the real executor on several ranks,
    the consumers on a side stream -- under every order of execution the mock runtime's stream model allows"""
    check(grid, ["--suite", "fast", "--expect-p2p", "1"], {"P3DFFT_B200_OVERLAP": str(chunks), "P3D_EMU_STREAMS": policy})


@pytest.mark.parametrize("grid,chunks,policy", [("1x2", 3, "lazy"), ("2x2", 2, "eager"), ("2x2", 4, "random:7"), ("2x1", 2, "random:2")])
def test_pipelined_group_with_the_split_flag_barrier(grid, chunks, policy):
    """P3DFFT_B200_OVERLAP=C with the flag barrier: every chunk barrier is a `signal` kernel on the main stream and a `wait`
    kernel in front of the consumer chunk on the side stream (api.cpp run_plan) -- also with a delayed rank and with the
    transforms repeated (the write-after-read rule across calls)"""
    env = {"P3DFFT_B200_OVERLAP": str(chunks), "P3DFFT_B200_FLAGBAR": "1", "P3D_EMU_STREAMS": policy}
    check(grid, ["--suite", "fast", "--expect-p2p", "1"], env)
    check(grid, ["--suite", "none", "--repeat"], dict(env, P3D_EMU_DELAY="1:3:2:250"))


@pytest.mark.parametrize("policy", ["eager", "random:3"])
def test_default_path_under_other_stream_orders(policy):
    check("2x2", ["--suite", "fast", "--expect-p2p", "1"], {"P3D_EMU_STREAMS": policy})
    check("1x2", ["--suite", "none", "--repeat"], {"P3D_EMU_STREAMS": policy, "P3D_EMU_DELAY": "1:3:2:250"})


def test_eight_ranks_2x4():
    """the headline grid (BASELINE config 3: 2 x 4) at emulation size"""
    check("2x4", ["--suite", "fast", "--expect-p2p", "1"])


def test_headline_lengths_2x4():
    """1024-point X, Y and Z stages (the lengths of the headline benchmark: split kernel, L2 prefetch, blocked tile order) on
    the 2x4 grid, one long dimension at a time"""
    check("2x4", ["--suite", "long-light", "--expect-p2p", "1"])


@pytest.mark.skipif(not os.environ.get("P3D_EMU_LONG"), reason="opt-in (P3D_EMU_LONG=1): ~1 min, 1024 x 32 x 1024 double and 2048 x 16 x 512 single on 8 ranks")
def test_headline_lengths_2x4_full():
    check("2x4", ["--suite", "long", "--expect-p2p", "1"])


@pytest.mark.parametrize("grid", ["3x2", "1x3"])
def test_grids_that_do_not_divide_the_mesh(grid):
    """64 / 3 pencils: ranks with 21 and 22 lines, partial X tiles, uneven blocks -- under the guard pages of the mock
    allocator (this is the case that exposed an L2 prefetch of xc2r_kernel reaching behind the work buffer)"""
    check(grid, ["--suite", "fast", "--expect-p2p", "1", "--aux"])


SWEEP_ENVS = [{"P3DFFT_B200_BULK": "1"}, {"P3DFFT_B200_BULK": "0"}, {"P3DFFT_B200_R32": "1"}, {"P3DFFT_B200_R32": "0"},
              {"P3DFFT_B200_BULK": "1", "P3DFFT_B200_FLAGBAR": "1", "P3DFFT_B200_OVERLAP": "4"}, {"P3DFFT_B200_ROWB": "64"},
              {"P3DFFT_B200_GENERIC": "1"}, {"P3DFFT_B200_SPLIT": "1"}]


@pytest.mark.skipif(not os.environ.get("P3D_EMU_LONG"), reason="opt-in (P3D_EMU_LONG=1): every switchable variant at the 1024-point lengths on 4 ranks, ~2 min")
@pytest.mark.parametrize("grid", ["2x2", "1x4"])
@pytest.mark.parametrize("env", SWEEP_ENVS, ids=lambda e: "+".join(k[12:] for k in e))
def test_switchable_variants_at_headline_lengths(grid, env):
    check(grid, ["--suite", "long-light", "--expect-p2p", "1"], env)


@pytest.mark.parametrize("grid,policy", [("1x2", "lazy"), ("2x2", "random:4")])
def test_asynchronous_calls_on_several_ranks(grid, policy):
    """p3dfft_b200_set_async: three forward and three backward transforms are only enqueued (device arrays), one
    p3dfft_b200_sync at the end; barriers and exchanges are stream work like the kernels"""
    check(grid, ["--suite", "none", "--repeat", "--async"], {"P3D_EMU_STREAMS": policy, "P3D_EMU_DELAY": "1:3:2:100"})


@pytest.mark.parametrize("grid,seed", [("2x2", 3), ("3x2", 5)])
def test_random_cases(grid, seed):
    """a dozen seeded random cases per grid (sizes for the specialised and the any-length kernels, pruning, every third-dimension
    variant, STRIDE1, two variables, both precisions); hundreds of such cases on nine grids passed when this was written"""
    check(grid, ["--suite", f"fuzz:{seed}:12"])
