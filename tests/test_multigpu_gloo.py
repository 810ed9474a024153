"""The N>1 host path on CPU: two gloo ranks partition the photons and sum their tallies with
one reduce, exactly the plumbing the GPU run uses with NCCL (multipleProcesses.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from mcbrat3d_b200 import multipleProcesses as mpx
    numProcs, thisProc = mpx.initializeProcesses(backend="gloo")
    assert (numProcs, thisProc) == (world, rank) and mpx.MasterProc == (rank == 0)
    first, count = mpx.photonRange(N, numProcs, thisProc)
    # a stand-in tally: each "photon id" deposits a deterministic weight in a column -- the sum
    # over ranks must not depend on the partition (what the counter-based RNG guarantees on the GPU)
    ids = np.arange(first, first + count, dtype=np.int64)
    tally = np.bincount(ids % 16, weights=(ids % 7 + 1).astype(np.float64), minlength=17)
    tally[16] = count                                       # photons started (last slot of the tally buffer)
    t = torch.from_numpy(tally.copy())
    out = mpx.sumAcrossProcesses(t)                        # in-place reduce to rank 0
    arr = mpx.sumAcrossProcesses(tally)                    # numpy overload
    mpx.synchronizeProcesses()
    if rank == 0:
        q.put((out.numpy().copy(), arr))
    mpx.finalizeProcesses()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_and_reduce_gloo(world):
    N = 100003
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    got_t, got_a = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ids = np.arange(N, dtype=np.int64)
    want = np.bincount(ids % 16, weights=(ids % 7 + 1).astype(np.float64), minlength=17)
    want[16] = N
    np.testing.assert_array_equal(got_t, want)
    np.testing.assert_array_equal(got_a, want)
