"""Host-side mirror of ``src/multipleProcesses_mpi.f95`` on ``torch.distributed``.

The reference farms photon batches to MPI ranks and sums tallies with ``MPI_REDUCE(SUM, root 0)``
(MPIW:70-251, called at DRV:1151-1166).  Here there is one process per GPU; the domain is
replicated in each GPU's HBM, photons are split by global photon id (counter-based RNG, so the
result does not depend on the split) and the only communication is ONE sum-reduce of the packed
f64 tally buffer at the end -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist

MasterProc = True          # MPIW:24; updated by initializeProcesses
_initialised_here = False


def initializeProcesses(backend: str = None) -> Tuple[int, int]:
    """``initializeProcesses`` (MPIW:29-52): returns (numProcs, thisProcNum).

    Reads RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT from the environment (torchrun).  A run
    without those variables is a single process and needs no communicator.
    """
    global MasterProc, _initialised_here
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
        _initialised_here = True
    MasterProc = rank == 0
    return world, rank


def numProcs() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def thisProc() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def synchronizeProcesses() -> None:
    """MPIW:54-60."""
    if dist.is_initialized():
        dist.barrier()


def finalizeProcesses() -> None:
    """MPIW:62-68."""
    global _initialised_here
    if dist.is_initialized() and _initialised_here:
        dist.destroy_process_group()
        _initialised_here = False


def sumAcrossProcesses(x, root: int = 0):
    """``sumAcrossProcesses`` (12 overloads, MPIW:70-251): MPI_REDUCE(SUM) to ``root``.

    Accepts a torch tensor (reduced in place; a CUDA tensor goes over NCCL/NVLink) or anything
    NumPy can view (a summed copy is returned).  Non-root ranks get their own input back, as
    with MPI_REDUCE their receive buffer is undefined.
    """
    if not dist.is_initialized():
        return x
    if isinstance(x, torch.Tensor):
        dist.reduce(x, dst=root, op=dist.ReduceOp.SUM)
        return x
    a = np.ascontiguousarray(x)
    t = torch.from_numpy(a.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.reduce(t, dst=root, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().reshape(a.shape)


def photonRange(totalPhotons: int, numProcs_: int, thisProc_: int) -> Tuple[int, int]:
    """Static partition of the global photon ids [0, N) into contiguous ranges, one per GPU.

    Replaces the master/worker batch farm (DRV:665-880, 903-1085): with per-photon RNG streams
    no dynamic hand-out is needed and every rank (there is no idle master) traces photons.
    Returns (firstPhotonId, count)."""
    base, extra = divmod(int(totalPhotons), int(numProcs_))
    first = thisProc_ * base + min(thisProc_, extra)
    return first, base + (1 if thisProc_ < extra else 0)


def photonShares(totalPhotons: int, rates) -> list:
    """Shares of ``totalPhotons`` proportional to the ranks' measured photon rates (any positive unit), contiguous in
    the global photon id: rank r traces ids [sum(shares[:r]), sum(shares[:r+1])).  The static analogue of the
    master handing batches to workers as they finish (DRV:665-1095): a GPU that runs a few per cent slower gets
    proportionally fewer photons instead of setting the step.  Non-finite or non-positive rates fall back to the mean."""
    r = [float(x) for x in rates]
    good = [x for x in r if x > 0.0 and x == x and x != float("inf")]
    mean = sum(good) / len(good) if good else 1.0
    r = [x if (x > 0.0 and x == x and x != float("inf")) else mean for x in r]
    total = int(totalPhotons)
    shares = [int(total * x / sum(r)) for x in r]
    shares[-1] = total - sum(shares[:-1])
    return shares


def initializeIntegratorProcesses(thisIntegrator) -> bool:
    """Give the integrator its own NCCL communicator through the C ABI (``mcb_comm_init``): the same entry points a
    compiled host uses (``fortran/mcbrat_cuda_mod.f90``, ``examples/i3rc_driver --ranks N``).  Rank 0 creates the
    128-byte id with ``mcb_comm_unique_id``; ``torch.distributed`` only carries it to the other ranks -- the job
    ``MPI_BCAST`` has in the Fortran host.  Returns False (and changes nothing) in a single-process run."""
    if not dist.is_initialized() or dist.get_world_size() < 2:
        return False
    g = thisIntegrator
    ident = (C.c_ubyte * 128)()
    if dist.get_rank() == 0:
        rc = g._lib.mcb_comm_unique_id(C.cast(ident, C.c_void_p))
        if rc != 0:
            raise RuntimeError("initializeProcesses: mcb_comm_unique_id failed (%d): is libnccl.so.2 loadable?" % rc)
    t = torch.tensor(list(ident), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda(g.device)
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    buf = (C.c_ubyte * 128).from_buffer_copy(raw)
    g._check(g._lib.mcb_comm_init(g.handle, dist.get_world_size(), dist.get_rank(), C.cast(buf, C.c_void_p)),
             "initializeProcesses")
    g._hasComm = True
    return True


class _DeviceBuffer:
    """Exposes a raw device pointer to torch through ``__cuda_array_interface__``."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def tallyTensor(thisIntegrator) -> torch.Tensor:
    """The integrator's packed f64 tally buffer as a CUDA tensor aliasing the library's memory
    (``mcb_tally_buffer``): the one message of the run."""
    ptr = C.c_void_p()
    n = C.c_int64(0)
    thisIntegrator._check(thisIntegrator._lib.mcb_tally_buffer(thisIntegrator.handle, C.byref(ptr), C.byref(n)),
                          "tallyTensor")
    return torch.as_tensor(_DeviceBuffer(ptr.value, n.value), device="cuda:%d" % thisIntegrator.device)


def sumTalliesAcrossProcesses(thisIntegrator, root: int = 0) -> None:
    """One NCCL sum-reduce of all tallies (replaces the nine reduces at DRV:1151-1166)."""
    if getattr(thisIntegrator, "_hasComm", False):               # the library's own communicator: ncclReduce on its stream
        thisIntegrator._check(thisIntegrator._lib.mcb_reduce_tallies(thisIntegrator.handle, int(root)), "sumTalliesAcrossProcesses")
        thisIntegrator._check(thisIntegrator._lib.mcb_synchronize(thisIntegrator.handle), "sumTalliesAcrossProcesses")
        return
    if not dist.is_initialized():
        return
    thisIntegrator._check(thisIntegrator._lib.mcb_synchronize(thisIntegrator.handle), "sumTalliesAcrossProcesses")
    t = tallyTensor(thisIntegrator)
    dist.reduce(t, dst=root, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize(thisIntegrator.device)


def statisticsTensor(thisIntegrator) -> torch.Tensor:
    """The device-side batch-statistics buffer (``mcb_stats_buffer``: first moments, second moments,
    totalNumPhotons, batchesCompleted) as a CUDA tensor aliasing the library's memory."""
    ptr = C.c_void_p()
    n = C.c_int64(0)
    thisIntegrator._check(thisIntegrator._lib.mcb_stats_buffer(thisIntegrator.handle, C.byref(ptr), C.byref(n)),
                          "statisticsTensor")
    return torch.as_tensor(_DeviceBuffer(ptr.value, n.value), device="cuda:%d" % thisIntegrator.device)


def sumStatisticsAcrossProcesses(thisIntegrator, root: int = 0) -> None:
    """Moments, photon counts and batch counts add across ranks exactly as the driver's
    ``sumAcrossProcesses`` calls do (DRV:1151-1166): one reduce of the moment buffer."""
    if getattr(thisIntegrator, "_hasComm", False):
        thisIntegrator._check(thisIntegrator._lib.mcb_reduce_statistics(thisIntegrator.handle, int(root)), "sumStatisticsAcrossProcesses")
        thisIntegrator._check(thisIntegrator._lib.mcb_synchronize(thisIntegrator.handle), "sumStatisticsAcrossProcesses")
        return
    if not dist.is_initialized():
        return
    thisIntegrator._check(thisIntegrator._lib.mcb_synchronize(thisIntegrator.handle), "sumStatisticsAcrossProcesses")
    t = statisticsTensor(thisIntegrator)
    dist.reduce(t, dst=root, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize(thisIntegrator.device)
