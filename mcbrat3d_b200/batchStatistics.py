"""The driver's per-batch statistics (``Drivers/monteCarloDriver.f95``), host side.

First and second moments over batches weighted by photons per batch (DRV:1023-1052) and the
final mean / standard error (DRV:1188-1228).  The Fortran driver itself is kept as the host in
a deployment (north star); this mirror exists so that the Python host and the tests form the
statistics exactly the way the reference does.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


class BatchStatistics:
    def __init__(self):
        self.moment1: Dict[str, np.ndarray] = {}
        self.moment2: Dict[str, np.ndarray] = {}
        self.totalNumPhotons = 0
        self.batchesCompleted = 0

    def accumulate(self, results: Dict[str, np.ndarray], numPhotonsProcessed: int) -> None:
        """DRV:1023-1052: stats(:,1) += x*n ; stats(:,2) += n*x**2 (double precision)."""
        n = float(numPhotonsProcessed)
        for k, v in results.items():
            x = np.asarray(v, dtype=np.float64)
            if k not in self.moment1:
                self.moment1[k] = np.zeros_like(x)
                self.moment2[k] = np.zeros_like(x)
            self.moment1[k] += x * n
            self.moment2[k] += n * x ** 2.0
        self.totalNumPhotons += int(numPhotonsProcessed)
        self.batchesCompleted += 1

    def merge(self, other: "BatchStatistics") -> None:
        """sumAcrossProcesses over workers (DRV:1151-1166)."""
        for k in other.moment1:
            if k not in self.moment1:
                self.moment1[k] = other.moment1[k].copy(); self.moment2[k] = other.moment2[k].copy()
            else:
                self.moment1[k] += other.moment1[k]; self.moment2[k] += other.moment2[k]
        self.totalNumPhotons += other.totalNumPhotons
        self.batchesCompleted += other.batchesCompleted

    def finalise(self, solarFlux: float = 1.0):
        """DRV:1188-1228: returns ({name: mean}, {name: standard error})."""
        mean, err = {}, {}
        for k in self.moment1:
            m1 = solarFlux * self.moment1[k] / self.totalNumPhotons
            m2 = solarFlux * (solarFlux * self.moment2[k] / self.totalNumPhotons)
            mean[k] = m1
            err[k] = np.sqrt(np.maximum(0.0, m2 - m1 ** 2.0) / (self.batchesCompleted - 1))
        return mean, err


# ------------------------------------------------------------------------------------------------
# the same loop and statistics on the GPU (csrc/mcb_stage.cu, mcb_run_batches / mcb_get_statistics)
# ------------------------------------------------------------------------------------------------
def resetDeviceStatistics(thisIntegrator, thisDomain=None) -> None:
    """Zero the device-side moments (the driver zeroes its *Stats arrays at DRV:610-660).  Pass the
    domain if it has not been staged on this integrator yet."""
    g = thisIntegrator
    if thisDomain is not None:
        from .monteCarloRadiativeTransfer import _stage_domain
        _stage_domain(g, thisDomain)
    g._check(g._lib.mcb_stats_reset(g.handle), "resetDeviceStatistics")


def computeRadiativeTransferBatches(thisIntegrator, thisDomain, randomNumbers, incomingPhotons,
                                    numPhotonsPerBatch: int, numBatches: int, synchronize: bool = True) -> int:
    """The driver's batch loop (DRV:949-1052) in one call: ``numBatches`` batches of ``numPhotonsPerBatch``
    photons taken from ``incomingPhotons``; after every batch the device folds the normalised results into
    the first and second moments.  Nothing is copied back until ``reportStatistics``.  Returns photons traced."""
    import ctypes as C

    from .monteCarloRadiativeTransfer import McbError, _stage_domain, _stage_source
    g = thisIntegrator
    _stage_domain(g, thisDomain)
    _stage_source(g, incomingPhotons)
    need = int(numPhotonsPerBatch) * int(numBatches)
    left = incomingPhotons.numberOfPhotons - (incomingPhotons.currentPhoton - 1)
    if incomingPhotons.currentPhoton < 1 or left < need:
        raise McbError("computeRadiativeTransfer: the photon stream holds %d photons, %d batches of %d need %d"
                       % (max(left, 0), numBatches, numPhotonsPerBatch, need))
    first = incomingPhotons.firstPhotonId + (incomingPhotons.currentPhoton - 1)
    done = C.c_int64(0)
    g._check(g._lib.mcb_run_batches(g.handle, int(numBatches), int(numPhotonsPerBatch), C.c_uint64(randomNumbers.seed),
                                    C.c_uint64(first), C.byref(done)), "computeRadiativeTransfer")
    incomingPhotons.currentPhoton += need
    if synchronize:
        g._check(g._lib.mcb_synchronize(g.handle), "computeRadiativeTransfer")
    return int(done.value)


def reportStatistics(thisIntegrator, solarFlux: float = 1.0, volumeAbsorption: bool = True, intensity: bool = None):
    """DRV:1188-1228 on the device: returns ({name: mean}, {name: standard error}, totalNumPhotons,
    batchesCompleted) with the names ``BatchStatistics.finalise`` uses."""
    import ctypes as C

    from . import _lib
    g = thisIntegrator
    nx, ny, nz = g.numX, g.numY, g.numZ
    nDir = 0 if g.intensityDirections is None else g.intensityDirections.shape[0]
    if intensity is None:
        intensity = nDir > 0
    mf = np.empty((2, 3)); up = np.empty((2, ny, nx)); dn = np.empty((2, ny, nx)); ab = np.empty((2, ny, nx))
    prof = np.empty((2, nz))
    vol = np.empty((2, nz, ny, nx)) if volumeAbsorption else None
    rad = np.empty((2, nDir, ny, nx)) if intensity else None
    tot = C.c_int64(0); nb = C.c_int64(0)
    d = C.c_double
    g._check(g._lib.mcb_get_statistics(g.handle, float(solarFlux), _lib.ptr(mf, d), _lib.ptr(up, d), _lib.ptr(dn, d),
                                       _lib.ptr(ab, d), _lib.ptr(prof, d), _lib.ptr(vol, d), _lib.ptr(rad, d),
                                       C.byref(tot), C.byref(nb)), "reportStatistics")
    mean = {"meanFluxUp": mf[0, 0], "meanFluxDown": mf[0, 1], "meanFluxAbsorbed": mf[0, 2], "fluxUp": up[0],
            "fluxDown": dn[0], "fluxAbsorbed": ab[0], "absorbedProfile": prof[0]}
    err = {"meanFluxUp": mf[1, 0], "meanFluxDown": mf[1, 1], "meanFluxAbsorbed": mf[1, 2], "fluxUp": up[1],
           "fluxDown": dn[1], "fluxAbsorbed": ab[1], "absorbedProfile": prof[1]}
    if volumeAbsorption:
        mean["volumeAbsorption"], err["volumeAbsorption"] = vol[0], vol[1]
    if intensity:
        mean["intensity"], err["intensity"] = rad[0], rad[1]
    return mean, err, int(tot.value), int(nb.value)
