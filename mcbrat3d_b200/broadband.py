"""Host-side mirror of the driver's spectral loop (``Drivers/monteCarloDriver.f95``) for broadband runs
(BASELINE config ``I3RC_bench_SW/LW``): which wavelength bin every photon belongs to, and the per-bin sequence
read_SSPTable -> [emission_weighting] -> new_PhotonStream -> computeRadiativeTransfer -> moment accumulation.

Everything of size O(cells) or O(photons) runs on the GPU through the C ABI:
  * the LW set-up pass (DRV:307-407): per bin ``read_SSPTable(setup=.true.)`` + ``emission_weighting`` for the
    emitted flux -> ``mcb_assemble_optics`` + ``mcb_build_thermal_source``;
  * the flux CDF over bins (Kahan-summed, DRV:417-433 / ``solar_Weighting`` EMI:149-208): O(numLambda), host;
  * ``getFrequencyDistr`` (EMI:552-573), one draw per photon -> ``mcb_frequency_distribution``;
  * the worker loop (DRV:903-1052) -> ``mcb_assemble_optics``, ``mcb_build_thermal_source``, ``mcb_run_batches``
    (the moments of DRV:1023-1052 never leave the device) and ``mcb_get_statistics`` (DRV:1188-1228).
The master/worker message passing of the driver (DRV:665-880) has no counterpart: photons are identified by global
id, each rank takes a contiguous share of every bin, and the moment buffer is summed once at the end.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from .batchStatistics import computeRadiativeTransferBatches, reportStatistics, resetDeviceStatistics
from .emissionAndBroadBandWeights import Weights, emission_weighting
from .monteCarloIllumination import new_PhotonStream
from .monteCarloRadiativeTransfer import specifyParameters
from .multipleProcesses import numProcs, photonRange, sumStatisticsAcrossProcesses, thisProc
from .opticalProperties import SSPTable, commonDomain, light_spd, read_SSPTable


def kahanCDF(contributions) -> Tuple[np.ndarray, float]:
    """Compensated running sum, then normalisation with the last entry forced to 1 (DRV:417-433; the same loop
    as ``solar_Weighting`` EMI:176-203).  Returns (CDF, total)."""
    cdf = np.array(contributions, dtype=np.float64)
    corr = 0.0
    for i in range(1, cdf.size):
        corrContr = cdf[i] - corr
        tempSum = cdf[i - 1] + corrContr
        corr = (tempSum - cdf[i - 1]) - corrContr
        cdf[i] = tempSum
    total = float(cdf[-1])
    cdf = cdf / total
    cdf[-1] = 1.0
    return cdf, total


def solar_Weighting(solarSourceFunction, lambdas, solarMu) -> Tuple[np.ndarray, float]:
    """``solar_Weighting`` (EMI:149-208) without an instrument response file: band widths from the half points
    between neighbouring wavelengths; returns (totalPowerCDF, totalFlux)."""
    lam = np.asarray(lambdas, dtype=np.float64); src = np.asarray(solarSourceFunction, dtype=np.float64)
    n = lam.size
    d = np.empty(n)
    d[0] = abs(lam[1] - lam[0])
    d[1:n - 1] = np.abs((lam[2:] - lam[:-2]) / 2.0)
    d[n - 1] = abs(lam[n - 1] - lam[n - 2])
    mu = np.float64(np.float32(solarMu))
    return kahanCDF(d * mu * src)


def bandWidths(lambdas) -> np.ndarray:
    """dLambda as the LW set-up loop forms it (DRV:336-360)."""
    lam = np.asarray(lambdas, dtype=np.float64)
    n = lam.size
    d = np.empty(n)
    d[0] = abs(lam[1] - lam[0])
    d[1:n - 1] = np.abs((lam[2:] - lam[:-2]) / 2.0)
    d[n - 1] = abs(lam[n - 1] - lam[n - 2])
    return d


def getFrequencyDistr(thisIntegrator, CDF, totalPhotons: int, seed: int) -> np.ndarray:
    """``getFrequencyDistr`` (EMI:552-573) on the device."""
    g = thisIntegrator
    cdf = np.ascontiguousarray(CDF, dtype=np.float64)
    out = np.zeros(cdf.size, dtype=np.int64)
    g._check(g._lib.mcb_frequency_distribution(g.handle, cdf.size, _lib.ptr(cdf, C.c_double), int(totalPhotons),
                                               C.c_uint64(seed), out.ctypes.data_as(C.POINTER(C.c_int64))),
             "getFrequencyDistr")
    return out


def runBroadband(thisIntegrator, tables: List[SSPTable], commonD: commonDomain, randomNumbers, totalPhotons: int,
                 numPhotonsPerBatch: int, solarMu: float = 0.5, solarAzimuth: float = 0.0,
                 solarSourceFunction=None, LW: bool = False, surfaceTemp: float = 290.0, calcRayl: bool = True,
                 minBatchesPerBin: int = 1, counterSink: Dict = None) -> Dict:
    """One broadband run on this rank's GPU (all ranks call it; rank r traces its share of every bin).

    ``counterSink`` (a dict) receives the event counters summed over the bins (one extra synchronisation per bin:
    for measurements, not for production runs).

    Returns a dict with ``mean`` / ``err`` (the driver's finalised statistics, W m^-2 when the source function is in
    W m^-2 um^-1), ``freqDistr``, ``solarFlux``, ``totalNumPhotons``, ``batchesCompleted``."""
    g = thisIntegrator
    commonD.temps.setflags(write=False)                    # the physical state is frozen for the run (uploaded once)
    nLambda = tables[0].f_grid.size
    lambdas = light_spd * 1e6 / np.asarray(tables[0].f_grid, dtype=np.float64)
    # ---- set-up: flux per bin -> CDF -> photons per bin (DRV:307-445 LW, 447-503 SW) ----
    if LW:
        dLam = bandWidths(lambdas)
        flux = np.empty(nLambda)
        for i in range(1, nLambda + 1):
            d = read_SSPTable(tables, i, commonD, setup=True, calcRayl=False, thisIntegrator=g)
            w = Weights()
            flux[i - 1] = emission_weighting(d, w, surfaceTemp, thisIntegrator=g) * dLam[i - 1]
        fluxCDF, solarFlux = kahanCDF(flux)
    else:
        if solarSourceFunction is None:
            raise ValueError("runBroadband: a solar source function is needed for a SW run")
        fluxCDF, solarFlux = solar_Weighting(solarSourceFunction, lambdas, solarMu)
    freqDistr = getFrequencyDistr(g, fluxCDF, totalPhotons, randomNumbers.seed ^ 0x5DEECE66D)
    # ---- the spectral loop (DRV:903-1052) ----
    specifyParameters(g, LW_flag=1.0 if LW else -1.0, buildTablesOnDevice=True)
    started = False
    world, rank = numProcs(), thisProc()
    for i in range(1, nLambda + 1):
        first, mine = photonRange(int(freqDistr[i - 1]), world, rank)
        # every rank must advance the global photon ids identically, whatever its own share is
        base = randomNumbers.nextPhotonId
        randomNumbers.nextPhotonId = base + int(freqDistr[i - 1])
        if mine <= 0:
            continue
        d = read_SSPTable(tables, i, commonD, setup=False, calcRayl=calcRayl, thisIntegrator=g)
        if not started:
            resetDeviceStatistics(g)
            started = True
        nb = max(minBatchesPerBin, -(-mine // int(numPhotonsPerBatch)))
        per = mine // nb                                   # equal batches; the remainder rides in one extra batch
        sub = type(randomNumbers)(randomNumbers.seed); sub.nextPhotonId = base + first
        if LW:
            w = Weights()
            emission_weighting(d, w, surfaceTemp, thisIntegrator=g)
            mk = lambda n: new_PhotonStream(theseWeights=w, numberOfPhotons=n, randomNumbers=sub)
        else:
            mk = lambda n: new_PhotonStream(solarMu, solarAzimuth, n, sub)
        if per > 0:
            computeRadiativeTransferBatches(g, d, sub, mk(per * nb), per, nb, synchronize=False)
        rest = mine - per * nb
        if counterSink is not None and per > 0:
            from .monteCarloRadiativeTransfer import getCounters
            for k, v in getCounters(g).items():
                counterSink[k] = counterSink.get(k, 0) + v
        if rest > 0:
            computeRadiativeTransferBatches(g, d, sub, mk(rest), rest, 1, synchronize=False)
    sumStatisticsAcrossProcesses(g)
    mean, err, tot, nb = reportStatistics(g, solarFlux=solarFlux)
    return dict(mean=mean, err=err, freqDistr=freqDistr, solarFlux=solarFlux, fluxCDF=fluxCDF, lambdas=lambdas,
                totalNumPhotons=tot, batchesCompleted=nb)
