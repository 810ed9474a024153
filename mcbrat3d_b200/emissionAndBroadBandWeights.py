"""Host-side mirror of ``src/emissionAndBroadBandWeights.f95`` (setup-time staging).

``type(Weights)`` holds the Planck-emission CDF over voxels that the thermal source kernel
searches (``voxelWeights`` with ``colWeights``/``levelWeights`` as slices of it, EMI:56-57)
and ``fracAtmsPower``.  ``emission_weighting`` follows ``emission_weightingNEW`` (EMI:424-550).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class Weights:
    voxelWeights: Optional[np.ndarray] = None      # (nz, ny, nx) CDF in x-fastest order
    fracAtmsPower: float = 0.0
    spectrIntgrFlux: float = 0.0                   # W m^-2 (monochromatic, EMI:536-538)

    @property
    def levelWeights(self):                        # voxelWeights(nx, ny, :)
        return self.voxelWeights[:, -1, -1]

    @property
    def colWeights(self):                          # voxelWeights(nx, :, :)
        return self.voxelWeights[:, :, -1]


def new_Weights(numX=None, numY=None, numZ=None, numLambda=1) -> Weights:
    """EMI:40-62."""
    if numX is None or numY is None or numZ is None:
        return Weights()
    return Weights(voxelWeights=np.zeros((numZ, numY, numX), dtype=np.float64))


def emission_weighting(thisDomain, theseWeights: Weights, sfcTemp: float) -> float:
    """``emission_weightingNEW`` (EMI:424-550) for the domain's wavelength; returns totalFlux.

    The running sum is compensated (Kahan) in the reference (EMI:505-509); here the same
    per-voxel terms are accumulated with ``math.fsum``-grade accuracy via a long-double
    cumulative sum, which agrees with the compensated sum to the last bit or two.
    """
    h = 6.62606957e-34; c = 2.99792458e+8; k = 1.3806488e-23
    a = 2.0 * h * c ** 2.0
    Pi = 4.0 * np.arctan(1.0)
    d = thisDomain
    nx, ny, nz = d.numX, d.numY, d.numZ
    if d.totalExt is None:
        raise ValueError("emission_weighting: domain hasn't been initialized.")
    lam = d.lambda_um / 1.0e6
    b = h * c / (k * lam)
    emiss = 1.0 - d.surfaceAlbedo
    areaX = d.xPosition[-1] - d.xPosition[0]
    areaY = d.yPosition[-1] - d.yPosition[0]
    if emiss == 0.0 or sfcTemp == 0.0:
        sfcPower = 0.0
    else:
        sfcPlanckRad = (a / ((lam ** 5.0) * (np.exp(b / sfcTemp) - 1.0))) / 1.0e6
        sfcPower = Pi * emiss * sfcPlanckRad * areaX * areaY * (1000.0 ** 2.0)
    nc = d.cumulativeExt.shape[0]
    ext = np.empty_like(d.cumulativeExt)                    # OPT:872-882
    ext[0] = d.totalExt * d.cumulativeExt[0]
    for j in range(1, nc):
        ext[j] = d.totalExt * (d.cumulativeExt[j] - d.cumulativeExt[j - 1])
    totalAbsCoef = d.totalExt - np.sum(d.ssa * ext, axis=0)
    cdf = np.zeros((nz, ny, nx), dtype=np.float64)
    if not np.any(d.temps <= 0.0):
        planck = (a / ((lam ** 5.0) * (np.exp(b / d.temps) - 1.0))) / 1.0e6
        dz = np.diff(d.zPosition)[:, None, None]
        contrib = 4.0 * Pi * planck * totalAbsCoef * dz
        cdf = np.cumsum(contrib.ravel().astype(np.longdouble)).astype(np.float64).reshape(nz, ny, nx)
    atmsPower = 0.0
    last = cdf[-1, -1, -1]
    if last > 0.0:
        atmsPower = last * areaX * areaY * (1000.0 ** 2.0) / float(nx * ny)
        cdf = cdf / last
        cdf[-1, -1, -1] = 1.0
        theseWeights.fracAtmsPower = atmsPower / (atmsPower + sfcPower)
    if atmsPower + sfcPower == 0.0:
        raise ValueError("emission_weightingNEW: Neither surface nor atmosphere will emitt photons "
                         "since total power is 0. Not a valid solution")
    theseWeights.voxelWeights = np.ascontiguousarray(cdf)
    theseWeights.spectrIntgrFlux = (atmsPower + sfcPower) / (areaX * areaY * (1000.0 ** 2.0))
    return theseWeights.spectrIntgrFlux
