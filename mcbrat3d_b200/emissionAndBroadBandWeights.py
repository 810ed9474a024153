"""Host-side mirror of ``src/emissionAndBroadBandWeights.f95`` (setup-time staging).

``type(Weights)`` holds the Planck-emission CDF over voxels that the thermal source kernel
searches (``voxelWeights`` with ``colWeights``/``levelWeights`` as slices of it, EMI:56-57)
and ``fracAtmsPower``.  ``emission_weighting`` follows ``emission_weightingNEW`` (EMI:424-550).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from .opticalProperties import new_token


@dataclass
class Weights:
    voxelWeights: Optional[np.ndarray] = None      # (nz, ny, nx) CDF in x-fastest order
    fracAtmsPower: float = 0.0
    spectrIntgrFlux: float = 0.0                   # W m^-2 (monochromatic, EMI:536-538)
    deviceOwner: Optional[object] = None           # the integrator whose HBM holds the CDF (device-built weights)
    token: int = 0                                 # staging-cache key: replaced by every emission_weighting call

    def __post_init__(self):
        self.token = new_token()

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


def emission_weighting_device(thisDomain, theseWeights: Weights, sfcTemp: float, thisIntegrator) -> float:
    """``emission_weightingNEW`` (EMI:424-550) on the GPU: the domain's optics are already in the
    integrator's HBM, the temperatures are uploaded, Planck emission and the prefix sum over cells run
    in ``csrc/mcb_stage.cu``; the CDF never visits the host (``fetchVoxelWeights`` reads it back on
    request).  Returns totalFlux."""
    import ctypes as C

    from . import _lib
    from .monteCarloRadiativeTransfer import _stage_domain
    g = thisIntegrator
    if thisDomain.totalExt is None and getattr(thisDomain, "deviceOwner", None) is None:
        raise ValueError("emission_weighting: domain hasn't been initialized.")
    _stage_domain(g, thisDomain)
    temps = np.ascontiguousarray(thisDomain.temps, dtype=np.float64)
    # The temperatures do not depend on the wavelength: a read-only array (runBroadband freezes the physical state of
    # a run) is uploaded once and NULL ("as before") is passed afterwards; the integrator keeps the array alive so
    # that its address cannot be recycled.
    tkey = (temps.ctypes.data, temps.shape)
    reuse = getattr(g, "_stagedTemps", None) == tkey and not temps.flags.writeable
    frac = C.c_double(0.0); flux = C.c_double(0.0)
    g._check(g._lib.mcb_build_thermal_source(g.handle, None if reuse else _lib.ptr(temps, C.c_double), float(thisDomain.lambda_um),
                                             float(sfcTemp), C.byref(frac), C.byref(flux)), "emission_weighting")
    g._stagedTemps = tkey if not temps.flags.writeable else None
    g._stagedTempsRef = temps
    theseWeights.voxelWeights = None
    theseWeights.fracAtmsPower = float(frac.value)
    theseWeights.spectrIntgrFlux = float(flux.value)
    theseWeights.deviceOwner = g
    theseWeights.token = new_token()
    g._stagedSource = ("bbemission-device", theseWeights.token)
    return theseWeights.spectrIntgrFlux


def fetchVoxelWeights(theseWeights: Weights) -> np.ndarray:
    """The CDF of device-built weights, copied back as ``(nz, ny, nx)``."""
    import ctypes as C

    from . import _lib
    g = theseWeights.deviceOwner
    if g is None:
        return theseWeights.voxelWeights
    out = np.empty((g.numZ, g.numY, g.numX), dtype=np.float64)
    g._check(g._lib.mcb_get_thermal_source(g.handle, None, _lib.ptr(out, C.c_double), out.size), "fetchVoxelWeights")
    return out


def emission_weighting(thisDomain, theseWeights: Weights, sfcTemp: float, thisIntegrator=None) -> float:
    """``emission_weightingNEW`` (EMI:424-550) for the domain's wavelength; returns totalFlux.

    With ``thisIntegrator`` the build runs on that integrator's GPU (``emission_weighting_device``);
    without it this is the NumPy staging producer.

    The running sum is compensated (Kahan) in the reference (EMI:505-509); here the same
    per-voxel terms are accumulated with ``math.fsum``-grade accuracy via a long-double
    cumulative sum, which agrees with the compensated sum to the last bit or two.
    """
    if thisIntegrator is not None:
        return emission_weighting_device(thisDomain, theseWeights, sfcTemp, thisIntegrator)
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
    theseWeights.token = new_token()
    theseWeights.deviceOwner = None
    theseWeights.spectrIntgrFlux = (atmsPower + sfcPower) / (areaX * areaY * (1000.0 ** 2.0))
    return theseWeights.spectrIntgrFlux
