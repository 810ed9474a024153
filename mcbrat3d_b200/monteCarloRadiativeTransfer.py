"""Host-side mirror of ``Integrators/monteCarloRadiativeTransfer.f95`` -- ``type(integrator)``.

Same public procedures as the reference module (INT:121-123): ``new_Integrator``,
``specifyParameters``, ``computeRadiativeTransfer``, ``reportResults``, ``isReady_Integrator``,
``finalize_Integrator``.  The bodies do what the Fortran ISO_C_BINDING shim does
(``fortran/mcbrat_cuda_mod.f90``): hand the domain's arrays to the CUDA library once and launch
the photon kernels through the C ABI.  The photon loop (``computeRT`` INT:393-841) never runs
on the host and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (MCB_ARITH_FAST, MCB_ARITH_REFERENCE, MCB_KERNEL_PARK, MCB_KERNEL_POOL, MCB_LAYOUT_BRICKS,
                   MCB_LAYOUT_LINEAR, McbError, mcb_counters, mcb_options)
from .monteCarloIllumination import morePhotonsExist, photonStream
from .opticalProperties import Domain
from .RandomNumbersForMC import randomNumberSequence

f32 = np.float32
Pi = f32(3.14159265358979312)                        # INT:31
defaultMinForwardTableSize = 9001                    # INT:24-25
defaultMinInverseTableSize = 9001
defaultHybridPhaseFunWidth = 7.0
maxHybridPhaseFunWidth = 30.0


def makeDirectionCosines(mu, phi) -> np.ndarray:
    """INT:1876-1894 in single precision (cos/sin correctly rounded)."""
    mu = f32(mu); phi = f32(phi)
    sinTheta = np.sqrt(f32(1.0) - mu * mu, dtype=f32)
    cosPhi = f32(np.cos(np.float64(phi))); sinPhi = f32(np.sin(np.float64(phi)))
    return np.array([sinTheta * cosPhi, sinTheta * sinPhi, mu], dtype=f32)


class integrator:
    """``type(integrator)`` (INT:40-117).  Owns one ``mcb_handle`` on one GPU."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.mcb_create(int(device), C.byref(self._h))
        if rc != 0:
            raise McbError("new_Integrator: mcb_create failed with code %d (no CUDA device? there is no CPU fallback)" % rc)
        self.device = int(device)
        self.readyToCompute = False
        self.computeIntensity = False
        self.minForwardTableSize = defaultMinForwardTableSize
        self.minInverseTableSize = defaultMinInverseTableSize
        self.hybridPhaseFunWidth = defaultHybridPhaseFunWidth
        self.options = mcb_options()
        self._lib.mcb_default_options(C.byref(self.options))
        self.intensityDirections: Optional[np.ndarray] = None       # (nDir, 3)
        self.numX = self.numY = self.numZ = 0
        self.numComps = 0
        self.buildTablesOnDevice = False                             # inverse tables: NumPy mirror, or csrc/mcb_stage.cu
        self._stagedDomain = None
        self._stagedTables = None
        self._stagedSource = None
        self.numPhotonsProcessed = 0

    # -- helpers ------------------------------------------------------------------------
    def _check(self, rc, where):
        _lib.check(self._lib, self._h, rc, where)

    def __del__(self):
        try:
            finalize_Integrator(self)
        except Exception:
            pass

    @property
    def handle(self):
        return self._h


def new_Integrator(atmosphere: Domain, device: int = 0) -> integrator:
    """``new_Integrator`` (INT:129-201): stage the grid, decide regular/irregular (quirk q1)."""
    new = integrator(device)
    new.numX, new.numY, new.numZ = atmosphere.numX, atmosphere.numY, atmosphere.numZ
    new._check(new._lib.mcb_set_grid(new._h, new.numX, new.numY, new.numZ,
                                     _lib.ptr(atmosphere.xPosition, C.c_double),
                                     _lib.ptr(atmosphere.yPosition, C.c_double),
                                     _lib.ptr(atmosphere.zPosition, C.c_double)), "new_Integrator")
    new._grid = (atmosphere.xPosition.copy(), atmosphere.yPosition.copy(), atmosphere.zPosition.copy())
    new.readyToCompute = True
    return new


def isReady_Integrator(thisIntegrator: integrator) -> bool:
    return bool(thisIntegrator.readyToCompute)


def finalize_Integrator(thisIntegrator: integrator) -> None:
    """INT:1486-1547."""
    if getattr(thisIntegrator, "_h", None) is not None and thisIntegrator._h:
        thisIntegrator._lib.mcb_destroy(thisIntegrator._h)
        thisIntegrator._h = C.c_void_p()
    thisIntegrator.readyToCompute = False


def specifyParameters(thisIntegrator: integrator, minForwardTableSize=None, minInverseTableSize=None,
                      intensityMus: Optional[Sequence[float]] = None, intensityPhis: Optional[Sequence[float]] = None,
                      computeIntensity: Optional[bool] = None, useRayTracing: Optional[bool] = None,
                      useRussianRoulette: Optional[bool] = None, useRussianRouletteForIntensity: Optional[bool] = None,
                      zetaMin: Optional[float] = None, useHybridPhaseFunsForIntenCalcs: Optional[bool] = None,
                      hybridPhaseFunWidth: Optional[float] = None, numOrdersOrigPhaseFunIntenCalcs: Optional[int] = None,
                      limitIntensityContributions: Optional[bool] = None, maxIntensityContribution: Optional[float] = None,
                      LW_flag: Optional[float] = None, numComps: Optional[int] = None,
                      arithmetic: Optional[int] = None, buildTablesOnDevice: Optional[bool] = None,
                      tuneKernel: Optional[int] = None, tuneLayout: Optional[int] = None,
                      tuneBlocksPerSM: Optional[int] = None, tuneParkThreshold: Optional[int] = None,
                      tuneLeCarry: Optional[int] = None, tuneExtMask: Optional[int] = None,
                      tuneBurst: Optional[int] = None, tuneLeap: Optional[int] = None,
                      tuneLeapLanes: Optional[int] = None) -> None:
    """``specifyParameters`` (INT:1046-1337): same optional arguments, same checks.

    ``arithmetic`` is one addition: ``MCB_ARITH_FAST`` (default) or ``MCB_ARITH_REFERENCE``;
    ``buildTablesOnDevice`` the other: the inverse phase-function tables of INV:66-174 are then built by
    ``mcb_build_inverse_table`` in HBM instead of by the NumPy mirror (the default).
    The ``tune*`` arguments are the measurement knobs of ``mcb_options`` (kernel variant, field layout, occupancy,
    ...; 0 = the library's own choice): tests and profiling scripts compare variants with them.
    ``surfaceBDRF`` and ``recScatOrd``/``numRecScatOrd`` are not supported (the driver never
    installs a BDRF, INT:667-674; the by-order tallies are commented out in the reference).
    """
    g = thisIntegrator
    o = g.options
    if (intensityMus is None) != (intensityPhis is None):
        raise ValueError("specifyParameters: Both or neither of intensityMus and intensityPhis must be supplied")
    if intensityMus is not None:
        mus = np.asarray(intensityMus, dtype=f32); phis = np.asarray(intensityPhis, dtype=f32)
        if mus.size != phis.size:
            raise ValueError("specifyParameters: intensityMus, intensityPhis must be the same length.")
        if np.any(mus < -1) or np.any(mus > 1):
            raise ValueError("specifyParameters: intensityMus must be between -1 and 1")
        if np.any(np.abs(mus) < np.finfo(f32).tiny):
            raise ValueError("specifyParameters: intensityMus can't be 0 (directly sideways)")
        if np.any(phis < 0) or np.any(phis > 360):
            raise ValueError("specifyParameters: intensityPhis must be between 0 and 360")
    if computeIntensity and intensityMus is None and g.intensityDirections is None:
        raise ValueError("specifyParameters: Can't compute intensity without specifying directions.")
    if zetaMin is not None and zetaMin < 0:
        raise ValueError("specifyParameters: zetaMin must be >= 0.")

    if useRayTracing is not None: o.useRayTracing = int(bool(useRayTracing))
    if minForwardTableSize is not None: g.minForwardTableSize = max(int(minForwardTableSize), defaultMinForwardTableSize)
    if minInverseTableSize is not None: g.minInverseTableSize = max(int(minInverseTableSize), defaultMinInverseTableSize)
    if useRussianRoulette is not None: o.useRussianRoulette = int(bool(useRussianRoulette))
    if useRussianRouletteForIntensity is not None: o.useRussianRouletteForIntensity = int(bool(useRussianRouletteForIntensity))
    if zetaMin is not None: o.zetaMin = float(zetaMin)
    if useHybridPhaseFunsForIntenCalcs is not None:
        o.useHybridPhaseFunsForIntenCalcs = int(bool(useHybridPhaseFunsForIntenCalcs))
        g._stagedTables = None
    if hybridPhaseFunWidth is not None:
        g.hybridPhaseFunWidth = (float(hybridPhaseFunWidth)
                                 if 0 < hybridPhaseFunWidth < maxHybridPhaseFunWidth else defaultHybridPhaseFunWidth)
        g._stagedTables = None
    if numOrdersOrigPhaseFunIntenCalcs is not None:
        o.numOrdersOrigPhaseFunIntenCalcs = int(numOrdersOrigPhaseFunIntenCalcs) if numOrdersOrigPhaseFunIntenCalcs >= 0 else 0
    if limitIntensityContributions is not None: o.limitIntensityContributions = int(bool(limitIntensityContributions))
    if maxIntensityContribution is not None and maxIntensityContribution > 0:
        o.maxIntensityContribution = float(maxIntensityContribution)
    if LW_flag is not None: o.LW_flag = float(LW_flag)
    if arithmetic is not None: o.arithmetic = int(arithmetic)
    for name, val in (("tuneKernel", tuneKernel), ("tuneLayout", tuneLayout), ("tuneBlocksPerSM", tuneBlocksPerSM),
                      ("tuneParkThreshold", tuneParkThreshold), ("tuneLeCarry", tuneLeCarry), ("tuneExtMask", tuneExtMask),
                      ("tuneBurst", tuneBurst), ("tuneLeap", tuneLeap), ("tuneLeapLanes", tuneLeapLanes)):
        if val is not None:
            setattr(o, name, int(val))
    if buildTablesOnDevice is not None:
        g.buildTablesOnDevice = bool(buildTablesOnDevice)
        g._stagedTables = None
    if intensityMus is not None:                                           # INT:1245-1271
        dirs = np.stack([makeDirectionCosines(m, f32(p) * Pi / f32(180.0)) for m, p in zip(mus, phis)], axis=0)
        g.intensityDirections = np.ascontiguousarray(dirs, dtype=f32)
        g.computeIntensity = True
        g._stagedTables = None
    if computeIntensity is not None and not computeIntensity and intensityMus is None:   # INT:1278-1284
        g.intensityDirections = None
        g.computeIntensity = False
    nDir = 0 if g.intensityDirections is None else g.intensityDirections.shape[0]
    g._check(g._lib.mcb_set_views(g._h, nDir, _lib.ptr(g.intensityDirections, C.c_float)), "specifyParameters")
    g._check(g._lib.mcb_set_options(g._h, C.byref(o)), "specifyParameters")


def _stage_domain(g: integrator, d: Domain) -> None:
    """What ``computeRT`` copies out of the domain every batch (INT:434-443) -- staged once."""
    if getattr(d, "deviceOwner", None) is not None:          # assembled in this integrator's HBM (read_SSPTable)
        key = ("device", d.token)
        if d.deviceOwner is not g or g._stagedDomain != key:
            raise McbError("computeRadiativeTransfer: the domain was assembled on another integrator or has been replaced")
    elif d.totalExt is None:
        d.getOpticalPropertiesByComponent()
    if getattr(d, "deviceOwner", None) is None:
        key = ("host", d.token, float(d.surfaceAlbedo))
    if g._stagedDomain != key:
        if (d.numX, d.numY, d.numZ) != (g.numX, g.numY, g.numZ):
            raise McbError("computeRadiativeTransfer: domain and integrator grids differ")
        nc = d.cumulativeExt.shape[0]
        g._check(g._lib.mcb_set_optics(g._h, nc, _lib.ptr(d.totalExt, C.c_double), _lib.ptr(d.cumulativeExt, C.c_double),
                                       _lib.ptr(d.ssa, C.c_double), _lib.ptr(d.phaseFunctionIndex, C.c_int32),
                                       float(d.surfaceAlbedo)), "computeRadiativeTransfer")
        g.numComps = nc
        g._stagedDomain = key
        g._stagedTables = None
    # INT:280-285: (re)tabulate only if missing or too coarse, then stage
    onDevice = g.buildTablesOnDevice
    if not onDevice:
        d.tabulateInversePhaseFunctions(g.minInverseTableSize)
    # forward tables: Legendre-stored components without the hybrid peak can be tabulated in HBM as well
    # (buildTablesOnDevice: every forward table -- Legendre moments or angle / value pairs, with or without the hybrid
    # peak -- is tabulated in HBM: mcb_build_forward_table_general)
    hybrid = bool(g.options.useHybridPhaseFunsForIntenCalcs)
    fwdOnDevice = [onDevice for tab in d.forwardTables] if g.computeIntensity else []
    if g.computeIntensity and not all(fwdOnDevice):
        d.tabulateForwardPhaseFunctions(g.minForwardTableSize, hybrid, g.hybridPhaseFunWidth)
    tkey = (key, onDevice, g.minInverseTableSize if onDevice else d.tableToken,
            (tuple(fwdOnDevice), g.minForwardTableSize, hybrid, g.hybridPhaseFunWidth, d.tableToken)
            if g.computeIntensity else None)
    if g._stagedTables != tkey:
        if onDevice:
            from .inversePhaseFunctions import inversion_inputs
            for c, tab in enumerate(d.forwardTables):
                if all(pf.storedAsLegendre() for pf in tab.phaseFunctions):        # Lobatto nodes and values in HBM too
                    nCoef = np.array([pf.legendreCoefficients.size for pf in tab.phaseFunctions], dtype=np.int32)
                    coefs = np.ascontiguousarray(np.concatenate([pf.legendreCoefficients for pf in tab.phaseFunctions]
                                                                + [np.zeros(0, f32)]), dtype=f32)
                    g._check(g._lib.mcb_build_inverse_table_legendre(g._h, c + 1, int(g.minInverseTableSize), len(tab.phaseFunctions),
                                                                     _lib.ptr(nCoef, C.c_int32), _lib.ptr(coefs, C.c_float)),
                             "tabulateInversePhaseFunctions")
                    continue
                pairs = [inversion_inputs(pf) for pf in tab.phaseFunctions]
                nAng = np.array([m.size for m, _ in pairs], dtype=np.int32)
                mus = np.ascontiguousarray(np.concatenate([m for m, _ in pairs]), dtype=f32)
                vals = np.ascontiguousarray(np.concatenate([v for _, v in pairs]), dtype=f32)
                g._check(g._lib.mcb_build_inverse_table(g._h, c + 1, int(g.minInverseTableSize), len(pairs),
                                                        _lib.ptr(nAng, C.c_int32), _lib.ptr(mus, C.c_float),
                                                        _lib.ptr(vals, C.c_float)), "tabulateInversePhaseFunctions")
        else:
            for c, T in enumerate(d.inversePhaseFunctions):
                g._check(g._lib.mcb_set_inverse_table(g._h, c + 1, T.shape[1], T.shape[0], _lib.ptr(T, C.c_float)),
                         "tabulateInversePhaseFunctions")
        if g.computeIntensity:
            for c, tab in enumerate(d.forwardTables):
                if fwdOnDevice[c]:
                    leg = [pf.storedAsLegendre() for pf in tab.phaseFunctions]
                    nCoef = np.array([pf.legendreCoefficients.size if l else 0 for pf, l in zip(tab.phaseFunctions, leg)], dtype=np.int32)
                    nAng = np.array([0 if l else pf.scatteringAngle.size for pf, l in zip(tab.phaseFunctions, leg)], dtype=np.int32)
                    z = [np.zeros(0, f32)]
                    coefs = np.ascontiguousarray(np.concatenate([pf.legendreCoefficients for pf, l in zip(tab.phaseFunctions, leg) if l] + z), dtype=f32)
                    angs = np.ascontiguousarray(np.concatenate([pf.scatteringAngle for pf, l in zip(tab.phaseFunctions, leg) if not l] + z), dtype=f32)
                    vals = np.ascontiguousarray(np.concatenate([pf.value for pf, l in zip(tab.phaseFunctions, leg) if not l] + z), dtype=f32)
                    g._check(g._lib.mcb_build_forward_table_general(
                        g._h, c + 1, int(g.minForwardTableSize), len(tab.phaseFunctions), _lib.ptr(nCoef, C.c_int32),
                        _lib.ptr(coefs, C.c_float), _lib.ptr(nAng, C.c_int32), _lib.ptr(angs, C.c_float), _lib.ptr(vals, C.c_float),
                        C.c_float(g.hybridPhaseFunWidth if hybrid else 0.0)), "tabulateForwardPhaseFunctions")
                else:
                    Pf, Po = d.tabulatedPhaseFunctions[c], d.tabulatedOrigPhaseFunctions[c]
                    g._check(g._lib.mcb_set_forward_table(g._h, c + 1, Pf.shape[1], Pf.shape[0], _lib.ptr(Pf, C.c_float),
                                                          _lib.ptr(Po, C.c_float)), "tabulateForwardPhaseFunctions")
        g._stagedTables = tkey


def _stage_source(g: integrator, photons: photonStream) -> None:
    if photons.kind == "directional":
        key = ("directional", photons.solarMu, photons.solarAzimuth)
        if g._stagedSource != key:
            g._check(g._lib.mcb_set_solar_source(g._h, photons.solarMu, photons.solarAzimuth), "new_PhotonStream")
    else:
        w = photons.weights
        if w.deviceOwner is not None:                      # built in this integrator's HBM (emission_weighting_device)
            key = ("bbemission-device", w.token)
            if w.deviceOwner is not g or g._stagedSource != key:
                raise McbError("new_PhotonStream: device-built weights belong to another integrator or were replaced")
            return
        key = ("bbemission", w.token, w.fracAtmsPower)
        if g._stagedSource != key:
            g._check(g._lib.mcb_set_thermal_source(g._h, float(w.fracAtmsPower), _lib.ptr(w.voxelWeights, C.c_double)),
                     "new_PhotonStream")
    g._stagedSource = key


def computeRadiativeTransfer(thisIntegrator: integrator, thisDomain: Domain, randomNumbers: randomNumberSequence,
                             incomingPhotons: photonStream, numPhotonsPerBatch: int, synchronize: bool = True) -> int:
    """``computeRadiativeTransfer`` (INT:209-391): one batch.  Returns ``numPhotonsProcessed``.

    Zeroes the tallies, makes sure the domain/tables/source are resident in HBM, launches the
    photon kernel for ``min(numPhotonsPerBatch, photons left in the stream)`` photons; the
    normalisation of INT:328-388 is applied when results are read (``reportResults``).
    """
    g = thisIntegrator
    if not isReady_Integrator(g):
        raise McbError("computeRadiativeTransfer: problem not completely specified.")
    _stage_domain(g, thisDomain)
    _stage_source(g, incomingPhotons)
    if not morePhotonsExist(incomingPhotons):
        raise McbError("computeRadiativeTransfer: Didn't process any photons.")
    left = incomingPhotons.numberOfPhotons - (incomingPhotons.currentPhoton - 1)
    n = int(min(int(numPhotonsPerBatch), left))
    first = incomingPhotons.firstPhotonId + (incomingPhotons.currentPhoton - 1)
    done = C.c_int64(0)
    g._check(g._lib.mcb_run_batch(g._h, n, C.c_uint64(randomNumbers.seed), C.c_uint64(first), C.byref(done)),
             "computeRadiativeTransfer")
    incomingPhotons.currentPhoton += n
    g.numPhotonsProcessed = int(done.value)
    if synchronize:
        g._check(g._lib.mcb_synchronize(g._h), "computeRadiativeTransfer")
    return int(done.value)


def reportResults(thisIntegrator: integrator, *, meanFluxUp=False, meanFluxDown=False, meanFluxAbsorbed=False,
                  fluxUp=False, fluxDown=False, fluxAbsorbed=False, absorbedProfile=False, volumeAbsorption=False,
                  meanIntensity=False, intensity=False, intensityByComponent=False,
                  numPhotonsForNormalisation: int = 0) -> Dict[str, np.ndarray]:
    """``reportResults`` (INT:845-1042): request outputs by keyword (Fortran ``optional``).

    Arrays come back in the reference's layout as NumPy arrays of shape ``(ny, nx)``,
    ``(nz, ny, nx)``, ``(nDir, ny, nx)``, ``(nc+1, nDir, ny, nx)``.  Means are plain
    ``sum()/numColumns`` in single precision (INT:881-884, 966, 990).
    """
    g = thisIntegrator
    nx, ny, nz = g.numX, g.numY, g.numZ
    nDir = 0 if g.intensityDirections is None else g.intensityDirections.shape[0]
    if (meanIntensity or intensity or intensityByComponent) and nDir == 0:
        raise McbError("reportResults: intensity information not available")
    want_up = meanFluxUp or fluxUp
    want_dn = meanFluxDown or fluxDown
    want_ab = meanFluxAbsorbed or fluxAbsorbed
    want_vol = absorbedProfile or volumeAbsorption
    want_int = meanIntensity or intensity
    a_up = np.empty((ny, nx), f32) if want_up else None
    a_dn = np.empty((ny, nx), f32) if want_dn else None
    a_ab = np.empty((ny, nx), f32) if want_ab else None
    a_vol = np.empty((nz, ny, nx), f32) if want_vol else None
    a_int = np.empty((nDir, ny, nx), f32) if want_int else None
    a_byc = np.empty((g.numComps + 1, nDir, ny, nx), f32) if intensityByComponent else None
    g._check(g._lib.mcb_get_results(g._h, int(numPhotonsForNormalisation), _lib.ptr(a_up, C.c_float),
                                    _lib.ptr(a_dn, C.c_float), _lib.ptr(a_ab, C.c_float), _lib.ptr(a_vol, C.c_float),
                                    _lib.ptr(a_int, C.c_float), _lib.ptr(a_byc, C.c_float)), "reportResults")
    numColumns = f32(nx * ny)
    out: Dict[str, np.ndarray] = {}
    if meanFluxUp: out["meanFluxUp"] = f32(a_up.sum(dtype=f32) / numColumns)
    if meanFluxDown: out["meanFluxDown"] = f32(a_dn.sum(dtype=f32) / numColumns)
    if meanFluxAbsorbed: out["meanFluxAbsorbed"] = f32(a_ab.sum(dtype=f32) / numColumns)
    if fluxUp: out["fluxUp"] = a_up
    if fluxDown: out["fluxDown"] = a_dn
    if fluxAbsorbed: out["fluxAbsorbed"] = a_ab
    if absorbedProfile: out["absorbedProfile"] = (a_vol.reshape(nz, -1).sum(axis=1, dtype=f32) / numColumns).astype(f32)
    if volumeAbsorption: out["volumeAbsorption"] = a_vol
    if meanIntensity: out["meanIntensity"] = (a_int.reshape(nDir, -1).sum(axis=1, dtype=f32) / numColumns).astype(f32)
    if intensity: out["intensity"] = a_int
    if intensityByComponent: out["intensityByComponent"] = a_byc
    return out


def getCounters(thisIntegrator: integrator) -> Dict[str, int]:
    """Event counters of the last batch (crossings, scatterings, ... -- algorithmic bytes)."""
    c = mcb_counters()
    thisIntegrator._check(thisIntegrator._lib.mcb_get_counters(thisIntegrator._h, C.byref(c)), "getCounters")
    return c.as_dict()


def gatherProbe(thisIntegrator: integrator, nbytes: int, loadsInFlight: int = 8, blocksPerSM: int = 8,
                iterations: int = 2000) -> float:
    """Measured ceiling of fully divergent sector gathers on this GPU (``csrc/mcb_probe.cu``): gathers per second
    inside a buffer of ``nbytes`` bytes, launched like the flux kernels.  The denominator of the gather roofline."""
    out = C.c_double(0.0)
    thisIntegrator._check(thisIntegrator._lib.mcb_debug_gather_probe(thisIntegrator.handle, int(nbytes), int(loadsInFlight),
                                                                     int(blocksPerSM), int(iterations), C.byref(out)),
                          "gatherProbe")
    return float(out.value)


def vacuumDistanceMap(thisIntegrator: integrator, domain) -> np.ndarray:
    """The (nz, ny, nx) map of Chebyshev distances to the nearest cell with extinction that the packed extinction
    field of ``domain`` is encoded with (``csrc/mcb_stage.cu``; the photon-pool kernels leap that far through vacuum)."""
    g = thisIntegrator
    _stage_domain(g, domain)
    out = np.zeros((domain.numZ, domain.numY, domain.numX), dtype=np.uint8)
    g._check(g._lib.mcb_debug_distance_map(g.handle, out.ctypes.data_as(C.POINTER(C.c_uint8)), out.size), "vacuumDistanceMap")
    return out


def lastBatchMilliseconds(thisIntegrator: integrator) -> float:
    ms = C.c_float(0)
    thisIntegrator._check(thisIntegrator._lib.mcb_last_batch_ms(thisIntegrator._h, C.byref(ms)), "lastBatchMilliseconds")
    return float(ms.value)


def tracePhotons(thisIntegrator: integrator, thisDomain: Domain, incomingPhotons: photonStream,
                 randomReals: np.ndarray, maxEventsPerPhoton: int = 256):
    """Fixed-random-number single-photon trace harness (north-star criterion (a)).

    ``randomReals`` has shape ``(nPhotons, stride)``: photon p is born from and transported
    with row p.  Returns ``(events, rawTallies)``; events is a structured array
    (``_lib.EVENT_DTYPE``) in photon order.  Always runs the reference-arithmetic kernel.
    """
    g = thisIntegrator
    _stage_domain(g, thisDomain)
    _stage_source(g, incomingPhotons)
    rn = np.ascontiguousarray(randomReals, dtype=f32)
    n, stride = rn.shape
    cap = int(n) * int(maxEventsPerPhoton)
    ev = np.zeros(cap, dtype=_lib.EVENT_DTYPE)
    nEv = C.c_int64(0)
    g._check(g._lib.mcb_run_trace(g._h, n, _lib.ptr(rn, C.c_float), stride, int(maxEventsPerPhoton),
                                  ev.ctypes.data_as(C.c_void_p), cap, C.byref(nEv)), "tracePhotons")
    dptr = C.c_void_p(); nd = C.c_int64(0)
    g._check(g._lib.mcb_tally_buffer(g._h, C.byref(dptr), C.byref(nd)), "tracePhotons")
    raw = np.empty(nd.value, dtype=np.float64)
    g._check(g._lib.mcb_get_raw_tallies(g._h, _lib.ptr(raw, C.c_double), nd.value), "tracePhotons")
    return ev[: min(nEv.value, cap)], raw
