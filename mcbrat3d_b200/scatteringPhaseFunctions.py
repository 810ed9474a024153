"""Host-side mirror of ``src/scatteringPhaseFunctions.f95`` (setup-time only).

``phaseFunction`` / ``phaseFunctionTable`` objects (SPF:32-55) stored either as Legendre
moments chi_l (l = 1.., P0 == 1 implied; value = sum (2l+1) chi_l P_l) or as angle/value
pairs, and ``getPhaseFunctionValues`` (SPF:448-650).  The photon kernels never see these
objects: they reach the device only as the tabulated matrices built in
``inversePhaseFunctions`` and ``opticalProperties``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .numericUtilities import computeLegendrePolynomials, findIndex

f32 = np.float32
Pi = f32(3.14159265358979312)          # SPF module constant (default real)


def _cos32(a):
    return np.cos(np.asarray(a, dtype=np.float64)).astype(f32)


@dataclass
class phaseFunction:
    """SPF:32-44.  Exactly one of (scatteringAngle, value) / legendreCoefficients is set."""
    scatteringAngle: Optional[np.ndarray] = None
    value: Optional[np.ndarray] = None
    legendreCoefficients: Optional[np.ndarray] = None
    extinction: float = 1.0
    singleScatteringAlbedo: float = 1.0
    description: str = ""

    def storedAsLegendre(self) -> bool:
        return self.legendreCoefficients is not None


def new_PhaseFunction(legendreCoefficients=None, scatteringAngle=None, value=None,
                      extinction=1.0, singleScatteringAlbedo=1.0, description="") -> phaseFunction:
    """``new_PhaseFunction`` (SPF:101-227): Legendre moments starting at P1, or angle/value pairs."""
    if legendreCoefficients is not None:
        lc = np.asarray(legendreCoefficients, dtype=f32)
        if lc.size > 1 and (lc[0] > 1 or lc[0] < -1):
            raise ValueError("newPhaseFunction: Asymmetery parameter out of bounds.")
        return phaseFunction(legendreCoefficients=lc, extinction=extinction,
                             singleScatteringAlbedo=singleScatteringAlbedo, description=description)
    ang = np.asarray(scatteringAngle, dtype=f32)
    val = np.asarray(value, dtype=f32)
    if ang.size != val.size:
        raise ValueError("newPhaseFunction: Number of scattering angles and phase function values must match.")
    if np.any(np.diff(ang) <= 0):
        raise ValueError("newPhaseFunction: Scattering angle must be increasing, unique.")
    if np.any(val < 0):
        raise ValueError("newPhaseFunction: Negative phase function values supplied.")
    return phaseFunction(scatteringAngle=ang, value=_normalize(ang, val), extinction=extinction,
                         singleScatteringAlbedo=singleScatteringAlbedo, description=description)


def _normalize(scatteringAngle, values):
    """``normalizePhaseFunction`` (SPF:1520-1536): integral over mu equals 2."""
    mu = _cos32(scatteringAngle)
    denom = f32(0)
    for k in range(len(values) - 1):
        denom = f32(denom + (mu[k + 1] - mu[k]) * (f32(0.5) * (values[k + 1] + values[k])))
    return (-values * f32(2.0) / denom).astype(f32)


@dataclass
class phaseFunctionTable:
    """SPF:46-55: a keyed series of phase functions."""
    phaseFunctions: List[phaseFunction] = field(default_factory=list)
    key: Optional[np.ndarray] = None
    description: str = ""

    @property
    def nEntries(self) -> int:
        return len(self.phaseFunctions)


def new_PhaseFunctionTable(phaseFunctions: Sequence[phaseFunction], key, tableDescription="") -> phaseFunctionTable:
    key = np.asarray(key, dtype=f32)
    if key.size != len(phaseFunctions):
        raise ValueError("newPhaseFunctionTable: Number of phase functions and key values must match.")
    if np.any(np.diff(key) <= 0):
        raise ValueError("newPhaseFunctionTable: Key values must be unique, increasing.")
    return phaseFunctionTable(list(phaseFunctions), key, tableDescription)


def getPhaseFunctionValues(pf, scatteringAngle) -> np.ndarray:
    """``getPhaseFunctionValues`` (one: SPF:448-531; table: SPF:533-650).

    Returns ``value(nAngles)`` for a phaseFunction or ``values(nAngles, nEntries)`` (Fortran
    order: angle fastest) for a phaseFunctionTable.
    """
    if isinstance(pf, phaseFunctionTable):
        cols = [getPhaseFunctionValues(p, scatteringAngle) for p in pf.phaseFunctions]
        return np.stack(cols, axis=1).astype(f32)
    ang = np.asarray(scatteringAngle, dtype=f32)
    if pf.storedAsLegendre():
        maxL = pf.legendreCoefficients.size
        if maxL == 0:                                # isotropic, SPF:486-491 (value 1/2, quirk q14)
            return np.full(ang.shape, f32(0.5), dtype=f32)
        P = computeLegendrePolynomials(maxL, _cos32(ang))
        # (/ 1., chi(:) /) * (/ (2l+1) /) evaluated in single precision
        coef = (np.concatenate(([f32(1.0)], pf.legendreCoefficients)).astype(f32) *
                (2 * np.arange(0, maxL + 1) + 1).astype(f32)).astype(f32)
        value = np.zeros(ang.shape, dtype=f32)
        for l in range(maxL + 1):                    # matmul, accumulated in order
            value = (value + coef[l] * P[l]).astype(f32)
        return value
    # tabulated: interpolate linearly in cos(angle), SPF:499-527
    nStored = pf.scatteringAngle.size
    idx = np.array([findIndex(a, pf.scatteringAngle) for a in ang], dtype=np.int64)
    idx = np.clip(idx, 1, nStored)
    ip1 = np.where(idx < nStored, idx + 1, idx)
    cs = _cos32(pf.scatteringAngle)
    dMu = np.where(idx < nStored, cs[ip1 - 1] - cs[idx - 1], np.finfo(f32).max).astype(f32)
    w = (f32(1.0) - (_cos32(ang) - cs[idx - 1]) / dMu).astype(f32)
    return (w * pf.value[idx - 1] + (f32(1.0) - w) * pf.value[ip1 - 1]).astype(f32)


def henyeyGreenstein(g: float, nLegendreCoefficients: int) -> phaseFunction:
    """HG phase function as the reference's generators build it: chi_l = g**l, l = 1..n
    (``Domain-Files/i3rcStepCloud.f95:55``, ``i3rcLandsatCloud.f95:57``)."""
    l = np.arange(1, nLegendreCoefficients + 1)
    return new_PhaseFunction(legendreCoefficients=(f32(g) ** l.astype(f32)).astype(f32),
                             description="Henyey-Greenstein g=%g" % g)


def rayleigh() -> phaseFunction:
    """Rayleigh phase function, ``calc_RayleighScattering`` OPT:2080-2081: LG = (0, 0.5)/(3, 5)."""
    return new_PhaseFunction(legendreCoefficients=np.array([0.0, 0.5], dtype=f32) / np.array([3.0, 5.0], dtype=f32),
                             description="Rayleigh Scattering")
