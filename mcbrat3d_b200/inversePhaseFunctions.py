"""Host-side mirror of ``src/inversePhaseFunctions.f95`` (setup-time only).

``computeInversePhaseFuncTable`` (INV:26-64) builds, for every entry of a phase-function
table, the scattering angle as a function of the cumulative probability at ``nSteps``
equally spaced probabilities (INV:66-174): trapezoid CDF in mu on the native angles or on
``max(nMoments, 2)`` Lobatto nodes, normalised, then the analytic inversion of the locally
linear phase function.  ``T(1) = pi`` (backscatter) ... ``T(nSteps) = 0`` (forward).

The result is what ``mcb_set_inverse_table`` stages into HBM.
"""
from __future__ import annotations

import numpy as np

from .numericUtilities import computeLobattoTerms, findIndex, spacing32
from .scatteringPhaseFunctions import getPhaseFunctionValues, phaseFunction, phaseFunctionTable

f32 = np.float32


def _acos32(x):
    return np.arccos(np.asarray(x, dtype=np.float64)).astype(f32)


def find_cdf_brackets(cdf: np.ndarray, nSteps: int) -> np.ndarray:
    """INV:131-135: ``indicies(i) = findIndex((i-1)/(nSteps-1), cdf, firstGuess = indicies(i-1))``.

    On a non-decreasing table the hunt/bisection of NUM:206-260 returns the unique index with
    ``cdf(index) <= p < cdf(index+1)`` whatever the first guess, which is one vectorised search; a table that
    decreases somewhere (a truncated Legendre series can go negative) is searched with the reference's own
    sequence of guesses."""
    if np.all(np.diff(cdf) >= 0):
        p = np.arange(nSteps, dtype=f32) / f32(nSteps - 1)
        return np.minimum(np.searchsorted(cdf, p, side="right"), cdf.size).astype(np.int64)
    indicies = np.empty(nSteps, dtype=np.int64)
    indicies[0] = findIndex(f32(0.0), cdf)
    for i in range(2, nSteps + 1):
        p = f32(f32(i - 1) / f32(nSteps - 1))
        indicies[i - 1] = findIndex(p, cdf, firstGuess=int(indicies[i - 2]))
    return indicies


def inversion_inputs(thisPhaseFunction: phaseFunction):
    """The (mus, values) pairs the inversion works on (INV:87-112): native angles reversed to increasing mu, or
    the phase function evaluated at ``max(nMoments, 2)`` Lobatto nodes."""
    if not thisPhaseFunction.storedAsLegendre():
        angles = thisPhaseFunction.scatteringAngle
        values = getPhaseFunctionValues(thisPhaseFunction, angles)
        mus = np.cos(angles[::-1].astype(np.float64)).astype(f32)
        return mus, np.ascontiguousarray(values[::-1], dtype=f32)
    nAngles = max(thisPhaseFunction.legendreCoefficients.size, 2)
    mus, _ = computeLobattoTerms(nAngles)
    values = getPhaseFunctionValues(thisPhaseFunction, _acos32(mus[::-1]))
    return np.ascontiguousarray(mus, dtype=f32), np.ascontiguousarray(values[::-1], dtype=f32)


def computeInversePhaseFunction(thisPhaseFunction: phaseFunction, nSteps: int) -> np.ndarray:
    """INV:66-174.  Returns ``inverseTable(nSteps)`` in single precision."""
    if not thisPhaseFunction.storedAsLegendre():
        angles = thisPhaseFunction.scatteringAngle
        nAngles = angles.size
        values = getPhaseFunctionValues(thisPhaseFunction, angles)
        mus = np.cos(angles[::-1].astype(np.float64)).astype(f32)
        values = values[::-1].copy()
    else:
        nMoments = thisPhaseFunction.legendreCoefficients.size
        nAngles = max(nMoments, 2)
        mus, _ = computeLobattoTerms(nAngles)
        values = getPhaseFunctionValues(thisPhaseFunction, _acos32(mus[::-1]))
        values = values[::-1].copy()

    cdf = np.zeros(nAngles, dtype=f32)
    for i in range(1, nAngles):                                   # INV:118-121
        cdf[i] = f32(cdf[i - 1] + (mus[i] - mus[i - 1]) * f32(0.5) * (values[i] + values[i - 1]))
    cdf = (cdf / cdf[nAngles - 1]).astype(f32)

    indicies = find_cdf_brackets(cdf, nSteps)

    out = np.zeros(nSteps, dtype=f32)
    i = np.arange(1, nSteps, dtype=np.int64)                      # INV:137-167, vectorised over i
    p = (i - 1).astype(f32) / f32(nSteps - 1)
    idx = indicies[:nSteps - 1]                                   # 1-based
    c0, c1 = cdf[idx - 1], cdf[idx]
    v0, v1 = values[idx - 1], values[idx]
    m0, m1 = mus[idx - 1], mus[idx]
    flat_cdf = (c1 - c0) <= spacing32(c0)
    flat_val = np.abs(v0 - v1) <= spacing32(v0)
    with np.errstate(all="ignore"):
        lin = m0 + (m1 - m0) * (p - c0) / (c1 - c0)
        rad = ((c1 - p) * v0 ** 2 + (p - c0) * v1 ** 2) / (c1 - c0)
        gen = m0 + (m1 - m0) / (v0 - v1) * (v0 - np.sqrt(rad.astype(f32)))
    arg = np.where(flat_cdf, m0, np.where(flat_val, lin, gen)).astype(f32)
    out[:nSteps - 1] = _acos32(np.clip(arg, f32(-1.0), f32(1.0)))
    out[nSteps - 1] = f32(0.0)                                    # INV:168
    return out


def computeInversePhaseFuncTable(forwardTable: phaseFunctionTable, nSteps: int) -> np.ndarray:
    """INV:26-64.  Returns ``inverseTable(nSteps, nEntries)`` as a C array of shape
    ``(nEntries, nSteps)`` (i.e. Fortran ``values(step, entry)``, step fastest)."""
    rows = [computeInversePhaseFunction(pf, nSteps) for pf in forwardTable.phaseFunctions]
    return np.ascontiguousarray(np.stack(rows, axis=0), dtype=f32)
