"""Host-side mirror of ``src/numericUtilities.f95`` (setup-time only, never on the photon path).

Table searches (``findIndex`` NUM:206-315, ``findCDFIndex`` NUM:317-348) and the Lobatto /
Legendre helpers (``computeLobattoTerms`` NUM:27-114, ``computeLegendrePolynomials``
NUM:187-205) that the phase-table builders use.  Single precision is kept where the
reference uses default ``real``.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def spacing32(x):
    """Fortran ``spacing`` for default reals (``spacing(0) = tiny``)."""
    x = np.abs(np.asarray(x, dtype=f32))
    s = np.nextafter(x, f32(np.inf), dtype=f32) - x
    tiny = np.finfo(f32).tiny
    return np.where((x == 0) | (s < tiny), tiny, s).astype(f32)


def spacing64(x):
    """Fortran ``spacing`` for ``real(8)``."""
    x = np.abs(np.asarray(x, dtype=np.float64))
    s = np.nextafter(x, np.inf) - x
    tiny = np.finfo(np.float64).tiny
    return np.where((x == 0) | (s < tiny), tiny, s)


def findIndex(value, table, firstGuess=None) -> int:
    """``findIndex`` (NUM:206-315): 1-based i with table(i) <= value < table(i+1).

    Hunt from ``firstGuess`` when given, then bisection.  As in the reference: 0 below the table
    (without a guess; with one the hunt never ends there), ``size(table)`` at or beyond its end
    with a guess, ``size(table) - 1`` without.
    """
    n = len(table)
    if firstGuess is not None:
        lower = int(firstGuess)
        inc = 1
        while True:
            upper = min(lower + inc, n)
            if lower == n or (table[lower - 1] <= value and table[upper - 1] > value):
                break
            if table[lower - 1] > value:
                upper = lower
                lower = max(upper - inc, 1)
            else:
                lower = upper
            inc *= 2
    else:
        lower, upper = 0, n
    while True:
        if lower == n or upper <= lower + 1:
            break
        mid = (lower + upper) // 2
        if value >= table[mid - 1]:
            lower = mid
        else:
            upper = mid
    return lower


def findCDFIndex(value, table) -> int:
    """``findCDFIndex`` (NUM:317-348): 1-based i with table(i-1) < value <= table(i)."""
    n = len(table)
    lower, upper = 0, n
    while True:
        if lower == n or upper <= lower + 1:
            break
        mid = (lower + upper) // 2
        if value > table[mid - 1]:
            lower = mid
        else:
            upper = mid
    return upper


def computeLegendrePolynomials(maxL: int, mus) -> np.ndarray:
    """P_0..P_maxL at ``mus`` by upward recursion in single precision (NUM:187-205)."""
    mus = np.asarray(mus, dtype=f32)
    P = np.empty((max(maxL, 1) + 1, mus.size), dtype=f32)
    P[0] = f32(1)
    P[1] = mus
    for l in range(1, maxL):
        P[l + 1] = ((f32(2 * l + 1) * mus) * P[l] - f32(l) * P[l - 1]) / f32(l + 1)
    return P[: maxL + 1]


_LOBATTO_CACHE = {}


def computeLobattoTerms(n: int):
    """Lobatto abscissas and weights on [-1, 1] by Newton iteration (NUM:27-114).  The result depends on ``n`` only, so
    it is kept: a many-wavelength run asks for the same few node counts at every wavelength and phase function."""
    n = int(n)
    if n not in _LOBATTO_CACHE:
        mus, weights = _computeLobattoTerms(n)
        mus.setflags(write=False); weights.setflags(write=False)
        _LOBATTO_CACHE[n] = (mus, weights)
    return _LOBATTO_CACHE[n]


def _computeLobattoTerms(n: int):
    relativeAccuracy = f32(3.0)
    maxIterations = 25
    pi = f32(np.arccos(np.float64(-1.0)))
    nTerms = n
    midPoint = (nTerms + 1) // 2
    mus = np.zeros(n, dtype=f32)
    weights = np.zeros(n, dtype=f32)
    m = midPoint - 1
    c1 = f32(1.0) if nTerms % 2 == 1 else f32(0.5)
    i = np.arange(1, m + 1, dtype=f32)
    trial = np.sin((pi * (i - c1) / f32(nTerms - 1.0 + 0.5)).astype(np.float64)).astype(f32)

    def newton(trial):
        P = computeLegendrePolynomials(nTerms - 1, trial)
        d1 = f32(nTerms - 1) * (trial * P[nTerms - 1] - P[nTerms - 2]) / (trial * trial - f32(1.0))
        d2 = (f32(2.0) * trial * d1 - f32(nTerms * (nTerms - 1)) * P[nTerms - 1]) / (f32(1.0) - trial * trial)
        return P, d1, d2

    if m > 0:
        P, d1, d2 = newton(trial)
        last = trial.copy()
        trial = (trial - d1 / d2).astype(f32)
        it = 0
        while True:
            moving = np.abs(trial - last) > relativeAccuracy * spacing32(trial)
            if not moving.any():
                break
            P, d1n, d2n = newton(trial)
            d1 = np.where(moving, d1n, d1)
            d2 = np.where(moving, d2n, d2)
            last = np.where(moving, trial, last)
            trial = np.where(moving, (trial - d1n / d2n).astype(f32), trial)
            it += 1
            if it > maxIterations:
                break
    mus[0] = f32(-1)
    weights[0] = f32(2.0) / f32(nTerms * (nTerms - 1))
    if m > 0:
        # mus(midPoint:2:-1) = -trialMus(:)
        mus[1:midPoint] = (-trial)[::-1]
        weights[1:midPoint] = (f32(2.0) / (f32(nTerms * (nTerms - 1)) * P[nTerms - 1] ** 2))[::-1]
    if nTerms % 2 == 0:
        mus[midPoint:nTerms] = -mus[:midPoint][::-1]
        weights[midPoint:nTerms] = weights[:midPoint][::-1]
    else:
        mus[midPoint - 1:nTerms] = -mus[:midPoint][::-1]
        weights[midPoint - 1:nTerms] = weights[:midPoint][::-1]
    return mus, weights
