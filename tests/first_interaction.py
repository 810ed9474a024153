"""Scenes and fixtures for the first-interaction tests (tests/independent_3d.py has the solver).

One small heterogeneous, periodic scene in three grid flavours — spacings exactly representable in single precision (the
reference's regular path, quirk q1), not representable (its irregular path: findIndex hunts, q2 / q3), and genuinely
stretched in x, y and z — built from RAW arrays twice: as a ``Medium`` for the deterministic solver and through the
package's ``Domain`` (the mirror of the reference's opticalProperties API) for the oracle and the CUDA library.
Two components: a cloud with lognormal extinction (an empty layer, 15 % empty cells, two Henyey-Greenstein entries,
per-cell single-scattering albedo 0.6-1) and a horizontally uniform Rayleigh "gas" over layers 2-5 (albedo 0.4)."""
import os

import numpy as np

from independent_3d import Medium
from mcbrat3d_b200.opticalProperties import Domain
from mcbrat3d_b200.scatteringPhaseFunctions import new_PhaseFunction, new_PhaseFunctionTable

KINDS = ("regular", "irregular", "stretched")
SOLAR_MU, SOLAR_AZIMUTH = 0.6, 30.0
VIEW_MUS, VIEW_PHIS = [1.0, 0.7, 0.4, -0.5], [0.0, 45.0, 200.0, 120.0]
NX, NY, NZ = 8, 6, 7
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def scene(kind="irregular", albedo=0.25, ssaScale=1.0, tiles=(1, 1), seed=3):
    """(Domain, Medium of ONE tile).  ``tiles`` repeats the scene periodically in x and y — the same physics on a
    larger grid (results fold back onto the tile), which is how the kernels for wide grids see this scene."""
    nx, ny, nz = NX, NY, NZ
    if kind == "regular":
        xE = 0.0625 * np.arange(nx + 1); yE = 0.03125 * np.arange(ny + 1); zE = 0.03125 * np.arange(nz + 1)
    elif kind == "irregular":
        xE = 0.05 * np.arange(nx + 1); yE = 0.03 * np.arange(ny + 1); zE = 0.04 * np.arange(nz + 1)
    elif kind == "stretched":
        xE = np.concatenate([[0.0], np.cumsum(0.05 * (1 + 0.3 * np.sin(1.0 + np.arange(nx))))])
        yE = np.concatenate([[0.0], np.cumsum(0.03 * (1 + 0.25 * np.cos(0.5 + np.arange(ny))))])
        zE = np.concatenate([[0.0], np.cumsum(0.025 * 1.15 ** np.arange(nz))])
    else:
        raise ValueError(kind)
    rng = np.random.default_rng(seed)
    ext1 = np.exp(rng.normal(2.2, 0.8, size=(nz, ny, nx)))
    ext1[3] = 0.0
    ext1[rng.random(ext1.shape) < 0.15] = 0.0
    ssa1 = rng.uniform(0.6, 1.0, size=ext1.shape) * ssaScale
    idx1 = np.where(ext1 > 12.0, 2, 1).astype(np.int32)
    idx1[ext1 == 0] = 0
    ssa1[ext1 == 0] = 0.0
    hg = [0.85 ** np.arange(1, 49), 0.5 ** np.arange(1, 25)]             # Legendre coefficients chi_l = g**l
    profile = np.array([3.0, 2.0, 1.5, 1.0])
    ext2 = np.zeros_like(ext1)
    ext2[1:5] = profile[:, None, None]
    ssa2 = np.where(ext2 > 0, 0.4 * ssaScale, 0.0)
    idx2 = (ext2 > 0).astype(np.int32)
    rayleigh = np.array([0.0, 0.1])                                      # 1 + 5 * 0.1 * P2 = 3/4 (1 + cos^2)
    med = Medium(xE, yE, zE, np.stack([ext1, ext2]), np.stack([ssa1, ssa2]), np.stack([idx1, idx2]), [hg, [rayleigh]], albedo)
    tx, ty = tiles
    if (tx, ty) != (1, 1):
        if kind == "stretched":
            xE = np.concatenate([[0.0], np.cumsum(np.tile(np.diff(xE), tx))])
            yE = np.concatenate([[0.0], np.cumsum(np.tile(np.diff(yE), ty))])
        else:
            xE = (xE[1] - xE[0]) * np.arange(nx * tx + 1); yE = (yE[1] - yE[0]) * np.arange(ny * ty + 1)
        ext1, ssa1, idx1 = (np.tile(a, (1, ty, tx)) for a in (ext1, ssa1, idx1))
    d = Domain(xE, yE, zE, surfaceAlbedo=albedo)
    d.addOpticalComponent("cloud", ext1, ssa1, idx1,
                          new_PhaseFunctionTable([new_PhaseFunction(legendreCoefficients=c) for c in hg], key=[1.0, 2.0]))
    d.addOpticalComponent("gas", profile, np.full(4, 0.4 * ssaScale), np.ones(4, np.int32),
                          new_PhaseFunctionTable([new_PhaseFunction(legendreCoefficients=rayleigh)], key=[0.0]), zLevelBase=2)
    d.getOpticalPropertiesByComponent()
    return d, med


def fold(a, tiles):
    """Sum an array whose last two axes are (ny * ty, nx * tx) over the tiles -> (..., ny, nx)."""
    tx, ty = tiles
    s = a.shape[:-2]
    return a.reshape(s + (ty, NY, tx, NX)).sum(axis=(-4, -2))


def fixture(kind):
    """The deterministic answers computed at high resolution by tests/golden/make_first_interaction.py:
    first (nz, ny, nx), surf (ny, nx): probabilities per photon; E1 (nDir, ny, nx) at ssaScale = 1; E0 (nDir, ny, nx) per
    unit surface albedo."""
    return np.load(os.path.join(GOLDEN, "first_interaction_%s.npz" % kind))
