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


def scene(kind="irregular", albedo=0.25, ssaScale=1.0, tiles=(1, 1), seed=3, temps=None, lambda_um=0.0):
    """(Domain, Medium of ONE tile).  ``tiles`` repeats the scene periodically in x and y — the same physics on a
    larger grid (results fold back onto the tile), which is how the kernels for wide grids see this scene."""
    nx, ny, nz = NX, NY, NZ
    if kind == "regular":
        xE = 0.0625 * np.arange(nx + 1); yE = 0.03125 * np.arange(ny + 1); zE = 0.03125 * np.arange(nz + 1)
    elif kind == "irregular":
        xE = 0.05 * np.arange(nx + 1); yE = 0.03 * np.arange(ny + 1); zE = 0.04 * np.arange(nz + 1)
    elif kind == "zstretched":                   # uniform columns, stretched layers (the usual LES grid)
        xE = 0.05 * np.arange(nx + 1); yE = 0.03 * np.arange(ny + 1)
        zE = np.concatenate([[0.0], np.cumsum(0.025 * 1.15 ** np.arange(nz))])
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
        temps = None if temps is None else np.tile(temps, (1, ty, tx))
    d = Domain(xE, yE, zE, temps=temps, surfaceAlbedo=albedo, lambda_um=lambda_um)
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


# ---- the first interaction taken out of an event trace (oracle: orc.trace; CUDA: tracePhotons -- the same record) -------
EV_BIRTH, EV_SCATTER, EV_SURFACE, EV_KILLED_SURFACE, EV_KILLED_ROULETTE, EV_LE = 1, 2, 3, 5, 6, 8


def first_interaction_of_trace(ev, n, nDir, med):
    """(first-collision counts per cell, uncollided surface arrivals per column, sum of first-order local-estimate
    contributions (nDir, ncol), same for the surface-reflected direct beam) of n traced photons."""
    ncol = med.nx * med.ny
    le = ev["kind"] == EV_LE
    major = ~le & (ev["kind"] != EV_BIRTH)
    c = np.cumsum(major)
    base = np.zeros(len(ev), np.int64)
    birth = ev["kind"] == EV_BIRTH
    base[birth] = c[birth]
    k = c - np.maximum.accumulate(base)                  # major events of this photon so far (events are in photon order)
    fe = ev[major & (k == 1)]                            # every photon's first event after its birth
    coll = (fe["kind"] == EV_SCATTER) | (fe["kind"] == EV_KILLED_ROULETTE)
    surf = (fe["kind"] == EV_SURFACE) | (fe["kind"] == EV_KILLED_SURFACE)
    assert coll.sum() + surf.sum() == n, np.unique(fe["kind"], return_counts=True)
    cell = (fe["ix"] - 1) + med.nx * ((fe["iy"] - 1) + med.ny * (fe["iz"] - 1))
    cells = np.bincount(cell[coll], minlength=med.sigma.size)
    cols = np.bincount(((fe["ix"] - 1) + med.nx * (fe["iy"] - 1))[surf], minlength=ncol)
    firstKind = np.zeros(n, np.int32)
    firstKind[fe["photon"]] = fe["kind"]
    out = []
    # local-estimate events precede their scattering event and follow their surface event (INT:681-700, 776-790)
    for sel in (le & (k == 0), le & (k == 1) & (ev["order"] == 1) & (firstKind[ev["photon"]] == EV_SURFACE)):
        e = ev[sel]
        idx = (e["component"] - 1) * ncol + (e["ix"] - 1) + med.nx * (e["iy"] - 1)
        out.append(np.bincount(idx, weights=e["weight"].astype(np.float64), minlength=nDir * ncol).reshape(nDir, ncol))
    return cells, cols, out[0], out[1]


def z_stats(z):
    return float(np.sqrt(np.mean(z ** 2))), float(z.mean()), float(np.abs(z).max())


def check_first_interaction(kind, albedo, rr, cells, cols, L1, L0, N):
    """cells / cols: first-collision and uncollided-arrival COUNTS of N photons; L1 / L0 (B, nDir, ncol): per-batch mean
    first-order / surface-reflected local-estimate contribution per photon.  Compared with the committed answers."""
    fx = fixture(kind)
    B = L1.shape[0]
    # where the first collision happens: multinomial counts against Beer's law along the slant paths
    p = fx["first"].ravel()
    assert cells[p == 0].sum() == 0                         # nothing collides in empty cells
    ok = N * p > 25
    rms, mean, worst = z_stats((cells - N * p)[ok] / np.sqrt(N * p * (1 - p))[ok])
    print("%s rr %d: first collisions over %d cells: z rms %.3f mean %+.3f max %.2f" % (kind, rr, ok.sum(), rms, mean, worst))
    assert ok.sum() > 250 and rms < 1.12 and abs(mean) < 4.0 / np.sqrt(ok.sum()) and worst < 4.8, (kind, rms, mean, worst)
    for axis, what in (((1, 2), "layers"), ((0, 2), "rows"), ((0, 1), "x-slabs")):   # aggregated: sigma ~ 0.3 % relative
        q = fx["first"].sum(axis=axis)
        zz = (cells.reshape(fx["first"].shape).sum(axis=axis) - N * q)[q > 0] / np.sqrt(N * q * (1 - q))[q > 0]
        assert np.abs(zz).max() < 4.0, (kind, what, zz)
    # where the uncollided beam lands
    q = fx["surf"].ravel()
    rms, mean, worst = z_stats((cols - N * q) / np.sqrt(N * q * (1 - q)))
    assert rms < 1.35 and abs(mean) < 0.6 and worst < 4.5, (kind, "surface", rms, mean, worst)
    assert abs(cols.sum() - N * q.sum()) < 4.0 * np.sqrt(N * q.sum())

    # first-order radiance and the surface-reflected direct beam, per view direction and exit column
    for i, mu in enumerate(VIEW_MUS):
        for what, L, E in (("E1", L1[:, i], fx["E1"][i].ravel()), ("E0", L0[:, i], albedo * fx["E0"][i].ravel())):
            if rr and mu < 0:                               # the roulette form only counts rays that reach the TOP (INT:1768-1800)
                assert L.sum() == 0.0
                continue
            m, se = L.mean(axis=0), L.std(axis=0, ddof=1) / np.sqrt(B)
            tot = L.sum(axis=1)
            zt = (tot.mean() - E.sum()) / (tot.std(ddof=1) / np.sqrt(B))
            print("%s rr %d view %d %s: total %.6g vs %.6g (rel %+.2e, z %+.2f)" % (kind, rr, i, what, tot.mean(), E.sum(), tot.mean() / E.sum() - 1, zt))
            assert abs(zt) < 4.2, (kind, what, i, tot.mean(), E.sum(), zt)
            if what == "E1":                                # (surface arrivals are too few per column and batch for a z-map)
                rms, mean, worst = z_stats((m - E) / se)    # Student t, 19 degrees of freedom: rms 1.06 if unbiased
                assert rms < 1.45 and abs(mean) < 0.65 and worst < 6.5, (kind, what, i, rms, mean, worst)
                assert abs(tot.mean() / E.sum() - 1.0) < 0.012


def check_absorber(name, first, surf, fx):
    """Pure absorber: per-batch first-collision probabilities per cell (NB, nz, ny, nx) and uncollided arrivals per
    column (NB, ny, nx) as the tallies give them, against the committed answers."""
    NB = first.shape[0]
    # quad: the fixture's own quadrature error relative to the largest entry (make_first_interaction.py prints it)
    for what, got, want, minP, quad in (("cells", first, fx["first"], 2e-4, 3e-5), ("surface", surf, fx["surf"], 0.0, 1.5e-4)):
        m, se = got.mean(axis=0), got.std(axis=0, ddof=1) / np.sqrt(NB)
        ok = want > minP
        z = (m - want)[ok] / np.sqrt(se[ok] ** 2 + (quad * want.max()) ** 2)
        rms, mean, worst = z_stats(z)                       # Student t, 15 degrees of freedom: rms 1.07 if unbiased
        print("%s %s: n %d rms %.3f mean %+.3f max %.2f" % (name, what, z.size, rms, mean, worst))
        assert rms < (1.25 if what == "cells" else 1.45) and abs(mean) < 4.5 / np.sqrt(z.size) and worst < 6.0, (name, what, rms, mean, worst)
    # aggregated (sigma ~ 2e-4 relative): the absorption profile and the total transmission
    layers, want = first.sum(axis=(2, 3)), fx["first"].sum(axis=(1, 2))
    z = (layers.mean(axis=0) - want) / np.sqrt(layers.var(axis=0, ddof=1) / NB + (3e-5 * want.max()) ** 2)
    print("%s layers: z %s rel %s" % (name, np.round(z, 2), np.round(layers.mean(axis=0) / want - 1, 5)))
    assert np.abs(z).max() < 4.5, (name, z)
    t = surf.sum(axis=(1, 2))
    assert abs(t.mean() - fx["surf"].sum()) < 4.5 * t.std(ddof=1) / np.sqrt(NB) + 1.5e-4 * fx["surf"].sum(), (name, t.mean(), fx["surf"].sum())
