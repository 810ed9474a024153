"""The oracle against an INDEPENDENT deterministic solver in a heterogeneous, periodic 3-D scene (tests/independent_3d.py).

The adding-doubling pins (tests/adding_doubling.py) are plane-parallel; this one is three-dimensional.  For a collimated
beam the first interaction has a closed form in any voxel medium: where the first collision happens (Beer's law along
slant paths through the periodically continued field), where the uncollided beam lands, and what first-order scattering
and the surface-reflected direct beam contribute to the radiance in every view direction and exit column.  The oracle's
event trace carries the scattering order of every event, so exactly these terms can be taken out of its photon
histories and compared — cell by cell, column by column — with numbers that come from sorted plane crossings and
mid-point / Gauss rules: no marcher, no random numbers, no shared code.  What this reaches that nothing else
independent does: the 3-D ray marcher with periodic wrap on the regular (q1), irregular (q2, q3) and stretched paths,
empty cells and layers, the component pick, the phase-function look-up and 1 / (4 pi |mu|) normalisation of the local
estimate, its Russian-roulette form (Iwabuchi 2006: an unbiased estimator of the same expectation), the Lambertian
surface term and the exit-column bookkeeping (INT:393-841, 1623-1832; OPT:1656-1815)."""
import numpy as np
import pytest

import first_interaction as fi
from independent_3d import Medium, phase_value

EV_BIRTH, EV_SCATTER, EV_SURFACE, EV_KILLED_SURFACE, EV_KILLED_ROULETTE, EV_LE = 1, 2, 3, 5, 6, 8


# ---- the solver itself: closed forms of a homogeneous slab ------------------------------------------------------------
def test_solver_reproduces_the_homogeneous_slab_closed_forms():
    nx, ny, nz = 3, 2, 5
    sigma, omega, g, albedo, mu0 = 7.0, 0.8, 0.6, 0.3, 0.55
    xE, yE = 0.07 * np.arange(nx + 1), 0.11 * np.arange(ny + 1)
    zE = np.concatenate([[0.0], np.cumsum([0.02, 0.05, 0.03, 0.06, 0.04])])
    lc = g ** np.arange(1, 33)
    shape = (1, nz, ny, nx)
    med = Medium(xE, yE, zE, np.full(shape, sigma), np.full(shape, omega), np.ones(shape, np.int32), [[lc]], albedo)
    tauStar = sigma * zE[-1]
    first, surf = med.first_collision(mu0, 70.0, m=5)
    tauAbove = sigma * (zE[-1] - zE[1:])
    want = np.exp(-tauAbove / mu0) * -np.expm1(-sigma * np.diff(zE) / mu0)
    np.testing.assert_allclose(first, np.broadcast_to(want[:, None, None] / (nx * ny), first.shape), rtol=1e-12)
    np.testing.assert_allclose(surf, np.exp(-tauStar / mu0) / (nx * ny), rtol=1e-12)
    sun = np.array([np.sqrt(1 - mu0 ** 2) * np.cos(np.deg2rad(70.0)), np.sqrt(1 - mu0 ** 2) * np.sin(np.deg2rad(70.0)), -mu0])
    for mu, phi in ((1.0, 0.0), (0.35, 200.0), (-0.8, 20.0)):
        view = np.array([np.sqrt(1 - mu ** 2) * np.cos(np.deg2rad(phi)), np.sqrt(1 - mu ** 2) * np.sin(np.deg2rad(phi)), mu])
        P = phase_value(lc, float(sun @ view))
        E1, E0 = med.first_order_radiance(mu0, 70.0, mu, phi, m=3, gauss=8)
        if mu > 0:      # reflected: omega P / (4 pi (mu + mu0)) (1 - exp(-tau* (1/mu + 1/mu0)))
            w1 = omega * P / (4 * np.pi * (mu + mu0)) * -np.expm1(-tauStar * (1 / mu + 1 / mu0))
            w0 = albedo / np.pi * np.exp(-tauStar / mu0 - tauStar / mu)
        else:           # transmitted diffuse radiance at the bottom
            k = 1 / mu0 - 1 / abs(mu)
            w1 = omega * P / (4 * np.pi * mu0 * abs(mu)) * np.exp(-tauStar / abs(mu)) * -np.expm1(-tauStar * k) / k
            w0 = albedo / np.pi * np.exp(-tauStar / mu0)             # the reference's "downward view of the surface": tau = 0
        np.testing.assert_allclose(E1, w1 / (nx * ny), rtol=1e-9)
        np.testing.assert_allclose(E0, w0 / (nx * ny), rtol=1e-9)


@pytest.mark.parametrize("kind", fi.KINDS)
def test_fixture_is_what_the_solver_computes(kind):
    """The committed high-resolution answers against a coarse recomputation (guards against a stale fixture)."""
    fx = fi.fixture(kind)
    _, med = fi.scene(kind, albedo=1.0)
    assert abs(fx["first"].sum() + fx["surf"].sum() - 1.0) < 1e-12
    first, surf = med.first_collision(fi.SOLAR_MU, fi.SOLAR_AZIMUTH, m=24)
    assert np.abs(first - fx["first"]).max() < 3e-3 * fx["first"].max()
    assert np.abs(surf - fx["surf"]).max() < 3e-3 * fx["surf"].max()
    assert (fx["first"] == 0).sum() == (med.sigma == 0).sum() > 20
    for i in (0, 3):
        E1, E0 = med.first_order_radiance(fi.SOLAR_MU, fi.SOLAR_AZIMUTH, fi.VIEW_MUS[i], fi.VIEW_PHIS[i], m=8, gauss=4)
        assert abs(E1.sum() / fx["E1"][i].sum() - 1.0) < 0.015 and np.abs(E1 - fx["E1"][i]).max() < 0.1 * fx["E1"][i].max()
        assert abs(E0.sum() / fx["E0"][i].sum() - 1.0) < 0.015 and np.abs(E0 - fx["E0"][i]).max() < 0.1 * fx["E0"][i].max()


# ---- the oracle's photon histories, first interaction taken out of the event trace --------------------------------------
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


@pytest.mark.parametrize("rr", [0, 1], ids=["le", "le_rr"])
@pytest.mark.parametrize("kind", fi.KINDS)
def test_oracle_first_interaction_matches_the_independent_solver(orc, kind, rr):
    albedo = 0.25
    dom, med = fi.scene(kind, albedo=albedo)
    fx = fi.fixture(kind)
    od = orc.OracleDomain(dom, tableSize=9001, forward=True)
    g = orc.OracleIntegrator(od, useRussianRouletteForIntensity=rr, zetaMin=0.3)
    g.set_views(fi.VIEW_MUS, fi.VIEW_PHIS)
    nDir, ncol = len(fi.VIEW_MUS), med.nx * med.ny
    B, n = 20, 15000
    N = B * n
    cells, cols = np.zeros(med.sigma.size), np.zeros(ncol)
    L1, L0 = np.zeros((B, nDir, ncol)), np.zeros((B, nDir, ncol))
    for b in range(B):
        # enough random numbers for the birth, the first leg, the first event and its view rays; the photon then runs out
        rn = np.random.default_rng(1000 * rr + 100 + b).random((n, 10 + 3 * nDir), dtype=np.float32)
        ev = g.trace(rn, 0, fi.SOLAR_MU, fi.SOLAR_AZIMUTH, maxEvents=n * (6 + 4 * nDir))
        c, s, l1, l0 = first_interaction_of_trace(ev, n, nDir, med)
        cells += c; cols += s; L1[b] = l1 / n; L0[b] = l0 / n

    # where the first collision happens: multinomial counts against Beer's law along the slant paths
    p = fx["first"].ravel()
    assert cells[p == 0].sum() == 0                         # nothing collides in empty cells
    ok = N * p > 25
    rms, mean, worst = z_stats((cells - N * p)[ok] / np.sqrt(N * p * (1 - p))[ok])
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
    for i, mu in enumerate(fi.VIEW_MUS):
        for what, L, E in (("E1", L1[:, i], fx["E1"][i].ravel()), ("E0", L0[:, i], albedo * fx["E0"][i].ravel())):
            if rr and mu < 0:                               # the roulette form only counts rays that reach the TOP (INT:1768-1800)
                assert L.sum() == 0.0
                continue
            m, se = L.mean(axis=0), L.std(axis=0, ddof=1) / np.sqrt(B)
            tot = L.sum(axis=1)
            zt = (tot.mean() - E.sum()) / (tot.std(ddof=1) / np.sqrt(B))
            assert abs(zt) < 4.2, (kind, what, i, tot.mean(), E.sum(), zt)
            if what == "E1":                                # (surface arrivals are too few per column and batch for a z-map)
                rms, mean, worst = z_stats((m - E) / se)    # Student t, 19 degrees of freedom: rms 1.06 if unbiased
                assert rms < 1.45 and abs(mean) < 0.65 and worst < 6.5, (kind, what, i, rms, mean, worst)
                assert abs(tot.mean() / E.sum() - 1.0) < 0.012
