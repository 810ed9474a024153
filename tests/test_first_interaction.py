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
        c, s, l1, l0 = fi.first_interaction_of_trace(ev, n, nDir, med)
        cells += c; cols += s; L1[b] = l1 / n; L0[b] = l0 / n

    fi.check_first_interaction(kind, albedo, rr, cells, cols, L1, L0, N)


def test_absorber_checker_accepts_an_unbiased_sampler_and_sees_a_small_bias():
    """The statistical check the GPU pure-absorber test applies (fi.check_absorber), exercised on CPU with multinomial
    samples of the committed probabilities at the GPU test's photon count: it passes an unbiased sampler and fails one
    that moves 0.2 % of one layer's collisions into the layer below (the power the GPU test has)."""
    fx = fi.fixture("irregular")
    nb, n = 16, 2_500_000
    for bias in (0.0, 2e-3):
        p = fx["first"].copy()
        moved = bias * p[5].sum()
        p[5] *= 1.0 - bias; p[4] *= 1.0 + moved / p[4].sum()
        pAll = np.concatenate([p.ravel(), fx["surf"].ravel()])
        counts = np.random.default_rng(8).multinomial(n, pAll / pAll.sum(), size=nb) / n
        first = counts[:, :p.size].reshape((nb,) + p.shape)
        surf = counts[:, p.size:].reshape((nb,) + fx["surf"].shape)
        if bias == 0.0:
            fi.check_absorber("synthetic", first, surf, fx)
        else:
            with pytest.raises(AssertionError):
                fi.check_absorber("synthetic biased", first, surf, fx)
