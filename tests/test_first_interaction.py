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
from independent_3d import Medium, emission_radiance, phase_value, thermal_source

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


# ---- thermal source: where photons are born and the emission they contribute at birth ---------------------------------------
@pytest.mark.parametrize("kind", ["regular", "irregular", "zstretched"])
def test_oracle_thermal_births_and_emission_radiance_match_the_independent_solver(orc, kind):
    """Kirchhoff's law and the emission integral in a 3-D scene with a temperature that differs from cell to cell: the
    share of the atmosphere in the emitted power (emission_weighting, EMI:424-550: equal to 1e-12), the cell every
    photon is born in (findCDFIndex on the voxel CDF, ILL:481-515), the surface births, and the local-estimate
    contribution at birth per view direction and exit column (INT:513-542, 1695-1696) against
    int kappa_abs B exp(-tau) ds + transmitted surface emission, integrated exactly piece by piece along sorted plane
    crossings (tests/independent_3d.py).  Grids with uniform columns only: the reference's voxel weights carry no cell
    area (EMI:489-496 weighs a cell by 4 pi B kappa dz).  One reference behaviour shows up and is asserted: a photon born
    AT the surface contributes nothing to a downward view (its view ray has no first step; the marcher signals an error
    and INT:1745-1751 drops the contribution) although a photon REFLECTED there does (tests above)."""
    albedo, lam, sfcT = 0.25, 10.0, 300.0
    temps = 250.0 + 40.0 * np.random.default_rng(5).random((fi.NZ, fi.NY, fi.NX))
    dom, med = fi.scene(kind, albedo=albedo, temps=temps, lambda_um=lam)
    od = orc.OracleDomain(dom, tableSize=9001, forward=True)
    frac, cdf, _ = od.emission_weighting(temps, lam, sfcT)
    pCell, pSfc, _, sfcTerm = thermal_source(med, temps, lam, sfcT)
    assert abs(frac - pCell.sum()) < 1e-12 and abs(pCell.sum() + pSfc - 1.0) < 1e-12
    g = orc.OracleIntegrator(od, LW_flag=1.0)
    g.set_views(fi.VIEW_MUS, fi.VIEW_PHIS)
    nDir, ncol = len(fi.VIEW_MUS), med.nx * med.ny
    B, n = 20, 15000
    N = B * n
    births, sfcBirths, L = np.zeros(med.sigma.size), np.zeros(ncol), np.zeros((B, nDir, ncol))
    for b in range(B):
        rn = np.random.default_rng(300 + b).random((n, 9), dtype=np.float32)       # birth: 7 numbers; the photon soon runs out
        ev = g.trace(rn, 1, 1.0, 0.0, fracAtmsPower=frac, voxelCDF=cdf, maxEvents=n * (4 + 3 * nDir))
        bi = ev[ev["kind"] == fi.EV_BIRTH]
        assert len(bi) == n
        atSurface = bi["z"] == 0.0
        cell = (bi["ix"] - 1) + med.nx * ((bi["iy"] - 1) + med.ny * (bi["iz"] - 1))
        births += np.bincount(cell[~atSurface], minlength=births.size)
        sfcBirths += np.bincount(((bi["ix"] - 1) + med.nx * (bi["iy"] - 1))[atSurface], minlength=ncol)
        le = ev[(ev["kind"] == fi.EV_LE) & (ev["order"] == 0)]
        idx = (le["component"] - 1) * ncol + (le["ix"] - 1) + med.nx * (le["iy"] - 1)
        L[b] = np.bincount(idx, weights=le["weight"].astype(np.float64), minlength=nDir * ncol).reshape(nDir, ncol) / n
    p = pCell.ravel()
    assert births[p == 0].sum() == 0                       # no photon is born where nothing absorbs
    ok = N * p > 25
    rms, mean, worst = fi.z_stats((births - N * p)[ok] / np.sqrt(N * p * (1 - p))[ok])
    print("%s: births over %d cells: z rms %.3f mean %+.3f max %.2f" % (kind, ok.sum(), rms, mean, worst))
    assert ok.sum() > 250 and rms < 1.12 and abs(mean) < 4.0 / np.sqrt(ok.sum()) and worst < 4.8, (kind, rms, mean, worst)
    q = pSfc * med.area.ravel()
    rms, mean, worst = fi.z_stats((sfcBirths - N * q) / np.sqrt(N * q * (1 - q)))
    assert rms < 1.35 and abs(mean) < 0.6 and worst < 4.5, (kind, "surface births", rms, mean, worst)
    for i, (mu, phi) in enumerate(zip(fi.VIEW_MUS, fi.VIEW_PHIS)):
        E = emission_radiance(med, temps, lam, sfcT, mu, phi, m=32).ravel()
        if mu < 0:
            E = E - sfcTerm * med.area.ravel()             # surface-born photons and downward views: see the docstring
        m, se = L[:, i].mean(axis=0), L[:, i].std(axis=0, ddof=1) / np.sqrt(B)
        tot = L[:, i].sum(axis=1)
        zt = (tot.mean() - E.sum()) / (tot.std(ddof=1) / np.sqrt(B))
        rms, mean, worst = fi.z_stats((m - E) / se)
        print("%s view %d: total %.6g vs %.6g (rel %+.2e, z %+.2f); columns z rms %.3f mean %+.3f max %.2f" % (
            kind, i, tot.mean(), E.sum(), tot.mean() / E.sum() - 1, zt, rms, mean, worst))
        assert abs(zt) < 4.2 and abs(tot.mean() / E.sum() - 1.0) < 0.01, (kind, i, tot.mean(), E.sum(), zt)
        assert rms < 1.45 and abs(mean) < 0.65 and worst < 6.5, (kind, i, rms, mean, worst)


def test_periodic_seam_fold_can_disagree_with_the_cell_and_seam_fix_repairs_it():
    """Single-precision emulation of what csrc/mcb_march.cuh does at an event on an edge-table grid: ray_position() folds the
    position with floor((p - x0) / L) computed as a rounded product, the marcher wraps the integer cell index.  On the
    stretched grid of these tests a position exactly AT the seam (p = L, cell nx wrapped to 0) is NOT folded -- position at
    the far end of the domain, cell at the near end: the state that blew up the radiances.  seam_fix moves
    the position by the one period that puts it next to its cell; positions that agree with their cell are left alone."""
    f32 = np.float32
    dom, _ = fi.scene("stretched")
    xE = np.asarray(dom.xPosition, dtype=f32)
    L, x0 = f32(xE[-1] - xE[0]), f32(xE[0])
    invL = f32(1.0) / L

    def fold(p):                                  # ray_position
        return f32(p - L * np.floor(f32(f32(p - x0) * invL)))

    def seam_fix(p, ix):                          # mcb_march.cuh::seam_fix
        e = f32(p - f32(0.5) * f32(xE[ix] + xE[ix + 1]))
        return f32(p - (L if e > f32(0.5) * L else (-L if e < f32(-0.5) * L else f32(0.0))))

    p = fold(L)                                   # the marcher says: ghost cell nx, i.e. cell 0 after the wrap
    assert p == L                                 # ... but the fold leaves the position at the far end
    assert abs(float(seam_fix(p, 0))) < 1e-6      # repaired: at the left edge of cell 0
    below = np.nextafter(L, f32(0))               # just inside the last cell: consistent, must stay
    assert fold(below) == below and seam_fix(fold(below), fi.NX - 1) == below
    rng = np.random.default_rng(0)
    for _ in range(2000):                         # consistent (position, cell) pairs anywhere are never moved
        ix = int(rng.integers(0, fi.NX))
        q = f32(xE[ix] + rng.random() * (xE[ix + 1] - xE[ix]))
        assert seam_fix(q, ix) == q
    for ix, q in ((fi.NX - 1, f32(-1e-9)), (0, L), (0, np.nextafter(L, f32(1)))):   # the three ways to be one period off
        fixed = float(seam_fix(q, ix))
        assert xE[ix] - 1e-6 <= fixed <= xE[ix + 1] + 1e-6
