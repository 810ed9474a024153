"""The CUDA kernels against the INDEPENDENT deterministic first-interaction answers (tests/independent_3d.py, fixtures under
tests/golden/first_interaction_*.npz) in a heterogeneous periodic 3-D scene — see tests/test_first_interaction.py for
what the solver is and for the same comparison on the oracle's photon histories.

The throughput kernels write no event trace, so the first interaction is isolated by the optics instead:

* **pure absorber** (every single-scattering albedo 0, black surface): a photon's history IS its first interaction.
  ``volumeAbsorption`` is the first-collision probability per cell and ``fluxDown`` the uncollided beam per column —
  Beer's law along slant paths through the periodically continued field, at 8e7 photons (sigma ~ 0.2 % per cell,
  ~ 2e-4 on the layer sums);
* **first-order radiances by Richardson extrapolation in the albedo**: with every albedo scaled by s the radiance is
  I(s) = s I1 + s^2 I2 + s^3 I3 ..., so 2 I(s) / s - I(2 s) / (2 s) = I1 - 2 s^2 I3: at s = 0.005 the remainder is
  below 1e-4 of I1.  Photon roulette is off for these runs (it would trade the s^2 term's size for variance), the
  local estimate runs in its plain and its Russian-roulette form.

* **thermal emission** (last test): a purely absorbing scene with a temperature that differs from cell to cell over a black
  surface -- every radiance is emission at birth, so the maps must equal the emission integral along the line of sight.

The scene is also run TILED 5 x 5 (same physics, results folded back onto one tile): 40 x 30 columns is what brings in
the photon-pool kernels with their ghost shell, bricked field and vacuum leaps."""
import numpy as np
import pytest

import first_interaction as fi
from independent_3d import emission_radiance, thermal_source
from mcbrat3d_b200.emissionAndBroadBandWeights import Weights, emission_weighting
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.monteCarloRadiativeTransfer import (MCB_ARITH_REFERENCE, MCB_KERNEL_PARK, MCB_KERNEL_POOL,
                                                       computeRadiativeTransfer, finalize_Integrator, getCounters,
                                                       new_Integrator, reportResults, specifyParameters, tracePhotons)
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence

pytestmark = pytest.mark.gpu
NB = 16


def run_batches(dom, n, seed, want, views=False, rr=False, **params):
    g = new_Integrator(dom)
    try:
        if views:
            specifyParameters(g, intensityMus=fi.VIEW_MUS, intensityPhis=fi.VIEW_PHIS, computeIntensity=True,
                              useRussianRouletteForIntensity=bool(rr), zetaMin=0.3)
        specifyParameters(g, minInverseTableSize=9001, minForwardTableSize=9001, **params)
        rs = new_RandomNumberSequence(list(seed))
        rows = {k: [] for k in want}
        for _ in range(NB):
            ps = new_PhotonStream(fi.SOLAR_MU, fi.SOLAR_AZIMUTH, n, rs)
            assert computeRadiativeTransfer(g, dom, rs, ps, n) == n
            res = reportResults(g, **{k: True for k in want})
            for k in want:
                rows[k].append(np.asarray(res[k], np.float64).copy())
        c = getCounters(g)
        assert c["bad"] <= (0 if params.get("arithmetic", 0) != MCB_ARITH_REFERENCE else 2e-5 * n), c
        return {k: np.array(v) for k, v in rows.items()}, c
    finally:
        finalize_Integrator(g)


def area_fraction(dom):
    x, y = np.asarray(dom.xPosition), np.asarray(dom.yPosition)
    return np.outer(np.diff(y), np.diff(x)) / ((x[-1] - x[0]) * (y[-1] - y[0]))


z_stats = fi.z_stats


@pytest.mark.parametrize("rr", [False, True], ids=["le", "le_rr"])
@pytest.mark.parametrize("kind", fi.KINDS)
def test_trace_kernel_first_interaction_matches_the_independent_solver(kind, rr):
    """The reference-arithmetic CUDA kernel's own event trace (mcbref::trace_kernel), taken apart exactly like the
    oracle's in tests/test_first_interaction.py: first-collision cells, landing columns, first-order radiance maps and
    the surface term against the deterministic answers."""
    albedo = 0.25
    dom, med = fi.scene(kind, albedo=albedo)
    nDir, ncol = len(fi.VIEW_MUS), med.nx * med.ny
    B, n = 20, 15000
    g = new_Integrator(dom)
    try:
        specifyParameters(g, intensityMus=fi.VIEW_MUS, intensityPhis=fi.VIEW_PHIS, computeIntensity=True,
                          useRussianRouletteForIntensity=rr, zetaMin=0.3, minInverseTableSize=9001, minForwardTableSize=9001)
        rs = new_RandomNumberSequence([10, 1, 0])
        ps = new_PhotonStream(fi.SOLAR_MU, fi.SOLAR_AZIMUTH, n, rs)
        cells, cols = np.zeros(med.sigma.size), np.zeros(ncol)
        L1, L0 = np.zeros((B, nDir, ncol)), np.zeros((B, nDir, ncol))
        for b in range(B):
            rn = np.random.default_rng(5000 + 1000 * rr + b).random((n, 10 + 3 * nDir), dtype=np.float32)
            ev, _ = tracePhotons(g, dom, ps, rn, maxEventsPerPhoton=8 + 6 * nDir)
            c, s, l1, l0 = fi.first_interaction_of_trace(ev, n, nDir, med)
            cells += c; cols += s; L1[b] = l1 / n; L0[b] = l0 / n
    finally:
        finalize_Integrator(g)
    fi.check_first_interaction(kind, albedo, int(rr), cells, cols, L1, L0, B * n)


ABSORBER = [("regular", (1, 1), {}), ("regular_pool", (2, 2), dict(tuneKernel=MCB_KERNEL_POOL)),
            ("irregular", (1, 1), {}), ("irregular_tiled", (5, 5), {}),
            ("irregular_tiled_park", (5, 5), dict(tuneKernel=MCB_KERNEL_PARK)),
            ("irregular_tiled_no_leaps", (5, 5), dict(tuneKernel=MCB_KERNEL_POOL, tuneLeap=-1)),
            ("stretched", (1, 1), {}), ("irregular_reference_arithmetic", (1, 1), dict(arithmetic=MCB_ARITH_REFERENCE))]


@pytest.mark.parametrize("name,tiles,params", ABSORBER, ids=[c[0] for c in ABSORBER])
def test_pure_absorber_matches_beers_law_along_slant_paths(name, tiles, params):
    kind = name.split("_")[0]
    dom, med = fi.scene(kind, albedo=0.0, ssaScale=0.0, tiles=tiles)
    fx = fi.fixture(kind)
    n = 5_000_000 if params.get("arithmetic", 0) != MCB_ARITH_REFERENCE else 500_000
    rows, c = run_batches(dom, n, (21, 4, 0), ("volumeAbsorption", "fluxDown", "fluxUp", "meanFluxAbsorbed", "meanFluxDown"), **params)
    assert c["photons"] == n
    af = area_fraction(dom)
    dz = np.diff(np.asarray(dom.zPosition))
    first = fi.fold(rows["volumeAbsorption"] * (dz[:, None, None] * 1000.0) * af, tiles)        # (NB, nz, ny, nx) probabilities
    surf = fi.fold(rows["fluxDown"] * af, tiles)
    assert np.all(rows["fluxUp"] == 0.0)
    assert np.all(first[:, fx["first"] == 0] == 0.0)                                            # nothing is absorbed in empty cells
    np.testing.assert_allclose(first.sum(axis=(1, 2, 3)) + surf.sum(axis=(1, 2)), 1.0, atol=1e-5)   # every photon ends somewhere
    fi.check_absorber(name, first, surf, fx)


RADIANCE = [("regular", (1, 1), {}), ("regular_park", (1, 1), dict(tuneKernel=MCB_KERNEL_PARK)),
            ("irregular_tiled", (5, 5), {}), ("irregular_tiled_park", (5, 5), dict(tuneKernel=MCB_KERNEL_PARK)),
            ("stretched", (1, 1), {})]


@pytest.mark.parametrize("rr", [False, True], ids=["le", "le_rr"])
@pytest.mark.parametrize("name,tiles,params", RADIANCE, ids=[c[0] for c in RADIANCE])
def test_first_order_radiances_match_the_independent_solver(name, tiles, params, rr):
    """(The "stretched-le" case is also the regression test of mcb_march.cuh::seam_fix: photon batch 8 of the second run
    holds the event at the periodic seam that, before the fix, put 1e11 ... inf into three radiance maps.)"""
    kind = name.split("_")[0]
    fx = fi.fixture(kind)
    s, n = 0.005, 2_000_000
    f = []
    for k, scale in enumerate((s, 2 * s)):
        dom, med = fi.scene(kind, albedo=0.0, ssaScale=scale, tiles=tiles)
        rows, c = run_batches(dom, n, (31 + k, 5, 0), ("intensity",), views=True, rr=rr, useRussianRoulette=False, **params)
        assert c["leRays"] > 0
        f.append(fi.fold(rows["intensity"] * area_fraction(dom), tiles) / scale)               # (NB, nDir, ny, nx): E[contribution] / s
    m = 2.0 * f[0].mean(axis=0) - f[1].mean(axis=0)
    se = np.sqrt(4.0 * f[0].var(axis=0, ddof=1) / NB + f[1].var(axis=0, ddof=1) / NB)
    t0, t1 = f[0].sum(axis=(2, 3)), f[1].sum(axis=(2, 3))                                       # per batch and direction
    tot = 2.0 * t0.mean(axis=0) - t1.mean(axis=0)
    tse = np.sqrt(4.0 * t0.var(axis=0, ddof=1) / NB + t1.var(axis=0, ddof=1) / NB)
    for i, mu in enumerate(fi.VIEW_MUS):
        if rr and mu < 0:                                   # the roulette form only counts view rays that reach the TOP (INT:1768-1800)
            assert f[0][:, i].sum() == 0.0 and f[1][:, i].sum() == 0.0
            continue
        E = fx["E1"][i]
        quadT, quadC = 3e-4 * E.sum(), 2e-3 * E.max()       # the fixture's quadrature error (totals / columns), see make_first_interaction.py
        zt = (tot[i] - E.sum()) / np.sqrt(tse[i] ** 2 + quadT ** 2)
        rms, mean, worst = z_stats(((m[i] - E) / np.sqrt(se[i] ** 2 + quadC ** 2)).ravel())
        print("%s rr %d view %d: total %.6g vs %.6g (rel %+.2e, z %+.2f); columns rms %.3f mean %+.3f max %.2f" % (
            name, rr, i, tot[i], E.sum(), tot[i] / E.sum() - 1, zt, rms, mean, worst))
        assert abs(zt) < 4.5 and abs(tot[i] / E.sum() - 1.0) < 0.01, (name, rr, i, tot[i], E.sum(), zt)
        assert rms < 1.4 and abs(mean) < 0.7 and worst < 6.5, (name, rr, i, rms, mean, worst)


THERMAL = [("regular", (1, 1), {}), ("irregular_tiled", (5, 5), {}), ("irregular_tiled_park", (5, 5), dict(tuneKernel=MCB_KERNEL_PARK)),
           ("zstretched", (1, 1), {}), ("irregular_reference_arithmetic", (1, 1), dict(arithmetic=MCB_ARITH_REFERENCE))]


@pytest.mark.parametrize("name,tiles,params", THERMAL, ids=[c[0] for c in THERMAL])
def test_thermal_emission_radiance_matches_the_emission_integral(name, tiles, params):
    """Thermal source in a purely absorbing 3-D scene with a temperature that differs from cell to cell, over a black
    (emissivity 1) surface: every radiance is emission at birth, so the maps must equal the emission integral
    int kappa B exp(-tau) ds + transmitted surface emission of tests/independent_3d.py (exact per piece), and the
    emission CDF built on the device must give Kirchhoff's share of the atmosphere.  Downward view: the reference drops
    the contribution of photons born AT the surface (tests/test_first_interaction.py); the reference-arithmetic kernel
    must do the same, the throughput kernels may count it (tau = 0: weight / pi) -- which one is printed."""
    kind = name.split("_")[0]
    lam, sfcT = 10.0, 300.0
    temps = 250.0 + 40.0 * np.random.default_rng(5).random((fi.NZ, fi.NY, fi.NX))
    dom, med = fi.scene(kind, albedo=0.0, ssaScale=0.0, tiles=tiles, temps=temps, lambda_um=lam)
    reference = params.get("arithmetic", 0) == MCB_ARITH_REFERENCE
    n = 1_000_000 if not reference else 250_000
    g = new_Integrator(dom)
    try:
        specifyParameters(g, intensityMus=fi.VIEW_MUS, intensityPhis=fi.VIEW_PHIS, computeIntensity=True,
                          useRussianRouletteForIntensity=False)
        specifyParameters(g, minInverseTableSize=9001, minForwardTableSize=9001, LW_flag=1.0, **params)
        w = Weights()
        emission_weighting(dom, w, sfcT, thisIntegrator=g)
        pCell, pSfc, _, sfcTerm = thermal_source(med, temps, lam, sfcT)
        assert abs(w.fracAtmsPower - pCell.sum()) < 1e-9
        rs = new_RandomNumberSequence([41, 6, 0])
        rows = []
        for _ in range(NB):
            ps = new_PhotonStream(theseWeights=w, numberOfPhotons=n, randomNumbers=rs)
            assert computeRadiativeTransfer(g, dom, rs, ps, n) == n
            rows.append(np.asarray(reportResults(g, intensity=True)["intensity"], np.float64).copy())
        assert getCounters(g)["bad"] <= (2e-5 * n if reference else 0)
    finally:
        finalize_Integrator(g)
    got = fi.fold(np.array(rows) * area_fraction(dom), tiles)                     # (NB, nDir, ny, nx): E[contribution at birth]
    m, se = got.mean(axis=0), got.std(axis=0, ddof=1) / np.sqrt(NB)
    tot = got.sum(axis=(2, 3))
    for i, (mu, phi) in enumerate(zip(fi.VIEW_MUS, fi.VIEW_PHIS)):
        E = emission_radiance(med, temps, lam, sfcT, mu, phi, m=32)
        if mu < 0:
            dropped = E - sfcTerm * med.area
            counts = abs(tot[:, i].mean() - E.sum()) < abs(tot[:, i].mean() - dropped.sum())
            print("%s downward view: surface-born photons %s" % (name, "COUNTED (tau = 0)" if counts else "dropped, as in the reference"))
            assert not (reference and counts)
            E = E if counts else dropped
        zt = (tot[:, i].mean() - E.sum()) / np.sqrt(tot[:, i].var(ddof=1) / NB + (2e-4 * E.sum()) ** 2)
        rms, mean, worst = z_stats(((m[i] - E) / np.sqrt(se[i] ** 2 + (1e-3 * E.max()) ** 2)).ravel())
        print("%s view %d: total %.6g vs %.6g (rel %+.2e, z %+.2f); columns rms %.3f mean %+.3f max %.2f" % (
            name, i, tot[:, i].mean(), E.sum(), tot[:, i].mean() / E.sum() - 1, zt, rms, mean, worst))
        assert abs(zt) < 4.5 and abs(tot[:, i].mean() / E.sum() - 1.0) < 0.01, (name, i, tot[:, i].mean(), E.sum(), zt)
        assert rms < 1.4 and abs(mean) < 0.7 and worst < 6.5, (name, i, rms, mean, worst)
