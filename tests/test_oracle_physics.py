"""Pin the oracle with physical invariants: the reference ships no fixtures, so
these closed-form properties are what anchors the restatement of computeRT / the marcher."""
import ctypes as C

import numpy as np
import pytest

from mcbrat3d_b200 import domains


def _fin(orc, st, key, tot, nb):
    m, e = orc.finalise(st[key], 1.0, tot, nb)
    return m, e


def test_energy_closure_solar(orc):
    """Fup + (1 - A) Fdn + Fabs = 1 (INT:221-223: absorption agrees with the flux divergence)."""
    d, case = domains.homogeneous_slab(ssa=0.99, albedo=0.2)
    g = orc.OracleIntegrator(orc.OracleDomain(d, tableSize=9001))
    nb = 16
    tot, st = g.run_batches(nb, 5000, solarMu=0.5, solarAzimuth=0.0)
    up, eu = _fin(orc, st, "meanFluxUpStats", tot, nb)
    dn, ed = _fin(orc, st, "meanFluxDownStats", tot, nb)
    ab, ea = _fin(orc, st, "meanFluxAbsorbedStats", tot, nb)
    closure = up[0] + 0.8 * dn[0] + ab[0]
    # Russian roulette makes the closure a statistical identity, not an exact one
    assert abs(closure - 1.0) < 4.0 * np.sqrt(eu[0] ** 2 + (0.8 * ed[0]) ** 2 + ea[0] ** 2) + 1e-3


def test_beer_law_direct_beam(orc):
    """Pure absorber: every photon is absorbed at its first event, so Fdn = exp(-tau/mu0)."""
    tau, mu0 = 2.0, 0.5
    d, case = domains.homogeneous_slab(ssa=0.0, tau=tau, albedo=0.0)
    g = orc.OracleIntegrator(orc.OracleDomain(d, tableSize=9001))
    nb = 20
    tot, st = g.run_batches(nb, 5000, solarMu=mu0, solarAzimuth=0.0)
    dn, ed = _fin(orc, st, "meanFluxDownStats", tot, nb)
    up, _ = _fin(orc, st, "meanFluxUpStats", tot, nb)
    ab, _ = _fin(orc, st, "meanFluxAbsorbedStats", tot, nb)
    want = np.exp(-tau / mu0)
    assert abs(dn[0] - want) < 4.0 * ed[0] + 1e-4
    assert up[0] == 0.0
    assert abs(dn[0] + ab[0] - 1.0) < 1e-5           # no roulette variance here: weight 1 or 0


def test_absorption_profile_sums_to_column_absorption(orc):
    """sum_k volumeAbsorption(:,:,k) dz_k 1000 = fluxAbsorbed (INT:361-364 normalisation)."""
    d, case = domains.homogeneous_slab(ssa=0.9, tau=4.0)
    od = orc.OracleDomain(d, tableSize=9001)
    g = orc.OracleIntegrator(od)
    lib = orc.load()
    rng = orc.orc_rng()
    key = (C.c_uint32 * 3)(10, 1, 0)
    lib.orc_rng_init_array(C.byref(rng), key, 3)
    lib.orc_photons_directional.restype = C.c_void_p
    lib.orc_photons_directional.argtypes = [C.c_float, C.c_float, C.c_int64, C.POINTER(orc.orc_rng)]
    lib.orc_compute_radiative_transfer.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(orc.orc_rng), C.c_void_p,
                                                   C.c_int64, C.c_int, C.POINTER(C.c_int64)]
    lib.orc_report_results.argtypes = [C.c_void_p] + [C.POINTER(C.c_float)] * 10
    ph = lib.orc_photons_directional(0.5, 0.0, 20000, C.byref(rng))
    done = C.c_int64(0)
    assert lib.orc_compute_radiative_transfer(g.ptr, od.ptr, C.byref(rng), ph, 20000, 1, C.byref(done)) == 0
    assert done.value == 20000
    n = 20
    fab = np.zeros((n, n), np.float32); vol = np.zeros((n, n, n), np.float32); prof = np.zeros(n, np.float32)
    mab = C.c_float(0)
    lib.orc_report_results(g.ptr, None, None, C.byref(mab), None, None, fab.ctypes.data_as(C.POINTER(C.c_float)),
                           prof.ctypes.data_as(C.POINTER(C.c_float)), vol.ctypes.data_as(C.POINTER(C.c_float)), None, None)
    dz = 0.0625
    np.testing.assert_allclose(vol.sum(axis=0) * dz * 1000.0, fab, rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(prof.sum() * dz * 1000.0, mab.value, rtol=2e-4)


def test_lw_isothermal_closure(orc):
    """Isothermal medium over a black surface at the same temperature: net absorbed - emitted = 0
    everywhere in expectation, i.e. mean fluxAbsorbed (with the -1 emission term, INT:505-508)
    balances against what leaves through the top."""
    d, case = domains.homogeneous_lw(ssa=0.5, tau=5.0, albedo=0.0, atmTemp=290.0, sfcTemp=290.0)
    od = orc.OracleDomain(d, tableSize=9001)
    frac, cdf, flux = od.emission_weighting(d.temps, d.lambda_um, 290.0)
    assert 0.0 < frac < 1.0 and cdf[-1, -1, -1] == 1.0 and np.all(np.diff(cdf.ravel()) >= 0)
    g = orc.OracleIntegrator(od, LW_flag=1.0)
    nb = 16
    tot, st = g.run_batches(nb, 5000, source=1, fracAtmsPower=frac, voxelCDF=cdf)
    up, eu = _fin(orc, st, "meanFluxUpStats", tot, nb)
    dn, ed = _fin(orc, st, "meanFluxDownStats", tot, nb)
    ab, ea = _fin(orc, st, "meanFluxAbsorbedStats", tot, nb)
    # photons: fraction `frac` born in the atmosphere (-1 each), all end absorbed, at the surface, or out the top
    # energy bookkeeping: (absorbed in atmosphere - emitted by atmosphere) + surface absorption + escape = surface emission
    closure = ab[0] + dn[0] + up[0]
    assert abs(closure - (1.0 - frac)) < 4.0 * np.sqrt(eu[0] ** 2 + ed[0] ** 2 + ea[0] ** 2) + 2e-3
    # isothermal + black surface: upwelling flux at the top equals the black-body flux, i.e. the
    # escaping fraction of emitted power is pi B / (total emitted per unit area)
    planck_flux_fraction = (1.0 - frac)          # surface emits pi*B*area; it is the (1-frac) share
    assert abs(up[0] - planck_flux_fraction) < 4.0 * eu[0] + 5e-3


def test_march_vertical_optical_depth(orc):
    """A vertical ray through the slab accumulates exactly tau; oblique rays tau/mu; x/y wrap."""
    lib = orc.load()
    d, _ = domains.homogeneous_slab(ssa=1.0, tau=10.0)
    od = orc.OracleDomain(d, tableSize=9001)
    for mu, phi in ((-1.0, 0.0), (-0.5, 0.3), (-0.2, 2.0), (0.7, 4.0)):
        dirv = np.zeros(3, np.float32)
        lib.orc_make_direction_cosines(C.c_float(mu), C.c_float(phi), dirv.ctypes.data_as(C.POINTER(C.c_float)))
        x, y = C.c_double(0.3), C.c_double(0.7)
        z = C.c_double(1.25 - 1e-9 if mu < 0 else 1e-9)
        ix, iy, iz = C.c_int(5), C.c_int(12), C.c_int(20 if mu < 0 else 1)
        path, cross = C.c_double(0), C.c_int64(0)
        ext = lib.orc_march(od.ptr, dirv.ctypes.data_as(C.POINTER(C.c_float)), C.byref(x), C.byref(y), C.byref(z),
                            C.byref(ix), C.byref(iy), C.byref(iz), 0, 0.0, C.byref(path), C.byref(cross))
        assert ext == pytest.approx(10.0 / abs(mu), rel=2e-6)
        assert path.value == pytest.approx(1.25 / abs(mu), rel=1e-6)
        assert (iz.value == 0) if mu < 0 else (iz.value == 21)
        assert 1 <= ix.value <= 20 and 1 <= iy.value <= 20
        assert 0.0 <= x.value <= 1.25 and 0.0 <= y.value <= 1.25


def test_march_stops_at_target(orc):
    lib = orc.load()
    d, _ = domains.homogeneous_slab(ssa=1.0, tau=10.0)
    od = orc.OracleDomain(d, tableSize=9001)
    dirv = np.array([0.0, 0.0, -1.0], np.float32)
    x, y, z = C.c_double(0.3), C.c_double(0.7), C.c_double(1.2)
    ix, iy, iz = C.c_int(5), C.c_int(12), C.c_int(20)
    path, cross = C.c_double(0), C.c_int64(0)
    ext = lib.orc_march(od.ptr, dirv.ctypes.data_as(C.POINTER(C.c_float)), C.byref(x), C.byref(y), C.byref(z),
                        C.byref(ix), C.byref(iy), C.byref(iz), 1, 3.0, C.byref(path), C.byref(cross))
    assert ext == 3.0
    assert z.value == pytest.approx(1.2 - 3.0 / 8.0, rel=1e-6)
    assert iz.value == int((1.2 - 0.375) / 0.0625) + 1


def test_regular_grid_detection_quirk_q1(orc):
    """new_Integrator keeps the spacing in default reals (INT:140, 163-181): a grid counts as
    regular only if its spacing is exactly representable in single precision."""
    d, _ = domains.homogeneous_slab()
    h = orc.OracleIntegrator(orc.OracleDomain(d, tableSize=9001)).head()
    assert h.xyRegularlySpaced == 1 and h.zRegularlySpaced == 1 and h.deltaX == 0.0625
    di, _ = domains.irregular_test_domain()
    hi = orc.OracleIntegrator(orc.OracleDomain(di, tableSize=9001)).head()
    assert hi.xyRegularlySpaced == 0 and hi.zRegularlySpaced == 0
    dl, _ = domains.step_cloud()
    hl = orc.OracleIntegrator(orc.OracleDomain(dl, tableSize=9001)).head()
    assert hl.xyRegularlySpaced == 1 and hl.zRegularlySpaced == 1 and hl.deltaX == 15.625


def test_max_cross_section_equals_ray_tracing_on_homogeneous_slab(orc):
    """INT:564-571, 709-710: on a homogeneous domain every Woodcock event is physical and the
    reference's stale cell indices are harmless, so the two transport modes sample the same problem."""
    dom, case = domains.homogeneous_slab(ssa=0.99)
    od = orc.OracleDomain(dom, tableSize=9001)
    res = {}
    for rt in (1, 0):
        og = orc.OracleIntegrator(od, useRayTracing=rt)
        tot, st = og.run_batches(16, 3000, source=0, solarMu=case["solarMu"], solarAzimuth=case["solarAzimuth"])
        res[rt] = [orc.finalise(st[q + "Stats"], 1.0, tot, 16) for q in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed")]
    for (m1, e1), (m0, e0) in zip(res[1], res[0]):
        assert abs(m1[0] - m0[0]) < 3.5 * np.hypot(e1[0], e0[0])
    closure = res[0][0][0][0] + 0.8 * res[0][1][0][0] + res[0][2][0][0]
    assert abs(closure - 1.0) < 5e-3


def test_make_periodic_is_single_precision(orc):
    """Quirk q15: with injected random numbers the x position after a maximum cross-section move is a
    float (makePeriodic returns default real), and mathematical collisions are traced."""
    dom, case = domains.irregular_test_domain()
    od = orc.OracleDomain(dom, tableSize=9001)
    og = orc.OracleIntegrator(od, useRayTracing=0)
    rn = np.random.default_rng(3).random((100, 600), dtype=np.float32)
    ev = og.trace(rn, 0, case["solarMu"], case["solarAzimuth"], maxEvents=100 * 2048)
    null = ev[ev["kind"] == 10]
    assert len(null) > 0
    assert np.array_equal(null["x"], null["x"].astype(np.float32).astype(np.float64))
    assert np.array_equal(null["y"], null["y"].astype(np.float32).astype(np.float64))


# (tau, omega, g, mu0, surface albedo): the plane-parallel table protocol of Drivers/planeParallel.f95:242, 269-272
PLANE_PARALLEL = [(10.0, 0.99, 0.85, 0.5, 0.2), (10.0, 1.0, 0.85, 0.5, 0.2), (1.0, 0.9, 0.0, 1.0, 0.0),
                  (2.0, 0.95, 0.6, 0.3, 0.5), (0.2, 1.0, 0.85, 0.8, 0.0)]


@pytest.mark.parametrize("tau,omega,g,mu0,albedo", PLANE_PARALLEL)
def test_plane_parallel_fluxes_match_adding_doubling(orc, tau, omega, g, mu0, albedo):
    """An INDEPENDENT pin of the oracle (SURVEY 8c substitute 3): on a horizontally homogeneous slab the 3-D Monte
    Carlo must reproduce the deterministic adding-doubling solution of the plane-parallel problem
    (tests/adding_doubling.py: no shared code, no random numbers; it is given the Legendre moments of the scattering
    law the inverse table encodes, see table_moments) -- reflected, transmitted and absorbed flux within
    4 sigma of the batch standard error (+2e-4 for the solver's angular quadrature and the 10001-entry phase table)."""
    from adding_doubling import slab_fluxes, table_moments
    d, case = domains.homogeneous_slab(ssa=omega, tau=tau, albedo=albedo, g=g, n=8, delta=0.125)
    og = orc.OracleIntegrator(orc.OracleDomain(d, tableSize=10001))
    nb = 40
    tot, st = og.run_batches(nb, 5000, solarMu=mu0, solarAzimuth=0.0, iseed=10, rank=1, thread=0)
    d.tabulateInversePhaseFunctions(10001)
    want = slab_fluxes(tau, omega, table_moments(d.inversePhaseFunctions[0]), mu0, albedo, nStreams=96)
    for name, w in zip(("meanFluxUpStats", "meanFluxDownStats", "meanFluxAbsorbedStats"), want):
        m, e = _fin(orc, st, name, tot, nb)
        assert abs(m[0] - w) < 4.0 * e[0] + 2e-4, (name, m[0], w, e[0])


@pytest.mark.parametrize("rr", [False, True], ids=["plain", "rr"])
@pytest.mark.parametrize("tau,omega,mu0,albedo", [(2.0, 0.9, 0.5, 0.3), (0.5, 1.0, 0.8, 0.0)])
def test_local_estimate_radiances_match_adding_doubling(orc, tau, omega, mu0, albedo, rr):
    """Independent pin of the local estimate (computeIntensityContribution INT:1623-1832, with and without the
    Russian-roulette variants): for ISOTROPIC scattering the radiance leaving a plane-parallel slab does not depend on
    azimuth, so the adding-doubling solver's azimuthally averaged radiance at the top is the answer for every view
    direction -- normalisation 1/(4 pi |mu_view|), the Lambertian surface term 1/pi and the extinction along the view
    rays included.  4 sigma of the batch standard error + 0.2 %."""
    from adding_doubling import slab_fluxes
    d, case = domains.homogeneous_slab(ssa=omega, tau=tau, albedo=albedo, g=0.0, n=8, delta=0.125)
    mus, phis = [1.0, 0.866, 0.5], [0.0, 0.0, 180.0]
    og = orc.OracleIntegrator(orc.OracleDomain(d, tableSize=10001, forward=True), useRussianRouletteForIntensity=int(rr), zetaMin=0.3)
    og.set_views(mus, phis)
    nb = 30
    tot, st = og.run_batches(nb, 4000, solarMu=mu0, solarAzimuth=0.0, iseed=10, rank=1, thread=0)
    cols = d.numX * d.numY
    m, e = orc.finalise(st["radianceStats"], 1.0, tot, nb)
    got = m.reshape(-1, cols).mean(axis=1); err = np.sqrt((e.reshape(-1, cols) ** 2).sum(axis=1)) / cols
    want = slab_fluxes(tau, omega, [1.0], mu0, albedo, muOut=mus)[3]
    assert (np.abs(got - want) < 4.0 * err + 2e-3 * want).all(), (got, want, err)


@pytest.mark.parametrize("name,layers", [("two_layers", [(1.0, 1.0, 4), (3.0, 0.8, 4)]),
                                         ("gap", [(2.0, 0.9, 3), (0.0, 0.0, 2), (4.0, 0.99, 3)])])
def test_layered_slab_fluxes_match_adding(orc, name, layers):
    """Vertical structure against the independent solver: a stack of different homogeneous layers (one case with an
    EMPTY layer in the middle: extinction 0, phase-function index 0) is marched cell by cell by the Monte Carlo and
    solved layer by layer with doubling + adding (tests/adding_doubling.py::layered_fluxes; the adding step reproduces
    the single-slab solution exactly when the layers are identical).  layers: (tau, omega, cells), TOP FIRST."""
    from adding_doubling import layered_fluxes, table_moments
    from mcbrat3d_b200.opticalProperties import Domain
    from mcbrat3d_b200.scatteringPhaseFunctions import henyeyGreenstein, new_PhaseFunctionTable
    n, delta, mu0, albedo = 8, 0.125, 0.5, 0.2
    edges = delta * np.arange(n + 1, dtype=np.float64)
    d = Domain(edges, edges, edges, surfaceAlbedo=albedo)
    ext = np.zeros((n, n, n)); ssa = np.zeros((n, n, n)); idx = np.zeros((n, n, n), np.int32)
    k = n
    for tau, omega, cells in layers:                                  # z index 0 is the bottom layer
        k -= cells
        ext[k:k + cells] = tau / (cells * delta); ssa[k:k + cells] = omega; idx[k:k + cells] = 1 if tau > 0 else 0
    assert k == 0
    d.addOpticalComponent("cloud", ext, ssa, idx, new_PhaseFunctionTable([henyeyGreenstein(0.85, 64)], key=[1.0]))
    d.getOpticalPropertiesByComponent()
    og = orc.OracleIntegrator(orc.OracleDomain(d, tableSize=10001))
    nb = 40
    tot, st = og.run_batches(nb, 5000, solarMu=mu0, solarAzimuth=0.0, iseed=10, rank=1, thread=0)
    d.tabulateInversePhaseFunctions(10001)
    want = layered_fluxes([(t, w) for t, w, _ in layers], table_moments(d.inversePhaseFunctions[0]), mu0, albedo, nStreams=96)
    for key, w in zip(("meanFluxUpStats", "meanFluxDownStats", "meanFluxAbsorbedStats"), want):
        m, e = _fin(orc, st, key, tot, nb)
        assert abs(m[0] - w) < 4.0 * e[0] + 2e-4, (key, m[0], w, e[0])


@pytest.mark.parametrize("name,layers,sfcTemp,albedo", [
    ("isothermal", [(5.0, 0.5, 290.0, 8)], 290.0, 0.0),
    ("lapse", [(1.0, 0.3, 230.0, 2), (2.0, 0.6, 260.0, 3), (1.5, 0.2, 285.0, 3)], 300.0, 0.1),
    ("warm_layer_aloft", [(0.5, 0.0, 300.0, 2), (0.0, 0.0, 250.0, 2), (3.0, 0.9, 250.0, 4)], 270.0, 0.3)])
def test_thermal_fluxes_match_adding_doubling_with_sources(orc, name, layers, sfcTemp, albedo):
    """The thermal source (emission_weighting EMI:424-550, new_PhotonStream_BBEmission ILL:431-522, the -1 emission
    bookkeeping INT:505-508) against doubling + adding WITH SOURCES (tests/adding_doubling.py::thermal_fluxes): layers of
    different temperature, optical depth and albedo (one of them empty) over an emitting, partly reflecting surface.
    Checked: the share of the atmosphere in the emitted power, the flux leaving the top, the flux reaching the surface
    and absorbed-minus-emitted in the atmosphere, all per emitted photon.  layers: (tau, omega, T, cells), TOP FIRST."""
    from adding_doubling import table_moments, thermal_fluxes
    from mcbrat3d_b200.opticalProperties import Domain
    from mcbrat3d_b200.scatteringPhaseFunctions import henyeyGreenstein, new_PhaseFunctionTable
    n, delta, lam = 8, 0.125, 10.0
    edges = delta * np.arange(n + 1, dtype=np.float64)
    temps = np.zeros((n, n, n)); ext = np.zeros((n, n, n)); ssa = np.zeros((n, n, n)); idx = np.zeros((n, n, n), np.int32)
    k = n
    for tau, omega, T, cells in layers:                                # z index 0 is the bottom layer
        k -= cells
        temps[k:k + cells] = T; ext[k:k + cells] = tau / (cells * delta); ssa[k:k + cells] = omega
        idx[k:k + cells] = 1 if tau > 0 else 0
    assert k == 0
    d = Domain(edges, edges, edges, temps=temps, surfaceAlbedo=albedo, lambda_um=lam)
    d.addOpticalComponent("cloud", ext, ssa, idx, new_PhaseFunctionTable([henyeyGreenstein(0.6, 32)], key=[1.0]))
    d.getOpticalPropertiesByComponent()
    od = orc.OracleDomain(d, tableSize=10001)
    frac, cdf, flux = od.emission_weighting(d.temps, d.lambda_um, sfcTemp)
    d.tabulateInversePhaseFunctions(10001)
    want = thermal_fluxes([(t, w, T) for t, w, T, _ in layers], table_moments(d.inversePhaseFunctions[0]), lam, sfcTemp,
                          albedo, nStreams=96)
    assert abs(frac - want[0]) < 1e-9
    og = orc.OracleIntegrator(od, LW_flag=1.0)
    nb = 40
    tot, st = og.run_batches(nb, 5000, source=1, fracAtmsPower=frac, voxelCDF=cdf)
    for key, w in zip(("meanFluxUpStats", "meanFluxDownStats", "meanFluxAbsorbedStats"), want[1:]):
        m, e = _fin(orc, st, key, tot, nb)
        assert abs(m[0] - w) < 4.0 * e[0] + 3e-4, (key, m[0], w, e[0])


@pytest.mark.parametrize("rr", [False, True], ids=["plain", "rr"])
@pytest.mark.parametrize("tau,omega,g,mu0,albedo", [(2.0, 0.9, 0.6, 0.5, 0.3), (5.0, 0.99, 0.85, 0.7, 0.0)])
def test_nadir_radiance_with_anisotropic_scattering_matches_adding_doubling(orc, tau, omega, g, mu0, albedo, rr):
    """The view along the vertical needs only the azimuthally averaged field, so the solver also pins the local estimate
    for a forward-peaked phase function (forward-table look-up by scattering angle, INT:1834-1873).  The photon loop
    scatters by the law the inverse table encodes while the local estimate evaluates the analytic function, so the
    solver is run with both sets of Legendre moments and their difference is added to the tolerance."""
    from adding_doubling import slab_fluxes, table_moments
    d, case = domains.homogeneous_slab(ssa=omega, tau=tau, albedo=albedo, g=g, n=8, delta=0.125)
    og = orc.OracleIntegrator(orc.OracleDomain(d, tableSize=10001, forward=True), useRussianRouletteForIntensity=int(rr), zetaMin=0.3)
    og.set_views([1.0], [0.0])
    nb = 30
    tot, st = og.run_batches(nb, 4000, solarMu=mu0, solarAzimuth=0.0, iseed=10, rank=1, thread=0)
    m, e = orc.finalise(st["radianceStats"], 1.0, tot, nb)
    got, err = m.mean(), np.sqrt((e ** 2).sum()) / e.size
    d.tabulateInversePhaseFunctions(10001)
    a = slab_fluxes(tau, omega, table_moments(d.inversePhaseFunctions[0]), mu0, albedo, nStreams=96, muOut=[1.0])[3][0]
    b = slab_fluxes(tau, omega, g ** np.arange(64), mu0, albedo, nStreams=96, muOut=[1.0])[3][0]
    assert abs(got - 0.5 * (a + b)) < 4.0 * err + abs(a - b) + 1e-3 * a, (got, a, b, err)


@pytest.mark.parametrize("grid", ["irregular", "stretched"])
def test_layered_fluxes_do_not_depend_on_the_3d_grid(orc, grid):
    """Multiple scattering through the 3-D machinery against the plane-parallel solver: a horizontally homogeneous stack
    (layers of different optical depth and albedo, one empty) laid out on grids the reference marches on its IRREGULAR
    path -- spacings not representable in single precision (q1-q3) and spacings that differ from cell to cell in x, y and
    z -- with 8 x 6 columns and an oblique sun at 30 degrees azimuth, so photons wrap around both periodic boundaries many
    times between scatterings.  Whatever the grid, the domain-mean fluxes must be the doubling + adding solution for
    the stack (tests/adding_doubling.py::layered_fluxes): 4 sigma + 2e-4 at 2e5 photons."""
    from adding_doubling import layered_fluxes, table_moments
    from mcbrat3d_b200.opticalProperties import Domain
    from mcbrat3d_b200.scatteringPhaseFunctions import henyeyGreenstein, new_PhaseFunctionTable
    nx, ny, nz, mu0, albedo = 8, 6, 7, 0.6, 0.3
    if grid == "irregular":
        xE = 0.05 * np.arange(nx + 1); yE = 0.03 * np.arange(ny + 1); zE = 0.04 * np.arange(nz + 1)
    else:
        xE = np.concatenate([[0.0], np.cumsum(0.05 * (1 + 0.3 * np.sin(1.0 + np.arange(nx))))])
        yE = np.concatenate([[0.0], np.cumsum(0.03 * (1 + 0.25 * np.cos(0.5 + np.arange(ny))))])
        zE = np.concatenate([[0.0], np.cumsum(0.025 * 1.15 ** np.arange(nz))])
    layers = [(1.5, 0.95, 2), (0.0, 0.0, 1), (2.5, 0.8, 3), (0.7, 1.0, 1)]            # (tau, omega, cells), TOP FIRST
    d = Domain(xE, yE, zE, surfaceAlbedo=albedo)
    ext = np.zeros((nz, ny, nx)); ssa = np.zeros((nz, ny, nx)); idx = np.zeros((nz, ny, nx), np.int32)
    k = nz
    for tau, omega, cells in layers:                                                  # z index 0 is the bottom layer
        k -= cells
        ext[k:k + cells] = tau / (zE[k + cells] - zE[k]); ssa[k:k + cells] = omega; idx[k:k + cells] = 1 if tau > 0 else 0
    assert k == 0
    d.addOpticalComponent("cloud", ext, ssa, idx, new_PhaseFunctionTable([henyeyGreenstein(0.85, 64)], key=[1.0]))
    d.getOpticalPropertiesByComponent()
    og = orc.OracleIntegrator(orc.OracleDomain(d, tableSize=10001))
    nb = 40
    tot, st = og.run_batches(nb, 5000, solarMu=mu0, solarAzimuth=30.0, iseed=10, rank=1, thread=0)
    d.tabulateInversePhaseFunctions(10001)
    want = layered_fluxes([(t, w) for t, w, _ in layers], table_moments(d.inversePhaseFunctions[0]), mu0, albedo, nStreams=96)
    for key, w in zip(("meanFluxUpStats", "meanFluxDownStats", "meanFluxAbsorbedStats"), want):
        m, e = _fin(orc, st, key, tot, nb)
        assert abs(m[0] - w) < 4.0 * e[0] + 2e-4, (grid, key, m[0], w, e[0])
    # ... and no column is special: the per-column maps scatter about the domain mean like noise
    for key in ("fluxUpStats", "fluxDownStats"):
        m, e = orc.finalise(st[key], 1.0, tot, nb)
        z = (m - m.mean()) / e
        assert np.sqrt(np.mean(z ** 2)) < 1.5 and np.abs(z).max() < 4.5, (grid, key, z)
