"""Device-side staging and read-back (csrc/mcb_stage.cu) against the oracle / a NumPy restatement:
argument checks of addOpticalComponent, the normalisation of INT:328-388 (bit-exact, single
precision), the excess redistribution of INT:294-322, and the emission CDF of EMI:498-522."""
import ctypes as C

import numpy as np
import pytest

from mcbrat3d_b200 import _lib, domains
from mcbrat3d_b200.emissionAndBroadBandWeights import Weights, emission_weighting, fetchVoxelWeights
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.monteCarloRadiativeTransfer import (computeRadiativeTransfer, finalize_Integrator, new_Integrator,
                                                       reportResults, specifyParameters)
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence

pytestmark = pytest.mark.gpu
f32 = np.float32


def _raw(g):
    dptr = C.c_void_p(); nd = C.c_int64(0)
    g._check(g._lib.mcb_tally_buffer(g.handle, C.byref(dptr), C.byref(nd)), "tally")
    raw = np.empty(nd.value, dtype=np.float64)
    g._check(g._lib.mcb_get_raw_tallies(g.handle, _lib.ptr(raw, C.c_double), nd.value), "raw")
    return raw


def _set_optics(g, d, **over):
    a = dict(totalExt=d.totalExt, cumulativeExt=d.cumulativeExt, ssa=d.ssa, phaseFunctionIndex=d.phaseFunctionIndex)
    a.update(over)
    return g._lib.mcb_set_optics(g.handle, d.cumulativeExt.shape[0], _lib.ptr(a["totalExt"], C.c_double),
                                 _lib.ptr(a["cumulativeExt"], C.c_double), _lib.ptr(a["ssa"], C.c_double),
                                 _lib.ptr(a["phaseFunctionIndex"], C.c_int32), 0.0)


def test_optics_argument_checks_run_on_device():
    d, _ = domains.step_cloud()
    d.getOpticalPropertiesByComponent()
    g = new_Integrator(d)
    buf = C.create_string_buffer(256)
    try:
        bad = d.totalExt.copy(); bad.ravel()[17] = -1.0
        assert _set_optics(g, d, totalExt=bad) != 0
        g._lib.mcb_last_error(g.handle, buf, 256); assert b"extinction must be >= 0" in buf.value
        bad = d.ssa.copy(); bad.ravel()[5] = 1.5
        assert _set_optics(g, d, ssa=bad) != 0
        g._lib.mcb_last_error(g.handle, buf, 256); assert b"singleScatteringAlbedo must be between 0 and 1" in buf.value
        bad = d.ssa.copy(); bad.ravel()[5] = np.nan
        assert _set_optics(g, d, ssa=bad) != 0
        bad = d.phaseFunctionIndex.copy(); bad.ravel()[3] = -2
        assert _set_optics(g, d, phaseFunctionIndex=bad) != 0
        g._lib.mcb_last_error(g.handle, buf, 256); assert b"phase function index is out of bounds" in buf.value
        # a failed staging leaves the problem unspecified, a good one repairs it
        done = C.c_int64(0)
        assert g._lib.mcb_set_solar_source(g.handle, 0.5, 0.0) == 0
        assert g._lib.mcb_run_batch(g.handle, 10, 1, 0, C.byref(done)) != 0
        assert _set_optics(g, d) == 0
    finally:
        finalize_Integrator(g)


@pytest.mark.parametrize("which", ["regular", "irregular"])
def test_normalisation_bit_exact(which):
    """reportResults' arrays == the reference's normalisation (INT:328-388) applied in NumPy to the raw
    f64 tallies, bit for bit (quirks q11, q12 included)."""
    if which == "regular":
        d, case = domains.step_cloud(ssa=0.9, solarMu=0.5)
    else:
        d, case = domains.irregular_test_domain()
    g = new_Integrator(d)
    try:
        specifyParameters(g, intensityMus=[1.0, 0.5], intensityPhis=[0.0, 180.0], computeIntensity=True)
        rs = new_RandomNumberSequence(3)
        n = 30011
        ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], n, rs)
        assert computeRadiativeTransfer(g, d, rs, ps, n) == n
        r = reportResults(g, fluxUp=True, fluxDown=True, fluxAbsorbed=True, volumeAbsorption=True, intensity=True,
                          intensityByComponent=True)
        raw = _raw(g)
        nx, ny, nz, nc, nDir = g.numX, g.numY, g.numZ, g.numComps, 2
        cols = nx * ny
        N = f32(n)
        if which == "regular":
            nppc = np.full((ny, nx), N / f32(nx * ny), dtype=f32)
        else:
            dx = np.diff(d.xPosition); dy = np.diff(d.yPosition)
            area = (d.xPosition[-1] - d.xPosition[0]) * (d.yPosition[-1] - d.yPosition[0])
            nppc = ((dy[:, None] * dx[None, :]) / area).astype(f32) * N
        o = 0
        for name in ("fluxUp", "fluxDown", "fluxAbsorbed"):
            want = raw[o:o + cols].astype(f32).reshape(ny, nx) / nppc
            assert np.array_equal(r[name], want), name
            o += cols
        vol = raw[o:o + cols * nz].astype(f32).reshape(nz, ny, nx); o += cols * nz
        dz = np.diff(d.zPosition)[:, None, None]
        want = (vol.astype(np.float64) / (nppc.astype(np.float64)[None] * dz * 1000.0)).astype(f32)
        assert np.array_equal(r["volumeAbsorption"], want)
        inten = raw[o:o + cols * nDir].astype(f32).reshape(nDir, ny, nx); o += cols * nDir
        assert np.array_equal(r["intensity"], inten / nppc[None])
        byc = raw[o:o + cols * nDir * (nc + 1)].astype(f32).reshape(nc + 1, nDir, ny, nx)
        want = byc.copy(); want[1:] = byc[1:] / nppc[None, None]
        assert np.array_equal(r["intensityByComponent"], want)
        assert r["intensity"].sum() > 0
    finally:
        finalize_Integrator(g)


def test_excess_redistribution_conserves_radiance():
    """limitIntensityContributions (INT:1815-1826, 294-322): capped contributions are handed back in
    proportion to each component's map, so the domain-mean radiance does not change."""
    d, case = domains.step_cloud(ssa=1.0, solarMu=0.5)
    out = []
    for limit in (False, True):
        g = new_Integrator(d)
        try:
            specifyParameters(g, intensityMus=[1.0, 0.5], intensityPhis=[0.0, 0.0], computeIntensity=True,
                              limitIntensityContributions=limit, maxIntensityContribution=0.05)
            rs = new_RandomNumberSequence(7)
            n = 100000
            ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], n, rs)
            computeRadiativeTransfer(g, d, rs, ps, n)
            r1 = reportResults(g, meanIntensity=True, intensityByComponent=True, intensity=True)
            r2 = reportResults(g, meanIntensity=True)                 # idempotent: the excess was zeroed
            assert np.array_equal(r1["meanIntensity"], r2["meanIntensity"])
            assert np.allclose(r1["intensity"], r1["intensityByComponent"].sum(axis=0), rtol=2e-5)
            out.append(r1["meanIntensity"].astype(np.float64))
        finally:
            finalize_Integrator(g)
    assert np.allclose(out[0], out[1], rtol=1e-5), out


@pytest.mark.parametrize("name", ["C4", "T_irr"])
def test_emission_cdf_built_on_device_matches_oracle(orc, name):
    d, case = domains.homogeneous_lw() if name == "C4" else domains.irregular_test_domain()
    d.getOpticalPropertiesByComponent()
    sfcTemp = case.get("surfaceTemp", 300.0)
    frac, cdf, flux = orc.OracleDomain(d, tableSize=9001).emission_weighting(d.temps, d.lambda_um, sfcTemp)
    g = new_Integrator(d)
    try:
        w = Weights()
        got_flux = emission_weighting(d, w, sfcTemp, thisIntegrator=g)
        got = fetchVoxelWeights(w)
        assert got.shape == cdf.shape
        assert got.ravel()[-1] == 1.0
        assert np.all(np.diff(got.ravel()) >= 0.0)
        assert np.allclose(got, cdf, rtol=1e-12, atol=0.0)
        assert abs(w.fracAtmsPower - frac) <= 1e-12 * abs(frac)
        assert abs(got_flux - flux) <= 1e-12 * abs(flux)
        # and it drives the thermal source: one batch, isothermal-ish closure is covered by test_gpu_stats
        specifyParameters(g, LW_flag=1.0)
        rs = new_RandomNumberSequence(11)
        ps = new_PhotonStream(theseWeights=w, numberOfPhotons=20000, randomNumbers=rs)
        assert computeRadiativeTransfer(g, d, rs, ps, 20000) == 20000
        # a cold cell switches the atmosphere off (EMI:498) -> only the surface emits
        cold = d.temps.copy(); cold.ravel()[3] = 0.0
        d2 = d; keep = d.temps; d2.temps = cold
        try:
            w2 = Weights()
            emission_weighting(d2, w2, sfcTemp, thisIntegrator=g)
            assert w2.fracAtmsPower == 0.0 and not fetchVoxelWeights(w2).any()
        finally:
            d2.temps = keep
        # a read-only temperature array is uploaded once: the next build passes NULL ("the temperatures staged
        # before") and must give the same CDF; a different array of the same shape is uploaded again
        d.temps.setflags(write=False)
        try:
            wa, wb = Weights(), Weights()
            emission_weighting(d, wa, sfcTemp, thisIntegrator=g)
            assert g._stagedTemps is not None
            fa = emission_weighting(d, wb, sfcTemp, thisIntegrator=g)                 # NULL path
            assert np.array_equal(fetchVoxelWeights(wb), got) and fa == got_flux
            warm = d.temps + 5.0; warm.setflags(write=False)
            keep = d.temps; d.temps = warm
            try:
                wc = Weights()
                assert emission_weighting(d, wc, sfcTemp, thisIntegrator=g) > got_flux
            finally:
                d.temps = keep
        finally:
            d.temps.setflags(write=True)
    finally:
        finalize_Integrator(g)


def test_thermal_source_without_temperatures_is_an_error():
    d, case = domains.homogeneous_lw()
    g = new_Integrator(d)
    try:
        from mcbrat3d_b200.monteCarloRadiativeTransfer import _stage_domain
        _stage_domain(g, d)
        frac = C.c_double(0.0); flux = C.c_double(0.0)
        assert g._lib.mcb_build_thermal_source(g.handle, None, 10.0, 300.0, C.byref(frac), C.byref(flux)) != 0
    finally:
        finalize_Integrator(g)


@pytest.mark.parametrize("views", [False, True], ids=["flux", "le"])
def test_device_batch_statistics_match_host_loop(views):
    """mcb_run_batches + mcb_get_statistics (DRV:949-1052, 1188-1228 on the device) against the host loop
    computeRadiativeTransfer -> reportResults -> BatchStatistics over the SAME photons."""
    from mcbrat3d_b200.batchStatistics import BatchStatistics, computeRadiativeTransferBatches, reportStatistics
    d, case = domains.step_cloud(ssa=0.95, solarMu=0.5)
    nb, n = 8, 20000
    g = new_Integrator(d)
    try:
        if views:
            specifyParameters(g, intensityMus=[1.0, 0.5], intensityPhis=[0.0, 180.0], computeIntensity=True)
        rs = new_RandomNumberSequence(5)
        bs = BatchStatistics()
        for _ in range(nb):
            ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], n, rs)
            done = computeRadiativeTransfer(g, d, rs, ps, n)
            bs.accumulate(reportResults(g, meanFluxUp=True, meanFluxDown=True, meanFluxAbsorbed=True, fluxUp=True,
                                        fluxDown=True, fluxAbsorbed=True, absorbedProfile=True, volumeAbsorption=True,
                                        intensity=views), done)
        hm, he = bs.finalise(2.5)
        rs2 = new_RandomNumberSequence(5)
        ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], nb * n, rs2)
        assert computeRadiativeTransferBatches(g, d, rs2, ps, n, nb) == nb * n
        dm, de, tot, done = reportStatistics(g, solarFlux=2.5)
        assert (tot, done) == (nb * n, nb)
        for k in hm:
            np.testing.assert_allclose(dm[k], hm[k], rtol=2e-6, atol=1e-12, err_msg=k)
            np.testing.assert_allclose(de[k], he[k], rtol=5e-3, atol=1e-8, err_msg=k + " stderr")
        assert dm["meanFluxUp"] > 0 and de["meanFluxUp"] > 0
        # a second call adds batches to the same moments
        ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], 2 * n, rs2)
        computeRadiativeTransferBatches(g, d, rs2, ps, n, 2)
        _, _, tot, done = reportStatistics(g)
        assert (tot, done) == ((nb + 2) * n, nb + 2)
    finally:
        finalize_Integrator(g)


def test_inverse_tables_built_on_device_match_oracle(orc):
    """computeInversePhaseFunction (INV:113-168) as a kernel, against the oracle's C
    restatement -- Legendre (Lobatto nodes), tabulated (native angles) and the 2-moment Rayleigh function."""
    from mcbrat3d_b200.inversePhaseFunctions import inversion_inputs
    d, case = domains.landsat_cloud(ssa=0.99, nxy=16, mie=True)
    g = new_Integrator(d)
    try:
        specifyParameters(g, minInverseTableSize=10001, buildTablesOnDevice=True)
        rs = new_RandomNumberSequence(2)
        ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], 50000, rs)
        assert computeRadiativeTransfer(g, d, rs, ps, 50000) == 50000      # the run uses the device-built tables
        r_dev = reportResults(g, meanFluxUp=True, meanFluxDown=True, meanFluxAbsorbed=True)
        total = mismatch = 0
        for c, tab in enumerate(d.forwardTables):
            nE = len(tab.phaseFunctions)
            got = np.empty((nE, 10001), dtype=f32)
            g._check(g._lib.mcb_get_inverse_table(g.handle, c + 1, _lib.ptr(got, C.c_float), got.size), "get")
            for e, pf in enumerate(tab.phaseFunctions):
                mus, vals = inversion_inputs(pf)
                want = orc.inverse_phase_function(mus, vals, 10001)
                assert want[0] == f32(np.pi) or abs(want[0] - np.pi) < 1e-6
                assert got[e, -1] == 0.0
                bad = got[e] != want
                mismatch += int(bad.sum()); total += want.size
                if bad.any():                                                # double-precision acos is 1 ulp, not exact
                    assert np.abs(got[e][bad] - want[bad]).max() <= 2.4e-7 * np.maximum(want[bad], 1.0).max()
        assert mismatch <= 1e-4 * total, (mismatch, total)
        # host-built tables give the same transport result (same photons, same tables up to the ulps above)
        specifyParameters(g, buildTablesOnDevice=False)
        rs = new_RandomNumberSequence(2)
        ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], 50000, rs)
        computeRadiativeTransfer(g, d, rs, ps, 50000)
        r_host = reportResults(g, meanFluxUp=True, meanFluxDown=True, meanFluxAbsorbed=True)
        for k in r_dev:
            assert abs(float(r_dev[k]) - float(r_host[k])) < 2e-4
    finally:
        finalize_Integrator(g)


def test_forward_tables_built_on_device_match_oracle(orc):
    """tabulateForwardPhaseFunctions (OPT:1872-1934) as a kernel for Legendre-stored tables (HG with 64 moments,
    the 2-moment Rayleigh function, a mixed two-component domain) against the oracle's C restatement."""
    for make in (lambda: domains.step_cloud(ssa=0.99, solarMu=0.5), lambda: domains.irregular_test_domain()):
        d, case = make()
        g = new_Integrator(d)
        try:
            specifyParameters(g, intensityMus=[1.0, 0.5], intensityPhis=[0.0, 90.0], computeIntensity=True,
                              minInverseTableSize=9001, minForwardTableSize=9001, buildTablesOnDevice=True)
            rs = new_RandomNumberSequence(4)
            ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], 20000, rs)
            assert computeRadiativeTransfer(g, d, rs, ps, 20000) == 20000
            r_dev = reportResults(g, meanIntensity=True)
            total = mismatch = 0
            for c, tab in enumerate(d.forwardTables):
                got = np.empty((len(tab.phaseFunctions), 9001), dtype=f32)
                g._check(g._lib.mcb_get_forward_table(g.handle, c + 1, _lib.ptr(got, C.c_float), got.size), "get")
                for e, pf in enumerate(tab.phaseFunctions):
                    want = orc.forward_phase_function(pf.legendreCoefficients, 9001)
                    bad = got[e] != want
                    mismatch += int(bad.sum()); total += want.size
                    np.testing.assert_allclose(got[e], want, rtol=2e-5, atol=1e-6)     # cos() of the device is 1 ulp, not exact
            assert mismatch <= 2e-3 * total, (mismatch, total)
            specifyParameters(g, buildTablesOnDevice=False)
            rs = new_RandomNumberSequence(4)
            ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], 20000, rs)
            computeRadiativeTransfer(g, d, rs, ps, 20000)
            r_host = reportResults(g, meanIntensity=True)
            np.testing.assert_allclose(r_dev["meanIntensity"], r_host["meanIntensity"], rtol=2e-3)
        finally:
            finalize_Integrator(g)


def test_every_forward_table_is_built_on_device_and_matches_oracle(orc):
    """The remainder of SURVEY 8(f) rank 2: forward tables of angle / value phase functions (SPF:499-527) and the hybrid
    tables with a Gaussian forward peak (computeHybridPhaseFunctions OPT:1936-2050: hunt + bisection for the transition
    angle) tabulated in HBM by mcb_build_forward_table_general, against the oracle's C restatements."""
    d, case = domains.landsat_cloud(ssa=0.99, nxy=16, mie=True)          # 16 angle / value entries + the Rayleigh moments
    nS = 9001
    angles = (np.arange(nS, dtype=f32) / f32(nS - 1) * f32(np.pi)).astype(f32)
    for width in (0.0, 7.0):
        g = new_Integrator(d)
        try:
            specifyParameters(g, intensityMus=[1.0, 0.5], intensityPhis=[0.0, 90.0], computeIntensity=True,
                              minInverseTableSize=9001, minForwardTableSize=nS, buildTablesOnDevice=True,
                              useHybridPhaseFunsForIntenCalcs=width > 0, hybridPhaseFunWidth=width if width > 0 else None)
            rs = new_RandomNumberSequence(4)
            ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], 20000, rs)
            assert computeRadiativeTransfer(g, d, rs, ps, 20000) == 20000
            r_dev = reportResults(g, meanIntensity=True)
            transitions = 0
            for c, tab in enumerate(d.forwardTables):
                got = np.empty((len(tab.phaseFunctions), nS), dtype=f32)
                g._check(g._lib.mcb_get_forward_table(g.handle, c + 1, _lib.ptr(got, C.c_float), got.size), "get")
                for e, pf in enumerate(tab.phaseFunctions):
                    if pf.storedAsLegendre():
                        orig = orc.phase_function_values(angles, legendreCoefficients=pf.legendreCoefficients)
                    else:
                        orig = orc.phase_function_values(angles, storedAngle=pf.scatteringAngle, storedValue=pf.value)
                    want, t = orc.hybrid_phase_function(angles, orig, width) if width > 0 else (orig, 0)
                    transitions += t > 0
                    # cos() / exp() of the device are 1 ulp, not exact: a few values differ in the last place, and the
                    # normalisation of the Gaussian part (a sum over the whole table) may carry that along
                    np.testing.assert_allclose(got[e], want, rtol=3e-5, atol=1e-6, err_msg="width %g comp %d entry %d" % (width, c, e))
                    if t > 0:
                        assert np.array_equal(got[e][t:], want[t:]) or np.abs(got[e][t:] - want[t:]).max() <= 2e-5 * np.abs(want).max()
            specifyParameters(g, buildTablesOnDevice=False)               # host mirror: same radiances
            rs = new_RandomNumberSequence(4)
            ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], 20000, rs)
            computeRadiativeTransfer(g, d, rs, ps, 20000)
            r_host = reportResults(g, meanIntensity=True)
            np.testing.assert_allclose(r_dev["meanIntensity"], r_host["meanIntensity"], rtol=2e-3)
        finally:
            finalize_Integrator(g)


def test_lobatto_inputs_built_on_device_match_oracle(orc):
    """computeLobattoTerms (NUM:27-114) + the phase function at the nodes (INV:97-112) in HBM: the inverse table of a
    Legendre-stored entry built by mcb_build_inverse_table_legendre equals the oracle's inversion of the oracle's inputs."""
    from mcbrat3d_b200.scatteringPhaseFunctions import henyeyGreenstein, new_PhaseFunctionTable, rayleigh
    d, case = domains.homogeneous_slab(ssa=0.99)
    g = new_Integrator(d)
    try:
        specifyParameters(g, minInverseTableSize=10001)
        rs = new_RandomNumberSequence(2)
        ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], 1000, rs)
        computeRadiativeTransfer(g, d, rs, ps, 1000)                       # stages the domain
        pfs = [henyeyGreenstein(0.85, 299), henyeyGreenstein(0.85, 64), henyeyGreenstein(0.6, 5), rayleigh(),
               henyeyGreenstein(0.9, 128), henyeyGreenstein(0.3, 1)]
        nCoef = np.array([pf.legendreCoefficients.size for pf in pfs], dtype=np.int32)
        coefs = np.ascontiguousarray(np.concatenate([pf.legendreCoefficients for pf in pfs]), dtype=f32)
        g._check(g._lib.mcb_build_inverse_table_legendre(g.handle, 1, 10001, len(pfs), _lib.ptr(nCoef, C.c_int32),
                                                         _lib.ptr(coefs, C.c_float)), "build")
        got = np.empty((len(pfs), 10001), dtype=f32)
        g._check(g._lib.mcb_get_inverse_table(g.handle, 1, _lib.ptr(got, C.c_float), got.size), "get")
        mismatch = 0
        for e, pf in enumerate(pfs):
            mus, vals = orc.inversion_inputs_legendre(pf.legendreCoefficients)
            want = orc.inverse_phase_function(mus, vals, 10001)
            bad = got[e] != want
            mismatch += int(bad.sum())
            if bad.any():
                assert np.abs(got[e][bad] - want[bad]).max() <= 2.4e-7 * np.maximum(want[bad], 1.0).max(), e
        assert mismatch <= 1e-4 * got.size, mismatch
    finally:
        finalize_Integrator(g)


def test_hybrid_tables_with_a_transition_match_oracle(orc):
    """Entries peaked enough for computeHybridPhaseFunctions to find a transition angle (HG g >= 0.9 for a 7-degree
    Gaussian), through the C ABI directly: Legendre-stored and angle / value entries in one table."""
    from mcbrat3d_b200.scatteringPhaseFunctions import henyeyGreenstein
    d, case = domains.homogeneous_slab(ssa=0.99)
    nS = 9001
    angles = (np.arange(nS, dtype=f32) / f32(nS - 1) * f32(np.pi)).astype(f32)
    ang = np.linspace(0.0, np.pi, 1441).astype(f32); ang[-1] = f32(np.pi)
    mu = np.cos(ang.astype(np.float64))
    hg = lambda gg: (1 - gg * gg) / (1 + gg * gg - 2 * gg * mu) ** 1.5
    tabulated = [(0.97 * hg(g1) + 0.03 * hg(-0.45)).astype(f32) for g1 in (0.9, 0.97)]
    legendre = [henyeyGreenstein(0.9, 64).legendreCoefficients, henyeyGreenstein(0.95, 256).legendreCoefficients,
                henyeyGreenstein(0.85, 64).legendreCoefficients]
    # entry order: Legendre, tabulated, Legendre, tabulated, Legendre
    kinds = ["l", "t", "l", "t", "l"]
    g = new_Integrator(d)
    try:
        rs = new_RandomNumberSequence(2)
        ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], 1000, rs)
        computeRadiativeTransfer(g, d, rs, ps, 1000)                       # stages the domain
        for width in (7.0, 2.0):
            li = iter(legendre); ti = iter(tabulated)
            entries = [next(li) if k == "l" else next(ti) for k in kinds]
            nCoef = np.array([e.size if k == "l" else 0 for e, k in zip(entries, kinds)], dtype=np.int32)
            nAng = np.array([0 if k == "l" else ang.size for k in kinds], dtype=np.int32)
            coefs = np.ascontiguousarray(np.concatenate([e for e, k in zip(entries, kinds) if k == "l"]), dtype=f32)
            angs = np.ascontiguousarray(np.concatenate([ang for k in kinds if k == "t"]), dtype=f32)
            vals = np.ascontiguousarray(np.concatenate([e for e, k in zip(entries, kinds) if k == "t"]), dtype=f32)
            g._check(g._lib.mcb_build_forward_table_general(g.handle, 1, nS, len(kinds), _lib.ptr(nCoef, C.c_int32),
                                                            _lib.ptr(coefs, C.c_float), _lib.ptr(nAng, C.c_int32),
                                                            _lib.ptr(angs, C.c_float), _lib.ptr(vals, C.c_float), C.c_float(width)), "build")
            got = np.empty((len(kinds), nS), dtype=f32)
            g._check(g._lib.mcb_get_forward_table(g.handle, 1, _lib.ptr(got, C.c_float), got.size), "get")
            found = 0
            for e, (ent, k) in enumerate(zip(entries, kinds)):
                orig = (orc.phase_function_values(angles, legendreCoefficients=ent) if k == "l" else
                        orc.phase_function_values(angles, storedAngle=ang, storedValue=ent))
                want, t = orc.hybrid_phase_function(angles, orig, width)
                found += t > 0
                # the transition index is decided by the sign of a difference of nearly equal numbers: the device's cos()
                # and exp() are 1 ulp, so it may land one table step away, where the two branches differ by < 1e-3
                changed = np.nonzero(np.abs(got[e] - want) > 3e-5 * np.abs(want) + 1e-6)[0]
                assert changed.size <= 2 and (changed.size == 0 or abs(int(changed[0]) - t) <= 2), (width, e, t, changed[:5])
                np.testing.assert_allclose(got[e], want, rtol=2e-3, atol=1e-6)
            assert found >= (4 if width == 7.0 else 1), (width, found)
    finally:
        finalize_Integrator(g)
