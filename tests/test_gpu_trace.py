"""North-star criterion (a): with injected random numbers the CUDA reference-arithmetic kernel
reproduces the oracle's cell indices, event sequence and table look-ups bit-exactly, and path
lengths / weights within 1e-6 relative.  Everything goes through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from common import assert_events_equal, injected_randoms, oracle_weights, stable_seed, trace_cases
from mcbrat3d_b200 import _lib
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.monteCarloRadiativeTransfer import (finalize_Integrator, new_Integrator, specifyParameters,
                                                       tracePhotons)
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence

pytestmark = pytest.mark.gpu

CASES = trace_cases()


@pytest.mark.parametrize("name,lazy,source", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("views,rr", [(False, False), (True, False), (True, True)], ids=["flux", "le", "le_rr"])
def test_trace_parity(orc, name, lazy, source, views, rr):
    dom, case = lazy.get()
    g = new_Integrator(dom)
    try:
        if views:
            mus = case.get("intensityMus", [1.0, 0.5]); phis = case.get("intensityPhis", [0.0, 0.0])
            specifyParameters(g, intensityMus=mus, intensityPhis=phis, computeIntensity=True,
                              useRussianRouletteForIntensity=rr, zetaMin=0.3)
        specifyParameters(g, minInverseTableSize=10001, minForwardTableSize=10001,
                          LW_flag=1.0 if source else -1.0)
        od = orc.OracleDomain(dom, tableSize=10001, forward=views)
        og = orc.OracleIntegrator(od, useRussianRouletteForIntensity=int(rr), zetaMin=0.3,
                                  LW_flag=1.0 if source else -1.0)
        if views:
            og.set_view_cosines(g.intensityDirections)
        n, stride = 1500, 400
        rn = injected_randoms(n, stride, seed=stable_seed(name))
        rs = new_RandomNumberSequence([10, 1, 0])
        if source == 0:
            ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], n, rs)
            want = og.trace(rn, 0, case["solarMu"], case["solarAzimuth"], maxEvents=n * 1024)
        else:
            w = oracle_weights(orc, od, dom, case.get("surfaceTemp", 300.0))
            ps = new_PhotonStream(theseWeights=w, numberOfPhotons=n, randomNumbers=rs)
            want = og.trace(rn, 1, fracAtmsPower=w.fracAtmsPower, voxelCDF=w.voxelWeights, maxEvents=n * 1024)
        got, raw = tracePhotons(g, dom, ps, rn, maxEventsPerPhoton=1024)
        assert len(want) > (5 if source == 0 else 2) * n      # thermal photons in an absorbing slab live a few events
        assert_events_equal(got, want, "%s views=%s rr=%s" % (name, views, rr))
        # the tallies of the traced photons agree too (f64 sums on the GPU, f32 in the oracle)
        ot = og.raw_tallies()
        np.testing.assert_allclose(raw[: ot.size], ot, rtol=2e-5, atol=2e-4)
        assert raw[-1] == n
    finally:
        finalize_Integrator(g)


@pytest.mark.parametrize("name,lazy,source", [c for c in CASES if c[0] in ("C1", "T_irr", "C2")],
                         ids=[c[0] for c in CASES if c[0] in ("C1", "T_irr", "C2")])
@pytest.mark.parametrize("views", [False, True], ids=["flux", "le"])
def test_trace_parity_max_cross_section(orc, name, lazy, source, views):
    """useRayTracing=.false. (INT:564-571, 578-585, 624-631, 709-710; makePeriodic returns default real,
    quirk q15): same bit-exact criterion, including the mathematical (null) collisions."""
    dom, case = lazy.get()
    g = new_Integrator(dom)
    try:
        if views:
            mus = case.get("intensityMus", [1.0, 0.5]); phis = case.get("intensityPhis", [0.0, 0.0])
            specifyParameters(g, intensityMus=mus, intensityPhis=phis, computeIntensity=True)
        specifyParameters(g, minInverseTableSize=10001, minForwardTableSize=10001, useRayTracing=False)
        od = orc.OracleDomain(dom, tableSize=10001, forward=views)
        og = orc.OracleIntegrator(od, useRayTracing=0)
        if views:
            og.set_view_cosines(g.intensityDirections)
        n, stride = 800, 600
        rn = injected_randoms(n, stride, seed=stable_seed(name, 7))
        rs = new_RandomNumberSequence([10, 1, 0])
        ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], n, rs)
        want = og.trace(rn, 0, case["solarMu"], case["solarAzimuth"], maxEvents=n * 2048)
        got, raw = tracePhotons(g, dom, ps, rn, maxEventsPerPhoton=2048)
        assert len(want) > 3 * n
        assert (want["kind"] == 10).any() or name == "C1"         # heterogeneous domains see null collisions
        assert_events_equal(got, want, "%s max-xsec views=%s" % (name, views))
        ot = og.raw_tallies()
        np.testing.assert_allclose(raw[: ot.size], ot, rtol=2e-5, atol=2e-4)
    finally:
        finalize_Integrator(g)


def test_trace_matches_committed_golden(orc):
    """The committed fixture (tests/golden/trace_T_irr.npz, written by tests/golden/make_golden.py
    from the oracle) pins both implementations against drift."""
    import os
    from mcbrat3d_b200 import domains
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "trace_T_irr.npz"))
    dom, case = domains.irregular_test_domain()
    g = new_Integrator(dom)
    try:
        specifyParameters(g, intensityMus=case["intensityMus"], intensityPhis=case["intensityPhis"],
                          computeIntensity=True, minInverseTableSize=9001, minForwardTableSize=9001)
        rs = new_RandomNumberSequence([10, 1, 0])
        rn = fx["rn"]
        ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], rn.shape[0], rs)
        got, _ = tracePhotons(g, dom, ps, rn, maxEventsPerPhoton=1024)
        assert_events_equal(got, fx["events"], "golden T_irr")
    finally:
        finalize_Integrator(g)


def test_philox_known_answer():
    """Philox4x32-10 KAT (Random123 kat_vectors): counter 0, key 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8;
    counter/key all-ones -> 408f276d 41c83b0e a20bc7c6 6d5451fd."""
    from mcbrat3d_b200 import domains
    dom, _ = domains.homogeneous_slab()
    g = new_Integrator(dom)
    try:
        out = (C.c_uint32 * 8)()
        assert g._lib.mcb_debug_philox(g.handle, 0, 0, 8, out) == 0
        assert [hex(v) for v in out[:4]] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
        # the next block bumps counter word 2 (draw block), not the photon id
        assert list(out[4:]) != list(out[:4])
    finally:
        finalize_Integrator(g)


@pytest.mark.parametrize("name,lazy,source", [c for c in CASES if c[0] in ("C2", "T_irr", "C3_mie")],
                         ids=[c[0] for c in CASES if c[0] in ("C2", "T_irr", "C3_mie")])
def test_trace_parity_hybrid_tables_and_contribution_limit(orc, name, lazy, source):
    """Local estimation with the hybrid (Gaussian forward peak) tables for scattering orders above
    numOrdersOrigPhaseFunIntenCalcs (INT:1715-1724, OPT:1936-2050) and with limited contributions
    (INT:1815-1826): bit-exact events, and the raw tallies including intensityExcess."""
    dom, case = lazy.get()
    g = new_Integrator(dom)
    try:
        mus = case.get("intensityMus", [1.0, 0.5]); phis = case.get("intensityPhis", [0.0, 0.0])
        specifyParameters(g, intensityMus=mus, intensityPhis=phis, computeIntensity=True,
                          useHybridPhaseFunsForIntenCalcs=True, hybridPhaseFunWidth=7.0, numOrdersOrigPhaseFunIntenCalcs=2,
                          limitIntensityContributions=True, maxIntensityContribution=0.05,
                          minInverseTableSize=10001, minForwardTableSize=10001)
        od = orc.OracleDomain(dom, tableSize=10001, forward=True, hybrid=True, hybridWidth=7.0)
        og = orc.OracleIntegrator(od, useHybridPhaseFunsForIntenCalcs=1, numOrdersOrigPhaseFunIntenCalcs=2,
                                  limitIntensityContributions=1, maxIntensityContribution=0.05)
        og.set_view_cosines(g.intensityDirections)
        n, stride = 1000, 400
        rn = injected_randoms(n, stride, seed=stable_seed(name, 11))
        rs = new_RandomNumberSequence([10, 1, 0])
        ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], n, rs)
        want = og.trace(rn, 0, case["solarMu"], case["solarAzimuth"], maxEvents=n * 1024)
        got, raw = tracePhotons(g, dom, ps, rn, maxEventsPerPhoton=1024)
        assert_events_equal(got, want, "%s hybrid+limit" % name)
        ot = og.raw_tallies()
        assert ot[-(len(mus) * (od.nc + 1)):].sum() > 0                      # some contribution was capped
        np.testing.assert_allclose(raw[: ot.size], ot, rtol=2e-5, atol=2e-4)
    finally:
        finalize_Integrator(g)
