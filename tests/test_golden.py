"""The oracle reproduces the committed fixtures bit for bit (CPU, no GPU needed)."""
import os

import numpy as np

from common import assert_events_equal
from mcbrat3d_b200 import domains
from mcbrat3d_b200.monteCarloRadiativeTransfer import makeDirectionCosines

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_oracle_trace_fixture(orc):
    fx = np.load(os.path.join(GOLDEN, "trace_T_irr.npz"))
    dom, case = domains.irregular_test_domain()
    od = orc.OracleDomain(dom, tableSize=9001, forward=True)
    og = orc.OracleIntegrator(od)
    f32 = np.float32
    dirs = np.stack([makeDirectionCosines(m, f32(p) * f32(3.14159265358979312) / f32(180.0))
                     for m, p in zip(case["intensityMus"], case["intensityPhis"])])
    og.set_view_cosines(dirs)
    ev = og.trace(fx["rn"], 0, case["solarMu"], case["solarAzimuth"], maxEvents=48 * 1024)
    assert_events_equal(ev, fx["events"], "oracle vs golden")
    assert np.array_equal(ev["weight"], fx["events"]["weight"])
    np.testing.assert_array_equal(og.raw_tallies(), fx["tallies"])
    kinds = set(ev["kind"].tolist())
    assert {1, 2, 3, 4, 8} <= kinds            # birth, scatter, surface, exit, local estimate all occur


def test_oracle_view_cosines_match_host_mirror(orc):
    """specifyParameters builds the view direction cosines in single precision (INT:1267-1269)."""
    dom, case = domains.step_cloud()
    og = orc.OracleIntegrator(orc.OracleDomain(dom, tableSize=9001, forward=True))
    og.set_views(case["intensityMus"], case["intensityPhis"])
    f32 = np.float32
    dirs = np.stack([makeDirectionCosines(m, f32(p) * f32(3.14159265358979312) / f32(180.0))
                     for m, p in zip(case["intensityMus"], case["intensityPhis"])])
    assert np.array_equal(og.view_cosines(), dirs)


def test_slab_flux_table_reproduces(orc):
    """MT19937-driven batches (seed (10,1,0), DRV:901) are deterministic: the committed
    tau/omega -> Fup, Fdn, Fabs table (planeParallel.f95:242 protocol) reproduces exactly."""
    want = np.load(os.path.join(GOLDEN, "slab_fluxes_mt19937.npy"))
    for row in want:
        d, c = domains.homogeneous_slab(ssa=row[0])
        g = orc.OracleIntegrator(orc.OracleDomain(d, tableSize=9001))
        tot, st = g.run_batches(10, 2000, solarMu=0.5, solarAzimuth=0.0, iseed=10, rank=1, thread=0)
        up, eu = orc.finalise(st["meanFluxUpStats"], 1.0, tot, 10)
        dn, ed = orc.finalise(st["meanFluxDownStats"], 1.0, tot, 10)
        ab, ea = orc.finalise(st["meanFluxAbsorbedStats"], 1.0, tot, 10)
        np.testing.assert_array_equal(np.array([up[0], eu[0], dn[0], ed[0], ab[0], ea[0]]), row[1:])
    # and the numbers are physical: conservative slab reflects more, closes to 1 with A = 0.2
    assert abs(want[0, 1] + 0.8 * want[0, 3] - 1.0) < 0.02
