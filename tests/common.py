"""Shared helpers for the parity tests (tests only)."""
import numpy as np

from mcbrat3d_b200 import domains
from mcbrat3d_b200.emissionAndBroadBandWeights import Weights

INT_FIELDS = ("photon", "kind", "ix", "iy", "iz", "component", "phaseIndex", "angleIndex", "order", "nrn")
FLT_FIELDS = ("weight", "tau", "path", "x", "y", "z")
REL_TOL = 1e-6          # north-star criterion (a): path lengths and weights within 1e-6 relative


def trace_cases():
    """(name, domain, case, source) for the fixed-random-number harness."""
    return [("C1", *domains.homogeneous_slab(ssa=0.99), 0),
            ("T_irr", *domains.irregular_test_domain(), 0),
            ("C2", *domains.step_cloud(ssa=0.99, solarMu=0.5), 0),
            ("T_irr_LW", *domains.irregular_test_domain(), 1)]


def injected_randoms(nPhotons, stride, seed):
    rng = np.random.default_rng(seed)
    rn = rng.random((nPhotons, stride), dtype=np.float32)
    # the reference's generator returns [0, 1] INCLUSIVE (RNG:286-300): exercise both ends
    rn[::7, 3] = 0.0
    rn[::11, 5] = 1.0
    rn[::13, 2] = 1.0
    rn[::17, 0] = 0.0
    rn[::19, 1] = 1.0
    return rn


def assert_events_equal(a, b, what=""):
    """Cell indices, event sequence, table look-ups: bit-exact.  Lengths/weights: 1e-6 relative."""
    assert len(a) == len(b), "%s: %d vs %d events" % (what, len(a), len(b))
    for f in INT_FIELDS:
        bad = np.nonzero(a[f] != b[f])[0]
        assert bad.size == 0, "%s: field %s differs at events %s" % (what, f, bad[:5])
    for f in FLT_FIELDS:
        x = a[f].astype(np.float64); y = b[f].astype(np.float64)
        ok = (x == y) | (np.abs(x - y) <= REL_TOL * np.maximum(np.abs(x), np.abs(y))) | (np.isnan(x) & np.isnan(y))
        assert ok.all(), "%s: field %s differs at events %s" % (what, f, np.nonzero(~ok)[0][:5])
    d = (a["dir"] == b["dir"]) | (np.abs(a["dir"] - b["dir"]) <= 1e-6) | (np.isnan(a["dir"]) & np.isnan(b["dir"]))
    assert d.all(), "%s: direction cosines differ" % what


def oracle_weights(orc, od, dom, sfcTemp=300.0):
    frac, cdf, flux = od.emission_weighting(dom.temps, dom.lambda_um, sfcTemp)
    return Weights(voxelWeights=cdf, fracAtmsPower=frac, spectrIntgrFlux=flux)
