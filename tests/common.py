"""Shared helpers for the parity tests (tests only)."""
import numpy as np

from mcbrat3d_b200 import domains
from mcbrat3d_b200.emissionAndBroadBandWeights import Weights

INT_FIELDS = ("photon", "kind", "ix", "iy", "iz", "component", "phaseIndex", "angleIndex", "order", "nrn")
FLT_FIELDS = ("weight", "tau", "path", "x", "y", "z")
REL_TOL = 1e-6          # north-star criterion (a): path lengths and weights within 1e-6 relative


class _Lazy:
    """A (domain, case) pair built on first use: the full-size scenes are not constructed at collection time."""

    def __init__(self, make):
        self._make, self._value = make, None

    def get(self):
        if self._value is None:
            self._value = self._make()
        return self._value


def trace_cases():
    """(name, lazy (domain, case), source) for the fixed-random-number harness.  The headline configurations are in:
    C3 at full size (z0 = 200, empty cells with phase index 0, 299-term HG), C3 with the 16-entry Mie-like table and a
    Rayleigh component (nc = 2), a small C5 (nc = 2, molecular background in every cell, albedo 0.05)."""
    return [("C1", _Lazy(lambda: domains.homogeneous_slab(ssa=0.99)), 0),
            ("T_irr", _Lazy(lambda: domains.irregular_test_domain()), 0),
            ("C2", _Lazy(lambda: domains.step_cloud(ssa=0.99, solarMu=0.5)), 0),
            ("T_irr_LW", _Lazy(lambda: domains.irregular_test_domain()), 1),
            ("C3", _Lazy(lambda: domains.landsat_cloud(ssa=0.99)), 0),
            ("C3_mie", _Lazy(lambda: domains.landsat_cloud(ssa=0.99, mie=True)), 0),
            ("C5_small", _Lazy(lambda: domains.bench_domain(nxy=40, nz=48)), 0),
            ("C4_LW", _Lazy(lambda: domains.homogeneous_lw()), 1)]


def stable_seed(name, salt=0):
    """A seed that does not change from run to run (str hashes are salted per process)."""
    import zlib
    return (zlib.crc32(name.encode()) + salt) % 1000


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
