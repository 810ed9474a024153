"""The oracle's scattering step taken out of its event trace and tested as a distribution, with nothing but the tables
and elementary geometry: the polar angle of every scattering against the inverse table it is drawn from
(computeScatteringAngle, INT:1594-1621), the azimuth against uniformity and independence of the incoming direction
(next_direct, INT:1921-1948; makeDirectionCosines), the new direction against unit length."""
import numpy as np
from scipy import stats

import first_interaction as fi


def _scatterings(orc, n=40000, stride=64, seed=3):
    dom, med = fi.scene("irregular", albedo=0.0)
    od = orc.OracleDomain(dom, tableSize=9001)
    g = orc.OracleIntegrator(od)
    rn = np.random.default_rng(seed).random((n, stride), dtype=np.float32)
    ev = g.trace(rn, 0, fi.SOLAR_MU, fi.SOLAR_AZIMUTH, maxEvents=n * 40)
    # consecutive (birth | scatter) -> scatter pairs of the same photon: direction before and after
    k = np.nonzero((ev["kind"][1:] == 2) & (ev["photon"][1:] == ev["photon"][:-1]) & np.isin(ev["kind"][:-1], (1, 2)))[0]
    return dom, ev[k], ev[k + 1]


def test_scattering_angles_follow_the_inverse_tables(orc):
    dom, before, after = _scatterings(orc)
    a, b = before["dir"].astype(np.float64), after["dir"].astype(np.float64)
    assert np.abs(np.linalg.norm(b, axis=1) - 1.0).max() < 5e-6             # rotations keep unit length (single precision)
    theta = np.arccos(np.clip((a * b).sum(axis=1), -1.0, 1.0))
    tables = dom.inversePhaseFunctions                                        # per component: (entries, steps) angles, equal steps in probability
    for comp, entry in ((1, 1), (1, 2), (2, 1)):
        sel = (after["component"] == comp) & (after["phaseIndex"] == entry)
        assert sel.sum() > 3000, (comp, entry, sel.sum())
        T = np.asarray(tables[comp - 1][entry - 1], dtype=np.float64)
        assert T[0] > T[-1]                                                   # tabulated from pi (probability 0) down to 0
        u = np.interp(theta[sel], T[::-1], np.linspace(1.0, 0.0, T.size))   # probability integral transform through the table
        ks = stats.kstest(u, "uniform")
        assert ks.pvalue > 1e-3, (comp, entry, ks)
        # ... and the forward-scattering entry really is forward-peaked: mean cosine close to the stored asymmetry
        if (comp, entry) == (1, 1):
            assert abs(np.cos(theta[sel]).mean() - 0.85) < 0.01


def test_scattering_azimuth_is_uniform_and_independent_of_the_incoming_direction(orc):
    _, before, after = _scatterings(orc, seed=4)
    a, b = before["dir"].astype(np.float64), after["dir"].astype(np.float64)
    # azimuth of the new direction about the old one, measured in a frame built from the old direction alone
    ref = np.where(np.abs(a[:, 2:3]) < 0.9, np.array([[0.0, 0.0, 1.0]]), np.array([[1.0, 0.0, 0.0]]))
    e1 = np.cross(a, ref); e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = np.cross(a, e1)
    sinT = np.linalg.norm(np.cross(a, b), axis=1)
    ok = sinT > 1e-3                                                         # the azimuth of a forward hit is noise
    phi = np.arctan2((b * e2).sum(axis=1), (b * e1).sum(axis=1))[ok]
    assert ok.sum() > 50000
    ks = stats.kstest((phi + np.pi) / (2.0 * np.pi), "uniform")
    assert ks.pvalue > 1e-3, ks
    # independence of the incoming direction: the same test in four classes of incoming vertical cosine
    for lo, hi in ((-1.0, -0.5), (-0.5, 0.0), (0.0, 0.5), (0.5, 1.0)):
        s = (a[ok, 2] >= lo) & (a[ok, 2] < hi)
        assert s.sum() > 3000
        assert stats.kstest((phi[s] + np.pi) / (2.0 * np.pi), "uniform").pvalue > 1e-3, (lo, hi)
