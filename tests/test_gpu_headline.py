"""The headline configurations under the statistical parity test at a power that resolves small biases.

(1) Full-size C3 (the configuration BASELINE.json's metric is quoted on) against the CPU oracle run on all host
    cores: >= 2e6 photons a side for the fluxes and the absorption profile, the five I3RC view directions with the
    Russian-roulette local estimate for the radiances (criterion (b): 3 sigma of the combined standard error, batch
    statistics formed as the driver does, DRV:1023-1052, 1188-1228).
(2) The throughput kernels against the reference-arithmetic kernel (which is trace-exact against the oracle) at 1e8
    photons: per-column fluxUp / fluxDown maps, the absorption profile and per-view radiance maps.  Independent seeds,
    so the per-column z-scores must look like unit normals: ensemble RMS < 1.1, mean compatible with 0 (a relative
    bias of 1e-3 in a flux map would move the mean z by 0.05), domain means within 3.5 sigma (sigma ~ 1e-4 relative).
"""
import os

import numpy as np
import pytest

from mcbrat3d_b200 import domains
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.monteCarloRadiativeTransfer import (MCB_ARITH_FAST, MCB_ARITH_REFERENCE, MCB_KERNEL_PARK, MCB_KERNEL_POOL,
                                                       computeRadiativeTransfer, finalize_Integrator, getCounters,
                                                       new_Integrator, reportResults, specifyParameters)
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence
from test_gpu_stats import assert_within

pytestmark = pytest.mark.gpu


def _gpu_rows(dom, case, nb, n, seed, views=False, want=(), **params):
    """nb batches of n photons; returns {name: (nb, ...) array} of the normalised results of every batch."""
    g = new_Integrator(dom)
    try:
        if views:
            specifyParameters(g, intensityMus=case["intensityMus"], intensityPhis=case["intensityPhis"], computeIntensity=True,
                              useRussianRouletteForIntensity=True, zetaMin=0.3)
        specifyParameters(g, minInverseTableSize=10001, minForwardTableSize=10001, **params)
        rs = new_RandomNumberSequence(list(seed))
        rows = {k: [] for k in want}
        for _ in range(nb):
            ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], n, rs)
            assert computeRadiativeTransfer(g, dom, rs, ps, n) == n
            res = reportResults(g, **{k: True for k in want})
            for k in want:
                rows[k].append(np.asarray(res[k], np.float64).copy())
        # The reference drops a photon whose march reports a non-positive step (OPT:1719-1722, INT:562-563): the
        # reference-arithmetic kernel reproduces that (a few per 1e6 photons on C3); the throughput kernels never do.
        bad = getCounters(g)["bad"]
        assert bad == 0 if params.get("arithmetic", MCB_ARITH_FAST) == MCB_ARITH_FAST else bad <= 2e-5 * n, bad
        return {k: np.array(v) for k, v in rows.items()}
    finally:
        finalize_Integrator(g)


def _mean_err(rows):
    return rows.mean(axis=0), rows.std(axis=0, ddof=1) / np.sqrt(rows.shape[0])


@pytest.mark.parametrize("views", [False, True], ids=["flux", "views"])
def test_full_size_c3_three_sigma_against_oracle(orc, views):
    dom, case = domains.landsat_cloud(ssa=0.99)
    workers = max(1, (os.cpu_count() or 2) - 1)
    n = 10000                                                          # photons per batch, as in the decks
    nb = max((240 if not views else 60) // workers, 2) * workers        # >= 2.4e6 photons (6e5 with five view rays per event)
    od = orc.OracleDomain(dom, tableSize=10001, forward=views)

    def make():
        g = orc.OracleIntegrator(od, useRussianRouletteForIntensity=1, zetaMin=0.3)
        if views:
            g.set_views(case["intensityMus"], case["intensityPhis"])
        return g
    tot, batches, st = orc.run_workers(make, workers, nb, n, solarMu=case["solarMu"], solarAzimuth=case["solarAzimuth"])
    assert tot == nb * n and batches == nb
    want = ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "absorbedProfile") + (("meanIntensity",) if views else ())
    rows = _gpu_rows(dom, case, nb, n, (10, 1, 0), views=views, want=want)
    for q in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed"):
        om, oe = orc.finalise(st[q + "Stats"], 1.0, tot, nb)
        gm, ge = _mean_err(rows[q])
        assert_within("C3 full " + q, gm, ge, om, oe, 3.0)
    om, oe = orc.finalise(st["absorbedProfileStats"], 1.0, tot, nb)
    gm, ge = _mean_err(rows["absorbedProfile"])
    z = assert_within("C3 full absorbedProfile", gm.ravel(), ge.ravel(), om, oe, 4.0)
    assert np.sqrt(np.mean(z[om > 0] ** 2)) < 1.5
    if views:
        cols = dom.numX * dom.numY
        m, e = orc.finalise(st["radianceStats"], 1.0, tot, nb)
        om = m.reshape(-1, cols).mean(axis=1)
        oe = np.sqrt((e.reshape(-1, cols) ** 2).sum(axis=1)) / cols
        gm, ge = _mean_err(rows["meanIntensity"])
        assert (om > 0).all()
        assert_within("C3 full meanIntensity", gm, ge, om, np.maximum(oe, ge), 3.0)


def _zmap(a, b):
    (ma, ea), (mb, eb) = _mean_err(a), _mean_err(b)
    sig = np.sqrt(ea ** 2 + eb ** 2)
    ok = sig > 0
    return ((ma - mb)[ok] / sig[ok]), ma, mb, ea, eb


_REF = {}
CASES = [("C3", lambda: domains.landsat_cloud(ssa=0.99), 100_000_000),
         ("C3_mie", lambda: domains.landsat_cloud(ssa=0.99, mie=True), 50_000_000),
         ("C5_small", lambda: domains.bench_domain(nxy=64, nz=64), 100_000_000)]


@pytest.mark.parametrize("kernel", [MCB_KERNEL_PARK, MCB_KERNEL_POOL], ids=["park", "pool"])
@pytest.mark.parametrize("name,make,photons", CASES, ids=[c[0] for c in CASES])
def test_throughput_kernels_match_reference_kernel_maps(name, make, photons, kernel):
    dom, case = make()
    nb = 16
    n = photons // nb
    want = ("fluxUp", "fluxDown", "absorbedProfile", "meanFluxUp", "meanFluxDown", "meanFluxAbsorbed")
    fast = _gpu_rows(dom, case, nb, n, (10, 1, 0), want=want, arithmetic=MCB_ARITH_FAST, tuneKernel=kernel)
    if name not in _REF:                                   # the reference-arithmetic side is shared by both kernels
        _REF[name] = _gpu_rows(dom, case, nb, n, (77, 3, 0), want=want, arithmetic=MCB_ARITH_REFERENCE)
    ref = _REF[name]
    for q in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed"):
        (ma, ea), (mb, eb) = _mean_err(fast[q]), _mean_err(ref[q])
        assert_within("%s %s fast vs reference kernel" % (name, q), ma, ea, mb, eb, 3.5)
    for q in ("fluxUp", "fluxDown"):
        z, ma, mb, ea, eb = _zmap(fast[q], ref[q])
        assert z.size > 0.5 * dom.numX * dom.numY
        rms, mean = float(np.sqrt(np.mean(z ** 2))), float(z.mean())
        assert rms < 1.1, (name, q, rms)
        assert abs(mean) < 5.0 / np.sqrt(z.size) + 0.02, (name, q, mean)
        assert np.abs(z).max() < 6.0, (name, q, np.abs(z).max())
    z, *_ = _zmap(fast["absorbedProfile"].reshape(nb, -1), ref["absorbedProfile"].reshape(nb, -1))
    assert np.abs(z).max() < 4.5 and np.sqrt(np.mean(z ** 2)) < 1.4, (name, "absorbedProfile", z)


def test_local_estimate_maps_match_reference_kernel_on_c3():
    """C3 + the five I3RC views (Russian-roulette local estimate): per-view radiance maps of the throughput kernel
    against the reference-arithmetic kernel, 2e7 photons a side."""
    dom, case = domains.landsat_cloud(ssa=0.99)
    nb, n = 16, 1_250_000
    want = ("intensity", "meanIntensity")
    fast = _gpu_rows(dom, case, nb, n, (10, 1, 0), views=True, want=want, arithmetic=MCB_ARITH_FAST)
    ref = _gpu_rows(dom, case, nb, n, (77, 3, 0), views=True, want=want, arithmetic=MCB_ARITH_REFERENCE)
    (ma, ea), (mb, eb) = _mean_err(fast["meanIntensity"]), _mean_err(ref["meanIntensity"])
    assert_within("C3 views meanIntensity fast vs reference kernel", ma, ea, mb, eb, 3.5)
    nDir = len(case["intensityMus"])
    a = fast["intensity"].reshape(nb, nDir, -1); b = ref["intensity"].reshape(nb, nDir, -1)
    for d in range(nDir):
        z, *_ = _zmap(a[:, d], b[:, d])
        rms, mean = float(np.sqrt(np.mean(z ** 2))), float(z.mean())
        assert rms < 1.1, (d, rms)
        assert abs(mean) < 5.0 / np.sqrt(z.size) + 0.02, (d, mean)


def test_column_compressed_storage_matches_reference_kernel_maps():
    """The pool flux kernel on column-compressed storage (what fields too large for L2 run on: C5) against the
    reference-arithmetic kernel on the small C5 scene, 1e8 photons, same criteria as above."""
    name, make, photons = CASES[2]
    dom, case = make()
    nb = 16
    n = photons // nb
    want = ("fluxUp", "fluxDown", "absorbedProfile", "meanFluxUp", "meanFluxDown", "meanFluxAbsorbed")
    fast = _gpu_rows(dom, case, nb, n, (10, 1, 0), want=want, arithmetic=MCB_ARITH_FAST, tuneKernel=MCB_KERNEL_POOL, tuneExtMask=2)
    if name not in _REF:
        _REF[name] = _gpu_rows(dom, case, nb, n, (77, 3, 0), want=want, arithmetic=MCB_ARITH_REFERENCE)
    ref = _REF[name]
    for q in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed"):
        (ma, ea), (mb, eb) = _mean_err(fast[q]), _mean_err(ref[q])
        assert_within("%s %s columns vs reference kernel" % (name, q), ma, ea, mb, eb, 3.5)
    for q in ("fluxUp", "fluxDown"):
        z, *_ = _zmap(fast[q], ref[q])
        rms, mean = float(np.sqrt(np.mean(z ** 2))), float(z.mean())
        assert rms < 1.1 and abs(mean) < 5.0 / np.sqrt(z.size) + 0.02 and np.abs(z).max() < 6.0, (q, rms, mean)
    z, *_ = _zmap(fast["absorbedProfile"].reshape(nb, -1), ref["absorbedProfile"].reshape(nb, -1))
    assert np.abs(z).max() < 4.5 and np.sqrt(np.mean(z ** 2)) < 1.4


def test_full_size_c5_default_path_matches_reference_kernel():
    """The configuration `bench.py --workload c5` measures, at full size (325 x 325 x 150, nc = 2), on the path the library
    picks by itself there -- layer-cropped field, column-compressed records and tally, clear-layer leaps -- against the
    reference-arithmetic kernel: 3.2e7 photons a side, domain means within 3.5 sigma, per-column flux maps and the
    absorption profile as unit-normal z-scores."""
    dom, case = domains.bench_domain()
    nb, n = 16, 2_000_000
    want = ("fluxUp", "fluxDown", "absorbedProfile", "meanFluxUp", "meanFluxDown", "meanFluxAbsorbed")
    fast = _gpu_rows(dom, case, nb, n, (10, 1, 0), want=want, arithmetic=MCB_ARITH_FAST)
    ref = _gpu_rows(dom, case, nb, n, (77, 3, 0), want=want, arithmetic=MCB_ARITH_REFERENCE)
    for q in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed"):
        (ma, ea), (mb, eb) = _mean_err(fast[q]), _mean_err(ref[q])
        assert_within("C5 full %s fast vs reference kernel" % q, ma, ea, mb, eb, 3.5)
    for q in ("fluxUp", "fluxDown"):
        z, *_ = _zmap(fast[q], ref[q])
        assert z.size > 0.5 * dom.numX * dom.numY
        rms, mean = float(np.sqrt(np.mean(z ** 2))), float(z.mean())
        assert rms < 1.1 and abs(mean) < 5.0 / np.sqrt(z.size) + 0.02 and np.abs(z).max() < 6.5, (q, rms, mean, np.abs(z).max())
    z, *_ = _zmap(fast["absorbedProfile"].reshape(nb, -1), ref["absorbedProfile"].reshape(nb, -1))
    assert np.abs(z).max() < 4.5 and np.sqrt(np.mean(z ** 2)) < 1.4
