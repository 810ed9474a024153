"""The photon-pool flux kernel (csrc/mcb_pool.cu) against the park/regroup kernel (csrc/mcb_fast.cu).

Both kernels give photon p the Philox stream (seed, p) and draw from it in the same order (one block per birth and
per event), start every leg from the same single-precision position and march it with the same burst code -- so with
the same seed they trace the SAME photon histories.  Only the order in which lanes pick photons up differs.  Hence:
event counters equal exactly, tallies equal up to f64 summation order (f32 where small grids privatise them in shared
memory).  (The pool kernels' vacuum leaps are switched off here -- a leap lands a ray where the cell-by-cell walk takes
it only up to rounding; tests/test_gpu_leap.py covers them.)  The park kernel's own parity with the oracle (test_gpu_stats.py) then carries over; the pool kernel is also
run through the oracle's 3-sigma test directly."""
import numpy as np
import pytest

from common import oracle_weights
from mcbrat3d_b200 import domains
from mcbrat3d_b200.emissionAndBroadBandWeights import Weights, emission_weighting
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.monteCarloRadiativeTransfer import (MCB_KERNEL_PARK, MCB_KERNEL_POOL, MCB_LAYOUT_LINEAR,
                                                       computeRadiativeTransfer, finalize_Integrator, getCounters,
                                                       new_Integrator, reportResults, specifyParameters)
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence

pytestmark = pytest.mark.gpu

CASES = [
    ("C3_small", lambda: domains.landsat_cloud(ssa=0.99, nxy=32), 400000, {}),
    ("C3_small_mie", lambda: domains.landsat_cloud(ssa=0.99, nxy=24, mie=True), 300000, {}),
    ("C3_small_linear", lambda: domains.landsat_cloud(ssa=0.99, nxy=32), 300000, dict(tuneLayout=MCB_LAYOUT_LINEAR)),
    ("C5_small_bitmap", lambda: domains.bench_domain(nxy=41, nz=47), 300000, dict(tuneExtMask=1)),
    ("C5_small_columns", lambda: domains.bench_domain(nxy=41, nz=47), 300000, dict(tuneExtMask=2)),   # column-compressed storage
    ("C5_columns_taller", lambda: domains.bench_domain(nxy=24, nz=96), 200000, dict(tuneExtMask=2)),
    ("C5_small", lambda: domains.bench_domain(nxy=40, nz=48), 300000, {}),
    ("C1", lambda: domains.homogeneous_slab(ssa=0.99), 400000, {}),                 # tallies privatised in shared memory
    ("C4_LW", lambda: domains.homogeneous_lw(), 400000, {}),                        # thermal source, emission bookkeeping
    ("reflecting", lambda: domains.homogeneous_slab(ssa=1.0, tau=2.0, albedo=0.8, n=9, delta=0.125), 300000, {}),
    ("tiny_launch", lambda: domains.landsat_cloud(ssa=0.99, nxy=16), 777, {}),      # fewer photons than one warp's pool... x12
]
# tuneBurst: 8 / 4 cells per burst with all gathers up front, 44 = eight cells with the gathers in two halves
VARIANTS = [dict(tuneBlocksPerSM=6, tuneBurst=8), dict(tuneBlocksPerSM=8, tuneBurst=8), dict(tuneBlocksPerSM=6, tuneBurst=4),
            dict(tuneBlocksPerSM=8, tuneBurst=44), dict(tuneBlocksPerSM=6, tuneBurst=44), dict(tuneBlocksPerSM=7, tuneBurst=44)]


def run(dom, case, n, **knobs):
    g = new_Integrator(dom)
    try:
        lw = case.get("LW_flag", -1.0) > 0
        specifyParameters(g, minInverseTableSize=10001, LW_flag=1.0 if lw else -1.0, **knobs)
        rs = new_RandomNumberSequence([10, 1, 0])
        if lw:
            w = Weights()
            emission_weighting(dom, w, case.get("surfaceTemp", 300.0), thisIntegrator=g)
            ps = new_PhotonStream(theseWeights=w, numberOfPhotons=n, randomNumbers=rs)
        else:
            ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], n, rs)
        assert computeRadiativeTransfer(g, dom, rs, ps, n) == n
        res = reportResults(g, fluxUp=True, fluxDown=True, fluxAbsorbed=True, volumeAbsorption=True, meanFluxUp=True,
                            meanFluxDown=True, meanFluxAbsorbed=True)
        return res, getCounters(g)
    finally:
        finalize_Integrator(g)


@pytest.mark.parametrize("name,make,n,knobs", CASES, ids=[c[0] for c in CASES])
def test_pool_traces_the_same_histories_as_the_park_kernel(name, make, n, knobs):
    dom, case = make()
    want, cw = run(dom, case, n, tuneKernel=MCB_KERNEL_PARK, **knobs)
    assert cw["photons"] == n and cw["bad"] == 0 and cw["crossings"] > n
    private = dom.numX * dom.numY <= 1024          # shared-memory f32 partial sums: summation order shows at 1e-6
    for variant in VARIANTS:
        got, cg = run(dom, case, n, tuneKernel=MCB_KERNEL_POOL, tuneLeap=-1, **knobs, **variant)   # leaps: test_gpu_leap.py
        # The crossings COUNTER is the one thing that may differ, and only between burst lengths: when a ray leaves the
        # domain in the middle of a burst the cells it entered are counted by comparing face distances with the distance to
        # the boundary, and rounding can count the first ghost cell too (a few per 1e5 crossings; no effect on the physics).
        diff = {k: (cg[k], cw[k]) for k in cg if cg[k] != cw[k] and not (k == "crossings" and variant["tuneBurst"] == 4)}
        assert not diff, (variant, diff)
        assert abs(cg["crossings"] - cw["crossings"]) <= 1e-3 * cw["crossings"]
        for k in want:
            np.testing.assert_allclose(np.asarray(got[k], np.float64), np.asarray(want[k], np.float64),
                                       rtol=3e-4 if private else 2e-5, atol=2e-5 if private else 1e-7, err_msg="%s %s" % (k, variant))


def test_pool_result_independent_of_batch_split():
    """One launch of N photons = ragged launches that add up to N (pool start-up and drain at every launch)."""
    import ctypes as C
    dom, case = domains.landsat_cloud(ssa=0.99, nxy=32)
    g = new_Integrator(dom)
    try:
        specifyParameters(g, minInverseTableSize=10001, tuneKernel=MCB_KERNEL_POOL)
        rs = new_RandomNumberSequence([10, 1, 0])
        n = 200000
        ps = new_PhotonStream(0.5, 0.0, n, rs)
        computeRadiativeTransfer(g, dom, rs, ps, n)
        want = dict(fluxUp=True, fluxDown=True, volumeAbsorption=True)
        whole = reportResults(g, **want)
        cw = getCounters(g)
        done = C.c_int64(0)
        parts = [0, 1, 64, 65, 1037, n // 3, n - 5, n]
        for i in range(len(parts) - 1):
            fn = g._lib.mcb_run_batch if i == 0 else g._lib.mcb_accumulate_batch
            assert fn(g.handle, parts[i + 1] - parts[i], C.c_uint64(rs.seed), C.c_uint64(parts[i]), C.byref(done)) == 0
        split = reportResults(g, **want)
        assert getCounters(g) == cw
        for k in whole:
            np.testing.assert_allclose(split[k], whole[k], rtol=2e-5, atol=1e-7, err_msg=k)
    finally:
        finalize_Integrator(g)


def test_pool_three_sigma_against_oracle(orc):
    """Criterion (b) for the pool kernel directly: C3-style scene with a Mie-like table and a Rayleigh component."""
    from test_gpu_stats import NB, assert_within, oracle_batches
    from mcbrat3d_b200.batchStatistics import BatchStatistics
    dom, case = domains.landsat_cloud(ssa=0.99, nxy=16, mie=True)
    n = 3000
    ores, _ = oracle_batches(orc, dom, case, NB, n, 0)
    g = new_Integrator(dom)
    try:
        specifyParameters(g, minInverseTableSize=10001, tuneKernel=MCB_KERNEL_POOL)
        rs = new_RandomNumberSequence([10, 1, 0])
        bs = BatchStatistics()
        for _ in range(NB):
            ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], n, rs)
            done = computeRadiativeTransfer(g, dom, rs, ps, n)
            bs.accumulate(reportResults(g, meanFluxUp=True, meanFluxDown=True, meanFluxAbsorbed=True, absorbedProfile=True,
                                        fluxUp=True), done)
        gm, ge = bs.finalise(1.0)
    finally:
        finalize_Integrator(g)
    for q in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed"):
        assert_within("pool " + q, gm[q], ge[q], ores[q][0], ores[q][1], 3.0)
    z = assert_within("pool absorbedProfile", gm["absorbedProfile"], ge["absorbedProfile"], ores["absorbedProfile"][0],
                      ores["absorbedProfile"][1], 4.0)
    assert np.sqrt(np.mean(z ** 2)) < 1.6


# ---- local estimation on the pool organisation (csrc/mcb_pool_le.cu) against the task-queue kernel (mcb_fast.cu) ----
LE_CASES = [
    ("C3_small_rr", lambda: domains.landsat_cloud(ssa=0.99, nxy=32), 60000, dict(useRussianRouletteForIntensity=True, zetaMin=0.3), None),
    ("C3_small_plain", lambda: domains.landsat_cloud(ssa=0.99, nxy=32), 40000, dict(useRussianRouletteForIntensity=False), None),
    ("C3_small_mie_rr", lambda: domains.landsat_cloud(ssa=0.99, nxy=24, mie=True), 40000,
     dict(useRussianRouletteForIntensity=True, zetaMin=0.3), None),
    ("C5_small_bitmap_rr", lambda: domains.bench_domain(nxy=24, nz=32), 40000,
     dict(useRussianRouletteForIntensity=True, zetaMin=0.3, tuneExtMask=1), (domains.I3RC_VIEWS_MU, domains.I3RC_VIEWS_PHI)),
    ("C3_small_hybrid_limit", lambda: domains.landsat_cloud(ssa=0.99, nxy=24, mie=True), 30000,
     dict(useRussianRouletteForIntensity=True, zetaMin=0.3, useHybridPhaseFunsForIntenCalcs=True, hybridPhaseFunWidth=7.0,
          numOrdersOrigPhaseFunIntenCalcs=2, limitIntensityContributions=True, maxIntensityContribution=0.05), None),
    ("C3_small_three_views", lambda: domains.landsat_cloud(ssa=0.99, nxy=32), 40000,
     dict(useRussianRouletteForIntensity=True, zetaMin=0.3), ([1.0, 0.5, -0.5], [0.0, 0.0, 180.0])),     # odd count, one looking down
    ("C4_LW_views", lambda: domains.homogeneous_lw(), 60000, dict(useRussianRouletteForIntensity=True, zetaMin=0.3),
     ([1.0, 0.5, -0.5], [0.0, 0.0, 90.0])),                                                               # births post requests too
    ("C2_step_cloud_rr", lambda: domains.step_cloud(ssa=0.99, solarMu=0.5), 60000,                       # 32 x 1 x 32: narrower
     dict(useRussianRouletteForIntensity=True, zetaMin=0.3), None),                                       # than the ghost shell
    ("C2_step_cloud_plain", lambda: domains.step_cloud(ssa=1.0, solarMu=1.0), 40000, dict(useRussianRouletteForIntensity=False), None),
    ("reflecting_plain", lambda: domains.homogeneous_slab(ssa=0.9, tau=2.0, albedo=0.5, n=9, delta=0.125), 40000,
     dict(useRussianRouletteForIntensity=False), ([1.0, 0.866], [0.0, 0.0])),                             # surface requests
]


def run_le(dom, case, n, views, **params):
    g = new_Integrator(dom)
    try:
        lw = case.get("LW_flag", -1.0) > 0
        mus, phis = views if views else (case["intensityMus"], case["intensityPhis"])
        specifyParameters(g, intensityMus=mus, intensityPhis=phis, computeIntensity=True, minInverseTableSize=10001,
                          minForwardTableSize=10001, LW_flag=1.0 if lw else -1.0, **params)
        rs = new_RandomNumberSequence([10, 1, 0])
        if lw:
            w = Weights()
            emission_weighting(dom, w, case.get("surfaceTemp", 300.0), thisIntegrator=g)
            ps = new_PhotonStream(theseWeights=w, numberOfPhotons=n, randomNumbers=rs)
        else:
            ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], n, rs)
        assert computeRadiativeTransfer(g, dom, rs, ps, n) == n
        res = reportResults(g, fluxUp=True, fluxDown=True, volumeAbsorption=True, intensity=True, intensityByComponent=True,
                            meanIntensity=True)
        return res, getCounters(g)
    finally:
        finalize_Integrator(g)


@pytest.mark.parametrize("name,make,n,params,views", LE_CASES, ids=[c[0] for c in LE_CASES])
def test_pool_le_traces_the_same_rays_as_the_queue_kernel(name, make, n, params, views):
    dom, case = make()
    want, cw = run_le(dom, case, n, views, tuneKernel=MCB_KERNEL_PARK, **params)
    up = np.asarray(views[0] if views else case["intensityMus"]) > 0       # with the Russian-roulette estimate a ray that ends
    assert cw["bad"] == 0 and cw["leRays"] > n                             # on a black surface contributes nothing (INT:1753-1813)
    assert (np.asarray(want["meanIntensity"])[up] > 0).all()
    private = dom.numX * dom.numY <= 1024
    for variant in (dict(tuneBlocksPerSM=5), dict(tuneBlocksPerSM=4), dict(tuneBlocksPerSM=5, tuneLayout=2)):
        got, cg = run_le(dom, case, n, views, tuneKernel=MCB_KERNEL_POOL, tuneLeap=-1, **params, **variant)
        diff = {k: (cg[k], cw[k]) for k in cg if cg[k] != cw[k]}
        assert not diff, (variant, diff)
        for k in want:
            np.testing.assert_allclose(np.asarray(got[k], np.float64), np.asarray(want[k], np.float64),
                                       rtol=3e-4 if private else 3e-5, atol=3e-5 if private else 1e-7, err_msg="%s %s" % (k, variant))
