"""Vacuum leaps of the photon-pool kernels (csrc/mcb_march.cuh march_leap, csrc/mcb_stage.cu dist_*_kernel).

The packed extinction field carries, in every cell without extinction, the Chebyshev distance D to the nearest cell
that has some; a ray whose cell lies D >= 4 cells deep in vacuum goes straight to the face where it leaves the cube of
D - 1 cells around it.  No optical depth accumulates in vacuum (OPT:1729-1738 adds 0 per cell), so the physics is the
cell-by-cell walk's: (1) the map is checked against a brute-force transform, (2) leaps on / off give the same photon
histories up to rounding at the landing faces -- event counters (including the cells-crossed counter, which a leap
advances by the faces it crosses) and tallies agree far inside the Monte Carlo noise, (3) the full-size 1e8-photon
maps of tests/test_gpu_headline.py run with leaps on (the default), (4) the leap counters, which only the
bounds-checked build keeps, show that the leaps are taken."""
import numpy as np
import pytest

from mcbrat3d_b200 import domains
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.monteCarloRadiativeTransfer import (MCB_KERNEL_POOL, computeRadiativeTransfer, finalize_Integrator,
                                                       getCounters, new_Integrator, reportResults, specifyParameters,
                                                       vacuumDistanceMap)
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence

pytestmark = pytest.mark.gpu


def brute_force_distance(ext, cap):
    """ext (nz, ny, nx).  Grow cubes around the occupied cells one shell at a time: periodic in x and y, nothing
    above the top or below the surface."""
    occ = np.asarray(ext, np.float32) != 0
    nz = occ.shape[0]
    D = np.where(occ, 0, cap).astype(np.int32)
    reach = np.concatenate([np.zeros_like(occ[:1]), occ, np.zeros_like(occ[:1])], axis=0)
    for s in range(1, cap):
        grown = reach.copy()
        for ax in (1, 2):                                   # periodic axes
            grown = grown | np.roll(grown, 1, axis=ax) | np.roll(grown, -1, axis=ax)
        up = np.zeros_like(grown); up[1:] = grown[:-1]
        dn = np.zeros_like(grown); dn[:-1] = grown[1:]
        grown = grown | up | dn
        new = grown[1:nz + 1] & ~reach[1:nz + 1]
        D[new] = s
        reach = grown
        reach[0] = reach[-1] = False
    return D


def sparse_domain(nx=19, ny=12, nz=23, seed=3):
    from mcbrat3d_b200.opticalProperties import Domain
    from mcbrat3d_b200.scatteringPhaseFunctions import henyeyGreenstein, new_PhaseFunctionTable
    rng = np.random.default_rng(seed)
    d = Domain(0.5 * np.arange(nx + 1), 0.25 * np.arange(ny + 1), 0.125 * np.arange(nz + 1), surfaceAlbedo=0.3)
    ext = np.where(rng.random((nz, ny, nx)) < 0.004, rng.uniform(2.0, 40.0, (nz, ny, nx)), 0.0)
    ext[14:] = 0.0                                         # vacuum above the highest cell
    ext[3:6, 2:5, 4:9] = 12.0                              # one solid block, the rest isolated cells
    ssa = np.where(ext > 0, 0.95, 0.0)
    idx = np.where(ext > 0, 1, 0).astype(np.int32)
    d.addOpticalComponent("cloud", ext, ssa, idx, new_PhaseFunctionTable([henyeyGreenstein(0.8, 32)], key=[1.0]))
    d.getOpticalPropertiesByComponent()
    return d, dict(name="sparse", solarMu=0.35, solarAzimuth=25.0)


MAPS = [("C3_small", lambda: domains.landsat_cloud(ssa=0.99, nxy=32)),
        ("sparse", sparse_domain),
        ("slab", lambda: domains.homogeneous_slab(ssa=0.99, n=9, delta=0.125))]


@pytest.mark.parametrize("name,make", MAPS, ids=[m[0] for m in MAPS])
def test_vacuum_distance_map_matches_brute_force(name, make):
    dom, case = make()
    g = new_Integrator(dom)
    try:
        got = vacuumDistanceMap(g, dom).astype(np.int32)
    finally:
        finalize_Integrator(g)
    cap = min(64, dom.numX, dom.numY)
    want = brute_force_distance(dom.totalExt.reshape(dom.numZ, dom.numY, dom.numX), cap)
    assert np.array_equal(got, want), (np.argwhere(got != want)[:5], got[got != want][:5], want[got != want][:5])
    if name != "slab":
        assert want.max() >= 4                              # there is something to leap through


def _run(dom, case, n, views=None, **knobs):
    g = new_Integrator(dom)
    try:
        if views:
            specifyParameters(g, intensityMus=views[0], intensityPhis=views[1], computeIntensity=True,
                              useRussianRouletteForIntensity=True, zetaMin=0.3)
        specifyParameters(g, minInverseTableSize=10001, minForwardTableSize=10001, tuneKernel=MCB_KERNEL_POOL, **knobs)
        rs = new_RandomNumberSequence([10, 1, 0])
        ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], n, rs)
        assert computeRadiativeTransfer(g, dom, rs, ps, n) == n
        want = dict(fluxUp=True, fluxDown=True, volumeAbsorption=True, meanFluxUp=True, meanFluxDown=True, meanFluxAbsorbed=True)
        if views:
            want.update(intensity=True, meanIntensity=True)
        return reportResults(g, **want), getCounters(g)
    finally:
        finalize_Integrator(g)


LEAP_CASES = [("C3_small", lambda: domains.landsat_cloud(ssa=0.99, nxy=32), 400000, None),
              ("C3_small_linear", lambda: domains.landsat_cloud(ssa=0.99, nxy=32), 300000, None),
              ("sparse", sparse_domain, 400000, None),
              ("C5_small_bitmap", lambda: domains.bench_domain(nxy=40, nz=96), 300000, None),   # whole clear layers, not vacuum
              ("C5_small_columns", lambda: domains.bench_domain(nxy=40, nz=96), 300000, None),  # the same on column-compressed storage
              ("C5_small_bitmap_views", lambda: domains.bench_domain(nxy=24, nz=96), 40000, ([1.0, 0.5, -0.5], [0.0, 0.0, 180.0])),
              ("C3_small_views", lambda: domains.landsat_cloud(ssa=0.99, nxy=32), 60000, (domains.I3RC_VIEWS_MU, domains.I3RC_VIEWS_PHI)),
              ("sparse_views", sparse_domain, 60000, ([1.0, 0.5, -0.6], [0.0, 70.0, 200.0]))]


@pytest.mark.parametrize("name,make,n,views", LEAP_CASES, ids=[c[0] for c in LEAP_CASES])
def test_leaps_trace_the_same_histories_up_to_rounding(name, make, n, views):
    dom, case = make()
    extra = dict(tuneLayout=1) if name.endswith("linear") else dict(tuneExtMask=1) if "bitmap" in name else dict(tuneExtMask=2) if "columns" in name else {}
    want, cw = _run(dom, case, n, views, tuneLeap=-1, **extra)
    assert cw["bad"] == 0
    for leap in (0, 2, 7):                                  # default distance, the smallest, a larger one
        got, cg = _run(dom, case, n, views, tuneLeap=leap, tuneLeapLanes=8 if leap == 7 else 0, **extra)   # one run with the warp gate
        assert cg["bad"] == 0 and cg["photons"] == cw["photons"]
        # Same random numbers, same directions; positions differ by an ulp of the coordinate once a leap has replaced
        # repeated additions by one multiplication, and a photon that passes within that of a cell edge takes another
        # cell there: a few photons per thousand end in another column, everything else is the same history.  The
        # counters move by a few 1e-5, the per-column maps far less than their noise (5 % per column at this size).
        # (crossings: where a burst runs through the top or the surface, rounding of the boundary distance can count the
        # first ghost cell as entered -- tests/test_gpu_pool.py; a leap lands on the boundary exactly -- hence up to one
        # cell per exit between the two counts)
        for k in ("crossings", "scatters", "surfaceHits", "leRays", "leCrossings"):
            slack = (cw["photons"] + cw["surfaceHits"] if k == "crossings" else cw["leRays"] if k == "leCrossings" else 0)
            assert abs(cg[k] - cw[k]) <= 2e-4 * max(cw[k], 1) + 3 + slack, (leap, k, cg[k], cw[k])
        for k in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed") + (("meanIntensity",) if views else ()):
            np.testing.assert_allclose(got[k], want[k], rtol=6e-3 if k == "meanIntensity" else 1e-3, atol=1e-6,
                                       err_msg="%s leap=%d" % (k, leap))
        for k in ("fluxUp", "fluxDown"):
            a, b = np.asarray(got[k], np.float64), np.asarray(want[k], np.float64)
            assert np.abs(a - b).sum() <= 2e-2 * np.abs(b).sum(), (leap, k)


COUNTER_SCRIPT = r"""
import json, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
from mcbrat3d_b200 import _lib
assert _lib.LIB_PATH.endswith("libmcbrat_cuda_dbg.so")
import test_gpu_leap as T
out = {}
for name, make, n, views in T.LEAP_CASES:
    dom, case = make()
    extra = dict(tuneLayout=1) if name.endswith("linear") else dict(tuneExtMask=1) if "bitmap" in name else dict(tuneExtMask=2) if "columns" in name else {}
    _, c = T._run(dom, case, min(n, 100000), views, **extra)
    out[name] = c
dom, case = T.domains.landsat_cloud(ssa=0.99, nxy=24, mie=True)      # Rayleigh background: no vacuum anywhere
_, out["no_vacuum"] = T._run(dom, case, 50000)
print("RESULT " + json.dumps(out))
"""


def test_leap_counters():
    """The counters leaps / leapCells exist in the bounds-checked build only (the warp reduction behind them costs the
    72-register flux kernel 9 %): every case above takes leaps, each crosses at least two cells, the cells are part of
    the crossings counters, nothing is read out of bounds -- and a scene without vacuum runs the kernels without the
    leap code (no leaps, same counters as ever)."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-c", COUNTER_SCRIPT % (root, os.path.join(root, "tests"))], capture_output=True, text=True,
                       env=dict(os.environ, MCB_LIB_DEBUG="1"), timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith("RESULT ")][0][7:])
    for name, c in res.items():
        assert c["bad"] == 0, (name, c)
        if name == "no_vacuum":
            assert c["leaps"] == 0 and c["leapCells"] == 0, c
            continue
        assert c["leaps"] > 0 and 2 * c["leaps"] <= c["leapCells"] < c["crossings"] + c["leCrossings"], (name, c)
        print("%s: %.2f leaps per photon, %.1f cells per leap, %.3f of the crossings" % (
            name, c["leaps"] / c["photons"], c["leapCells"] / c["leaps"], c["leapCells"] / (c["crossings"] + c["leCrossings"])))


def test_an_empty_domain_is_crossed_in_leaps():
    """Nothing but vacuum: every photon goes straight to the surface (flux down exactly 1), 30 % come back up and leave
    through the top -- leaps on or off, and the leaps land on the boundaries exactly (same crossings up to the one cell
    per exit that a burst may count beyond the boundary)."""
    from mcbrat3d_b200.opticalProperties import Domain
    from mcbrat3d_b200.scatteringPhaseFunctions import henyeyGreenstein, new_PhaseFunctionTable
    nx, ny, nz = 24, 16, 40
    d = Domain(0.25 * np.arange(nx + 1), 0.25 * np.arange(ny + 1), 0.125 * np.arange(nz + 1), surfaceAlbedo=0.3)
    z = np.zeros((nz, ny, nx))
    d.addOpticalComponent("nothing", z, z, np.zeros((nz, ny, nx), np.int32), new_PhaseFunctionTable([henyeyGreenstein(0.8, 16)], key=[1.0]))
    d.getOpticalPropertiesByComponent()
    case = dict(solarMu=0.4, solarAzimuth=33.0)
    n = 200000
    res = {}
    for leap in (-1, 0):
        r, c = _run(d, case, n, tuneLeap=leap)
        assert c["bad"] == 0 and c["scatters"] == 0 and c["surfaceHits"] == n, c
        assert abs(float(r["meanFluxDown"]) - 1.0) < 1e-6 and abs(float(r["meanFluxUp"]) - 0.3) < 1e-6, r
        assert abs(float(r["meanFluxAbsorbed"])) < 1e-9
        res[leap] = c
    assert abs(res[0]["crossings"] - res[-1]["crossings"]) <= 2 * n + 2e-4 * res[-1]["crossings"], res


def test_thermal_source_below_vacuum_leaps():
    """Emission bookkeeping with leaps: an emitting, scattering layer under an empty upper half of the domain, thermal
    source (births inside the layer and at the surface, INT:504-508 decrements where a photon is born).  Leaps on / off
    and the park/regroup kernel agree to the rounding-level differences of the other cases."""
    from mcbrat3d_b200.emissionAndBroadBandWeights import Weights, emission_weighting
    from mcbrat3d_b200.monteCarloRadiativeTransfer import MCB_KERNEL_PARK
    from mcbrat3d_b200.opticalProperties import Domain
    from mcbrat3d_b200.scatteringPhaseFunctions import henyeyGreenstein, new_PhaseFunctionTable
    n3 = 24
    edges = 0.0625 * np.arange(n3 + 1, dtype=np.float64)
    d = Domain(edges, edges, edges, temps=np.full((n3, n3, n3), 285.0), surfaceAlbedo=0.1, lambda_um=10.0)
    rng = np.random.default_rng(12)
    ext = np.zeros((n3, n3, n3)); ext[:10] = rng.uniform(2.0, 12.0, (10, n3, n3)); ext[4:7, 5:9, 3:8] = 0.0    # a hole as well
    ssa = np.where(ext > 0, 0.6, 0.0); idx = np.where(ext > 0, 1, 0).astype(np.int32)
    d.addOpticalComponent("cloud", ext, ssa, idx, new_PhaseFunctionTable([henyeyGreenstein(0.8, 32)], key=[1.0]))
    d.getOpticalPropertiesByComponent()
    n = 400000
    out = {}
    for tag, knobs in (("park", dict(tuneKernel=MCB_KERNEL_PARK)), ("pool", dict(tuneKernel=MCB_KERNEL_POOL, tuneLeap=-1)),
                       ("leap", dict(tuneKernel=MCB_KERNEL_POOL))):
        g = new_Integrator(d)
        try:
            specifyParameters(g, minInverseTableSize=10001, LW_flag=1.0, **knobs)
            rs = new_RandomNumberSequence([10, 1, 0])
            w = Weights()
            emission_weighting(d, w, 300.0, thisIntegrator=g)
            ps = new_PhotonStream(theseWeights=w, numberOfPhotons=n, randomNumbers=rs)
            assert computeRadiativeTransfer(g, d, rs, ps, n) == n
            out[tag] = (reportResults(g, meanFluxUp=True, meanFluxDown=True, meanFluxAbsorbed=True, absorbedProfile=True),
                        getCounters(g))
        finally:
            finalize_Integrator(g)
    (rp, cp), (ro, co), (rl, cl) = out["park"], out["pool"], out["leap"]
    assert cp["bad"] == co["bad"] == cl["bad"] == 0 and co["scatters"] == cp["scatters"]
    for k in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed"):
        assert abs(float(ro[k]) - float(rp[k])) < 2e-6, k                       # same histories: summation order only
        assert abs(float(rl[k]) - float(rp[k])) < 1e-3 * max(abs(float(rp[k])), 0.1), (k, float(rl[k]), float(rp[k]))
    assert abs(cl["scatters"] - cp["scatters"]) <= 3e-4 * cp["scatters"] + 3
    np.testing.assert_allclose(rl["absorbedProfile"], rp["absorbedProfile"], rtol=5e-3, atol=2e-4)
    assert float(rp["meanFluxUp"]) > 0 and abs(float(rp["absorbedProfile"][12:].sum())) < 1e-12   # nothing is absorbed in vacuum
