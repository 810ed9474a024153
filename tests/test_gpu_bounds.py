"""Every gather / tally index of the throughput kernel stays inside its array (there is no compute-sanitizer on the
pool): the cases run through the bounds-checked build (csrc/Makefile target dbg, -DMCB_BOUNDS_CHECK), which counts
violations in the `bad` counter instead of making the access."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, sys
sys.path.insert(0, %r)
from mcbrat3d_b200 import _lib, domains
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.monteCarloRadiativeTransfer import *
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence
assert _lib.LIB_PATH.endswith("libmcbrat_cuda_dbg.so")
sys.path.insert(0, %r)
from test_gpu_stats import _tiny_domain
cases = {"C3_small_mie": (domains.landsat_cloud(ssa=0.99, nxy=24, mie=True), False, 150000),
         "C2_views": (domains.step_cloud(ssa=0.99, solarMu=0.5), True, 60000),
         "T_irr_views": (domains.irregular_test_domain(), True, 60000),
         "T_irr_stretched_views": (domains.irregular_test_domain(stretched=True), True, 60000),
         "C5_small": (domains.bench_domain(nxy=40, nz=48), False, 100000),
         "C5_small_odd": (domains.bench_domain(nxy=41, nz=47), False, 100000),
         "C5_small_odd_bitmap": (domains.bench_domain(nxy=41, nz=47), False, 100000),
         "single_column": ((_tiny_domain(1, 1, 6), dict(solarMu=0.5, solarAzimuth=0.0)), False, 60000),
         "narrow": ((_tiny_domain(2, 9, 3), dict(solarMu=0.3, solarAzimuth=315.0)), False, 60000)}
cases["C5_small_bitmap"] = cases["C5_small"]
cases["C5_small_bitmap_views"] = (domains.bench_domain(nxy=24, nz=32), True, 30000)
# thermal source + views: births post local-estimate requests inside the event phase (ADVICE r1: high)
cases["C4_LW_views"] = ((domains.homogeneous_lw()[0], dict(lw=True, surfaceTemp=300.0, intensityMus=[1.0, 0.5, -0.5],
                                                          intensityPhis=[0.0, 0.0, 90.0])), True, 60000)
cases["T_irr_LW_views"] = ((domains.irregular_test_domain()[0], dict(lw=True, surfaceTemp=290.0, intensityMus=[1.0, 0.7, -0.5],
                                                                    intensityPhis=[0.0, 45.0, 200.0])), True, 60000)
for name in ("C3_small_mie", "C5_small", "C5_small_odd", "C5_small_odd_bitmap"):      # the photon-pool kernel as well
    cases[name + "_pool"] = cases[name]
cases["C5_small_odd_columns_pool"] = cases["C5_small_odd"]                              # ... on column-compressed storage
cases["C5_taller_columns_pool"] = (domains.bench_domain(nxy=24, nz=96), False, 100000)
cases["C3_small_views_pool"] = (domains.landsat_cloud(ssa=0.99, nxy=24, mie=True), True, 30000)        # ... and its view rays
cases["C5_small_bitmap_views_pool"] = cases["C5_small_bitmap_views"]
cases["C4_LW_views_pool"] = cases["C4_LW_views"]
cases["C2_views_pool"] = cases["C2_views"]                                                # narrow grid on the pool LE kernel
cases["narrow_views_pool"] = ((_tiny_domain(2, 9, 3), dict(solarMu=0.3, solarAzimuth=315.0)), True, 40000)
out = {}
from mcbrat3d_b200.emissionAndBroadBandWeights import Weights, emission_weighting
for name, ((dom, case), views, n) in cases.items():
    g = new_Integrator(dom)
    if views:
        specifyParameters(g, intensityMus=case.get("intensityMus", [1.0, 0.5]), intensityPhis=case.get("intensityPhis", [0.0, 0.0]),
                          computeIntensity=True, useRussianRouletteForIntensity=True, zetaMin=0.3)
    specifyParameters(g, minInverseTableSize=9001, minForwardTableSize=9001, LW_flag=1.0 if case.get("lw") else -1.0,
                      tuneExtMask=1 if "bitmap" in name else 2 if "columns" in name else 0,   # bitmap / column-compressed variants
                      tuneKernel=MCB_KERNEL_POOL if name.endswith("_pool") else MCB_KERNEL_PARK)
    rs = new_RandomNumberSequence([3, 1, 0])
    if case.get("lw"):
        w = Weights()
        emission_weighting(dom, w, case["surfaceTemp"], thisIntegrator=g)
        ps = new_PhotonStream(theseWeights=w, numberOfPhotons=n, randomNumbers=rs)
    else:
        ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], n, rs)
    computeRadiativeTransfer(g, dom, rs, ps, n)
    c = getCounters(g)
    out[name] = dict(bad=c["bad"], photons=c["photons"], crossings=c["crossings"], leCrossings=c["leCrossings"])
    finalize_Integrator(g)
print("RESULT " + json.dumps(out))
"""


@pytest.mark.gpu
def test_no_out_of_bounds_access_in_the_fast_kernel():
    env = dict(os.environ, MCB_LIB_DEBUG="1")
    p = subprocess.run([sys.executable, "-c", SCRIPT % (ROOT, os.path.join(ROOT, "tests"))], capture_output=True, text=True,
                       env=env, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith("RESULT ")][0][7:])
    for name, c in res.items():
        assert c["bad"] == 0, (name, c)
        assert c["crossings"] > c["photons"]
    assert res["C2_views"]["leCrossings"] > 0 and res["T_irr_views"]["leCrossings"] > 0
    assert res["C4_LW_views"]["leCrossings"] > 0 and res["T_irr_LW_views"]["leCrossings"] > 0
