"""Short single-GPU run for ncu: a few batches of the bench workload (C3 Landsat) at reduced photon count."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcbrat3d_b200 import domains
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.monteCarloRadiativeTransfer import *
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence

ap = argparse.ArgumentParser()
ap.add_argument("--case", default="c3")
ap.add_argument("--photons", type=int, default=4000000)
ap.add_argument("--batches", type=int, default=3)
ap.add_argument("--views", action="store_true")
ap.add_argument("--arith", type=int, default=0)
for knob in ("kernel", "layout", "blocks-per-sm", "park-threshold", "le-carry", "ext-mask", "burst", "leap", "leap-lanes"):   # mcb_options.tune*
    ap.add_argument("--" + knob, type=int, default=0)
ap.add_argument("--tag", default="")
a = ap.parse_args()
dom, case = {"c3": lambda: domains.landsat_cloud(ssa=0.99), "c1": lambda: domains.homogeneous_slab(ssa=0.99),
             "c2": lambda: domains.step_cloud(ssa=0.99, solarMu=0.5), "c3mie": lambda: domains.landsat_cloud(ssa=0.99, mie=True),
             "c4": lambda: domains.homogeneous_lw(),
             "c5": lambda: domains.bench_domain()}[a.case]()
g = new_Integrator(dom)
if a.views:
    specifyParameters(g, intensityMus=case["intensityMus"], intensityPhis=case["intensityPhis"], computeIntensity=True,
                      useRussianRouletteForIntensity=True, zetaMin=0.3)
specifyParameters(g, minInverseTableSize=10001, minForwardTableSize=10001, arithmetic=a.arith,
                  LW_flag=case.get("LW_flag", -1.0), tuneKernel=a.kernel, tuneLayout=a.layout, tuneBlocksPerSM=a.blocks_per_sm,
                  tuneParkThreshold=a.park_threshold, tuneLeCarry=a.le_carry, tuneExtMask=a.ext_mask, tuneBurst=a.burst, tuneLeap=a.leap, tuneLeapLanes=a.leap_lanes)
rs = new_RandomNumberSequence([10, 1, 0])
weights = None
if case.get("LW_flag", -1.0) > 0:
    from mcbrat3d_b200.emissionAndBroadBandWeights import Weights, emission_weighting
    weights = Weights()
    emission_weighting(dom, weights, case.get("surfaceTemp", 300.0), thisIntegrator=g)
best = 0.0
for b in range(a.batches):
    ps = (new_PhotonStream(case["solarMu"], case["solarAzimuth"], a.photons, rs) if weights is None else
          new_PhotonStream(theseWeights=weights, numberOfPhotons=a.photons, randomNumbers=rs))
    computeRadiativeTransfer(g, dom, rs, ps, a.photons)
    ms = lastBatchMilliseconds(g)
    c = getCounters(g)
    print("batch %d: %.3f ms  %.4g photons/s  %.4g crossings/s  crossings/photon %.1f scatters/photon %.2f bad %d  leaps/photon %.2f cells/leap %.1f" % (
        b, ms, a.photons / ms * 1e3, c["crossings"] / ms * 1e3, c["crossings"] / a.photons, c["scatters"] / a.photons, c["bad"],
        c["leaps"] / a.photons, c["leapCells"] / max(1, c["leaps"])))
    best = max(best, a.photons / ms * 1e3)
print("BEST %s case=%s views=%d kernel=%d layout=%d occ=%d burst=%d park=%d mask=%d leap=%d/%d photons=%d: %.4g photons/s  leaps/photon %.2f cells/leap %.1f" % (
    a.tag, a.case, a.views, a.kernel, a.layout, a.blocks_per_sm, a.burst, a.park_threshold, a.ext_mask, a.leap, a.leap_lanes, a.photons, best,
    c["leaps"] / a.photons, c["leapCells"] / max(1, c["leaps"])))
r = reportResults(g, meanFluxUp=True, meanFluxDown=True, meanFluxAbsorbed=True)
print({k: float(v) for k, v in r.items()})
