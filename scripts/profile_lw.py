"""Thermal-source (LW) timing on the C5-size bench domain at 10 um: emission CDF built on the device, flux-only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcbrat3d_b200 import domains
from mcbrat3d_b200.emissionAndBroadBandWeights import Weights, emission_weighting
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.monteCarloRadiativeTransfer import *
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence

small = len(sys.argv) > 1 and sys.argv[1] == "small"
dom, case = domains.bench_domain(nxy=64, nz=48, ssa=0.5) if small else domains.bench_domain(ssa=0.5)
dom.lambda_um = 10.0
g = new_Integrator(dom)
specifyParameters(g, minInverseTableSize=10001, LW_flag=1.0)
w = Weights()
emission_weighting(dom, w, 295.0, thisIntegrator=g)
rs = new_RandomNumberSequence([10, 1, 0])
n = 8000000
for b in range(3):
    ps = new_PhotonStream(theseWeights=w, numberOfPhotons=n, randomNumbers=rs)
    computeRadiativeTransfer(g, dom, rs, ps, n)
    ms = lastBatchMilliseconds(g); c = getCounters(g)
    print("batch %d: %.3f ms  %.4g photons/s  crossings/photon %.1f scatters/photon %.2f bad %d fracAtms %.3f" % (
        b, ms, n / ms * 1e3, c["crossings"] / n, c["scatters"] / n, c["bad"], w.fracAtmsPower))
print({k: float(v) for k, v in reportResults(g, meanFluxUp=True, meanFluxDown=True, meanFluxAbsorbed=True).items()})
