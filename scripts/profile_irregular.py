"""The C3 scene on a grid whose spacings are not exactly representable in single precision (30.003 x 30.003 x 20.002):
new_Integrator's regular-grid test (INT:163-181, quirk q1) fails, so the irregular-grid variant of the throughput kernel
runs (edges in shared memory, x-fastest field)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcbrat3d_b200 import domains
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.monteCarloRadiativeTransfer import *
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence

dom, case = domains.landsat_cloud(ssa=0.99)
dom.xPosition = dom.xPosition * 1.0001; dom.yPosition = dom.yPosition * 1.0001
dom.zPosition = 200.0 + (dom.zPosition - 200.0) * 1.0001
if len(sys.argv) > 1 and sys.argv[1] == "stretched":        # genuinely stretched layers: the edge-table kernel variant
    dz = 20.0 * (0.6 + 0.8 * np.arange(dom.numZ) / (dom.numZ - 1))
    dom.zPosition = np.concatenate([[200.0], 200.0 + np.cumsum(dz)])
g = new_Integrator(dom)
specifyParameters(g, minInverseTableSize=10001)

rs = new_RandomNumberSequence([10, 1, 0])
n = 8000000
for b in range(3):
    ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], n, rs)
    computeRadiativeTransfer(g, dom, rs, ps, n)
    ms = lastBatchMilliseconds(g); c = getCounters(g)
    print("batch %d: %.3f ms  %.4g photons/s  %.4g crossings/s  crossings/photon %.1f bad %d" % (
        b, ms, n / ms * 1e3, c["crossings"] / ms * 1e3, c["crossings"] / n, c["bad"]))
print({k: float(v) for k, v in reportResults(g, meanFluxUp=True, meanFluxDown=True, meanFluxAbsorbed=True).items()})
