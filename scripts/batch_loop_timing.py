"""Driver-style batch loop (DRV:949-1052) on the device at the decks' batch sizes: time per loop and per report."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcbrat3d_b200 import domains
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.monteCarloRadiativeTransfer import *
from mcbrat3d_b200.batchStatistics import *
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence
dom, case = domains.landsat_cloud(ssa=0.99)
g = new_Integrator(dom); specifyParameters(g, minInverseTableSize=10001)
rs = new_RandomNumberSequence([10, 1, 0])
for n, nb in ((10000, 100), (100000, 100), (1000000, 32), (10000000, 8), (10000000, 8)):
    ps = new_PhotonStream(0.5, 0.0, n * nb, rs)
    resetDeviceStatistics(g, dom)
    t0 = time.perf_counter(); computeRadiativeTransferBatches(g, dom, rs, ps, n, nb); t1 = time.perf_counter()
    m, e, tot, done = reportStatistics(g); t2 = time.perf_counter()
    print("batches %d x %d: loop %.1f ms (%.3g photons/s), report %.1f ms, meanFluxUp %.5f +- %.5f" % (
        nb, n, (t1 - t0) * 1e3, n * nb / (t1 - t0), (t2 - t1) * 1e3, m["meanFluxUp"], e["meanFluxUp"]))
