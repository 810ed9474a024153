"""Per-call cost of the device-side staging paths at C5 size (325 x 325 x 160, 3 components)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcbrat3d_b200 import domains
from mcbrat3d_b200.monteCarloRadiativeTransfer import new_Integrator, specifyParameters, _stage_domain
from mcbrat3d_b200.opticalProperties import read_SSPTable
common, tables, case = domains.broadband_problem(nxy=325, nz=160, nLambda=8)
d0 = read_SSPTable(tables, 1, common, setup=True)
g = new_Integrator(d0)
specifyParameters(g, minInverseTableSize=9001, buildTablesOnDevice=True, tuneExtMask=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for i in range(1, 7):
    t0 = time.perf_counter(); d = read_SSPTable(tables, i, common, calcRayl=True, thisIntegrator=g); t1 = time.perf_counter()
    _stage_domain(g, d); t2 = time.perf_counter()
    print("bin %d: read_SSPTable on device %.2f ms, inverse tables on device %.2f ms" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
t0 = time.perf_counter(); dh = read_SSPTable(tables, 2, common, calcRayl=True); t1 = time.perf_counter()
print("NumPy read_SSPTable (host mirror): %.1f ms" % ((t1 - t0) * 1e3))
g2 = new_Integrator(dh); specifyParameters(g2, minInverseTableSize=9001)
t0 = time.perf_counter(); _stage_domain(g2, dh); t1 = time.perf_counter()
print("mcb_set_optics upload of the dense arrays + host inverse tables: %.1f ms" % ((t1 - t0) * 1e3))
