"""Writes the committed fixtures under tests/golden/ from the CPU oracle.

The reference cannot be built or run in this image (no Fortran compiler, MPI or netCDF) and
ships no golden vectors, so these fixtures are ORACLE outputs: they pin the oracle and the CUDA
kernels against drift, they do not pin the oracle to the reference ("parity unpinned").
Run from the repository root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import injected_randoms          # noqa: E402
from mcbrat3d_b200 import domains             # noqa: E402
from mcbrat3d_b200.monteCarloRadiativeTransfer import makeDirectionCosines  # noqa: E402
from oracle import oracle as orc              # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    dom, case = domains.irregular_test_domain()
    od = orc.OracleDomain(dom, tableSize=9001, forward=True)
    og = orc.OracleIntegrator(od)
    f32 = np.float32
    Pi = f32(3.14159265358979312)
    dirs = np.stack([makeDirectionCosines(m, f32(p) * Pi / f32(180.0))
                     for m, p in zip(case["intensityMus"], case["intensityPhis"])])
    og.set_view_cosines(dirs)
    rn = injected_randoms(48, 160, seed=2024)
    ev = og.trace(rn, 0, case["solarMu"], case["solarAzimuth"], maxEvents=48 * 1024)
    np.savez_compressed(os.path.join(HERE, "trace_T_irr.npz"), rn=rn, events=ev, tallies=og.raw_tallies())
    print("trace_T_irr.npz:", len(ev), "events")

    # small-sample flux table of the homogeneous slab (the planeParallel.f95:242 protocol)
    rows = []
    for ssa in (1.0, 0.99):
        d, c = domains.homogeneous_slab(ssa=ssa)
        g = orc.OracleIntegrator(orc.OracleDomain(d, tableSize=9001))
        tot, st = g.run_batches(10, 2000, solarMu=0.5, solarAzimuth=0.0, iseed=10, rank=1, thread=0)
        up, eu = orc.finalise(st["meanFluxUpStats"], 1.0, tot, 10)
        dn, ed = orc.finalise(st["meanFluxDownStats"], 1.0, tot, 10)
        ab, ea = orc.finalise(st["meanFluxAbsorbedStats"], 1.0, tot, 10)
        rows.append([ssa, up[0], eu[0], dn[0], ed[0], ab[0], ea[0]])
    np.save(os.path.join(HERE, "slab_fluxes_mt19937.npy"), np.array(rows))
    print(np.array(rows))


if __name__ == "__main__":
    main()
