#!/usr/bin/env python
"""Case files for oracle/_ref/ref_trace_driver (layout: ref_trace_driver.f90 header) from the repo's synthetic domains.

TEST INFRASTRUCTURE (oracle/).  The same `Domain` objects the GPU tests use are flattened into the reference's own
constructor arguments: grid edges, per-component extinction / albedo / phase-function index in Fortran order, and the
phase functions as Legendre coefficients or angle / value pairs."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def cases():
    from mcbrat3d_b200 import domains
    return {
        "C1": (domains.homogeneous_slab(ssa=0.99), False),
        "C2_views": (domains.step_cloud(ssa=0.99, solarMu=0.5), True),
        "T_irr": (domains.irregular_test_domain(), False),
        "T_irr_views": (domains.irregular_test_domain(), True),
        "C3_small": (domains.landsat_cloud(ssa=0.99, nxy=16), False),
        "C3_small_mie_views": (domains.landsat_cloud(ssa=0.99, nxy=16, mie=True), True),
        "C5_small": (domains.bench_domain(nxy=24, nz=32), False),
    }


def write_case(path, dom, case, views, numBatches, photonsPerBatch, iseed=10, rank=1, mode=0, nS=10001, useRR=1, zetaMin=0.3):
    """Returns the dict of everything written (the fixture keeps it next to the reference's answers)."""
    nx, ny, nz = dom.numX, dom.numY, dom.numZ
    mus = np.asarray(case.get("intensityMus", [1.0, 0.5]) if views else [], np.float32)
    phis = np.asarray(case.get("intensityPhis", [0.0, 0.0]) if views else [], np.float32)
    with open(path, "wb") as f:
        f.write(struct.pack("<10i", nx, ny, nz, len(dom.components), mus.size, iseed, rank, mode, nS, useRR))
        f.write(struct.pack("<2q", numBatches, photonsPerBatch))
        f.write(struct.pack("<d", dom.surfaceAlbedo))
        f.write(struct.pack("<3f", case["solarMu"], case["solarAzimuth"], zetaMin))
        for e in (dom.xPosition, dom.yPosition, dom.zPosition):
            f.write(np.ascontiguousarray(e, "<f8").tobytes())
        f.write(mus.astype("<f4").tobytes()); f.write(phis.astype("<f4").tobytes())
        for c in dom.components:
            nzc = c.extinction.shape[0]
            f.write(struct.pack("<4i", c.zLevelBase, 1 if c.horizontallyUniform else 0, nzc, c.table.nEntries))
            for pf in c.table.phaseFunctions:
                if pf.storedAsLegendre():
                    co = np.ascontiguousarray(pf.legendreCoefficients, "<f4")
                    f.write(struct.pack("<i", co.size)); f.write(co.tobytes())
                else:
                    a = np.ascontiguousarray(pf.scatteringAngle, "<f4"); v = np.ascontiguousarray(pf.value, "<f4")
                    f.write(struct.pack("<i", -a.size)); f.write(a.tobytes()); f.write(v.tobytes())
            # host arrays are (nzc, ny, nx) in C order = (nx, ny, nzc) in Fortran order: written as they lie
            f.write(np.ascontiguousarray(c.extinction, "<f8").tobytes())
            f.write(np.ascontiguousarray(c.singleScatteringAlbedo, "<f8").tobytes())
            f.write(np.ascontiguousarray(c.phaseFunctionIndex, "<i4").tobytes())
    return dict(nx=nx, ny=ny, nz=nz, nDir=mus.size, iseed=iseed, rank=rank, nS=nS, useRR=useRR, zetaMin=zetaMin,
                numBatches=numBatches, photonsPerBatch=photonsPerBatch, solarMu=case["solarMu"], solarAzimuth=case["solarAzimuth"],
                intensityMus=mus, intensityPhis=phis)


def read_fingerprints(path):
    """The driver's mode-0 output -> (batch, processed, array id, index, value) arrays."""
    raw = open(path, "rb").read()
    off = 0
    batch, proc, aid, idx, val = [], [], [], [], []
    while off < len(raw):
        b, p, n = struct.unpack_from("<3i", raw, off); off += 12
        rec = np.frombuffer(raw, dtype=np.dtype([("a", "<i4"), ("i", "<i4"), ("v", "<f4")]), count=n, offset=off); off += 12 * n
        batch += [b] * n; proc += [p] * n
        aid.append(rec["a"]); idx.append(rec["i"]); val.append(rec["v"])
        if n == 0:                                      # a photon that left no trace still counts as a batch
            batch.append(b); proc.append(p); aid.append(np.zeros(1, "<i4")); idx.append(np.zeros(1, "<i4")); val.append(np.zeros(1, "<f4"))
    return (np.array(batch, np.int32), np.array(proc, np.int32), np.concatenate(aid), np.concatenate(idx), np.concatenate(val))


if __name__ == "__main__":
    out = os.path.join(ROOT, "oracle", "_ref", "cases")
    os.makedirs(out, exist_ok=True)
    for name, ((dom, case), views) in cases().items():
        write_case(os.path.join(out, name + ".case"), dom, case, views, 1500, 1)
        print("wrote", name)
