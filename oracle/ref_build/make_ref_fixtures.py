#!/usr/bin/env python
"""Runs oracle/_ref/ref_trace_driver (the unmodified reference, one photon per batch) on every case of
make_ref_cases.py and commits what it answered as tests/golden/ref_trace_<case>.npz -- the fixtures that pin the C
oracle to the Fortran (tests/test_ref_fixtures.py).  TEST INFRASTRUCTURE; needs `make -C oracle/ref_build` first."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_ref_cases import cases, read_fingerprints, write_case  # noqa: E402

DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_trace_driver")


def main(nPhotons=1500):
    if not os.path.exists(DRIVER):
        raise SystemExit("oracle/_ref/ref_trace_driver is not built (no Fortran compiler?): see oracle/ref_build/Makefile")
    work = os.path.join(ROOT, "oracle", "_ref", "cases")
    os.makedirs(work, exist_ok=True)
    for name, ((dom, case), views) in cases().items():
        cf, of = os.path.join(work, name + ".case"), os.path.join(work, name + ".out")
        meta = write_case(cf, dom, case, views, nPhotons, 1)
        subprocess.check_call([DRIVER, cf, of])
        batch, proc, aid, idx, val = read_fingerprints(of)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_trace_%s.npz" % name), batch=batch, processed=proc,
                            array=aid, index=idx, value=val, **{k: np.asarray(v) for k, v in meta.items()})
        print("%s: %d photons, %d non-zero tally entries" % (name, nPhotons, aid.size))


if __name__ == "__main__":
    main()
