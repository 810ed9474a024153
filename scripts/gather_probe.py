"""Measured ceiling of divergent sector gathers (csrc/mcb_probe.cu) for a few working-set sizes and launch shapes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcbrat3d_b200 import domains
from mcbrat3d_b200.monteCarloRadiativeTransfer import gatherProbe, new_Integrator

dom, _ = domains.homogeneous_slab(n=8, delta=0.125)
g = new_Integrator(dom)
out = []
import sys
full = "--full" in sys.argv
sizes = (16 << 10, 128 << 10, 2 << 20, 16 << 20, 93 << 20, 1024 << 20) if full else (16 << 20, 93 << 20)
for nbytes in sizes:
    for occ in ((6, 8, 16) if full else (8,)):
        for inflight in ((1, 8, -1, -8) if full else (8,)):
            r = gatherProbe(g, nbytes, inflight, occ, 2000 if nbytes < (1 << 30) else 600)
            out.append(dict(buffer_bytes=nbytes, blocks_per_sm=occ, loads_in_flight=inflight, gathers_per_s=r,
                            sector_gbs=r * 32 / 1e9))
            print("buffer %8.3f MB  %2d CTAs/SM  %2d loads in flight (%s): %.4g gathers/s = %.0f GB/s of 32-byte sectors" % (
                nbytes / 2.0 ** 20, occ, abs(inflight), "16 B" if inflight < 0 else "4 B", r, r * 32 / 1e9), flush=True)
print("JSON " + json.dumps(out))
