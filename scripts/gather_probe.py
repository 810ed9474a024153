"""Measured ceiling of divergent sector gathers (csrc/mcb_probe.cu) for a few working-set sizes and launch shapes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcbrat3d_b200 import domains
from mcbrat3d_b200.monteCarloRadiativeTransfer import gatherProbe, new_Integrator

dom, _ = domains.homogeneous_slab(n=8, delta=0.125)
g = new_Integrator(dom)
out = []
for mb in (16, 93, 1024):
    for occ in (6, 8, 16):
        for inflight in (1, 4, 8, 16):
            r = gatherProbe(g, mb << 20, inflight, occ, 2000 if mb < 1024 else 600)
            out.append(dict(buffer_mb=mb, blocks_per_sm=occ, loads_in_flight=inflight, gathers_per_s=r,
                            sector_gbs=r * 32 / 1e9))
            print("buffer %5d MB  %2d CTAs/SM  %2d loads in flight: %.4g gathers/s = %.0f GB/s of 32-byte sectors" % (
                mb, occ, inflight, r, r * 32 / 1e9), flush=True)
print("JSON " + json.dumps(out))
