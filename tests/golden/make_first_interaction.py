"""Writes tests/golden/first_interaction_<kind>.npz: the deterministic first-interaction answers of
tests/independent_3d.py at a resolution whose quadrature error is far below any Monte Carlo noise the tests reach
(printed below as the difference to the next coarser rule).  These are NOT reference outputs and NOT oracle outputs:
they come from an independent solver of the transfer equation's first-order terms (no random numbers).

    python tests/golden/make_first_interaction.py [kind ...]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import first_interaction as fi  # noqa: E402

for kind in (sys.argv[1:] or fi.KINDS):
    t0 = time.time()
    _, med = fi.scene(kind, albedo=1.0)
    first, surf = med.first_collision(fi.SOLAR_MU, fi.SOLAR_AZIMUTH, m=96)
    f2, s2 = med.first_collision(fi.SOLAR_MU, fi.SOLAR_AZIMUTH, m=48)
    print(kind, "first collision: closure %.3e, m=48 vs m=96: %.1e (cells, rel. to the largest) %.1e (surface)" % (
        first.sum() + surf.sum() - 1.0, np.abs(first - f2).max() / first.max(), np.abs(surf - s2).max() / surf.max()), flush=True)
    E1, E0 = [], []
    for mu, phi in zip(fi.VIEW_MUS, fi.VIEW_PHIS):
        a1, a0 = med.first_order_radiance(fi.SOLAR_MU, fi.SOLAR_AZIMUTH, mu, phi, m=48, gauss=4, sub=2)
        b1, b0 = med.first_order_radiance(fi.SOLAR_MU, fi.SOLAR_AZIMUTH, mu, phi, m=24, gauss=6)
        print(kind, "view mu %.2f phi %.0f: E1 total %.7g (coarser rule %+.1e relative, columns %.1e of the largest); E0 total %.7g (%+.1e, %.1e)"
              % (mu, phi, a1.sum(), b1.sum() / a1.sum() - 1, np.abs(a1 - b1).max() / a1.max(), a0.sum(), b0.sum() / a0.sum() - 1,
                 np.abs(a0 - b0).max() / a0.max()), flush=True)
        E1.append(a1); E0.append(a0)
    np.savez_compressed(os.path.join(HERE, "first_interaction_%s.npz" % kind), first=first, surf=surf, E1=np.array(E1), E0=np.array(E0),
                        solar=np.array([fi.SOLAR_MU, fi.SOLAR_AZIMUTH]), viewMus=np.array(fi.VIEW_MUS), viewPhis=np.array(fi.VIEW_PHIS))
    print(kind, "done in %.0f s" % (time.time() - t0), flush=True)
