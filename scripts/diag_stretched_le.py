"""Diagnostic (GPU): the garbage radiance seen once on the stretched grid with the plain local estimate."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import first_interaction as fi
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.monteCarloRadiativeTransfer import (computeRadiativeTransfer, finalize_Integrator, getCounters,
                                                       new_Integrator, reportResults, specifyParameters)
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence

def run(tag, kind="stretched"):
  roulette = tag == "roulette"
  mus, phis = fi.VIEW_MUS, fi.VIEW_PHIS
  if tag == "nadir": mus, phis = [1.0], [0.0]
  if tag == "down": mus, phis = [-0.5], [120.0]
  if tag == "slant": mus, phis = [0.7, 0.4], [45.0, 200.0]
  n = 1_000_000 if tag == "n1e6" else 2_000_000
  for k, scale in enumerate((0.005, 0.01)):
      dom, med = fi.scene(kind, albedo=0.0, ssaScale=scale)
      g = new_Integrator(dom)
      specifyParameters(g, intensityMus=mus, intensityPhis=phis, computeIntensity=True, useRussianRouletteForIntensity=False, zetaMin=0.3)
      specifyParameters(g, minInverseTableSize=9001, minForwardTableSize=9001, useRussianRoulette=roulette)
      rs = new_RandomNumberSequence([31 + k, 5, 0])
      for b in range(16):
          ps = new_PhotonStream(fi.SOLAR_MU, fi.SOLAR_AZIMUTH, n, rs)
          computeRadiativeTransfer(g, dom, rs, ps, n)
          res = reportResults(g, intensity=True, intensityByComponent=True, fluxUp=True, fluxDown=True, volumeAbsorption=True)
          I = np.asarray(res["intensity"], np.float64)
          c = getCounters(g)
          odd = np.argwhere((I > 50 * scale) | (I < -0.5 * scale) | ~np.isfinite(I))
          other = [k2 for k2 in ("fluxUp", "fluxDown", "volumeAbsorption") if not np.isfinite(res[k2]).all() or np.abs(res[k2]).max() > 1e4]
          if len(odd) or other or c["bad"]:
              byc = np.asarray(res["intensityByComponent"], np.float64)
              print(tag, kind, "scale", scale, "batch", b, "bad", c["bad"], "odd entries (dir, iy, ix):", odd[:6].tolist(),
                    "values", [float(I[tuple(o)]) for o in odd[:6]], "by component", [[float(byc[cc][tuple(o)]) for cc in range(byc.shape[0])] for o in odd[:3]],
                    "other", other, flush=True)
      print(tag, kind, "scale", scale, "done; counters", {k2: c[k2] for k2 in ("photons", "scatters", "bad", "leRays", "leCrossings", "crossings")},
            "mean intensity", np.round(I.reshape(len(mus), -1).mean(axis=1) / scale, 5).tolist(), flush=True)
      finalize_Integrator(g)


for spec in sys.argv[1:]:
    run(*spec.split(":"))
