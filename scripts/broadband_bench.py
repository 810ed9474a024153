"""C5-style broadband run (BASELINE config I3RC_bench_SW / _LW) on one GPU or under torchrun: numLambda wavelength
bins, per-bin optics assembled on the device, photons allocated to bins on the device, device-side statistics.
Prints the set-up cost, the spectral loop time and whole-run photons/s."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcbrat3d_b200 import domains, multipleProcesses as mpx
from mcbrat3d_b200.broadband import runBroadband
from mcbrat3d_b200.monteCarloRadiativeTransfer import new_Integrator, specifyParameters, getCounters
from mcbrat3d_b200.opticalProperties import read_SSPTable
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence

ap = argparse.ArgumentParser()
ap.add_argument("--nxy", type=int, default=325); ap.add_argument("--nz", type=int, default=160)
ap.add_argument("--lambdas", type=int, default=32); ap.add_argument("--photons", type=float, default=3.2e8)
ap.add_argument("--batch", type=float, default=5e6); ap.add_argument("--lw", action="store_true")
ap.add_argument("--ext-mask", type=int, default=0)       # mcb_options.tuneExtMask: 1 = bitmap marcher, 0 / 2 = the library's choice
a = ap.parse_args()
world, rank = mpx.initializeProcesses()
local = int(os.environ.get("LOCAL_RANK", "0"))
t0 = time.perf_counter()
common, tables, case = domains.broadband_problem(nxy=a.nxy, nz=a.nz, nLambda=a.lambdas, lw=a.lw)
t1 = time.perf_counter()
d0 = read_SSPTable(tables, 1, common, setup=True)                    # grid only: new_Integrator needs the edges
g = new_Integrator(d0, device=local)
specifyParameters(g, minInverseTableSize=9001, tuneExtMask=a.ext_mask)
rs = new_RandomNumberSequence([10, 0, 0])
src = 2.0e3 * np.exp(-((np.linspace(0.45, 2.1, a.lambdas) - 0.5) / 0.6) ** 2)
if world > 1:                                                           # create the NCCL communicator outside the timed run
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.all_reduce(torch.zeros(1, device="cuda"))
    torch.cuda.synchronize()
mpx.synchronizeProcesses()
t2 = time.perf_counter()
out = runBroadband(g, tables, common, rs, int(a.photons), int(a.batch), solarMu=0.5, solarSourceFunction=None if a.lw else src,
                   LW=a.lw, surfaceTemp=case["surfaceTemp"], calcRayl=not a.lw)
t3 = time.perf_counter()
if rank == 0:
    m, e = out["mean"], out["err"]
    print("grid %dx%dx%d, %d bins, %s, %d GPU(s): problem generation %.1f s; spectral run %.2f s for %.3g photons = %.3g photons/s "
          "(includes physical-state upload, per-bin assembly%s, photon allocation, statistics)" % (
              a.nxy, a.nxy, a.nz, a.lambdas, "LW" if a.lw else "SW", world, t1 - t0, t3 - t2, out["totalNumPhotons"],
              out["totalNumPhotons"] / (t3 - t2), " + emission CDF + LW set-up pass" if a.lw else ""))
    print("bins used %d, batches %d, solarFlux %.4g; mean flux up %.4g +- %.2g, down %.4g +- %.2g, absorbed %.4g +- %.2g" % (
        int((out["freqDistr"] > 0).sum()), out["batchesCompleted"], out["solarFlux"], m["meanFluxUp"], e["meanFluxUp"],
        m["meanFluxDown"], e["meanFluxDown"], m["meanFluxAbsorbed"], e["meanFluxAbsorbed"]))
mpx.finalizeProcesses()
