"""First-contact GPU script: Philox KAT, trace parity, quick statistics and timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
from mcbrat3d_b200 import domains, _lib
from mcbrat3d_b200.monteCarloRadiativeTransfer import *
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence
from mcbrat3d_b200.emissionAndBroadBandWeights import Weights, emission_weighting
from oracle import oracle as orc

def cmp_events(a, b, label):
    n = min(len(a), len(b))
    print(label, 'events gpu', len(a), 'oracle', len(b))
    bad = 0
    for f in ('photon','kind','ix','iy','iz','component','phaseIndex','angleIndex','order','nrn'):
        m = (a[f][:n] != b[f][:n]).sum()
        if m: print('  MISMATCH', f, m, 'first at', np.nonzero(a[f][:n] != b[f][:n])[0][:5]); bad += m
    for f in ('weight','tau','path','x','y','z'):
        x = a[f][:n].astype(np.float64); y = b[f][:n].astype(np.float64)
        rel = np.abs(x-y)/np.maximum(np.abs(y),1e-30)
        rel[(x==y)] = 0
        print('  ', f, 'max rel', rel.max(), 'nonzero', (rel>0).sum())
    dd = np.abs(a['dir'][:n]-b['dir'][:n]).max()
    print('   dir max abs', dd)
    return bad

lib = _lib.load()
d, case = domains.homogeneous_slab(ssa=0.99)
g = new_Integrator(d)
out = (C.c_uint32*8)()
lib.mcb_debug_philox(g.handle, 0, 0, 8, out)
print('philox', [hex(v) for v in out])   # 6627e8d5 e169c58d bc57ac4c 9b00dbd8

rng = np.random.default_rng(1)
for name, (dom, cs), src in [('C1', domains.homogeneous_slab(ssa=0.99), 0), ('Tirr', domains.irregular_test_domain(), 0),
                        ('C2', domains.step_cloud(ssa=0.99, solarMu=0.5), 0), ('TirrLW', domains.irregular_test_domain(), 1)]:
    for views in (False, True):
      for rr in (False, True):
        if not views and rr: continue
        gi = new_Integrator(dom)
        kw = {}
        if views:
            mus = cs.get('intensityMus', [1.0, 0.5]); phis = cs.get('intensityPhis', [0., 0.])
            specifyParameters(gi, intensityMus=mus, intensityPhis=phis, computeIntensity=True, numComps=1,
                              useRussianRouletteForIntensity=rr, zetaMin=0.3)
        specifyParameters(gi, minInverseTableSize=10001, minForwardTableSize=10001, LW_flag=1.0 if src else -1.0)
        rs = new_RandomNumberSequence([10,1,0])
        N = 2000; stride = 600
        rn = rng.random((N, stride), dtype=np.float32)
        rn[::7, 3] = 0.0; rn[::11, 5] = 1.0
        od = orc.OracleDomain(dom, tableSize=10001, forward=views)
        og = orc.OracleIntegrator(od, useRussianRouletteForIntensity=int(rr), zetaMin=0.3, LW_flag=1.0 if src else -1.0)
        if views: og.set_view_cosines(gi.intensityDirections)
        if src == 0:
            ps = new_PhotonStream(cs['solarMu'], cs['solarAzimuth'], N, rs)
            oe = og.trace(rn, 0, cs['solarMu'], cs['solarAzimuth'])
        else:
            w = Weights(); emission_weighting(dom, w, 300.0)
            frac, cdf, flux = od.emission_weighting(dom.temps, dom.lambda_um, 300.0)
            print('frac', frac, w.fracAtmsPower, 'cdf diff', np.abs(cdf-w.voxelWeights).max())
            w.voxelWeights = cdf; w.fracAtmsPower = frac
            ps = new_PhotonStream(theseWeights=w, numberOfPhotons=N, randomNumbers=rs)
            oe = og.trace(rn, 1, fracAtmsPower=frac, voxelCDF=cdf)
        t = time.time()
        ge, raw = tracePhotons(gi, dom, ps, rn, maxEventsPerPhoton=1024)
        bad = cmp_events(ge, oe, '%s views=%s rr=%s' % (name, views, rr))
        ot = og.raw_tallies()
        nt = min(len(ot), len(raw)-1)
        print('   tallies max abs diff', np.abs(raw[:nt]-ot[:nt]).max(), 'sum', raw[:nt].sum(), ot[:nt].sum(), 'time', time.time()-t)
        finalize_Integrator(gi)

# statistics + timing
for name, (dom, cs) in [('C1', domains.homogeneous_slab(ssa=0.99)), ('C2', domains.step_cloud(ssa=0.99, solarMu=0.5))]:
    for arith in (MCB_ARITH_REFERENCE, MCB_ARITH_FAST):
        gi = new_Integrator(dom)
        specifyParameters(gi, minInverseTableSize=10001, arithmetic=arith)
        rs = new_RandomNumberSequence([10,1,0])
        N = 2000000
        ps = new_PhotonStream(cs['solarMu'], cs['solarAzimuth'], N, rs)
        computeRadiativeTransfer(gi, dom, rs, ps, N)
        res = reportResults(gi, meanFluxUp=True, meanFluxDown=True, meanFluxAbsorbed=True)
        ms = lastBatchMilliseconds(gi)
        print(name, 'arith', arith, {k: float(v) for k,v in res.items()}, 'closure', float(res['meanFluxUp']+(1-dom.surfaceAlbedo)*res['meanFluxDown']+res['meanFluxAbsorbed']),
              'ms', ms, 'photons/s', N/ms*1e3, getCounters(gi))
        finalize_Integrator(gi)
