import sys; sys.path.insert(0,'/root/repo')
from mcbrat3d_b200 import domains
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.monteCarloRadiativeTransfer import *
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence
for name,mk in (("c3",lambda: domains.landsat_cloud(ssa=0.99)),("c2",lambda: domains.step_cloud(ssa=0.99, solarMu=0.5))):
    dom,case=mk()
    g=new_Integrator(dom)
    specifyParameters(g, intensityMus=case["intensityMus"], intensityPhis=case["intensityPhis"], computeIntensity=True, useRussianRouletteForIntensity=True, zetaMin=0.3, minInverseTableSize=10001, minForwardTableSize=10001)
    rs=new_RandomNumberSequence([10,1,0]); n=2000000
    for b in range(2):
        ps=new_PhotonStream(case["solarMu"],case["solarAzimuth"],n,rs); computeRadiativeTransfer(g,dom,rs,ps,n)
    c=getCounters(g); ms=lastBatchMilliseconds(g)
    print(name, "ms",ms, "leRays/photon",c["leRays"]/n,"leCross/ray",c["leCrossings"]/c["leRays"],"leCross/photon",c["leCrossings"]/n,"cross/photon",c["crossings"]/n, "total crossings/s", (c["leCrossings"]+c["crossings"])/ms*1e3)
