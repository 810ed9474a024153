"""Broadband runs (BASELINE config I3RC_bench_SW/LW): spectral photon allocation (getFrequencyDistr EMI:552-573),
the flux CDF (DRV:417-433, solar_Weighting EMI:149-208) and the driver's spectral loop with device-side statistics."""
import numpy as np
import pytest

from mcbrat3d_b200 import domains
from mcbrat3d_b200.broadband import bandWidths, kahanCDF, solar_Weighting


def test_flux_cdf_and_band_widths():
    lam = np.array([0.4, 0.5, 0.7, 1.0, 1.6])
    src = np.array([1.6, 1.9, 1.4, 0.7, 0.2]) * 1e3
    cdf, total = solar_Weighting(src, lam, 0.5)
    d = bandWidths(lam)
    assert np.allclose(d, [0.1, 0.15, 0.25, 0.45, 0.6])
    assert cdf[-1] == 1.0 and np.all(np.diff(cdf) > 0)
    assert abs(total - float(np.sum(d * 0.5 * src))) < 1e-9 * total
    c2, t2 = kahanCDF([1e16, 1.0, 1.0, 1.0, 1.0])                      # the compensated sum keeps the small terms
    assert t2 == 1e16 + 4.0


def test_oracle_frequency_distribution(orc):
    cdf = np.array([0.1, 0.1, 0.55, 1.0])                               # an empty bin (flat CDF step) gets nothing
    n = 200000
    got = orc.frequency_distribution(cdf, n)
    assert got.sum() == n and got[1] == 0
    p = np.array([0.1, 0.0, 0.45, 0.45])
    z = (got - n * p) / np.sqrt(np.maximum(n * p * (1 - p), 1.0))
    assert np.abs(z).max() < 4.5


@pytest.mark.gpu
def test_device_frequency_distribution(orc):
    from mcbrat3d_b200.broadband import getFrequencyDistr
    from mcbrat3d_b200.monteCarloRadiativeTransfer import finalize_Integrator, new_Integrator
    d, _ = domains.homogeneous_slab()
    g = new_Integrator(d)
    try:
        rng = np.random.default_rng(1)
        p = rng.random(37); p[5] = 0.0; p /= p.sum()
        cdf = np.cumsum(p); cdf[-1] = 1.0
        n = 3000001                                                      # not a multiple of 4
        got = getFrequencyDistr(g, cdf, n, seed=123)
        assert got.sum() == n and got[5] == 0
        assert np.array_equal(got, getFrequencyDistr(g, cdf, n, seed=123))            # deterministic
        assert not np.array_equal(got, getFrequencyDistr(g, cdf, n, seed=124))
        ref = orc.frequency_distribution(cdf, n)                         # MT19937 sample of the same multinomial
        sig = np.sqrt(np.maximum(2 * n * p * (1 - p), 1.0))
        assert np.abs((got - ref) / sig).max() < 4.5
        assert np.abs((got - n * p) / np.sqrt(np.maximum(n * p * (1 - p), 1.0))).max() < 4.5
        big = getFrequencyDistr(g, cdf, 2 * 10 ** 9, seed=5)             # deck-sized allocation in one call
        assert big.sum() == 2 * 10 ** 9
    finally:
        finalize_Integrator(g)


@pytest.mark.gpu
@pytest.mark.parametrize("lw", [False, True], ids=["SW", "LW"])
def test_broadband_run(lw):
    """The spectral loop end to end on the device, against a host loop over the same bins with the same photon
    ids (NumPy-assembled domains uploaded through mcb_set_optics, host-side BatchStatistics)."""
    from mcbrat3d_b200.batchStatistics import BatchStatistics
    from mcbrat3d_b200.broadband import runBroadband
    from mcbrat3d_b200.emissionAndBroadBandWeights import Weights, emission_weighting
    from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
    from mcbrat3d_b200.monteCarloRadiativeTransfer import (computeRadiativeTransfer, finalize_Integrator, new_Integrator,
                                                           reportResults, specifyParameters)
    from mcbrat3d_b200.opticalProperties import read_SSPTable
    from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence
    common, tables, case = domains.broadband_problem(nLambda=5, lw=lw)
    nLambda = 5
    src = np.array([1.7, 1.3, 0.6, 0.3, 0.1]) * 1e3
    total, per = 400000, 25000
    d0 = read_SSPTable(tables, 1, common, calcRayl=not lw)
    g = new_Integrator(d0)
    try:
        specifyParameters(g, minInverseTableSize=9001)
        rs = new_RandomNumberSequence([10, 1, 0])
        out = runBroadband(g, tables, common, rs, total, per, solarMu=0.5, solarSourceFunction=None if lw else src,
                           LW=lw, surfaceTemp=case["surfaceTemp"], calcRayl=not lw)
        assert out["freqDistr"].sum() == total and out["totalNumPhotons"] == total
        assert out["batchesCompleted"] >= nLambda
        assert out["solarFlux"] > 0 and np.all(np.isfinite(out["err"]["absorbedProfile"]))
        # host loop over the same bins / photon ids
        g2 = new_Integrator(d0)
        try:
            specifyParameters(g2, minInverseTableSize=9001, LW_flag=1.0 if lw else -1.0)
            bs = BatchStatistics()
            rs2 = new_RandomNumberSequence([10, 1, 0])
            for i in range(1, nLambda + 1):
                n = int(out["freqDistr"][i - 1])
                if n == 0:
                    continue
                d = read_SSPTable(tables, i, common, calcRayl=not lw)
                w = None
                if lw:
                    w = Weights(); emission_weighting(d, w, case["surfaceTemp"])
                nb = -(-n // per); each = n // nb; sizes = [each] * nb + ([n - each * nb] if n - each * nb else [])
                for m in sizes:
                    ps = (new_PhotonStream(theseWeights=w, numberOfPhotons=m, randomNumbers=rs2) if lw
                          else new_PhotonStream(0.5, 0.0, m, rs2))
                    done = computeRadiativeTransfer(g2, d, rs2, ps, m)
                    bs.accumulate(reportResults(g2, meanFluxUp=True, meanFluxDown=True, meanFluxAbsorbed=True,
                                                absorbedProfile=True, fluxUp=True), done)
            hm, he = bs.finalise(out["solarFlux"])
        finally:
            finalize_Integrator(g2)
        for k in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "absorbedProfile", "fluxUp"):
            np.testing.assert_allclose(out["mean"][k], hm[k], rtol=5e-5, atol=1e-6 * out["solarFlux"], err_msg=k)
        assert out["batchesCompleted"] == bs.batchesCompleted
    finally:
        finalize_Integrator(g)
