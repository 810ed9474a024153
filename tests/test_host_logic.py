"""Host-side mirrors of the reference modules (no GPU): domain assembly, parameter checks,
photon-stream bookkeeping, batch statistics, photon partition."""
import numpy as np
import pytest

from mcbrat3d_b200 import domains
from mcbrat3d_b200.batchStatistics import BatchStatistics
from mcbrat3d_b200.emissionAndBroadBandWeights import Weights, emission_weighting
from mcbrat3d_b200.monteCarloIllumination import morePhotonsExist, new_PhotonStream
from mcbrat3d_b200.multipleProcesses import photonRange
from mcbrat3d_b200.opticalProperties import Domain
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence
from mcbrat3d_b200.scatteringPhaseFunctions import henyeyGreenstein, new_PhaseFunctionTable, rayleigh


def test_domain_assembly_cumulative_extinction():
    """getOpticalPropertiesByComponent (OPT:1022-1061): last cumulative fraction is 1 wherever
    there is extinction (OPT:1052-1053), partial-height uniform components are expanded."""
    d, _ = domains.irregular_test_domain()
    nc = d.cumulativeExt.shape[0]
    assert nc == 2 and d.totalExt.shape == (8, 10, 12)
    has = d.totalExt > 0
    assert np.allclose(d.cumulativeExt[nc - 1][has], 1.0)
    assert np.all(np.diff(d.cumulativeExt, axis=0) >= 0)
    # the gas component lives on levels 2..4 only
    gas = d.totalExt * (d.cumulativeExt[1] - d.cumulativeExt[0])
    assert np.allclose(gas[1:4], np.array([3.0, 2.0, 1.0])[:, None, None]) and np.all(gas[4:] == 0) and np.all(gas[0] == 0)
    assert np.all(d.phaseFunctionIndex[0][d.totalExt * d.cumulativeExt[0] == 0] == 0)


def test_domain_rejects_bad_input():
    with pytest.raises(ValueError, match="increasing"):
        Domain([0, 1, 1], [0, 1], [0, 1])
    d = Domain([0, 1, 2], [0, 1], [0, 1, 2])
    tab = new_PhaseFunctionTable([henyeyGreenstein(0.8, 8)], key=[1.0])
    with pytest.raises(ValueError, match="extinction must be >= 0"):
        d.addOpticalComponent("c", -np.ones((2, 1, 2)), np.ones((2, 1, 2)), np.ones((2, 1, 2), np.int32), tab)
    with pytest.raises(ValueError, match="singleScatteringAlbedo"):
        d.addOpticalComponent("c", np.ones((2, 1, 2)), 2 * np.ones((2, 1, 2)), np.ones((2, 1, 2), np.int32), tab)
    with pytest.raises(ValueError, match="phase function index"):
        d.addOpticalComponent("c", np.ones((2, 1, 2)), np.ones((2, 1, 2)), 2 * np.ones((2, 1, 2), np.int32), tab)
    with pytest.raises(ValueError, match="vertical extent"):
        d.addOpticalComponent("c", np.ones(2), np.ones(2), np.ones(2, np.int32), tab, zLevelBase=2)
    with pytest.raises(ValueError, match="no optical components"):
        d.getOpticalPropertiesByComponent()


def test_photon_stream_bookkeeping():
    rs = new_RandomNumberSequence([10, 1, 0])
    a = new_PhotonStream(0.5, 0.0, 1000, rs)
    b = new_PhotonStream(0.5, 0.0, 500, rs)
    assert (a.firstPhotonId, b.firstPhotonId, rs.nextPhotonId) == (0, 1000, 1500)
    assert morePhotonsExist(a)
    a.currentPhoton = 1001
    assert not morePhotonsExist(a)
    with pytest.raises(ValueError, match="solarMu"):
        new_PhotonStream(0.0, 0.0, 10, rs)
    with pytest.raises(ValueError, match="solarAzimuth"):
        new_PhotonStream(0.5, 400.0, 10, rs)
    assert new_RandomNumberSequence([10, 1, 0]).seed != new_RandomNumberSequence([10, 2, 0]).seed
    assert new_RandomNumberSequence([10, 1, 0]).seed == rs.seed


def test_emission_weighting_matches_oracle(orc):
    d, _ = domains.homogeneous_lw()
    w = Weights()
    flux = emission_weighting(d, w, 300.0)
    frac, cdf, oflux = orc.OracleDomain(d, tableSize=9001).emission_weighting(d.temps, d.lambda_um, 300.0)
    assert w.fracAtmsPower == pytest.approx(frac, rel=1e-12)
    assert flux == pytest.approx(oflux, rel=1e-12)
    np.testing.assert_allclose(w.voxelWeights, cdf, rtol=0, atol=4e-16)
    assert np.array_equal(w.levelWeights, w.voxelWeights[:, -1, -1]) and w.colWeights.shape == (20, 20)


def test_batch_statistics_match_driver_formulas(orc):
    """BatchStatistics == orc_finalise_stats == DRV:1023-1052, 1188-1228."""
    rng = np.random.default_rng(3)
    bs = BatchStatistics()
    xs = rng.random((12, 5)); ns = rng.integers(900, 1100, 12)
    for x, n in zip(xs, ns):
        bs.accumulate({"flux": x}, int(n))
    mean, err = bs.finalise(solarFlux=1.7)
    m, e = orc.finalise(np.concatenate([bs.moment1["flux"], bs.moment2["flux"]]), 1.7, bs.totalNumPhotons, 12)
    np.testing.assert_allclose(mean["flux"], m, rtol=1e-14)
    np.testing.assert_allclose(err["flux"], e, rtol=1e-12)
    half = BatchStatistics(); other = BatchStatistics()
    for i, (x, n) in enumerate(zip(xs, ns)):
        (half if i % 2 else other).accumulate({"flux": x}, int(n))
    half.merge(other)
    m2, e2 = half.finalise(1.7)
    np.testing.assert_allclose(m2["flux"], mean["flux"], rtol=1e-13)
    np.testing.assert_allclose(e2["flux"], err["flux"], rtol=1e-10)


def test_photon_range_partition():
    for N in (0, 1, 7, 1000, 10 ** 9 + 3):
        for G in (1, 2, 3, 8):
            ranges = [photonRange(N, G, r) for r in range(G)]
            assert ranges[0][0] == 0 and sum(c for _, c in ranges) == N
            for (f0, c0), (f1, _) in zip(ranges, ranges[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in ranges) - min(c for _, c in ranges) <= 1


def test_rate_proportional_photon_shares():
    from mcbrat3d_b200.multipleProcesses import photonShares
    n = 8 * 125_000_000
    eq = photonShares(n, [1.0] * 8)
    assert sum(eq) == n and max(eq) - min(eq) <= 8
    rates = [1.0, 1.0, 0.95, 1.0, 1.02, 1.0, 1.0, 0.99]
    sh = photonShares(n, rates)
    assert sum(sh) == n and all(s > 0 for s in sh)
    for s_, r in zip(sh, rates):
        assert abs(s_ - n * r / sum(rates)) <= 8
    # every rank finishes at the same time: share / rate is constant
    t = [s_ / r for s_, r in zip(sh, rates)]
    assert max(t) / min(t) - 1.0 < 1e-6
    assert sum(photonShares(10, [0.0, float("nan"), 1.0])) == 10 and min(photonShares(10, [0.0, float("nan"), 1.0])) >= 3
    assert photonShares(7, [3.0]) == [7]


def test_synthetic_domains_shapes():
    d, c = domains.step_cloud()
    assert (d.numX, d.numY, d.numZ) == (32, 1, 32) and d.xPosition[1] == 15.625
    tau = (d.totalExt * np.diff(d.zPosition)[:, None, None]).sum(axis=0)[0]
    assert np.allclose(tau[:16], 2.0, rtol=1e-6) and np.allclose(tau[16:], 18.0, rtol=1e-6)
    d3, c3 = domains.landsat_cloud(nxy=32)
    assert d3.numZ == 119 and d3.zPosition[0] == 200.0 and d3.xPosition[1] == 30.0
    assert (d3.totalExt[:, :, :].sum(axis=0) == 0).mean() == pytest.approx(0.2, abs=0.05)   # ~20 % clear columns
    assert np.all(d3.phaseFunctionIndex[0][d3.totalExt == 0] == 0)
