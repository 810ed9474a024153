"""per-wavelength optical-property assembly (read_SSPTable's inner loops OPT:204-299 +
getOpticalPropertiesByComponent OPT:1022-1061).  CPU: the oracle's C restatement against the NumPy host mirror.
GPU: the device kernel (mcb_assemble_optics) against the oracle, bit for bit, and a multi-wavelength run."""
import ctypes as C

import numpy as np
import pytest

from mcbrat3d_b200 import domains
from mcbrat3d_b200.opticalProperties import calc_RayleighScattering, fetchOpticalProperties, read_SSPTable


def oracle_components(tables, li, common, lam, rayl):
    t = tables[0]
    comps = [dict(kind=0, physIndex=1, zLevelBase=1, key=t.components[0].key, ext=t.components[0].extinctionT[li - 1],
                  ssa=t.components[0].singleScatteringAlbedoT[li - 1]),
             dict(kind=1, zLevelBase=1, ext=t.components[1].xsec[li - 1])]
    if rayl:
        e, s, i, _ = calc_RayleighScattering(lam, common.rho[:, 0, 0], common.numConc[:, 0, 0])
        comps.append(dict(kind=2, zLevelBase=1, ext=e, ssa=s, idx=i))
    return comps


@pytest.mark.parametrize("li,rayl,setup", [(1, False, False), (3, True, False), (6, True, True)])
def test_oracle_matches_numpy_mirror(orc, li, rayl, setup):
    common, tables, case = domains.broadband_problem()
    d = read_SSPTable(tables, li, common, setup=setup, calcRayl=rayl)
    comps = oracle_components(tables, li, common, d.lambda_um, rayl and not setup)
    rc, total, cum, ssa, idx = orc.assemble_optics(d.numX, d.numY, d.numZ, common.massConc, common.Reff,
                                                   common.numConc[:, 0, 0], comps, setup=setup)
    assert rc == 0
    assert np.array_equal(total, d.totalExt) and np.array_equal(cum, d.cumulativeExt)
    assert np.array_equal(ssa, d.ssa) and np.array_equal(idx, d.phaseFunctionIndex)
    cloudy = total > 0
    assert np.allclose(cum[-1][cloudy], 1.0, rtol=0, atol=4e-16)          # OPT:1052-1053: last cumulative fraction is 1
    if setup:
        assert np.all(idx[0] == 1)                                         # OPT:278: no phase-function work in the setup pass
    else:
        assert idx[0].max() > 1


def test_effective_radius_outside_table_is_an_error(orc):
    common, tables, case = domains.broadband_problem()
    common.Reff[5, 3, 3, 0] = 40.0; common.massConc[5, 3, 3, 0] = 0.1
    with pytest.raises(ValueError, match="Effective radius outside of table range"):        # OPT:288-289
        read_SSPTable(tables, 1, common)
    comps = oracle_components(tables, 1, common, 0.45, False)
    rc, *_ = orc.assemble_optics(24, 24, 20, common.massConc, common.Reff, common.numConc[:, 0, 0], comps)
    assert rc == 1


@pytest.mark.gpu
@pytest.mark.parametrize("li,rayl,setup", [(1, False, False), (3, True, False), (6, True, True)])
def test_device_assembly_bit_exact(orc, li, rayl, setup):
    from mcbrat3d_b200.monteCarloRadiativeTransfer import finalize_Integrator, new_Integrator
    common, tables, case = domains.broadband_problem()
    host = read_SSPTable(tables, li, common, setup=setup, calcRayl=rayl)
    g = new_Integrator(host)
    try:
        dev = read_SSPTable(tables, li, common, setup=setup, calcRayl=rayl, thisIntegrator=g)
        assert dev.totalExt is None and dev.deviceOwner is g
        fetchOpticalProperties(dev)
        comps = oracle_components(tables, li, common, host.lambda_um, rayl and not setup)
        rc, total, cum, ssa, idx = orc.assemble_optics(host.numX, host.numY, host.numZ, common.massConc, common.Reff,
                                                       common.numConc[:, 0, 0], comps, setup=setup)
        assert np.array_equal(dev.totalExt, total) and np.array_equal(dev.cumulativeExt, cum)
        assert np.array_equal(dev.ssa, ssa) and np.array_equal(dev.phaseFunctionIndex, idx)
    finally:
        finalize_Integrator(g)


@pytest.mark.gpu
def test_device_assembly_errors():
    from mcbrat3d_b200._lib import McbError
    from mcbrat3d_b200.monteCarloRadiativeTransfer import finalize_Integrator, new_Integrator
    common, tables, case = domains.broadband_problem()
    host = read_SSPTable(tables, 1, common)
    g = new_Integrator(host)
    try:
        common.Reff[5, 3, 3, 0] = 40.0; common.massConc[5, 3, 3, 0] = 0.1
        with pytest.raises(McbError, match="Effective radius outside of table range"):
            read_SSPTable(tables, 1, common, thisIntegrator=g)
    finally:
        finalize_Integrator(g)


@pytest.mark.gpu
def test_multi_wavelength_run_matches_host_assembled_domains():
    """The C5 loop: every wavelength bin assembled on the device and traced; the same bins assembled in NumPy,
    uploaded with mcb_set_optics and traced with the same photons give the same fluxes."""
    from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
    from mcbrat3d_b200.monteCarloRadiativeTransfer import (computeRadiativeTransfer, finalize_Integrator, new_Integrator,
                                                           reportResults, specifyParameters)
    from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence
    common, tables, case = domains.broadband_problem()
    n = 100000
    out = {}
    for where in ("device", "host"):
        d0 = read_SSPTable(tables, 1, common, calcRayl=True)
        g = new_Integrator(d0)
        try:
            specifyParameters(g, minInverseTableSize=9001)
            rs = new_RandomNumberSequence([10, 1, 0])
            res = []
            for li in range(1, case["numLambda"] + 1):
                d = read_SSPTable(tables, li, common, calcRayl=True, thisIntegrator=g if where == "device" else None)
                ps = new_PhotonStream(case["solarMu"], case["solarAzimuth"], n, rs)
                assert computeRadiativeTransfer(g, d, rs, ps, n) == n
                r = reportResults(g, meanFluxUp=True, meanFluxDown=True, meanFluxAbsorbed=True, absorbedProfile=True)
                closure = float(r["meanFluxUp"]) + (1.0 - d.surfaceAlbedo) * float(r["meanFluxDown"]) + float(r["meanFluxAbsorbed"])
                assert abs(closure - 1.0) < 3e-3
                res.append(r)
            out[where] = res
        finally:
            finalize_Integrator(g)
    for a, b in zip(out["device"], out["host"]):
        for k in a:
            np.testing.assert_allclose(a[k], b[k], rtol=2e-5, atol=1e-7, err_msg=k)
    ups = [float(r["meanFluxUp"]) for r in out["device"]]
    assert max(ups) - min(ups) > 1e-3                       # the bins really differ
