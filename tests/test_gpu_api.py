"""Error behaviour through the C ABI mirrors the reference's status messages."""
import ctypes as C

import numpy as np
import pytest

from mcbrat3d_b200 import _lib, domains
from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
from mcbrat3d_b200.monteCarloRadiativeTransfer import (computeRadiativeTransfer, finalize_Integrator, new_Integrator,
                                                       reportResults, specifyParameters)
from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence

pytestmark = pytest.mark.gpu


def test_not_completely_specified():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.mcb_create(0, C.byref(h)) == 0
    done = C.c_int64(0)
    assert lib.mcb_run_batch(h, 10, 1, 0, C.byref(done)) != 0
    buf = C.create_string_buffer(256)
    lib.mcb_last_error(h, buf, 256)
    assert b"problem not completely specified" in buf.value         # INT:238
    e = np.array([0.0, 1.0, 2.0])
    assert lib.mcb_set_grid(h, 2, 2, 2, _lib.ptr(e[::-1].copy(), C.c_double), _lib.ptr(e, C.c_double), _lib.ptr(e, C.c_double)) != 0
    lib.mcb_last_error(h, buf, 256)
    assert b"Positions must be increasing" in buf.value             # OPT:526-529
    assert lib.mcb_destroy(h) == 0


def test_parameter_checks_and_missing_tables():
    dom, case = domains.homogeneous_slab()
    g = new_Integrator(dom)
    try:
        with pytest.raises(ValueError, match="intensityMus can't be 0"):
            specifyParameters(g, intensityMus=[0.0], intensityPhis=[0.0])
        with pytest.raises(ValueError, match="Both or neither"):
            specifyParameters(g, intensityMus=[0.5])
        specifyParameters(g, useRayTracing=False)           # maximum cross-section (INT:564-571) is accepted
        specifyParameters(g, useRayTracing=True)
        with pytest.raises(_lib.McbError, match="intensity information not available"):
            reportResults(g, meanIntensity=True)
        rs = new_RandomNumberSequence(1)
        ps = new_PhotonStream(0.5, 0.0, 0, rs)
        with pytest.raises(_lib.McbError, match="Didn't process any photons"):     # INT:835-836
            computeRadiativeTransfer(g, dom, rs, ps, 100)
        # reaching the run without an inverse table is an error, not a silent default
        lib, h = g._lib, g.handle
        d2, _ = domains.step_cloud()
        g2 = new_Integrator(d2)
        assert lib.mcb_set_optics(g2.handle, 1, _lib.ptr(d2.totalExt, C.c_double), _lib.ptr(d2.cumulativeExt, C.c_double),
                                  _lib.ptr(d2.ssa, C.c_double), _lib.ptr(d2.phaseFunctionIndex, C.c_int32), 0.0) == 0
        assert lib.mcb_set_solar_source(g2.handle, 0.5, 0.0) == 0
        done = C.c_int64(0)
        assert lib.mcb_run_batch(g2.handle, 10, 1, 0, C.byref(done)) != 0
        buf = C.create_string_buffer(256)
        lib.mcb_last_error(g2.handle, buf, 256)
        assert b"no inverse phase function table" in buf.value
        finalize_Integrator(g2)
    finally:
        finalize_Integrator(g)


def test_photons_beyond_stream_are_not_traced():
    """computeRT stops when the stream runs out (INT:465): asking for more than the stream holds
    processes only what is left."""
    dom, case = domains.homogeneous_slab()
    g = new_Integrator(dom)
    try:
        rs = new_RandomNumberSequence(5)
        ps = new_PhotonStream(0.5, 0.0, 3000, rs)
        assert computeRadiativeTransfer(g, dom, rs, ps, 2000) == 2000
        assert computeRadiativeTransfer(g, dom, rs, ps, 2000) == 1000
        with pytest.raises(_lib.McbError):
            computeRadiativeTransfer(g, dom, rs, ps, 2000)
    finally:
        finalize_Integrator(g)
