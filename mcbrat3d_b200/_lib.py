"""ctypes binding of ``csrc/libmcbrat_cuda.so`` -- the C ABI in ``include/mcbrat_cuda.h``.

This is the same ABI the Fortran ISO_C_BINDING shim binds (``fortran/mcbrat_cuda_mod.f90``).
There is no fallback: if the library is missing or no CUDA device is present the calls
raise, they never route to a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCB_LIB_DEBUG=1 loads the bounds-checked build of the throughput kernel (csrc/Makefile target dbg)
# MCB_LIB_VARIANT=name loads csrc/libmcbrat_cuda_<name>.so (kernel A/B measurements: csrc/Makefile target variant)
LIB_PATH = os.path.join(_HERE, "csrc", "libmcbrat_cuda_dbg.so" if os.environ.get("MCB_LIB_DEBUG") == "1"
                        else "libmcbrat_cuda_%s.so" % os.environ["MCB_LIB_VARIANT"] if os.environ.get("MCB_LIB_VARIANT")
                        else "libmcbrat_cuda.so")

MCB_ARITH_FAST = 0
MCB_ARITH_REFERENCE = 1


class mcb_options(C.Structure):
    _fields_ = [("useRayTracing", C.c_int32), ("useRussianRoulette", C.c_int32),
                ("russianRouletteW", C.c_float), ("useRussianRouletteForIntensity", C.c_int32),
                ("zetaMin", C.c_float), ("useHybridPhaseFunsForIntenCalcs", C.c_int32),
                ("numOrdersOrigPhaseFunIntenCalcs", C.c_int32), ("limitIntensityContributions", C.c_int32),
                ("maxIntensityContribution", C.c_float), ("LW_flag", C.c_float),
                ("arithmetic", C.c_int32),
                # measurement knobs (0 = the library's own choice): see include/mcbrat_cuda.h
                ("tuneKernel", C.c_int32), ("tuneLayout", C.c_int32), ("tuneBlocksPerSM", C.c_int32),
                ("tuneParkThreshold", C.c_int32), ("tuneLeCarry", C.c_int32), ("tuneExtMask", C.c_int32),
                ("tuneBurst", C.c_int32), ("tuneLeap", C.c_int32), ("tuneLeapLanes", C.c_int32), ("reserved", C.c_int32 * 1)]


MCB_KERNEL_PARK, MCB_KERNEL_POOL = 1, 2
MCB_LAYOUT_LINEAR, MCB_LAYOUT_BRICKS = 1, 2


class mcb_component(C.Structure):
    _fields_ = [("kind", C.c_int32), ("physIndex", C.c_int32), ("nTable", C.c_int32), ("zLevelBase", C.c_int32),
                ("key", C.POINTER(C.c_float)), ("ext", C.POINTER(C.c_double)), ("ssa", C.POINTER(C.c_double)),
                ("phaseIdx", C.POINTER(C.c_int32))]


MCB_COMP_VOLEXT, MCB_COMP_ABSXSEC, MCB_COMP_PROFILE = 0, 1, 2


class mcb_counters(C.Structure):
    _fields_ = ([(n, C.c_int64) for n in ("photons", "crossings", "scatters", "surfaceHits", "topExits", "bad",
                                          "leRays", "leCrossings", "rouletteKills", "surfaceKills", "leaps", "leapCells")]
                + [("reserved", C.c_int64 * 4)])

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_ if n != "reserved"}


# numpy view of mcb_event / orc_event (96 bytes)
EVENT_DTYPE = np.dtype([("photon", "<i4"), ("kind", "<i4"), ("ix", "<i4"), ("iy", "<i4"), ("iz", "<i4"),
                        ("component", "<i4"), ("phaseIndex", "<i4"), ("angleIndex", "<i4"), ("order", "<i4"),
                        ("nrn", "<i4"), ("weight", "<f4"), ("tau", "<f4"), ("path", "<f8"),
                        ("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("dir", "<f4", (3,)), ("pad", "<i4")])
assert EVENT_DTYPE.itemsize == 96

# every symbol include/mcbrat_cuda.h declares
EXPORTS = ["mcb_create", "mcb_destroy", "mcb_last_error", "mcb_version", "mcb_set_stream", "mcb_synchronize",
           "mcb_set_grid", "mcb_set_optics", "mcb_set_physical", "mcb_assemble_optics", "mcb_get_optics", "mcb_set_inverse_table", "mcb_build_inverse_table", "mcb_build_inverse_table_legendre", "mcb_get_inverse_table", "mcb_build_forward_table", "mcb_build_forward_table_general", "mcb_get_forward_table", "mcb_set_forward_table", "mcb_set_views",
           "mcb_default_options", "mcb_set_options", "mcb_set_solar_source", "mcb_set_thermal_source",
           "mcb_build_thermal_source", "mcb_get_thermal_source", "mcb_frequency_distribution", "mcb_run_batch", "mcb_accumulate_batch", "mcb_stats_reset", "mcb_run_batches",
           "mcb_stats_buffer", "mcb_get_statistics", "mcb_last_batch_ms",
           "mcb_get_counters", "mcb_get_results", "mcb_tally_buffer", "mcb_get_raw_tallies", "mcb_run_trace",
           "mcb_debug_philox", "mcb_debug_gather_probe", "mcb_debug_distance_map", "mcb_comm_unique_id", "mcb_comm_init", "mcb_comm_info",
           "mcb_reduce_tallies", "mcb_reduce_statistics", "mcb_comm_destroy"]

_lib: Optional[C.CDLL] = None

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p


def load() -> C.CDLL:
    """Load the CUDA library (raises if it has not been built: no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("mcbrat3d_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; "
                           "g.build()'` or `make -C mcbrat3d_b200/csrc`; there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.mcb_create.argtypes = [C.c_int, C.POINTER(_vp)]
    lib.mcb_destroy.argtypes = [_vp]
    lib.mcb_last_error.argtypes = [_vp, C.c_char_p, C.c_int]
    lib.mcb_set_stream.argtypes = [_vp, _vp]
    lib.mcb_synchronize.argtypes = [_vp]
    lib.mcb_set_grid.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]
    lib.mcb_set_optics.argtypes = [_vp, C.c_int, _dp, _dp, _dp, _ip, C.c_double]
    lib.mcb_set_physical.argtypes = [_vp, C.c_int, _dp, _dp, _dp]
    lib.mcb_assemble_optics.argtypes = [_vp, C.c_int, C.POINTER(mcb_component), C.c_int, C.c_double]
    lib.mcb_get_optics.argtypes = [_vp, _dp, _dp, _dp, _ip]
    lib.mcb_set_inverse_table.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _fp]
    lib.mcb_build_inverse_table.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _ip, _fp, _fp]
    lib.mcb_get_inverse_table.argtypes = [_vp, C.c_int, _fp, C.c_int64]
    lib.mcb_build_forward_table.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _ip, _fp]
    lib.mcb_build_inverse_table_legendre.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _ip, _fp]
    lib.mcb_build_forward_table_general.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _ip, _fp, _ip, _fp, _fp, C.c_float]
    lib.mcb_get_forward_table.argtypes = [_vp, C.c_int, _fp, C.c_int64]
    lib.mcb_set_forward_table.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _fp, _fp]
    lib.mcb_set_views.argtypes = [_vp, C.c_int, _fp]
    lib.mcb_default_options.argtypes = [C.POINTER(mcb_options)]
    lib.mcb_default_options.restype = None
    lib.mcb_set_options.argtypes = [_vp, C.POINTER(mcb_options)]
    lib.mcb_set_solar_source.argtypes = [_vp, C.c_float, C.c_float]
    lib.mcb_set_thermal_source.argtypes = [_vp, C.c_double, _dp]
    lib.mcb_build_thermal_source.argtypes = [_vp, _dp, C.c_double, C.c_double, _dp, _dp]
    lib.mcb_get_thermal_source.argtypes = [_vp, _dp, _dp, C.c_int64]
    lib.mcb_frequency_distribution.argtypes = [_vp, C.c_int, _dp, C.c_int64, C.c_uint64, C.POINTER(C.c_int64)]
    lib.mcb_run_batch.argtypes = [_vp, C.c_int64, C.c_uint64, C.c_uint64, C.POINTER(C.c_int64)]
    lib.mcb_accumulate_batch.argtypes = [_vp, C.c_int64, C.c_uint64, C.c_uint64, C.POINTER(C.c_int64)]
    lib.mcb_stats_reset.argtypes = [_vp]
    lib.mcb_run_batches.argtypes = [_vp, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, C.POINTER(C.c_int64)]
    lib.mcb_stats_buffer.argtypes = [_vp, C.POINTER(_vp), C.POINTER(C.c_int64)]
    lib.mcb_get_statistics.argtypes = [_vp, C.c_double, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int64)]
    lib.mcb_last_batch_ms.argtypes = [_vp, _fp]
    lib.mcb_get_counters.argtypes = [_vp, C.POINTER(mcb_counters)]
    lib.mcb_get_results.argtypes = [_vp, C.c_int64, _fp, _fp, _fp, _fp, _fp, _fp]
    lib.mcb_tally_buffer.argtypes = [_vp, C.POINTER(_vp), C.POINTER(C.c_int64)]
    lib.mcb_get_raw_tallies.argtypes = [_vp, _dp, C.c_int64]
    lib.mcb_run_trace.argtypes = [_vp, C.c_int64, _fp, C.c_int64, C.c_int32, _vp, C.c_int64, C.POINTER(C.c_int64)]
    lib.mcb_debug_philox.argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint32)]
    lib.mcb_debug_gather_probe.argtypes = [_vp, C.c_int64, C.c_int, C.c_int, C.c_int, _dp]
    lib.mcb_debug_distance_map.argtypes = [_vp, C.POINTER(C.c_uint8), C.c_int64]
    lib.mcb_comm_unique_id.argtypes = [_vp]
    lib.mcb_comm_init.argtypes = [_vp, C.c_int, C.c_int, _vp]
    lib.mcb_comm_info.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.mcb_reduce_tallies.argtypes = [_vp, C.c_int]
    lib.mcb_reduce_statistics.argtypes = [_vp, C.c_int]
    lib.mcb_comm_destroy.argtypes = [_vp]
    for name in EXPORTS:
        if name != "mcb_default_options":
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def ptr(a: Optional[np.ndarray], ctype):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))


class McbError(RuntimeError):
    """A failure reported through the C ABI (the Fortran shim maps it to setStateToFailure)."""


def check(lib, handle, rc, where):
    if rc != 0:
        buf = C.create_string_buffer(512)
        if handle:
            lib.mcb_last_error(handle, buf, 512)
        raise McbError("%s: %s" % (where, buf.value.decode() or ("error code %d" % rc)))
