"""ctypes binding of the CPU ORACLE (``oracle/libmcb_oracle.so``) -- TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker or the reported CPU
baseline.  The product package ``mcbrat3d_b200`` never imports it.
PARITY STATUS: parity unpinned (see ``mcb_oracle.h``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcb_oracle.so")

EVENT_DTYPE = np.dtype([("photon", "<i4"), ("kind", "<i4"), ("ix", "<i4"), ("iy", "<i4"), ("iz", "<i4"),
                        ("component", "<i4"), ("phaseIndex", "<i4"), ("angleIndex", "<i4"), ("order", "<i4"),
                        ("nrn", "<i4"), ("weight", "<f4"), ("tau", "<f4"), ("path", "<f8"),
                        ("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("dir", "<f4", (3,)), ("pad", "<i4")])


class orc_options(C.Structure):
    _fields_ = [("useRayTracing", C.c_int), ("useRussianRoulette", C.c_int), ("RussianRouletteW", C.c_float),
                ("useRussianRouletteForIntensity", C.c_int), ("zetaMin", C.c_float),
                ("useHybridPhaseFunsForIntenCalcs", C.c_int), ("numOrdersOrigPhaseFunIntenCalcs", C.c_int),
                ("limitIntensityContributions", C.c_int), ("maxIntensityContribution", C.c_float),
                ("LW_flag", C.c_float)]


class orc_component(C.Structure):
    _fields_ = [("kind", C.c_int32), ("physIndex", C.c_int32), ("nTable", C.c_int32), ("zLevelBase", C.c_int32),
                ("key", C.POINTER(C.c_float)), ("ext", C.POINTER(C.c_double)), ("ssa", C.POINTER(C.c_double)),
                ("phaseIdx", C.POINTER(C.c_int32))]


class orc_counters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("photons", "crossings", "scatters", "surfaceHits", "topExits", "bad",
                                         "leRays", "leCrossings", "rouletteKills", "rnDrawn")]


class orc_rng(C.Structure):
    _fields_ = [("mode", C.c_int), ("mt", C.c_uint32 * 624), ("mti", C.c_int), ("inj", C.POINTER(C.c_float)),
                ("ninj", C.c_int64), ("pos", C.c_int64), ("exhausted", C.c_int), ("ndrawn", C.c_int64)]


class orc_stats(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in
                ("meanFluxUpStats", "meanFluxDownStats", "meanFluxAbsorbedStats", "fluxUpStats", "fluxDownStats",
                 "fluxAbsorbedStats", "absorbedProfileStats", "absorbedVolumeStats", "radianceStats")]


class _integrator_head(C.Structure):          # leading members of orc_integrator we read back
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nc", C.c_int),
                ("xyRegularlySpaced", C.c_int), ("zRegularlySpaced", C.c_int),
                ("deltaX", C.c_double), ("deltaY", C.c_double), ("deltaZ", C.c_double),
                ("x0", C.c_double), ("y0", C.c_double), ("z0", C.c_double),
                ("xPosition", C.c_void_p), ("yPosition", C.c_void_p), ("zPosition", C.c_void_p),
                ("opt", orc_options), ("computeIntensity", C.c_int), ("nDir", C.c_int),
                ("intensityDirections", C.POINTER(C.c_float)),
                ("fluxUp", C.POINTER(C.c_float)), ("fluxDown", C.POINTER(C.c_float)),
                ("fluxAbsorbed", C.POINTER(C.c_float)), ("volumeAbsorption", C.POINTER(C.c_float)),
                ("intensity", C.POINTER(C.c_float)), ("intensityByComponent", C.POINTER(C.c_float)),
                ("intensityExcess", C.POINTER(C.c_float)), ("cnt", orc_counters),
                ("trace", C.c_void_p), ("traceCap", C.c_int64), ("traceN", C.c_int64), ("tracePhoton0", C.c_int32)]


_lib = None
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, seconds).  Building the checker is not using it."""
    src = os.path.join(_HERE, "mcb_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "mcb_oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmcb_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(LIB_PATH)
    lib.orc_domain_new.restype = C.c_void_p
    lib.orc_domain_new.argtypes = [C.c_int] * 4 + [_dp] * 6 + [C.POINTER(C.c_int32), C.c_double]
    lib.orc_domain_set_inverse.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _fp]
    lib.orc_domain_set_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _fp, _fp]
    lib.orc_domain_free.argtypes = [C.c_void_p]
    lib.orc_integrator_new.restype = C.c_void_p
    lib.orc_integrator_new.argtypes = [C.c_void_p]
    lib.orc_integrator_set_options.argtypes = [C.c_void_p, C.POINTER(orc_options)]
    lib.orc_integrator_set_views.argtypes = [C.c_void_p, C.c_int, _fp, _fp]
    lib.orc_integrator_set_view_cosines.argtypes = [C.c_void_p, C.c_int, _fp]
    lib.orc_integrator_set_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.orc_integrator_free.argtypes = [C.c_void_p]
    lib.orc_default_options.argtypes = [C.POINTER(orc_options)]
    lib.orc_make_direction_cosines.argtypes = [C.c_float, C.c_float, _fp]
    lib.orc_rng_init_scalar.argtypes = [C.POINTER(orc_rng), C.c_uint32]
    lib.orc_rng_init_array.argtypes = [C.POINTER(orc_rng), C.POINTER(C.c_uint32), C.c_int]
    lib.orc_rng_int.argtypes = [C.POINTER(orc_rng)]
    lib.orc_rng_int.restype = C.c_uint32
    lib.orc_rng_real.argtypes = [C.POINTER(orc_rng)]
    lib.orc_rng_real.restype = C.c_float
    lib.orc_rng_double.argtypes = [C.POINTER(orc_rng)]
    lib.orc_rng_double.restype = C.c_double
    lib.orc_findIndexDouble.argtypes = [C.c_double, _dp, C.c_int, C.c_int]
    lib.orc_findIndexMixed.argtypes = [C.c_float, _dp, C.c_int, C.c_int]
    lib.orc_findCDFIndex.argtypes = [C.c_float, _dp, C.c_int, C.c_int]
    lib.orc_emission_weighting.restype = C.c_double
    lib.orc_emission_weighting.argtypes = [C.c_void_p, _dp, C.c_double, C.c_double, _dp, _dp]
    lib.orc_trace_photons.restype = C.c_int64
    lib.orc_trace_photons.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_double, _dp,
                                      C.c_int64, _fp, C.c_int64]
    lib.orc_photons_directional.restype = C.c_void_p
    lib.orc_photons_directional.argtypes = [C.c_float, C.c_float, C.c_int64, C.POINTER(orc_rng)]
    lib.orc_photons_free.argtypes = [C.c_void_p]
    lib.orc_compute_radiative_transfer.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(orc_rng), C.c_void_p, C.c_int64, C.c_int,
                                                   C.POINTER(C.c_int64)]
    lib.orc_report_results.restype = None
    lib.orc_report_results.argtypes = [C.c_void_p] + [_fp] * 10
    lib.orc_run_batches.restype = C.c_int64
    lib.orc_run_batches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_double, _dp,
                                    C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.POINTER(orc_stats)]
    lib.orc_assemble_optics.argtypes = [C.c_int] * 4 + [_dp, _dp, _dp, C.c_int, C.POINTER(orc_component), C.c_int,
                                        _dp, _dp, _dp, C.POINTER(C.c_int32)]
    lib.orc_frequency_distribution.argtypes = [C.c_int, _dp, C.c_int64, C.POINTER(orc_rng), C.POINTER(C.c_int64)]
    lib.orc_frequency_distribution.restype = None
    lib.orc_inverse_phase_function.argtypes = [C.c_int, _fp, _fp, C.c_int, _fp]
    lib.orc_inverse_phase_function.restype = None
    lib.orc_forward_phase_function.argtypes = [C.c_int, _fp, C.c_int, _fp]
    lib.orc_forward_phase_function.restype = None
    lib.orc_lobatto_terms.argtypes = [C.c_int, _fp, _fp]
    lib.orc_lobatto_terms.restype = None
    lib.orc_phase_function_values.argtypes = [C.c_int, _fp, C.c_int, _fp, _fp, C.c_int, _fp, _fp]
    lib.orc_phase_function_values.restype = None
    lib.orc_inversion_inputs_legendre.argtypes = [C.c_int, _fp, _fp, _fp]
    lib.orc_inversion_inputs_legendre.restype = None
    lib.orc_hybrid_phase_function.argtypes = [C.c_int, _fp, _fp, C.c_float, _fp]
    lib.orc_hybrid_phase_function.restype = C.c_int
    lib.orc_finalise_stats.argtypes = [_dp, C.c_int64, C.c_double, C.c_int64, C.c_int64]
    lib.orc_finalise_stats.restype = None
    lib.orc_march.restype = C.c_float
    lib.orc_march.argtypes = [C.c_void_p, _fp, _dp, _dp, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                              C.c_int, C.c_float, _dp, C.POINTER(C.c_int64)]
    _lib = lib
    return lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class OracleDomain:
    """Wraps ``orc_domain``; built from the host-side ``Domain`` mirror's dense arrays."""

    def __init__(self, d, tableSize=9001, forward=False, hybrid=False, hybridWidth=7.0):
        self.lib = load()
        if d.totalExt is None:
            d.getOpticalPropertiesByComponent()
        self.nx, self.ny, self.nz = d.numX, d.numY, d.numZ
        self.nc = d.cumulativeExt.shape[0]
        self.ptr = self.lib.orc_domain_new(self.nx, self.ny, self.nz, self.nc, _p(d.xPosition, C.c_double),
                                           _p(d.yPosition, C.c_double), _p(d.zPosition, C.c_double),
                                           _p(d.totalExt, C.c_double), _p(d.cumulativeExt, C.c_double),
                                           _p(d.ssa, C.c_double), _p(d.phaseFunctionIndex, C.c_int32),
                                           float(d.surfaceAlbedo))
        d.tabulateInversePhaseFunctions(tableSize)
        for c, T in enumerate(d.inversePhaseFunctions):
            self.lib.orc_domain_set_inverse(self.ptr, c + 1, T.shape[1], T.shape[0], _p(T, C.c_float))
        if forward:
            d.tabulateForwardPhaseFunctions(tableSize, hybrid, hybridWidth)
            for c, (Pf, Po) in enumerate(zip(d.tabulatedPhaseFunctions, d.tabulatedOrigPhaseFunctions)):
                self.lib.orc_domain_set_forward(self.ptr, c + 1, Pf.shape[1], Pf.shape[0], _p(Pf, C.c_float),
                                                _p(Po, C.c_float))

    def emission_weighting(self, temps, lambda_um, sfcTemp):
        cdf = np.zeros((self.nz, self.ny, self.nx), dtype=np.float64)
        flux = C.c_double(0)
        temps = np.ascontiguousarray(temps, dtype=np.float64)
        frac = self.lib.orc_emission_weighting(self.ptr, _p(temps, C.c_double), float(lambda_um), float(sfcTemp),
                                               _p(cdf, C.c_double), C.byref(flux))
        return float(frac), cdf, float(flux.value)

    def __del__(self):
        try:
            self.lib.orc_domain_free(self.ptr)
        except Exception:
            pass


class OracleIntegrator:
    def __init__(self, dom: OracleDomain, **opts):
        self.lib = dom.lib
        self.dom = dom
        self.ptr = self.lib.orc_integrator_new(dom.ptr)
        self.opt = orc_options()
        self.lib.orc_default_options(C.byref(self.opt))
        self.nDir = 0
        self.set_options(**opts)

    def set_options(self, **opts):
        for k, v in opts.items():
            if not hasattr(self.opt, k):
                raise AttributeError(k)
            setattr(self.opt, k, v)
        self.lib.orc_integrator_set_options(self.ptr, C.byref(self.opt))

    def head(self) -> _integrator_head:
        return _integrator_head.from_address(self.ptr)

    def set_views(self, mus, phisDeg):
        mus = np.ascontiguousarray(mus, dtype=np.float32); ph = np.ascontiguousarray(phisDeg, dtype=np.float32)
        self.nDir = mus.size
        self.lib.orc_integrator_set_views(self.ptr, self.nDir, _p(mus, C.c_float), _p(ph, C.c_float))

    def set_view_cosines(self, dirCos):
        dc = np.ascontiguousarray(dirCos, dtype=np.float32)
        self.nDir = dc.shape[0]
        self.lib.orc_integrator_set_view_cosines(self.ptr, self.nDir, _p(dc, C.c_float))

    def view_cosines(self):
        h = self.head()
        return np.ctypeslib.as_array(h.intensityDirections, shape=(self.nDir, 3)).copy()

    def raw_tallies(self):
        """Current tallies in the packed order of the CUDA library's tally buffer (as f64)."""
        h = self.head()
        cols = self.dom.nx * self.dom.ny
        parts = [np.ctypeslib.as_array(h.fluxUp, shape=(cols,)), np.ctypeslib.as_array(h.fluxDown, shape=(cols,)),
                 np.ctypeslib.as_array(h.fluxAbsorbed, shape=(cols,)),
                 np.ctypeslib.as_array(h.volumeAbsorption, shape=(cols * self.dom.nz,))]
        if self.nDir:
            parts.append(np.ctypeslib.as_array(h.intensity, shape=(cols * self.nDir,)))
            parts.append(np.ctypeslib.as_array(h.intensityByComponent, shape=(cols * self.nDir * (self.dom.nc + 1),)))
            parts.append(np.ctypeslib.as_array(h.intensityExcess, shape=(self.nDir * (self.dom.nc + 1),)))
        return np.concatenate([p.astype(np.float64) for p in parts])

    def counters(self):
        c = self.head().cnt
        return {n: int(getattr(c, n)) for n, _ in orc_counters._fields_}

    def trace(self, randomReals, source=0, solarMu=1.0, solarAzimuth=0.0, fracAtmsPower=0.0, voxelCDF=None,
              maxEvents=None):
        rn = np.ascontiguousarray(randomReals, dtype=np.float32)
        n, stride = rn.shape
        cap = int(maxEvents or n * 256)
        ev = np.zeros(cap, dtype=EVENT_DTYPE)
        self.lib.orc_integrator_set_trace(self.ptr, ev.ctypes.data_as(C.c_void_p), cap)
        self.lib.orc_trace_photons(self.ptr, self.dom.ptr, int(source), float(solarMu), float(solarAzimuth),
                                   float(fracAtmsPower), _p(voxelCDF, C.c_double), n, _p(rn, C.c_float), stride)
        nEv = int(self.head().traceN)
        self.lib.orc_integrator_set_trace(self.ptr, None, 0)
        return ev[: min(nEv, cap)]

    def photon_fingerprints(self, numBatches, photonsPerBatch, solarMu, solarAzimuth, iseed=10, rank=1):
        """What oracle/ref_build/ref_trace_driver.f90 records for the Fortran reference: the reference's batch loop
        (new_PhotonStream + computeRadiativeTransfer + reportResults, DRV:956-1011) on ONE MT19937 stream seeded
        (iseed, rank, 0), and after every batch every non-zero entry of fluxUp / fluxDown / fluxAbsorbed /
        volumeAbsorption / intensity as (batch, photons processed, array id 1..5, 1-based Fortran index, value)."""
        nx, ny, nz = self.dom.nx, self.dom.ny, self.dom.nz
        cols = nx * ny
        r = orc_rng()
        key = (C.c_uint32 * 3)(int(iseed), int(rank), 0)
        self.lib.orc_rng_init_array(C.byref(r), key, 3)
        arrs = [np.zeros(cols, np.float32), np.zeros(cols, np.float32), np.zeros(cols, np.float32),
                np.zeros(cols * nz, np.float32), np.zeros(cols * max(self.nDir, 1), np.float32)]
        out = [[], [], [], [], []]
        for b in range(1, int(numBatches) + 1):
            ph = self.lib.orc_photons_directional(float(solarMu), float(solarAzimuth), int(photonsPerBatch), C.byref(r))
            done = C.c_int64(0)
            self.lib.orc_compute_radiative_transfer(self.ptr, self.dom.ptr, C.byref(r), ph, int(photonsPerBatch), 1, C.byref(done))
            self.lib.orc_photons_free(ph)
            self.lib.orc_report_results(self.ptr, None, None, None, _p(arrs[0], C.c_float), _p(arrs[1], C.c_float),
                                        _p(arrs[2], C.c_float), None, _p(arrs[3], C.c_float), None,
                                        _p(arrs[4], C.c_float) if self.nDir else None)
            n0 = len(out[0])
            for a, arr in enumerate(arrs if self.nDir else arrs[:4]):
                nzi = np.nonzero(arr)[0]
                out[0] += [b] * nzi.size; out[1] += [int(done.value)] * nzi.size
                out[2].append(np.full(nzi.size, a + 1, np.int32)); out[3].append((nzi + 1).astype(np.int32)); out[4].append(arr[nzi].copy())
            if len(out[0]) == n0:                       # a photon that left no trace still shows up as a batch
                out[0].append(b); out[1].append(int(done.value))
                out[2].append(np.zeros(1, np.int32)); out[3].append(np.zeros(1, np.int32)); out[4].append(np.zeros(1, np.float32))
        return (np.array(out[0], np.int32), np.array(out[1], np.int32), np.concatenate(out[2]), np.concatenate(out[3]),
                np.concatenate(out[4]))

    def run_batches(self, numBatches, numPhotonsPerBatch, source=0, solarMu=1.0, solarAzimuth=0.0,
                    fracAtmsPower=0.0, voxelCDF=None, iseed=10, rank=1, thread=0, volume=False):
        """DRV:949-1052 for one worker; returns (totalPhotons, dict of un-finalised moment arrays)."""
        nx, ny, nz = self.dom.nx, self.dom.ny, self.dom.nz
        cols = nx * ny
        arrs = dict(meanFluxUpStats=np.zeros(2), meanFluxDownStats=np.zeros(2), meanFluxAbsorbedStats=np.zeros(2),
                    fluxUpStats=np.zeros(cols * 2), fluxDownStats=np.zeros(cols * 2),
                    fluxAbsorbedStats=np.zeros(cols * 2), absorbedProfileStats=np.zeros(nz * 2),
                    absorbedVolumeStats=np.zeros(cols * nz * 2) if volume else None,
                    radianceStats=np.zeros(cols * self.nDir * 2) if self.nDir else None)
        st = orc_stats(**{k: _p(v, C.c_double) for k, v in arrs.items()})
        total = self.lib.orc_run_batches(self.ptr, self.dom.ptr, int(source), float(solarMu), float(solarAzimuth),
                                         float(fracAtmsPower), _p(voxelCDF, C.c_double), int(iseed), int(rank),
                                         int(thread), int(numBatches), int(numPhotonsPerBatch), C.byref(st))
        return int(total), {k: v for k, v in arrs.items() if v is not None}

    def __del__(self):
        try:
            self.lib.orc_integrator_free(self.ptr)
        except Exception:
            pass


def assemble_optics(nx, ny, nz, massConc, Reff, numConc, comps, setup=False):
    """read_SSPTable's loops + getOpticalPropertiesByComponent (OPT:204-299, 1022-1061).  ``comps`` is a list of
    dicts with kind / physIndex / zLevelBase / ext and, by kind, key / ssa / idx (the reference's per-lambda reads).
    Returns (rc, totalExt, cumExt, ssa, phaseIdx) in the domain's layout."""
    lib = load()
    mc = np.ascontiguousarray(massConc, dtype=np.float64); re = np.ascontiguousarray(Reff, dtype=np.float64)
    nPhys = mc.shape[-1] if mc.ndim == 4 else 0
    nco = None if numConc is None else np.ascontiguousarray(numConc, dtype=np.float64)
    arr = (orc_component * len(comps))()
    keep = []
    for i, q in enumerate(comps):
        ext = np.ascontiguousarray(q["ext"], dtype=np.float64); keep.append(ext)
        arr[i].kind = q["kind"]; arr[i].physIndex = q.get("physIndex", 0); arr[i].zLevelBase = q.get("zLevelBase", 1)
        arr[i].nTable = ext.size; arr[i].ext = _p(ext, C.c_double)
        if "ssa" in q:
            a = np.ascontiguousarray(q["ssa"], dtype=np.float64); keep.append(a); arr[i].ssa = _p(a, C.c_double)
        if "key" in q:
            a = np.ascontiguousarray(q["key"], dtype=np.float32); keep.append(a); arr[i].key = _p(a, C.c_float)
        if "idx" in q:
            a = np.ascontiguousarray(q["idx"], dtype=np.int32); keep.append(a); arr[i].phaseIdx = _p(a, C.c_int32)
    nc = len(comps)
    total = np.empty((nz, ny, nx)); cum = np.empty((nc, nz, ny, nx)); ssa = np.empty((nc, nz, ny, nx))
    idx = np.empty((nc, nz, ny, nx), np.int32)
    rc = lib.orc_assemble_optics(nx, ny, nz, nPhys, _p(mc, C.c_double), _p(re, C.c_double), _p(nco, C.c_double), nc, arr,
                                 int(bool(setup)), _p(total, C.c_double), _p(cum, C.c_double), _p(ssa, C.c_double),
                                 idx.ctypes.data_as(C.POINTER(C.c_int32)))
    return rc, total, cum, ssa, idx


def frequency_distribution(CDF, totalPhotons, seed=(10, 0, 0)):
    """getFrequencyDistr (EMI:552-573) with the reference's MT19937 seeded by init_by_array(seed)."""
    lib = load()
    cdf = np.ascontiguousarray(CDF, dtype=np.float64)
    r = orc_rng()
    key = (C.c_uint32 * len(seed))(*[int(s) for s in seed])
    lib.orc_rng_init_array(C.byref(r), key, len(seed))
    out = np.zeros(cdf.size, dtype=np.int64)
    lib.orc_frequency_distribution(cdf.size, _p(cdf, C.c_double), int(totalPhotons), C.byref(r),
                                   out.ctypes.data_as(C.POINTER(C.c_int64)))
    return out


def inverse_phase_function(mus, values, nSteps):
    """computeInversePhaseFunction INV:113-168 on (mus, values) increasing in mu."""
    lib = load()
    m = np.ascontiguousarray(mus, dtype=np.float32); v = np.ascontiguousarray(values, dtype=np.float32)
    out = np.empty(int(nSteps), dtype=np.float32)
    lib.orc_inverse_phase_function(m.size, _p(m, C.c_float), _p(v, C.c_float), int(nSteps), _p(out, C.c_float))
    return out


def forward_phase_function(legendreCoefficients, nSteps):
    """tabulateForwardPhaseFunctions OPT:1912-1913 for one Legendre-stored phase function."""
    lib = load()
    c = np.ascontiguousarray(legendreCoefficients, dtype=np.float32)
    out = np.empty(int(nSteps), dtype=np.float32)
    lib.orc_forward_phase_function(c.size, _p(c, C.c_float), int(nSteps), _p(out, C.c_float))
    return out


def lobatto_terms(n):
    """computeLobattoTerms NUM:27-114: (mus, weights)."""
    mus = np.empty(int(n), dtype=np.float32); w = np.empty(int(n), dtype=np.float32)
    load().orc_lobatto_terms(int(n), _p(mus, C.c_float), _p(w, C.c_float))
    return mus, w


def phase_function_values(angles, legendreCoefficients=None, storedAngle=None, storedValue=None):
    """getPhaseFunctionValues_one SPF:448-531 for a Legendre-stored or an angle / value phase function."""
    a = np.ascontiguousarray(angles, dtype=np.float32)
    out = np.empty(a.size, dtype=np.float32)
    if storedAngle is None:
        c = np.ascontiguousarray(legendreCoefficients, dtype=np.float32)
        load().orc_phase_function_values(c.size, _p(c, C.c_float), 0, None, None, a.size, _p(a, C.c_float), _p(out, C.c_float))
    else:
        sa = np.ascontiguousarray(storedAngle, dtype=np.float32); sv = np.ascontiguousarray(storedValue, dtype=np.float32)
        load().orc_phase_function_values(0, None, sa.size, _p(sa, C.c_float), _p(sv, C.c_float), a.size, _p(a, C.c_float),
                                         _p(out, C.c_float))
    return out


def inversion_inputs_legendre(legendreCoefficients):
    """INV:97-112: the Lobatto abscissas and the phase function values there that computeInversePhaseFunction inverts."""
    c = np.ascontiguousarray(legendreCoefficients, dtype=np.float32)
    n = max(c.size, 2)
    mus = np.empty(n, dtype=np.float32); v = np.empty(n, dtype=np.float32)
    load().orc_inversion_inputs_legendre(c.size, _p(c, C.c_float), _p(mus, C.c_float), _p(v, C.c_float))
    return mus, v


def hybrid_phase_function(angles, values, gaussianWidth):
    """computeHybridPhaseFunctions OPT:1936-2050 for one entry: (new values, transition index)."""
    a = np.ascontiguousarray(angles, dtype=np.float32); v = np.ascontiguousarray(values, dtype=np.float32)
    out = np.empty(a.size, dtype=np.float32)
    t = load().orc_hybrid_phase_function(a.size, _p(a, C.c_float), _p(v, C.c_float), float(gaussianWidth), _p(out, C.c_float))
    return out, int(t)


def finalise(stats: np.ndarray, solarFlux: float, totalNumPhotons: int, batchesCompleted: int):
    """DRV:1188-1228: (first moments | second moments) -> (means, standard errors)."""
    s = np.ascontiguousarray(stats, dtype=np.float64).copy()
    n = s.size // 2
    load().orc_finalise_stats(_p(s, C.c_double), n, float(solarFlux), int(totalNumPhotons), int(batchesCompleted))
    return s[:n], s[n:]


def run_workers(make_integrator, workers: int, numBatches: int, numPhotonsPerBatch: int, **kw):
    """The reference's 1 master + W workers layout (DRV:665-1095): W independent workers, each an
    MT19937 stream seeded (iseed, rank, 0), moments summed as sumAcrossProcesses does (DRV:1151-1166).
    ctypes releases the GIL, so the workers run on W host cores."""
    def one(rank):
        g = make_integrator()
        per = numBatches // workers + (1 if rank <= numBatches % workers else 0)
        return g.run_batches(per, numPhotonsPerBatch, rank=rank, **kw) + (per,)
    with ThreadPoolExecutor(max_workers=workers) as ex:
        results = list(ex.map(one, range(1, workers + 1)))
    total = sum(r[0] for r in results)
    batches = sum(r[2] for r in results)
    summed = {k: sum(r[1][k] for r in results) for k in results[0][1]}
    return total, batches, summed
