"""Host-side mirror of ``src/opticalProperties.f95`` -- the ``domain`` object.

This is staging code: it assembles the dense per-cell arrays the photon kernels read
(``getOpticalPropertiesByComponent`` OPT:966-1072) and the tabulated phase-function
matrices (``tabulateInversePhaseFunctions`` OPT:1817-1870, ``tabulateForwardPhaseFunctions``
OPT:1872-1934, hybrid OPT:1936-2050).  The ray marcher ``accumulateExtinctionAlongPath``
(OPT:1656-1815) lives on the device (``csrc/mcb_reference.cu`` / ``csrc/mcb_fast.cu``).

Arrays keep the reference's Fortran layout: ``a[ix + nx*(iy + ny*iz)]`` (x fastest); they
are held as NumPy arrays of shape ``(nz, ny, nx)`` / ``(nc, nz, ny, nx)`` so that the C order
of the buffer IS the Fortran order.  File I/O (netCDF readers/writers) is out of scope.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .inversePhaseFunctions import computeInversePhaseFuncTable
from .numericUtilities import findIndex, spacing64
from .scatteringPhaseFunctions import getPhaseFunctionValues, phaseFunctionTable

f32 = np.float32
Pi = f32(3.14159265358979312)          # OPT:26

# Staging caches (what is already in HBM) are keyed on these tokens, never on id(): a token is unique for the life of
# the process and is replaced whenever the object's staged content changes.
_tokens = itertools.count(1)


def new_token() -> int:
    return next(_tokens)


@dataclass
class opticalComponent:
    """OPT:42-60."""
    name: str
    zLevelBase: int
    horizontallyUniform: bool
    extinction: np.ndarray                 # (nzc, ny, nx) or (nzc,) when uniform
    singleScatteringAlbedo: np.ndarray
    phaseFunctionIndex: np.ndarray
    table: phaseFunctionTable


class Domain:
    """``type(domain)`` (OPT:77-111)."""

    def __init__(self, xPosition, yPosition, zPosition, temps=None, surfaceAlbedo=0.0,
                 lambda_um=0.0):
        self.xPosition = np.ascontiguousarray(xPosition, dtype=np.float64)
        self.yPosition = np.ascontiguousarray(yPosition, dtype=np.float64)
        self.zPosition = np.ascontiguousarray(zPosition, dtype=np.float64)
        for p in (self.xPosition, self.yPosition, self.zPosition):
            if np.any(np.diff(p) <= 0):
                raise ValueError("new_Domain: Positions must be increasing, unique.")
        nx, ny, nz = self.numX, self.numY, self.numZ
        self.temps = (np.zeros((nz, ny, nx)) if temps is None
                      else np.ascontiguousarray(temps, dtype=np.float64).reshape(nz, ny, nx))
        self.surfaceAlbedo = float(surfaceAlbedo)
        self.lambda_um = float(lambda_um)
        self.token = new_token()           # identifies the dense optical arrays (replaced when they change)
        self.tableToken = new_token()      # identifies the tabulated phase functions

        def regular(p, tol):
            return bool(np.all(np.abs(np.diff(p) - (p[1] - p[0])) <= tol * spacing64(p[1:])))
        self.xyRegularlySpaced = regular(self.xPosition, 2) and regular(self.yPosition, 2)   # OPT:541-545
        self.zRegularlySpaced = regular(self.zPosition, 2)
        self.components: List[opticalComponent] = []
        self.deviceOwner = None            # integrator whose HBM holds the dense arrays (read_SSPTable on the device)
        self.totalExt: Optional[np.ndarray] = None
        self.cumulativeExt: Optional[np.ndarray] = None
        self.ssa: Optional[np.ndarray] = None
        self.phaseFunctionIndex: Optional[np.ndarray] = None
        self.forwardTables: List[phaseFunctionTable] = []
        self.inversePhaseFunctions: List[Optional[np.ndarray]] = []
        self.tabulatedPhaseFunctions: List[Optional[np.ndarray]] = []
        self.tabulatedOrigPhaseFunctions: List[Optional[np.ndarray]] = []

    # -- getInfo_Domain (OPT:796-962) ---------------------------------------------------
    @property
    def numX(self): return self.xPosition.size - 1
    @property
    def numY(self): return self.yPosition.size - 1
    @property
    def numZ(self): return self.zPosition.size - 1
    @property
    def numberOfComponents(self): return len(self.forwardTables) or len(self.components)

    def getInfo_Domain(self):
        """The subset of ``getInfo_Domain`` the integrator asks for (INT:434-443, 1668-1673)."""
        return dict(numX=self.numX, numY=self.numY, numZ=self.numZ, albedo=self.surfaceAlbedo,
                    numberOfComponents=self.numberOfComponents,
                    xPosition=self.xPosition, yPosition=self.yPosition, zPosition=self.zPosition,
                    temps=self.temps, totalExt=self.totalExt, cumExt=self.cumulativeExt,
                    ssa=self.ssa, phaseFuncI=self.phaseFunctionIndex,
                    inversePhaseFuncs=self.inversePhaseFunctions,
                    tabPhase=self.tabulatedPhaseFunctions,
                    tabOrigPhase=self.tabulatedOrigPhaseFunctions)

    # -- addOpticalComponent (OPT:557-665) ----------------------------------------------
    def addOpticalComponent(self, componentName, extinction, singleScatteringAlbedo,
                            phaseFunctionIndex, phaseFunctions: phaseFunctionTable, zLevelBase=1):
        ext = np.asarray(extinction, dtype=np.float64)
        ssa = np.asarray(singleScatteringAlbedo, dtype=np.float64)
        idx = np.asarray(phaseFunctionIndex, dtype=np.int32)
        uniform = ext.ndim == 1
        nzc = ext.shape[0]
        if ssa.shape != ext.shape or idx.shape != ext.shape:
            raise ValueError("addOpticalComponent: optical property grids must be the same size.")
        if not uniform and ext.shape[1:] != (self.numY, self.numX):
            raise ValueError("addOpticalComponent: arrays don't span the horizontal extent of the domain.")
        if zLevelBase < 1 or zLevelBase + nzc - 1 > self.numZ:
            raise ValueError("addOpticalComponent: arrays don't fit the vertical extent of the domain.")
        if np.any(ext < 0):
            raise ValueError("addOpticalComponent: extinction must be >= 0.")
        if np.any(ssa < 0) or np.any(ssa > 1):
            raise ValueError("addOpticalComponent: singleScatteringAlbedo must be between 0 and 1")
        if np.any(idx < 0) or np.any(idx > phaseFunctions.nEntries):
            raise ValueError("addOpticalComponent: phase function index is out of bounds")
        self.components.append(opticalComponent(componentName, int(zLevelBase), uniform, ext, ssa, idx,
                                                phaseFunctions))
        self.totalExt = None               # must be re-assembled
        self.token = new_token()

    # -- getOpticalPropertiesByComponent (OPT:966-1072) ---------------------------------
    def getOpticalPropertiesByComponent(self):
        if not self.components:
            raise ValueError("getOpticalPropertiesByComponent: domain contains no optical components.")
        nx, ny, nz, nc = self.numX, self.numY, self.numZ, len(self.components)
        cum = np.zeros((nc, nz, ny, nx), dtype=np.float64)
        ssa = np.zeros((nc, nz, ny, nx), dtype=np.float64)
        idx = np.zeros((nc, nz, ny, nx), dtype=np.int32)
        self.forwardTables = []
        for i, c in enumerate(self.components):
            lo = c.zLevelBase - 1
            hi = lo + c.extinction.shape[0]
            if c.horizontallyUniform:
                cum[i, lo:hi] = c.extinction[:, None, None]
                ssa[i, lo:hi] = c.singleScatteringAlbedo[:, None, None]
                idx[i, lo:hi] = c.phaseFunctionIndex[:, None, None]
            else:
                cum[i, lo:hi] = c.extinction
                ssa[i, lo:hi] = c.singleScatteringAlbedo
                idx[i, lo:hi] = c.phaseFunctionIndex
            self.forwardTables.append(c.table)
        for i in range(1, nc):                                           # OPT:1055-1057
            cum[i] = cum[i] + cum[i - 1]
        total = cum[nc - 1].copy()
        mask = total > np.finfo(np.float64).tiny                         # OPT:1059-1061
        for i in range(nc):
            np.divide(cum[i], total, out=cum[i], where=mask)
        self.totalExt, self.cumulativeExt, self.ssa, self.phaseFunctionIndex = total, cum, ssa, idx
        self.token = new_token(); self.tableToken = new_token()
        self.inversePhaseFunctions = [None] * nc
        self.tabulatedPhaseFunctions = [None] * nc
        self.tabulatedOrigPhaseFunctions = [None] * nc
        return self

    # -- tabulateInversePhaseFunctions (OPT:1817-1870) -----------------------------------
    def tabulateInversePhaseFunctions(self, tableSize: int):
        for i, tab in enumerate(self.forwardTables):
            cur = self.inversePhaseFunctions[i]
            if cur is not None and cur.shape[1] >= tableSize:
                continue
            self.inversePhaseFunctions[i] = computeInversePhaseFuncTable(tab, tableSize)
            self.tableToken = new_token()

    def touch(self):
        """Call after editing ``totalExt`` / ``cumulativeExt`` / ``ssa`` / ``phaseFunctionIndex`` or ``surfaceAlbedo`` in
        place: the next ``computeRadiativeTransfer`` stages the arrays again."""
        self.token = new_token()

    # -- tabulateForwardPhaseFunctions (OPT:1872-1934) -----------------------------------
    def tabulateForwardPhaseFunctions(self, tableSize: int, hybrid: bool = False, hybridWidth: float = 7.0):
        for i, tab in enumerate(self.forwardTables):
            cur = self.tabulatedPhaseFunctions[i]
            if cur is not None and cur.shape[1] >= tableSize:
                continue
            nSteps = tableSize
            angles = (np.arange(nSteps).astype(f32) / f32(nSteps - 1) * Pi).astype(f32)     # OPT:1912
            values = getPhaseFunctionValues(tab, angles)                 # (nSteps, nEntries)
            orig = np.ascontiguousarray(values.T, dtype=f32)             # (nEntries, nSteps)
            self.tabulatedOrigPhaseFunctions[i] = orig
            self.tableToken = new_token()
            if hybrid and hybridWidth > 0:
                self.tabulatedPhaseFunctions[i] = computeHybridPhaseFunctions(angles, orig, f32(hybridWidth))
            else:
                self.tabulatedPhaseFunctions[i] = orig.copy()


# ---------------------------------------------------------------------------------------
# computeHybridPhaseFunctions (OPT:1936-2050)
# ---------------------------------------------------------------------------------------
def _dot32(a, b):
    """dot_product in single precision, accumulated in order (a compiler that keeps IEEE semantics does not
    reassociate the reduction)."""
    prod = (np.asarray(a, dtype=f32) * np.asarray(b, dtype=f32)).astype(f32)
    return f32(np.add.accumulate(prod, dtype=f32)[-1]) if prod.size else f32(0.0)


def _computeNormalization(angleCosines, values, gaussianValues, t):     # OPT:2027-2050 (t is 1-based)
    n = angleCosines.size
    ig = _dot32(f32(0.5) * (gaussianValues[0:t - 1] + gaussianValues[1:t]),
                angleCosines[0:t - 1] - angleCosines[1:t])
    io = _dot32(f32(0.5) * (values[t - 1:n - 1] + values[t:n]),
                angleCosines[t - 1:n - 1] - angleCosines[t:n])
    if io >= f32(2.0):
        return f32(1.0) / ig
    return (f32(2.0) - io) / ig


def _phaseFuncDiff(angleCosines, values, gaussianValues, t):            # OPT:2011-2025
    P0 = _computeNormalization(angleCosines, values, gaussianValues, t)
    return f32(P0 * gaussianValues[t - 1] - values[t - 1])


def computeHybridPhaseFunctions(angles, values, GaussianWidth):
    """OPT:1936-2009.  ``values`` has shape (nEntries, nAngles); a Gaussian of the given
    width (degrees) replaces the forward peak, continuous with the original."""
    angles = np.asarray(angles, dtype=f32)
    nAngles = angles.size
    angleCosines = np.cos(angles.astype(np.float64)).astype(f32)
    w = f32(GaussianWidth * Pi / f32(180))
    gaussianValues = np.exp(-((angles / w).astype(f32) ** 2).astype(np.float64)).astype(f32)
    newValues = np.array(values, dtype=f32, copy=True)
    for e in range(values.shape[0]):
        v = values[e]
        lower = findIndex(w, angles) + 1
        if lower >= nAngles - 2:
            break
        lowDiff = _phaseFuncDiff(angleCosines, v, gaussianValues, lower)
        inc = 1
        noRoot = False
        while True:
            upper = min(lower + inc, nAngles - 1)
            upDiff = _phaseFuncDiff(angleCosines, v, gaussianValues, upper)
            if lower == nAngles - 1:
                noRoot = True
                break
            if lowDiff * upDiff < 0:
                break
            lower = upper
            lowDiff = upDiff
            inc *= 2
        if noRoot:
            continue
        while upper > lower + 1:
            mid = (lower + upper) // 2
            midDiff = _phaseFuncDiff(angleCosines, v, gaussianValues, mid)
            if midDiff * upDiff < 0:
                lower, lowDiff = mid, midDiff
            else:
                upper, upDiff = mid, midDiff
        t = lower
        P0 = _computeNormalization(angleCosines, v, gaussianValues, t)
        newValues[e, :t] = P0 * gaussianValues[:t]
        newValues[e, t:] = v[t:]
    return newValues


# ---------------------------------------------------------------------------------------
# type(commonDomain) (OPT:63-75) and read_SSPTable (OPT:147-343)
# ---------------------------------------------------------------------------------------
light_spd = 2.99792458E8                   # OPT:27 [m/s]


@dataclass
class commonDomain:
    """OPT:63-75: what every wavelength shares.  ``massConc``/``Reff`` are (nz, ny, nx, nPhys) so that the C order
    of the buffer is the reference's ``(component, x, y, z)``; ``rho``/``numConc`` are (nz, ny, nx) of which the
    reference only ever reads column (1,1,:) (OPT:223, 324)."""
    xPosition: np.ndarray
    yPosition: np.ndarray
    zPosition: np.ndarray
    temps: np.ndarray
    massConc: np.ndarray
    Reff: np.ndarray
    numConc: Optional[np.ndarray] = None
    rho: Optional[np.ndarray] = None
    token: int = field(default_factory=new_token)      # staging-cache key (replace after editing the arrays in place)


@dataclass
class SSPComponent:
    """One ``ComponentN_*`` group of a single-scattering-property file (OPT:200-246)."""
    name: str
    extType: str                              # "volExt" | "absXsec"
    zLevelBase: int = 1
    key: Optional[np.ndarray] = None          # phaseFunctionKeyT(nReff), default real
    extinctionT: Optional[np.ndarray] = None  # (nLambda, nReff)
    singleScatteringAlbedoT: Optional[np.ndarray] = None
    xsec: Optional[np.ndarray] = None         # (nLambda, nz)
    tables: Optional[List[phaseFunctionTable]] = None   # per lambda: read_PhaseFunctionTable (OPT:257)


@dataclass
class SSPTable:
    """In-memory image of one SSP file (the netCDF container itself is out of scope)."""
    f_grid: np.ndarray                        # frequencies [Hz], one per lambda index
    surfaceAlbedo: np.ndarray
    components: List[SSPComponent]


def calc_RayleighScattering(lambda_um, rho, N):
    """OPT:2052-2086: returns (ext, ssa, phaseInd, table) for one column."""
    from .scatteringPhaseFunctions import new_PhaseFunctionTable, rayleigh
    f = 1.060816681; rho0 = 1.275
    lam = float(lambda_um)
    Pi8 = np.float64(Pi)                                        # the module's Pi is default real (OPT:26)
    mr1 = 6.4328E-5 + (2.94981E-2 / (146 - (lam ** (-2)))) + (2.554E-4 / (41 - (lam ** (-2))))
    rho = np.asarray(rho, dtype=np.float64); N = np.asarray(N, dtype=np.float64)
    ext = (32.0E27) * f * (Pi8 ** 3) * (rho ** 2) * (mr1 ** 2) / (3.0 * N * (rho0 ** 2) * (lam ** 4))
    return ext, np.ones_like(ext), np.ones(ext.size, np.int32), new_PhaseFunctionTable([rayleigh()], key=[0.0])


def _null_table():
    from .scatteringPhaseFunctions import new_PhaseFunction, new_PhaseFunctionTable
    return new_PhaseFunctionTable([new_PhaseFunction(legendreCoefficients=np.array([0.0, 0.0], dtype=f32))], key=[0.0])


def read_SSPTable(tables: List[SSPTable], lambdaIndex: int, commonD: commonDomain, setup: bool = False,
                  calcRayl: bool = False, thisIntegrator=None) -> "Domain":
    """``read_SSPTable`` (OPT:147-343) for one wavelength (``lambdaIndex`` is 1-based): cloud / aerosol components are
    interpolated in effective radius from the file's tables, gas components are ``xsec * numConc * 1000``, Rayleigh
    scattering is added on request, then ``getOpticalPropertiesByComponent`` assembles the dense arrays.

    Without ``thisIntegrator`` this is the NumPy staging producer.  With it, the per-cell loops run on that
    integrator's GPU (``mcb_set_physical`` once per run, ``mcb_assemble_optics`` per wavelength): the returned Domain
    carries the component tables and metadata, its dense arrays live in HBM only."""
    li = int(lambdaIndex) - 1
    first = tables[0]
    lam = (light_spd * (10 ** 6)) / float(first.f_grid[li])                      # OPT:199 [microns]
    d = Domain(commonD.xPosition, commonD.yPosition, commonD.zPosition, temps=commonD.temps,
               surfaceAlbedo=float(first.surfaceAlbedo[li]), lambda_um=lam)
    nz, ny, nx = d.numZ, d.numY, d.numX
    descr = []                                                                    # for the device path
    comp = 1; gasComp = 0
    for t in tables:
        for c in t.components:
            if c.extType == "absXsec":                                            # OPT:203-221
                gasComp += 1
                ext = np.asarray(c.xsec[li], dtype=np.float64) * commonD.numConc[:, 0, 0] * 1000.0
                if thisIntegrator is None:
                    d.addOpticalComponent(c.name, ext, np.zeros(nz), np.ones(nz, np.int32), _null_table(),
                                          zLevelBase=c.zLevelBase)
                descr.append(dict(kind=1, physIndex=0, zLevelBase=c.zLevelBase, table=_null_table(),
                                  ext=np.ascontiguousarray(c.xsec[li], dtype=np.float64)))
            elif c.extType == "volExt":                                           # OPT:222-293
                key = np.asarray(c.key, dtype=f32)
                extT = np.ascontiguousarray(c.extinctionT[li], dtype=np.float64)
                ssaT = np.ascontiguousarray(c.singleScatteringAlbedoT[li], dtype=np.float64)
                table = _null_table() if setup else c.tables[li]
                p = comp - gasComp - 1
                if thisIntegrator is None:
                    m = commonD.massConc[..., p]; re = commonD.Reff[..., p]
                    inside = (m > 0.0) & (re < np.float64(key.max())) & (re >= np.float64(key.min()))
                    if np.any((m > 0.0) & ~inside):
                        raise ValueError("read_SSPTable: Effective radius outside of table range")
                    key8 = key.astype(np.float64)
                    il = np.clip(np.searchsorted(key8, re, side="right"), 1, key.size - 1)      # findIndex: key(il) <= Reff < key(il+1)
                    f = (re - key8[il - 1]) / (key[il] - key[il - 1]).astype(np.float64)         # OPT:272 (f32 difference)
                    ext = np.where(inside, m * ((1 - f) * extT[il - 1] + f * extT[il]), 0.0)
                    ssa = np.where(inside, (1 - f) * ssaT[il - 1] + f * ssaT[il], 0.0)
                    idx = np.ones((nz, ny, nx), np.int32)
                    if not setup:
                        idx = np.where(inside, np.where(f < 0.5, il, il + 1), 1).astype(np.int32)
                    d.addOpticalComponent(c.name, ext, ssa, idx, table, zLevelBase=c.zLevelBase)
                descr.append(dict(kind=0, physIndex=p + 1, zLevelBase=c.zLevelBase, table=table, key=key, ext=extT, ssa=ssaT))
            else:
                raise ValueError("read_SSPTable: unrecognizable extType")
            comp += 1
    if calcRayl and not setup:                                                    # OPT:320-338
        ext, ssa, idx, table = calc_RayleighScattering(lam, commonD.rho[:, 0, 0], commonD.numConc[:, 0, 0])
        if thisIntegrator is None:
            d.addOpticalComponent("Rayleigh Scattering", ext, ssa, idx, table, zLevelBase=1)
        descr.append(dict(kind=2, physIndex=0, zLevelBase=1, table=table, ext=np.ascontiguousarray(ext),
                          ssa=np.ascontiguousarray(ssa), idx=np.ascontiguousarray(idx, dtype=np.int32)))
    if thisIntegrator is None:
        return d.getOpticalPropertiesByComponent()
    _assemble_on_device(thisIntegrator, d, commonD, descr, setup)
    return d


def _assemble_on_device(g, d: "Domain", commonD: commonDomain, descr, setup: bool) -> None:
    import ctypes as C

    from . import _lib
    if (d.numX, d.numY, d.numZ) != (g.numX, g.numY, g.numZ):
        raise ValueError("read_SSPTable: domain and integrator grids differ")
    if getattr(g, "_stagedPhysical", None) != commonD.token:                      # once per run
        mc = np.ascontiguousarray(commonD.massConc, dtype=np.float64)
        re = np.ascontiguousarray(commonD.Reff, dtype=np.float64)
        nPhys = mc.shape[-1] if mc.ndim == 4 else 0
        nconc = None if commonD.numConc is None else np.ascontiguousarray(commonD.numConc[:, 0, 0], dtype=np.float64)
        g._check(g._lib.mcb_set_physical(g.handle, nPhys, _lib.ptr(mc, C.c_double), _lib.ptr(re, C.c_double),
                                         _lib.ptr(nconc, C.c_double)), "read_SSPTable")
        g._stagedPhysical = commonD.token
    arr = (_lib.mcb_component * len(descr))()
    keep = []
    for i, q in enumerate(descr):
        arr[i].kind = q["kind"]; arr[i].physIndex = q["physIndex"]; arr[i].zLevelBase = q["zLevelBase"]
        arr[i].nTable = q["ext"].size
        arr[i].ext = _lib.ptr(q["ext"], C.c_double)
        if "ssa" in q: arr[i].ssa = _lib.ptr(q["ssa"], C.c_double)
        if "key" in q: arr[i].key = _lib.ptr(q["key"], C.c_float)
        if "idx" in q: arr[i].phaseIdx = _lib.ptr(q["idx"], C.c_int32)
        keep.append(q)
    g._check(g._lib.mcb_assemble_optics(g.handle, len(descr), arr, int(bool(setup)), float(d.surfaceAlbedo)), "read_SSPTable")
    nc = len(descr)
    d.forwardTables = [q["table"] for q in descr]
    d.inversePhaseFunctions = [None] * nc
    d.tabulatedPhaseFunctions = [None] * nc
    d.tabulatedOrigPhaseFunctions = [None] * nc
    d.deviceOwner = g
    d.token = new_token(); d.tableToken = new_token()
    g.numComps = nc
    g._stagedDomain = ("device", d.token)
    g._stagedTables = None


def fetchOpticalProperties(d: "Domain") -> "Domain":
    """Copy the dense arrays of a device-assembled Domain back to the host (tests, output)."""
    import ctypes as C

    from . import _lib
    g = getattr(d, "deviceOwner", None)
    if g is None:
        return d
    nc = len(d.forwardTables); nz, ny, nx = d.numZ, d.numY, d.numX
    d.totalExt = np.empty((nz, ny, nx)); d.cumulativeExt = np.empty((nc, nz, ny, nx))
    d.ssa = np.empty((nc, nz, ny, nx)); d.phaseFunctionIndex = np.empty((nc, nz, ny, nx), np.int32)
    g._check(g._lib.mcb_get_optics(g.handle, _lib.ptr(d.totalExt, C.c_double), _lib.ptr(d.cumulativeExt, C.c_double),
                                   _lib.ptr(d.ssa, C.c_double), _lib.ptr(d.phaseFunctionIndex, C.c_int32)),
             "fetchOpticalProperties")
    return d
