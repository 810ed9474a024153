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

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .inversePhaseFunctions import computeInversePhaseFuncTable
from .numericUtilities import findIndex, spacing64
from .scatteringPhaseFunctions import getPhaseFunctionValues, phaseFunctionTable

f32 = np.float32
Pi = f32(3.14159265358979312)          # OPT:26


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

        def regular(p, tol):
            return bool(np.all(np.abs(np.diff(p) - (p[1] - p[0])) <= tol * spacing64(p[1:])))
        self.xyRegularlySpaced = regular(self.xPosition, 2) and regular(self.yPosition, 2)   # OPT:541-545
        self.zRegularlySpaced = regular(self.zPosition, 2)
        self.components: List[opticalComponent] = []
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
            if hybrid and hybridWidth > 0:
                self.tabulatedPhaseFunctions[i] = computeHybridPhaseFunctions(angles, orig, f32(hybridWidth))
            else:
                self.tabulatedPhaseFunctions[i] = orig.copy()


# ---------------------------------------------------------------------------------------
# computeHybridPhaseFunctions (OPT:1936-2050)
# ---------------------------------------------------------------------------------------
def _dot32(a, b):
    return f32(np.dot(np.asarray(a, dtype=f32), np.asarray(b, dtype=f32)))


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
