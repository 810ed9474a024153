"""``writeResults_ASCII`` of ``Drivers/monteCarloDriver.f95`` (DRV:1324-1495), format-exact.

The four ASCII tables the driver writes at the end of a run -- pixel fluxes, the absorption profile, the 3-D
absorption field and the pixel radiances -- with the reference's header lines, loop orders and Fortran edit
descriptors (``F7.3``, ``F9.4``, ``E13.6``, ``A60``, ``I10``, ``L1``) reproduced character for character, including
the row of asterisks a value that does not fit its field is replaced by.  Statistics arrive as the reference holds
them, ``stats(..., 1:2)``: here arrays whose FIRST axis is (mean, standard error) and whose remaining axes are the
C-order view of the Fortran array -- ``(2, ny, nx)``, ``(2, nz)``, ``(2, nz, ny, nx)``, ``(2, nDir, ny, nx)`` --
which is exactly what ``mcb_get_statistics`` / ``batchStatistics.reportStatistics`` deliver.

The netCDF writer (DRV:1498-1807) needs the netCDF library and stays out of scope.
"""
from __future__ import annotations

import math
from decimal import ROUND_HALF_EVEN, Context, Decimal
from typing import Optional, Sequence

import numpy as np


# ---- Fortran edit descriptors ---------------------------------------------------------------------------------
_EXACT = Context(prec=1200)          # a double has at most 767 significant decimal digits: every quantize below is exact


def _F(value, w: int, d: int) -> str:
    """``Fw.d``: fixed notation rounded from the exact binary value (round-half-even on ties, as gfortran and glibc
    do), right-justified; the optional leading zero is dropped when the field is one character short; ``w``
    asterisks when it still does not fit."""
    v = float(value)
    if math.isnan(v):
        return "NaN".rjust(w)
    if math.isinf(v):
        s = "-Infinity" if v < 0 else ("+Infinity" if w >= 9 else "Infinity")
        s = s if len(s) <= w else ("-Inf" if v < 0 else "Inf")
        return s.rjust(w) if len(s) <= w else "*" * w
    q = Decimal(v).quantize(Decimal(1).scaleb(-d), rounding=ROUND_HALF_EVEN, context=_EXACT)
    s = format(abs(q), "f")
    if v < 0 or (v == 0 and math.copysign(1.0, v) < 0):
        s = "-" + s
    if len(s) > w and (s.startswith("0.") or s.startswith("-0.")):
        s = s.replace("0.", ".", 1)
    return s.rjust(w) if len(s) <= w else "*" * w


def _E(value, w: int, d: int) -> str:
    """``Ew.d``: 0.ddddddE+ee (three-digit exponents as +eee without the E)."""
    v = float(value)
    if v == 0.0:
        mant, ex = Decimal(0).quantize(Decimal(1).scaleb(-d)), 0
    else:
        ex = int(math.floor(math.log10(abs(v)))) + 1
        mant = (Decimal(abs(v)).scaleb(-ex, context=_EXACT)).quantize(Decimal(1).scaleb(-d), rounding=ROUND_HALF_EVEN, context=_EXACT)
        if mant >= 1:                                   # 0.9999995 rounds up to 1.000000
            ex += 1
            mant = (Decimal(abs(v)).scaleb(-ex, context=_EXACT)).quantize(Decimal(1).scaleb(-d), rounding=ROUND_HALF_EVEN, context=_EXACT)
    es = ("E%+03d" % ex) if abs(ex) < 100 else ("%+04d" % ex)
    s = ("-" if v < 0 else "") + format(mant, "f") + es
    if len(s) > w and (s.startswith("0.") or s.startswith("-0.")):
        s = s.replace("0.", ".", 1)
    return s.rjust(w) if len(s) <= w else "*" * w


def _A(text: str, w: Optional[int] = None) -> str:
    """``A`` / ``Aw``: right-justified when shorter than the field, the LEFTMOST w characters when longer."""
    if w is None:
        return text
    return text[:w] if len(text) >= w else text.rjust(w)


def _I(value, w: int) -> str:
    s = str(int(value))
    return s.rjust(w) if len(s) <= w else "*" * w


def _L(flag) -> str:
    return "T" if flag else "F"


def _header(kind: str, domainFileName: str, totalNumPhotons: int, useRayTracing, useRussianRoulette,
            useHybridPhaseFunsForIntenCalcs, hybridPhaseFunWidth, solarFlux, solarMu, solarAzimuth, surfaceAlbedo,
            radiance=None):
    # the driver passes its character(len=256) variable: A60 prints the leftmost 60 characters, i.e. the name
    # left-justified and blank-padded (a shorter actual argument would be right-justified instead)
    lines = ["!   I3RC Monte Carlo 3D Solar Radiative Transfer: " + kind,
             "!  Property_File=" + _A(domainFileName.ljust(256), 60),
             "!  Num_Photons=" + _I(totalNumPhotons, 10),
             "!  PhotonTracing=" + _L(useRayTracing) + "    Russian_Roulette=" + _L(useRussianRoulette),
             "!  Hybrid_Phase_Func_for_Radiance=" + _L(useHybridPhaseFunsForIntenCalcs)
             + "   Gaussian_Phase_Func_Width_deg=" + _F(np.float32(hybridPhaseFunWidth), 5, 2)]
    if radiance is not None:                             # DRV:1463-1466: module variables of the driver
        lines.append("!  Intensity_uses_Russian_Roulette=" + _L(radiance["useRussianRouletteForIntensity"])
                     + "   Intensity_Russian_Roulette_zeta_min=" + _F(np.float32(radiance["zetaMin"]), 5, 2))
        lines.append("!  limited_intensity_contributions=" + _L(radiance["limitIntensityContributions"])
                     + "   max_intensity_contribution=" + _F(np.float32(radiance["maxIntensityContribution"]), 5, 2))
    lines.append("!  Solar_Flux=" + _E(solarFlux, 13, 6) + "   Solar_Mu=" + _F(np.float32(solarMu), 10, 7)
                 + "   Solar_Phi=" + _F(np.float32(solarAzimuth), 7, 3))
    lines.append("!  Lambertian_Surface_Albedo=" + _F(surfaceAlbedo, 7, 4))
    return lines


def _pair(stats, idx) -> str:
    """``2(1X,F9.4)`` of stats(idx, 1:2)."""
    return " " + _F(stats[(0,) + idx], 9, 4) + " " + _F(stats[(1,) + idx], 9, 4)


def formatResults_ASCII(domainFileName: str, totalNumPhotons: int, numBatches: int, useRayTracing: bool,
                        useRussianRoulette: bool, useHybridPhaseFunsForIntenCalcs: bool, hybridPhaseFunWidth: float,
                        solarFlux: float, solarMu: float, solarAzimuth: float, surfaceAlbedo: float,
                        xPosition, yPosition, zPosition,
                        meanFluxUpStats=None, meanFluxDownStats=None, meanFluxAbsorbedStats=None,
                        fluxUpStats=None, fluxDownStats=None, fluxAbsorbedStats=None,
                        absorbedProfileStats=None, absorbedVolumeStats=None,
                        intensityMus: Sequence[float] = (), intensityPhis: Sequence[float] = (), RadianceStats=None,
                        useRussianRouletteForIntensity: bool = True, zetaMin: float = 0.3,
                        limitIntensityContributions: bool = False, maxIntensityContribution: float = 77.0):
    """The text of the four files as a dict {"flux" | "absProf" | "absVolume" | "rad": str}; a table whose
    statistics are not given is left out (the driver skips a table whose file name is blank, DRV:1375, 1410, ...).
    The defaults of the last four arguments are the driver's (DRV:78-82)."""
    x = np.asarray(xPosition, np.float64); y = np.asarray(yPosition, np.float64); z = np.asarray(zPosition, np.float64)
    nx, ny, nz = x.size - 1, y.size - 1, z.size - 1
    xc = [(x[i] + x[i + 1]) / 2.0 for i in range(nx)]            # sum(xPosition(i:i+1))/2.
    yc = [(y[j] + y[j + 1]) / 2.0 for j in range(ny)]
    head = (domainFileName, totalNumPhotons, useRayTracing, useRussianRoulette, useHybridPhaseFunsForIntenCalcs,
            hybridPhaseFunWidth, solarFlux, solarMu, solarAzimuth, surfaceAlbedo)
    out = {}
    if fluxUpStats is not None:                                                       # DRV:1375-1403
        up = np.asarray(fluxUpStats, np.float64).reshape(2, ny, nx)
        dn = np.asarray(fluxDownStats, np.float64).reshape(2, ny, nx)
        ab = np.asarray(fluxAbsorbedStats, np.float64).reshape(2, ny, nx)
        L = _header("Flux", *head)
        L.append("!  Output_Type= Pixel Flux")
        L.append("!  Upwelling_Level=" + _F(z[nz], 7, 3) + "   Downwelling_level=" + _F(z[0], 7, 3))
        L.append("!   X      Y           Flux_Up             Flux_Down            Flux_Absorbed ")
        L.append("!                  Mean     StdErr       Mean     StdErr       Mean     StdErr")
        L.append(_A("!  Average:   ", 14) + "".join(" " + _pair(np.asarray(s, np.float64).reshape(2), ())
                                                    for s in (meanFluxUpStats, meanFluxDownStats, meanFluxAbsorbedStats)))
        for j in range(ny):
            for i in range(nx):
                L.append(_F(xc[i], 7, 3) + _F(yc[j], 7, 3) + "".join(" " + _pair(s, (j, i)) for s in (up, dn, ab)))
        out["flux"] = "\n".join(L) + "\n"
    if absorbedProfileStats is not None:                                              # DRV:1410-1431
        prof = np.asarray(absorbedProfileStats, np.float64).reshape(2, nz)
        L = _header("Absorption Profile", *head)
        L.append("!  Output_Type= Absorption Profile")
        L.append("!   Z    Absorbed_Flux (flux/km) ")
        L.append("!          Mean     StdErr ")
        for k in range(nz):
            L.append(_F(0.5 * (z[k] + z[k + 1]), 7, 3) + " " + _pair(prof, (k,)))
        out["absProf"] = "\n".join(L) + "\n"
    if absorbedVolumeStats is not None:                                               # DRV:1437-1463: i outermost, k innermost
        vol = np.asarray(absorbedVolumeStats, np.float64).reshape(2, nz, ny, nx)
        L = _header("3D Absorption Field", *head)
        L.append("!  Output_Type= Volume Absorption ")
        L.append("!    X       Y        Z       Absorbed_Flux (flux/km)")
        L.append("!                               Mean     StdErr ")
        zc = [(z[k] + z[k + 1]) / 2.0 for k in range(nz)]
        for i in range(nx):
            for j in range(ny):
                for k in range(nz):
                    L.append(_F(xc[i], 7, 3) + " " + _F(yc[j], 7, 3) + " " + _F(zc[k], 7, 3) + " " + _pair(vol, (k, j, i)))
        out["absVolume"] = "\n".join(L) + "\n"
    if RadianceStats is not None:                                                     # DRV:1468-1493
        mus = np.asarray(intensityMus, np.float32); phis = np.asarray(intensityPhis, np.float32)
        numRadDir = int(np.count_nonzero(np.abs(mus) > 0))
        rad = np.asarray(RadianceStats, np.float64).reshape(2, -1, ny, nx)
        L = _header("Radiance", *head, radiance=dict(useRussianRouletteForIntensity=useRussianRouletteForIntensity, zetaMin=zetaMin,
                                                     limitIntensityContributions=limitIntensityContributions,
                                                     maxIntensityContribution=maxIntensityContribution))
        L.append("!  Output_Type= Pixel Radiance")
        L.append("!  RADIANCE AT Z=" + _F(z[nz], 7, 3) + "   NXO=" + _I(nx, 4) + "   NYO=" + _I(ny, 4) + "   NDIR=" + _I(numRadDir, 4))
        L.append("!   X      Y         Radiance (Mean, StdErr)")
        for k in range(numRadDir):
            L.append("! " + " " + _F(mus[k], 8, 5) + " " + _F(phis[k], 6, 2) + "  " + "<- (mu,phi)")
            for j in range(ny):
                for i in range(nx):
                    L.append(_F(xc[i], 7, 3) + _F(yc[j], 7, 3) + " " + _F(rad[0, k, j, i], 9, 4) + " " + _F(rad[1, k, j, i], 9, 4))
        out["rad"] = "\n".join(L) + "\n"
    return out


def writeResults_ASCII(domainFileName, totalNumPhotons, numBatches, useRayTracing, useRussianRoulette,
                       useHybridPhaseFunsForIntenCalcs, hybridPhaseFunWidth, solarFlux, solarMu, solarAzimuth, surfaceAlbedo,
                       xPosition, yPosition, zPosition,
                       outputFluxFile="", meanFluxUpStats=None, meanFluxDownStats=None, meanFluxAbsorbedStats=None,
                       fluxUpStats=None, fluxDownStats=None, fluxAbsorbedStats=None,
                       outputAbsProfFile="", absorbedProfileStats=None, outputAbsVolumeFile="", absorbedVolumeStats=None,
                       outputRadFile="", intensityMus=(), intensityPhis=(), RadianceStats=None, **radianceOptions) -> None:
    """Same argument order as DRV:1324-1334; a blank file name skips that table (``len_trim(...) > 0``)."""
    text = formatResults_ASCII(
        domainFileName, totalNumPhotons, numBatches, useRayTracing, useRussianRoulette, useHybridPhaseFunsForIntenCalcs,
        hybridPhaseFunWidth, solarFlux, solarMu, solarAzimuth, surfaceAlbedo, xPosition, yPosition, zPosition,
        meanFluxUpStats, meanFluxDownStats, meanFluxAbsorbedStats,
        fluxUpStats if outputFluxFile.strip() else None, fluxDownStats, fluxAbsorbedStats,
        absorbedProfileStats if outputAbsProfFile.strip() else None,
        absorbedVolumeStats if outputAbsVolumeFile.strip() else None,
        intensityMus, intensityPhis, RadianceStats if outputRadFile.strip() else None, **radianceOptions)
    for key, name in (("flux", outputFluxFile), ("absProf", outputAbsProfFile), ("absVolume", outputAbsVolumeFile),
                      ("rad", outputRadFile)):
        if key in text:
            with open(name.strip(), "w") as f:
                f.write(text[key])
