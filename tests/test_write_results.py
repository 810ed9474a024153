"""writeResults_ASCII (DRV:1324-1495): the four ASCII tables, character for character.  The expected lines below
were derived by hand from the reference's format strings -- '(2(F7.3),3(1X,2(1X,F9.4)))' and friends -- not from
this implementation."""
import os
import subprocess

import numpy as np

from mcbrat3d_b200.writeResults import _A, _E, _F, _I, formatResults_ASCII, writeResults_ASCII

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_edit_descriptors():
    assert _F(0.5, 9, 4) == "   0.5000"
    assert _F(12.34567, 9, 4) == "  12.3457"
    assert _F(-0.00001, 9, 4) == "  -0.0000"
    assert _F(123456.0, 9, 4) == "*********"            # does not fit: a row of asterisks
    assert _F(-1234.5, 9, 4) == "*********"
    assert _F(9999.99996, 9, 4) == "*********"          # rounds up to 10000.0000: 10 characters
    assert _F(0.0, 7, 3) == "  0.000"
    assert _F(0.125, 5, 2) == " 0.12"                   # tie on an exactly representable value: round half to even
    assert _F(0.375, 5, 2) == " 0.38"
    assert _F(-0.5, 4, 2) == "-.50"                     # the optional leading zero goes when the field is one short
    assert _F(np.float32(0.3), 5, 2) == " 0.30"
    assert _F(np.float32(0.5), 10, 7) == " 0.5000000"
    assert _F(3.402823466e38, 5, 2) == "*****"
    assert _E(1.0, 13, 6) == " 0.100000E+01"
    assert _E(1367.0, 13, 6) == " 0.136700E+04"
    assert _E(-0.00123456789, 13, 6) == "-0.123457E-02"
    assert _E(0.0, 13, 6) == " 0.000000E+00"
    assert _E(0.99999996, 13, 6) == " 0.100000E+01"
    assert _I(3200000, 10) == "   3200000" and _I(10 ** 10, 10) == "**********"
    assert _A("abc", 5) == "  abc" and _A("abcdefg", 5) == "abcde"


def _problem():
    x = np.array([0.0, 0.5, 1.0]); y = np.array([0.0, 2.0]); z = np.array([0.0, 0.25, 0.5])
    nx, ny, nz, nd = 2, 1, 2, 2
    up = np.array([[[0.25, 0.5]], [[0.001, 0.002]]])                        # (2, ny, nx)
    dn = np.array([[[0.125, 1.0]], [[0.0005, 0.00004]]])
    ab = np.array([[[12.34567, 0.0]], [[0.0, 0.0]]])
    prof = np.array([[0.75, 1.5], [0.01, 0.02]])                           # (2, nz)
    vol = np.arange(2 * nz * ny * nx, dtype=np.float64).reshape(2, nz, ny, nx) / 8.0
    rad = np.arange(2 * nd * ny * nx, dtype=np.float64).reshape(2, nd, ny, nx) / 16.0
    return dict(domainFileName="step.dom", totalNumPhotons=3200000, numBatches=32, useRayTracing=True, useRussianRoulette=True,
                useHybridPhaseFunsForIntenCalcs=False, hybridPhaseFunWidth=7.0, solarFlux=1.0, solarMu=0.5, solarAzimuth=30.0,
                surfaceAlbedo=0.2, xPosition=x, yPosition=y, zPosition=z,
                meanFluxUpStats=[0.375, 0.0015], meanFluxDownStats=[0.5625, 0.0003], meanFluxAbsorbedStats=[6.1728, 0.0],
                fluxUpStats=up, fluxDownStats=dn, fluxAbsorbedStats=ab, absorbedProfileStats=prof, absorbedVolumeStats=vol,
                intensityMus=[1.0, 0.866], intensityPhis=[0.0, 180.0], RadianceStats=rad)


HEAD = ["!  Property_File=" + "step.dom" + " " * 52,
        "!  Num_Photons=   3200000",
        "!  PhotonTracing=T    Russian_Roulette=T",
        "!  Hybrid_Phase_Func_for_Radiance=F   Gaussian_Phase_Func_Width_deg= 7.00"]
SOLAR = ["!  Solar_Flux= 0.100000E+01   Solar_Mu= 0.5000000   Solar_Phi= 30.000",
         "!  Lambertian_Surface_Albedo= 0.2000"]


def test_flux_table():
    text = formatResults_ASCII(**_problem())["flux"].split("\n")
    assert text[:13] == ["!   I3RC Monte Carlo 3D Solar Radiative Transfer: Flux"] + HEAD + SOLAR + [
        "!  Output_Type= Pixel Flux",
        "!  Upwelling_Level=  0.500   Downwelling_level=  0.000",
        "!   X      Y           Flux_Up             Flux_Down            Flux_Absorbed ",
        "!                  Mean     StdErr       Mean     StdErr       Mean     StdErr",
        "!  Average:        0.3750    0.0015     0.5625    0.0003     6.1728    0.0000",
        "  0.250  1.000     0.2500    0.0010     0.1250    0.0005    12.3457    0.0000"]
    assert text[13] == "  0.750  1.000     0.5000    0.0020     1.0000    0.0000     0.0000    0.0000"
    assert text[14] == "" and len(text) == 15


def test_absorption_tables():
    out = formatResults_ASCII(**_problem())
    prof = out["absProf"].split("\n")
    assert prof[0] == "!   I3RC Monte Carlo 3D Solar Radiative Transfer: Absorption Profile"
    assert prof[7:] == ["!  Output_Type= Absorption Profile", "!   Z    Absorbed_Flux (flux/km) ", "!          Mean     StdErr ",
                        "  0.125     0.7500    0.0100", "  0.375     1.5000    0.0200", ""]
    vol = out["absVolume"].split("\n")
    assert vol[7:10] == ["!  Output_Type= Volume Absorption ", "!    X       Y        Z       Absorbed_Flux (flux/km)",
                         "!                               Mean     StdErr "]
    # i outermost, k innermost (DRV:1453-1461); vol[s, k, j, i] = (i + 2 k + 4 s) / 8
    assert vol[10:14] == ["  0.250   1.000   0.125     0.0000    0.5000", "  0.250   1.000   0.375     0.2500    0.7500",
                          "  0.750   1.000   0.125     0.1250    0.6250", "  0.750   1.000   0.375     0.3750    0.8750"]


def test_radiance_table():
    rad = formatResults_ASCII(**_problem())["rad"].split("\n")
    assert rad[:5] == ["!   I3RC Monte Carlo 3D Solar Radiative Transfer: Radiance"] + HEAD
    assert rad[5:7] == ["!  Intensity_uses_Russian_Roulette=T   Intensity_Russian_Roulette_zeta_min= 0.30",
                        "!  limited_intensity_contributions=F   max_intensity_contribution=77.00"]
    assert rad[7:9] == SOLAR
    assert rad[9:12] == ["!  Output_Type= Pixel Radiance", "!  RADIANCE AT Z=  0.500   NXO=   2   NYO=   1   NDIR=   2",
                         "!   X      Y         Radiance (Mean, StdErr)"]
    # rad[s, d, j, i] = (i + 2 d + 4 s) / 16
    assert rad[12:] == ["!   1.00000   0.00  <- (mu,phi)", "  0.250  1.000    0.0000    0.2500", "  0.750  1.000    0.0625    0.3125",
                        "!   0.86600 180.00  <- (mu,phi)", "  0.250  1.000    0.1250    0.3750", "  0.750  1.000    0.1875    0.4375", ""]


def test_blank_file_names_skip_tables(tmp_path):
    p = _problem()
    writeResults_ASCII(p["domainFileName"], p["totalNumPhotons"], p["numBatches"], True, True, False, 7.0, 1.0, 0.5, 30.0, 0.2,
                       p["xPosition"], p["yPosition"], p["zPosition"],
                       outputFluxFile=str(tmp_path / "flux.out"), meanFluxUpStats=p["meanFluxUpStats"],
                       meanFluxDownStats=p["meanFluxDownStats"], meanFluxAbsorbedStats=p["meanFluxAbsorbedStats"],
                       fluxUpStats=p["fluxUpStats"], fluxDownStats=p["fluxDownStats"], fluxAbsorbedStats=p["fluxAbsorbedStats"],
                       outputAbsProfFile="   ", absorbedProfileStats=p["absorbedProfileStats"])
    assert sorted(os.listdir(tmp_path)) == ["flux.out"]
    assert open(tmp_path / "flux.out").read() == formatResults_ASCII(**p)["flux"]


def test_compiled_host_writes_the_same_files(tmp_path):
    """host/mcbrat_host.hpp::writeResults_ASCII (the compiled host) against this module, byte for byte, on values that
    exercise rounding ties, negative zeros, overflowing fields and long file names."""
    src = os.path.join(ROOT, "tests", "cpp", "write_ascii_check.cpp")
    exe = str(tmp_path / "write_ascii_check")
    csrc = os.path.join(ROOT, "mcbrat3d_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, src, "-L" + csrc, "-lmcbrat_cuda", "-Wl,-rpath," + csrc])
    subprocess.check_call([exe, str(tmp_path)])
    rng = np.random.default_rng(5)
    nx, ny, nz, nd = 3, 2, 4, 3
    x = 0.0625 * np.arange(nx + 1); y = 0.5 * np.arange(ny + 1) + 100.0; z = np.array([0.2, 0.45, 0.7, 1.7, 9.95])

    def lcg(n):                                      # the same deterministic values the C++ program generates
        v = np.empty(n); s = 12345
        for i in range(n):
            s = (s * 1103515245 + 12345) % (1 << 31)
            v[i] = (s / float(1 << 31) - 0.3) * (10.0 ** ((i % 7) - 3))
        return v
    vals = lcg(2 * (3 + 3 * nx * ny + nz + nx * ny * nz + nd * nx * ny))
    o = 0

    def take(shape):
        nonlocal o
        n = int(np.prod(shape)); a = vals[o:o + n].reshape(shape); o += n
        return a
    mf = take((2, 3)); up = take((2, ny, nx)); dn = take((2, ny, nx)); ab = take((2, ny, nx)); prof = take((2, nz))
    vol = take((2, nz, ny, nx)); rad = take((2, nd, ny, nx))
    up[0, 0, 0] = 0.00005; up[1, 0, 0] = -0.00001; dn[0, 0, 0] = 99999.99996; dn[1, 0, 0] = 2.5e-5
    name = "a_rather_long_domain_file_name_that_exceeds_sixty_characters_by_a_good_margin.dom"
    want = formatResults_ASCII(name, 12345678901, 17, False, True, True, 3.25, 1367.0, 0.8660254, 275.5, 0.05, x, y, z,
                               mf[:, 0], mf[:, 1], mf[:, 2], up, dn, ab, prof, vol, [1.0, -0.5, 0.25], [0.0, 90.0, 359.99], rad,
                               useRussianRouletteForIntensity=False, zetaMin=0.15, limitIntensityContributions=True,
                               maxIntensityContribution=3.402823466e38)
    for key, fname in (("flux", "flux.out"), ("absProf", "absprof.out"), ("absVolume", "absvol.out"), ("rad", "rad.out")):
        got = open(tmp_path / fname).read()
        assert got == want[key], (key, [(a, b) for a, b in zip(got.split("\n"), want[key].split("\n")) if a != b][:3])
