"""Synthetic stand-ins for the reference's test domains.

The reference's own inputs (namelists, netCDF domains, the I3RC ``scene43`` Landsat data) are
not in its repository, so each case is rebuilt from the generator sources under
``Domain-Files/`` and the decks under ``run/``.  Every function returns ``(Domain, case)``
where ``case`` carries the illumination and algorithm settings of that deck.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from .opticalProperties import Domain
from .scatteringPhaseFunctions import (henyeyGreenstein, new_PhaseFunction, new_PhaseFunctionTable, rayleigh)

f32 = np.float32

I3RC_VIEWS_MU = [1.0, 0.866, 0.866, 0.5, 0.5]          # C2 / C3 radiance directions
I3RC_VIEWS_PHI = [0.0, 0.0, 180.0, 0.0, 180.0]


def _hg_table(g=0.85, nLegendre=64):
    return new_PhaseFunctionTable([henyeyGreenstein(g, nLegendre)], key=[1.0],
                                  tableDescription="Henyey-Greenstein with g = %g" % g)


def homogeneous_slab(ssa=0.99, tau=10.0, albedo=0.2, n=20, delta=0.0625, g=0.85, nLegendre=64,
                     temperature=0.0, lambda_um=0.0) -> Tuple[Domain, Dict]:
    """C1 ``I3RC_mono_SWhomog``: n x n x n cells of ``delta`` km (f32-exact => regular path, q1),
    uniform extinction tau/(n*delta), one HG component; mu0 = 0.5; fluxes only."""
    edges = delta * np.arange(n + 1, dtype=np.float64)
    d = Domain(edges, edges, edges, temps=np.full((n, n, n), temperature), surfaceAlbedo=albedo, lambda_um=lambda_um)
    ext = np.full((n, n, n), tau / (n * delta))
    d.addOpticalComponent("cloud", ext, np.full((n, n, n), ssa), np.ones((n, n, n), np.int32), _hg_table(g, nLegendre))
    d.getOpticalPropertiesByComponent()
    return d, dict(name="C1_SWhomog", solarMu=0.5, solarAzimuth=0.0, LW_flag=-1.0, numPhotonsPerBatch=10000,
                   numBatches=100, iseed=10)


def homogeneous_lw(ssa=0.5, tau=10.0, albedo=0.1, n=20, delta=0.0625, atmTemp=290.0, sfcTemp=300.0,
                   lambda_um=10.0) -> Tuple[Domain, Dict]:
    """C4 ``I3RC_mono_LWhomog``: the C1 grid with a thermal source (z0 = 0 required, q4)."""
    d, case = homogeneous_slab(ssa=ssa, tau=tau, albedo=albedo, n=n, delta=delta, temperature=atmTemp,
                               lambda_um=lambda_um)
    case.update(name="C4_LWhomog", LW_flag=1.0, surfaceTemp=sfcTemp)
    return d, case


def step_cloud(ssa=1.0, solarMu=1.0) -> Tuple[Domain, Dict]:
    """C2 I3RC step cloud (``Domain-Files/i3rcStepCloud.f95:27-84``): 32 x 1 x 32 cells over
    500 x 500 x 250 (generator units), tau = 2 (columns 1-16) / 18 (17-32), HG g = 0.85, 64 terms."""
    nColumns, nLayers = 32, 32
    deltaX = f32(500.0) / f32(nColumns)
    deltaZ = f32(250.0) / f32(nLayers)
    x = (deltaX * np.arange(nColumns + 1, dtype=f32)).astype(np.float64)
    z = (deltaZ * np.arange(nLayers + 1, dtype=f32)).astype(np.float64)
    d = Domain(x, [0.0, 500.0], z, surfaceAlbedo=0.0)
    tau = np.concatenate([np.full(nColumns // 2, 2.0, f32), np.full(nColumns // 2, 18.0, f32)])
    ext = np.broadcast_to((tau / f32(250.0)).astype(np.float64)[None, None, :], (nLayers, 1, nColumns)).copy()
    d.addOpticalComponent("cloud", ext, np.full(ext.shape, ssa), np.ones(ext.shape, np.int32), _hg_table(0.85, 64))
    d.getOpticalPropertiesByComponent()
    return d, dict(name="C2_stepcloud", solarMu=solarMu, solarAzimuth=0.0, LW_flag=-1.0,
                   intensityMus=I3RC_VIEWS_MU, intensityPhis=I3RC_VIEWS_PHI,
                   useRussianRouletteForIntensity=True, zetaMin=0.3, numPhotonsPerBatch=10000)


def _gaussian_field(n, rng, slope=-5.0 / 3.0):
    """Periodic 2-D Gaussian field whose 1-D power spectrum falls as k**slope (unit variance)."""
    k = np.fft.fftfreq(n) * n
    kx, ky = np.meshgrid(k, k, indexing="ij")
    kk = np.sqrt(kx * kx + ky * ky)
    kk[0, 0] = 1.0
    amp = kk ** ((slope - 1.0) / 2.0)
    amp[0, 0] = 0.0
    noise = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    field = np.real(np.fft.ifft2(noise * amp))
    return (field - field.mean()) / field.std()


def landsat_cloud(ssa=1.0, nxy=128, mie=False, seed=43, nLegendre=299) -> Tuple[Domain, Dict]:
    """C3 I3RC Landsat cloud (``Domain-Files/i3rcLandsatCloud.f95:27-123``): 128 x 128 x 119 cells of
    30 x 30 x 20 (generator units), z0 = 200.  The ``scene43`` optical-depth / thickness files are
    not distributed with the reference, so the field is synthetic: lognormal optical depth (mean 11,
    sigma_ln 0.9, k^-5/3 spectrum, 20 % clear), thickness 20*round(15*sqrt(tau)) clipped to
    [20, 2380], uniform extinction tau/h in the lowest h/20 layers of each column and empty cells
    (extinction 0, phase index 0) above -- the same construction as the generator.

    ``mie=True`` adds structure the HG case lacks: a 16-entry phase table keyed by effective radius
    (double-HG surrogates) chosen per column, plus a Rayleigh component (nc = 2).
    """
    deltaXY, deltaZ, maxThickness = 30.0, 20, 2380
    nLayers = (maxThickness + deltaZ // 2) // deltaZ
    rng = np.random.default_rng(seed)
    gfield = _gaussian_field(nxy, rng)
    sig = 0.9
    tau = np.exp(np.log(11.0) - 0.5 * sig * sig + sig * gfield)
    tau[gfield < np.quantile(gfield, 0.20)] = 0.0
    tau = np.round(tau, 2).astype(f32)                                     # file format f7.2
    thick = np.clip(deltaZ * np.round(15.0 * np.sqrt(tau.astype(np.float64))), deltaZ, maxThickness)
    nlev = np.rint(thick / deltaZ).astype(int)
    x = deltaXY * np.arange(nxy + 1, dtype=np.float64)
    z = deltaZ * np.arange(nLayers + 1, dtype=np.float64) + 200.0
    d = Domain(x, x, z, surfaceAlbedo=0.0)
    # arrays are (nz, ny, nx); tau[i, j] is column (x = i, y = j)
    ext = np.zeros((nLayers, nxy, nxy)); ssaA = np.zeros_like(ext); idx = np.zeros(ext.shape, np.int32)
    lev = np.arange(nLayers)[:, None, None]
    cloudy = (lev < nlev.T[None]) & (tau.T[None] > 0)
    colext = np.where(tau > 0, tau.astype(np.float64) / (nlev * deltaZ), 0.0).T
    ext[:] = np.where(cloudy, colext[None], 0.0)
    ssaA[cloudy] = ssa
    if not mie:
        idx[cloudy] = 1
        table = _hg_table(0.85, nLegendre)
        d.addOpticalComponent("cloud", ext, ssaA, idx, table)
    else:
        nE = 16
        reff = np.linspace(5.0, 20.0, nE)
        pfs = []
        ang = np.linspace(0.0, np.pi, 721).astype(f32)
        ang[-1] = f32(np.pi)
        for re in reff:                                                    # double-HG surrogates of Mie functions
            g1 = 0.80 + 0.006 * (re - 5.0); g2 = -0.45; f = 0.97
            mu = np.cos(ang.astype(np.float64))
            hg = lambda gg: (1 - gg * gg) / (1 + gg * gg - 2 * gg * mu) ** 1.5
            pfs.append(new_PhaseFunction(scatteringAngle=ang, value=(f * hg(g1) + (1 - f) * hg(g2)).astype(f32),
                                         description="double-HG reff=%.1f" % re))
        table = new_PhaseFunctionTable(pfs, key=reff, tableDescription="Mie surrogate table")
        colidx = 1 + np.clip(np.rint((np.sqrt(tau.astype(np.float64)) / np.sqrt(max(tau.max(), 1e-6))) * (nE - 1)), 0, nE - 1).astype(np.int32)
        idx[:] = np.where(cloudy, colidx.T[None], 0)
        d.addOpticalComponent("cloud", ext, ssaA, idx, table)
        # Rayleigh: horizontally uniform, exponential profile, conservative (OPT:2075-2081)
        zc = 0.5 * (z[1:] + z[:-1])
        rext = 1.2e-5 * np.exp(-(zc - 200.0) / 8000.0)
        d.addOpticalComponent("rayleigh", rext, np.ones(nLayers), np.ones(nLayers, np.int32),
                              new_PhaseFunctionTable([rayleigh()], key=[0.0], tableDescription="Rayleigh Scattering"))
    d.getOpticalPropertiesByComponent()
    return d, dict(name="C3_landsat" + ("_mie" if mie else "_hg"), solarMu=0.5, solarAzimuth=0.0, LW_flag=-1.0,
                   intensityMus=I3RC_VIEWS_MU, intensityPhis=I3RC_VIEWS_PHI,
                   useRussianRouletteForIntensity=True, zetaMin=0.3, numPhotonsPerBatch=10000,
                   meanOpticalDepth=float(tau.mean()))


def irregular_test_domain(albedo=0.3, stretched=False) -> Tuple[Domain, Dict]:
    """T-irr (trace harness): 12 x 10 x 8 cells with spacings 0.05 / 0.03 / 0.02 km that are NOT
    exactly representable in single precision, so new_Integrator takes the irregular path
    (quirks q1-q3); two components (one horizontally uniform, partial height), a layer with no
    extinction, absorbing cloud, reflecting surface."""
    nx, ny, nz = 12, 10, 8
    x = 0.05 * np.arange(nx + 1); y = 0.03 * np.arange(ny + 1); z = 0.02 * np.arange(nz + 1)
    if stretched:                       # genuinely non-uniform spacings (the throughput kernel's edge-table variant)
        x = np.concatenate([[0.0], np.cumsum(0.05 * (1.0 + 0.3 * np.sin(1.0 + np.arange(nx))))])
        y = np.concatenate([[0.0], np.cumsum(0.03 * (1.0 + 0.25 * np.cos(0.5 + np.arange(ny))))])
        z = np.concatenate([[0.0], np.cumsum(0.012 * 1.15 ** np.arange(nz))])
    d = Domain(x, y, z, temps=np.full((nz, ny, nx), 280.0), surfaceAlbedo=albedo, lambda_um=10.0)
    rng = np.random.default_rng(7)
    ext = rng.uniform(5.0, 60.0, size=(nz, ny, nx))
    ext[4] = 0.0                                           # a layer with no cloud extinction
    ssa = np.full(ext.shape, 0.9)
    idx = np.ones(ext.shape, np.int32)
    idx[ext == 0] = 0
    ssa[ext == 0] = 0.0
    pfs = [henyeyGreenstein(0.85, 64), henyeyGreenstein(0.6, 32)]
    idx[(ext > 30.0)] = 2
    d.addOpticalComponent("cloud", ext, ssa, idx, new_PhaseFunctionTable(pfs, key=[1.0, 2.0]))
    # gas-like component: horizontally uniform, levels 2..4 only, weakly scattering Rayleigh
    d.addOpticalComponent("gas", np.array([3.0, 2.0, 1.0]), np.array([0.3, 0.3, 0.3]), np.ones(3, np.int32),
                          new_PhaseFunctionTable([rayleigh()], key=[0.0]), zLevelBase=2)
    d.getOpticalPropertiesByComponent()
    return d, dict(name="T_irr_stretched" if stretched else "T_irr", solarMu=0.6, solarAzimuth=30.0, LW_flag=-1.0,
                   intensityMus=[1.0, 0.7, -0.5], intensityPhis=[0.0, 45.0, 200.0])


def bench_domain(nxy=325, nz=150, seed=5, ssa=0.999) -> Tuple[Domain, Dict]:
    """C5 ``I3RC_bench_SW``: 325 x 325 x 150 cells of 0.0625 x 0.0625 x 0.03125 km (f32-exact), a
    C3-style synthetic cloud field between 0.5 and 2.5 km over a thin molecular background."""
    dxy, dz = 0.0625, 0.03125
    rng = np.random.default_rng(seed)
    n2 = 512
    gfield = _gaussian_field(n2, rng)[:nxy, :nxy]
    tau = np.exp(np.log(8.0) - 0.405 + 0.9 * gfield)
    tau[gfield < np.quantile(gfield, 0.3)] = 0.0
    x = dxy * np.arange(nxy + 1, dtype=np.float64)
    z = dz * np.arange(nz + 1, dtype=np.float64)
    d = Domain(x, x, z, temps=np.broadcast_to((288.0 - 6.5 * 0.5 * (z[1:] + z[:-1]))[:, None, None], (nz, nxy, nxy)).copy(),
               surfaceAlbedo=0.05, lambda_um=0.55)
    base = 16
    nlev = np.clip(np.rint(12.0 * np.sqrt(tau)), 1, 64).astype(int)
    lev = np.arange(nz)[:, None, None]
    cloudy = (lev >= base) & (lev < base + nlev.T[None]) & (tau.T[None] > 0)
    colext = np.where(tau > 0, tau / (nlev * dz), 0.0).T
    ext = np.where(cloudy, colext[None], 0.0)
    ssaA = np.where(cloudy, ssa, 0.0)
    idx = np.where(cloudy, 1, 0).astype(np.int32)
    d.addOpticalComponent("cloud", ext, ssaA, idx, _hg_table(0.85, 128))
    zc = 0.5 * (z[1:] + z[:-1])
    d.addOpticalComponent("rayleigh", 0.012 * np.exp(-zc / 8.0), np.ones(nz), np.ones(nz, np.int32),
                          new_PhaseFunctionTable([rayleigh()], key=[0.0]))
    d.getOpticalPropertiesByComponent()
    return d, dict(name="C5_bench", solarMu=0.5, solarAzimuth=0.0, LW_flag=-1.0, numPhotonsPerBatch=10000)


def bench_problem(nxy=325, nz=150, seed=5, ssa=0.999):
    """The C5 scene of ``bench_domain`` described the way the driver holds it (DRV:903-947): the wavelength-independent
    physical state of ``type(commonDomain)`` -- here the cloud's mass concentration and effective radius, the molecular
    number concentration and density profiles -- plus a one-wavelength SSP table (0.55 um).  ``read_SSPTable`` turns it
    into the dense optical arrays, on the host (NumPy mirror) or in HBM (``mcb_set_physical`` once per run, then
    ``mcb_assemble_optics`` with a few hundred bytes per wavelength).  Returns (commonDomain, [SSPTable], case)."""
    from .opticalProperties import SSPComponent, SSPTable, commonDomain, light_spd
    d, case = bench_domain(nxy=nxy, nz=nz, seed=seed, ssa=ssa)
    cloud = d.components[0]
    zc = 0.5 * (d.zPosition[1:] + d.zPosition[:-1])
    mass = cloud.extinction[..., None].copy()                       # unit mass extinction coefficient: ext = massConc * 1
    reff = np.where(cloud.extinction > 0, 10.0, 0.0)[..., None].copy()
    numConc = np.broadcast_to((2.55e25 * np.exp(-zc / 8.0))[:, None, None], (nz, nxy, nxy)).copy()
    rho = np.broadcast_to((1.225 * np.exp(-zc / 8.0))[:, None, None], (nz, nxy, nxy)).copy()
    common = commonDomain(d.xPosition, d.yPosition, d.zPosition, d.temps, mass, reff, numConc, rho)
    key = np.array([5.0, 15.0], dtype=f32)
    table = SSPTable(np.array([light_spd * 1e6 / 0.55]), np.array([d.surfaceAlbedo]),
                     [SSPComponent("cloud", "volExt", 1, key, np.ones((1, 2)), np.full((1, 2), ssa),
                                   tables=[new_PhaseFunctionTable([henyeyGreenstein(0.85, 128), henyeyGreenstein(0.85, 128)],
                                                                  key=key.astype(np.float64))])])
    case = dict(case, name="C5_bench_physical", calcRayl=True)
    return common, [table], case


def broadband_problem(nxy=24, nz=20, nLambda=6, seed=11, lw=False):
    """C5-style multi-wavelength input (``run/I3RC_bench_SW.deck`` / ``_LW.deck``): the wavelength-independent physical
    state of ``type(commonDomain)`` (liquid-water mass concentration and effective radius of one cloud component,
    molecular number concentration and density profiles) plus an in-memory SSP table with ``nLambda`` wavelength bins
    (per bin: extinction / albedo of the cloud as functions of effective radius, a keyed set of phase functions, a gas
    absorption cross-section profile, a surface albedo).  Returns (commonDomain, [SSPTable], case)."""
    from .opticalProperties import SSPComponent, SSPTable, commonDomain, light_spd
    rng = np.random.default_rng(seed)
    dxy, dz = 0.0625, 2.5 / nz                       # 2.5 km deep (0.125-km layers at nz = 20; f32-exact for nz = 20, 40, 80, 160)
    x = dxy * np.arange(nxy + 1, dtype=np.float64)
    z = dz * np.arange(nz + 1, dtype=np.float64)
    zc = 0.5 * (z[1:] + z[:-1])
    n2 = 64
    while n2 < nxy:
        n2 *= 2
    g = _gaussian_field(n2, rng)[:nxy, :nxy]
    lwp = np.where(g > np.quantile(g, 0.35), np.exp(0.6 * g), 0.0)               # relative liquid water path per column
    base, top = nz // 4, nz // 4 + max(2, nz // 3)
    mass = np.zeros((nz, nxy, nxy)); reff = np.zeros((nz, nxy, nxy))
    for k in range(base, top):
        adiab = (k - base + 0.5) / (top - base)                                  # adiabatic-like growth with height
        mass[k] = 0.06 * lwp.T * adiab                                           # g m^-3
        reff[k] = np.where(lwp.T > 0, 6.0 + 9.0 * adiab ** (1.0 / 3.0) + 0.5 * rng.random((nxy, nxy)), 0.0)
    numConc = np.broadcast_to((2.55e25 * np.exp(-zc / 8.0))[:, None, None], (nz, nxy, nxy)).copy()
    rho = np.broadcast_to((1.225 * np.exp(-zc / 8.0))[:, None, None], (nz, nxy, nxy)).copy()
    temps = np.broadcast_to((288.0 - 6.5 * zc)[:, None, None], (nz, nxy, nxy)).copy()
    common = commonDomain(x, x, z, temps, mass[..., None].copy(), reff[..., None].copy(), numConc, rho)
    lam = np.linspace(8.0, 12.0, nLambda) if lw else np.linspace(0.45, 2.1, nLambda)      # microns
    freq = light_spd * 1e6 / lam
    key = np.array([4.0, 6.0, 8.0, 10.0, 13.0, 16.0, 20.0], dtype=f32)
    extT = np.empty((nLambda, key.size)); ssaT = np.empty((nLambda, key.size)); tabs = []
    for i in range(nLambda):
        qext = 2.0 + 0.6 * (lam[i] / key.astype(np.float64)) ** 0.7                      # extinction efficiency ~ 2
        extT[i] = 750.0 * qext / key.astype(np.float64)                                  # km^-1 per (g m^-3): 3 Q / (4 rho_w r)
        absorb = (0.02 + 0.4 * (lam[i] / 12.0) ** 2) if lw else 1e-4 * np.exp(3.2 * (lam[i] - 0.45))
        ssaT[i] = np.clip(1.0 - absorb * np.sqrt(key.astype(np.float64) / 10.0), 0.3, 1.0)
        tabs.append(new_PhaseFunctionTable([henyeyGreenstein(min(0.8 + 0.005 * float(k), 0.9), 48) for k in key],
                                           key=key.astype(np.float64)))
    xsec = np.outer(1e-29 * (1.0 + np.sin(np.arange(nLambda)) ** 2), np.exp(-zc / 2.0))  # m^2 per molecule, (nLambda, nz)
    albedo = np.full(nLambda, 0.02 if lw else 0.06) + 0.01 * np.arange(nLambda) / max(1, nLambda - 1)
    table = SSPTable(freq, albedo, [SSPComponent("cloud", "volExt", 1, key, extT, ssaT, tables=tabs),
                                    SSPComponent("gas", "absXsec", 1, xsec=xsec)])
    return common, [table], dict(name="C5_broadband", solarMu=0.5, solarAzimuth=0.0, LW_flag=1.0 if lw else -1.0,
                                 numLambda=nLambda, surfaceTemp=290.0)
