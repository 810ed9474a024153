"""Independent plane-parallel solver (tests only): the adding-doubling method for the azimuthally averaged radiance
field of a homogeneous slab over a Lambertian surface, written from the textbook equations (van de Hulst 1963;
Hansen & Travis 1974, section 4) -- it shares no code, no random numbers and no algorithm with the Monte Carlo path.

It is the deterministic answer for the protocol of ``Drivers/planeParallel.f95:242, 269-272`` (tau, omega, g, theta0
-> Fup, Fdn): SURVEY section 8c lists it as the substitute pin for a reference that ships no golden vectors.  Fluxes
are fractions of the incident solar flux on a horizontal surface, which is the normalisation of
``computeRadiativeTransfer`` (INT:328-350).

Discretisation.  Nodes: N Gauss-Legendre points on (0, 1] for each hemisphere plus the solar zenith cosine as an extra
node of zero weight.  State vector: the hemispheric flux carried by each node, y_i = w_i mu_i I_i; the direct beam is
the unit vector on the solar node.  For an optically thin layer d
    T_ij = delta_ij (1 - d / mu_i) + w_i omega d P(mu_i,  mu_j) / (2 mu_j)
    R_ij =                           w_i omega d P(mu_i, -mu_j) / (2 mu_j)
with P the azimuth-averaged phase function sum_l (2l+1) chi_l P_l(mu) P_l(mu'), renormalised on the quadrature so that
each column conserves energy.  Doubling: R2 = R + T (1 - R R)^-1 R T, T2 = T (1 - R R)^-1 T.
"""
import numpy as np


def _legendre_matrix(lmax, x):
    P = np.empty((lmax + 1, x.size))
    P[0] = 1.0
    if lmax >= 1:
        P[1] = x
    for l in range(1, lmax):
        P[l + 1] = ((2 * l + 1) * x * P[l] - l * P[l - 1]) / (l + 1)
    return P


def slab_fluxes(tau, omega, chi, mu0, albedo=0.0, nStreams=64, thin=1.0e-5, muOut=()):
    """(Fup at the top, Fdown at the surface incl. the direct beam, absorbed in the slab), per unit incident flux.
    ``chi``: Legendre coefficients chi_l of the phase function, chi_0 = 1 (Henyey-Greenstein: g**l).
    With ``muOut`` a fourth value is returned: the azimuthally averaged upward radiance at the top, per unit incident
    flux and per steradian, at those zenith cosines (extra nodes of negligible weight: their flux share divided by
    weight, cosine and 2 pi)."""
    chi = np.asarray(chi, dtype=np.float64)
    x, w = np.polynomial.legendre.leggauss(nStreams)
    eps = 1.0e-9
    muOut = np.asarray(muOut, dtype=np.float64).reshape(-1)
    mu = np.concatenate([0.5 * (x + 1.0), muOut, [mu0]])
    wt = np.concatenate([0.5 * w, np.full(muOut.size, eps), [0.0]])
    L = chi.size - 1
    Pl = _legendre_matrix(L, mu)
    fac = (2 * np.arange(L + 1) + 1) * chi
    Pf = (Pl * fac[:, None]).T @ Pl                                            # P(mu_i,  mu_j)
    Pb = (Pl * (fac * (-1.0) ** np.arange(L + 1))[:, None]).T @ Pl             # P(mu_i, -mu_j)
    norm = 0.5 * (wt[:, None] * (Pf + Pb)).sum(axis=0)                         # should be 1 per column
    Pf = Pf / norm[None, :]
    Pb = Pb / norm[None, :]
    # doublings: the starting layer is `thin` times the smallest node cosine (the single-scattering start is first order)
    n = max(int(np.ceil(np.log2(max(tau, 1e-300) / (thin * mu.min())))), 0)
    d = tau / 2.0 ** n
    T = np.diag(np.exp(-d / mu)) + wt[:, None] * omega * d * Pf / (2.0 * mu[None, :])
    R = wt[:, None] * omega * d * Pb / (2.0 * mu[None, :])
    I = np.eye(mu.size)
    for _ in range(n):
        G = np.linalg.solve(I - R @ R, np.concatenate([T, R @ T], axis=1))
        TG, TGR = T @ G[:, :mu.size], T @ G[:, mu.size:]
        R, T = R + TGR, TG
    # Lambertian surface: the reflected flux is shared among the nodes like w_i mu_i
    lam = wt * mu / (wt * mu).sum()
    Rs = albedo * np.outer(lam, np.ones(mu.size))
    y0 = np.zeros(mu.size); y0[-1] = 1.0
    down = np.linalg.solve(I - R @ Rs, T @ y0)                                 # flux reaching the surface, all orders
    up = R @ y0 + T @ (Rs @ down)
    fup, fdn = up.sum(), down.sum()
    if muOut.size:
        sl = slice(nStreams, nStreams + muOut.size)
        return fup, fdn, 1.0 - fup - (1.0 - albedo) * fdn, up[sl] / (wt[sl] * mu[sl] * 2.0 * np.pi)
    return fup, fdn, 1.0 - fup - (1.0 - albedo) * fdn


def table_moments(inverseTable, lmax=127):
    """Legendre moments chi_l of the phase function the photon loop actually samples: computeScatteringAngle
    (INT:1594-1621) takes entry int(RN * nS) + 1 of the inverse table, i.e. every entry but the last with equal
    probability.  The reference builds that table from a CDF on only max(nMoments, 2) Lobatto nodes (INV:107-112), so
    the sampled function is not the analytic one (Henyey-Greenstein g = 0.85, 64 terms -> g_eff = 0.8517); the
    deterministic solver must be given the same scattering law to be a test of the transport."""
    theta = np.asarray(inverseTable, dtype=np.float64).reshape(-1)
    return _legendre_matrix(lmax, np.cos(theta[:-1])).mean(axis=1)


def layered_fluxes(layers, chi, mu0, albedo=0.0, nStreams=64, thin=1.0e-5):
    """Fluxes of a stack of homogeneous layers, TOP FIRST, each (tau, omega), all with the phase function ``chi``:
    every layer by doubling, the stack by adding (R12 = R1 + T1 (1 - R2 R1)^-1 R2 T1, T12 = T2 (1 - R1 R2)^-1 T1; a
    homogeneous layer reflects alike from both sides).  Returns (Fup at the top, Fdown at the surface, absorbed)."""
    chi = np.asarray(chi, dtype=np.float64)
    x, w = np.polynomial.legendre.leggauss(nStreams)
    mu = np.concatenate([0.5 * (x + 1.0), [mu0]])
    wt = np.concatenate([0.5 * w, [0.0]])
    L = chi.size - 1
    Pl = _legendre_matrix(L, mu)
    fac = (2 * np.arange(L + 1) + 1) * chi
    Pf = (Pl * fac[:, None]).T @ Pl
    Pb = (Pl * (fac * (-1.0) ** np.arange(L + 1))[:, None]).T @ Pl
    norm = 0.5 * (wt[:, None] * (Pf + Pb)).sum(axis=0)
    Pf, Pb = Pf / norm[None, :], Pb / norm[None, :]
    I = np.eye(mu.size)

    def layer(tau, omega):
        n = max(int(np.ceil(np.log2(max(tau, 1e-300) / (thin * mu.min())))), 0)
        d = tau / 2.0 ** n
        T = np.diag(np.exp(-d / mu)) + wt[:, None] * omega * d * Pf / (2.0 * mu[None, :])
        R = wt[:, None] * omega * d * Pb / (2.0 * mu[None, :])
        for _ in range(n):
            G = np.linalg.solve(I - R @ R, np.concatenate([T, R @ T], axis=1))
            R, T = R + T @ G[:, mu.size:], T @ G[:, :mu.size]
        return R, T
    # R: reflection of the stack seen from above, Rb: seen from below, T: transmission downwards, Tb: upwards
    R, T = layer(*layers[0]); Rb, Tb = R.copy(), T.copy()
    for tau, omega in layers[1:]:
        R2, T2 = layer(tau, omega)
        A = np.linalg.inv(I - Rb @ R2)              # bounces between the stack's underside and the new layer
        B = np.linalg.inv(I - R2 @ Rb)
        R, T, Rb, Tb = (R + Tb @ np.linalg.solve(I - R2 @ Rb, R2 @ T), T2 @ A @ T,
                        R2 + T2 @ A @ Rb @ T2, Tb @ B @ T2)
    lam = wt * mu / (wt * mu).sum()
    Rs = albedo * np.outer(lam, np.ones(mu.size))
    y0 = np.zeros(mu.size); y0[-1] = 1.0
    down = np.linalg.solve(I - Rb @ Rs, T @ y0)
    up = R @ y0 + Tb @ (Rs @ down)
    fup, fdn = up.sum(), down.sum()
    return fup, fdn, 1.0 - fup - (1.0 - albedo) * fdn


def planck(lambda_um, T):
    """Spectral radiance up to the constant factors the Monte Carlo's weights share (EMI:503): only ratios matter."""
    hP, cL, kB = 6.62606957e-34, 2.99792458e+8, 1.3806488e-23
    lam = lambda_um / 1.0e6
    return (2.0 * hP * cL * cL) / (lam ** 5 * (np.exp(hP * cL / (kB * lam * T)) - 1.0))


def thermal_fluxes(layers, chi, lambda_um, surfaceTemp, albedo=0.0, nStreams=64, thin=1.0e-5):
    """Thermal emission of a stack of homogeneous layers, TOP FIRST, each (tau, omega, temperature), over a Lambertian
    surface of emissivity 1 - albedo: doubling and adding WITH SOURCES.  A thin layer d emits 2 pi B (1 - omega) d into
    each hemisphere, isotropically (node shares w_i); two identical layers combine to S2 = S + T (1 - R)^-1 S; a layer
    2 added under a stack 1 gives D = (1 - Rb1 R2)^-1 (Sd1 + Rb1 Su2), U = Su2 + R2 D, Su = Su1 + Tb1 U, Sd = Sd2 + T2 D.
    Returns the Monte Carlo's normalised quantities (per emitted photon): (fracAtmsPower, Fup at the top, Fdown at
    the surface, absorbed minus emitted in the atmosphere)."""
    chi = np.asarray(chi, dtype=np.float64)
    x, w = np.polynomial.legendre.leggauss(nStreams)
    mu, wt = 0.5 * (x + 1.0), 0.5 * w
    L = chi.size - 1
    Pl = _legendre_matrix(L, mu)
    fac = (2 * np.arange(L + 1) + 1) * chi
    Pf = (Pl * fac[:, None]).T @ Pl
    Pb = (Pl * (fac * (-1.0) ** np.arange(L + 1))[:, None]).T @ Pl
    norm = 0.5 * (wt[:, None] * (Pf + Pb)).sum(axis=0)
    Pf, Pb = Pf / norm[None, :], Pb / norm[None, :]
    I = np.eye(mu.size)

    def layer(tau, omega, T):
        n = max(int(np.ceil(np.log2(max(tau, 1e-300) / (thin * mu.min())))), 0)
        d = tau / 2.0 ** n
        Tm = np.diag(np.exp(-d / mu)) + wt[:, None] * omega * d * Pf / (2.0 * mu[None, :])
        R = wt[:, None] * omega * d * Pb / (2.0 * mu[None, :])
        S = 2.0 * np.pi * wt * planck(lambda_um, T) * (1.0 - omega) * d
        for _ in range(n):
            S = S + Tm @ np.linalg.solve(I - R, S)
            G = np.linalg.solve(I - R @ R, np.concatenate([Tm, R @ Tm], axis=1))
            R, Tm = R + Tm @ G[:, mu.size:], Tm @ G[:, :mu.size]
        return R, Tm, S
    R, T, S = layer(*layers[0])
    Rb, Tb, Su, Sd = R.copy(), T.copy(), S.copy(), S.copy()
    for tau, omega, temp in layers[1:]:
        R2, T2, S2 = layer(tau, omega, temp)
        D = np.linalg.solve(I - Rb @ R2, Sd + Rb @ S2)
        U = S2 + R2 @ D
        Su, Sd = Su + Tb @ U, S2 + T2 @ D
        A = np.linalg.inv(I - Rb @ R2); B = np.linalg.inv(I - R2 @ Rb)
        R, T, Rb, Tb = (R + Tb @ np.linalg.solve(I - R2 @ Rb, R2 @ T), T2 @ A @ T, R2 + T2 @ A @ Rb @ T2, Tb @ B @ T2)
    lam = wt * mu / (wt * mu).sum()
    eps = 1.0 - albedo
    ys = np.pi * eps * planck(lambda_um, surfaceTemp) * lam
    Rs = albedo * np.outer(lam, np.ones(mu.size))
    Ds = np.linalg.solve(I - Rb @ Rs, Sd + Rb @ ys)
    Us = ys + Rs @ Ds
    fup, fdn = (Su + Tb @ Us).sum(), Ds.sum()
    eAtm = sum(4.0 * np.pi * planck(lambda_um, t) * (1.0 - o) * tau for tau, o, t in layers)
    eSfc = ys.sum()
    tot = eAtm + eSfc
    return eAtm / tot, fup / tot, fdn / tot, (eSfc - fup - eps * fdn) / tot
