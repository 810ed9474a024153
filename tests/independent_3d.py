"""An INDEPENDENT deterministic answer for the first interaction of a collimated beam with a 3-D voxel medium that is
periodic in x and y — the part of the photon path that has a closed form in three dimensions.

Test infrastructure.  Nothing here shares code, structure or arithmetic with ``oracle/`` or with the CUDA kernels:
there is no cell-to-cell marcher (OPT:1656-1815 steps from face to face and fixes positions up with ``spacing()``);
instead every plane crossing of a straight line is computed in one go in double precision, the crossing parameters are
SORTED, and the cell of each piece is read off its mid-point.  No random numbers are used: the averages over entry /
exit points are mid-point rules on a regular sub-grid of every column.

What it returns, for a medium given by raw arrays (cell edges, per-component extinction, single-scattering albedo and
Legendre coefficients):

* ``first_collision``: the probability that a photon entering uniformly at the top along the solar direction has its
  first collision in a given cell, and the probability that it reaches the surface uncollided in a given column
  (Beer's law along slant paths through a heterogeneous, periodically continued medium);
* ``first_order_radiance``: the expected local-estimate contribution of FIRST-order scattering (and of the direct
  beam reflected by the Lambertian surface) per photon, per exit column and view direction:
  with exit point e in the column, s the path along the view line through e,
      E1 = (area_col / A) * mean_e  int ds  sum_c sigma_c(s) omega_c(s) P_c(Theta) / (4 pi mu0)
                                            * exp(-tau_sun(s)) * exp(-tau_view(s)),
      E0 = (area_col / A) * mean_e  albedo / pi * exp(-tau_sun(surface point)) * exp(-tau_view(surface point -> e)),
  which is what INT:1623-1752 tallies at the first scattering (photon weight omega_c, phase function over
  4 pi |mu_view|, transmission to the boundary; exit column = where the view ray leaves) divided by the number of
  photons — derived from the transfer equation, not from the code.
"""
from __future__ import annotations

import numpy as np
from numpy.polynomial import legendre as _leg


def direction(mu, phiDeg, down=False):
    """Unit vector with |z| = mu and azimuth phi; ``down`` for the solar beam (travelling towards the surface)."""
    s = np.sqrt(max(0.0, 1.0 - mu * mu))
    p = np.deg2rad(phiDeg)
    return np.array([s * np.cos(p), s * np.sin(p), -abs(mu) if down else mu], dtype=np.float64)


def phase_value(legendreCoefficients, cosTheta):
    """P(Theta) = sum_l (2l + 1) chi_l P_l(cos Theta), chi_0 = 1 (normalised to 4 pi over the sphere)."""
    chi = np.concatenate([[1.0], np.asarray(legendreCoefficients, dtype=np.float64)])
    return float(_leg.legval(cosTheta, chi * (2.0 * np.arange(chi.size) + 1.0)))


class Medium:
    def __init__(self, xE, yE, zE, ext, ssa, phaseIdx, phaseTables, albedo=0.0):
        """ext, ssa: (nc, nz, ny, nx); phaseIdx: (nc, nz, ny, nx) 1-based entry into phaseTables[c] (a list of Legendre
        coefficient arrays), 0 where the component is absent."""
        self.xE, self.yE, self.zE = (np.asarray(a, dtype=np.float64) for a in (xE, yE, zE))
        self.nx, self.ny, self.nz = self.xE.size - 1, self.yE.size - 1, self.zE.size - 1
        self.ext = np.asarray(ext, dtype=np.float64)
        self.ssa = np.asarray(ssa, dtype=np.float64)
        self.idx = np.asarray(phaseIdx)
        self.tables = phaseTables
        self.albedo = float(albedo)
        self.sigma = self.ext.sum(axis=0).ravel()                # total extinction, flat z-major / x-fastest
        self.Lx, self.Ly = self.xE[-1] - self.xE[0], self.yE[-1] - self.yE[0]
        self.area = np.outer(np.diff(self.yE), np.diff(self.xE)) / (self.Lx * self.Ly)      # (ny, nx), sums to 1

    # ------------------------------------------------------------------------------------------------------------
    def _planes(self, E, L, p0, d, T):
        """Crossing parameters of the periodically continued planes of one horizontal axis, clipped into [0, T]."""
        if d == 0.0:
            return np.zeros((p0.size, 0))
        m = int(np.ceil(abs(d) * T.max() / L)) + 1
        planes = (E[:-1][None, :] + L * np.arange(-m, m + 1)[:, None]).ravel()
        return np.clip((planes[None, :] - p0[:, None]) / d, 0.0, T[:, None])

    def pieces(self, p0, d):
        """Straight lines from ``p0`` (N, 3) along the common direction ``d`` until they leave the slab through the top
        or the bottom.  Returns (breakpoints t (N, K + 1) ascending, flat cell index of every piece (N, K), T)."""
        p0 = np.array(p0, dtype=np.float64)
        p0[:, 0] = self.xE[0] + np.mod(p0[:, 0] - self.xE[0], self.Lx)
        p0[:, 1] = self.yE[0] + np.mod(p0[:, 1] - self.yE[0], self.Ly)
        if d[2] > 0:
            T = (self.zE[-1] - p0[:, 2]) / d[2]
        else:
            T = (self.zE[0] - p0[:, 2]) / d[2]
        tz = np.clip((self.zE[None, :] - p0[:, 2:3]) / d[2], 0.0, T[:, None])
        t = np.concatenate([np.zeros((p0.shape[0], 1)), tz, self._planes(self.xE, self.Lx, p0[:, 0], d[0], T),
                            self._planes(self.yE, self.Ly, p0[:, 1], d[1], T), T[:, None]], axis=1)
        t.sort(axis=1)
        tm = 0.5 * (t[:, 1:] + t[:, :-1])
        return t, self.cell_of(p0[:, None, :] + tm[:, :, None] * d[None, None, :]), T

    def column_of(self, x, y):
        ix = np.clip(np.searchsorted(self.xE, self.xE[0] + np.mod(x - self.xE[0], self.Lx), side="right") - 1, 0, self.nx - 1)
        iy = np.clip(np.searchsorted(self.yE, self.yE[0] + np.mod(y - self.yE[0], self.Ly), side="right") - 1, 0, self.ny - 1)
        return ix, iy

    def cell_of(self, p):
        ix, iy = self.column_of(p[..., 0], p[..., 1])
        iz = np.clip(np.searchsorted(self.zE, p[..., 2], side="right") - 1, 0, self.nz - 1)
        return ix + self.nx * (iy + self.ny * iz)

    def optical_depth(self, p0, d, chunk=100000):
        """Optical depth from every p0 to the boundary along d."""
        out = np.empty(p0.shape[0])
        for i in range(0, p0.shape[0], chunk):
            t, cell, _ = self.pieces(p0[i:i + chunk], d)
            out[i:i + chunk] = (np.diff(t, axis=1) * self.sigma[cell]).sum(axis=1)
        return out

    def sub_grid(self, m, z):
        """m x m mid-points in every column at height z: (points (ncol * m * m, 3), column index of every point)."""
        f = (np.arange(m) + 0.5) / m
        xs = (self.xE[:-1, None] + np.diff(self.xE)[:, None] * f[None, :]).ravel()            # nx * m
        ys = (self.yE[:-1, None] + np.diff(self.yE)[:, None] * f[None, :]).ravel()
        X, Y = np.meshgrid(xs, ys)                                                            # (ny m, nx m)
        col = (np.arange(self.nx * m) // m)[None, :] + self.nx * (np.arange(self.ny * m) // m)[:, None]
        p = np.stack([X.ravel(), Y.ravel(), np.full(X.size, z)], axis=1)
        return p, col.ravel()

    # ------------------------------------------------------------------------------------------------------------
    def first_collision(self, mu0, phi0, m=24):
        """(probability of the first collision per cell (nz, ny, nx), probability of reaching the surface uncollided per
        column (ny, nx)).  Lines are parametrised by their LANDING point, so the column sums are exact mid-point
        rules of a continuous integrand."""
        sun = direction(mu0, phi0, down=True)
        p, col = self.sub_grid(m, self.zE[0])
        w = self.area.ravel()[col] / (m * m)
        t, cell, _ = self.pieces(p, -sun)                            # from the landing point up to the top
        dtau = np.diff(t, axis=1) * self.sigma[cell]
        above = dtau[:, ::-1].cumsum(axis=1)[:, ::-1] - dtau         # optical depth between the top and the piece
        dep = np.exp(-above) * -np.expm1(-dtau) * w[:, None]
        first = np.bincount(cell.ravel(), weights=dep.ravel(), minlength=self.sigma.size).reshape(self.nz, self.ny, self.nx)
        surf = np.bincount(col, weights=w * np.exp(-dtau.sum(axis=1)), minlength=self.nx * self.ny).reshape(self.ny, self.nx)
        return first, surf

    def _source(self, cosTheta):
        """sum_c sigma_c omega_c P_c(Theta) per cell (flat)."""
        S = np.zeros(self.sigma.size)
        for c in range(self.ext.shape[0]):
            P = np.array([0.0] + [phase_value(lc, cosTheta) for lc in self.tables[c]])
            S += (self.ext[c] * self.ssa[c] * P[self.idx[c]]).ravel()
        return S

    def first_order_radiance(self, mu0, phi0, muV, phiV, m=6, gauss=3, sub=1):
        """(E1, E0), each (ny, nx): expected first-order and surface-reflected direct-beam local-estimate contribution
        per photon, by exit column of the view ray (module docstring)."""
        sun = direction(mu0, phi0, down=True)
        view = direction(muV, phiV)
        S = self._source(float(sun @ view))
        zExit = self.zE[-1] if view[2] > 0 else self.zE[0]
        e, col = self.sub_grid(m, zExit)
        w = self.area.ravel()[col] / (m * m)
        t, cell, T = self.pieces(e, -view)                           # from the exit point back along the line of sight
        dt = np.diff(t, axis=1)
        dtau = dt * self.sigma[cell]
        before = dtau.cumsum(axis=1) - dtau                          # view optical depth between the exit and the piece
        gx, gw = np.polynomial.legendre.leggauss(gauss)
        # composite rule: ``sub`` equal parts per piece (tau_sun has kinks along the line of sight wherever the solar
        # ray through the point passes a cell edge, so more parts pay better than a higher order)
        gx = ((np.arange(sub)[:, None] + 0.5 * (gx[None, :] + 1.0)) / sub).ravel()
        gw = np.tile(0.5 * gw / sub, sub)
        r, k = np.nonzero(dt > 1e-14 * T[:, None])                   # the pieces that exist (the rest is padding)
        line = np.zeros(e.shape[0])
        for a, b in zip(gx, gw):                                     # Gauss points inside every piece
            q = e[r] - (t[r, k] + a * dt[r, k])[:, None] * view[None, :]
            tauSun = self.optical_depth(q, -sun)
            line += np.bincount(r, weights=b * dt[r, k] * S[cell[r, k]] * np.exp(-(before[r, k] + a * dtau[r, k]) - tauSun),
                                minlength=e.shape[0])
        E1 = np.bincount(col, weights=w * line / (4.0 * np.pi * abs(mu0)), minlength=self.nx * self.ny)
        # the direct beam reflected by the Lambertian surface (INT:1688-1694: weight * albedo / pi, no 1 / mu_view).  A
        # downward view "ray" leaves through the surface at once: the reference tallies albedo / pi there (tau = 0).
        foot = e - T[:, None] * view[None, :] if view[2] > 0 else e.copy()
        foot[:, 2] = self.zE[0]
        tauView = dtau.sum(axis=1) if view[2] > 0 else 0.0
        E0 = np.bincount(col, weights=w * self.albedo / np.pi * np.exp(-self.optical_depth(foot, -sun) - tauView),
                         minlength=self.nx * self.ny)
        return E1.reshape(self.ny, self.nx), E0.reshape(self.ny, self.nx)


# ---- thermal emission: where photons are born and what they contribute at birth (zeroth order) ----------------------------
def planck(lambda_um, T):
    """Spectral radiance of a black body (any constant factor cancels: only ratios of powers are used)."""
    h, c, k = 6.62606957e-34, 2.99792458e8, 1.3806488e-23
    lam = lambda_um * 1.0e-6
    return 2.0 * h * c * c / lam ** 5 / np.expm1(h * c / (lam * k * np.asarray(T, dtype=np.float64)))


def thermal_source(med, temps, lambda_um, sfcTemp):
    """Kirchhoff: a cell emits 4 pi kappa_abs B(T) V isotropically, the Lambertian surface pi (1 - albedo) B(T_s) A.
    Returns (probability that a photon is born in each cell (nz, ny, nx), probability that it is born at the surface,
    kappa_abs B per cell / p and emissivity B_s / p, with p the total power per unit area)."""
    kappa = (med.ext * (1.0 - med.ssa)).sum(axis=0)                        # (nz, ny, nx), per km
    B = planck(lambda_um, temps)
    dz = np.diff(med.zE)[:, None, None]
    cellPower = 4.0 * np.pi * kappa * B * dz * med.area[None, :, :]          # per unit domain area
    sfcPower = np.pi * (1.0 - med.albedo) * float(planck(lambda_um, sfcTemp))
    p = cellPower.sum() + sfcPower
    return cellPower / p, sfcPower / p, kappa * B / p, (1.0 - med.albedo) * float(planck(lambda_um, sfcTemp)) / p


def emission_radiance(med, temps, lambda_um, sfcTemp, muV, phiV, m=24):
    """Expected local-estimate contribution AT BIRTH per photon, by exit column (ny, nx): the emission integral
    int kappa_abs B exp(-tau) ds along the line of sight plus the transmitted surface emission, over the total power
    (INT:513-542, 1695-1696: isotropic emission contributes weight / (4 pi |mu|) exp(-tau), the surface weight / pi
    exp(-tau); a downward view "ray" from the surface leaves at once, tau = 0).  Every piece is integrated exactly
    (source and extinction are constant in a cell)."""
    _, _, src, sfc = thermal_source(med, temps, lambda_um, sfcTemp)
    src = src.ravel()
    view = direction(muV, phiV)
    e, col = med.sub_grid(m, med.zE[-1] if view[2] > 0 else med.zE[0])
    w = med.area.ravel()[col] / (m * m)
    t, cell, _ = med.pieces(e, -view)
    dt = np.diff(t, axis=1)
    sig = med.sigma[cell]
    dtau = dt * sig
    before = dtau.cumsum(axis=1) - dtau
    seg = np.where(sig > 0, -np.expm1(-dtau) / np.where(sig > 0, sig, 1.0), dt)      # int_0^L exp(-sigma s) ds
    line = (src[cell] * np.exp(-before) * seg).sum(axis=1)
    line += sfc * (np.exp(-dtau.sum(axis=1)) if view[2] > 0 else 1.0)
    return np.bincount(col, weights=w * line, minlength=med.nx * med.ny).reshape(med.ny, med.nx)
