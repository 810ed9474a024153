"""Phase-table builders (inversePhaseFunctions.f95:66-174, opticalProperties.f95:1872-2050)."""
import numpy as np

from mcbrat3d_b200.inversePhaseFunctions import computeInversePhaseFunction, computeInversePhaseFuncTable
from mcbrat3d_b200.numericUtilities import computeLegendrePolynomials, computeLobattoTerms
from mcbrat3d_b200.opticalProperties import computeHybridPhaseFunctions
from mcbrat3d_b200.scatteringPhaseFunctions import (getPhaseFunctionValues, henyeyGreenstein, new_PhaseFunction,
                                                    new_PhaseFunctionTable, rayleigh)


def test_lobatto_nodes_and_weights():
    for n in (2, 3, 8, 64, 299):
        mus, w = computeLobattoTerms(n)
        assert mus[0] == -1 and mus[-1] == 1 and np.all(np.diff(mus) > 0)
        assert abs(w.sum() - 2.0) < 1e-4
        assert np.allclose(mus, -mus[::-1], atol=1e-6)
        k = min(2 * n - 3, 12)                                 # exact for polynomials up to 2n-3
        assert abs(np.dot(w.astype(float), mus.astype(float) ** (k - k % 2)) - 2.0 / (k - k % 2 + 1)) < 2e-4


def test_legendre_recursion():
    mus = np.linspace(-1, 1, 11).astype(np.float32)
    P = computeLegendrePolynomials(4, mus)
    m = mus.astype(float)
    assert np.allclose(P[2], 0.5 * (3 * m ** 2 - 1), atol=1e-6)
    assert np.allclose(P[4], (35 * m ** 4 - 30 * m ** 2 + 3) / 8, atol=1e-6)


def test_inverse_table_hg():
    g = 0.85
    T = computeInversePhaseFunction(henyeyGreenstein(g, 64), 10001)
    assert T.dtype == np.float32 and T.shape == (10001,)
    assert abs(T[0] - np.pi) < 1e-6 and T[-1] == 0.0          # back-scatter first, forward last (INV:168)
    assert np.all(np.diff(T) <= 1e-6)
    # sampling the table uniformly reproduces the asymmetry parameter of the (truncated) HG
    assert abs(np.cos(T.astype(float)).mean() - g) < 5e-3


def test_inverse_table_isotropic_and_rayleigh():
    T = computeInversePhaseFunction(new_PhaseFunction(legendreCoefficients=np.zeros(0, np.float32)), 9001)
    p = np.arange(9001) / 9000.0
    assert np.allclose(np.cos(T.astype(float)), -1 + 2 * p, atol=2e-4)     # isotropic: mu uniform
    Tr = computeInversePhaseFunction(rayleigh(), 9001)
    assert abs(np.cos(Tr.astype(float)).mean()) < 2e-3
    # Reference quirk: a Legendre phase function is sampled on max(nMoments, 2) Lobatto nodes
    # (INV:107-112), so the 2-moment Rayleigh function is inverted from a 2-point CDF and its
    # inverse table is the isotropic one (<mu^2> = 1/3, not the 2/5 of true Rayleigh scattering).
    assert abs((np.cos(Tr.astype(float)) ** 2).mean() - 1.0 / 3.0) < 2e-3


def test_inverse_table_tabulated_phase_function():
    ang = np.linspace(0, np.pi, 361).astype(np.float32); ang[-1] = np.float32(np.pi)
    mu = np.cos(ang.astype(float))
    g = 0.6
    val = ((1 - g * g) / (1 + g * g - 2 * g * mu) ** 1.5).astype(np.float32)
    pf = new_PhaseFunction(scatteringAngle=ang, value=val)
    T = computeInversePhaseFuncTable(new_PhaseFunctionTable([pf, henyeyGreenstein(g, 48)], key=[1.0, 2.0]), 9001)
    assert T.shape == (2, 9001)
    assert abs(np.cos(T[0].astype(float)).mean() - g) < 5e-3
    assert abs(np.cos(T[1].astype(float)).mean() - g) < 5e-3


def test_forward_values_normalised():
    ang = (np.arange(9001).astype(np.float32) / np.float32(9000) * np.float32(np.pi)).astype(np.float32)
    for pf in (henyeyGreenstein(0.85, 64), rayleigh()):
        v = getPhaseFunctionValues(pf, ang).astype(float)
        mu = np.cos(ang.astype(float))
        integral = np.sum(0.5 * (v[1:] + v[:-1]) * (mu[:-1] - mu[1:]))
        assert abs(integral - 2.0) < 2e-3
    assert np.all(getPhaseFunctionValues(new_PhaseFunction(legendreCoefficients=np.zeros(0, np.float32)), ang) == 0.5)


def test_hybrid_phase_function_keeps_normalisation_and_tail():
    """computeHybridPhaseFunctions (OPT:1936-2009) hunts UPWARD from the Gaussian width for the
    angle where the normalised Gaussian meets the phase function; it finds one for Mie-like
    functions with a narrow diffraction peak and leaves smooth functions (HG g=0.85) untouched."""
    n = 9001
    ang = (np.arange(n).astype(np.float32) / np.float32(n - 1) * np.float32(np.pi)).astype(np.float32)
    smooth = getPhaseFunctionValues(henyeyGreenstein(0.85, 64), ang)[None, :]
    assert np.array_equal(computeHybridPhaseFunctions(ang, smooth, 7.0), smooth)
    mu = np.cos(ang.astype(float))
    hg = lambda g: (1 - g * g) / (1 + g * g - 2 * g * mu) ** 1.5
    v = (0.5 * hg(0.995) + 0.5 * hg(0.6)).astype(np.float32)[None, :]
    h = computeHybridPhaseFunctions(ang, v, 7.0)
    assert h.shape == v.shape
    changed = np.nonzero(h[0] != v[0])[0]
    assert changed.size > 0
    t = changed.max() + 1
    assert t < n // 4 and np.array_equal(h[0, t:], v[0, t:])
    integ = lambda f: np.sum(0.5 * (f[1:] + f[:-1]) * (mu[:-1] - mu[1:]))
    assert abs(integ(h[0].astype(float)) - integ(v[0].astype(float))) < 5e-3
    assert h[0, 0] < v[0, 0]                                    # the forward peak is flattened


def test_oracle_inverse_builder_matches_numpy_mirror(orc):
    """INV:113-168 restated in C (oracle) and vectorised in NumPy (host mirror): identical tables."""
    from mcbrat3d_b200 import domains
    from mcbrat3d_b200.inversePhaseFunctions import computeInversePhaseFunction, inversion_inputs
    d, _ = domains.landsat_cloud(ssa=0.99, nxy=16, mie=True)
    n = 0
    for tab in d.forwardTables:
        for pf in tab.phaseFunctions[::4]:
            mus, vals = inversion_inputs(pf)
            assert np.array_equal(orc.inverse_phase_function(mus, vals, 9001), computeInversePhaseFunction(pf, 9001))
            n += 1
    assert n >= 4
    # a CDF that decreases somewhere takes the reference's sequential hunt in both
    mus = np.linspace(-1, 1, 9).astype(np.float32)
    vals = np.array([1, 2, -1.5, 0.5, 3, 1, 0.2, 2, 4], dtype=np.float32)
    from mcbrat3d_b200.inversePhaseFunctions import find_cdf_brackets
    cdf = np.zeros(9, np.float32)
    for i in range(1, 9):
        cdf[i] = np.float32(cdf[i - 1] + (mus[i] - mus[i - 1]) * np.float32(0.5) * (vals[i] + vals[i - 1]))
    cdf = (cdf / cdf[-1]).astype(np.float32)
    assert np.any(np.diff(cdf) < 0)
    assert find_cdf_brackets(cdf, 101).min() >= 1


def test_oracle_forward_builder_matches_numpy_mirror(orc):
    """OPT:1912-1913 + SPF:480-498 + NUM:187-205 restated in C (oracle) and in NumPy (host mirror): identical tables."""
    from mcbrat3d_b200 import domains
    for make in (lambda: domains.step_cloud(), lambda: domains.irregular_test_domain()):
        d, _ = make()
        d.tabulateForwardPhaseFunctions(9001)
        for c, tab in enumerate(d.forwardTables):
            for e, pf in enumerate(tab.phaseFunctions):
                assert np.array_equal(orc.forward_phase_function(pf.legendreCoefficients, 9001), d.tabulatedOrigPhaseFunctions[c][e])
    assert np.all(orc.forward_phase_function(np.zeros(0, np.float32), 11) == 0.5)          # quirk q14


def test_oracle_lobatto_values_and_hybrid_equal_the_host_mirror(orc):
    """The C restatements of the remaining table producers -- computeLobattoTerms (NUM:27-114), getPhaseFunctionValues
    for both storage kinds (SPF:448-531), the inversion inputs of INV:97-112 and computeHybridPhaseFunctions
    (OPT:1936-2050) -- against the NumPy mirror, bit for bit."""
    from mcbrat3d_b200.inversePhaseFunctions import inversion_inputs
    from mcbrat3d_b200.numericUtilities import computeLobattoTerms
    from mcbrat3d_b200.opticalProperties import computeHybridPhaseFunctions
    from mcbrat3d_b200.scatteringPhaseFunctions import (getPhaseFunctionValues, henyeyGreenstein, new_PhaseFunction, rayleigh)
    f32 = np.float32
    for n in (2, 3, 4, 5, 16, 64, 65, 128, 299):
        mus, w = orc.lobatto_terms(n)
        m2, w2 = computeLobattoTerms(n)
        assert np.array_equal(mus, m2) and np.array_equal(w, w2), n
        assert mus[0] == -1 and mus[-1] == 1 and np.all(np.diff(mus) > 0) and abs(float(w.sum()) - 2.0) < 1e-4
    for pf in (henyeyGreenstein(0.85, 64), henyeyGreenstein(0.85, 299), rayleigh(), henyeyGreenstein(0.6, 3), henyeyGreenstein(0.3, 1)):
        mus, vals = orc.inversion_inputs_legendre(pf.legendreCoefficients)
        m2, v2 = inversion_inputs(pf)
        assert np.array_equal(mus, m2) and np.array_equal(vals, v2)
    nS = 9001
    angles = (np.arange(nS, dtype=f32) / f32(nS - 1) * f32(np.pi)).astype(f32)
    ang = np.linspace(0.0, np.pi, 721).astype(f32); ang[-1] = f32(np.pi)
    mu = np.cos(ang.astype(np.float64))
    hg = lambda gg: (1 - gg * gg) / (1 + gg * gg - 2 * gg * mu) ** 1.5
    tab = new_PhaseFunction(scatteringAngle=ang, value=(0.97 * hg(0.9) + 0.03 * hg(-0.45)).astype(f32))
    a = orc.phase_function_values(angles, storedAngle=tab.scatteringAngle, storedValue=tab.value)
    assert np.array_equal(a, getPhaseFunctionValues(tab, angles))
    found = 0
    for pf, width in ((henyeyGreenstein(0.9, 64), 7.0), (henyeyGreenstein(0.95, 256), 7.0), (henyeyGreenstein(0.85, 64), 7.0),
                      (tab, 7.0), (tab, 2.0)):
        orig = getPhaseFunctionValues(pf, angles)
        want = computeHybridPhaseFunctions(angles, orig[None, :], width)[0]
        got, t = orc.hybrid_phase_function(angles, orig, width)
        assert np.array_equal(got, want), (width, t)
        if t > 0:                                        # Gaussian below the transition, the original above, still normalised
            found += 1
            assert np.array_equal(got[t:], orig[t:]) and not np.array_equal(got[:t], orig[:t])
            cs = np.cos(angles.astype(np.float64))
            integral = float(np.sum(0.5 * (got[:-1] + got[1:]).astype(np.float64) * (cs[:-1] - cs[1:])))
            assert abs(integral - 2.0) < 2e-3, integral
    assert found >= 3
