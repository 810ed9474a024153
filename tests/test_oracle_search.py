"""findIndex / findCDFIndex (numericUtilities.f95:206-348): oracle vs host mirror vs brute force."""
import ctypes as C

import numpy as np

from mcbrat3d_b200.numericUtilities import findCDFIndex, findIndex


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_findIndex_matches_definition(orc):
    lib = orc.load()
    rng = np.random.default_rng(0)
    for n in (2, 3, 5, 17, 120):
        table = np.sort(rng.random(n)) * 10.0
        for v in np.concatenate([rng.random(50) * 12.0 - 1.0, table]):
            want = int(np.searchsorted(table, v, side="right"))          # table(i) <= v < table(i+1)
            # without a first guess the bisection starts at (0, n) and can return at most n-1
            assert lib.orc_findIndexDouble(float(v), _p(table), n, 0) == min(want, n - 1)
            assert findIndex(v, table) == min(want, n - 1)
            if v < table[0]:
                continue        # the hunt never terminates below the table, in the reference too (NUM:224-237)
            for guess in (1, n // 2 + 1, n):
                got = lib.orc_findIndexDouble(float(v), _p(table), n, guess)
                assert got == findIndex(v, table, firstGuess=guess)
                if table[0] <= v < table[-1] and guess < n:     # guess == n returns n at once (NUM:224-226)
                    assert got == want


def test_findIndex_component_pick(orc):
    """component = findIndex(RN, (/ 0, cumExt(:) /)) (INT:759-760) incl. RN = 0 and RN = 1."""
    lib = orc.load()
    table = np.array([0.0, 0.25, 0.6, 1.0])
    cases = {0.0: 1, 0.2: 1, 0.25: 2, 0.5: 2, 0.6: 3, 0.99: 3, 1.0: 3}
    for v, want in cases.items():
        assert lib.orc_findIndexMixed(C.c_float(v), _p(table), 4, 0) == want
    one = np.array([0.0, 1.0])
    for v in (0.0, 0.5, 1.0):
        assert lib.orc_findIndexMixed(C.c_float(v), _p(one), 2, 0) == 1


def test_findCDFIndex(orc):
    lib = orc.load()
    rng = np.random.default_rng(1)
    for n in (1, 2, 7, 64):
        cdf = np.cumsum(rng.random(n)); cdf /= cdf[-1]; cdf[-1] = 1.0
        for v in np.concatenate([rng.random(64).astype(np.float32), np.float32([0.0, 1.0]), cdf.astype(np.float32)]):
            want = max(int(np.searchsorted(cdf, np.float64(v), side="left")) + 1, 1)   # table(i-1) < v <= table(i)
            want = min(want, n)
            assert lib.orc_findCDFIndex(C.c_float(v), _p(cdf), n, 1) == want
            assert findCDFIndex(np.float64(v), cdf) == want
    # strided access = the colWeights / levelWeights pointer slices (EMI:56-57)
    a = np.cumsum(rng.random(40)).reshape(8, 5)
    col = np.ascontiguousarray(a[:, 4])
    for v in (0.1, 3.0, 9.0):
        assert lib.orc_findCDFIndex(C.c_float(v), _p(a.ravel()[4:]), 8, 5) == lib.orc_findCDFIndex(C.c_float(v), _p(col), 8, 1)
