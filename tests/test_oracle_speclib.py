"""Known-answer tests for the oracle's speclib restatement (oracle/speclib.c).  The Fortran cannot be
compiled here, so the pins are mathematical: closed-form GLL nodes/weights for small n, numpy Legendre
roots for every n, exactness of D on polynomials, interpolation property of hgll."""
import numpy as np
import pytest
from oracle import capi as c


def test_closed_forms():
    z, w = c.zwgll(2)
    assert np.array_equal(z, [-1.0, 1.0]) and np.allclose(w, [1.0, 1.0], atol=0, rtol=1e-15)
    z, w = c.zwgll(3)
    assert np.allclose(z, [-1, 0, 1], atol=1e-16) and np.allclose(w, [1 / 3, 4 / 3, 1 / 3], rtol=1e-15)
    z, w = c.zwgll(4)
    assert np.allclose(z, [-1, -1 / np.sqrt(5), 1 / np.sqrt(5), 1], rtol=1e-15)
    assert np.allclose(w, [1 / 6, 5 / 6, 5 / 6, 1 / 6], rtol=1e-15)
    z, w = c.zwgll(5)
    assert np.allclose(z, [-1, -np.sqrt(3 / 7), 0, np.sqrt(3 / 7), 1], atol=2e-16)
    assert np.allclose(w, [1 / 10, 49 / 90, 32 / 45, 49 / 90, 1 / 10], rtol=1e-15)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16])
def test_against_legendre(n):
    N = n - 1
    z, w = c.zwgll(n)
    P = np.polynomial.legendre.Legendre.basis(N)
    ref = np.concatenate([[-1.0], np.sort(P.deriv().roots().real), [1.0]]) if N > 1 else np.array([-1.0, 1.0])
    assert np.abs(z - ref).max() < 5e-15
    assert np.abs(w - 2.0 / (N * (N + 1) * P(ref) ** 2)).max() < 5e-15
    D = c.dgll(z.copy(), n)
    for k in range(n):  # D differentiates x^k exactly for k <= N
        exact = k * z ** (k - 1) if k > 0 else np.zeros(n)
        assert np.abs(D @ z ** k - exact).max() < 1e-11
    assert np.abs(D.sum(axis=1)).max() < 1e-12  # constants in the null space


def test_hgll_interpolation():
    for nf, nc in [(8, 5), (8, 2), (5, 2), (10, 7), (4, 2)]:
        zf, _ = c.zwgll(nf)
        zc, _ = c.zwgll(nc)
        J = np.array([[c.hgll(j + 1, zf[i], zc.copy(), nc) for j in range(nc)] for i in range(nf)])
        assert np.abs(J.sum(axis=1) - 1).max() < 1e-13          # partition of unity
        for k in range(nc):                                      # exact on polynomials of degree < nc
            assert np.abs(J @ zc ** k - zf ** k).max() < 1e-12
        # cardinality at coincident nodes (end points)
        assert J[0, 0] == 1.0 and J[-1, -1] == 1.0


def test_glibc_rand_is_libc():
    """o_rand_fill calls the real libc rand(); its first values for seed 1 are the documented ones."""
    import ctypes as C
    a = np.zeros(4)
    c.lib().o_rand_fill(c.ptr(a), 4, C.c_uint(1))
    assert np.array_equal((a * 2147483647).round().astype(np.int64), [1804289383, 846930886, 1681692777, 1714636915])
