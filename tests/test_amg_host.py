"""Host-side AMG setup of the product (csrc/host/amg.hpp, C++) against the oracle's restatement
(oracle/amg.py, numpy/scipy): identical C/F splittings on every level, interpolation and Galerkin operators
to 1e-12, Chebyshev data to 1e-10, same number of levels.  No GPU needed."""
import ctypes as C
import numpy as np
import pytest
import scipy.sparse as sp
from oracle import amg as oamg


def product_hierarchy(prfdd, A, cheby_order=2, max_coarse=9, coarsening=-1):
    L = prfdd.lib()
    A = A.tocsr(); A.sort_indices()
    ptr, col, val = A.indptr.astype(np.int32), A.indices.astype(np.int32), np.ascontiguousarray(A.data, dtype=np.float64)
    h = C.c_void_p()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.prfdd_amg_host_setup_ex(C.byref(h), C.c_int(A.shape[0]), vp(ptr), vp(col), vp(val), C.c_int(cheby_order), C.c_int(max_coarse), C.c_int(coarsening)) == 0
    out = []
    for l in range(L.prfdd_amg_host_num_levels(h)):
        sz = (C.c_int * 4)()
        L.prfdd_amg_host_level_sizes(h, C.c_int(l), sz)
        n, nnzA, ncP, nnzP = list(sz)
        ap, ac, av = np.zeros(n + 1, np.int32), np.zeros(nnzA, np.int32), np.zeros(nnzA)
        L.prfdd_amg_host_get_matrix(h, C.c_int(l), C.c_int(0), vp(ap), vp(ac), vp(av))
        lev = dict(n=n, A=sp.csr_matrix((av, ac, ap), shape=(n, n)))
        cf = np.zeros(n if nnzP else 0, np.int8); ds = np.zeros(n); coefs = np.zeros(cheby_order); eigs = (C.c_double * 2)()
        L.prfdd_amg_host_get_vectors(h, C.c_int(l), vp(cf) if nnzP else None, vp(ds), vp(coefs), eigs)
        lev.update(cf=cf, ds=ds, coefs=coefs, eigs=(eigs[0], eigs[1]))
        if nnzP:
            pp, pc, pv = np.zeros(n + 1, np.int32), np.zeros(nnzP, np.int32), np.zeros(nnzP)
            L.prfdd_amg_host_get_matrix(h, C.c_int(l), C.c_int(1), vp(pp), vp(pc), vp(pv))
            lev["P"] = sp.csr_matrix((pv, pc, pp), shape=(n, ncP))
        out.append(lev)
    nlast = out[-1]["n"]
    Ainv = np.zeros(nlast * nlast)
    L.prfdd_amg_host_get_coarse_inverse(h, vp(Ainv))
    L.prfdd_amg_host_destroy(h)
    return out, Ainv.reshape(nlast, nlast)


def laplace3d(n, aniso=1.0):
    I = sp.eye(n); T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    return (sp.kron(sp.kron(T, I), I) + sp.kron(sp.kron(I, T), I) + aniso * sp.kron(sp.kron(I, I), T)).tocsr()


def random_spd(n, seed):
    rng = np.random.default_rng(seed)
    G = sp.random(n, n, density=6.0 / n, random_state=seed, format="csr")
    G = -(abs(G) + abs(G).T)
    G.setdiag(0); G.eliminate_zeros()
    d = -np.asarray(G.sum(axis=1)).ravel() + rng.uniform(0.01, 0.2, n)
    return (G + sp.diags(d)).tocsr()


@pytest.mark.parametrize("name,A", [("lap3d_12", laplace3d(12)), ("lap3d_aniso", laplace3d(9, 0.01)), ("rand_spd", random_spd(1500, 3))])
@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("coarsening", ["pmis", "hmis"])
def test_hierarchy_matches_oracle(prfdd, name, A, order, coarsening):
    oamg.set_coarsening(coarsening)
    try:
        Ho = oamg.Hierarchy(A, cheby_order=order)
    finally:
        oamg.set_coarsening("hmis")
    Hp, Ainv = product_hierarchy(prfdd, A, order, coarsening=0 if coarsening == "pmis" else 1)
    assert len(Hp) == Ho.num_levels
    for lo, lp in zip(Ho.levels, Hp):
        assert lo.n == lp["n"]
        assert abs(lo.A - lp["A"]).max() <= 1e-12 * abs(lo.A).max()
        assert np.allclose(lo.ds, lp["ds"], rtol=1e-14)
        assert np.allclose(lo.coefs, lp["coefs"], rtol=1e-9), (lo.coefs, lp["coefs"])
        if hasattr(lo, "P"):
            assert np.array_equal(lo.cf, lp["cf"])
            assert lo.P.shape == lp["P"].shape
            assert (lo.P != 0).nnz == (lp["P"] != 0).nnz
            assert abs(lo.P - lp["P"]).max() <= 1e-12
            assert np.asarray(lp["P"].getnnz(axis=1)).max() <= 4
    assert np.abs(Ainv - Ho.levels[-1].Ainv).max() <= 1e-9 * np.abs(Ainv).max()


def test_hmis_is_a_valid_first_pass_colouring():
    """properties of the Ruge-Stueben first pass (HMIS on one process): every F point with strong connections depends strongly on
    at least one C point; no two C points ... are NOT forbidden to be strongly connected in general, but on the 7-point Laplacian
    the pass yields the red-black colouring (half the points, no C-C connection)"""
    A = laplace3d(10)
    S = oamg.strength(A)
    cf = oamg.rs_first_pass(S)
    n = A.shape[0]
    assert set(np.unique(cf)) == {-1, 1}
    isC = cf == 1
    dep_on_C = np.asarray((S.astype(np.float64) @ isC.astype(np.float64))).ravel()
    assert np.all(dep_on_C[~isC] >= 1)
    Sc = S[isC][:, isC]
    assert Sc.nnz == 0 and abs(isC.sum() - n / 2) <= 1


def test_vcycle_contracts():
    """SURVEY 8c pin 5: V-cycle residual contraction < 1 on the recipe of subdomain.tpp:3707-3855 (random u*, f = A u*)."""
    A = laplace3d(14)
    H = oamg.Hierarchy(A, cheby_order=2)
    rng = np.random.default_rng(0)
    f = A @ rng.random(A.shape[0])
    x = np.zeros_like(f); r = f.copy(); hist = [np.linalg.norm(r)]
    for _ in range(8):
        x += H.vcycle(r); r = f - A @ x; hist.append(np.linalg.norm(r))
    rates = np.array(hist[1:]) / np.array(hist[:-1])
    assert rates.max() < 0.35


def test_cheby_coefficients_are_the_chebyshev_residual_polynomial():
    lo, up = 0.3, 1.7
    for order in (1, 2, 3, 4):
        c = oamg.cheby_coefs(lo, up, order)
        x = np.linspace(lo, up, 201)
        res = 1 - x * np.polyval(c[::-1], x)
        theta, delta = (up + lo) / 2, (up - lo) / 2
        Tk = np.polynomial.chebyshev.Chebyshev.basis(order)
        assert np.allclose(res, Tk((theta - x) / delta) / Tk(theta / delta), atol=1e-12)
    # order 2 closed form (hypre par_cheby.c "case 1"): c0 = -4 theta/den, c1 = 2/den, den = delta^2 - 2 theta^2
    theta, delta = (up + lo) / 2, (up - lo) / 2
    den = delta * delta - 2 * theta * theta
    assert np.allclose(oamg.cheby_coefs(lo, up, 2), [-4 * theta / den, 2 / den])
