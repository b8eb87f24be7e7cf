"""Pins oracle/kernels.c against the REFERENCE'S OWN kernels: oracle/_ref/libref_okl_{2,3}d.so is the
four .okl files of the reference compiled for the CPU the way OCCA Serial runs them (oracle/build_ref.py).
Every comparison is bit-for-bit.  Skipped only if oracle/_ref was never built (no /root/reference)."""
import ctypes as C
import numpy as np
import pytest
from oracle import capi as c

P = c.ptr


def _ref(dim):
    r = c.ref(dim)
    if r is None:
        pytest.skip("oracle/_ref not built")
    return r


def _rand(rng, n):
    return np.ascontiguousarray(rng.standard_normal(n))


@pytest.mark.parametrize("dim,N,E", [(2, 3, 5), (2, 7, 3), (3, 2, 4), (3, 7, 2), (3, 4, 3)])
def test_domain_operator(dim, N, E):
    R, L = _ref(dim), c.lib()
    rng = np.random.default_rng(dim * 100 + N)
    n = N + 1
    npts = E * n ** dim
    z, _ = c.zwgll(n)
    D = np.ascontiguousarray(c.dgll(z, n).ravel())
    u = _rand(rng, npts)
    G = [_rand(rng, npts) for _ in range(6)]
    outs = []
    for lib, pre in ((R, "domain_"), (L, "o_")):
        gdu = [np.zeros(npts) for _ in range(dim)]
        Au = np.zeros(npts)
        a1 = [c.ptr_table(gdu), P(u), P(D), c.ptr_table(G), C.c_int(npts), C.c_int(N)]
        a2 = [P(Au), c.ptr_table(gdu), P(D), C.c_int(npts), C.c_int(N)]
        if lib is L:
            a1.append(C.c_int(dim)); a2.append(C.c_int(dim))
        getattr(lib, pre + "stiffness_matrix_1")(*a1)
        getattr(lib, pre + "stiffness_matrix_2")(*a2)
        outs.append((gdu, Au))
    for a, b in zip(outs[0][0], outs[1][0]):
        assert np.array_equal(a, b)
    assert np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000])
def test_domain_vector_kernels(n):
    R, L = _ref(3), c.lib()
    rng = np.random.default_rng(n)
    nb = (n + 127) // 128
    a, b, cc, d, m = (_rand(rng, n) for _ in range(5))
    m = (m > 0).astype(np.float64)
    def both(rname, lname, nout, *args):
        o1, o2 = np.zeros(nout), np.zeros(nout)
        getattr(R, rname)(P(o1), *args)
        getattr(L, lname)(P(o2), *args)
        assert np.array_equal(o1, o2), rname
    both("domain_residual_norm", "o_residual_norm", nb, P(a), P(b), P(m), C.c_int(n), C.c_int(nb))
    both("domain_projection_inner_products", "o_projection_inner_products", 2 * nb, P(a), P(b), P(cc), P(d), C.c_int(n), C.c_int(nb))
    both("domain_inner_product_flexible", "o_inner_product_flexible", nb, P(a), P(b), P(cc), C.c_int(n), C.c_int(nb))
    both("domain_inner_product", "o_inner_product_mask", nb, P(a), P(b), P(m), C.c_int(n), C.c_int(nb))
    both("subdomain_inner_product", "o_sub_inner_product", nb, P(a), P(b), C.c_int(n), C.c_int(nb))
    both("subdomain_weighted_inner_product", "o_sub_weighted_inner_product", nb, P(a), P(b), P(m), C.c_int(n), C.c_int(nb))
    both("subdomain_projection_inner_products", "o_sub_projection_inner_products", 2 * nb, P(a), P(b), P(cc), P(d), P(m), C.c_int(n), C.c_int(nb))
    both("subdomain_search_update_inner_product", "o_sub_search_update_inner_product", nb, P(a), P(b), P(cc), P(m), C.c_int(n), C.c_int(nb))
    # element-wise updates
    for rname, lname in (("domain_solution_and_residual_update", "o_solution_and_residual_update"),
                         ("subdomain_solution_and_residual_update", "o_solution_and_residual_update")):
        u1, u2, r1, r2 = a.copy(), a.copy(), np.zeros(n), np.zeros(n)
        getattr(R, rname)(P(u1), P(r1), P(b), P(cc), P(d), C.c_double(0.37), C.c_int(n))
        getattr(L, lname)(P(u2), P(r2), P(b), P(cc), P(d), C.c_double(0.37), C.c_int(n))
        assert np.array_equal(u1, u2) and np.array_equal(r1, r2)
    p1, p2, r1, r2 = a.copy(), a.copy(), np.zeros(n), np.zeros(n)
    R.domain_residual_and_search_update(P(p1), P(r1), P(b), P(cc), C.c_double(-1.7), C.c_int(n))
    L.o_residual_and_search_update(P(p2), P(r2), P(b), P(cc), C.c_double(-1.7), C.c_int(n))
    assert np.array_equal(p1, p2) and np.array_equal(r1, r2)
    u1, u2, r1, r2 = np.ones(n), np.ones(n), np.zeros(n), np.zeros(n)
    R.domain_initialize_arrays(P(u1), P(r1), P(a), C.c_int(n))
    L.o_initialize_arrays(P(u2), P(r2), P(a), C.c_int(n))
    assert np.array_equal(u1, u2) and np.array_equal(r1, r2)
    # math.okl
    o1, o2 = np.zeros(n), np.zeros(n)
    R.math_vector_vector_addition(P(o1), C.c_double(1.3), P(a), C.c_double(-0.2), P(b), C.c_int(n))
    L.o_vector_vector_addition(P(o2), C.c_double(1.3), P(a), C.c_double(-0.2), P(b), C.c_int(n))
    assert np.array_equal(o1, o2)
    R.math_vector_scaling(P(o1), C.c_double(1.3), P(a), C.c_int(n)); L.o_vector_scaling(P(o2), C.c_double(1.3), P(a), C.c_int(n))
    assert np.array_equal(o1, o2)
    x1, x2 = a.copy() + 3, a.copy() + 3
    R.math_invert_vector_elements(P(x1), C.c_int(n)); L.o_invert_vector_elements(P(x2), C.c_int(n))
    assert np.array_equal(x1, x2)
    x1, x2 = np.zeros(n + 5), np.zeros(n + 5)
    R.math_set_to_value(P(x1), C.c_double(2.5), C.c_int(n), C.c_int(3)); L.o_set_to_value(P(x2), C.c_double(2.5), C.c_int(n), C.c_int(3))
    assert np.array_equal(x1, x2)


@pytest.mark.parametrize("dim,ladder", [(2, [7, 4, 1]), (3, [7, 4, 1]), (3, [3, 1]), (2, [5, 2, 1])])
def test_subdomain_operator_mixed_degree(dim, ladder):
    R, L = _ref(dim), c.lib()
    rng = np.random.default_rng(7)
    pd = np.array(ladder, dtype=np.float64)
    R.ref_set_poly_degree(P(pd), C.c_int(len(ladder)))
    # region: 2 elements of every level, fine first
    offs, verts, levels = [], [], []
    o = 0
    for l, N in enumerate(ladder):
        for _ in range(2):
            npe = (N + 1) ** dim
            offs += [o] * npe; verts += list(range(npe)); levels += [l] * npe
            o += npe
    npts = o
    offs, verts, levels = (np.array(x, dtype=np.int32) for x in (offs, verts, levels))
    Ds = []
    for N in ladder:
        z, _ = c.zwgll(N + 1)
        Ds.append(np.ascontiguousarray(c.dgll(z, N + 1).ravel()))
    u = _rand(rng, npts); G = [_rand(rng, npts) for _ in range(6)]
    res = []
    for lib, pre in ((R, "subdomain_"), (L, "o_sub_")):
        gdu = [np.zeros(npts) for _ in range(dim)]
        Au = np.zeros(npts)
        a1 = [c.ptr_table(gdu), P(u), c.ptr_table(Ds), P(offs), P(verts), P(levels), c.ptr_table(G), C.c_int(npts)]
        a2 = [P(Au), c.ptr_table(gdu), c.ptr_table(Ds), P(offs), P(verts), P(levels), C.c_int(npts)]
        if lib is L:
            a1 += [P(pd), C.c_int(dim)]; a2 += [P(pd), C.c_int(dim)]
        getattr(lib, pre + "stiffness_matrix_1")(*a1)
        getattr(lib, pre + "stiffness_matrix_2")(*a2)
        res.append(Au)
    assert np.array_equal(res[0], res[1])


@pytest.mark.parametrize("dim,nf,nc", [(2, 8, 5), (2, 5, 2), (3, 8, 5), (3, 5, 2), (3, 8, 2), (3, 10, 7)])
def test_restriction(dim, nf, nc):
    R, L = _ref(dim), c.lib()
    rng = np.random.default_rng(nf * 10 + nc)
    E = 3
    zf, _ = c.zwgll(nf); zc, _ = c.zwgll(nc)
    J = np.ascontiguousarray(np.array([[c.hgll(j + 1, zf[i], zc.copy(), nc) for j in range(nc)] for i in range(nf)]).ravel())
    u = _rand(rng, E * nf ** dim)
    outs = []
    for lib, pre in ((R, "subdomain_"), (L, "o_")):
        ex = [C.c_int(dim)] if lib is L else []
        if dim == 2:
            t1 = np.zeros(E * nf * nc); uc = np.zeros(E * nc * nc)
            getattr(lib, pre + "restriction_1")(P(t1), P(J), P(u), C.c_int(t1.size), C.c_int(nf), C.c_int(nc), *ex)
            getattr(lib, pre + "restriction_2")(P(uc), P(J), P(t1), C.c_int(uc.size), C.c_int(nf), C.c_int(nc), *ex)
        else:
            t1 = np.zeros(E * nf * nf * nc); t2 = np.zeros(E * nf * nc * nc); uc = np.zeros(E * nc ** 3)
            getattr(lib, pre + "restriction_1")(P(t1), P(J), P(u), C.c_int(t1.size), C.c_int(nf), C.c_int(nc), *ex)
            getattr(lib, pre + "restriction_2")(P(t2), P(J), P(t1), C.c_int(t2.size), C.c_int(nf), C.c_int(nc), *ex)
            getattr(lib, pre + "restriction_3")(P(uc), P(J), P(t2), C.c_int(uc.size), C.c_int(nf), C.c_int(nc))
        outs.append(uc)
    assert np.array_equal(outs[0], outs[1])
    # restriction = interpolation^T (pin 3 of SURVEY 8c)
    Jm = J.reshape(nf, nc)
    K = np.kron(Jm, Jm) if dim == 2 else np.kron(Jm, np.kron(Jm, Jm))
    for e in range(E):
        assert np.allclose(outs[1][e * nc ** dim:(e + 1) * nc ** dim], K.T @ u[e * nf ** dim:(e + 1) * nf ** dim], rtol=1e-12, atol=1e-12)


def test_csr_kernels():
    R, L = _ref(3), c.lib()
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    A = sp.random(200, 150, density=0.05, random_state=5, format="csr")
    ptr, col, val = A.indptr.astype(np.int32), A.indices.astype(np.int32), np.ascontiguousarray(A.data)
    u = _rand(rng, 150); w = _rand(rng, 200)
    o1, o2 = np.zeros(200), np.zeros(200)
    R.csr_multiply(P(o1), P(ptr), P(col), P(val), P(u), C.c_int(200)); L.o_csr_multiply(P(o2), P(ptr), P(col), P(val), P(u), C.c_int(200))
    assert np.array_equal(o1, o2) and np.allclose(o1, A @ u)
    R.csr_multiply_weight(P(o1), P(ptr), P(col), P(val), P(u), P(w), C.c_int(200)); L.o_csr_multiply_weight(P(o2), P(ptr), P(col), P(val), P(u), P(w), C.c_int(200))
    assert np.array_equal(o1, o2)
    o1[:] = 0; o2[:] = 0
    R.csr_multiply_range(P(o1), P(ptr), P(col), P(val), P(u), C.c_int(10), C.c_int(20)); L.o_csr_multiply_range(P(o2), P(ptr), P(col), P(val), P(u), C.c_int(10), C.c_int(20))
    assert np.array_equal(o1, o2) and o1[21] == 0 and o1[9] == 0
