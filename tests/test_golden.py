"""Golden vectors: outputs of the REFERENCE'S OWN kernels (the four .okl files compiled for the CPU, tests/golden/make_golden.py)
on seeded inputs, committed as tests/golden/okl_reference_vectors.npz.
  * CPU (not gpu): the oracle's C restatement reproduces them bit for bit -- the oracle stays pinned where /root/reference and
    oracle/_ref do not exist (the GPU box);
  * -m gpu: the CUDA kernels, called through the C ABI, reproduce them to 1e-13 of scale (FMA contraction and summation
    order differ from scalar CPU loops; nothing else does)."""
import ctypes as C
import os
import numpy as np
import pytest

from oracle import capi as oc

P = oc.ptr
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "okl_reference_vectors.npz"))
OPS = [(3, 7), (3, 4), (3, 1), (3, 9), (2, 7), (2, 3)]
RESTR = [(3, 8, 5), (3, 5, 2), (3, 10, 7), (3, 16, 9), (2, 8, 5), (2, 5, 2)]
VECS = [1, 129, 1000, 5000]
TOL = 1e-13


# ------------------------------------------------------------------ CPU: oracle == reference outputs, bit for bit
@pytest.mark.parametrize("dim,N", OPS)
def test_oracle_operator_matches_reference_vectors(dim, N):
    L = oc.lib()
    k = "op_%dd_N%d" % (dim, N)
    u, Gs, D, Au = GOLD[k + "_u"], [np.ascontiguousarray(g) for g in GOLD[k + "_G"]], GOLD[k + "_D"], GOLD[k + "_Au"]
    npts = u.size
    gdu = [np.zeros(npts) for _ in range(dim)]; out = np.zeros(npts)
    L.o_stiffness_matrix_1(oc.ptr_table(gdu), P(u), P(D), oc.ptr_table(Gs), C.c_int(npts), C.c_int(N), C.c_int(dim))
    L.o_stiffness_matrix_2(P(out), oc.ptr_table(gdu), P(D), C.c_int(npts), C.c_int(N), C.c_int(dim))
    assert np.array_equal(out, Au)
    z, _ = oc.zwgll(N + 1)
    assert np.array_equal(np.ascontiguousarray(oc.dgll(z, N + 1).ravel()), D)     # speclib D the fixture was made with


@pytest.mark.parametrize("dim,nf,nc", RESTR)
def test_oracle_restriction_matches_reference_vectors(dim, nf, nc):
    L = oc.lib()
    k = "restr_%dd_%d_%d" % (dim, nf, nc)
    J, u, uc = GOLD[k + "_J"], GOLD[k + "_u"], GOLD[k + "_uc"]
    E = u.size // nf ** dim
    ex = [C.c_int(dim)]
    if dim == 2:
        t1 = np.zeros(E * nf * nc); out = np.zeros(E * nc * nc)
        L.o_restriction_1(P(t1), P(J), P(u), C.c_int(t1.size), C.c_int(nf), C.c_int(nc), *ex)
        L.o_restriction_2(P(out), P(J), P(t1), C.c_int(out.size), C.c_int(nf), C.c_int(nc), *ex)
    else:
        t1 = np.zeros(E * nf * nf * nc); t2 = np.zeros(E * nf * nc * nc); out = np.zeros(E * nc ** 3)
        L.o_restriction_1(P(t1), P(J), P(u), C.c_int(t1.size), C.c_int(nf), C.c_int(nc), *ex)
        L.o_restriction_2(P(t2), P(J), P(t1), C.c_int(t2.size), C.c_int(nf), C.c_int(nc), *ex)
        L.o_restriction_3(P(out), P(J), P(t2), C.c_int(out.size), C.c_int(nf), C.c_int(nc))
    assert np.array_equal(out, uc)


@pytest.mark.parametrize("n", VECS)
def test_oracle_vector_kernels_match_reference_vectors(n):
    L = oc.lib()
    k = "vec_%d" % n
    a, b, cc, d, m = (np.ascontiguousarray(x) for x in GOLD[k + "_in"])
    nb = (n + 127) // 128
    def blocks(name, nout, *args):
        o = np.zeros(nout); getattr(L, name)(P(o), *args); return o
    assert np.array_equal(blocks("o_residual_norm", nb, P(a), P(b), P(m), C.c_int(n), C.c_int(nb)), GOLD[k + "_residual_norm"])
    assert np.array_equal(blocks("o_projection_inner_products", 2 * nb, P(a), P(b), P(cc), P(d), C.c_int(n), C.c_int(nb)), GOLD[k + "_projection"])
    assert np.array_equal(blocks("o_inner_product_flexible", nb, P(a), P(b), P(cc), C.c_int(n), C.c_int(nb)), GOLD[k + "_flexible"])
    assert np.array_equal(blocks("o_inner_product_mask", nb, P(a), P(b), P(m), C.c_int(n), C.c_int(nb)), GOLD[k + "_inner_mask"])
    assert np.array_equal(blocks("o_sub_weighted_inner_product", nb, P(a), P(b), P(m), C.c_int(n), C.c_int(nb)), GOLD[k + "_sub_weighted"])
    assert np.array_equal(blocks("o_sub_projection_inner_products", 2 * nb, P(a), P(b), P(cc), P(d), P(m), C.c_int(n), C.c_int(nb)), GOLD[k + "_sub_projection"])
    assert np.array_equal(blocks("o_sub_search_update_inner_product", nb, P(a), P(b), P(cc), P(m), C.c_int(n), C.c_int(nb)), GOLD[k + "_sub_search"])
    u1, r1 = a.copy(), np.zeros(n)
    L.o_solution_and_residual_update(P(u1), P(r1), P(b), P(cc), P(d), C.c_double(0.37), C.c_int(n))
    assert np.array_equal(u1, GOLD[k + "_sru_u"]) and np.array_equal(r1, GOLD[k + "_sru_r1"])
    p1, r1 = a.copy(), np.zeros(n)
    L.o_residual_and_search_update(P(p1), P(r1), P(b), P(cc), C.c_double(-1.7), C.c_int(n))
    assert np.array_equal(p1, GOLD[k + "_rsu_p"]) and np.array_equal(r1, GOLD[k + "_rsu_r"])
    o1 = np.zeros(n)
    L.o_vector_vector_addition(P(o1), C.c_double(1.3), P(a), C.c_double(-0.2), P(b), C.c_int(n))
    assert np.array_equal(o1, GOLD[k + "_axpby"])


def test_oracle_csr_and_region_operator_match_reference_vectors():
    L = oc.lib()
    ptr, col, val, u, w = (np.ascontiguousarray(GOLD["csr_" + x]) for x in ("ptr", "col", "val", "u", "w"))
    nr = ptr.size - 1
    o = np.zeros(nr); L.o_csr_multiply(P(o), P(ptr), P(col), P(val), P(u), C.c_int(nr)); assert np.array_equal(o, GOLD["csr_multiply"])
    o = np.zeros(nr); L.o_csr_multiply_weight(P(o), P(ptr), P(col), P(val), P(u), P(w), C.c_int(nr)); assert np.array_equal(o, GOLD["csr_multiply_weight"])
    o = np.zeros(nr); L.o_csr_multiply_range(P(o), P(ptr), P(col), P(val), P(u), C.c_int(100), C.c_int(555)); assert np.array_equal(o, GOLD["csr_multiply_range"])
    # mixed-degree region operator
    dim, ladder = 3, [int(x) for x in GOLD["region_ladder"]]
    u, Gs, Au = GOLD["region_u"], [np.ascontiguousarray(g) for g in GOLD["region_G"]], GOLD["region_Au"]
    npts = u.size
    offs, verts, levels = _region_maps(ladder, dim)
    Ds = [_D(N + 1) for N in ladder]
    pd = np.array(ladder, dtype=np.float64)
    gdu = [np.zeros(npts) for _ in range(dim)]; out = np.zeros(npts)
    L.o_sub_stiffness_matrix_1(oc.ptr_table(gdu), P(u), oc.ptr_table(Ds), P(offs), P(verts), P(levels), oc.ptr_table(Gs), C.c_int(npts), P(pd), C.c_int(dim))
    L.o_sub_stiffness_matrix_2(P(out), oc.ptr_table(gdu), oc.ptr_table(Ds), P(offs), P(verts), P(levels), C.c_int(npts), P(pd), C.c_int(dim))
    assert np.array_equal(out, Au)


def _D(n):
    z, _ = oc.zwgll(n)
    return np.ascontiguousarray(oc.dgll(z, n).ravel())


def _region_maps(ladder, dim):
    offs, verts, levels, o = [], [], [], 0
    for l, N in enumerate(ladder):
        for _ in range(2):
            npe = (N + 1) ** dim
            offs += [o] * npe; verts += list(range(npe)); levels += [l] * npe
            o += npe
    return tuple(np.array(x, dtype=np.int32) for x in (offs, verts, levels))


# ------------------------------------------------------------------ GPU: CUDA kernels vs reference outputs
@pytest.fixture(scope="module")
def G(prfdd):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gpu_util as g
    g.lib = prfdd.lib()
    g.ws = g.WS(g.lib)
    return g


@pytest.mark.gpu
@pytest.mark.parametrize("dim,N", OPS)
def test_cuda_operator_matches_reference_vectors(G, dim, N):
    k = "op_%dd_N%d" % (dim, N)
    u, Gs, D, Au = GOLD[k + "_u"], GOLD[k + "_G"], GOLD[k + "_D"], GOLD[k + "_Au"]
    n = N + 1
    E = u.size // n ** dim
    du, dD, dG, dAu = G.dev(u), G.dev(D), [G.dev(np.ascontiguousarray(g)) for g in Gs], G.dev(np.full(u.size, np.nan))
    scale = (np.abs(np.stack(Gs)).max() * np.abs(u).max() * np.abs(D).max() ** 2) * n * n
    Dh = np.ascontiguousarray(D, dtype=np.float64)
    for host_d in (None, Dh.ctypes.data_as(C.c_void_p)):      # generic kernel, then the constant-bank / bulk-async kernels
        assert G.lib.prfdd_stiffness_matrix_hd(G.p(dAu), G.p(du), G.p(dD), host_d, G.ptr_array(dG), C.c_int(E), C.c_int(n), C.c_int(dim), G.stream()) == 0
        G.sync()
        assert np.abs(G.host(dAu) - Au).max() <= 50 * TOL * scale


@pytest.mark.gpu
@pytest.mark.parametrize("dim,nf,nc", RESTR)
def test_cuda_restriction_matches_reference_vectors(G, dim, nf, nc):
    k = "restr_%dd_%d_%d" % (dim, nf, nc)
    J, u, uc = GOLD[k + "_J"], GOLD[k + "_u"], GOLD[k + "_uc"]
    E = u.size // nf ** dim
    dJ, du, dout = G.dev(J), G.dev(u), G.dev(np.full(uc.size, np.nan))
    assert G.lib.prfdd_restriction(G.p(dout), G.p(dJ), G.p(du), C.c_int(E), C.c_int(nf), C.c_int(nc), C.c_int(dim), G.stream()) == 0
    G.sync()
    assert np.abs(G.host(dout) - uc).max() <= 100 * TOL * np.abs(uc).max()


@pytest.mark.gpu
@pytest.mark.parametrize("n", VECS)
def test_cuda_vector_kernels_match_reference_vectors(G, n):
    lib = G.lib
    k = "vec_%d" % n
    a, b, cc, d, m = (np.ascontiguousarray(x) for x in GOLD[k + "_in"])
    da, db, dc, dd, dm = (G.dev(x) for x in (a, b, cc, d, m))
    nb = (n + 127) // 128
    out = G.dev(np.zeros(8))
    ssum = lambda blk: float(np.sum(blk))                      # the reference's host loop over the block partials (domain.tpp:924-926)

    def check(ref, terms, idx=0):
        G.sync()
        assert abs(G.host(out)[idx] - ref) <= 4e-13 * np.abs(terms).sum() + 1e-300

    assert lib.prfdd_residual_norm(G.ws.h, G.p(out), G.p(da), G.p(db), G.p(dm), C.c_int(n), G.stream()) == 0
    check(ssum(GOLD[k + "_residual_norm"]), a * b * m)
    assert lib.prfdd_projection_inner_products(G.ws.h, G.p(out), G.p(da), G.p(db), G.p(dc), G.p(dd), C.c_int(n), G.stream()) == 0
    check(ssum(GOLD[k + "_projection"][:nb]), a * b, 0); check(ssum(GOLD[k + "_projection"][nb:]), cc * d, 1)
    assert lib.prfdd_inner_product_flexible(G.ws.h, G.p(out), G.p(da), G.p(db), G.p(dc), C.c_int(n), G.stream()) == 0
    check(ssum(GOLD[k + "_flexible"]), (b - a) * cc)
    assert lib.prfdd_inner_product(G.ws.h, G.p(out), G.p(da), G.p(db), G.p(dm), C.c_int(n), G.stream()) == 0
    check(ssum(GOLD[k + "_inner_mask"]), a * b * m)
    assert lib.prfdd_weighted_inner_product(G.ws.h, G.p(out), G.p(da), G.p(db), G.p(dm), C.c_int(n), G.stream()) == 0
    check(ssum(GOLD[k + "_sub_weighted"]), a * b * m)
    assert lib.prfdd_weighted_projection_inner_products(G.ws.h, G.p(out), G.p(da), G.p(db), G.p(dc), G.p(dd), G.p(dm), C.c_int(n), G.stream()) == 0
    check(ssum(GOLD[k + "_sub_projection"][:nb]), a * b * m, 0); check(ssum(GOLD[k + "_sub_projection"][nb:]), cc * d * m, 1)
    assert lib.prfdd_search_update_inner_product(G.ws.h, G.p(out), G.p(da), G.p(db), G.p(dc), G.p(dm), C.c_int(n), G.stream()) == 0
    check(ssum(GOLD[k + "_sub_search"]), (b - a) * cc * m)
    du, dr1 = G.dev(a), G.dev(np.zeros(n))
    assert lib.prfdd_solution_and_residual_update(G.p(du), G.p(dr1), G.p(db), G.p(dc), G.p(dd), C.c_double(0.37), C.c_int(n), G.stream()) == 0
    G.sync()
    assert np.abs(G.host(du) - GOLD[k + "_sru_u"]).max() <= 4e-15 and np.abs(G.host(dr1) - GOLD[k + "_sru_r1"]).max() <= 4e-15
    dp, dr = G.dev(a), G.dev(np.zeros(n))
    assert lib.prfdd_residual_and_search_update(G.p(dp), G.p(dr), G.p(db), G.p(dc), C.c_double(-1.7), C.c_int(n), G.stream()) == 0
    G.sync()
    assert np.abs(G.host(dp) - GOLD[k + "_rsu_p"]).max() <= 4e-15 and np.array_equal(G.host(dr), GOLD[k + "_rsu_r"])
    do = G.dev(np.zeros(n))
    assert lib.prfdd_vector_vector_addition(G.p(do), C.c_double(1.3), G.p(da), C.c_double(-0.2), G.p(db), C.c_int(n), G.stream()) == 0
    G.sync(); assert np.abs(G.host(do) - GOLD[k + "_axpby"]).max() <= 4e-15


@pytest.mark.gpu
def test_cuda_csr_and_region_operator_match_reference_vectors(G):
    lib = G.lib
    ptr, col, val, u, w = (np.ascontiguousarray(GOLD["csr_" + x]) for x in ("ptr", "col", "val", "u", "w"))
    nr = ptr.size - 1
    dptr, dcol, dval, du, dw = (G.dev(x) for x in (ptr, col, val, u, w))
    tol = 8e-15 * np.abs(GOLD["csr_multiply"]).max() * 20
    for tpr in (1, 4, 32):
        out = G.dev(np.zeros(nr))
        assert lib.prfdd_csr_multiply(G.p(out), G.p(dptr), G.p(dcol), G.p(dval), G.p(du), C.c_int(nr), C.c_int(tpr), G.stream()) == 0
        G.sync(); assert np.abs(G.host(out) - GOLD["csr_multiply"]).max() <= tol
        assert lib.prfdd_csr_multiply_weight(G.p(out), G.p(dptr), G.p(dcol), G.p(dval), G.p(du), G.p(dw), C.c_int(nr), C.c_int(tpr), G.stream()) == 0
        G.sync(); assert np.abs(G.host(out) - GOLD["csr_multiply_weight"]).max() <= tol * np.abs(w).max()
        out = G.dev(np.zeros(nr))
        assert lib.prfdd_csr_multiply_range(G.p(out), G.p(dptr), G.p(dcol), G.p(dval), G.p(du), C.c_int(100), C.c_int(555), C.c_int(tpr), G.stream()) == 0
        G.sync(); assert np.abs(G.host(out) - GOLD["csr_multiply_range"]).max() <= tol
    dim, ladder = 3, [int(x) for x in GOLD["region_ladder"]]
    u, Gs, Au = GOLD["region_u"], GOLD["region_G"], GOLD["region_Au"]
    first, o = [], 0
    for N in ladder:
        first.append(o); o += 2 * (N + 1) ** dim
    du, dG, dDs, dAu = G.dev(u), [G.dev(np.ascontiguousarray(g)) for g in Gs], [G.dev(_D(N + 1)) for N in ladder], G.dev(np.zeros(u.size))
    fp = (C.c_int * 3)(*first); ne = (C.c_int * 3)(2, 2, 2); nn = (C.c_int * 3)(*[N + 1 for N in ladder])
    assert lib.prfdd_stiffness_matrix_region(G.p(dAu), G.p(du), G.ptr_array(dG), C.c_int(3), fp, ne, nn, G.ptr_array(dDs), C.c_int(dim), G.stream()) == 0
    G.sync()
    assert np.abs(G.host(dAu) - Au).max() <= 50 * TOL * np.abs(Au).max()
