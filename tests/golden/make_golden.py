"""Generates tests/golden/okl_reference_vectors.npz: inputs and outputs of the REFERENCE'S OWN kernels (the four .okl files of
/root/reference compiled for the CPU by oracle/build_ref.py into oracle/_ref/) on seeded inputs.  Run in the build container,
where /root/reference exists:   python tests/golden/make_golden.py
The fixtures travel to the GPU box; tests/test_golden.py checks the oracle against them bit for bit (CPU) and the CUDA kernels
against them (-m gpu), so the CUDA path is compared with reference outputs directly, not only with the restatement."""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import capi as c  # noqa: E402

P = c.ptr
out = {}


def rnd(rng, n):
    return np.ascontiguousarray(rng.standard_normal(n))


def operator_cases():
    for dim, N, E in ((3, 7, 2), (3, 4, 2), (3, 1, 5), (3, 9, 1), (2, 7, 4), (2, 3, 3)):
        R = c.ref(dim)
        rng = np.random.default_rng(1000 * dim + N)
        n = N + 1
        npts = E * n ** dim
        z, _ = c.zwgll(n)
        D = np.ascontiguousarray(c.dgll(z, n).ravel())
        u = rnd(rng, npts); G = [rnd(rng, npts) for _ in range(6)]
        gdu = [np.zeros(npts) for _ in range(dim)]; Au = np.zeros(npts)
        R.domain_stiffness_matrix_1(c.ptr_table(gdu), P(u), P(D), c.ptr_table(G), C.c_int(npts), C.c_int(N))
        R.domain_stiffness_matrix_2(P(Au), c.ptr_table(gdu), P(D), C.c_int(npts), C.c_int(N))
        k = "op_%dd_N%d" % (dim, N)
        out[k + "_u"] = u; out[k + "_G"] = np.stack(G); out[k + "_D"] = D; out[k + "_Au"] = Au; out[k + "_E"] = np.array(E)


def restriction_cases():
    for dim, nf, nc in ((3, 8, 5), (3, 5, 2), (3, 10, 7), (3, 16, 9), (2, 8, 5), (2, 5, 2)):
        R = c.ref(dim)
        rng = np.random.default_rng(nf * 10 + nc)
        E = 3
        zf, _ = c.zwgll(nf); zc, _ = c.zwgll(nc)
        J = np.ascontiguousarray(np.array([[c.hgll(j + 1, zf[i], zc.copy(), nc) for j in range(nc)] for i in range(nf)]).ravel())
        u = rnd(rng, E * nf ** dim)
        if dim == 2:
            t1 = np.zeros(E * nf * nc); uc = np.zeros(E * nc * nc)
            R.subdomain_restriction_1(P(t1), P(J), P(u), C.c_int(t1.size), C.c_int(nf), C.c_int(nc))
            R.subdomain_restriction_2(P(uc), P(J), P(t1), C.c_int(uc.size), C.c_int(nf), C.c_int(nc))
        else:
            t1 = np.zeros(E * nf * nf * nc); t2 = np.zeros(E * nf * nc * nc); uc = np.zeros(E * nc ** 3)
            R.subdomain_restriction_1(P(t1), P(J), P(u), C.c_int(t1.size), C.c_int(nf), C.c_int(nc))
            R.subdomain_restriction_2(P(t2), P(J), P(t1), C.c_int(t2.size), C.c_int(nf), C.c_int(nc))
            R.subdomain_restriction_3(P(uc), P(J), P(t2), C.c_int(uc.size), C.c_int(nf), C.c_int(nc))
        k = "restr_%dd_%d_%d" % (dim, nf, nc)
        out[k + "_J"] = J; out[k + "_u"] = u; out[k + "_uc"] = uc


def vector_cases():
    R = c.ref(3)
    for n in (1, 129, 1000, 5000):
        rng = np.random.default_rng(n)
        nb = (n + 127) // 128
        a, b, cc, d = (rnd(rng, n) for _ in range(4))
        m = (rnd(rng, n) > 0).astype(np.float64)
        k = "vec_%d" % n
        out[k + "_in"] = np.stack([a, b, cc, d, m])
        def blocks(name, nout, *args):
            o = np.zeros(nout); getattr(R, name)(P(o), *args); return o
        out[k + "_residual_norm"] = blocks("domain_residual_norm", nb, P(a), P(b), P(m), C.c_int(n), C.c_int(nb))
        out[k + "_projection"] = blocks("domain_projection_inner_products", 2 * nb, P(a), P(b), P(cc), P(d), C.c_int(n), C.c_int(nb))
        out[k + "_flexible"] = blocks("domain_inner_product_flexible", nb, P(a), P(b), P(cc), C.c_int(n), C.c_int(nb))
        out[k + "_inner_mask"] = blocks("domain_inner_product", nb, P(a), P(b), P(m), C.c_int(n), C.c_int(nb))
        out[k + "_sub_weighted"] = blocks("subdomain_weighted_inner_product", nb, P(a), P(b), P(m), C.c_int(n), C.c_int(nb))
        out[k + "_sub_projection"] = blocks("subdomain_projection_inner_products", 2 * nb, P(a), P(b), P(cc), P(d), P(m), C.c_int(n), C.c_int(nb))
        out[k + "_sub_search"] = blocks("subdomain_search_update_inner_product", nb, P(a), P(b), P(cc), P(m), C.c_int(n), C.c_int(nb))
        u1, r1 = a.copy(), np.zeros(n)
        R.domain_solution_and_residual_update(P(u1), P(r1), P(b), P(cc), P(d), C.c_double(0.37), C.c_int(n))
        out[k + "_sru_u"] = u1; out[k + "_sru_r1"] = r1
        p1, r1 = a.copy(), np.zeros(n)
        R.domain_residual_and_search_update(P(p1), P(r1), P(b), P(cc), C.c_double(-1.7), C.c_int(n))
        out[k + "_rsu_p"] = p1; out[k + "_rsu_r"] = r1
        o1 = np.zeros(n)
        R.math_vector_vector_addition(P(o1), C.c_double(1.3), P(a), C.c_double(-0.2), P(b), C.c_int(n))
        out[k + "_axpby"] = o1


def csr_case():
    import scipy.sparse as sp
    R = c.ref(3)
    rng = np.random.default_rng(3)
    A = sp.random(900, 700, density=0.02, random_state=5, format="csr"); A.sort_indices()
    ptr, col, val = A.indptr.astype(np.int32), A.indices.astype(np.int32), np.ascontiguousarray(A.data)
    u = rnd(rng, 700); w = rnd(rng, 900)
    o1, o2, o3 = np.zeros(900), np.zeros(900), np.zeros(900)
    R.csr_multiply(P(o1), P(ptr), P(col), P(val), P(u), C.c_int(900))
    R.csr_multiply_weight(P(o2), P(ptr), P(col), P(val), P(u), P(w), C.c_int(900))
    R.csr_multiply_range(P(o3), P(ptr), P(col), P(val), P(u), C.c_int(100), C.c_int(555))
    out.update(csr_ptr=ptr, csr_col=col, csr_val=val, csr_u=u, csr_w=w, csr_multiply=o1, csr_multiply_weight=o2, csr_multiply_range=o3)


def region_case():
    """mixed-degree region operator (subdomain.okl:4-101), ladder 7/4/1, 2 elements per level, 3D"""
    dim, ladder = 3, [7, 4, 1]
    R = c.ref(dim)
    rng = np.random.default_rng(7)
    pd = np.array(ladder, dtype=np.float64)
    R.ref_set_poly_degree(P(pd), C.c_int(len(ladder)))
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
    u = rnd(rng, npts); G = [rnd(rng, npts) for _ in range(6)]
    gdu = [np.zeros(npts) for _ in range(dim)]; Au = np.zeros(npts)
    R.subdomain_stiffness_matrix_1(c.ptr_table(gdu), P(u), c.ptr_table(Ds), P(offs), P(verts), P(levels), c.ptr_table(G), C.c_int(npts))
    R.subdomain_stiffness_matrix_2(P(Au), c.ptr_table(gdu), c.ptr_table(Ds), P(offs), P(verts), P(levels), C.c_int(npts))
    out.update(region_u=u, region_G=np.stack(G), region_Au=Au, region_ladder=np.array(ladder), region_levels=levels)


if __name__ == "__main__":
    assert c.ref(3) is not None and c.ref(2) is not None, "oracle/_ref not built: python oracle/build_ref.py (needs /root/reference)"
    operator_cases(); restriction_cases(); vector_cases(); csr_case(); region_case()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "okl_reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%d arrays, %.1f KB" % (len(out), os.path.getsize(path) / 1024))
