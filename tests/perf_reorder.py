"""Experiment (not a test): does a storage permutation of the AMG levels (Morton order of the dof coordinates) speed the
V-cycle SpMVs up?  Builds the real c2 low-order matrix, the product's host AMG hierarchy, and times cheby_step per level with
the natural ordering and with Morton / lexicographic orderings.   python tests/perf_reorder.py [nel]"""
import ctypes as C
import os
import sys
import tempfile
import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import polynomial_reduction_with_full_domain_decomposition_preconditioner_b200 as pr  # noqa: E402
from test_amg_host import product_hierarchy  # noqa: E402
from perf_micro import timeit_batch as timeit, P, L, sh, PEAK  # noqa: E402


def tpr_of(A):
    avg = A.nnz / A.shape[0]
    return 1 if avg <= 10 else 2 if avg <= 18 else 16 if (avg > 60 and A.shape[0] < 50000) else 8


def morton(ix, iy, iz):
    key = np.zeros(ix.shape, np.int64)
    for b in range(10):
        key |= ((ix >> b) & 1) << (3 * b) | ((iy >> b) & 1) << (3 * b + 1) | ((iz >> b) & 1) << (3 * b + 2)
    return key


def time_level(A, name):
    A = A.tocsr(); A.sort_indices()
    nr = A.shape[0]
    ptr, col, val = (torch.from_numpy(a).cuda() for a in (A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)))
    x = torch.rand(nr, dtype=torch.float64, device="cuda"); y = torch.empty(nr, dtype=torch.float64, device="cuda")
    r = torch.rand(nr, dtype=torch.float64, device="cuda"); ds = torch.rand(nr, dtype=torch.float64, device="cuda"); uu = torch.zeros(nr, dtype=torch.float64, device="cuda")
    gb = (12.0 * A.nnz + 4.0 * (nr + 1) + 8.0 * nr * 5) / 1e9
    out = []
    xh = x.cpu().numpy(); ref = A @ xh
    t0 = tpr_of(A)
    for tpr in sorted({1, t0, max(1, t0 // 2), min(32, t0 * 2)}):
        L.prfdd_csr_multiply(P(y), P(ptr), P(col), P(val), P(x), C.c_int(nr), C.c_int(tpr), sh)
        torch.cuda.synchronize()
        err = np.abs(y.cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err < 1e-13, (tpr, err)
        med, mn = timeit(lambda: L.prfdd_cheby_step(P(uu), P(y), P(ptr), P(col), P(val), P(x), P(r), P(ds), C.c_double(0.5), C.c_int(1), C.c_int(0), C.c_int(nr), C.c_int(tpr), sh))
        out.append("tpr%d %.1fus (%.2f)" % (tpr, med * 1e3, gb / (med * 1e-3) / PEAK))
    print("  %-14s rows %8d nnz/row %5.1f: %s" % (name, nr, A.nnz / nr, "  ".join(out)), flush=True)


def main():
    nel = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    N, r = 7, 3
    d = tempfile.mkdtemp()
    pr.mesh_generate_box(d, 3, nel, N, 1, 0.0, reduction=r)
    S = pr.Solver(d, poly_degree=N, poly_reduction=r)
    nd = S.query("NUM_DOFS")
    A = sp.csr_matrix((S.get_array("A_FEM_VAL"), S.get_array("A_FEM_COL"), S.get_array("A_FEM_PTR")), shape=(nd, nd))
    dof = S.get_array("SUB_DOF_NUM"); eid = S.get_array("SUB_ELEMENT_IDS")
    n3 = (N + 1) ** 3
    xyz = []
    for c in "xyz":
        a = np.fromfile(os.path.join(d, "lx1_%d" % (N + 1), "%s_0.%d.dat" % (c, N)), dtype=np.float64).reshape(-1, n3)[eid].ravel()
        v = np.zeros(nd); m = dof > 0; v[dof[m] - 1] = a[m]
        xyz.append(v)
    del S
    torch.cuda.empty_cache()
    H, _ = product_hierarchy(pr, A, 2, 9)
    coords = np.stack(xyz, 1)
    for l, lev in enumerate(H[:4]):
        Al = lev["A"]
        print("level %d" % l)
        time_level(Al, "natural")
        if os.environ.get("PRFDD_REORDER_NATURAL_ONLY"):
            if "cf" in lev and len(lev["cf"]):
                coords = coords[lev["cf"] > 0]
            continue
        q = [np.unique(np.round(coords[:, k], 9), return_inverse=True)[1].astype(np.int64) for k in range(3)]
        for name, key in (("lexicographic", (q[2] * 4096 + q[1]) * 4096 + q[0]), ("morton", morton(*q)),
                          ("morton/2", morton(q[0] // 2, q[1] // 2, q[2] // 2) * 8 + morton(q[0] % 2, q[1] % 2, q[2] % 2))):
            perm = np.argsort(key, kind="stable")
            time_level(Al[perm][:, perm], name)
        rng = np.random.default_rng(0)
        perm = rng.permutation(Al.shape[0])
        time_level(Al[perm][:, perm], "random")
        if "cf" in lev and len(lev["cf"]):
            coords = coords[lev["cf"] > 0]


if __name__ == "__main__":
    main()
