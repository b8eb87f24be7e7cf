"""Experiment (not a test): streamed SpMV (k_spmv_stream) against the sub-warp-per-row kernel on the real c2 AMG levels.
    python tests/perf_stream.py [nel]"""
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
from perf_micro import timeit, P, L, sh, PEAK  # noqa: E402
from perf_reorder import tpr_of  # noqa: E402


def row_blocks(ptr):
    nr = len(ptr) - 1
    nb = C.c_int(0)
    rc = L.prfdd_csr_row_blocks(ptr.ctypes.data_as(C.c_void_p), C.c_int(nr), None, C.byref(nb))
    if rc:
        return None
    rb = np.zeros(nb.value + 1, np.int32)
    L.prfdd_csr_row_blocks(ptr.ctypes.data_as(C.c_void_p), C.c_int(nr), rb.ctypes.data_as(C.c_void_p), C.byref(nb))
    return rb


def main():
    nel = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    N, r = 7, 3
    d = tempfile.mkdtemp()
    pr.mesh_generate_box(d, 3, nel, N, 1, 0.0, reduction=r)
    S = pr.Solver(d, poly_degree=N, poly_reduction=r)
    nd = S.query("NUM_DOFS")
    A = sp.csr_matrix((S.get_array("A_FEM_VAL"), S.get_array("A_FEM_COL"), S.get_array("A_FEM_PTR")), shape=(nd, nd))
    del S
    torch.cuda.empty_cache()
    H, _ = product_hierarchy(pr, A, 2, 9)
    only = os.environ.get("PRFDD_STREAM_ONLY")   # "level:tpr" -> just that case (for ncu)
    for l, lev in enumerate(H[:5]):
        if only and l != int(only.split(":")[0]):
            continue
        Al = lev["A"].tocsr(); Al.sort_indices()
        nr = Al.shape[0]
        hp = Al.indptr.astype(np.int32)
        ptr, col, val = (torch.from_numpy(a).cuda() for a in (hp, Al.indices.astype(np.int32), Al.data.astype(np.float64)))
        xh = np.random.default_rng(l).standard_normal(nr)
        x = torch.from_numpy(xh).cuda(); y = torch.empty(nr, dtype=torch.float64, device="cuda")
        rr = torch.rand(nr, dtype=torch.float64, device="cuda"); ds = torch.rand(nr, dtype=torch.float64, device="cuda"); uu = torch.zeros(nr, dtype=torch.float64, device="cuda")
        gb = (12.0 * Al.nnz + 4.0 * (nr + 1) + 8.0 * nr * 5) / 1e9
        rb = row_blocks(hp)
        print("level %d rows %d nnz/row %.1f blocks %s" % (l, nr, Al.nnz / nr, None if rb is None else len(rb) - 1), flush=True)
        t0 = tpr_of(Al)
        med, _ = timeit(lambda: L.prfdd_cheby_step(P(uu), P(y), P(ptr), P(col), P(val), P(x), P(rr), P(ds), C.c_double(0.5), C.c_int(1), C.c_int(0), C.c_int(nr), C.c_int(t0), sh))
        print("   rowwise tpr%-2d %.1fus (%.2f)" % (t0, med * 1e3, gb / (med * 1e-3) / PEAK), flush=True)
        if rb is None:
            continue
        rbd = torch.from_numpy(rb).cuda()
        ref = Al @ xh
        for tpr in ((int(only.split(":")[1]),) if only else (1, 2, 4, 8, 16, 32)):
            rc = L.prfdd_csr_multiply_stream(P(y), P(ptr), P(col), P(val), P(x), P(rbd), C.c_int(len(rb) - 1), C.c_int(Al.nnz), C.c_int(tpr), sh)
            torch.cuda.synchronize()
            err = np.abs(y.cpu().numpy() - ref).max() / np.abs(ref).max()
            med, _ = timeit(lambda: L.prfdd_cheby_step_stream(P(uu), P(y), P(ptr), P(col), P(val), P(x), P(rr), P(ds), C.c_double(0.5), C.c_int(1), C.c_int(0), P(rbd), C.c_int(len(rb) - 1), C.c_int(Al.nnz), C.c_int(tpr), sh))
            print("   stream  tpr%-2d %.1fus (%.2f)  rc %d err %.1e" % (tpr, med * 1e3, gb / (med * 1e-3) / PEAK, rc, err), flush=True)


if __name__ == "__main__":
    main()
