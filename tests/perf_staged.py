"""Experiment (not a test): row-group SpMV (variant 0) against the warp-staged SpMV (variant 1) on every level of the real c2
low-order AMG hierarchy, through the descriptor entry points, batch-timed with CUDA events over rotating copies of the matrix
(working set > 2 x L2, so no launch finds its matrix in L2).   python tests/perf_staged.py [nel] [eps]"""
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
from test_gpu_kernels import csr_descriptor  # noqa: E402
import gpu_util as G  # noqa: E402

L = pr.lib()
G.lib = L
PEAK = 6451.8
stream = torch.cuda.Stream()
sh = C.c_void_p(stream.cuda_stream)


def P(t):
    return C.c_void_p(t.data_ptr())


def batch(fn, reps=24, warm=4):
    with torch.cuda.stream(stream):
        for i in range(warm):
            fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(reps):
            fn(i)
        e1.record(stream)
        stream.synchronize()
    return e0.elapsed_time(e1) / reps


def time_level(A, name, shapes):
    A = A.tocsr(); A.sort_indices()
    nr = A.shape[0]
    ptr, col, val = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    nbytes = 12.0 * A.nnz + 4.0 * (nr + 1) + 32.0 * nr
    ncopy = max(2, int(np.ceil(300e6 / nbytes)))
    ncopy = min(ncopy, 16)
    copies = []
    for _ in range(ncopy):
        copies.append(tuple(torch.from_numpy(a).cuda() for a in (ptr, col, val)) + tuple(torch.rand(nr, dtype=torch.float64, device="cuda") for _ in range(5)))
    xh = copies[0][3].cpu().numpy(); ref = A @ xh
    out = []
    for variant, tpr, rpg, ctas in shapes:
        descs = [csr_descriptor(G, ptr, c[0], c[1], c[2], tpr, rpg) if tpr else csr_descriptor(G, ptr, c[0], c[1], c[2]) for c in copies]
        L.prfdd_csr_set_spmv_variant(C.c_int(variant), C.c_int(ctas))
        y = copies[0][4]
        assert L.prfdd_csrm_multiply(P(y), C.byref(descs[0][0]), P(copies[0][3]), sh) == 0
        torch.cuda.synchronize()
        err = np.abs(y.cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err < 1e-13, (variant, tpr, err)

        def step(i):
            c = copies[i % ncopy]; D = descs[i % ncopy][0]
            rc = L.prfdd_csrm_cheby_step(P(c[7]), P(c[4]), C.byref(D), P(c[3]), P(c[5]), P(c[6]), C.c_double(0.5), C.c_int(0), C.c_int(0), sh)
            assert rc == 0
        t = batch(step)
        D = descs[0][0]
        out.append("v%d tpr%d rpg%d cap%d ctas%d: %.1f us (%.2f)" % (variant, D.threads_per_row, D.stage_rows_per_lane_group, D.stage_cap, ctas, t * 1e3, nbytes / 1e9 / (t * 1e-3) / PEAK))
    L.prfdd_csr_set_spmv_variant(C.c_int(1), C.c_int(0))
    print("  %-8s rows %8d nnz/row %5.1f copies %d (%.0f MB each)\n      %s" % (name, nr, A.nnz / nr, ncopy, nbytes / 1e6, "\n      ".join(out)), flush=True)


def main():
    nel = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    eps = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    N, r = 7, 3
    d = tempfile.mkdtemp()
    pr.mesh_generate_box(d, 3, nel, N, 1, eps, reduction=r)
    S = pr.Solver(d, poly_degree=N, poly_reduction=r)
    nd = S.query("NUM_DOFS")
    A = sp.csr_matrix((S.get_array("A_FEM_VAL"), S.get_array("A_FEM_COL"), S.get_array("A_FEM_PTR")), shape=(nd, nd))
    del S
    torch.cuda.empty_cache()
    H, _ = product_hierarchy(pr, A, 2, 9)
    for l, lev in enumerate(H[:4]):
        Al = lev["A"]
        avg = Al.nnz / Al.shape[0]
        shapes = [(0, 0, 0, 0), (1, 0, 0, 0), (1, 0, 0, 2), (1, 0, 0, 3), (1, 0, 0, 4)]
        if avg > 10:
            shapes += [(1, 4, 2, 0), (1, 4, 1, 0), (1, 8, 1, 0), (1, 8, 2, 0), (1, 16, 1, 0)]
        else:
            shapes += [(1, 2, 1, 0), (1, 4, 2, 0)]
        print("level %d" % l)
        time_level(Al, "A", shapes)
        if "P" in lev and lev["P"] is not None and l < 3:
            Pm = lev["P"].tocsr()
            time_level(Pm, "P", [(0, 0, 0, 0), (1, 0, 0, 0)])
            time_level(Pm.T.tocsr(), "R", [(0, 0, 0, 0), (1, 0, 0, 0)])


if __name__ == "__main__":
    main()
