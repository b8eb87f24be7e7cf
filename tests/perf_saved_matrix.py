"""Experiment (not a test): cheby_step on a CSR matrix saved by perf_multi_levels.py (PRFDD_SAVE_LEVEL0=file.npz), single process so
that ncu can wrap it.   python tests/perf_saved_matrix.py file.npz [tpr_arg ...]"""
import ctypes as C
import os
import sys
import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from perf_micro import timeit_batch, P, L, sh, PEAK  # noqa: E402

z = np.load(sys.argv[1])
A = sp.csr_matrix((z["data"], z["indices"], z["indptr"]))
nr = A.shape[0]
ptr, col, val = (torch.from_numpy(a).cuda() for a in (A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)))
x = torch.rand(nr, dtype=torch.float64, device="cuda"); y = torch.empty(nr, dtype=torch.float64, device="cuda")
rr = torch.rand(nr, dtype=torch.float64, device="cuda"); ds = torch.rand(nr, dtype=torch.float64, device="cuda"); uu = torch.zeros(nr, dtype=torch.float64, device="cuda")
gb = (12.0 * A.nnz + 4.0 * (nr + 1) + 8.0 * nr * 5) / 1e9
for arg in [int(v) for v in sys.argv[2:]] or [1]:
    t, _ = timeit_batch(lambda: L.prfdd_cheby_step(P(uu), P(y), P(ptr), P(col), P(val), P(x), P(rr), P(ds), C.c_double(0.5), C.c_int(1), C.c_int(0), C.c_int(nr), C.c_int(arg), sh), reps=5, warm=2)
    print("tpr %d: %.1f us (%.2f)" % (arg, t * 1e3, gb / (t * 1e-3) / PEAK), flush=True)
rl = np.diff(A.indptr)
n = nr
for arg in [int(v) for v in sys.argv[2:]] or [1]:
    out = []
    for q in range(8):
        lo, hi = q * n // 8, (q + 1) * n // 8 - 1
        t, _ = timeit_batch(lambda: L.prfdd_csr_multiply_range(P(y), P(ptr), P(col), P(val), P(x), C.c_int(lo), C.c_int(hi), C.c_int(arg), sh), reps=10, warm=2)
        out.append("%.1f" % (t * 1e3))
    print("tpr %d eighths (us): %s" % (arg, " ".join(out)), flush=True)
print("max row length per eighth:", [int(rl[q * n // 8:(q + 1) * n // 8].max()) for q in range(8)])
print("rows > 16 per eighth:", [int((rl[q * n // 8:(q + 1) * n // 8] > 16).sum()) for q in range(8)])
