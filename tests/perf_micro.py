"""Kernel micro-benchmarks on the GPU box (not a test): CUDA-event timing with an L2 flush between launches.
    python tests/perf_micro.py [ax|spmv|stream]"""
import ctypes as C
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import polynomial_reduction_with_full_domain_decomposition_preconditioner_b200 as pr  # noqa: E402

L = pr.lib()
stream = torch.cuda.Stream()
sh = C.c_void_p(stream.cuda_stream)
flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
PEAK = 6451.8


def P(t):
    return C.c_void_p(t.data_ptr())


def timeit(fn, reps=20, warm=3):
    ts = []
    with torch.cuda.stream(stream):
        for it in range(reps + warm):
            flush.fill_(float(it))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream)
            stream.synchronize()
            if it >= warm:
                ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def timeit_batch(fn, reps=20, warm=3):
    """`reps` back-to-back launches between one event pair, no L2 flush: for operands larger than L2 (a flush leaves dirty lines
    whose write-back is charged to the kernel, and a per-launch event pair adds ~6-10 us); returns (avg, avg)"""
    with torch.cuda.stream(stream):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        stream.synchronize()
    t = e0.elapsed_time(e1) / reps
    return t, t


def bench_ax():
    E, n = 4096, 8
    Pn = E * n ** 3
    u = torch.rand(Pn, dtype=torch.float64, device="cuda")
    G = [torch.rand(Pn, dtype=torch.float64, device="cuda") for _ in range(6)]
    Au = torch.empty(Pn, dtype=torch.float64, device="cuda")
    z = np.zeros(n); w = np.zeros(n); D = np.zeros(n * n)
    L.prfdd_zwgll(z.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), C.c_int(n))
    L.prfdd_dgll(D.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), C.c_int(n))
    Dd = torch.from_numpy(D).cuda()
    gp = (C.c_void_p * 6)(*[t.data_ptr() for t in G])
    med, mn = timeit(lambda: L.prfdd_stiffness_matrix(P(Au), P(u), P(Dd), gp, C.c_int(E), C.c_int(n), C.c_int(3), sh))
    gb = 64.0 * Pn / 1e9
    print("ax3d n=8 E=4096 variant=%s: median %.2f us (%.0f GB/s, %.2f of peak), min %.2f us" % (os.environ.get("PRFDD_AX_VARIANT", "0"), med * 1e3, gb / (med * 1e-3), gb / (med * 1e-3) / PEAK, mn * 1e3))


def bench_spmv():
    import scipy.sparse as sp
    n = 111
    I = sp.eye(n); T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    A = (sp.kron(sp.kron(T, I), I) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(I, I), T)).tocsr(); A.sort_indices()
    nr = A.shape[0]
    ptr, col, val = (torch.from_numpy(a).cuda() for a in (A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)))
    x = torch.rand(nr, dtype=torch.float64, device="cuda"); y = torch.empty(nr, dtype=torch.float64, device="cuda")
    r = torch.rand(nr, dtype=torch.float64, device="cuda"); ds = torch.rand(nr, dtype=torch.float64, device="cuda"); uu = torch.zeros(nr, dtype=torch.float64, device="cuda")
    gb = (12.0 * A.nnz + 4.0 * (nr + 1) + 8.0 * nr * 2) / 1e9
    for tpr in (1, 2, 4, 8):
        med, mn = timeit(lambda: L.prfdd_csr_multiply(P(y), P(ptr), P(col), P(val), P(x), C.c_int(nr), C.c_int(tpr), sh))
        print("csr_multiply 7pt 111^3 tpr=%d: median %.2f us (%.0f GB/s, %.2f of peak)" % (tpr, med * 1e3, gb / (med * 1e-3), gb / (med * 1e-3) / PEAK))
    gb2 = (12.0 * A.nnz + 4.0 * (nr + 1) + 8.0 * nr * 5) / 1e9
    for tpr in (2, 4, 8):
        med, mn = timeit(lambda: L.prfdd_cheby_step(P(uu), P(y), P(ptr), P(col), P(val), P(x), P(r), P(ds), C.c_double(0.5), C.c_int(1), C.c_int(0), C.c_int(nr), C.c_int(tpr), sh))
        print("cheby_step   7pt 111^3 tpr=%d: median %.2f us (%.0f GB/s, %.2f of peak)" % (tpr, med * 1e3, gb2 / (med * 1e-3), gb2 / (med * 1e-3) / PEAK))


def bench_spmv_levels():
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    for nr, k in ((480000, 26), (98000, 66), (13000, 80), (1800, 60)):
        # banded random pattern: k entries per row within a window of 4k columns around the diagonal
        cols = (np.arange(nr)[:, None] + rng.integers(-2 * k, 2 * k, size=(nr, k))) % nr
        cols.sort(axis=1)
        ptr = (np.arange(nr + 1) * k).astype(np.int32)
        col = cols.ravel().astype(np.int32); val = rng.standard_normal(nr * k)
        dptr, dcol, dval = (torch.from_numpy(a).cuda() for a in (ptr, col, val))
        x = torch.rand(nr, dtype=torch.float64, device="cuda"); y = torch.empty(nr, dtype=torch.float64, device="cuda")
        gb = (12.0 * nr * k + 4.0 * (nr + 1) + 16.0 * nr) / 1e9
        out = []
        for tpr in (1, 2, 4, 8, 16, 32):
            med, mn = timeit(lambda: L.prfdd_csr_multiply(P(y), P(dptr), P(dcol), P(dval), P(x), C.c_int(nr), C.c_int(tpr), sh), reps=10)
            out.append("tpr%d %.1fus(%.0fGB/s)" % (tpr, med * 1e3, gb / (med * 1e-3)))
        print("rows %d nnz/row %d: %s" % (nr, k, "  ".join(out)))


def bench_stream():
    n = 2097152
    a, b, c, d = (torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(4))
    med, mn = timeit(lambda: L.prfdd_vector_vector_addition(P(c), C.c_double(1.0), P(a), C.c_double(2.0), P(b), C.c_int(n), sh))
    print("axpby 2M: median %.2f us (%.0f GB/s)" % (med * 1e3, 24.0 * n / 1e9 / (med * 1e-3)))
    big = 64 * 1024 * 1024
    a, b, c = (torch.rand(big, dtype=torch.float64, device="cuda") for _ in range(3))
    med, mn = timeit(lambda: L.prfdd_vector_vector_addition(P(c), C.c_double(1.0), P(a), C.c_double(2.0), P(b), C.c_int(big), sh), reps=5)
    print("axpby 64M: median %.2f us (%.0f GB/s)" % (med * 1e3, 24.0 * big / 1e9 / (med * 1e-3)))
    med, mn = timeit(lambda: c.copy_(a), reps=5)
    print("torch copy 64M doubles (on torch stream): median %.2f us (%.0f GB/s)" % (med * 1e3, 16.0 * big / 1e9 / (med * 1e-3)))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("ax", "all"):
        bench_ax()
    if what in ("spmv", "all"):
        bench_spmv()
    if what in ("levels", "all"):
        bench_spmv_levels()
    if what in ("stream", "all"):
        bench_stream()
