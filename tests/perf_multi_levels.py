"""Experiment (not a test): row-length statistics and per-level SpMV timings of the MULTI-rank composite hierarchy (rank 0 of a
torchrun launch; the other ranks idle).   torchrun ... tests/perf_multi_levels.py"""
import ctypes as C
import os
import sys
import tempfile
import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import polynomial_reduction_with_full_domain_decomposition_preconditioner_b200 as pr  # noqa: E402
from bench import layout, NEL_PER_GPU, N_DEG, REDUCTION, TOL  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
holder = [None, None]
if rank == 0:
    nccl = C.CDLL("libnccl.so.2")
    buf = C.create_string_buffer(128)
    assert nccl.ncclGetUniqueId(buf) == 0
    holder = [bytes(buf.raw), tempfile.mkdtemp(prefix="prfdd_ml_")]
dist.broadcast_object_list(holder, src=0)
uid, d = holder
nel = tuple(NEL_PER_GPU * p for p in layout(world))
if rank == 0:
    pr.mesh_generate_box(d, 3, nel, N_DEG, world, 0.0, reduction=REDUCTION)
dist.barrier()
S = pr.Solver(d, poly_degree=N_DEG, poly_reduction=REDUCTION, outer_tolerance=TOL, proc_id=rank, num_procs=world, nccl_unique_id=uid)
if rank == 0:
    nd = S.query("NUM_DOFS")
    A = sp.csr_matrix((S.get_array("A_FEM_VAL"), S.get_array("A_FEM_COL"), S.get_array("A_FEM_PTR")), shape=(nd, nd))
del S
torch.cuda.empty_cache()
dist.barrier()
if rank == 0:
    from test_amg_host import product_hierarchy
    from perf_reorder import time_level
    H, _ = product_hierarchy(pr, A, 2, 9)
    for l, lev in enumerate(H[:4]):
        Al = lev["A"].tocsr()
        rl = np.diff(Al.indptr)
        print("level %d rows %d avg %.1f max %d p50 %d p90 %d p99 %d p99.9 %d; rows > 2*avg: %d (%.2f%% of rows, %.1f%% of entries); bandwidth p50 %d p99 %d" % (
            l, Al.shape[0], rl.mean(), rl.max(), *np.percentile(rl, [50, 90, 99, 99.9]).astype(int), (rl > 2 * rl.mean()).sum(),
            100.0 * (rl > 2 * rl.mean()).mean(), 100.0 * rl[rl > 2 * rl.mean()].sum() / rl.sum(),
            *np.percentile(np.abs(Al.indices - np.repeat(np.arange(Al.shape[0]), rl)), [50, 99]).astype(int)), flush=True)
        time_level(Al, "natural")
        n = Al.shape[0]
        pad = (-n) % 32
        wl = np.concatenate([rl, np.zeros(pad, rl.dtype)]).reshape(-1, 32)
        print("   warp cost model (tpr 1): sum of per-warp max row length / sum of per-warp mean = %.2f; warps with max > 2*avg: %.1f%%" % (
            wl.max(1).sum() / wl.mean(1).sum(), 100.0 * (wl.max(1) > 2 * rl.mean()).mean()), flush=True)
        if l == 0:
            if os.environ.get("PRFDD_SAVE_LEVEL0"):
                np.savez(os.environ["PRFDD_SAVE_LEVEL0"], indptr=Al.indptr, indices=Al.indices, data=Al.data)
            from perf_micro import timeit_batch, P, L, sh
            ptr, col, val = (torch.from_numpy(a).cuda() for a in (Al.indptr.astype(np.int32), Al.indices.astype(np.int32), Al.data.astype(np.float64)))
            x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.zeros(n, dtype=torch.float64, device="cuda")
            for lo, hi in ((0, n - 1), (0, n // 2 - 1), (n // 2, n - 1), (0, n // 4 - 1), (n // 4, n // 2 - 1), (n // 2, 3 * n // 4 - 1), (3 * n // 4, n - 1)):
                t, _ = timeit_batch(lambda: L.prfdd_csr_multiply_range(P(y), P(ptr), P(col), P(val), P(x), C.c_int(lo), C.c_int(hi), C.c_int(1), sh))
                nz = int(Al.indptr[hi + 1] - Al.indptr[lo])
                print("   rows [%d, %d]: %.1f us, %.2f ns/row, entries/row %.2f, max row %d" % (lo, hi, t * 1e3, t * 1e6 / (hi - lo + 1), nz / (hi - lo + 1), rl[lo:hi + 1].max()), flush=True)
            order = np.argsort(rl, kind="stable")
            time_level(Al[order][:, order].tocsr(), "rows sorted by length")
dist.barrier()
dist.destroy_process_group()
