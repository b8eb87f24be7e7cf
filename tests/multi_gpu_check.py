"""Multi-GPU parity check, launched one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py [--pc 1]
Every rank builds its share of the solver through the C ABI (NCCL communicator from a broadcast ncclUniqueId), rank 0
runs the oracle with N simulated ranks on the same mesh files and compares: boundary ids and node maps bit-exact per rank,
iteration count, residual history, per-rank solution."""
import argparse
import ctypes as C
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pc", type=int, default=0)
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--nel", default="4", help="elements per side, or nx,ny[,nz]")
    ap.add_argument("--degree", dest="N", type=int, default=5)
    ap.add_argument("--reduction", dest="r", type=int, default=2)
    ap.add_argument("--eps", type=float, default=0.04)
    ap.add_argument("--golden-only", action="store_true", help="compare with tests/golden/solve_histories.json only (no oracle run on the GPU box)")
    a = ap.parse_args()
    a.nel = tuple(int(x) for x in a.nel.split(",")) if "," in a.nel else int(a.nel)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import polynomial_reduction_with_full_domain_decomposition_preconditioner_b200 as pr
    holder = [None, None]
    if rank == 0:
        nccl = C.CDLL("libnccl.so.2")
        buf = C.create_string_buffer(128)
        assert nccl.ncclGetUniqueId(buf) == 0
        holder = [bytes(buf.raw), tempfile.mkdtemp(prefix="prfdd_mg_")]
        pr.mesh_generate_box(holder[1], a.dim, a.nel, a.N, world, a.eps, reduction=a.r)
    dist.broadcast_object_list(holder, src=0)
    uid, d = holder
    S = pr.Solver(d, poly_degree=a.N, poly_reduction=a.r, use_preconditioner=a.pc, proc_id=rank, num_procs=world, nccl_unique_id=uid)
    S.setup_problem(4)
    ok = True
    for solver_id in (0, 1):
        nit, hist = S.solve(solver_id)
        mine = dict(rank=rank, nit=nit, hist=hist, u=S.get_array("U"), nop=S.get_array("NODE_OF_POINT"), bn=S.get_array("BOUNDARY_NODES"),
                    aw=S.get_array("ASSEMBLED_WEIGHT"), nodes=S.query("NUM_GLOBAL_NODES"))
        if a.pc and solver_id == 0:
            mine.update(sub=dict(ids=S.get_array("SUB_ELEMENT_IDS"), deg=S.get_array("SUB_ELEMENT_DEGREE"), dof=S.get_array("SUB_DOF_NUM"),
                                 qptr=S.get_array("SUB_Q_PTR"), qcol=S.get_array("SUB_Q_COL"), qval=S.get_array("SUB_Q_VAL"),
                                 sizes=[S.query(k) for k in ("SUB_NUM_POINTS", "SUB_NUM_DOFS", "SUB_NUM_EXTENDED_DOFS", "SUP_NUM_DOFS", "SUP_NUM_EXTENDED_DOFS", "NUM_VALUES", "NUM_DOFS")],
                                 aptr=S.get_array("A_FEM_PTR"), acol=S.get_array("A_FEM_COL"), aval=S.get_array("A_FEM_VAL"), rows=S.get_array("AMG_LEVEL_ROWS"),
                                 nw=S.get_array("NORM_WEIGHT"), iw=S.get_array("INNER_WEIGHT")))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0 and a.golden_only:
            import json
            gold = [c for c in json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solve_histories.json")))
                    if (c["dim"], c["nel"], c["N"], c["r"], c["eps"], c["ranks"], c["solver"]) == (a.dim, list(a.nel) if isinstance(a.nel, tuple) else a.nel, a.N, a.r, a.eps, world, solver_id)]
            if solver_id == 0:
                assert gold, "no golden record for this case"
            for c in gold:
                ref = np.array(c["history"])
                assert gathered[0]["nit"] == c["iterations"], (gathered[0]["nit"], c["iterations"])
                assert np.abs(gathered[0]["hist"] - ref).max() <= 1e-9 * ref[0], np.abs(gathered[0]["hist"] - ref).max() / ref[0]
                print("solver_id %d: GPU iters %d = golden (oracle, build container) iters %d, max history diff / r0 %.2e, final relative residual %.2e"
                      % (solver_id, gathered[0]["nit"], c["iterations"], np.abs(gathered[0]["hist"] - ref).max() / ref[0], gathered[0]["hist"][-1] / gathered[0]["hist"][0]), flush=True)
        elif rank == 0:
            from oracle import domain as od
            W = od.DomainWorld(d, a.N, world)
            Sd = None
            if a.pc:
                from oracle import subdomain as osub
                Sd = osub.SubdomainWorld(W, d, a.N, a.r)
            else:
                W.use_preconditioner = False
            us = W.initial_function(4)
            f = W.new_vector(); W.stiffness_matrix(f, us)
            u = W.new_vector()
            (W.flexible_conjugate_gradient if solver_id == 0 else W.generalized_minimum_residual)(u, f, Sd)
            for g in gathered:
                R = W.ranks[g["rank"]]
                assert np.array_equal(g["nop"], R.local_node_idx), "node map differs on rank %d" % g["rank"]
                assert np.array_equal(g["bn"], R.boundary_nodes), "boundary ids differ on rank %d" % g["rank"]
                assert np.abs(g["aw"] - R.assembled_weight).max() < 1e-15
                assert g["nodes"] == W.num_global_nodes()
            if a.pc and solver_id == 0:
                import scipy.sparse as sp
                for g in gathered:
                    So, sub = Sd.ranks[g["rank"]], g["sub"]
                    assert np.array_equal(sub["ids"], So.elem_id) and np.array_equal(sub["deg"], So.elem_degree), "region differs on rank %d" % g["rank"]
                    assert np.array_equal(sub["dof"], So.dof_num), "region dof numbering differs on rank %d" % g["rank"]
                    assert sub["sizes"] == [So.num_points, So.sub_num_dofs, So.sub_num_extended_dofs, So.sup_num_dofs, So.sup_num_extended_dofs, So.num_values, So.num_dofs], (sub["sizes"], g["rank"])
                    assert np.array_equal(sub["qptr"], So.Q.ptr) and np.array_equal(sub["qcol"], So.Q.col), "region Q structure differs on rank %d" % g["rank"]
                    assert np.abs(sub["qval"] - So.Q.val).max() < 1e-14
                    assert np.array_equal(sub["nw"], So.norm_weight) and np.array_equal(sub["iw"], So.inner_weight)
                    A = sp.csr_matrix((sub["aval"], sub["acol"], sub["aptr"]), shape=(So.num_dofs, So.num_dofs))
                    assert (A != 0).nnz == (So.A_fem != 0).nnz, ((A != 0).nnz, (So.A_fem != 0).nnz)
                    assert abs(A - So.A_fem).max() <= 1e-11 * abs(So.A_fem).max(), abs(A - So.A_fem).max()
                    assert np.array_equal(sub["rows"], [L.n for L in So.amg.levels]), (sub["rows"], [L.n for L in So.amg.levels])
                print("per-rank region maps, Q, weights, low-order FEM matrix, AMG level sizes: identical to the oracle on all %d ranks" % world, flush=True)
            tol_it = 0 if a.pc else 2
            assert abs(gathered[0]["nit"] - W.num_iterations) <= tol_it, (gathered[0]["nit"], W.num_iterations)
            m = min(len(W.history), gathered[0]["hist"].size)
            rel = np.abs(gathered[0]["hist"][:m] - np.array(W.history[:m])) / np.array(W.history[:m])
            num = sum(np.linalg.norm(g["u"] - u[g["rank"]]) ** 2 for g in gathered) ** 0.5
            den = sum(np.linalg.norm(x) ** 2 for x in u) ** 0.5
            print("solver_id %d: GPU iters %d, oracle iters %d, max rel history diff %.2e, rel L2 solution diff %.2e"
                  % (solver_id, gathered[0]["nit"], W.num_iterations, rel.max(), num / den), flush=True)
            assert rel.max() < (1e-8 if a.pc else 5e-2)
            assert num / den < (1e-10 if a.pc else 1e-6)
            if a.pc:
                # the same numbers computed by the oracle in the build container (tests/golden/solve_histories.json), if this case is there
                import json
                gold = [c for c in json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solve_histories.json")))
                        if (c["dim"], c["nel"], c["N"], c["r"], c["eps"], c["ranks"], c["solver"]) == (a.dim, list(a.nel) if isinstance(a.nel, tuple) else a.nel, a.N, a.r, a.eps, world, solver_id)]
                for c in gold:
                    ref = np.array(c["history"])
                    assert gathered[0]["nit"] == c["iterations"] and np.abs(gathered[0]["hist"] - ref).max() <= 1e-9 * ref[0]
                    print("golden history (build container) matched: %d iterations" % c["iterations"], flush=True)
    S.close()
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK_OK world=%d pc=%d" % (world, a.pc), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
