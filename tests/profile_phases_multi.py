"""Phase breakdown of one multi-GPU solve (bench.py's weak-scaling workload: 16^3 elements of N=7 per GPU) with the reference's
timer keys.  Enabling the Timer fences every phase and disables graph replay: read the SHARES.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/profile_phases_multi.py"""
import ctypes as C
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import polynomial_reduction_with_full_domain_decomposition_preconditioner_b200 as pr
from bench import layout, NEL_PER_GPU, N_DEG, REDUCTION, TOL

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
uid, d = None, None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    holder = [None, None]
    if rank == 0:
        nccl = C.CDLL("libnccl.so.2")
        buf = C.create_string_buffer(128)
        assert nccl.ncclGetUniqueId(buf) == 0
        holder = [bytes(buf.raw), tempfile.mkdtemp(prefix="prfdd_ph_")]
    dist.broadcast_object_list(holder, src=0)
    uid, d = holder
else:
    d = tempfile.mkdtemp(prefix="prfdd_ph_")
P3 = layout(world)
nel = tuple(NEL_PER_GPU * p for p in P3)
if rank == 0:
    pr.mesh_generate_box(d, 3, nel, N_DEG, world, 0.0, reduction=REDUCTION)
if world > 1:
    dist.barrier()
stream = torch.cuda.Stream()
S = pr.Solver(d, stream=stream.cuda_stream, poly_degree=N_DEG, poly_reduction=REDUCTION, outer_tolerance=TOL, proc_id=rank, num_procs=world, nccl_unique_id=uid)
S.setup_problem(4)
for _ in range(3):
    S.solve(0)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter(); nit, hist = S.solve(0); torch.cuda.synchronize(); t_graph = time.perf_counter() - t0
S.timer("__enable__")
S.solve(0)
keys = ["domain.operator_application", "domain.inner_products", "domain.residual_norm", "domain.vector_operations", "subdomain.stitching",
        "subdomain.tree_construction.gpu_to_gpu", "subdomain.tree_construction.subdomain", "subdomain.tree_construction.assemble_coarse",
        "subdomain.tree_construction.superdomain", "subdomain.tree_exchange.subdomain", "subdomain.tree_exchange.superdomain",
        "subdomain.preconditioner", "subdomain.preconditioner.assemble_subdomain", "subdomain.preconditioner.assemble_composite",
        "subdomain.preconditioner.down_leg_gpu", "subdomain.preconditioner.unassemble_composite", "subdomain.preconditioner.unassemble_subdomain",
        "subdomain.operator_application", "subdomain.inner_products", "subdomain.residual_norm", "subdomain.vector_operations"]
lines = ["rank %d: graph solve %.2f ms, %d iterations, %d launches per preconditioner application; sizes: points %d sub_dofs %d ext %d sup_ext %d values %d dofs %d; AMG rows %s" % (
    rank, 1e3 * t_graph, len(hist) - 1, S.query("GPU_LAUNCHES_PER_PRECOND"), S.query("SUB_NUM_POINTS"), S.query("SUB_NUM_DOFS"), S.query("SUB_NUM_EXTENDED_DOFS"),
    S.query("SUP_NUM_EXTENDED_DOFS"), S.query("NUM_VALUES"), S.query("NUM_DOFS"), list(S.get_array("AMG_LEVEL_ROWS")))]
for k in keys:
    v = S.timer(k)
    if v >= 0:
        lines.append("   %-52s %9.3f ms" % (k, 1e3 * v))
S.timer("__disable__")
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, lines)
else:
    out = [lines]
if rank == 0:
    for o in out[:2] + out[-1:]:
        print("\n".join(o))
if world > 1:
    dist.destroy_process_group()
