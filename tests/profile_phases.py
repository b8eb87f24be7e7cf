"""Phase breakdown of one solve with the reference's timer keys (timer.tpp / poisson.cpp:253-401): enabling the Timer
brackets every phase with a stream synchronise (and disables graph replay), so absolute numbers carry sync overhead;
the SHARES show where the time goes.   python tests/profile_phases.py [nel]"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import polynomial_reduction_with_full_domain_decomposition_preconditioner_b200 as pr

nel = int(sys.argv[1]) if len(sys.argv) > 1 else 16
d = tempfile.mkdtemp()
pr.mesh_generate_box(d, 3, nel, 7, 1, 0.0, reduction=3)
stream = torch.cuda.Stream()
S = pr.Solver(d, stream=stream.cuda_stream, poly_degree=7, poly_reduction=3, outer_tolerance=1e-8)
S.setup_problem(4)
print("AMG level rows", S.get_array("AMG_LEVEL_ROWS"), "nnz", S.get_array("AMG_LEVEL_NNZ"))
for _ in range(3):
    S.solve(0)
torch.cuda.synchronize(); t0 = time.perf_counter(); nit, hist = S.solve(0); torch.cuda.synchronize(); t_graph = time.perf_counter() - t0
print("graph solve: %.2f ms, %d iterations, launches per preconditioner application %d" % (1e3 * t_graph, len(hist) - 1, S.query("GPU_LAUNCHES_PER_PRECOND")))
S.timer("__enable__")
S.solve(0)
keys = ["domain.operator_application", "domain.inner_products", "domain.residual_norm", "domain.vector_operations", "subdomain.stitching",
        "subdomain.tree_construction.gpu_to_gpu", "subdomain.preconditioner", "subdomain.preconditioner.assemble_subdomain", "subdomain.preconditioner.down_leg_gpu",
        "subdomain.preconditioner.unassemble_subdomain", "subdomain.operator_application", "subdomain.inner_products", "subdomain.residual_norm", "subdomain.vector_operations"]
tot = 0.0
rows = []
for k in keys:
    v = S.timer(k)
    if v >= 0:
        rows.append((k, v))
for k, v in rows:
    print("%-50s %9.3f ms" % (k, 1e3 * v))
S.timer("__disable__")
