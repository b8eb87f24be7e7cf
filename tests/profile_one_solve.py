"""One solve between cudaProfilerStart/Stop, for `ncu --profile-from-start off` (launch list or --set full of one kernel).
   python tests/profile_one_solve.py [nel]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import polynomial_reduction_with_full_domain_decomposition_preconditioner_b200 as pr

nel = int(sys.argv[1]) if len(sys.argv) > 1 else 16
d = tempfile.mkdtemp()
pr.mesh_generate_box(d, 3, nel, 7, 1, 0.0, reduction=3)
stream = torch.cuda.Stream()
S = pr.Solver(d, stream=stream.cuda_stream, poly_degree=7, poly_reduction=3, outer_tolerance=1e-8)
S.setup_problem(4)
for _ in range(2):
    S.solve(0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
nit, hist = S.solve(0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("iterations", len(hist) - 1, "final relative residual", hist[-1] / hist[0])
S.close()
