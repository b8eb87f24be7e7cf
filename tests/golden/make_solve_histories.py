"""Generates tests/golden/solve_histories.json: iteration counts and residual histories of the ORACLE's PR-FDD solves on small seeded
meshes, computed in the build container.  The GPU tests compare the CUDA path with these numbers as well as with the oracle
re-run on the GPU box, so a host-library difference between the two machines cannot hide behind the oracle.
    python tests/golden/make_solve_histories.py"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import meshgen, domain, subdomain  # noqa: E402

CASES = [  # dim, nel, N, r, eps, ranks, outer solver (0 FCG / 1 GMRES)
    (2, 8, 7, 3, 0.05, 1, 0), (3, 3, 4, 3, 0.05, 1, 0), (3, 4, 7, 3, 0.0, 1, 0), (3, 4, 7, 3, 0.0, 1, 1),
    (3, 2, 9, 3, 0.03, 1, 0), (2, 8, 4, 3, 0.04, 2, 0), (3, 4, 3, 2, 0.04, 2, 0),   # the 2-rank cases are those of tests/test_gpu_multi.py
]

# `python make_solve_histories.py dim nel N r eps ranks solver` computes ONE extra case and adds it to the file (the 8-rank c4 case
# takes the oracle many minutes: tests/multi_gpu_check.py --golden-only compares an 8-GPU run with it without re-running the oracle)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "solve_histories.json")
EXTRA = [(3, 8, 9, 3, 0.03, 8, 0)]   # BASELINE configs[3] (c4) on 8 ranks: 2x2x2 blocks of 4^3 elements, ladder 9/6/3/1
out = []
if len(sys.argv) == 8:
    a = sys.argv[1:]
    CASES = [(int(a[0]), int(a[1]), int(a[2]), int(a[3]), float(a[4]), int(a[5]), int(a[6]))]
    out = [c for c in json.load(open(OUT)) if (c["dim"], c["nel"], c["N"], c["r"], c["eps"], c["ranks"], c["solver"]) != CASES[0]]
for dim, nel, N, r, eps, ranks, solver in CASES:
    d = tempfile.mkdtemp()
    for n in subdomain.ladder(N, r):
        meshgen.generate(d, dim, nel, n, nranks=ranks, eps=eps)
    W = domain.DomainWorld(d, N, ranks)
    Sd = subdomain.SubdomainWorld(W, d, N, r)
    us = W.initial_function(4); f = W.new_vector(); W.stiffness_matrix(f, us); u = W.new_vector()
    (W.flexible_conjugate_gradient if solver == 0 else W.generalized_minimum_residual)(u, f, Sd)
    out.append(dict(dim=dim, nel=nel, N=N, r=r, eps=eps, ranks=ranks, solver=solver, iterations=int(W.num_iterations),
                    history=[float(h) for h in W.history]))
    print(out[-1]["dim"], nel, N, r, ranks, solver, "->", W.num_iterations, "iterations")
json.dump(out, open(OUT, "w"), indent=1)
