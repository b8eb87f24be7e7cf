"""Generates tests/golden/bench_histories.json: the ORACLE's PR-FDD PCG solves at the sizes bench.py measures
(BASELINE configs c2 / c3 and the weak- and strong-scaling meshes the driver runs), computed offline in the build
container -- a full-size oracle solve takes minutes to hours, far more than the GPU box's bench budget.  bench.py
compares the CUDA path's iteration count and residual history with these records inside the timed job and
prints the outcome in its "parity" block, at every GPU count.

    python tests/golden/make_bench_histories.py c2                   # one case by name
    python tests/golden/make_bench_histories.py w2 w4 w8 s1 s2 s4    # several
Existing records of other cases are kept; a case is keyed by (nel, N, r, eps, ranks, tolerance, coarsening)."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import meshgen, domain, subdomain, amg  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_histories.json")
TOL = 1.0e-8
# name: (nel tuple, N, r, eps, ranks)
CASES = {
    "c1": ((16, 16), 7, 3, 0.0, 1),             # BASELINE configs[0], 2D
    "c2": ((16, 16, 16), 7, 3, 0.0, 1),         # BASELINE configs[1]
    "c2k": ((16, 16, 16), 7, 3, 0.3, 1),        # c2 on the Kershaw-style deformed mesh (profile.sh:5-11 uses eps = 0.3)
    "w2": ((32, 16, 16), 7, 3, 0.0, 2),         # weak scaling, 16^3 per GPU
    "w4": ((32, 32, 16), 7, 3, 0.0, 4),
    "w8": ((32, 32, 32), 7, 3, 0.0, 8),         # = BASELINE configs[2] on 8 GPUs
    "s1": ((32, 32, 32), 7, 3, 0.0, 1),         # strong scaling of the fixed 32^3 mesh (configs[2] as written)
    "s2": ((32, 32, 32), 7, 3, 0.0, 2),
    "s4": ((32, 32, 32), 7, 3, 0.0, 4),
    "t2": ((8, 4, 4), 7, 3, 0.0, 2),            # small multi-rank cases for the tests
    "t1": ((4, 4, 4), 7, 3, 0.0, 1),
    "c4b": ((8, 8, 8), 9, 3, 0.0, 1),           # BASELINE configs[3] degrees (ladder 9/6/3/1) at bench size:  bench.py --degree 9 --reduction 3 --nel-per-gpu 8
    "c4w2": ((16, 8, 8), 9, 3, 0.0, 2),         # the same two on 2 ranks (bench.py --gpus 2 ... --nel-per-gpu 8): rings at every ladder degree
    "c5w2": ((16, 8, 8), 15, 7, 0.0, 2),
    "c5b": ((8, 8, 8), 15, 7, 0.0, 1),          # BASELINE configs[4] degree (ladder 15/8/1):                 bench.py --degree 15 --reduction 7 --nel-per-gpu 8
}


def key_of(rec):
    return (tuple(rec["nel"]), rec["N"], rec["r"], rec["eps"], rec["ranks"], rec["tolerance"], rec.get("coarsening", "hmis"))


def run(name, coarsening):
    nel, N, r, eps, ranks = CASES[name]
    dim = len(nel)
    d = tempfile.mkdtemp(prefix="prfdd_gold_")
    t0 = time.time()
    for n in subdomain.ladder(N, r):
        meshgen.generate(d, dim, nel, n, nranks=ranks, eps=eps)
    W = domain.DomainWorld(d, N, ranks)
    W.tolerance = TOL
    if hasattr(amg, "set_coarsening"):
        amg.set_coarsening(coarsening)
    Sd = subdomain.SubdomainWorld(W, d, N, r)
    t1 = time.time()
    us = W.initial_function(4); f = W.new_vector(); W.stiffness_matrix(f, us); u = W.new_vector()
    W.flexible_conjugate_gradient(u, f, Sd)
    t2 = time.time()
    rec = dict(name=name, dim=dim, nel=list(nel), N=N, r=r, eps=eps, ranks=ranks, tolerance=TOL, coarsening=coarsening,
               iterations=int(W.num_iterations), history=[float(h) for h in W.history], oracle_setup_s=round(t1 - t0, 1), oracle_solve_s=round(t2 - t1, 1))
    print(name, coarsening, "->", rec["iterations"], "iterations", rec["history"], "setup %.0f s solve %.0f s" % (t1 - t0, t2 - t1), flush=True)
    return rec


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("--")]
    coarsening = "hmis"
    for a in sys.argv[1:]:
        if a.startswith("--coarsening="):
            coarsening = a.split("=")[1]
    for name in names:
        rec = run(name, coarsening)
        recs = json.load(open(OUT)) if os.path.exists(OUT) else []   # re-read: several generators may run side by side
        recs = [x for x in recs if key_of(x) != key_of(rec)] + [rec]
        recs.sort(key=lambda x: (x["name"], x.get("coarsening", "")))
        json.dump(recs, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
