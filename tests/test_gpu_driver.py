"""-m gpu: the `poisson` driver binary (reference command line, poisson.cpp:40-81; log lines poisson.cpp:226-251 and the
"Timings:" table poisson.cpp:385-401) run as a subprocess on one GPU; its iteration count is checked against the oracle."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu

from oracle import domain as odomain, subdomain as osub  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINARY = os.path.join(ROOT, "polynomial_reduction_with_full_domain_decomposition_preconditioner_b200", "poisson")


@pytest.mark.parametrize("solver_id", [0, 1])
def test_poisson_driver(prfdd, tmp_path, solver_id):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    assert os.path.exists(BINARY), "driver binary missing: run __graft_entry__.build()"
    d = str(tmp_path)
    prfdd.mesh_generate_box(d, 3, 3, 4, 1, 0.03, reduction=3)
    env = dict(os.environ, PRFDD_TIMINGS="1", PRFDD_TOLERANCE="1e-8", PRFDD_OUTPUT=os.path.join(d, "domain"))
    out = subprocess.run([BINARY, d, "4", "3", "1", "0", str(solver_id)], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    W = odomain.DomainWorld(d, 4, 1)
    W.tolerance = 1.0e-8
    Sd = osub.SubdomainWorld(W, d, 4, 3)
    us = W.initial_function(4); f = W.new_vector(); W.stiffness_matrix(f, us); u = W.new_vector()
    if solver_id == 0:
        W.flexible_conjugate_gradient(u, f, Sd)
    else:
        W.generalized_minimum_residual(u, f, Sd)
    m = re.search(r"^Iterations: (\d+)$", out.stdout, re.M)
    assert m and int(m.group(1)) == W.num_iterations, out.stdout
    assert ('Solver type: "%s"' % ("FCG" if solver_id == 0 else "GMRES")) in out.stdout
    for row in ("Total", "Inner products", "Residual norm", "Vector operations", "Operator application", "Tree construction",
                "Tree exchange", "Subdomain stitching", "Subdomain solver"):
        assert re.search(r"^%s\s+=\s+[0-9.]+ s" % row, out.stdout, re.M), (row, out.stdout)
    total = float(re.search(r"^Total\s+=\s+([0-9.]+) s", out.stdout, re.M).group(1))
    assert total > 0.0
    # field output (u_star, f, u on the low-order cell mesh): 27 elements of 5^3 points, 4^3 cells each
    head = open(os.path.join(d, "domain_0.vtk")).read(400)
    assert "DATASET UNSTRUCTURED_GRID" in head and "POINTS %d double" % (27 * 125) in head
    text = open(os.path.join(d, "domain_0.vtk")).read()
    assert "CELLS %d %d" % (27 * 64, 27 * 64 * 9) in text and all(("SCALARS %s double 1" % k) in text for k in ("u_star", "f", "u"))
    # Subdomain::output (subdomain.tpp:4648-4791): the rank's region; on one rank = the 27 own elements at degree 4, fields degree / f / u
    sub = open(os.path.join(d, "domain_subdomain_0.vtk")).read()
    assert "POINTS %d double" % (27 * 125) in sub and "CELLS %d %d" % (27 * 64, 27 * 64 * 9) in sub
    assert all(("SCALARS %s double 1" % k) in sub for k in ("degree", "f", "u"))
    vals = sub.split("SCALARS degree double 1\nLOOKUP_TABLE default\n")[1].split("SCALARS")[0].split()
    assert len(vals) == 27 * 125 and set(vals) == {"4"}
    uvals = [float(v) for v in sub.split("SCALARS u double 1\nLOOKUP_TABLE default\n")[1].split()]
    assert len(uvals) == 27 * 125 and any(abs(v) > 0 for v in uvals)


def test_usage_message():
    out = subprocess.run([BINARY], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "ERROR: Use as 'poisson <directory>" in out.stdout   # quit() exits EXIT_SUCCESS (config.hpp:57-62)
