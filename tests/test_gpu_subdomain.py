"""-m gpu: the PR-FDD preconditioned solve on the GPU (one rank) against the oracle on the same mesh files.
Bit-exact integer maps (region dof numbering, Q structure, AMG level sizes), low-order FEM matrix to 1e-12,
every building block (composite operator, V-cycle, low-order preconditioner, one inner Krylov solve) to 1e-10,
and the headline: same outer PCG iteration count, same residual history, solution within 1e-10 relative L2
(the north-star tolerance), for both inner solvers and both outer drivers."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

from oracle import domain as odomain, subdomain as osub  # noqa: E402

CASES = [(2, 4, 4, 3, 0.0), (2, 8, 7, 3, 0.05), (2, 16, 7, 3, 0.0), (3, 3, 4, 3, 0.05), (3, 2, 7, 6, 0.03), (3, 4, 7, 3, 0.0), (3, 3, 2, 1, 0.04), (2, 5, 1, 1, 0.1),
         (3, 2, 9, 3, 0.03), (3, 2, 15, 7, 0.02), (2, 3, 15, 7, 0.0)]   # configs 4 / 5 degrees: ladders 9/6/3/1 and 15/8/1


def _need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.mark.parametrize("dim,nel,N,r,eps", CASES)
def test_preconditioned_parity(prfdd, tmp_path, dim, nel, N, r, eps):
    _need_gpu()
    d = str(tmp_path)
    prfdd.mesh_generate_box(d, dim, nel, N, 1, eps, reduction=r)
    W = odomain.DomainWorld(d, N, 1)
    Sd = osub.SubdomainWorld(W, d, N, r)
    So = Sd.ranks[0]
    S = prfdd.Solver(d, poly_degree=N, poly_reduction=r)
    # integer maps: bit exact
    assert S.query("SUB_NUM_POINTS") == So.num_points and S.query("SUB_NUM_DOFS") == So.sub_num_dofs
    assert S.query("SUB_NUM_EXTENDED_DOFS") == So.sub_num_extended_dofs and S.query("NUM_DOFS") == So.num_dofs
    assert S.query("NUM_VALUES") == So.num_values and S.query("SUP_NUM_EXTENDED_DOFS") == 0
    assert np.array_equal(S.get_array("SUB_DOF_NUM"), So.dof_num)
    assert np.array_equal(S.get_array("SUB_ELEMENT_IDS"), So.elem_id) and np.array_equal(S.get_array("SUB_ELEMENT_DEGREE"), So.elem_degree)
    assert np.array_equal(S.get_array("SUB_Q_PTR"), So.Q.ptr) and np.array_equal(S.get_array("SUB_Q_COL"), So.Q.col)
    assert np.array_equal(S.get_array("SUB_Q_VAL"), So.Q.val)
    assert np.array_equal(S.get_array("NORM_WEIGHT"), So.norm_weight) and np.array_equal(S.get_array("INNER_WEIGHT"), So.inner_weight)
    # low-order FEM matrix and AMG hierarchy
    A = sp.csr_matrix((S.get_array("A_FEM_VAL"), S.get_array("A_FEM_COL"), S.get_array("A_FEM_PTR")), shape=(So.num_dofs, So.num_dofs))
    assert (A != 0).nnz == (So.A_fem != 0).nnz
    assert abs(A - So.A_fem).max() <= 1e-12 * abs(So.A_fem).max()
    assert np.array_equal(S.get_array("AMG_LEVEL_ROWS"), [L.n for L in So.amg.levels])
    assert np.array_equal(S.get_array("AMG_LEVEL_NNZ"), [L.A.nnz for L in So.amg.levels])
    assert np.allclose(S.get_array("AMG_CHEBY_COEFS"), np.concatenate([L.coefs for L in So.amg.levels]), rtol=1e-8)
    # building blocks
    rng = np.random.default_rng(3)
    nvl = So.num_values
    x = rng.standard_normal(nvl)
    ref = np.zeros(nvl); Sd.stiffness_matrix(So, ref, x)
    assert np.abs(S.apply("SUB_STIFFNESS", x) - ref).max() <= 1e-12 * np.abs(ref).max()
    b = rng.standard_normal(So.num_dofs)
    ref = So.amg.vcycle(b, 1)
    got = S.apply("VCYCLE", b)
    assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max()
    ref = np.zeros(nvl); Sd.low_order_preconditioner(So, ref, x)
    got = S.apply("LOW_ORDER", x)
    assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max()
    P_ = W.ranks[0].num_local_points
    rr = rng.standard_normal(P_)
    for ptype, fn in ((1, Sd.generalized_minimum_residual), (0, Sd.flexible_conjugate_gradient)):
        zo = [np.zeros(P_)]; fn(zo, [rr])
        Sx = prfdd.Solver(d, poly_degree=N, poly_reduction=r, preconditioner_type=ptype)
        for _ in range(2):   # second call replays the captured CUDA graph
            zg = Sx.apply("PRECONDITIONER", rr)
            assert np.linalg.norm(zg - zo[0]) <= 1e-9 * np.linalg.norm(zo[0])
        Sx.close()
    # the solve
    S.setup_problem(4)
    us = W.initial_function(4)
    f = W.new_vector(); W.stiffness_matrix(f, us)
    for ptype in (1, 0):
        W.preconditioner_type = ptype
        Sx = prfdd.Solver(d, poly_degree=N, poly_reduction=r, preconditioner_type=ptype)
        Sx.setup_problem(4)
        for solver_id, drv in ((0, W.flexible_conjugate_gradient), (1, W.generalized_minimum_residual)):
            u = W.new_vector(); drv(u, f, Sd)
            nit, hist = Sx.solve(solver_id)
            assert nit == W.num_iterations, (nit, W.num_iterations)
            assert np.abs(hist - np.array(W.history)).max() <= 1e-9 * W.history[0]
            ug = Sx.get_array("U")
            assert np.linalg.norm(ug - u[0]) <= 1e-10 * np.linalg.norm(u[0])
            assert np.linalg.norm(ug - us[0]) <= 1e-5 * np.linalg.norm(us[0])
            assert hist[-1] / hist[0] < 1e-7
        Sx.close()
    # graph and no-graph paths agree bit for bit
    S1 = prfdd.Solver(d, poly_degree=N, poly_reduction=r, use_cuda_graph=0); S1.setup_problem(4)
    S2 = prfdd.Solver(d, poly_degree=N, poly_reduction=r, use_cuda_graph=1); S2.setup_problem(4)
    n1, h1 = S1.solve(0); n2, h2 = S2.solve(0)
    assert n1 == n2 and np.array_equal(h1, h2) and np.array_equal(S1.get_array("U"), S2.get_array("U"))
    assert S2.query("GPU_LAUNCHES_PER_PRECOND") > 0
    S1.close(); S2.close(); S.close()


def _golden_cases():
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solve_histories.json")))
    return [c for c in g if c["ranks"] == 1]


@pytest.mark.parametrize("case", _golden_cases(), ids=lambda c: "%dd_nel%d_N%d_r%d_s%d" % (c["dim"], c["nel"], c["N"], c["r"], c["solver"]))
def test_solve_matches_golden_history(prfdd, tmp_path, case):
    """iteration count and residual history against tests/golden/solve_histories.json (the oracle's numbers computed in the build
    container, tests/golden/make_solve_histories.py) -- independent of the oracle re-run on this machine"""
    _need_gpu()
    d = str(tmp_path)
    prfdd.mesh_generate_box(d, case["dim"], case["nel"], case["N"], 1, case["eps"], reduction=case["r"])
    S = prfdd.Solver(d, poly_degree=case["N"], poly_reduction=case["r"])
    S.setup_problem(4)
    nit, hist = S.solve(case["solver"])
    ref = np.array(case["history"])
    assert nit == case["iterations"] and len(hist) == len(ref)
    assert np.abs(hist - ref).max() <= 1e-9 * ref[0]
    S.close()


@pytest.mark.parametrize("dim,nel,N,r,eps", [(3, 4, 7, 3, 0.0), (3, 3, 4, 3, 0.05), (2, 16, 7, 3, 0.0)])
def test_fp32_vcycle_parity(prfdd, tmp_path, dim, nel, N, r, eps):
    """SURVEY 8f N3, the AMG half: `Float float` (AMG/config.hpp:4) -- FP32 hierarchy, smoother and cycle under the FP64 Krylov
    methods.  The oracle restates it with its own FP32 loops (liboracle_f32.so); FP32 sums are order dependent, so the bar is the
    FP32 one: one V-cycle to 2e-5 of its norm, the SAME outer iteration count as the FP32 oracle and as the FP64 solve, residual
    history to 1e-4, and the converged solution as close to the manufactured one as the FP64 solve's."""
    _need_gpu()
    d = str(tmp_path)
    prfdd.mesh_generate_box(d, dim, nel, N, 1, eps, reduction=r)
    W = odomain.DomainWorld(d, N, 1)
    Sd32 = osub.SubdomainWorld(W, d, N, r, amg_precision="float")
    So = Sd32.ranks[0]
    S = prfdd.Solver(d, poly_degree=N, poly_reduction=r, amg_precision=1)
    assert np.array_equal(S.get_array("AMG_LEVEL_ROWS"), [L.n for L in So.amg.levels])
    rng = np.random.default_rng(5)
    b = rng.standard_normal(So.num_dofs)
    ref = So.amg.vcycle(b, 1).astype(np.float64)
    got = S.apply("VCYCLE", b)
    assert np.linalg.norm(got - ref) <= 2e-5 * np.linalg.norm(ref)
    S.setup_problem(4)
    us = W.initial_function(4)
    f = W.new_vector(); W.stiffness_matrix(f, us)
    u = W.new_vector(); W.flexible_conjugate_gradient(u, f, Sd32)
    h32 = np.array(W.history)
    nit, hist = S.solve(0)
    assert nit == W.num_iterations and len(hist) == len(h32)
    assert np.max(np.abs(hist - h32) / h32) <= 1e-4
    ug = S.get_array("U")
    S64 = prfdd.Solver(d, poly_degree=N, poly_reduction=r)
    S64.setup_problem(4)
    n64, h64 = S64.solve(0)
    assert n64 == nit                                     # the flexible outer method does not notice the FP32 preconditioner
    e32 = np.linalg.norm(ug - us[0]) / np.linalg.norm(us[0])
    e64 = np.linalg.norm(S64.get_array("U") - us[0]) / np.linalg.norm(us[0])
    assert e32 <= 2.0 * e64 + 1e-9
    S.close(); S64.close()


@pytest.mark.parametrize("dim,nel,N,r,eps,inner", [(3, 4, 7, 3, 0.0, 1), (2, 8, 7, 3, 0.05, 1), (3, 3, 4, 3, 0.05, 0)])
def test_outer_fcg_as_one_graph(prfdd, tmp_path, dim, nel, N, r, eps, inner):
    """prfdd_options.device_outer_loop: from the second solve on the same buffers the outer flexible CG runs as ONE CUDA graph whose
    loop is a conditional WHILE node (convergence test on the device, no host round trip per iteration).  Same launches in the same
    order as the host-driven loop: iteration count, residual history and solution must be IDENTICAL to the first (host-driven)
    solve, bit for bit, and to a solver that never uses the graph."""
    _need_gpu()
    d = str(tmp_path)
    prfdd.mesh_generate_box(d, dim, nel, N, 1, eps, reduction=r)
    S = prfdd.Solver(d, poly_degree=N, poly_reduction=r, preconditioner_type=inner, outer_tolerance=1e-9, device_outer_loop=1)
    S.setup_problem(4)
    nit1, h1 = S.solve(0); u1 = S.get_array("U").copy()          # host-driven loop (warms every lazily built piece of state)
    l0 = prfdd.lib().prfdd_launch_count()
    nit2, h2 = S.solve(0); u2 = S.get_array("U").copy()          # builds and runs the graph
    l1 = prfdd.lib().prfdd_launch_count()
    nit3, h3 = S.solve(0); u3 = S.get_array("U").copy()          # replays it
    l2 = prfdd.lib().prfdd_launch_count()
    assert nit1 == nit2 == nit3 and nit1 >= 2
    assert np.array_equal(h1, h2) and np.array_equal(h1, h3)
    assert np.array_equal(u1, u2) and np.array_equal(u1, u3)
    assert l1 - l0 == l2 - l1 > 0                                 # the launches inside the graph are counted
    S.close()
    T = prfdd.Solver(d, poly_degree=N, poly_reduction=r, preconditioner_type=inner, outer_tolerance=1e-9, device_outer_loop=0)
    T.setup_problem(4)
    for _ in range(2):
        nit, h = T.solve(0)
        assert nit == nit1 and np.array_equal(h, h1) and np.array_equal(T.get_array("U"), u1)
    # iteration cap: the graph stops where the host loop stops
    T.close()
    for loop in (0, 1):
        Q = prfdd.Solver(d, poly_degree=N, poly_reduction=r, preconditioner_type=inner, outer_tolerance=1e-30, outer_max_iterations=2, device_outer_loop=loop)
        Q.setup_problem(4)
        res = [Q.solve(0) for _ in range(2)]
        assert res[0][0] == res[1][0] == 2 and np.array_equal(res[0][1], res[1][1]) and len(res[1][1]) == 3
        if loop == 0:
            cap_hist = res[0][1]
        else:
            assert np.array_equal(res[1][1], cap_hist)
        Q.close()


def test_solve_host_matches_the_device_resident_solve(prfdd, tmp_path):
    """prfdd_solver_solve_host (right-hand side from and solution to HOST buffers inside the call): identical iterates to the
    device-resident solve, bit for bit, for page-locked and pageable buffers and both outer drivers"""
    _need_gpu()
    import torch
    d = str(tmp_path)
    prfdd.mesh_generate_box(d, 3, 4, 7, 1, 0.0, reduction=3)
    S = prfdd.Solver(d, poly_degree=7, poly_reduction=3, outer_tolerance=1e-9)
    S.setup_problem(4)
    f = S.get_array("F")
    for solver_id in (0, 1):
        nit0, h0 = S.solve(solver_id)
        u0 = S.get_array("U").copy()
        fh = torch.from_numpy(f.copy()).pin_memory(); uh = torch.full_like(fh, 7.0).pin_memory()
        nit, h = S.solve_host(fh.numpy(), uh.numpy(), solver_id)
        assert nit == nit0 and np.array_equal(h, h0) and np.array_equal(uh.numpy(), u0)
        up = np.full_like(f, 5.0)                               # pageable: staged through the solver's pinned buffers
        nit, h = S.solve_host(f.copy(), up, solver_id)
        assert nit == nit0 and np.array_equal(up, u0)
    S.close()
