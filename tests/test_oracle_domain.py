"""Oracle-only checks of the Domain restatement (SURVEY 8c pins 1, 2, 6): Q^T = transpose(Q), QQ^T 1 =
multiplicity, boundary-first node order, operator == assembled sparse matrix properties (symmetry,
constants in the null space), partition independence of the unpreconditioned solve."""
import numpy as np
import pytest
from oracle import meshgen, domain


def _world(tmp, dim, nel, N, nr, eps):
    meshgen.generate(tmp, dim, nel, N, nranks=nr, eps=eps)
    W = domain.DomainWorld(tmp, N, nr)
    W.use_preconditioner = False
    return W


@pytest.mark.parametrize("dim,nel,N,nr", [(2, 4, 3, 1), (2, 4, 4, 2), (3, 2, 3, 2), (3, 4, 2, 8)])
def test_gather_scatter_maps(tmp_path, dim, nel, N, nr):
    W = _world(str(tmp_path), dim, nel, N, nr, 0.05)
    for r in W.ranks:
        Q, Qt = r.Q.to_scipy(), r.Qt.to_scipy()
        assert (Q.T != Qt).nnz == 0
        assert np.array_equal(np.asarray(Q.sum(axis=1)).ravel(), np.ones(r.num_local_points))
        # boundary nodes come first: every point flagged as process boundary maps below num_bdary_nodes
        if nr == 1:
            assert r.num_bdary_nodes == 0
        else:
            assert 0 < r.num_bdary_nodes < r.num_local_nodes
    # 1/assembled_weight == global multiplicity (= node_degree of the mesh files)
    allg = np.concatenate([r.glo_num for r in W.ranks])
    ids, counts = np.unique(allg, return_counts=True)
    mult = dict(zip(ids.tolist(), counts.tolist()))
    for r in W.ranks:
        for p in range(0, r.num_local_points, 7):
            assert abs(1.0 / r.assembled_weight[r.local_node_idx[p]] - mult[int(r.glo_num[p])]) < 1e-12


@pytest.mark.parametrize("dim,nel,N", [(2, 3, 4), (3, 2, 3)])
def test_operator_properties(tmp_path, dim, nel, N):
    W = _world(str(tmp_path), dim, nel, N, 1, 0.08)
    rng = np.random.default_rng(0)
    P = W.ranks[0].num_local_points
    u, v = [rng.standard_normal(P)], [rng.standard_normal(P)]
    Au, Av = W.new_vector(), W.new_vector()
    W.stiffness_matrix(Au, u); W.stiffness_matrix(Av, v)
    assert abs(Au[0] @ v[0] - u[0] @ Av[0]) < 1e-11 * abs(Au[0] @ v[0])        # symmetric
    one, A1 = [np.ones(P)], W.new_vector()
    W.stiffness_matrix(A1, one)
    assert np.abs(A1[0]).max() < 1e-11                                         # constants in the null space (element-wise)
    assert Au[0] @ u[0] > 0


def test_partition_independence(tmp_path):
    """Without the preconditioner the algorithm is partition independent: 1, 2 and 4 simulated ranks give
    the same iteration count and the same solution at the same global nodes."""
    res = {}
    for nr in (1, 2, 4):
        d = str(tmp_path / ("r%d" % nr))
        W = _world(d, 3, 4, 3, nr, 0.05)
        # partition-independent RHS: f = A u*, u* = smooth function (function_id 0)
        us = W.initial_function(0)
        f = W.new_vector(); W.stiffness_matrix(f, us)
        u = W.new_vector()
        W.flexible_conjugate_gradient(u, f)
        sol = {}
        for r, ur in zip(W.ranks, u):
            for g, val in zip(r.glo_num.tolist(), ur.tolist()):
                sol[g] = val
        res[nr] = (W.num_iterations, np.array(W.history), sol)
    assert res[1][0] == res[2][0] == res[4][0]
    for nr in (2, 4):
        assert np.allclose(res[nr][1], res[1][1], rtol=1e-9)
        keys = sorted(res[1][2])
        a = np.array([res[1][2][k] for k in keys]); b = np.array([res[nr][2][k] for k in keys])
        assert np.abs(a - b).max() < 1e-9 * np.abs(a).max()


def test_manufactured_solution_both_drivers(tmp_path):
    W = _world(str(tmp_path), 2, 4, 5, 2, 0.1)
    us = W.initial_function(4)
    f = W.new_vector(); W.stiffness_matrix(f, us)
    for drv in (W.flexible_conjugate_gradient, W.generalized_minimum_residual):
        u = W.new_vector()
        drv(u, f)
        err = np.sqrt(sum(((a - b) ** 2).sum() for a, b in zip(u, us)) / sum((b ** 2).sum() for b in us))
        assert W.history[-1] / W.history[0] < 1e-7 and err < 1e-5
        assert all(l.startswith("Iter ") for l in W.log)
