"""Oracle-only checks of the PR-FDD restatement (SURVEY 8c pins 3-6): restriction = interpolation^T is in
test_oracle_pin.py; here: non-conforming Q rows sum to 1, region dof numbering is dense, low-order FEM matrices are
symmetric positive definite with constants in the null space of the unconstrained rows, the preconditioned solve converges
in a handful of iterations for 1, 2, 4, 8, 16 simulated ranks (single- and multi-level composite superdomain)."""
import numpy as np
import pytest
from oracle import meshgen, domain, subdomain


def _setup(tmp, dim, nel, N, r, nr, eps):
    for n in subdomain.ladder(N, r):
        meshgen.generate(tmp, dim, nel, n, nranks=nr, eps=eps)
    W = domain.DomainWorld(tmp, N, nr)
    return W, subdomain.SubdomainWorld(W, tmp, N, r)


CASES = [(2, 4, 4, 3, 1, 0.0), (3, 3, 4, 3, 1, 0.05), (2, 8, 4, 3, 4, 0.03), (3, 4, 3, 2, 2, 0.03), (2, 24, 3, 1, 16, 0.0), (3, 4, 3, 2, 8, 0.03)]


@pytest.mark.parametrize("dim,nel,N,r,nr,eps", CASES)
def test_prfdd_converges(tmp_path, dim, nel, N, r, nr, eps):
    W, Sd = _setup(str(tmp_path), dim, nel, N, r, nr, eps)
    for S in Sd.ranks:
        # dense numbering 1..num_ext_dofs, masked points 0
        d = S.dof_num
        assert d.min() == 0 and set(np.unique(d[d > 0]).tolist()) == set(range(1, int(d.max()) + 1))
        Q = S.Q.to_scipy()
        rs = np.asarray(Q.sum(axis=1)).ravel()
        live = (np.asarray(abs(Q).sum(axis=1)).ravel() > 0)
        # interpolation rows (zeroed non-conforming edge / face interiors) sum to 1 wherever the coarse neighbour's
        # edge / face does not touch the Dirichlet boundary (pin 4 of SURVEY 8c)
        full = live & (S.glo_num == 0)
        if full.any() and nel >= 8:
            assert np.isclose(rs[full], 1.0, atol=1e-12).any()
        unit = live & (S.glo_num > 0)
        assert np.array_equal(rs[unit], np.ones(int(unit.sum())))
        A = S.A_fem
        assert abs(A - A.T).max() < 1e-12 * abs(A).max()
        if A.shape[0] < 1500:
            assert np.linalg.eigvalsh(A.toarray())[0] > 0
        assert S.num_dofs == S.sub_num_dofs + S.sup_num_dofs - S.num_interface_dofs
        assert S.norm_weight.sum() == S.num_dofs
    us = W.initial_function(4)
    f = W.new_vector(); W.stiffness_matrix(f, us)
    for ptype in (1, 0):
        W.preconditioner_type = ptype
        u = W.new_vector()
        W.flexible_conjugate_gradient(u, f, Sd)
        err = np.sqrt(sum(((a - b) ** 2).sum() for a, b in zip(u, us)) / sum((b ** 2).sum() for b in us))
        assert W.num_iterations <= 8 and W.history[-1] / W.history[0] < 1e-7 and err < 1e-5
    Wn = domain.DomainWorld(str(tmp_path), N, nr)
    Wn.use_preconditioner = False
    u = Wn.new_vector(); Wn.flexible_conjugate_gradient(u, f)
    assert Wn.num_iterations > 4 * W.num_iterations      # the preconditioner does its job


def test_regions_follow_the_ladder(tmp_path):
    W, Sd = _setup(str(tmp_path), 2, 16, 7, 3, 4, 0.0)    # ladder 7, 4, 1: rings at 7, 4, 1 then extended at 1
    S = Sd.ranks[0]
    own = W.ranks[0].num_local_elements
    deg = S.elem_degree
    assert list(deg[:own]) == [7] * own
    assert sorted(set(deg[own:S.num_subdomain_elems].tolist())) == [1, 4, 7]
    assert np.all(np.diff(deg[:S.num_subdomain_elems]) <= 0)          # non-increasing: runs of equal degree
    assert np.all(deg[S.num_subdomain_elems:] == 1)
    assert S.num_superdomain_elems + S.num_subdomain_elems == 16 * 16
    # a 4-rank block of 8x8 own elements + 3 rings (clipped by the domain boundary) = 11x11
    assert S.num_subdomain_elems == 11 * 11 and S.num_subdomain_extended_elems == 12 * 12


@pytest.mark.parametrize("dim,nel,N,r,nr", [(2, 8, 7, 3, 1), (3, 4, 4, 3, 1), (3, 4, 3, 2, 2)])
def test_vcycle_contracts(tmp_path, dim, nel, N, r, nr):
    """pin 5 of SURVEY 8c (the reference's own `#if 0` self-test, subdomain.tpp:3707-3855): one V-cycle of the low-order
    hierarchy reduces the residual of A_fem x = b, and the stationary iteration x += V(b - A x) converges"""
    W, Sd = _setup(str(tmp_path), dim, nel, N, r, nr, 0.02)
    rng = np.random.default_rng(11)
    for S in Sd.ranks:
        A = S.A_fem
        b = rng.standard_normal(A.shape[0])
        x = np.zeros_like(b)
        res = [np.linalg.norm(b)]
        for _ in range(6):
            x = x + S.amg.vcycle(b - A @ x, 1)
            res.append(np.linalg.norm(b - A @ x))
        rates = np.array(res[1:]) / np.array(res[:-1])
        assert rates.max() < 0.6, rates
        assert res[-1] / res[0] < 1e-3
