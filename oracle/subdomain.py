"""
ORACLE (test infrastructure, NOT product code).

CPU restatement of the reference's PR-FDD preconditioner `Subdomain<double>` (subdomain.hpp / subdomain.tpp):
the constructor (subdomain.tpp:86-3933) and the run-time methods (3942-4646): tree_operator, the
composite operator, the rank-local inner Krylov solves (flexible GMRES(4) default, flexible CG), the
low-order (P1 simplex FEM) preconditioner with its Chebyshev-smoothed AMG V-cycle.

Ranks are simulated in one process (see oracle/domain.py).  "Pulling" a region element's data from its
owner through gslib (subdomain.tpp:644-805) is a direct lookup in the owner's Domain at that degree.
HYPRE is replaced by oracle/amg.py (see its header: SETUP parity unpinned).

Status of the restatement:
  * num_procs == 1: complete (the region is the rank's own elements at degree N, the superdomain is empty --
    a path the reference never exercised, SURVEY.md 8e -- the zero-sized superdomain operators are
    special-cased exactly where the reference would hand HYPRE empty matrices, subdomain.tpp:2426-2431).
  * num_procs > 1: see oracle/subdomain_multi.py (regions with overlap rings at ladder degrees,
    non-conforming Q, superdomain composite grid).

File:line citations are into /root/reference.
"""
import ctypes as C
import math
import numpy as np
import scipy.sparse as sp

from . import capi as _c
from . import amg as _amg
from .domain import CSRMatrix, DomainWorld, BLOCK_SIZE, _L

P = _c.ptr
EPSILON = 1.0e-12     # Subdomain::epsilon (subdomain.hpp:233)


def ladder(N, r):
    out = [N]
    while out[-1] > 1:
        out.append(max(out[-1] - r, 1))          # subdomain.tpp:98-108
    return out


def ranking(data):
    """dense ranking of subdomain.tpp:881-918: equal values get equal ranks; the value 0 gets rank 0."""
    data = np.asarray(data, dtype=np.float64)
    if data.size == 0:
        return data.copy()
    uniq, inv = np.unique(data, return_inverse=True)
    base = 0.0 if uniq[0] == 0.0 else 1.0
    return (inv + base).astype(np.float64)


# the 2 triangles / 6 tetrahedra of a GLL cell (subdomain.tpp:2853-2883)
TRIS = [[(0, 0, 0), (1, 0, 0), (1, 1, 0)], [(1, 1, 0), (0, 1, 0), (0, 0, 0)]]
TETS = [[(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 0, 1)],
        [(1, 0, 0), (0, 1, 0), (1, 1, 0), (1, 0, 1)],
        [(0, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 1)],
        [(1, 0, 1), (1, 1, 0), (1, 1, 1), (0, 1, 0)],
        [(0, 0, 1), (1, 0, 1), (0, 1, 1), (0, 1, 0)],
        [(1, 0, 1), (1, 1, 1), (0, 1, 1), (0, 1, 0)]]
D_FEM_2D = [np.array([-1.0, 1.0, 0.0] * 3).reshape(3, 3), np.array([-1.0, 0.0, 1.0] * 3).reshape(3, 3)]
D_FEM_3D = [np.array([1.0, 0.0, 0.0, -1.0] * 4).reshape(4, 4), np.array([0.0, 1.0, 0.0, -1.0] * 4).reshape(4, 4),
            np.array([0.0, 0.0, 1.0, -1.0] * 4).reshape(4, 4)]


def low_order_element_matrix(dim, N, x, y, z):
    """A_e (num_points x num_points, dense) of one element of degree N > 1: P1 stiffness on the simplices of
    every GLL cell, accumulated in the reference's loop order (subdomain.tpp:2934-3038), entries with
    |v| <= epsilon of a single simplex matrix dropped before accumulation (3025)."""
    n = N + 1
    npts = n ** dim
    simplices = TRIS if dim == 2 else TETS
    nv = dim + 1
    Dfem = D_FEM_2D if dim == 2 else D_FEM_3D
    weight = 6.0 if dim == 2 else 24.0
    S = N
    sz_range = range(S) if dim == 3 else range(1)
    # local indices of all simplices in loop order (s_z, s_y, s_x, simplex, vertex)
    loc = []
    for s_z in sz_range:
        for s_y in range(S):
            for s_x in range(S):
                for smp in simplices:
                    if dim == 2:
                        loc.append([(s_x + i) + (s_y + j) * n for (i, j, k) in smp])
                    else:
                        loc.append([(s_x + i) + (s_y + j) * n + (s_z + k) * n * n for (i, j, k) in smp])
    loc = np.array(loc, dtype=np.int64)                     # (T, nv)
    xs, ys = x[loc], y[loc]
    zs = z[loc] if dim == 3 else None
    T = loc.shape[0]
    H = np.zeros((T, dim, dim))
    if dim == 2:
        H[:, 0, 0] = xs[:, 1] - xs[:, 0]; H[:, 0, 1] = xs[:, 2] - xs[:, 0]
        H[:, 1, 0] = ys[:, 1] - ys[:, 0]; H[:, 1, 1] = ys[:, 2] - ys[:, 0]
        det = H[:, 0, 0] * H[:, 1, 1] - H[:, 0, 1] * H[:, 1, 0]
        inv = np.zeros_like(H)
        inv[:, 0, 0] = (1.0 / det) * H[:, 1, 1]
        inv[:, 0, 1] = -(1.0 / det) * H[:, 0, 1]
        inv[:, 1, 0] = -(1.0 / det) * H[:, 1, 0]
        inv[:, 1, 1] = (1.0 / det) * H[:, 0, 0]
    else:
        for c, arr in enumerate((xs, ys, zs)):
            H[:, c, 0] = arr[:, 0] - arr[:, 3]; H[:, c, 1] = arr[:, 1] - arr[:, 3]; H[:, c, 2] = arr[:, 2] - arr[:, 3]
        A = H.reshape(T, 9)
        det = A[:, 0] * (A[:, 4] * A[:, 8] - A[:, 5] * A[:, 7]) - A[:, 1] * (A[:, 3] * A[:, 8] - A[:, 5] * A[:, 6]) + A[:, 2] * (A[:, 3] * A[:, 7] - A[:, 4] * A[:, 6])
        inv = np.zeros((T, 9))
        r = 1.0 / det
        inv[:, 0] = r * (A[:, 4] * A[:, 8] - A[:, 7] * A[:, 5])
        inv[:, 1] = r * (A[:, 2] * A[:, 7] - A[:, 8] * A[:, 1])
        inv[:, 2] = r * (A[:, 1] * A[:, 5] - A[:, 4] * A[:, 2])
        inv[:, 3] = r * (A[:, 5] * A[:, 6] - A[:, 8] * A[:, 3])
        inv[:, 4] = r * (A[:, 0] * A[:, 8] - A[:, 6] * A[:, 2])
        inv[:, 5] = r * (A[:, 2] * A[:, 3] - A[:, 5] * A[:, 0])
        inv[:, 6] = r * (A[:, 3] * A[:, 7] - A[:, 6] * A[:, 4])
        inv[:, 7] = r * (A[:, 1] * A[:, 6] - A[:, 7] * A[:, 0])
        inv[:, 8] = r * (A[:, 0] * A[:, 4] - A[:, 3] * A[:, 1])
        inv = inv.reshape(T, 3, 3)
    # G[m][n] = sum_k (det/weight) * inv[m][k] * inv[n][k]     (tpp:2985-2999)
    Gmn = np.zeros((T, dim, dim))
    for m in range(dim):
        for nn in range(dim):
            g = np.zeros(T)
            for k in range(dim):
                g = g + (det / weight) * inv[:, m, k] * inv[:, nn, k]
            Gmn[:, m, nn] = g
    # A_t[i][j] = sum_{m,n} sum_q D[m][q][i] * (G[m][n] * D[n][q][j])   (tpp:3001-3019); the quadrature
    # points q all carry the same G, the sum over q is kept in the reference's order
    At = np.zeros((T, nv, nv))
    for m in range(dim):
        for nn in range(dim):
            for q in range(nv):
                for i in range(nv):
                    dmi = Dfem[m][q, i]
                    if dmi == 0.0:
                        continue
                    for j in range(nv):
                        dnj = Dfem[nn][q, j]
                        if dnj == 0.0:
                            continue
                        At[:, i, j] = At[:, i, j] + dmi * (Gmn[:, m, nn] * dnj)
    keep = np.abs(At) > EPSILON
    rows = np.broadcast_to(loc[:, :, None], (T, nv, nv))[keep]
    cols = np.broadcast_to(loc[:, None, :], (T, nv, nv))[keep]
    Ae = np.zeros((npts, npts))
    np.add.at(Ae, (rows, cols), At[keep])                   # sequential accumulation in loop order
    return Ae


def q1_element_matrix(dim, D2, G):
    """N = 1 elements use the Q1 SEM matrix D^T G D (subdomain.tpp:1715-1826, 3040-3124). G: (6, nverts)."""
    nv = 2 ** dim
    Dm = [np.zeros((nv, nv)) for _ in range(dim)]
    d = D2.reshape(2, 2)
    if dim == 2:
        for k in range(2):
            for i in range(2):
                for j in range(2):
                    Dm[0][(i + k * 2), (j + k * 2)] = d[i, j]
        for i in range(2):
            for j in range(2):
                for k in range(2):
                    Dm[1][(i * 2 + k), (j * 2 + k)] = d[i, j]
        GD1 = G[0][:, None] * Dm[0] + G[2][:, None] * Dm[1]
        GD2 = G[2][:, None] * Dm[0] + G[1][:, None] * Dm[1]
        return Dm[0].T @ GD1 + Dm[1].T @ GD2
    for p in range(2):
        for q in range(2):
            for i in range(2):
                for j in range(2):
                    Dm[0].flat[(i + (p * 2 + q) * 2) * 8 + (j + (p * 2 + q) * 2)] = d[i, j]
                    Dm[1].flat[(i * 8 + j) * 2 + ((p + p * 8) * (2 * 2) + (q + q * 8))] = d[i, j]
                    Dm[2].flat[(i * 8 + j) * (2 * 2) + (p + q * 2) * (1 + 8)] = d[i, j]
    GD1 = G[0][:, None] * Dm[0] + G[3][:, None] * Dm[1] + G[4][:, None] * Dm[2]
    GD2 = G[3][:, None] * Dm[0] + G[1][:, None] * Dm[1] + G[5][:, None] * Dm[2]
    GD3 = G[4][:, None] * Dm[0] + G[5][:, None] * Dm[1] + G[2][:, None] * Dm[2]
    return Dm[0].T @ GD1 + Dm[1].T @ GD2 + Dm[2].T @ GD3


class Region:
    """a rank's subdomain (or superdomain) region: element list + per-point arrays, region order."""
    pass


class SubdomainRank:
    pass


class SubdomainWorld:
    """All ranks' Subdomain<double> objects.  Constructor == subdomain.tpp:86-3933."""

    num_vectors = 4            # subdomain.hpp:229
    max_iterations = 4         # subdomain.hpp:230
    use_preconditioner = True
    tolerance = 1.0e-12        # subdomain.hpp:232
    num_vcycles = 1
    cheby_order = 2
    level_cutoff = 5

    def __init__(self, domain_world, directory, poly_degree, poly_reduction, subdomain_overlap=1, superdomain_overlap=1, **kw):
        for k, v in kw.items():
            setattr(self, k, v)
        self.W = domain_world
        self.num_procs = domain_world.num_procs
        self.dim = domain_world.dim
        self.poly_degree = ladder(poly_degree, poly_reduction)
        self.num_levels = len(self.poly_degree)
        self.domains = {poly_degree: domain_world}
        for N in self.poly_degree[1:]:
            self.domains[N] = DomainWorld(directory, N, self.num_procs)     # poisson.cpp:176-199
        self.subdomain_overlap, self.superdomain_overlap = subdomain_overlap, superdomain_overlap
        # reference operators per level (tpp:129-196)
        self.r_gll, self.D_hat, self.J_cf = [], [], {}
        for N in self.poly_degree:
            z, _ = _c.zwgll(N + 1)
            self.r_gll.append(z)
        for lf in range(self.num_levels - 1):
            for lc in range(lf + 1, self.num_levels):
                nf, nc = self.poly_degree[lf] + 1, self.poly_degree[lc] + 1
                J = np.zeros(nf * nc)
                for i in range(nf):
                    for j in range(1, nc + 1):
                        J[i * nc + (j - 1)] = _c.hgll(j, self.r_gll[lf][i], self.r_gll[lc], nc)
                self.J_cf[(self.poly_degree[lc], self.poly_degree[lf])] = J
        for l, N in enumerate(self.poly_degree):
            self.D_hat.append(np.ascontiguousarray(_c.dgll(self.r_gll[l], N + 1).ravel()))  # snaps the mid node to 0 (PNLEG)
        self.num_iterations = 0
        if self.num_procs == 1:
            self.ranks = [self._build_single_rank()]
        else:
            from . import subdomain_multi
            self.ranks = subdomain_multi.build(self)

    # ------------------------------------------------------------------------------------------
    def _build_single_rank(self):
        dim, N = self.dim, self.poly_degree[0]
        dr = self.W.ranks[0]
        S = SubdomainRank()
        S.proc_id = 0
        npe = (N + 1) ** dim
        E = dr.num_local_elements
        # region = own elements at degree N (tpp:468-474); no rings, no extended, no superdomain
        S.elem_id = np.arange(E, dtype=np.int32)
        S.elem_degree = np.full(E, N, dtype=np.int32)
        S.elem_offset = (np.arange(E) * npe).astype(np.int32)
        S.num_subdomain_elems = S.num_subdomain_extended_elems = E
        S.num_points = E * npe
        S.mask = dr.dirichlet_mask.copy()
        S.geom_fact = [g.copy() for g in dr.geom_fact]
        S.x, S.y, S.z = dr.x.copy(), dr.y.copy(), dr.z.copy()
        # numbering (tpp:920-1176): global offset of level 0 is 0, no non-conforming entities, no interface
        glo = ranking(dr.glo_num.astype(np.float64))
        S.glo_num = glo.astype(np.int64)
        S.dof_num = ranking(S.glo_num.astype(np.float64) * S.mask).astype(np.int64)
        # region Q (tpp:1496-1585): one 1.0 per unmasked point
        S.sub_num_dofs = int(S.dof_num.max())
        S.Q = CSRMatrix(S.num_points, S.sub_num_dofs)
        nz = np.flatnonzero(S.dof_num > 0)
        S.Q.add_entries(nz, S.dof_num[nz] - 1, np.ones(nz.size))
        S.Q.assemble()
        S.Qt = S.Q.transpose()
        S.sub_num_extended_dofs = S.Q.num_cols
        # per-point lookups of the variable-degree operator (tpp:1603-1630)
        S.offset = np.repeat(S.elem_offset, npe).astype(np.int32)
        S.vertex = np.tile(np.arange(npe, dtype=np.int32), E)
        S.level = np.zeros(S.num_points, dtype=np.int32)
        # empty superdomain
        S.sup_num_dofs = S.sup_num_extended_dofs = 0
        S.num_interface_dofs = 0
        S.A_sup = None; S.Pt = None; S.Qt_coarse = None
        S.num_dofs = S.sub_num_dofs + S.sup_num_dofs - S.num_interface_dofs        # tpp:2583
        S.num_values = S.num_points + S.sup_num_extended_dofs                     # tpp:3858
        next_ = S.sub_num_extended_dofs + S.sup_num_extended_dofs
        # interface operators (tpp:2653-2729): identities here
        S.Q_int = CSRMatrix(next_, S.num_dofs); S.Q_int.add_entries(np.arange(next_), np.arange(next_), np.ones(next_)); S.Q_int.assemble()
        S.Qt_int = CSRMatrix(S.num_dofs, next_); S.Qt_int.add_entries(np.arange(S.sub_num_dofs), np.arange(S.sub_num_dofs), np.ones(S.sub_num_dofs)); S.Qt_int.assemble()
        S.QQt_int = CSRMatrix(next_, next_); S.QQt_int.add_entries(np.arange(S.sub_num_dofs), np.arange(S.sub_num_dofs), np.ones(S.sub_num_dofs)); S.QQt_int.assemble()
        # weights (tpp:2731-2747)
        S.norm_weight = np.ones(next_)
        S.norm_weight[S.sub_num_dofs:S.sub_num_extended_dofs] = 0.0
        S.inner_weight = np.zeros(S.num_values)
        S.Q.multiply(S.inner_weight, S.norm_weight)
        S.inner_weight[S.inner_weight > 0.0] = 1.0
        # low-order preconditioner (tpp:2749-3549)
        if self.use_preconditioner:
            S.A_sub_fem = self._assemble_fem_conforming(S)
            S.A_fem = S.A_sub_fem                                                   # Q_int is the identity (tpp:3414-3472)
            S.amg = _amg.Hierarchy(S.A_fem, cheby_order=self.cheby_order, dtype=np.float32 if getattr(self, "amg_precision", "double") == "float" else np.float64)
        self._alloc(S)
        return S

    def _assemble_fem_conforming(self, S):
        """A_sub_fem for a region without non-conforming entities (J_e is a boolean selection, tpp:3287-3297)."""
        dim = self.dim
        rows, cols, vals = [], [], []
        for e in range(S.elem_id.size):
            N = int(S.elem_degree[e]); npe = (N + 1) ** dim
            o = int(S.elem_offset[e])
            sl = slice(o, o + npe)
            if N > 1:
                Ae = low_order_element_matrix(dim, N, S.x[sl], S.y[sl], S.z[sl])
            else:
                G = np.array([g[sl] for g in S.geom_fact])
                Ae = q1_element_matrix(dim, self.D_hat[self.poly_degree.index(1)], G)
                Ae = np.where(np.abs(Ae) > EPSILON, Ae, 0.0)                        # tpp:3077, 3120
            dof = S.dof_num[sl]
            ii, jj = np.nonzero(np.abs(Ae) > EPSILON)                               # tpp:3395
            ok = (dof[ii] > 0) & (dof[jj] > 0)
            rows.append(dof[ii[ok]] - 1); cols.append(dof[jj[ok]] - 1); vals.append(Ae[ii[ok], jj[ok]])
        n = S.sub_num_extended_dofs
        M = CSRMatrix(n, n)
        M.sparse_tolerance = -1.0           # HYPRE_IJMatrixAddToValues drops nothing
        M.add_entries(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))
        M.assemble()
        return M.to_scipy()

    def _alloc(self, S):
        nvl = S.num_values
        S.f = np.zeros(nvl); S.u_k = np.zeros(nvl); S.r_k = np.zeros(nvl); S.r_kp1 = np.zeros(nvl)
        S.q_k = np.zeros(nvl); S.z_k = np.zeros(nvl); S.p_k = np.zeros(nvl)
        S.V = [np.zeros(nvl) for _ in range(self.num_vectors + 1)]
        S.Z = [np.zeros(nvl) for _ in range(self.num_vectors)]
        size = max(nvl, S.sub_num_extended_dofs + S.sup_num_extended_dofs, 1)
        S.work = [np.zeros(size + 8) for _ in range(max(self.dim, 2))]
        S.gdu = [np.zeros(max(S.num_points, 1)) for _ in range(self.dim)]
        S.pd = np.array(self.poly_degree, dtype=np.float64)

    # ------------------------------------------------------------------------------------------
    # run-time methods, one rank at a time (no communication after tree_operator)
    # ------------------------------------------------------------------------------------------
    def stiffness_matrix(self, S, Au, u):
        """subdomain.tpp:3942-3967: A_sup SpMV on the superdomain dofs + variable-degree SEM apply on the region points."""
        L = _L()
        npt = S.num_points
        if S.sup_num_extended_dofs > 0:
            S.A_sup.multiply(Au[npt:], u[npt:])
        gdu = _c.ptr_table(S.gdu)
        L.o_sub_stiffness_matrix_1(gdu, P(u), _c.ptr_table(self.D_hat), P(S.offset), P(S.vertex), P(S.level), _c.ptr_table(S.geom_fact), C.c_int(npt), P(S.pd), C.c_int(self.dim))
        L.o_sub_stiffness_matrix_2(P(Au), gdu, _c.ptr_table(self.D_hat), P(S.offset), P(S.vertex), P(S.level), C.c_int(npt), P(S.pd), C.c_int(self.dim))

    def direct_stiffness_summation(self, S, QQtu, u):
        """subdomain.tpp:3969-3985."""
        npt, ne, ns = S.num_points, S.sub_num_extended_dofs, S.sup_num_extended_dofs
        w0, w1 = S.work[0], S.work[1]
        S.Qt.multiply(w0, u[:npt])
        w0[ne:ne + ns] = u[npt:npt + ns]
        S.QQt_int.multiply(w1, w0)
        S.Q.multiply(QQtu[:npt], w1)
        QQtu[npt:npt + ns] = w1[ne:ne + ns]

    def low_order_preconditioner(self, S, z, r):
        """subdomain.tpp:3987-4159: z = Q Q_int Vcycle(A_fem) Qt_int Qt r."""
        npt, ne, ns = S.num_points, S.sub_num_extended_dofs, S.sup_num_extended_dofs
        w0, w1 = S.work[0], S.work[1]
        S.Qt.multiply(w0, r[:npt])
        w0[ne:ne + ns] = r[npt:npt + ns]
        S.Qt_int.multiply(w1, w0)
        f_fem = w1[:S.num_dofs].copy()
        u_fem = S.amg.vcycle(f_fem, self.num_vcycles)
        w1[:S.num_dofs] = u_fem
        S.Q_int.multiply(w0, w1)
        S.Q.multiply(z[:npt], w0)
        z[npt:npt + ns] = w0[ne:ne + ns]

    def _wdot(self, S, a, b, w, nvals):
        L = _L()
        nb = (nvals + BLOCK_SIZE - 1) // BLOCK_SIZE
        blk = np.zeros(max(nb, 1))
        L.o_sub_weighted_inner_product(P(blk), P(a), P(b), P(w), C.c_int(nvals), C.c_int(nb))
        return L.o_serial_sum(P(blk), C.c_int(nb))

    def residual_norm(self, S, r):
        """subdomain.tpp:4491-4515."""
        npt, ne, ns = S.num_points, S.sub_num_extended_dofs, S.sup_num_extended_dofs
        w1 = S.work[1]
        S.Qt.multiply_weight(w1, r[:npt], S.norm_weight)
        w1[ne:ne + ns] = r[npt:npt + ns]
        return math.sqrt(self._wdot(S, w1, w1, S.norm_weight, ne + ns))

    def assembled_inner_product(self, S, u, v):
        """subdomain.tpp:4277-4307."""
        npt, ne, ns = S.num_points, S.sub_num_extended_dofs, S.sup_num_extended_dofs
        w0, w1 = S.work[0], S.work[1]
        S.Qt.multiply_weight(w0, u[:npt], S.norm_weight)
        w0[ne:ne + ns] = u[npt:npt + ns]
        S.Qt.multiply_weight(w1, v[:npt], S.norm_weight)
        w1[ne:ne + ns] = v[npt:npt + ns]
        return self._wdot(S, w0, w1, S.norm_weight, ne + ns)

    def tree_operator(self, Tu_list, u_list):
        """subdomain.tpp:4566-4646.  Tu[rank] (num_values) from the outer residual u[rank] (own points)."""
        if self.num_procs == 1:
            S = self.ranks[0]
            L = _L()
            npts0 = self.W.ranks[0].num_local_points
            L.o_copy_from_domain_data(P(S.work[0]), P(u_list[0]), C.c_int(npts0))       # tpp:4571
            # the ladder restrictions (tpp:4576-4609) feed only other ranks' regions and the (empty) superdomain
            Tu_list[0][:S.num_points] = S.work[0][:S.num_points]                        # own elements at degree N
            return
        from . import subdomain_multi
        subdomain_multi.tree_operator(self, Tu_list, u_list)

    # -- inner solvers: called by DomainWorld._precondition with per-rank lists ------------------
    def generalized_minimum_residual(self, u_l, f_l):
        """subdomain.tpp:4309-4489 on every rank."""
        L = _L()
        self.tree_operator([S.f for S in self.ranks], f_l)
        for S, ul in zip(self.ranks, u_l):
            nvl = S.num_values
            nv = self.num_vectors
            H = [[0.0] * nv for _ in range(nv)]
            c_g = [0.0] * nv; s_g = [0.0] * nv; gamma = [0.0] * (nv + 1)
            L.o_initialize_arrays(P(S.u_k), P(S.r_k), P(S.f), C.c_int(nvl))
            r_0_norm = self.residual_norm(S, S.r_k)
            converged = False
            it = 0
            while it < self.max_iterations:
                if it > 0:
                    self.stiffness_matrix(S, S.r_k, S.u_k)
                    L.o_vector_vector_addition(P(S.r_k), C.c_double(1.0), P(S.f), C.c_double(-1.0), P(S.r_k), C.c_int(nvl))
                    gamma[0] = self.residual_norm(S, S.r_k)
                else:
                    gamma[0] = r_0_norm
                L.o_vector_scaling(P(S.V[0]), C.c_double(1.0 / gamma[0]), P(S.r_k), C.c_int(nvl))
                j = 0
                while j < nv:
                    it += 1
                    if self.use_preconditioner:
                        self.low_order_preconditioner(S, S.Z[j], S.V[j])
                    else:
                        self.direct_stiffness_summation(S, S.Z[j], S.V[j])
                    self.stiffness_matrix(S, S.q_k, S.Z[j])
                    for i in range(j + 1):
                        H[i][j] = self.assembled_inner_product(S, S.q_k, S.V[i])
                    for i in range(j + 1):
                        L.o_vector_vector_addition(P(S.q_k), C.c_double(1.0), P(S.q_k), C.c_double(-H[i][j]), P(S.V[i]), C.c_int(nvl))
                    for i in range(j):
                        h_ij = H[i][j]
                        H[i][j] = c_g[i] * h_ij + s_g[i] * H[i + 1][j]
                        H[i + 1][j] = -s_g[i] * h_ij + c_g[i] * H[i + 1][j]
                    alpha_j = self.residual_norm(S, S.q_k)
                    if abs(alpha_j) == 0.0:
                        converged = True
                        break
                    beta_j = math.sqrt(H[j][j] * H[j][j] + alpha_j * alpha_j)
                    gamma_j = 1.0 / beta_j
                    c_g[j] = H[j][j] * gamma_j
                    s_g[j] = alpha_j * gamma_j
                    H[j][j] = beta_j
                    gamma[j + 1] = -s_g[j] * gamma[j]
                    gamma[j] = c_g[j] * gamma[j]
                    r_norm = abs(gamma[j + 1])
                    if r_norm < self.tolerance:             # use_relative = false (subdomain.hpp:244, tpp:4440-4447)
                        converged = True
                        break
                    if it >= self.max_iterations:
                        converged = True
                        break
                    L.o_vector_scaling(P(S.V[j + 1]), C.c_double(1.0 / alpha_j), P(S.q_k), C.c_int(nvl))
                    j += 1
                if j == nv:
                    j -= 1
                for k in range(j, -1, -1):
                    gamma_k = gamma[k]
                    for i in range(j, k, -1):
                        gamma_k -= H[k][i] * c_g[i]
                    c_g[k] = gamma_k / H[k][k]
                for i in range(j + 1):
                    L.o_vector_vector_addition(P(S.u_k), C.c_double(1.0), P(S.u_k), C.c_double(c_g[i]), P(S.Z[i]), C.c_int(nvl))
                if converged:
                    break
            npts0 = self.W.ranks[S.proc_id].num_local_points
            L.o_copy_to_domain_data(P(ul), P(S.u_k), C.c_int(npts0))               # tpp:4485
            self.num_iterations += it

    def flexible_conjugate_gradient(self, u_l, f_l):
        """subdomain.tpp:4161-4268 on every rank."""
        L = _L()
        self.tree_operator([S.r_k for S in self.ranks], f_l)
        for S, ul in zip(self.ranks, u_l):
            nvl = S.num_values
            nb = (nvl + BLOCK_SIZE - 1) // BLOCK_SIZE
            L.o_set_to_value(P(S.u_k), C.c_double(0.0), C.c_int(nvl), C.c_int(0))
            r_0_norm = self.residual_norm(S, S.r_k)
            pre = self.low_order_preconditioner if self.use_preconditioner else self.direct_stiffness_summation
            pre(S, S.z_k, S.r_k)
            S.p_k[:] = S.z_k
            it = 0
            while it < self.max_iterations:
                self.stiffness_matrix(S, S.q_k, S.p_k)
                blk = np.zeros(2 * nb)
                L.o_sub_projection_inner_products(P(blk), P(S.z_k), P(S.r_k), P(S.p_k), P(S.q_k), P(S.inner_weight), C.c_int(nvl), C.c_int(nb))
                gamma_k = L.o_serial_sum(P(blk), C.c_int(nb)); theta_k = L.o_serial_sum(P(blk[nb:]), C.c_int(nb))
                alpha_k = gamma_k / theta_k
                L.o_solution_and_residual_update(P(S.u_k), P(S.r_kp1), P(S.r_k), P(S.p_k), P(S.q_k), C.c_double(alpha_k), C.c_int(nvl))
                r_norm = self.residual_norm(S, S.r_kp1)
                it += 1
                if r_norm < self.tolerance:                 # use_relative = false
                    break
                if it == self.max_iterations:
                    break
                pre(S, S.z_k, S.r_kp1)
                blk = np.zeros(nb)
                L.o_sub_search_update_inner_product(P(blk), P(S.r_k), P(S.r_kp1), P(S.z_k), P(S.inner_weight), C.c_int(nvl), C.c_int(nb))
                theta_k = L.o_serial_sum(P(blk), C.c_int(nb))
                beta_k = theta_k / gamma_k
                L.o_residual_and_search_update(P(S.p_k), P(S.r_k), P(S.z_k), P(S.r_kp1), C.c_double(beta_k), C.c_int(nvl))
            self.num_iterations += it
            npts0 = self.W.ranks[S.proc_id].num_local_points
            L.o_copy_to_domain_data(P(ul), P(S.u_k), C.c_int(npts0))               # tpp:4266
            del r_0_norm
