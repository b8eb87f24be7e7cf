"""
ORACLE (test infrastructure, NOT product code).

CPU restatement of the reference's `Domain<double>` (domain.hpp / domain.tpp) and `CSR_Matrix<double>`
(csr_matrix.tpp): mesh reader, process-boundary detection, gather-scatter operators Q / Q^T,
direct-stiffness summation, the matrix-free SEM Laplacian, the manufactured right-hand side, and the
two outer Krylov drivers (flexible CG = the north-star path, flexible GMRES(20)).

MPI ranks are simulated inside one process: a `DomainWorld` holds one `DomainRank` per rank and
every algorithm is written SPMD -- "for every rank do the local step; then the collective" -- so
the arithmetic each rank performs, and the order it performs it in, is what the reference's rank
performs.  Collectives: MPI_Allreduce(SUM) = sum in rank order; gslib gs_add = for every id held
by several ranks, every holder receives the sum (taken in rank order) of all holders' values.

Device kernels are the C restatements in oracle/kernels.c (pinned bit-for-bit against the
reference's own OKL compiled for CPU, tests/test_oracle_pin.py), called through ctypes on numpy
arrays; the 128-wide block partials are summed serially on the "host" exactly like
domain.tpp:924-926, 958-964, 989-992.

File:line citations are into /root/reference.
"""
import ctypes as C
import math
import os
import numpy as np

from . import capi as _c

BLOCK_SIZE = 128  # config.hpp:38-40


def _L():
    L = _c.lib()
    L.o_serial_sum.restype = C.c_double
    return L


P = _c.ptr


# ------------------------------------------------------------------------------------------
# CSR_Matrix  (csr_matrix.tpp)
# ------------------------------------------------------------------------------------------
class CSRMatrix:
    """COO -> CSR with the reference's rules: entries with |v| <= 1e-12 are dropped AT INSERTION
    (csr_matrix.tpp:61-64, 79-80); sort by (row, col); duplicates are summed (93-180)."""

    sparse_tolerance = 1.0e-12

    def __init__(self, num_rows=0, num_cols=0):
        self.initialize(num_rows, num_cols)

    def initialize(self, num_rows, num_cols):
        self.num_rows, self.num_cols, self.num_nnz = int(num_rows), int(num_cols), 0
        self._r, self._c, self._v = [], [], []
        self.ptr = self.col = self.val = None

    def add_entry(self, row, col, val):
        if row < 0 or row >= self.num_rows or col < 0 or col >= self.num_cols:
            raise IndexError("Entry at (%d, %d) is outside the matrix of size (%d, %d)" % (row, col, self.num_rows, self.num_cols))
        if abs(val) > self.sparse_tolerance:
            self._r.append(row); self._c.append(col); self._v.append(val)

    def add_entries(self, rows, cols, vals):
        """vectorised add_entry (same drop rule, same insertion order)."""
        rows = np.asarray(rows, dtype=np.int64); cols = np.asarray(cols, dtype=np.int64)
        vals = np.asarray(vals, dtype=np.float64)
        if rows.size == 0:
            return
        if rows.min() < 0 or rows.max() >= self.num_rows or cols.min() < 0 or cols.max() >= self.num_cols:
            raise IndexError("entry outside the matrix")
        keep = np.abs(vals) > self.sparse_tolerance
        self._r.append(rows[keep]); self._c.append(cols[keep]); self._v.append(vals[keep])

    def assemble(self):
        def cat(lst, dt):
            parts = [np.atleast_1d(np.asarray(a, dtype=dt)) for a in lst]
            return np.concatenate(parts) if parts else np.zeros(0, dtype=dt)
        r = cat(self._r, np.int64); c = cat(self._c, np.int64); v = cat(self._v, np.float64)
        self._r, self._c, self._v = [], [], []
        if self.num_rows == 0 or self.num_cols == 0 or r.size == 0:
            return  # csr_matrix.tpp:96: early return, nothing allocated
        order = np.lexsort((c, r))  # stable: duplicates keep insertion order
        r, c, v = r[order], c[order], v[order]
        new = np.ones(r.size, dtype=bool)
        new[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
        starts = np.flatnonzero(new)
        self.num_nnz = int(starts.size)
        val = v[starts].copy()
        if starts.size != r.size:  # sum duplicates left to right in sorted order (tpp:148-165)
            seg = np.cumsum(new) - 1
            dup = np.flatnonzero(~new)
            for k in dup:
                val[seg[k]] += v[k]
        self.col = c[starts].astype(np.int32)
        self.val = val
        ptr = np.zeros(self.num_rows + 1, dtype=np.int64)
        np.add.at(ptr, r[starts] + 1, 1)
        self.ptr = np.cumsum(ptr).astype(np.int32)

    def transpose(self):
        At = CSRMatrix(self.num_cols, self.num_rows)
        if self.num_rows == 0 or self.num_cols == 0 or self.ptr is None:
            return At
        rows = np.repeat(np.arange(self.num_rows), np.diff(self.ptr))
        At.add_entries(self.col, rows, self.val)  # csr_matrix.tpp:243-251 (drop rule applies again)
        At.assemble()
        return At

    def diagonal(self):
        d = np.zeros(self.num_rows)
        for i in range(self.num_rows):
            for j in range(self.ptr[i], self.ptr[i + 1]):
                if self.col[j] == i:
                    d[i] = self.val[j]
                    break
        return d

    # the three SpMV launchers (csr_matrix.tpp:301-341) -> csr_matrix.okl
    def multiply(self, Au, u):
        if self.num_rows == 0 or self.num_cols == 0:
            return
        if self.ptr is None:
            raise RuntimeError("multiply on an unassembled CSR matrix (reference would dereference null)")
        _L().o_csr_multiply(P(Au), P(self.ptr), P(self.col), P(self.val), P(u), C.c_int(self.num_rows))

    def multiply_weight(self, Au, u, weight):
        if self.num_rows == 0 or self.num_cols == 0:
            return
        _L().o_csr_multiply_weight(P(Au), P(self.ptr), P(self.col), P(self.val), P(u), P(weight), C.c_int(self.num_rows))

    def to_scipy(self):
        import scipy.sparse as sp
        if self.ptr is None:
            return sp.csr_matrix((self.num_rows, self.num_cols))
        return sp.csr_matrix((self.val, self.col, self.ptr), shape=(self.num_rows, self.num_cols))


# ------------------------------------------------------------------------------------------
# Element record (element.hpp:18-55) -- arrays kept per rank instead of per element
# ------------------------------------------------------------------------------------------
class DomainRank:
    """What one MPI rank's Domain<double> holds after initialize() (domain.tpp:31-371)."""

    def __init__(self, directory, poly_degree, proc_id):
        self.directory, self.poly_degree, self.proc_id = directory, poly_degree, proc_id
        N = poly_degree
        d = os.path.join(directory, "lx1_%d" % (N + 1))
        def fn(name):
            return os.path.join(d, "%s_%d.%d.dat" % (name, proc_id, N))
        with open(fn("size")) as f:  # domain.tpp:45-47
            self.dim, n_x, n_y, n_z, self.num_local_elements = [int(t) for t in f.read().split()[:5]]
        dim, E = self.dim, self.num_local_elements
        self.n = N + 1
        self.num_elem_points = (N + 1) ** dim
        self.num_local_points = E * self.num_elem_points
        npts = self.num_local_points
        self.x = np.fromfile(fn("x"), dtype=np.float64, count=npts)
        self.y = np.fromfile(fn("y"), dtype=np.float64, count=npts) if dim >= 2 else np.zeros(npts)
        self.z = np.fromfile(fn("z"), dtype=np.float64, count=npts) if dim >= 3 else np.zeros(npts)
        self.glo_num = np.fromfile(fn("glo_num"), dtype=np.int64, count=npts)
        node_degree = np.fromfile(fn("node_degree"), dtype=np.int32, count=npts)
        self.dirichlet_mask = np.fromfile(fn("p_mask"), dtype=np.float64, count=npts)
        self.geom_fact = [np.fromfile(fn("g_%d" % (g + 1)), dtype=np.float64, count=npts) for g in range(6)]
        for a in [self.x, self.glo_num, node_degree, self.dirichlet_mask] + self.geom_fact:
            if a.size != npts:
                raise IOError("ERROR: There was a problem reading Nek5000 data")

        # process-boundary nodes first, then first-touch order (domain.tpp:236-281)
        uniq, first_idx, inv, counts = np.unique(self.glo_num, return_index=True, return_inverse=True, return_counts=True)
        flagged = counts[inv] != node_degree                       # tpp:257, evaluated per point
        first_flag = np.full(uniq.size, npts, dtype=np.int64)
        np.minimum.at(first_flag, inv[flagged], np.flatnonzero(flagged))
        is_b = first_flag < npts
        b_nodes = np.flatnonzero(is_b)
        b_nodes = b_nodes[np.argsort(first_flag[b_nodes], kind="stable")]
        i_nodes = np.flatnonzero(~is_b)
        i_nodes = i_nodes[np.argsort(first_idx[i_nodes], kind="stable")]
        self.num_bdary_nodes = int(b_nodes.size)
        self.num_local_nodes = int(uniq.size)
        local_idx_of_uniq = np.empty(uniq.size, dtype=np.int64)
        local_idx_of_uniq[b_nodes] = np.arange(b_nodes.size)
        local_idx_of_uniq[i_nodes] = b_nodes.size + np.arange(i_nodes.size)
        self.local_node_idx = local_idx_of_uniq[inv].astype(np.int32)   # point -> local node
        self.boundary_nodes = uniq[b_nodes].astype(np.int64)            # gs ids (tpp:261, 284)

        # Q (points x nodes, one 1.0 per row) and Qt (tpp:286-294)
        self.Q = CSRMatrix(npts, self.num_local_nodes)
        self.Q.add_entries(np.arange(npts), self.local_node_idx, np.ones(npts))
        self.Q.assemble()
        self.Qt = self.Q.transpose()

        # D_hat (tpp:304-316)
        z, w = _c.zwgll(self.n)
        self.D_hat = np.ascontiguousarray(_c.dgll(z, self.n).ravel())

        self.num_blocks = (npts + BLOCK_SIZE - 1) // BLOCK_SIZE
        self.work = [np.zeros(max(npts, 2 * self.num_blocks)) for _ in range(max(dim, 2))]
        self.gdu = [np.zeros(npts) for _ in range(dim)]
        self.assembled_weight = None


class DomainWorld:
    """All ranks' Domain<double> objects plus the collectives between them."""

    num_vectors = 20          # domain.hpp:113
    max_iterations = 500      # domain.hpp:114
    preconditioner_type = 1   # domain.hpp:115
    use_preconditioner = True
    tolerance = 1.0e-07       # domain.hpp:118 (double)

    def __init__(self, directory, poly_degree, num_procs=1):
        self.num_procs = num_procs
        self.poly_degree = poly_degree
        self.ranks = [DomainRank(directory, poly_degree, p) for p in range(num_procs)]
        self.dim = self.ranks[0].dim
        self.num_total_elements = sum(r.num_local_elements for r in self.ranks)  # tpp:49-50
        self.log = []
        self.history = []
        self.num_iterations = 0
        # gslib handle over boundary ids (tpp:283-284): id -> [(rank, slot)]
        table = {}
        for r in self.ranks:
            for s, gid in enumerate(r.boundary_nodes.tolist()):
                table.setdefault(gid, []).append((r.proc_id, s))
        self.gs_groups = [v for v in table.values() if len(v) > 1]
        # assembled_weight = 1 / (Qt 1 (+) gs_add)  (tpp:296-302)
        L = _L()
        for r in self.ranks:
            one = np.ones(r.num_local_points)
            r.assembled_weight = np.zeros(r.num_local_nodes)
            r.Qt.multiply(r.assembled_weight, one)
        self._gs_add([r.assembled_weight for r in self.ranks])
        for r in self.ranks:
            L.o_invert_vector_elements(P(r.assembled_weight), C.c_int(r.num_local_nodes))

    # -- collectives ------------------------------------------------------------------------
    def _gs_add(self, node_vecs):
        """gslib_gs(..., gs_add, ...) on the first num_bdary_nodes entries of every rank (tpp:590-594)."""
        for grp in self.gs_groups:
            s = 0.0
            for (p, slot) in grp:
                s += node_vecs[p][slot]
            for (p, slot) in grp:
                node_vecs[p][slot] = s

    @staticmethod
    def _allreduce(vals):
        s = 0.0
        for v in vals:
            s += v
        return s

    def rstdout(self, line):
        self.log.append(line)

    def new_vector(self):
        return [np.zeros(r.num_local_points) for r in self.ranks]

    # -- member functions -------------------------------------------------------------------
    def initial_function(self, function_id=4):
        """domain.tpp:527-580.  function_id 4 = glibc rand() with the default seed, per rank."""
        u = self.new_vector()
        L = _L()
        for r, ur in zip(self.ranks, u):
            if function_id == 4:
                L.o_rand_fill(P(ur), C.c_int(r.num_local_points), C.c_uint(1))
            elif function_id == 0:
                if self.dim == 2:
                    ur[:] = np.sin(math.pi * r.x) * np.sin(math.pi * r.y)
                else:
                    ur[:] = np.sin(math.pi * r.x) * np.sin(math.pi * r.y) * np.sin(math.pi * r.z)
            else:
                raise NotImplementedError("function_id %d" % function_id)
        self.direct_stiffness_summation(u, u, True, True)
        return u

    def direct_stiffness_summation(self, QQtu, u, apply_dirichlet_mask=True, apply_assembled_weight=False):
        """domain.tpp:582-600."""
        for r, ur in zip(self.ranks, u):
            if apply_assembled_weight:
                r.Qt.multiply_weight(r.work[0], ur, r.assembled_weight)
            else:
                r.Qt.multiply(r.work[0], ur)
        self._gs_add([r.work[0] for r in self.ranks])
        for r, out in zip(self.ranks, QQtu):
            if apply_dirichlet_mask:
                r.Q.multiply_weight(out, r.work[0], r.dirichlet_mask)
            else:
                r.Q.multiply(out, r.work[0])

    def stiffness_matrix(self, Au, u, apply_dssum=False):
        """domain.tpp:602-609 -> domain.okl:5-98."""
        L = _L()
        for r, Aur, ur in zip(self.ranks, Au, u):
            gdu = _c.ptr_table(r.gdu)
            G = _c.ptr_table(r.geom_fact)
            L.o_stiffness_matrix_1(gdu, P(ur), P(r.D_hat), G, C.c_int(r.num_local_points), C.c_int(r.poly_degree), C.c_int(r.dim))
            L.o_stiffness_matrix_2(P(Aur), gdu, P(r.D_hat), C.c_int(r.num_local_points), C.c_int(r.poly_degree), C.c_int(r.dim))
        if apply_dssum:
            self.direct_stiffness_summation(Au, Au, True, False)

    # -- reductions (domain.tpp:916-996) ----------------------------------------------------
    def residual_norm(self, r_vec):
        L = _L()
        tmp = [rk.work[1][:rk.num_local_points] for rk in self.ranks]
        self.direct_stiffness_summation(tmp, r_vec)
        part = []
        for rk, rr, qq in zip(self.ranks, r_vec, tmp):
            blk = np.zeros(rk.num_blocks)
            L.o_residual_norm(P(blk), P(rr), P(qq), P(rk.dirichlet_mask), C.c_int(rk.num_local_points), C.c_int(rk.num_blocks))
            part.append(L.o_serial_sum(P(blk), C.c_int(rk.num_blocks)))
        return math.sqrt(self._allreduce(part))

    def assembled_inner_product(self, u, v):
        L = _L()
        tmp = [rk.work[1][:rk.num_local_points] for rk in self.ranks]
        self.direct_stiffness_summation(tmp, v)
        part = []
        for rk, uu, qq in zip(self.ranks, u, tmp):
            blk = np.zeros(rk.num_blocks)
            L.o_inner_product_mask(P(blk), P(uu), P(qq), P(rk.dirichlet_mask), C.c_int(rk.num_local_points), C.c_int(rk.num_blocks))
            part.append(L.o_serial_sum(P(blk), C.c_int(rk.num_blocks)))
        return self._allreduce(part)

    def projection_inner_products(self, z, r_, p, q):
        L = _L()
        g, t = [], []
        for rk, zz, rr, pp, qq in zip(self.ranks, z, r_, p, q):
            nb = rk.num_blocks
            blk = np.zeros(2 * nb)
            L.o_projection_inner_products(P(blk), P(zz), P(rr), P(pp), P(qq), C.c_int(rk.num_local_points), C.c_int(nb))
            g.append(L.o_serial_sum(P(blk), C.c_int(nb)))
            t.append(L.o_serial_sum(P(blk[nb:]), C.c_int(nb)))
        return self._allreduce(g), self._allreduce(t)

    def inner_product_flexible(self, r_k, r_kp1, z):
        L = _L()
        part = []
        for rk, a, b, zz in zip(self.ranks, r_k, r_kp1, z):
            blk = np.zeros(rk.num_blocks)
            L.o_inner_product_flexible(P(blk), P(a), P(b), P(zz), C.c_int(rk.num_local_points), C.c_int(rk.num_blocks))
            part.append(L.o_serial_sum(P(blk), C.c_int(rk.num_blocks)))
        return self._allreduce(part)

    # -- element-wise updates ----------------------------------------------------------------
    def _each(self, fn, *vecs):
        for i, rk in enumerate(self.ranks):
            fn(rk, *[v[i] for v in vecs])

    def _precondition(self, z, r_vec, subdomain):
        """domain.tpp:637-651 / 697-711."""
        if self.use_preconditioner and subdomain is not None:
            if self.preconditioner_type == 0:
                subdomain.flexible_conjugate_gradient(z, r_vec)
            else:
                subdomain.generalized_minimum_residual(z, r_vec)
            self.direct_stiffness_summation(z, z, True, True)   # "subdomain stitching"
        else:
            self.direct_stiffness_summation(z, r_vec)

    # -- outer solvers ----------------------------------------------------------------------
    def flexible_conjugate_gradient(self, u, f, subdomain=None, use_relative=True, max_iterations=None):
        """domain.tpp:611-725 -- the PCG hot loop."""
        L = _L()
        if max_iterations is None:
            max_iterations = self.max_iterations
        r_k, r_kp1, q_k, z_k, p_k = (self.new_vector() for _ in range(5))
        u_k = u
        self._each(lambda rk, a, b, c: L.o_initialize_arrays(P(a), P(b), P(c), C.c_int(rk.num_local_points)), u_k, r_k, f)
        r_0_norm = self.residual_norm(r_k)
        self.history = [r_0_norm]
        self.rstdout("Iter %2d: | residual_norm = %24.16g | relative_residual_norm = %24.16g | " % (0, r_0_norm, 1.0))
        self._precondition(z_k, r_k, subdomain)
        for a, b in zip(p_k, z_k):
            a[:] = b
        self.num_iterations = 0
        for it in range(max_iterations):
            self.stiffness_matrix(q_k, p_k)
            gamma_k, theta_k = self.projection_inner_products(z_k, r_k, p_k, q_k)
            alpha_k = gamma_k / theta_k
            self._each(lambda rk, a, b, c, d, e: L.o_solution_and_residual_update(P(a), P(b), P(c), P(d), P(e), C.c_double(alpha_k), C.c_int(rk.num_local_points)),
                       u_k, r_kp1, r_k, p_k, q_k)
            r_norm = self.residual_norm(r_kp1)
            self.history.append(r_norm)
            self.rstdout("Iter %2d: | residual_norm = %24.16g | relative_residual_norm = %24.16g | " % (it + 1, r_norm, r_norm / r_0_norm))
            if use_relative:
                if r_norm / r_0_norm < self.tolerance:
                    break
            else:
                if r_norm < self.tolerance:
                    break
            if math.isnan(r_norm):
                break
            self._precondition(z_k, r_kp1, subdomain)
            theta_k = self.inner_product_flexible(r_k, r_kp1, z_k)
            beta_k = theta_k / gamma_k
            self._each(lambda rk, a, b, c, d: L.o_residual_and_search_update(P(a), P(b), P(c), P(d), C.c_double(beta_k), C.c_int(rk.num_local_points)),
                       p_k, r_k, z_k, r_kp1)
            self.num_iterations += 1
        return u_k

    def generalized_minimum_residual(self, u, f, subdomain=None, use_relative=True, max_iterations=None):
        """domain.tpp:727-914 -- flexible GMRES(20), one-pass classical Gram-Schmidt, Givens."""
        L = _L()
        if max_iterations is None:
            max_iterations = self.max_iterations
        nv = self.num_vectors
        r_k, q_k = self.new_vector(), self.new_vector()
        V = [self.new_vector() for _ in range(nv + 1)]
        Z = [self.new_vector() for _ in range(nv)]
        H = [[0.0] * nv for _ in range(nv)]
        c_g = [0.0] * nv; s_g = [0.0] * nv; gamma = [0.0] * (nv + 1)
        u_k = u
        self._each(lambda rk, a, b, c: L.o_initialize_arrays(P(a), P(b), P(c), C.c_int(rk.num_local_points)), u_k, r_k, f)
        r_0_norm = self.residual_norm(r_k)
        self.history = [r_0_norm]
        self.rstdout("Iter %2d: | residual_norm = %24.16g | relative_residual_norm = %24.16g | " % (0, r_0_norm, 1.0))

        def axpby(out, a, x, b, y):
            self._each(lambda rk, o, xx, yy: L.o_vector_vector_addition(P(o), C.c_double(a), P(xx), C.c_double(b), P(yy), C.c_int(rk.num_local_points)), out, x, y)

        def scal(out, a, x):
            self._each(lambda rk, o, xx: L.o_vector_scaling(P(o), C.c_double(a), P(xx), C.c_int(rk.num_local_points)), out, x)

        converged = False
        it = 0
        while it < max_iterations:
            if it > 0:
                self.stiffness_matrix(r_k, u_k)
                axpby(r_k, 1.0, f, -1.0, r_k)
                gamma[0] = self.residual_norm(r_k)
            else:
                gamma[0] = r_0_norm
            scal(V[0], 1.0 / gamma[0], r_k)
            j = 0
            while j < nv:
                self._precondition(Z[j], V[j], subdomain)
                self.stiffness_matrix(q_k, Z[j])
                for i in range(j + 1):
                    H[i][j] = self.assembled_inner_product(q_k, V[i])
                for i in range(j + 1):
                    axpby(q_k, 1.0, q_k, -H[i][j], V[i])
                for i in range(j):
                    h_ij = H[i][j]
                    H[i][j] = c_g[i] * h_ij + s_g[i] * H[i + 1][j]
                    H[i + 1][j] = -s_g[i] * h_ij + c_g[i] * H[i + 1][j]
                alpha_j = self.residual_norm(q_k)
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
                self.history.append(r_norm)
                self.rstdout("Iter %2d: | residual_norm = %24.16g | relative_residual_norm = %24.16g | " % (it + 1, r_norm, r_norm / r_0_norm))
                if (r_norm / r_0_norm < self.tolerance) if use_relative else (r_norm < self.tolerance):
                    converged = True
                    break
                if it >= max_iterations:
                    converged = True
                    break
                if math.isnan(r_norm):
                    converged = True
                    break
                scal(V[j + 1], 1.0 / alpha_j, q_k)
                it += 1
                j += 1
            if j == nv:
                j -= 1
            for k in range(j, -1, -1):
                gamma_k = gamma[k]
                for i in range(j, k, -1):
                    gamma_k -= H[k][i] * c_g[i]
                c_g[k] = gamma_k / H[k][k]
            for i in range(j + 1):
                axpby(u_k, 1.0, u_k, c_g[i], Z[i])
            if converged:
                break
        self.num_iterations = it
        return u_k

    # -- helpers for tests ------------------------------------------------------------------
    def num_global_nodes(self):
        return int(len(set(np.concatenate([r.glo_num for r in self.ranks]).tolist())))
