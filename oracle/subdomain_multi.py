"""
ORACLE (test infrastructure, NOT product code).

Multi-rank part of the PR-FDD constructor restatement (subdomain.tpp:198-3549) and of tree_operator
(subdomain.tpp:4566-4646): overlap rings at the ladder degrees, extended elements, the superdomain and its
AMG-composite coarsening, non-conforming region Q, interface maps, low-order FEM with hanging nodes.

Everything the reference gathers on every rank ("TODO: still based on global data stored locally",
subdomain.tpp:198-430, 1632-1660) is computed once here and shared by the simulated ranks.
HYPRE -> oracle/amg.py (setup parity unpinned, see there).  One documented deviation: the reference finds the
C-points of an AMG level as "rows of P with exactly one entry" (subdomain.tpp:1987-1990, 2026-2028); here the C/F
marker of the coarsening is used, which is what that test stands for (a truncated F-row with one entry would
otherwise be mistaken for a C-point).
"""
import ctypes as C
import numpy as np
import scipy.sparse as sp

from . import capi as _c
from . import amg as _amg
from .domain import CSRMatrix, _L

P = _c.ptr
EPSILON = 1.0e-12

EDGE_PAIRS_2D = [(0, 1), (2, 3), (0, 2), (1, 3)]
EDGE_PAIRS_3D = [(0, 1), (2, 3), (0, 2), (1, 3), (4, 5), (6, 7), (4, 6), (5, 7), (0, 4), (1, 5), (2, 6), (3, 7)]
FACE_QUADS = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 4, 5), (2, 3, 6, 7), (0, 2, 4, 6), (1, 3, 5, 7)]
# edges bounding each face, in the order the reference fills face_conn (subdomain.tpp:3205-3264):
# (k + 0*n_j, k + (n_j-1)*n_j, 0 + k*n_j, (n_j-1) + k*n_j)
FACE_EDGES = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 4, 8, 9), (1, 5, 10, 11), (2, 6, 8, 10), (3, 7, 9, 11)]


def corner_indices(dim, n):
    c = [0, n - 1, n * (n - 1), n * n - 1]
    if dim == 3:
        c = c + [x + n * n * (n - 1) for x in c]
    return c


def edge_points(dim, n, eid):
    """local point indices along edge eid, in the reference's parameterisation (subdomain.tpp:1197-1308)."""
    k = np.arange(n)
    if dim == 2:
        return [k, k + (n - 1) * n, k * n, (n - 1) + k * n][eid]
    nn = n * n
    return [k, k + (n - 1) * n, k * n, (n - 1) + k * n,
            k + (n - 1) * nn, k + (n - 1) * n + (n - 1) * nn, k * n + (n - 1) * nn, (n - 1) + k * n + (n - 1) * nn,
            k * nn, (n - 1) + k * nn, (n - 1) * n + k * nn, (n - 1) + (n - 1) * n + k * nn][eid]


def face_points(n, fid):
    """local point indices of face fid as a flat n*n list (subdomain.tpp:1366-1431)."""
    a = np.arange(n)
    nn = n * n
    if fid == 0:
        return (a[None, :] + a[:, None] * n).ravel()                   # i + j*n, index i + j*n
    if fid == 1:
        return (a[None, :] + a[:, None] * n + (n - 1) * nn).ravel()
    if fid == 2:
        return (a[None, :] + a[:, None] * nn).ravel()                  # i + k*nn, index i + k*n
    if fid == 3:
        return (a[None, :] + (n - 1) * n + a[:, None] * nn).ravel()
    if fid == 4:
        return (a[None, :] * n + a[:, None] * nn).ravel()              # j*n + k*nn, index j + k*n
    return ((n - 1) + a[None, :] * n + a[:, None] * nn).ravel()


class Region:
    pass


def _ranking(data):
    from .subdomain import ranking
    return ranking(data)


# ----------------------------------------------------------------------------------------------
# global data (identical on every rank)
# ----------------------------------------------------------------------------------------------
class GlobalMesh:
    def __init__(self, world):
        W = world.W
        dim = world.dim
        self.dim = dim
        self.nverts = 4 if dim == 2 else 8
        N = world.poly_degree[0]
        n = N + 1
        self.proc_count = [r.num_local_elements for r in W.ranks]
        self.proc_offset = np.concatenate([[0], np.cumsum(self.proc_count)[:-1]]).astype(int)
        self.total = int(sum(self.proc_count))
        cidx = corner_indices(dim, n)
        gm = []
        for r in W.ranks:
            g = r.glo_num.reshape(r.num_local_elements, -1)
            gm.append(g[:, cidx])
        self.geometry_mesh = np.concatenate(gm, axis=0)            # (total, nverts)
        self.partition = [(p, e) for p in range(world.num_procs) for e in range(self.proc_count[p])]
        T = self.total
        # connectivity (subdomain.tpp:282-430)
        verts = {}
        for e in range(T):
            for v in range(self.nverts):
                verts.setdefault(int(self.geometry_mesh[e, v]), set()).add(e)
        self.vert_conn = [[set(verts[int(self.geometry_mesh[e, v])]) - {e} for v in range(self.nverts)] for e in range(T)]
        pairs = EDGE_PAIRS_2D if dim == 2 else EDGE_PAIRS_3D
        edges = {}
        for e in range(T):
            for (a, b) in pairs:
                key = tuple(sorted((int(self.geometry_mesh[e, a]), int(self.geometry_mesh[e, b]))))
                edges.setdefault(key, set()).add(e)
        self.edge_conn = [[set(edges[tuple(sorted((int(self.geometry_mesh[e, a]), int(self.geometry_mesh[e, b]))))]) - {e} for (a, b) in pairs] for e in range(T)]
        self.face_conn = [[] for _ in range(T)]
        if dim == 3:
            faces = {}
            for e in range(T):
                for q in FACE_QUADS:
                    key = tuple(sorted(int(self.geometry_mesh[e, c]) for c in q))
                    faces.setdefault(key, set()).add(e)
            self.face_conn = [[set(faces[tuple(sorted(int(self.geometry_mesh[e, c]) for c in q))]) - {e} for q in FACE_QUADS] for e in range(T)]
        # expander: element adjacency incl. self (subdomain.tpp:432-453)
        self.adj = []
        for e in range(T):
            s = {e}
            for c in self.vert_conn[e]:
                s |= c
            for c in self.edge_conn[e]:
                s |= c
            for c in self.face_conn[e]:
                s |= c
            self.adj.append(s)

        # coarse (N = 1) global data (subdomain.tpp:1632-1713)
        Wc = world.domains[1]
        nv = self.nverts
        self.geom_fact_coarse = [np.concatenate([r.geom_fact[g] for r in Wc.ranks]) for g in range(6)]   # (total*nv)
        masked_glo = np.concatenate([np.where(r.dirichlet_mask > 0.0, r.glo_num, 0) for r in Wc.ranks]).astype(np.int64)
        self.glo_num_coarse = masked_glo.copy()
        # integer dense ranking (tpp:1666-1704)
        uniq, inv = np.unique(masked_glo, return_inverse=True)
        base = 0 if uniq[0] == 0 else 1
        self.dof_num_coarse = (inv + base).astype(np.int64)
        self.num_coarse_dofs = int(self.dof_num_coarse.max())
        self.Qt_coarse = CSRMatrix(self.num_coarse_dofs, T * nv)
        nz = np.flatnonzero(self.dof_num_coarse > 0)
        self.Qt_coarse.add_entries(self.dof_num_coarse[nz] - 1, nz, np.ones(nz.size))
        self.Qt_coarse.assemble()
        # global N = 1 stiffness matrix (tpp:1715-1849)
        from .subdomain import q1_element_matrix
        D2 = world.D_hat[world.num_levels - 1]
        rows, cols, vals = [], [], []
        for e in range(T):
            G = np.array([self.geom_fact_coarse[g][e * nv:(e + 1) * nv] for g in range(6)])
            Ae = q1_element_matrix(dim, D2, G)
            dof = self.dof_num_coarse[e * nv:(e + 1) * nv]
            ii, jj = np.nonzero(np.abs(Ae) > EPSILON)
            ok = (dof[ii] > 0) & (dof[jj] > 0)
            rows.append(dof[ii[ok]] - 1); cols.append(dof[jj[ok]] - 1); vals.append(Ae[ii[ok], jj[ok]])
        M = CSRMatrix(self.num_coarse_dofs, self.num_coarse_dofs)
        M.sparse_tolerance = -1.0
        M.add_entries(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))
        M.assemble()
        self.A_coarse = M.to_scipy()
        # BoomerAMG #1: coarsen 10, interp 6, max coarse size 1, theta 0.25 (tpp:1851-1858)
        self.amg_coarse = _amg.Hierarchy(self.A_coarse, max_coarse=1, with_smoother=False)


# ----------------------------------------------------------------------------------------------
# per-rank construction
# ----------------------------------------------------------------------------------------------
def build(world):
    from .subdomain import SubdomainRank
    G = GlobalMesh(world)
    world.global_mesh = G
    # J_cf_fem: hat-function interpolation for hanging nodes (tpp:2754-2783); r_gll as left by dgll_ (PNLEG snap)
    world.J_cf_fem = {}
    for lf in range(world.num_levels - 1):
        for lc in range(lf + 1, world.num_levels):
            Nf, Nc = world.poly_degree[lf], world.poly_degree[lc]
            nf, nc = Nf + 1, Nc + 1
            rf, rc = world.r_gll[lf], world.r_gll[lc]
            J = np.zeros(nf * nc)
            J[0] = 1.0
            for i in range(1, Nf):
                for j in range(Nc):
                    if rc[j] <= rf[i] <= rc[j + 1]:
                        J[i * nc + j] = (rc[j + 1] - rf[i]) / (rc[j + 1] - rc[j])
                        J[i * nc + j + 1] = (rf[i] - rc[j]) / (rc[j + 1] - rc[j])
            J[(nf - 1) * nc + (nc - 1)] = 1.0
            world.J_cf_fem[(Nc, Nf)] = J
    ranks = []
    for p in range(world.num_procs):
        S = SubdomainRank()
        S.proc_id = p
        _build_rank(world, G, S)
        world._alloc(S)
        ranks.append(S)
    return ranks


def _make_region(world, G, elem_ids, degrees):
    """pull mask / geometry / glo_num / coordinates of the region elements from their owners (tpp:644-805)."""
    R = Region()
    dim = world.dim
    R.elem_id = np.array(elem_ids, dtype=np.int32)
    R.elem_degree = np.array(degrees, dtype=np.int32)
    npe = (R.elem_degree.astype(np.int64) + 1) ** dim
    R.elem_offset = np.concatenate([[0], np.cumsum(npe)[:-1]]).astype(np.int32) if len(elem_ids) else np.zeros(0, np.int32)
    R.num_points = int(npe.sum())
    R.mask = np.zeros(R.num_points); R.x = np.zeros(R.num_points); R.y = np.zeros(R.num_points); R.z = np.zeros(R.num_points)
    R.glo_num = np.zeros(R.num_points, dtype=np.int64)
    R.dof_num = np.zeros(R.num_points, dtype=np.int64)
    R.geom_fact = [np.zeros(R.num_points) for _ in range(6)]
    for k, (gid, deg) in enumerate(zip(elem_ids, degrees)):
        owner, le = G.partition[gid]
        dr = world.domains[int(deg)].ranks[owner]
        n_ = int(npe[k])
        src = slice(le * n_, (le + 1) * n_)
        dst = slice(int(R.elem_offset[k]), int(R.elem_offset[k]) + n_)
        R.mask[dst] = dr.dirichlet_mask[src]
        R.x[dst] = dr.x[src]; R.y[dst] = dr.y[src]; R.z[dst] = dr.z[src]
        R.glo_num[dst] = dr.glo_num[src]
        for g in range(6):
            R.geom_fact[g][dst] = dr.geom_fact[g][src]
    return R


def _pts(R, k):
    o = int(R.elem_offset[k])
    n_ = (int(R.elem_degree[k]) + 1) ** R.dim
    return slice(o, o + n_)


def _build_rank(world, G, S):
    dim = world.dim
    p = S.proc_id
    T = G.total
    ladder = world.poly_degree
    nl = world.num_levels
    nv = G.nverts
    nloc = G.proc_count[p]
    off = int(G.proc_offset[p])

    # ---- computational regions (tpp:455-553) -------------------------------------------------
    sub_ids, sub_deg = list(range(off, off + nloc)), [ladder[0]] * nloc
    marked = np.zeros(T, dtype=bool)
    marked[off:off + nloc] = True
    reach = set(range(off, off + nloc))
    overlap = world.subdomain_overlap
    for l in range(nl):
        for _ in range(overlap):
            new = set()
            for e in reach:
                new |= G.adj[e]
            reach = new
        for e in range(T):
            if e in reach and not marked[e]:
                marked[e] = True
                sub_ids.append(e); sub_deg.append(ladder[l])
        if overlap == 0:
            overlap = 1
    S.num_subdomain_elems = len(sub_ids)
    new = set()
    for e in reach:
        new |= G.adj[e]
    reach1 = new
    sup_ids, sup_deg = [], []
    for e in range(T):
        if not marked[e]:
            if e in reach1:
                sub_ids.append(e); sub_deg.append(ladder[nl - 1])
            sup_ids.append(e); sup_deg.append(ladder[nl - 1])
    S.num_subdomain_extended_elems = len(sub_ids)
    S.num_superdomain_elems = len(sup_ids)
    near_sup = set()
    for e in range(T):
        if not marked[e]:
            near_sup |= G.adj[e]
    for e in range(T):
        if marked[e] and e in near_sup:
            sup_ids.append(e); sup_deg.append(ladder[nl - 1])
    S.num_superdomain_extended_elems = len(sup_ids)
    subdomain_partition = {e: k for k, e in enumerate(sub_ids)}

    sub = _make_region(world, G, sub_ids, sub_deg); sub.dim = dim
    sup = _make_region(world, G, sup_ids, sup_deg); sup.dim = dim
    S.sub_region, S.sup_region = sub, sup

    # ---- interface nodes (tpp:810-843) -------------------------------------------------------
    sub_glo = set()
    for k in range(S.num_subdomain_elems):
        if sub.elem_degree[k] == 1:
            sl = _pts(sub, k)
            sub_glo |= set(sub.glo_num[sl][sub.mask[sl] > 0.0].tolist())
    interface = set()
    for k in range(S.num_superdomain_elems):
        sl = _pts(sup, k)
        interface |= (set(sup.glo_num[sl].tolist()) & sub_glo)
    S.interface_glo_num = interface
    if interface:
        iface = np.array(sorted(interface), dtype=np.int64)
        for R in (sub, sup):
            hit = np.isin(R.glo_num, iface)
            R.dof_num[hit] = R.glo_num[hit]

    # ---- connectivity inside each region (tpp:845-878) ---------------------------------------
    for R, ids in ((sub, sub_ids), (sup, sup_ids)):
        mapping = {e: k for k, e in enumerate(ids)}
        R.vert_conn = [[sorted(mapping[x] for x in c if x in mapping) for c in G.vert_conn[e]] for e in ids]
        R.edge_conn = [[sorted(mapping[x] for x in c if x in mapping) for c in G.edge_conn[e]] for e in ids]
        R.face_conn = [[sorted(mapping[x] for x in c if x in mapping) for c in G.face_conn[e]] for e in ids]

    # ---- global numbering of the subdomain region (tpp:920-1098) -----------------------------
    global_offset = {ladder[0]: 0}
    for l in range(1, nl):
        global_offset[ladder[l]] = global_offset[ladder[l - 1]] + T * (ladder[l - 1] + 1) ** dim
    for k in range(len(sub_ids)):
        sl = _pts(sub, k)
        n = int(sub.elem_degree[k]) + 1
        g = sub.glo_num[sl]
        cidx = corner_indices(dim, n)
        corners = g[cidx].copy()
        g = g + global_offset[int(sub.elem_degree[k])]
        g[cidx] = corners
        sub.glo_num[sl] = g
    for k in range(len(sub_ids)):
        n = int(sub.elem_degree[k]) + 1
        sl = _pts(sub, k)
        g = sub.glo_num[sl]
        for eid, nbrs in enumerate(sub.edge_conn[k]):
            if any(sub.elem_degree[j] < sub.elem_degree[k] for j in nbrs):
                g[edge_points(dim, n, eid)[1:n - 1]] = 0
        if dim == 3:
            for fid, nbrs in enumerate(sub.face_conn[k]):
                if any(sub.elem_degree[j] < sub.elem_degree[k] for j in nbrs):
                    fp = face_points(n, fid).reshape(n, n)
                    g[fp[1:n - 1, 1:n - 1].ravel()] = 0
        sub.glo_num[sl] = g

    # ---- interface second to last, extended last (tpp:1100-1149) -----------------------------
    def elem_mask(R, cond):
        m = np.zeros(R.num_points, dtype=bool)
        for k in range(R.elem_id.size):
            if cond(k):
                m[_pts(R, k)] = True
        return m
    mx = int(sub.glo_num.max())
    sel = elem_mask(sub, lambda k: sub.elem_degree[k] == 1) & (sub.dof_num > 0)
    sub.glo_num[sel] += mx
    mx = int(sub.glo_num.max())
    sel = elem_mask(sub, lambda k: k >= S.num_subdomain_elems) & (sub.mask > 0.0) & (sub.dof_num == 0)
    sub.glo_num[sel] += mx
    if sup.num_points > 0:
        mx = int(sup.glo_num.max())
        sel = (sup.mask > 0.0) & (sup.dof_num == 0)
        sup.glo_num[sel] += mx
        mx = int(sup.glo_num.max())
        sel = elem_mask(sup, lambda k: k >= S.num_superdomain_elems) & (sup.mask > 0.0) & (sup.dof_num == 0)
        sup.glo_num[sel] += mx
    for R in (sub, sup):
        if R.num_points == 0:
            continue
        R.glo_num = _ranking(R.glo_num.astype(np.float64)).astype(np.int64)       # tpp:1157-1165
        R.dof_num = _ranking(R.glo_num.astype(np.float64) * R.mask).astype(np.int64)  # tpp:1167-1175

    # ---- region Q with interpolation rows on non-conforming edges / faces (tpp:1496-1585) -----
    for R in (sub, sup):
        R.Q = _build_Q(world, R)
        R.Qt = R.Q.transpose()

    # ---- subdomain operator (tpp:1587-1630) ---------------------------------------------------
    S.elem_id, S.elem_degree, S.elem_offset = sub.elem_id, sub.elem_degree, sub.elem_offset
    S.num_points = sub.num_points
    S.mask, S.geom_fact, S.x, S.y, S.z = sub.mask, sub.geom_fact, sub.x, sub.y, sub.z
    S.glo_num, S.dof_num = sub.glo_num, sub.dof_num
    S.Q, S.Qt = sub.Q, sub.Qt
    S.sub_num_dofs = 0
    for k in range(S.num_subdomain_elems):
        S.sub_num_dofs = max(S.sub_num_dofs, int(sub.dof_num[_pts(sub, k)].max()))
    S.sub_num_extended_dofs = sub.Q.num_cols
    level_of = {d: l for l, d in enumerate(ladder)}
    S.offset = np.concatenate([np.full((int(d) + 1) ** dim, o, dtype=np.int32) for o, d in zip(sub.elem_offset, sub.elem_degree)])
    S.vertex = np.concatenate([np.arange((int(d) + 1) ** dim, dtype=np.int32) for d in sub.elem_degree])
    S.level = np.concatenate([np.full((int(d) + 1) ** dim, level_of[int(d)], dtype=np.int32) for d in sub.elem_degree])

    # ---- superdomain composite grid (tpp:1860-2579) -------------------------------------------
    _build_superdomain(world, G, S, sub_ids, sup_ids, subdomain_partition)

    # ---- interface operator, weights (tpp:2581-2747) -------------------------------------------
    _build_interface(world, G, S, sub_ids, sup_ids, subdomain_partition)

    # ---- low-order FEM + AMG (tpp:2749-3549) ---------------------------------------------------
    if world.use_preconditioner:
        S.A_sub_fem = _assemble_fem(world, S, sub)
        ne, nd = S.sub_num_extended_dofs, S.num_dofs
        x = np.arange(nd, dtype=np.float64)
        m = np.zeros(ne + S.sup_num_extended_dofs)
        S.Q_int.multiply(m, x)                                                   # tpp:3414-3417
        m = m.astype(np.int64)
        rows, cols, vals = [], [], []
        A = S.A_sub_fem.tocsr()
        for i in range(S.sub_num_dofs):
            for ptr_ in range(A.indptr[i], A.indptr[i + 1]):
                rows.append(m[i]); cols.append(m[A.indices[ptr_]]); vals.append(A.data[ptr_])
        if S.sup_num_dofs > 0:
            As = S.A_sup
            for i in range(S.num_interface_dofs, S.sup_num_dofs):
                for ptr_ in range(As.ptr[i], As.ptr[i + 1]):
                    rows.append(m[ne + i]); cols.append(m[ne + As.col[ptr_]]); vals.append(As.val[ptr_])
        M = CSRMatrix(nd, nd)
        M.sparse_tolerance = -1.0
        M.add_entries(rows, cols, vals)
        M.assemble()
        S.A_fem = M.to_scipy()
        S.amg = _amg.Hierarchy(S.A_fem, cheby_order=world.cheby_order, dtype=np.float32 if getattr(world, "amg_precision", "double") == "float" else np.float64)


def _matching(dim, R, ki, kj, kind, idx):
    """local indices on element kj of the edge (kind 0) / face (kind 1) `idx` of element ki, identified through the
    corner ids and assumed identically oriented (subdomain.tpp:1179-1494)."""
    ni, nj = int(R.elem_degree[ki]) + 1, int(R.elem_degree[kj]) + 1
    gi = R.glo_num[_pts(R, ki)]
    gj = R.glo_num[_pts(R, kj)]
    if kind == 0:
        pi = edge_points(dim, ni, idx)
        ends = {int(gi[pi[0]]), int(gi[pi[-1]])}
        for eid in range(4 if dim == 2 else 12):
            # the reference tests the candidate edges of elem_j in the order x-edges, y-edges (bottom), then top, then z
            pass
        order = [0, 1, 2, 3] if dim == 2 else [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]
        for eid in order:
            pj = edge_points(dim, nj, eid)
            if int(gj[pj[0]]) in ends and int(gj[pj[-1]]) in ends:
                return pi, pj
        raise RuntimeError("matching_edge: no matching edge")
    pi = face_points(ni, idx)
    cs = {int(gi[pi[0]]), int(gi[pi[ni - 1]]), int(gi[pi[(ni - 1) * ni]]), int(gi[pi[ni * ni - 1]])}
    for fid in range(6):
        pj = face_points(nj, fid)
        if all(int(gj[pj[c]]) in cs for c in (0, nj - 1, (nj - 1) * nj, nj * nj - 1)):
            return pi, pj
    raise RuntimeError("matching_face: no matching face")


def _min_degree_edge_neighbor(R, k, eid):
    e_j, N_j = -1, int(R.elem_degree[k])
    for e in R.edge_conn[k][eid]:
        if R.elem_degree[e] < N_j:
            e_j, N_j = e, int(R.elem_degree[e])
    return e_j, N_j


def _build_Q(world, R):
    dim = world.dim
    ndofs = int(R.dof_num.max()) if R.num_points else 0
    Q = CSRMatrix(R.num_points, ndofs)
    for k in range(R.elem_id.size):
        sl = _pts(R, k)
        o = int(R.elem_offset[k])
        Ni = int(R.elem_degree[k]); ni = Ni + 1
        dof = R.dof_num[sl]
        nz = np.flatnonzero(dof > 0)
        Q.add_entries(o + nz, dof[nz] - 1, np.ones(nz.size))
        for eid in range(len(R.edge_conn[k])):
            e_j, Nj = _min_degree_edge_neighbor(R, k, eid)
            if e_j < 0:
                continue
            nj = Nj + 1
            pi, pj = _matching(dim, R, k, e_j, 0, eid)
            J = world.J_cf[(Nj, Ni)]
            dj = R.dof_num[_pts(R, e_j)]
            for i in range(1, ni - 1):
                for j in range(nj):
                    if dj[pj[j]] > 0:
                        Q.add_entry(o + int(pi[i]), int(dj[pj[j]]) - 1, J[i * nj + j])
        if dim == 3:
            for fid in range(6):
                for e_j in R.face_conn[k][fid]:
                    Nj = int(R.elem_degree[e_j]); nj = Nj + 1
                    if Ni > Nj:
                        pi, pj = _matching(dim, R, k, e_j, 1, fid)
                        J = world.J_cf[(Nj, Ni)]
                        dj = R.dof_num[_pts(R, e_j)]
                        for j in range(1, ni - 1):
                            for i in range(1, ni - 1):
                                for q in range(nj):
                                    for p_ in range(nj):
                                        if dj[pj[p_ + q * nj]] > 0:
                                            Q.add_entry(o + int(pi[i + j * ni]), int(dj[pj[p_ + q * nj]]) - 1, J[i * nj + p_] * J[j * nj + q])
    Q.assemble()
    return Q


def _csr(M):
    M = M.tocsr(); M.sort_indices()
    return M


def _build_superdomain(world, G, S, sub_ids, sup_ids, subdomain_partition):
    nv = G.nverts
    ncd = G.num_coarse_dofs
    dofc, gloc = G.dof_num_coarse, G.glo_num_coarse
    # dof markers (tpp:1860-1905)
    dof_marker = np.zeros(ncd, dtype=np.int64)
    for k in range(S.num_subdomain_elems):
        eid = sub_ids[k]
        for v in range(nv):
            dof = int(dofc[eid * nv + v]); glo = int(gloc[eid * nv + v])
            if dof > 0:
                dof_marker[dof - 1] = 1
            if glo in S.interface_glo_num:
                dof_marker[dof - 1] = 2
    for k in range(S.num_subdomain_elems, S.num_subdomain_extended_elems):
        eid = sub_ids[k]
        for v in range(nv):
            dof = int(dofc[eid * nv + v])
            if dof > 0 and dof_marker[dof - 1] == 0:
                dof_marker[dof - 1] = 3
    for k in range(S.num_superdomain_elems, S.num_superdomain_extended_elems):
        eid = sup_ids[k]
        for v in range(nv):
            dof = int(dofc[eid * nv + v])
            if dof > 0 and dof_marker[dof - 1] == 1:
                dof_marker[dof - 1] = 4
    S.dof_marker = dof_marker

    H = G.amg_coarse
    nlev = H.num_levels
    A = [_csr(L.A) for L in H.levels]
    Pm = [_csr(L.P) for L in H.levels[:-1]]
    cfm = [L.cf for L in H.levels[:-1]]
    num_nodes = [L.n for L in H.levels]
    D = [np.zeros(n) for n in num_nodes]
    D[0][dof_marker > 0] = 1.0
    ncl = 0
    ov = world.superdomain_overlap
    for l in range(nlev):
        ncl = l + 1
        w = D[l].copy()
        pat = A[l].copy(); pat.data[:] = 1.0
        for _ in range(ov):
            w = pat @ w
        if ov == 0:
            ov = 1
        if l == nlev - 1:
            w[:] = 1.0
        D[l][(D[l] == 0.0) & (w > 0.0)] = 2.0
        if not np.any(D[l] == 0.0):
            break
        if l < nlev - 1:
            crow = np.flatnonzero((cfm[l] == 1) & (D[l] > 0.0))
            D[l + 1][Pm[l].indices[Pm[l].indptr[crow]]] = 1.0
    num_local = [int((D[l] == 1.0).sum()) for l in range(ncl)]
    num_overlap = [int((D[l] == 2.0).sum()) for l in range(ncl)]
    num_remaining = [int((D[l] == 0.0).sum()) for l in range(ncl)]
    num_comp_overlap = list(num_overlap)
    num_comp_overlap[0] += num_local[0]

    def c_rows(l):
        rows = np.flatnonzero(cfm[l] == 1)
        return rows, Pm[l].indices[Pm[l].indptr[rows]]

    nodes_to_fine = [np.arange(num_nodes[0])]
    for l in range(ncl - 1):
        ntf = np.zeros(num_nodes[l + 1], dtype=np.int64)
        rows, cols = c_rows(l)
        ntf[cols] = nodes_to_fine[l][rows]
        nodes_to_fine.append(ntf)
    nodes_to_dofs = [np.full(num_nodes[l], -1, dtype=np.int64) for l in range(ncl)]
    dof_end = 0
    for marker in (1, 2, 3, 4):
        idx = np.flatnonzero(dof_marker == marker)
        nodes_to_dofs[0][idx] = dof_end + np.arange(idx.size)
        dof_end += idx.size
    dof_end = num_local[0]
    idx = np.flatnonzero(D[0] == 2.0)
    nodes_to_dofs[0][idx] = dof_end + np.arange(idx.size)
    offset = num_local[0] + num_overlap[0]
    for l in range(ncl - 1):
        idx = np.flatnonzero(D[l + 1] == 2.0)
        nodes_to_dofs[0][nodes_to_fine[l + 1][idx]] = offset + np.arange(idx.size)
        offset += idx.size
    for l in range(ncl - 1):
        rows, cols = c_rows(l)
        nodes_to_dofs[l + 1][cols] = nodes_to_dofs[l][rows]
    num_dofs = offset

    P_c = [None] * max(ncl - 1, 0)
    R_c = [None] * max(ncl - 1, 0)
    for l in range(ncl - 1, 0, -1):
        Pl = Pm[l - 1]
        nf_, ncn = num_nodes[l - 1], num_nodes[l]
        fine = np.full(nf_, -1, dtype=np.int64)
        if l - 1 == 0:
            de = 0
            for marker in (1, 2, 3, 4):
                idx = np.flatnonzero(dof_marker == marker)
                fine[idx] = de + np.arange(idx.size); de += idx.size
            idx = np.flatnonzero(D[0] == 2.0)
            fine[idx] = de + np.arange(idx.size)
            de = num_local[0] + num_overlap[0]
            idx = np.flatnonzero(D[0] == 0.0)
            fine[idx] = de + np.arange(idx.size)
        else:
            idx = np.flatnonzero(D[l - 1] == 2.0)
            fine[idx] = np.arange(idx.size)
            idx = np.flatnonzero(D[l - 1] == 0.0)
            fine[idx] = num_overlap[l - 1] + np.arange(idx.size)
        coarse = np.full(ncn, -1, dtype=np.int64)
        de = (num_local[0] + num_overlap[0]) if l - 1 == 0 else num_overlap[l - 1]
        idx = np.flatnonzero((D[l] == 2.0) | (D[l] == 0.0))
        coarse[idx] = de + np.arange(idx.size)
        bound = (num_local[0] + num_overlap[0]) if l - 1 == 0 else num_overlap[l - 1]
        rows, cols = c_rows(l - 1)
        for r_, c_ in zip(rows, cols):          # sequential, later rows overwrite (tpp:2160-2175)
            if fine[r_] < bound:
                coarse[c_] = fine[r_]
        num_fine = num_overlap[l - 1] + (num_local[0] if l - 1 == 0 else 0)
        if l - 1 == 0:
            shape = (num_nodes[0], num_local[0] + num_overlap[0] + num_overlap[l] + num_remaining[l])
        else:
            shape = (num_overlap[l - 1] + num_remaining[l - 1], num_overlap[l - 1] + num_overlap[l] + num_remaining[l])
        rr, cc, vv = [], [], []
        for row in range(nf_):
            fr = fine[row]
            if fr < 0:
                continue
            if fr < num_fine:
                rr.append(fr); cc.append(fr); vv.append(1.0)
            else:
                for ptr_ in range(Pl.indptr[row], Pl.indptr[row + 1]):
                    col = coarse[Pl.indices[ptr_]]
                    if col >= 0:
                        rr.append(fr); cc.append(col); vv.append(Pl.data[ptr_])
        P_c[l - 1] = sp.csr_matrix((vv, (rr, cc)), shape=shape)
        keep = np.flatnonzero(fine >= 0)
        R_c[l - 1] = sp.csr_matrix((np.ones(keep.size), (np.arange(keep.size), fine[keep])), shape=(keep.size, keep.size))

    if ncl > 1:
        for l in range(ncl - 2, 0, -1):
            Pc = _csr(P_c[l - 1])
            nco = num_comp_overlap[l - 1]
            lower = Pc[nco:, :]
            P21 = lower[:, :nco]
            P22 = lower[:, nco:]
            RlPl = R_c[l] @ P_c[l]
            P22n = P22 @ RlPl
            top = sp.hstack([sp.identity(nco, format="csr"), sp.csr_matrix((nco, P22n.shape[1]))])
            bot = sp.hstack([P21, P22n])
            P_c[l - 1] = _csr(sp.vstack([top, bot]))
        Pfull = _csr(R_c[0] @ P_c[0])
    else:
        Pfull = sp.csr_matrix((np.ones(num_dofs), (np.arange(num_dofs), nodes_to_dofs[0][:num_dofs])), shape=(num_dofs, num_dofs))
    PtAP = _csr(Pfull.T @ A[0] @ Pfull)

    marker_count = [int((dof_marker == m).sum()) for m in (1, 2, 3, 4)] + [0]
    marker_offset = [0] * 5
    for m in range(1, 5):
        marker_offset[m] = marker_offset[m - 1] + marker_count[m - 1]
    nrows = PtAP.shape[0]
    marker_count[4] = nrows - marker_offset[4]
    R_sup = np.full(nrows, -1, dtype=np.int64)
    dof = 0
    for rng in (range(marker_offset[1], marker_offset[3]), range(marker_offset[4], nrows), range(marker_offset[3], marker_offset[4])):
        for i in rng:
            R_sup[i] = dof; dof += 1
    ncols = marker_count[1] + marker_count[2] + marker_count[3] + marker_count[4]
    A_sup = CSRMatrix(ncols, ncols)
    coo = PtAP.tocoo()
    ok = (R_sup[coo.row] >= 0) & (R_sup[coo.col] >= 0)
    A_sup.add_entries(R_sup[coo.row[ok]], R_sup[coo.col[ok]], coo.data[ok])      # CSR_Matrix::add_entry drops |v| <= 1e-12 (tpp:2563-2575)
    A_sup.assemble()
    Pt = CSRMatrix(dof, Pfull.shape[0])
    coo = Pfull.tocoo()
    ok = R_sup[coo.col] >= 0
    Pt.add_entries(R_sup[coo.col[ok]], coo.row[ok], coo.data[ok])                  # tpp:2549-2561
    Pt.assemble()

    dof_sup = nodes_to_dofs[0].copy()
    dof_sup[dof_marker == 1] = -1
    dof_max = int(nodes_to_dofs[0].max())
    dof_sup[dof_marker == 4] += dof_max
    uniq, inv = np.unique(dof_sup, return_inverse=True)
    base = 0 if uniq[0] == -1 else 1
    S.dof_sup = (inv + base).astype(np.int64)

    S.sup_num_dofs = ncols - marker_count[3]
    S.sup_num_extended_dofs = dof
    S.A_sup, S.Pt = A_sup, Pt
    S.Qt_coarse = G.Qt_coarse
    S.marker_count = marker_count
    S.num_comp_levels = ncl


def _build_interface(world, G, S, sub_ids, sup_ids, subdomain_partition):
    nv = G.nverts
    sub, sup = S.sub_region, S.sup_region
    dofc = G.dof_num_coarse
    S.num_interface_dofs = len(S.interface_glo_num)
    S.num_dofs = S.sub_num_dofs + S.sup_num_dofs - S.num_interface_dofs
    ni_ = S.num_interface_dofs
    shift = S.sub_num_dofs - ni_
    sub_map = {}
    for k in range(S.num_subdomain_elems):
        for d in sub.dof_num[_pts(sub, k)]:
            if d > 0:
                sub_map[int(d)] = int(d)
    for k in range(S.num_subdomain_elems, S.num_subdomain_extended_elems):
        eid = sub_ids[k]
        dn = sub.dof_num[_pts(sub, k)]
        for v in range(nv):
            dof = int(dofc[eid * nv + v])
            if dof > 0 and S.dof_sup[dof - 1] > 0:
                sub_map[int(dn[v])] = int(S.dof_sup[dof - 1]) + shift
    sup_map = {}
    for k in range(S.num_superdomain_elems):
        eid = sup_ids[k]
        for v in range(nv):
            dof = int(dofc[eid * nv + v])
            if dof > 0:
                sup_map[int(S.dof_sup[dof - 1])] = int(S.dof_sup[dof - 1]) + shift
    for k in range(S.num_superdomain_elems, S.num_superdomain_extended_elems):
        eid = sup_ids[k]
        sk = subdomain_partition[eid]
        sdn = sub.dof_num[_pts(sub, sk)]
        for v in range(nv):
            dof = int(dofc[eid * nv + v])
            if dof > 0 and S.dof_marker[dof - 1] == 4:
                sup_map[int(S.dof_sup[dof - 1])] = int(sdn[v])
    ne, ns = S.sub_num_extended_dofs, S.sup_num_extended_dofs
    S.Q_int = CSRMatrix(ne + ns, S.num_dofs)
    for i in range(ne):
        S.Q_int.add_entry(i, sub_map[i + 1] - 1, 1.0)
    for i in range(ns):
        S.Q_int.add_entry(ne + i, sup_map[i + 1] - 1, 1.0)
    S.Q_int.assemble()
    S.Qt_int = CSRMatrix(S.num_dofs, ne + ns)
    for i in range(S.sub_num_dofs):
        S.Qt_int.add_entry(i, i, 1.0)
    for i in range(S.sup_num_dofs - ni_):
        S.Qt_int.add_entry(S.sub_num_dofs + i, ne + ni_ + i, 1.0)
    S.Qt_int.assemble()
    S.QQt_int = CSRMatrix(ne + ns, ne + ns)
    seen = np.zeros(ne + ns)
    for i in range(S.sub_num_dofs):
        S.QQt_int.add_entry(i, i, 1.0); seen[i] = 1.0
    for k in range(S.num_subdomain_elems, S.num_subdomain_extended_elems):
        eid = sub_ids[k]
        dn = sub.dof_num[_pts(sub, k)]
        for v in range(nv):
            if dn[v] > 0 and seen[dn[v] - 1] == 0:
                S.QQt_int.add_entry(int(dn[v]) - 1, ne + int(S.dof_sup[dofc[eid * nv + v] - 1]) - 1, 1.0)
                seen[dn[v] - 1] = 1.0
    for i in range(ni_):
        S.QQt_int.add_entry(ne + i, S.sub_num_dofs - ni_ + i, 1.0); seen[ne + i] = 1.0
    for i in range(ni_, S.sup_num_dofs):
        S.QQt_int.add_entry(ne + i, ne + i, 1.0); seen[ne + i] = 1.0
    for k in range(S.num_superdomain_elems, S.num_superdomain_extended_elems):
        eid = sup_ids[k]
        sk = subdomain_partition[eid]
        sdn = sub.dof_num[_pts(sub, sk)]
        for v in range(nv):
            if dofc[eid * nv + v] > 0:
                dof = int(S.dof_sup[dofc[eid * nv + v] - 1])
                if seen[ne + dof - 1] == 0:
                    S.QQt_int.add_entry(ne + dof - 1, int(sdn[v]) - 1, 1.0)
                    seen[ne + dof - 1] = 1.0
    S.QQt_int.assemble()
    # weights (tpp:2731-2747)
    S.norm_weight = np.ones(ne + ns)
    S.norm_weight[S.sub_num_dofs:ne] = 0.0
    S.norm_weight[ne:ne + ni_] = 0.0
    S.norm_weight[ne + S.sup_num_dofs:ne + ns] = 0.0
    S.num_values = S.num_points + ns
    S.inner_weight = np.zeros(S.num_values)
    S.Q.multiply(S.inner_weight, S.norm_weight)
    S.inner_weight[S.num_points:] = S.norm_weight[ne:]
    S.inner_weight[S.inner_weight > 0.0] = 1.0


def _assemble_fem(world, S, sub):
    """A_sub_fem with hanging-node elimination J_e^T A_e J_e per element (subdomain.tpp:2913-3412)."""
    from .subdomain import low_order_element_matrix, q1_element_matrix
    dim = world.dim
    nedges = 4 if dim == 2 else 12
    rows, cols, vals = [], [], []
    for k in range(sub.elem_id.size):
        Ni = int(sub.elem_degree[k]); ni = Ni + 1
        sl = _pts(sub, k)
        npts = ni ** dim
        if Ni > 1:
            Ae = low_order_element_matrix(dim, Ni, sub.x[sl], sub.y[sl], sub.z[sl])
        else:
            Gf = np.array([g[sl] for g in sub.geom_fact])
            Ae = q1_element_matrix(dim, world.D_hat[world.num_levels - 1], Gf)
            Ae = np.where(np.abs(Ae) > EPSILON, Ae, 0.0)
        glo = sub.glo_num[sl]; dofn = sub.dof_num[sl]
        # column numbering of J_e: own points with glo_num > 0 first, then the coarse neighbours' edge / face interiors
        rank = 1
        vfirst = np.zeros(npts, dtype=np.int64)
        for v in range(npts):
            if glo[v] > 0:
                vfirst[v] = rank; rank += 1
        vert = [(int(vfirst[v]), int(dofn[v])) for v in range(npts)]
        edge = [None] * nedges
        for eid in range(nedges):
            e_j, Nj = _min_degree_edge_neighbor(sub, k, eid)
            if e_j < 0:
                continue
            nj = Nj + 1
            pi, pj = _matching(dim, sub, k, e_j, 0, eid)
            dj = sub.dof_num[_pts(sub, e_j)]
            lst = [None] * nj
            lst[0] = vert[int(pi[0])]; lst[nj - 1] = vert[int(pi[ni - 1])]
            for t in range(1, nj - 1):
                lst[t] = (rank, int(dj[pj[t]])); rank += 1
            edge[eid] = (pi, lst)
        face = [None] * (6 if dim == 3 else 0)
        if dim == 3:
            for fid in range(6):
                for e_j in sub.face_conn[k][fid]:
                    Nj = int(sub.elem_degree[e_j]); nj = Nj + 1
                    if Ni > Nj:
                        pi, pj = _matching(dim, sub, k, e_j, 1, fid)
                        dj = sub.dof_num[_pts(sub, e_j)]
                        lst = [None] * (nj * nj)
                        lst[0] = vert[int(pi[0])]; lst[nj - 1] = vert[int(pi[ni - 1])]
                        lst[(nj - 1) * nj] = vert[int(pi[(ni - 1) * ni])]; lst[nj * nj - 1] = vert[int(pi[ni * ni - 1])]
                        e0, e1, e2, e3 = FACE_EDGES[fid]
                        for t in range(1, nj - 1):
                            lst[t] = edge[e0][1][t]
                            lst[t + (nj - 1) * nj] = edge[e1][1][t]
                            lst[t * nj] = edge[e2][1][t]
                            lst[(nj - 1) + t * nj] = edge[e3][1][t]
                        for j in range(1, nj - 1):
                            for i in range(1, nj - 1):
                                lst[i + j * nj] = (rank, int(dj[pj[i + j * nj]])); rank += 1
                        face[fid] = (pi, lst)
        ncols = rank - 1
        Je = np.zeros((npts, ncols))
        for v in range(npts):
            if vert[v][0] > 0:
                Je[v, vert[v][0] - 1] += 1.0
        for eid in range(nedges):
            if edge[eid] is None:
                continue
            pi, lst = edge[eid]
            nj = len(lst)
            Jf = world.J_cf_fem[(nj - 1, Ni)]
            for i in range(1, ni - 1):
                for j in range(nj):
                    val = Jf[i * nj + j]
                    if abs(val) > EPSILON:
                        Je[int(pi[i]), lst[j][0] - 1] += val
        for fid in range(len(face)):
            if face[fid] is None:
                continue
            pi, lst = face[fid]
            nj = int(round(np.sqrt(len(lst))))
            Jf = world.J_cf_fem[(nj - 1, Ni)]
            for j in range(1, ni - 1):
                for i in range(1, ni - 1):
                    for q in range(nj):
                        for p_ in range(nj):
                            val = Jf[i * nj + p_] * Jf[j * nj + q]
                            if abs(val) > EPSILON:
                                Je[int(pi[i + j * ni]), lst[p_ + q * nj][0] - 1] += val
        JtAJ = Je.T @ Ae @ Je
        dcol = np.zeros(ncols, dtype=np.int64)
        for v in range(npts):
            if vert[v][0] > 0:
                dcol[vert[v][0] - 1] = vert[v][1]
        for eid in range(nedges):
            if edge[eid] is not None:
                for (r_, d_) in edge[eid][1]:
                    dcol[r_ - 1] = d_
        for fid in range(len(face)):
            if face[fid] is not None:
                for (r_, d_) in face[fid][1]:
                    dcol[r_ - 1] = d_
        ii, jj = np.nonzero(np.abs(JtAJ) > EPSILON)
        ok = (dcol[ii] > 0) & (dcol[jj] > 0)
        rows.append(dcol[ii[ok]] - 1); cols.append(dcol[jj[ok]] - 1); vals.append(JtAJ[ii[ok], jj[ok]])
    n = S.sub_num_extended_dofs
    M = CSRMatrix(n, n)
    M.sparse_tolerance = -1.0
    M.add_entries(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))
    M.assemble()
    return M.to_scipy()


# ----------------------------------------------------------------------------------------------
# tree operator (subdomain.tpp:4566-4646)
# ----------------------------------------------------------------------------------------------
def tree_operator(world, Tu_list, u_list):
    L = _L()
    dim = world.dim
    G = world.global_mesh
    nl = world.num_levels
    # every rank: cast + ladder restrictions of its own elements (tpp:4571-4609)
    level_data = []
    for p in range(world.num_procs):
        E = G.proc_count[p]
        lev = [np.ascontiguousarray(u_list[p], dtype=np.float64).copy()]
        for l in range(nl - 1):
            nf, nc = world.poly_degree[l] + 1, world.poly_degree[l + 1] + 1
            J = world.J_cf[(world.poly_degree[l + 1], world.poly_degree[l])]
            uf = lev[l]
            if dim == 2:
                t1 = np.zeros(E * nf * nc); uc = np.zeros(E * nc * nc)
                L.o_restriction_1(P(t1), P(J), P(uf), C.c_int(t1.size), C.c_int(nf), C.c_int(nc), C.c_int(2))
                L.o_restriction_2(P(uc), P(J), P(t1), C.c_int(uc.size), C.c_int(nf), C.c_int(nc), C.c_int(2))
            else:
                t1 = np.zeros(E * nf * nf * nc); t2 = np.zeros(E * nf * nc * nc); uc = np.zeros(E * nc ** 3)
                L.o_restriction_1(P(t1), P(J), P(uf), C.c_int(t1.size), C.c_int(nf), C.c_int(nc), C.c_int(3))
                L.o_restriction_2(P(t2), P(J), P(t1), C.c_int(t2.size), C.c_int(nf), C.c_int(nc), C.c_int(3))
                L.o_restriction_3(P(uc), P(J), P(t2), C.c_int(uc.size), C.c_int(nf), C.c_int(nc))
            lev.append(uc)
        level_data.append(lev)
    # MPI_Allgatherv of the N = 1 level (tpp:4619-4622)
    coarse_all = np.concatenate([level_data[p][nl - 1] for p in range(world.num_procs)])
    level_of = {d: l for l, d in enumerate(world.poly_degree)}
    for S, Tu in zip(world.ranks, Tu_list):
        # gslib_gs: every region slot receives the owner's value at the element's level (tpp:4625-4631)
        for k in range(S.elem_id.size):
            owner, le = G.partition[int(S.elem_id[k])]
            d = int(S.elem_degree[k]); n_ = (d + 1) ** dim
            o = int(S.elem_offset[k])
            Tu[o:o + n_] = level_data[owner][level_of[d]][le * n_:(le + 1) * n_]
        # superdomain: assemble the global coarse residual, restrict to the composite dofs (tpp:4634-4645)
        if S.sup_num_extended_dofs > 0:
            w1 = np.zeros(G.num_coarse_dofs)
            G.Qt_coarse.multiply(w1, coarse_all)
            out = np.zeros(S.sup_num_extended_dofs)
            S.Pt.multiply(out, w1)
            Tu[S.num_points:S.num_points + S.sup_num_extended_dofs] = out
