"""
ORACLE tooling (test infrastructure, NOT product code).

Independent numpy generator of synthetic box meshes in the reference's exact on-disk format
(the Nek5000 dumps read at /root/reference/domain.tpp:43-224, BINARY_INPUT true):

    <dir>/lx1_<N+1>/size_<p>.<N>.dat          text  "dim n_x n_y n_z num_local_elements"
    <dir>/lx1_<N+1>/{x,y,z}_<p>.<N>.dat        raw f64, element-major, n^dim per element
    <dir>/lx1_<N+1>/glo_num_<p>.<N>.dat        raw i64 (1-based global node ids)
    <dir>/lx1_<N+1>/node_degree_<p>.<N>.dat    raw i32 (global multiplicity of the node)
    <dir>/lx1_<N+1>/p_mask_<p>.<N>.dat         raw f64 (0 on the Dirichlet boundary, 1 inside)
    <dir>/lx1_<N+1>/g_<1..6>_<p>.<N>.dat       raw f64 geometric factors, order [11,22,33,12,13,23]
                                               (2D: [11,22,12,0,0,0]); all six are always read

Conventions the reference imposes silently:
  * corner glo_nums are identical in every degree directory (subdomain.tpp:930-966), so
    vertices are numbered first (1..Nv), all other nodes after them;
  * element order on a rank is lexicographic inside the rank's block; the global element id
    is rank-major (proc_offset[p] + e, subdomain.tpp:219-280).
Geometric factors follow Nek5000's isoparametric recipe: G_ab = w_i w_j w_k * J * (grad r_a . grad r_b)
with the Jacobian obtained by spectral differentiation of the coordinates.

Used by the tests to cross-check the product's C++ generator (same format, same integers,
floats to 1e-13) -- never by the product.
"""
import os
import numpy as np

from . import capi as _c


def gll(n):
    """GLL nodes, weights and D (D[i, j] = dl_j/dxi(xi_i)) from the oracle speclib restatement."""
    z, w = _c.zwgll(n)
    D = _c.dgll(z.copy(), n)
    return z, w, D


def rank_layout(nranks, dim):
    """px,py,pz with px*py*pz == nranks, as cubic as possible, x fastest growing last (2 -> 2x1x1 ...)."""
    p = [1, 1, 1]
    d = 0
    r = nranks
    while r > 1:
        assert r % 2 == 0, "rank count must be a power of two"
        p[d % dim] *= 2
        r //= 2
        d += 1
    return tuple(p)


def deform(X, eps):
    """Smooth boundary-preserving deformation of [0,1]^dim (identity when eps == 0)."""
    if eps == 0.0:
        return X
    s = np.ones_like(X[0])
    for c in X:
        s = s * np.sin(np.pi * c)
    out = []
    for d, c in enumerate(X):
        out.append(c + eps * (0.5 + 0.25 * d) * s)
    return out


def generate(directory, dim, nel, N, nranks=1, eps=0.0, write=True):
    """
    nel: elements per side (int or tuple).  Returns a list (one per rank) of dicts with the arrays.
    """
    if isinstance(nel, int):
        nel = (nel,) * dim
    nel = tuple(nel) + (1,) * (3 - dim)
    n = N + 1
    z, w, D = gll(n)
    P = rank_layout(nranks, dim)
    for d in range(3):
        assert nel[d] % P[d] == 0

    # global node grid
    gshape = [nel[d] * N + 1 if d < dim else 1 for d in range(3)]
    idx = np.indices(gshape[::-1])  # [k, j, i] ordering -> idx[0]=k
    K, J, I = idx[0], idx[1], idx[2]
    is_vert = (I % N == 0) & (J % N == 0) & (K % N == 0)
    ids = np.zeros(gshape[::-1], dtype=np.int64)
    nvx, nvy = nel[0] + 1, nel[1] + 1
    vid = 1 + (I // N) + (J // N) * nvx + (K // N) * nvx * nvy
    nv = int(is_vert.sum())
    ids[is_vert] = vid[is_vert]
    nonv = ~is_vert
    ids[nonv] = nv + np.cumsum(nonv.ravel())[nonv.ravel()]
    # multiplicity: product over directions of (2 if shared interior element boundary else 1)
    def mult1(ix, ne):
        m = np.ones_like(ix)
        m[(ix % N == 0) & (ix > 0) & (ix < ne * N)] = 2
        return m
    mult = mult1(I, nel[0]) * (mult1(J, nel[1]) if dim >= 2 else 1) * (mult1(K, nel[2]) if dim >= 3 else 1)
    on_bdry = (I == 0) | (I == nel[0] * N)
    if dim >= 2:
        on_bdry |= (J == 0) | (J == nel[1] * N)
    if dim >= 3:
        on_bdry |= (K == 0) | (K == nel[2] * N)

    # reference coordinates on the global grid
    def coord1(ix, ne):
        e = np.minimum(ix // N, ne - 1)
        l = ix - e * N
        return (e + 0.5 * (z[l] + 1.0)) / ne
    Xg = [coord1(I, nel[0])]
    if dim >= 2:
        Xg.append(coord1(J, nel[1]))
    if dim >= 3:
        Xg.append(coord1(K, nel[2]))
    Xg = deform(Xg, eps)

    out = []
    npts = n ** dim
    bl = [nel[d] // P[d] for d in range(3)]
    for p in range(nranks):
        pc = (p % P[0], (p // P[0]) % P[1], p // (P[0] * P[1]))
        E = bl[0] * bl[1] * bl[2]
        x = np.zeros((E, npts)); y = np.zeros((E, npts)); zc = np.zeros((E, npts))
        glo = np.zeros((E, npts), dtype=np.int64)
        deg = np.zeros((E, npts), dtype=np.int32)
        mask = np.zeros((E, npts))
        G = np.zeros((6, E, npts))
        e = 0
        for ez in range(bl[2]):
            for ey in range(bl[1]):
                for ex in range(bl[0]):
                    g0 = [(pc[0] * bl[0] + ex) * N, (pc[1] * bl[1] + ey) * N, (pc[2] * bl[2] + ez) * N]
                    sl = (slice(g0[2], g0[2] + (n if dim >= 3 else 1)),
                          slice(g0[1], g0[1] + (n if dim >= 2 else 1)),
                          slice(g0[0], g0[0] + n))
                    glo[e] = ids[sl].ravel()
                    deg[e] = mult[sl].ravel()
                    mask[e] = np.where(on_bdry[sl], 0.0, 1.0).ravel()
                    xe = [c[sl].reshape((n,) * dim) for c in Xg]  # index order [k][j][i]
                    x[e] = xe[0].ravel()
                    if dim >= 2:
                        y[e] = xe[1].ravel()
                    if dim >= 3:
                        zc[e] = xe[2].ravel()
                    G[:, e, :] = _geom(xe, D, w, dim)
                    e += 1
        rec = dict(dim=dim, n=n, E=E, x=x, y=y, z=zc, glo_num=glo, node_degree=deg, p_mask=mask, G=G)
        out.append(rec)
        if write:
            _write(directory, p, N, rec)
    return out


def _geom(xe, D, w, dim):
    n = D.shape[0]
    if dim == 2:
        # xe[c][j, i]; d/dr acts on i (last axis), d/ds on j
        xr = xe[0] @ D.T; xs = D @ xe[0]
        yr = xe[1] @ D.T; ys = D @ xe[1]
        jac = xr * ys - xs * yr
        # jac * grad r = ( ys, -xs ), jac * grad s = ( -yr, xr )
        rx, ry = ys, -xs
        sx, sy = -yr, xr
        W = np.outer(w, w)
        G = np.zeros((6, n * n))
        G[0] = (W * (rx * rx + ry * ry) / jac).ravel()
        G[1] = (W * (sx * sx + sy * sy) / jac).ravel()
        G[2] = (W * (rx * sx + ry * sy) / jac).ravel()
        return G
    X = xe
    def dr(a): return np.einsum('im,kjm->kji', D, a)
    def ds(a): return np.einsum('jm,kmi->kji', D, a)
    def dt(a): return np.einsum('km,mji->kji', D, a)
    xr, xs, xt = dr(X[0]), ds(X[0]), dt(X[0])
    yr, ys, yt = dr(X[1]), ds(X[1]), dt(X[1])
    zr, zs, zt = dr(X[2]), ds(X[2]), dt(X[2])
    jac = xr * (ys * zt - yt * zs) - xs * (yr * zt - yt * zr) + xt * (yr * zs - ys * zr)
    # cofactors: jac * d r_a / d x_c
    rx = ys * zt - yt * zs; ry = xt * zs - xs * zt; rz = xs * yt - xt * ys
    sx = yt * zr - yr * zt; sy = xr * zt - xt * zr; sz = xt * yr - xr * yt
    tx = yr * zs - ys * zr; ty = xs * zr - xr * zs; tz = xr * ys - xs * yr
    W = w[:, None, None] * w[None, :, None] * w[None, None, :]
    G = np.zeros((6, n ** 3))
    G[0] = (W * (rx * rx + ry * ry + rz * rz) / jac).ravel()
    G[1] = (W * (sx * sx + sy * sy + sz * sz) / jac).ravel()
    G[2] = (W * (tx * tx + ty * ty + tz * tz) / jac).ravel()
    G[3] = (W * (rx * sx + ry * sy + rz * sz) / jac).ravel()
    G[4] = (W * (rx * tx + ry * ty + rz * tz) / jac).ravel()
    G[5] = (W * (sx * tx + sy * ty + sz * tz) / jac).ravel()
    return G


def _write(directory, p, N, r):
    d = os.path.join(directory, "lx1_%d" % (N + 1))
    os.makedirs(d, exist_ok=True)
    n = r["n"]
    with open(os.path.join(d, "size_%d.%d.dat" % (p, N)), "w") as f:
        f.write("%d %d %d %d %d\n" % (r["dim"], n, n, n if r["dim"] == 3 else 1, r["E"]))
    def w(name, a):
        a.tofile(os.path.join(d, "%s_%d.%d.dat" % (name, p, N)))
    w("x", r["x"]); w("y", r["y"]); w("z", r["z"])
    w("glo_num", r["glo_num"]); w("node_degree", r["node_degree"]); w("p_mask", r["p_mask"])
    for g in range(6):
        w("g_%d" % (g + 1), r["G"][g])


def ladder(N, r):
    """poly-degree ladder N, N-r, ..., clamped to 1 (subdomain.tpp:98-108)."""
    out = [N]
    while out[-1] > 1:
        out.append(max(out[-1] - r, 1))
    return out
