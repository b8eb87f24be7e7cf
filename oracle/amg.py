"""
ORACLE (test infrastructure, NOT product code).

Deterministic BoomerAMG-style setup used wherever the reference calls HYPRE (subdomain.tpp:1851-1858
and 3480-3549) and the Chebyshev-smoothed V-cycle of subdomain.tpp:3987-4159.

*** parity unpinned for the SETUP ***  HYPRE is a third-party dependency absent from /root/reference
(a "Master" checkout, Makefile:30-31; no pinned version), so its coarsening / interpolation / Chebyshev
setup cannot be reproduced bit for bit.  What is restated here is the *published algorithm family* HYPRE's
defaults select, made deterministic:
  * classical strength of connection, theta = 0.25 (negative couplings)            [Ruge-Stueben 1987]
  * coarsening, selectable (set_coarsening):
      "pmis"  PMIS (the parallel-independent-set half of HMIS)                       [De Sterck, Yang, Heys 2006]
              with hashed (not random) tie-breaking measures
      "hmis"  HMIS, coarsen type 10, the value the reference requests (subdomain.tpp:1853, default of 3480-3489): on one
              process it is the FIRST pass of the Ruge-Stueben colouring alone (oracle/amg_rs.c), ties first-in first-out
  * extended+i interpolation (interp type 6), truncated to P_max = 4 per row         [De Sterck, Falgout, Nolting, Yang 2008]
  * Galerkin coarse operators R A P with R = P^T; coarsest level solved exactly (hypre_GaussElimSolve)
  * Chebyshev smoother, hypre's scaled variant: ds = 1/sqrt(diag), spectrum of ds A ds estimated by
    10 CG/Lanczos steps from a hashed start vector, upper = 1.1 max_eig, lower = 0.3*(upper - min_eig) + min_eig,
    coefficients of the order-k residual polynomial (hypre's eig_est = 10, cheby_fraction = 0.3)
The V-cycle APPLY follows the reference line by line and is pinned through oracle/kernels.c.
The product implements the same setup in C++ (csrc/host/amg.hpp); tests compare the two hierarchies
(C/F splittings exactly, matrices to 1e-12).
"""
import ctypes as C
import os
import numpy as np
import scipy.sparse as sp

from . import capi as _c

P_ = _c.ptr
MARGIN = 1.0e-10   # relative margin that makes threshold / truncation decisions robust to rounding noise


def splitmix64(x):
    """vectorised splitmix64 on uint64 arrays (wraps mod 2^64)."""
    x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def hashed_unit(idx, salt):
    """deterministic pseudo-random numbers in [0,1): top 53 bits of splitmix64(idx + salt*2^32)."""
    with np.errstate(over="ignore"):
        h = splitmix64(np.asarray(idx, dtype=np.uint64) + (np.uint64(salt) << np.uint64(32)))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def strength(A, theta=0.25):
    """S[i,j] = 1 iff j != i and -a_ij >= theta * max_k(-a_ik) > 0 (row i strongly depends on j)."""
    A = A.tocsr()
    n = A.shape[0]
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    off = rows != A.indices
    neg = np.where(off, -A.data, 0.0)
    rowmax = np.zeros(n)
    np.maximum.at(rowmax, rows, neg)
    strong = off & (neg > 0.0) & (neg >= theta * rowmax[rows] * (1.0 - MARGIN))
    S = sp.csr_matrix((np.ones(int(strong.sum()), dtype=np.int8), (rows[strong], A.indices[strong])), shape=A.shape)
    S.sort_indices()
    return S


COARSENING = os.environ.get("PRFDD_AMG_COARSENING", "hmis").lower()   # default: what the reference requests from HYPRE


def set_coarsening(name):
    """ "pmis" or "hmis" for every Hierarchy built from now on"""
    global COARSENING
    assert name in ("pmis", "hmis")
    COARSENING = name


def rs_first_pass(S):
    """HMIS on one process = first Ruge-Stueben pass (oracle/amg_rs.c).  Returns cf: +1 C, -1 F."""
    S = S.tocsr(); S.sort_indices()
    n = S.shape[0]
    cf = np.zeros(n, dtype=np.int8)
    ptr, col = S.indptr.astype(np.int32), S.indices.astype(np.int32)
    _c.lib().oracle_rs_first_pass(C.c_int(n), P_(ptr), P_(col), P_(cf))
    return cf


def pmis(S, salt):
    """PMIS C/F splitting.  Returns cf: +1 C, -1 F.  Measures = (#points depending on i, hash(i))."""
    n = S.shape[0]
    St = S.T.tocsr()
    count = np.diff(St.indptr).astype(np.int64)            # how many points strongly depend on i
    with np.errstate(over="ignore"):
        h = splitmix64(np.arange(n, dtype=np.uint64) + (np.uint64(salt) << np.uint64(32)))
    order = np.lexsort((np.arange(n), h, count))             # ascending (count, hash, index)
    rank = np.empty(n, dtype=np.int64)
    rank[order] = np.arange(n)
    Gsym = (S + St).tocsr()
    Gsym.sort_indices()
    cf = np.zeros(n, dtype=np.int8)
    nodep = np.diff(S.indptr) == 0
    cf[(count == 0)] = -1                                   # nobody depends on it: F (hypre: measure < 1)
    # points that depend on nobody and that nobody depends on stay F with an empty interpolation row
    und = cf == 0
    Sr = S.tocsr()
    rows_g = np.repeat(np.arange(n), np.diff(Gsym.indptr))
    rows_s = np.repeat(np.arange(n), np.diff(Sr.indptr))
    while und.any():
        nb = np.where(und[Gsym.indices], rank[Gsym.indices], -1)
        mx = np.full(n, -1, dtype=np.int64)
        np.maximum.at(mx, rows_g, nb)
        newc = und & (rank > mx)
        cf[newc] = 1
        und &= ~newc
        # undecided points that strongly depend on a C point become F
        hit = np.zeros(n, dtype=bool)
        dep_c = cf[Sr.indices] == 1
        np.logical_or.at(hit, rows_s[dep_c], True)
        newf = und & hit
        cf[newf] = -1
        und &= ~newf
    del nodep
    return cf


def interp_extpi(A, S, cf, pmax=4):
    """extended+i interpolation (distance two), truncated to pmax entries per row, rescaled."""
    A = A.tocsr(); A.sort_indices()
    n = A.shape[0]
    isC = cf == 1
    cidx = np.full(n, -1, dtype=np.int64)
    cidx[isC] = np.arange(int(isC.sum()))
    nc = int(isC.sum())
    Fmask = sp.diags((~isC).astype(np.float64))
    Cmask = sp.diags(isC.astype(np.float64))
    Sf = S.astype(np.float64)
    Fs = (Fmask @ Sf @ Fmask).tocsr()      # strong F neighbours of F rows
    Cs = (Fmask @ Sf @ Cmask).tocsr()      # strong C neighbours of F rows
    Fs.eliminate_zeros(); Cs.eliminate_zeros()
    Chat = (Cs + Fs @ ((Sf @ Cmask).tocsr()))
    Chat.data[:] = 1.0
    Chat = Chat.tocsr(); Chat.sort_indices()
    diag = A.diagonal()
    Aoff = A - sp.diags(diag)
    Aoff = Aoff.tocsr(); Aoff.eliminate_zeros()
    # abar_kl = a_kl where sign(a_kl) != sign(a_kk), l != k
    rows = np.repeat(np.arange(n), np.diff(Aoff.indptr))
    keep = np.sign(Aoff.data) != np.sign(diag[rows])
    Abar = sp.csr_matrix((np.where(keep, Aoff.data, 0.0), Aoff.indices.copy(), Aoff.indptr.copy()), shape=A.shape)
    Abar.eliminate_zeros()
    ChatI = (Chat + Fmask).tocsr()          # C^_i U {i} for F rows
    ChatI.data[:] = 1.0
    # edge quantities on (i,k), k in F_i^s: d_ik = sum_{l in C^_i U {i}} abar_kl and a_ik
    Fs.sort_indices()
    erow = np.repeat(np.arange(n), np.diff(Fs.indptr)); ecol = Fs.indices
    M = (ChatI @ Abar.T).tocsr()
    if erow.size:
        dvals = np.asarray(M[erow, ecol]).ravel()
        avals = np.asarray(A[erow, ecol]).ravel()
    else:
        dvals = np.zeros(0); avals = np.zeros(0)
    good = dvals != 0.0     # edges with d_ik == 0 are lumped into the diagonal
    B = sp.csr_matrix((np.where(good, avals / np.where(good, dvals, 1.0), 0.0), ecol.copy(), Fs.indptr.copy()), shape=A.shape)
    lump = sp.csr_matrix((np.where(good, 0.0, avals), ecol.copy(), Fs.indptr.copy()), shape=A.shape)
    BA = (B @ Abar).tocsr()
    Num = (Chat.multiply(Aoff) + Chat.multiply(BA)).tocsr()
    # diagonal: a_ii + weak neighbours outside C^_i + sum_k B_ik abar_ki + lumped
    Sp = Sf.copy(); Sp.data[:] = 1.0
    Apat = Aoff.copy(); Apat.data[:] = 1.0
    weak_out = Aoff.multiply(Apat - Apat.multiply(Sp) - Apat.multiply(Chat) + Apat.multiply(Sp).multiply(Chat))
    # (pattern arithmetic: A-neighbours that are neither strong nor in C^_i; strong C neighbours are in C^_i)
    atil = diag + np.asarray(weak_out.sum(axis=1)).ravel() + BA.diagonal() + np.asarray(lump.sum(axis=1)).ravel()
    atil_safe = np.where(atil != 0.0, atil, 1.0)
    W = (sp.diags(-1.0 / atil_safe) @ Num).tocsr()
    W.sort_indices()
    # assemble P rows: C rows identity; F rows truncated
    indptr = [0]; indices = []; data = []
    Wp, Wi, Wd = W.indptr, W.indices, W.data
    for i in range(n):
        if isC[i]:
            indices.append(cidx[i]); data.append(1.0)
        else:
            s, e = Wp[i], Wp[i + 1]
            cols = Wi[s:e]; vals = Wd[s:e]
            nz = vals != 0.0
            cols, vals = cols[nz], vals[nz]
            if cols.size > pmax:
                total = vals.sum()
                absv = np.abs(vals)
                chosen = []
                avail = np.ones(cols.size, dtype=bool)
                for _ in range(pmax):
                    best = -1; bestv = -1.0
                    for t in range(cols.size):       # index order, strict-by-margin replacement
                        if avail[t] and absv[t] > bestv * (1.0 + MARGIN):
                            best, bestv = t, absv[t]
                    chosen.append(best); avail[best] = False
                chosen.sort()
                kept = vals[chosen].sum()
                scale = total / kept if kept != 0.0 else 1.0
                cols, vals = cols[chosen], vals[chosen] * scale
            for cc, vv in zip(cols, vals):
                indices.append(cidx[cc]); data.append(vv)
        indptr.append(len(indices))
    Pm = sp.csr_matrix((np.array(data), np.array(indices, dtype=np.int64), np.array(indptr)), shape=(n, nc))
    Pm.sort_indices()
    return Pm


def cheby_setup(A, order, salt, eig_iters=10, fraction=0.3):
    """hypre-style scaled Chebyshev: returns ds, coefs (length `order`), (max_eig, min_eig)."""
    A = A.tocsr()
    n = A.shape[0]
    d = A.diagonal()
    ds = 1.0 / np.sqrt(d)
    # CG / Lanczos on ds A ds  (hypre_ParCSRMaxEigEstimateCG)
    r = 2.0 * hashed_unit(np.arange(n), salt) - 1.0
    x = np.zeros(n)
    p = r.copy()
    rho = float(r @ r)
    alphas, betas = [], []
    for it in range(min(eig_iters, n)):
        q = ds * (A @ (ds * p))
        pq = float(p @ q)
        if pq == 0.0:
            break
        alpha = rho / pq
        x += alpha * p
        r = r - alpha * q
        rho_new = float(r @ r)
        beta = rho_new / rho
        alphas.append(alpha); betas.append(beta)
        if rho_new == 0.0:
            break
        p = r + beta * p
        rho = rho_new
    m = len(alphas)
    T = np.zeros((m, m))
    for i in range(m):
        T[i, i] = 1.0 / alphas[i] + (betas[i - 1] / alphas[i - 1] if i > 0 else 0.0)
        if i + 1 < m:
            T[i, i + 1] = T[i + 1, i] = np.sqrt(betas[i]) / alphas[i]
    ev = np.linalg.eigvalsh(T)
    max_eig, min_eig = float(ev[-1]), float(ev[0])
    upper = max_eig * 1.1
    lower = (upper - min_eig) * fraction + min_eig
    coefs = cheby_coefs(lower, upper, order)
    return ds, coefs, (max_eig, min_eig)


def cheby_coefs(lower, upper, order):
    """monomial coefficients c[0..order-1] of p with 1 - x p(x) = T_order((theta - x)/delta) / T_order(theta/delta)."""
    theta = 0.5 * (upper + lower)
    delta = 0.5 * (upper - lower)
    # T_k((theta - x)/delta) as a polynomial in x, by the three-term recurrence
    t0 = np.array([1.0])
    t1 = np.array([theta / delta, -1.0 / delta])
    for _ in range(order - 1):
        t2 = 2.0 * np.convolve(np.array([theta / delta, -1.0 / delta]), t1)
        t2[:t0.size] -= t0
        t0, t1 = t1, t2
    tk = t1
    # T_order(theta/delta) = value at x = 0
    scale = tk[0]
    q = tk / scale              # q(x) = 1 - x p(x)
    p = -q[1:]                  # p(x) = (1 - q(x)) / x
    return np.ascontiguousarray(p[:order])


def _csr_arrays(M, dtype=np.float64):
    """int32 / value-type views of a scipy CSR matrix for the C kernels, converted once per matrix (not per call)"""
    key = "_oracle_arrays_%s" % np.dtype(dtype).name
    c = getattr(M, key, None)
    if c is None:
        c = (np.ascontiguousarray(M.indptr, dtype=np.int32), np.ascontiguousarray(M.indices, dtype=np.int32), np.ascontiguousarray(M.data, dtype=dtype))
        setattr(M, key, c)
    return c


class Level:
    pass


class Hierarchy:
    """levels[l]: A, P (l -> l+1 interpolation, rows = fine), R = P^T, ds, coefs, cf."""

    def __init__(self, A0, cheby_order=2, max_coarse=9, theta=0.25, pmax=4, max_levels=25, with_smoother=True, dtype=np.float64):
        # dtype float32: the reference's `Float float` (AMG/config.hpp:4) -- HYPRE's set-up stays double, the extraction into the
        # amg:: classes casts matrices, ds and the Chebyshev coefficients to Float (subdomain.tpp:3491-3549), the cycle runs in Float
        self.dtype = np.dtype(dtype).type
        self.levels = []
        A = A0.tocsr().astype(np.float64); A.sort_indices()
        l = 0
        while True:
            L = Level()
            L.A = A
            L.n = A.shape[0]
            if with_smoother and L.n > 0:
                L.ds, L.coefs, L.eigs = cheby_setup(A, cheby_order, salt=1000 + l)
            self.levels.append(L)
            if L.n <= max_coarse or l + 1 >= max_levels:
                break
            S = strength(A, theta)
            cf = rs_first_pass(S) if COARSENING.startswith("h") else pmis(S, salt=l)
            nc = int((cf == 1).sum())
            if nc == 0 or nc == L.n:
                break
            Pm = interp_extpi(A, S, cf, pmax)
            L.S, L.cf, L.P = S, cf, Pm
            L.R = Pm.T.tocsr(); L.R.sort_indices()
            A = (L.R @ A @ Pm).tocsr(); A.sort_indices()
            l += 1
        last = self.levels[-1]
        last.Ainv = np.linalg.inv(last.A.toarray()) if last.n > 0 else np.zeros((0, 0))
        self.cheby_order = cheby_order
        if self.dtype is np.float32:
            for L in self.levels:
                if hasattr(L, "ds"):
                    L.ds32 = L.ds.astype(np.float32)
            last.Ainv32 = last.Ainv.astype(np.float32)

    @property
    def num_levels(self):
        return len(self.levels)

    # ---- V-cycle: subdomain.tpp:4015-4139 (down leg, coarse solve, up leg) ------------------
    def _smooth(self, L, u, f, u_is_zero):
        """scaled_residual + polynomial_evaluation x (order-1) + update_field (subdomain.tpp:3652-3657), host branches."""
        f32 = self.dtype is np.float32
        K = _c.lib32() if f32 else _c.lib()
        cs = C.c_float if f32 else C.c_double
        ds = L.ds32 if f32 else L.ds
        A = L.A
        ptr, col, val = _csr_arrays(A, self.dtype)
        n = L.n
        r = np.zeros(n, self.dtype); w = np.zeros(n, self.dtype); v = np.zeros(n, self.dtype)
        k = self.cheby_order
        K.o_scaled_residual(P_(r), P_(w), P_(ptr), P_(col), P_(val), P_(u), P_(f), P_(ds), cs(L.coefs[k - 1]), C.c_int(n))
        for pidx in range(k - 2, -1, -1):
            K.o_polynomial_evaluation(P_(w), P_(v), P_(ptr), P_(col), P_(val), P_(r), P_(ds), cs(L.coefs[pidx]), C.c_int(n))
        K.o_update_field(P_(u), P_(w), P_(ds), C.c_int(n))

    def _matvec(self, M, y, x, alpha, beta):
        f32 = self.dtype is np.float32
        K = _c.lib32() if f32 else _c.lib()
        cs = C.c_float if f32 else C.c_double
        ptr, col, val = _csr_arrays(M, self.dtype)
        K.o_amg_matvec(P_(y), P_(ptr), P_(col), P_(val), P_(x), cs(alpha), cs(beta), C.c_int(M.shape[0]))

    def vcycle(self, f0, num_vcycles=1):
        nl = self.num_levels
        f = [None] * nl; u = [None] * nl
        dt = self.dtype
        f[0] = np.ascontiguousarray(f0, dtype=dt).copy()
        u[0] = np.zeros(self.levels[0].n, dt)
        for _ in range(num_vcycles):
            for l in range(nl - 1):
                L = self.levels[l]
                if l > 0:
                    u[l] = np.zeros(L.n, dt)
                self._smooth(L, u[l], f[l], True)
                v = f[l].copy()
                self._matvec(L.A, v, u[l], -1.0, 1.0)            # residual (tpp:3660-3661)
                f[l + 1] = np.zeros(self.levels[l + 1].n, dt)
                self._matvec(L.R, f[l + 1], v, 1.0, 0.0)          # restrict (tpp:3666-3672)
            last = self.levels[-1]
            u[nl - 1] = (last.Ainv32 if dt is np.float32 else last.Ainv) @ f[nl - 1]   # hypre_GaussElimSolve (tpp:4084)
            for l in range(nl - 1, 0, -1):
                L = self.levels[l - 1]
                self._matvec(L.P, u[l - 1], u[l], 1.0, 1.0)       # coarse-grid correction (tpp:4098)
                self._smooth(L, u[l - 1], f[l - 1], False)
        return u[0]
