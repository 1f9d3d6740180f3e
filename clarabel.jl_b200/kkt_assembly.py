"""One-time symbolic build of the upper-triangular KKT matrix and its value maps.

Mirror of `_assemble_kkt_matrix` (src/kktsolvers/direct-ldl/directldl_kkt_assembly.jl:15-50,
colcounts :52-101, fill :104-175), the CSC helpers (src/utils/csc_assembly.jl) and
`LDLDataMap` / `SOCExpansionMap` (src/kktsolvers/direct-ldl/directldl_datamaps.jl:8-22,170-214),
for shape = :triu (what the QDLDL and B200 engines request).

    K = [ triu(P)+0*I    A'        .   ]      columns 0..n-1      : P column, then the diagonal
        [      .       -Hs(0s)   [v u] ]      columns n..n+m-1    : A' entries, then the cone block
        [      .          .        D   ]      columns n+m..N-1    : v, u per sparse SOC, then D

Every column is filled in ascending row order with the diagonal LAST (directldl_kkt_assembly.jl
:161-165), so the result is the canonical sorted CSC; destinations are computed in closed form
instead of the reference's running column pointers (same result, no O(nnz) Python loop).
All indices here are 0-based (the reference is 1-based).
"""
import numpy as np
import scipy.sparse as sp


class LDLDataMap:
    """0-based value maps into K.data (directldl_datamaps.jl:170-214)."""
    def __init__(self):
        self.P = self.A = self.Hsblocks = None
        self.soc_u = self.soc_v = self.soc_D = None      # concatenated over sparse SOCs
        self.gp_q = self.gp_r = self.gp_p = self.gp_D = None   # ... over generalised power cones
        self.diagP = self.diag_full = None


def assemble_kkt_matrix(P, A, cones):
    """Returns (K: scipy csc upper-triangular with structural zeros kept, map: LDLDataMap)."""
    m, n = A.shape
    p = cones.p
    N = n + m + p
    Pp, Pi = P.indptr.astype(np.int64), P.indices.astype(np.int64)
    Ap, Ai = A.indptr.astype(np.int64), A.indices.astype(np.int64)
    nnzP, nnzA = len(Pi), len(Ai)

    # --- column counts (directldl_kkt_assembly.jl:52-101)
    cc = np.zeros(N, dtype=np.int64)
    pcnt = np.diff(Pp)
    last_is_diag = np.zeros(n, dtype=bool)
    ne = pcnt > 0
    last_is_diag[ne] = Pi[Pp[1:][ne] - 1] == np.nonzero(ne)[0]
    missing = ~last_is_diag                                 # _csc_colcount_missing_diag
    cc[:n] = pcnt + missing
    Arowcnt = np.bincount(Ai, minlength=m).astype(np.int64)  # A' : one per entry of row
    cc[n:n + m] = Arowcnt
    rc, rb = cones.rng_cones, cones.rng_blocks
    ncone = len(cones.specs)
    # cone blocks
    blk_in_col = np.zeros(m, dtype=np.int64)                # Hs entries in column n+i
    for i in range(ncone):
        a, b = rc[i], rc[i + 1]
        if cones.Hs_is_diagonal[i]:
            blk_in_col[a:b] = 1
        else:
            blk_in_col[a:b] = np.arange(1, b - a + 1)       # dense triu triangle
    cc[n:n + m] += blk_in_col
    # expansion columns follow in cone order (directldl_kkt_assembly.jl:86-99): v,u per sparse SOC,
    # q,r,p per generalised power cone; pcol[i] = first expansion column of cone i
    pdims = getattr(cones, "pdims", np.where(cones.is_sparse, 2, 0))
    pcol = n + m + np.concatenate([[0], np.cumsum(pdims)]).astype(np.int64)
    sparse_cones = np.nonzero(cones.is_sparse)[0]
    genpow_cones = np.nonzero(getattr(cones, "is_genpow", np.zeros(ncone, dtype=bool)))[0]
    for i in sparse_cones:
        dim = rc[i + 1] - rc[i]
        cc[pcol[i]] = dim + 1                                # v column + D
        cc[pcol[i] + 1] = dim + 1                            # u column + D
    for i in genpow_cones:
        dim = int(rc[i + 1] - rc[i]); dim2 = cones.specs[i][3]; dim1 = dim - dim2
        cc[pcol[i]] = dim1 + 1                               # q column + D
        cc[pcol[i] + 1] = dim2 + 1                           # r column + D
        cc[pcol[i] + 2] = dim + 1                            # p column + D
    colptr = np.concatenate([[0], np.cumsum(cc)]).astype(np.int64)
    nnzK = int(colptr[-1])
    rowval = np.empty(nnzK, dtype=np.int64)
    nzval = np.zeros(nnzK, dtype=np.float64)

    mp = LDLDataMap()
    # --- P block (csc_assembly.jl:125-143 with shape :N) then missing diagonal (:207-220)
    Pcol = np.repeat(np.arange(n, dtype=np.int64), pcnt)
    mp.P = colptr[Pcol] + (np.arange(nnzP, dtype=np.int64) - Pp[Pcol])
    rowval[mp.P] = Pi
    nzval[mp.P] = P.data
    md = np.nonzero(missing)[0]
    dest = colptr[md + 1] - 1
    rowval[dest] = md
    # --- A' block: entry j of A (row r, column c) goes to K[c, n+r]; within K column n+r the
    # entries appear in ascending c because the fill iterates A's columns in order.
    Acol = np.repeat(np.arange(n, dtype=np.int64), np.diff(Ap))
    order = np.argsort(Ai, kind="stable")                   # groups by row, ascending column
    rowptrA = np.concatenate([[0], np.cumsum(Arowcnt)])
    rank = np.empty(nnzA, dtype=np.int64)
    rank[order] = np.arange(nnzA, dtype=np.int64) - rowptrA[Ai[order]]
    mp.A = colptr[n + Ai] + rank
    rowval[mp.A] = Acol
    nzval[mp.A] = A.data
    # --- cone blocks (structural zeros), packed-triu order for dense blocks
    Hs = np.empty(int(rb[-1]), dtype=np.int64)
    base = colptr[n:n + m] + Arowcnt                        # first Hs slot of each column
    for i in range(ncone):
        a, b = int(rc[i]), int(rc[i + 1])
        dim = b - a
        if cones.Hs_is_diagonal[i]:
            dst = base[a:b]
            Hs[rb[i]:rb[i + 1]] = dst
            rowval[dst] = n + np.arange(a, b)
        else:
            ti, tj = np.tril_indices(dim)                   # (col=ti, row=tj): col-major upper
            dst = base[a + ti] + tj
            Hs[rb[i]:rb[i + 1]] = dst
            rowval[dst] = n + a + tj
    mp.Hsblocks = Hs
    # --- sparse SOC expansion columns: v first, then u (directldl_datamaps.jl:42-59)
    us, vs, Ds = [], [], []
    for i in sparse_cones:
        a, b = int(rc[i]), int(rc[i + 1])
        dim = b - a
        cv, cu = int(pcol[i]), int(pcol[i]) + 1
        dv = colptr[cv] + np.arange(dim); du = colptr[cu] + np.arange(dim)
        rowval[dv] = n + np.arange(a, b); rowval[du] = n + np.arange(a, b)
        rowval[colptr[cv] + dim] = cv; rowval[colptr[cu] + dim] = cu
        vs.append(dv); us.append(du)
        Ds.append(np.array([colptr[cv] + dim, colptr[cu] + dim], dtype=np.int64))
    z = np.zeros(0, dtype=np.int64)
    mp.soc_u = np.concatenate(us) if us else z
    mp.soc_v = np.concatenate(vs) if vs else z
    mp.soc_D = np.concatenate(Ds) if Ds else z
    # --- generalised power cone expansion columns: q (rows of u), r (rows of w), p (all rows)
    # (directldl_datamaps.jl:124-144)
    qs, rs, ps, gD = [], [], [], []
    for i in genpow_cones:
        a, b = int(rc[i]), int(rc[i + 1])
        dim = b - a; dim2 = cones.specs[i][3]; dim1 = dim - dim2
        cq, cr, cp_ = int(pcol[i]), int(pcol[i]) + 1, int(pcol[i]) + 2
        dq = colptr[cq] + np.arange(dim1); dr = colptr[cr] + np.arange(dim2); dp = colptr[cp_] + np.arange(dim)
        rowval[dq] = n + a + np.arange(dim1)
        rowval[dr] = n + a + dim1 + np.arange(dim2)
        rowval[dp] = n + a + np.arange(dim)
        for c_, cnt in ((cq, dim1), (cr, dim2), (cp_, dim)):
            rowval[colptr[c_] + cnt] = c_
        qs.append(dq); rs.append(dr); ps.append(dp)
        gD.append(np.array([colptr[cq] + dim1, colptr[cr] + dim2, colptr[cp_] + dim], dtype=np.int64))
    mp.gp_q = np.concatenate(qs) if qs else z
    mp.gp_r = np.concatenate(rs) if rs else z
    mp.gp_p = np.concatenate(ps) if ps else z
    mp.gp_D = np.concatenate(gD) if gD else z
    mp.diag_full = colptr[1:] - 1
    mp.diagP = colptr[1:n + 1] - 1
    K = sp.csc_matrix((nzval, rowval, colptr), shape=(N, N))
    return K, mp


def fill_Dsigns(m, n, p, cones=None):
    """_fill_Dsigns! (kktsolver_directldl.jl:112-126); expansion signs are (-1,+1) per sparse SOC
    and (-1,-1,+1) per generalised power cone (directldl_datamaps.jl:21,98), in cone order."""
    D = np.ones(n + m + p, dtype=np.int64)
    D[n:n + m] = -1
    if cones is None or not getattr(cones, "is_genpow", np.zeros(1, dtype=bool)).any():
        D[n + m::2] = -1
        return D
    k = n + m
    for pd in cones.pdims:
        if pd == 2:
            D[k] = -1
        elif pd == 3:
            D[k] = -1; D[k + 1] = -1
        k += int(pd)
    return D
