"""Synthetic instance generators for the BASELINE configs C1..C5.

Every generator is seeded and returns (P, q, A, b, cones) in the reference's `Solver(P,q,A,b,
cones)` convention (P upper-triangular CSC, A CSC, cones in reference API names), plus a
`to_reference_json` writer that emits the reference's own fixture schema (src/json.jl:118-156)
so the same instances can be replayed through Clarabel.jl by anyone with Julia.

All instances are built strictly feasible: b = A x0 + s0 with s0 in the interior of the cone.
"""
import json
import numpy as np
import scipy.sparse as sp

from .cones import (NonnegativeConeT, SecondOrderConeT, PSDTriangleConeT, ZeroConeT,  # noqa: F401
                    triangular_number)


def _triu(P):
    return sp.triu(sp.csc_matrix(P), format="csc")


def c1_random_qp(n=1000, m=2000, seed=0):
    """C1: random sparse QP, Nonneg cone only (BASELINE.json configs[0])."""
    rng = np.random.default_rng(seed)
    # P = diag(d) + M'M with 2-3 nonzeros per row of M  -> nnz(triu P) ~ 3n
    k = n
    rows = np.repeat(np.arange(k), 3)
    cols = rng.integers(0, n, size=3 * k)
    keep = rng.random(3 * k) < 0.85
    M = sp.csc_matrix((rng.standard_normal(3 * k)[keep], (rows[keep], cols[keep])), shape=(k, n))
    P = _triu(M.T @ M + sp.diags(0.1 + rng.random(n)))
    # A: ~3.5 nnz per row
    nper = rng.integers(3, 5, size=m)
    r = np.repeat(np.arange(m), nper)
    c = rng.integers(0, n, size=len(r))
    A = sp.csc_matrix((rng.standard_normal(len(r)), (r, c)), shape=(m, n))
    A.sum_duplicates()
    x0 = rng.standard_normal(n)
    b = A @ x0 + np.abs(rng.standard_normal(m)) + 0.1
    q = rng.standard_normal(n)
    return P, q, A, b, [NonnegativeConeT(m)]


def c2_portfolio(n=100000, seed=1, nnz_per_asset=25):
    """C2: factor-model portfolio QP, lifted: vars (x in R^a, y in R^k), P = blkdiag(diag(d), I_k),
    rows +-(F'x - y) <= 0, +-(1'x - 1) <= 0, x >= 0, x <= u.  All NonnegativeConeT."""
    rng = np.random.default_rng(seed)
    a = int(round(n / 1.01)); k = n - a
    r = np.repeat(np.arange(a), nnz_per_asset)
    c = rng.integers(0, k, size=len(r))
    F = sp.csc_matrix((rng.standard_normal(len(r)) / np.sqrt(nnz_per_asset), (r, c)), shape=(a, k))
    F.sum_duplicates()
    d = 0.05 + rng.random(a)
    gamma = 1.0
    P = sp.diags(np.concatenate([gamma * d, gamma * np.ones(k)]), format="csc")
    mu = 0.1 * rng.standard_normal(a)
    q = np.concatenate([-mu, np.zeros(k)])
    Ft = F.T.tocsc()
    Ik = sp.identity(k, format="csc"); Ia = sp.identity(a, format="csc")
    ones = sp.csc_matrix(np.ones((1, a)))
    Zk1 = sp.csc_matrix((1, k)); Zak = sp.csc_matrix((a, k))
    A = sp.vstack([sp.hstack([Ft, -Ik]), sp.hstack([-Ft, Ik]),
                   sp.hstack([ones, Zk1]), sp.hstack([-ones, Zk1]),
                   sp.hstack([-Ia, Zak]), sp.hstack([Ia, Zak])]).tocsc()
    u = np.full(a, 10.0 / a)
    slack = 1e-3
    b = np.concatenate([np.full(k, slack), np.full(k, slack), [1.0 + slack], [-1.0 + slack],
                        np.zeros(a), u])
    m = A.shape[0]
    return P, q, A, b, [NonnegativeConeT(m)]


def c3_socp(n=500000, ncones=10000, nn_rows=None, seed=2, window=24, dmin=4, dmax=64):
    """C3: `ncones` second-order cones of dim U{dmin..dmax} (dim<=4 take the dense block, the
    rest the diagonal + rank-2 expansion) mixed with Nonneg bound-type rows; rows have ~4 nnz
    with locality (row touches variables inside a sliding window) so fill stays bounded."""
    rng = np.random.default_rng(seed)
    nn_rows = n if nn_rows is None else nn_rows
    dims = rng.integers(dmin, dmax + 1, size=ncones)
    msoc = int(dims.sum())
    m = nn_rows + msoc
    # row -> window centre: NN rows sweep the variables once, SOC rows sweep them once
    centre = np.concatenate([np.linspace(0, n - 1, nn_rows), np.linspace(0, n - 1, msoc)])
    nper = 4
    r = np.repeat(np.arange(m), nper)
    off = rng.integers(-window, window + 1, size=len(r))
    c = np.clip(np.round(np.repeat(centre, nper)).astype(np.int64) + off, 0, n - 1)
    A = sp.csc_matrix((rng.standard_normal(len(r)), (r, c)), shape=(m, n))
    A.sum_duplicates()
    P = sp.diags(0.5 + rng.random(n), format="csc")
    x0 = rng.standard_normal(n)
    s0 = np.empty(m)
    s0[:nn_rows] = 0.1 + np.abs(rng.standard_normal(nn_rows))
    ptr = nn_rows + np.concatenate([[0], np.cumsum(dims)])
    tail = rng.standard_normal(msoc)
    s0[nn_rows:] = tail
    heads = ptr[:-1]
    sq = np.add.reduceat(tail * tail, heads - nn_rows) - tail[heads - nn_rows] ** 2
    s0[heads] = np.sqrt(sq) + 0.5 + rng.random(ncones)
    b = A @ x0 + s0
    q = rng.standard_normal(n)
    cones = [NonnegativeConeT(nn_rows)] + [SecondOrderConeT(int(d)) for d in dims]
    return P, q, A, b, cones


def c4_sdp(ncones=200, side=50, n=50000, seed=3, vars_per_cone=400, nnz_per_row=2):
    """C4: SDP already in decomposed form: `ncones` PSDTriangleConeT(side) blocks; cone j reads a
    window of `vars_per_cone` variables that overlaps its neighbours' windows (the coupling
    chordal decomposition would create through shared x-columns), a few nnz per row."""
    rng = np.random.default_rng(seed)
    ne = triangular_number(side)
    m = ncones * ne
    stride = max(1, (n - vars_per_cone) // max(1, ncones - 1))
    rows, cols, vals = [], [], []
    for j in range(ncones):
        lo = min(j * stride, n - vars_per_cone)
        r = np.repeat(np.arange(j * ne, (j + 1) * ne), nnz_per_row)
        c = lo + rng.integers(0, vars_per_cone, size=len(r))
        rows.append(r); cols.append(c); vals.append(rng.standard_normal(len(r)))
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(m, n))
    A.sum_duplicates()
    P = sp.diags(0.5 + rng.random(n), format="csc")
    x0 = 0.1 * rng.standard_normal(n)
    # s0 = svec(S0), S0 = G G'/side + I  (strictly PD)
    ti, tj = np.tril_indices(side)
    s0 = np.empty(m)
    for j in range(ncones):
        G = rng.standard_normal((side, side))
        S = G @ G.T / side + np.eye(side)
        v = S[tj, ti] * np.where(ti == tj, 1.0, np.sqrt(2.0))
        s0[j * ne:(j + 1) * ne] = v
    b = A @ x0 + s0
    q = rng.standard_normal(n)
    return P, q, A, b, [PSDTriangleConeT(side)] * ncones


def c5_block_angular(nblocks=64, grid=125, nlink=2000, link_nnz=64, seed=4):
    """C5: block-angular large sparse QP.  `nblocks` diagonal blocks, each a grid x grid 2-D
    lattice sub-QP (P = lattice Laplacian + diag; two stencil rows per variable), coupled by
    `nlink` linking rows that each touch `link_nnz` variables spread over all blocks.
    n = nblocks*grid^2, m = 2n + nlink.  Default: n = 1e6, m = 2.002e6."""
    rng = np.random.default_rng(seed)
    g = grid; nb = g * g; n = nblocks * nb
    idx = np.arange(nb).reshape(g, g)
    # lattice edges inside a block
    e_r = np.concatenate([idx[:, :-1].ravel(), idx[:-1, :].ravel()])
    e_c = np.concatenate([idx[:, 1:].ravel(), idx[1:, :].ravel()])
    ne = len(e_r)
    Pr, Pc, Pv = [], [], []
    diag = np.zeros(n)
    for bl in range(nblocks):
        w = 0.2 + rng.random(ne)
        Pr.append(bl * nb + e_r); Pc.append(bl * nb + e_c); Pv.append(-w)
        np.add.at(diag, bl * nb + e_r, w); np.add.at(diag, bl * nb + e_c, w)
    diag += 0.1 + rng.random(n)
    P = sp.csc_matrix((np.concatenate(Pv + [diag]),
                       (np.concatenate(Pr + [np.arange(n)]), np.concatenate(Pc + [np.arange(n)]))),
                      shape=(n, n))
    P = _triu(P)
    # two rows per variable: 5-point stencil (centre + 4 neighbours, clipped at block borders)
    ii, jj = np.divmod(np.arange(nb), g)
    nbrs = [np.arange(nb),
            np.where(ii > 0, np.arange(nb) - g, -1), np.where(ii < g - 1, np.arange(nb) + g, -1),
            np.where(jj > 0, np.arange(nb) - 1, -1), np.where(jj < g - 1, np.arange(nb) + 1, -1)]
    Ar, Ac, Av = [], [], []
    for rep in range(2):
        for bl in range(nblocks):
            for t, nb_t in enumerate(nbrs):
                sel = nb_t >= 0
                rr = rep * n + bl * nb + np.arange(nb)[sel]
                Ar.append(rr); Ac.append(bl * nb + nb_t[sel])
                v = rng.standard_normal(int(sel.sum())) * (1.0 if t == 0 else 0.3)
                Av.append(v)
    # linking rows
    lr = np.repeat(2 * n + np.arange(nlink), link_nnz)
    lc = rng.integers(0, n, size=len(lr))
    Ar.append(lr); Ac.append(lc); Av.append(rng.standard_normal(len(lr)) / np.sqrt(link_nnz))
    m = 2 * n + nlink
    A = sp.csc_matrix((np.concatenate(Av), (np.concatenate(Ar), np.concatenate(Ac))), shape=(m, n))
    A.sum_duplicates()
    x0 = rng.standard_normal(n)
    b = A @ x0 + 0.1 + np.abs(rng.standard_normal(m))
    q = rng.standard_normal(n)
    return P, q, A, b, [NonnegativeConeT(m)]


CONFIGS = {
    "C1": lambda: c1_random_qp(),
    "C2": lambda: c2_portfolio(),
    "C3": lambda: c3_socp(),
    "C4": lambda: c4_sdp(),
    "C5": lambda: c5_block_angular(),
}


def socp_lasso(nfeat=8, seed=12345):
    """Lasso regression as an SOCP with the shape of the reference's test instance
    (test/OptTests/socp-lasso.jl:6-56: 2*nfeat + 3 + m variables, rows NN(m + 2), NN(2 nfeat),
    SOC(m + 2) with m = 50 nfeat).  The reference draws its data from Julia's MersenneTwister
    stream, which cannot be reproduced here: same construction, numpy random data.
        variables (t, v, u, w1, w2, y):  min t + mu sum(u) + |x|^2/2  (P = I as in the reference)"""
    rng = np.random.default_rng(seed)
    n = nfeat; m = 50 * n
    F = rng.random((m, n))
    vtrue = np.where(rng.random(n) < 0.1, rng.random(n), 0.0)
    bb = F @ vtrue + 0.1 * rng.random(m)
    mu = 0.1 * np.abs(F.T @ bb).max()
    Z = np.zeros; I = np.eye
    A1 = -np.block([[np.ones((1, 1)), Z((1, 2 * n + 1)), np.ones((1, 1)), Z((1, m))],
                    [-np.ones((1, 1)), Z((1, 2 * n)), np.ones((1, 1)), Z((1, m + 1))],
                    [Z((m, 1)), -2 * F, Z((m, n + 2)), I(m)]])
    A2 = -np.block([[Z((n, 1)), I(n), -I(n), Z((n, m + 2))],
                    [Z((n, 1)), -I(n), -I(n), Z((n, m + 2))]])
    A3 = -np.block([[Z((1, 2 * n + 1)), -np.ones((1, 1)), Z((1, m + 1))],
                    [Z((1, 2 * n + 2)), -np.ones((1, 1)), Z((1, m))],
                    [Z((m, 2 * n + 3)), -I(m)]])
    b = np.concatenate([[1.0, 1.0], -2 * bb, np.zeros(2 * n), np.zeros(m + 2)])
    c = np.concatenate([[1.0], np.zeros(n), mu * np.ones(n), np.zeros(m + 2)])
    P = sp.identity(len(c), format="csc")
    A = sp.csc_matrix(np.vstack([A1, A2, A3]))
    cones = [NonnegativeConeT(m + 2), NonnegativeConeT(2 * n), SecondOrderConeT(m + 2)]
    return P, c, A, b, cones


def _cone_to_json(c):
    """lower(cone) (src/json.jl:138-151)."""
    if c[0] == "PowerConeT":
        return {c[0]: c[2]}
    if c[0] == "ExponentialConeT":
        return {c[0]: []}
    if c[0] == "GenPowerConeT":
        return {c[0]: [list(c[2]), c[3]]}
    return {c[0]: c[1]}


def _cone_from_json(d):
    """parse(dict, SupportedCone) (src/json.jl:196-216)."""
    from . import cones as C
    (key, val), = d.items()
    if key == "GenPowerConeT":
        return C.GenPowerConeT(val[0], int(val[1]))
    if key == "ExponentialConeT":
        return C.ExponentialConeT()
    if key == "PowerConeT":
        return C.PowerConeT(float(val))
    return getattr(C, key)(int(val))


def to_reference_json(path, P, q, A, b, cones, settings=None):
    """Write the instance in the reference's JSON schema (src/json.jl:118-156): CSC matrices as
    {m,n,colptr,rowval,nzval} with 0-based indices, cones as [{"NonnegativeConeT": k}, ...]."""
    def mat(M):
        M = sp.csc_matrix(M); M.sort_indices()
        return dict(m=int(M.shape[0]), n=int(M.shape[1]), colptr=M.indptr.tolist(),
                    rowval=M.indices.tolist(), nzval=M.data.tolist())
    doc = dict(settings=settings or {}, P=mat(_triu(P)), q=np.asarray(q, dtype=float).tolist(), A=mat(A),
               b=np.asarray(b, dtype=float).tolist(), cones=[_cone_to_json(c) for c in cones])
    with open(path, "w") as f:
        json.dump(doc, f)


def from_reference_json(path):
    """Read a fixture written by the reference's `save_to_file` or by `to_reference_json`
    (load_from_file, src/json.jl:58-80).  Returns (P, q, A, b, cones, settings_dict)."""
    with open(path) as f:
        doc = json.load(f)

    def mat(d):
        return sp.csc_matrix((np.asarray(d["nzval"], dtype=float), np.asarray(d["rowval"], dtype=np.int64),
                              np.asarray(d["colptr"], dtype=np.int64)), shape=(int(d["m"]), int(d["n"])))
    return (mat(doc["P"]), np.asarray(doc["q"], dtype=float), mat(doc["A"]), np.asarray(doc["b"], dtype=float),
            [_cone_from_json(c) for c in doc["cones"]], dict(doc.get("settings") or {}))
