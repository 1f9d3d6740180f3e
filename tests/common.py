"""Shared helpers for the tests: instance builders and KKT-level fixtures."""
import numpy as np
import scipy.sparse as sp


def small_instances(cb):
    pr = cb.problems
    return {
        "C1": lambda: pr.c1_random_qp(),
        "C1s": lambda: pr.c1_random_qp(n=120, m=260, seed=7),
        "C2s": lambda: pr.c2_portfolio(n=3000),
        "C3s": lambda: pr.c3_socp(n=3000, ncones=60),
        "C4s": lambda: pr.c4_sdp(ncones=6, side=8, n=200, vars_per_cone=50),
        "C4m": lambda: pr.c4_sdp(ncones=4, side=22, n=300, vars_per_cone=120),
        "C5s": lambda: pr.c5_block_angular(nblocks=4, grid=12, nlink=10, link_nnz=8),
    }


def kkt_fixture(cb, gen, seed=0):
    """Assembled KKT (triu) with quasidefinite random values in the cone blocks."""
    from clarabel_jl_b200 import kkt_assembly as ka
    P, q, A, b, K = gen()
    st = cb.Settings()
    data = cb.problemdata.ProblemData(P, q, A, b, K, st)
    cones = cb.CompositeCone(data.cones)
    KKT, mp = ka.assemble_kkt_matrix(data.P, data.A, cones)
    N = KKT.shape[0]
    rng = np.random.default_rng(seed)
    Ds = ka.fill_Dsigns(data.m, data.n, cones.p)
    KKT.data[mp.diag_full] += np.where(Ds > 0, 1.0 + rng.random(N), -1.0 - rng.random(N))
    KKT.data[mp.soc_u] = 0.1 * rng.standard_normal(len(mp.soc_u))
    KKT.data[mp.soc_v] = 0.1 * rng.standard_normal(len(mp.soc_v))
    # dense (non-diagonal) cone blocks: make them negative definite-ish via small off-diagonals
    diag_set = set(mp.diag_full.tolist())
    off = np.array([k for k in mp.Hsblocks if k not in diag_set], dtype=np.int64)
    if len(off):
        KKT.data[off] = 0.02 * rng.standard_normal(len(off))
    return KKT, mp, Ds, data, cones


def sym_full(K):
    return (K + sp.triu(K, 1).T).tocsc()
