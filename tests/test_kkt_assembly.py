"""CPU: KKT assembly + LDLDataMap against an independent scipy construction and the
reference's structural invariants (directldl_kkt_assembly.jl:34-41 nnz formula, :161-165
diagonal-last, map disjointness directldl_datamaps.jl:177-180)."""
import numpy as np
import scipy.sparse as sp
import pytest
from common import small_instances


@pytest.mark.parametrize("name", ["C1s", "C2s", "C3s", "C4s", "C5s"])
def test_assembly_matches_scipy_bmat(cb, name):
    from clarabel_jl_b200 import kkt_assembly as ka
    P, q, A, b, K = small_instances(cb)[name]()
    data = cb.problemdata.ProblemData(P, q, A, b, K, cb.Settings())
    cones = cb.CompositeCone(data.cones)
    KKT, mp = ka.assemble_kkt_matrix(data.P, data.A, cones)
    n, m, p = data.n, data.m, cones.p
    N = n + m + p
    assert KKT.shape == (N, N)
    # canonical CSC: sorted rows, diagonal last in every column, upper triangular
    for j in range(N):
        rows = KKT.indices[KKT.indptr[j]:KKT.indptr[j + 1]]
        assert np.all(np.diff(rows) > 0) and rows[-1] == j
    assert np.array_equal(mp.diag_full, KKT.indptr[1:] - 1)
    # nnz formula
    nnz_diagP = int((data.P.diagonal() != 0).sum()) if data.P.nnz else 0
    nnz_diagP = sum(1 for j in range(n) if data.P.indptr[j + 1] > data.P.indptr[j]
                    and data.P.indices[data.P.indptr[j + 1] - 1] == j)
    nnz_vec = len(mp.soc_u) + len(mp.soc_v)
    assert KKT.nnz == data.P.nnz + n - nnz_diagP + data.A.nnz + len(mp.Hsblocks) + nnz_vec + p
    # maps are disjoint and cover everything except the structural-zero P diagonal
    allidx = np.concatenate([mp.P, mp.A, mp.Hsblocks, mp.soc_u, mp.soc_v, mp.soc_D])
    assert len(np.unique(allidx)) == len(allidx)
    # values: top-left = triu(P), top-right = A'
    Kd = KKT.toarray()
    assert np.array_equal(Kd[:n, :n], sp.triu(data.P).toarray())
    assert np.array_equal(Kd[:n, n:n + m], data.A.T.toarray())
    assert np.array_equal(KKT.data[mp.P], data.P.data)
    assert np.array_equal(KKT.data[mp.A], data.A.data)
    # writing through the Hs map reproduces block-diagonal placement
    vals = np.arange(1, len(mp.Hsblocks) + 1, dtype=float)
    K2 = KKT.copy(); K2.data[mp.Hsblocks] = vals
    D2 = K2.toarray()[n:n + m, n:n + m]
    for i in range(len(cones.specs)):
        a, b_ = cones.rng_cones[i], cones.rng_cones[i + 1]
        blk = vals[cones.rng_blocks[i]:cones.rng_blocks[i + 1]]
        sub = D2[a:b_, a:b_]
        if cones.Hs_is_diagonal[i]:
            assert np.array_equal(np.diag(sub), blk)
        else:
            ti, tj = np.tril_indices(b_ - a)
            assert np.array_equal(sub[tj, ti], blk)        # packed triu, column-major
    Ds = ka.fill_Dsigns(m, n, p)
    assert np.all(Ds[:n] == 1) and np.all(Ds[n:n + m] == -1)
    assert np.all(Ds[n + m::2] == -1) and np.all(Ds[n + m + 1::2] == 1)
