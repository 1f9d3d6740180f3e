"""CPU: the QDLDL restatement (oracle/qdldl_oracle.c) checked at the LDL boundary itself.

The reference has no tests at this boundary and QDLDL.jl is not vendored, so these
are *independent-method* checks, not reference goldens: the factors must reconstruct P K P', the
solve must agree with a dense LAPACK solve, the inertia must match Dsigns, and the
update/scale/refactor cycle (directldl_qdldl.jl:54-79) must equal a fresh factorisation."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle.qdldl import QDLDLFactorisation, amd_order


def _quasidefinite(rng, n, m, density=0.15):
    Ph = sp.random(n, n, density, random_state=np.random.RandomState(rng.integers(1 << 30)))
    Pm = (Ph @ Ph.T + sp.identity(n) * 0.1).tocsc()
    Am = sp.random(m, n, density * 2, random_state=np.random.RandomState(rng.integers(1 << 30))).tocsc()
    K = sp.bmat([[Pm, Am.T], [Am, -sp.identity(m) * (0.5 + rng.random())]]).tocsc()
    signs = np.r_[np.ones(n, dtype=np.int64), -np.ones(m, dtype=np.int64)]
    return K, signs


def _dense_L(F):
    Lp, Li, Lx, D, Dinv = F.factors()
    n = F.n
    Lm = sp.csc_matrix((Lx, Li, Lp), shape=(n, n)).toarray() + np.eye(n)
    return Lm, D, Dinv


@pytest.mark.parametrize("n,m,seed", [(1, 0, 0), (5, 3, 1), (20, 12, 2), (60, 45, 3), (150, 90, 4)])
def test_factors_reconstruct_and_solve(n, m, seed):
    rng = np.random.default_rng(seed)
    K, signs = _quasidefinite(rng, n, m)
    F = QDLDLFactorisation(sp.triu(K).tocsc(), signs)
    assert F.refactor()
    Lm, D, Dinv = _dense_L(F)
    p = F.perm
    Kp = K.toarray()[np.ix_(p, p)]
    assert np.allclose(Lm @ np.diag(D) @ Lm.T, Kp, rtol=1e-11, atol=1e-11)
    assert np.allclose(D * Dinv, 1.0, rtol=1e-15)
    assert np.array_equal(np.sign(D), signs[p])               # quasidefinite inertia, any ordering
    assert np.all(np.tril(Lm, -1)[np.triu_indices(n + m, 0)] == 0)
    b = rng.standard_normal(n + m)
    x = b.copy(); F.solve(x)
    assert np.allclose(x, np.linalg.solve(K.toarray(), b), rtol=1e-9, atol=1e-10)
    assert F.regularize_count == 0
    # factor flop count in the CHOLMOD convention: sum_j (nnz(L_j) + 1)^2
    Lp = F.factors()[0]
    assert F.sum_lnz_sq == float(((np.diff(Lp).astype(float) + 1) ** 2).sum())


def test_published_qdldl_example_matrix():
    """10x10 quasidefinite example (upper-triangular CSC) solved against a dense solve."""
    Ap = np.array([0, 1, 2, 4, 5, 6, 8, 10, 12, 14, 17])
    Ai = np.array([0, 1, 1, 2, 3, 4, 1, 5, 0, 6, 3, 7, 6, 8, 1, 2, 9])
    Ax = np.array([1.0, 0.460641, -0.121189, 0.417928, 0.177828, 0.1, -0.0290058, -1.0, 0.350321,
                   -0.441092, -0.0845395, -0.316228, 0.178663, -0.299077, 0.182452, -1.56506, -0.1])
    U = sp.csc_matrix((Ax, Ai, Ap), shape=(10, 10))
    K = (U + sp.triu(U, 1).T).toarray()
    b = np.arange(1.0, 11.0)
    xs = np.linalg.solve(K, b)
    for perm in (np.arange(10), amd_order(U), np.arange(10)[::-1].copy()):
        F = QDLDLFactorisation(U, np.sign(np.diag(K)).astype(np.int64), perm=perm, regularize=False)
        assert F.refactor()
        x = b.copy(); F.solve(x)
        assert np.allclose(x, xs, rtol=1e-10, atol=1e-12)


def test_update_scale_refactor_cycle_equals_fresh_factor():
    rng = np.random.default_rng(7)
    K, signs = _quasidefinite(rng, 30, 20)
    U = sp.triu(K).tocsc(); U.sort_indices()
    F = QDLDLFactorisation(U, signs); assert F.refactor()
    # overwrite a scattered subset, scale another (directldl_qdldl.jl:54, :66)
    idx = rng.choice(U.nnz, size=U.nnz // 3, replace=False)
    vals = U.data[idx] * (1.0 + 0.1 * rng.standard_normal(len(idx)))
    diag_idx = U.indptr[1:] - 1
    # keep the diagonal's sign so the matrix stays quasidefinite
    isdiag = np.isin(idx, diag_idx); vals[isdiag] = np.abs(vals[isdiag]) * np.sign(U.data[idx][isdiag])
    sidx = np.setdiff1d(np.arange(U.nnz), idx)[:15]
    U2 = U.copy(); U2.data[idx] = vals; U2.data[sidx] *= 0.75
    F.update_values(idx, vals); F.scale_values(sidx, 0.75); assert F.refactor()
    G = QDLDLFactorisation(U2, signs, perm=F.perm); assert G.refactor()
    for a, b in zip(F.factors(), G.factors()):
        assert np.array_equal(a, b)                            # same arithmetic, bit for bit
    r = rng.standard_normal(50); x = r.copy(); F.solve(x)
    K2 = (U2 + sp.triu(U2, 1).T).toarray()
    assert np.allclose(K2 @ x, r, atol=1e-9)


def test_dynamic_regularisation_and_failure_flag():
    # zero pivot in a +1 slot -> replaced by +delta; in a -1 slot -> -delta (settings.jl:122-124)
    U = sp.csc_matrix((np.array([0.0, 1.0, -1.0, 0.0]), np.array([0, 0, 1, 2]), np.array([0, 1, 3, 4])), shape=(3, 3))
    signs = np.array([1, -1, -1], dtype=np.int64)
    F = QDLDLFactorisation(U, signs, perm=np.arange(3), eps=1e-13, delta=2e-7)
    assert F.refactor()
    D = F.factors()[3]
    assert D[0] == 2e-7 and D[2] == -2e-7 and F.regularize_count == 2
    assert D[1] == -1.0 - 1.0 / 2e-7
    # without regularisation the zero pivot makes Dinv non-finite -> refactor! reports failure (:79)
    G = QDLDLFactorisation(U, signs, perm=np.arange(3), regularize=False)
    assert not G.refactor()
    # wrong-signed pivot is also replaced
    U3 = sp.csc_matrix(np.diag([1.0, 0.5]))
    H = QDLDLFactorisation(sp.triu(U3).tocsc(), np.array([1, -1], dtype=np.int64), perm=np.arange(2))
    assert H.refactor() and H.factors()[3][1] == -2e-7 and H.regularize_count == 1
