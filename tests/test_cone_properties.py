"""CPU: identity-based cone property tests carried over from the reference's unit tests
(test/UnitTests/test_coneops_secondordercone.jl:31-91, test_coneops_psdtrianglecone.jl:112-251):
they pin the host-side scaling state the KKT backends read (w, eta, d, u, v, R) and the oracle's
get_Hs! without depending on Julia's RNG stream."""
import numpy as np
import pytest


def _soc_point(rng, dim):
    x = rng.standard_normal(dim)
    x[0] = np.linalg.norm(x[1:]) + 0.1 + rng.random()
    return x


@pytest.mark.parametrize("dim", [3, 4, 5, 9, 40])
def test_soc_scaling_identities(cb, dim):
    from oracle.kktsolver_oracle import get_Hs
    rng = np.random.default_rng(dim)
    cones = cb.CompositeCone([cb.SecondOrderConeT(dim)])
    s, z = _soc_point(rng, dim), _soc_point(rng, dim)
    assert cones.update_scaling(s, z, 1.0)
    w, eta = cones.w.copy(), cones.soc_eta[0]
    J = np.diag(np.r_[1.0, -np.ones(dim - 1)])
    # w is a unit hyperbolic vector
    assert abs(w @ J @ w - 1.0) < 1e-12
    W2 = eta ** 2 * (2 * np.outer(w, w) - J)                      # W'W
    # lambda = W z = W^{-T} s  and  (W'W) z = s
    assert np.allclose(W2 @ z, s, rtol=1e-10, atol=1e-10)
    y = np.zeros(dim); cones.mul_Hs(y, z)
    assert np.allclose(y, s, rtol=1e-10, atol=1e-10)
    lam = np.zeros(dim); cones.mul_W("N", lam, z)
    lam2 = np.zeros(dim); cones.mul_Winv("T", lam2, s)
    assert np.allclose(lam, cones.lam, atol=1e-10) and np.allclose(lam2, cones.lam, atol=1e-10)
    # sparse representation: eta^2 (D + uu' - vv') == eta^2 (2ww' - J)   (test_..._secondordercone.jl:60-66)
    u, v, d = cones.soc_u, cones.soc_v, cones.soc_d[0]
    D = np.eye(dim); D[0, 0] = d
    assert np.allclose(eta ** 2 * (D + np.outer(u, u) - np.outer(v, v)), W2, rtol=1e-12, atol=1e-12)
    # oracle get_Hs!: packed block reproduces W'W (dense form) or its diagonal part (sparse form)
    Hs = np.zeros(int(cones.rng_blocks[-1])); get_Hs(cones, Hs)
    if dim <= 4:
        ti, tj = np.tril_indices(dim)
        assert np.allclose(Hs, W2[tj, ti], rtol=1e-12, atol=1e-12)
    else:
        assert np.allclose(Hs, eta ** 2 * np.diag(D), rtol=1e-14)


@pytest.mark.parametrize("n", [2, 3, 6])
def test_psd_scaling_identities(cb, n):
    from oracle.kktsolver_oracle import get_Hs, skron_triu, skron_triu_loops
    rng = np.random.default_rng(100 + n)
    cones = cb.CompositeCone([cb.PSDTriangleConeT(n)])
    g = cones.psd_groups[0]

    def rand_pd():
        G = rng.standard_normal((n, n)); return G @ G.T + 0.5 * np.eye(n)
    S, Z = rand_pd(), rand_pd()
    svec = lambda M: cones._psd_svec(g, M[None])[0]
    s, z = svec(S), svec(Z)
    assert cones.update_scaling(s, z, 1.0)
    R, Rinv, lam = g["R"][0], g["Rinv"][0], g["lam"][0]
    # R'ZR = Lambda, Rinv S Rinv' = Lambda, R Rinv = I     (test_..._psdtrianglecone.jl:133-137)
    assert np.allclose(R.T @ Z @ R, np.diag(lam), atol=1e-9)
    assert np.allclose(Rinv @ S @ Rinv.T, np.diag(lam), atol=1e-9)
    assert np.allclose(R @ Rinv, np.eye(n), atol=1e-9)
    # W'W z = s
    y = np.zeros(len(z)); cones.mul_Hs(y, z)
    assert np.allclose(y, s, rtol=1e-8, atol=1e-8)
    # Symmetric(unpack(get_Hs!)) v == W'W v     (test_..._psdtrianglecone.jl:236-249)
    Hs = np.zeros(int(cones.rng_blocks[-1])); get_Hs(cones, Hs)
    ne = len(z)
    H = np.zeros((ne, ne)); ti, tj = np.tril_indices(ne); H[tj, ti] = Hs; H = H + np.triu(H, 1).T
    v = rng.standard_normal(ne)
    y = np.zeros(ne); cones.mul_Hs(y, v)
    assert np.allclose(H @ v, y, rtol=1e-8, atol=1e-8)
    # vectorised skron == literal loop form of skron! (coneops_psdtrianglecone.jl:502-540)
    A = R @ R.T
    assert np.allclose(skron_triu(A), skron_triu_loops(A), rtol=1e-13, atol=1e-13)


def test_nn_and_composite_ranges(cb):
    specs = [cb.ZeroConeT(2), cb.NonnegativeConeT(3), cb.SecondOrderConeT(4), cb.SecondOrderConeT(6),
             cb.PSDTriangleConeT(3)]
    cones = cb.CompositeCone(specs)
    # rng_blocks: diag for Zero/NN/sparse SOC, packed triangle otherwise (compositecone_type.jl:126-141)
    assert cones.rng_cones.tolist() == [0, 2, 5, 9, 15, 21]
    assert cones.rng_blocks.tolist() == [0, 2, 5, 15, 21, 42]
    assert cones.p == 2 and cones.degree == 3 + 1 + 1 + 3
    rng = np.random.default_rng(0)
    s = np.abs(rng.standard_normal(cones.numel)) + 0.5
    z = np.abs(rng.standard_normal(cones.numel)) + 0.5
    for a, b in ((5, 9), (9, 15)):
        s[a] += np.linalg.norm(s[a + 1:b]); z[a] += np.linalg.norm(z[a + 1:b])
    S3 = np.eye(3) * 2 + 0.1; Z3 = np.eye(3) * 3 - 0.1
    g = cones.psd_groups[0]
    s[15:21] = cones._psd_svec(g, S3[None])[0]; z[15:21] = cones._psd_svec(g, Z3[None])[0]
    assert cones.update_scaling(s, z, 1.0)
    assert np.allclose(cones.w[2:5] ** 2, s[2:5] / z[2:5])
    # collapsing (cone_api.jl:96-152)
    col = cb.cones.cones_new_collapsed([cb.NonnegativeConeT(2), cb.SecondOrderConeT(1), cb.PSDTriangleConeT(1),
                                        cb.ZeroConeT(0), cb.NonnegativeConeT(3), cb.SecondOrderConeT(3)])
    assert col == [cb.NonnegativeConeT(7), cb.SecondOrderConeT(3)]


def test_cones_new_collapsed_reference_cases(cb):
    """test/UnitTests/test_cones_new_collapsed.jl, case by case."""
    Z, NN, SOC, PSDT, EXP = cb.ZeroConeT, cb.NonnegativeConeT, cb.SecondOrderConeT, cb.PSDTriangleConeT, cb.ExponentialConeT
    col = cb.cones.cones_new_collapsed
    assert col([NN(3), SOC(4), EXP()]) == [NN(3), SOC(4), EXP()]                              # :3-13
    assert col([NN(3), NN(2), SOC(4)]) == [NN(5), SOC(4)]                                     # :14-26
    assert col([NN(3), Z(0), SOC(4), NN(0)]) == [NN(3), SOC(4)]                               # :27-40
    assert col([SOC(1), SOC(4)]) == [NN(1), SOC(4)]                                           # :41-52
    assert col([PSDT(1), SOC(4)]) == [NN(1), SOC(4)]                                          # :53-64
    assert col([SOC(1), NN(3), NN(2), EXP(), NN(0), SOC(1)]) == [NN(6), EXP(), NN(1)]         # :65-81
    assert col([NN(3), NN(2), Z(0), SOC(1), PSDT(1), SOC(4), NN(0)]) == [NN(7), SOC(4)]       # :82-98
