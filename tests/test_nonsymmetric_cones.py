"""CPU: nonsymmetric cone operations (clarabel.jl_b200/nonsymmetric.py) checked against finite
differences of their own barrier functions, the conjugacy identity and the secant equations of the
primal-dual scaling; then the three reference known answers that involve these cones
(test/OptTests/basic_exp.jl:58-76, basic_pow.jl:58-66, basic_genpow.jl:53-61)."""
import numpy as np
import pytest
import scipy.sparse as sp

from clarabel_jl_b200 import nonsymmetric as ns


def _fd_grad(f, x, h=1e-6):
    g = np.zeros(len(x))
    for i in range(len(x)):
        e = np.zeros(len(x)); e[i] = h
        g[i] = (f(x + e) - f(x - e)) / (2 * h)
    return g


def _interior_points(cone, rng, k=6):
    """random strictly interior primal/dual pairs obtained by walking from the unit point"""
    z0, s0 = cone.unit_initialization()
    out = []
    for _ in range(k):
        for _try in range(100):
            s = s0 + 0.4 * rng.standard_normal(len(s0)) * np.maximum(1.0, np.abs(s0))
            z = z0 + 0.4 * rng.standard_normal(len(z0)) * np.maximum(1.0, np.abs(z0))
            if cone.is_primal_feasible(s) and cone.is_dual_feasible(z):
                out.append((s, z)); break
    assert len(out) >= 3
    return out


CONES3 = [("exp", lambda: ns.ExponentialCone()), ("pow0.6", lambda: ns.PowerCone(0.6)),
          ("pow0.1", lambda: ns.PowerCone(0.1)), ("pow0.5", lambda: ns.PowerCone(0.5))]


@pytest.mark.parametrize("name,mk", CONES3)
def test_dual_barrier_derivatives(name, mk):
    rng = np.random.default_rng(1)
    cone = mk()
    for s, z in _interior_points(cone, rng):
        grad, H = cone.dual_grad_hess(z)
        assert np.allclose(grad, _fd_grad(cone.barrier_dual, z), rtol=1e-6, atol=1e-7)
        Hfd = np.array([_fd_grad(lambda x, i=i: cone.dual_grad_hess(x)[0][i], z) for i in range(3)])
        assert np.allclose(H, Hfd, rtol=1e-5, atol=1e-6)
        assert np.allclose(H, H.T) and np.all(np.linalg.eigvalsh(H) > 0)
        # logarithmic homogeneity: <grad f*(z), z> = -nu
        assert abs(grad @ z + 3.0) < 1e-9
        # third-order term: eta = 1/2 D^3 f*(z)[u, v]
        u, v = rng.standard_normal(3), rng.standard_normal(3)
        h = 1e-5
        D3 = (cone.dual_grad_hess(z + h * u)[1] - cone.dual_grad_hess(z - h * u)[1]) @ v / (2 * h)
        assert np.allclose(cone._eta(z, u, v), 0.5 * D3, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name,mk", CONES3)
def test_primal_gradient_is_the_conjugate_map(name, mk):
    rng = np.random.default_rng(2)
    cone = mk()
    for s, z in _interior_points(cone, rng):
        g = cone.gradient_primal(s)
        assert cone.is_dual_feasible(-g)
        assert np.allclose(-cone.dual_grad_hess(-g)[0], s, rtol=1e-9, atol=1e-10)      # -f*'(-f'(s)) = s
        assert abs(g @ s + 3.0) < 1e-9
        assert np.allclose(g, _fd_grad(cone.barrier_primal, s), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name,mk", CONES3)
def test_primal_dual_scaling_satisfies_the_secant_equations(name, mk):
    rng = np.random.default_rng(3)
    cone = mk()
    for s, z in _interior_points(cone, rng):
        mu = float(s @ z) / 3
        assert cone.update_scaling(s, z, mu, ns.PRIMAL_DUAL)
        Hs = cone.Hs
        assert np.allclose(Hs, Hs.T)
        zt = -cone.gradient_primal(s); st = -cone.grad
        if not np.allclose(Hs, mu * cone.H_dual):              # not the central-path fallback
            assert np.allclose(Hs @ z, s, rtol=1e-8, atol=1e-9)
            assert np.allclose(Hs @ zt, st, rtol=1e-7, atol=1e-8)
            assert np.all(np.linalg.eigvalsh(Hs) > 0)
        assert cone.update_scaling(s, z, mu, ns.DUAL)
        assert np.allclose(cone.Hs, mu * cone.H_dual)
        H = cone.Hs                                             # pack_triu: column-major upper triangle
        assert np.array_equal(cone.hs_triu(), [H[0, 0], H[0, 1], H[1, 1], H[0, 2], H[1, 2], H[2, 2]])


def test_wright_omega():
    for beta in (1.0, 1.3, 2.0, 4.2, 10.0, 1e3):
        w = ns._wright_omega(beta)
        assert abs(w + np.log(w) - beta) < 1e-12 * max(1.0, beta)


@pytest.mark.parametrize("alpha,dim2", [((0.6, 0.4), 1), ((0.1, 0.9), 1), ((0.2, 0.3, 0.5), 2)])
def test_genpow_cone(alpha, dim2):
    rng = np.random.default_rng(4)
    cone = ns.GenPowerCone(alpha, dim2)
    for s, z in _interior_points(cone, rng):
        mu = 0.7
        assert cone.update_scaling(s, z, mu, ns.DUAL)
        assert np.allclose(cone.grad, _fd_grad(cone.barrier_dual, z), rtol=1e-6, atol=1e-7)
        # rank-3 representation == mu * Hessian of the dual barrier
        Hfd = np.array([_fd_grad(lambda x, i=i: _grad_of(cone, x)[i], z) for i in range(cone.dim)])
        Hrep = np.column_stack([cone.mul_Hs(e) for e in np.eye(cone.dim)])
        assert np.allclose(Hrep, mu * Hfd, rtol=1e-5, atol=1e-6)
        D = np.diag(cone.hs_diag())
        q = np.concatenate([cone.q, np.zeros(cone.dim2)]); r = np.concatenate([np.zeros(cone.dim1), cone.r])
        assert np.allclose(Hrep, D + mu * (np.outer(cone.p, cone.p) - np.outer(q, q) - np.outer(r, r)))
        g = cone.gradient_primal(s)
        assert cone.is_dual_feasible(-g)
        assert np.allclose(-_grad_of(cone, -g), s, rtol=1e-9, atol=1e-10)
        assert abs(g @ s + cone.degree) < 1e-9


def _grad_of(cone, z):
    c2 = ns.GenPowerCone(cone.alpha, cone.dim2)
    assert c2.update_scaling(z, z, 1.0, ns.DUAL)
    return c2.grad


# ---------------------------------------------------------------- reference known answers
def _exp_problem(cb):
    A1 = np.hstack([np.ones((1, 3)), np.zeros((1, 4))])
    A2 = np.hstack([np.zeros((3, 2)), -np.eye(3), np.zeros((3, 2))])
    A3 = np.zeros((3, 7)); A3[0, 0] = -1; A3[1, 2] = -1; A3[2, 4] = -1
    c = np.array([1.0, 0.5, -2.0, -0.1, 1.0, 3.0, 0.0])
    P = sp.identity(7, format="csc") * 0.1
    A = sp.csc_matrix(np.vstack([A1, A2, A3]))
    b = np.array([10.0, 0, 0, 0, 0, 0, 0])
    return P, c, A, b, [cb.ZeroConeT(1), cb.NonnegativeConeT(3), cb.ExponentialConeT()]


def _pow_problem(cb):
    P = sp.csc_matrix((6, 6)); q = np.zeros(6); q[2] = q[5] = -1
    A = -sp.csc_matrix(np.vstack([np.eye(6), [[1.0, 2, 0, 3, 0, 0]], [[0, 0, 0, 0, 1.0, 0]]]))
    b = np.concatenate([np.zeros(6), [-3.0], [-1.0]])
    return P, q, A, b, [cb.PowerConeT(0.6), cb.PowerConeT(0.1), cb.ZeroConeT(1), cb.ZeroConeT(1)]


def _genpow_problem(cb):
    P = sp.csc_matrix((6, 6)); q = np.zeros(6); q[2] = q[5] = -1
    A = sp.csc_matrix(np.vstack([-np.eye(6), [[1.0, 2, 0, 3, 0, 0]], [[0, 0, 0, 0, 1.0, 0]]]))
    b = np.array([0, 0, 0, 0, 0, 0, 3.0, 1.0])
    return P, q, A, b, [cb.GenPowerConeT([0.6, 0.4], 1), cb.GenPowerConeT([0.1, 0.9], 1), cb.ZeroConeT(2)]


def test_reference_exp_known_answer(cb):
    sol = cb.Solver(*_exp_problem(cb), cb.Settings(direct_solve_method="qdldl")).solve()
    assert sol.status_name == "SOLVED"
    xref = np.array([-9.425995201329599, 4.828561507482018, 14.59743362204262, 1.0000012112102774,
                     7.65314081561849, -29.99999978458479, -0.0])
    assert np.linalg.norm(sol.x - xref) < 1e-3                     # basic_exp.jl:66-75
    assert abs(sol.obj_val - (-54.41243965302268)) < 1e-3          # :76


def test_reference_pow_known_answer(cb):
    sol = cb.Solver(*_pow_problem(cb), cb.Settings(direct_solve_method="qdldl")).solve()
    assert sol.status_name == "SOLVED"
    assert abs(sol.obj_val - (-1.8458)) < 1e-3                     # basic_pow.jl:65


def test_reference_genpow_known_answer(cb):
    sol = cb.Solver(*_genpow_problem(cb), cb.Settings(direct_solve_method="qdldl")).solve()
    assert sol.status_name == "SOLVED"
    assert abs(sol.obj_val - (-1.8458)) < 1e-3                     # basic_genpow.jl:60


def test_b200_nonsymmetric_payload_equals_oracle_kkt_values(cb):
    """What the B200 backend sends through update_values! for the nonsymmetric cones must be
    exactly what the reference-path oracle writes into K at the same positions."""
    from clarabel_jl_b200 import kkt_assembly as ka
    from clarabel_jl_b200.kktsolver_b200 import nonsym_update_index, nonsym_update_values
    from oracle.kktsolver_oracle import OracleDirectLDLKKTSolver
    specs = [cb.NonnegativeConeT(2), cb.GenPowerConeT([0.3, 0.7], 2), cb.ExponentialConeT(),
             cb.SecondOrderConeT(6), cb.PowerConeT(0.25), cb.GenPowerConeT([0.5, 0.2, 0.3], 1)]
    cones = cb.CompositeCone(specs)
    m = cones.numel; n = 5
    rng = np.random.default_rng(0)
    A = sp.random(m, n, 0.5, random_state=np.random.RandomState(0), format="csc") + sp.csc_matrix((m, n))
    P = sp.identity(n, format="csc")
    z = np.zeros(m); s = np.zeros(m)
    cones.unit_initialization(z, s)
    s += 0.05 * rng.standard_normal(m) * (np.abs(s) > 0); z += 0.05 * rng.standard_normal(m) * (np.abs(z) > 0)
    assert cones.update_scaling(s, z, 0.8, 0)
    assert cones.p == 3 + 2 + 3
    st = cb.Settings(direct_solve_method="qdldl")
    ks = OracleDirectLDLKKTSolver(P, A.tocsc(), cones, m, n, st)
    assert ks.update(cones)
    K, mp = ka.assemble_kkt_matrix(P, A.tocsc(), cones)
    idx, vals = nonsym_update_index(mp, cones), nonsym_update_values(mp, cones)
    assert len(idx) == len(vals) == len(set(idx.tolist()))
    Kref = ks.KKT.data.copy()
    Kref[mp.diag_full] = ks.KKT.data[mp.diag_full]          # (unregularised copy is what KKT holds)
    assert np.array_equal(Kref[idx], vals)
    # Dsigns follow the expansion columns in cone order: genpow (-1,-1,+1), SOC (-1,+1), genpow
    Ds = ka.fill_Dsigns(m, n, cones.p, cones)
    assert Ds[n + m:].tolist() == [-1, -1, 1, -1, 1, -1, -1, 1]
    # and the assembled pattern is a valid sorted upper-triangular CSC
    Kc = K.copy(); Kc.data[:] = 1.0
    assert (sp.tril(Kc, -1)).nnz == 0 and Kc.has_sorted_indices


def _mixed_problem(cb):
    """min x3 + t + u  s.t.  x3 >= exp(x1),  t >= |(x1 - 1, 0.5)|,  [[u, x1], [x1, 1]] >= 0"""
    r2 = np.sqrt(2.0)
    A = np.zeros((9, 4)); b = np.zeros(9)
    A[0, 0] = -1.0; b[1] = 1.0; A[2, 1] = -1.0                  # (x1, 1, x3) in K_exp
    A[3, 2] = -1.0; A[4, 0] = -1.0; b[4] = -1.0; b[5] = 0.5     # (t, x1 - 1, 0.5) in SOC
    A[6, 3] = -1.0; A[7, 0] = -r2; b[8] = 1.0                   # svec [[u, x1], [x1, 1]] in PSD
    q = np.array([0.0, 1.0, 1.0, 1.0])
    K = [cb.ExponentialConeT(), cb.SecondOrderConeT(3), cb.PSDTriangleConeT(2)]
    return sp.csc_matrix((4, 4)), q, sp.csc_matrix(A), b, K


def test_mixed_symmetric_and_nonsymmetric_cones(cb):
    """exp + SOC + PSD cones in one problem (unit initialisation, barrier, step-length and
    ds-shift paths of every cone type together); the optimal value is
    min_x1 exp(x1) + sqrt((x1-1)^2 + 0.25) + x1^2."""
    from scipy.optimize import minimize_scalar
    sol = cb.Solver(*_mixed_problem(cb), cb.Settings(direct_solve_method="qdldl")).solve()
    f = lambda x: np.exp(x) + np.sqrt((x - 1) ** 2 + 0.25) + x * x
    ref = minimize_scalar(f, bounds=(-2, 2), method="bounded", options=dict(xatol=1e-12))
    assert sol.status_name == "SOLVED"
    assert abs(sol.obj_val - ref.fun) < 1e-6 and abs(sol.x[0] - ref.x) < 1e-4


def _sdp_chordal_problem(cb):
    """The explicit CSC instance of test/OptTests/sdp_chordal.jl:6-76 (NN(1), PSD(6), Pow(1/3),
    Pow(1/2)); the reference solves it under every chordal-decomposition setting, which is an
    equivalent reformulation -- here it is solved undecomposed (decomposition is out of scope)."""
    r2 = np.sqrt(2.0)
    colptr = np.array([0, 1, 4, 5, 8, 9, 10, 13, 16])
    rowval = np.array([24, 7, 10, 22, 8, 12, 15, 25, 9, 13, 18, 21, 26, 0, 23, 27])
    nzval = np.array([-1.0, -r2, -1.0, -1.0, -r2, -r2, -1.0, -1.0, -r2, -r2, -r2, -1.0, -1.0, -1.0, -1.0, -1.0])
    A = sp.csc_matrix((nzval, rowval, colptr), shape=(28, 8))
    b = np.zeros(28); b[1:7] = [3.0, 2 * r2, 2.0, r2, r2, 3.0]
    c = np.zeros(8); c[0] = -1.0
    K = [cb.NonnegativeConeT(1), cb.PSDTriangleConeT(6), cb.PowerConeT(0.3333333333333333), cb.PowerConeT(0.5)]
    return sp.csc_matrix((8, 8)), c, A, b, K


def test_reference_sdp_chordal_instance_undecomposed(cb):
    sol = cb.Solver(*_sdp_chordal_problem(cb), cb.Settings(direct_solve_method="qdldl")).solve()
    assert sol.status_name == "SOLVED"                              # sdp_chordal.jl:101
