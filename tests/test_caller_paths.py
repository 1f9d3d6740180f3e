"""CPU: the data-update path of the caller (Solver.update_P/A/q/b -> kktsolver_update_P!/A!),
scenario by scenario as test/OptTests/data_updating.jl (tolerance 1e-7 on the solutions of the
updated solver vs a fresh solver on the new data).  Indices in the (index, value) form are 0-based
here (the reference is 1-based)."""
import numpy as np
import pytest
import scipy.sparse as sp

TOL = 1e-7


def _data(cb):
    P = sp.csc_matrix(np.array([[4.0, 1.0], [1.0, 2.0]]))
    q = np.array([1.0, 1.0])
    A0 = sp.identity(2, format="csc")
    A = sp.vstack([-A0, A0]).tocsc()
    b = np.array([1.0, 1.0, 1.0, 1.0])                       # [-l; u] with l = -1, u = 1
    return P, q, A, b, [cb.NonnegativeConeT(2), cb.NonnegativeConeT(2)]


def _fresh(cb, method, P, q, A, b, K):
    return cb.Solver(P, q, A, b, K, cb.Settings(direct_solve_method=method)).solve()


SCENARIOS = ["P_matrix", "P_vector", "P_pairs", "A_matrix", "A_vector", "A_pairs",
             "q_vector", "q_pairs", "b_vector", "b_pairs", "all_at_once", "noop"]


def run_scenario(cb, method, scenario, tol=TOL):
    P, q, A, b, K = _data(cb)
    s1 = cb.Solver(P, q, A, b, K, cb.Settings(direct_solve_method=method))
    s1.solve()
    P2, q2, A2, b2 = P.copy(), q.copy(), A.copy(), b.copy()
    if scenario == "P_matrix":                               # data_updating.jl:34-57
        P2 = sp.csc_matrix(np.array([[100.0, 1.0], [1.0, 2.0]]))
        s1.update_P(sp.triu(P2).tocsc())
    elif scenario == "P_vector":                             # :59-77
        P2 = sp.csc_matrix(np.array([[100.0, 1.0], [1.0, 2.0]]))
        s1.update_P(sp.triu(P2).tocsc().data)
    elif scenario == "P_pairs":                              # :79-98 (1-based [2,3] -> 0-based [1,2])
        s1.update_P(zip([1, 2], [3.0, 5.0]))
        P2 = sp.csc_matrix(np.array([[4.0, 3.0], [0.0, 5.0]]))
    elif scenario in ("A_matrix", "A_vector"):               # :100-146
        A2 = A.tolil(); A2[1, 1] = -1000.0; A2 = A2.tocsc()
        s1.update_A(A2 if scenario == "A_matrix" else A2.data)
    elif scenario == "A_pairs":                              # :148-168
        s1.update_A(zip([1, 2], [0.5, -0.5]))
        A2 = A.copy(); A2.data[[1, 2]] = [0.5, -0.5]
    elif scenario == "q_vector":                             # :170-190
        q2[0] = 10.0; s1.update_q(q2)
    elif scenario == "q_pairs":                              # :192-212
        s1.update_q(zip([1], [10.0])); q2[1] = 10.0
    elif scenario == "b_vector":                             # :214-234
        b2[:] = 0.0; s1.update_b(b2)
    elif scenario == "b_pairs":                              # :236-256
        s1.update_b(zip([1, 3], [0.0, 0.0])); b2[[1, 3]] = 0.0
    elif scenario == "all_at_once":                          # update_data! :22-37
        P2 = sp.csc_matrix(np.array([[5.0, 0.5], [0.5, 3.0]])); q2 = np.array([-1.0, 2.0])
        A2 = A.copy(); A2.data *= 1.5; b2 = b * 0.5
        s1.update_data(sp.triu(P2).tocsc(), q2, A2, b2)
    elif scenario == "noop":                                 # nothing / empty input: no action
        s1.update_data(None, None, None, None); s1.update_P(np.zeros(0)); s1.update_b([])
    sol1 = s1.solve()
    sol2 = _fresh(cb, method, P2, q2, A2, b2, K)
    assert sol1.status_name == sol2.status_name == "SOLVED"
    assert np.linalg.norm(sol1.x - sol2.x) < tol
    assert abs(sol1.obj_val - sol2.obj_val) < 1e-6


@pytest.mark.parametrize("scenario", SCENARIOS)
def test_data_updating_oracle_backend(cb, scenario):
    run_scenario(cb, "qdldl", scenario)


def test_update_rejects_wrong_pattern_or_length(cb):
    P, q, A, b, K = _data(cb)
    s = cb.Solver(P, q, A, b, K, cb.Settings(direct_solve_method="qdldl"))
    with pytest.raises(ValueError):
        s.update_P(np.ones(7))
    with pytest.raises(ValueError):
        s.update_A(sp.csc_matrix(np.ones((4, 2))))
    with pytest.raises(ValueError):
        s.update_q(np.ones(3))
    with pytest.raises(ValueError):
        s.update_P(zip([9], [1.0]))


# ---------------------------------------------------------------- equilibration bounds
# test/UnitTests/test_equilibration_bounds.jl:24-85 (the scaling decides the K values the path sees)
def _equil_data(cb):
    P = sp.csc_matrix(np.array([[4.0, 1.0], [1.0, 2.0]]))
    c = np.array([1.0, 1.0])
    A0 = sp.csc_matrix(np.array([[1.0, 1.0], [1.0, 0.0], [0.0, 1.0]]))
    A = sp.vstack([-A0, A0]).tolil()
    b = np.array([-1.0, 0.0, 0.0, 1.0, 0.7, 0.7])
    return P.tolil(), c, A, b, [cb.NonnegativeConeT(3), cb.NonnegativeConeT(3)]


@pytest.mark.parametrize("case", ["lower", "upper", "zero_rows"])
def test_equilibration_bounds(cb, case):
    P, c, A, b, K = _equil_data(cb)
    st = cb.Settings(direct_solve_method="qdldl")
    if case == "lower":
        P[0, 0] = 1e-15
    elif case == "upper":
        A[0, 0] = 1e15
    else:
        A = A.tocsc(); A.data[:] = 0.0
    s = cb.Solver(sp.triu(P.tocsc()).tocsc(), c, A.tocsc(), b, K, st)
    d, e = s.data.d, s.data.e
    if case == "zero_rows":
        assert np.all(e == 1.0)
    else:
        assert d.min() >= st.equilibrate_min_scaling and e.min() >= st.equilibrate_min_scaling
        assert d.max() <= st.equilibrate_max_scaling and e.max() <= st.equilibrate_max_scaling


# ---------------------------------------------------------------- presolver (test/OptTests/presolve.jl)
def _presolve_data(cb):
    I3 = sp.identity(3, format="csc")
    A = (sp.vstack([I3, -I3]) * 2.0).tocsc()
    return I3.copy(), np.array([3.0, -2.0, 1.0]), A, np.ones(6), [cb.NonnegativeConeT(3), cb.NonnegativeConeT(3)]


def test_presolver_single_unbounded_constraint(cb):
    P, c, A, b, K = _presolve_data(cb)
    b[3] = 1e30
    s = cb.Solver(P, c, A, b, K, cb.Settings(direct_solve_method="qdldl"))
    sol = s.solve()
    assert sol.status_name == "SOLVED" and len(s.variables.z) == 5          # presolve.jl:34-44
    assert sol.z[3] == 0.0 and sol.s[3] == 1e20 and len(sol.z) == 6
    with pytest.raises(RuntimeError):
        s.update_b(np.ones(6))                                             # data_updating.jl: not allowed when presolved


def test_presolver_redundant_cone_and_all_redundant(cb):
    P, c, A, b, K = _presolve_data(cb)
    b[:3] = 1e30
    s = cb.Solver(P, c, A, b, K, cb.Settings(direct_solve_method="qdldl"))
    sol = s.solve()
    assert sol.status_name == "SOLVED" and len(s.variables.z) == 3          # :46-60
    assert np.all(sol.z[:3] == 0.0) and np.all(sol.s[:3] == 1e20)
    assert np.linalg.norm(sol.x - np.array([-0.5, 2.0, -0.5])) < 1e-3
    b[:] = 1e30
    s = cb.Solver(P, c, A, b, K, cb.Settings(direct_solve_method="qdldl"))
    sol = s.solve()
    assert sol.status_name == "SOLVED" and len(s.variables.z) == 0          # :62-74
    assert np.linalg.norm(sol.x + c) < 1e-3


# ---------------------------------------------------------------- JSON fixtures (test/UnitTests/test_json.jl)
def test_json_roundtrip(cb, tmp_path):
    P = sp.csc_matrix(np.array([[4.0, 1.0], [1.0, 2.0]]))
    c = np.array([1.0, 1.0])
    A = sp.csc_matrix(np.array([[1.0, 1.0], [1.0, 0.0], [0.0, 1.0]]))
    b = np.ones(3)
    K = [cb.NonnegativeConeT(1), cb.ZeroConeT(1), cb.NonnegativeConeT(1)]
    s1 = cb.Solver(P, c, A, b, K, cb.Settings(direct_solve_method="qdldl"))
    f = str(tmp_path / "problem.json")
    s1.save_to_file(f)
    s2 = cb.Solver.load_from_file(f)
    assert s2.settings.direct_solve_method == "qdldl"
    x1, x2 = s1.solve(), s2.solve()
    assert x1.status_name == x2.status_name and np.allclose(x1.x, x2.x, atol=1e-10)   # test_json.jl:24-25
    st = cb.Settings(direct_solve_method="qdldl", max_iter=1)
    s3 = cb.Solver.load_from_file(f, st)
    assert s3.solve().status_name == "MAX_ITERATIONS"                                  # :29-33
    # every cone type survives the schema (src/json.jl:138-151, 196-216)
    from clarabel_jl_b200 import problems
    K2 = [cb.ZeroConeT(1), cb.NonnegativeConeT(2), cb.SecondOrderConeT(3), cb.PSDTriangleConeT(2),
          cb.ExponentialConeT(), cb.PowerConeT(0.3), cb.GenPowerConeT([0.25, 0.75], 2)]
    m = sum(3 if k[0] == "PSDTriangleConeT" else k[1] for k in K2)
    f2 = str(tmp_path / "cones.json")
    problems.to_reference_json(f2, P, c, sp.csc_matrix((m, 2)), np.zeros(m), K2)
    assert problems.from_reference_json(f2)[4] == K2


def test_committed_c1_fixture_is_the_generated_instance(cb):
    """tests/golden/C1.json (tools/export_fixtures.py) is BASELINE config C1 in the reference's own
    fixture format; it must stay in sync with the generator the bench uses."""
    import os
    from clarabel_jl_b200 import problems
    P, q, A, b, K, _ = problems.from_reference_json(os.path.join(os.path.dirname(__file__), "golden", "C1.json"))
    P0, q0, A0, b0, K0 = problems.c1_random_qp()
    assert K == list(K0) and np.array_equal(q, q0) and np.array_equal(b, b0)
    assert (abs(sp.triu(P0) - P)).max() == 0 and (abs(sp.csc_matrix(A0) - A)).max() == 0


def test_socp_lasso_shape_instance(cb):
    """test/OptTests/socp-lasso.jl:58-65: 419 variables, 820 rows, a 402-dimensional second-order
    cone (sparse expansion), status SOLVED."""
    from clarabel_jl_b200 import problems
    P, c, A, b, K = problems.socp_lasso()
    assert A.shape == (820, 419) and [k[1] for k in K] == [402, 16, 402]
    s = cb.Solver(P, c, A, b, K, cb.Settings(direct_solve_method="qdldl"))
    assert s.cones.p == 2                                       # NN cones collapse, one expanded SOC
    assert s.solve().status_name == "SOLVED"
