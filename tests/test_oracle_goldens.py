"""CPU: the oracle (restated IP loop + QDLDL-algorithm engine) reproduces every golden answer the
reference's own tests hold for this path (test/OptTests/*.jl; atol = 1e-3 as in the reference)."""
import numpy as np
import pytest
import reference_cases as rc


def _cases():
    import clarabel_jl_b200  # noqa: F401  (cone constructors)
    return rc.cases()


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_oracle_reproduces_reference_goldens(cb, case):
    st = cb.Settings(direct_solve_method="qdldl")
    s = cb.Solver(case["P"], case["q"], case["A"], case["b"], case["cones"], st)
    sol = s.solve()
    assert sol.status_name == case["status"]
    tol = 1e-3
    if case.get("x") is not None:
        assert np.linalg.norm(sol.x - np.asarray(case["x"])) < tol
    if case.get("obj") is not None:
        assert abs(sol.obj_val - case["obj"]) < tol
    if case.get("obj_dual") is not None:
        assert abs(sol.obj_val_dual - case["obj_dual"]) < tol
    if case["status"].endswith("INFEASIBLE"):
        assert np.isnan(sol.obj_val) and np.isnan(sol.obj_val_dual)
