"""GPU: problems with nonsymmetric cones through the B200 backend.  The exp / pow / genpow blocks
are host-computed and reach the device through update_values!; everything else
(regularisation, factorisation, solves, refinement) is the same CUDA path as for symmetric cones."""
import numpy as np
import pytest

from test_nonsymmetric_cones import (_exp_problem, _pow_problem, _genpow_problem, _mixed_problem,
                                      _sdp_chordal_problem)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("make", [_exp_problem, _pow_problem, _genpow_problem, _mixed_problem, _sdp_chordal_problem])
def test_nonsymmetric_problems_match_the_oracle(cb, make):
    so = cb.Solver(*make(cb), cb.Settings(direct_solve_method="qdldl")).solve()
    sg_solver = cb.Solver(*make(cb), cb.Settings(direct_solve_method="b200"))
    sg = sg_solver.solve()
    assert sg.status_name == so.status_name == "SOLVED"
    # the backtracking line searches of these cones take discrete steps (0.8^k): a last-bit
    # difference in a KKT solve may move one of them, so the iteration count gets some slack
    assert abs(sg.iterations - so.iterations) <= 2
    assert abs(sg.obj_val - so.obj_val) <= 1e-6 * max(1.0, abs(so.obj_val))
    assert np.allclose(sg.x, so.x, rtol=1e-5, atol=1e-5)
    assert sg_solver.kktsystem.kktsolver.ldl.timers()["nlaunch"] > 0


def test_lasso_shape_instance_matches_the_oracle(cb):
    """test/OptTests/socp-lasso.jl shape (one 402-dimensional expanded SOC) through the B200 backend."""
    from clarabel_jl_b200 import problems
    so = cb.Solver(*problems.socp_lasso(), cb.Settings(direct_solve_method="qdldl")).solve()
    sg = cb.Solver(*problems.socp_lasso(), cb.Settings(direct_solve_method="b200")).solve()
    assert sg.status_name == so.status_name == "SOLVED" and sg.iterations == so.iterations
    assert abs(sg.obj_val - so.obj_val) <= 1e-6 * max(1.0, abs(so.obj_val))
