"""CPU tests of the goal-space cost family (reference: icem/environments/abstract_environments.py:115-123
MaskedGoalSpaceEnvironmentInterface.cost_fn = FetchReach, icem/environments/robotics.py:150-164 FetchPickAndPlace):
the NumPy restatement against the reference's own functions (fixture tests/golden/costs_goal_distance.npz, re-checked
live when /root/reference exists) and the product-side cost function of the stand-in envs."""
import os

import numpy as np
import pytest

from oracle import costs_np
from oracle.make_golden_costs import GOAL_CASES, goal_inputs


def test_oracle_costs_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "costs_goal_distance.npz"))
    for name, _cls, gi, ai, sparse, thr, shaped in GOAL_CASES:
        o = goal_inputs(gi[-1] + 1)
        np.testing.assert_array_equal(costs_np.goal_distance_cost(o, gi, ai, sparse, thr, shaped), g[name], err_msg=name)
        if sparse:      # the thresholds split the samples
            assert 0.05 < np.mean(g[name] >= 1.0) < 0.95, name


@pytest.mark.skipif(not os.path.isdir("/root/reference/icem"), reason="needs the reference sources")
def test_oracle_costs_match_reference_live():
    from oracle.make_golden_costs import reference_goal_costs
    ref = reference_goal_costs()
    for name, _cls, gi, ai, sparse, thr, shaped in GOAL_CASES:
        o = goal_inputs(gi[-1] + 1)
        np.testing.assert_array_equal(costs_np.goal_distance_cost(o, gi, ai, sparse, thr, shaped), ref[name])


def test_env_cost_function_equals_the_oracle():
    from icem_b200 import envs
    for name, _cls, gi, ai, sparse, thr, shaped in GOAL_CASES:
        o = goal_inputs(gi[-1] + 1)
        got = envs.goal_distance_cost_fn(o, None, o, goal_index=gi[0], achieved_index=ai[0], sparse=sparse,
                                         threshold=thr, shaped=shaped)
        np.testing.assert_array_equal(got, costs_np.goal_distance_cost(o, gi, ai, sparse, thr, shaped))
    env = envs.DenseStandInEnv(act_dim=4, bound=1.0, cost="goal_distance", obs_dim=13,
                               cost_params=dict(goal_index=10, achieved_index=0, sparse=False, threshold=0.05))
    assert env.cuda_cost_spec() == ("goal_distance", False, dict(goal_index=10, achieved_index=0, sparse=False,
                                                                  threshold=0.05))
    o = goal_inputs(13)
    np.testing.assert_array_equal(env.cost_fn(o, None, o), costs_np.goal_distance_cost(o, [10, 11, 12], [0, 1, 2]))
