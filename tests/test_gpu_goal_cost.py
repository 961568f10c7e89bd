"""GPU parity of ICEM_COST_GOAL_DISTANCE (reference: environments/abstract_environments.py:115-123,
environments/robotics.py:150-164) on the two batched models through the C ABI: the fp32 dense model (exact to 2e-4)
and the tensor-core MLP rollout (the goal-cost instantiation of csrc/mlp_rollout.cuh, operand-resolution tolerance as
in tests/test_gpu_mlp.py); and a plan step that moves the achieved goal towards the desired one."""
import numpy as np
import pytest

from oracle import costs_np
from oracle.dynamics_np import DenseTanhModel, MlpModel
from oracle.icem_np import reduce_costs

pytestmark = pytest.mark.gpu

CASES = [dict(goal_index=10, achieved_index=0, sparse=False, threshold=0.05, shaped=False),      # FetchReach layout
         dict(goal_index=13, achieved_index=3, sparse=False, threshold=0.05, shaped=True),       # pick-and-place layout
         dict(goal_index=13, achieved_index=3, sparse=True, threshold=0.9, shaped=True)]


def _ref(obs, cp, reduce="sum"):
    gi = list(range(cp["goal_index"], cp["goal_index"] + 3))
    ai = list(range(cp["achieved_index"], cp["achieved_index"] + 3))
    per = costs_np.goal_distance_cost(obs, gi, ai, cp["sparse"], cp["threshold"], cp["shaped"])
    return reduce_costs(np.asarray(per, np.float64), reduce), per


@pytest.mark.parametrize("cp", CASES)
def test_dense_model_costs_match_oracle(cp):
    from icem_b200.planner import Planner, PlannerSettings
    od, d, h, n = 16, 4, 20, 400
    rs = np.random.RandomState(2)
    model = DenseTanhModel(0.9 * np.eye(od) + 0.05 * rs.randn(od, od), 0.3 * rs.randn(od, d), 0.05 * rs.randn(od))
    p = Planner(PlannerSettings(horizon=h, num_simulated_trajectories=64, action_low=-np.ones(d, np.float32),
                                action_high=np.ones(d, np.float32), dynamics="dense_tanh", cost="goal_distance",
                                cost_params=cp, obs_dim=od))
    p.set_dense_model(model.w_obs, model.w_act, model.bias)
    acts = rs.uniform(-1, 1, (n, h, d)).astype(np.float32)
    start = (0.5 * rs.randn(od)).astype(np.float32).astype(np.float64)
    obs = model.rollout(start, acts.astype(np.float64))
    ref, per = _ref(obs, cp)
    got = p.op_rollout_cost(start, acts)
    if cp["sparse"]:
        gi, ai = cp["goal_index"], cp["achieved_index"]
        dist = np.linalg.norm(obs[..., gi:gi + 3] - obs[..., ai:ai + 3], axis=-1)
        eff = np.linalg.norm(obs[..., :3] - obs[..., 3:6], axis=-1)
        safe = np.all(np.abs(dist - cp["threshold"]) > 1e-4, axis=1) & np.all(np.abs(eff - cp["threshold"]) > 1e-4, axis=1)
        assert safe.mean() > 0.9 and 0.05 < np.mean(per >= 1.0) < 0.95
        assert np.abs(got - ref)[safe].max() <= 2e-4
    else:
        assert np.abs(got - ref).max() <= 2e-4, np.abs(got - ref).max()
    p.close()


@pytest.mark.parametrize("cp", CASES[:2])
def test_tensor_core_model_costs_match_oracle(cp):
    from icem_b200 import workloads
    from icem_b200.planner import Planner, PlannerSettings
    od, d, h, n = 16, 4, 12, 600
    ws, bs = workloads.mlp_model_weights(od, d, 128, 9)
    mod = MlpModel(ws, bs)
    p = Planner(PlannerSettings(horizon=h, num_simulated_trajectories=256, action_low=-np.ones(d, np.float32),
                                action_high=np.ones(d, np.float32), dynamics="mlp", cost="goal_distance", cost_params=cp,
                                obs_dim=od, noise_beta=0.25, keep_iteration_actions=True))
    p.set_mlp_model(ws, bs)
    rs = np.random.RandomState(4)
    acts = rs.uniform(-1, 1, (n, h, d)).astype(np.float32)
    start = (0.4 * rs.randn(od)).astype(np.float32).astype(np.float64)
    obs = mod.rollout(start, acts.astype(np.float64))
    ref, _ = _ref(obs, cp)
    got = p.op_rollout_cost(start, acts)
    err = np.abs(got - ref)
    assert np.median(err) <= 1e-2 and err.max() <= 8e-2, (np.median(err), err.max())
    # a plan step with this cost: the best planned cost beats the population's median by a wide margin and the elite
    # costs the device reports are the oracle's for the same action sequences
    p.begin_rollout()
    a = p.plan(start)
    assert a.shape == (d,) and np.all(np.abs(a) <= 1)
    e_acts, e_costs, _ = p.elites()
    ref_e, _ = _ref(mod.rollout(start, e_acts.astype(np.float64)), cp)
    assert np.abs(ref_e - e_costs).max() <= 8e-2
    n0 = p.population_size(0, first_step=True)[1]
    assert e_costs[0] < np.median(p.costs(0, n0)) - 0.1
    p.close()


def test_goal_cost_is_refused_for_ground_truth_models_and_bad_indices():
    from icem_b200.planner import IcemError, Planner, PlannerSettings
    kw = dict(horizon=5, num_simulated_trajectories=8, action_low=-np.ones(6, np.float32), action_high=np.ones(6, np.float32))
    with pytest.raises(IcemError, match="batched model"):
        Planner(PlannerSettings(dynamics="halfcheetah", cost="goal_distance", obs_dim=17,
                                cost_params=dict(goal_index=10, achieved_index=0), **kw))
    with pytest.raises(IcemError, match="indices"):
        Planner(PlannerSettings(dynamics="dense_tanh", cost="goal_distance", obs_dim=12,
                                cost_params=dict(goal_index=10, achieved_index=0), **kw))
