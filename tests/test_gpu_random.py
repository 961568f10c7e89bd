"""GPU parity of the random-shooting planner mode (reference: icem/controllers/mpc.py::MpcRandom)
through the C ABI, on the oracle's exact uniform draws and against the fixtures recorded from the reference class
(tests/golden/random_*.npz).

Tolerances: actions |d| <= 2e-7 * (high - low) + 1 float32 ulp (one fp32 fma vs float64 then cast); costs 2e-4;
the best index is exact whenever the oracle's gap between the two cheapest trajectories exceeds 20x that."""
import os

import numpy as np
import pytest

from oracle import cases, costs_np
from oracle.random_np import RandomConfig, RandomOracle
from tests.util import elite_gap

pytestmark = pytest.mark.gpu
COST_TOL = 2e-4


def _planner(case, c, model, **over):
    from icem_b200.planner import Planner, PlannerSettings
    kw = dict(horizon=c["horizon"], num_simulated_trajectories=c["num_simulated_trajectories"],
              action_low=c["action_low"], action_high=c["action_high"], dynamics="dense_tanh", cost=case["cost"],
              obs_dim=model.obs_dim, penalise_flipping=case["penalise_flipping"],
              cost_along_trajectory=c["cost_along_trajectory"], opt_iterations=1, elites_size=2, planner="random",
              action_change_frequency=c["action_change_frequency"], keep_iteration_actions=True)
    kw.update(over)
    p = Planner(PlannerSettings(**kw))
    p.set_dense_model(model.w_obs, model.w_act, model.bias)
    return p


@pytest.mark.parametrize("name", sorted(cases.RANDOM_CASES))
def test_random_plan_steps_match_oracle_and_reference_golden(name, golden_dir):
    case = cases.RANDOM_CASES[name]
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    c = cases.random_config(case)
    model = case["model"]()
    p = _planner(case, c, model)
    cost = lambda o, a: costs_np.halfcheetah_cost(o, a, case["penalise_flipping"])
    np.random.seed(case["seed"])
    orc = RandomOracle(RandomConfig(**c), model.rollout, cost, record_actions=True)
    obs = np.asarray(case["start_obs"], np.float64).copy()
    p.begin_rollout()
    n = c["num_simulated_trajectories"]
    span = (c["action_high"] - c["action_low"]).astype(np.float64)
    for s in range(case["steps"]):
        tr = orc.get_action(obs)
        it = tr.iterations[0]
        p.inject_noise(0, it.noise[0][0], None)
        act = p.plan(obs)
        assert p.population_size(0, first_step=(s == 0)) == (n, n)
        a_dev = p.actions(0, n)
        assert np.all(np.abs(a_dev - it.actions) <= 2e-7 * span + 1.2e-7), np.abs(a_dev - it.actions).max()
        assert np.all(np.abs(a_dev - g[f"s{s}_i0_actions"]) <= 2e-7 * span + 1.2e-7)
        assert np.all(a_dev >= c["action_low"]) and np.all(a_dev <= c["action_high"])
        assert np.abs(p.costs(0, n) - it.costs).max() <= COST_TOL
        if elite_gap(it.costs, 1) > 20 * COST_TOL:
            assert p.iteration_record(0)["elite_idx"][0] == it.elite_idx[0] == g[f"s{s}_i0_elite_idx"][0]
            assert np.abs(act - tr.action).max() <= 2e-7 * span.max() + 1.2e-7
            assert np.abs(act - g[f"s{s}_action"]).max() <= 2e-7 * span.max() + 1.2e-7
        obs = model.step(obs[None], tr.action[None])[0]
    p.close()


def test_random_production_draws_are_piecewise_constant_uniform():
    """Philox mode: the population is the flattened call sequence of MpcRandom.sample(): segment 0 (the action drawn
    at construction) serves `freq` calls, every later draw freq + 1 calls, and the sequence runs on across plan
    steps and across beginning_of_rollout (mpc.py:95-107); values are uniform inside the Box."""
    case = cases.RANDOM_CASES["random_cheetah"]
    c = cases.random_config(case)
    model = case["model"]()
    n, h, f = 4096, c["horizon"], 4
    p = _planner(case, c, model, num_simulated_trajectories=n, action_change_frequency=f, seed=11)
    obs = np.asarray(case["start_obs"], np.float64)
    p.begin_rollout()
    flat = []
    for step in range(2):
        if step == 1:
            p.begin_rollout()
        act = p.plan(obs)
        a = p.actions(0, n)
        flat.append(a.reshape(n * h, -1))
        best = int(np.argmin(p.costs(0, n)))
        np.testing.assert_array_equal(act.astype(np.float32), a[best, 0])
    flat = np.concatenate(flat)
    calls = np.arange(len(flat))
    seg = np.where(calls < f, 0, 1 + (calls - f) // (f + 1))
    change = np.any(flat[1:] != flat[:-1], axis=1)
    np.testing.assert_array_equal(change, seg[1:] != seg[:-1])
    vals = flat[np.concatenate([[True], change])]              # one row per drawn action
    assert len(vals) == seg[-1] + 1
    assert np.all(vals >= -1) and np.all(vals <= 1)
    assert np.abs(vals.mean(0)).max() < 0.02 and np.abs(vals.var(0) - 1 / 3).max() < 0.02
    assert abs(np.corrcoef(vals[1:, 0], vals[:-1, 0])[0, 1]) < 0.02
    assert abs(np.corrcoef(vals[:, 0], vals[:, 1])[0, 1]) < 0.02
    p.close()


def test_random_controller_class_mirrors_reference(capsys):
    from icem_b200 import envs
    from icem_b200.controller import MpcRandomB200
    from icem_b200.models import CudaGroundTruthModel
    env = envs.make_env("HalfCheetah")
    env.seed(3)
    ctrl = MpcRandomB200(env=env, forward_model=CudaGroundTruthModel(env=env), horizon=30,
                         num_simulated_trajectories=2048, cost_along_trajectory="sum", seed=5,
                         action_sampler_params=dict(action_change_frequency=5))
    ob = env.reset()
    ctrl.beginning_of_rollout(observation=ob, state=env.get_GT_state(), mode="train")
    x0 = env.get_GT_state()[1]
    for _ in range(25):
        ac = ctrl.get_action(ob, state=env.get_GT_state())
        assert ac.shape == (6,) and np.all(np.abs(ac) <= 1.0)
        ob, _, _, _ = env.step(ac)
    assert env.get_GT_state()[1] - x0 > 0.0          # random shooting with 2048 candidates still moves forward
    ctrl.end_of_rollout(0.0, 0.0, "train")
    with pytest.raises(AssertionError):               # mpc.py:92
        MpcRandomB200(env=env, forward_model=CudaGroundTruthModel(env=env), horizon=5,
                      num_simulated_trajectories=64, cost_along_trajectory="sum",
                      action_sampler_params=dict(action_change_frequency=5))
    ctrl.close()
    env.close()
