"""GPU parity of the vanilla-CEM planner mode (reference: icem/controllers/mpc.py::MpcCemStd) through
the C ABI: truncated-normal sampling on the oracle's exact uniform draws, elite lists bit-exact per iteration, the
`bounds_like_levine` / `execute_best_elite` / `shift_means` switches, and the fixtures recorded from the UNMODIFIED
reference (tests/golden/cemstd_*.npz).

Tolerances: sampled actions |d| <= 2e-5 * (high - low) (fp32 normcdf / normcdfinv vs scipy's float64 log-space
quantile); costs 3e-4; mean / std 1e-5; elite indices exact whenever the oracle's gaps exceed 20x the cost
tolerance."""
import os

import numpy as np
import pytest

from oracle import cases, costs_np
from oracle.cem_std_np import CemStdConfig, CemStdOracle
from tests.util import elite_gap, stack_noise

pytestmark = pytest.mark.gpu
COST_TOL = 3e-4


def _setup(case):
    from icem_b200.planner import Planner, PlannerSettings
    c = cases.cem_std_config(case)
    model = case["model"]()
    p = Planner(PlannerSettings(
        horizon=c["horizon"], num_simulated_trajectories=c["num_simulated_trajectories"], action_low=c["action_low"],
        action_high=c["action_high"], dynamics="dense_tanh", cost=case["cost"], obs_dim=model.obs_dim,
        penalise_flipping=case["penalise_flipping"], cost_along_trajectory=c["cost_along_trajectory"],
        alpha=c["alpha"], elites_size=c["elites_size"], opt_iterations=c["opt_iterations"], init_std=c["init_std"],
        planner="cem_std", execute_best_elite=c["execute_best_elite"], shift_means=c["shift_means"],
        bounds_like_levine=c["bounds_like_levine"], keep_iteration_actions=True))
    p.set_dense_model(model.w_obs, model.w_act, model.bias)
    cost = (lambda o, a: costs_np.halfcheetah_cost(o, a, case["penalise_flipping"])) \
        if case["cost"] == "halfcheetah" else costs_np.humanoid_standup_cost
    orc = CemStdOracle(CemStdConfig(**c), model.rollout, cost, record_actions=True)
    return p, orc, model, c


@pytest.mark.parametrize("name", sorted(cases.CEM_STD_CASES))
def test_cem_std_plan_steps_match_oracle_and_reference_golden(name, golden_dir):
    case = cases.CEM_STD_CASES[name]
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    p, orc, model, c = _setup(case)
    np.random.seed(case["seed"])
    obs = np.asarray(case["start_obs"], np.float64).copy()
    orc.beginning_of_rollout()
    p.begin_rollout()
    np.testing.assert_allclose(p.std(), orc.std, atol=1e-7)              # levine clamp at reset
    k = orc.cfg.num_elites
    span = (orc.cfg.action_high - orc.cfg.action_low).astype(np.float64)
    n = c["num_simulated_trajectories"]
    for s in range(case["steps"]):
        tr = orc.get_action(obs)
        for i, it in enumerate(tr.iterations):
            u, _ = stack_noise(it.noise)
            p.inject_noise(i, u, None)
        act = p.plan(obs)
        for i, it in enumerate(tr.iterations):
            rec = p.iteration_record(i)
            assert p.population_size(i, first_step=(s == 0)) == (n, n)    # no decay, no extra rows
            a_dev = p.actions(i, n)
            assert np.all(np.abs(a_dev - it.actions) <= 2e-5 * span + 1e-6), np.abs(a_dev - it.actions).max()
            assert np.all(a_dev >= orc.cfg.action_low - 1e-6) and np.all(a_dev <= orc.cfg.action_high + 1e-6)
            assert np.abs(p.costs(i, n) - it.costs).max() <= COST_TOL
            if elite_gap(it.costs, k) > 20 * COST_TOL:
                np.testing.assert_array_equal(rec["elite_idx"], it.elite_idx)
                np.testing.assert_array_equal(rec["elite_idx"], g[f"s{s}_i{i}_elite_idx"])
            assert np.abs(rec["mean"] - it.mean).max() <= 1e-5
            assert np.abs(rec["std"] - it.std).max() <= 1e-5
        assert np.abs(act - tr.action).max() <= 2e-5 * span.max() + 1e-6
        assert np.abs(act - g[f"s{s}_action"]).max() <= 2e-5 * span.max() + 1e-6
        assert np.abs(p.mean() - tr.mean_after_shift).max() <= 1e-5
        assert np.abs(p.std() - tr.std_after_reset).max() <= 1e-6
        obs = model.step(obs[None], tr.action[None])[0]
    p.close()


def test_cem_std_controller_class_on_ground_truth_env(capsys):
    """MpcCemStdB200 driven like RolloutManager drives a controller, production (Philox) uniforms, HalfCheetah GT."""
    from icem_b200 import envs
    from icem_b200.controller import MpcCemStdB200
    from icem_b200.models import CudaGroundTruthModel
    env = envs.make_env("HalfCheetah")
    env.seed(2)
    ctrl = MpcCemStdB200(env=env, forward_model=CudaGroundTruthModel(env=env), horizon=30,
                         num_simulated_trajectories=256, cost_along_trajectory="sum", seed=9,
                         action_sampler_params=dict(alpha=0.1, elites_size=10, opt_iterations=3, init_std=0.5,
                                                    shift_means=True, execute_best_elite=True,
                                                    bounds_like_levine=False))
    ob = env.reset()
    ctrl.beginning_of_rollout(observation=ob, state=env.get_GT_state(), mode="train")
    assert "CEM-Standard using 23040 evaluations per step" in capsys.readouterr().out     # mpc.py:174-175
    x0 = env.get_GT_state()[1]
    for _ in range(20):
        ac = ctrl.get_action(ob, state=env.get_GT_state())
        assert np.all(np.abs(ac) <= 1.0 + 1e-6)
        ob, _, _, _ = env.step(ac)
    assert env.get_GT_state()[1] - x0 > 0.3          # the vanilla-CEM baseline also drives the cheetah forward
    assert len(ctrl.elite_samples) == 10
    ctrl.close()
    env.close()
