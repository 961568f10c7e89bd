"""Episode-level parallelism on the device (reference: misc/rollout_utils.py:129-152): many planner
handles launched before any is collected give bit-identical results to planning the same episodes one by one, and
the asynchronous halves of the C ABI keep their state machine."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PARAMS = dict(horizon=30, num_simulated_trajectories=40, factor_decrease_num=1.25, cost_along_trajectory="sum",
              action_sampler_params=dict(alpha=0.1, elites_size=10, fraction_elites_reused=0.3, init_std=0.5,
                                         keep_previous_elites=True, shift_elites_over_time=True,
                                         use_mean_actions=True, opt_iterations=3, noise_beta=0.25))


def test_batched_episodes_equal_sequential_episodes(capsys):
    from icem_b200.batched import make_episode_batch
    B, T = 6, 4
    batch = make_episode_batch("HalfCheetah", B, PARAMS, seed=3)
    eps = batch.run(T)
    solo = make_episode_batch("HalfCheetah", B, PARAMS, seed=3)
    for i in range(B):                       # same episodes, planned one at a time through get_action
        env, ctrl = solo.envs[i], solo.controllers[i]
        ob = env.reset_with_mode("train")
        ctrl.beginning_of_rollout(observation=ob, state=env.get_GT_state(), mode="train")
        for t in range(T):
            a = ctrl.get_action(ob, state=env.get_GT_state())
            np.testing.assert_array_equal(a, eps[i]["actions"][t])
            np.testing.assert_array_equal(ob, eps[i]["observations"][t])
            ob, r, _, _ = env.step(a)
            assert r == eps[i]["rewards"][t]
    assert eps[0]["actions"].shape == (T, 6)
    assert not np.array_equal(eps[0]["actions"], eps[1]["actions"])      # independent episodes
    batch.close()
    solo.close()


def test_async_halves_state_machine():
    from icem_b200 import workloads
    from icem_b200.planner import IcemError, Planner
    name = "halfcheetah_gt_n4096"
    p = Planner(workloads.planner_settings(name, scale_population=1 / 32))
    st = workloads.start_state(name)
    p.begin_rollout()
    with pytest.raises(IcemError, match="no plan step in flight"):
        p.plan_finish()
    p.plan_async(st)
    with pytest.raises(IcemError, match="already in flight"):
        p.plan_async(st)
    a = p.plan_finish()
    q = Planner(workloads.planner_settings(name, scale_population=1 / 32))
    q.begin_rollout()
    np.testing.assert_array_equal(a, q.plan(st))
    p.close()
    q.close()
