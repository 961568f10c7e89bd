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


@pytest.mark.parametrize("name,scale", [("halfcheetah_gt_n4096", 1 / 32), ("dense_tanh_cheetah_n4096", 1 / 64),
                                        ("humanoid_standup_gt_n16384", 1 / 256)])
def test_multi_problem_handle_equals_single_problem_handles(name, scale):
    """icem_plan_batch: problem i of a num_problems = B handle plans bit-identically to a single-problem handle
    created with seed + i, over several closed-loop steps (shifted / kept elites, per-problem distributions)."""
    import dataclasses
    from icem_b200 import workloads
    from icem_b200.planner import IcemError, Planner
    B = 5
    w = workloads.get_workload(name)

    def make(**over):
        s = dataclasses.replace(workloads.planner_settings(name, scale_population=scale, seed=7), **over)
        p = Planner(s)
        if w.get("dense"):
            p.set_dense_model(*workloads.dense_model_weights(*w["dense"]))
        p.begin_rollout()
        return p
    batch = make(num_problems=B)
    singles = [make(seed=7 + i) for i in range(B)]
    states = np.stack([workloads.start_state(name, seed=i) for i in range(B)])
    with pytest.raises(IcemError, match="icem_plan_batch"):
        batch.plan(states[0])
    for step in range(3):
        acts = batch.plan_batch(states)
        assert acts.shape == (B, w["act_dim"])
        for i, p in enumerate(singles):
            a = p.plan(states[i])
            np.testing.assert_array_equal(acts[i], a)
            batch.set_active_problem(i)
            np.testing.assert_array_equal(batch.mean(), p.mean())
            np.testing.assert_array_equal(batch.elites()[0], p.elites()[0])
            np.testing.assert_array_equal(batch.iteration_record(1)["elite_idx"], p.iteration_record(1)["elite_idx"])
            states[i] = p.sim_step(states[i], a)[0]
        assert not np.array_equal(acts[0], acts[1])
    batch.close()
    for p in singles:
        p.close()


def test_fused_episode_batch_runs_like_the_streamed_one():
    from icem_b200.batched import make_episode_batch, make_fused_episode_batch
    B, T = 4, 3
    fused = make_fused_episode_batch("HalfCheetah", B, PARAMS, seed=3)
    eps = fused.run(T)
    streamed = make_episode_batch("HalfCheetah", B, PARAMS, seed=3)
    ref = streamed.run(T)
    for i in range(B):           # same env seeds, planner seeds 1003 + i in both
        np.testing.assert_array_equal(eps[i]["actions"], ref[i]["actions"])
        np.testing.assert_array_equal(eps[i]["rewards"], ref[i]["rewards"])
    fused.close()
    streamed.close()


def test_batched_transitions_equal_single_transitions():
    """icem_sim_step_batch: n transitions in one launch == n icem_sim_step calls, bit for bit."""
    from icem_b200 import workloads
    from icem_b200.planner import Planner
    name = "humanoid_standup_gt_n16384"
    p = Planner(workloads.planner_settings(name, scale_population=1 / 256))
    rs = np.random.RandomState(0)
    states = np.stack([workloads.start_state(name, seed=i) for i in range(7)])
    acts = rs.uniform(-0.4, 0.4, (7, 17))
    got = p.sim_step_batch(states, acts)
    for i in range(7):
        np.testing.assert_array_equal(got[i], p.sim_step(states[i], acts[i])[0])
    p.close()
