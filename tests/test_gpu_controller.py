"""GPU tests of the plugin boundary: MpcICemB200 driven exactly like RolloutManager._sample drives a controller
(icem/misc/rollout_utils.py:155-227) -- beginning_of_rollout / get_action / env.step / end_of_rollout -- on the
device-simulated stand-in environments, without the reference package (absent on the GPU box)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SAMPLER = dict(alpha=0.1, elites_size=10, fraction_elites_reused=0.3, init_std=0.5, keep_previous_elites=True,
               shift_elites_over_time=True, use_mean_actions=True, opt_iterations=3, noise_beta=0.25)


def _make(env_name="HalfCheetah", n=128, **over):
    from icem_b200 import envs
    from icem_b200.controller import MpcICemB200
    from icem_b200.models import CudaGroundTruthModel
    env = envs.make_env(env_name)
    env.seed(1)
    model = CudaGroundTruthModel(env=env, num_parallel=8)
    kw = dict(env=env, forward_model=model, horizon=30, num_simulated_trajectories=n, factor_decrease_num=1.25,
              cost_along_trajectory="sum", do_visualize_plan=False, verbose=False,
              action_sampler_params=dict(SAMPLER, noise_beta=0.25 if env_name == "HalfCheetah" else 2.0), seed=5)
    kw.update(over)
    return MpcICemB200(**kw), env


@pytest.mark.parametrize("env_name", ["HalfCheetah", "HumanoidStandup"])
def test_rollout_loop_like_rollout_manager(env_name, capsys):
    ctrl, env = _make(env_name)
    assert ctrl.has_state and not ctrl.needs_data and not ctrl.needs_training
    with pytest.raises(AttributeError, match=r"beginning_of_rollout\(\) needs to be called before"):
        ctrl.get_action(np.zeros(env.observation_space.shape[0]), state=None)
    ob = env.reset_with_mode("train")
    ctrl.beginning_of_rollout(observation=ob, state=env.get_GT_state(), mode="train")
    assert "iCEM using" in capsys.readouterr().out                      # controllers/icem.py:42-43
    assert ctrl.model_evals_per_timestep == (128 + 102 + 81) * 30
    assert len(ctrl.elite_samples) == 0 and not ctrl.elite_samples      # empty RolloutBuffer before the first step
    h, d = 30, env.action_space.shape[0]
    np.testing.assert_allclose(ctrl.mean, 0.0)
    np.testing.assert_allclose(ctrl.std, 0.5 * env.action_space.high[0], rtol=1e-6)
    ret = 0.0
    for t in range(6):
        state = env.get_GT_state()
        ac = ctrl.get_action(ob, state=state, mode="train")
        assert ac.shape == (d,) and ac.dtype == np.float64
        assert np.all(ac >= env.action_space.low - 1e-6) and np.all(ac <= env.action_space.high + 1e-6)
        ob, rew, done, _ = env.step(ac)
        ret += rew
        assert ctrl.mean.shape == (h, d) and ctrl.std.shape == (h, d)
        el = ctrl.elite_samples
        assert len(el) == 10 and el[0]["actions"].shape == (h, d)
        np.testing.assert_allclose(el[0]["actions"][0], ac, atol=1e-6)   # executed = first action of the best elite
        np.testing.assert_allclose(ctrl.std, 0.5 * env.action_space.high[0], rtol=1e-6)   # std reset (icem.py:175)
    ctrl.end_of_rollout(total_time=6, total_return=ret, mode="train")
    assert np.isfinite(ret)
    # a new rollout resets the distribution and the elites
    ctrl.beginning_of_rollout(observation=env.reset(), state=env.get_GT_state(), mode="train")
    np.testing.assert_allclose(ctrl.mean, 0.0)
    assert len(ctrl.elite_samples) == 0
    ctrl.close()
    env.close()


def test_constructor_errors_mirror_reference():
    from icem_b200 import envs
    from icem_b200.controller import MpcICemB200
    with pytest.raises(ValueError, match="At least two trajectories needed!"):            # controllers/mpc.py:30-31
        _make(n=1)
    with pytest.raises(NotImplementedError, match="to compute cost along trajectory"):    # abstract_controller.py:88-91
        _make(cost_along_trajectory="median")
    with pytest.warns(UserWarning, match="Setting num_elites to 2"):                      # controllers/icem.py:238-240
        c, e = _make(n=3)
        c.close(); e.close()

    class NotCuda:
        pass
    env = envs.make_env("HalfCheetah")
    with pytest.raises(TypeError, match="no CPU fallback"):
        MpcICemB200(env=env, forward_model=NotCuda(), horizon=5, num_simulated_trajectories=8,
                    cost_along_trajectory="sum", action_sampler_params=SAMPLER)


def test_verbose_prints_reference_format(capsys):
    ctrl, env = _make(verbose=True)
    ob = env.reset()
    ctrl.beginning_of_rollout(observation=ob, state=env.get_GT_state(), mode="train")
    ctrl.get_action(ob, state=env.get_GT_state())
    out = capsys.readouterr().out
    assert "iter 0:128 --- best cost:" in out and "iter 2:81 --- best cost:" in out     # controllers/icem.py:156-158
    ctrl.close()
    env.close()


@pytest.mark.parametrize("env_name", ["HalfCheetah", "HumanoidStandup"])
def test_elite_samples_carry_the_rollout_fields_of_the_reference(env_name):
    """`elite_samples` are full rollouts (observations, next_observations, actions, rewards) like the
    reference's (models/abstract_models.py:28-53), materialised lazily from the device model: replaying an elite's
    actions in the env from the same state visits the same observations, and the summed step costs are the elite's
    planner cost."""
    ctrl, env = _make(env_name)
    ob = env.reset()
    ctrl.beginning_of_rollout(observation=ob, state=env.get_GT_state(), mode="train")
    for _ in range(2):
        state = env.get_GT_state()
        ac = ctrl.get_action(ob, state=state)
        el = ctrl.elite_samples
        assert el is ctrl.elite_samples                       # cached until the next plan step
        obs_dim = env.observation_space.shape[0]
        assert el.as_array("observations").shape == (10, 30, obs_dim)
        assert el.as_array("next_observations").shape == (10, 30, obs_dim)
        np.testing.assert_allclose(el[0]["observations"][0], ob, atol=1e-5)
        np.testing.assert_array_equal(el[3]["observations"][1:], el[3]["next_observations"][:-1])
        _, costs, _ = ctrl._planner.elites()
        for j in (0, 9):
            np.testing.assert_allclose(-np.sum(el[j]["rewards"]), costs[j], atol=2e-3, rtol=1e-4)
        # replay elite 0 in a second env instance from the same state
        twin = type(env)(name=env.name)
        twin.reset()
        twin.set_GT_state(state)
        for t in range(5):
            o, *_ = twin.step(el[0]["actions"][t])
            np.testing.assert_allclose(o, el[0]["next_observations"][t], atol=2e-4, rtol=1e-4)
        twin.close()
        ob, *_ = env.step(ac)
    ctrl.close()
    env.close()


def test_do_visualize_plan_gets_the_best_planned_trajectory(monkeypatch):
    """icem.py:179-183: with do_visualize_plan set, visualize_plan receives the best trajectory's observations and
    actions and the model state."""
    ctrl, env = _make("HalfCheetah", do_visualize_plan="all")
    seen = {}
    monkeypatch.setattr(type(ctrl), "visualize_plan", lambda self, *, obs, state, acts: seen.update(
        obs=np.array(obs), state=np.array(state), acts=np.array(acts)), raising=False)
    ob = env.reset()
    ctrl.beginning_of_rollout(observation=ob, state=env.get_GT_state(), mode="train")
    st = env.get_GT_state()
    ac = ctrl.get_action(ob, state=st)
    assert seen["obs"].shape == (30, 17) and seen["acts"].shape == (30, 6)
    np.testing.assert_allclose(seen["acts"][0], ac, atol=1e-6)
    np.testing.assert_allclose(seen["obs"][0], ob, atol=1e-5)
    np.testing.assert_array_equal(seen["state"], st)
    ctrl.close()
    env.close()


def test_use_env_reward_as_cost_is_the_cost_path_on_envs_that_declare_it():
    """abstract_controller.py:75-76: costs = -rewards.  The stand-in envs return reward = -cost_fn, so the flag
    selects the same numbers; an env that does not declare that is refused (no silent substitution)."""
    a, env_a = _make("HalfCheetah", use_env_reward_as_cost=True)
    b, env_b = _make("HalfCheetah")
    ob = env_a.reset()
    env_b.set_GT_state(env_a.get_GT_state())
    for c, e in ((a, env_a), (b, env_b)):
        c.beginning_of_rollout(observation=ob, state=e.get_GT_state(), mode="train")
    np.testing.assert_array_equal(a.get_action(ob, state=env_a.get_GT_state()),
                                  b.get_action(ob, state=env_b.get_GT_state()))
    env_a.reward_is_negative_cost = False
    with pytest.raises(NotImplementedError, match="reward_is_negative_cost"):
        from icem_b200.controller import MpcICemB200
        from icem_b200.models import CudaGroundTruthModel
        MpcICemB200(env=env_a, forward_model=CudaGroundTruthModel(env=env_a), horizon=30,
                    num_simulated_trajectories=16, cost_along_trajectory="sum", use_env_reward_as_cost=True,
                    action_sampler_params=SAMPLER)
    for c, e in ((a, env_a), (b, env_b)):
        c.close(); e.close()
