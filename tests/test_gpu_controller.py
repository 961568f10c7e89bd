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
