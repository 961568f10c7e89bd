"""GPU parity of the Hopper / Ant additions: generic articulated dynamics id + the locomotion cost
(reads next_obs, so every step is simulated) through the C ABI against the float64 oracle on the same tables, and the
plugin classes on the stand-in envs.  Tolerances as in tests/test_gpu_articulated.py; the cost divides an x difference
by dt (0.008 / 0.05), so a state error of 2e-5 is a per-step cost error of up to 5e-3 (Hopper)."""
import numpy as np
import pytest

from oracle import costs_np
from oracle.articulated_np import make_model
from oracle.icem_np import reduce_costs

pytestmark = pytest.mark.gpu

ENVS = {"hopper": ("Hopper", costs_np.hopper_cost), "ant": ("Ant", costs_np.ant_cost),
        "humanoid": ("Humanoid", costs_np.humanoid_cost)}


def _planner(robot, n=64, **over):
    from icem_b200 import envs
    from icem_b200.planner import Planner, PlannerSettings
    from icem_b200.robots import get_model
    m = get_model(robot)
    env_cls = getattr(envs, ENVS[robot][0])
    kw = dict(horizon=20, num_simulated_trajectories=n, action_low=-m.ctrl_limit * np.ones(m.nu, np.float32),
              action_high=m.ctrl_limit * np.ones(m.nu, np.float32), dynamics="articulated", articulated_model=m, obs_offset=0,
              cost="locomotion", cost_params=dict(env_cls.cost_params, dt=env_cls.dt), obs_dim=m.nq + m.nv,
              opt_iterations=3, factor_decrease_num=1.25, noise_beta=0.25, keep_iteration_actions=True)
    kw.update(over)
    return Planner(PlannerSettings(**kw)), m


@pytest.mark.parametrize("robot", sorted(ENVS))
def test_env_step_matches_oracle(robot):
    p, m = _planner(robot)
    mod = make_model(robot)
    rs = np.random.RandomState(0)
    st = np.concatenate([m.qpos0, 0.1 * rs.randn(m.nv)])
    for t in range(40):
        u = rs.uniform(-1, 1, m.nu) * m.ctrl_limit * (1.5 if t % 7 == 0 else 1.0)
        ref = mod.step_state(st[None], u[None])[0]
        got, obs, _ = p.sim_step(st, u, obs_dim=m.nq + m.nv)
        assert np.abs(got - ref).max() <= 2e-4, (t, np.abs(got - ref).max())
        np.testing.assert_allclose(obs, got, atol=1e-6)
        st = ref
    p.close()


@pytest.mark.parametrize("robot", sorted(ENVS))
@pytest.mark.parametrize("reduce", ["sum", "best"])
def test_rollout_costs_match_oracle(robot, reduce):
    p, m = _planner(robot, cost_along_trajectory=reduce)
    mod = make_model(robot)
    rs = np.random.RandomState(3)
    n, h = 96, 20
    acts = (m.ctrl_limit * rs.uniform(-1, 1, (n, h, m.nu))).astype(np.float32)
    start = np.concatenate([m.qpos0, 0.05 * rs.randn(m.nv)]).astype(np.float32).astype(np.float64)
    obs = mod.rollout(start, acts.astype(np.float64), with_final=True)
    if robot == "hopper":
        obs[..., 6:] = np.clip(obs[..., 6:], -10, 10)          # gym's Hopper observation clips the velocities
    per_step = ENVS[robot][1](obs[:, :-1], acts.astype(np.float64), obs[:, 1:])
    ref = reduce_costs(per_step, reduce)
    got = p.op_rollout_cost(start, acts)
    d = np.abs(got - ref)
    # a trajectory whose height crosses the healthy threshold within fp32 error of it flips a 100 / 200 penalty
    z = obs[:, :-1, 1 if robot == "hopper" else 2]
    lo = {"hopper": 0.7, "ant": 0.2, "humanoid": 2.0}[robot]
    safe = np.all(np.abs(z - lo) > 5e-3, axis=1) & np.all(np.abs(z - 1.0) > 5e-3, axis=1)
    assert safe.mean() > 0.7
    assert np.median(d[safe]) <= 2e-3, np.median(d[safe])
    # the summed velocity term telescopes to (x_h - x_0) / dt: the final-position error is amplified 1/dt = 125x
    # (Hopper) / 20x (Ant), and the Ant's 100 substeps of contact switching spread fp32 rounding to ~1e-2 in x
    tol = 5e-2 if robot == "hopper" else 0.3
    assert np.mean(d[safe] <= tol) >= 0.95, np.sort(d[safe])[-8:]
    p.close()


def test_unhealthy_penalty_and_state_bound():
    """The unhealthy branch of the Hopper cost (mujoco.py:196-212, 227): a height range nothing satisfies penalises
    every step by 200; a state bound below the joint angles does the same; both equal the oracle."""
    from icem_b200 import envs
    mod = make_model("hopper")
    rs = np.random.RandomState(5)
    acts = rs.uniform(-1, 1, (32, 20, 3)).astype(np.float32)
    for over, op in ((dict(z_lo=2.0), dict(costs_np.HOPPER, healthy_z_range=(2.0, float("inf")))),
                     (dict(state_bound=1e-4), dict(costs_np.HOPPER, healthy_state_range=(-1e-4, 1e-4)))):
        p, m = _planner("hopper", cost_params=dict(envs.Hopper.cost_params, dt=envs.Hopper.dt, **over))
        start = np.concatenate([m.qpos0, 0.05 * rs.randn(m.nv)]).astype(np.float32).astype(np.float64)
        obs = mod.rollout(start, acts.astype(np.float64), with_final=True)
        obs[..., 6:] = np.clip(obs[..., 6:], -10, 10)          # gym's Hopper observation clips the velocities
        per_step = costs_np.hopper_cost(obs[:, :-1], acts.astype(np.float64), obs[:, 1:], op)
        assert np.all(per_step > 150)
        got = p.op_rollout_cost(start, acts)
        assert np.abs(got - per_step.sum(1)).max() <= 5e-2
        p.close()


@pytest.mark.parametrize("env_name", ["Hopper", "Ant", "Humanoid"])
def test_controller_on_locomotion_envs(env_name):
    from icem_b200 import envs
    from icem_b200.controller import MpcICemB200
    from icem_b200.models import CudaGroundTruthModel
    env = envs.make_env(env_name)
    env.seed(4)
    ctrl = MpcICemB200(env=env, forward_model=CudaGroundTruthModel(env=env), horizon=20,
                       num_simulated_trajectories=256, factor_decrease_num=1.25, cost_along_trajectory="sum", seed=2,
                       action_sampler_params=dict(alpha=0.1, elites_size=10, fraction_elites_reused=0.3, init_std=0.5,
                                                  keep_previous_elites=True, shift_elites_over_time=True,
                                                  use_mean_actions=True, opt_iterations=3, noise_beta=0.25))
    ob = env.reset()
    ctrl.beginning_of_rollout(observation=ob, state=env.get_GT_state(), mode="train")
    x0 = env.get_GT_state()[1]
    ret = 0.0
    for t in range(15):
        ac = ctrl.get_action(ob, state=env.get_GT_state())
        assert np.all(np.abs(ac) <= env.action_space.high + 1e-6)
        if t == 3:         # the elites' lazily materialised rollouts reproduce the planner's costs
            el = ctrl.elite_samples
            _, costs, _ = ctrl._planner.elites()
            assert el.as_array("observations").shape == (10, 20, env.observation_space.shape[0])
            np.testing.assert_allclose(-np.sum(el[0]["rewards"]), costs[0], atol=5e-2, rtol=1e-3)
        ob, r, _, _ = env.step(ac)
        ret += r
    assert np.isfinite(ret)
    assert env.get_GT_state()[1] - x0 > -0.05       # planning for forward velocity does not walk backwards
    ctrl.close()
    env.close()
