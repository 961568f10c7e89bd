"""GPU parity of the Reacher addition (reference icem/environments/mujoco.py:346-368): the contact-free
two-root robot table (arm + target on undriven slides) and ICEM_COST_REACHER through the C ABI against the float64
oracle, and MpcICemB200 on the stand-in env.  Tolerances: one env step |d| <= 2e-4 (velocities reach 20 rad/s: gear
200 on armature 1), h = 20 costs |d| <= 2e-4 median / 2e-3 at the 95th percentile."""
import numpy as np
import pytest

from oracle import costs_np
from oracle.articulated_np import make_model
from oracle.icem_np import reduce_costs

pytestmark = pytest.mark.gpu


def _planner(n=64, **over):
    from icem_b200.planner import Planner, PlannerSettings
    from icem_b200.robots import get_model
    m = get_model("reacher")
    kw = dict(horizon=20, num_simulated_trajectories=n, action_low=-np.ones(2, np.float32),
              action_high=np.ones(2, np.float32), dynamics="articulated", articulated_model=m, obs_offset=0,
              cost="reacher", cost_params=dict(reach=(0.1, 0.11, 0.0, 0.0)), obs_dim=11, opt_iterations=3,
              factor_decrease_num=1.25, noise_beta=2.0, keep_iteration_actions=True)
    kw.update(over)
    return Planner(PlannerSettings(**kw)), m


def _start(rs):
    return np.concatenate([rs.uniform(-1.5, 1.5, 2), rs.uniform(-0.15, 0.15, 2), rs.uniform(-0.5, 0.5, 2), np.zeros(2)])


def test_env_step_matches_oracle():
    p, m = _planner()
    mod = make_model("reacher", obs_skip=0)
    rs = np.random.RandomState(0)
    st = _start(rs)
    for t in range(40):
        u = rs.uniform(-1, 1, 2) * (1.5 if t % 7 == 0 else 1.0)          # also beyond the control range (clipped)
        ref = mod.step_state(st[None], np.clip(u, -1, 1)[None])[0]
        got, _, _ = p.sim_step(st, u)
        assert np.abs(got - ref).max() <= 2e-4, (t, np.abs(got - ref).max())
        np.testing.assert_array_equal(got[2:4], st[2:4].astype(np.float32).astype(np.float64))    # the target stays put
        st = ref
    p.close()


@pytest.mark.parametrize("reduce", ["sum", "final"])
def test_rollout_costs_match_oracle(reduce):
    p, m = _planner(cost_along_trajectory=reduce)
    mod = make_model("reacher", obs_skip=0)
    rs = np.random.RandomState(3)
    n, h = 200, 20
    acts = rs.uniform(-1, 1, (n, h, 2)).astype(np.float32)
    start = _start(rs).astype(np.float32).astype(np.float64)
    states = mod.rollout(start, acts.astype(np.float64))
    per_step = costs_np.reacher_cost(costs_np.reacher_observation(states))
    ref = reduce_costs(per_step, reduce)
    got = p.op_rollout_cost(start, acts)
    d = np.abs(got - ref)
    assert np.median(d) <= 2e-4, np.median(d)
    assert np.mean(d <= 2e-3) >= 0.95, np.sort(d)[-8:]
    p.close()


def test_plan_steps_select_the_oracle_elites():
    """A full plan step (production noise) scored again by the oracle: the device's elite set of the last iteration is
    the oracle's top-k of the same population whenever the oracle's gap at k exceeds 4x the measured error."""
    p, m = _planner(n=256)
    mod = make_model("reacher", obs_skip=0)
    rs = np.random.RandomState(5)
    start = _start(rs).astype(np.float32).astype(np.float64)
    p.begin_rollout()
    p.plan(start)
    n_last = p.population_size(2, first_step=True)[1]
    acts, costs = p.actions(2, n_last), p.costs(2, n_last)
    states = mod.rollout(start, acts.astype(np.float64))
    ref = costs_np.reacher_cost(costs_np.reacher_observation(states)).sum(axis=1)
    err = np.abs(costs - ref)
    assert np.median(err) <= 2e-4
    k = 10
    order = np.argsort(ref, kind="stable")
    gap = ref[order[k]] - ref[order[k - 1]]
    if gap > 4 * err[order[: k + 2]].max():
        assert set(np.argsort(costs, kind="stable")[:k].tolist()) == set(order[:k].tolist())
    p.close()


def test_controller_reaches_the_target():
    """MpcICemB200 + CudaGroundTruthModel on the Reacher stand-in: 40 closed-loop steps bring the fingertip to the
    target (cost = distance, reference mujoco.py:366-368); `elite_samples` carries gym's 11-wide observations."""
    from icem_b200 import envs
    from icem_b200.controller import MpcICemB200
    from icem_b200.models import CudaGroundTruthModel
    env = envs.make_env("Reacher")
    env.seed(4)
    obs = env.reset()
    fm = CudaGroundTruthModel(env=env)
    ctrl = MpcICemB200(env=env, forward_model=fm, horizon=15, num_simulated_trajectories=256, factor_decrease_num=1.25,
                       cost_along_trajectory="sum", seed=2,
                       action_sampler_params=dict(alpha=0.1, elites_size=10, opt_iterations=3, init_std=0.5,
                                                  use_mean_actions=True, keep_previous_elites=True,
                                                  shift_elites_over_time=True, fraction_elites_reused=0.3,
                                                  noise_beta=2.0))
    d0 = float(np.linalg.norm(obs[-3:]))
    ctrl.beginning_of_rollout(observation=obs, state=env.get_GT_state(), mode="train")
    for t in range(40):
        a = ctrl.get_action(obs, env.get_GT_state())
        if t == 3:
            es = ctrl.elite_samples
            o = es.as_array("observations")
            assert o.shape == (10, 15, 11)
            np.testing.assert_allclose(o[:, 0], np.broadcast_to(obs, (10, 11)), atol=1e-6)
            np.testing.assert_allclose(o[..., 0] ** 2 + o[..., 2] ** 2, 1.0, atol=1e-6)      # (cos q0, sin q0)
        obs, r, _, _ = env.step(a)
        assert np.isfinite(r) and r <= 0.0                  # reward = -distance of the pre-step observation
    d1 = float(np.linalg.norm(obs[-3:]))
    assert d0 > 0.05 and d1 < 0.25 * d0 and d1 < 0.03, (d0, d1)
    ctrl.close()
