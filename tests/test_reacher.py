"""CPU tests of the Reacher addition (reference: icem/environments/mujoco.py:346-368): cost restatement
against the reference's own function (fixture tests/golden/costs_reacher.npz, re-checked live when /root/reference
exists), the closed-form fingertip against the oracle's rigid-body kinematics of the same robot table, the stand-in
env's observation layout and reset, and the physics of the float64 model (undriven target stays put, actuated arm
accelerates as torque / inertia predicts)."""
import os

import numpy as np
import pytest

from oracle import costs_np
from oracle.make_golden_costs import reacher_inputs


def test_oracle_cost_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "costs_reacher.npz"))
    o, a = reacher_inputs()
    np.testing.assert_array_equal(costs_np.reacher_cost(o, a, o), g["reacher"])
    assert g["reacher"].shape == (50, 7) and np.all(g["reacher"] >= 0)


@pytest.mark.skipif(not os.path.isdir("/root/reference/icem"), reason="needs the reference sources")
def test_oracle_cost_matches_reference_live():
    from oracle.make_golden_costs import reference_reacher_costs
    o, a = reacher_inputs()
    np.testing.assert_array_equal(costs_np.reacher_cost(o, a, o), reference_reacher_costs())


def test_closed_form_fingertip_equals_the_rigid_body_kinematics():
    """The device cost forms |fingertip - target| from (q0, q1, q2, q3) in closed form; the same quantity from the
    oracle's kinematics of the robot table: world position of the fingertip geom on body1 minus the target body's."""
    from oracle.articulated_np import make_model
    mod = make_model("reacher", obs_skip=0)
    m = mod.m
    rs = np.random.RandomState(0)
    q = np.concatenate([rs.uniform(-3, 3, (64, 2)), rs.uniform(-0.2, 0.2, (64, 2))], axis=1)
    qd = np.zeros((64, 4))
    parts = mod.qacc(q, qd, np.zeros((64, 2)), return_parts=True)
    tip = parts["pb"][:, 1] + np.einsum("pij,j->pi", parts["Rb"][:, 1], np.array([0.11, 0.0, 0.0]))
    diff = tip - parts["pb"][:, 2]
    obs = costs_np.reacher_observation(np.concatenate([q, qd], axis=1))
    np.testing.assert_allclose(obs[:, -3:], diff, atol=1e-14)
    np.testing.assert_allclose(costs_np.reacher_cost(obs), np.linalg.norm(diff, axis=-1), atol=1e-14)
    assert m.nc == 0 and m.nq == 4 and m.nu == 2


def test_standin_env_layout_reset_and_cost():
    from icem_b200 import envs
    env = envs.make_env("Reacher")
    env.seed(3)
    o = env.reset()
    st = env.get_GT_state()
    assert o.shape == (11,) and st.shape == (9,) and st[0] == 0.0
    np.testing.assert_array_equal(o, costs_np.reacher_observation(st[1:]))
    assert np.linalg.norm(st[3:5]) < 0.2 and np.all(np.abs(st[1:3]) <= 0.1) and np.all(st[7:] == 0)
    assert env.cost_fn(o, np.zeros(2), o) == pytest.approx(np.linalg.norm(o[-3:]))
    rs = np.random.RandomState(1)
    obs = rs.randn(5, 7, 11)
    np.testing.assert_array_equal(env.cost_fn(obs, None, obs), costs_np.reacher_cost(obs))
    np.testing.assert_array_equal(envs.Reacher.observation_from_state(rs.randn(4, 6, 8)).shape, (4, 6, 11))
    # environments/mujoco.py:359-364: the observation determines angles (mod 2 pi), target and arm velocities
    env.set_state_from_observation(o)
    np.testing.assert_allclose(env.get_GT_state()[1:7], st[1:7], atol=1e-12)
    assert env.cuda_cost_spec()[0] == "reacher" and env.cuda_dynamics == "articulated"


def test_float64_model_physics():
    from oracle.articulated_np import make_model
    mod = make_model("reacher", obs_skip=0)
    m = mod.m
    st = np.array([[0.3, -0.5, 0.13, -0.07, 0.0, 0.0, 0.0, 0.0]])
    # no control: nothing moves (gravity is along the hinge axes, the target carries no force)
    s = st.copy()
    for _ in range(10):
        s = mod.step_state(s, np.zeros((1, 2)))
    np.testing.assert_allclose(s, st, atol=1e-12)
    # joint1 alone, arm at rest: qacc1 = gear / (armature + inertia of link 1 about its joint) within the coupling to
    # joint0 -- check against the mass matrix the oracle builds
    parts = mod.qacc(st[:, :4], st[:, 4:], np.array([[0.0, 1.0]]), return_parts=True)
    M = parts["M"][0]
    # (semi-implicit Euler: the joint dampers enter the solve as M + dt * diag(damping), articulated_np.qacc)
    want = np.linalg.solve(M + m.dt * np.diag(m.dof_damping), np.array([0.0, 200.0, 0.0, 0.0]))
    np.testing.assert_allclose(parts["qacc"][0], want, rtol=1e-10, atol=1e-10)
    assert M[0, 0] > 1.0 and M[1, 1] > 1.0 and abs(M[0, 2]) < 1e-15      # armature 1; the target is decoupled
    # the target never moves whatever the arm does
    s = st.copy()
    rs = np.random.RandomState(0)
    for _ in range(20):
        s = mod.step_state(s, rs.uniform(-1, 1, (1, 2)))
    np.testing.assert_array_equal(s[0, 2:4], st[0, 2:4])
    assert np.abs(s[0, :2] - st[0, :2]).max() > 0.1
