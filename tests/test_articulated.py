"""CPU tests of the articulated-body oracle (oracle/articulated_np.py) and the model tables (icem_b200/robots.py).

The ground-truth dynamics has no external pin (MuJoCo absent, SURVEY F4: parity unpinned), so the oracle is pinned by
physics: conservation laws and closed-form cases that any correct rigid-body integrator must satisfy, plus a
committed regression fixture so the model cannot drift silently."""
import copy
import os

import numpy as np
import pytest

from icem_b200 import robots
from oracle.articulated_np import ArticulatedModel, make_model


def _frictionless(name, dt):
    m = copy.deepcopy(robots.get_model(name))
    m.dof_damping[:] = 0
    m.dof_stiffness[:] = 0
    m.dof_limited[:] = 0
    m.dt = dt
    return m


def _lifted_state(m, rs, P):
    q = np.tile(m.qpos0, (P, 1))
    q[:, 2 if m.dof_type[0] == robots.FREE_TRANS else 1] = 10.0          # far above the floor: no contacts
    simple = np.isin(m.dof_type, (robots.SLIDE, robots.HINGE))
    adr = m.dof_qadr[simple]
    q[:, adr[2:]] += rs.uniform(-0.3, 0.3, (P, len(adr) - 2)) if m.dof_type[0] != robots.FREE_TRANS else 0
    if m.dof_type[0] == robots.FREE_TRANS:
        q[:, adr] += rs.uniform(-0.3, 0.3, (P, len(adr)))
    return q, 1.5 * rs.randn(P, m.nv)


@pytest.mark.parametrize("name", ["halfcheetah", "humanoid_standup", "hopper", "ant", "humanoid"])
def test_tables(name):
    m = robots.get_model(name)
    dims = {"halfcheetah": (7, 9, 9, 6, 16), "humanoid_standup": (13, 24, 23, 17, 29), "hopper": (4, 6, 6, 3, 8),
            "ant": (9, 15, 14, 8, 25), "humanoid": (13, 24, 23, 17, 29)}[name]
    assert (m.nb, m.nq, m.nv, m.nu, m.nc) == dims
    assert np.all(m.body_parent < np.arange(m.nb))
    assert np.all(np.diff(m.con_body) >= 0)
    assert np.all(m.body_mass > 0)
    if name == "halfcheetah":
        assert abs(m.body_mass.sum() - 14.0) < 1e-9          # settotalmass
        assert list(m.dof_gear[3:]) == [120, 90, 60, 120, 60, 30]
    elif name == "hopper":
        assert 15 < m.body_mass.sum() < 16.5 and list(m.dof_gear[3:]) == [200, 200, 200]      # gym hopper.xml
        assert m.nsub == 4 and abs(m.dt - 0.002) < 1e-12
    elif name == "ant":
        assert 0.8 < m.body_mass.sum() < 1.0 and sorted(m.dof_gear[m.dof_act >= 0]) == [150] * 8   # density 5
        # actuator order of ant.xml: hip_4, ankle_4, hip_1, ankle_1, hip_2, ankle_2, hip_3, ankle_3
        assert list(m.dof_act[6:]) == [2, 3, 4, 5, 6, 7, 0, 1]
    else:
        assert 35 < m.body_mass.sum() < 50
        assert sorted(m.dof_gear[m.dof_act >= 0]) == sorted([100] * 7 + [300] * 2 + [200] * 2 + [25] * 6)
    # inertia tensors are positive definite
    I = m.body_inertia
    for b in range(m.nb):
        T = np.array([[I[b, 0], I[b, 3], I[b, 4]], [I[b, 3], I[b, 1], I[b, 5]], [I[b, 4], I[b, 5], I[b, 2]]])
        assert np.all(np.linalg.eigvalsh(T) > 0)


@pytest.mark.parametrize("name", ["halfcheetah", "humanoid_standup", "hopper", "ant"])
def test_energy_drift_is_first_order_in_dt(name):
    """Free flight, no dissipation: total energy is conserved up to the integrator's O(dt) error (halving dt
    halves the drift), which exercises the mass matrix and every Coriolis / centrifugal term."""
    drifts = []
    for dt, steps in ((4e-4, 100), (2e-4, 200)):
        m = _frictionless(name, dt)
        mod = ArticulatedModel(m)
        q, qd = _lifted_state(m, np.random.RandomState(0), 3)
        k0, p0 = mod.energy(q, qd)
        for _ in range(steps):
            q, qd = mod.integrate(q, qd, mod.qacc(q, qd, np.zeros((3, m.nu))))
        k1, p1 = mod.energy(q, qd)
        drifts.append(np.abs(k1 + p1 - k0 - p0) / k0)
    assert np.all(drifts[0] < 5e-3)
    assert np.all(drifts[1] < 0.65 * drifts[0] + 1e-6)


def test_momentum_conserved_without_gravity():
    m = _frictionless("humanoid_standup", 1e-3)
    m.gravity = 0.0
    mod = ArticulatedModel(m)
    q, qd = _lifted_state(m, np.random.RandomState(1), 2)

    def momentum(q, qd):
        parts = mod.qacc(q, qd, np.zeros((2, m.nu)), return_parts=True)
        lin = np.einsum("b,pba->pa", m.body_mass, parts["vcb"])
        ang = (np.einsum("pbij,pbj->pi", parts["Iw"], parts["wb"])
               + np.einsum("b,pba->pa", m.body_mass, np.cross(parts["cb"], parts["vcb"])))
        return lin, ang

    l0, a0 = momentum(q, qd)
    for _ in range(100):
        q, qd = mod.integrate(q, qd, mod.qacc(q, qd, np.zeros((2, m.nu))))
    l1, a1 = momentum(q, qd)
    # conserved up to the O(dt) error of evaluating M(q) at the start of each step
    np.testing.assert_allclose(l1, l0, atol=1e-3 * np.abs(l0).max())
    np.testing.assert_allclose(a1, a0, atol=2e-2 * np.abs(a0).max())


def test_free_fall_and_actuator_sign():
    mod = make_model("halfcheetah")
    m = mod.m
    q = m.qpos0[None].copy()
    q[0, 1] = 5.0
    acc = mod.qacc(q, np.zeros((1, m.nv)), np.zeros((1, m.nu)))[0]
    np.testing.assert_allclose(acc[:3], [0, -m.gravity, 0], atol=1e-9)
    np.testing.assert_allclose(acc[3:], 0, atol=1e-9)
    u = np.zeros((1, m.nu))
    u[0, 0] = 1.0                                   # bthigh motor, gear 120
    acc = mod.qacc(q, np.zeros((1, m.nv)), u)[0]
    assert acc[3] > 50                              # positive control accelerates its joint positively
    u[0, 0] = 5.0                                   # control is clipped to ctrl_limit
    np.testing.assert_allclose(mod.qacc(q, np.zeros((1, m.nv)), u)[0], acc)


def test_rest_on_floor_supports_weight():
    """Dropped with zero control, both robots settle; the contact forces then carry the total weight."""
    for name in ("halfcheetah", "humanoid_standup", "ant"):
        mod = make_model(name)
        m = mod.m
        st = np.concatenate([m.qpos0, np.zeros(m.nv)])[None]
        for _ in range(100 if name == "ant" else 60):
            st = mod.step_state(st, np.zeros((1, m.nu)))
        parts = mod.qacc(st[:, :m.nq], st[:, m.nq:], np.zeros((1, m.nu)), return_parts=True)
        # the ant keeps creeping at ~0.15 m/s: its ankles start outside their range, the limit springs load the feet
        # tangentially and the friction law is viscous below the Coulomb cap (20 N s/m is the explicit-stability limit
        # for this light robot)
        assert np.abs(st[0, m.nq:]).max() < (0.2 if name == "ant" else 0.05)
        assert abs(parts["fn"].sum() / (m.body_mass.sum() * m.gravity) - 1.0) < 0.02


def test_stable_under_bang_bang_controls():
    for name in ("halfcheetah", "humanoid_standup", "hopper", "ant"):
        mod = make_model(name)
        m = mod.m
        rs = np.random.RandomState(2)
        st = np.tile(np.concatenate([m.qpos0, np.zeros(m.nv)]), (48, 1))
        for t in range(24):
            if t % 4 == 0:
                u = rs.choice([-m.ctrl_limit, m.ctrl_limit], (48, m.nu))
            st = mod.step_state(st, u)
        assert np.isfinite(st).all() and np.abs(st[:, m.nq:]).max() < 150


@pytest.mark.parametrize("name", ["halfcheetah", "humanoid_standup", "hopper", "ant"])
def test_regression_fixture(name, golden_dir):
    """tests/golden/articulated_<name>.npz (oracle/make_golden_articulated.py): seeded rollouts of THIS oracle; pins
    the model + integrator against silent drift (it is a self-pin, not an external one)."""
    g = np.load(os.path.join(golden_dir, f"articulated_{name}.npz"))
    mod = make_model(name)
    obs = mod.rollout(g["start"], g["actions"])
    np.testing.assert_allclose(obs, g["observations"], rtol=0, atol=1e-9)
