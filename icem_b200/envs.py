"""Stand-in environments for the B200 path.

gym / mujoco-py / MuJoCo are not installable here, so the two in-scope environments are provided as
`GroundTruthSupportEnv`-shaped classes (reference contract: icem/environments/abstract_environments.py:140-178,
what the callers touch is listed in SURVEY Appendix D item 4) whose `step` runs ONE transition of the same device
model the planner rolls out (`icem_sim_step`, include/icem_b200.h).  The articulated-body model is this repo's own
(icem_b200/csrc/dyn_articulated.cuh); it is NOT MuJoCo -- see DESIGN.md "Ground-truth dynamics".

Cost functions: each env carries `cost_fn` with the reference's signature and arithmetic
(environments/mujoco.py:67-99, :259-277) for host-side use by other controllers; the CUDA controller reads
`cuda_cost_spec()` and evaluates the same formula inside the rollout kernel.
"""
import math

import numpy as np

import sys as _sys

from .api import forget_failed_reference_import
from .planner import Planner, PlannerSettings

_before = set(_sys.modules)
try:   # the reference's env base class when the reference is importable (launcher), else a minimal stand-in
    from environments.abstract_environments import GroundTruthSupportEnv as _EnvBase  # noqa
    _HAVE_REF_ENV = True
except ImportError:   # pragma: no cover - exercised on boxes without the reference
    forget_failed_reference_import(_before)
    _HAVE_REF_ENV = False

    class _EnvBase:
        def __init__(self, *, name, **kwargs):
            self.name = name
            self.init_kwargs = {}

        def store_init_arguments(self, all_parameters):
            forbidden = {"self", "__class__"}
            self.init_kwargs = {k: v for k, v in all_parameters.items() if k not in forbidden}
            self.init_kwargs.update(self.init_kwargs.pop("kwargs", {}))


class Box:
    """gym.spaces.Box surface the controller reads: low / high / shape (float32 like gym)."""

    def __init__(self, low, high):
        self.low = np.asarray(low, np.float32)
        self.high = np.asarray(high, np.float32)
        self.shape = self.low.shape
        self.dtype = np.dtype(np.float32)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(np.float32)


# ---- cost functions restated from the reference (host side; the kernels evaluate the same formulas) ------------
def halfcheetah_cost_fn(observation, action, next_obs=None, penalise_flipping=True):
    """environments/mujoco.py:67-99."""
    observation, action = np.asarray(observation), np.asarray(action)
    if observation.shape[-1] == 18:
        root_angle, velocity = observation[..., 2], observation[..., 9]
    elif observation.shape[-1] == 17:
        root_angle, velocity = observation[..., 1], observation[..., 8]
    else:
        raise AttributeError(f"Got state of dimension {observation.shape[-1]}. Possible dimensions are 17 or 18.")
    scores = np.zeros(action.shape[:-1])
    if penalise_flipping:
        scores = scores + (root_angle > math.pi / 2) * 10
        scores = scores + (root_angle < -math.pi / 2) * 10
    return scores + 0.1 * np.sum(action ** 2, axis=-1) - velocity


def humanoid_standup_cost_fn(observation, action, next_obs=None):
    """environments/mujoco.py:259-277."""
    observation, action = np.asarray(observation), np.asarray(action)
    return -observation[..., 2] + 0.1 * np.square(action).sum(axis=-1)


def goal_distance_cost_fn(observation, action, next_obs, *, goal_index, achieved_index, sparse=False, threshold=0.05,
                          shaped=False):
    """environments/abstract_environments.py:115-123 / environments/robotics.py:150-164: distance between the desired
    goal and the achieved goal inside the observation (+ 0.1 x end effector to box when `shaped`), dense or thresholded."""
    o = np.asarray(observation)
    dist = np.linalg.norm(o[..., goal_index:goal_index + 3] - o[..., achieved_index:achieved_index + 3], axis=-1)
    eff = np.linalg.norm(o[..., :3] - o[..., 3:6], axis=-1) if shaped else 0
    if sparse:
        return np.asarray(dist > threshold, dtype=np.float32) + np.asarray(eff > threshold, dtype=np.float32) * 0.1
    return dist + eff * 0.1


class DenseStandInEnv(_EnvBase):
    """Environment whose true dynamics IS the dense-tanh model (obs' = tanh(W_o obs + W_a a + b)); used for the
    memory-side workloads and the plumbing tests.  state == observation."""

    def __init__(self, *, name="dense", act_dim, bound, cost, obs_dim, penalise_flipping=False, weights=None,
                 cost_params=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.cost_params = cost_params      # cost="goal_distance": goal_index, achieved_index, sparse, threshold, shaped
        self.store_init_arguments(locals())
        self.action_space = Box(-bound * np.ones(act_dim), bound * np.ones(act_dim))
        self.observation_space = Box(-np.ones(obs_dim), np.ones(obs_dim))
        self.cost = cost
        self.penalise_flipping = penalise_flipping
        self.weights = weights
        self._obs = np.zeros(obs_dim)
        self._rs = np.random.RandomState(0)

    def cuda_cost_spec(self):
        if self.cost == "goal_distance":
            return self.cost, False, dict(self.cost_params)
        return self.cost, self.penalise_flipping

    def cost_fn(self, observation, action, next_obs):
        if self.cost == "goal_distance":
            return goal_distance_cost_fn(observation, action, next_obs, **self.cost_params)
        if self.cost == "halfcheetah":
            return halfcheetah_cost_fn(observation, action, next_obs, self.penalise_flipping)
        return humanoid_standup_cost_fn(observation, action, next_obs)

    def seed(self, seed=None):
        self._rs = np.random.RandomState(seed)
        return [seed]

    def reset(self):
        self._obs = 0.1 * self._rs.randn(self.observation_space.shape[0])
        return self._obs.copy()

    def reset_with_mode(self, mode):
        return self.reset()

    def step(self, action):
        if self.weights is None:
            raise RuntimeError("DenseStandInEnv needs weights=(w_obs, w_act, bias) to step")
        w_o, w_a, b = self.weights
        cost = float(self.cost_fn(self._obs, np.asarray(action, np.float64), None))
        self._obs = np.tanh(w_o @ self._obs + w_a @ np.asarray(action, np.float64) + (0 if b is None else b))
        return self._obs.copy(), -cost, False, {}

    def get_GT_state(self):
        return self._obs.copy()

    def set_GT_state(self, state):
        self._obs = np.asarray(state, np.float64).copy()

    def set_state_from_observation(self, observation):
        self.set_GT_state(observation)

    def close(self):
        pass


class MlpStandInEnv(DenseStandInEnv):
    """Environment whose true dynamics IS the MLP forward model (float64 host evaluation of one transition)."""

    def __init__(self, *, name="mlp", act_dim, bound, cost, obs_dim, penalise_flipping=False, mlp=None,
                 cost_params=None, **kwargs):
        super().__init__(name=name, act_dim=act_dim, bound=bound, cost=cost, obs_dim=obs_dim,
                         penalise_flipping=penalise_flipping, cost_params=cost_params, **kwargs)
        self.mlp = mlp

    def step(self, action):
        ws, bs = self.mlp
        a = np.asarray(action, np.float64)
        cost = float(self.cost_fn(self._obs, a, None))
        x = np.concatenate([self._obs, a])
        for l, (w, b) in enumerate(zip(ws, bs)):
            x = np.asarray(w, np.float64) @ x + np.asarray(b, np.float64)
            if l + 1 < len(ws):
                x = np.tanh(x)
        self._obs = self._obs + x
        return self._obs.copy(), -cost, False, {}


# ---- articulated ground-truth stand-ins --------------------------------------------------------------------------
_ARTICULATED = {
    # name -> (dynamics id string, nq, nv, act_dim, ctrl bound, obs_dim (reference layout), reset noise)
    "HalfCheetah": dict(dynamics="halfcheetah", nq=9, nv=9, act_dim=6, bound=1.0),
    "HumanoidStandup": dict(dynamics="humanoid_standup", nq=24, nv=23, act_dim=17, bound=0.4),
    # generic articulated models (ICEM_DYN_ARTICULATED + tables of robots.py)
    "Hopper": dict(dynamics="articulated", robot="hopper", nq=6, nv=6, act_dim=3, bound=1.0),
    "Ant": dict(dynamics="articulated", robot="ant", nq=15, nv=14, act_dim=8, bound=1.0),
    "Humanoid": dict(dynamics="articulated", robot="humanoid", nq=24, nv=23, act_dim=17, bound=0.4),
    "Reacher": dict(dynamics="articulated", robot="reacher", nq=4, nv=4, act_dim=2, bound=1.0),
}


def halfcheetah_qpos0():
    from .robots import get_model
    return get_model("halfcheetah").qpos0.copy()


def humanoid_standup_qpos0():
    """Initial pose of the lying humanoid: root at z=0.105, rotated -90 deg about y (face up), joints at zero."""
    from .robots import get_model
    return get_model("humanoid_standup").qpos0.copy()


class _DeviceSimEnv(_EnvBase):
    """Common part: one device planner handle used only for `icem_sim_step` (single transitions)."""
    kind = None
    reward_is_negative_cost = True     # step() returns -cost_fn(obs, action): use_env_reward_as_cost == cost path

    def __init__(self, *, name, device=0, integrator="euler", **kwargs):
        super().__init__(name=name, **kwargs)
        spec = _ARTICULATED[self.kind]
        self.spec = spec
        self.device = device
        # env_params "integrator": "rk4" = the four-stage Runge-Kutta substep gym's XML files select (robots.get_model)
        self.integrator = integrator
        self.cuda_dynamics = spec["dynamics"]
        b = spec["bound"]
        self.action_space = Box(-b * np.ones(spec["act_dim"]), b * np.ones(spec["act_dim"]))
        self._sim = None
        self._rs = np.random.RandomState(0)
        self._t = 0.0
        self._state = np.concatenate([self._qpos0(), np.zeros(spec["nv"])])

    # -- device model --------------------------------------------------------------------------------
    def _simulator(self):
        if self._sim is None:      # created lazily: after fork()s of the host process
            spec = self.cuda_cost_spec()
            self._sim = Planner(PlannerSettings(
                horizon=2, num_simulated_trajectories=2, action_low=self.action_space.low,
                action_high=self.action_space.high, dynamics=self.cuda_dynamics, cost=spec[0],
                cost_params=spec[2] if len(spec) > 2 else None, articulated_model=self.cuda_articulated_model(),
                integrator=self.integrator,
                obs_offset=0 if self.cuda_dynamics == "articulated" else None,
                obs_dim=self.observation_space.shape[0], device=self.device))
        return self._sim

    def cuda_articulated_model(self):
        """robots.CompiledModel for dynamics="articulated" (None: the built-in tables of the dynamics id)."""
        from .robots import get_model
        if self.cuda_dynamics != "articulated":
            return None if self.integrator == "euler" else get_model(self.cuda_dynamics, integrator=self.integrator)
        return get_model(self.spec["robot"], integrator=self.integrator)

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_sim"] = None
        return d

    def seed(self, seed=None):
        self._rs = np.random.RandomState(seed)
        return [seed]

    def reset_with_mode(self, mode):
        return self.reset()

    def get_GT_state(self):
        # MjSimState.flatten() layout [time, qpos, qvel] (environments/mujoco.py:37-38)
        return np.concatenate([[self._t], self._state])

    def set_GT_state(self, state):
        state = np.asarray(state, np.float64)
        self._t = float(state[0])
        self._state = state[1:].copy()

    def set_state_from_observation(self, observation):
        raise NotImplementedError("the observation does not determine the state; use env states")

    def step(self, action):
        a = np.clip(np.asarray(action, np.float64), self.action_space.low, self.action_space.high)
        obs = self._obs()
        nxt, _, _ = self._simulator().sim_step(self._state, a)
        self._state = nxt
        self._t += self.dt
        nobs = self._obs()
        cost = float(self.cost_fn(obs, a, nobs))
        return nobs, -cost, False, {}

    def simulate(self, state, action):
        self.set_GT_state(state)
        obs, r, _, _ = self.step(action)
        return obs, self.get_GT_state(), r

    def simulate_state(self, state, action):
        """(next_obs, next GT state, reward) of one transition from `state` = [time, qpos, qvel] on the device model;
        the env itself (its own state, time, RNG) is left untouched."""
        saved = (self._state.copy(), self._t)
        try:
            return self.simulate(state, action)
        finally:
            self._state, self._t = saved

    def close(self):
        if self._sim is not None:
            self._sim.close()
            self._sim = None


class HalfCheetahMaybeWithPosition(_DeviceSimEnv):
    """Stand-in for environments/mujoco.py:48-131.  obs = qpos[1:] ++ qvel (17) or qpos ++ qvel (18)."""
    kind = "HalfCheetah"
    dt = 0.05

    def __init__(self, *, name="HalfCheetah", penalise_flipping=True, exclude_current_positions_from_observation=True,
                 device=0, **kwargs):
        self.penalise_flipping = penalise_flipping
        self.exclude_pos = exclude_current_positions_from_observation
        self.observation_space = Box(-np.inf * np.ones(17 if self.exclude_pos else 18),
                                     np.inf * np.ones(17 if self.exclude_pos else 18))
        super().__init__(name=name, device=device, **kwargs)
        self.store_init_arguments(locals())

    def _qpos0(self):
        return halfcheetah_qpos0()

    def cuda_cost_spec(self):
        return "halfcheetah", self.penalise_flipping

    def cost_fn(self, observation, action, next_obs):
        return halfcheetah_cost_fn(observation, action, next_obs, self.penalise_flipping)

    def _obs(self):
        return self._state[1:].copy() if self.exclude_pos else self._state.copy()

    def reset(self):
        # gym half_cheetah_v3 reset noise
        qpos = self._qpos0() + self._rs.uniform(-0.1, 0.1, 9)
        qvel = 0.1 * self._rs.randn(9)
        self._state = np.concatenate([qpos, qvel])
        self._t = 0.0
        return self._obs()


class HumanoidStandup(_DeviceSimEnv):
    """Stand-in for environments/mujoco.py:228-277.  Same 378-wide observation layout as the reference
    (mujoco.py:241-252: qpos(24) ++ qvel(23) ++ cinert(140) ++ cvel(84) ++ qfrc_actuator(23) ++ cfrc_ext(84)); the cost
    reads only obs[2] (= qpos[2], root height), and the device model carries qpos / qvel only, so the four trailing
    blocks (331 entries) are zeros -- consumers of `elite_samples["observations"]` see the reference's shape."""
    kind = "HumanoidStandup"
    dt = 0.015
    obs_pad = 331

    def __init__(self, *, name="HumanoidStandup", device=0, **kwargs):
        n = 47 + self.obs_pad
        self.observation_space = Box(-np.inf * np.ones(n), np.inf * np.ones(n))
        super().__init__(name=name, device=device, **kwargs)
        self.store_init_arguments(locals())

    def _qpos0(self):
        return humanoid_standup_qpos0()

    def cuda_cost_spec(self):
        return "humanoid_standup", False

    def cost_fn(self, observation, action, next_obs):
        return humanoid_standup_cost_fn(observation, action, next_obs)

    def _obs(self):
        return np.concatenate([self._state, np.zeros(self.obs_pad)])

    def reset(self):
        c = 0.01
        qpos = self._qpos0() + self._rs.uniform(-c, c, 24)
        qpos[3:7] /= np.linalg.norm(qpos[3:7])
        qvel = self._rs.uniform(-c, c, 23)
        self._state = np.concatenate([qpos, qvel])
        self._t = 0.0
        return self._obs()


def locomotion_cost_fn(observation, action, next_obs, *, dt, ctrl_weight, unhealthy_weight, z_index, z_lo, z_hi,
                       z_strict, state_bound, velocity_index=-1, forward_weight=1.0):
    """environments/mujoco.py:153-176 (Ant) / :196-231 (Hopper) / :314-343 (Humanoid: the x velocity is read from the
    observation instead of differenced), vectorised over leading dims.  Hopper's
    `np.logical_and(healthy_state, healthy_z, healthy_angle)` passes the angle test as the OUT argument, so the
    angle range never takes part (kept that way)."""
    observation, action, next_obs = np.asarray(observation), np.asarray(action), np.asarray(next_obs)
    z = observation[..., z_index]
    healthy = (z_lo < z) * (z < z_hi) if z_strict else (z_lo <= z) * (z <= z_hi)
    if state_bound > 0:
        st = observation[..., 2:]
        healthy = healthy * np.all(np.logical_and(-state_bound < st, st < state_bound), axis=-1)
    unhealthy = 1 - np.isfinite(observation).all(axis=-1) * healthy
    if velocity_index >= 0:
        x_velocity = observation[..., velocity_index]
    else:
        x_velocity = (next_obs[..., 0] - observation[..., 0]) / dt
    return -forward_weight * x_velocity + unhealthy_weight * unhealthy + ctrl_weight * np.sum(np.square(action), axis=-1)


class _LocomotionEnv(_DeviceSimEnv):
    """Hopper / Ant stand-ins: obs = qpos ++ qvel (+ zeros where the reference appends contact forces)."""
    cost_params = None
    obs_pad = 0

    def __init__(self, *, name, device=0, **kwargs):
        spec = _ARTICULATED[self.kind]
        n = spec["nq"] + spec["nv"] + self.obs_pad
        self.observation_space = Box(-np.inf * np.ones(n), np.inf * np.ones(n))
        super().__init__(name=name, device=device, **kwargs)
        self.store_init_arguments(locals())

    def _qpos0(self):
        from .robots import get_model
        return get_model(self.spec["robot"]).qpos0.copy()

    def cuda_cost_spec(self):
        return "locomotion", False, dict(self.cost_params, dt=self.dt)

    def cost_fn(self, observation, action, next_obs):
        return locomotion_cost_fn(observation, action, next_obs, dt=self.dt, **self.cost_params)

    def _obs(self):
        return np.concatenate([self._state, np.zeros(self.obs_pad)])


class Hopper(_LocomotionEnv):
    """Stand-in for environments/mujoco.py:179-231 (gym Hopper-v3 with exclude_current_positions_from_observation
    false: obs = qpos(6) ++ clip(qvel(6), -10, 10) = 12, like gym's hopper_v3._get_obs)."""
    kind = "Hopper"
    dt = 0.008
    cost_params = dict(ctrl_weight=1e-3, unhealthy_weight=200.0, z_index=1, z_lo=0.7, z_hi=float("inf"), z_strict=True,
                       state_bound=100.0)

    def _obs(self):
        return np.concatenate([self._state[:6], np.clip(self._state[6:], -10.0, 10.0)])

    def reset(self):
        c = 5e-3            # gym hopper_v3 reset_noise_scale
        self._state = np.concatenate([self._qpos0() + self._rs.uniform(-c, c, 6), self._rs.uniform(-c, c, 6)])
        self._t = 0.0
        return self._obs()


class Ant(_LocomotionEnv):
    """Stand-in for environments/mujoco.py:134-176 (gym Ant-v3, 113-wide observation: qpos(15) ++ qvel(14) ++ 84
    clipped contact forces; the cost reads obs[0], obs[2] and finiteness only, so the force block is zeros)."""
    kind = "Ant"
    dt = 0.05
    obs_pad = 84
    cost_params = dict(ctrl_weight=0.5, unhealthy_weight=100.0, z_index=2, z_lo=0.2, z_hi=1.0, z_strict=False,
                       state_bound=0.0)

    def reset(self):
        c = 0.1             # gym ant_v3 reset_noise_scale
        qpos = self._qpos0() + self._rs.uniform(-c, c, 15)
        qpos[3:7] /= np.linalg.norm(qpos[3:7])
        self._state = np.concatenate([qpos, c * self._rs.randn(14)])
        self._t = 0.0
        return self._obs()


class Humanoid(_LocomotionEnv):
    """Stand-in for environments/mujoco.py:279-343 (gym Humanoid-v3 with exclude_current_positions_from_observation
    false: 378-wide observation = qpos(24) ++ qvel(23) ++ inertia / velocity / force blocks the cost never reads,
    zeros here).  Cost: -1.25 * obs[nq] + 100 * [z outside (1, 2)] + 0.1 |a|^2."""
    kind = "Humanoid"
    dt = 0.015
    obs_pad = 331
    cost_params = dict(ctrl_weight=0.1, unhealthy_weight=100.0, z_index=2, z_lo=1.0, z_hi=2.0, z_strict=True,
                       state_bound=0.0, velocity_index=24, forward_weight=1.25)

    def reset(self):
        c = 0.01            # gym humanoid_v3 reset_noise_scale
        qpos = self._qpos0() + self._rs.uniform(-c, c, 24)
        qpos[3:7] /= np.linalg.norm(qpos[3:7])
        self._state = np.concatenate([qpos, self._rs.uniform(-c, c, 23)])
        self._t = 0.0
        return self._obs()


def reacher_cost_fn(observation, action, next_obs):
    """environments/mujoco.py:366-368: |fingertip - target| = norm of the last three observation entries."""
    return np.linalg.norm(np.asarray(observation)[..., -3:], axis=-1)


class Reacher(_DeviceSimEnv):
    """Stand-in for environments/mujoco.py:346-368 (gym Reacher-v2): 11-wide observation
    [cos q0, cos q1, sin q0, sin q1, target x, target y, qvel0, qvel1, fingertip - target (3)]; state = [time, qpos(4),
    qvel(4)] with qpos[2:4] the target's slide joints.  The device model carries the state; `observation_from_state`
    rebuilds gym's layout from it (cost on the device: ICEM_COST_REACHER from the same forward kinematics)."""
    kind = "Reacher"
    dt = 0.02
    reach = (0.1, 0.11, 0.0, 0.0)      # link lengths, target world position at q2 = q3 = 0 (robots.reacher)

    def __init__(self, *, name="Reacher", device=0, frame_skip=None, **kwargs):
        if frame_skip not in (None, 2):
            raise NotImplementedError("the device Reacher model is compiled with gym's frame_skip = 2")
        self.observation_space = Box(-np.inf * np.ones(11), np.inf * np.ones(11))
        super().__init__(name=name, device=device, **kwargs)
        self.store_init_arguments(locals())

    def _qpos0(self):
        return np.zeros(4)

    def cuda_cost_spec(self):
        return "reacher", False, dict(reach=self.reach)

    def cost_fn(self, observation, action, next_obs):
        return reacher_cost_fn(observation, action, next_obs)

    @classmethod
    def observation_from_state(cls, state):
        """[..., 8] device states (qpos ++ qvel) -> [..., 11] observations in gym's layout (reacher.py::_get_obs)."""
        st = np.asarray(state, np.float64)
        q0, q1, tx, ty = st[..., 0], st[..., 1], st[..., 2], st[..., 3]
        l1, l2, ox, oy = cls.reach
        fx = l1 * np.cos(q0) + l2 * np.cos(q0 + q1)
        fy = l1 * np.sin(q0) + l2 * np.sin(q0 + q1)
        return np.stack([np.cos(q0), np.cos(q1), np.sin(q0), np.sin(q1), tx, ty, st[..., 4], st[..., 5],
                         fx - (ox + tx), fy - (oy + ty), np.zeros_like(q0)], axis=-1)

    def _obs(self):
        return self.observation_from_state(self._state)

    def set_state_from_observation(self, observation):      # environments/mujoco.py:359-364
        o = np.asarray(observation, np.float64)
        self._state = np.array([np.arctan2(o[2], o[0]), np.arctan2(o[3], o[1]), o[4], o[5], o[6], o[7], 0.0, 0.0])

    def reset(self):
        # gym reacher.py::reset_model: arm angles U(-0.1, 0.1), goal uniform in the disc of radius 0.2, arm velocities
        # U(-0.005, 0.005), target at rest
        q = self._rs.uniform(-0.1, 0.1, 4)
        while True:
            goal = self._rs.uniform(-0.2, 0.2, 2)
            if np.linalg.norm(goal) < 0.2:
                break
        q[2:] = goal
        qd = np.concatenate([self._rs.uniform(-0.005, 0.005, 2), np.zeros(2)])
        self._state = np.concatenate([q, qd])
        self._t = 0.0
        return self._obs()


def make_env(kind, device=0, **kwargs):
    if kind == "Reacher":
        return Reacher(name=kind, device=device, **kwargs)
    if kind == "Humanoid":
        return Humanoid(name=kind, device=device, **kwargs)
    if kind == "Hopper":
        return Hopper(name=kind, device=device, **kwargs)
    if kind == "Ant":
        return Ant(name=kind, device=device, **kwargs)
    if kind == "HalfCheetah":
        return HalfCheetahMaybeWithPosition(name=kind, device=device, **kwargs)
    if kind == "HumanoidStandup":
        return HumanoidStandup(name=kind, device=device, **kwargs)
    raise KeyError(kind)
