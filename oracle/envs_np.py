"""CPU stand-in environments backed by the NumPy articulated oracle.  TEST INFRASTRUCTURE ONLY.

Purpose: BASELINE configs[0] -- run the reference's UNCHANGED `icem/main.py` + `MpcICem` +
`ParallelGroundTruthModel` on a CPU (no GPU, no MuJoCo) to prove the plumbing (test T7), and to
give the CPU baseline a real `GroundTruthSupportEnv` to step.  `make_module()` returns a module object exporting
`HalfCheetahMaybeWithPosition` / `HumanoidStandup` (the names icem/environments/__init__.py:37-41 looks up in
`environments.mujoco`); the classes subclass the reference's own `GroundTruthSupportEnv`, so the reference must be
importable (oracle/ref_loader.install_reference()).
"""
import sys
import types

import numpy as np

from oracle import costs_np
from oracle.articulated_np import make_model


def make_module():
    from environments.abstract_environments import GroundTruthSupportEnv   # the reference's ABC
    from gym import spaces

    class _OracleEnv(GroundTruthSupportEnv):
        robot = None
        dt = 0.05

        def __init__(self, *, name, **kwargs):
            super().__init__(name=name, **kwargs)
            self._model = make_model(self.robot)
            m = self._model.m
            self.action_space = spaces.Box(-m.ctrl_limit * np.ones(m.nu, np.float32),
                                           m.ctrl_limit * np.ones(m.nu, np.float32))
            n = self._model.obs_dim
            self.observation_space = spaces.Box(-np.inf * np.ones(n, np.float32), np.inf * np.ones(n, np.float32))
            self._rs = np.random.RandomState(0)
            self._t = 0.0
            self._state = np.concatenate([m.qpos0, np.zeros(m.nv)])

        def seed(self, seed=None):
            self._rs = np.random.RandomState(seed)
            return [seed]

        def _obs(self):
            return self._model.observe(self._state).copy()

        def step(self, action):
            a = np.clip(np.asarray(action, np.float64), self.action_space.low, self.action_space.high)
            cost = float(self.cost_fn(self._obs(), a, None))
            self._state = self._model.step_state(self._state[None], a[None])[0]
            self._t += self.dt
            return self._obs(), -cost, False, {}

        def get_GT_state(self):
            return np.concatenate([[self._t], self._state])      # MjSimState.flatten() layout (mujoco.py:37-38)

        def set_GT_state(self, state):
            state = np.asarray(state, np.float64)
            self._t, self._state = float(state[0]), state[1:].copy()

        def set_state_from_observation(self, observation):
            raise NotImplementedError

        def close(self):
            pass

    class HalfCheetahMaybeWithPosition(_OracleEnv):
        robot = "halfcheetah"
        dt = 0.05

        def __init__(self, *, name, penalise_flipping=True, exclude_current_positions_from_observation=True,
                     **kwargs):
            self.penalise_flipping = penalise_flipping
            super().__init__(name=name, **kwargs)
            self.store_init_arguments(locals())
            assert exclude_current_positions_from_observation, "the stand-in implements the 17-wide observation"

        def cost_fn(self, observation, action, next_obs):
            return costs_np.halfcheetah_cost(np.asarray(observation), np.asarray(action), self.penalise_flipping)

        def reset(self):
            m = self._model.m
            self._state = np.concatenate([m.qpos0 + self._rs.uniform(-0.1, 0.1, m.nq), 0.1 * self._rs.randn(m.nv)])
            self._t = 0.0
            return self._obs()

    class HumanoidStandup(_OracleEnv):
        robot = "humanoid_standup"
        dt = 0.015

        def __init__(self, *, name, **kwargs):
            super().__init__(name=name, **kwargs)
            self.store_init_arguments(locals())

        def cost_fn(self, observation, action, next_obs):
            return costs_np.humanoid_standup_cost(np.asarray(observation), np.asarray(action))

        def reset(self):
            m = self._model.m
            q = m.qpos0 + self._rs.uniform(-0.01, 0.01, m.nq)
            q[3:7] /= np.linalg.norm(q[3:7])
            self._state = np.concatenate([q, self._rs.uniform(-0.01, 0.01, m.nv)])
            self._t = 0.0
            return self._obs()

    mod = types.ModuleType("environments.mujoco")
    mod.HalfCheetahMaybeWithPosition = HalfCheetahMaybeWithPosition
    mod.HumanoidStandup = HumanoidStandup
    return mod


def install():
    """Make `env_from_string("HalfCheetah")` of the reference resolve to the oracle-backed stand-ins."""
    import environments  # noqa: F401
    sys.modules["environments.mujoco"] = make_module()
