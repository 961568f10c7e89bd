"""NumPy float64 restatement of the reference's random-shooting controller `MpcRandom`
(/root/reference/icem/controllers/mpc.py:86-138, with `RndController.__init__`, controllers/random.py:5-9).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Pinned against the imported reference class by oracle/make_golden.py -> tests/golden/random_*.npz.  As shipped the
reference class is abstract (it never defines `StatefulController.end_of_rollout`), so the fixture generator
instantiates a subclass that adds that one no-op method and nothing else.

Draws: `env.action_space.sample()` is gym's `Box.sample`; with the stand-in `gym` package of oracle/shims it is
`np.random.uniform(low, high).astype(float32)` on the global legacy RNG, i.e. low + (high - low) * u with one
`random_sample()` per action dimension.  The u behind every action entry is recorded so the device can consume
identical draws.
"""
from dataclasses import dataclass
from typing import Callable

import numpy as np

from oracle.icem_np import IterationTrace, PlanTrace, reduce_costs


@dataclass
class RandomConfig:
    horizon: int
    num_simulated_trajectories: int
    action_low: np.ndarray
    action_high: np.ndarray
    action_change_frequency: int
    cost_along_trajectory: str = "sum"

    def __post_init__(self):
        self.action_low = np.asarray(self.action_low, np.float32)
        self.action_high = np.asarray(self.action_high, np.float32)
        assert self.action_change_frequency < self.horizon          # mpc.py:92

    @property
    def act_dim(self):
        return int(self.action_low.shape[0])


class RandomOracle:
    def __init__(self, cfg: RandomConfig, rollout_fn: Callable, cost_fn: Callable, record_actions=False):
        self.cfg = cfg
        self.rollout_fn = rollout_fn
        self.cost_fn = cost_fn
        self.record_actions = record_actions
        self._box_sample()                      # RndController.__init__: previous_action (random.py:8), unused
        self.current_action, self.current_u = self._box_sample()     # mpc.py:90
        self.counter = 0

    def _box_sample(self):
        c = self.cfg
        u = np.random.random_sample(c.act_dim)
        lo, hi = c.action_low.astype(np.float64), c.action_high.astype(np.float64)
        return (lo + (hi - lo) * u).astype(np.float32), u

    def _sample(self):                          # mpc.py:95-101
        if self.counter < self.cfg.action_change_frequency:
            self.counter += 1
        else:
            self.current_action, self.current_u = self._box_sample()
            self.counter = 0
        return self.current_action, self.current_u

    def beginning_of_rollout(self):             # mpc.py:69-73: nothing of the sampler is reset
        pass

    def get_action(self, obs) -> PlanTrace:     # mpc.py:109-138
        c = self.cfg
        acts = np.empty((c.num_simulated_trajectories, c.horizon, c.act_dim), np.float32)
        us = np.empty((c.num_simulated_trajectories, c.horizon, c.act_dim), np.float64)
        for n in range(c.num_simulated_trajectories):       # mpc.py:105: rows outer, time inner
            for t in range(c.horizon):
                acts[n, t], us[n, t] = self._sample()
        a64 = acts.astype(np.float64)
        observations = self.rollout_fn(obs, a64)
        costs = reduce_costs(self.cost_fn(observations, a64), c.cost_along_trajectory)
        best = int(np.argmin(costs))
        it = IterationTrace(population=len(costs), num_fresh=len(costs), costs=costs, elite_idx=np.array([best]),
                            elite_costs=costs[[best]], mean=None, std=None, noise=[(us, None)],
                            actions=a64 if self.record_actions else None)
        return PlanTrace(action=a64[best, 0].copy(), iterations=[it], mean_after_shift=None, std_after_reset=None)
