"""NumPy float64 restatement of the reference's vanilla CEM controller `MpcCemStd`
(/root/reference/icem/controllers/mpc.py:142-327).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Pinned against the imported, unmodified reference by oracle/make_golden.py -> tests/golden/cemstd_*.npz.
Sampling: the reference calls `scipy.stats.truncnorm.rvs(lower, upper, loc, scale, size)` (mpc.py:194-198), which
draws `np.random.uniform(size)` once and maps it through `truncnorm.ppf`; the uniforms are recorded so the device
can consume identical draws.  scipy is a genuine third-party dependency present in this image (here and on the GPU
box), so `truncnorm.ppf` itself is called, not restated.
"""
from dataclasses import dataclass
from typing import Callable, List

import numpy as np
from scipy.stats import truncnorm

from oracle.icem_np import IterationTrace, PlanTrace, reduce_costs


@dataclass
class CemStdConfig:
    """Keyword surface of MpcCemStd (mpc.py:150-156, 303-313) flattened."""
    horizon: int
    num_simulated_trajectories: int
    action_low: np.ndarray
    action_high: np.ndarray
    cost_along_trajectory: str = "sum"
    alpha: float = 0.1
    elites_size: int = 10
    opt_iterations: int = 3
    init_std: float = 0.5
    shift_means: bool = True
    execute_best_elite: bool = True
    bounds_like_levine: bool = False

    def __post_init__(self):
        self.action_low = np.asarray(self.action_low)
        self.action_high = np.asarray(self.action_high)

    @property
    def act_dim(self):
        return int(self.action_low.shape[0])

    @property
    def num_elites(self):      # mpc.py:316-320
        return max(2, min(self.elites_size, self.num_simulated_trajectories // 2))


class CemStdOracle:
    def __init__(self, cfg: CemStdConfig, rollout_fn: Callable, cost_fn: Callable, record_actions=False):
        self.cfg = cfg
        self.rollout_fn = rollout_fn
        self.cost_fn = cost_fn
        self.record_actions = record_actions
        self.was_reset = False

    def _init_std(self):       # mpc.py:180-185
        c = self.cfg
        return np.ones((c.horizon, c.act_dim)) * (c.action_high - c.action_low) / 2.0 * c.init_std

    def _update_bounds(self):  # mpc.py:290-301
        c = self.cfg
        if c.bounds_like_levine:
            lb_dist, ub_dist = self.mean - c.action_low, c.action_high - self.mean
            self.std = np.maximum(1e-8, np.minimum(np.minimum(lb_dist / 2, ub_dist / 2), self.std))
            self.lower, self.upper = -2, 2
        else:
            self.lower = (c.action_low - self.mean) / (self.std + 1e-8)
            self.upper = (c.action_high - self.mean) / (self.std + 1e-8)

    def beginning_of_rollout(self):   # mpc.py:163-175
        c = self.cfg
        self.mean = np.zeros((c.horizon, c.act_dim)) + (c.action_high + c.action_low) / 2.0
        self.std = self._init_std()
        self.elite_actions = None
        self._update_bounds()
        self.was_reset = True

    def _sample(self, n, noise_log):  # mpc.py:187-198
        u = np.random.uniform(size=(n,) + self.mean.shape)
        noise_log.append((u.copy(), None))
        return truncnorm.ppf(u, self.lower, self.upper) * self.std[None] + self.mean[None]

    def get_action(self, start_state) -> PlanTrace:   # mpc.py:200-262
        c = self.cfg
        if not self.was_reset:
            raise AttributeError("beginning_of_rollout() needs to be called before")
        iters: List[IterationTrace] = []
        actions = costs = None
        best = None
        for _ in range(c.opt_iterations):
            noise_log = []
            actions = self._sample(c.num_simulated_trajectories, noise_log)
            obs = self.rollout_fn(start_state, actions)
            costs = reduce_costs(self.cost_fn(obs, actions), c.cost_along_trajectory)
            best = int(np.argmin(costs))
            elite_idx = np.argsort(costs, kind="stable")[: c.num_elites]      # mpc.py:276 (see icem_np.py on ties)
            self.elite_actions = actions[elite_idx]
            self.mean = (1 - c.alpha) * self.elite_actions.mean(axis=0) + c.alpha * self.mean
            self.std = (1 - c.alpha) * self.elite_actions.std(axis=0) + c.alpha * self.std
            self._update_bounds()
            iters.append(IterationTrace(
                population=len(costs), num_fresh=len(costs), costs=costs.copy(), elite_idx=elite_idx.copy(),
                elite_costs=costs[elite_idx].copy(), mean=self.mean.copy(), std=self.std.copy(),
                actions=actions.copy() if self.record_actions else None, noise=noise_log))
        executed = actions[best][0].copy() if c.execute_best_elite else self.mean[0].copy()   # mpc.py:237-240
        if c.shift_means:                                                    # mpc.py:243-248
            self.mean[:-1] = self.mean[1:]
            self.mean[-1] = self.mean[-1] * 0 if c.bounds_like_levine else self.mean[-1]
        else:
            self.mean = np.zeros((c.horizon, c.act_dim))
        self.std = self._init_std()                                          # mpc.py:251-252
        self._update_bounds()
        return PlanTrace(action=executed, iterations=iters, mean_after_shift=self.mean.copy(),
                         std_after_reset=self.std.copy())
