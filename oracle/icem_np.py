"""NumPy float64 restatement of the iCEM plan step (`MpcICem.get_action` and what it calls).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the CUDA path and the
"port" CPU baseline.  It travels to the GPU box (where /root/reference does not exist), so it
must not import the reference; it is pinned against the imported, unmodified reference by
`oracle/make_golden.py` -> `tests/golden/*.npz` and (when /root/reference is present) live in
`tests/test_oracle.py`.

All citations are relative to /root/reference/icem/.
"""
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

from oracle.shims import colorednoise as _cn


@dataclass
class ICemConfig:
    """Keyword surface of `MpcICem` (controllers/icem.py:22,213-233; controllers/mpc.py:22;
    controllers/abstract_controller.py:64-65) flattened."""
    horizon: int
    num_simulated_trajectories: int
    action_low: np.ndarray
    action_high: np.ndarray
    factor_decrease_num: float = 1.0
    cost_along_trajectory: str = "sum"
    alpha: float = 0.1
    elites_size: int = 10
    opt_iterations: int = 3
    init_std: float = 0.5
    use_mean_actions: bool = True
    keep_previous_elites: bool = True
    shift_elites_over_time: bool = True
    fraction_elites_reused: float = 0.3
    noise_beta: float = 1.0

    def __post_init__(self):
        # gym Box bounds are float32 (icem.py:50,56-57 do their arithmetic on them as such);
        # keep a float32 input as float32 so (high+low)/2 and (high-low)/2 round like the reference.
        self.action_low = np.asarray(self.action_low)
        self.action_high = np.asarray(self.action_high)
        if self.action_low.dtype != np.float32:
            self.action_low = self.action_low.astype(np.float64)
            self.action_high = self.action_high.astype(np.float64)
        if self.num_simulated_trajectories < 2:  # controllers/mpc.py:30-31
            raise ValueError("At least two trajectories needed!")

    @property
    def act_dim(self):
        return int(self.action_low.shape[0])

    @property
    def num_elites(self):  # controllers/icem.py:237-240
        return max(2, min(self.elites_size, self.num_simulated_trajectories // 2))


def population_schedule(cfg: ICemConfig) -> List[int]:
    """controllers/icem.py:123-127 -- compounded integer truncation, floor 2*elites_size."""
    sizes, n = [], cfg.num_simulated_trajectories
    for i in range(cfg.opt_iterations):
        if i > 0:
            n = max(cfg.elites_size * 2, int(n / cfg.factor_decrease_num))
        sizes.append(n)
    return sizes


def reduce_costs(costs_path: np.ndarray, how: str) -> np.ndarray:
    """controllers/abstract_controller.py:82-91 on a [p,h] array."""
    if how == "sum":
        return np.sum(costs_path, axis=1)
    if how == "best":
        return np.amin(costs_path, axis=1)
    if how == "final":
        return costs_path[:, -1]
    raise NotImplementedError(f"Implement method {how} to compute cost along trajectory")


@dataclass
class IterationTrace:
    population: int                 # p = fresh + shifted/kept extras
    num_fresh: int
    costs: np.ndarray               # [p]
    elite_idx: np.ndarray           # [k] best first
    elite_costs: np.ndarray         # [k]
    mean: np.ndarray                # [h,d] after refit
    std: np.ndarray                 # [h,d] after refit
    actions: Optional[np.ndarray] = None   # [p,h,d] if record_actions
    noise: list = field(default_factory=list)   # [(zr, zi), ...] unit draws in reference order


@dataclass
class PlanTrace:
    action: np.ndarray
    iterations: List[IterationTrace]
    mean_after_shift: np.ndarray
    std_after_reset: np.ndarray


class ICemOracle:
    """State machine equivalent to `MpcICem` (controllers/icem.py:16-247) for a model given as

        rollout_fn(start_state, actions[p,h,d]) -> observations[p,h,obs_dim]  (PRE-action obs, F9)
        cost_fn(observations[p,h,obs_dim], actions[p,h,d]) -> [p,h]

    Random numbers come from the legacy global `np.random` exactly like the reference
    (misc/seeding.py:18), through the restated `colorednoise` (oracle/shims/colorednoise.py).
    """

    def __init__(self, cfg: ICemConfig, rollout_fn: Callable, cost_fn: Callable,
                 record_actions=False):
        self.cfg = cfg
        self.rollout_fn = rollout_fn
        self.cost_fn = cost_fn
        self.record_actions = record_actions
        self.was_reset = False
        self.mean = self.std = None
        self.elite_actions = None   # [k,h,d] best first (== elite_samples.as_array("actions"))
        self.elite_obs = None
        self.elite_costs = None

    # controllers/icem.py:48-59
    def _init_mean(self):
        c = self.cfg
        return np.zeros((c.horizon, c.act_dim)) + (c.action_high + c.action_low) / 2.0

    def _init_std(self):
        c = self.cfg
        return np.ones((c.horizon, c.act_dim)) * (c.action_high - c.action_low) / 2.0 * c.init_std

    # controllers/icem.py:31-43
    def beginning_of_rollout(self):
        self.mean = self._init_mean()
        self.std = self._init_std()
        self.elite_actions = None
        self.elite_obs = None
        self.elite_costs = None
        self.was_reset = True

    # controllers/icem.py:61-82
    def _sample(self, num_traj, noise_log):
        c = self.cfg
        if c.noise_beta > 0:
            rec_prev, _cn.RECORDER = _cn.RECORDER, []
            try:
                samples = _cn.powerlaw_psd_gaussian(
                    c.noise_beta, size=(num_traj, c.act_dim, c.horizon)).transpose([0, 2, 1])
                noise_log.extend(_cn.RECORDER)
            finally:
                _cn.RECORDER = rec_prev
        else:
            z = np.random.randn(num_traj, c.horizon, c.act_dim)
            noise_log.append((z.copy(), None))
            samples = z
        return np.clip(samples * self.std + self.mean, c.action_low, c.action_high)

    def _n_keep(self):
        return int(self.elite_actions.shape[0] * self.cfg.fraction_elites_reused)

    # controllers/icem.py:106-189
    def get_action(self, start_state) -> PlanTrace:
        c = self.cfg
        if not self.was_reset:
            raise AttributeError("beginning_of_rollout() needs to be called before")
        iters = []
        num = c.num_simulated_trajectories
        pop_actions = pop_obs = costs = None
        best = None
        for i in range(c.opt_iterations):
            if i > 0:
                num = max(c.elites_size * 2, int(num / c.factor_decrease_num))
            noise_log = []
            actions = self._sample(num, noise_log)                      # :129 -> :84-89
            if c.use_mean_actions and i == c.opt_iterations - 1:
                actions[0] = self.mean
            if i == 0 and c.shift_elites_over_time and self.elite_actions is not None:   # :131-137
                n_keep = self._n_keep()
                reused = self.elite_actions[:n_keep, 1:]                 # :97-100
                last = self._sample(n_keep, noise_log)[:, -1:]           # :102 (full draw, keep t=h-1)
                actions = np.concatenate([actions, np.concatenate([reused, last], axis=1)], axis=0)
            obs = self.rollout_fn(start_state, actions)                  # :139
            pop_actions, pop_obs = actions, obs
            if i > 0 and c.keep_previous_elites:                         # :143-145 (not re-simulated)
                n_keep = self._n_keep()
                pop_actions = np.concatenate([actions, self.elite_actions[:n_keep]], axis=0)
                pop_obs = np.concatenate([obs, self.elite_obs[:n_keep]], axis=0)
            costs = reduce_costs(self.cost_fn(pop_obs, pop_actions), c.cost_along_trajectory)   # :147
            best = int(np.argmin(costs))                                 # :149
            # update_distributions, :194-211.  The reference's argsort is the default (unstable)
            # kind; the canonical order used by the CUDA path is ascending (cost, index), i.e.
            # kind="stable".  They coincide whenever the k+1 smallest costs are distinct.
            elite_idx = np.argsort(costs, kind="stable")[: c.num_elites]
            self.elite_actions = pop_actions[elite_idx]
            self.elite_obs = pop_obs[elite_idx]
            self.elite_costs = costs[elite_idx]
            new_mean = self.elite_actions.mean(axis=0)
            new_std = self.elite_actions.std(axis=0)
            self.mean = (1 - c.alpha) * new_mean + c.alpha * self.mean
            self.std = (1 - c.alpha) * new_std + c.alpha * self.std
            iters.append(IterationTrace(
                population=len(costs), num_fresh=num, costs=costs.copy(), elite_idx=elite_idx.copy(),
                elite_costs=self.elite_costs.copy(), mean=self.mean.copy(), std=self.std.copy(),
                actions=pop_actions.copy() if self.record_actions else None, noise=noise_log))
        executed = pop_actions[best][0].copy()                           # :163
        self.mean[:-1] = self.mean[1:]                                   # :167-171 (last row kept)
        self.std = self._init_std()                                      # :175
        return PlanTrace(action=executed, iterations=iters,
                         mean_after_shift=self.mean.copy(), std_after_reset=self.std.copy())


def trajectories_per_plan_step(cfg: ICemConfig, first_step: bool) -> int:
    """Simulated trajectories per plan step: sum_i N_i + shifted extras at t>0."""
    n = sum(population_schedule(cfg))
    if (not first_step) and cfg.shift_elites_over_time:
        n += int(cfg.num_elites * cfg.fraction_elites_reused)
    return n
