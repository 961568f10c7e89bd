"""CPU baseline: the NumPy oracle port of the reference's plan step timed on host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py): used by bench.py's `cpu_baseline` leg and by
`bench.py --impl reference`.  /root/reference does not exist on the GPU box, so this is the "port" arm: the
float64 restatement (oracle/icem_np.py) of `MpcICem.get_action` driving the oracle's NumPy dynamics.

Parallelism mirrors the reference's only strategy (icem/models/gt_par_model.py:66-94): the population is split
with np.array_split over `cores` forked worker processes, each rolls out its chunk, observations are gathered.

Bounded sample: a plan step at the workload's full population can take minutes on a CPU, so the population is
scaled down (same horizon / dims / iterations / hyper-parameters) until one plan step fits the time budget; the
reported value is trajectories/s of that sample and the scale is stated in `sample`.
"""
import multiprocessing as mp
import os
import time

import numpy as np

from oracle import costs_np
from oracle.icem_np import ICemConfig, ICemOracle, trajectories_per_plan_step

_WORKER_MODEL = None


def _worker_rollout(args):
    start, actions = args
    return _WORKER_MODEL.rollout(start, actions)


class _ParallelRollout:
    def __init__(self, model, cores):
        global _WORKER_MODEL
        self.model = model
        self.cores = cores
        self.pool = None
        if cores > 1:
            _WORKER_MODEL = model
            self.pool = mp.get_context("fork").Pool(cores)

    def __call__(self, start, actions):
        if self.pool is None or len(actions) < 2 * self.cores:
            return self.model.rollout(start, actions)
        chunks = np.array_split(actions, self.cores)          # gt_par_model.py:77-82
        outs = self.pool.map(_worker_rollout, [(start, c) for c in chunks])
        return np.concatenate(outs, axis=0)

    def close(self):
        if self.pool is not None:
            self.pool.terminate()
            self.pool = None


def build_oracle(name, population=None):
    """(oracle, model, start_state, cfg) for a named workload of icem_b200/workloads.py."""
    from icem_b200 import workloads    # sizes / hyper-parameters only
    w = workloads.get_workload(name)
    s = workloads.planner_settings(name)
    if w.get("dense"):
        from oracle.dynamics_np import DenseTanhModel
        model = DenseTanhModel(*workloads.dense_model_weights(*w["dense"]))
    elif w.get("mlp"):
        from oracle.dynamics_np import MlpModel
        model = MlpModel(*workloads.mlp_model_weights(*w["mlp"]))
    else:
        from oracle.articulated_np import make_model
        model = make_model(s.dynamics)
    cfg = ICemConfig(
        horizon=s.horizon, num_simulated_trajectories=int(population or s.num_simulated_trajectories),
        action_low=np.asarray(s.action_low, np.float32), action_high=np.asarray(s.action_high, np.float32),
        factor_decrease_num=s.factor_decrease_num, cost_along_trajectory=s.cost_along_trajectory, alpha=s.alpha,
        elites_size=s.elites_size, opt_iterations=s.opt_iterations, init_std=s.init_std,
        use_mean_actions=s.use_mean_actions, keep_previous_elites=s.keep_previous_elites,
        shift_elites_over_time=s.shift_elites_over_time, fraction_elites_reused=s.fraction_elites_reused,
        noise_beta=s.noise_beta)
    if s.cost == "halfcheetah":
        cost = lambda o, a: costs_np.halfcheetah_cost(o[..., : s.obs_dim], a, s.penalise_flipping)
    else:
        cost = costs_np.humanoid_standup_cost
    start = workloads.start_state(name, seed=0)
    return cfg, model, cost, start


def run(name, cores=1, budget_s=15.0, steps=None, warmup=0):
    """Time the oracle port on a bounded sample of workload `name`.  Returns the `cpu_baseline` object."""
    cores = max(1, int(cores))
    cfg_full, model, cost, start = build_oracle(name)
    n_full = cfg_full.num_simulated_trajectories
    roll = _ParallelRollout(model, cores)
    try:
        # calibrate the per-trajectory rollout cost on a small batch (also warms the worker pool)
        rs = np.random.RandomState(0)
        n_cal = max(32, 8 * cores)
        acts = rs.uniform(cfg_full.action_low, cfg_full.action_high, (n_cal, cfg_full.horizon, cfg_full.act_dim))
        roll(start, acts[: 2 * cores])
        t0 = time.perf_counter()
        roll(start, acts)
        per_traj = (time.perf_counter() - t0) / n_cal
        want_steps = int(steps) if steps else 2
        ratio = trajectories_per_plan_step(cfg_full, first_step=False) / float(n_full)    # trajectories per unit N
        n_fit = int(budget_s / (want_steps + warmup) / max(per_traj * ratio, 1e-12))
        n_sample = int(min(n_full, max(2 * cfg_full.elites_size + 4, n_fit)))
        cfg, _, _, _ = build_oracle(name, population=n_sample)
        orc = ICemOracle(cfg, roll, cost)
        np.random.seed(0)
        orc.beginning_of_rollout()
        orc.get_action(start) if warmup else None
        state = start
        t_sum, n_traj = 0.0, 0
        for i in range(want_steps):
            first = (i == 0 and not warmup)
            t0 = time.perf_counter()
            orc.get_action(state)
            t_sum += time.perf_counter() - t0
            n_traj += trajectories_per_plan_step(cfg, first_step=first)
        value = n_traj / t_sum
    finally:
        roll.close()
    return {"value": value, "unit": "trajectories/s", "cores": cores, "kind": "port",
            "sample": f"{want_steps} plan step(s) of workload {name} with the population scaled "
                      f"{n_full}->{n_sample} (all other settings unchanged), float64 NumPy oracle, "
                      f"{cores} process(es) splitting the population like ParallelGroundTruthModel",
            "ms_per_step": 1e3 * t_sum / want_steps, "population": n_sample,
            "host_cpu_count": os.cpu_count()}
