"""CPU baseline: the reference's plan step timed on host cores -- the UNMODIFIED `MpcICem` when the reference
sources are present (`run_reference`, kind "reference"), else the NumPy oracle port (`run`, kind "port").

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


# ------------------------------------------------------------------------------------------------------------------
# kind "reference": the UNMODIFIED reference controller (icem/controllers/icem.py) imported from the reference sources
# (oracle/ref_loader.py: /root/reference in the build container, baseline/_ref/icem on a GPU box), driven through the
# reference's two model paths:
#   (a) "batched"  : a `ForwardModelWithDefaults` (models/abstract_models.py:8-53) whose predict() steps the whole
#                    population at once with the float64 oracle dynamics -- one process; this is the path a learned
#                    model takes and where the reference spends its time in Rollout / RolloutBuffer construction;
#   (b) "parallel" : `ParallelGroundTruthModel` (models/gt_par_model.py:17-100) with num_parallel = host cores over a
#                    `GroundTruthSupportEnv` stand-in backed by the same oracle dynamics -- the reference's own
#                    ground-truth configuration (settings/defaults/gt_default_env.json).
def _reference_controller_params(s, population):
    return dict(
        horizon=s.horizon, num_simulated_trajectories=int(population), factor_decrease_num=s.factor_decrease_num,
        cost_along_trajectory=s.cost_along_trajectory, do_visualize_plan=False, verbose=False,
        action_sampler_params=dict(alpha=s.alpha, elites_size=s.elites_size, opt_iterations=s.opt_iterations,
                                   init_std=s.init_std, use_mean_actions=s.use_mean_actions,
                                   keep_previous_elites=s.keep_previous_elites,
                                   shift_elites_over_time=s.shift_elites_over_time,
                                   fraction_elites_reused=s.fraction_elites_reused, noise_beta=s.noise_beta))


def _time_reference_arm(ctrl, env, steps, traj_per_step):
    """beginning_of_rollout + `steps` x get_action on the env's reset state; (seconds, trajectories)."""
    import contextlib
    import sys
    obs = env.reset()
    state = env.get_GT_state()
    with contextlib.redirect_stdout(sys.stderr):
        ctrl.beginning_of_rollout(observation=obs, state=state, mode="train")
        t_sum, n_traj = 0.0, 0
        for i in range(steps):
            t0 = time.perf_counter()
            act = ctrl.get_action(obs, state=state, mode="train")
            t_sum += time.perf_counter() - t0
            n_traj += traj_per_step(i == 0)
            obs, _, _, _ = env.step(np.asarray(act))
            state = env.get_GT_state()
    return t_sum, n_traj


def run_reference(name, cores=1, budget_s=25.0, steps=2):
    """Time the unmodified reference `MpcICem` on a bounded sample of workload `name` (both arms above)."""
    from icem_b200 import workloads
    from oracle import envs_np, ref_loader
    ref = ref_loader.load_reference()
    envs_np.install()
    from environments import env_from_string
    from models.gt_par_model import ParallelGroundTruthModel
    w = workloads.get_workload(name)
    s = workloads.planner_settings(name)
    if not w.get("env"):
        raise ValueError("the reference arm is defined for the ground-truth workloads")
    cores = max(1, int(cores))
    n_full = s.num_simulated_trajectories
    env_kwargs = dict(penalise_flipping=True, exclude_current_positions_from_observation=True) \
        if w["env"] == "HalfCheetah" else {}
    env = env_from_string(w["env"], **env_kwargs)
    env.seed(0)
    model = env._model
    obs_skip = model.obs_skip

    class _BatchedOracleModel(ref.abstract_models.ForwardModelWithDefaults):
        def predict(self, *, observations, states, actions):
            obs = np.asarray(observations, np.float64)
            st = np.concatenate([np.zeros((len(obs), obs_skip)), obs], axis=-1)   # dropped x does not enter the dynamics
            nxt = model.step_state(st, np.asarray(actions, np.float64))
            return nxt[:, obs_skip:], states, np.zeros((len(obs), 1))

        def train(self, buffer):
            pass

        def save(self, path):
            pass

        def load(self, path):
            pass

        def reset(self, observation):
            return None

        def got_actual_observation_and_env_state(self, *, observation, env_state=None, model_state=None):
            return None

    def traj_counter(cfg):
        return lambda first: trajectories_per_plan_step(cfg, first_step=first)

    arms = {}
    # calibrate: seconds per trajectory of the vectorised oracle dynamics (arm a) and of one env.step loop (arm b)
    rs = np.random.RandomState(0)
    lo, hi = np.asarray(s.action_low, np.float64), np.asarray(s.action_high, np.float64)
    st0 = np.concatenate([model.m.qpos0, np.zeros(model.m.nv)])
    n_cal = 256
    t0 = time.perf_counter()
    model.rollout(st0, rs.uniform(lo, hi, (n_cal, s.horizon, len(lo))))
    per_traj_batched = 3.0 * (time.perf_counter() - t0) / n_cal        # x3: the reference's Rollout bookkeeping
    t0 = time.perf_counter()
    model.rollout(st0, rs.uniform(lo, hi, (2, s.horizon, len(lo))))
    per_traj_seq = 3.0 * (time.perf_counter() - t0) / 2                # x3: env.step / cost_fn / pickling per trajectory
    ratio = trajectories_per_plan_step(build_oracle(name)[0], first_step=False) / float(n_full)
    for arm, per_traj, share in (("batched", per_traj_batched, 0.35), ("parallel", per_traj_seq / cores, 0.65)):
        n_fit = int(share * budget_s / max(steps, 1) / max(per_traj * ratio, 1e-12))
        n_sample = int(min(n_full, max(2 * s.elites_size + 4, n_fit)))
        cfg = build_oracle(name, population=n_sample)[0]
        if arm == "batched":
            fm = _BatchedOracleModel(env=env)
        else:
            fm = ParallelGroundTruthModel(env=env, num_parallel=cores)
        try:
            ctrl = ref.icem.MpcICem(env=env, forward_model=fm, **_reference_controller_params(s, n_sample))
            np.random.seed(0)
            t_sum, n_traj = _time_reference_arm(ctrl, env, steps, traj_counter(cfg))
        finally:
            if arm == "parallel":          # the reference's workers have no shutdown command (daemon processes)
                for p_ in fm.ps:
                    p_.terminate()
                for p_ in fm.ps:
                    p_.join(timeout=5)
        arms[arm] = {"value": n_traj / t_sum, "ms_per_step": 1e3 * t_sum / steps, "population": n_sample,
                     "steps": steps, "processes": 1 if arm == "batched" else cores}
    best = max(arms, key=lambda a: arms[a]["value"])
    return {"value": arms[best]["value"], "unit": "trajectories/s", "cores": arms[best]["processes"],
            "kind": "reference", "arm": best, "arms": arms,
            "sample": f"unmodified reference MpcICem.get_action, {steps} plan step(s) of workload {name}: population "
                      f"scaled {n_full}->{arms['batched']['population']} (batched ForwardModelWithDefaults, 1 process) "
                      f"and {n_full}->{arms['parallel']['population']} (ParallelGroundTruthModel, {cores} processes); "
                      f"all other settings unchanged; dynamics = this repo's float64 oracle (MuJoCo is not "
                      f"installable); value = the faster arm ({best})",
            "ms_per_step": arms[best]["ms_per_step"], "population": arms[best]["population"],
            "steps": steps, "host_cpu_count": os.cpu_count(), "reference_root": ref_loader.REFERENCE_ROOT}
