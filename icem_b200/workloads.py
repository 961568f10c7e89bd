"""Named workloads = the BASELINE.json configs, as PlannerSettings + synthetic start states.

Only sizes / hyper-parameters / seeds live here (resolved values of the reference's
settings/<env>/i-cem-blitz.json); no oracle code.
"""
import numpy as np

from .planner import PlannerSettings

_BLITZ = dict(alpha=0.1, elites_size=10, fraction_elites_reused=0.3, init_std=0.5, keep_previous_elites=True,
              shift_elites_over_time=True, use_mean_actions=True, factor_decrease_num=1.25, horizon=30,
              cost_along_trajectory="sum")


def mlp_model_weights(obs_dim, act_dim, hidden, seed):
    """Random-init 2-hidden-layer MLP (no checkpoint exists: the reference has no learned model):
    ([W1, W2, W3], [b1, b2, b3]) with W_l [out, in]; scaled so that 12-step rollouts stay O(1)."""
    rs = np.random.RandomState(seed)
    dims = [obs_dim + act_dim, hidden, hidden, obs_dim]
    ws, bs = [], []
    for i in range(3):
        ws.append((rs.randn(dims[i + 1], dims[i]) / np.sqrt(dims[i])).astype(np.float32))
        bs.append((0.05 * rs.randn(dims[i + 1])).astype(np.float32))
    ws[2] *= 0.3
    return ws, bs


def dense_model_weights(obs_dim, act_dim, seed):
    """Deterministic toy dense model (SURVEY Appendix C recipe)."""
    rs = np.random.RandomState(seed)
    a = 0.95 * np.eye(obs_dim) + 0.02 * rs.randn(obs_dim, obs_dim)
    b = 0.1 * rs.randn(obs_dim, act_dim)
    return a, b, np.zeros(obs_dim)


WORKLOADS = {
    # BASELINE configs[1]: HalfCheetah GT dynamics, h=30, beta=0.25, N=4096, 5 CEM iterations
    "halfcheetah_gt_n4096": dict(
        settings=dict(_BLITZ, num_simulated_trajectories=4096, opt_iterations=5, noise_beta=0.25,
                      dynamics="halfcheetah", cost="halfcheetah", obs_dim=17, penalise_flipping=True),
        act_dim=6, bound=1.0, env="HalfCheetah"),
    # BASELINE configs[2]: Humanoid Standup GT dynamics, h=30, beta=2.0, N=16384, elite reuse 0.3, 3 iterations
    "humanoid_standup_gt_n16384": dict(
        settings=dict(_BLITZ, num_simulated_trajectories=16384, opt_iterations=3, noise_beta=2.0,
                      dynamics="humanoid_standup", cost="humanoid_standup", obs_dim=47),
        act_dim=17, bound=0.4, env="HumanoidStandup"),
    # the same two with the four-stage Runge-Kutta substep gym's XML files ask of MuJoCo (4 dynamics evaluations per
    # substep instead of 1; robots.get_model(..., integrator="rk4"))
    "halfcheetah_gt_n4096_rk4": dict(
        settings=dict(_BLITZ, num_simulated_trajectories=4096, opt_iterations=5, noise_beta=0.25,
                      dynamics="halfcheetah", cost="halfcheetah", obs_dim=17, penalise_flipping=True,
                      integrator="rk4"),
        act_dim=6, bound=1.0, env="HalfCheetah", env_kwargs=dict(integrator="rk4")),
    "humanoid_standup_gt_n16384_rk4": dict(
        settings=dict(_BLITZ, num_simulated_trajectories=16384, opt_iterations=3, noise_beta=2.0,
                      dynamics="humanoid_standup", cost="humanoid_standup", obs_dim=47, integrator="rk4"),
        act_dim=17, bound=0.4, env="HumanoidStandup", env_kwargs=dict(integrator="rk4")),
    # the population of BASELINE configs[4] on ONE GPU (the strong-scaling base), Euler and Runge-Kutta
    "humanoid_standup_gt_n262144": dict(
        settings=dict(_BLITZ, num_simulated_trajectories=262144, opt_iterations=3, noise_beta=2.0,
                      dynamics="humanoid_standup", cost="humanoid_standup", obs_dim=47),
        act_dim=17, bound=0.4, env="HumanoidStandup"),
    "humanoid_standup_gt_n262144_rk4": dict(
        settings=dict(_BLITZ, num_simulated_trajectories=262144, opt_iterations=3, noise_beta=2.0,
                      dynamics="humanoid_standup", cost="humanoid_standup", obs_dim=47, integrator="rk4"),
        act_dim=17, bound=0.4, env="HumanoidStandup", env_kwargs=dict(integrator="rk4")),
    # one shard of BASELINE configs[4] (N=262144 over 8 GPUs): 32768 trajectories per GPU
    "humanoid_standup_gt_shard32768": dict(
        settings=dict(_BLITZ, num_simulated_trajectories=32768, opt_iterations=3, noise_beta=2.0,
                      dynamics="humanoid_standup", cost="humanoid_standup", obs_dim=47),
        act_dim=17, bound=0.4, env="HumanoidStandup"),
    # memory-side variant: same sampler / top-k / refit path with a trivially cheap dense forward model
    "dense_tanh_cheetah_n4096": dict(
        settings=dict(_BLITZ, num_simulated_trajectories=4096, opt_iterations=5, noise_beta=0.25,
                      dynamics="dense_tanh", cost="halfcheetah", obs_dim=17, penalise_flipping=True),
        act_dim=6, bound=1.0, env=None, dense=(17, 6, 7)),
    # BASELINE configs[3]: cheetah-sized learned MLP forward model (obs 18 + act 6 -> 256 -> 256 -> 18), h=12,
    # N=65536: tensor-core rollout path
    "mlp_cheetah_n65536": dict(
        settings=dict(_BLITZ, horizon=12, num_simulated_trajectories=65536, opt_iterations=3, noise_beta=0.25,
                      dynamics="mlp", cost="halfcheetah", obs_dim=18, penalise_flipping=True),
        act_dim=6, bound=1.0, env=None, mlp=(18, 6, 256, 21)),
    "dense_tanh_humanoid_n16384": dict(
        settings=dict(_BLITZ, num_simulated_trajectories=16384, opt_iterations=3, noise_beta=2.0,
                      dynamics="dense_tanh", cost="humanoid_standup", obs_dim=47),
        act_dim=17, bound=0.4, env=None, dense=(47, 17, 11)),
}


def get_workload(name):
    if name not in WORKLOADS:
        raise KeyError(f"unknown workload {name!r}; choose from {sorted(WORKLOADS)}")
    return WORKLOADS[name]


def planner_settings(name, world_size=1, rank=0, device=0, seed=0, scale_population=1) -> PlannerSettings:
    w = get_workload(name)
    s = dict(w["settings"])
    s["num_simulated_trajectories"] = int(s["num_simulated_trajectories"] * scale_population)
    low = -w["bound"] * np.ones(w["act_dim"], np.float32)
    return PlannerSettings(action_low=low, action_high=-low, world_size=world_size, rank=rank, device=device,
                           seed=seed, **s)


def populations(settings: PlannerSettings):
    """N_i per CEM iteration (icem/controllers/icem.py:123-127)."""
    out, n = [], settings.num_simulated_trajectories
    for i in range(settings.opt_iterations):
        if i > 0:
            n = max(settings.elites_size * 2, int(n / settings.factor_decrease_num))
        out.append(n)
    return out


def trajectories_per_step(settings: PlannerSettings, first_step=False):
    k = max(2, min(settings.elites_size, settings.num_simulated_trajectories // 2))
    n = sum(populations(settings))
    if not first_step and settings.shift_elites_over_time:
        n += int(k * settings.fraction_elites_reused)
    return n


def start_state(name, seed=0):
    """Synthetic start state mirroring the gym reset noise."""
    w = get_workload(name)
    rs = np.random.RandomState(1000 + seed)
    if w["env"] == "HalfCheetah":
        qpos = rs.uniform(-0.1, 0.1, 9)
        qvel = 0.1 * rs.randn(9)
        return np.concatenate([qpos, qvel])
    if w["env"] == "HumanoidStandup":
        from .envs import humanoid_standup_qpos0
        qpos = humanoid_standup_qpos0() + rs.uniform(-0.01, 0.01, 24)
        qpos[3:7] /= np.linalg.norm(qpos[3:7])          # a valid orientation (MuJoCo normalises the quaternion)
        qvel = rs.uniform(-0.01, 0.01, 23)
        return np.concatenate([qpos, qvel])
    obs_dim = w["dense"][0] if w.get("dense") else w["mlp"][0]
    return 0.1 * rs.randn(obs_dim)
