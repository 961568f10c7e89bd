"""Shared helpers for the parity tests (oracle side + planner construction)."""
import numpy as np

from oracle import cases, costs_np
from oracle.icem_np import ICemConfig, ICemOracle


def oracle_for_case(case, record_actions=False):
    model = case["model"]()
    cfg = ICemConfig(**cases.controller_config(case))
    if case["cost"] == "halfcheetah":
        cost = lambda o, a: costs_np.halfcheetah_cost(o, a, case["penalise_flipping"])
    else:
        cost = costs_np.humanoid_standup_cost
    return model, cfg, ICemOracle(cfg, model.rollout, cost, record_actions=record_actions)


def planner_settings_for_case(case, **over):
    from icem_b200.planner import PlannerSettings
    c = cases.controller_config(case)
    model = case["model"]()
    s = PlannerSettings(
        horizon=c["horizon"], num_simulated_trajectories=c["num_simulated_trajectories"],
        action_low=c["action_low"], action_high=c["action_high"], dynamics="dense_tanh", cost=case["cost"],
        obs_dim=model.obs_dim, penalise_flipping=case["penalise_flipping"],
        factor_decrease_num=c["factor_decrease_num"], cost_along_trajectory=c["cost_along_trajectory"],
        alpha=c["alpha"], elites_size=c["elites_size"], opt_iterations=c["opt_iterations"], init_std=c["init_std"],
        use_mean_actions=c["use_mean_actions"], keep_previous_elites=c["keep_previous_elites"],
        shift_elites_over_time=c["shift_elites_over_time"], fraction_elites_reused=c["fraction_elites_reused"],
        noise_beta=c["noise_beta"], keep_iteration_actions=True)
    for k, v in over.items():
        setattr(s, k, v)
    return s, model


def stack_noise(noise_log):
    """[(zr, zi), ...] of one iteration (fresh rows, then shifted-elite rows) -> row-stacked arrays."""
    zr = np.concatenate([n[0] for n in noise_log], axis=0)
    zi = None if noise_log[0][1] is None else np.concatenate([n[1] for n in noise_log], axis=0)
    return zr, zi


def elite_gap(costs, k):
    """smallest gap between consecutive sorted costs among the first k+1 (the margin that decides the
    elite ORDER and the elite SET)."""
    s = np.sort(costs)[: k + 1]
    return float(np.min(np.diff(s)))
