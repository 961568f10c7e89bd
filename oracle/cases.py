"""Seeded parity cases shared by the golden generator and the tests.  TEST INFRASTRUCTURE ONLY.

Every case is a (controller settings, toy dense-tanh model, start observation, seed) tuple at a
size the reference itself finishes in seconds.  Settings are the resolved
`settings/<env>/i-cem-blitz.json` values unless a case overrides them.
"""
import numpy as np

from oracle.dynamics_np import DenseTanhModel


def _sampler(**over):
    p = dict(alpha=0.1, elites_size=10, fraction_elites_reused=0.3, init_std=0.5,
             keep_previous_elites=True, noise_beta=0.25, opt_iterations=3,
             shift_elites_over_time=True, use_mean_actions=True)
    p.update(over)
    return p


def _ctrl(n, sampler, **over):
    p = dict(num_simulated_trajectories=n, factor_decrease_num=1.25, horizon=30,
             cost_along_trajectory="sum", action_sampler_params=sampler,
             do_visualize_plan=False, verbose=False)
    p.update(over)
    return p


def _humanoid_like_model(obs_dim=47, act_dim=17, seed=11):
    rs = np.random.RandomState(seed)
    a = 0.9 * np.eye(obs_dim) + 0.03 * rs.randn(obs_dim, obs_dim)
    b = 0.3 * rs.randn(obs_dim, act_dim)
    c = 0.05 * rs.randn(obs_dim)
    return DenseTanhModel(a, b, c)


CASES = {
    # the survey-time known-answer trace
    "appendix_c": dict(
        model=lambda: DenseTanhModel.appendix_c(), cost="halfcheetah", penalise_flipping=False,
        low=-np.ones(6), high=np.ones(6), ctrl=_ctrl(40, _sampler()),
        start_obs=np.zeros(17), seed=123, steps=3),
    # BASELINE configs[0]: N=128, 3 CEM iterations, h=30, beta=0.25 (HalfCheetah cost incl. flip penalty)
    "cheetah_n128": dict(
        model=lambda: DenseTanhModel.appendix_c(seed=3), cost="halfcheetah", penalise_flipping=True,
        low=-np.ones(6), high=np.ones(6), ctrl=_ctrl(128, _sampler()),
        start_obs=0.1 * np.random.RandomState(1000).randn(17), seed=0, steps=3),
    # configs[1] shape at reduced N: 5 CEM iterations
    "cheetah_5iter": dict(
        model=lambda: DenseTanhModel.appendix_c(seed=5), cost="halfcheetah", penalise_flipping=True,
        low=-np.ones(6), high=np.ones(6), ctrl=_ctrl(96, _sampler(opt_iterations=5)),
        start_obs=0.1 * np.random.RandomState(1001).randn(17), seed=1, steps=3),
    # configs[2] shape at reduced N: d=17, beta=2.0, bounds +-0.4, HumanoidStandup cost (obs[2])
    "humanoid_n128": dict(
        model=_humanoid_like_model, cost="humanoid_standup", penalise_flipping=False,
        low=-0.4 * np.ones(17), high=0.4 * np.ones(17),
        ctrl=_ctrl(128, _sampler(noise_beta=2.0)),
        start_obs=0.1 * np.random.RandomState(1002).randn(47), seed=2, steps=3),
    # feature flags off / other reductions / white noise / decay floor / h=12
    "flags_off_best": dict(
        model=lambda: DenseTanhModel.appendix_c(seed=6), cost="halfcheetah", penalise_flipping=True,
        low=-np.ones(6), high=np.ones(6),
        ctrl=_ctrl(64, _sampler(keep_previous_elites=False, shift_elites_over_time=False,
                                use_mean_actions=False, noise_beta=1.0),
                   cost_along_trajectory="best"),
        start_obs=0.1 * np.random.RandomState(1003).randn(17), seed=3, steps=2),
    "white_final_floor": dict(
        model=lambda: DenseTanhModel.appendix_c(seed=8), cost="halfcheetah", penalise_flipping=True,
        low=np.array([-1, -0.5, -1, -2, -1, -1.0]), high=np.array([1, 0.5, 2, 1, 1, 1.0]),
        ctrl=_ctrl(24, _sampler(noise_beta=0.0, elites_size=8, opt_iterations=4, alpha=0.3,
                                fraction_elites_reused=0.5),
                   cost_along_trajectory="final", factor_decrease_num=2.0, horizon=12),
        start_obs=0.1 * np.random.RandomState(1004).randn(17), seed=4, steps=3),
}


def controller_config(case):
    """Flatten a case's settings into oracle.icem_np.ICemConfig kwargs."""
    c = dict(case["ctrl"])
    s = dict(c.pop("action_sampler_params"))
    c.pop("do_visualize_plan", None)
    c.pop("verbose", None)
    c.update(s)
    c["action_low"] = np.asarray(case["low"], np.float32)     # gym Box dtype
    c["action_high"] = np.asarray(case["high"], np.float32)
    return c


def _cem_sampler(**over):
    p = dict(alpha=0.1, elites_size=10, opt_iterations=3, init_std=0.5, shift_means=True, execute_best_elite=True,
             bounds_like_levine=False)
    p.update(over)
    return p


# vanilla CEM (controllers/mpc.py::MpcCemStd): the paper's baseline, SURVEY 8(f)-1
CEM_STD_CASES = {
    "cemstd_cheetah": dict(
        model=lambda: DenseTanhModel.appendix_c(seed=3), cost="halfcheetah", penalise_flipping=True,
        low=-np.ones(6), high=np.ones(6),
        ctrl=dict(num_simulated_trajectories=96, horizon=30, cost_along_trajectory="sum",
                  action_sampler_params=_cem_sampler(), do_visualize_plan=False, verbose=False),
        start_obs=0.1 * np.random.RandomState(1010).randn(17), seed=10, steps=3),
    "cemstd_levine_mean": dict(
        model=_humanoid_like_model, cost="humanoid_standup", penalise_flipping=False,
        low=-0.4 * np.ones(17), high=0.4 * np.ones(17),
        ctrl=dict(num_simulated_trajectories=64, horizon=12, cost_along_trajectory="sum",
                  action_sampler_params=_cem_sampler(bounds_like_levine=True, execute_best_elite=False, alpha=0.25,
                                                     elites_size=8, opt_iterations=4),
                  do_visualize_plan=False, verbose=False),
        start_obs=0.1 * np.random.RandomState(1011).randn(47), seed=11, steps=3),
    "cemstd_noshift": dict(
        model=lambda: DenseTanhModel.appendix_c(seed=9), cost="halfcheetah", penalise_flipping=True,
        low=np.array([-1, -0.5, -1, -2, -1, -1.0]), high=np.array([1, 0.5, 2, 1, 1, 1.0]),
        ctrl=dict(num_simulated_trajectories=40, horizon=20, cost_along_trajectory="best",
                  action_sampler_params=_cem_sampler(shift_means=False, init_std=0.3),
                  do_visualize_plan=False, verbose=False),
        start_obs=0.1 * np.random.RandomState(1012).randn(17), seed=12, steps=2),
}


def cem_std_config(case):
    c = dict(case["ctrl"])
    s = dict(c.pop("action_sampler_params"))
    c.pop("do_visualize_plan", None)
    c.pop("verbose", None)
    c.update(s)
    c["action_low"] = np.asarray(case["low"], np.float32)
    c["action_high"] = np.asarray(case["high"], np.float32)
    return c


# ---- MpcRandom (controllers/mpc.py:86-138) -----------------------------------------------------------------
RANDOM_CASES = {
    "random_cheetah": dict(
        model=lambda: DenseTanhModel.appendix_c(seed=4), cost="halfcheetah", penalise_flipping=True,
        low=-np.ones(6), high=np.ones(6),
        ctrl=dict(num_simulated_trajectories=50, horizon=30, cost_along_trajectory="sum",
                  action_sampler_params=dict(action_change_frequency=4), do_visualize_plan=False, verbose=False),
        start_obs=0.1 * np.random.RandomState(1020).randn(17), seed=20, steps=3),
    "random_bounds_final": dict(
        model=lambda: DenseTanhModel.appendix_c(seed=6), cost="halfcheetah", penalise_flipping=False,
        low=np.array([-1, -0.5, -1, -2, -1, -1.0]), high=np.array([1, 0.5, 2, 1, 1, 1.0]),
        ctrl=dict(num_simulated_trajectories=33, horizon=7, cost_along_trajectory="final",
                  action_sampler_params=dict(action_change_frequency=0), do_visualize_plan=False, verbose=False),
        start_obs=0.1 * np.random.RandomState(1021).randn(17), seed=21, steps=4),
}


def random_config(case):
    c = dict(case["ctrl"])
    s = dict(c.pop("action_sampler_params"))
    c.pop("do_visualize_plan", None)
    c.pop("verbose", None)
    c.update(s)
    c["action_low"] = np.asarray(case["low"], np.float32)
    c["action_high"] = np.asarray(case["high"], np.float32)
    return c
