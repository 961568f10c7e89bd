"""CPU tests: the NumPy oracle (oracle/icem_np.py) against fixtures recorded from the UNMODIFIED
reference controller (tests/golden/*.npz, made by oracle/make_golden.py), plus a live re-check
when /root/reference is present."""
import os

import numpy as np
import pytest

from oracle import cases, costs_np, ref_loader
from oracle.icem_np import ICemConfig, ICemOracle, population_schedule, trajectories_per_plan_step
from oracle.shims import colorednoise as cn

COSTS = {"halfcheetah": costs_np.halfcheetah_cost, "humanoid_standup": costs_np.humanoid_standup_cost}


def run_oracle_case(case, record_actions=False):
    model = case["model"]()
    cfg = ICemConfig(**cases.controller_config(case))
    if case["cost"] == "halfcheetah":
        cost = lambda o, a: costs_np.halfcheetah_cost(o, a, case["penalise_flipping"])
    else:
        cost = COSTS[case["cost"]]
    orc = ICemOracle(cfg, model.rollout, cost, record_actions=record_actions)
    np.random.seed(case["seed"])
    obs = np.asarray(case["start_obs"], np.float64).copy()
    orc.beginning_of_rollout()
    traces = []
    for _ in range(case["steps"]):
        tr = orc.get_action(obs)
        traces.append(tr)
        obs = model.step(obs[None], tr.action[None])[0]
    return traces, float(np.random.randn())


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_oracle_matches_reference_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    traces, next_randn = run_oracle_case(cases.CASES[name])
    assert next_randn == float(g["next_randn"])          # identical RNG consumption
    assert len(traces) == int(g["num_steps"])
    for s, tr in enumerate(traces):
        np.testing.assert_allclose(tr.action, g[f"s{s}_action"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(tr.mean_after_shift, g[f"s{s}_mean_after_shift"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(tr.std_after_reset, g[f"s{s}_std_after_reset"], rtol=0, atol=1e-12)
        assert len(tr.iterations) == int(g[f"s{s}_num_iters"])
        for i, it in enumerate(tr.iterations):
            np.testing.assert_allclose(it.costs, g[f"s{s}_i{i}_costs"], rtol=0, atol=1e-10)
            np.testing.assert_array_equal(it.elite_idx, g[f"s{s}_i{i}_elite_idx"])   # bit-exact, in order
            np.testing.assert_allclose(it.mean, g[f"s{s}_i{i}_mean"], rtol=0, atol=1e-12)
            np.testing.assert_allclose(it.std, g[f"s{s}_i{i}_std"], rtol=0, atol=1e-12)


def test_appendix_c_known_answers(golden_dir):
    """SURVEY Appendix C numbers, typed in independently of the fixture file."""
    traces, next_randn = run_oracle_case(cases.CASES["appendix_c"])
    np.testing.assert_allclose(
        traces[0].action, [0.00573912, 0.03802868, 0.12162141, 0.30960938, -0.1226788, -0.08239668], atol=1e-8)
    assert [it.population for it in traces[0].iterations] == [40, 35, 28]
    assert [it.population for it in traces[1].iterations] == [43, 35, 28]
    assert list(traces[0].iterations[0].elite_idx) == [16, 20, 8, 29, 31, 33, 15, 0, 28, 21]
    assert list(traces[2].iterations[0].elite_idx) == [40, 41, 42, 17, 2, 0, 34, 35, 11, 33]
    np.testing.assert_allclose(traces[2].iterations[2].costs.min(), -13.22418276, atol=1e-7)
    assert abs(next_randn - (-1.409332708115941)) < 1e-15


@pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")
def test_oracle_matches_live_reference():
    from oracle import ref_harness
    case = cases.CASES["cheetah_n128"]
    steps, next_ref = ref_harness.run_reference_episode(
        case["model"](), case["cost"], case["ctrl"], case["low"], case["high"], case["start_obs"],
        case["seed"], 2, case["penalise_flipping"])
    short = dict(case, steps=2)
    traces, next_orc = run_oracle_case(short)
    assert next_ref == next_orc
    for st, tr in zip(steps, traces):
        np.testing.assert_allclose(tr.action, st["action"], atol=1e-12)
        for a, b in zip(tr.iterations, st["iterations"]):
            np.testing.assert_array_equal(a.elite_idx, b["elite_idx"])
            np.testing.assert_allclose(a.costs, b["costs"], atol=1e-10)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")
def test_costs_match_reference_cost_fn():
    ref_loader.install_reference()
    import environments.mujoco as m

    class Dummy:
        penalise_flipping = True
    rs = np.random.RandomState(0)
    for od in (17, 18):
        o = 2.5 * rs.randn(7, 30, od)
        a = rs.randn(7, 30, 6)
        np.testing.assert_array_equal(m.HalfCheetahMaybeWithPosition.cost_fn(Dummy(), o, a),
                                      costs_np.halfcheetah_cost(o, a, True))
    o = rs.randn(7, 30, 378)
    a = rs.randn(7, 30, 17)
    np.testing.assert_array_equal(m.HumanoidStandup.cost_fn(Dummy(), o, a, None),
                                  costs_np.humanoid_standup_cost(o, a))


def test_population_schedule_configs():
    """SURVEY section 8 population sizes for the BASELINE configs."""
    def sched(n, iters):
        return population_schedule(ICemConfig(horizon=30, num_simulated_trajectories=n, action_low=[-1], action_high=[1],
                                              factor_decrease_num=1.25, opt_iterations=iters))
    assert sched(128, 3) == [128, 102, 81]
    assert sched(4096, 5) == [4096, 3276, 2620, 2096, 1676]
    assert sched(16384, 3) == [16384, 13107, 10485]
    assert sched(262144, 3) == [262144, 209715, 167772]
    assert sched(40, 3) == [40, 32, 25]
    cfg = ICemConfig(horizon=30, num_simulated_trajectories=40, action_low=[-1], action_high=[1],
                     factor_decrease_num=1.25, opt_iterations=3)
    assert trajectories_per_plan_step(cfg, True) == 97
    assert trajectories_per_plan_step(cfg, False) == 100


def test_colorednoise_draw_equivalence_and_constants():
    """normal(scale=s) == standard_normal*s bit-for-bit; Appendix A constants; matrix form == irfft."""
    s, sigma = cn.spectrum_scale(0.25, 30)
    np.testing.assert_allclose(s[:4], [1.5298, 1.5298, 1.4029, 1.3335], atol=5e-5)
    assert abs(sigma - 0.309708) < 1e-6
    assert abs(cn.spectrum_scale(2.0, 30)[1] - 2.511658) < 1e-6
    assert abs(cn.spectrum_scale(0.25, 12)[1] - 0.462438) < 1e-6
    np.random.seed(5)
    a = np.random.normal(scale=s, size=(3, 4, 16))
    np.random.seed(5)
    b = np.random.standard_normal((3, 4, 16)) * s
    np.testing.assert_array_equal(a, b)
    np.random.seed(9)
    rec_prev, cn.RECORDER = cn.RECORDER, []
    y = cn.powerlaw_psd_gaussian(0.25, (5, 6, 30))
    (zr, zi), = cn.RECORDER
    cn.RECORDER = rec_prev
    np.testing.assert_array_equal(y, cn.synthesize(zr, zi, 0.25, 30))
    assert 0.9 < y.var() < 1.2


def test_min_trajectories_error():
    with pytest.raises(ValueError, match="At least two trajectories"):
        ICemConfig(horizon=3, num_simulated_trajectories=1, action_low=[-1], action_high=[1])


# ---- vanilla CEM (MpcCemStd) ----------------------------------------------------------------------
def _run_cem_std_case(case, record_actions=False):
    from oracle.cem_std_np import CemStdConfig, CemStdOracle
    model = case["model"]()
    cfg = CemStdConfig(**cases.cem_std_config(case))
    if case["cost"] == "halfcheetah":
        cost = lambda o, a: costs_np.halfcheetah_cost(o, a, case["penalise_flipping"])
    else:
        cost = COSTS[case["cost"]]
    orc = CemStdOracle(cfg, model.rollout, cost, record_actions=record_actions)
    np.random.seed(case["seed"])
    obs = np.asarray(case["start_obs"], np.float64).copy()
    orc.beginning_of_rollout()
    traces = []
    for _ in range(case["steps"]):
        tr = orc.get_action(obs)
        traces.append(tr)
        obs = model.step(obs[None], tr.action[None])[0]
    return traces, float(np.random.randn())


@pytest.mark.parametrize("name", sorted(cases.CEM_STD_CASES))
def test_cem_std_oracle_matches_reference_golden(name, golden_dir):
    """oracle/cem_std_np.py against fixtures recorded from the UNMODIFIED reference MpcCemStd."""
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    traces, next_randn = _run_cem_std_case(cases.CEM_STD_CASES[name])
    assert next_randn == float(g["next_randn"])
    for s, tr in enumerate(traces):
        np.testing.assert_allclose(tr.action, g[f"s{s}_action"], rtol=0, atol=1e-10)
        np.testing.assert_allclose(tr.mean_after_shift, g[f"s{s}_mean_after_shift"], rtol=0, atol=1e-10)
        np.testing.assert_allclose(tr.std_after_reset, g[f"s{s}_std_after_reset"], rtol=0, atol=1e-10)
        for i, it in enumerate(tr.iterations):
            np.testing.assert_allclose(it.costs, g[f"s{s}_i{i}_costs"], rtol=0, atol=1e-8)
            np.testing.assert_array_equal(it.elite_idx, g[f"s{s}_i{i}_elite_idx"])
            np.testing.assert_allclose(it.mean, g[f"s{s}_i{i}_mean"], rtol=0, atol=1e-10)
            np.testing.assert_allclose(it.std, g[f"s{s}_i{i}_std"], rtol=0, atol=1e-10)


def test_truncnorm_rvs_is_ppf_of_uniform_draws():
    """What "identical RNG state" means for MpcCemStd: truncnorm.rvs == ppf(np.random.uniform) draw for draw."""
    from scipy.stats import truncnorm
    lo, hi = np.array([-1.5, -0.2, 0.3]), np.array([0.5, 2.0, 4.0])
    np.random.seed(4)
    a = truncnorm.rvs(lo, hi, loc=0.1, scale=0.7, size=(5, 3))
    np.random.seed(4)
    u = np.random.uniform(size=(5, 3))
    np.testing.assert_allclose(a, truncnorm.ppf(u, lo, hi) * 0.7 + 0.1, rtol=0, atol=1e-14)


# ---- random shooting (MpcRandom) ------------------------------------------------------------------
def _run_random_case(case, record_actions=False):
    from oracle.random_np import RandomConfig, RandomOracle
    model = case["model"]()
    cost = lambda o, a: costs_np.halfcheetah_cost(o, a, case["penalise_flipping"])
    np.random.seed(case["seed"])                    # the constructor draws (random.py:8, mpc.py:90)
    orc = RandomOracle(RandomConfig(**cases.random_config(case)), model.rollout, cost, record_actions=record_actions)
    obs = np.asarray(case["start_obs"], np.float64).copy()
    orc.beginning_of_rollout()
    traces = []
    for _ in range(case["steps"]):
        tr = orc.get_action(obs)
        traces.append(tr)
        obs = model.step(obs[None], tr.action[None])[0]
    return traces, float(np.random.randn())


@pytest.mark.parametrize("name", sorted(cases.RANDOM_CASES))
def test_random_oracle_matches_reference_golden(name, golden_dir):
    """oracle/random_np.py against fixtures recorded from the reference MpcRandom (+ the one no-op method it lacks):
    the sampled populations (held actions running across rows and plan steps), costs, best index, executed action
    and the total number of RNG draws."""
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    traces, next_randn = _run_random_case(cases.RANDOM_CASES[name], record_actions=True)
    assert next_randn == float(g["next_randn"])
    for s, tr in enumerate(traces):
        np.testing.assert_array_equal(tr.action.astype(np.float32), g[f"s{s}_action"].astype(np.float32))
        it = tr.iterations[0]
        np.testing.assert_array_equal(it.actions.astype(np.float32), g[f"s{s}_i0_actions"])
        np.testing.assert_allclose(it.costs, g[f"s{s}_i0_costs"], rtol=0, atol=1e-10)
        np.testing.assert_array_equal(it.elite_idx, g[f"s{s}_i0_elite_idx"])
        # the recorded uniforms reproduce the actions: what the device is fed in parity mode
        u, _ = it.noise[0]
        c = cases.random_config(cases.RANDOM_CASES[name])
        lo, hi = c["action_low"].astype(np.float64), c["action_high"].astype(np.float64)
        np.testing.assert_array_equal((lo + (hi - lo) * u).astype(np.float32), g[f"s{s}_i0_actions"])


@pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")
@pytest.mark.parametrize("fuzz_seed", range(24))
def test_oracle_matches_live_reference_on_random_configurations(fuzz_seed):
    """Beyond the committed fixtures: seeded random controller settings (population, horizon incl. odd, action dim,
    iterations, elite count / reuse fraction, all three feature flags, white and coloured noise, decay factor, cost
    reduction, momentum) run through the imported reference and through oracle/icem_np.py: identical RNG consumption,
    elite lists, costs and executed actions over 3 closed-loop steps."""
    import warnings
    from oracle import ref_harness
    from oracle.dynamics_np import DenseTanhModel
    rs = np.random.RandomState(100 + fuzz_seed)
    d = int(rs.randint(2, 6))
    sampler = dict(alpha=float(rs.choice([0.0, 0.1, 0.5])), elites_size=int(rs.randint(2, 12)),
                   fraction_elites_reused=float(rs.choice([0.0, 0.3, 0.5, 1.0])), init_std=float(rs.uniform(0.2, 0.8)),
                   keep_previous_elites=bool(rs.randint(2)), shift_elites_over_time=bool(rs.randint(2)),
                   use_mean_actions=bool(rs.randint(2)), opt_iterations=int(rs.randint(1, 5)),
                   noise_beta=float(rs.choice([0.0, 0.5, 1.0, 3.0])))
    ctrl = dict(num_simulated_trajectories=int(rs.randint(4, 80)), factor_decrease_num=float(rs.choice([1.0, 1.25, 2.0])),
                horizon=int(rs.randint(3, 17)), cost_along_trajectory=str(rs.choice(["sum", "best", "final"])),
                action_sampler_params=sampler, do_visualize_plan=False, verbose=False)
    a = 0.9 * np.eye(17) + 0.05 * rs.randn(17, 17)
    b = 0.3 * rs.randn(17, d)
    model_fn = lambda: DenseTanhModel(a, b, np.zeros(17))
    lo = -rs.uniform(0.3, 1.5, d)
    hi = rs.uniform(0.3, 1.5, d)
    case = dict(model=model_fn, cost="halfcheetah", penalise_flipping=bool(rs.randint(2)), low=lo, high=hi, ctrl=ctrl,
                start_obs=0.2 * rs.randn(17), seed=int(rs.randint(1000)), steps=3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # "Number of trajectories is too low ... Setting num_elites to 2"
        steps, next_ref = ref_harness.run_reference_episode(
            case["model"](), case["cost"], case["ctrl"], case["low"], case["high"], case["start_obs"], case["seed"],
            case["steps"], case["penalise_flipping"])
        traces, next_orc = run_oracle_case(case)
    assert next_ref == next_orc, (ctrl, "RNG consumption differs")
    for st, tr in zip(steps, traces):
        np.testing.assert_allclose(tr.action, st["action"], atol=1e-12, err_msg=str(ctrl))
        np.testing.assert_allclose(tr.mean_after_shift, st["mean_after_shift"], atol=1e-12)
        assert len(tr.iterations) == len(st["iterations"])
        for x, y in zip(tr.iterations, st["iterations"]):
            np.testing.assert_allclose(x.costs, y["costs"], atol=1e-10)
            if len(np.unique(np.round(y["costs"], 12))) == len(y["costs"]):      # no exact ties: order is defined
                np.testing.assert_array_equal(x.elite_idx, y["elite_idx"])
            np.testing.assert_allclose(x.mean, y["mean"], atol=1e-12)
            np.testing.assert_allclose(x.std, y["std"], atol=1e-12)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")
@pytest.mark.parametrize("fuzz_seed", range(8))
def test_cem_std_and_random_oracles_match_live_reference_on_random_configurations(fuzz_seed):
    """Same live fuzzing for the two baseline controllers (mpc.py::MpcCemStd, mpc.py::MpcRandom)."""
    from oracle import ref_harness
    from oracle.dynamics_np import DenseTanhModel
    rs = np.random.RandomState(300 + fuzz_seed)
    d = int(rs.randint(2, 6))
    a = 0.9 * np.eye(17) + 0.05 * rs.randn(17, 17)
    b = 0.3 * rs.randn(17, d)
    model_fn = lambda: DenseTanhModel(a, b, np.zeros(17))
    lo, hi = -rs.uniform(0.3, 1.5, d), rs.uniform(0.3, 1.5, d)
    base = dict(model=model_fn, cost="halfcheetah", penalise_flipping=bool(rs.randint(2)), low=lo, high=hi,
                start_obs=0.2 * rs.randn(17), seed=int(rs.randint(1000)), steps=3)
    h = int(rs.randint(3, 15))
    n = int(rs.randint(4, 60))
    reduce = str(rs.choice(["sum", "best", "final"]))
    # ---- MpcCemStd
    sampler = dict(alpha=float(rs.choice([0.0, 0.1, 0.4])), elites_size=int(rs.randint(2, 10)),
                   opt_iterations=int(rs.randint(1, 4)), init_std=float(rs.uniform(0.2, 0.8)),
                   shift_means=bool(rs.randint(2)), execute_best_elite=bool(rs.randint(2)),
                   bounds_like_levine=bool(rs.randint(2)))
    case = dict(base, ctrl=dict(num_simulated_trajectories=n, horizon=h, cost_along_trajectory=reduce,
                                action_sampler_params=sampler, do_visualize_plan=False, verbose=False))
    steps, next_ref = ref_harness.run_reference_episode(
        case["model"](), case["cost"], case["ctrl"], lo, hi, case["start_obs"], case["seed"], 3,
        case["penalise_flipping"], controller="MpcCemStd")
    traces, next_orc = _run_cem_std_case(case)
    assert next_ref == next_orc
    for st, tr in zip(steps, traces):
        np.testing.assert_allclose(tr.action, st["action"], atol=1e-10, err_msg=str(sampler))
        for x, y in zip(tr.iterations, st["iterations"]):
            np.testing.assert_allclose(x.costs, y["costs"], atol=1e-8)
            np.testing.assert_allclose(x.mean, y["mean"], atol=1e-10)
            np.testing.assert_allclose(x.std, y["std"], atol=1e-10)
    # ---- MpcRandom
    case = dict(base, ctrl=dict(num_simulated_trajectories=n, horizon=h, cost_along_trajectory=reduce,
                                action_sampler_params=dict(action_change_frequency=int(rs.randint(0, h))),
                                do_visualize_plan=False, verbose=False))
    steps, next_ref = ref_harness.run_reference_episode(
        case["model"](), case["cost"], case["ctrl"], lo, hi, case["start_obs"], case["seed"], 3,
        case["penalise_flipping"], controller="MpcRandom")
    traces, next_orc = _run_random_case(case, record_actions=True)
    assert next_ref == next_orc
    for st, tr in zip(steps, traces):
        np.testing.assert_array_equal(tr.action.astype(np.float32), st["action"].astype(np.float32))
        np.testing.assert_array_equal(tr.iterations[0].actions.astype(np.float32),
                                      st["iterations"][0]["actions"].astype(np.float32))
        np.testing.assert_allclose(tr.iterations[0].costs, st["iterations"][0]["costs"], atol=1e-10)
