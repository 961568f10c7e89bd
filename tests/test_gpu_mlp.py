"""GPU parity of the tensor-core MLP rollout (csrc/mlp_rollout.cuh: tcgen05.mma + TMEM, fp16 operands, fp32
accumulation) against (a) the NumPy oracle that restates the operand roundings (oracle/dynamics_np.py::MlpModel) and
(b) the fp32 model without any restatement (MlpModelF32 == torch fp32 nn.Sequential): the precision reference.

Tolerance of (a): operand resolution (written for the round-1 bf16 kernel; the fp16 kernel sits well inside it).  The oracle reproduces every bf16 rounding point but not the tensor core's accumulation
order nor tanh.approx (relative error ~5e-4); a value that lands within that error of a bf16 rounding boundary
rounds the other way (1 bf16 ulp = 0.4 %), so per-step observations agree to ~1e-2 absolute after 11 steps of an
O(1) state and h=12 costs to a few 1e-2, not to fp32 precision.  Elite SETS must agree except for candidates whose
cost gap to the k-th elite is inside that band."""
import numpy as np
import pytest

from oracle import costs_np
from oracle.dynamics_np import MlpModel, bf16_round
from oracle.icem_np import reduce_costs

pytestmark = pytest.mark.gpu


def _planner(hidden, n=256, h=12, obs_dim=18, d=6, seed=21):
    from icem_b200 import workloads
    from icem_b200.planner import Planner, PlannerSettings
    ws, bs = workloads.mlp_model_weights(obs_dim, d, hidden, seed)
    p = Planner(PlannerSettings(horizon=h, num_simulated_trajectories=n, action_low=-np.ones(d, np.float32),
                                action_high=np.ones(d, np.float32), dynamics="mlp", cost="halfcheetah",
                                obs_dim=obs_dim, penalise_flipping=True, factor_decrease_num=1.25,
                                noise_beta=0.25, keep_iteration_actions=True))
    p.set_mlp_model(ws, bs)
    return p, MlpModel(ws, bs)


def test_bf16_round_helper():
    x = np.array([1.0, 1.00390625, 1.001953125, -2.5, 3.1415927, 1e-8, 0.0], np.float32)
    import torch
    ref = torch.tensor(x).to(torch.bfloat16).to(torch.float32).numpy()
    np.testing.assert_array_equal(bf16_round(x).astype(np.float32), ref)


@pytest.mark.parametrize("hidden", [64, 128, 256])
def test_single_transition_matches_oracle(hidden):
    p, mod = _planner(hidden)
    rs = np.random.RandomState(0)
    for _ in range(5):
        st = 0.5 * rs.randn(18)
        a = rs.uniform(-1, 1, 6)
        got, _, _ = p.sim_step(st, a)
        ref = mod.step(st.astype(np.float32).astype(np.float64)[None], a.astype(np.float32).astype(np.float64)[None])[0]
        assert np.abs(got - ref).max() <= 5e-3, np.abs(got - ref).max()
    p.close()


@pytest.mark.parametrize("hidden,n", [(256, 1000), (128, 300), (64, 129)])
def test_tensor_core_rollout_costs_match_oracle(hidden, n):
    p, mod = _planner(hidden)
    rs = np.random.RandomState(1)
    acts = rs.uniform(-1, 1, (n, 12, 6)).astype(np.float32)
    start = (0.3 * rs.randn(18)).astype(np.float32).astype(np.float64)
    obs = mod.rollout(start, acts.astype(np.float64))
    ref = reduce_costs(costs_np.halfcheetah_cost(obs, acts.astype(np.float64), True), "sum")
    got = p.op_rollout_cost(start, acts)
    d = np.abs(got - ref)
    # exclude trajectories grazing the discontinuous flip threshold
    safe = np.all(np.abs(np.abs(obs[..., 2]) - np.pi / 2) > 5e-2, axis=1)
    assert safe.mean() > 0.7
    assert np.median(d[safe]) <= 1e-2, np.median(d[safe])
    assert np.max(d[safe]) <= 8e-2, np.max(d[safe])
    # ranking: the device's 10 best are all inside the oracle's best 10 + near-ties
    k = 10
    order = np.argsort(ref, kind="stable")
    band = ref[order[k - 1]] + 2 * np.max(d[safe])
    dev = np.argsort(got, kind="stable")[:k]
    assert np.all(ref[dev] <= band), (ref[dev], band)
    p.close()


def test_plan_step_runs_and_improves_cost():
    """End to end through the graph-captured plan step with the tensor-core model: iteration-wise best cost is
    non-increasing (kept elites) and the executed action is inside the bounds."""
    p, mod = _planner(256, n=2048)
    p.begin_rollout()
    st = 0.1 * np.random.RandomState(5).randn(18)
    for step in range(3):
        a = p.plan(st)
        assert a.shape == (6,) and np.all(np.abs(a) <= 1.0)
        best = [p.iteration_record(i)["elite_costs"][0] for i in range(3)]
        assert best[1] <= best[0] + 1e-6 and best[2] <= best[1] + 1e-6
        # the elite costs the device reports agree with the oracle's evaluation of the same action sequences
        acts, costs, _ = p.elites()
        obs = mod.rollout(st.astype(np.float32).astype(np.float64), acts.astype(np.float64))
        ref = reduce_costs(costs_np.halfcheetah_cost(obs, acts.astype(np.float64), True), "sum")
        assert np.abs(ref - costs).max() <= 8e-2
        st, _, _ = p.sim_step(st, a)
    p.close()


def test_series_sampler_inside_the_plan_step():
    """The MLP path samples with the stand-alone thread-per-series kernel (csrc/sampler.cuh).  Inside the graph-captured
    plan step it must honour: injected draws (icem.py:73-79), shifted elites appended after the fresh rows at
    iteration 0 of every later step (icem.py:91-104, 131-137), the mean written over row 0 on the last iteration
    (icem.py:87-88), and Philox production draws with the reference sampler's statistics."""
    from oracle.shims import colorednoise as cn
    h, d, n = 12, 6, 2048
    p, _ = _planner(256, n=n)
    p.begin_rollout()
    rs = np.random.RandomState(3)
    K = h // 2 + 1
    zr, zi = rs.standard_normal((n, d, K)), rs.standard_normal((n, d, K))
    p.inject_noise(0, zr, zi)
    for i in (1, 2):          # parity mode needs every iteration's draws
        ni = p.population_size(i, first_step=True)[1]
        p.inject_noise(i, rs.standard_normal((ni, d, K)), rs.standard_normal((ni, d, K)))
    st = 0.1 * rs.randn(18)
    a0 = p.plan(st)
    y = cn.synthesize(zr.copy(), zi.copy(), 0.25, h).transpose(0, 2, 1)
    ref = np.clip(y * 0.5 + 0.0, -1, 1)
    got = p.actions(0, n)
    assert np.abs(got - ref).max() <= 4e-6 + 1e-7
    # last iteration: row 0 is the mean the iteration sampled from (the refit result of the one before)
    n2 = p.population_size(2, first_step=True)[1]
    np.testing.assert_array_equal(p.actions(2, n2)[0], p.iteration_record(1)["mean"].astype(np.float32))
    elites, _, _ = p.elites()
    # second step, production noise: 3 shifted elites follow the n fresh rows
    st, _, _ = p.sim_step(st, a0)
    mean_before = p.mean().astype(np.float64)          # already time-shifted (icem.py:167-171)
    p.plan(st)
    rows = p.population_size(0, first_step=False)[1]
    assert rows == n + 3
    acts = p.actions(0, rows)
    np.testing.assert_array_equal(acts[n:, :-1], elites[:3, 1:])
    assert np.all(np.abs(acts[n:, -1]) <= 1.0) and np.abs(acts[n:, -1] - mean_before[-1]).max() > 1e-3
    # fresh rows: unit-variance colored noise around the shifted mean with the reset std 0.5, clipped
    np.random.seed(0)
    refn = cn.powerlaw_psd_gaussian(0.25, (n, d, h)).transpose(0, 2, 1)
    ref_a = np.clip(refn * 0.5 + mean_before, -1, 1)
    fresh = acts[:n].astype(np.float64)
    assert np.abs(fresh.mean(0) - ref_a.mean(0)).max() < 0.06
    assert np.abs(fresh.std(0) - ref_a.std(0)).max() < 0.05
    z, zref = fresh - mean_before, ref_a - mean_before
    ac = lambda x: np.mean(x[:, 1:] * x[:, :-1]) / np.mean(x * x)
    assert abs(ac(z) - ac(zref)) < 0.04
    assert abs((np.abs(fresh) == 1.0).mean() - (np.abs(ref_a) == 1.0).mean()) < 0.02
    # no two series share draws: correlation between action dims / neighbouring rows ~ 0
    assert abs(np.mean(z[:, :, 0] * z[:, :, 1]) / np.mean(z * z)) < 0.05
    assert abs(np.mean(z[1:, :, 0] * z[:-1, :, 0]) / np.mean(z * z)) < 0.05
    p.close()


@pytest.mark.parametrize("h,d,obs_dim,cost", [(30, 17, 8, "humanoid_standup"), (30, 6, 18, "halfcheetah"),
                                              (12, 6, 18, "halfcheetah"), (20, 5, 18, "halfcheetah")])
def test_production_normals_of_the_series_sampler(h, d, obs_dim, cost):
    """The unit normals behind the production draws of csrc/sampler.cuh (Philox4x32-7, three Box-Muller pairs per call
    from 21-bit fields, SFU approximations): undo the synthesis with an rfft and test the recovered per-bin real /
    imaginary parts for what `np.random.normal` gives the reference (icem.py:73-75 via colorednoise): zero mean, equal
    variance of real and imaginary parts, Gaussian shape (KS, kurtosis, 3- and 4-sigma tail mass), no correlation between
    bins, between the two members of a Box-Muller pair, or between the series of one call group.  The first three
    shapes run the compile-time-shaped kernels, the last one the run-time-shaped one."""
    from scipy import stats
    from icem_b200 import workloads
    from icem_b200.planner import Planner, PlannerSettings
    n = 8192
    ws, bs = workloads.mlp_model_weights(obs_dim, d, 64, 3)
    p = Planner(PlannerSettings(horizon=h, num_simulated_trajectories=n, action_low=-np.ones(d, np.float32),
                                action_high=np.ones(d, np.float32), dynamics="mlp", cost=cost, obs_dim=obs_dim,
                                noise_beta=1.0, init_std=0.02, opt_iterations=1, use_mean_actions=False, seed=11,
                                keep_iteration_actions=True))
    p.set_mlp_model(ws, bs)
    p.begin_rollout()
    p.plan(np.zeros(obs_dim))
    y = p.actions(0, n).astype(np.float64) / 0.02          # [n, h, d]: nothing is clipped at 50 sigma
    assert np.abs(y).max() < 40
    Y = np.fft.rfft(y.transpose(0, 2, 1), axis=-1)          # [n, d, K]
    K = h // 2 + 1
    re, im = Y.real, Y.imag
    assert np.abs(im[..., 0]).max() < 1e-3 and np.abs(im[..., -1]).max() < 1e-3
    comps = [re[..., k] for k in range(K)] + [im[..., k] for k in range(1, K - 1)]      # the h normals of a series
    z = np.stack([c / c.std() for c in comps], axis=-1)     # [n, d, h] unit normals up to the known per-bin scale
    for k in range(1, K - 1):                               # real and imaginary parts carry the same scale
        assert abs(re[..., k].std() / im[..., k].std() - 1) < 0.02
    flat = z.reshape(-1)
    assert abs(flat.mean()) < 4 / np.sqrt(flat.size)
    assert abs(stats.kurtosis(flat)) < 0.03
    assert abs(stats.skew(flat)) < 0.01
    for thr, mass in ((3.0, 2.6998e-3), (4.0, 6.334e-5)):
        got = np.mean(np.abs(flat) > thr)
        assert abs(got - mass) < 5 * np.sqrt(mass / flat.size) + 0.02 * mass, (thr, got)
    assert stats.kstest(flat[::7], "norm").pvalue > 1e-3
    # independence: between bins, between the members of a pair (real / imaginary part of one bin), between the action
    # dims of a trajectory and between neighbouring trajectories
    c = np.corrcoef(z.reshape(-1, h), rowvar=False)
    assert np.abs(c - np.eye(h)).max() < 5 / np.sqrt(n * d)
    assert abs(np.mean(z[:, 0, :] * z[:, 1, :])) < 5 / np.sqrt(n * h)
    assert abs(np.mean(z[1:, 0, :] * z[:-1, 0, :])) < 5 / np.sqrt(n * h)
    # squared magnitudes of a pair are independent exponentials (radius and angle come from disjoint bit fields)
    r2 = z[..., 1:K - 1] ** 2 + z[..., K:] ** 2
    ang = np.arctan2(z[..., K:], z[..., 1:K - 1])
    assert abs(np.corrcoef(r2.reshape(-1), np.cos(ang).reshape(-1))[0, 1]) < 5 / np.sqrt(r2.size)
    assert stats.kstest(r2.reshape(-1)[::5] / 2, "expon").pvalue > 1e-3
    assert stats.kstest((ang.reshape(-1)[::5] + np.pi) / (2 * np.pi), "uniform").pvalue > 1e-3
    p.close()


def test_tensor_core_ranking_against_the_fp32_model():
    """The precision reference of the tensor-core path is the fp32 model (oracle/dynamics_np.py::MlpModelF32 == a torch
    fp32 nn.Sequential), NOT the oracle that restates the kernel's operand roundings: along 4 closed-loop steps of the
    fp32 plan the device scores the same candidates at every CEM iteration (scripts/mlp_precision_report.py) and must
    pick essentially the same elites and the same best candidate.  Full-size numbers (N = 65536, 20 steps) are in
    profiles/r2_mlp_f16_vs_fp32.json."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    from mlp_precision_report import run
    rep = run(4096, 4, quiet=True)
    assert rep["cost_err_median"] <= 5e-3, rep["cost_err_median"]
    assert rep["elite_overlap"]["mean"] >= 0.85, rep["elite_overlap"]
    assert rep["executed_action_abs_diff"]["same_best_candidate_share"] >= 0.75, rep["executed_action_abs_diff"]
