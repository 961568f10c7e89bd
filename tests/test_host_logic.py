"""CPU tests of the host-side logic above the C ABI that needs no device: the settings -> icem_config_t mapping, the
population schedule, the stand-in envs' observation layouts / cost specs, and bench.py's peak lookup."""
import importlib.util
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_population_schedule_matches_the_reference_decay():
    """icem/controllers/icem.py:126-127 (compounded truncation) at the BASELINE configs."""
    from icem_b200 import workloads
    expect = {"halfcheetah_gt_n4096": [4096, 3276, 2620, 2096, 1676],
              "humanoid_standup_gt_n16384": [16384, 13107, 10485],
              "mlp_cheetah_n65536": [65536, 52428, 41942]}
    for name, pops in expect.items():
        s = workloads.planner_settings(name)
        assert workloads.populations(s) == pops
        assert workloads.trajectories_per_step(s, first_step=True) == sum(pops)
        assert workloads.trajectories_per_step(s, first_step=False) == sum(pops) + 3      # int(10 * 0.3) shifted elites
    s = workloads.planner_settings("humanoid_standup_gt_n16384", scale_population=16)        # BASELINE configs[4]
    assert workloads.populations(s) == [262144, 209715, 167772]


def test_cost_parameter_mapping():
    from icem_b200 import envs
    from icem_b200.planner import _cost_fields
    f = _cost_fields(dict(envs.Hopper.cost_params, dt=envs.Hopper.dt))
    assert f["cost_z_index"] == 1 and f["cost_z_strict"] == 1 and f["cost_velocity_index1"] == 0
    assert f["cost_dt"] == 0.008 and f["cost_state_bound"] == 100.0 and f["cost_unhealthy_weight"] == 200.0
    assert np.isfinite(np.float32(f["cost_z_hi"])) and f["cost_z_hi"] > 1e38          # open range end, float32-safe
    f = _cost_fields(dict(envs.Humanoid.cost_params, dt=envs.Humanoid.dt))
    assert f["cost_velocity_index1"] == 25 and f["cost_forward_weight"] == 1.25 and f["cost_z_strict"] == 1
    f = _cost_fields(None)
    assert f["cost_dt"] == 0.0 and f["cost_velocity_index1"] == 0 and f["cost_forward_weight"] == 1.0


def test_standin_env_layouts_without_a_device():
    """Construction, reset, observation widths and cost specs need no GPU (the device model is created lazily)."""
    from icem_b200 import envs
    widths = {"HalfCheetah": 17, "HumanoidStandup": 378, "Hopper": 12, "Ant": 113, "Humanoid": 378}   # reference widths
    for name, w in widths.items():
        env = envs.make_env(name)
        env.seed(1)
        ob = env.reset()
        assert ob.shape == (w,) == env.observation_space.shape
        st = env.get_GT_state()
        assert st[0] == 0.0 and len(st) == 1 + env.spec["nq"] + env.spec["nv"]          # [time, qpos, qvel]
        env.set_GT_state(st)
        np.testing.assert_array_equal(env.get_GT_state(), st)
        spec = env.cuda_cost_spec()
        assert spec[0] in ("halfcheetah", "humanoid_standup", "locomotion")
        if spec[0] == "locomotion":
            assert spec[2]["dt"] == env.dt and env.cuda_dynamics == "articulated"
            assert env.cuda_articulated_model().nu == env.action_space.shape[0]
        assert env.reward_is_negative_cost
    hop = envs.make_env("Hopper")
    hop._state = np.concatenate([np.zeros(6), [50.0, -50.0, 1.0, 2.0, 3.0, 4.0]])
    np.testing.assert_array_equal(hop._obs()[6:], [10.0, -10.0, 1.0, 2.0, 3.0, 4.0])     # gym clips the velocities
    ant = envs.make_env("Ant")
    assert np.all(ant.reset()[29:] == 0.0)


def test_bench_peak_lookup_tolerates_key_spellings(tmp_path, monkeypatch):
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    monkeypatch.setattr(b, "ROOT", str(tmp_path))
    assert b.read_peaks() == (6650.0, "fallback (B200_PROFILING.md)") and b.read_bf16_peak() == 1590.0
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps(
        {"hbm_gbs": 6547.2, "bf16_tflops_burst": 1645.3, "bf16_tflops_sustained": 1386.1}))
    assert b.read_peaks()[0] == 6547.2 and b.read_bf16_peak() == 1645.3
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps(
        {"hbm_copy_GBps": 6547.2, "bf16": {"burst_tflops": 1645.3, "sustained_tflops": 1386.1}}))
    assert b.read_peaks()[0] == 6547.2 and b.read_bf16_peak() == 1645.3


def test_importing_the_package_does_not_shadow_the_reference():
    """An unrelated `environments` namespace package exists in this image's site path; importing icem_b200 without
    the reference on sys.path must not leave it (or any other top-level name the reference uses) cached in
    sys.modules, or a later import of the reference would resolve to the wrong package."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import icem_b200.envs, icem_b200.models, icem_b200.controller\n"
            "bad = [m for m in sys.modules if m.split('.')[0] in ('environments', 'controllers', 'misc', 'models')]\n"
            "assert not bad, bad\n" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)


def test_trainer_minibatch_schedule_and_transition_extraction():
    """Host side of forward_model.train(rollout_buffer) (icem/main.py:209-210): minibatch rows and the (input, target)
    layout, no device involved."""
    from icem_b200 import api
    from icem_b200.trainer import epoch_indices, transitions_from_buffer
    idx = epoch_indices(1000, 256, 3, seed=4)
    assert idx.shape == (9, 256) and idx.dtype == np.int32
    for e in range(3):                                 # inside an epoch no row repeats
        assert np.unique(idx[3 * e:3 * e + 3]).size == 768
    np.testing.assert_array_equal(idx, epoch_indices(1000, 256, 3, seed=4))
    assert not np.array_equal(idx, epoch_indices(1000, 256, 3, seed=5))
    small = epoch_indices(10, 64, 2, seed=0)           # fewer rows than a batch: drawn with replacement
    assert small.shape == (2, 64) and small.min() >= 0 and small.max() < 10
    full = epoch_indices(10, 4, 1, seed=0, drop_last=False)
    assert full.shape == (3, 4) and set(full.ravel().tolist()) == set(range(10))
    rs = np.random.RandomState(0)
    rollouts = [api.EliteRollout(observations=rs.randn(5, 3), next_observations=rs.randn(5, 3), actions=rs.randn(5, 2),
                                 rewards=np.zeros(5)) for _ in range(4)]
    x, t = transitions_from_buffer(api.EliteBuffer(rollouts))
    assert x.shape == (20, 5) and t.shape == (20, 3)
    np.testing.assert_array_equal(x[5:10, :3], rollouts[1]["observations"])
    np.testing.assert_array_equal(x[5:10, 3:], rollouts[1]["actions"])
    np.testing.assert_array_equal(t[5:10], rollouts[1]["next_observations"] - rollouts[1]["observations"])
    import pytest
    with pytest.raises(ValueError):
        transitions_from_buffer(api.EliteBuffer([]))
