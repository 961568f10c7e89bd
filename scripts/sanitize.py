#!/usr/bin/env python
"""Small instance of every kernel of the library, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python scripts/sanitize.py
    compute-sanitizer --tool racecheck python scripts/sanitize.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icem_b200 import workloads  # noqa: E402
from icem_b200.planner import Planner, PlannerSettings  # noqa: E402

which = sys.argv[1:] or ["dense", "cheetah", "humanoid", "mlp", "cemstd", "random", "ops", "sampler", "trainer", "reacher"]
if "dense" in which:
    name = "dense_tanh_cheetah_n4096"
    w = workloads.get_workload(name)
    p = Planner(workloads.planner_settings(name, scale_population=1 / 32))
    p.set_dense_model(*workloads.dense_model_weights(*w["dense"]))
    p.begin_rollout()
    for _ in range(2):
        p.plan(workloads.start_state(name))
    p.close()
    print("dense ok")
for key, name in (("cheetah", "halfcheetah_gt_n4096"), ("humanoid", "humanoid_standup_gt_n16384")):
    if key in which:
        p = Planner(workloads.planner_settings(name, scale_population=1 / 128))
        p.begin_rollout()
        st = workloads.start_state(name)
        for _ in range(2):
            a = p.plan(st)
        p.rollout_observations(st, p.elites()[0][:2], 17 if key == "cheetah" else 47)
        p.close()
        print(key, "ok")
if "mlp" in which:
    name = "mlp_cheetah_n65536"
    w = workloads.get_workload(name)
    p = Planner(workloads.planner_settings(name, scale_population=1 / 128))
    p.set_mlp_model(*workloads.mlp_model_weights(*w["mlp"]))
    p.begin_rollout()
    for _ in range(2):
        p.plan(workloads.start_state(name))
    p.close()
    print("mlp ok")
for key, extra in (("cemstd", dict(planner="cem_std", opt_iterations=2)),
                   ("random", dict(planner="random", opt_iterations=1, action_change_frequency=3))):
    if key in which:
        p = Planner(PlannerSettings(horizon=10, num_simulated_trajectories=64, action_low=-np.ones(6),
                                    action_high=np.ones(6), obs_dim=17, **extra))
        p.set_dense_model(0.5 * np.eye(17), 0.1 * np.ones((17, 6)))
        p.begin_rollout()
        for _ in range(2):
            p.plan(np.zeros(17))
        p.close()
        print(key, "ok")
if "ops" in which:
    rs = np.random.RandomState(0)
    p = Planner(PlannerSettings(horizon=30, num_simulated_trajectories=64, action_low=-np.ones(17),
                                action_high=np.ones(17), obs_dim=17, noise_beta=2.0))
    p.set_dense_model(0.5 * np.eye(17), 0.1 * np.ones((17, 17)))
    p.op_sample(rs.randn(100, 17, 16), rs.randn(100, 17, 16), np.zeros((30, 17)), np.ones((30, 17)))
    p.op_topk(rs.randn(5000).astype(np.float32), 10)
    p.close()
    print("ops ok")
if "sampler" in which:
    # stand-alone sampler, production noise: the three compile-time-shaped kernels and the run-time-shaped one, several
    # row batches per CTA
    for h, d in ((30, 17), (30, 6), (12, 6), (20, 5)):
        p = Planner(PlannerSettings(horizon=h, num_simulated_trajectories=64, action_low=-np.ones(d),
                                    action_high=np.ones(d), obs_dim=17, noise_beta=1.0))
        p.set_dense_model(0.5 * np.eye(17), 0.1 * np.ones((17, d)))
        p.begin_rollout()
        p.plan(np.zeros(17))
        p.bench_op("sample", 20011, reps=1, flush_l2=False)
        p.close()
    print("sampler ok")
if "trainer" in which:
    from icem_b200.trainer import MlpTrainer, epoch_indices
    rs = np.random.RandomState(0)
    ws, bs = workloads.mlp_model_weights(18, 6, 64, 1)
    tr = MlpTrainer(24, 64, 18)
    tr.set_weights(ws, bs)
    tr.set_data(rs.randn(700, 24), rs.randn(700, 18))
    tr.fit(epoch_indices(700, 300, 2, 0), lr=1e-3, weight_decay=1e-3)
    tr.predict(rs.randn(33, 24))
    tr.close()
    print("trainer ok")
if "reacher" in which:
    from icem_b200.robots import get_model
    m = get_model("reacher")
    p = Planner(PlannerSettings(horizon=10, num_simulated_trajectories=64, action_low=-np.ones(2), action_high=np.ones(2),
                                dynamics="articulated", articulated_model=m, obs_offset=0, cost="reacher",
                                cost_params=dict(reach=(0.1, 0.11, 0.0, 0.0)), obs_dim=11, opt_iterations=2))
    p.begin_rollout()
    st = np.array([0.3, -0.5, 0.1, -0.1, 0.0, 0.0, 0.0, 0.0])
    for _ in range(2):
        p.plan(st)
    p.sim_step(st, np.array([0.5, -0.5]))
    p.close()
    print("reacher ok")
