#!/usr/bin/env python
"""What the bf16 operands of the tensor-core MLP rollout cost the planner, against the fp32 model.

Two planners run in lockstep on the SAME Gaussian draws (parity mode) from the same states, closed loop:
  fp32 plan : the float64 NumPy restatement of MpcICem (oracle/icem_np.py, pinned to the reference) rolling out the
              fp32 MLP (oracle/dynamics_np.py::MlpModelF32 == torch fp32 nn.Sequential);
  bf16 plan : the device planner, tcgen05 bf16 x bf16 -> fp32 rollout (csrc/mlp_rollout.cuh).
Per plan step: overlap of the last iteration's elite index sets (|E_dev & E_ref| / k), |executed action difference|,
cost error on the fp32 elites.  The shared state advances with the fp32 plan's action through the fp32 model.

    python scripts/mlp_precision_report.py [--n 65536] [--steps 20] > profiles/r2_mlp_bf16_vs_fp32.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n, steps, seed=0, hidden=256, quiet=False):
    from icem_b200 import workloads
    from icem_b200.planner import Planner, PlannerSettings
    from oracle import costs_np
    from oracle.dynamics_np import MlpModelF32
    from oracle.icem_np import ICemConfig, ICemOracle
    from tests.util import stack_noise
    od, ad, h = 18, 6, 12
    ws, bs = workloads.mlp_model_weights(od, ad, hidden, 21)
    low, high = -np.ones(ad, np.float32), np.ones(ad, np.float32)
    p = Planner(PlannerSettings(horizon=h, num_simulated_trajectories=n, action_low=low, action_high=high,
                                dynamics="mlp", cost="halfcheetah", obs_dim=od, penalise_flipping=True,
                                factor_decrease_num=1.25, noise_beta=0.25))
    p.set_mlp_model(ws, bs)
    mod = MlpModelF32(ws, bs)
    cfg = ICemConfig(horizon=h, num_simulated_trajectories=n, action_low=low, action_high=high,
                     factor_decrease_num=1.25, opt_iterations=3, noise_beta=0.25)
    cost = lambda o, a: costs_np.halfcheetah_cost(o, a, True)
    orc = ICemOracle(cfg, mod.rollout, cost)
    np.random.seed(seed)
    rs = np.random.RandomState(seed + 1)
    state = 0.1 * rs.randn(od)
    orc.beginning_of_rollout()
    p.begin_rollout()
    k = cfg.num_elites
    rows = []
    for s in range(steps):
        t0 = time.time()
        tr = orc.get_action(state)
        for i, it in enumerate(tr.iterations):
            zr, zi = stack_noise(it.noise)
            p.inject_noise(i, zr, zi)
        act = p.plan(state)
        last = tr.iterations[-1]
        rec = p.iteration_record(cfg.opt_iterations - 1)
        overlap = len(set(rec["elite_idx"].tolist()) & set(np.asarray(last.elite_idx).tolist())) / float(k)
        first = tr.iterations[0]
        rec0 = p.iteration_record(0)
        overlap0 = len(set(rec0["elite_idx"].tolist()) & set(np.asarray(first.elite_idx).tolist())) / float(k)
        rows.append(dict(step=s, elite_overlap_last_iteration=overlap, elite_overlap_first_iteration=overlap0,
                         action_abs_diff_max=float(np.abs(act - tr.action).max()),
                         best_cost_fp32=float(np.min(last.costs)), best_cost_bf16=float(rec["elite_costs"][0]),
                         seconds=time.time() - t0))
        if not quiet:
            sys.stderr.write(json.dumps(rows[-1]) + "\n")
        state = mod.step(state[None], tr.action[None])[0]
    p.close()
    ov = np.array([r["elite_overlap_last_iteration"] for r in rows])
    ov0 = np.array([r["elite_overlap_first_iteration"] for r in rows])
    da = np.array([r["action_abs_diff_max"] for r in rows])
    return dict(population=n, steps=steps, hidden=hidden, horizon=h,
                elite_overlap_last_iteration=dict(mean=float(ov.mean()), min=float(ov.min()), median=float(np.median(ov))),
                elite_overlap_first_iteration=dict(mean=float(ov0.mean()), min=float(ov0.min())),
                executed_action_abs_diff=dict(median=float(np.median(da)), max=float(da.max()),
                                              action_range=2.0),
                note="first iteration = both planners sample from the identical distribution; later iterations and "
                     "steps also carry the drift of the two planners' own mean / std",
                per_step=rows)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    print(json.dumps(run(a.n, a.steps), indent=1))
