#!/usr/bin/env python
"""What the reduced-precision operands of the tensor-core MLP rollout cost the planner, against the fp32 model
(VERDICT r1 #8).

The fp32 plan -- the float64 NumPy restatement of MpcICem (oracle/icem_np.py, pinned to the reference) rolling out the
fp32 MLP (oracle/dynamics_np.py::MlpModelF32 == torch fp32 nn.Sequential) -- runs closed loop.  At EVERY CEM iteration
of every plan step the device scores the very same candidate action sequences with the tensor-core kernel
(icem_op_rollout_cost: tcgen05 fp16 x fp16 -> fp32, tanh.approx) from the same state, and the two rankings are compared:
  elite overlap   |top-k by device cost  &  top-k by fp32 cost| / k        (k = 10)
  action diff     |first action of the device's best candidate - of the fp32 best candidate|
  cost error      |device cost - fp32 cost| (median; candidates grazing the discontinuous flip penalty excluded)
Same candidates on both sides: the numbers isolate arithmetic precision from the drift two separately evolving
planners would add.

    python scripts/mlp_precision_report.py [--n 65536] [--steps 20] > profiles/r2_mlp_f16_vs_fp32.json
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
    orc = ICemOracle(cfg, mod.rollout, cost, record_actions=True)
    np.random.seed(seed)
    rs = np.random.RandomState(seed + 1)
    state = 0.1 * rs.randn(od)
    orc.beginning_of_rollout()
    k = cfg.num_elites
    rows = []
    for s in range(steps):
        t0 = time.time()
        tr = orc.get_action(state)
        st32 = state.astype(np.float32).astype(np.float64)
        for i, it in enumerate(tr.iterations):
            acts = np.asarray(it.actions, np.float32)            # every candidate the fp32 plan scored (kept elites too)
            ref = np.asarray(it.costs, np.float64)
            dev = p.op_rollout_cost(st32, acts).astype(np.float64)
            o_ref, o_dev = np.argsort(ref, kind="stable"), np.argsort(dev, kind="stable")
            overlap = len(set(o_ref[:k].tolist()) & set(o_dev[:k].tolist())) / float(k)
            err = np.abs(dev - ref)
            smooth = err < 5.0                                    # a flip-penalty crossing moves a cost by 10
            rows.append(dict(step=s, iteration=i, candidates=int(len(ref)), elite_overlap=overlap,
                             same_best=bool(o_ref[0] == o_dev[0]),
                             action_abs_diff_max=float(np.abs(acts[o_ref[0], 0] - acts[o_dev[0], 0]).max()),
                             cost_err_median=float(np.median(err[smooth])), cost_err_p99=float(np.percentile(err[smooth], 99)),
                             flip_crossings=int((~smooth).sum()),
                             elite_gap_k=float(ref[o_ref[k]] - ref[o_ref[k - 1]]),
                             seconds=time.time() - t0))
            if not quiet:
                sys.stderr.write(json.dumps(rows[-1]) + "\n")
        state = mod.step(state[None], tr.action[None])[0]
    p.close()
    ov = np.array([r["elite_overlap"] for r in rows])
    last = np.array([r["elite_overlap"] for r in rows if r["iteration"] == cfg.opt_iterations - 1])
    da = np.array([r["action_abs_diff_max"] for r in rows if r["iteration"] == cfg.opt_iterations - 1])
    return dict(population=n, steps=steps, hidden=hidden, horizon=h, operands="fp16 x fp16 -> fp32, tanh.approx",
                elite_overlap=dict(mean=float(ov.mean()), min=float(ov.min()), median=float(np.median(ov)),
                                   share_of_iterations_with_at_least_9_of_10=float((ov >= 0.9).mean())),
                elite_overlap_last_iteration=dict(mean=float(last.mean()), min=float(last.min())),
                executed_action_abs_diff=dict(median=float(np.median(da)), max=float(da.max()),
                                              same_best_candidate_share=float(np.mean(
                                                  [r["same_best"] for r in rows if r["iteration"] == cfg.opt_iterations - 1])),
                                              action_range=2.0),
                cost_err_median=float(np.median([r["cost_err_median"] for r in rows])),
                per_iteration=rows)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    print(json.dumps(run(a.n, a.steps), indent=1))
