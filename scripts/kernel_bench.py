#!/usr/bin/env python
"""Per-kernel timings on the GPU box: sampler only / fused sample->rollout->cost / select+refit, CUDA events, L2
flushed between launches; HBM roofline fractions from the algorithmic bytes of SURVEY 8(d).

    python scripts/kernel_bench.py > profiles/r1_kernels.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icem_b200 import workloads  # noqa: E402
from icem_b200.planner import Planner  # noqa: E402


def main():
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    out = {"peak_hbm_gbs": peak, "kernels": []}
    for name, n in (("dense_tanh_humanoid_n16384", 262144), ("dense_tanh_cheetah_n4096", 262144),
                    ("humanoid_standup_gt_n16384", 16384), ("halfcheetah_gt_n4096", 16384)):
        w = workloads.get_workload(name)
        s = workloads.planner_settings(name, scale_population=max(1, n // w["settings"]["num_simulated_trajectories"]))
        p = Planner(s)
        if w.get("dense"):
            p.set_dense_model(*workloads.dense_model_weights(*w["dense"]))
        p.begin_rollout()
        p.plan(workloads.start_state(name))
        h, d = s.horizon, w["act_dim"]
        b_act = 4 * h * d
        for op, bytes_per in (("sample", b_act), ("fused", b_act + 4), ("select", 4)):
            if op == "sample" and not w.get("dense"):
                continue
            ms = p.bench_op(op, n, reps=10)
            gbs = bytes_per * n / (ms * 1e-3) / 1e9
            out["kernels"].append({"workload": name, "kernel": op, "n": n, "ms": ms, "trajectories_per_s": n / (ms * 1e-3),
                                   "algorithmic_bytes_per_trajectory": bytes_per, "achieved_gbs": gbs,
                                   "frac_of_measured_hbm": gbs / peak})
        p.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
