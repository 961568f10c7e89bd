#!/usr/bin/env python
"""Closed-loop behaviour of the planner at BASELINE's populations on the device models: what the executed actions do to
the robot over an episode (the reference measures the same thing as the episode return, icem/main.py:195-200).

    python scripts/closed_loop_demo.py > profiles/r2_closed_loop_demo.json

HumanoidStandup (configs[2], N = 16384): height of the root over 100 env steps (cost = -height + 0.1 |a|^2).
HalfCheetah (configs[1], N = 4096): forward distance over 100 env steps (cost = -velocity + ...).
Each line: Euler and RK4 integrators; wall-clock per closed-loop step (plan + env transition, host buffers)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icem_b200 import workloads  # noqa: E402
from icem_b200.planner import Planner  # noqa: E402


def run(name, steps, watch):
    s = workloads.planner_settings(name, seed=1)
    p = Planner(s)
    p.begin_rollout()
    state = workloads.start_state(name, seed=1)
    trace, costs = [float(state[watch])], []
    t0 = time.perf_counter()
    for _ in range(steps):
        a = p.plan(state)
        state, _, _ = p.sim_step(state, a)
        trace.append(float(state[watch]))
        costs.append(float(p.iteration_record(s.opt_iterations - 1)["elite_costs"][0]))
    dt = (time.perf_counter() - t0) / steps
    p.close()
    assert np.isfinite(state).all()
    return dict(workload=name, population=s.num_simulated_trajectories, opt_iterations=s.opt_iterations, env_steps=steps,
                ms_per_closed_loop_step=1e3 * dt, watched_state_index=watch, start=trace[0], end=trace[-1],
                max=max(trace), every_10th=[round(v, 4) for v in trace[::10]],
                best_planned_cost_first_last=[costs[0], costs[-1]])


if __name__ == "__main__":
    out = []
    for name, watch in (("humanoid_standup_gt_n16384", 2), ("humanoid_standup_gt_n16384_rk4", 2),
                        ("halfcheetah_gt_n4096", 0), ("halfcheetah_gt_n4096_rk4", 0)):
        out.append(run(name, 100, watch))
        print(json.dumps(out[-1]), file=sys.stderr)
    print(json.dumps(out, indent=1))
