#!/usr/bin/env python
"""Throughput of B independent episodes planned side by side at the reference's DEFAULT budget (40 -> 32 -> 25
trajectories per step, settings/defaults/i-cem-blitz.json) -- SURVEY 8f-4.

    python scripts/batched_bench.py [--env HumanoidStandup] [--episodes 1 8 32 64 128] [--steps 5]
"""
import argparse
import contextlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icem_b200.batched import make_episode_batch, make_fused_episode_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="HumanoidStandup")
    ap.add_argument("--episodes", type=int, nargs="+", default=[1, 8, 32, 64, 128])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--fused", action="store_true", help="one num_problems=B handle instead of B handles/streams")
    args = ap.parse_args()
    beta = 2.0 if args.env == "HumanoidStandup" else 0.25
    params = dict(horizon=30, num_simulated_trajectories=40, factor_decrease_num=1.25, cost_along_trajectory="sum",
                  action_sampler_params=dict(alpha=0.1, elites_size=10, fraction_elites_reused=0.3, init_std=0.5,
                                             keep_previous_elites=True, shift_elites_over_time=True,
                                             use_mean_actions=True, opt_iterations=3, noise_beta=beta))
    traj_per_step = 40 + 32 + 25 + 3
    out = []
    for B in args.episodes:
        with contextlib.redirect_stdout(sys.stderr):
            batch = (make_fused_episode_batch if args.fused else make_episode_batch)(args.env, B, params, seed=1)
            batch.reset()
            for _ in range(2):
                batch.step()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                batch.plan()
            dt = (time.perf_counter() - t0) / args.steps
            batch.close()
        out.append(dict(env=args.env, episodes=B, mode="fused (one handle, grid.y = episode)" if args.fused
                        else "streams (one handle per episode)", ms_per_plan_step_all_episodes=1e3 * dt,
                        plan_steps_per_s=B / dt, trajectories_per_s=B * traj_per_step / dt))
        print(json.dumps(out[-1]))


if __name__ == "__main__":
    main()
