"""Regression fixtures of the articulated-body oracle: tests/golden/articulated_<name>.npz.  TEST INFRASTRUCTURE.

    python -m oracle.make_golden_articulated

These are SELF-pins (seeded rollouts of oracle/articulated_np.py on icem_b200/robots.py tables): MuJoCo, the
reference's actual ground truth, is unavailable, so there is no external vector to pin against."""
import os

import numpy as np

from oracle.articulated_np import make_model

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def case(name, n=6, h=12, seed=5):
    mod = make_model(name)
    m = mod.m
    rs = np.random.RandomState(seed)
    start = np.concatenate([m.qpos0, 0.1 * rs.randn(m.nv)])
    actions = rs.uniform(-m.ctrl_limit, m.ctrl_limit, (n, h, m.nu))
    return mod, start, actions


def main():
    for name in ("halfcheetah", "humanoid_standup", "hopper", "ant"):
        mod, start, actions = case(name)
        obs = mod.rollout(start, actions)
        np.savez_compressed(os.path.join(OUT, f"articulated_{name}.npz"), start=start, actions=actions,
                            observations=obs)
        print(name, obs.shape, float(np.abs(obs).max()))


if __name__ == "__main__":
    main()
