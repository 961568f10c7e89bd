"""Per-step cost functions of the two in-scope environments, restated in NumPy float64.

TEST INFRASTRUCTURE ONLY.  Citations relative to /root/reference/icem/.
Signature: cost(observations[..., obs_dim], actions[..., d]) -> [...]; `next_obs` is ignored by
both reference functions, so it is not an argument here.
"""
import math

import numpy as np


def halfcheetah_cost(obs, act, penalise_flipping=True):
    """environments/mujoco.py:67-99."""
    if obs.shape[-1] == 18:
        root_angle, velocity = obs[..., 2], obs[..., 9]
    elif obs.shape[-1] == 17:
        root_angle, velocity = obs[..., 1], obs[..., 8]
    else:
        raise ValueError(f"Got state of dimension {obs.shape[-1]}. Possible dimensions are 17 or 18.")
    scores = np.zeros(act.shape[:-1])
    if penalise_flipping:
        scores += (root_angle > math.pi / 2) * 10
        scores += (root_angle < -math.pi / 2) * 10
    scores += 0.1 * np.sum(act ** 2, axis=-1)
    scores -= velocity
    return scores


def humanoid_standup_cost(obs, act):
    """environments/mujoco.py:259-277 (obs[2] = root z because _get_obs keeps the full qpos, :241-252)."""
    return -obs[..., 2] + 0.1 * np.square(act).sum(axis=-1)
