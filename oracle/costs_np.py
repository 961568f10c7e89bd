"""Per-step cost functions of the two in-scope environments, restated in NumPy float64.

TEST INFRASTRUCTURE ONLY.  Citations relative to /root/reference/icem/.
Signature: cost(observations[..., obs_dim], actions[..., d]) -> [...]; `next_obs` is ignored by
the HalfCheetah / HumanoidStandup functions, so it is not an argument there; the Hopper / Ant costs
 read it and take it explicitly.
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


HOPPER = dict(dt=0.008, ctrl_weight=1e-3, healthy_state_range=(-100.0, 100.0), healthy_z_range=(0.7, float("inf")),
              healthy_angle_range=(-0.2, 0.2))       # gym Hopper-v3 defaults (frame_skip 4 x timestep 0.002)
ANT = dict(dt=0.05, ctrl_weight=0.5, healthy_z_range=(0.2, 1.0))       # gym Ant-v3 defaults


def hopper_cost(obs, act, next_obs, p=HOPPER):
    """environments/mujoco.py:196-231.  Line 208 is `np.logical_and(healthy_state, healthy_z, healthy_angle)`:
    the third positional argument of a ufunc is `out`, so the angle test is overwritten, not combined."""
    z, state = obs[..., 1], obs[..., 2:]
    lo, hi = p["healthy_state_range"]
    healthy_state = np.all(np.logical_and(lo < state, state < hi), axis=-1)
    healthy_z = (p["healthy_z_range"][0] < z) * (z < p["healthy_z_range"][1])
    is_healthy = np.logical_and(healthy_state, healthy_z)
    unhealthy = 1 - np.isfinite(obs).all(axis=-1) * is_healthy
    x_velocity = (next_obs[..., 0] - obs[..., 0]) / p["dt"]
    return -x_velocity + 200 * unhealthy + p["ctrl_weight"] * np.sum(np.square(act), axis=-1)


def ant_cost(obs, act, next_obs, p=ANT):
    """environments/mujoco.py:146-176."""
    lo, hi = p["healthy_z_range"]
    unhealthy = 1 - np.isfinite(obs).all(axis=-1) * (lo <= obs[..., 2]) * (obs[..., 2] <= hi)
    x_velocity = (next_obs[..., 0] - obs[..., 0]) / p["dt"]
    return -x_velocity + 100 * unhealthy + p["ctrl_weight"] * np.sum(np.square(act), axis=-1)


HUMANOID = dict(nq=24, exclude_current_positions=False, forward_weight=1.25, ctrl_weight=0.1,
                healthy_z_range=(1.0, 2.0))       # gym Humanoid-v3 defaults


def humanoid_cost(obs, act, next_obs=None, p=HUMANOID):
    """environments/mujoco.py:301-343: the velocity is READ from the observation (index nq with positions kept, nq - 2
    without), the height test is strict, the penalty is the literal 100."""
    z = obs[..., 0] if p["exclude_current_positions"] else obs[..., 2]
    lo, hi = p["healthy_z_range"]
    unhealthy = 1 - np.isfinite(obs).all(axis=-1) * ((lo < z) * (z < hi))
    x_velocity = obs[..., p["nq"] - 2] if p["exclude_current_positions"] else obs[..., p["nq"]]
    return -p["forward_weight"] * x_velocity + 100 * unhealthy + p["ctrl_weight"] * np.sum(np.square(act), axis=-1)


REACHER = dict(l1=0.1, l2=0.11, target0=(0.0, 0.0))      # gym reacher.xml link lengths; target world position at q2 = q3 = 0


def reacher_cost(obs, act=None, next_obs=None):
    """environments/mujoco.py:366-368: |fingertip - target| = norm of the last three observation entries."""
    return np.linalg.norm(np.asarray(obs)[..., -3:], axis=-1)


def reacher_observation(state, p=REACHER):
    """gym reacher.py::_get_obs from the state (qpos(4) ++ qvel(4)): [cos q0, cos q1, sin q0, sin q1, target x, y,
    qvel0, qvel1, fingertip - target (3)] with the fingertip from the planar forward kinematics."""
    st = np.asarray(state, np.float64)
    q0, q1, tx, ty = st[..., 0], st[..., 1], st[..., 2], st[..., 3]
    fx = p["l1"] * np.cos(q0) + p["l2"] * np.cos(q0 + q1)
    fy = p["l1"] * np.sin(q0) + p["l2"] * np.sin(q0 + q1)
    return np.stack([np.cos(q0), np.cos(q1), np.sin(q0), np.sin(q1), tx, ty, st[..., 4], st[..., 5],
                     fx - (p["target0"][0] + tx), fy - (p["target0"][1] + ty), np.zeros_like(q0)], axis=-1)


def goal_distance_cost(obs, goal_idx, achieved_idx, sparse=False, threshold=0.05, shaped=False):
    """environments/abstract_environments.py:115-123 (MaskedGoalSpaceEnvironmentInterface.cost_fn: FetchReach) and,
    with `shaped`, environments/robotics.py:150-164 (FetchPickAndPlace.cost_fn, the sparse branch of which always adds
    the end-effector term -- dist_end_eff_to_box is 0 there unless shaped_reward)."""
    obs = np.asarray(obs)
    dist = np.linalg.norm(np.take(obs, goal_idx, axis=-1) - np.take(obs, achieved_idx, axis=-1), axis=-1)
    eff = np.linalg.norm(obs[..., :3] - obs[..., 3:6], axis=-1) if shaped else 0
    if sparse:
        return np.asarray(dist > threshold, dtype=np.float32) + np.asarray(eff > threshold, dtype=np.float32) * 0.1
    return dist + eff * 0.1
