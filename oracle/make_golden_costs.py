"""tests/golden/costs_locomotion.npz: the reference's OWN Hopper / Ant cost functions
(/root/reference/icem/environments/mujoco.py:146-176, 196-231) evaluated on seeded inputs.  TEST INFRASTRUCTURE.

    python -m oracle.make_golden_costs        # this container only (needs /root/reference)

The env classes cannot be constructed here (gym / mujoco-py absent), so `cost_fn`, `unhealthy_states` and
`are_states_unhealthy` are called unbound on a stub carrying the gym-v3 default attributes they read."""
import os
import types

import numpy as np

from oracle import costs_np, ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def inputs():
    rs = np.random.RandomState(7)
    ho = rs.randn(60, 9, 12)
    ho[..., 1] += 1.0                       # heights around the 0.7 threshold
    ho[3, 2, 5] = 150.0                     # outside the +-100 state range
    ho[4, 1, 7] = np.nan
    ho[5, :, 2] = 0.5                       # angle outside +-0.2: must NOT matter (logical_and out= quirk)
    hn = ho + 0.01 * rs.randn(*ho.shape)
    ha = rs.uniform(-1, 1, (60, 9, 3))
    ao = rs.randn(40, 5, 113)
    ao[..., 2] = rs.uniform(0.0, 1.2, (40, 5))
    ao[0, 0, 2], ao[0, 1, 2] = 0.2, 1.0     # the closed interval ends are healthy
    ao[2, 1, 50] = np.inf
    an = ao + 0.01 * rs.randn(*ao.shape)
    aa = rs.uniform(-1, 1, (40, 5, 8))
    return (ho, ha, hn), (ao, aa, an)


def humanoid_inputs():
    rs = np.random.RandomState(8)
    o = rs.randn(30, 6, 378)
    o[..., 2] = rs.uniform(0.8, 2.2, (30, 6))
    o[0, 0, 2], o[0, 1, 2] = 1.0, 2.0          # the open interval ends are unhealthy
    o[1, 2, 300] = np.nan
    return o, rs.uniform(-0.4, 0.4, (30, 6, 17)), o + 0.01 * rs.randn(*o.shape)


def reference_humanoid_costs():
    ref_loader.load_reference()
    import environments.mujoco as rm
    hp = costs_np.HUMANOID
    st = types.SimpleNamespace(_exclude_current_positions_from_observation=hp["exclude_current_positions"],
                               _healthy_z_range=hp["healthy_z_range"], _forward_reward_weight=hp["forward_weight"],
                               _ctrl_cost_weight=hp["ctrl_weight"], model=types.SimpleNamespace(nq=hp["nq"]))
    st.unhealthy_states = lambda s: rm.Humanoid.unhealthy_states(st, s)
    o, a, n = humanoid_inputs()
    return rm.Humanoid.cost_fn(st, o.copy(), a, n)


def reference_costs():
    ref_loader.load_reference()
    import environments.mujoco as rm
    hp, ap = costs_np.HOPPER, costs_np.ANT
    hs = types.SimpleNamespace(dt=hp["dt"], _ctrl_cost_weight=hp["ctrl_weight"],
                               _healthy_state_range=hp["healthy_state_range"], _healthy_z_range=hp["healthy_z_range"],
                               _healthy_angle_range=hp["healthy_angle_range"])
    hs.unhealthy_states = lambda st: rm.Hopper.unhealthy_states(hs, st)
    as_ = types.SimpleNamespace(dt=ap["dt"], _ctrl_cost_weight=ap["ctrl_weight"], _healthy_z_range=ap["healthy_z_range"])
    as_.are_states_unhealthy = lambda st: rm.Ant.are_states_unhealthy(as_, st)
    (ho, ha, hn), (ao, aa, an) = inputs()
    return rm.Hopper.cost_fn(hs, ho.copy(), ha, hn), rm.Ant.cost_fn(as_, ao.copy(), aa, an)


def reacher_inputs():
    rs = np.random.RandomState(9)
    return rs.randn(50, 7, 11), rs.uniform(-1, 1, (50, 7, 2))


def reference_reacher_costs():
    """environments/mujoco.py:366-368, called unbound (the method reads nothing from `self`)."""
    ref_loader.load_reference()
    import environments.mujoco as rm
    o, a = reacher_inputs()
    return rm.Reacher.cost_fn(None, o.copy(), a, o)


GOAL_CASES = [  # (name, reference class, goal_idx, achieved_idx, sparse, threshold, shaped)
    ("fetch_pick_dense_shaped", "FetchPickAndPlace", list(range(25, 28)), [3, 4, 5], False, 0.05, True),
    ("fetch_pick_dense", "FetchPickAndPlace", list(range(25, 28)), [3, 4, 5], False, 0.05, False),
    ("fetch_pick_sparse_shaped", "FetchPickAndPlace", list(range(25, 28)), [3, 4, 5], True, 2.2, True),
    ("fetch_reach_dense", "FetchReach", list(range(10, 13)), [0, 1, 2], False, 0.05, False),
    ("fetch_reach_sparse", "FetchReach", list(range(10, 13)), [0, 1, 2], True, 2.2, False),
]


def goal_inputs(width):
    return np.random.RandomState(10 + width).randn(40, 6, width)


def reference_goal_costs():
    """environments/robotics.py:150-164 (FetchPickAndPlace) and abstract_environments.py:115-123 (FetchReach inherits
    MaskedGoalSpaceEnvironmentInterface.cost_fn), called unbound on a stub carrying the attributes they read."""
    ref_loader.load_reference()
    import environments.robotics as rr
    out = {}
    for name, cls_name, gi, ai, sparse, thr, shaped in GOAL_CASES:
        cls = getattr(rr, cls_name)
        st = types.SimpleNamespace(goal_idx=np.asarray(gi), achieved_goal_idx=ai, sparse=sparse, threshold=thr,
                                   shaped_reward=shaped)
        st.goal_from_observation = lambda o, st=st, cls=cls: cls.goal_from_observation(st, o)
        st.achieved_goal_from_observation = lambda o, st=st, cls=cls: cls.achieved_goal_from_observation(st, o)
        o = goal_inputs(gi[-1] + 1)
        out[name] = np.asarray(cls.cost_fn(st, o.reshape(-1, o.shape[-1]).copy(), None, None)).reshape(o.shape[:-1])
    return out


def main():
    np.savez_compressed(os.path.join(OUT, "costs_goal_distance.npz"), **reference_goal_costs())
    h, a = reference_costs()
    hu = reference_humanoid_costs()
    np.savez_compressed(os.path.join(OUT, "costs_reacher.npz"), reacher=reference_reacher_costs())
    np.savez_compressed(os.path.join(OUT, "costs_locomotion.npz"), hopper=h, ant=a, humanoid=hu)
    print("humanoid", hu.shape, float(np.mean(hu > 50)))
    print("hopper", h.shape, float(np.nanmean(h > 100)), "ant", a.shape, float(np.mean(a > 50)))


if __name__ == "__main__":
    main()
