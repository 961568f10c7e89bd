"""dev check: device articulated step vs oracle (run on the GPU box)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from icem_b200.planner import Planner, PlannerSettings
from icem_b200.robots import get_model
from oracle.articulated_np import make_model
from oracle import costs_np
from oracle.icem_np import reduce_costs

for name, cost in [("halfcheetah", "halfcheetah"), ("humanoid_standup", "humanoid_standup")]:
    m = get_model(name)
    mod = make_model(name)
    lim = m.ctrl_limit
    p = Planner(PlannerSettings(horizon=30, num_simulated_trajectories=64, action_low=-lim * np.ones(m.nu),
                                action_high=lim * np.ones(m.nu), dynamics=name, cost=cost,
                                obs_dim=17 if name == "halfcheetah" else 47))
    rs = np.random.RandomState(0)
    st = np.concatenate([m.qpos0, np.zeros(m.nv)])
    errs = []
    for t in range(60):
        u = rs.uniform(-lim, lim, m.nu)
        ref = mod.step_state(st[None], u[None])[0]
        got, _, _ = p.sim_step(st, u)
        errs.append(np.abs(got - ref).max())
        st = ref
    print(name, "single-step max err over 60 closed-loop steps:", np.max(errs), "median", np.median(errs))
    # rollout costs
    n = 256
    acts = rs.uniform(-lim, lim, (n, 30, m.nu)).astype(np.float32)
    start = np.concatenate([m.qpos0, np.zeros(m.nv)])
    start[m.nq:] = 0.1 * rs.randn(m.nv)
    t0 = time.time()
    obs = mod.rollout(start.astype(np.float32).astype(np.float64), acts.astype(np.float64))
    t1 = time.time()
    if cost == "halfcheetah":
        c = costs_np.halfcheetah_cost(obs, acts.astype(np.float64), True)
    else:
        c = costs_np.humanoid_standup_cost(obs, acts.astype(np.float64))
    ref = reduce_costs(c, "sum")
    got = p.op_rollout_cost(start, acts)
    d = np.abs(got - ref)
    print(name, "rollout cost: max abs diff", d.max(), "median", np.median(d), "cost range", ref.min(), ref.max(),
          "oracle s", t1 - t0)
    k = 10
    print("  elite idx oracle", np.argsort(ref, kind="stable")[:k], "\n  elite idx device", np.argsort(got, kind="stable")[:k])
    p.close()
