"""One-step state error of the device engines against the float64 oracle over contact-rich states (diagnostic)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icem_b200 import robots, workloads
from icem_b200.planner import Planner, PlannerSettings
from oracle.articulated_np import make_model
name = sys.argv[1] if len(sys.argv) > 1 else "humanoid_standup"
m = robots.get_model(name)
lim = m.ctrl_limit
s = PlannerSettings(horizon=30, num_simulated_trajectories=64, action_low=-lim * np.ones(m.nu, np.float32),
                    action_high=lim * np.ones(m.nu, np.float32), dynamics=name,
                    cost="halfcheetah" if name == "halfcheetah" else "humanoid_standup",
                    obs_dim=17 if name == "halfcheetah" else 47)
p = Planner(s)
mod = make_model(name, obs_skip=0)
rs = np.random.RandomState(0)
# states along oracle rollouts with random actions
n, h = 256, 24
st = np.broadcast_to(np.concatenate([m.qpos0, 0.1 * rs.randn(m.nv)]), (n, m.nq + m.nv)).copy()
states, actions = [], []
for t in range(h):
    a = rs.uniform(-lim, lim, (n, m.nu))
    if t % 4 == 3:
        states.append(st.astype(np.float32).astype(np.float64)); actions.append(a.astype(np.float32).astype(np.float64))
    st = mod.step_state(st, a)
S = np.concatenate(states); A = np.concatenate(actions)
ref = mod.step_state(S, A)
got = p.sim_step_batch(S, A)
e = np.abs(got - ref)
per = e.max(axis=1)
print(os.environ.get("ICEM_B200_ENGINE", "chain"), name, "states", len(S), "one-step err: median %.2e p90 %.2e p99 %.2e max %.2e" % (np.median(per), np.percentile(per, 90), np.percentile(per, 99), per.max()))
np.set_printoptions(precision=1, linewidth=220)
print("  per-dof p99 qd err", np.percentile(e[:, m.nq:], 99, axis=0))
