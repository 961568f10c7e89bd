"""Cost error of the device engines against the float64 oracle on rows of a full-size population (diagnostic)."""
import dataclasses, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icem_b200 import workloads
from icem_b200.planner import Planner
from oracle import costs_np
from oracle.articulated_np import make_model
from oracle.icem_np import reduce_costs
for name, robot in (("halfcheetah_gt_n4096", "halfcheetah"), ("humanoid_standup_gt_n16384", "humanoid_standup")):
    s = dataclasses.replace(workloads.planner_settings(name, seed=5), keep_iteration_actions=True)
    p = Planner(s)
    start = workloads.start_state(name, seed=2)
    p.begin_rollout(); p.plan(start)
    for it in (0, s.opt_iterations - 1):
        n = p.population_size(it, first_step=True)[1]
        acts = p.actions(it, n); costs = p.costs(it, n)
        order = np.argsort(costs, kind="stable")
        rows = np.unique(np.concatenate([order[:32], np.random.RandomState(0).choice(n, 256, replace=False)]))
        mod = make_model(robot)
        obs = mod.rollout(np.asarray(start, np.float32).astype(np.float64), acts[rows].astype(np.float64))
        per = costs_np.halfcheetah_cost(obs, acts[rows].astype(np.float64), True) if robot == "halfcheetah" else costs_np.humanoid_standup_cost(obs, acts[rows].astype(np.float64))
        ref = reduce_costs(per, "sum")
        d = np.abs(costs[rows] - ref)
        print(os.environ.get("ICEM_B200_ENGINE", "chain"), name, "iter", it, "median %.2e p90 %.2e p99 %.2e max %.2e" % (np.median(d), np.percentile(d, 90), np.percentile(d, 99), d.max()), "cost range", ref.min(), ref.max())
    p.close()
