import dataclasses, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icem_b200 import workloads
from icem_b200.planner import Planner
name = "humanoid_standup_gt_n16384"
s = dataclasses.replace(workloads.planner_settings(name, seed=5), keep_iteration_actions=True)
p = Planner(s)
start = workloads.start_state(name, seed=2)
p.begin_rollout(); p.plan(start)
it = s.opt_iterations - 1
n = p.population_size(it, first_step=True)[1]
acts = p.actions(it, n); costs = p.costs(it, n)
order = np.argsort(costs, kind="stable")
rows = np.unique(np.concatenate([order[:32], np.random.RandomState(0).choice(n, 256, replace=False)]))
rc = p.op_rollout_cost(start, acts[rows])
np.savez(os.path.join("gpurun_out", "fullsize_dump_" + os.environ.get("ICEM_B200_ENGINE", "chain") + ".npz"), start=start, rows=rows, acts=acts[rows], costs=costs[rows], rollout_only=rc)
print("dumped", len(rows))
