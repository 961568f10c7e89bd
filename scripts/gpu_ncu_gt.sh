mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rollout_kernel" -c 1 -f -o gpurun_out/prof_rollout_humanoid_gt_v6 python -c "
import sys; sys.path.insert(0,'.')
from icem_b200 import workloads
from icem_b200.planner import Planner
name='humanoid_standup_gt_n16384'
s=workloads.planner_settings(name)
p=Planner(s); p.begin_rollout()
p.plan(workloads.start_state(name))
" > gpurun_out/ncu_gt.log 2>&1; tail -2 gpurun_out/ncu_gt.log
