set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"mlp_rollout" -c 2 -f -o gpurun_out/prof_mlp3 python -c "
import sys; sys.path.insert(0,'.')
from icem_b200 import workloads
from icem_b200.planner import Planner
name='mlp_cheetah_n65536'
w=workloads.get_workload(name); s=workloads.planner_settings(name)
p=Planner(s); p.set_mlp_model(*workloads.mlp_model_weights(*w['mlp'])); p.begin_rollout()
p.bench_op('rollout', 65536, reps=1, flush_l2=False)
" > gpurun_out/ncu_mlp.log 2>&1; tail -2 gpurun_out/ncu_mlp.log
