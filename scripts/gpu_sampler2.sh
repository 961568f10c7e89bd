set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mlp.py -m gpu -x -q 2>&1 | tail -8
./scripts/ubench/tanh_rate
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"colored_sampler" -c 2 -f -o gpurun_out/prof_series_sampler python -c "
import sys; sys.path.insert(0,'.')
from icem_b200 import workloads
from icem_b200.planner import Planner
name='dense_tanh_humanoid_n16384'
w=workloads.get_workload(name); s=workloads.planner_settings(name, scale_population=16)
p=Planner(s); p.set_dense_model(*workloads.dense_model_weights(*w['dense'])); p.begin_rollout()
p.bench_op('sample', 262144, reps=1, flush_l2=False)
" > gpurun_out/ncu_sampler.log 2>&1; tail -2 gpurun_out/ncu_sampler.log
