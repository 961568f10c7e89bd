set -x
mkdir -p gpurun_out
python scripts/kernel_bench.py > gpurun_out/kernels.json 2> gpurun_out/kernels.err; tail -3 gpurun_out/kernels.err; python -c "
import json; d=json.load(open('gpurun_out/kernels.json'))
for k in d['kernels']: print(k['workload'], k['kernel'], k['n'], round(k['ms'],4),'ms', round(k['achieved_gbs'],1),'GB/s', round(100*k['frac_of_measured_hbm'],2),'%')"
# ncu: sampler-only and select kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"rollout_kernel|select_kernel" -c 6 -o gpurun_out/prof_kernels python -c "
import sys; sys.path.insert(0,'.')
from icem_b200 import workloads
from icem_b200.planner import Planner
name='dense_tanh_humanoid_n16384'
w=workloads.get_workload(name); s=workloads.planner_settings(name, scale_population=16)
p=Planner(s); p.set_dense_model(*workloads.dense_model_weights(*w['dense'])); p.begin_rollout()
p.bench_op('sample', 262144, reps=1, flush_l2=False)
p.bench_op('select', 262144, reps=1, flush_l2=False)
" > gpurun_out/ncu_kernels.log 2>&1; tail -2 gpurun_out/ncu_kernels.log
