# end-of-round evidence: tests, smoke, benches (all workloads), reference arm, launch list, ncu --set full of the dominant kernels
set -x
TAG=${1:-v9}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_humanoid_gt_$TAG.json 2>gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_humanoid_gt_$TAG.json
python bench.py --workload halfcheetah_gt_n4096 --no-cpu-baseline > gpurun_out/bench_halfcheetah_gt_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_halfcheetah_gt_$TAG.json
python bench.py --workload mlp_cheetah_n65536 > gpurun_out/bench_mlp_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_mlp_$TAG.json
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_reference_$TAG.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference_$TAG.json
python scripts/kernel_bench.py > gpurun_out/kernels_$TAG.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_humanoid_gt_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rollout_kernel" -s 3 -c 1 -f -o gpurun_out/prof_rollout_humanoid_gt_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"mlp_rollout_kernel" -c 1 -f -o gpurun_out/prof_mlp_$TAG python bench.py --workload mlp_cheetah_n65536 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"colored_sampler" -c 1 -f -o gpurun_out/prof_sampler_$TAG python -c "
import sys; sys.path.insert(0,'.')
from icem_b200 import workloads
from icem_b200.planner import Planner
name='dense_tanh_humanoid_n16384'
w=workloads.get_workload(name); s=workloads.planner_settings(name, scale_population=16)
p=Planner(s); p.set_dense_model(*workloads.dense_model_weights(*w['dense'])); p.begin_rollout()
p.bench_op('sample', 262144, reps=1, flush_l2=False)" > /dev/null 2>&1
timeout 150 compute-sanitizer --tool synccheck --print-limit 5 python scripts/sanitize.py cheetah humanoid mlp > gpurun_out/sanitize_synccheck_$TAG.log 2>&1; grep -E "ok$|ERROR SUMMARY" gpurun_out/sanitize_synccheck_$TAG.log
ls -la gpurun_out | tail -12
