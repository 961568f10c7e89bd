# usage: bash scripts/gpu_perf.sh <tag>   -- articulated parity tests, GT benches, one ncu full capture of the fused kernel
set -x
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_articulated.py tests/test_gpu_multi.py -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_humanoid_gt_$TAG.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log; cat gpurun_out/bench_humanoid_gt_$TAG.json
python bench.py --workload halfcheetah_gt_n4096 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cheetah_gt_$TAG.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_cheetah_gt_$TAG.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 6 -c 1 -o gpurun_out/prof_rollout_humanoid_gt_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
