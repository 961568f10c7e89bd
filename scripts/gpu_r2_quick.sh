set -x
TAG=${1:-b}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_articulated.py tests/test_gpu_fullsize.py tests/test_gpu_locomotion.py tests/test_gpu_batched.py -x -q 2>&1 | tail -5
python bench.py --no-cpu-baseline > gpurun_out/r2_bench_humanoid_chain_$TAG.json 2> gpurun_out/r2_bench.err; cut -c1-200 gpurun_out/r2_bench_humanoid_chain_$TAG.json; tail -3 gpurun_out/r2_bench.err
python bench.py --workload halfcheetah_gt_n4096 --no-cpu-baseline > gpurun_out/r2_bench_cheetah_chain_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_cheetah_chain_$TAG.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chain_rollout_kernel" -s 3 -c 1 -f -o gpurun_out/r2_prof_chain_humanoid_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
