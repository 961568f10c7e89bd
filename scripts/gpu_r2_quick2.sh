set -x
TAG=${1:-d}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_articulated.py tests/test_gpu_fullsize.py tests/test_gpu_locomotion.py tests/test_gpu_batched.py -x -q 2>&1 | tail -5
python bench.py --no-cpu-baseline > gpurun_out/r2_bench_humanoid_chain_$TAG.json 2> gpurun_out/r2_bench.err; cut -c1-200 gpurun_out/r2_bench_humanoid_chain_$TAG.json; tail -3 gpurun_out/r2_bench.err
python bench.py --workload humanoid_standup_gt_n16384_rk4 --no-cpu-baseline --steps 10 > gpurun_out/r2_bench_humanoid_rk4_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_humanoid_rk4_$TAG.json
python bench.py --workload halfcheetah_gt_n4096 --no-cpu-baseline > gpurun_out/r2_bench_cheetah_chain_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_cheetah_chain_$TAG.json
python bench.py --workload halfcheetah_gt_n4096_rk4 --no-cpu-baseline --steps 10 > gpurun_out/r2_bench_cheetah_rk4_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_cheetah_rk4_$TAG.json
bash scripts/gpu_r2_sweep.sh 2>/dev/null | grep "^warps"
