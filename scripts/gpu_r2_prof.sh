set -x
TAG=${1:-a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chain_rollout_kernel" -s 3 -c 1 -f -o gpurun_out/r2_prof_chain_humanoid_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chain_rollout_kernel" -s 5 -c 1 -f -o gpurun_out/r2_prof_chain_cheetah_$TAG python bench.py --workload halfcheetah_gt_n4096 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -5
