set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp.py -x -q 2>&1 | tail -3 | sed "s/^/T: /"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_rollout -s 3 -c 1 -o gpurun_out/prof_mlp2 python bench.py --workload mlp_cheetah_n65536 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mlp.log 2>&1; tail -1 gpurun_out/ncu_mlp.log | cut -c1-200
