set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_articulated.py tests/test_gpu_fullsize.py tests/test_gpu_locomotion.py -x -q 2>&1 | tail -25
python bench.py --no-cpu-baseline > gpurun_out/r2_bench_humanoid_chain_a.json 2> gpurun_out/r2_bench.err; cut -c1-300 gpurun_out/r2_bench_humanoid_chain_a.json; tail -3 gpurun_out/r2_bench.err
python bench.py --workload halfcheetah_gt_n4096 --no-cpu-baseline > gpurun_out/r2_bench_cheetah_chain_a.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_cheetah_chain_a.json
ICEM_B200_ENGINE=warp python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2_bench_humanoid_warp_a.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_humanoid_warp_a.json
