set -x
mkdir -p gpurun_out
python bench.py --workload mlp_cheetah_n65536 --steps 20 --warmup 3 > gpurun_out/bench_mlp.json 2> gpurun_out/bench_mlp.err; tail -3 gpurun_out/bench_mlp.err; cat gpurun_out/bench_mlp.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_rollout -s 3 -c 1 -o gpurun_out/prof_mlp python bench.py --workload mlp_cheetah_n65536 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mlp.log 2>&1; tail -2 gpurun_out/ncu_mlp.log | cut -c1-200
