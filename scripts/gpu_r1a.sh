set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --workload dense_tanh_humanoid_n16384 --steps 20 --warmup 3 > gpurun_out/bench_dense_humanoid.json 2> gpurun_out/bench_dense_humanoid.err; tail -3 gpurun_out/bench_dense_humanoid.err; cat gpurun_out/bench_dense_humanoid.json
python bench.py --workload dense_tanh_cheetah_n4096 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dense_cheetah.json 2> gpurun_out/bench_dense_cheetah.err; cat gpurun_out/bench_dense_cheetah.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_dense_humanoid.csv python bench.py --workload dense_tanh_humanoid_n16384 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 6 -c 2 -o gpurun_out/prof_rollout_dense_humanoid python bench.py --workload dense_tanh_humanoid_n16384 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
