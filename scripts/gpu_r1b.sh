set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_humanoid_gt.json 2> gpurun_out/bench_humanoid_gt.err; tail -3 gpurun_out/bench_humanoid_gt.err; cat gpurun_out/bench_humanoid_gt.json
python bench.py --workload halfcheetah_gt_n4096 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cheetah_gt.json 2> gpurun_out/bench_cheetah_gt.err; tail -3 gpurun_out/bench_cheetah_gt.err; cat gpurun_out/bench_cheetah_gt.json
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_humanoid_gt.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 6 -c 1 -o gpurun_out/prof_rollout_humanoid_gt python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
