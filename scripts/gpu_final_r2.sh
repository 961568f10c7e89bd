# end-of-round evidence (round 2): tests, smoke, benches (all workloads + both integrators), reference arm, kernel
# timings, launch list, ncu --set full of the dominant kernels, sanitizer
set -x
TAG=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2_bench_humanoid_gt_$TAG.json 2>gpurun_out/bench.err; cut -c1-300 gpurun_out/r2_bench_humanoid_gt_$TAG.json
python bench.py --workload humanoid_standup_gt_n16384_rk4 --no-cpu-baseline > gpurun_out/r2_bench_humanoid_gt_rk4_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_humanoid_gt_rk4_$TAG.json
python bench.py --workload humanoid_standup_gt_n262144_rk4 --no-cpu-baseline --steps 5 > gpurun_out/r2_bench_humanoid_gt_n262144_rk4_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_humanoid_gt_n262144_rk4_$TAG.json
python bench.py --workload halfcheetah_gt_n4096 --no-cpu-baseline > gpurun_out/r2_bench_halfcheetah_gt_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_halfcheetah_gt_$TAG.json
python bench.py --workload halfcheetah_gt_n4096_rk4 --no-cpu-baseline > gpurun_out/r2_bench_halfcheetah_gt_rk4_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_halfcheetah_gt_rk4_$TAG.json
python bench.py --workload mlp_cheetah_n65536 --no-cpu-baseline > gpurun_out/r2_bench_mlp_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_mlp_$TAG.json
python bench.py --workload dense_tanh_humanoid_n16384 --no-cpu-baseline > gpurun_out/r2_bench_dense_tanh_humanoid_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_dense_tanh_humanoid_$TAG.json
ICEM_B200_ENGINE=warp python bench.py --no-cpu-baseline --no-strong --steps 5 > gpurun_out/r2_bench_humanoid_gt_warp_engine_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_humanoid_gt_warp_engine_$TAG.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference_$TAG.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_reference_$TAG.json
python scripts/kernel_bench.py > gpurun_out/r2_kernels_$TAG.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_humanoid_gt_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chain_rollout_kernel" -s 3 -c 1 -f -o gpurun_out/r2_prof_chain_humanoid_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-strong > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"chain_rollout_kernel" -s 5 -c 1 -f -o gpurun_out/r2_prof_chain_cheetah_$TAG python bench.py --workload halfcheetah_gt_n4096 --steps 1 --warmup 3 --no-cpu-baseline --no-strong > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"mlp_rollout_kernel" -c 1 -f -o gpurun_out/r2_prof_mlp_$TAG python bench.py --workload mlp_cheetah_n65536 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# summaries here (the box has ncu); only the headline kernel's report travels back (64 MiB limit on gpurun_out)
python scripts/ncu_summary.py gpurun_out/r2_prof_chain_humanoid_$TAG.ncu-rep > gpurun_out/r2_ncu_full_chain_rollout_humanoid_gt_n16384_$TAG.txt
python scripts/ncu_summary.py gpurun_out/r2_prof_chain_cheetah_$TAG.ncu-rep > gpurun_out/r2_ncu_full_chain_rollout_halfcheetah_gt_n4096_$TAG.txt
python scripts/ncu_summary.py gpurun_out/r2_prof_mlp_$TAG.ncu-rep > gpurun_out/r2_ncu_full_mlp_rollout_n65536_$TAG.txt
rm -f gpurun_out/r2_prof_chain_cheetah_$TAG.ncu-rep gpurun_out/r2_prof_mlp_$TAG.ncu-rep
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize.py cheetah humanoid sampler trainer reacher > gpurun_out/r2_sanitize_memcheck_$TAG.log 2>&1; grep -E "ok$|ERROR SUMMARY" gpurun_out/r2_sanitize_memcheck_$TAG.log
timeout 300 compute-sanitizer --tool synccheck --print-limit 5 python scripts/sanitize.py cheetah humanoid mlp sampler trainer reacher > gpurun_out/r2_sanitize_synccheck_$TAG.log 2>&1; grep -E "ok$|ERROR SUMMARY" gpurun_out/r2_sanitize_synccheck_$TAG.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python scripts/sanitize.py cheetah humanoid sampler trainer reacher > gpurun_out/r2_sanitize_racecheck_$TAG.log 2>&1; grep -E "ok$|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2_sanitize_racecheck_$TAG.log
ls -la gpurun_out | tail -12
