# planar instantiation of the chain engine: parity tests, then A/B timings (ICEM_B200_PLANAR=0 = spatial instantiation)
set -x
TAG=${1:-p}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_articulated.py tests/test_gpu_locomotion.py tests/test_gpu_fullsize.py tests/test_gpu_batched.py tests/test_gpu_cem_std.py tests/test_gpu_controller.py -x -q 2>&1 | tail -4
for V in planar spatial; do
  if [ $V = spatial ]; then export ICEM_B200_PLANAR=0; fi
  python bench.py --workload halfcheetah_gt_n4096 --no-cpu-baseline > gpurun_out/r2_bench_halfcheetah_gt_${V}_$TAG.json 2>/dev/null
  python bench.py --workload halfcheetah_gt_n4096_rk4 --no-cpu-baseline > gpurun_out/r2_bench_halfcheetah_gt_rk4_${V}_$TAG.json 2>/dev/null
  python -c "
import json
for w in ('', '_rk4'):
    d=json.load(open('gpurun_out/r2_bench_halfcheetah_gt%s_${V}_$TAG.json' % w)); print('$V', w, round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms_avg'])"
done
unset ICEM_B200_PLANAR
python scripts/kernel_bench.py > gpurun_out/r2_kernels_$TAG.json 2>/dev/null; grep -B2 -A6 '"halfcheetah_gt_n4096"' gpurun_out/r2_kernels_$TAG.json | grep -E "kernel|ms\"|traj"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"chain_rollout_kernel" -s 5 -c 1 -f -o gpurun_out/r2_prof_chain_cheetah_planar_$TAG python bench.py --workload halfcheetah_gt_n4096 --steps 1 --warmup 3 --no-cpu-baseline --no-strong > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r2_prof_chain_cheetah_planar_$TAG.ncu-rep > gpurun_out/r2_ncu_full_chain_rollout_halfcheetah_gt_n4096_planar_$TAG.txt; head -16 gpurun_out/r2_ncu_full_chain_rollout_halfcheetah_gt_n4096_planar_$TAG.txt
