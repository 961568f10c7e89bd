# MLP kernel A/B: packed f16x2 tanh (variant library) vs fp32 tanh; ubench of the tanh variants
set -x
TAG=${1:-m}
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tanh_rate scripts/ubench/tanh_rate.cu && /tmp/tanh_rate | tee gpurun_out/r2_ubench_tanh_rate.txt
for V in base tanh16; do
  if [ $V = tanh16 ]; then export ICEM_B200_LIB=$PWD/icem_b200/lib/libicem_b200_tanh16.so; fi
  python -m pytest tests/test_gpu_mlp.py -x -q 2>&1 | tail -3
  python bench.py --workload mlp_cheetah_n65536 --no-cpu-baseline > gpurun_out/r2_bench_mlp_${V}_$TAG.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_mlp_${V}_$TAG.json'));print('$V', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['roofline']['frac'])"
  python scripts/mlp_precision_report.py --n 65536 --steps 10 > gpurun_out/r2_mlp_precision_${V}_$TAG.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r2_mlp_precision_${V}_$TAG.json'));print('$V', {k:d[k] for k in ('cost_err_median','elite_overlap','executed_action_abs_diff') if k in d})"
done
