# new sampler: parity tests + kernel timings
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_mlp.py -m gpu -x -q 2>&1 | tail -15
python scripts/kernel_bench.py > gpurun_out/kernels.json 2> gpurun_out/kernels.err; tail -3 gpurun_out/kernels.err; python -c "
import json; d=json.load(open('gpurun_out/kernels.json'))
for k in d['kernels']: print(k['workload'], k['kernel'], k['n'], round(k['ms'],4),'ms', round(k['achieved_gbs'],1),'GB/s', round(100*k['frac_of_measured_hbm'],2),'%')"
python bench.py --workload mlp_cheetah_n65536 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_mlp.json | cut -c1-400
