# stand-alone sampler iteration: parity (injected draws), production-noise statistics, per-kernel timings, ncu of the sampler
set -x
TAG=${1:-s}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_mlp.py -x -q 2>&1 | tail -5
python scripts/kernel_bench.py > gpurun_out/r2_kernels_$TAG.json 2>gpurun_out/kb.err; grep -A8 '"sample"' gpurun_out/r2_kernels_$TAG.json | grep -E "workload|ms|frac"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"colored_sampler_kernel" -c 1 -f -o gpurun_out/r2_prof_sampler_$TAG python scripts/kernel_bench.py > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r2_prof_sampler_$TAG.ncu-rep > gpurun_out/r2_ncu_full_series_sampler_n262144_$TAG.txt; head -30 gpurun_out/r2_ncu_full_series_sampler_n262144_$TAG.txt
