set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
bash scripts/gpu_main_dropin.sh 2>&1 | tail -12
python bench.py > gpurun_out/r2_bench_default_e.json 2> gpurun_out/r2_bench_default_e.err; cut -c1-300 gpurun_out/r2_bench_default_e.json; tail -2 gpurun_out/r2_bench_default_e.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference_e.json 2>/dev/null; cut -c1-400 gpurun_out/r2_bench_reference_e.json
