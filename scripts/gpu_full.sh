# full validation: GPU tests, smoke, default bench, reference arm, launch list
set -x
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err; tail -2 gpurun_out/bench_default_$TAG.err; cat gpurun_out/bench_default_$TAG.json
python bench.py --workload halfcheetah_gt_n4096 --no-cpu-baseline > gpurun_out/bench_cheetah_$TAG.json 2>/dev/null; cat gpurun_out/bench_cheetah_$TAG.json
