mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_mlp.py -m gpu -x -q 2>&1 | tail -4
timeout 120 python bench.py --workload mlp_cheetah_n65536 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_mlp.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['kernel_ms_avg'])"
ICEM_B200_LIB=icem_b200/lib/libicem_b200_trace.so timeout 120 python scripts/mlp_trace.py > gpurun_out/trace2.txt 2>&1; cat gpurun_out/trace2.txt
