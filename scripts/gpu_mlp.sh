set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp.py -x -q 2>&1 | tail -3 | sed "s/^/T: /"
python bench.py --workload mlp_cheetah_n65536 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_mlp.json; python -c "
import json; d=json.load(open('gpurun_out/bench_mlp.json')); print('T:', round(d['value']/1e6,1),'M traj/s', round(d['ms_per_step'],3),'ms/step; mlp kernel', round(d['roofline']['kernel_ms_avg'],4),'ms', round(d['roofline']['achieved'],1),'TFLOP/s', round(100*d['roofline']['frac'],1),'%')"
