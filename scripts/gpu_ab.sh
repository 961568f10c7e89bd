python -m pytest tests/test_gpu_batched.py tests/test_gpu_random.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_avg'])"; done
