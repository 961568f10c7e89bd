python -m pytest tests/test_gpu_articulated.py tests/test_gpu_batched.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_avg'])"; done
timeout 200 compute-sanitizer --tool synccheck --print-limit 5 python scripts/sanitize.py cheetah humanoid > gpurun_out/sanitize_synccheck.log 2>&1; grep -E "ok$|ERROR SUMMARY" gpurun_out/sanitize_synccheck.log
