mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])"
python scripts/batched_bench.py --env HumanoidStandup --fused --episodes 1 8 32 64 128 256 2>/dev/null | tee gpurun_out/batched_humanoid_fused.jsonl
python scripts/batched_bench.py --env HalfCheetah --fused --episodes 32 128 256 2>/dev/null | tee gpurun_out/batched_cheetah_fused.jsonl
python scripts/batched_bench.py --env HalfCheetah --episodes 32 128 2>/dev/null | tee gpurun_out/batched_cheetah_streams.jsonl
timeout 300 compute-sanitizer --tool synccheck --print-limit 5 python scripts/sanitize.py cheetah humanoid mlp ops > gpurun_out/sanitize_synccheck.log 2>&1; grep -E "ok$|ERROR SUMMARY" gpurun_out/sanitize_synccheck.log
