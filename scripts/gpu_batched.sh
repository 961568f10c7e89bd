mkdir -p gpurun_out
python -m pytest tests/test_gpu_batched.py tests/test_gpu_controller.py -m gpu -x -q 2>&1 | tail -6
python scripts/batched_bench.py --env HumanoidStandup 2>/dev/null | tee gpurun_out/batched_humanoid.jsonl
python scripts/batched_bench.py --env HalfCheetah 2>/dev/null | tee gpurun_out/batched_cheetah.jsonl
