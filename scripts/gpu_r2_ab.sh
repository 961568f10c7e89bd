# quick A/B of the two ground-truth lines after an engine change (tests of the articulated paths first)
python -m pytest tests/test_gpu_articulated.py tests/test_gpu_locomotion.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
for W in humanoid_standup_gt_n16384 halfcheetah_gt_n4096 humanoid_standup_gt_n16384_rk4; do python bench.py --workload $W --no-cpu-baseline --no-strong 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms_avg'])"; done
