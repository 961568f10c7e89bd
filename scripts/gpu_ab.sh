python -m pytest tests/test_gpu_parity.py tests/test_gpu_cem_std.py tests/test_gpu_random.py tests/test_gpu_batched.py tests/test_gpu_multi.py tests/test_gpu_mlp.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys; sys.path.insert(0,'.')
from icem_b200 import workloads
from icem_b200.planner import Planner
name='dense_tanh_humanoid_n16384'
w=workloads.get_workload(name); s=workloads.planner_settings(name, scale_population=16)
p=Planner(s); p.set_dense_model(*workloads.dense_model_weights(*w['dense'])); p.begin_rollout(); p.plan(workloads.start_state(name))
for n in (16384, 65536, 262144):
    print('select n', n, 'ms', round(p.bench_op('select', n, reps=20), 4))
PY
python bench.py --workload mlp_cheetah_n65536 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('mlp', d['value'], d['ms_per_step'])"
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'])"
