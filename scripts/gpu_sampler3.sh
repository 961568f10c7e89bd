set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mlp.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
python - <<'PY'
import sys; sys.path.insert(0,'.')
from icem_b200 import workloads
from icem_b200.planner import Planner
for name in ('dense_tanh_humanoid_n16384','dense_tanh_cheetah_n4096'):
    w=workloads.get_workload(name); s=workloads.planner_settings(name, scale_population=16 if 'humanoid' in name else 64)
    p=Planner(s); p.set_dense_model(*workloads.dense_model_weights(*w['dense'])); p.begin_rollout()
    ms=p.bench_op('sample', 262144, reps=20)
    b=4*30*w['act_dim']*262144
    print(name, 'sample ms', round(ms,4), 'GB/s', round(b/ms/1e6,1))
    p.close()
PY
