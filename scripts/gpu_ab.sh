python -m pytest tests/test_gpu_parity.py tests/test_gpu_cem_std.py tests/test_gpu_random.py tests/test_gpu_batched.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
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
python - <<'PY'
import sys, time; sys.path.insert(0,'.')
import contextlib
from icem_b200.batched import make_fused_episode_batch
params = dict(horizon=30, num_simulated_trajectories=40, factor_decrease_num=1.25, cost_along_trajectory="sum",
              action_sampler_params=dict(alpha=0.1, elites_size=10, fraction_elites_reused=0.3, init_std=0.5,
                                         keep_previous_elites=True, shift_elites_over_time=True,
                                         use_mean_actions=True, opt_iterations=3, noise_beta=2.0))
with contextlib.redirect_stdout(sys.stderr):
    b = make_fused_episode_batch("HumanoidStandup", 256, params, seed=1); b.reset(); b.step()
t0 = time.perf_counter()
for _ in range(5): b.step()
dt = (time.perf_counter() - t0) / 5
print('fused 256 episodes: full step (plan + env transitions) ms', round(1e3*dt, 2), 'env steps/s', round(256/dt))
PY
