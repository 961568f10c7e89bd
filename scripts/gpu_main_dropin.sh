#!/bin/bash
# T7 on the GPU box: the reference's UNCHANGED icem/main.py (staged copy, scripts/stage_reference.sh) drives the
# "mpc-icem-b200" controller + "CudaGroundTruthModel" for 20 env steps of examples/halfcheetah_icem_b200.json.
# Log -> gpurun_out/r2_main_dropin.log (copied to profiles/ afterwards).
set -x
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${ICEM_REFERENCE_ROOT:-$ROOT/baseline/_ref/icem}"
mkdir -p "$ROOT/gpurun_out" /tmp/dropin && cd /tmp/dropin
python - <<PY
import json
s = json.load(open("$ROOT/examples/halfcheetah_icem_b200.json"))
s["rollout_params"]["task_horizon"] = 20
s["model_dir"] = "/tmp/dropin/results"
json.dump(s, open("/tmp/dropin/settings.json", "w"))
PY
( time PYTHONPATH="$ROOT" python -m icem_b200.launch /tmp/dropin/settings.json --reference "$REF" --shims "$ROOT/oracle/shims" ) > "$ROOT/gpurun_out/r2_main_dropin.log" 2>&1
echo "exit code: $?" >> "$ROOT/gpurun_out/r2_main_dropin.log"
ls -la /tmp/dropin/results /tmp/dropin/results/checkpoints_latest/ >> "$ROOT/gpurun_out/r2_main_dropin.log" 2>&1
tail -25 "$ROOT/gpurun_out/r2_main_dropin.log"
