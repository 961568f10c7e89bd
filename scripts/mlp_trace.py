#!/usr/bin/env python
"""Debug: per-step cycle stamps of CTA 0 of mlp_rollout_kernel (library built with -DICEM_MLP_TRACE).

    ICEM_B200_LIB=icem_b200/lib/libicem_b200_trace.so python scripts/mlp_trace.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icem_b200 import _lib, workloads  # noqa: E402
from icem_b200.planner import Planner  # noqa: E402

name = "mlp_cheetah_n65536"
w = workloads.get_workload(name)
p = Planner(workloads.planner_settings(name))
p.set_mlp_model(*workloads.mlp_model_weights(*w["mlp"]))
p.begin_rollout()
p.bench_op("rollout", 65536, reps=1, flush_l2=False)
lib = _lib.load()
buf = np.zeros(64 * 32, np.int64)
lib.icem_debug_mlp_trace.argtypes = [C.c_void_p, C.c_int]
rc = lib.icem_debug_mlp_trace(buf.ctypes.data, buf.size)
assert rc == 0, rc
t = buf.reshape(64, 32)
names2 = {0: "x", 1: "d1A", 2: "e1A", 3: "d1B", 4: "e1B", 5: "d2A", 6: "e2A", 7: "d2B", 8: "e2B", 9: "d3A", 10: "oA", 11: "d3B",
          12: "oB", 16: "I:xA", 17: "I:xB", 18: "I:h1A", 19: "I:h1B", 20: "I:h2A", 21: "I:h2B"}
names = {0: "x_arrive", 1: "d1_ready", 2: "e1s0", 3: "e1s1", 4: "e1s2", 5: "e1s3", 7: "d2_ready", 8: "e2s0", 9: "e2s1",
         10: "e2s2", 11: "e2s3", 13: "d3_ready", 14: "step_end", 16: "I:x", 17: "I:a0", 18: "I:a1", 19: "I:a2", 20: "I:a3",
         21: "I:b0", 22: "I:b1", 23: "I:b2", 24: "I:b3", 25: "I:commit3"}
    base = t[step, 0]
    ev = sorted((t[step, k] - base, v) for k, v in names.items() if t[step, k])
    print("step", step, " ".join(f"{v}@{c}" for c, v in ev), "| next x_arrive @", t[step + 1, 0] - base)
