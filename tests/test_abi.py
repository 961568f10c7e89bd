"""CPU tests of the drop-in boundary: the C-ABI shared library builds, loads without a GPU, and exports exactly the
symbols include/icem_b200.h declares; no compute is called.  Also: the product package never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "icem_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(icem_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from icem_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/icem_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in icem_b200/_lib.py"
    assert set(_lib.SIGNATURES) <= set(names), set(_lib.SIGNATURES) - set(names)
    loaded = _lib.load()
    assert loaded.icem_abi_version() == _lib.ICEM_ABI_VERSION


def test_struct_layouts_match_the_header():
    """sizeof(icem_config_t) / sizeof(icem_articulated_model_t) as gcc sees the header == the ctypes mirrors."""
    from icem_b200 import _lib
    code = ('#include "%s"\n#include <stdio.h>\nint main(){printf("%%zu %%zu\\n", sizeof(icem_config_t), '
            'sizeof(icem_articulated_model_t));return 0;}\n' % HEADER)
    exe = os.path.join(ROOT, "build", "abi_sizes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-o", exe], input=code, text=True, check=True)
    a, b = map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    assert a == ctypes.sizeof(_lib.IcemConfig)
    assert b == ctypes.sizeof(_lib.IcemArticulatedModel)


def test_create_fails_loudly_without_a_gpu():
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("asserts the no-GPU failure mode")
    from icem_b200.planner import IcemError, Planner, PlannerSettings
    with pytest.raises(IcemError):
        Planner(PlannerSettings(horizon=5, num_simulated_trajectories=8, action_low=-np.ones(2), action_high=np.ones(2)))


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "icem_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
    code = "import sys; sys.path.insert(0, %r); import icem_b200, icem_b200.planner, icem_b200.controller, " \
           "icem_b200.models, icem_b200.envs, icem_b200.workloads, icem_b200.robots, icem_b200.distributed; " \
           "assert not [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]" % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)
