"""CPU parity of the branch-parallel articulated engine: icem_b200/csrc/dyn_chain.cuh -- the source the CUDA kernels
run -- is compiled for the HOST (tests/chain_host/chain_host.cpp, one thread per lane of a group, float32) and compared
with the independent float64 oracle (oracle/articulated_np.py) on the same tables.  Test infrastructure only: the
product package never loads this library.

Tolerances (float32 articulated-body algorithm vs float64 Jacobian formulation + LAPACK solve):
  one env step from the same state   |d state| <= 2e-4   (HalfCheetah / Hopper / Humanoid*: measured <= 8e-5)
                                     Ant: <= 5e-4 (soft contacts on a 0.4 kg torso sphere amplify rounding)
  free-running h-step rollouts       stay within 5e-3 of the oracle for the first 10 control steps
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from icem_b200 import robots
from icem_b200._lib import fptr
from icem_b200.planner import articulated_model_struct
from oracle.articulated_np import make_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "chain_host", "chain_host.cpp")
OUT = os.path.join(ROOT, "build", "libchain_host.so")
DEPS = [SRC, os.path.join(ROOT, "icem_b200", "csrc", "dyn_chain.cuh"), os.path.join(ROOT, "include", "icem_b200.h")]


@pytest.fixture(scope="module")
def host_lib():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in DEPS):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", SRC, "-o", OUT], check=True)
    lib = C.CDLL(OUT)
    lib.chain_host_describe.restype = C.c_int
    lib.chain_host_rollout.restype = C.c_int
    return lib


def _describe(lib, name):
    m = robots.get_model(name)
    st, keep = articulated_model_struct(m)
    out = (C.c_int * 8)()
    why = C.create_string_buffer(128)
    rc = lib.chain_host_describe(C.byref(st), m.nu, out, why, 128)
    return rc, list(out), why.value.decode()


def test_decomposition_of_the_reference_robots(host_lib):
    """trunk + limb chains: lanes per trajectory, trunk bodies, limbs, nodes after fusing jointless bodies."""
    want = {"halfcheetah": (2, 1, 2, 7), "humanoid_standup": (4, 3, 4, 11), "hopper": (1, 1, 1, 4),
            "ant": (4, 1, 4, 9), "humanoid": (4, 3, 4, 11)}
    for name, (lanes, trunk, limbs, nodes) in want.items():
        rc, out, why = _describe(host_lib, name)
        assert rc == 0, (name, why)
        assert tuple(out[:4]) == (lanes, trunk, limbs, nodes), (name, out)
        assert out[6] * 4 <= 32 * 1024, "scratch per warp must stay small enough for >= 6 warps per SM"


@pytest.mark.parametrize("integrator", ["euler", "rk4"])
@pytest.mark.parametrize("name,tol", [("halfcheetah", 2e-4), ("humanoid_standup", 2e-4), ("hopper", 2e-4),
                                      ("ant", 5e-4), ("humanoid", 2e-4)])
def test_host_compiled_engine_matches_float64_oracle(host_lib, name, tol, integrator):
    m = robots.get_model(name, integrator=integrator)
    st, keep = articulated_model_struct(m)
    mod = make_model(name, obs_skip=0, integrator=integrator)
    rs = np.random.RandomState(5)
    n, h = 12, 10
    start = np.concatenate([m.qpos0, 0.1 * rs.randn(m.nv)])
    acts = rs.uniform(-1.3 * m.ctrl_limit, 1.3 * m.ctrl_limit, (n, h, m.nu)).astype(np.float32)   # also beyond the clip
    states = np.zeros((n, h + 1, m.nq + m.nv))
    rc = host_lib.chain_host_rollout(C.byref(st), m.nu, {"euler": 0, "rk4": 1}[integrator], n, h, start.ctypes.data_as(C.POINTER(C.c_double)),
                                     fptr(acts), states.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    np.testing.assert_allclose(states[:, 0], np.broadcast_to(start.astype(np.float32), (n, start.size)), atol=0)
    # teacher forcing: one oracle step from every state the engine visited
    worst = 0.0
    for t in range(h):
        ref = mod.step_state(states[:, t], acts[:, t].astype(np.float64))
        worst = max(worst, float(np.abs(ref - states[:, t + 1]).max()))
    assert worst <= tol, worst
    # free running
    s = np.broadcast_to(start.astype(np.float32).astype(np.float64), (n, start.size)).copy()
    for t in range(h):
        s = mod.step_state(s, acts[:, t].astype(np.float64))
    assert np.abs(s - states[:, h]).max() <= (5e-2 if name == "ant" else 5e-3)


@pytest.mark.parametrize("integrator", ["euler", "rk4"])
@pytest.mark.parametrize("name", ["halfcheetah", "hopper"])
def test_planar_instantiation_matches_oracle_and_spatial_engine(host_lib, name, integrator):
    """Planar robots run ChainLane<..., PLANAR = true> (3-vectors instead of 6-vectors): same tolerance against the
    float64 oracle as the spatial instantiation, and the two instantiations agree to fp32 rounding on one step."""
    m = robots.get_model(name, integrator=integrator)
    st, keep = articulated_model_struct(m)
    host_lib.chain_host_is_planar.restype = C.c_int
    host_lib.chain_host_rollout_engine.restype = C.c_int
    assert host_lib.chain_host_is_planar(C.byref(st), m.nu) == 1
    for other in ("humanoid_standup", "ant", "humanoid"):
        so, _k = articulated_model_struct(robots.get_model(other))
        assert host_lib.chain_host_is_planar(C.byref(so), robots.get_model(other).nu) == 0
    mod = make_model(name, obs_skip=0, integrator=integrator)
    rs = np.random.RandomState(7)
    n, h = 16, 12
    start = np.concatenate([m.qpos0, 0.2 * rs.randn(m.nv)])
    acts = rs.uniform(-1.3 * m.ctrl_limit, 1.3 * m.ctrl_limit, (n, h, m.nu)).astype(np.float32)
    out = {}
    for planar in (0, 1):
        states = np.zeros((n, h + 1, m.nq + m.nv))
        rc = host_lib.chain_host_rollout_engine(C.byref(st), m.nu, {"euler": 0, "rk4": 1}[integrator], planar, n, h,
                                                start.ctypes.data_as(C.POINTER(C.c_double)), fptr(acts),
                                                states.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        worst = 0.0
        for t in range(h):
            ref = mod.step_state(states[:, t], acts[:, t].astype(np.float64))
            worst = max(worst, float(np.abs(ref - states[:, t + 1]).max()))
        assert worst <= 2e-4, (planar, worst)
        out[planar] = states
    # first step from the identical start state: the two instantiations differ by fp32 rounding only
    assert np.abs(out[0][:, 1] - out[1][:, 1]).max() <= 2e-5
    # a non-planar robot is refused by the planar engine
    so, _k = articulated_model_struct(robots.get_model("ant"))
    ant = robots.get_model("ant")
    rc = host_lib.chain_host_rollout_engine(C.byref(so), ant.nu, 0, 1, 1, 1,
                                            np.zeros(ant.nq + ant.nv).ctypes.data_as(C.POINTER(C.c_double)),
                                            fptr(np.zeros((1, 1, ant.nu), np.float32)),
                                            np.zeros((1, 2, ant.nq + ant.nv)).ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 2


def test_twelve_humanoid_warps_fit_in_shared_memory(host_lib):
    """The launch of the fused kernel holds up to 12 warps per SM for HumanoidStandup (csrc/planner.cu::launch_chain_impl,
    rollout_chain.cuh::chain_rollout_smem_bytes): model + sampler tables + 12 warp scratches must stay inside the 227 KB
    a CTA may use -- a few hundred bytes of slack, so growing ChainModel or a record silently costs a warp (and a whole
    round at N = 13107 rows per launch).  Same arithmetic as chain_rollout_smem_bytes<true> for h = 30, d = 17."""
    rc, out, why = _describe(host_lib, "humanoid_standup")
    assert rc == 0, why
    warp_floats, model_bytes = out[6], out[7]
    h, d, K = 30, 17, 16
    gs = 2 * K + 1
    fixed = ((model_bytes // 4 + 3) & ~3) + ((h * gs + 3) & ~3) + 2 * ((h * d + 3) & ~3) + 2 * ((d + 3) & ~3)
    stride = (h * d + 3) & ~3
    sample_floats = 2 * stride + ((d * gs + 3) & ~3)
    per_warp = max(sample_floats, (warp_floats + 3) & ~3)
    assert (fixed + 12 * per_warp) * 4 <= 227 * 1024, ((fixed + 12 * per_warp) * 4, model_bytes, warp_floats)
