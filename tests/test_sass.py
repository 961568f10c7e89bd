"""CPU guard on the compiled kernels (no GPU needed: nvcc cross-compiles, cuobjdump reads the cubin): the resource
budgets the occupancy of each hot kernel relies on, and the instructions that prove which hardware paths are used.

The fused ground-truth kernel is sensitive to its register allocation (+-3 % from three live
registers, +9 % from outlining the rollout), so a change that silently alters these numbers should fail here, before
any GPU time is spent."""
import re
import subprocess

import pytest


@pytest.fixture(scope="module")
def lib():
    from icem_b200 import build
    return build.build()


def _res_usage(lib):
    out = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True, check=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+)", out):
        res[m.group(1)] = dict(reg=int(m.group(2)), stack=int(m.group(3)), shared=int(m.group(4)))
    return res


def _find(res, *parts):
    hits = [k for k in res if all(p in k for p in parts)]
    assert len(hits) == 1, (parts, hits)
    return res[hits[0]]


def test_register_budgets_of_the_hot_kernels(lib):
    res = _res_usage(lib)
    # fused sample -> rollout -> cost, ground-truth dynamics: 80 registers = 3 CTAs of 8 warps per SM
    for nv in (12, 24):
        k = _find(res, "rollout_kernel", f"ArticulatedILi{nv}EEELb1ELb1ELb0")
        assert k["reg"] <= 80 and k["stack"] <= 192, k
    # stand-alone sampler: 3 CTAs per SM at K <= 16 (256 threads / 80 registers for the run-time-shaped kernel, 192
    # threads / 96 for the compile-time-shaped ones), no stack
    k = _find(res, "colored_sampler_kernelILi16ELi0ELi0E")
    assert k["reg"] <= 80 and k["stack"] == 0, k
    for shape in ("ILi16ELi30ELi17E", "ILi16ELi30ELi6E", "ILi8ELi12ELi6E"):
        k = _find(res, "colored_sampler_kernel" + shape)
        assert k["reg"] <= 96 and k["stack"] == 0, k
    # tensor-core MLP rollout: 17 warps -> 96 registers at most (65536 / (5 warps * 32 lanes) per scheduler)
    k = _find(res, "mlp_rollout_kernelILb0E")
    assert k["reg"] <= 96, k
    # branch-parallel articulated engine: up to 12 warps per SM (384 threads) -> at most 170 registers, no spills
    for g in (1, 2, 4):
        k = _find(res, "chain_rollout_kernel", f"ILi{g}ELb1ELb1ELb0ELb0")
        assert k["reg"] <= 170 and k["stack"] <= 64, k
    for g in (1, 2):       # planar instantiation (HalfCheetah, Hopper)
        k = _find(res, "chain_rollout_kernel", f"ILi{g}ELb1ELb1ELb0ELb1")
        assert k["reg"] <= 170 and k["stack"] <= 64, k
    k = _find(res, "select_kernel")
    assert k["reg"] <= 64 and k["shared"] <= 16 * 1024, k


def test_sass_shows_the_blackwell_paths(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs = {}
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)

    def body(*parts):
        hits = [k for k in funcs if all(p in k for p in parts)]
        assert len(hits) == 1, (parts, hits)
        return "\n".join(funcs[hits[0]])
    mlp = body("mlp_rollout_kernelILb0E")
    for mnemonic in ("UTCHMMA", "LDTM", "UTCBAR", "MUFU.TANH", "SYNCS"):      # tcgen05.mma / .ld / .commit, mbarriers
        assert mnemonic in mlp, mnemonic
    chain = body("chain_rollout_kernel", "ILi4ELb1ELb1ELb0ELb0")
    assert "UBLKCP" in chain                                                  # TMA bulk store of the sampled tiles
    assert "SHFL.BFLY" in chain                                               # junction sums inside a lane group
    assert "MUFU.RCP" in chain and "LDS" in chain
    fused = body("rollout_kernel", "ArticulatedILi24EEELb1ELb1ELb0")
    assert "UBLKCP" in fused                                                  # TMA bulk store of the action tile
    assert "CALL.REL.NOINC" in fused or "CALL.ABS.NOINC" in fused             # the rollout is compiled out of line
    dense = body("rollout_kernel", "DenseTanhELb1ELb1ELb0")
    assert "BAR.SYNC" in dense                                                # (cheap model: rollout inlined)
    loads = body("rollout_kernel", "ArticulatedILi24EEELb0ELb1ELb0")
    assert "UBLKCP" in loads and "SYNCS" in loads                             # TMA bulk loads on mbarriers
    sampler = body("colored_sampler_kernelILi16ELi0ELi0E")                    # run-time shape: table rows from smem
    assert "UBLKCP" in sampler and "MUFU.LG2" in sampler and "LDS.128" in sampler
    sampler = body("colored_sampler_kernelILi16ELi30ELi17E")                  # BASELINE shape: packed-fp32 fold whose
    assert "UBLKCP" in sampler and "FFMA2" in sampler and "LDCU" in sampler  # table operands are kernel-parameter constants
    assert "LDS.128" not in sampler and "LDL" not in sampler and "STL" not in sampler
    select = body("select_kernel")
    assert "REDUX" in select or "CREDUX" in select                            # 64-bit warp minima by two REDUX.MIN
