"""Build libicem_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m icem_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libicem_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_input():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(os.path.dirname(HERE), "include", "icem_b200.h"))
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False, out=None, defines=()):
    """`out` / `defines`: A/B variants of the library for kernel experiments (loaded with ICEM_B200_LIB=<path>)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    target = out or LIB_PATH
    if not force and os.path.exists(target) and os.path.getmtime(target) >= _newest_input():
        return target
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = ([nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + sources()
           + ["-o", target, "-ldl"])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libicem_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
