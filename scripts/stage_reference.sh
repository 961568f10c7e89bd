#!/bin/bash
# Stage the UNMODIFIED reference sources where a GPU box can see them: baseline/_ref/icem is git-ignored (never
# committed) but travels with gpurun snapshots, like the pip-installed reference of the base bench contract would.
# The vendored mj_envs suite (27 MB of meshes, out of scope) is left out.  oracle/ref_loader.py and
# icem_b200/launch.py find the copy when /root/reference is absent.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference/icem}"
if [ ! -f "$SRC/main.py" ]; then echo "no reference under $SRC" >&2; exit 1; fi
mkdir -p "$ROOT/baseline/_ref"
rm -rf "$ROOT/baseline/_ref/icem"
mkdir -p "$ROOT/baseline/_ref/icem"
(cd "$SRC" && tar --exclude='environments/mj_envs' --exclude='__pycache__' -cf - .) | (cd "$ROOT/baseline/_ref/icem" && tar -xf -)
du -sh "$ROOT/baseline/_ref/icem"
