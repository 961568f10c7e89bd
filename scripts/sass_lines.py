#!/usr/bin/env python
"""Attribute executed warp instructions of an ncu capture to CUDA source lines (no GPU needed).

    python scripts/sass_lines.py <report.ncu-rep> <function-substring> [top]

ncu's `--page source --csv` lists SASS instructions with their executed counts in program order; `nvdisasm -g`
of the same cubin lists the same instructions with `//## File ..., line N` markers.  The two are zipped by order."""
import csv
import re
import subprocess
import sys
import os
import tempfile
from collections import defaultdict

rep, func = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "icem_b200", "lib", "libicem_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# locate function
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and func in l)
lines = []
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//--------------------- .text"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append((cur, l.strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ie = hdr.index("Instructions Executed")
sm = hdr.index("# Samples")
sass = rows[2:]
if len(sass) != len(lines):
    print(f"warning: {len(sass)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
agg = defaultdict(lambda: [0, 0])
tot = 0
for (loc, txt), r in zip(lines, sass):
    n = int(r[ie]); s = int(r[sm])
    agg[loc][0] += n; agg[loc][1] += s
    tot += n
print(f"total warp instructions executed: {tot}")
src_cache = {}
def src(loc):
    f, ln = loc
    for d in ("icem_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][ln - 1].strip()[:90] if ln - 1 < len(src_cache[p]) else ""
    return ""
for loc, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100.0 * n / tot:6.2f}%  inst {n:>12d}  samples {s:>7d}  {loc[0]}:{loc[1]:<4d} {src(loc)}")
