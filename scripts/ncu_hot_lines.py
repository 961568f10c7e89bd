#!/usr/bin/env python
"""Warp-stall samples and executed instructions of an ncu capture (--set full --import-source on, -lineinfo build)
attributed to CUDA source lines; no GPU needed.

    python scripts/ncu_hot_lines.py <report.ncu-rep> [top] > profiles/<name>_hot_lines.txt
"""
import collections
import csv
import subprocess
import sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    sec = hdr = None
    samples, executed, text = collections.Counter(), collections.Counter(), {}
    kernel = ""
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            sec = r[1]
            continue
        if r[0] == "Function Name":
            kernel = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or sec is None:
            continue
        try:
            ln = int(r[0])
            s, e = int(r[hdr.index("# Samples")] or 0), int(r[hdr.index("Instructions Executed")] or 0)
        except ValueError:
            continue
        key = (sec.split("/")[-1], ln)
        samples[key] += s
        executed[key] += e
        text[key] = r[1].strip()[:110]
    ts, te = sum(samples.values()), sum(executed.values())
    print(f"# {rep}\n# kernel: {kernel}\n# {ts} warp-stall samples, {te} warp instructions attributed to source lines")
    per_file = collections.Counter()
    for (f, _), v in samples.items():
        per_file[f] += v
    print("# samples per file: " + ", ".join(f"{f} {100 * v / ts:.1f} %" for f, v in per_file.most_common(5)))
    print("# file:line  samples%  instructions%  source")
    for k, v in samples.most_common(top):
        print(f"{k[0]}:{k[1]:<5d} {100 * v / ts:6.2f} {100 * executed[k] / te:6.2f}  {text[k]}")


if __name__ == "__main__":
    main()
