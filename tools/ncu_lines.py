#!/usr/bin/env python
"""Aggregate an ncu source page (`ncu -i rep --page source --csv --print-source sass,cuda`) by CUDA
source line: warp-level instructions executed, average active threads, stall samples.
usage: ncu_lines.py report.ncu-rep [kernel-substring] [top-N]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        fname = rows[i][1]
        func = rows[i + 1][1]
        hdr = rows[i + 2]
        j = i + 3
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            body.append(rows[j])
            j += 1
        i = j
        if want not in func:
            continue
        c = {h: k for k, h in enumerate(hdr)}
        # the first "Source" column is the CUDA line, the second the SASS text
        src_cols = [k for k, h in enumerate(hdr) if h == "Source"]
        agg = collections.defaultdict(lambda: [0, 0, 0, 0])
        tot_i = tot_t = tot_s = 0
        for r in body:
            if len(r) < len(hdr):
                continue
            try:
                ie = int(float(r[c["Instructions Executed"]] or 0)); te = int(float(r[c["Thread Instructions Executed"]] or 0))
                sm = int(float(r[c["# Samples"]] or 0))
            except ValueError:
                continue
            key = (r[c["Line No"]], r[src_cols[0]].strip()[:110])
            a = agg[key]
            a[0] += ie; a[1] += te; a[2] += sm; a[3] += 1
            tot_i += ie; tot_t += te; tot_s += sm
        print(f"=== {func}  ({fname})  warp-inst {tot_i:.3e}  thread-inst {tot_t:.3e}  avg threads {tot_t / max(tot_i, 1):.2f}  samples {tot_s}")
        for (ln, text), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"  {100 * a[0] / max(tot_i, 1):5.1f}% inst  {100 * a[2] / max(tot_s, 1):5.1f}% stall  thr {a[1] / max(a[0], 1):5.1f}  sass {a[3]:4d}  L{ln}: {text}")
    else:
        i += 1
