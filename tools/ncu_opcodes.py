#!/usr/bin/env python
"""Opcode census of one kernel from an ncu report's SASS source page: warp-level instructions executed per opcode,
L1 wavefronts (shared) / tag requests (global), and the stall-sample split.
usage: ncu_opcodes.py report.ncu-rep kernel-substring [top-N]"""
import collections
import csv
import subprocess
import sys

rep, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        func, hdr, j, body = rows[i][1], rows[i + 1], i + 2, []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j]); j += 1
        i = j
        if want not in func:
            continue
        c = {h: k for k, h in enumerate(hdr)}
        agg, extra, stalls = collections.Counter(), collections.Counter(), collections.Counter()
        for r in body:
            if len(r) < len(hdr):
                continue
            toks = r[c["Source"]].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            op = op.split(".")[0].rstrip(";")
            try:
                ie = int(float(r[c["Instructions Executed"]] or 0))
            except ValueError:
                continue
            agg[op] += ie
            for col in ("L1 Tag Requests Global", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"):
                try:
                    extra[(op, col)] += int(float(r[c[col]] or 0))
                except ValueError:
                    pass
            for h in hdr:
                if h.startswith("stall_") and "Not Issued" not in h:
                    try:
                        stalls[h] += int(float(r[c[h]] or 0))
                    except ValueError:
                        pass
        tot = sum(agg.values())
        print(f"=== {func}: {tot:.4e} warp-level instructions")
        for op, v in agg.most_common(top):
            ex = {k[1]: vv for k, vv in extra.items() if k[0] == op and vv}
            print(f"  {op:10s} {v:13d} {100 * v / tot:5.1f}%  {ex if ex else ''}")
        ts = sum(stalls.values())
        print("  stall samples:", ", ".join(f"{k[6:]} {100 * v / ts:.1f}%" for k, v in stalls.most_common(10)))
    else:
        i += 1
