#!/usr/bin/env python
"""Tuning helper: runs bench.py (sponza, reduced spp) over a grid of k_trace scheduling parameters and prints
the per-stage times.  usage: sweep_trace.py "P1,P2,..." "R1,R2,..." [spp]"""
import json
import os
import subprocess
import sys

prims = [int(x) for x in sys.argv[1].split(",")]
refills = [int(x) for x in sys.argv[2].split(",")]
spp = sys.argv[3] if len(sys.argv) > 3 else "64"
extra = dict(kv.split("=") for kv in sys.argv[4:])  # more LJ_* environment settings
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in prims:
    for r in refills:
        env = dict(os.environ, LJ_PRIM_MIN_LANES=str(p), LJ_REFILL=str(r), **extra)
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--spp", spp, "--steps", "2", "--warmup", "1", "--no-cpu-baseline"],
                             capture_output=True, text=True, env=env).stdout.strip().splitlines()
        try:
            j = json.loads(out[-1])
            st = j["stage_ms_per_step"]
            print(f"{extra} prim_min={p:2d} refill={r:2d}  Msamples/s={j['value']:7.1f}  extend={st['extend_ms']:7.1f} shadow={st['shadow_ms']:6.1f} "
                  f"shade={st['shade_ms']:6.1f} regen={st['regen_ms']:6.1f} render={st['render_ms']:7.1f} nodes/ray={j['node_steps_per_ray']:.2f} prims/ray={j['prim_tests_per_ray']:.2f}", flush=True)
        except Exception as e:
            print(p, r, "failed", e, out[-1:] if out else "")
