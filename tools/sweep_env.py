#!/usr/bin/env python
"""Tuning helper: runs bench.py once per environment setting and prints the per-stage times.
usage: sweep_env.py VAR v1,v2,... [bench args ...]   e.g.  sweep_env.py LJ_CHUNK 32,64,128 --spp 128 --workload sponza"""
import json
import os
import subprocess
import sys

var, vals, rest = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for v in vals:
    env = dict(os.environ)
    if v != "-":
        env[var] = v
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "2", "--warmup", "1", "--no-cpu-baseline"] + rest,
                         capture_output=True, text=True, env=env).stdout.strip().splitlines()
    try:
        j = json.loads(out[-1])
        st = j["stage_ms_per_step"]
        print(f"{var}={v:>4}  Msamples/s={j['value']:7.1f} e2e={j['e2e']['value']:7.1f}  extend={st['extend_ms']:7.1f} shadow={st['shadow_ms']:6.1f} "
              f"shade={st['shade_ms']:6.1f} regen={st['regen_ms']:6.1f} render={st['render_ms']:7.1f}", flush=True)
    except Exception as e:
        print(var, v, "failed", e, out[-1:] if out else "")
