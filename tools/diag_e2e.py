#!/usr/bin/env python
"""Diagnostic: host-side phases of a fresh-scene render (LJ_PROFILE_HOST=1 prints lj_render's own phases)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lajolla_public_b200 as lj, oracle_lib
from lajolla_public_b200 import ljs
name, spp = sys.argv[1], int(sys.argv[2])
desc = ljs.load(oracle_lib.scene_ljs(name))
for i in range(3):
    t0 = time.perf_counter(); sc = lj.Scene(desc); t1 = time.perf_counter()
    img = sc.render(spp=spp); t2 = time.perf_counter(); st = sc.last_stats
    sc.close(); t3 = time.perf_counter()
    print(f"{name} #{i}: create {1e3*(t1-t0):.1f} ms, lj_render {1e3*(t2-t1):.1f} ms (device loop {st.render_ms:.1f} ms, waves {st.waves}), close {1e3*(t3-t2):.1f} ms", flush=True)
