#!/usr/bin/env python
"""Diagnostic: determinism of a render and additivity of sample blocks on the GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lajolla_public_b200 as lj, oracle_lib
name = sys.argv[1] if len(sys.argv) > 1 else "cbox"
sc = lj.parse_scene(oracle_lib.scene_ljs(name))
def cmp(tag, a, b):
    d = np.abs(a - b); rel = d / np.maximum(np.abs(b), 1e-3)
    print(f"{tag}: max abs {d.max():.3e} max rel {rel.max():.3e} n(rel>1e-4) {(rel > 1e-4).sum()} of {d.size}; sums {a.sum():.6e} {b.sum():.6e}", flush=True)
full = sc.render(spp=8, normalize=False); st = sc.last_stats
print("full: samples", st.samples, "closest", st.closest_rays, "shadow", st.shadow_rays, "bounces", st.bounces, "waves", st.waves)
full2 = sc.render(spp=8, normalize=False); st2 = sc.last_stats
print("full2: samples", st2.samples, "closest", st2.closest_rays, "shadow", st2.shadow_rays, "bounces", st2.bounces, "waves", st2.waves)
cmp("full vs full2", full, full2)
parts = []
for b in range(0, 8, 2):
    parts.append(sc.render(spp=8, sample_begin=b, sample_end=b + 2, normalize=False)); s = sc.last_stats
    print(" part", b, "samples", s.samples, "closest", s.closest_rays, "shadow", s.shadow_rays, "bounces", s.bounces, "waves", s.waves)
cmp("sum(parts) vs full", sum(parts), full)
half = sc.render(spp=8, sample_begin=0, sample_end=4, normalize=False) + sc.render(spp=8, sample_begin=4, sample_end=8, normalize=False)
cmp("two halves vs full", half, full)
p0 = sc.render(spp=8, sample_begin=0, sample_end=2, normalize=False)
cmp("part0 again vs part0", p0, parts[0])
