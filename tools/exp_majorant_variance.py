#!/usr/bin/env python
"""Variance and cost of the volpath estimator with block-wise majorants (default) against the global majorant
(LJ_MAJ_BLOCK=0): mean per-pixel variance of the mean, render time, and their product (inverse efficiency)."""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import numpy as np
import lajolla_public_b200 as lj
import oracle_lib
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for name in ("hetvol", "hetvol_colored"):
    rows = {}
    for B in ("0", "4", "8", "16"):
        os.environ["LJ_MAJ_BLOCK"] = B
        sc = lj.parse_scene(oracle_lib.scene_ljs(name))
        sc.render(spp=8)
        img, var = sc.render(spp=spp, variance=True)
        ms = sc.last_stats.render_ms
        sc.close()
        rows[B] = (float(var.mean()), float(np.median(var)), ms, img.mean(axis=(0, 1)))
    v0, _, ms0, _ = rows["0"]
    for B, (v, med, ms, m) in rows.items():
        print(f"{name} spp={spp} LJ_MAJ_BLOCK={B:>2}: mean var {v:.3e} (x{v / v0:.2f})  median var {med:.3e}  {ms:8.1f} ms (x{ms / ms0:.2f})  var*time x{v * ms / (v0 * ms0):.2f}  mean {m}", flush=True)
