#!/usr/bin/env python
"""Characterise the fp32-vs-fp64 error tail of the device BSDF code: per material type, the fraction
of eval / pdf queries whose relative error against the reference's double-precision material.o exceeds 1e-5 / 1e-4 /
1e-3, and the worst queries with their inputs.  Runs on the GPU library, or with --hostsim on the g++ build of the
same device sources (no GPU needed).
usage: bsdf_tail.py [--hostsim] [--n 200000] scene [scene ...]"""
import argparse
import contextlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lajolla_public_b200 as lj
import oracle_lib
import parity_checks as pc

MAT = ["lambertian", "roughplastic", "roughdielectric", "disney_diffuse", "disney_metal", "disney_glass", "disney_clearcoat", "disney_sheen", "disney_bsdf"]

ap = argparse.ArgumentParser()
ap.add_argument("--hostsim", action="store_true")
ap.add_argument("--n", type=int, default=200000)
ap.add_argument("--worst", type=int, default=10)
ap.add_argument("scenes", nargs="+")
args = ap.parse_args()
if args.hostsim:
    import hostsim_lib

for name in args.scenes:
    with (hostsim_lib.simulated() if args.hostsim else contextlib.nullcontext()):
        sc = lj.parse_scene(oracle_lib.scene_ljs(name))
        ref = oracle_lib.RefScene(oracle_lib.scene_xml(name), threads=4)
        rays = pc.primary_rays(ref, args.n)
        q = pc.make_bsdf_queries(ref, rays)
        r1, r2 = sc.bsdf(q), ref.bsdf(q)
    mat_type = np.array([sc.desc.materials[m].type for m in q["vertex"]["material_id"]])
    n_l = (q["vertex"]["frame_n"].astype(np.float64))
    cos_i = (q["dir_in"] * n_l).sum(axis=1)
    cos_o = (q["dir_out"] * n_l).sum(axis=1)
    for t in np.unique(mat_type):
        sel = mat_type == t
        for what in ("f", "pdf"):
            a = r1[what][sel].astype(np.float64).reshape(sel.sum(), -1)
            b = r2[what][sel].astype(np.float64).reshape(sel.sum(), -1)
            nz = (np.abs(b).max(axis=1) > 1e-12) & (np.abs(a).max(axis=1) > 1e-12)
            if not nz.any():
                continue
            e = (np.abs(a - b) / np.maximum(np.abs(b), 1e-9)).max(axis=1)
            e = np.where(nz, e, 0)
            row = dict(scene=name, material=MAT[t], quantity=what, n=int(nz.sum()), median=float(np.median(e[nz])),
                       frac_gt_1e5=float((e[nz] > 1e-5).mean()), frac_gt_1e4=float((e[nz] > 1e-4).mean()), frac_gt_1e3=float((e[nz] > 1e-3).mean()),
                       max=float(e.max()))
            idx = np.flatnonzero(sel)[np.argsort(-e)[:args.worst]]
            row["worst"] = [dict(err=float(e[np.flatnonzero(sel) == i][0]), cos_in=float(cos_i[i]), cos_out=float(cos_o[i]),
                                 half_dot_n=float(((q["dir_in"][i] + q["dir_out"][i]) / max(np.linalg.norm(q["dir_in"][i] + q["dir_out"][i]), 1e-30) * n_l[i]).sum()),
                                 device=np.atleast_1d(r1[what][i]).tolist(), reference=np.atleast_1d(r2[what][i]).tolist()) for i in idx]
            print(json.dumps(row), flush=True)
