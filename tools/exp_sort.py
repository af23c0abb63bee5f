#!/usr/bin/env python
"""Experiment: how much does ray ORDER matter to the persistent closest-hit kernel?  Traces the same batch of bounce
rays (sponza) in random order and sorted by (direction octant, Morton cell of the origin), through k_trace_q<0>.
usage (GPU box): python tools/exp_sort.py [scene] [log2 n]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lajolla_public_b200 as lj
from lajolla_public_b200 import abi
import oracle_lib

name = sys.argv[1] if len(sys.argv) > 1 else "sponza"
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 21)
sc = lj.parse_scene(oracle_lib.scene_ljs(name))
info = sc.info()
eps = info.shadow_epsilon
rng = np.random.default_rng(1)


def bounce(rays):
    v = sc.intersect(rays)
    v = v[v["shape_id"] >= 0]
    d = rng.normal(size=(v.shape[0], 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    flip = (d * v["geometric_normal"]).sum(axis=1) < 0
    d[flip] *= -1
    return lj.make_rays(v["position"], d.astype(np.float32), tnear=np.float32(eps), tfar=np.inf)


def sort_key(rays, bits):
    lo, hi = np.array(info.bounds_lo), np.array(info.bounds_hi)
    g = np.clip(((rays["org"] - lo) / (hi - lo) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    m = np.zeros(rays.shape[0], dtype=np.int64)
    for b in range(bits):
        for a in range(3):
            m |= ((g[:, a] >> b) & 1) << (3 * b + a)
    octant = (rays["dir"][:, 0] < 0) * 1 + (rays["dir"][:, 1] < 0) * 2 + (rays["dir"][:, 2] < 0) * 4
    return (octant.astype(np.int64) << (3 * bits)) | m


def timed(rays, label, kernel=abi.LJ_TRACE_WAVEFRONT):
    best = 1e9
    for _ in range(3):
        _, ms = sc.intersect_hits(rays, want_ms=True, kernel=kernel)
        best = min(best, ms)
    print(f"  {label:32s} {best:8.3f} ms  {rays.shape[0] / best / 1e6:7.2f} Grays/s", flush=True)


r0 = sc.sample_primary(rng.random((n, 2)).astype(np.float32))
r1 = bounce(r0)
r2 = bounce(r1)
for label, rays in (("primary (random pixels)", r0), ("bounce 1", r1), ("bounce 2", r2)):
    print(f"{name}: {label}, {rays.shape[0]} rays")
    perm = rng.permutation(rays.shape[0])
    timed(rays[perm], "random order")
    timed(rays[perm], "random order, lane kernel", abi.LJ_TRACE_WAVEFRONT_LANE)
    for bits in (3, 4, 5, 7):
        k = sort_key(rays, bits)
        timed(rays[np.argsort(k, kind="stable")], f"sorted octant + {bits}-bit morton")
    k = sort_key(rays, 5) & ((1 << 15) - 1)
    timed(rays[np.argsort(k, kind="stable")], "sorted 5-bit morton only")
