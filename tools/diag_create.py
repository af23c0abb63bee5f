#!/usr/bin/env python
"""Diagnostic: where lj_scene_create spends its time (host marshalling vs upload vs BVH/mip build)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import lajolla_public_b200 as lj, oracle_lib
from lajolla_public_b200 import ljs, abi
name = sys.argv[1] if len(sys.argv) > 1 else "sponza"
desc = ljs.load(oracle_lib.scene_ljs(name))
lib = lj.load_library(); abi.check(lib.lj_init(None, 1))
for i in range(6):
    t0 = time.perf_counter(); cdesc, keep = ljs.to_c(desc); t1 = time.perf_counter()
    h = C.c_void_p(); abi.check(lib.lj_scene_create(C.byref(cdesc), C.byref(h))); t2 = time.perf_counter()
    info = abi.lj_scene_info(); lib.lj_scene_get_info(h, C.byref(info))
    if i % 2 == 1:  # render in between like the e2e loop does
        sc = lj.Scene.__new__(lj.Scene); sc._lib = lib; sc._h = h; sc.desc = desc; sc.width, sc.height = desc.camera.width, desc.camera.height
        sc.render(spp=8)
    t3 = time.perf_counter(); lib.lj_scene_destroy(h); t4 = time.perf_counter()
    print(f"create #{i}: to_c {1e3*(t1-t0):.1f} ms, lj_scene_create {1e3*(t2-t1):.1f} ms (upload {info.upload_ms:.1f} prep {info.prep_ms:.1f} bvh {info.bvh_build_ms:.1f}), destroy {1e3*(t4-t3):.1f} ms", flush=True)
