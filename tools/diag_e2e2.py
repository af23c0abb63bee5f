#!/usr/bin/env python
"""Diagnostic: e2e step times inside a bench-like process (torch loaded, first scene alive), with / without NVML sampling."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import lajolla_public_b200 as lj, oracle_lib
from lajolla_public_b200 import ljs
import bench
name, spp, use_nvml = sys.argv[1], int(sys.argv[2]), sys.argv[3] == "1"
desc = ljs.load(oracle_lib.scene_ljs(name))
torch.cuda.set_device(0)
scene = lj.Scene(desc)
film = torch.zeros((scene.height, scene.width, 3), dtype=torch.float32, device="cuda")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
if use_nvml:
    smp = bench.ClockSampler(0); smp.start()
for _ in range(4):
    scene.render_device(film.data_ptr(), stream.cuda_stream, spp=spp)
torch.cuda.synchronize()
if use_nvml:
    print("clocks", smp.summary())
for i in range(6):
    t0 = time.perf_counter(); sc = lj.Scene(desc); t1 = time.perf_counter()
    img = sc.render(spp=spp); t2 = time.perf_counter(); st = sc.last_stats
    sc.close(); t3 = time.perf_counter()
    print(f"nvml={use_nvml} {name} #{i}: create {1e3*(t1-t0):.1f} ms, lj_render {1e3*(t2-t1):.1f} ms (device loop {st.render_ms:.1f} ms), close {1e3*(t3-t2):.1f} ms", flush=True)
