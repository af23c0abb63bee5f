#!/usr/bin/env python
"""bench.py -- headline benchmark: Msamples/s (camera paths/s) of the rendering hot path.

A "step" is one full render of the workload scene: every pixel x spp camera paths through the
wavefront path tracer (regen / extend / shade / shadow kernels of libljb200.so).  Workload at N=1 is
BASELINE.json configs[3]: scenes/sponza/sponza.xml, 768x575, path integrator, 1024 spp.
  value  : whole-job Msamples/s with the scene resident in HBM (lj_render_device into device memory)
  e2e    : the same through the host-buffer C ABI: lj_scene_create (H2D of the flat scene + GPU BVH/mip
           build) + lj_render (D2H of the w*h*3 fp32 image) inside the timed region
  N > 1  : one process per GPU (torchrun).  Default "strong" scaling: the image and its sample count are FIXED
           (BASELINE config 4: sponza at 1024 spp on 1/2/4/8 GPUs); rank r renders its contiguous block of the
           per-pixel sample indices (--split spp, default) or its interleaved share of the 8x4-pixel tiles
           (--split tiles), then one NCCL sum-reduce of the fp32 film.  --scaling weak keeps the per-GPU
           sample count fixed instead (image spp grows with N).
  --impl reference : lajolla's own CPU render() (unmodified reference objects + the Embree-API shim,
           oracle/_ref) on the host cores, on a bounded spp sample of the same scene.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (oracle scene key, spp at full size, cpu sample spp)
    "sponza": ("sponza", 1024, 4),
    "cbox": ("cbox", 64, 8),
    "veach_mi": ("veach_mi", 256, 8),
    "disney_bsdf": ("disney_bsdf", 256, 4),
    "volpath_test6": ("volpath_test6", 1024, 4),
    "hetvol": ("hetvol", 1024, 1),
    "hetvol_colored": ("hetvol_colored", 1024, 1),
    "vol_cbox_teapot": ("vol_cbox_teapot", 1024, 2),
}
VOLPATH = {"volpath_test6", "hetvol", "hetvol_colored", "vol_cbox_teapot"}
BYTES_PER_EXTENSION_RAY = 104  # 2R + 2H, R = 32 B ray, H = 20 B hit


def workload_name(key, w, h, full_spp):
    """The same string in both arms (b200 and --impl reference)."""
    return f"{key} {w}x{h} {'volpath' if key in VOLPATH else 'path'} integrator, {full_spp} spp"


# dram__bytes_read.sum + dram__bytes_write.sum of one k_trace_q<0> launch per ray it traced, from the ncu --set full
# capture profiles/r02y_sponza_metrics.txt (sponza, steady-state wave of a 4 Mi-slot pool: 217.6 + 50.4 MB for 3.9 M
# rays); a launch of the bench's larger pool moves proportionally more.  Other workloads have no capture of that kernel.
NCU_TRAFFIC_BYTES_PER_RAY = {"sponza": (217.648640e6 + 50.367488e6) / 3.9036e6}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks + throttle reasons through NVML while the timed region runs (in-process: forking
    nvidia-smi from a process that holds a CUDA context stalls the thread that is queueing kernels)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(x) for x in vis.split(",") if x.strip().isdigit()]
            phys = ids[gpu_index] if gpu_index < len(ids) else gpu_index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        n = self.nvml
        if n is None:
            return
        while not self.stop_flag.is_set():
            try:
                self.rows.append((n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM),
                                  n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM),
                                  n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)))
            except Exception:
                pass
            self.stop_flag.wait(0.5)  # NVML queries take driver locks the kernel-queueing thread also needs: keep them rare

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=3)
        n = self.nvml
        sm = sorted(r[0] for r in self.rows)
        reasons = set()
        if n is not None:
            names = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap}
            for r in self.rows:
                for nm, bit in names.items():
                    if r[2] & bit:
                        reasons.add(nm)
        return {"sm_mhz": float(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": float(max(r[1] for r in self.rows)) if self.rows else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def scene_paths(workload):
    import oracle_lib
    key = WORKLOADS[workload][0]
    return oracle_lib.scene_ljs(key), oracle_lib.scene_xml(key)


def run_reference(args):
    """Reference arm: the reference's render() on the host cores (oracle/_ref)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle_lib
    key, full_spp, cpu_spp = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    ref = oracle_lib.RefScene(oracle_lib.scene_xml(key), threads=cores)
    spp = args.cpu_spp or cpu_spp
    secs = []
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)  # the reference prints a progress line per tile
    try:
        shares = []
        for i in range(args.warmup + args.steps):
            c0 = oracle_lib.ray_counters()
            img, s, share = oracle_lib.timed_render(ref, spp, cores)
            c1 = oracle_lib.ray_counters()
            if i >= args.warmup:
                secs.append((s, c1[0] - c0[0] + c1[1] - c0[1]))
                shares.append(share)
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
    h, w = img.shape[:2]
    total = sum(s for s, _ in secs)
    value = w * h * spp * len(secs) / total / 1e6
    mrays = sum(r for _, r in secs) / total / 1e6
    sample = f"{key} {w}x{h} at {spp} spp per step (Msamples/s is spp-invariant), lajolla+shim" + (" + handout overlay (lajolla_ref_hw)" if key in VOLPATH or key.startswith("disney") else "")
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(secs), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(key, w, h, full_spp), "timed_sample_spp": spp},
        "mrays_per_s": mrays,
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "reference", "sample": sample,
                         "shim_share": sum(shares) / max(len(shares), 1),
                         "shim_share_note": "fraction of the CPU time (all threads) spent inside the Embree-API shim's rtcIntersect1 / rtcOccluded1; real Embree would shrink this part"},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sponza", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override the workload's spp (makes the run a non-headline one)")
    ap.add_argument("--cpu-spp", type=int, default=0)
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--split", default="tiles", choices=["spp", "tiles"])
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"], help="diagnostic: gloo reduces the film through host memory")
    ap.add_argument("--no-reduce", action="store_true", help="diagnostic: skip the film reduction (the image stays split across ranks)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import lajolla_public_b200 as lj
    from lajolla_public_b200 import ljs

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on stdout, which must carry exactly one JSON line
    if world > 1:
        # The one collective of a step is a 5 MB film reduce: in-switch reduction (NVLS) buys it nothing, and a process
        # in which NCCL has set NVLS up runs k_trace_q<0> 18 % slower (measured, profiles/r02e_nccl_nvls.txt: 353 ->
        # 396 Msamples/s at N=2 with NVLS off, = 2 x the single-GPU rate).  An explicit setting in the environment wins.
        os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
        import torch.distributed as dist

    def init_dist():
        if dist is None:
            return
        if args.backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group("gloo")

    # Order matters (measured, profiles/r02z4_nccl.txt): device memory allocated AFTER the NCCL communicator exists makes
    # the random-access traversal kernel 8-15 % slower (k_trace_q<0>: 586 -> 630..680 ms per 512 spp, whatever the NCCL
    # transport settings; shade / shadow / regen unaffected), memory allocated before it does not.  So the scene, the
    # path pool and the traversal scratch are allocated -- one untimed render of this rank's share -- before
    # init_process_group, the way an application loads its scene before it sets up communication.  The later scenes of
    # the e2e leg reuse those blocks (the library recycles them).  LJ_BENCH_EARLY_DIST=1 restores the other order (A/B).
    early = os.environ.get("LJ_BENCH_EARLY_DIST") == "1"
    if early:
        init_dist()
    key, full_spp, cpu_spp = WORKLOADS[args.workload]
    spp = args.spp or full_spp
    # The scene goes through the product's own front end (lajolla_public_b200/host: XML + mesh / image / volume
    # loaders, `lajolla --dump-ljs`) from the scene file itself; the reference parser's dump (.ljs) is only the
    # fallback when the scene file did not travel with the checkout.  Outside the timed region either way.
    ljs_path, xml_path = scene_paths(args.workload)
    if os.path.exists(xml_path) and os.path.exists(lj.LAJOLLA_CLI):
        desc = lj.load_scene_description(xml_path)
        scene_source = "scene XML parsed by the product front end (lajolla --dump-ljs)"
    else:
        desc = ljs.load(ljs_path)
        scene_source = "reference scene flattened to .ljs"
    scene = lj.Scene(desc, device=local_rank)
    info = scene.info()
    w, h = scene.width, scene.height
    npix = w * h
    strong = args.scaling == "strong"
    total_spp = spp if strong else spp * world
    film = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()  # a non-blocking stream of our own: the legacy default stream serialises against every other stream
    torch.cuda.set_stream(stream)

    from lajolla_public_b200 import partition
    tile_stride, tile_offset = 0, 0
    if args.split == "tiles" and world > 1:
        s_begin, s_end = 0, total_spp              # every sample of this rank's interleaved tiles
        tile_stride, tile_offset = world, rank
    elif strong:
        s_begin, s_end = partition.sample_ranges(total_spp, world)[rank]
    else:
        _, s_begin, s_end = partition.weak_range(rank, world, spp)

    def step():
        # every rank renders its share of the total_spp-sample image into its own film (raw sums)
        st = scene.render_device(film.data_ptr(), stream.cuda_stream, spp=total_spp, sample_begin=s_begin,
                                 sample_end=s_end, normalize=False, pool_paths=args.pool,
                                 tile_stride=tile_stride, tile_offset=tile_offset)
        if not args.no_reduce:
            partition.reduce_film(film, dist, 0)  # the one collective
        return st

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if not early and dist is not None:
        scene.render_device(film.data_ptr(), stream.cuda_stream, spp=total_spp, sample_begin=s_begin, sample_end=s_end, normalize=False,
                            pool_paths=args.pool, tile_stride=tile_stride, tile_offset=tile_offset)
        torch.cuda.synchronize()
        init_dist()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agg = {"extend_ms": 0.0, "shadow_ms": 0.0, "shade_ms": 0.0, "regen_ms": 0.0, "render_ms": 0.0, "closest": 0, "shadow": 0,
           "bounces": 0, "launches": 0, "extend_launches": 0, "samples": 0, "waves": 0, "node_steps": 0, "prim_tests": 0, "node_passes": 0, "prim_passes": 0}
    e0.record(stream)
    for _ in range(args.steps):
        st = step()
        agg["extend_ms"] += st.extend_ms; agg["shadow_ms"] += st.shadow_ms; agg["shade_ms"] += st.shade_ms
        agg["regen_ms"] += st.regen_ms; agg["render_ms"] += st.render_ms
        agg["closest"] += st.closest_rays; agg["shadow"] += st.shadow_rays; agg["bounces"] += st.bounces
        agg["launches"] += st.kernel_launches + (1 if dist is not None else 0)
        agg["extend_launches"] += st.extend_launches; agg["samples"] += st.samples; agg["waves"] += st.waves
        agg["node_steps"] += st.node_steps; agg["prim_tests"] += st.prim_tests
        agg["node_passes"] += st.node_passes; agg["prim_passes"] += st.prim_passes
        agg["pool_paths"] = int(st.pool_paths)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.summary()

    # ---- e2e through the host-buffer C ABI (rank-local; max over ranks)
    cdesc_bytes = sum(im.nbytes for im in desc.images) + sum(
        (s.positions.nbytes + s.indices.nbytes + (s.normals.nbytes if s.normals is not None else 0) +
         (s.uvs.nbytes if s.uvs is not None else 0)) for s in desc.shapes if s.type == 1)
    e2e_steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    e2e_parts = {"create_s": 0.0, "render_s": 0.0, "close_s": 0.0, "device_render_ms": 0.0}
    for _ in range(e2e_steps):
        ta = time.perf_counter()
        sc2 = lj.Scene(desc, device=local_rank)
        tb = time.perf_counter()
        img = sc2.render(spp=total_spp, sample_begin=s_begin, sample_end=s_end, normalize=True, pool_paths=args.pool,
                         tile_stride=tile_stride, tile_offset=tile_offset)
        tc = time.perf_counter()
        e2e_parts["device_render_ms"] += sc2.last_stats.render_ms / e2e_steps
        sc2.close()
        td = time.perf_counter()
        e2e_parts["create_s"] += (tb - ta) / e2e_steps; e2e_parts["render_s"] += (tc - tb) / e2e_steps; e2e_parts["close_s"] += (td - tc) / e2e_steps
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = npix * total_spp * e2e_steps / float(e2e_s.item()) / 1e6

    # L2-side bound: BVH nodes and primitives are L2-resident; measured L2 copy bandwidth of this GPU
    # (two 24 MB buffers, both inside the 126 MB L2) against the bytes the traversal kernels fetch, counted live.
    l2_peak = hbm_read = None
    if rank == 0:
        l2_peak = lj.measure_read_bandwidth(32 << 20, 200)    # 32 MB working set, L2-resident
        hbm_read = lj.measure_read_bandwidth(4 << 30, 3)      # 4 GB, streams from HBM
    if rank == 0:
        peak, peak_kind = load_peaks()
        value = npix * total_spp * args.steps / (ms_total / 1e3) / 1e6
        ext_bytes = BYTES_PER_EXTENSION_RAY * agg["closest"]
        achieved = ext_bytes / (agg["extend_ms"] / 1e3) / 1e9 if agg["extend_ms"] > 0 else 0.0
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(key, w, h, full_spp), "spp_per_gpu": (s_end - s_begin) if tile_stride == 0 else total_spp, "image_spp": total_spp,
                       "l2_policy": f"path pool ({agg.get('pool_paths', 0)} slots x {336 if key in VOLPATH else 144} B) streams through HBM every wave, larger than the 126 MB L2",
                       "scene_source": scene_source, "parallelism": f"{args.split}-split x{world} + NCCL reduce (NCCL_NVLS_ENABLE={os.environ.get('NCCL_NVLS_ENABLE', 'default')})" if world > 1 else "single GPU"},
            "mrays_per_s": (agg["closest"] + agg["shadow"]) * world / (ms_total / 1e3) / 1e6,
            "rays_per_sample": (agg["closest"] + agg["shadow"]) / max(agg["samples"], 1),
            "mean_bounces": agg["bounces"] / max(agg["samples"], 1),
            "node_steps_per_ray": agg["node_steps"] / max(agg["closest"] + agg["shadow"], 1),
            "prim_tests_per_ray": agg["prim_tests"] / max(agg["closest"] + agg["shadow"], 1),
            "simd_efficiency": {"node_step": agg["node_steps"] / max(32 * agg["node_passes"], 1), "prim_step": agg["prim_tests"] / max(32 * agg["prim_passes"], 1),
                                "note": "lanes doing the step / 32, per warp pass of the traversal kernels (counted by the kernels)"},
            "stage_ms_per_step": {k: agg[k] / args.steps for k in ("regen_ms", "extend_ms", "shade_ms", "shadow_ms", "render_ms")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(cdesc_bytes),
                    "d2h_bytes_per_step": int(npix * 3 * 4), "includes": "lj_scene_create (upload + GPU BVH/mip build) + lj_render + D2H", "parts": e2e_parts},
            "gpu_launches": int(agg["launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_trace_q<0> (closest-hit extension)" if key not in VOLPATH else "k_trace<0> (closest-hit extension)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES_PER_RAY[key] * agg["closest"] / max(agg["extend_launches"], 1) if key in NCU_TRAFFIC_BYTES_PER_RAY else None,
                         "traffic_unit": "bytes per launch: ncu dram read+write per ray (profiles/r02y_sponza_metrics.txt) x rays per launch of this run", "peak_kind": peak_kind,
                         "algorithmic_bytes_per_ray": BYTES_PER_EXTENSION_RAY,
                         "rays_per_launch": agg["closest"] / max(agg["extend_launches"], 1),
                         "avg_launch_ms": agg["extend_ms"] / max(agg["extend_launches"], 1)},
            "roofline_l2": {"bound": "l2", "kernel": "k_trace_q<0|1> (path) / k_trace<0|2|3> (volpath)", "unit": "GB/s", "peak": l2_peak,
                            "peak_kind": "measured here: read-only 16-byte-load stream over a 32 MB working set, all SMs (lj_measure_read_bandwidth)",
                            "hbm_read_gbs_same_probe": hbm_read,
                            "achieved": (agg["node_steps"] * 80 + agg["prim_tests"] * 48) / max((agg["extend_ms"] + agg["shadow_ms"]) / 1e3, 1e-9) / 1e9,
                            "frac": (agg["node_steps"] * 80 + agg["prim_tests"] * 48) / max((agg["extend_ms"] + agg["shadow_ms"]) / 1e3, 1e-9) / 1e9 / l2_peak,
                            "bytes": "80 B per wide-node step + 48 B per primitive test, both counted by the kernels"},
            "bvh": {"prims": info.num_prims, "nodes": info.num_bvh_nodes, "width": info.bvh_width, "build_ms": info.bvh_build_ms,
                    "sah": info.sah_cost},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                import oracle_lib
                cores = os.cpu_count() or 1
                ref = oracle_lib.RefScene(oracle_lib.scene_xml(key), threads=cores)
                cs = args.cpu_spp or cpu_spp
                devnull = os.open(os.devnull, os.O_WRONLY)
                saved = os.dup(1)
                os.dup2(devnull, 1)
                try:
                    _, secs, shim_share = oracle_lib.timed_render(ref, cs, cores)
                finally:
                    os.dup2(saved, 1)
                    os.close(devnull)
                line["cpu_baseline"] = {"value": npix * cs / secs / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "reference",
                                        "sample": f"{key} {w}x{h} at {cs} spp, unmodified lajolla sources + Embree-API shim (lajolla+shim)",
                                        "shim_share": shim_share}
            except Exception as e:  # the oracle is optional at bench time
                line["cpu_baseline"] = {"value": None, "unit": "Msamples/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
