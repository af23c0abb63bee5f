"""lajolla_public_b200 -- host-side mirror of lajolla's rendering interface over the CUDA C ABI.

Names follow the reference (BachiLi/lajolla_public): `parse_scene` (parsers/parse_scene.h:9),
`render` (render.h:9), `intersect` / `occluded` (intersection.h:39-46), `eval` / `pdf_sample_bsdf` /
`sample_bsdf` (material.h:119-163), `sample_primary` (camera.h:27-28).  Everything executes in
libljb200.so on a B200 (include/lajolla_b200.h); this package only marshals buffers.  There is no
CPU implementation behind it: without the built library or without a CUDA device calls raise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import abi, ljs
from .abi import LajollaError, load_library

__all__ = ["Scene", "parse_scene", "load_scene_description", "render", "LajollaError", "load_library", "abi", "ljs"]

RAY_DTYPE = np.dtype([("org", "<f4", 3), ("tnear", "<f4"), ("dir", "<f4", 3), ("tfar", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("shape_id", "<i4"), ("primitive_id", "<i4")])
VERTEX_DTYPE = np.dtype([("position", "<f4", 3), ("geometric_normal", "<f4", 3), ("frame_x", "<f4", 3),
                         ("frame_y", "<f4", 3), ("frame_n", "<f4", 3), ("st", "<f4", 2), ("uv", "<f4", 2),
                         ("uv_screen_size", "<f4"), ("mean_curvature", "<f4"), ("ray_radius", "<f4"),
                         ("shape_id", "<i4"), ("primitive_id", "<i4"), ("material_id", "<i4"),
                         ("interior_medium_id", "<i4"), ("exterior_medium_id", "<i4")])
BSDF_QUERY_DTYPE = np.dtype([("vertex", VERTEX_DTYPE), ("dir_in", "<f4", 3), ("dir_out", "<f4", 3),
                             ("rnd_uv", "<f4", 2), ("rnd_w", "<f4"), ("transport", "<i4")])
BSDF_RESULT_DTYPE = np.dtype([("f", "<f4", 3), ("pdf", "<f4"), ("sampled", "<i4"), ("s_dir_out", "<f4", 3),
                              ("s_eta", "<f4"), ("s_roughness", "<f4")])
LIGHT_QUERY_DTYPE = np.dtype([("ref_point", "<f4", 3), ("rnd_uv", "<f4", 2), ("rnd_w", "<f4"), ("light_w", "<f4")])
LIGHT_RESULT_DTYPE = np.dtype([("light_id", "<i4"), ("position", "<f4", 3), ("normal", "<f4", 3), ("pmf", "<f4"),
                               ("pdf", "<f4"), ("emission", "<f4", 3)])
MEDIUM_QUERY_DTYPE = np.dtype([("org", "<f4", 3), ("tfar", "<f4"), ("dir", "<f4", 3), ("t", "<f4"), ("rnd", "<f4", 2),
                               ("medium_id", "<i4"), ("_pad", "<i4")])
WALK_QUERY_DTYPE = np.dtype([("origin", "<f4", 3), ("medium_id", "<i4"), ("light_point", "<f4", 3), ("seed", "<u4"), ("c", "<f4", 3),
                             ("pdf_nee", "<f4"), ("pdf_dir", "<f4"), ("budget", "<i4"), ("_pad", "<i4", 2)])
MEDIUM_RESULT_DTYPE = np.dtype([("majorant", "<f4", 3), ("sigma_a", "<f4", 3), ("sigma_s", "<f4", 3), ("phase_dir", "<f4", 3),
                                ("phase_eval", "<f4"), ("phase_pdf", "<f4")])

assert RAY_DTYPE.itemsize == C.sizeof(abi.lj_ray) and HIT_DTYPE.itemsize == C.sizeof(abi.lj_hit)
assert VERTEX_DTYPE.itemsize == C.sizeof(abi.lj_vertex)
assert BSDF_QUERY_DTYPE.itemsize == C.sizeof(abi.lj_bsdf_query)
assert BSDF_RESULT_DTYPE.itemsize == C.sizeof(abi.lj_bsdf_result)
assert LIGHT_QUERY_DTYPE.itemsize == C.sizeof(abi.lj_light_query)
assert LIGHT_RESULT_DTYPE.itemsize == C.sizeof(abi.lj_light_result)
assert MEDIUM_QUERY_DTYPE.itemsize == C.sizeof(abi.lj_medium_query)
assert MEDIUM_RESULT_DTYPE.itemsize == C.sizeof(abi.lj_medium_result)
MEDIUM_BOUND_DTYPE = np.dtype([("majorant", "<f4", 3), ("t_exit", "<f4"), ("sigma_t", "<f4", 3), ("local", "<i4")])
assert MEDIUM_BOUND_DTYPE.itemsize == C.sizeof(abi.lj_medium_bound)
assert WALK_QUERY_DTYPE.itemsize == C.sizeof(abi.lj_walk_query)


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def make_rays(org, dir, tnear=0.0, tfar=np.inf):
    org = np.asarray(org, dtype=np.float32).reshape(-1, 3)
    r = np.zeros(org.shape[0], dtype=RAY_DTYPE)
    r["org"] = org
    r["dir"] = np.asarray(dir, dtype=np.float32).reshape(-1, 3)
    r["tnear"] = tnear
    r["tfar"] = tfar
    return r


class Scene:
    """Device-resident scene (the reference's `Scene`, scene.h:39-88): BVH, mip chains and sampling
    tables are built on the GPU by lj_scene_create."""

    def __init__(self, desc: ljs.SceneDesc, device=0):
        """device: one CUDA device index, or a list -- the first is the primary device, the scene is replicated on the
        others and render(num_gpus=...) splits the work among them inside the library."""
        self._lib = load_library()
        ids = [int(device)] if np.isscalar(device) else [int(d) for d in device]
        self.devices = ids
        arr = (C.c_int32 * len(ids))(*ids)
        abi.check(self._lib.lj_init(arr, len(ids)))
        cdesc, keep = ljs.to_c(desc)
        h = C.c_void_p()
        abi.check(self._lib.lj_scene_create(C.byref(cdesc), C.byref(h)))
        del keep
        self._h = h
        self.desc = desc
        self.width, self.height = desc.camera.width, desc.camera.height
        self.last_stats = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lj_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- S1
    def render(self, spp=0, sample_begin=0, sample_end=0, normalize=True, pool_paths=0, seed=0, variance=False,
               tile_stride=0, tile_offset=0, num_gpus=1, split=0, reduce=0):
        opts = abi.lj_render_opts(spp, sample_begin, sample_end, 1 if normalize else 0, pool_paths, seed, None,
                                  tile_stride, tile_offset, num_gpus, split, reduce, 0)
        out = np.empty((self.height, self.width, 3), dtype=np.float32)
        var = None
        if variance:
            var = np.empty((self.height, self.width, 3), dtype=np.float32)
            opts.variance_out = _ptr(var, C.c_float)
        stats = abi.lj_stats()
        abi.check(self._lib.lj_render(self._h, C.byref(opts), _ptr(out, C.c_float), C.byref(stats)))
        self.last_stats = stats
        return (out, var) if variance else out

    def render_device(self, d_out_ptr, stream_ptr=None, spp=0, sample_begin=0, sample_end=0, normalize=False, pool_paths=0, seed=0,
                      tile_stride=0, tile_offset=0):
        """Render into DEVICE memory (w*h*3 fp32 at d_out_ptr) on the given cudaStream_t."""
        opts = abi.lj_render_opts(spp, sample_begin, sample_end, 1 if normalize else 0, pool_paths, seed, None,
                                  tile_stride, tile_offset, 1, 0, 0, 0)
        stats = abi.lj_stats()
        abi.check(self._lib.lj_render_device(self._h, C.byref(opts), C.c_void_p(d_out_ptr), C.c_void_p(stream_ptr or 0), C.byref(stats)))
        self.last_stats = stats
        return stats

    # ---- S2
    def intersect_hits(self, rays, want_ms=False, kernel=abi.LJ_TRACE_PLAIN, pool_paths=0, slot_stride=1):
        """rtcIntersect1-equivalent for a batch.  kernel: abi.LJ_TRACE_PLAIN (one thread per ray), LJ_TRACE_WAVEFRONT /
        LJ_TRACE_WAVEFRONT_LANE (the persistent kernels of the renderer, rays loaded into the path pool)."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        ms = C.c_double(0)
        opts = abi.lj_trace_opts(kernel, pool_paths, slot_stride, 0)
        abi.check(self._lib.lj_trace_closest_ex(self._h, _ptr(rays, abi.lj_ray), rays.shape[0], C.byref(opts), _ptr(hits, abi.lj_hit), C.byref(ms)))
        return (hits, ms.value) if want_ms else hits

    def occluded(self, rays, want_ms=False, kernel=abi.LJ_TRACE_PLAIN, pool_paths=0, slot_stride=1):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        occ = np.zeros(rays.shape[0], dtype=np.uint8)
        ms = C.c_double(0)
        opts = abi.lj_trace_opts(kernel, pool_paths, slot_stride, 0)
        abi.check(self._lib.lj_trace_any_ex(self._h, _ptr(rays, abi.lj_ray), rays.shape[0], C.byref(opts), _ptr(occ, C.c_uint8), C.byref(ms)))
        return (occ.astype(bool), ms.value) if want_ms else occ.astype(bool)

    def intersect(self, rays, ray_diff=None):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        out = np.zeros(rays.shape[0], dtype=VERTEX_DTYPE)
        rd = None
        if ray_diff is not None:
            rd = np.ascontiguousarray(ray_diff, dtype=np.float32).reshape(-1, 2)
        abi.check(self._lib.lj_intersect(self._h, _ptr(rays, abi.lj_ray), _ptr(rd, C.c_float) if rd is not None else None,
                                         rays.shape[0], _ptr(out, abi.lj_vertex)))
        return out

    def bsdf(self, queries):
        q = np.ascontiguousarray(queries, dtype=BSDF_QUERY_DTYPE)
        out = np.zeros(q.shape[0], dtype=BSDF_RESULT_DTYPE)
        abi.check(self._lib.lj_bsdf_batch(self._h, _ptr(q, abi.lj_bsdf_query), q.shape[0], _ptr(out, abi.lj_bsdf_result)))
        return out

    def sample_lights(self, queries):
        q = np.ascontiguousarray(queries, dtype=LIGHT_QUERY_DTYPE)
        out = np.zeros(q.shape[0], dtype=LIGHT_RESULT_DTYPE)
        abi.check(self._lib.lj_light_batch(self._h, _ptr(q, abi.lj_light_query), q.shape[0], _ptr(out, abi.lj_light_result)))
        return out

    def medium(self, queries):
        """get_majorant / get_sigma_a / get_sigma_s (medium.h:25-27) and the phase function (phase_function.h:18-29)."""
        q = np.ascontiguousarray(queries, dtype=MEDIUM_QUERY_DTYPE)
        out = np.zeros(q.shape[0], dtype=MEDIUM_RESULT_DTYPE)
        abi.check(self._lib.lj_medium_batch(self._h, _ptr(q, abi.lj_medium_query), q.shape[0], _ptr(out, abi.lj_medium_result)))
        return out

    def medium_bounds(self, queries):
        """The bound the tracking loops use at org + t * dir (block-wise majorant of a grid medium), where it stops
        holding, and sigma_t at the point."""
        q = np.ascontiguousarray(queries, dtype=MEDIUM_QUERY_DTYPE)
        out = np.zeros(q.shape[0], dtype=MEDIUM_BOUND_DTYPE)
        abi.check(self._lib.lj_medium_bound_batch(self._h, _ptr(q, abi.lj_medium_query), q.shape[0], _ptr(out, abi.lj_medium_bound)))
        return out

    def nee_walks(self, queries, kernel=abi.LJ_TRACE_PLAIN, pool_paths=0, slot_stride=1, want_ms=False, walk_rounds=0):
        """The volpath integrator's NEE walk for a batch of (origin, light point) queries: (n, 3) contributions."""
        q = np.ascontiguousarray(queries, dtype=WALK_QUERY_DTYPE)
        out = np.zeros((q.shape[0], 3), dtype=np.float32)
        ms = C.c_double(0)
        opts = abi.lj_trace_opts(kernel, pool_paths, slot_stride, walk_rounds)
        abi.check(self._lib.lj_nee_walk_batch(self._h, _ptr(q, abi.lj_walk_query), q.shape[0], C.byref(opts), _ptr(out, C.c_float), C.byref(ms)))
        return (out, ms.value) if want_ms else out

    def sample_primary(self, screen_pos):
        xy = np.ascontiguousarray(screen_pos, dtype=np.float32).reshape(-1, 2)
        rays = np.zeros(xy.shape[0], dtype=RAY_DTYPE)
        abi.check(self._lib.lj_camera_rays(self._h, _ptr(xy, C.c_float), xy.shape[0], _ptr(rays, abi.lj_ray)))
        return rays

    def eval_texture(self, material_id, slot, uv_footprint):
        q = np.ascontiguousarray(uv_footprint, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((q.shape[0], 3), dtype=np.float32)
        abi.check(self._lib.lj_texture_batch(self._h, material_id, slot, _ptr(q, C.c_float), q.shape[0], _ptr(out, C.c_float)))
        return out

    def info(self):
        i = abi.lj_scene_info()
        abi.check(self._lib.lj_scene_get_info(self._h, C.byref(i)))
        return i

    def light_table(self):
        n = len(self.desc.lights)
        pmf = np.zeros(n, dtype=np.float32)
        cdf = np.zeros(n + 1, dtype=np.float32)
        abi.check(self._lib.lj_scene_get_light_table(self._h, _ptr(pmf, C.c_float), _ptr(cdf, C.c_float)))
        return pmf, cdf

    def mip_level(self, image_id, level):
        ch = self.desc.images[image_id].shape[2]
        w, h = C.c_int32(0), C.c_int32(0)
        abi.check(self._lib.lj_scene_get_mip_level(self._h, ch, image_id, level, C.byref(w), C.byref(h), None))
        data = np.zeros((h.value, w.value, ch), dtype=np.float32)
        abi.check(self._lib.lj_scene_get_mip_level(self._h, ch, image_id, level, C.byref(w), C.byref(h), _ptr(data, C.c_float)))
        return data


def pcg32(first_stream, n_streams, n_draws, seed=0):
    lib = load_library()
    u = np.zeros((n_streams, n_draws), dtype=np.uint32)
    f = np.zeros((n_streams, n_draws), dtype=np.float32)
    abi.check(lib.lj_pcg32_batch(first_stream, seed, n_streams, n_draws, _ptr(u, C.c_uint32), _ptr(f, C.c_float)))
    return u, f


def measure_read_bandwidth(nbytes, iters=20):
    """GB/s of a read-only 16-byte-load stream over `nbytes` (<= L2: the L2 read rate; >> L2: HBM)."""
    lib = load_library()
    v = C.c_double(0)
    abi.check(lib.lj_measure_read_bandwidth(int(nbytes), int(iters), C.byref(v)))
    return v.value


LAJOLLA_CLI = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lajolla")


def load_scene_description(path) -> ljs.SceneDesc:
    """The flat description of a scene file: `.ljs` containers are read directly, Mitsuba-style XML goes through the
    host front end (`lajolla --dump-ljs`, lajolla_public_b200/host: XML + OBJ / serialized / PLY / image / volume
    loaders in C++).  Needs no GPU."""
    path = str(path)
    if path.endswith(".ljs"):
        return ljs.load(path)
    if not os.path.exists(LAJOLLA_CLI):
        raise ImportError(f"{LAJOLLA_CLI} not built (make -C lajolla_public_b200)")
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "scene.ljs")
        r = subprocess.run([LAJOLLA_CLI, "--dump-ljs", out, path], capture_output=True, text=True)
        if r.returncode != 0:
            raise LajollaError(abi.LJ_ERR_INVALID, r.stderr.strip() or "scene parse failed")
        return ljs.load(out)


def parse_scene(path, device=0) -> Scene:
    """parse_scene() of the reference (parsers/parse_scene.cpp:1602): scene file -> device-resident Scene."""
    return Scene(load_scene_description(path), device)


def render(scene: Scene, **kw):
    """Image3 render(const Scene&) (render.cpp:155): returns (h, w, 3) float32."""
    return scene.render(**kw)
