"""ctypes access to oracle/_ref/libljoracle.so -- the CPU oracle (reference objects + Embree-API shim).

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg only.
"""
import atexit
import ctypes as C
import os

import numpy as np

import lajolla_public_b200 as lj
from lajolla_public_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_ref", "libljoracle.so")
SCENES = os.path.join(ROOT, "oracle", "_ref", "scenes")
LJS_DIR = os.path.join(ROOT, "oracle", "_ref", "ljs")

_lib = None


def available():
    return os.path.exists(ORACLE_SO)


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(ORACLE_SO)
        l.ljo_scene_load.restype = C.c_void_p
        l.ljo_scene_load.argtypes = [C.c_char_p, C.c_int]
        l.ljo_scene_free.argtypes = [C.c_void_p]
        l.ljo_set_spp.argtypes = [C.c_void_p, C.c_int]
        l.ljo_set_integrator.argtypes = [C.c_void_p, C.c_int]
        l.ljo_scene_dump.argtypes = [C.c_void_p, C.c_char_p]
        l.ljo_scene_info.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        l.ljo_light_table.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        l.ljo_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.ljo_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.ljo_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.ljo_bsdf.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.ljo_light.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.ljo_medium.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.ljo_num_media.restype = C.c_int
        l.ljo_num_media.argtypes = [C.c_void_p]
        l.ljo_camera_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.ljo_texture.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        l.ljo_mip_level.restype = C.c_int
        l.ljo_mip_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
        l.ljo_pcg32.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        l.ljo_render.restype = C.c_double
        l.ljo_render.argtypes = [C.c_void_p, C.c_void_p]
        l.ljo_film_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        l.ljshim_get_counters.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
        l.ljshim_get_ticks.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
        atexit.register(l.ljo_shutdown)
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefScene:
    """The reference's own Scene, built by the reference's own parse_scene()."""

    def __init__(self, xml_path, threads=0):
        threads = threads or os.cpu_count() or 1
        self.h = lib().ljo_scene_load(os.fsencode(xml_path), threads)
        if not self.h:
            raise RuntimeError(f"reference parse_scene failed for {xml_path}")

    def close(self):
        if self.h:
            lib().ljo_scene_free(self.h)
            self.h = None

    def dump(self, out_path):
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        if lib().ljo_scene_dump(self.h, os.fsencode(out_path)) != 0:
            raise RuntimeError("dump failed")

    def set_integrator(self, integrator):
        lib().ljo_set_integrator(self.h, integrator)

    def set_spp(self, spp):
        lib().ljo_set_spp(self.h, spp)

    def info(self):
        b = (C.c_double * 4)()
        e = C.c_double()
        c = (C.c_int * 3)()
        lib().ljo_scene_info(self.h, b, C.byref(e), c)
        return dict(radius=b[0], center=(b[1], b[2], b[3]), eps=e.value, shapes=c[0], lights=c[1], materials=c[2])

    def light_table(self):
        n = self.info()["lights"]
        pmf = (C.c_double * n)()
        cdf = (C.c_double * (n + 1))()
        lib().ljo_light_table(self.h, pmf, cdf)
        return np.array(pmf), np.array(cdf)

    def intersect_hits(self, rays):
        rays = np.ascontiguousarray(rays, dtype=lj.RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=lj.HIT_DTYPE)
        lib().ljo_trace_closest(self.h, _p(rays), rays.shape[0], _p(hits))
        return hits

    def occluded(self, rays):
        rays = np.ascontiguousarray(rays, dtype=lj.RAY_DTYPE)
        occ = np.zeros(rays.shape[0], dtype=np.uint8)
        lib().ljo_trace_any(self.h, _p(rays), rays.shape[0], _p(occ))
        return occ.astype(bool)

    def intersect(self, rays, ray_diff=None):
        rays = np.ascontiguousarray(rays, dtype=lj.RAY_DTYPE)
        out = np.zeros(rays.shape[0], dtype=lj.VERTEX_DTYPE)
        rd = None if ray_diff is None else np.ascontiguousarray(ray_diff, dtype=np.float32)
        lib().ljo_intersect(self.h, _p(rays), _p(rd) if rd is not None else None, rays.shape[0], _p(out))
        return out

    def bsdf(self, queries):
        q = np.ascontiguousarray(queries, dtype=lj.BSDF_QUERY_DTYPE)
        out = np.zeros(q.shape[0], dtype=lj.BSDF_RESULT_DTYPE)
        lib().ljo_bsdf(self.h, _p(q), q.shape[0], _p(out))
        return out

    def sample_lights(self, queries):
        q = np.ascontiguousarray(queries, dtype=lj.LIGHT_QUERY_DTYPE)
        out = np.zeros(q.shape[0], dtype=lj.LIGHT_RESULT_DTYPE)
        lib().ljo_light(self.h, _p(q), q.shape[0], _p(out))
        return out

    def num_media(self):
        return lib().ljo_num_media(self.h)

    def medium(self, queries):
        q = np.ascontiguousarray(queries, dtype=lj.MEDIUM_QUERY_DTYPE)
        out = np.zeros(q.shape[0], dtype=lj.MEDIUM_RESULT_DTYPE)
        lib().ljo_medium(self.h, _p(q), q.shape[0], _p(out))
        return out

    def sample_primary(self, screen_pos):
        xy = np.ascontiguousarray(screen_pos, dtype=np.float32).reshape(-1, 2)
        rays = np.zeros(xy.shape[0], dtype=lj.RAY_DTYPE)
        lib().ljo_camera_rays(self.h, _p(xy), xy.shape[0], _p(rays))
        return rays

    def eval_texture(self, material_id, uv_footprint):
        q = np.ascontiguousarray(uv_footprint, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((q.shape[0], 3), dtype=np.float32)
        lib().ljo_texture(self.h, material_id, _p(q), q.shape[0], _p(out))
        return out

    def mip_level(self, image3_id, level):
        w, h = C.c_int(0), C.c_int(0)
        n = lib().ljo_mip_level(self.h, image3_id, level, C.byref(w), C.byref(h), None)
        if level >= n:
            return None
        data = np.zeros((h.value, w.value, 3), dtype=np.float32)
        lib().ljo_mip_level(self.h, image3_id, level, C.byref(w), C.byref(h), _p(data))
        return data

    def render(self, spp=None):
        """The reference's render() (render.cpp:155) -> ((h,w,3) float32, seconds inside render())."""
        if spp is not None:
            self.set_spp(spp)
        w, h, s = C.c_int(0), C.c_int(0), C.c_int(0)
        lib().ljo_film_size(self.h, C.byref(w), C.byref(h), C.byref(s))
        out = np.zeros((h.value, w.value, 3), dtype=np.float32)
        secs = lib().ljo_render(self.h, _p(out))
        return out, secs


def pcg32(first_stream, n_streams, n_draws, seed=0x31e241f862a1fb5e):
    u = np.zeros((n_streams, n_draws), dtype=np.uint32)
    f = np.zeros((n_streams, n_draws), dtype=np.float64)
    lib().ljo_pcg32(first_stream, seed, n_streams, n_draws, _p(u), _p(f))
    return u, f


def ray_counters():
    a, b = C.c_ulonglong(0), C.c_ulonglong(0)
    lib().ljshim_get_counters(C.byref(a), C.byref(b))
    return a.value, b.value


def shim_ticks():
    """(time-stamp ticks spent inside rtcIntersect1 / rtcOccluded1 summed over threads, the counter now)."""
    a, b = C.c_ulonglong(0), C.c_ulonglong(0)
    lib().ljshim_get_ticks(C.byref(a), C.byref(b))
    return a.value, b.value


def timed_render(ref, spp, threads):
    """ref.render(spp) plus the share of the CPU time that went into the Embree-API shim's ray casts:
    ticks inside the shim (all threads) / (wall ticks * threads)."""
    i0, n0 = shim_ticks()
    img, secs = ref.render(spp=spp)
    i1, n1 = shim_ticks()
    share = (i1 - i0) / max((n1 - n0) * max(threads, 1), 1)
    return img, secs, min(share, 1.0)


# scene name -> xml path relative to SCENES
SCENE_XML = {
    "cbox": "cbox/cbox.xml",
    "veach_mi": "veach_mi/mi.xml",
    "sponza": "sponza/sponza.xml",
    "matpreview": "matpreview/matpreview.xml",
    "pixel_filter_test": "pixel_filter_test/pixel_filter_test.xml",
    "pixel_filter_box": "pixel_filter_test/pixel_filter_box.xml",    # generated by oracle/Makefile (box / tent rfilter)
    "pixel_filter_tent": "pixel_filter_test/pixel_filter_tent.xml",
    "disney_bsdf": "disney_bsdf_test/disney_bsdf.xml",
    "disney_diffuse": "disney_bsdf_test/disney_diffuse.xml",
    "disney_metal": "disney_bsdf_test/disney_metal.xml",
    "disney_glass": "disney_bsdf_test/disney_glass.xml",
    "disney_clearcoat": "disney_bsdf_test/disney_clearcoat.xml",
    "disney_sheen": "disney_bsdf_test/disney_sheen.xml",
    "disney_bsdf_array": "disney_bsdf_test/disney_bsdf_array.xml",
    "simple_sphere": "disney_bsdf_test/simple_sphere.xml",
    "volpath_test1": "volpath_test/volpath_test1.xml",
    "volpath_test2": "volpath_test/volpath_test2.xml",
    "volpath_test3": "volpath_test/volpath_test3.xml",
    "volpath_test4": "volpath_test/volpath_test4.xml",
    "volpath_test4_2": "volpath_test/volpath_test4_2.xml",
    "volpath_test5": "volpath_test/volpath_test5.xml",
    "volpath_test5_2": "volpath_test/volpath_test5_2.xml",
    "volpath_test6": "volpath_test/volpath_test6.xml",
    "vol_cbox": "volpath_test/vol_cbox.xml",
    "vol_cbox_teapot": "volpath_test/vol_cbox_teapot.xml",
    "hetvol": "volpath_test/hetvol.xml",
    "hetvol_colored": "volpath_test/hetvol_colored.xml",
}


HANDOUTS = os.path.join(ROOT, "oracle", "_ref", "handouts")

# scene name -> the handout's render of it (handouts/imgs/*.png, copied to oracle/_ref/handouts by oracle/Makefile)
HANDOUT_IMAGE = {
    "cbox": "cbox.png", "veach_mi": "veach_mis.png", "sponza": "sponza.png", "matpreview": "matpreview.png",
    "pixel_filter_test": "gaussian.png", "pixel_filter_box": "box.png", "pixel_filter_tent": "tent.png",
    "disney_diffuse": "disney_diffuse.png", "disney_metal": "disney_metal.png", "disney_clearcoat": "disney_clearcoat.png",
    "disney_glass": "disney_glass.png", "disney_sheen": "disney_sheen.png", "disney_bsdf_array": "disney_bsdf.png",
    "volpath_test1": "volpath_1.png", "volpath_test2": "volpath_2.png", "volpath_test3": "volpath_3.png",
    "volpath_test4": "volpath_4.png", "volpath_test4_2": "volpath_4_2.png", "volpath_test5": "volpath_5.png",
    "volpath_test5_2": "volpath_5_2.png", "volpath_test6": "volpath_6.png", "vol_cbox": "volpath_5_cbox.png",
    "vol_cbox_teapot": "volpath_5_cbox_teapot.png", "hetvol": "hetvol.png", "hetvol_colored": "colored_smoke.png",
}


def handout_image(name):
    """The handout's LDR render of a scene as float sRGB in [0, 1], (h, w, 3)."""
    from PIL import Image
    return np.asarray(Image.open(os.path.join(HANDOUTS, HANDOUT_IMAGE[name])).convert("RGB"), dtype=np.float64) / 255.0


def scene_xml(name):
    return os.path.join(SCENES, SCENE_XML[name])


def scene_ljs(name):
    """Path of the oracle-dumped flat scene (created on demand from the reference's own parser)."""
    out = os.path.join(LJS_DIR, name + ".ljs")
    if not os.path.exists(out):
        s = RefScene(scene_xml(name), threads=1)
        s.dump(out)
        s.close()
    return out
