"""Golden vectors: answers of the reference's own object code (oracle/_ref/libljoracle.so) to fixed batches of
queries, recorded once by tests/golden/make_golden.py and committed under tests/golden/*.npz.

`RecordingRef` wraps an oracle_lib.RefScene and stores (method, query bytes) -> answer for every call the parity
checks make; `GoldenScene` replays those answers with the same interface, so tests/parity_checks.py runs unchanged
against either the live oracle or the committed fixtures (which need neither /root/reference nor the oracle build).
"""
import hashlib
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ARRAY_METHODS = ("intersect_hits", "occluded", "intersect", "bsdf", "sample_lights", "sample_primary", "medium", "eval_texture")


def _key(method, *arrays):
    h = hashlib.sha1(method.encode())
    for a in arrays:
        if a is None:
            h.update(b"none")
        elif isinstance(a, (int, np.integer)):
            h.update(str(int(a)).encode())
        else:
            a = np.ascontiguousarray(a)
            h.update(str(a.dtype.itemsize).encode() + str(a.shape).encode() + a.tobytes())
    return method + "_" + h.hexdigest()[:16]


def _norm(method, args):
    """Bring query arguments to the dtype the C ABI sees, so recorder and replayer hash the same bytes."""
    import lajolla_public_b200 as lj
    dt = {"intersect_hits": lj.RAY_DTYPE, "occluded": lj.RAY_DTYPE, "intersect": lj.RAY_DTYPE, "bsdf": lj.BSDF_QUERY_DTYPE,
          "sample_lights": lj.LIGHT_QUERY_DTYPE, "medium": lj.MEDIUM_QUERY_DTYPE}
    out = list(args)
    if method in dt:
        out[0] = np.ascontiguousarray(out[0], dtype=dt[method])
        if method == "intersect" and len(out) > 1 and out[1] is not None:
            out[1] = np.ascontiguousarray(out[1], dtype=np.float32).reshape(-1, 2)
    elif method == "sample_primary":
        out[0] = np.ascontiguousarray(out[0], dtype=np.float32).reshape(-1, 2)
    elif method == "eval_texture":
        out[1] = np.ascontiguousarray(out[1], dtype=np.float32).reshape(-1, 3)
    return out


class RecordingRef:
    def __init__(self, ref):
        self.ref = ref
        self.store = {}
        self.meta = {"info": ref.info(), "num_media": ref.num_media()}
        pmf, cdf = ref.light_table()
        self.store["light_pmf"], self.store["light_cdf"] = pmf, cdf

    def __getattr__(self, name):
        if name not in ARRAY_METHODS:
            raise AttributeError(name)

        def call(*args):
            args = _norm(name, args)
            res = getattr(self.ref, name)(*args)
            self.store[_key(name, *args)] = np.asarray(res)
            return res
        return call

    def info(self):
        return self.meta["info"]

    def num_media(self):
        return self.meta["num_media"]

    def light_table(self):
        return self.store["light_pmf"], self.store["light_cdf"]

    def save(self, name):
        os.makedirs(GOLDEN_DIR, exist_ok=True)
        arrays = dict(self.store)
        arrays["meta_json"] = np.frombuffer(json.dumps(self.meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **arrays)


class GoldenScene:
    """Replays tests/golden/<name>.npz with the RefScene interface."""

    def __init__(self, name):
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        self.z = np.load(path)
        self.meta = json.loads(bytes(self.z["meta_json"]).decode())
        self.calls = 0

    def __getattr__(self, name):
        if name not in ARRAY_METHODS:
            raise AttributeError(name)

        def call(*args):
            k = _key(name, *_norm(name, args))
            if k not in self.z:
                raise KeyError(f"no golden answer recorded for this {name} batch ({k}); regenerate with tests/golden/make_golden.py")
            self.calls += 1
            res = self.z[k]
            return res.astype(bool) if name == "occluded" else res
        return call

    def keys(self, method):
        return [k for k in self.z.files if k.startswith(method + "_")]

    def info(self):
        i = dict(self.meta["info"])
        i["center"] = tuple(i["center"])
        return i

    def num_media(self):
        return self.meta["num_media"]

    def light_table(self):
        return self.z["light_pmf"], self.z["light_cdf"]


def run_checks(sc, ref, name, n=384):
    """The batch of parity checks whose oracle answers the fixtures hold (same calls for recording and replay)."""
    import parity_checks as pc
    out = {}
    pc.check_scene_info(sc, ref)
    pc.check_light_table(sc, ref)
    rays = pc.primary_rays(ref, n)
    out["primary"] = pc.check_ray_parity(sc, ref, rays, min_agree=0.997)  # 1 ray of 384 may sit on an edge tie
    out["bounce"] = pc.check_ray_parity(sc, ref, pc.bounce_rays(ref, rays), min_agree=0.997)
    out["occlusion"] = pc.check_occlusion_parity(sc, ref, pc.shadow_rays(ref, rays), min_agree=0.997)
    rd = np.tile(np.array([0.0, 0.25 / 768], dtype=np.float32), (rays.shape[0], 1))
    pc.check_vertex_parity(sc, ref, rays, rd)
    pc.check_camera_parity(sc, ref, n=512)
    v = ref.intersect(rays)
    pc.check_light_parity(sc, ref, v["position"][v["shape_id"] >= 0])
    out["bsdf"] = pc.check_bsdf_parity(sc, ref, pc.make_bsdf_queries(ref, rays))
    pc.check_bsdf_parity(sc, ref, pc.fixed_material_queries(ref.info()["materials"]))
    if ref.num_media() > 0:
        out["media"] = pc.check_medium_parity(sc, ref, n=512)
    return out
