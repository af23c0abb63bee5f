"""ctypes mirror of include/lajolla_b200.h (the C ABI of libljb200.so).

Struct layouts must match the header field for field; tests/test_abi.py checks sizes and that the
library exports every declared symbol.  The library is CUDA-only: loading it does not need a GPU,
calling it does (LJ_ERR_NO_DEVICE otherwise) -- there is no CPU fallback.
"""
import ctypes as C
import os

LJ_OK, LJ_ERR_INVALID, LJ_ERR_CUDA, LJ_ERR_NO_DEVICE, LJ_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
LJ_NUM_TEX_SLOTS = 12

f32, i32, u8, u32, u64, i64, f64 = C.c_float, C.c_int32, C.c_uint8, C.c_uint32, C.c_uint64, C.c_int64, C.c_double
pf32 = C.POINTER(f32)
pi32 = C.POINTER(i32)


class lj_image_desc(C.Structure):
    _fields_ = [("width", i32), ("height", i32), ("channels", i32), ("_pad", i32), ("data", pf32)]


class lj_texture_desc(C.Structure):
    _fields_ = [("kind", i32), ("image_id", i32), ("value", f32 * 3), ("color1", f32 * 3),
                ("uscale", f32), ("vscale", f32), ("uoffset", f32), ("voffset", f32)]


class lj_material_desc(C.Structure):
    _fields_ = [("type", i32), ("eta", f32), ("tex", lj_texture_desc * LJ_NUM_TEX_SLOTS)]


class lj_shape_desc(C.Structure):
    _fields_ = [("type", i32), ("material_id", i32), ("area_light_id", i32), ("interior_medium_id", i32),
                ("exterior_medium_id", i32), ("center", f32 * 3), ("radius", f32),
                ("num_vertices", i32), ("num_triangles", i32),
                ("positions", pf32), ("indices", pi32), ("normals", pf32), ("uvs", pf32)]


class lj_light_desc(C.Structure):
    _fields_ = [("type", i32), ("shape_id", i32), ("intensity", f32 * 3), ("values", lj_texture_desc),
                ("to_world", f32 * 16), ("to_local", f32 * 16), ("scale", f32)]


class lj_volume_desc(C.Structure):
    _fields_ = [("is_grid", i32), ("res", i32 * 3), ("value", f32 * 3), ("p_min", f32 * 3), ("p_max", f32 * 3),
                ("scale", f32), ("data", pf32)]


class lj_medium_desc(C.Structure):
    _fields_ = [("type", i32), ("phase_type", i32), ("phase_g", f32), ("sigma_a", f32 * 3), ("sigma_s", f32 * 3),
                ("albedo", lj_volume_desc), ("density", lj_volume_desc)]


class lj_camera_desc(C.Structure):
    _fields_ = [("cam_to_world", f32 * 16), ("world_to_cam", f32 * 16), ("sample_to_cam", f32 * 16),
                ("cam_to_sample", f32 * 16), ("width", i32), ("height", i32), ("filter_type", i32),
                ("filter_param", f32), ("medium_id", i32)]


class lj_options_desc(C.Structure):
    _fields_ = [("integrator", i32), ("samples_per_pixel", i32), ("max_depth", i32), ("rr_depth", i32),
                ("vol_path_version", i32), ("max_null_collisions", i32)]


class lj_scene_desc(C.Structure):
    _fields_ = [("camera", lj_camera_desc), ("options", lj_options_desc),
                ("num_images", i32), ("num_materials", i32), ("num_shapes", i32), ("num_lights", i32),
                ("num_media", i32), ("envmap_light_id", i32),
                ("images", C.POINTER(lj_image_desc)), ("materials", C.POINTER(lj_material_desc)),
                ("shapes", C.POINTER(lj_shape_desc)), ("lights", C.POINTER(lj_light_desc)),
                ("media", C.POINTER(lj_medium_desc))]


class lj_render_opts(C.Structure):
    _fields_ = [("spp", i32), ("sample_begin", i32), ("sample_end", i32), ("normalize", i32), ("pool_paths", i32),
                ("seed", u64), ("variance_out", pf32), ("tile_stride", i32), ("tile_offset", i32),
                ("num_gpus", i32), ("split", i32), ("reduce", i32), ("_pad", i32)]


class lj_stats(C.Structure):
    _fields_ = [("render_ms", f64), ("extend_ms", f64), ("shadow_ms", f64), ("shade_ms", f64), ("regen_ms", f64),
                ("samples", u64), ("closest_rays", u64), ("shadow_rays", u64), ("bounces", u64),
                ("kernel_launches", u64), ("waves", u64),
                ("extend_launches", u64), ("shadow_launches", u64), ("shade_launches", u64), ("regen_launches", u64),
                ("node_steps", u64), ("prim_tests", u64), ("node_passes", u64), ("prim_passes", u64),
                ("pool_paths", u64), ("gpus_used", i32), ("_pad", i32), ("reduce_ms", f64)]


class lj_ray(C.Structure):
    _fields_ = [("org", f32 * 3), ("tnear", f32), ("dir", f32 * 3), ("tfar", f32)]


class lj_trace_opts(C.Structure):
    _fields_ = [("kernel", i32), ("pool_paths", i32), ("slot_stride", i32), ("walk_rounds", i32)]


LJ_TRACE_PLAIN, LJ_TRACE_WAVEFRONT, LJ_TRACE_WAVEFRONT_LANE, LJ_TRACE_WALK_WHOLE, LJ_TRACE_WALK_STEP, LJ_TRACE_WALK_STAGED = 0, 1, 2, 3, 4, 5


class lj_walk_query(C.Structure):
    _fields_ = [("origin", f32 * 3), ("medium_id", i32), ("light_point", f32 * 3), ("seed", u32), ("c", f32 * 3), ("pdf_nee", f32),
                ("pdf_dir", f32), ("budget", i32), ("_pad", i32 * 2)]


class lj_hit(C.Structure):
    _fields_ = [("t", f32), ("u", f32), ("v", f32), ("shape_id", i32), ("primitive_id", i32)]


class lj_vertex(C.Structure):
    _fields_ = [("position", f32 * 3), ("geometric_normal", f32 * 3), ("frame_x", f32 * 3), ("frame_y", f32 * 3),
                ("frame_n", f32 * 3), ("st", f32 * 2), ("uv", f32 * 2), ("uv_screen_size", f32),
                ("mean_curvature", f32), ("ray_radius", f32), ("shape_id", i32), ("primitive_id", i32),
                ("material_id", i32), ("interior_medium_id", i32), ("exterior_medium_id", i32)]


class lj_bsdf_query(C.Structure):
    _fields_ = [("vertex", lj_vertex), ("dir_in", f32 * 3), ("dir_out", f32 * 3), ("rnd_uv", f32 * 2),
                ("rnd_w", f32), ("transport", i32)]


class lj_bsdf_result(C.Structure):
    _fields_ = [("f", f32 * 3), ("pdf", f32), ("sampled", i32), ("s_dir_out", f32 * 3), ("s_eta", f32),
                ("s_roughness", f32)]


class lj_light_query(C.Structure):
    _fields_ = [("ref_point", f32 * 3), ("rnd_uv", f32 * 2), ("rnd_w", f32), ("light_w", f32)]


class lj_medium_query(C.Structure):
    _fields_ = [("org", f32 * 3), ("tfar", f32), ("dir", f32 * 3), ("t", f32), ("rnd", f32 * 2), ("medium_id", i32), ("_pad", i32)]


class lj_medium_result(C.Structure):
    _fields_ = [("majorant", f32 * 3), ("sigma_a", f32 * 3), ("sigma_s", f32 * 3), ("phase_dir", f32 * 3),
                ("phase_eval", f32), ("phase_pdf", f32)]


class lj_medium_bound(C.Structure):
    _fields_ = [("majorant", f32 * 3), ("t_exit", f32), ("sigma_t", f32 * 3), ("local", i32)]


class lj_light_result(C.Structure):
    _fields_ = [("light_id", i32), ("position", f32 * 3), ("normal", f32 * 3), ("pmf", f32), ("pdf", f32),
                ("emission", f32 * 3)]


class lj_scene_info(C.Structure):
    _fields_ = [("num_prims", i32), ("num_triangles", i32), ("num_spheres", i32), ("num_bvh_nodes", i32),
                ("bvh_width", i32), ("bvh_depth", i32), ("bounds_lo", f32 * 3), ("bounds_hi", f32 * 3), ("bsphere_radius", f32),
                ("bsphere_center", f32 * 3), ("shadow_epsilon", f32), ("bvh_build_ms", f64), ("upload_ms", f64),
                ("prep_ms", f64), ("sah_cost", f64), ("device_bytes", i64), ("num_prim_refs", i32), ("_pad", i32)]


# symbol -> (restype, argtypes); every function include/lajolla_b200.h declares.
PROTOTYPES = {
    "lj_init": (C.c_int, [pi32, C.c_int]),
    "lj_last_error": (C.c_char_p, []),
    "lj_scene_create": (C.c_int, [C.POINTER(lj_scene_desc), C.POINTER(C.c_void_p)]),
    "lj_scene_destroy": (None, [C.c_void_p]),
    "lj_render": (C.c_int, [C.c_void_p, C.POINTER(lj_render_opts), pf32, C.POINTER(lj_stats)]),
    "lj_render_device": (C.c_int, [C.c_void_p, C.POINTER(lj_render_opts), C.c_void_p, C.c_void_p, C.POINTER(lj_stats)]),
    "lj_trace_closest": (C.c_int, [C.c_void_p, C.POINTER(lj_ray), i64, C.POINTER(lj_hit), C.POINTER(f64)]),
    "lj_trace_any": (C.c_int, [C.c_void_p, C.POINTER(lj_ray), i64, C.POINTER(u8), C.POINTER(f64)]),
    "lj_trace_closest_ex": (C.c_int, [C.c_void_p, C.POINTER(lj_ray), i64, C.POINTER(lj_trace_opts), C.POINTER(lj_hit), C.POINTER(f64)]),
    "lj_trace_any_ex": (C.c_int, [C.c_void_p, C.POINTER(lj_ray), i64, C.POINTER(lj_trace_opts), C.POINTER(u8), C.POINTER(f64)]),
    "lj_nee_walk_batch": (C.c_int, [C.c_void_p, C.POINTER(lj_walk_query), i64, C.POINTER(lj_trace_opts), pf32, C.POINTER(f64)]),
    "lj_intersect": (C.c_int, [C.c_void_p, C.POINTER(lj_ray), pf32, i64, C.POINTER(lj_vertex)]),
    "lj_bsdf_batch": (C.c_int, [C.c_void_p, C.POINTER(lj_bsdf_query), i64, C.POINTER(lj_bsdf_result)]),
    "lj_light_batch": (C.c_int, [C.c_void_p, C.POINTER(lj_light_query), i64, C.POINTER(lj_light_result)]),
    "lj_medium_batch": (C.c_int, [C.c_void_p, C.POINTER(lj_medium_query), i64, C.POINTER(lj_medium_result)]),
    "lj_medium_bound_batch": (C.c_int, [C.c_void_p, C.POINTER(lj_medium_query), i64, C.POINTER(lj_medium_bound)]),
    "lj_camera_rays": (C.c_int, [C.c_void_p, pf32, i64, C.POINTER(lj_ray)]),
    "lj_texture_batch": (C.c_int, [C.c_void_p, i32, i32, pf32, i64, pf32]),
    "lj_pcg32_batch": (C.c_int, [u64, u64, i32, i32, C.POINTER(u32), pf32]),
    "lj_scene_get_info": (C.c_int, [C.c_void_p, C.POINTER(lj_scene_info)]),
    "lj_scene_get_light_table": (C.c_int, [C.c_void_p, pf32, pf32]),
    "lj_measure_read_bandwidth": (C.c_int, [i64, i32, C.POINTER(f64)]),
    "lj_scene_get_mip_level": (C.c_int, [C.c_void_p, i32, i32, i32, pi32, pi32, pf32]),
}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libljb200.so")
_lib = None


class LajollaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libljb200 error {code}: {msg}")
        self.code = code


def load_library():
    """Load libljb200.so (built in-tree by `make -C lajolla_public_b200`).  Fails loudly if it is missing:
    there is no Python or CPU implementation behind this package."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LJ_LIB", LIB_PATH)  # A/B builds of the CUDA library for tuning runs
    if not os.path.exists(path):
        raise ImportError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). lajolla_public_b200 has no fallback implementation.")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != LJ_OK:
        raise LajollaError(code, load_library().lj_last_error().decode("utf-8", "replace"))
