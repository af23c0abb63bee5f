"""The C-ABI shared library loads without a GPU and exports every symbol include/lajolla_b200.h declares."""
import ctypes as C
import os
import re
import subprocess

import pytest

from lajolla_public_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lajolla_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lj_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    names = declared_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"libljb200.so does not export {n}"
        assert n in abi.PROTOTYPES, f"abi.py has no prototype for {n}"
    assert sorted(abi.PROTOTYPES) == names


def test_struct_sizes_match_the_header(tmp_path):
    structs = ["lj_image_desc", "lj_texture_desc", "lj_material_desc", "lj_shape_desc", "lj_light_desc", "lj_volume_desc",
               "lj_medium_desc", "lj_camera_desc", "lj_options_desc", "lj_scene_desc", "lj_render_opts", "lj_stats", "lj_ray",
               "lj_hit", "lj_vertex", "lj_bsdf_query", "lj_bsdf_result", "lj_light_query", "lj_light_result", "lj_scene_info",
               "lj_medium_query", "lj_medium_result", "lj_medium_bound", "lj_walk_query", "lj_trace_opts"]
    prog = '#include <stdio.h>\n#include "lajolla_b200.h"\nint main(void){' + "".join(
        f'printf("{s} %zu\\n", sizeof({s}));' for s in structs) + "return 0;}"
    c = tmp_path / "sizes.c"
    c.write_text(prog)
    exe = tmp_path / "sizes"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    for line in out.strip().splitlines():
        name, size = line.split()
        assert C.sizeof(getattr(abi, name)) == int(size), name


def test_no_device_is_an_error_not_a_fallback():
    """Without a CUDA device the library must refuse (LJ_ERR_NO_DEVICE), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = abi.load_library()
    assert lib.lj_init(None, 1) == abi.LJ_ERR_NO_DEVICE
    assert b"no CPU path" in lib.lj_last_error()
    import lajolla_public_b200 as lj
    from lajolla_public_b200 import ljs
    with pytest.raises(lj.LajollaError):
        lj.Scene(ljs.SceneDesc())
