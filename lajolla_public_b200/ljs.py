"""Reader/writer of the `.ljs` flat-scene container and its conversion to the C-ABI `lj_scene_desc`.

`.ljs` is a little-endian dump of exactly the fields of include/lajolla_b200.h's description structs
(what the reference's parse_scene() + Scene::Scene produce, scene.cpp:4-53), written by the host
front end (`lajolla --dump-ljs`) and, for parity tests, by the oracle's dumper over the reference's
own parser (oracle/ref_glue.cpp: ljo_scene_dump).  Layout, all int32/float32:

  "LJS1" version
  camera : cam_to_world[16] world_to_cam[16] sample_to_cam[16] cam_to_sample[16] w h filter_type filter_param medium_id
  options: integrator spp max_depth rr_depth vol_path_version max_null_collisions
  counts : num_images num_materials num_shapes num_lights num_media envmap_light_id
  images : (w h channels, w*h*channels floats)*            3-channel images first, then 1-channel
  materials: (type eta, 12 x texture)*                     texture = kind image_id value[3] color1[3] us vs uo vo
  shapes : (type material light interior exterior center[3] radius nv nt has_normals has_uvs
            positions[3nv] indices[3nt] normals[3nv]? uvs[2nv]?)*
  lights : (type shape_id intensity[3] texture to_world[16] to_local[16] scale)*
  media  : (type phase_type g sigma_a[3] sigma_s[3] [volume albedo, volume density if heterogeneous])*
           volume = is_grid nx ny nz value_or_max[3] p_min[3] p_max[3] scale [3*nx*ny*nz floats if grid]
"""
import ctypes as C
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import abi


@dataclass
class Texture:
    kind: int = 0
    image_id: int = -1
    value: tuple = (0.0, 0.0, 0.0)
    color1: tuple = (0.0, 0.0, 0.0)
    uscale: float = 1.0
    vscale: float = 1.0
    uoffset: float = 0.0
    voffset: float = 0.0


@dataclass
class Material:
    type: int = 0
    eta: float = 1.0
    tex: List[Texture] = field(default_factory=lambda: [Texture() for _ in range(abi.LJ_NUM_TEX_SLOTS)])


@dataclass
class Shape:
    type: int = 0
    material_id: int = -1
    area_light_id: int = -1
    interior_medium_id: int = -1
    exterior_medium_id: int = -1
    center: tuple = (0.0, 0.0, 0.0)
    radius: float = 0.0
    positions: Optional[np.ndarray] = None  # (nv,3) f32
    indices: Optional[np.ndarray] = None    # (nt,3) i32
    normals: Optional[np.ndarray] = None
    uvs: Optional[np.ndarray] = None


@dataclass
class Light:
    type: int = 0
    shape_id: int = -1
    intensity: tuple = (0.0, 0.0, 0.0)
    values: Texture = field(default_factory=Texture)
    to_world: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    to_local: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    scale: float = 1.0


@dataclass
class Volume:
    is_grid: int = 0
    res: tuple = (0, 0, 0)
    value: tuple = (0.0, 0.0, 0.0)
    p_min: tuple = (0.0, 0.0, 0.0)
    p_max: tuple = (0.0, 0.0, 0.0)
    scale: float = 1.0
    data: Optional[np.ndarray] = None  # (nz,ny,nx,3) f32


@dataclass
class Medium:
    type: int = 0
    phase_type: int = 0
    phase_g: float = 0.0
    sigma_a: tuple = (0.0, 0.0, 0.0)
    sigma_s: tuple = (0.0, 0.0, 0.0)
    albedo: Volume = field(default_factory=Volume)
    density: Volume = field(default_factory=Volume)


@dataclass
class Camera:
    cam_to_world: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    world_to_cam: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    sample_to_cam: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    cam_to_sample: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    width: int = 256
    height: int = 256
    filter_type: int = 0
    filter_param: float = 1.0
    medium_id: int = -1


@dataclass
class Options:
    integrator: int = 5
    samples_per_pixel: int = 4
    max_depth: int = -1
    rr_depth: int = 5
    vol_path_version: int = 0
    max_null_collisions: int = 1000


@dataclass
class SceneDesc:
    camera: Camera = field(default_factory=Camera)
    options: Options = field(default_factory=Options)
    images: List[np.ndarray] = field(default_factory=list)  # (h,w,c) f32
    materials: List[Material] = field(default_factory=list)
    shapes: List[Shape] = field(default_factory=list)
    lights: List[Light] = field(default_factory=list)
    media: List[Medium] = field(default_factory=list)
    envmap_light_id: int = -1


class _Reader:
    def __init__(self, buf):
        self.b = memoryview(buf)
        self.o = 0

    def i32(self, n=None):
        if n is None:
            v = struct.unpack_from("<i", self.b, self.o)[0]
            self.o += 4
            return v
        a = np.frombuffer(self.b, dtype="<i4", count=n, offset=self.o).copy()
        self.o += 4 * n
        return a

    def f32(self, n=None):
        if n is None:
            v = struct.unpack_from("<f", self.b, self.o)[0]
            self.o += 4
            return v
        a = np.frombuffer(self.b, dtype="<f4", count=n, offset=self.o).copy()
        self.o += 4 * n
        return a

    def texture(self):
        kind, image_id = self.i32(), self.i32()
        v = self.f32(10)
        return Texture(kind, image_id, tuple(v[0:3]), tuple(v[3:6]), float(v[6]), float(v[7]), float(v[8]), float(v[9]))

    def volume(self):
        is_grid = self.i32()
        res = tuple(int(x) for x in self.i32(3))
        v = self.f32(10)
        vol = Volume(is_grid, res, tuple(v[0:3]), tuple(v[3:6]), tuple(v[6:9]), float(v[9]))
        if is_grid:
            n = res[0] * res[1] * res[2]
            vol.data = self.f32(3 * n).reshape(res[2], res[1], res[0], 3)
        return vol


def load(path) -> SceneDesc:
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:4] != b"LJS1":
        raise ValueError(f"{path}: not an .ljs file")
    r = _Reader(buf)
    r.o = 4
    version = r.i32()
    if version != 1:
        raise ValueError(f"{path}: unsupported .ljs version {version}")
    s = SceneDesc()
    c = s.camera
    c.cam_to_world = r.f32(16).reshape(4, 4)
    c.world_to_cam = r.f32(16).reshape(4, 4)
    c.sample_to_cam = r.f32(16).reshape(4, 4)
    c.cam_to_sample = r.f32(16).reshape(4, 4)
    c.width, c.height, c.filter_type = r.i32(), r.i32(), r.i32()
    c.filter_param = r.f32()
    c.medium_id = r.i32()
    o = s.options
    o.integrator, o.samples_per_pixel, o.max_depth, o.rr_depth, o.vol_path_version, o.max_null_collisions = (r.i32() for _ in range(6))
    n_img, n_mat, n_shape, n_light, n_med, s.envmap_light_id = (r.i32() for _ in range(6))
    for _ in range(n_img):
        w, h, ch = r.i32(), r.i32(), r.i32()
        s.images.append(r.f32(w * h * ch).reshape(h, w, ch))
    for _ in range(n_mat):
        m = Material(r.i32(), r.f32())
        m.tex = [r.texture() for _ in range(abi.LJ_NUM_TEX_SLOTS)]
        s.materials.append(m)
    for _ in range(n_shape):
        sh = Shape(r.i32(), r.i32(), r.i32(), r.i32(), r.i32())
        sh.center = tuple(r.f32(3))
        sh.radius = r.f32()
        nv, nt, has_n, has_uv = r.i32(), r.i32(), r.i32(), r.i32()
        if sh.type == 1:
            sh.positions = r.f32(3 * nv).reshape(nv, 3)
            sh.indices = r.i32(3 * nt).reshape(nt, 3)
            if has_n:
                sh.normals = r.f32(3 * nv).reshape(nv, 3)
            if has_uv:
                sh.uvs = r.f32(2 * nv).reshape(nv, 2)
        s.shapes.append(sh)
    for _ in range(n_light):
        l = Light(r.i32(), r.i32())
        l.intensity = tuple(r.f32(3))
        l.values = r.texture()
        l.to_world = r.f32(16).reshape(4, 4)
        l.to_local = r.f32(16).reshape(4, 4)
        l.scale = r.f32()
        s.lights.append(l)
    for _ in range(n_med):
        m = Medium(r.i32(), r.i32(), r.f32())
        m.sigma_a = tuple(r.f32(3))
        m.sigma_s = tuple(r.f32(3))
        if m.type == 1:
            m.albedo = r.volume()
            m.density = r.volume()
        s.media.append(m)
    if r.o != len(buf):
        raise ValueError(f"{path}: {len(buf) - r.o} trailing bytes")
    return s


def _tex_c(t: Texture):
    d = abi.lj_texture_desc()
    d.kind, d.image_id = t.kind, t.image_id
    d.value[:] = t.value
    d.color1[:] = t.color1
    d.uscale, d.vscale, d.uoffset, d.voffset = t.uscale, t.vscale, t.uoffset, t.voffset
    return d


def _fptr(a):
    return a.ctypes.data_as(abi.pf32)


def _vol_c(v: Volume, keep):
    d = abi.lj_volume_desc()
    d.is_grid = v.is_grid
    d.res[:] = v.res
    d.value[:] = v.value
    d.p_min[:] = v.p_min
    d.p_max[:] = v.p_max
    d.scale = v.scale
    if v.is_grid:
        a = np.ascontiguousarray(v.data, dtype=np.float32)
        keep.append(a)
        d.data = _fptr(a)
    return d


def to_c(s: SceneDesc):
    """SceneDesc -> (lj_scene_desc, keepalive list).  The caller must hold `keepalive` until
    lj_scene_create has returned (the library copies everything)."""
    keep = []
    d = abi.lj_scene_desc()
    c = s.camera
    for name in ("cam_to_world", "world_to_cam", "sample_to_cam", "cam_to_sample"):
        getattr(d.camera, name)[:] = [float(x) for x in np.asarray(getattr(c, name), dtype=np.float32).reshape(16)]
    d.camera.width, d.camera.height, d.camera.filter_type = c.width, c.height, c.filter_type
    d.camera.filter_param, d.camera.medium_id = c.filter_param, c.medium_id
    o = s.options
    d.options.integrator, d.options.samples_per_pixel, d.options.max_depth = o.integrator, o.samples_per_pixel, o.max_depth
    d.options.rr_depth, d.options.vol_path_version, d.options.max_null_collisions = o.rr_depth, o.vol_path_version, o.max_null_collisions
    d.num_images, d.num_materials, d.num_shapes = len(s.images), len(s.materials), len(s.shapes)
    d.num_lights, d.num_media, d.envmap_light_id = len(s.lights), len(s.media), s.envmap_light_id
    imgs = (abi.lj_image_desc * max(len(s.images), 1))()
    for i, im in enumerate(s.images):
        a = np.ascontiguousarray(im, dtype=np.float32)
        keep.append(a)
        imgs[i].height, imgs[i].width, imgs[i].channels = a.shape
        imgs[i].data = _fptr(a)
    mats = (abi.lj_material_desc * max(len(s.materials), 1))()
    for i, m in enumerate(s.materials):
        mats[i].type, mats[i].eta = m.type, m.eta
        for k in range(abi.LJ_NUM_TEX_SLOTS):
            mats[i].tex[k] = _tex_c(m.tex[k])
    shapes = (abi.lj_shape_desc * max(len(s.shapes), 1))()
    for i, sh in enumerate(s.shapes):
        e = shapes[i]
        e.type, e.material_id, e.area_light_id = sh.type, sh.material_id, sh.area_light_id
        e.interior_medium_id, e.exterior_medium_id = sh.interior_medium_id, sh.exterior_medium_id
        e.center[:] = sh.center
        e.radius = sh.radius
        if sh.type == 1:
            p = np.ascontiguousarray(sh.positions, dtype=np.float32)
            ix = np.ascontiguousarray(sh.indices, dtype=np.int32)
            keep += [p, ix]
            e.num_vertices, e.num_triangles = p.shape[0], ix.shape[0]
            e.positions, e.indices = _fptr(p), ix.ctypes.data_as(abi.pi32)
            if sh.normals is not None:
                n = np.ascontiguousarray(sh.normals, dtype=np.float32)
                keep.append(n)
                e.normals = _fptr(n)
            if sh.uvs is not None:
                u = np.ascontiguousarray(sh.uvs, dtype=np.float32)
                keep.append(u)
                e.uvs = _fptr(u)
    lights = (abi.lj_light_desc * max(len(s.lights), 1))()
    for i, l in enumerate(s.lights):
        e = lights[i]
        e.type, e.shape_id = l.type, l.shape_id
        e.intensity[:] = l.intensity
        e.values = _tex_c(l.values)
        e.to_world[:] = [float(x) for x in np.asarray(l.to_world, dtype=np.float32).reshape(16)]
        e.to_local[:] = [float(x) for x in np.asarray(l.to_local, dtype=np.float32).reshape(16)]
        e.scale = l.scale
    media = (abi.lj_medium_desc * max(len(s.media), 1))()
    for i, m in enumerate(s.media):
        e = media[i]
        e.type, e.phase_type, e.phase_g = m.type, m.phase_type, m.phase_g
        e.sigma_a[:] = m.sigma_a
        e.sigma_s[:] = m.sigma_s
        e.albedo = _vol_c(m.albedo, keep)
        e.density = _vol_c(m.density, keep)
    d.images = C.cast(imgs, C.POINTER(abi.lj_image_desc))
    d.materials = C.cast(mats, C.POINTER(abi.lj_material_desc))
    d.shapes = C.cast(shapes, C.POINTER(abi.lj_shape_desc))
    d.lights = C.cast(lights, C.POINTER(abi.lj_light_desc))
    d.media = C.cast(media, C.POINTER(abi.lj_medium_desc))
    keep += [imgs, mats, shapes, lights, media]
    return d, keep
