"""GPU parity tests: the CUDA kernels, through the C ABI, against the oracle on the same inputs."""
import json
import os

import numpy as np
import pytest

import lajolla_public_b200 as lj
import parity_checks as pc

pytestmark = pytest.mark.gpu
SCENES = ["cbox", "veach_mi", "sponza"]
_cache = {}
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def pair(oracle, name):
    if name not in _cache:
        _cache[name] = (lj.parse_scene(oracle.scene_ljs(name)), oracle.RefScene(oracle.scene_xml(name)))
    return _cache[name]


def record(name, data):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "parity.jsonl"), "a") as f:
        f.write(json.dumps({"test": name, **data}) + "\n")


def test_pcg32_bit_exact(oracle):
    u, f = lj.pcg32(0, 4096, 64)
    ru, rf = oracle.pcg32(0, 4096, 64)
    assert np.array_equal(u, ru)
    assert np.all(np.abs(f - rf) < 2.0 ** -23)
    u2, _ = lj.pcg32(2 ** 40 + 17, 256, 16, seed=42)
    assert np.array_equal(u2, oracle.pcg32(2 ** 40 + 17, 256, 16, seed=42)[0])


@pytest.mark.parametrize("name", SCENES)
def test_scene_tables(oracle, name):
    sc, ref = pair(oracle, name)
    pc.check_scene_info(sc, ref)
    pc.check_light_table(sc, ref)


@pytest.mark.parametrize("name", SCENES + ["disney_bsdf", "vol_cbox_teapot", "hetvol"])
def test_ray_parity(oracle, name):
    """Parity test 1 of north_star: same primitive for >= 99.99 % of rays, t within 1e-5 relative."""
    sc, ref = pair(oracle, name)
    n = 1 << 18 if name != "sponza" else 1 << 17
    rays = pc.primary_rays(ref, n)
    r1 = pc.check_ray_parity(sc, ref, rays)
    r2 = pc.check_ray_parity(sc, ref, pc.bounce_rays(ref, rays))
    occ = pc.check_occlusion_parity(sc, ref, pc.shadow_rays(ref, rays))
    record("ray_parity", dict(scene=name, primary=r1, bounce=r2, occlusion_agree=occ))


@pytest.mark.parametrize("name", ["cbox", "sponza", "disney_bsdf", "vol_cbox_teapot", "veach_mi"])
def test_wavefront_kernels_ray_parity(oracle, name):
    """Parity test 1 on the kernels that actually render: k_trace_q<0|1> (queue form, the path integrator's) and
    k_trace<0|1> (lane form) fed through the path pool exactly as lj_render launches them, with pools that are full,
    sparse (every 7th / 37th slot), smaller than the batch (several rounds) and not a multiple of the fetch chunk."""
    from lajolla_public_b200 import abi
    sc, ref = pair(oracle, name)
    n = 1 << 17
    rays = pc.primary_rays(ref, n + 13)
    Q, L = abi.LJ_TRACE_WAVEFRONT, abi.LJ_TRACE_WAVEFRONT_LANE
    cfg = [(Q, 0, 1), (Q, 0, 7), (Q, 1 << 15, 1), (Q, 50000, 37), (Q, 1000, 1), (L, 0, 1), (L, 0, 37), (L, 50000, 3)]
    r1 = pc.check_wavefront_trace(sc, ref, rays, False, cfg)
    r2 = pc.check_wavefront_trace(sc, ref, pc.bounce_rays(ref, rays), False, cfg)
    r3 = pc.check_wavefront_trace(sc, ref, pc.shadow_rays(ref, rays), True, cfg)
    record("wavefront_ray_parity", dict(scene=name, primary=r1, bounce=r2, shadow=r3))


@pytest.mark.parametrize("name", ["pixel_filter_test", "pixel_filter_box", "pixel_filter_tent"])
def test_pixel_filters(oracle, name):
    """filters/{gaussian,box,tent}.inl on the device: sample_primary ray parity (the filter offset is a function of
    the sub-pixel position, camera.cpp:27-33) and a render of the 1000x-checkerboard floor, whose pixel means depend
    on the filter footprint, against the reference's render()."""
    sc, ref = pair(oracle, name)
    pc.check_camera_parity(sc, ref)
    img, var = sc.render(spp=64, variance=True)
    ref_img, _ = ref.render(spp=64)
    st = pc.image_stats(img, ref_img)
    record("pixel_filter", dict(scene=name, **{k: (float(v) if np.isscalar(v) else v) for k, v in st.items() if np.isscalar(v)}))
    m, rm = img.mean(axis=(0, 1)), ref_img.mean(axis=(0, 1))
    assert np.allclose(m, rm, rtol=0.02), (m, rm)
    # the filters differ visibly on this scene: a box / tent render must be closer to its own reference than to the
    # gaussian scene's (guards against the filter type being ignored)
    if name != "pixel_filter_test":
        _, gref = pair(oracle, "pixel_filter_test")
        g_img, _ = gref.render(spp=64)
        blur = lambda a: a.reshape(60, 8, 80, 8, 3).mean(axis=(1, 3))
        assert np.abs(blur(img) - blur(ref_img)).mean() <= np.abs(blur(img) - blur(g_img)).mean() + 1e-4


@pytest.mark.parametrize("name", ["volpath_test6", "vol_cbox_teapot", "hetvol", "hetvol_colored"])
def test_walk_kernels_parity(oracle, name):
    """The volpath NEE-walk kernels k_trace<2> (whole tracking loops), k_trace<3> (one collision per pass, traversal
    and tracking phases voted per warp) and the staged form (k_walk_begin -> [k_trace_q<0> over the walk view ->
    k_walk_track] x rounds -> k_walk_finish) against the serial walk, bit for bit, with full, sparse and multi-round pools."""
    from lajolla_public_b200 import abi
    sc, ref = pair(oracle, name)
    W, S, G = abi.LJ_TRACE_WALK_WHOLE, abi.LJ_TRACE_WALK_STEP, abi.LJ_TRACE_WALK_STAGED
    cfg = [(W, 0, 1), (W, 0, 7), (W, 20000, 1), (S, 0, 1), (S, 0, 37), (S, 30000, 3), (S, 1000, 1), (G, 0, 1), (G, 0, 37), (G, 30000, 3), (G, 1000, 1),
           (G, 0, 1, 1), (G, 30000, 3, 2)]  # (one / two rounds: most multi-segment walks end in k_walk_finish)
    record("walk_parity", dict(scene=name, **pc.check_walk_parity(sc, ref, 1 << 16, cfg)))


@pytest.mark.parametrize("name", SCENES)
def test_vertex_camera_light_parity(oracle, name):
    sc, ref = pair(oracle, name)
    rays = pc.primary_rays(ref, 1 << 15)
    rd = np.tile(np.array([0.0, 0.25 / 768], dtype=np.float32), (rays.shape[0], 1))
    pc.check_vertex_parity(sc, ref, rays, rd)
    pc.check_camera_parity(sc, ref)
    v = ref.intersect(rays)
    record("light_parity", dict(scene=name, **pc.check_light_parity(sc, ref, v["position"][v["shape_id"] >= 0])))


@pytest.mark.parametrize("name", ["cbox", "veach_mi", "sponza", "matpreview", "disney_bsdf", "disney_glass", "disney_clearcoat",
                                  "disney_sheen", "volpath_test5_2"])
def test_bsdf_parity(oracle, name):
    """Parity test 2 of north_star: BSDF eval / pdf / sample on fixed inputs."""
    sc, ref = pair(oracle, name)
    rays = pc.primary_rays(ref, 1 << 16)
    r = pc.check_bsdf_parity(sc, ref, pc.make_bsdf_queries(ref, rays))
    pc.check_bsdf_parity(sc, ref, pc.make_bsdf_queries(ref, pc.bounce_rays(ref, rays), seed=9, transport=1))
    pc.check_bsdf_parity(sc, ref, pc.fixed_material_queries(ref.info()["materials"]))
    record("bsdf_parity", dict(scene=name, **r))


def test_texture_and_mip_parity(oracle):
    sc, ref = pair(oracle, "sponza")
    rng = np.random.default_rng(3)
    q = np.concatenate([rng.random((20000, 2)) * 4 - 1, 10 ** rng.uniform(-5, -0.5, (20000, 1))], axis=1).astype(np.float32)
    q[:200, 2] = 0
    for m in range(0, 20, 3):
        a, b = sc.eval_texture(m, 0, q), ref.eval_texture(m, q)
        assert np.abs(a - b).max() < 2e-4, m
    for img in range(3):
        for lvl in range(4):
            a, b = sc.mip_level(img, lvl), ref.mip_level(img, lvl)
            assert a.shape == b.shape and np.abs(a - b).max() < 1e-6


@pytest.mark.parametrize("name,spp", [("cbox", 64), ("veach_mi", 256), ("sponza", 64)])
def test_image_parity(oracle, name, spp):
    """Parity test 3 of north_star: image agreement with the reference's own CPU render() at equal spp.
    The statistical test is calibrated against its own null: the device renders the scene twice with different
    seeds (A, B) and the z statistics of A - B (same estimator by construction) give the tail fractions that
    Monte Carlo noise alone produces at this spp; A - reference must not exceed them by more than a small margin.
      * per-channel image mean within 1 %
      * fraction of pixel channels with |z| > 3:       reference <= 1.25 x null + 0.3 %
      * fraction of 8x8 tile means with |z| > 4:       reference <= null + 0.2 %   (bias detector)
      * rms of the tile z:                             reference <= 1.15 x null
      * relMSE below the noise floor at these spp for the two low-variance scenes (0.05 cbox,
        0.1 sponza; veach_mi's specular highlights make relMSE a firefly detector, recorded only).
    The reference keeps no per-pixel variance; both renders draw from the same estimator, so var_ref = var_gpu is
    used -- in the null too (B's own variance is ignored), because a rare bright sample inflates the variance
    estimate of the render it falls in and the two statistics must be blind to it in the same way."""
    sc, ref = pair(oracle, name)
    img, var = sc.render(spp=spp, variance=True)
    st = sc.last_stats
    img_b, var_b = sc.render(spp=spp, variance=True, seed=0x5eed5eed5eed)
    h, w = img.shape[:2]
    assert st.samples == w * h * spp
    ref_img, secs = ref.render(spp=spp)
    null = pc.image_stats(img_b, img, var, var)  # B plays the reference: same rule, var of A on both sides
    s = pc.image_stats(img, ref_img, var, var)
    bound = {"cbox": 0.05, "veach_mi": float("inf"), "sponza": 0.1}[name]
    record("image_parity", dict(scene=name, spp=spp, gpu_ms=st.render_ms, ref_s=secs, **s,
                                null={k: null[k] for k in ("relmse", "frac_z_gt_3", "block_frac_z_gt_4", "block_z_rms")},
                                gpu_msamples=st.samples / st.render_ms / 1e3, ref_msamples=w * h * spp / secs / 1e6,
                                rays=st.closest_rays + st.shadow_rays, bounces=st.bounces, waves=st.waves))
    if os.environ.get("LJ_SAVE_IMAGES"):  # gpurun_out is capped at 64 MiB
        np.save(os.path.join(OUT, f"img_{name}_gpu.npy"), img.astype(np.float16))
        np.save(os.path.join(OUT, f"img_{name}_ref.npy"), ref_img.astype(np.float16))
    assert np.all(np.isfinite(img))
    assert np.allclose(s["mean"], s["ref_mean"], rtol=0.01), s
    assert s["relmse"] < bound, s
    assert s["frac_z_gt_3"] <= 1.25 * null["frac_z_gt_3"] + 0.003, (s, null)
    assert s["block_frac_z_gt_4"] <= null["block_frac_z_gt_4"] + 0.002, (s, null)
    assert s["block_z_rms"] <= 1.15 * null["block_z_rms"], (s, null)


@pytest.mark.parametrize("name", ["volpath_test6", "hetvol", "hetvol_colored", "vol_cbox_teapot"])
def test_medium_parity(oracle, name):
    sc, ref = pair(oracle, name)
    record("medium_parity", dict(scene=name, media=pc.check_medium_parity(sc, ref)))


def test_light_parity_facing_the_pole(oracle):
    """See tests/test_hostsim.py: sphere-light sampling frames around -z (lj_common.h coordinate_system)."""
    sc, ref = pair(oracle, "volpath_test5_2")
    rays = ref.sample_primary((0.45 + 0.1 * np.random.default_rng(11).random((1 << 15, 2))).astype(np.float32))
    v = ref.intersect(rays)
    record("light_parity", dict(scene="volpath_test5_2", **pc.check_light_parity(sc, ref, v["position"][v["shape_id"] >= 0])))


# Configs 3 and 5 of BASELINE.json.  The public reference ships these estimators as homework stubs; the image on the
# CPU side comes from the reference renderer with OUR handout restatement linked in (oracle/overlay, "lajolla_ref_hw"):
# PARITY UNPINNED by any reference code -- what this checks is device code == CPU restatement of the same handout.
@pytest.mark.parametrize("name,spp", [("disney_bsdf", 64), ("disney_glass", 32), ("disney_metal", 32),
                                      ("volpath_test4_2", 32), ("volpath_test5_2", 32), ("volpath_test6", 64), ("vol_cbox_teapot", 16),
                                      ("hetvol", 16), ("hetvol_colored", 16)])
def test_image_parity_homework_configs(oracle, name, spp):
    """Same null-calibrated statistics as test_image_parity (image mean within 1.5 %, tail fractions of the pixel
    and 8x8-tile z statistics no worse than what two device renders with different seeds give each other)."""
    sc, ref = pair(oracle, name)
    img, var = sc.render(spp=spp, variance=True)
    st = sc.last_stats
    img_b, _ = sc.render(spp=spp, variance=True, seed=0x5eed5eed5eed)
    h, w = img.shape[:2]
    assert st.samples == w * h * spp
    ref_img, secs = ref.render(spp=spp)
    null = pc.image_stats(img_b, img, var, var)
    s = pc.image_stats(img, ref_img, var, var)
    record("image_parity_hw", dict(scene=name, spp=spp, gpu_ms=st.render_ms, ref_s=secs, **s,
                                   null={k: null[k] for k in ("relmse", "frac_z_gt_3", "block_frac_z_gt_4", "block_z_rms")},
                                   gpu_msamples=st.samples / st.render_ms / 1e3, ref_msamples=w * h * spp / secs / 1e6,
                                   rays=st.closest_rays + st.shadow_rays, bounces=st.bounces, waves=st.waves))
    if os.environ.get("LJ_SAVE_IMAGES"):  # gpurun_out is capped at 64 MiB
        np.save(os.path.join(OUT, f"img_{name}_gpu.npy"), img.astype(np.float16))
        np.save(os.path.join(OUT, f"img_{name}_ref.npy"), ref_img.astype(np.float16))
    assert np.all(np.isfinite(img))
    # The lobe mixture of homework1.tex has unbounded f / pdf where a direction sampled from one lobe lies under the
    # flipped shading horizon (profiles/r02_disney_fireflies.txt: identical weight distribution on the device and in the
    # oracle, offending direction pairs listed): single samples of 1e4..1e7 occur on both sides.  The mean is therefore
    # compared on DISPLAY-REFERRED images (clamp to [0, 1], sRGB: what a viewer shows and what the handout's figures
    # hold); the z statistics below are per pixel / per tile and one firefly cannot move them.
    import flip
    m_gpu, m_ref = flip.tonemap(img).mean(axis=(0, 1)), flip.tonemap(ref_img).mean(axis=(0, 1))
    assert np.allclose(m_gpu, m_ref, rtol=0.01), (m_gpu, m_ref, s)
    assert s["frac_z_gt_3"] <= 1.25 * null["frac_z_gt_3"] + 0.003, (s, null)
    assert s["block_frac_z_gt_4"] <= 1.25 * null["block_frac_z_gt_4"] + 0.003, (s, null)
    assert s["block_z_rms"] <= 1.15 * null["block_z_rms"] + 0.02, (s, null)


@pytest.mark.parametrize("name", ["hetvol", "hetvol_colored"])
def test_block_majorants_bound_the_medium(oracle, name):
    sc, _ = pair(oracle, name)
    r = pc.check_block_majorants(sc)
    assert r, "scene has no grid medium"
    record("block_majorants", dict(scene=name, **r))


@pytest.mark.parametrize("name,spp", [("hetvol", 32), ("hetvol_colored", 32)])
def test_local_majorants_keep_the_expectation(oracle, name, spp):
    """The tracking loops bound a grid medium block by block (lj_media.h) where the reference uses one global majorant
    (medium.cpp:27-29).  Same estimator, different random sequence: (i) the NEE walks (ratio tracking + MIS weight)
    have the same mean contribution with and without the majorant grid, (ii) so do the images (mean within 1 %, tile z
    statistics no worse than two renders of the same kind give each other)."""
    old = os.environ.get("LJ_MAJ_BLOCK")
    try:
        os.environ["LJ_MAJ_BLOCK"] = "0"
        sc_global = lj.parse_scene(oracle.scene_ljs(name))
    finally:
        if old is None:
            os.environ.pop("LJ_MAJ_BLOCK", None)
        else:
            os.environ["LJ_MAJ_BLOCK"] = old
    sc, ref = pair(oracle, name)
    q = pc.make_walk_queries(sc, ref, 1 << 18)
    q["medium_id"] = np.where(q["medium_id"] < 0, 0, q["medium_id"])  # every walk starts inside a medium
    a, b = sc.nee_walks(q).astype(np.float64), sc_global.nee_walks(q).astype(np.float64)
    assert np.all(np.isfinite(a)) and np.all(np.isfinite(b))
    se = np.sqrt(a.var(axis=0) / len(a) + b.var(axis=0) / len(b))
    zw = (a.mean(axis=0) - b.mean(axis=0)) / np.maximum(se, 1e-30)
    img, var = sc.render(spp=spp, variance=True)
    st = sc.last_stats
    img_g, var_g = sc_global.render(spp=spp, variance=True)
    st_g = sc_global.last_stats
    img_g2, _ = sc_global.render(spp=spp, variance=True, seed=0x5eed5eed5eed)
    null = pc.image_stats(img_g2, img_g, var_g, var_g)
    s = pc.image_stats(img, img_g, var, var_g)
    record("local_majorants", dict(scene=name, spp=spp, walk_mean=a.mean(axis=0).tolist(), walk_mean_global=b.mean(axis=0).tolist(), walk_z=zw.tolist(),
                                   ms=st.render_ms, ms_global=st_g.render_ms, **s, null={k: null[k] for k in ("frac_z_gt_3", "block_frac_z_gt_4", "block_z_rms")}))
    sc_global.close()
    assert np.abs(zw).max() < 4.5, (zw, a.mean(axis=0), b.mean(axis=0))
    assert np.allclose(img.mean(axis=(0, 1)), img_g.mean(axis=(0, 1)), rtol=0.01), s
    assert s["frac_z_gt_3"] <= 1.25 * null["frac_z_gt_3"] + 0.003, (s, null)
    assert s["block_frac_z_gt_4"] <= 1.25 * null["block_frac_z_gt_4"] + 0.003, (s, null)
    assert s["block_z_rms"] <= 1.15 * null["block_z_rms"] + 0.02, (s, null)


@pytest.mark.parametrize("name", ["cbox", "veach_mi", "matpreview", "disney_bsdf", "volpath_test6", "hetvol"])
def test_golden_vectors(oracle, name):
    """The CUDA kernels against the committed fixtures (tests/golden, answers of the reference's object code)."""
    import golden_lib
    sc, _ = pair(oracle, name)
    g = golden_lib.GoldenScene(name)
    record("golden", dict(scene=name, **golden_lib.run_checks(sc, g, name)))
    assert g.calls >= 10


def test_render_against_golden_tiles(oracle):
    import golden_lib
    z = np.load(golden_lib.GOLDEN_DIR + "/cbox_render.npz")
    sc, _ = pair(oracle, "cbox")
    img = sc.render(spp=64)
    tiles = img.reshape(32, 16, 32, 16, 3).mean(axis=(1, 3))
    gold = z["tiles"]
    assert np.allclose(tiles.mean(axis=(0, 1)), gold.mean(axis=(0, 1)), rtol=0.01)
    lum_t, lum_g = tiles.sum(axis=2), gold.sum(axis=2)
    lit = lum_g > 0.05 * lum_g.mean()
    assert np.median(np.abs(lum_t[lit] - lum_g[lit]) / lum_g[lit]) < 0.03


@pytest.mark.parametrize("name", ["cbox", "veach_mi", "sponza"])
def test_aux_integrators_pixel_exact(oracle, name):
    """The deterministic debug integrators (render.cpp:12-69): device image == reference image, pixel by pixel
    (depth 1e-4 relative, shading normals 1e-5 median, mean curvature, mip level)."""
    from lajolla_public_b200 import ljs
    desc = ljs.load(oracle.scene_ljs(name))
    ref = oracle.RefScene(oracle.scene_xml(name))
    record("aux_parity", dict(scene=name, **pc.check_aux_parity(lambda d: lj.Scene(d), ref, desc)))


def test_sample_range_split_is_additive(oracle):
    """Multi-GPU partition property: spp blocks rendered separately sum to the single render."""
    sc, _ = pair(oracle, "cbox")
    full = sc.render(spp=8, normalize=False)
    parts = sum(sc.render(spp=8, sample_begin=b, sample_end=b + 2, normalize=False) for b in range(0, 8, 2))
    assert np.allclose(full, parts, rtol=1e-4, atol=1e-4)


def test_full_size_properties(oracle):
    """BASELINE-size run (sponza 768x575): every sample accounted for, film finite and non-negative,
    ray count consistent with the path-depth bookkeeping (rays = samples + bounces + shadow rays)."""
    sc, _ = pair(oracle, "sponza")
    img = sc.render(spp=128)
    st = sc.last_stats
    assert st.samples == 768 * 575 * 128
    assert np.all(np.isfinite(img)) and img.min() >= 0
    assert st.closest_rays >= st.samples and st.closest_rays <= st.samples + st.bounces
    assert st.shadow_rays <= st.bounces
    record("full_size", dict(scene="sponza", spp=128, ms=st.render_ms, msamples=st.samples / st.render_ms / 1e3,
                             mrays=(st.closest_rays + st.shadow_rays) / st.render_ms / 1e3,
                             extend_ms=st.extend_ms, shade_ms=st.shade_ms, shadow_ms=st.shadow_ms, regen_ms=st.regen_ms, waves=st.waves))
