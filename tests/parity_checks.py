"""Parity checks shared by the host-simulation tests (CPU, -m "not gpu") and the GPU tests (-m gpu).
Each takes `scene` (lajolla_public_b200.Scene: the code under test, through the C ABI) and `ref`
(oracle_lib.RefScene: the reference's own object code) built from the same scene file.

Tolerances (north_star): hit primitive equal for >= 99.99 % of rays, t within 1e-5 relative;
BSDF eval/pdf/sample within 1e-5 relative -- that figure assumes equal arithmetic; the device path is
fp32 against the reference's fp64, so the per-quantity bounds below are the fp32 round-off of each
formula (stated per check), and the 1e-5 bound is asserted on the median error.
"""
import numpy as np

import lajolla_public_b200 as lj


def rel_err(a, b, floor=1e-6):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def primary_rays(ref, n, seed=1):
    rng = np.random.default_rng(seed)
    return ref.sample_primary(rng.random((n, 2)).astype(np.float32))


def bounce_rays(ref, rays, seed=2):
    """Second-generation rays: cosine-ish random directions leaving the reference's hit points."""
    rng = np.random.default_rng(seed)
    v = ref.intersect(rays)
    ok = v["shape_id"] >= 0
    v = v[ok]
    d = rng.normal(size=(v.shape[0], 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    n = v["geometric_normal"].astype(np.float64)
    flip = (d * n).sum(axis=1) < 0
    d[flip] *= -1
    eps = ref.info()["eps"]
    return lj.make_rays(v["position"], d.astype(np.float32), tnear=np.float32(eps), tfar=np.inf)


def check_ray_parity(scene, ref, rays, min_agree=0.9999, t_rel=1e-5):
    h1, h2 = scene.intersect_hits(rays), ref.intersect_hits(rays)
    same = (h1["shape_id"] == h2["shape_id"]) & (h1["primitive_id"] == h2["primitive_id"])
    agree = same.mean()
    assert agree >= min_agree, f"hit primitive agreement {agree:.6f} < {min_agree}"
    hit = same & (h2["shape_id"] >= 0)
    # t within 1e-5 relative for every agreeing ray (the device re-evaluates the winning primitive's t in
    # fp64, lj_bvh.h refine_hit_t; the only slack is the fp32 storage of t, 1e-9 R).
    R = ref.info()["radius"]
    t2 = h2["t"][hit].astype(np.float64)
    aerr = np.abs(h1["t"][hit].astype(np.float64) - t2)
    assert np.all(aerr <= t_rel * t2 + 1e-9 * R), f"t error {(aerr - t_rel * t2).max() / R:.3e} R beyond {t_rel} relative"
    terr = aerr / np.maximum(t2, 1e-2 * R)
    # barycentrics / sphere (u,v): the hit-point round-off (~1e-7 R) divided by the triangle's size, so small
    # triangles amplify it: median below 1e-5, 99.9 % below 2e-3.
    uverr = np.maximum(np.abs(h1["u"][hit] - h2["u"][hit]), np.abs(h1["v"][hit] - h2["v"][hit]))
    if uverr.size:
        assert np.median(uverr) < 1e-5 and np.quantile(uverr, 0.999) < 2e-3, (np.median(uverr), uverr.max())
    return dict(agree=float(agree), t_rel_max=float(terr.max()), n=int(rays.shape[0]), hit_frac=float((h2["shape_id"] >= 0).mean()))


def check_occlusion_parity(scene, ref, rays, min_agree=0.9999):
    o1, o2 = scene.occluded(rays), ref.occluded(rays)
    agree = (o1 == o2).mean()
    assert agree >= min_agree, f"occlusion agreement {agree:.6f}"
    return float(agree)


def check_wavefront_trace(scene, ref, rays, shadow, configs, min_agree=0.9999):
    """The persistent traversal kernels of the renderer (rays loaded into the path pool, lj_trace_*_ex) against
    (i) the plain per-thread loop: bit-identical primitive, t, u, v / occlusion flag -- the step functions and the
    equal-t tie policy are shared, so scheduling must not change any answer -- and (ii) the oracle, at the
    north_star bars (through check_ray_parity / check_occlusion_parity of the plain result they are identical to).
    configs: (kernel, pool_paths, slot_stride) tuples -- full, sparse and non-multiple-of-chunk pools."""
    from lajolla_public_b200 import abi
    out = {}
    if shadow:
        base = scene.occluded(rays)
        assert (base == ref.occluded(rays)).mean() >= min_agree
        for kernel, pool, stride in configs:
            got = scene.occluded(rays, kernel=kernel, pool_paths=pool, slot_stride=stride)
            assert np.array_equal(got, base), f"occlusion differs from the plain loop: kernel {kernel} pool {pool} stride {stride}: {(got != base).sum()} of {base.size}"
        out["occluded_frac"] = float(base.mean())
    else:
        base = scene.intersect_hits(rays)
        h2 = ref.intersect_hits(rays)
        same = (base["shape_id"] == h2["shape_id"]) & (base["primitive_id"] == h2["primitive_id"])
        assert same.mean() >= min_agree
        for kernel, pool, stride in configs:
            got = scene.intersect_hits(rays, kernel=kernel, pool_paths=pool, slot_stride=stride)
            for f in ("shape_id", "primitive_id"):
                assert np.array_equal(got[f], base[f]), f"{f} differs from the plain loop: kernel {kernel} pool {pool} stride {stride}: {(got[f] != base[f]).sum()} of {base.size}"
            for f in ("t", "u", "v"):
                assert np.array_equal(got[f].view(np.uint32), base[f].view(np.uint32)), f"{f} not bit-identical: kernel {kernel} pool {pool} stride {stride}"
        out["hit_frac"] = float((base["shape_id"] >= 0).mean())
    out["n"] = int(rays.shape[0])
    out["configs"] = len(configs)
    return out


def make_walk_queries(scene, ref, n, seed=12):
    """NEE walks as the volpath integrator issues them: from points inside the scene's media (and from surface points)
    towards sampled light points, with random budgets."""
    rng = np.random.default_rng(seed)
    info = scene.info()
    lo, hi = np.array(info.bounds_lo), np.array(info.bounds_hi)
    origin = (lo + rng.random((n, 3)) * (hi - lo)).astype(np.float32)
    lq = np.zeros(n, dtype=lj.LIGHT_QUERY_DTYPE)
    lq["ref_point"] = origin
    lq["rnd_uv"] = rng.random((n, 2))
    lq["rnd_w"] = rng.random(n)
    lq["light_w"] = rng.random(n)
    ls = ref.sample_lights(lq)
    q = np.zeros(n, dtype=lj.WALK_QUERY_DTYPE)
    q["origin"] = origin
    q["light_point"] = ls["position"]
    n_media = len(scene.desc.media)
    q["medium_id"] = rng.integers(-1, max(n_media, 1), size=n) if n_media else -1
    q["seed"] = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    q["c"] = rng.random((n, 3)) + 0.1
    q["pdf_nee"] = rng.random(n) + 0.05
    q["pdf_dir"] = rng.random(n) * 2
    q["budget"] = np.where(rng.random(n) < 0.3, rng.integers(0, 4, size=n), -1)
    return q


def check_walk_parity(scene, ref, n, configs):
    """k_trace<2> / k_trace<3> (the persistent NEE-walk kernels: closest-hit segment -> ratio tracking -> index-matched
    test -> next segment, walks handed to lanes through the sh_mask bit words) against the plain one-thread-per-walk
    loop over the same step functions: bit-identical contributions for every pool shape."""
    q = make_walk_queries(scene, ref, n)
    base = scene.nee_walks(q)
    assert np.all(np.isfinite(base))
    for cfg in configs:  # (kernel, pool_paths, slot_stride[, rounds of the staged walk before its whole-loop finish kernel])
        kernel, pool, stride = cfg[:3]
        rounds = cfg[3] if len(cfg) > 3 else 0
        got = scene.nee_walks(q, kernel=kernel, pool_paths=pool, slot_stride=stride, walk_rounds=rounds)
        bad = got.view(np.uint32) != base.view(np.uint32)
        assert not bad.any(), f"walk kernel {kernel} pool {pool} stride {stride} rounds {rounds}: {int(bad.any(axis=1).sum())} of {n} walks differ, e.g. {got[bad.any(axis=1)][:2]} vs {base[bad.any(axis=1)][:2]}"
    return dict(n=int(n), unblocked=float((np.abs(base).max(axis=1) > 0).mean()), mean=float(base.mean()), configs=len(configs))


def shadow_rays(ref, rays, seed=3):
    """Segments from hit points towards sampled light points, as path_tracing.h:124-128 builds them."""
    rng = np.random.default_rng(seed)
    v = ref.intersect(rays)
    v = v[v["shape_id"] >= 0]
    q = np.zeros(v.shape[0], dtype=lj.LIGHT_QUERY_DTYPE)
    q["ref_point"] = v["position"]
    q["rnd_uv"] = rng.random((v.shape[0], 2))
    q["rnd_w"] = rng.random(v.shape[0])
    q["light_w"] = rng.random(v.shape[0])
    ls = ref.sample_lights(q)
    d = ls["position"].astype(np.float64) - v["position"].astype(np.float64)
    dist = np.linalg.norm(d, axis=1)
    ok = dist > 0
    d = (d[ok] / dist[ok, None]).astype(np.float32)
    eps = ref.info()["eps"]
    return lj.make_rays(v["position"][ok], d, tnear=np.float32(eps), tfar=((1 - eps) * dist[ok]).astype(np.float32))


def check_vertex_parity(scene, ref, rays, ray_diff=None):
    """intersect() -> PathVertex fields (intersection.cpp:37-62)."""
    v1, v2 = scene.intersect(rays, ray_diff), ref.intersect(rays, ray_diff)
    same = (v1["shape_id"] == v2["shape_id"]) & (v1["primitive_id"] == v2["primitive_id"]) & (v2["shape_id"] >= 0)
    assert same.sum() > 0
    a, b = v1[same], v2[same]
    assert np.array_equal(a["material_id"], b["material_id"])
    scale = np.abs(b["position"]).max()
    assert np.abs(a["position"] - b["position"]).max() <= 2e-5 * scale
    for f in ("geometric_normal", "frame_n"):
        assert np.abs(a[f] - b[f]).max() < 2e-3, f  # normals of sliver triangles amplify fp32 vertex round-off
    # tangents can flip sign only with the normal; compare via |dot|
    dots = np.abs((a["frame_x"].astype(np.float64) * b["frame_x"]).sum(axis=1))
    assert np.median(dots) > 1 - 1e-5
    assert np.median(rel_err(a["uv"], b["uv"], 1e-3)) < 1e-5
    assert np.median(rel_err(a["uv_screen_size"], b["uv_screen_size"], 1e-9)) < 1e-4
    # degenerate uv triangles make upstream's dn/du infinite or NaN (triangle_mesh.inl:154-160 runs even when
    # det == 0); the device code must be non-finite on the same vertices, and agree where both are finite
    # (with FMA contraction on the GPU a determinant that is exactly 0 in fp64 can come out as a tiny non-zero,
    # turning NaN into a huge finite value: allow 2 % of such flips.  mean_curvature only feeds the ray
    # differential spread, which upstream never consumes -- envmap.inl:63-72 ignores its footprint argument.)
    fa, fb = np.isfinite(a["mean_curvature"]), np.isfinite(b["mean_curvature"])
    assert (fa == fb).mean() > 0.98
    ok = fa & fb & (np.abs(b["mean_curvature"]) < 1e3)
    assert np.median(rel_err(a["mean_curvature"][ok], b["mean_curvature"][ok], 1e-3)) < 1e-4
    return int(same.sum())


def make_bsdf_queries(ref, rays, seed=4, transport=0):
    """(vertex, dir_in, dir_out, u, w) tuples on the reference's own hit vertices, plus the fixed tuple of
    the reference's tests/materials.cpp:57-62 on every material (rnd=(0.3,0.4), w=0.6,
    dir_in=normalize(0.3,0.4,0.5), n=(0,0,1))."""
    rng = np.random.default_rng(seed)
    v = ref.intersect(rays)
    v = v[(v["shape_id"] >= 0) & (v["material_id"] >= 0)]
    n = v.shape[0]
    q = np.zeros(n, dtype=lj.BSDF_QUERY_DTYPE)
    q["vertex"] = v

    def rand_dirs(k):
        d = rng.normal(size=(k, 3))
        return d / np.linalg.norm(d, axis=1, keepdims=True)

    ng = v["geometric_normal"].astype(np.float64)
    wi = rand_dirs(n)
    wo = rand_dirs(n)
    # mostly upper-hemisphere pairs (the interesting case), a quarter left fully random (transmission / below)
    up = rng.random(n) < 0.75
    s = np.sign((wi * ng).sum(axis=1))
    wi[up & (s < 0)] *= -1
    s = np.sign((wo * ng).sum(axis=1))
    wo[up & (s < 0)] *= -1
    q["dir_in"], q["dir_out"] = wi, wo
    q["rnd_uv"] = rng.random((n, 2))
    q["rnd_w"] = rng.random(n)
    q["transport"] = transport
    return q


def fixed_material_queries(n_materials):
    q = np.zeros(n_materials, dtype=lj.BSDF_QUERY_DTYPE)
    for m in range(n_materials):
        vx = q["vertex"][m]
        vx["geometric_normal"] = (0, 0, 1)
        vx["frame_x"], vx["frame_y"], vx["frame_n"] = (1, 0, 0), (0, 1, 0), (0, 0, 1)
        vx["material_id"] = m
        vx["shape_id"] = 0
        d = np.array([0.3, 0.4, 0.5])
        q["dir_in"][m] = d / np.linalg.norm(d)
        o = np.array([-0.2, 0.1, 0.6])
        q["dir_out"][m] = o / np.linalg.norm(o)
        q["rnd_uv"][m] = (0.3, 0.4)
        q["rnd_w"][m] = 0.6
    return q


def _tail_ok(e, med_tol, out, key):
    """north_star's 1e-5 on the median; the fp32 tail characterised in profiles/r02_bsdf_tail.txt: at most 0.1 % of the
    queries above 1e-4 (grazing directions and values 1e-6 of the peak), none above 2e-3."""
    out[key + "_med"], out[key + "_max"] = float(np.median(e)), float(e.max())
    out[key + "_gt1e4"] = float((e > 1e-4).mean())
    assert np.median(e) <= med_tol and (e > 1e-4).mean() <= 1e-3 and e.max() <= 2e-3, out


def check_bsdf_parity(scene, ref, q, med_tol=1e-5):
    r1, r2 = scene.bsdf(q), ref.bsdf(q)
    out = {}
    nz = (np.abs(r2["f"]).max(axis=1) > 1e-12)
    # a query exactly on a support boundary (n.wo == 0 within fp32) may flip between zero / non-zero
    z1 = (np.abs(r1["f"]).max(axis=1) > 1e-12)
    assert (nz == z1).mean() > 0.999
    both = nz & z1
    if both.any():
        e = rel_err(r1["f"][both], r2["f"][both], 1e-9).max(axis=1)
        _tail_ok(e, med_tol, out, "f")
    pz = (r2["pdf"] > 1e-12) & (r1["pdf"] > 1e-12)
    if pz.any():
        e = rel_err(r1["pdf"][pz], r2["pdf"][pz], 1e-9)
        _tail_ok(e, med_tol, out, "pdf")
    assert (r1["sampled"] == r2["sampled"]).mean() > 0.999
    s = (r1["sampled"] == 1) & (r2["sampled"] == 1)
    # lobe selection (w < spec_prob, w <= F) can flip for a query within fp32 of the threshold
    same_lobe = s & (np.abs(r1["s_eta"] - r2["s_eta"]) < 1e-3) & (np.abs(r1["s_roughness"] - r2["s_roughness"]) < 1e-3)
    assert same_lobe.sum() >= 0.999 * s.sum()
    if same_lobe.any():
        e = np.abs(r1["s_dir_out"][same_lobe].astype(np.float64) - r2["s_dir_out"][same_lobe]).max(axis=1)
        out["dir_med"], out["dir_max"] = float(np.median(e)), float(e.max())
        assert np.median(e) <= med_tol and np.quantile(e, 0.99) <= 5e-3, out
    return out


def check_light_parity(scene, ref, ref_points, seed=5):
    rng = np.random.default_rng(seed)
    n = ref_points.shape[0]
    q = np.zeros(n, dtype=lj.LIGHT_QUERY_DTYPE)
    q["ref_point"] = ref_points
    q["rnd_uv"] = rng.random((n, 2))
    q["rnd_w"] = rng.random(n)
    q["light_w"] = rng.random(n)
    r1, r2 = scene.sample_lights(q), ref.sample_lights(q)
    same = r1["light_id"] == r2["light_id"]
    assert same.mean() > 0.999  # a light_w within fp32 of a cdf entry may pick the neighbour
    scale = max(np.abs(r2["position"]).max(), 1e-3)
    a, b = r1[same], r2[same]
    perr = np.abs(a["position"] - b["position"]).max(axis=1) / scale
    # cone sampling of a sphere: sin(alpha) = sqrt(1 - cos^2) costs half the fp32 digits for points facing the
    # reference point, a tangential error of ~3e-4 radii at worst
    assert np.median(perr) < 1e-5 and np.quantile(perr, 0.999) < 3e-4, (np.median(perr), perr.max())
    assert np.median(np.abs(a["normal"] - b["normal"]).max(axis=1)) < 1e-5
    assert np.median(rel_err(a["pmf"], b["pmf"])) < 1e-6
    # a reference point ON the light can be handed its own position back: dir = normalize(0), the reference's pdf is
    # then 1e13..1e17 garbage and the device's NaN (both rejected by the estimators' p1 > 0 / G > 0 tests)
    sep = np.linalg.norm(b["position"].astype(np.float64) - q["ref_point"][same], axis=1) > 1e-4 * scale
    pz = (b["pdf"] > 0) & sep
    e = rel_err(a["pdf"][pz], b["pdf"][pz], 1e-12)
    assert np.median(e) < 1e-5 and np.quantile(e, 0.99) < 5e-3, (np.median(e), e.max())
    # one-sided emitters (diffuse_area_light.inl:16): leave out grazing queries whose cosine is fp32 noise
    d = b["position"].astype(np.float64) - q["ref_point"][same]
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-30)
    grazing = (np.abs((d * b["normal"]).sum(axis=1)) < 1e-4) | ~sep
    ez = np.abs(b["emission"]).max(axis=1) > 0
    agree_zero = ((np.abs(a["emission"]).max(axis=1) > 0) == ez)[~grazing].mean()
    assert agree_zero > 0.999
    both = ez & (np.abs(a["emission"]).max(axis=1) > 0)
    if both.any():
        assert np.median(rel_err(a["emission"][both], b["emission"][both])) < 1e-5
    return dict(pos_med=float(np.median(perr)), pdf_med=float(np.median(e)), n=int(n))


def check_camera_parity(scene, ref, n=4096, seed=6):
    rng = np.random.default_rng(seed)
    xy = rng.random((n, 2)).astype(np.float32)
    r1, r2 = scene.sample_primary(xy), ref.sample_primary(xy)
    assert np.abs(r1["org"] - r2["org"]).max() <= 1e-5 * max(np.abs(r2["org"]).max(), 1)
    # (x*w - floor(x*w)) in fp32 loses log2(w) bits before the filter warp: 1e-4 absolute on unit directions,
    # 1e-5 on the median
    d = np.abs(r1["dir"] - r2["dir"]).max(axis=1)
    assert np.median(d) < 1e-5 and d.max() < 5e-4, (np.median(d), d.max())
    return float(d.max())


def check_light_table(scene, ref):
    pmf1, cdf1 = scene.light_table()
    pmf2, cdf2 = ref.light_table()
    assert np.allclose(pmf1, pmf2, rtol=1e-5, atol=1e-7)
    assert np.allclose(cdf1[:-1], cdf2[:-1], rtol=1e-5, atol=1e-7)


def check_scene_info(scene, ref):
    i, r = scene.info(), ref.info()
    assert abs(i.bsphere_radius - r["radius"]) <= 1e-6 * r["radius"]
    assert np.allclose(list(i.bsphere_center), r["center"], rtol=1e-6, atol=1e-6 * r["radius"])
    assert abs(i.shadow_epsilon - r["eps"]) <= 1e-6 * r["eps"]


def image_stats(img, ref_img, var=None, ref_var=None, block=8):
    """relMSE and, when both variance-of-the-mean buffers exist, two z statistics of img - ref_img:
    per pixel channel (|z| > 3 fraction) and per `block` x `block` tile mean (|z| > 4 fraction; averaging 64
    pixels brings the heavy-tailed light-transport noise close to Gaussian, so this one detects small biases).
    A floor of (1e-4 * value)^2 on the variance keeps constant pixels (directly visible emitters) from turning
    fp32-vs-fp64 round-off into infinite z."""
    a = img.astype(np.float64)
    b = ref_img.astype(np.float64)
    d = a - b
    out = dict(relmse=float(np.mean(d ** 2 / (b ** 2 + 1e-2))), mean=a.mean(axis=(0, 1)).tolist(), ref_mean=b.mean(axis=(0, 1)).tolist())
    if var is not None and ref_var is not None:
        v = var.astype(np.float64) + ref_var.astype(np.float64) + (1e-4 * np.maximum(np.abs(a), np.abs(b))) ** 2 + 1e-20
        z = d / np.sqrt(v)
        out["frac_z_gt_3"] = float((np.abs(z) > 3).mean())
        h, w = (a.shape[0] // block) * block, (a.shape[1] // block) * block
        db = d[:h, :w].reshape(h // block, block, w // block, block, 3).mean(axis=(1, 3))
        vb = v[:h, :w].reshape(h // block, block, w // block, block, 3).sum(axis=(1, 3)) / block ** 4
        zb = db / np.sqrt(vb)
        out["block_frac_z_gt_4"] = float((np.abs(zb) > 4).mean())
        out["block_z_rms"] = float(np.sqrt(np.mean(np.clip(zb, -6, 6) ** 2)))  # clipped: one zero-variance tile must not decide
    return out


def check_block_majorants(scene, n=4000, seed=9, max_blocks=400):
    """The block-wise majorants of the tracking loops (lj_media.h, lj_medium_bound_batch) are a valid bound: along random
    rays through every grid medium, walked block by block the way flight_step / ratio_step walk them,
      * t_exit is strictly ahead of t and the walk leaves the grid box after a bounded number of blocks,
      * sigma_t at random points of [t, t_exit) never exceeds the block's majorant (all three channels),
      * the majorant never exceeds the medium's global one (get_majorant, medium.cpp:27-29)."""
    rng = np.random.default_rng(seed)
    out = {}
    for mid, m in enumerate(scene.desc.media):
        if not m.density.is_grid:
            continue
        lo, hi = np.array(m.density.p_min, dtype=np.float64), np.array(m.density.p_max, dtype=np.float64)
        org = (lo + (hi - lo) * rng.uniform(-0.3, 1.3, (n, 3))).astype(np.float32)
        tgt = lo + (hi - lo) * rng.uniform(0.0, 1.0, (n, 3))
        d = tgt - org
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        q = np.zeros(n, dtype=lj.MEDIUM_QUERY_DTYPE)
        q["org"], q["dir"], q["tfar"], q["medium_id"] = org, d, np.inf, mid
        glob = scene.medium(q)["majorant"]
        t = np.zeros(n, dtype=np.float32)
        alive = np.ones(n, dtype=bool)
        worst, blocks, inside_blocks = 0.0, 0, 0
        for _ in range(max_blocks):
            if not alive.any():
                break
            qa = q[alive].copy()
            qa["t"] = t[alive]
            b = scene.medium_bounds(qa)
            assert np.all(b["local"] == 1)
            assert np.all(b["t_exit"] > qa["t"]), "a block step that does not advance"
            assert np.all(b["majorant"] <= glob[alive] * (1 + 1e-5) + 1e-12), "block majorant above the global one"
            fin = np.isfinite(b["t_exit"])
            # sigma_t at random points of the block's stretch of the ray
            for _k in range(3):
                u = rng.random(len(qa)).astype(np.float32)
                qs = qa[fin].copy()
                qs["t"] = qa["t"][fin] + u[fin] * (b["t_exit"][fin] - qa["t"][fin]) * np.float32(0.999)
                bs = scene.medium_bounds(qs)
                excess = (bs["sigma_t"] - b["majorant"][fin]).max() if len(qs) else 0.0
                worst = max(worst, float(excess))
            inside_blocks += int((b["majorant"].max(axis=1) > 0).sum())
            blocks += len(qa)
            idx = np.nonzero(alive)[0]
            t[idx] = b["t_exit"]
            alive[idx[~fin]] = False
        assert not alive.any(), f"{int(alive.sum())} rays still inside the grid after {max_blocks} blocks"
        assert worst <= 0.0, f"sigma_t exceeds the block majorant by {worst}"
        out[f"medium{mid}"] = dict(rays=int(n), block_steps=int(blocks), nonempty_block_steps=int(inside_blocks))
    return out


def check_medium_parity(scene, ref, n=20000, seed=6):
    """get_majorant / get_sigma_a / get_sigma_s (medium.cpp:27-37, volume.h:45-81,125-144) and the phase function
    (phase_functions/*.inl) on points in and around each medium's grid box."""
    rng = np.random.default_rng(seed)
    out = {}
    for mid in range(ref.num_media()):
        m = scene.desc.media[mid]
        lo, hi = np.array(m.density.p_min, dtype=np.float64), np.array(m.density.p_max, dtype=np.float64)
        if not m.density.is_grid:
            lo, hi = np.full(3, -2.0), np.full(3, 2.0)
        q = np.zeros(n, dtype=lj.MEDIUM_QUERY_DTYPE)
        q["org"] = lo + (hi - lo) * rng.uniform(-0.2, 1.2, (n, 3))
        d = rng.normal(size=(n, 3))
        q["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True)
        q["tfar"] = np.where(rng.random(n) < 0.5, np.inf, rng.uniform(0, 3, n))
        q["t"] = rng.uniform(0, 0.3, n)
        q["rnd"] = rng.random((n, 2))
        q["medium_id"] = mid
        a, b = scene.medium(q), ref.medium(q)
        # the slab test decides on fp32-rounded box distances: a ray grazing the box can flip
        assert (np.abs(a["majorant"] - b["majorant"]).max(axis=1) <= 1e-6 * np.abs(b["majorant"]).max()).mean() > 0.999
        scale = max(np.abs(b["sigma_s"]).max() + np.abs(b["sigma_a"]).max(), 1e-6)
        # trilinear weights are fp32 fractions of grid coordinates up to 128: 1e-5 of the medium's peak density
        assert np.abs(a["sigma_s"] - b["sigma_s"]).max() < 2e-5 * scale and np.abs(a["sigma_a"] - b["sigma_a"]).max() < 2e-5 * scale
        assert np.abs(a["phase_dir"] - b["phase_dir"]).max() < 1e-5
        assert np.median(rel_err(a["phase_eval"], b["phase_eval"])) < 1e-6 and rel_err(a["phase_eval"], b["phase_eval"]).max() < 1e-4
        assert rel_err(a["phase_pdf"], b["phase_pdf"]).max() < 1e-4
        out[mid] = dict(inside=float((np.abs(b["sigma_s"]).max(axis=1) > 0).mean()))
    return out


AUX_INTEGRATORS = {"depth": 0, "shading_normal": 1, "mean_curvature": 2, "ray_differential": 3, "mipmap_level": 4}


def check_aux_parity(make_scene, ref, desc):
    """The auxiliary integrators (render.cpp:12-69) are deterministic -- one ray through each pixel centre -- so the
    device image is compared with the reference's pixel by pixel.  `make_scene(desc)` builds the scene under test."""
    import copy
    out = {}
    for name, code in AUX_INTEGRATORS.items():
        d = copy.copy(desc)
        d.options = copy.copy(desc.options)
        d.options.integrator = code
        sc = make_scene(d)
        img = sc.render()
        sc.close()
        ref.set_integrator(code)
        ref_img, _ = ref.render()
        a, b = img.astype(np.float64), ref_img.astype(np.float64)
        hit_a, hit_b = np.abs(a).sum(axis=2) > 0, np.abs(b).sum(axis=2) > 0
        if name == "depth":
            assert (hit_a == hit_b).mean() > 0.9995   # silhouette pixels may fall on either side
            both = hit_a & hit_b
            e = np.abs(a - b)[both] / np.maximum(b[both], 1e-6)
            assert np.quantile(e, 0.999) < 1e-4, np.quantile(e, [0.5, 0.999, 1])  # a pixel straddling an edge sees another surface
            out[name] = float(np.median(e))
        elif name == "shading_normal":
            e = np.abs(a - b).max(axis=2)
            assert np.median(e) < 1e-5 and np.quantile(e, 0.995) < 2e-3, np.quantile(e, [0.5, 0.995, 1])
            out[name] = float(np.median(e))
        elif name == "mean_curvature":
            fa, fb = np.isfinite(a), np.isfinite(b)
            assert (fa == fb).mean() > 0.98
            ok = fa & fb & (np.abs(b) < 1e3)
            e = np.abs(a - b)[ok] / np.maximum(np.abs(b[ok]), 1e-3)
            assert np.median(e) < 1e-4 and np.quantile(e, 0.99) < 5e-2, np.quantile(e, [0.5, 0.99])
            out[name] = float(np.median(e))
        elif name == "ray_differential":
            assert np.abs(a - b).max() < 1e-9 + 1e-6 * np.abs(b).max()
            out[name] = 0.0
        else:  # mip level = log2(footprint): absolute
            both = hit_a & hit_b
            assert (hit_a == hit_b).mean() > 0.995
            e = np.abs(a - b)[both]
            if e.size:
                assert np.median(e) < 1e-4 and np.quantile(e, 0.99) < 5e-2, np.quantile(e, [0.5, 0.99, 1])
                out[name] = float(np.median(e))
    ref.set_integrator(desc.options.integrator)
    return out
