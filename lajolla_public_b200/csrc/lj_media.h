// lj_media.h -- participating media and phase functions in device memory.
// fp32 restatement of the reference's medium.cpp:27-37, media/homogeneous.inl:1-11,
// media/heterogeneous.inl:3-21, volume.h:45-81 (trilinear GridVolume lookup), :125-144 (slab test),
// phase_functions/isotropic.inl:1-15, phase_functions/henyeygreenstein.inl:3-48, and of the two
// tracking loops of the course handout the public reference leaves as homework
// (handouts/homework2.tex:713-758 chromatic delta tracking, :771-810 ratio tracking; our CPU
// restatement of the same text is oracle/overlay/hw_vol_path_tracing.h).
#pragma once
#include "lj_pcg.h"
#include "lj_scene_dev.h"

namespace lj {

// volume.h:45-81.  Zero outside the grid box, (res - 1) scaling, int() truncation, scale applied last.
LJ_HD V3 volume_lookup(const DevVolume &v, V3 p) {
    if (!v.is_grid) return mk3(v.value[0], v.value[1], v.value[2]);
    float px = (p.x - v.p_min[0]) / (v.p_max[0] - v.p_min[0]);
    float py = (p.y - v.p_min[1]) / (v.p_max[1] - v.p_min[1]);
    float pz = (p.z - v.p_min[2]) / (v.p_max[2] - v.p_min[2]);
    if (px < 0 || px > 1 || py < 0 || py > 1 || pz < 0 || pz > 1) return mk3(0);
    const int nx = v.res[0], ny = v.res[1], nz = v.res[2];
    px *= (float)(nx - 1); py *= (float)(ny - 1); pz *= (float)(nz - 1);
    int x0 = clampi((int)px, 0, nx - 1), y0 = clampi((int)py, 0, ny - 1), z0 = clampi((int)pz, 0, nz - 1);
    int x1 = clampi(x0 + 1, 0, nx - 1), y1 = clampi(y0 + 1, 0, ny - 1), z1 = clampi(z0 + 1, 0, nz - 1);
    float dx = px - x0, dy = py - y0, dz = pz - z0;
    const V4 *d = v.data;
    V3 v000 = xyz(ld4(&d[(z0 * ny + y0) * nx + x0])), v001 = xyz(ld4(&d[(z0 * ny + y0) * nx + x1]));
    V3 v010 = xyz(ld4(&d[(z0 * ny + y1) * nx + x0])), v011 = xyz(ld4(&d[(z0 * ny + y1) * nx + x1]));
    V3 v100 = xyz(ld4(&d[(z1 * ny + y0) * nx + x0])), v101 = xyz(ld4(&d[(z1 * ny + y0) * nx + x1]));
    V3 v110 = xyz(ld4(&d[(z1 * ny + y1) * nx + x0])), v111 = xyz(ld4(&d[(z1 * ny + y1) * nx + x1]));
    V3 r = v000 * ((1 - dx) * (1 - dy) * (1 - dz)) + v001 * (dx * (1 - dy) * (1 - dz)) +
           v010 * ((1 - dx) * dy * (1 - dz)) + v011 * (dx * dy * (1 - dz)) +
           v100 * ((1 - dx) * (1 - dy) * dz) + v101 * (dx * (1 - dy) * dz) +
           v110 * ((1 - dx) * dy * dz) + v111 * (dx * dy * dz);
    return r * v.scale;
}

// volume.h:125-144: does [0, tfar] of the ray overlap the grid box (a constant volume is everywhere)
LJ_HD bool volume_intersect(const DevVolume &v, V3 o, V3 d, float tfar) {
    if (!v.is_grid) return true;
    float t0 = 0, t1 = tfar;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    for (int i = 0; i < 3; i++) {
        float tn = (v.p_min[i] - oo[i]) / dd[i];
        float tf = (v.p_max[i] - oo[i]) / dd[i];
        if (tn > tf) { float s = tn; tn = tf; tf = s; }
        t0 = tn > t0 ? tn : t0;
        t1 = tf < t1 ? tf : t1;
        if (t0 > t1) return false;
    }
    return true;
}

LJ_HD V3 volume_max(const DevVolume &v) {  // volume.h get_max_value
    if (!v.is_grid) return mk3(v.value[0], v.value[1], v.value[2]);
    return mk3(v.max_data[0], v.max_data[1], v.max_data[2]) * v.scale;
}

// medium.cpp:27-37
LJ_HD V3 medium_majorant(const DevMedium &m, V3 o, V3 d, float tfar) {
    if (m.type == 0) return mk3(m.sigma_a[0] + m.sigma_s[0], m.sigma_a[1] + m.sigma_s[1], m.sigma_a[2] + m.sigma_s[2]);
    return volume_intersect(m.density, o, d, tfar) ? volume_max(m.density) : mk3(0);
}
LJ_HD void medium_sigmas(const DevMedium &m, V3 p, V3 &sigma_a, V3 &sigma_s) {
    if (m.type == 0) {
        sigma_a = mk3(m.sigma_a[0], m.sigma_a[1], m.sigma_a[2]);
        sigma_s = mk3(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]);
        return;
    }
    V3 density = volume_lookup(m.density, p), albedo = volume_lookup(m.albedo, p);
    sigma_s = density * albedo;
    sigma_a = density * (mk3(1) - albedo);
}

// ---- phase functions.  dir_in points away from the scattering point (hence +2g cos, henyeygreenstein.inl:3-7)
LJ_HD float phase_eval(const DevMedium &m, V3 dir_in, V3 dir_out) {
    if (m.phase_type == 0) return kInvFourPi;
    float g = m.phase_g;
    float base = 1 + g * g + 2 * g * dot(dir_in, dir_out);
    return kInvFourPi * (1 - g * g) / (base * sqrtf(base));
}
LJ_HD float phase_pdf(const DevMedium &m, V3 dir_in, V3 dir_out) { return phase_eval(m, dir_in, dir_out); }
LJ_HD V3 phase_sample(const DevMedium &m, V3 dir_in, V2 u) {
    float g = m.phase_g;
    if (m.phase_type == 0 || fabsf(g) < 1e-3f) {
        float z = 1 - 2 * u.x;
        float r = sqrtf(fmaxf(0.f, 1 - z * z));
        float phi = 2 * kPi * u.y;
        return mk3(r * cosf(phi), r * sinf(phi), z);
    }
    float tmp = (g * g - 1) / (2 * u.x * g - (g + 1));
    float cos_el = (tmp * tmp - (1 + g * g)) / (2 * g);
    float sin_el = sqrtf(fmaxf(1 - cos_el * cos_el, 0.f));
    float az = 2 * kPi * u.y;
    Frame f = make_frame(dir_in);
    return to_world(f, mk3(sin_el * cosf(az), sin_el * sinf(az), cos_el));
}

// Both tracking loops multiply three running products (transmittance and the two pdfs) by exp(-majorant * t) at
// every step.  Only ratios of those products are ever consumed -- transmittance / avg(pdf), and the two pdfs
// against each other inside the MIS weight -- so the factor exp(-min(majorant) * t) common to all channels of all
// three cancels exactly.  It is dropped here: in fp32 exp(-100 * 1.4) underflows where the reference's double does
// not (hetvol: majorant 100 over an empty stretch of the grid), which would silently zero long segments.
LJ_HD V3 tracking_exp(V3 majorant_minus_min, float t) {
    return mk3(majorant_minus_min.x > 0 ? expf(-majorant_minus_min.x * t) : 1.f,
               majorant_minus_min.y > 0 ? expf(-majorant_minus_min.y * t) : 1.f,
               majorant_minus_min.z > 0 ? expf(-majorant_minus_min.z * t) : 1.f);
}

LJ_HD int tracking_channel(float u) { return clampi((int)(u * 3), 0, 2); }

// ---- local majorants ------------------------------------------------------------------------------------------
// The reference bounds a heterogeneous medium by ONE number, the maximum of the density grid (medium.cpp:27-29), so a
// ray through hetvol's box takes ~100 null collisions per segment although most of the box is empty.  The tracking
// loops below bound it piecewise instead: DevVolume::maj holds the maximum of the grid over blocks of maj_block^3
// voxel cells, a segment is tracked block by block with the block's own majorant, and a block that holds nothing is
// crossed in one step.  Delta / ratio tracking are exact for ANY bound of the density (memorylessness of the
// exponential: a tentative collision beyond the block face is discarded and redrawn in the next block), so the
// estimators of homework2.tex:713-810 keep their expectation -- what changes is the random sequence, not the image
// (tests: same mean as the global-majorant run, LJ_MAJ_BLOCK=0, and as the oracle's).  The three running products
// stay consistent because both loops evaluate the same majorant function of the position.
LJ_HD bool medium_has_local_majorant(const DevMedium &m) { return m.type != 0 && m.density.is_grid && m.density.maj != nullptr; }

// Majorant (before the common scale) of the block that holds the point o + d t, and the parameter t_exit > t at which
// the ray leaves it; outside the grid box: 0 and the parameter at which the ray enters the box (infinity if never).
LJ_HD V3 volume_block_majorant(const DevVolume &v, V3 o, V3 d, float t, float &t_exit) {
    const float hi[3] = {(float)(v.res[0] - 1), (float)(v.res[1] - 1), (float)(v.res[2] - 1)};
    const V3 p = o + d * t;
    const float pp[3] = {p.x, p.y, p.z}, dd[3] = {d.x, d.y, d.z};
    float q[3], dq[3], fastest = 0;
    bool inside = true;
    for (int i = 0; i < 3; i++) {  // voxel coordinates, as volume_lookup
        const float s = hi[i] / (v.p_max[i] - v.p_min[i]);
        q[i] = (pp[i] - v.p_min[i]) * s;
        dq[i] = dd[i] * s;
        fastest = fmaxf(fastest, fabsf(dq[i]));
        inside = inside && q[i] >= 0 && q[i] <= hi[i];
    }
    if (!(fastest > 0)) { t_exit = LJ_INF; return mk3(0); }
    const float nudge = 1e-3f / fastest;  // lands a thousandth of a voxel beyond the face (the blocks are dilated by a node)
    float t_in = 0;  // distance from the point to the grid box along the ray (0: the point is inside)
    if (!inside) {
        float t0 = 0, t1 = LJ_INF;
        for (int i = 0; i < 3; i++) {
            if (dq[i] != 0) {
                float tn = (0 - q[i]) / dq[i], tf = (hi[i] - q[i]) / dq[i];
                if (tn > tf) { float sw = tn; tn = tf; tf = sw; }
                t0 = fmaxf(t0, tn); t1 = fminf(t1, tf);
            } else if (q[i] < 0 || q[i] > hi[i]) {
                t1 = -1;
            }
        }
        if (!(t0 <= t1)) { t_exit = LJ_INF; return mk3(0); }  // the ray never enters the box
        // empty space up to a nudge short of the entry face (the bound 0 must not reach into the grid); from there on
        // the stretch belongs to the block the ray enters
        if (t0 > 2 * nudge) { t_exit = t + t0 - nudge; return mk3(0); }
        t_in = t0;
        for (int i = 0; i < 3; i++) q[i] = fminf(fmaxf(q[i] + dq[i] * t0, 0.f), hi[i]);
    }
    const float B = (float)v.maj_block;
    int cell[3];
    float te = LJ_INF;
    for (int i = 0; i < 3; i++) {
        cell[i] = clampi((int)(q[i] / B), 0, v.maj_res[i] - 1);
        const float lo_face = (float)cell[i] * B, hi_face = fminf((float)(cell[i] + 1) * B, hi[i]);
        if (dq[i] > 0) te = fminf(te, (hi_face - q[i]) / dq[i]);
        else if (dq[i] < 0) te = fminf(te, (lo_face - q[i]) / dq[i]);
    }
    t_exit = t + t_in + fmaxf(te, 0.f) + nudge;
    // One bound for the three channels (the largest): with per-channel block majorants a channel whose density is low
    // in this block is sampled with few tentative collisions while the other channels' weights keep their full range,
    // and the chromatic estimator's variance grows (measured on hetvol_colored: x1.5 .. x8, profiles/r02n_*); with a
    // common bound every weight (majorant - sigma_t) / majorant stays in [0, 1].
    return mk3(max3(xyz(ld4(&v.maj[(cell[2] * v.maj_res[1] + cell[1]) * v.maj_res[0] + cell[0]])))) * v.scale;
}

// sigma_t = sigma_a + sigma_s at p.  Heterogeneous: density * albedo + density * (1 - albedo) is the density itself
// (heterogeneous.inl:3-21); the tracking loops only need this sum, so they skip the albedo grid's eight voxels.
LJ_HD V3 medium_sigma_t(const DevMedium &m, V3 p) {
    if (m.type == 0) return mk3(m.sigma_a[0] + m.sigma_s[0], m.sigma_a[1] + m.sigma_s[1], m.sigma_a[2] + m.sigma_s[2]);
    return volume_lookup(m.density, p);
}
LJ_HD V3 safe_ratio(V3 a, V3 b) { return mk3(b.x > 0 ? a.x / b.x : 0.f, b.y > 0 ? a.y / b.y : 0.f, b.z > 0 ? a.z / b.z : 0.f); }

// The two tracking loops, cut into single steps so the persistent kernels (wavefront.cu k_flight, k_trace<2|3>) can
// run ONE step per lane per warp iteration and refill lanes whose segment ended, instead of letting 31 lanes wait
// for the longest loop (measured on hetvol: 4 of 32 lanes active otherwise).  A step is one tentative collision in
// the current majorant block, or the move into the next block.
struct TrackState {
    V3 majorant, maj_rel;  // maj_rel = majorant - min(majorant), see tracking_exp
    float maj_c, max_maj;  // majorant of the sampled channel, max over channels
    float accum_t;
    float t_block;         // the majorant holds up to this ray parameter (infinity: to the end of the segment)
    int channel, it;
};
LJ_HD void track_set_majorant(TrackState &ts, V3 majorant) {
    ts.majorant = majorant;
    ts.max_maj = max3(majorant);
    ts.maj_c = comp(majorant, ts.channel);
    ts.maj_rel = majorant - mk3(min3(majorant));
}
// Draws the channel.  Returns false when there is nothing to track (global majorant: the sampled channel has none).
// GRID = false: the caller knows the scene has no grid medium (the block logic is compiled out).
template <bool GRID = true>
LJ_HD bool track_begin(const DevMedium &m, V3 o, V3 d, float ray_tfar, Pcg &rng, TrackState &ts) {
    const bool local = GRID && medium_has_local_majorant(m);
    const V3 global = local ? mk3(0) : medium_majorant(m, o, d, ray_tfar);
    ts.channel = tracking_channel(pcg_uniform(rng));
    ts.accum_t = 0;
    ts.it = 0;
    if (local) {
        track_set_majorant(ts, volume_block_majorant(m.density, o, d, 0.f, ts.t_block));
        return true;
    }
    track_set_majorant(ts, global);
    ts.t_block = LJ_INF;
    return ts.maj_c > 0;
}
// the segment goes on beyond the current block: step into the next one (counted against the iteration limit, so a
// ray that cannot advance still ends)
LJ_HD void track_next_block(const DevMedium &m, V3 o, V3 d, TrackState &ts) {
    track_set_majorant(ts, volume_block_majorant(m.density, o, d, ts.accum_t, ts.t_block));
    ts.it++;
}

enum { kTrackContinue = 0, kTrackScatter = 1, kTrackEnd = 2 };

// homework2.tex:713-758: one step of the chromatic delta tracking free flight over [0, t_hit].
template <bool GRID = true>
LJ_HD int flight_step(const DevMedium &m, V3 o, V3 d, float t_hit, int max_null, Pcg &rng, TrackState &ts,
                      V3 &transmittance, V3 &trans_dir_pdf, V3 &trans_nee_pdf) {
    if (ts.it >= max_null) return kTrackEnd;
    const float t_end = GRID ? fminf(ts.t_block, t_hit) : t_hit;
    float t = LJ_INF;
    if (ts.maj_c > 0) t = -logf(1 - pcg_uniform(rng)) / ts.maj_c;
    float dt = t_end - ts.accum_t;
    ts.accum_t = fminf(ts.accum_t + t, t_end);
    if (t < dt) {
        V3 sigma_t = medium_sigma_t(m, o + d * ts.accum_t);
        V3 real_prob = safe_ratio(sigma_t, ts.majorant);
        V3 e = tracking_exp(ts.maj_rel, t);
        if (pcg_uniform(rng) < comp(real_prob, ts.channel)) {
            transmittance *= e / ts.max_maj;
            trans_dir_pdf *= e * ts.majorant * real_prob / ts.max_maj;
            return kTrackScatter;
        }
        transmittance *= e * (ts.majorant - sigma_t) / ts.max_maj;
        trans_dir_pdf *= e * ts.majorant * (mk3(1) - real_prob) / ts.max_maj;
        trans_nee_pdf *= e * ts.majorant / ts.max_maj;
        ts.it++;
        return kTrackContinue;
    }
    V3 e = tracking_exp(ts.maj_rel, dt);
    transmittance *= e;
    trans_dir_pdf *= e;
    trans_nee_pdf *= e;
    if (!GRID || ts.t_block >= t_hit) return kTrackEnd;
    track_next_block(m, o, d, ts);
    return kTrackContinue;
}

// homework2.tex:771-810: one step of ratio tracking over the shadow segment [0, next_t].
template <bool GRID = true>
LJ_HD int ratio_step(const DevMedium &m, V3 o, V3 d, float next_t, int max_null, Pcg &rng, TrackState &ts,
                     V3 &T_light, V3 &p_trans_nee, V3 &p_trans_dir) {
    if (ts.it >= max_null) return kTrackEnd;
    const float t_end = GRID ? fminf(ts.t_block, next_t) : next_t;
    float t = LJ_INF;
    if (ts.maj_c > 0) t = -logf(1 - pcg_uniform(rng)) / ts.maj_c;
    float dt = t_end - ts.accum_t;
    ts.accum_t = fminf(ts.accum_t + t, t_end);
    if (t < dt) {
        V3 sigma_t = medium_sigma_t(m, o + d * ts.accum_t);
        V3 real_prob = safe_ratio(sigma_t, ts.majorant);
        V3 e = tracking_exp(ts.maj_rel, t);
        T_light *= e * (ts.majorant - sigma_t) / ts.max_maj;
        p_trans_nee *= e * ts.majorant / ts.max_maj;
        p_trans_dir *= e * ts.majorant * (mk3(1) - real_prob) / ts.max_maj;
        if (max3(T_light) <= 0) return kTrackEnd;
        ts.it++;
        return kTrackContinue;
    }
    V3 e = tracking_exp(ts.maj_rel, dt);
    T_light *= e;
    p_trans_nee *= e;
    p_trans_dir *= e;
    if (!GRID || ts.t_block >= next_t) return kTrackEnd;
    track_next_block(m, o, d, ts);
    return kTrackContinue;
}

// Whole loops (query seam, host simulation, and anything that is not a persistent kernel).
// Returns true on a real collision at distance accum_t; the three running products are updated in place.
LJ_HD bool free_flight(const DevMedium &m, V3 o, V3 d, float ray_tfar, float t_hit, int max_null, Pcg &rng,
                       V3 &transmittance, V3 &trans_dir_pdf, V3 &trans_nee_pdf, float &accum_t) {
    TrackState ts;
    accum_t = 0;
    if (!track_begin(m, o, d, ray_tfar, rng, ts)) return false;
    int r;
    do { r = flight_step(m, o, d, t_hit, max_null, rng, ts, transmittance, trans_dir_pdf, trans_nee_pdf); } while (r == kTrackContinue);
    accum_t = ts.accum_t;
    return r == kTrackScatter;
}
LJ_HD void ratio_track(const DevMedium &m, V3 o, V3 d, float ray_tfar, float next_t, int max_null, Pcg &rng,
                       V3 &T_light, V3 &p_trans_nee, V3 &p_trans_dir) {
    TrackState ts;
    if (!track_begin(m, o, d, ray_tfar, rng, ts)) return;
    while (ratio_step(m, o, d, next_t, max_null, rng, ts, T_light, p_trans_nee, p_trans_dir) == kTrackContinue) {}
}

}  // namespace lj
