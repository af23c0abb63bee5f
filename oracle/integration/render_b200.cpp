// render_b200.cpp -- the binding a lajolla maintainer adds to put libljb200.so behind render().
// Image3 render_b200(const Scene&) has render()'s signature (render.h:9): it flattens the parsed Scene into the C ABI's
// lj_scene_desc, calls lj_scene_create + lj_render, and returns the Image3.  oracle/Makefile compiles this file against
// the reference's own headers and links it with the reference's UNMODIFIED parser and main.cpp (compiled with
// -Drender=render_b200) into oracle/_ref/lajolla_b200_ref: the reference program with only its hot path swapped.
// Lives under oracle/ because it includes reference headers; it is an integration example, not part of the product.
#include <cstring>
#include <functional>
#include <vector>

#include "flexception.h"
#include "image.h"
#include "scene.h"

#include "lajolla_b200.h"

namespace {

struct Flat {
    lj_scene_desc desc;
    std::vector<lj_image_desc> images;
    std::vector<lj_material_desc> materials;
    std::vector<lj_shape_desc> shapes;
    std::vector<lj_light_desc> lights;
    std::vector<lj_medium_desc> media;
    std::vector<std::vector<float>> floats;
    std::vector<std::vector<int32_t>> ints;
    int n3 = 0;  // 1-channel images follow the 3-channel ones in the flat numbering
};

void put16(float dst[16], const Matrix4x4 &m) {
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) dst[4 * r + c] = (float)m(r, c);
}
void put3(float dst[3], const Vector3 &v) { dst[0] = (float)v.x; dst[1] = (float)v.y; dst[2] = (float)v.z; }
void put3(float dst[3], Real v) { dst[0] = dst[1] = dst[2] = (float)v; }

lj_texture_desc empty_texture() {
    lj_texture_desc t;
    memset(&t, 0, sizeof(t));
    t.kind = LJ_TEX_CONSTANT; t.image_id = -1; t.uscale = t.vscale = 1;
    return t;
}
template <typename T>
lj_texture_desc texture(const Texture<T> &tex, const Flat &f) {  // texture.h:84-113
    lj_texture_desc t = empty_texture();
    if (auto *c = std::get_if<ConstantTexture<T>>(&tex)) {
        put3(t.value, c->value); put3(t.color1, c->value);
    } else if (auto *im = std::get_if<ImageTexture<T>>(&tex)) {
        t.kind = LJ_TEX_IMAGE;
        t.image_id = im->texture_id + (std::is_same<T, Real>::value ? f.n3 : 0);
        t.uscale = (float)im->uscale; t.vscale = (float)im->vscale; t.uoffset = (float)im->uoffset; t.voffset = (float)im->voffset;
    } else if (auto *cb = std::get_if<CheckerboardTexture<T>>(&tex)) {
        t.kind = LJ_TEX_CHECKERBOARD;
        put3(t.value, cb->color0); put3(t.color1, cb->color1);
        t.uscale = (float)cb->uscale; t.vscale = (float)cb->vscale; t.uoffset = (float)cb->uoffset; t.voffset = (float)cb->voffset;
    }
    return t;
}

struct MaterialVisitor {  // material.h:10-110 -> texture slots LJ_SLOT_* of lajolla_b200.h
    const Flat &f;
    lj_material_desc blank(int type, Real eta) const {
        lj_material_desc m;
        memset(&m, 0, sizeof(m));
        m.type = type; m.eta = (float)eta;
        for (auto &t : m.tex) t = empty_texture();
        return m;
    }
    lj_material_desc operator()(const Lambertian &x) const { auto m = blank(LJ_MAT_LAMBERTIAN, 1); m.tex[LJ_SLOT_REFLECTANCE] = texture(x.reflectance, f); return m; }
    lj_material_desc operator()(const RoughPlastic &x) const {
        auto m = blank(LJ_MAT_ROUGHPLASTIC, x.eta);
        m.tex[LJ_SLOT_DIFFUSE_REFLECTANCE] = texture(x.diffuse_reflectance, f); m.tex[LJ_SLOT_SPECULAR_REFLECTANCE] = texture(x.specular_reflectance, f);
        m.tex[LJ_SLOT_ROUGHNESS] = texture(x.roughness, f);
        return m;
    }
    lj_material_desc operator()(const RoughDielectric &x) const {
        auto m = blank(LJ_MAT_ROUGHDIELECTRIC, x.eta);
        m.tex[LJ_SLOT_SPECULAR_TRANSMITTANCE] = texture(x.specular_transmittance, f); m.tex[LJ_SLOT_SPECULAR_REFLECTANCE] = texture(x.specular_reflectance, f);
        m.tex[LJ_SLOT_ROUGHNESS] = texture(x.roughness, f);
        return m;
    }
    lj_material_desc operator()(const DisneyDiffuse &x) const {
        auto m = blank(LJ_MAT_DISNEY_DIFFUSE, 1);
        m.tex[LJ_SLOT_BASE_COLOR] = texture(x.base_color, f); m.tex[LJ_SLOT_SUBSURFACE] = texture(x.subsurface, f); m.tex[LJ_SLOT_ROUGHNESS] = texture(x.roughness, f);
        return m;
    }
    lj_material_desc operator()(const DisneyMetal &x) const {
        auto m = blank(LJ_MAT_DISNEY_METAL, 1);
        m.tex[LJ_SLOT_BASE_COLOR] = texture(x.base_color, f); m.tex[LJ_SLOT_ROUGHNESS] = texture(x.roughness, f); m.tex[LJ_SLOT_ANISOTROPIC] = texture(x.anisotropic, f);
        return m;
    }
    lj_material_desc operator()(const DisneyGlass &x) const {
        auto m = blank(LJ_MAT_DISNEY_GLASS, x.eta);
        m.tex[LJ_SLOT_BASE_COLOR] = texture(x.base_color, f); m.tex[LJ_SLOT_ROUGHNESS] = texture(x.roughness, f); m.tex[LJ_SLOT_ANISOTROPIC] = texture(x.anisotropic, f);
        return m;
    }
    lj_material_desc operator()(const DisneyClearcoat &x) const { auto m = blank(LJ_MAT_DISNEY_CLEARCOAT, 1); m.tex[LJ_SLOT_CLEARCOAT_GLOSS] = texture(x.clearcoat_gloss, f); return m; }
    lj_material_desc operator()(const DisneySheen &x) const {
        auto m = blank(LJ_MAT_DISNEY_SHEEN, 1);
        m.tex[LJ_SLOT_BASE_COLOR] = texture(x.base_color, f); m.tex[LJ_SLOT_SHEEN_TINT] = texture(x.sheen_tint, f);
        return m;
    }
    lj_material_desc operator()(const DisneyBSDF &x) const {
        auto m = blank(LJ_MAT_DISNEY_BSDF, x.eta);
        m.tex[LJ_SLOT_BASE_COLOR] = texture(x.base_color, f); m.tex[LJ_SLOT_SUBSURFACE] = texture(x.subsurface, f);
        m.tex[LJ_SLOT_ROUGHNESS] = texture(x.roughness, f); m.tex[LJ_SLOT_ANISOTROPIC] = texture(x.anisotropic, f);
        m.tex[LJ_SLOT_CLEARCOAT_GLOSS] = texture(x.clearcoat_gloss, f); m.tex[LJ_SLOT_SHEEN_TINT] = texture(x.sheen_tint, f);
        m.tex[LJ_SLOT_SPECULAR_TRANSMISSION] = texture(x.specular_transmission, f); m.tex[LJ_SLOT_METALLIC] = texture(x.metallic, f);
        m.tex[LJ_SLOT_SPECULAR] = texture(x.specular, f); m.tex[LJ_SLOT_SPECULAR_TINT] = texture(x.specular_tint, f);
        m.tex[LJ_SLOT_SHEEN] = texture(x.sheen, f); m.tex[LJ_SLOT_CLEARCOAT] = texture(x.clearcoat, f);
        return m;
    }
};

lj_volume_desc volume(const VolumeSpectrum &v, Flat &f) {  // volume.h
    lj_volume_desc o;
    memset(&o, 0, sizeof(o));
    o.scale = 1;
    if (auto *c = std::get_if<ConstantVolume<Spectrum>>(&v)) {
        put3(o.value, c->value);
    } else {
        const GridVolume<Spectrum> &g = std::get<GridVolume<Spectrum>>(v);
        o.is_grid = 1;
        o.res[0] = g.resolution.x; o.res[1] = g.resolution.y; o.res[2] = g.resolution.z;
        put3(o.value, g.max_data); put3(o.p_min, g.p_min); put3(o.p_max, g.p_max);
        o.scale = (float)g.scale;
        std::vector<float> data(3 * g.data.size());
        for (size_t i = 0; i < g.data.size(); i++) { data[3 * i] = (float)g.data[i].x; data[3 * i + 1] = (float)g.data[i].y; data[3 * i + 2] = (float)g.data[i].z; }
        f.floats.push_back(std::move(data));
        o.data = f.floats.back().data();
    }
    return o;
}

void flatten(const Scene &s, Flat &f) {
    memset(&f.desc, 0, sizeof(f.desc));
    f.floats.reserve(s.texture_pool.image3s.size() + s.texture_pool.image1s.size() + 3 * s.shapes.size() + 2 * s.media.size() + 4);
    f.ints.reserve(s.shapes.size() + 1);
    lj_scene_desc &d = f.desc;
    const Camera &c = s.camera;  // camera.h:10-24
    put16(d.camera.cam_to_world, c.cam_to_world); put16(d.camera.world_to_cam, c.world_to_cam);
    put16(d.camera.sample_to_cam, c.sample_to_cam); put16(d.camera.cam_to_sample, c.cam_to_sample);
    d.camera.width = c.width; d.camera.height = c.height; d.camera.medium_id = c.medium_id;
    if (auto *b = std::get_if<Box>(&c.filter)) { d.camera.filter_type = LJ_FILTER_BOX; d.camera.filter_param = (float)b->width; }
    else if (auto *t = std::get_if<Tent>(&c.filter)) { d.camera.filter_type = LJ_FILTER_TENT; d.camera.filter_param = (float)t->width; }
    else { d.camera.filter_type = LJ_FILTER_GAUSSIAN; d.camera.filter_param = (float)std::get<Gaussian>(c.filter).stddev; }
    d.options.integrator = (int)s.options.integrator;  // scene.h:14-31: the enum values are LJ_INT_*
    d.options.samples_per_pixel = s.options.samples_per_pixel; d.options.max_depth = s.options.max_depth; d.options.rr_depth = s.options.rr_depth;
    d.options.vol_path_version = s.options.vol_path_version; d.options.max_null_collisions = s.options.max_null_collisions;
    f.n3 = (int)s.texture_pool.image3s.size();
    for (const Mipmap3 &m : s.texture_pool.image3s) {  // mip level 0 only: the library rebuilds the chain on the device
        const Image3 &im = m.images[0];
        std::vector<float> px(3 * im.data.size());
        for (size_t i = 0; i < im.data.size(); i++) { px[3 * i] = (float)im.data[i].x; px[3 * i + 1] = (float)im.data[i].y; px[3 * i + 2] = (float)im.data[i].z; }
        f.floats.push_back(std::move(px));
        f.images.push_back(lj_image_desc{im.width, im.height, 3, 0, f.floats.back().data()});
    }
    for (const Mipmap1 &m : s.texture_pool.image1s) {
        const Image1 &im = m.images[0];
        f.floats.push_back(std::vector<float>(im.data.begin(), im.data.end()));
        f.images.push_back(lj_image_desc{im.width, im.height, 1, 0, f.floats.back().data()});
    }
    for (const Material &m : s.materials) f.materials.push_back(std::visit(MaterialVisitor{f}, m));
    for (const Shape &sh : s.shapes) {  // shape.h:26-54
        lj_shape_desc o;
        memset(&o, 0, sizeof(o));
        std::visit([&](const auto &x) { o.material_id = x.material_id; o.area_light_id = x.area_light_id; o.interior_medium_id = x.interior_medium_id; o.exterior_medium_id = x.exterior_medium_id; }, sh);
        if (auto *sp = std::get_if<Sphere>(&sh)) {
            o.type = LJ_SHAPE_SPHERE; put3(o.center, sp->position); o.radius = (float)sp->radius;
        } else {
            const TriangleMesh &m = std::get<TriangleMesh>(sh);
            o.type = LJ_SHAPE_MESH; o.num_vertices = (int)m.positions.size(); o.num_triangles = (int)m.indices.size();
            auto flat3 = [&](const std::vector<Vector3> &v) { std::vector<float> a(3 * v.size()); for (size_t i = 0; i < v.size(); i++) { a[3 * i] = (float)v[i].x; a[3 * i + 1] = (float)v[i].y; a[3 * i + 2] = (float)v[i].z; } f.floats.push_back(std::move(a)); return f.floats.back().data(); };
            o.positions = flat3(m.positions);
            std::vector<int32_t> idx(3 * m.indices.size());
            for (size_t i = 0; i < m.indices.size(); i++) { idx[3 * i] = m.indices[i][0]; idx[3 * i + 1] = m.indices[i][1]; idx[3 * i + 2] = m.indices[i][2]; }
            f.ints.push_back(std::move(idx));
            o.indices = f.ints.back().data();
            if (!m.normals.empty()) o.normals = flat3(m.normals);
            if (!m.uvs.empty()) { std::vector<float> a(2 * m.uvs.size()); for (size_t i = 0; i < m.uvs.size(); i++) { a[2 * i] = (float)m.uvs[i].x; a[2 * i + 1] = (float)m.uvs[i].y; } f.floats.push_back(std::move(a)); o.uvs = f.floats.back().data(); }
        }
        f.shapes.push_back(o);
    }
    for (const Light &l : s.lights) {  // light.h:14-27
        lj_light_desc o;
        memset(&o, 0, sizeof(o));
        o.values = empty_texture();
        Matrix4x4 id = Matrix4x4::identity();
        put16(o.to_world, id); put16(o.to_local, id);
        o.scale = 1;
        if (auto *a = std::get_if<DiffuseAreaLight>(&l)) { o.type = LJ_LIGHT_AREA; o.shape_id = a->shape_id; put3(o.intensity, a->intensity); }
        else { const Envmap &e = std::get<Envmap>(l); o.type = LJ_LIGHT_ENVMAP; o.shape_id = -1; o.values = texture(e.values, f); put16(o.to_world, e.to_world); put16(o.to_local, e.to_local); o.scale = (float)e.scale; }
        f.lights.push_back(o);
    }
    for (const Medium &m : s.media) {  // medium.h, phase_function.h
        lj_medium_desc o;
        memset(&o, 0, sizeof(o));
        PhaseFunction pf = get_phase_function(m);
        if (auto *hg = std::get_if<HenyeyGreenstein>(&pf)) { o.phase_type = LJ_PHASE_HG; o.phase_g = (float)hg->g; }
        if (auto *hm = std::get_if<HomogeneousMedium>(&m)) { o.type = LJ_MEDIUM_HOMOGENEOUS; put3(o.sigma_a, hm->sigma_a); put3(o.sigma_s, hm->sigma_s); }
        else { const HeterogeneousMedium &het = std::get<HeterogeneousMedium>(m); o.type = LJ_MEDIUM_HETEROGENEOUS; o.albedo = volume(het.albedo, f); o.density = volume(het.density, f); }
        f.media.push_back(o);
    }
    d.num_images = (int)f.images.size(); d.num_materials = (int)f.materials.size(); d.num_shapes = (int)f.shapes.size();
    d.num_lights = (int)f.lights.size(); d.num_media = (int)f.media.size(); d.envmap_light_id = s.envmap_light_id;
    d.images = f.images.data(); d.materials = f.materials.data(); d.shapes = f.shapes.data(); d.lights = f.lights.data(); d.media = f.media.data();
}

}  // namespace

Image3 render_b200(const Scene &scene) {
    Flat flat;
    flatten(scene, flat);
    if (lj_init(nullptr, 1) != LJ_OK) Error(lj_last_error());  // flexception.h:8-24
    lj_scene *dev = nullptr;
    if (lj_scene_create(&flat.desc, &dev) != LJ_OK) Error(lj_last_error());
    const int w = scene.camera.width, h = scene.camera.height;
    std::vector<float> rgb((size_t)w * h * 3);
    lj_render_opts opts;
    memset(&opts, 0, sizeof(opts));  // spp 0: the scene's sampleCount; one GPU
    opts.normalize = 1;              // render.cpp:94 divides by spp
    opts.num_gpus = 1;
    lj_stats stats;
    int rc = lj_render(dev, &opts, rgb.data(), &stats);
    lj_scene_destroy(dev);
    if (rc != LJ_OK) Error(lj_last_error());
    Image3 img(w, h);  // image.h:13-39: row-major, top row first
    for (int i = 0; i < w * h; i++) img.data[i] = Vector3{rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
    return img;
}
