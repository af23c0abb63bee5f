// lj_host_scene.cpp -- Mitsuba-style XML -> HostScene -> lj_scene_desc / .ljs (see lj_host_scene.h).
// Behaviour follows the reference's parser (parsers/parse_scene.cpp; cited per function) including the quirks the
// renderer's results depend on: numbers go through float (std::stof), a single-value
// reflectance spectrum is white, alpha becomes roughness = sqrt(alpha), `direct` is a depth-2 path integrator, point
// and directional emitters become small area lights, meshes without normals get angle-weighted vertex normals.
#include "lj_host_scene.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <stdexcept>

#include "lj_image_io.h"
#include "lj_xml.h"

namespace ljhost {

namespace {

[[noreturn]] void fail(const std::string &what) { throw std::runtime_error(what); }

typedef std::map<std::string, std::string> DefaultMap;

struct ParsedTexture {  // parse_scene.cpp:27-33: images are only loaded when a material references the texture
    bool bitmap = true;
    std::string filename;
    Vec3 color0, color1;
    double uscale = 1, vscale = 1, uoffset = 0, voffset = 0;
};

struct Ctx {
    std::string dir;  // directory of the XML file
    DefaultMap defaults;
    std::map<std::string, ParsedTexture> textures;
    std::map<std::string, int> material_ids, medium_ids, image1_ids, image3_ids;
    int inline_spectrum = 0, inline_float = 0, inline_alpha = 0;
    HostScene scene;
    std::string path(const std::string &file) const { return (!file.empty() && file[0] == '/') || dir.empty() ? file : dir + "/" + file; }
};

// ---- value parsing (parse_scene.cpp:43-207).  "$name" is replaced from the <default> map.
const std::string &subst(const std::string &v, const DefaultMap &d) {
    if (!v.empty() && v[0] == '$') {
        auto it = d.find(v.substr(1));
        if (it == d.end()) fail("Reference default variable " + v + " not found.");
        return it->second;
    }
    return v;
}
double to_float(const std::string &s) {  // std::stof: float precision, leading blanks skipped, trailing text ignored
    const char *b = s.c_str();
    char *e = nullptr;
    float f = strtof(b, &e);
    if (e == b) fail("invalid number: '" + s + "'");
    return (double)f;
}
int to_int(const std::string &s) {
    const char *b = s.c_str();
    char *e = nullptr;
    long v = strtol(b, &e, 10);
    if (e == b) fail("invalid integer: '" + s + "'");
    return (int)v;
}
double parse_float(const std::string &v, const DefaultMap &d) { return to_float(subst(v, d)); }
int parse_integer(const std::string &v, const DefaultMap &d) { return to_int(subst(v, d)); }
bool parse_boolean(const std::string &v, const DefaultMap &d) {
    const std::string &s = subst(v, d);
    if (s == "true") return true;
    if (s == "false") return false;
    fail("parse_boolean failed");
}
std::vector<std::string> split_list(const std::string &s) {  // separators: runs of ',' and ' '
    std::vector<std::string> out;
    std::string cur;
    for (char c : s) {
        if (c == ',' || c == ' ') { if (!cur.empty()) { out.push_back(cur); cur.clear(); } }
        else cur.push_back(c);
    }
    if (!cur.empty()) out.push_back(cur);
    return out;
}
Vec3 parse_vector3(const std::string &v, const DefaultMap &d) {
    std::vector<std::string> l = split_list(subst(v, d));
    if (l.size() == 1) { double x = to_float(l[0]); return {x, x, x}; }
    if (l.size() == 3) return {to_float(l[0]), to_float(l[1]), to_float(l[2])};
    fail("parse_vector3 failed");
}
Vec3 parse_srgb(const std::string &v, const DefaultMap &d) {
    const std::string &s = subst(v, d);
    if (s.size() != 7 || s[0] != '#') fail("Unknown SRGB format: " + s);
    char *e = nullptr;
    long enc = strtol(s.c_str() + 1, &e, 16);
    if (*e != '\0') fail("Invalid SRGB value: " + s);
    return {(double)(((enc & 0xFF0000) >> 16) / 255.0f), (double)(((enc & 0x00FF00) >> 8) / 255.0f), (double)((enc & 0x0000FF) / 255.0f)};
}
std::vector<std::pair<double, double>> parse_spectrum(const std::string &v, const DefaultMap &d) {
    std::vector<std::string> l = split_list(subst(v, d));
    std::vector<std::pair<double, double>> s;
    if (l.size() == 1 && l[0].find(':') == std::string::npos) {
        s.emplace_back(-1.0, to_float(l[0]));  // one uniform value
    } else {
        for (auto &tok : l) {
            size_t c = tok.find(':');
            if (c == std::string::npos) fail("parse_spectrum failed");
            s.emplace_back(to_float(tok.substr(0, c)), to_float(tok.substr(c + 1)));
        }
    }
    return s;
}
Mat4 parse_matrix(const std::string &v, const DefaultMap &d) {
    std::vector<std::string> l = split_list(subst(v, d));
    if (l.size() != 16) fail("parse_matrix4x4 failed");
    Mat4 m;
    for (int i = 0; i < 16; i++) m.m[i / 4][i % 4] = to_float(l[i]);
    return m;
}

// ---- spectra -> linear RGB (spectrum.h:44-118): CIE 1931 fits of Wyman et al., 400..700 nm in 1 nm steps
double x_fit(double w) {
    double t1 = (w - 442.0) * (w < 442.0 ? 0.0624 : 0.0374), t2 = (w - 599.8) * (w < 599.8 ? 0.0264 : 0.0323), t3 = (w - 501.1) * (w < 501.1 ? 0.0490 : 0.0382);
    return 0.362 * exp(-0.5 * t1 * t1) + 1.056 * exp(-0.5 * t2 * t2) - 0.065 * exp(-0.5 * t3 * t3);
}
double y_fit(double w) {
    double t1 = (w - 568.8) * (w < 568.8 ? 0.0213 : 0.0247), t2 = (w - 530.9) * (w < 530.9 ? 0.0613 : 0.0322);
    return 0.821 * exp(-0.5 * t1 * t1) + 0.286 * exp(-0.5 * t2 * t2);
}
double z_fit(double w) {
    double t1 = (w - 437.0) * (w < 437.0 ? 0.0845 : 0.0278), t2 = (w - 459.0) * (w < 459.0 ? 0.0385 : 0.0725);
    return 1.217 * exp(-0.5 * t1 * t1) + 0.681 * exp(-0.5 * t2 * t2);
}
Vec3 integrate_xyz(const std::vector<std::pair<double, double>> &data) {
    const double cie_y_integral = 106.856895, w_beg = 400, w_end = 700;
    if (data.empty()) return {0, 0, 0};
    Vec3 ret{0, 0, 0};
    int pos = 0;
    const int n = (int)data.size();
    for (double w = w_beg; w <= w_end; w += 1.0) {
        while (pos < n - 1 && !((data[pos].first <= w && data[pos + 1].first > w) || data[0].first > w)) pos++;
        double m;
        if (pos < n - 1 && data[0].first <= w) {
            double cd = data[pos].second, nd = data[std::min(pos + 1, n - 1)].second, cw = data[pos].first, nw = data[std::min(pos + 1, n - 1)].first;
            m = cd * (nw - w) / (nw - cw) + nd * (w - cw) / (nw - cw);
        } else {
            m = data[pos].second;
        }
        ret = ret + Vec3{x_fit(w), y_fit(w), z_fit(w)} * m;
    }
    return ret * ((w_end - w_beg) / (cie_y_integral * (w_end - w_beg)));
}
Vec3 xyz_to_rgb(Vec3 c) {
    return {3.240479 * c.x - 1.537150 * c.y - 0.498535 * c.z, -0.969256 * c.x + 1.875991 * c.y + 0.041556 * c.z,
            0.055648 * c.x - 0.204043 * c.y + 1.057311 * c.z};
}
Vec3 srgb_to_rgb(Vec3 s) {
    Vec3 r;
    for (int i = 0; i < 3; i++) r[i] = s[i] <= 0.04045 ? s[i] / 12.92 : pow((s[i] + 0.055) / 1.055, 2.4);
    return r;
}
double avg(Vec3 v) { return (v.x + v.y + v.z) / 3; }

// parse_color (parse_scene.cpp:286-311): a one-value spectrum is WHITE whatever the value
Vec3 parse_color(const XmlNode &n, const DefaultMap &d) {
    if (n.name == "spectrum") {
        auto spec = parse_spectrum(n.attr("value"), d);
        if (spec.size() > 1) return xyz_to_rgb(integrate_xyz(spec));
        if (spec.size() == 1) return {1, 1, 1};
        return {0, 0, 0};
    }
    if (n.name == "rgb") return parse_vector3(n.attr("value"), d);
    if (n.name == "srgb") return srgb_to_rgb(parse_srgb(n.attr("value"), d));
    if (n.name == "float") { double v = parse_float(n.attr("value"), d); return {v, v, v}; }
    fail("Unknown color type:" + n.name);
}
// parse_intensity (:494-520): a one-value RADIANCE spectrum is scaled illuminant E... (XYZ 0.9505, 1, 1.0888)
Vec3 parse_intensity(const XmlNode &n, const DefaultMap &d) {
    if (n.name == "spectrum") {
        auto spec = parse_spectrum(n.attr("value"), d);
        if (spec.size() == 1) return xyz_to_rgb(Vec3{0.9505, 1.0, 1.0888} * spec[0].second);
        return xyz_to_rgb(integrate_xyz(spec));
    }
    if (n.name == "rgb") return parse_vector3(n.attr("value"), d);
    if (n.name == "srgb") return srgb_to_rgb(parse_srgb(n.attr("value"), d));
    return {1, 1, 1};
}

std::string lower(std::string s) { for (auto &c : s) c = (char)tolower((unsigned char)c); return s; }

// parse_transform (:208-284): each operation multiplies from the left, in document order
Mat4 parse_transform(const XmlNode &node, const DefaultMap &d) {
    Mat4 t = Mat4::identity();
    for (auto &cp : node.children) {
        const XmlNode &c = *cp;
        std::string name = lower(c.name);
        auto xyz = [&](double def) {
            Vec3 v{def, def, def};
            if (c.has("x")) v.x = parse_float(c.attr("x"), d);
            if (c.has("y")) v.y = parse_float(c.attr("y"), d);
            if (c.has("z")) v.z = parse_float(c.attr("z"), d);
            return v;
        };
        if (name == "scale") {
            Vec3 v = xyz(1.0);
            if (c.has("value")) v = parse_vector3(c.attr("value"), d);
            t = scale(v) * t;
        } else if (name == "translate") {
            Vec3 v = xyz(0.0);
            if (c.has("value")) v = parse_vector3(c.attr("value"), d);
            t = translate(v) * t;
        } else if (name == "rotate") {
            Vec3 axis = xyz(0.0);
            double angle = c.has("angle") ? parse_float(c.attr("angle"), d) : 0.0;
            t = rotate(angle, axis) * t;
        } else if (name == "lookat") {
            t = look_at(parse_vector3(c.attr("origin"), d), parse_vector3(c.attr("target"), d), parse_vector3(c.attr("up"), d)) * t;
        } else if (name == "matrix") {
            t = parse_matrix(c.attr("value"), d) * t;
        }
    }
    return t;
}

// parse_texture (:313-383)
ParsedTexture parse_texture(const XmlNode &node, const DefaultMap &d) {
    ParsedTexture t;
    const std::string &type = node.attr("type");
    if (type == "bitmap") t.bitmap = true;
    else if (type == "checkerboard") { t.bitmap = false; t.color0 = {0.4, 0.4, 0.4}; t.color1 = {0.2, 0.2, 0.2}; }
    else fail("Unknown texture type: " + type);
    for (auto &cp : node.children) {
        const XmlNode &c = *cp;
        const std::string &name = c.attr("name");
        if (t.bitmap && name == "filename") t.filename = subst(c.attr("value"), d);
        else if (!t.bitmap && name == "color0") t.color0 = parse_color(c, d);
        else if (!t.bitmap && name == "color1") t.color1 = parse_color(c, d);
        else if (name == "uvscale") t.uscale = t.vscale = parse_float(c.attr("value"), d);
        else if (name == "uscale") t.uscale = parse_float(c.attr("value"), d);
        else if (name == "vscale") t.vscale = parse_float(c.attr("value"), d);
        else if (name == "uoffset") t.uoffset = parse_float(c.attr("value"), d);
        else if (name == "voffset") t.voffset = parse_float(c.attr("value"), d);
    }
    return t;
}

// ---- texture pool (texture.h:13-66): images are cached by texture NAME, not by file
int pool_image3(Ctx &cx, const std::string &name, const std::string &file) {
    auto it = cx.image3_ids.find(name);
    if (it != cx.image3_ids.end()) return it->second;
    ImageF im = read_image(cx.path(file), 3);
    int id = (int)cx.scene.images3.size();
    cx.image3_ids[name] = id;
    cx.scene.images3.push_back(HostImage{im.width, im.height, 3, std::move(im.data)});
    return id;
}
int pool_image1(Ctx &cx, const std::string &name, const ImageF &im) {
    auto it = cx.image1_ids.find(name);
    if (it != cx.image1_ids.end()) return it->second;
    int id = (int)cx.scene.images1.size();
    cx.image1_ids[name] = id;
    cx.scene.images1.push_back(HostImage{im.width, im.height, 1, im.data});
    return id;
}
int pool_image1(Ctx &cx, const std::string &name, const std::string &file) {
    auto it = cx.image1_ids.find(name);
    if (it != cx.image1_ids.end()) return it->second;
    return pool_image1(cx, name, read_image(cx.path(file), 1));
}
bool pool_has(const Ctx &cx, const std::string &name) { return cx.image1_ids.count(name) || cx.image3_ids.count(name); }

HostTexture constant_texture(Vec3 v) {
    HostTexture t;
    for (int i = 0; i < 3; i++) t.value[i] = t.color1[i] = v[i];
    return t;
}
HostTexture constant_texture(double v) { return constant_texture(Vec3{v, v, v}); }
HostTexture checker_texture(Vec3 c0, Vec3 c1, const ParsedTexture &p) {
    HostTexture t;
    t.kind = LJ_TEX_CHECKERBOARD;
    for (int i = 0; i < 3; i++) { t.value[i] = c0[i]; t.color1[i] = c1[i]; }
    t.uscale = p.uscale; t.vscale = p.vscale; t.uoffset = p.uoffset; t.voffset = p.voffset;
    return t;
}
HostTexture image_texture(int id, bool one_channel, double us, double vs, double uo, double vo) {
    HostTexture t;
    t.kind = LJ_TEX_IMAGE;
    t.image_id = id;
    t.one_channel = one_channel;
    t.uscale = us; t.vscale = vs; t.uoffset = uo; t.voffset = vo;
    return t;
}
// the name an inline <texture> gets in the pool (:431-437): first unused "$inline_..._textureN"
std::string inline_name(Ctx &cx, const char *prefix, int &counter) {
    while (pool_has(cx, prefix + std::to_string(counter))) counter++;
    return prefix + std::to_string(counter);
}
const ParsedTexture &find_texture(const Ctx &cx, const std::string &id) {
    auto it = cx.textures.find(id);
    if (it == cx.textures.end()) fail("Texture not found. ID = " + id);
    return it->second;
}

// parse_spectrum_texture (:385-448)
HostTexture parse_spectrum_texture(Ctx &cx, const XmlNode &n) {
    const DefaultMap &d = cx.defaults;
    if (n.name == "spectrum" || n.name == "rgb" || n.name == "srgb") return constant_texture(parse_color(n, d));
    if (n.name == "ref") {
        const std::string &id = n.attr("id");
        const ParsedTexture &t = find_texture(cx, id);
        if (t.bitmap) return image_texture(pool_image3(cx, id, t.filename), false, t.uscale, t.vscale, t.uoffset, t.voffset);
        return checker_texture(t.color0, t.color1, t);
    }
    if (n.name == "texture") {
        ParsedTexture t = parse_texture(n, d);
        std::string name = inline_name(cx, "$inline_spectrum_texture", cx.inline_spectrum);
        if (t.bitmap) return image_texture(pool_image3(cx, name, t.filename), false, t.uscale, t.vscale, t.uoffset, t.voffset);
        return checker_texture(t.color0, t.color1, t);
    }
    fail("Unknown spectrum texture type:" + n.name);
}
// parse_float_texture (:450-492)
HostTexture parse_float_texture(Ctx &cx, const XmlNode &n) {
    const DefaultMap &d = cx.defaults;
    auto checker1 = [](const ParsedTexture &t) { double a = avg(t.color0), b = avg(t.color1); return checker_texture(Vec3{a, a, a}, Vec3{b, b, b}, t); };
    if (n.name == "ref") {
        const std::string &id = n.attr("id");
        const ParsedTexture &t = find_texture(cx, id);
        if (t.bitmap) return image_texture(pool_image1(cx, id, t.filename), true, t.uscale, t.vscale, t.uoffset, t.voffset);
        return checker1(t);
    }
    if (n.name == "float") return constant_texture(parse_float(n.attr("value"), d));
    if (n.name == "texture") {
        ParsedTexture t = parse_texture(n, d);
        std::string name = inline_name(cx, "$inline_float_texture", cx.inline_float);
        if (t.bitmap) return image_texture(pool_image1(cx, name, t.filename), true, t.uscale, t.vscale, t.uoffset, t.voffset);
        return checker1(t);
    }
    fail("Unknown float texture type:" + n.name);
}
// alpha_to_roughness (:849-913): roughness = sqrt(alpha).  A referenced bitmap is converted texel by texel (and loses
// its uv offsets, :876); an INLINE bitmap is pooled unconverted under its temporary name (:897-898) -- both kept.
HostTexture alpha_to_roughness(Ctx &cx, const XmlNode &n) {
    const DefaultMap &d = cx.defaults;
    auto checker_sqrt = [](const ParsedTexture &t) { double a = sqrt(avg(t.color0)), b = sqrt(avg(t.color1)); return checker_texture(Vec3{a, a, a}, Vec3{b, b, b}, t); };
    if (n.name == "ref") {
        const std::string &id = n.attr("id");
        const ParsedTexture &t = find_texture(cx, id);
        if (!t.bitmap) return checker_sqrt(t);
        ImageF alpha = read_image(cx.path(t.filename), 1);
        for (auto &v : alpha.data) v = (float)sqrt((double)v);
        return image_texture(pool_image1(cx, id, alpha), true, t.uscale, t.vscale, 0, 0);
    }
    if (n.name == "float") return constant_texture(sqrt(parse_float(n.attr("value"), d)));
    if (n.name == "texture") {
        ParsedTexture t = parse_texture(n, d);
        std::string name = inline_name(cx, "$inline_alpha_texture", cx.inline_alpha);
        if (!t.bitmap) return checker_sqrt(t);
        return image_texture(pool_image1(cx, name, t.filename), true, t.uscale, t.vscale, t.uoffset, t.voffset);
    }
    fail("Unknown float texture type:" + n.name);
}

// parse_bsdf (:915-1174).  Returns the id ("" if none) and the material.
std::pair<std::string, HostMaterial> parse_bsdf(Ctx &cx, const XmlNode &node, const std::string &parent_id = "") {
    const DefaultMap &d = cx.defaults;
    const std::string &type = node.attr("type");
    std::string id = node.has("id") ? node.attr("id") : parent_id;
    HostMaterial m;
    const Vec3 grey{0.5, 0.5, 0.5}, white{1, 1, 1};
    auto is = [](const std::string &name, const char *a, const char *b = nullptr, const char *c = nullptr, const char *e = nullptr) {
        return name == a || (b && name == b) || (c && name == c) || (e && name == e);
    };
    if (type == "twosided") {  // every BSDF is two-sided here: descend
        for (auto &cp : node.children) if (cp->name == "bsdf") return parse_bsdf(cx, *cp, id);
        return {"", HostMaterial()};
    } else if (type == "diffuse") {
        m.type = LJ_MAT_LAMBERTIAN;
        m.tex[LJ_SLOT_REFLECTANCE] = constant_texture(grey);
        for (auto &cp : node.children) if (cp->attr("name") == "reflectance") m.tex[LJ_SLOT_REFLECTANCE] = parse_spectrum_texture(cx, *cp);
    } else if (type == "roughplastic" || type == "plastic" || type == "roughdielectric" || type == "dielectric") {
        const bool plastic = type == "roughplastic" || type == "plastic";
        m.type = plastic ? LJ_MAT_ROUGHPLASTIC : LJ_MAT_ROUGHDIELECTRIC;
        // slots: plastic 0 diffuse / 1 specular / 2 roughness; dielectric 0 transmittance / 1 reflectance / 2 roughness
        m.tex[0] = constant_texture(plastic ? grey : white);
        m.tex[1] = constant_texture(white);
        m.tex[LJ_SLOT_ROUGHNESS] = constant_texture((type == "plastic" || type == "dielectric") ? 0.01 : 0.1);
        double int_ior = plastic ? 1.49 : 1.5046, ext_ior = 1.000277;
        for (auto &cp : node.children) {
            const std::string &name = cp->attr("name");
            if (plastic && is(name, "diffuseReflectance", "diffuse_reflectance")) m.tex[0] = parse_spectrum_texture(cx, *cp);
            else if (!plastic && is(name, "specularTransmittance", "specular_transmittance")) m.tex[0] = parse_spectrum_texture(cx, *cp);
            else if (is(name, "specularReflectance", "specular_reflectance")) m.tex[1] = parse_spectrum_texture(cx, *cp);
            else if (name == "alpha") m.tex[LJ_SLOT_ROUGHNESS] = alpha_to_roughness(cx, *cp);
            else if (name == "roughness") m.tex[LJ_SLOT_ROUGHNESS] = parse_float_texture(cx, *cp);
            else if (is(name, "intIOR", "int_ior")) int_ior = parse_float(cp->attr("value"), d);
            else if (is(name, "extIOR", "ext_ior")) ext_ior = parse_float(cp->attr("value"), d);
        }
        m.eta = int_ior / ext_ior;
    } else if (type == "disneydiffuse") {
        m.type = LJ_MAT_DISNEY_DIFFUSE;
        m.tex[LJ_SLOT_BASE_COLOR] = constant_texture(grey);
        m.tex[LJ_SLOT_ROUGHNESS] = constant_texture(0.5);
        m.tex[LJ_SLOT_SUBSURFACE] = constant_texture(0.0);
        for (auto &cp : node.children) {
            const std::string &name = cp->attr("name");
            if (is(name, "baseColor", "base_color")) m.tex[LJ_SLOT_BASE_COLOR] = parse_spectrum_texture(cx, *cp);
            else if (name == "roughness") m.tex[LJ_SLOT_ROUGHNESS] = parse_float_texture(cx, *cp);
            else if (name == "subsurface") m.tex[LJ_SLOT_SUBSURFACE] = parse_float_texture(cx, *cp);
        }
    } else if (type == "disneymetal" || type == "disneyglass") {
        m.type = type == "disneymetal" ? LJ_MAT_DISNEY_METAL : LJ_MAT_DISNEY_GLASS;
        m.tex[LJ_SLOT_BASE_COLOR] = constant_texture(grey);
        m.tex[LJ_SLOT_ROUGHNESS] = constant_texture(0.5);
        m.tex[LJ_SLOT_ANISOTROPIC] = constant_texture(0.0);
        if (m.type == LJ_MAT_DISNEY_GLASS) m.eta = 1.5;
        for (auto &cp : node.children) {
            const std::string &name = cp->attr("name");
            if (is(name, "baseColor", "base_color")) m.tex[LJ_SLOT_BASE_COLOR] = parse_spectrum_texture(cx, *cp);
            else if (name == "roughness") m.tex[LJ_SLOT_ROUGHNESS] = parse_float_texture(cx, *cp);
            else if (name == "anisotropic") m.tex[LJ_SLOT_ANISOTROPIC] = parse_float_texture(cx, *cp);
            else if (m.type == LJ_MAT_DISNEY_GLASS && name == "eta") m.eta = parse_float(cp->attr("value"), d);
        }
    } else if (type == "disneyclearcoat") {
        m.type = LJ_MAT_DISNEY_CLEARCOAT;
        m.tex[LJ_SLOT_CLEARCOAT_GLOSS] = constant_texture(1.0);
        for (auto &cp : node.children) if (cp->attr("name") == "clearcoatGloss") m.tex[LJ_SLOT_CLEARCOAT_GLOSS] = parse_float_texture(cx, *cp);
    } else if (type == "disneysheen") {
        m.type = LJ_MAT_DISNEY_SHEEN;
        m.tex[LJ_SLOT_BASE_COLOR] = constant_texture(grey);
        m.tex[LJ_SLOT_SHEEN_TINT] = constant_texture(0.5);
        for (auto &cp : node.children) {
            const std::string &name = cp->attr("name");
            if (is(name, "baseColor", "base_color")) m.tex[LJ_SLOT_BASE_COLOR] = parse_spectrum_texture(cx, *cp);
            else if (is(name, "sheenTint", "sheen_tint")) m.tex[LJ_SLOT_SHEEN_TINT] = parse_float_texture(cx, *cp);
        }
    } else if (type == "disneybsdf" || type == "principled") {
        m.type = LJ_MAT_DISNEY_BSDF;
        m.eta = 1.5;
        m.tex[LJ_SLOT_BASE_COLOR] = constant_texture(grey);
        m.tex[LJ_SLOT_SPECULAR_TRANSMISSION] = constant_texture(0.0);
        m.tex[LJ_SLOT_METALLIC] = constant_texture(0.0);
        m.tex[LJ_SLOT_SUBSURFACE] = constant_texture(0.0);
        m.tex[LJ_SLOT_SPECULAR] = constant_texture(0.5);
        m.tex[LJ_SLOT_ROUGHNESS] = constant_texture(0.5);
        m.tex[LJ_SLOT_SPECULAR_TINT] = constant_texture(0.0);
        m.tex[LJ_SLOT_ANISOTROPIC] = constant_texture(0.0);
        m.tex[LJ_SLOT_SHEEN] = constant_texture(0.0);
        m.tex[LJ_SLOT_SHEEN_TINT] = constant_texture(0.5);
        m.tex[LJ_SLOT_CLEARCOAT] = constant_texture(0.0);
        m.tex[LJ_SLOT_CLEARCOAT_GLOSS] = constant_texture(1.0);
        for (auto &cp : node.children) {
            const std::string &name = cp->attr("name");
            if (is(name, "baseColor", "base_color")) m.tex[LJ_SLOT_BASE_COLOR] = parse_spectrum_texture(cx, *cp);
            else if (is(name, "specularTransmission", "specular_transmission", "specTrans", "spec_trans")) m.tex[LJ_SLOT_SPECULAR_TRANSMISSION] = parse_float_texture(cx, *cp);
            else if (name == "metallic") m.tex[LJ_SLOT_METALLIC] = parse_float_texture(cx, *cp);
            else if (name == "subsurface") m.tex[LJ_SLOT_SUBSURFACE] = parse_float_texture(cx, *cp);
            else if (name == "specular") m.tex[LJ_SLOT_SPECULAR] = parse_float_texture(cx, *cp);
            else if (name == "roughness") m.tex[LJ_SLOT_ROUGHNESS] = parse_float_texture(cx, *cp);
            else if (is(name, "specularTint", "specular_tint", "specTint", "spec_tint")) m.tex[LJ_SLOT_SPECULAR_TINT] = parse_float_texture(cx, *cp);
            else if (name == "anisotropic") m.tex[LJ_SLOT_ANISOTROPIC] = parse_float_texture(cx, *cp);
            else if (name == "sheen") m.tex[LJ_SLOT_SHEEN] = parse_float_texture(cx, *cp);
            else if (is(name, "sheenTint", "sheen_tint")) m.tex[LJ_SLOT_SHEEN_TINT] = parse_float_texture(cx, *cp);
            else if (name == "clearcoat") m.tex[LJ_SLOT_CLEARCOAT] = parse_float_texture(cx, *cp);
            else if (is(name, "clearcoatGloss", "clearcoat_gloss")) m.tex[LJ_SLOT_CLEARCOAT_GLOSS] = parse_float_texture(cx, *cp);
            else if (name == "eta") m.eta = parse_float(cp->attr("value"), d);
        }
    } else if (type == "null") {  // no pass-through upstream: a black diffuse surface (:1166-1169)
        m.type = LJ_MAT_LAMBERTIAN;
        m.tex[LJ_SLOT_REFLECTANCE] = constant_texture(Vec3{0, 0, 0});
    } else {
        fail("Unknown BSDF: " + type);
    }
    return {id, m};
}

// parse_volume_spectrum / parse_phase_function / parse_medium (:642-746)
HostVolume parse_volume(Ctx &cx, const XmlNode &node) {
    const std::string &type = node.attr("type");
    HostVolume v;
    if (type == "constvolume") {
        for (auto &cp : node.children)
            if (cp->attr("name") == "value") { Vec3 c = parse_color(*cp, cx.defaults); for (int i = 0; i < 3; i++) v.value[i] = c[i]; }
    } else if (type == "gridvolume") {
        std::string filename;
        for (auto &cp : node.children) if (cp->attr("name") == "filename") filename = subst(cp->attr("value"), cx.defaults);
        if (filename.empty()) fail("Empty filename for a gridvolume.");
        v.is_grid = true;
        v.grid = load_volume(cx.path(filename));
    } else {
        fail("Unknown volume type:" + type);
    }
    return v;
}
std::pair<std::string, HostMedium> parse_medium(Ctx &cx, const XmlNode &node) {
    const DefaultMap &d = cx.defaults;
    HostMedium m;
    const std::string &type = node.attr("type");
    std::string id = node.has("id") ? node.attr("id") : "";
    auto phase = [&](const XmlNode &p) {
        const std::string &pt = p.attr("type");
        if (pt == "isotropic") { m.phase_type = LJ_PHASE_ISOTROPIC; m.phase_g = 0; }
        else if (pt == "hg") {
            m.phase_type = LJ_PHASE_HG; m.phase_g = 0;
            for (auto &cp : p.children) if (cp->attr("name") == "g") m.phase_g = parse_float(cp->attr("value"), d);
        } else fail("Unrecognized phase function:" + pt);
    };
    double scale = 1;
    if (type == "homogeneous") {
        m.type = LJ_MEDIUM_HOMOGENEOUS;
        Vec3 sa{0.5, 0.5, 0.5}, ss{0.5, 0.5, 0.5};
        for (auto &cp : node.children) {
            const std::string &name = cp->attr("name");
            if (name == "sigmaA" || name == "sigma_a") sa = parse_color(*cp, d);
            else if (name == "sigmaS" || name == "sigma_s") ss = parse_color(*cp, d);
            else if (name == "scale") scale = parse_float(cp->attr("value"), d);
            else if (cp->name == "phase") phase(*cp);
        }
        for (int i = 0; i < 3; i++) { m.sigma_a[i] = sa[i] * scale; m.sigma_s[i] = ss[i] * scale; }
    } else if (type == "heterogeneous") {
        m.type = LJ_MEDIUM_HETEROGENEOUS;
        for (int i = 0; i < 3; i++) m.albedo.value[i] = m.density.value[i] = 1;
        for (auto &cp : node.children) {
            const std::string &name = cp->attr("name");
            if (name == "albedo") m.albedo = parse_volume(cx, *cp);
            else if (name == "density") m.density = parse_volume(cx, *cp);
            else if (name == "scale") scale = parse_float(cp->attr("value"), d);
            else if (cp->name == "phase") phase(*cp);
        }
        // the scale applies to the density only (:739-740): a grid keeps it as a factor, a constant volume is multiplied
        if (m.density.is_grid) m.density.scale = scale;
        else for (int i = 0; i < 3; i++) m.density.value[i] *= scale;
    } else {
        fail("Unknown medium type:" + type);
    }
    return {id, m};
}

// parse_sensor + parse_film (:592-641, :748-847) and Camera::Camera (camera.cpp:7-21)
void parse_sensor(Ctx &cx, const XmlNode &node) {
    const DefaultMap &d = cx.defaults;
    HostScene &s = cx.scene;
    double fov = 45.0;
    Mat4 to_world = Mat4::identity();
    int width = 256, height = 256;
    std::string filename = "image.exr";
    int filter_type = LJ_FILTER_BOX;
    double filter_param = 1;
    enum { AX_X, AX_Y, AX_DIAGONAL, AX_SMALLER, AX_LARGER } axis = AX_X;
    int sample_count = 4, medium_id = -1;
    if (node.attr("type") != "perspective") fail("Unsupported sensor: " + node.attr("type"));
    for (auto &cp : node.children) {
        const std::string &name = cp->attr("name");
        if (name == "fov") fov = parse_float(cp->attr("value"), d);
        else if (name == "toWorld" || name == "to_world") to_world = parse_transform(*cp, d);
        else if (name == "fovAxis" || name == "fov_axis") {
            const std::string &v = cp->attr("value");
            if (v == "x") axis = AX_X; else if (v == "y") axis = AX_Y; else if (v == "diagonal") axis = AX_DIAGONAL;
            else if (v == "smaller") axis = AX_SMALLER; else if (v == "larger") axis = AX_LARGER;
            else fail("Unknown fovAxis value: " + v);
        }
    }
    for (auto &cp : node.children) {
        const XmlNode &c = *cp;
        if (c.name == "film") {
            width = 256; height = 256; filename = "image.exr"; filter_type = LJ_FILTER_BOX; filter_param = 1;
            for (auto &gp : c.children) {
                const std::string &name = gp->attr("name");
                if (name == "width") width = parse_integer(gp->attr("value"), d);
                else if (name == "height") height = parse_integer(gp->attr("value"), d);
                else if (name == "filename") filename = subst(gp->attr("value"), d);
                if (gp->name == "rfilter") {
                    const std::string &ft = gp->attr("type");
                    auto param = [&](const char *key, double def) {
                        double v = def;
                        for (auto &hp : gp->children) if (hp->attr("name") == key) v = parse_float(hp->attr("value"), d);
                        return v;
                    };
                    if (ft == "box") { filter_type = LJ_FILTER_BOX; filter_param = param("width", 1.0); }
                    else if (ft == "tent") { filter_type = LJ_FILTER_TENT; filter_param = param("width", 2.0); }
                    else if (ft == "gaussian") { filter_type = LJ_FILTER_GAUSSIAN; filter_param = param("stddev", 0.5); }
                }
            }
        } else if (c.name == "sampler") {
            if (c.attr("type") != "independent") std::cerr << "Warning: the renderer currently only supports independent samplers." << std::endl;
            for (auto &gp : c.children) {
                const std::string &name = gp->attr("name");
                if (name == "sampleCount" || name == "sample_count") sample_count = parse_integer(gp->attr("value"), d);
            }
        } else if (c.name == "ref") {
            if (!c.has("id")) fail("Medium reference not specified.");
            auto it = cx.medium_ids.find(c.attr("id"));
            if (it == cx.medium_ids.end()) fail("Medium reference " + c.attr("id") + " not found.");
            medium_id = it->second;
        } else if (c.name == "medium") {
            auto pm = parse_medium(cx, c);
            if (!pm.first.empty()) cx.medium_ids[pm.first] = (int)s.media.size();
            medium_id = (int)s.media.size();
            s.media.push_back(pm.second);
        }
    }
    // to a horizontal field of view (as Mitsuba's sensor.cpp)
    if (axis == AX_Y || (axis == AX_SMALLER && height < width) || (axis == AX_LARGER && width < height)) {
        double aspect = width / (double)height;
        fov = degrees(2 * atan(tan(radians(fov) / 2) * aspect));
    } else if (axis == AX_DIAGONAL) {
        double aspect = width / (double)height;
        double diagonal = 2 * tan(radians(fov) / 2);
        double w = diagonal / sqrt(1 + 1 / (aspect * aspect));
        fov = degrees(2 * atan(w / 2));
    }
    s.cam_to_world = to_world;
    s.world_to_cam = inverse(to_world);
    s.width = width; s.height = height;
    s.filter_type = filter_type; s.filter_param = filter_param;
    s.camera_medium_id = medium_id;
    double aspect = (double)width / (double)height;
    s.cam_to_sample = scale(Vec3{-0.5, -0.5 * aspect, 1.0}) * translate(Vec3{-1.0, -1.0 / aspect, 0.0}) * perspective(fov);
    s.sample_to_cam = inverse(s.cam_to_sample);
    s.output_filename = filename;
    s.samples_per_pixel = sample_count;
}

// parse_integrator (:539-590)
void parse_integrator(Ctx &cx, const XmlNode &node) {
    const DefaultMap &d = cx.defaults;
    HostScene &s = cx.scene;
    // a fresh RenderOptions: everything back to its default, the sample count included (:541, then :1442 may set it again)
    s.integrator = LJ_INT_PATH; s.samples_per_pixel = 4; s.max_depth = -1; s.rr_depth = 5; s.vol_path_version = 0; s.max_null_collisions = 1000;
    const std::string &type = node.attr("type");
    if (type == "path") {
        for (auto &cp : node.children) {
            const std::string &name = cp->attr("name");
            if (name == "maxDepth") s.max_depth = parse_integer(cp->attr("value"), d);
            else if (name == "rrDepth") s.rr_depth = parse_integer(cp->attr("value"), d);
        }
    } else if (type == "volpath") {
        s.integrator = LJ_INT_VOLPATH;
        for (auto &cp : node.children) {
            const std::string &name = cp->attr("name");
            if (name == "maxDepth" || name == "max_depth") s.max_depth = parse_integer(cp->attr("value"), d);
            else if (name == "rrDepth" || name == "rr_depth") s.rr_depth = parse_integer(cp->attr("value"), d);
            else if (name == "version") s.vol_path_version = parse_integer(cp->attr("value"), d);
            else if (name == "maxNullCollisions" || name == "max_null_collisions") s.max_null_collisions = parse_integer(cp->attr("value"), d);
        }
    } else if (type == "direct") { s.max_depth = 2; }
    else if (type == "depth") s.integrator = LJ_INT_DEPTH;
    else if (type == "shadingNormal" || type == "shading_normal") s.integrator = LJ_INT_SHADING_NORMAL;
    else if (type == "meanCurvature" || type == "mean_curvature") s.integrator = LJ_INT_MEAN_CURVATURE;
    else if (type == "rayDifferential" || type == "ray_differential") s.integrator = LJ_INT_RAY_DIFFERENTIAL;
    else if (type == "mipmapLevel" || type == "mipmap_level") s.integrator = LJ_INT_MIPMAP_LEVEL;
    else fail("Unsupported integrator: " + type);
}

// parse_shape (:1176-1407)
void parse_shape(Ctx &cx, const XmlNode &node) {
    const DefaultMap &d = cx.defaults;
    HostScene &s = cx.scene;
    HostShape sh;
    int material_id = -1, interior = -1, exterior = -1;
    for (auto &cp : node.children) {
        const XmlNode &c = *cp;
        if (c.name == "ref") {
            const std::string &nv = c.attr("name");
            if (!c.has("id")) fail("Material/medium reference id not specified.");
            const std::string &id = c.attr("id");
            if (nv == "interior" || nv == "exterior") {
                auto it = cx.medium_ids.find(id);
                if (it == cx.medium_ids.end()) fail("Medium reference " + id + " not found.");
                (nv == "interior" ? interior : exterior) = it->second;
            } else {
                auto it = cx.material_ids.find(id);
                if (it == cx.material_ids.end()) fail("Material reference " + id + " not found.");
                material_id = it->second;
            }
        } else if (c.name == "bsdf") {
            auto pm = parse_bsdf(cx, c);
            if (!pm.first.empty()) cx.material_ids[pm.first] = (int)s.materials.size();
            material_id = (int)s.materials.size();
            s.materials.push_back(pm.second);
        } else if (c.name == "medium") {
            auto pm = parse_medium(cx, c);
            if (!pm.first.empty()) cx.medium_ids[pm.first] = (int)s.media.size();
            const std::string &nv = c.attr("name");
            if (nv == "interior") interior = (int)s.media.size();
            else if (nv == "exterior") exterior = (int)s.media.size();
            else fail("Unrecognized medium name: " + nv);
            s.media.push_back(pm.second);
        }
    }
    const std::string &type = node.attr("type");
    if (type == "obj" || type == "serialized" || type == "ply") {
        std::string filename;
        Mat4 to_world = Mat4::identity();
        bool face_normals = false;
        int shape_index = 0;
        for (auto &cp : node.children) {
            const std::string &name = cp->attr("name");
            if (name == "filename") filename = subst(cp->attr("value"), d);
            else if (name == "toWorld" || name == "to_world") { if (cp->name == "transform") to_world = parse_transform(*cp, d); }
            else if ((name == "shapeIndex" || name == "shape_index") && type != "obj") shape_index = parse_integer(cp->attr("value"), d);
            else if (name == "faceNormals" || name == "face_normals") face_normals = parse_boolean(cp->attr("value"), d);
        }
        sh.type = LJ_SHAPE_MESH;
        if (type == "obj") sh.mesh = load_obj(cx.path(filename), to_world);
        else if (type == "serialized") sh.mesh = load_serialized(cx.path(filename), shape_index, to_world);
        else sh.mesh = load_ply(cx.path(filename), to_world);
        if (face_normals) sh.mesh.normals.clear();
        else if (sh.mesh.normals.empty()) sh.mesh.normals = compute_vertex_normals(sh.mesh.positions, sh.mesh.indices);
    } else if (type == "sphere") {
        sh.type = LJ_SHAPE_SPHERE;
        sh.center = {0, 0, 0};
        sh.radius = 1;
        for (auto &cp : node.children) {
            const std::string &name = cp->attr("name");
            if (name == "center") sh.center = {parse_float(cp->attr("x"), d), parse_float(cp->attr("y"), d), parse_float(cp->attr("z"), d)};
            else if (name == "radius") sh.radius = parse_float(cp->attr("value"), d);
        }
    } else if (type == "rectangle") {
        sh.type = LJ_SHAPE_MESH;
        Mat4 to_world = Mat4::identity();
        bool flip = false;
        sh.mesh.positions = {Vec3{-1, -1, 0}, Vec3{1, -1, 0}, Vec3{1, 1, 0}, Vec3{-1, 1, 0}};
        sh.mesh.indices = {0, 1, 2, 0, 2, 3};
        sh.mesh.uvs = {Vec2{0, 0}, Vec2{1, 0}, Vec2{1, 1}, Vec2{0, 1}};
        sh.mesh.normals.assign(4, Vec3{0, 0, 1});
        for (auto &cp : node.children) {
            const std::string &name = cp->attr("name");
            if (name == "toWorld" || name == "to_world") { if (cp->name == "transform") to_world = parse_transform(*cp, d); }
            else if (name == "flipNormals" || name == "flip_normals") flip = parse_boolean(cp->attr("value"), d);
        }
        if (flip) for (auto &n : sh.mesh.normals) n = -n;
        for (auto &p : sh.mesh.positions) p = xform_point(to_world, p);
        const Mat4 inv = inverse(to_world);
        for (auto &n : sh.mesh.normals) n = xform_normal(inv, n);
    } else {
        fail("Unknown shape:" + type);
    }
    sh.material_id = material_id;
    sh.interior_medium_id = interior;
    sh.exterior_medium_id = exterior;
    for (auto &cp : node.children) {
        if (cp->name != "emitter") continue;
        Vec3 radiance{1, 1, 1};
        for (auto &gp : cp->children) if (gp->attr("name") == "radiance") radiance = parse_intensity(*gp, d);
        sh.area_light_id = (int)s.lights.size();
        HostLight l;
        l.type = LJ_LIGHT_AREA;
        l.shape_id = (int)s.shapes.size();
        for (int i = 0; i < 3; i++) l.intensity[i] = radiance[i];
        s.lights.push_back(l);
    }
    s.shapes.push_back(std::move(sh));
}

// frame.h:11-21
void coordinate_system(Vec3 n, Vec3 &a, Vec3 &b) {
    if (n.z < -1 + 1e-6) { a = {0, -1, 0}; b = {-1, 0, 0}; return; }
    double s = 1 / (1 + n.z), t = -n.x * n.y * s;
    a = {1 - n.x * n.x * s, t, -n.x};
    b = {t, 1 - n.y * n.y * s, -n.y};
}

// top-level <emitter> (:1469-1578)
void parse_emitter(Ctx &cx, const XmlNode &node) {
    const DefaultMap &d = cx.defaults;
    HostScene &s = cx.scene;
    const std::string &type = node.attr("type");
    auto black_material = [&]() {
        HostMaterial m;
        m.type = LJ_MAT_LAMBERTIAN;
        m.tex[LJ_SLOT_REFLECTANCE] = constant_texture(Vec3{0, 0, 0});
        s.materials.push_back(m);
        return (int)s.materials.size() - 1;
    };
    auto add_light = [&](HostShape &sh, Vec3 intensity) {
        sh.material_id = black_material();
        sh.area_light_id = (int)s.lights.size();
        HostLight l;
        l.type = LJ_LIGHT_AREA;
        l.shape_id = (int)s.shapes.size();
        for (int i = 0; i < 3; i++) l.intensity[i] = intensity[i];
        s.lights.push_back(l);
        s.shapes.push_back(std::move(sh));
    };
    auto xyz_into = [&](const XmlNode &n, Vec3 &v) {
        if (n.has("x")) v.x = parse_float(n.attr("x"), d);
        if (n.has("y")) v.y = parse_float(n.attr("y"), d);
        if (n.has("z")) v.z = parse_float(n.attr("z"), d);
    };
    if (type == "envmap") {
        std::string filename;
        double scale_v = 1;
        Mat4 to_world = Mat4::identity();
        for (auto &gp : node.children) {
            const std::string &name = gp->attr("name");
            if (name == "filename") filename = subst(gp->attr("value"), d);
            else if (name == "toWorld" || name == "to_world") to_world = parse_transform(*gp, d);
            else if (name == "scale") scale_v = parse_float(gp->attr("value"), d);
        }
        if (filename.empty()) fail("Filename unspecified for envmap.");
        HostLight l;
        l.type = LJ_LIGHT_ENVMAP;
        l.shape_id = -1;
        l.values = image_texture(pool_image3(cx, "__envmap_texture__", filename), false, 1, 1, 0, 0);
        l.to_world = to_world;
        l.to_local = inverse(to_world);
        l.scale = scale_v;
        s.lights.push_back(l);
        s.envmap_light_id = (int)s.lights.size() - 1;
    } else if (type == "point") {
        std::cout << "[Warning] converting a point light into a small spherical light." << std::endl;
        Vec3 position{0, 0, 0}, intensity{1, 1, 1};
        for (auto &gp : node.children) {
            const std::string &name = gp->attr("name");
            if (name == "position") xyz_into(*gp, position);
            else if (name == "intensity") intensity = parse_intensity(*gp, d);
        }
        HostShape sh;
        sh.type = LJ_SHAPE_SPHERE;
        sh.center = position;
        sh.radius = 1e-4;
        intensity = intensity * ((4 * kPi) / (4 * kPi * sh.radius * sh.radius));  // c_FOURPI / surface_area
        add_light(sh, intensity);
    } else if (type == "directional") {
        std::cout << "[Warning] converting a directional light into a small spherical light." << std::endl;
        Vec3 direction{0, 0, 1}, intensity{1, 1, 1};
        for (auto &gp : node.children) {
            const std::string &name = gp->attr("name");
            if (name == "direction") xyz_into(*gp, direction);
            else if (name == "toWorld" || name == "to_world") direction = xform_vector(parse_transform(*gp, d), direction);
            else if (name == "irradiance") intensity = parse_intensity(*gp, d);
        }
        direction = normalize(direction);
        Vec3 t, b;
        coordinate_system(-direction, t, b);
        const double len = 1e-3, dist = 1e3;
        HostShape sh;
        sh.type = LJ_SHAPE_MESH;
        sh.mesh.positions = {0.5 * len * (-t - b) - dist * direction, 0.5 * len * (t - b) - dist * direction,
                             0.5 * len * (t + b) - dist * direction, 0.5 * len * (-t + b) - dist * direction};
        sh.mesh.indices = {0, 1, 2, 0, 2, 3};
        sh.mesh.normals.assign(4, direction);
        intensity = intensity * ((dist * dist) / (len * len));
        add_light(sh, intensity);
    } else {
        fail("Unknown emitter type:" + type);
    }
}

}  // namespace

HostScene parse_scene_file(const std::string &xml_path) {
    std::unique_ptr<XmlNode> root = xml_load_file(xml_path);
    // pugixml's doc.child("scene"): the root element must be <scene>
    if (root->name != "scene") fail("Parse error: no <scene> element in " + xml_path);
    Ctx cx;
    size_t slash = xml_path.find_last_of('/');
    cx.dir = slash == std::string::npos ? "" : xml_path.substr(0, slash);
    HostScene &s = cx.scene;
    for (auto &cp : root->children) {
        const XmlNode &c = *cp;
        if (c.name == "default") {
            if (c.has("name") && c.has("value")) cx.defaults[c.attr("name")] = c.attr("value");
        } else if (c.name == "integrator") {
            parse_integrator(cx, c);
        } else if (c.name == "sensor") {
            parse_sensor(cx, c);
        } else if (c.name == "bsdf") {
            auto pm = parse_bsdf(cx, c);
            if (!pm.first.empty()) { cx.material_ids[pm.first] = (int)s.materials.size(); s.materials.push_back(pm.second); }
        } else if (c.name == "shape") {
            parse_shape(cx, c);
        } else if (c.name == "texture") {
            const std::string &id = c.attr("id");
            if (cx.textures.count(id)) fail("Duplicated texture ID:" + id);
            cx.textures[id] = parse_texture(c, cx.defaults);
        } else if (c.name == "emitter") {
            parse_emitter(cx, c);
        } else if (c.name == "medium") {
            auto pm = parse_medium(cx, c);
            if (!pm.first.empty()) { cx.medium_ids[pm.first] = (int)s.media.size(); s.media.push_back(pm.second); }
        }
    }
    return std::move(cx.scene);
}

// ---- HostScene -> lj_scene_desc
namespace {
void m44(float *dst, const Mat4 &m) { for (int i = 0; i < 16; i++) dst[i] = (float)m.m[i / 4][i % 4]; }
lj_texture_desc flat_texture(const HostTexture &t, int n3) {
    lj_texture_desc o;
    memset(&o, 0, sizeof(o));
    o.kind = t.kind;
    o.image_id = t.kind == LJ_TEX_IMAGE ? t.image_id + (t.one_channel ? n3 : 0) : -1;
    for (int i = 0; i < 3; i++) { o.value[i] = (float)t.value[i]; o.color1[i] = (float)t.color1[i]; }
    o.uscale = (float)t.uscale; o.vscale = (float)t.vscale; o.uoffset = (float)t.uoffset; o.voffset = (float)t.voffset;
    return o;
}
lj_volume_desc flat_volume(const HostVolume &v, FlatScene &out) {
    lj_volume_desc o;
    memset(&o, 0, sizeof(o));
    o.is_grid = v.is_grid ? 1 : 0;
    o.scale = v.is_grid ? (float)v.scale : 1.f;
    if (v.is_grid) {
        for (int i = 0; i < 3; i++) { o.res[i] = v.grid.res[i]; o.value[i] = v.grid.max_data[i]; o.p_min[i] = v.grid.p_min[i]; o.p_max[i] = v.grid.p_max[i]; }
        out.floats.push_back(v.grid.data);
        o.data = out.floats.back().data();
    } else {
        for (int i = 0; i < 3; i++) o.value[i] = (float)v.value[i];
    }
    return o;
}
}  // namespace

void to_flat(const HostScene &s, FlatScene &out) {
    out = FlatScene();
    lj_scene_desc &d = out.desc;
    memset(&d, 0, sizeof(d));
    m44(d.camera.cam_to_world, s.cam_to_world); m44(d.camera.world_to_cam, s.world_to_cam);
    m44(d.camera.sample_to_cam, s.sample_to_cam); m44(d.camera.cam_to_sample, s.cam_to_sample);
    d.camera.width = s.width; d.camera.height = s.height;
    d.camera.filter_type = s.filter_type; d.camera.filter_param = (float)s.filter_param;
    d.camera.medium_id = s.camera_medium_id;
    d.options.integrator = s.integrator; d.options.samples_per_pixel = s.samples_per_pixel; d.options.max_depth = s.max_depth;
    d.options.rr_depth = s.rr_depth; d.options.vol_path_version = s.vol_path_version; d.options.max_null_collisions = s.max_null_collisions;
    const int n3 = (int)s.images3.size();
    // (reserve: the description keeps pointers into these vectors)
    out.floats.reserve(s.images3.size() + s.images1.size() + 3 * s.shapes.size() + 2 * s.media.size() + 8);
    out.ints.reserve(s.shapes.size() + 1);
    for (int pass = 0; pass < 2; pass++)
        for (const HostImage &im : pass == 0 ? s.images3 : s.images1) {
            lj_image_desc o;
            memset(&o, 0, sizeof(o));
            o.width = im.width; o.height = im.height; o.channels = im.channels;
            out.floats.push_back(im.data);
            o.data = out.floats.back().data();
            out.images.push_back(o);
        }
    for (const HostMaterial &m : s.materials) {
        lj_material_desc o;
        memset(&o, 0, sizeof(o));
        o.type = m.type; o.eta = (float)m.eta;
        for (int k = 0; k < LJ_NUM_TEX_SLOTS; k++) o.tex[k] = flat_texture(m.tex[k], n3);
        out.materials.push_back(o);
    }
    for (const HostShape &sh : s.shapes) {
        lj_shape_desc o;
        memset(&o, 0, sizeof(o));
        o.type = sh.type; o.material_id = sh.material_id; o.area_light_id = sh.area_light_id;
        o.interior_medium_id = sh.interior_medium_id; o.exterior_medium_id = sh.exterior_medium_id;
        if (sh.type == LJ_SHAPE_SPHERE) {
            for (int i = 0; i < 3; i++) o.center[i] = (float)sh.center[i];
            o.radius = (float)sh.radius;
        } else {
            const Mesh &m = sh.mesh;
            o.num_vertices = (int)m.positions.size(); o.num_triangles = (int)m.indices.size() / 3;
            std::vector<float> p(3 * m.positions.size());
            for (size_t i = 0; i < m.positions.size(); i++) for (int c = 0; c < 3; c++) p[3 * i + c] = (float)m.positions[i][c];
            out.floats.push_back(std::move(p));
            o.positions = out.floats.back().data();
            out.ints.push_back(std::vector<int32_t>(m.indices.begin(), m.indices.end()));
            o.indices = out.ints.back().data();
            if (!m.normals.empty()) {
                std::vector<float> n(3 * m.normals.size());
                for (size_t i = 0; i < m.normals.size(); i++) for (int c = 0; c < 3; c++) n[3 * i + c] = (float)m.normals[i][c];
                out.floats.push_back(std::move(n));
                o.normals = out.floats.back().data();
            }
            if (!m.uvs.empty()) {
                std::vector<float> uv(2 * m.uvs.size());
                for (size_t i = 0; i < m.uvs.size(); i++) { uv[2 * i] = (float)m.uvs[i].x; uv[2 * i + 1] = (float)m.uvs[i].y; }
                out.floats.push_back(std::move(uv));
                o.uvs = out.floats.back().data();
            }
        }
        out.shapes.push_back(o);
    }
    for (const HostLight &l : s.lights) {
        lj_light_desc o;
        memset(&o, 0, sizeof(o));
        o.type = l.type; o.shape_id = l.shape_id;
        for (int i = 0; i < 3; i++) o.intensity[i] = (float)l.intensity[i];
        HostTexture empty;
        o.values = flat_texture(l.type == LJ_LIGHT_ENVMAP ? l.values : empty, n3);
        m44(o.to_world, l.to_world); m44(o.to_local, l.to_local);
        o.scale = (float)l.scale;
        out.lights.push_back(o);
    }
    for (const HostMedium &m : s.media) {
        lj_medium_desc o;
        memset(&o, 0, sizeof(o));
        o.type = m.type; o.phase_type = m.phase_type; o.phase_g = (float)m.phase_g;
        if (m.type == LJ_MEDIUM_HOMOGENEOUS) for (int i = 0; i < 3; i++) { o.sigma_a[i] = (float)m.sigma_a[i]; o.sigma_s[i] = (float)m.sigma_s[i]; }
        else { o.albedo = flat_volume(m.albedo, out); o.density = flat_volume(m.density, out); }
        out.media.push_back(o);
    }
    d.num_images = (int)out.images.size(); d.num_materials = (int)out.materials.size(); d.num_shapes = (int)out.shapes.size();
    d.num_lights = (int)out.lights.size(); d.num_media = (int)out.media.size();
    d.envmap_light_id = s.envmap_light_id;
    d.images = out.images.data(); d.materials = out.materials.data(); d.shapes = out.shapes.data();
    d.lights = out.lights.data(); d.media = out.media.data();
}

// ---- .ljs container (layout: lajolla_public_b200/ljs.py)
void write_ljs(const FlatScene &flat, const std::string &path) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) fail("cannot write " + path);
    auto i32 = [&](int32_t v) { fwrite(&v, 4, 1, f); };
    auto f32 = [&](float v) { fwrite(&v, 4, 1, f); };
    auto arr = [&](const float *p, size_t n) { fwrite(p, 4, n, f); };
    auto tex = [&](const lj_texture_desc &t) { i32(t.kind); i32(t.image_id); arr(t.value, 3); arr(t.color1, 3); f32(t.uscale); f32(t.vscale); f32(t.uoffset); f32(t.voffset); };
    const lj_scene_desc &d = flat.desc;
    fwrite("LJS1", 4, 1, f);
    i32(1);
    arr(d.camera.cam_to_world, 16); arr(d.camera.world_to_cam, 16); arr(d.camera.sample_to_cam, 16); arr(d.camera.cam_to_sample, 16);
    i32(d.camera.width); i32(d.camera.height); i32(d.camera.filter_type); f32(d.camera.filter_param); i32(d.camera.medium_id);
    i32(d.options.integrator); i32(d.options.samples_per_pixel); i32(d.options.max_depth); i32(d.options.rr_depth);
    i32(d.options.vol_path_version); i32(d.options.max_null_collisions);
    i32(d.num_images); i32(d.num_materials); i32(d.num_shapes); i32(d.num_lights); i32(d.num_media); i32(d.envmap_light_id);
    for (const lj_image_desc &im : flat.images) { i32(im.width); i32(im.height); i32(im.channels); arr(im.data, (size_t)im.width * im.height * im.channels); }
    for (const lj_material_desc &m : flat.materials) { i32(m.type); f32(m.eta); for (int k = 0; k < LJ_NUM_TEX_SLOTS; k++) tex(m.tex[k]); }
    for (const lj_shape_desc &s : flat.shapes) {
        i32(s.type); i32(s.material_id); i32(s.area_light_id); i32(s.interior_medium_id); i32(s.exterior_medium_id);
        arr(s.center, 3); f32(s.radius);
        // (the container stores the VERTEX count and the TRIANGLE count, then the flags)
        i32(s.num_vertices); i32(s.num_triangles); i32(s.normals != nullptr); i32(s.uvs != nullptr);
        if (s.type == LJ_SHAPE_MESH) {
            arr(s.positions, 3 * (size_t)s.num_vertices);
            fwrite(s.indices, 4, 3 * (size_t)s.num_triangles, f);
            if (s.normals) arr(s.normals, 3 * (size_t)s.num_vertices);
            if (s.uvs) arr(s.uvs, 2 * (size_t)s.num_vertices);
        }
    }
    for (const lj_light_desc &l : flat.lights) { i32(l.type); i32(l.shape_id); arr(l.intensity, 3); tex(l.values); arr(l.to_world, 16); arr(l.to_local, 16); f32(l.scale); }
    auto vol = [&](const lj_volume_desc &v) {
        i32(v.is_grid); i32(v.res[0]); i32(v.res[1]); i32(v.res[2]); arr(v.value, 3); arr(v.p_min, 3); arr(v.p_max, 3); f32(v.scale);
        if (v.is_grid) arr(v.data, 3 * (size_t)v.res[0] * v.res[1] * v.res[2]);
    };
    for (const lj_medium_desc &m : flat.media) {
        i32(m.type); i32(m.phase_type); f32(m.phase_g); arr(m.sigma_a, 3); arr(m.sigma_s, 3);
        if (m.type == LJ_MEDIUM_HETEROGENEOUS) { vol(m.albedo); vol(m.density); }
    }
    fclose(f);
}

}  // namespace ljhost
