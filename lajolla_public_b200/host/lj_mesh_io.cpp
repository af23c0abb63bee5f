// lj_mesh_io.cpp -- see lj_mesh_io.h.
#include "lj_mesh_io.h"

#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

namespace ljhost {

namespace {
[[noreturn]] void fail(const std::string &what) { throw std::runtime_error(what); }

std::string trim(const std::string &s) {
    size_t b = 0, e = s.size();
    while (b < e && isspace((unsigned char)s[b])) b++;
    while (e > b && isspace((unsigned char)s[e - 1])) e--;
    return s.substr(b, e - b);
}

// "v", "v/vt", "v//vn", "v/vt/vn" -> zero-based indices, -1 where absent (parse_obj.cpp:31-64)
struct ObjKey {
    int v, vt, vn;
    bool operator<(const ObjKey &o) const { return v != o.v ? v < o.v : (vt != o.vt ? vt < o.vt : vn < o.vn); }
};
ObjKey parse_face_vertex(const std::string &s) {
    int id[3] = {0, 0, 0}, k = 0;
    size_t b = 0;
    while (k < 3) {
        size_t e = s.find('/', b);
        std::string tok = s.substr(b, e == std::string::npos ? std::string::npos : e - b);
        id[k++] = tok.empty() ? 0 : std::stoi(tok);
        if (e == std::string::npos) break;
        b = e + 1;
    }
    return ObjKey{id[0] - 1, id[1] - 1, id[2] - 1};
}
}  // namespace

Mesh load_obj(const std::string &path, const Mat4 &to_world) {
    std::ifstream ifs(path.c_str());
    if (!ifs.is_open()) fail("Unable to open the obj file " + path);
    std::vector<Vec3> pos_pool, nor_pool;
    std::vector<Vec2> st_pool;
    std::map<ObjKey, int> vertex_map;
    Mesh mesh;
    const Mat4 inv = inverse(to_world);
    auto vertex_id = [&](const ObjKey &k) {
        auto it = vertex_map.find(k);
        if (it != vertex_map.end()) return it->second;
        if (k.v < 0 || k.v >= (int)pos_pool.size()) fail("obj: vertex index out of range in " + path);
        int id = (int)mesh.positions.size();
        mesh.positions.push_back(xform_point(to_world, pos_pool[k.v]));
        if (k.vt != -1) { if (k.vt < 0 || k.vt >= (int)st_pool.size()) fail("obj: vt index out of range"); mesh.uvs.push_back(st_pool[k.vt]); }
        if (k.vn != -1) { if (k.vn < 0 || k.vn >= (int)nor_pool.size()) fail("obj: vn index out of range"); mesh.normals.push_back(xform_normal(inv, nor_pool[k.vn])); }
        vertex_map[k] = id;
        return id;
    };
    std::string line;
    while (std::getline(ifs, line)) {
        line = trim(line);
        if (line.empty() || line[0] == '#') continue;
        std::stringstream ss(line);
        std::string token;
        ss >> token;
        if (token == "v") {
            double x = 0, y = 0, z = 0, w = 1;
            ss >> x >> y >> z;
            if (!(ss >> w)) w = 1;
            pos_pool.push_back(Vec3{x, y, z} / w);
        } else if (token == "vt") {
            double s = 0, t = 0;
            ss >> s >> t;
            st_pool.push_back(Vec2{s, 1 - t});  // parse_obj.cpp:120
        } else if (token == "vn") {
            double x = 0, y = 0, z = 0;
            ss >> x >> y >> z;
            nor_pool.push_back(normalize(Vec3{x, y, z}));
        } else if (token == "f") {
            std::string i0, i1, i2, i3, i4;
            ss >> i0 >> i1 >> i2;
            int a = vertex_id(parse_face_vertex(i0)), b = vertex_id(parse_face_vertex(i1)), c = vertex_id(parse_face_vertex(i2));
            mesh.indices.push_back(a); mesh.indices.push_back(b); mesh.indices.push_back(c);
            if (ss >> i3) {  // quads are fan-split (parse_obj.cpp:162-176)
                int d = vertex_id(parse_face_vertex(i3));
                mesh.indices.push_back(a); mesh.indices.push_back(c); mesh.indices.push_back(d);
            }
            if (ss >> i4) fail("The object file contains n-gon (n>4) that we do not support.");
        }
    }
    return mesh;
}

namespace {
// sequential reader over one zlib stream that starts at a file offset
struct ZReader {
    FILE *f;
    z_stream zs;
    unsigned char in[32768];
    explicit ZReader(FILE *file) : f(file) {
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, 15) != Z_OK) fail("inflateInit failed");
    }
    ~ZReader() { inflateEnd(&zs); }
    void read(void *dst, size_t n) {
        zs.next_out = (unsigned char *)dst;
        zs.avail_out = (uInt)n;
        while (zs.avail_out > 0) {
            if (zs.avail_in == 0) {
                size_t got = fread(in, 1, sizeof(in), f);
                if (got == 0) fail("serialized: read less data than expected");
                zs.next_in = in;
                zs.avail_in = (uInt)got;
            }
            int r = inflate(&zs, Z_NO_FLUSH);
            if (r == Z_STREAM_END) { if (zs.avail_out > 0) fail("serialized: attempting to read past the end of the stream"); break; }
            if (r != Z_OK) fail("serialized: inflate error");
        }
    }
    template <typename T> T get() { T v; read(&v, sizeof(T)); return v; }
};
}  // namespace

Mesh load_serialized(const std::string &path, int shape_index, const Mat4 &to_world) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) fail("cannot open " + path);
    struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{f};
    uint16_t magic = 0, version = 0;
    if (fread(&magic, 2, 1, f) != 1 || fread(&version, 2, 1, f) != 1) fail("serialized: short file " + path);
    if (shape_index > 0) {  // offset table at the end: u64 offsets (v4) or u32 (v3), then u32 count (load_serialized.cpp:103-121)
        uint32_t count = 0;
        fseek(f, -4, SEEK_END);
        if (fread(&count, 4, 1, f) != 1) fail("serialized: no shape table");
        if (shape_index >= (int)count) fail("serialized: shape index out of range in " + path);
        uint64_t offset = 0;
        if (version == 4) {
            fseek(f, -(long)(8 * (count - shape_index)) - 4, SEEK_END);
            if (fread(&offset, 8, 1, f) != 1) fail("serialized: bad table");
        } else {
            uint32_t o32 = 0;
            fseek(f, -(long)(4 * (count - shape_index + 1)), SEEK_END);
            if (fread(&o32, 4, 1, f) != 1) fail("serialized: bad table");
            offset = o32;
        }
        fseek(f, (long)offset + 4, SEEK_SET);  // skip that shape's own magic + version
    }
    ZReader zs(f);
    uint32_t flags = zs.get<uint32_t>();
    if (version == 4) { char c; do { c = zs.get<char>(); } while (c != '\0'); }  // shape name
    uint64_t nv = zs.get<uint64_t>(), nt = zs.get<uint64_t>();
    const bool dbl = (flags & 0x2000) != 0, has_normals = (flags & 0x0001) != 0, has_uvs = (flags & 0x0002) != 0, has_colors = (flags & 0x0008) != 0;
    auto real = [&]() { return dbl ? zs.get<double>() : (double)zs.get<float>(); };
    Mesh mesh;
    mesh.positions.resize(nv);
    for (auto &p : mesh.positions) { double x = real(), y = real(), z = real(); p = xform_point(to_world, Vec3{x, y, z}); }
    if (has_normals) {
        const Mat4 inv = inverse(to_world);
        mesh.normals.resize(nv);
        for (auto &n : mesh.normals) { double x = real(), y = real(), z = real(); n = xform_normal(inv, Vec3{x, y, z}); }
    }
    if (has_uvs) {
        mesh.uvs.resize(nv);
        for (auto &uv : mesh.uvs) { double u = real(), v = real(); uv = Vec2{u, v}; }
    }
    if (has_colors) for (uint64_t i = 0; i < 3 * nv; i++) real();  // ignored (load_serialized.cpp:237-244)
    mesh.indices.resize(3 * nt);
    for (auto &i : mesh.indices) i = zs.get<int32_t>();
    return mesh;
}

Mesh load_ply(const std::string &path, const Mat4 &to_world) {
    std::ifstream ifs(path.c_str(), std::ios::binary);
    if (!ifs.is_open()) fail("cannot open " + path);
    std::string line;
    std::getline(ifs, line);
    if (trim(line) != "ply") fail("not a PLY file: " + path);
    struct Prop { std::string name, type, count_type; bool is_list; };
    struct Elem { std::string name; size_t count; std::vector<Prop> props; };
    std::vector<Elem> elems;
    bool ascii = true;
    while (std::getline(ifs, line)) {
        std::stringstream ss(trim(line));
        std::string tok;
        ss >> tok;
        if (tok == "format") { std::string fmt; ss >> fmt; if (fmt == "ascii") ascii = true; else if (fmt == "binary_little_endian") ascii = false; else fail("PLY: unsupported format " + fmt); }
        else if (tok == "element") { Elem e; ss >> e.name >> e.count; elems.push_back(e); }
        else if (tok == "property") {
            if (elems.empty()) fail("PLY: property before element");
            Prop p; std::string t; ss >> t;
            if (t == "list") { p.is_list = true; ss >> p.count_type >> p.type >> p.name; } else { p.is_list = false; p.type = t; ss >> p.name; }
            elems.back().props.push_back(p);
        } else if (tok == "end_header") break;
    }
    auto type_size = [](const std::string &t) {
        if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
        if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
        if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
        if (t == "double" || t == "float64") return 8;
        fail("PLY: unknown type " + t);
    };
    auto read_value = [&](const std::string &t) -> double {
        if (ascii) { double v; ifs >> v; return v; }
        unsigned char b[8];
        int n = type_size(t);
        ifs.read((char *)b, n);
        if (t == "char" || t == "int8") return (double)(int8_t)b[0];
        if (t == "uchar" || t == "uint8") return (double)b[0];
        if (t == "short" || t == "int16") { int16_t v; memcpy(&v, b, 2); return v; }
        if (t == "ushort" || t == "uint16") { uint16_t v; memcpy(&v, b, 2); return v; }
        if (t == "int" || t == "int32") { int32_t v; memcpy(&v, b, 4); return v; }
        if (t == "uint" || t == "uint32") { uint32_t v; memcpy(&v, b, 4); return v; }
        if (t == "float" || t == "float32") { float v; memcpy(&v, b, 4); return v; }
        double v; memcpy(&v, b, 8); return v;
    };
    Mesh mesh;
    std::vector<Vec3> raw_normals;
    bool have_pos = false, have_n = false, have_uv = false, have_faces = false;
    for (const Elem &e : elems) {
        int ix = -1, iy = -1, iz = -1, inx = -1, iny = -1, inz = -1, iu = -1, iv = -1;
        for (size_t k = 0; k < e.props.size(); k++) {
            const std::string &n = e.props[k].name;
            if (n == "x") ix = (int)k; else if (n == "y") iy = (int)k; else if (n == "z") iz = (int)k;
            else if (n == "nx") inx = (int)k; else if (n == "ny") iny = (int)k; else if (n == "nz") inz = (int)k;
            else if (n == "u") iu = (int)k; else if (n == "v") iv = (int)k;
        }
        for (size_t i = 0; i < e.count; i++) {
            std::vector<double> vals(e.props.size(), 0.0);
            for (size_t k = 0; k < e.props.size(); k++) {
                const Prop &p = e.props[k];
                if (!p.is_list) { vals[k] = read_value(p.type); continue; }
                int n = (int)read_value(p.count_type);
                std::vector<int> idx(n);
                for (int j = 0; j < n; j++) idx[j] = (int)read_value(p.type);
                if (e.name == "face" && (p.name == "vertex_indices" || p.name == "vertex_index")) {
                    if (n != 3) fail("PLY: only triangles are supported");  // parse_ply.cpp reads 3 indices per face
                    mesh.indices.insert(mesh.indices.end(), idx.begin(), idx.end());
                    have_faces = true;
                }
            }
            if (e.name == "vertex") {
                if (ix >= 0 && iy >= 0 && iz >= 0) { mesh.positions.push_back(xform_point(to_world, Vec3{vals[ix], vals[iy], vals[iz]})); have_pos = true; }
                if (inx >= 0 && iny >= 0 && inz >= 0) { raw_normals.push_back(Vec3{vals[inx], vals[iny], vals[inz]}); have_n = true; }
                if (iu >= 0 && iv >= 0) { mesh.uvs.push_back(Vec2{vals[iu], vals[iv]}); have_uv = true; }
            }
        }
    }
    if (!have_pos) fail("Vertex positions not found in " + path);
    if (!have_faces) fail("Vertex indices not found in " + path);
    if (have_n) { const Mat4 inv = inverse(to_world); for (auto &n : raw_normals) mesh.normals.push_back(xform_normal(inv, n)); }
    (void)have_uv;
    return mesh;
}

// Nelson Max angle-weighted vertex normals; the obtuse branch keeps the reference's (pi - 2) factor (shape_utils.h:10)
static double unit_angle(Vec3 u, Vec3 v) {
    if (dot(u, v) < 0) return (kPi - 2) * std::asin(0.5 * length(v + u));
    return 2 * std::asin(0.5 * length(v - u));
}
std::vector<Vec3> compute_vertex_normals(const std::vector<Vec3> &vertices, const std::vector<int> &indices) {
    std::vector<Vec3> normals(vertices.size(), Vec3{0, 0, 0});
    for (size_t t = 0; t + 2 < indices.size(); t += 3) {
        Vec3 n{0, 0, 0};
        for (int i = 0; i < 3; i++) {
            const Vec3 &v0 = vertices[indices[t + i]], &v1 = vertices[indices[t + (i + 1) % 3]], &v2 = vertices[indices[t + (i + 2) % 3]];
            Vec3 side1 = v1 - v0, side2 = v2 - v0;
            if (i == 0) {
                n = cross(side1, side2);
                double l = length(n);
                if (l == 0) break;
                n = n / l;
            }
            double angle = unit_angle(normalize(side1), normalize(side2));
            normals[indices[t + i]] = normals[indices[t + i]] + n * angle;
        }
    }
    for (auto &n : normals) {
        double l = length(n);
        n = l != 0 ? n / l : Vec3{0, 0, 0};
    }
    return normals;
}

VolumeGrid load_volume(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) fail("cannot open " + path);
    struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{f};
    char header[4];
    if (fread(header, 1, 4, f) != 4 || header[0] != 'V' || header[1] != 'O' || header[2] != 'L' || header[3] != 3)
        fail("Error loading volume from a file (incorrect header). Filename:" + path);
    int32_t type = 0, channels = 0;
    VolumeGrid g;
    if (fread(&type, 4, 1, f) != 1 || type != 1) fail("Unsupported volume format (only support Float32). Filename:" + path);
    if (fread(g.res, 4, 3, f) != 3 || fread(&channels, 4, 1, f) != 1) fail("volume: short header " + path);
    if (channels != 1 && channels != 3) fail("Unsupported volume format (wrong number of channels). Filename:" + path);
    if (fread(g.p_min, 4, 3, f) != 3 || fread(g.p_max, 4, 3, f) != 3) fail("volume: short header " + path);
    size_t n = (size_t)g.res[0] * g.res[1] * g.res[2];
    std::vector<float> raw(n * channels, 0.f);
    size_t got = fread(raw.data(), 4, raw.size(), f);
    (void)got;  // a short file leaves zeros, like the reference's unchecked read
    g.data.resize(3 * n);
    for (size_t i = 0; i < n; i++)
        for (int c = 0; c < 3; c++) {
            float v = channels == 1 ? raw[i] : raw[3 * i + c];
            g.data[3 * i + c] = v;
            g.max_data[c] = std::max(g.max_data[c], v);
        }
    return g;
}

}  // namespace ljhost
