// scene.cpp — OBJ/MTL ingest with tobj 4.0.2 semantics and the reference's mesh
// preparation (src/primitives.rs, src/renderer.rs:371-410).  Host-only C++.
// Compiled with -ffp-contract=off: Rust never fuses a*b+c, so neither may we.
#include "scene.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <unordered_map>

namespace rc {

// ---------------------------------------------------------------- glam-style f32 math
static inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline Vec3 operator*(float s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline Vec3 operator/(Vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
static inline float dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline Vec3 cross(Vec3 a, Vec3 b)
{
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
static inline Vec3 normalize(Vec3 a)
{
    float rl = 1.0f / std::sqrt(dot(a, a));  // length().recip()
    return {a.x * rl, a.y * rl, a.z * rl};
}
static inline bool is_nan(Vec3 a) { return std::isnan(a.x) || std::isnan(a.y) || std::isnan(a.z); }

// ---------------------------------------------------------------- text helpers
static std::string trim(const std::string& s)
{
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

static std::vector<std::string> split_ws(const std::string& s)
{
    std::vector<std::string> out;
    size_t i = 0, n = s.size();
    while (i < n) {
        while (i < n && isspace((unsigned char)s[i])) i++;
        size_t j = i;
        while (j < n && !isspace((unsigned char)s[j])) j++;
        if (j > i) out.emplace_back(s.substr(i, j - i));
        i = j;
    }
    return out;
}

static bool parse_f32(const std::string& tok, float& out)
{
    char* end = nullptr;
    out = strtof(tok.c_str(), &end);  // correctly rounded, like str::parse::<f32>()
    return end != tok.c_str();
}

static bool parse_floatn(const std::vector<std::string>& w, size_t first, size_t n, std::vector<float>& dst)
{
    if (w.size() < first + n) return false;
    float tmp[4];
    for (size_t i = 0; i < n; i++)
        if (!parse_f32(w[first + i], tmp[i])) return false;
    dst.insert(dst.end(), tmp, tmp + n);
    return true;
}

// ---------------------------------------------------------------- MTL (tobj::load_mtl_buf)
static bool load_mtl(const std::string& path, std::vector<TobjMaterial>& mats, std::map<std::string, size_t>& name_map,
                     size_t offset, LoadError& err)
{
    std::ifstream in(path);
    if (!in) { err.message = "cannot open material library " + path; return false; }
    std::string raw;
    bool have = false;
    TobjMaterial cur;
    auto flush = [&]() {
        if (have) { name_map[cur.name] = offset + mats.size(); mats.push_back(cur); }
    };
    while (std::getline(in, raw)) {
        std::string line = trim(raw);
        auto w = split_ws(line);
        if (w.empty() || w[0][0] == '#') continue;
        const std::string& key = w[0];
        std::string rest = trim(line.substr(key.size()));
        if (key == "newmtl") {
            flush();
            cur = TobjMaterial();
            cur.name = rest;
            have = true;
            continue;
        }
        if (!have) continue;
        auto vec3 = [&](std::optional<Vec3>& dst) {
            std::vector<float> f;
            if (parse_floatn(w, 1, 3, f)) dst = Vec3{f[0], f[1], f[2]};
        };
        if (key == "Ka") vec3(cur.ambient);
        else if (key == "Kd") vec3(cur.diffuse);
        else if (key == "Ks") vec3(cur.specular);
        else if (key == "Ns") { float v; if (w.size() > 1 && parse_f32(w[1], v)) cur.shininess = v; }
        else if (key == "map_Kd") { if (!rest.empty()) cur.diffuse_texture = rest; }
        else if (key == "map_Bump" || key == "map_bump" || key == "bump") { if (!rest.empty()) cur.normal_texture = rest; }
        else if (key == "Ke") {  // tobj keeps this in unknown_param; the reference ignores it; GI emission
            std::vector<float> f;
            if (parse_floatn(w, 1, 3, f)) cur.emission = Vec3{f[0], f[1], f[2]};
        }
    }
    flush();
    return true;
}

// ---------------------------------------------------------------- OBJ (tobj::load_obj)
namespace {
constexpr int64_t MISSING = -1;
struct VertexIndices {
    int64_t v, vt, vn;
    bool operator==(const VertexIndices& o) const { return v == o.v && vt == o.vt && vn == o.vn; }
};
struct VIHash {
    size_t operator()(const VertexIndices& k) const
    {
        uint64_t h = 1469598103934665603ull;
        for (int64_t x : {k.v, k.vt, k.vn}) { h ^= (uint64_t)x + 0x9e3779b97f4a7c15ull; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
using Face = std::vector<VertexIndices>;
}  // namespace

static bool parse_vertex_indices(const std::string& word, size_t npos, size_t ntex, size_t nnorm, VertexIndices& out)
{
    int64_t vals[3] = {MISSING, MISSING, MISSING};
    const size_t sizes[3] = {npos, ntex, nnorm};
    size_t start = 0;
    for (int i = 0; i < 3; i++) {
        size_t slash = word.find('/', start);
        std::string part = word.substr(start, slash == std::string::npos ? std::string::npos : slash - start);
        if (!part.empty()) {
            char* end = nullptr;
            long long x = strtoll(part.c_str(), &end, 10);
            if (end == part.c_str()) return false;
            vals[i] = x < 0 ? (int64_t)sizes[i] + x : x - 1;
        }
        if (slash == std::string::npos) break;
        start = slash + 1;
    }
    out = {vals[0], vals[1], vals[2]};
    return true;
}

// tobj export_faces, single_index path: first-appearance (v,vt,vn) re-indexing within a
// model; fan triangulation; points/lines become zero-area triangles.
static bool export_faces(const std::vector<float>& pos, const std::vector<float>& vcol, const std::vector<float>& tex,
                         const std::vector<float>& nrm, const std::vector<Face>& faces,
                         std::optional<size_t> mat_id, TobjMesh& mesh, LoadError& err)
{
    std::unordered_map<VertexIndices, uint32_t, VIHash> index_map;
    mesh = TobjMesh();
    mesh.material_id = mat_id;
    bool ok = true;
    auto add = [&](const VertexIndices& vi) {
        auto it = index_map.find(vi);
        if (it != index_map.end()) { mesh.indices.push_back(it->second); return; }
        if (vi.v < 0 || (size_t)vi.v * 3 + 2 >= pos.size()) { ok = false; err.message = "face vertex out of bounds"; return; }
        for (int k = 0; k < 3; k++) mesh.positions.push_back(pos[(size_t)vi.v * 3 + k]);
        if (!tex.empty() && vi.vt != MISSING) {
            if ((size_t)vi.vt * 2 + 1 >= tex.size()) { ok = false; err.message = "face texcoord out of bounds"; return; }
            mesh.texcoords.push_back(tex[(size_t)vi.vt * 2]);
            mesh.texcoords.push_back(tex[(size_t)vi.vt * 2 + 1]);
        }
        if (!nrm.empty() && vi.vn != MISSING) {
            if ((size_t)vi.vn * 3 + 2 >= nrm.size()) { ok = false; err.message = "face normal out of bounds"; return; }
            for (int k = 0; k < 3; k++) mesh.normals.push_back(nrm[(size_t)vi.vn * 3 + k]);
        }
        if (!vcol.empty() && (size_t)vi.v * 3 + 2 < vcol.size())
            for (int k = 0; k < 3; k++) mesh.vertex_color.push_back(vcol[(size_t)vi.v * 3 + k]);
        uint32_t next = (uint32_t)index_map.size();
        mesh.indices.push_back(next);
        index_map.emplace(vi, next);
    };
    for (const Face& f : faces) {
        switch (f.size()) {
        case 0: break;
        case 1: add(f[0]); add(f[0]); add(f[0]); break;
        case 2: add(f[0]); add(f[1]); add(f[1]); break;
        case 3: add(f[0]); add(f[1]); add(f[2]); break;
        case 4: add(f[0]); add(f[1]); add(f[2]); add(f[0]); add(f[2]); add(f[3]); break;
        default: {
            const VertexIndices& a = f[0];
            size_t b = 1;
            for (size_t c = 2; c < f.size(); c++) { add(a); add(f[b]); add(f[c]); b = c; }
        }
        }
        if (!ok) return false;
    }
    return true;
}

static bool load_obj(const std::string& path, std::vector<TobjModel>& models, std::vector<TobjMaterial>& materials,
                     LoadError& err)
{
    std::ifstream in(path);
    if (!in) { err.message = "cannot open " + path; return false; }
    std::string dir;
    {
        size_t s = path.find_last_of('/');
        dir = s == std::string::npos ? std::string(".") : path.substr(0, s);
    }
    std::vector<float> pos, vcol, tex, nrm;
    std::vector<Face> faces;
    std::map<std::string, size_t> mat_map;
    std::string name = "unnamed_object";
    std::optional<size_t> mat_id;
    auto push_model = [&]() -> bool {
        TobjModel m;
        m.name = name;
        if (!export_faces(pos, vcol, tex, nrm, faces, mat_id, m.mesh, err)) return false;
        models.push_back(std::move(m));
        faces.clear();
        return true;
    };
    std::string raw;
    while (std::getline(in, raw)) {
        std::string line = trim(raw);
        auto w = split_ws(line);
        if (w.empty()) continue;
        const std::string& key = w[0];
        if (key == "v") {
            if (!parse_floatn(w, 1, 3, pos)) { err.message = "position parse error"; return false; }
            if (w.size() >= 7) parse_floatn(w, 4, 3, vcol);
        } else if (key == "vt") {
            if (!parse_floatn(w, 1, 2, tex)) { err.message = "texcoord parse error"; return false; }
        } else if (key == "vn") {
            if (!parse_floatn(w, 1, 3, nrm)) { err.message = "normal parse error"; return false; }
        } else if (key == "f" || key == "l") {
            Face f;
            for (size_t i = 1; i < w.size(); i++) {
                VertexIndices vi;
                if (!parse_vertex_indices(w[i], pos.size() / 3, tex.size() / 2, nrm.size() / 3, vi)) {
                    err.message = "face parse error"; return false;
                }
                f.push_back(vi);
            }
            faces.push_back(std::move(f));
        } else if (key == "o" || key == "g") {
            if (!faces.empty() && !push_model()) return false;
            name = trim(line.substr(1));
            if (name.empty()) name = "unnamed_object";
        } else if (key == "mtllib") {
            std::string lib = trim(line.substr(6));
            // a missing .mtl makes the whole load fail (materials? at src/primitives.rs:132)
            if (!load_mtl(dir + "/" + lib, materials, mat_map, 0, err)) return false;
        } else if (key == "usemtl") {
            std::string mat_name = trim(line.substr(6));
            if (mat_name.empty()) { err.message = "material parse error"; return false; }
            std::optional<size_t> new_mat;
            auto it = mat_map.find(mat_name);
            if (it != mat_map.end()) new_mat = it->second;
            if (mat_id != new_mat && !faces.empty() && !push_model()) return false;
            mat_id = new_mat;
        }
    }
    return push_model();
}

// ---------------------------------------------------------------- ObjScene
bool ObjScene::load(const std::string& path, const std::function<bool(const TobjMaterial&)>& light_predicate,
                    std::vector<ObjScene>& out, std::optional<Vec3>& light, LoadError& err)
{
    std::vector<TobjModel> models;
    std::vector<TobjMaterial> mats;
    if (!load_obj(path, models, mats, err)) return false;
    std::vector<std::shared_ptr<TobjMaterial>> shared;
    for (auto& m : mats) shared.push_back(std::make_shared<TobjMaterial>(m));
    light.reset();
    for (const TobjModel& md : models) {   // first model whose material satisfies the predicate
        if (!md.mesh.material_id || *md.mesh.material_id >= shared.size()) continue;
        if (!light_predicate(*shared[*md.mesh.material_id])) continue;
        Vec3 s{0.f, 0.f, 0.f};
        size_t n = md.mesh.positions.size() / 3;
        for (size_t i = 0; i < n; i++)
            s = s + Vec3{md.mesh.positions[3 * i], md.mesh.positions[3 * i + 1], md.mesh.positions[3 * i + 2]};
        light = s / (float)n;
        break;
    }
    std::string dir;
    size_t s = path.find_last_of('/');
    dir = s == std::string::npos ? std::string(".") : path.substr(0, s);
    out.clear();
    for (TobjModel& m : models) {
        ObjScene sc;
        sc.obj_dir = dir;
        if (m.mesh.material_id && *m.mesh.material_id < shared.size()) sc.materials = shared[*m.mesh.material_id];
        sc.model = std::move(m);
        out.push_back(std::move(sc));
    }
    return true;
}

std::vector<Vec3> ObjScene::vertices() const
{
    const auto& p = model.mesh.positions;
    std::vector<Vec3> v(p.size() / 3);
    for (size_t i = 0; i < v.size(); i++) v[i] = {p[3 * i], p[3 * i + 1], p[3 * i + 2]};
    return v;
}

std::vector<Vec3> ObjScene::vertex_colors() const
{
    const auto& p = model.mesh.vertex_color;
    std::vector<Vec3> v(p.size() / 3);
    for (size_t i = 0; i < v.size(); i++) v[i] = {p[3 * i], p[3 * i + 1], p[3 * i + 2]};
    return v;
}

std::vector<Vec3> ObjScene::normals() const
{
    const auto& p = model.mesh.normals;
    std::vector<Vec3> v(p.size() / 3);
    for (size_t i = 0; i < v.size(); i++) v[i] = {p[3 * i], p[3 * i + 1], p[3 * i + 2]};
    return v;
}

std::vector<Vec2> ObjScene::texcoords() const
{
    if (model.mesh.positions.size() / 3 != model.mesh.texcoords.size() / 2) return {};
    const auto& p = model.mesh.texcoords;
    std::vector<Vec2> v(p.size() / 2);
    for (size_t i = 0; i < v.size(); i++) v[i] = {p[2 * i], p[2 * i + 1]};
    return v;
}

std::vector<uint32_t> ObjScene::indices() const
{
    const auto& ix = model.mesh.indices;
    std::vector<uint32_t> out(ix.size());
    for (size_t t = 0; t + 2 < ix.size(); t += 3) { out[t] = ix[t + 2]; out[t + 1] = ix[t + 1]; out[t + 2] = ix[t]; }
    return out;
}

void ObjScene::tbn(std::vector<Vec3>& T, std::vector<Vec3>& B, std::vector<Vec3>& N) const
{
    std::vector<Vec3> pos = vertices();
    std::vector<Vec2> uv = texcoords();
    if (uv.size() != pos.size()) uv.assign(pos.size(), Vec2{0.f, 0.f});
    const size_t nv = pos.size();
    T.assign(nv, Vec3{0, 0, 0});
    B.assign(nv, Vec3{0, 0, 0});
    N.assign(nv, Vec3{0, 0, 0});
    std::vector<int> count(nv, 0);
    std::vector<uint32_t> ix = indices();
    for (size_t t = 0; t + 2 < ix.size(); t += 3) {
        uint32_t c0 = ix[t], c1 = ix[t + 1], c2 = ix[t + 2];
        Vec3 dp1 = pos[c1] - pos[c0], dp2 = pos[c2] - pos[c0];
        const float s = 2048.0f;  // 2.0f32.powi(11)
        Vec2 d1{(uv[c1].x - uv[c0].x) * s, (uv[c1].y - uv[c0].y) * s};
        Vec2 d2{(uv[c2].x - uv[c0].x) * s, (uv[c2].y - uv[c0].y) * s};
        // glam Mat2::inverse of mat2(col0 = d1, col1 = d2)
        float det = d1.x * d2.y - d1.y * d2.x;
        float inv = 1.0f / det;
        float r00 = d2.y * inv, r01 = d1.y * -inv, r10 = d2.x * -inv, r11 = d1.x * inv;
        Vec3 tangent = r00 * dp1 - r01 * dp2;
        Vec3 bitangent = (-r10) * dp1 + r11 * dp2;
        Vec3 normal = normalize(cross(bitangent, tangent));
        if (!is_nan(tangent) && !is_nan(bitangent) && !is_nan(normal)) {
            for (uint32_t c : {c0, c1, c2}) {
                T[c] = T[c] + tangent;
                B[c] = B[c] + bitangent;
                N[c] = N[c] + normal;
                count[c]++;
            }
        }
    }
    for (size_t i = 0; i < nv; i++) {
        if (count[i] > 0) {
            float c = (float)count[i];
            T[i] = normalize(T[i] / c);
            B[i] = normalize(B[i] / c);
            N[i] = normalize(N[i] / c);
        } else {
            T[i] = {1, 0, 0};
            B[i] = {0, 1, 0};
            N[i] = {0, 0, 1};
        }
    }
}

std::vector<float> ObjScene::vertex_stream() const
{
    std::vector<Vec3> pos = vertices(), col = vertex_colors(), nrm = normals(), T, B, N;
    std::vector<Vec2> uv = texcoords();
    tbn(T, B, N);
    const size_t nv = pos.size();
    std::vector<float> out(nv * 17);
    for (size_t i = 0; i < nv; i++) {
        float* o = &out[i * 17];
        Vec3 c = i < col.size() ? col[i] : Vec3{1, 1, 1};
        // zip_longest(normals, tbn normal): OBJ normal while it lasts, then the tbn normal, then Z
        Vec3 n = i < nrm.size() ? nrm[i] : (i < N.size() ? N[i] : Vec3{0, 0, 1});
        Vec3 t = i < T.size() ? T[i] : Vec3{1, 0, 0};
        Vec3 b = i < B.size() ? B[i] : Vec3{0, 1, 0};
        Vec2 tc = i < uv.size() ? uv[i] : Vec2{0, 0};
        o[0] = pos[i].x; o[1] = pos[i].y; o[2] = pos[i].z;
        o[3] = c.x; o[4] = c.y; o[5] = c.z;
        o[6] = n.x; o[7] = n.y; o[8] = n.z;
        o[9] = t.x; o[10] = t.y; o[11] = t.z;
        o[12] = b.x; o[13] = b.y; o[14] = b.z;
        o[15] = tc.x; o[16] = tc.y;
    }
    return out;
}

std::optional<Material> ObjScene::material(bool decode_textures) const
{
    if (!materials) return std::nullopt;
    const TobjMaterial& e = *materials;
    Material m;
    m.ambient = e.ambient;
    m.diffuse = e.diffuse;
    m.specular = e.specular;
    m.shininess = e.shininess;
    m.emission = e.emission;
    if (decode_textures) {
        std::string why;
        if (e.diffuse_texture) {
            m.color_texture = load_image_rgba8(obj_dir + "/" + *e.diffuse_texture, &why);
            if (!m.color_texture) fprintf(stderr, "[rc_b200] warn: failed to open color texture: %s\n", why.c_str());
        }
        if (e.normal_texture) {
            m.normal_texture = load_image_rgba8(obj_dir + "/" + *e.normal_texture, &why);
            if (!m.normal_texture) fprintf(stderr, "[rc_b200] warn: failed to open normal texture: %s\n", why.c_str());
        }
    }
    return m;
}

UniformMaterial to_uniform(const std::optional<Material>& m)
{
    UniformMaterial u;
    memset(&u, 0, sizeof(u));
    auto put = [](float* dst, const std::optional<Vec3>& v) {
        if (v) { dst[0] = v->x; dst[1] = v->y; dst[2] = v->z; dst[3] = 1.0f; }
    };
    if (m) {
        put(u.ambient, m->ambient);
        put(u.diffuse, m->diffuse);
        put(u.specular, m->specular);
        u.shininess = m->shininess ? *m->shininess : 1.0f;
    } else {
        u.shininess = 1.0f;  // Material::default()
    }
    return u;
}

}  // namespace rc
