// scene.h — host-side scene ingest, mirroring the reference's ObjScene / Material /
// Scene-trait surface (src/primitives.rs) in C++.  Rust is not available in this
// image, so this is the "host side above the C ABI" for the compiled reference.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <optional>
#include <string>
#include <vector>

namespace rc {

struct Vec2 { float x, y; };
struct Vec3 { float x, y, z; };

// tobj::Material subset the reference reads (src/primitives.rs:386-415) + Ke for GI emission.
struct TobjMaterial {
    std::string name;
    std::optional<Vec3> ambient, diffuse, specular;
    std::optional<float> shininess;
    std::optional<std::string> diffuse_texture, normal_texture;
    Vec3 emission{0.f, 0.f, 0.f};
};

// tobj::Mesh with LoadOptions{triangulate, single_index} (src/primitives.rs:104-113).
struct TobjMesh {
    std::vector<float> positions, normals, texcoords, vertex_color;
    std::vector<uint32_t> indices;
    std::optional<size_t> material_id;
};
struct TobjModel { TobjMesh mesh; std::string name; };

struct Image { uint32_t width = 0, height = 0; std::vector<uint8_t> rgba; };  // to_rgba8()

// src/primitives.rs:75-83
struct Material {
    std::optional<Vec3> ambient, diffuse, specular;
    std::optional<float> shininess;
    std::shared_ptr<Image> color_texture, normal_texture;
    Vec3 emission{0.f, 0.f, 0.f};
};

// src/primitives.rs:37-46 (64 bytes)
struct UniformMaterial {
    float ambient[4], diffuse[4], specular[4];
    float shininess;
    uint32_t _padding[3];
};
UniformMaterial to_uniform(const std::optional<Material>& m);  // :48-73

struct LoadError { std::string message; };

// src/primitives.rs:115-416
class ObjScene {
public:
    TobjModel model;
    std::string obj_dir;
    std::shared_ptr<TobjMaterial> materials;

    // ObjScene::load(path, light_predicate) -> (Vec<Self>, Option<Vec3>)   :122-175
    static bool load(const std::string& path, const std::function<bool(const TobjMaterial&)>& light_predicate,
                     std::vector<ObjScene>& out, std::optional<Vec3>& light, LoadError& err);

    std::vector<Vec3> vertices() const;        // :218-225
    std::vector<Vec3> vertex_colors() const;   // :227-234
    std::vector<Vec3> normals() const;         // :236-243
    void tbn(std::vector<Vec3>& t, std::vector<Vec3>& b, std::vector<Vec3>& n) const;  // :245-354
    std::vector<Vec2> texcoords() const;       // :356-367
    std::vector<uint32_t> indices() const;     // :369-376 (winding reversed)
    uint32_t vertex_count() const { return (uint32_t)model.mesh.indices.size(); }      // :378-380
    const std::string& name() const { return model.name; }
    std::optional<Material> material(bool decode_textures = true) const;               // :386-415

    // The 17-float interleaved stream DefaultRenderer::new builds (src/renderer.rs:371-410).
    std::vector<float> vertex_stream() const;
};

// image decode stand-in for image::ImageReader::open(p).decode().to_rgba8()
// (src/primitives.rs:391-404): "<file>.rgba8" sidecar, else built-in PNG / JPEG decoders.
std::shared_ptr<Image> load_image_rgba8(const std::string& path, std::string* why);

}  // namespace rc
