// rc_b200.hpp — header-only C++17 host façade over the C ABI of librc_b200.so (include/rc_b200.h).
//
// The reference (jw910731/RadianceCascade) is a Rust crate; Rust is not available in this image, so the host
// side above the C ABI is C++ and keeps the reference's type names, argument meaning and error behaviour:
//
//   rc::Camera, rc::Projection, rc::UniformCamera      src/camera.rs:9-80
//   rc::CameraController                               src/camera.rs:82-200 (scripted: no winit events)
//   rc::UniformLight                                   src/primitives.rs:14-35
//   rc::AppState                                       src/app.rs:9-37
//   rc::RenderStage<T>                                 src/app.rs:3-7   (the reference's only plugin seam)
//   rc::DefaultRenderer : RenderStage<AppState>        src/renderer.rs:156-632
//   rc::ObjScene                                       src/primitives.rs:122-175, 218-415
//
// Nothing here computes a pixel or a matrix: every number comes out of librc_b200.so.  Where the reference
// panics (`.unwrap()`, src/renderer.rs:176) this façade throws rc::Error carrying the rc_status and the
// library's message; there is no CPU fallback (RC_ERR_NO_DEVICE without a CUDA device).
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "rc_b200.h"

namespace rc {

struct Error : std::runtime_error {
    rc_status status;
    Error(rc_status s, const std::string& what) : std::runtime_error(what), status(s) {}
};

using Vec3 = std::array<float, 3>;
using Mat4 = std::array<float, 16>;   // column-major (glam)

// ------------------------------------------------------------------ src/camera.rs
// SAFE_FRAC_PI_2 (src/camera.rs:25): f32 pi/2 minus f32 1e-4
inline float safe_frac_pi_2() { return 1.57079632679489661923f - 0.0001f; }

// src/camera.rs:27-53.  yaw / pitch are used as radians exactly as stored (AppState::new passes degrees:
// SURVEY Appendix B quirk 1 — reproduced, not fixed).
struct Camera {
    Vec3 position{0.f, 0.f, 0.f};
    float yaw = 0.f, pitch = 0.f;

    Camera() = default;
    Camera(Vec3 pos, float yaw_, float pitch_) : position(pos), yaw(yaw_), pitch(pitch_) {}

    Mat4 calc_matrix() const   // src/camera.rs:43-52
    {
        Mat4 m{};
        rc_camera_view_matrix(position.data(), yaw, pitch, m.data());
        return m;
    }
};

// src/camera.rs:55-80; the constructor takes fovy in DEGREES and stores radians (Projection::new).
struct Projection {
    float aspect = 1.f, fovy = 0.78539816339744830962f, znear = 0.1f, zfar = 100.f;

    Projection() = default;
    Projection(uint32_t width, uint32_t height, float fovy_deg, float znear_, float zfar_)
        : aspect((float)width / (float)height), fovy(fovy_deg * (3.14159265358979323846f / 180.0f)), znear(znear_), zfar(zfar_) {}

    void resize(uint32_t width, uint32_t height) { aspect = (float)width / (float)height; }   // src/camera.rs:73-75

    Mat4 calc_matrix() const   // src/camera.rs:77-79
    {
        Mat4 m{};
        rc_projection_matrix(fovy, aspect, znear, zfar, m.data());
        return m;
    }
};

// src/camera.rs:9-23: 80 bytes = proj*view (column-major) + (eye, 1)
struct UniformCamera : rc_camera {
    static UniformCamera from_camera_project(const Camera& c, const Projection& p)
    {
        UniformCamera u{};
        rc_uniform_camera(c.position.data(), c.yaw, c.pitch, p.fovy, p.aspect, p.znear, p.zfar, &u);
        return u;
    }
    // synthetic camera paths: the same look_to_rh arithmetic with dir = normalize(target - position)
    static UniformCamera look_at(const Vec3& position, const Vec3& target, const Projection& p)
    {
        UniformCamera u{};
        rc_uniform_camera_look_at(position.data(), target.data(), p.fovy, p.aspect, p.znear, p.zfar, &u);
        return u;
    }
};
static_assert(sizeof(UniformCamera) == 80, "UniformCamera must keep the reference's 80-byte layout");

// src/primitives.rs:14-35: (x, y, z, 1)
struct UniformLight : rc_light {
    UniformLight() : rc_light{{0.f, 0.f, 0.f, 1.f}} {}
    explicit UniformLight(const Vec3& p) : rc_light{{p[0], p[1], p[2], 1.f}} {}
};
static_assert(sizeof(UniformLight) == 16, "UniformLight must keep the reference's 16-byte layout");

// src/camera.rs:82-200 without the winit event types: the driver sets the same eleven fields the event
// handlers set (process_keyboard / process_mouse / process_scroll) and calls update_camera once per frame.
class CameraController {
public:
    enum class Key { Forward, Backward, Left, Right, Up, Down };   // W, S, A, D, Space, ShiftLeft (src/camera.rs:125-151)

    CameraController() = default;
    CameraController(float speed, float sensitivity) : speed_(speed), sensitivity_(sensitivity) {}

    bool process_keyboard(Key k, bool pressed)
    {
        const float amount = pressed ? 1.0f : 0.0f;
        switch (k) {
        case Key::Forward: forward_ = amount; return true;
        case Key::Backward: backward_ = amount; return true;
        case Key::Left: left_ = amount; return true;
        case Key::Right: right_ = amount; return true;
        case Key::Up: up_ = amount; return true;
        case Key::Down: down_ = amount; return true;
        }
        return false;
    }
    void process_mouse(double dx, double dy) { rot_h_ = (float)dx; rot_v_ = (float)dy; }
    // LineDelta(_, lines): one line counts as 100 pixels, sign flipped (src/camera.rs:161-167)
    void process_scroll_lines(float lines) { scroll_ = -(lines * 100.0f); }
    void process_scroll_pixels(double y) { scroll_ = -(float)y; }

    // src/camera.rs:170-199; dt in seconds.  Plain f32 arithmetic in the reference's order.
    void update_camera(Camera& cam, float dt)
    {
        const float ys = std::sin(cam.yaw), yc = std::cos(cam.yaw);
        const Vec3 fwd = normalized({yc, 0.0f, ys}), right = normalized({-ys, 0.0f, yc});
        add_scaled(cam.position, fwd, forward_ - backward_, speed_, dt);
        add_scaled(cam.position, right, right_ - left_, speed_, dt);
        const float ps = std::sin(cam.pitch), pc = std::cos(cam.pitch);
        const Vec3 toward = normalized({pc * yc, ps, pc * ys});
        add_scaled(cam.position, toward, scroll_, speed_, sensitivity_, dt);
        scroll_ = 0.0f;
        cam.position[1] += ((up_ - down_) * speed_) * dt;
        cam.yaw += (rot_h_ * sensitivity_) * dt;
        cam.pitch += (-rot_v_ * sensitivity_) * dt;
        rot_h_ = rot_v_ = 0.0f;
        const float lim = safe_frac_pi_2();
        if (cam.pitch < -lim) cam.pitch = -lim;
        else if (cam.pitch > lim) cam.pitch = lim;
    }

private:
    static Vec3 normalized(Vec3 v)   // glam Vec3::normalize: v * (1 / sqrt(dot))
    {
        const float r = 1.0f / std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
        return {v[0] * r, v[1] * r, v[2] * r};
    }
    // `position += dir * a * b * c` multiplies the vector by one scalar at a time, left to right
    static void add_scaled(Vec3& p, const Vec3& d, float a, float b, float c)
    {
        for (int i = 0; i < 3; i++) p[i] += ((d[i] * a) * b) * c;
    }
    static void add_scaled(Vec3& p, const Vec3& d, float a, float b, float c, float e)
    {
        for (int i = 0; i < 3; i++) p[i] += (((d[i] * a) * b) * c) * e;
    }

    float left_ = 0.f, right_ = 0.f, forward_ = 0.f, backward_ = 0.f, up_ = 0.f, down_ = 0.f;
    float rot_h_ = 0.f, rot_v_ = 0.f, scroll_ = 0.f, speed_ = 0.f, sensitivity_ = 0.f;
};

// ------------------------------------------------------------------ src/app.rs
// AppState::new (src/app.rs:24-37) minus the egui text boxes; `extra_lights` and `uniform_camera` are additions
// (the reference supports one light and always derives the uniform from camera + projection).
struct AppState {
    Camera camera{{0.0f, 5.0f, 10.0f}, -90.0f, -20.0f};
    Projection projection{1, 1, 45.0f, 0.1f, 100.0f};
    CameraController camera_controller{4.0f, 0.4f};
    bool enable_normal_map = true, normal_map_changed = false, given_light_position = false;
    Vec3 light_position{0.f, 0.f, 0.f};
    std::vector<Vec3> extra_lights;
    std::optional<UniformCamera> uniform_camera;
};

// The reference's plugin seam (src/app.rs:3-7).  wgpu's TextureView / CommandEncoder / Device / Queue become
// a caller-owned CUDA stream (opaque pointer, nullptr = the stage's own) and plain sizes.
template <class T>
struct RenderStage {
    virtual ~RenderStage() = default;
    virtual void render(T& state, void* stream) = 0;
    virtual void resize(uint32_t width, uint32_t height) = 0;
    virtual void update(const T& state) = 0;
};

// include/rc_spec.h parameters; zeros mean the defaults
struct CascadeConfig {
    uint32_t probe_spacing0 = 0, dir_res0 = 0, num_levels = 0;
    float interval0 = 0.f, t_far = 0.f, normal_offset = 0.f;
    Vec3 sky{0.f, 0.f, 0.f};
    uint32_t flags = 0;
    uint32_t tile_x0 = 0, tile_y0 = 0, tile_w = 0, tile_h = 0;
};

// ≙ DefaultRenderer (src/renderer.rs:156-632) behind rc_create / rc_update / rc_resize / rc_render
class DefaultRenderer : public RenderStage<AppState> {
public:
    // DefaultRenderer::new(device, config, queue, state, path)   src/renderer.rs:168-174
    DefaultRenderer(int device, uint32_t width, uint32_t height, AppState& state, const std::string& path,
                    const CascadeConfig& cc = {}, const char* resource_root = nullptr)
    {
        rc_config cfg{};
        cfg.struct_size = (uint32_t)sizeof(rc_config);
        cfg.width = width; cfg.height = height; cfg.device = device;
        cfg.scene_path = path.c_str();
        cfg.resource_root = resource_root;
        cfg.probe_spacing0 = cc.probe_spacing0; cfg.dir_res0 = cc.dir_res0; cfg.num_levels = cc.num_levels;
        cfg.interval0 = cc.interval0; cfg.t_far = cc.t_far; cfg.normal_offset = cc.normal_offset;
        for (int i = 0; i < 3; i++) cfg.sky[i] = cc.sky[i];
        cfg.flags = cc.flags;
        cfg.tile_x0 = cc.tile_x0; cfg.tile_y0 = cc.tile_y0; cfg.tile_w = cc.tile_w; cfg.tile_h = cc.tile_h;
        const rc_status st = rc_create(&cfg, &ctx_);
        if (st != RC_OK) {
            const char* m = rc_last_error(nullptr);
            throw Error(st, std::string("rc_create: ") + (m ? m : "failed"));
        }
        state.given_light_position = scene_info().light_from_obj != 0;   // src/renderer.rs:177
    }
    ~DefaultRenderer() override { if (ctx_) rc_destroy(ctx_); }
    DefaultRenderer(const DefaultRenderer&) = delete;
    DefaultRenderer& operator=(const DefaultRenderer&) = delete;
    DefaultRenderer(DefaultRenderer&& o) noexcept : ctx_(o.ctx_) { o.ctx_ = nullptr; }

    // RenderStage::update + the two queue.write_buffer calls of AppInternal::update (src/window/app.rs:112-131)
    void update(const AppState& s) override
    {
        const UniformCamera uc = s.uniform_camera ? *s.uniform_camera : UniformCamera::from_camera_project(s.camera, s.projection);
        std::vector<UniformLight> lights;
        lights.emplace_back(s.light_position);
        for (const Vec3& p : s.extra_lights) lights.emplace_back(p);
        check(rc_update(ctx_, &uc, lights.data(), (uint32_t)lights.size(), s.enable_normal_map ? RC_UPD_ENABLE_NORMAL_MAP : 0u), "rc_update");
    }
    // RenderStage::resize (src/renderer.rs:615-618)
    void resize(uint32_t width, uint32_t height) override { check(rc_resize(ctx_, width, height), "rc_resize"); }
    // multi-GPU: move this context's screen-space tile inside the frame (w == h == 0: the full frame)
    void set_tile(uint32_t x0, uint32_t y0, uint32_t w, uint32_t h) { check(rc_set_tile(ctx_, x0, y0, w, h), "rc_set_tile"); }
    // RenderStage::render (src/renderer.rs:559-613): enqueues only, the caller owns the stream
    void render(AppState&, void* stream = nullptr) override { check(rc_render(ctx_, stream), "rc_render"); }

    void synchronize() { check(rc_synchronize(ctx_), "rc_synchronize"); }

    size_t target_bytes(rc_target which) const
    {
        size_t n = 0;
        check(rc_target_bytes(ctx_, which, &n), "rc_target_bytes");
        return n;
    }
    void read_target(rc_target which, void* dst, size_t bytes) { check(rc_read_target(ctx_, which, dst, bytes), "rc_read_target"); }
    std::vector<uint8_t> read_target(rc_target which)
    {
        std::vector<uint8_t> out(target_bytes(which));
        read_target(which, out.data(), out.size());
        return out;
    }

    std::array<float, RC_STAGE_COUNT> stage_times()
    {
        std::array<float, RC_STAGE_COUNT> ms{};
        check(rc_stage_times(ctx_, ms.data(), RC_STAGE_COUNT), "rc_stage_times");
        return ms;
    }
    std::vector<rc_level_info> levels() const
    {
        rc_level_info lv[16];
        uint32_t n = 0;
        check(rc_get_levels(ctx_, lv, 16, &n), "rc_get_levels");
        return std::vector<rc_level_info>(lv, lv + n);
    }
    rc_scene_info scene_info() const
    {
        rc_scene_info i{};
        check(rc_get_scene_info(ctx_, &i), "rc_get_scene_info");
        return i;
    }
    std::array<uint32_t, 4> tile() const
    {
        std::array<uint32_t, 4> t{};
        check(rc_get_tile(ctx_, t.data()), "rc_get_tile");
        return t;
    }
    std::vector<uint32_t> rays_marched()
    {
        std::vector<uint32_t> r(levels().size());
        check(rc_rays_marched(ctx_, r.data(), (uint32_t)r.size()), "rc_rays_marched");
        return r;
    }
    uint32_t launch_count() const
    {
        uint32_t n = 0;
        check(rc_launch_count(ctx_, &n), "rc_launch_count");
        return n;
    }
    void set_tuning(const char* key, int value) { check(rc_set_tuning(ctx_, key, value), "rc_set_tuning"); }
    rc_ctx* handle() const { return ctx_; }

private:
    void check(rc_status st, const char* what) const
    {
        if (st == RC_OK) return;
        const char* m = rc_last_error(ctx_);
        throw Error(st, std::string(what) + ": " + (m ? m : "failed"));
    }
    rc_ctx* ctx_ = nullptr;
};

// Device-free scene ingest (≙ ObjScene::load + the per-model preparation of DefaultRenderer::new,
// src/primitives.rs:122-175, src/renderer.rs:370-497).  One ObjScene here holds all models of the file.
class ObjScene {
public:
    static ObjScene load(const std::string& path, bool no_textures = false)
    {
        rc_scene* s = nullptr;
        const rc_status st = rc_scene_load(path.c_str(), no_textures ? RC_CFG_NO_TEXTURES : 0u, &s);
        if (st != RC_OK) {
            const char* m = rc_last_error(nullptr);
            throw Error(st, std::string("rc_scene_load: ") + (m ? m : "failed"));
        }
        return ObjScene(s);
    }
    ~ObjScene() { if (s_) rc_scene_free(s_); }
    ObjScene(ObjScene&& o) noexcept : s_(o.s_) { o.s_ = nullptr; }
    ObjScene(const ObjScene&) = delete;
    ObjScene& operator=(const ObjScene&) = delete;

    rc_scene_info info() const
    {
        rc_scene_info i{};
        if (rc_scene_get_info(s_, &i) != RC_OK) throw Error(RC_ERR_INVALID_ARG, "rc_scene_get_info");
        return i;
    }
    // (17-float interleaved vertex stream, winding-reversed index buffer) of model m — src/renderer.rs:371-420
    std::pair<std::vector<float>, std::vector<uint32_t>> model_stream(uint32_t m)
    {
        uint32_t nv = 0, ni = 0;
        if (rc_scene_model_stream(s_, m, nullptr, 0, nullptr, 0, &nv, &ni) != RC_OK) throw Error(RC_ERR_INVALID_ARG, "rc_scene_model_stream");
        std::vector<float> v((size_t)nv * 17);
        std::vector<uint32_t> i(ni);
        if (rc_scene_model_stream(s_, m, v.data(), v.size() * 4, i.data(), i.size() * 4, &nv, &ni) != RC_OK)
            throw Error(RC_ERR_INVALID_ARG, "rc_scene_model_stream");
        return {std::move(v), std::move(i)};
    }
    // UniformMaterial (16 floats, src/primitives.rs:37-73), enable_bit, Ke
    struct Material { std::array<float, 16> uniform; uint32_t enable_bit; Vec3 ke; };
    Material model_material(uint32_t m) const
    {
        float raw[20];
        if (rc_scene_model_material(s_, m, raw, sizeof(raw)) != RC_OK) throw Error(RC_ERR_INVALID_ARG, "rc_scene_model_material");
        Material out{};
        std::memcpy(out.uniform.data(), raw, 64);
        std::memcpy(&out.enable_bit, raw + 16, 4);
        out.ke = {raw[17], raw[18], raw[19]};
        return out;
    }
    std::string model_name(uint32_t m) const
    {
        char buf[512] = {0};
        rc_scene_model_name(s_, m, buf, sizeof(buf));
        return buf;
    }

private:
    explicit ObjScene(rc_scene* s) : s_(s) {}
    rc_scene* s_ = nullptr;
};

}  // namespace rc
