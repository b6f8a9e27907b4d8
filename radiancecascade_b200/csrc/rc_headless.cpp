// rc_headless.cpp — headless frame driver over the C++ façade (include/rc_b200.hpp).
//
// Stands in for the reference's entry point and window loop — src/main.rs:12-22 and
// App::handle_redraw (src/window/app.rs:221-267) — with the winit surface replaced by an offscreen
// target: per frame  controller.update_camera -> stage.update -> stage.render -> read the target,
// in the reference's order, along a synthetic camera path instead of keyboard / mouse input.
// Prints one JSON line per frame (CUDA-event stage times from the library) and a summary line.
// No GPU => rc::Error(RC_ERR_NO_DEVICE), exit code 3: there is no CPU fallback.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/rc_b200.hpp"

namespace {

struct Options {
    std::string scene, camera = "orbit", light = "bench", target = "irradiance", out_raw, out_img;
    uint32_t width = 1360, height = 1360;   // the reference's initial surface (src/window/app.rs:204-210)
    int frames = 1, first_frame = 0, orbit_frames = 64, device = 0, warmup = 0;
    float fovy = 45.0f, dt = 1.0f / 60.0f;
    rc::Vec3 eye{0.f, 5.f, 10.f}, look{0.f, 0.f, 0.f};
    bool have_eye = false, normal_map = true, room_lights = false, quiet = false;
    float walk[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // forward, right, up (-1..1), mouse dx, mouse dy per frame
    rc::CascadeConfig cascade;
};

void usage(FILE* f)
{
    fprintf(f,
            "usage: rc_headless --scene FILE.obj [options]\n"
            "  --size WxH            surface size (default 1360x1360, the reference's initial surface)\n"
            "  --frames N            frames to render (default 1);  --first-frame K;  --warmup N (untimed, not written)\n"
            "  --camera orbit        SURVEY 8d orbit about the scene's bounding box, 64 azimuth steps (default)\n"
            "  --camera default      AppState::new camera driven by the scripted CameraController (--walk)\n"
            "  --camera lookat --eye x,y,z --target x,y,z\n"
            "  --walk f,r,u,dx,dy    held controller input: forward/right/up in -1..1, mouse dx/dy per frame;  --dt seconds\n"
            "  --fovy DEG            vertical field of view (default 45)\n"
            "  --light bench | origin | room | x,y,z     bench = bbox centre + 0.4*height (default); origin = the\n"
            "                        reference's default (0,0,0); room = four lights at the upper quarter points\n"
            "  --no-normal-map       AppState::enable_normal_map = false\n"
            "  --spacing P0  --dirs D0  --levels N       cascade parameters (include/rc_spec.h), 0 = default\n"
            "  --read TARGET         irradiance (default) | composite | direct_srgb8 | direct | albedo | depth | normal | prim\n"
            "  --out-raw PREFIX      write every frame's target bytes verbatim to PREFIX_%%04d.bin\n"
            "  --out-image PREFIX    8-bit targets -> PREFIX_%%04d.ppm, irradiance/direct/albedo -> PREFIX_%%04d.pfm\n"
            "  --device N            CUDA ordinal (default 0);  --quiet  only the summary line\n");
}

bool parse_vec3(const char* s, rc::Vec3& v) { return sscanf(s, "%f,%f,%f", &v[0], &v[1], &v[2]) == 3; }

bool parse(int argc, char** argv, Options& o)
{
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto need = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { fprintf(stderr, "rc_headless: %s needs a value\n", name); return nullptr; }
            return argv[++i];
        };
        const char* v = nullptr;
        if (a == "--help" || a == "-h") { usage(stdout); exit(0); }
        else if (a == "--no-normal-map") o.normal_map = false;
        else if (a == "--quiet") o.quiet = true;
        else if (a == "--scene") { if (!(v = need("--scene"))) return false; o.scene = v; }
        else if (a == "--size") { if (!(v = need("--size")) || sscanf(v, "%ux%u", &o.width, &o.height) != 2) return false; }
        else if (a == "--frames") { if (!(v = need("--frames"))) return false; o.frames = atoi(v); }
        else if (a == "--first-frame") { if (!(v = need("--first-frame"))) return false; o.first_frame = atoi(v); }
        else if (a == "--warmup") { if (!(v = need("--warmup"))) return false; o.warmup = atoi(v); }
        else if (a == "--camera") { if (!(v = need("--camera"))) return false; o.camera = v; }
        else if (a == "--eye") { if (!(v = need("--eye")) || !parse_vec3(v, o.eye)) return false; o.have_eye = true; }
        else if (a == "--target") { if (!(v = need("--target")) || !parse_vec3(v, o.look)) return false; }
        else if (a == "--walk") { if (!(v = need("--walk")) || sscanf(v, "%f,%f,%f,%f,%f", &o.walk[0], &o.walk[1], &o.walk[2], &o.walk[3], &o.walk[4]) != 5) return false; }
        else if (a == "--dt") { if (!(v = need("--dt"))) return false; o.dt = (float)atof(v); }
        else if (a == "--fovy") { if (!(v = need("--fovy"))) return false; o.fovy = (float)atof(v); }
        else if (a == "--light") { if (!(v = need("--light"))) return false; o.light = v; }
        else if (a == "--spacing") { if (!(v = need("--spacing"))) return false; o.cascade.probe_spacing0 = (uint32_t)atoi(v); }
        else if (a == "--dirs") { if (!(v = need("--dirs"))) return false; o.cascade.dir_res0 = (uint32_t)atoi(v); }
        else if (a == "--levels") { if (!(v = need("--levels"))) return false; o.cascade.num_levels = (uint32_t)atoi(v); }
        else if (a == "--read") { if (!(v = need("--read"))) return false; o.target = v; }
        else if (a == "--out-raw") { if (!(v = need("--out-raw"))) return false; o.out_raw = v; }
        else if (a == "--out-image") { if (!(v = need("--out-image"))) return false; o.out_img = v; }
        else if (a == "--device") { if (!(v = need("--device"))) return false; o.device = atoi(v); }
        else { fprintf(stderr, "rc_headless: unknown option %s\n", a.c_str()); return false; }
    }
    if (o.scene.empty()) { fprintf(stderr, "rc_headless: --scene is required\n"); return false; }
    if (o.frames < 1 || o.width == 0 || o.height == 0) { fprintf(stderr, "rc_headless: bad --frames / --size\n"); return false; }
    if (o.camera != "orbit" && o.camera != "default" && o.camera != "lookat") { fprintf(stderr, "rc_headless: bad --camera\n"); return false; }
    return true;
}

struct TargetDesc { const char* name; rc_target id; int kind; };   // kind: 0 half4, 1 bgra8, 2 f32, 3 u32
const TargetDesc kTargets[] = {
    {"irradiance", RC_TARGET_IRRADIANCE, 0}, {"direct", RC_TARGET_DIRECT, 0}, {"albedo", RC_TARGET_ALBEDO, 0},
    {"composite", RC_TARGET_COMPOSITE, 1}, {"direct_srgb8", RC_TARGET_DIRECT_SRGB8, 1},
    {"depth", RC_TARGET_DEPTH, 2}, {"normal", RC_TARGET_NORMAL, 3}, {"prim", RC_TARGET_PRIM, 3},
};

float half_to_float(uint16_t h)
{
    const uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31u, m = h & 1023u;
    uint32_t bits;
    if (e == 0) {
        if (m == 0) bits = s;
        else {   // subnormal: renormalise
            int sh = 0;
            uint32_t mm = m;
            while (!(mm & 1024u)) { mm <<= 1; sh++; }
            bits = s | ((uint32_t)(113 - sh) << 23) | ((mm & 1023u) << 13);
        }
    } else if (e == 31) bits = s | 0x7f800000u | (m << 13);
    else bits = s | ((e + 112u) << 23) | (m << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

bool write_file(const std::string& path, const void* data, size_t n, const std::string& header = "")
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "rc_headless: cannot write %s\n", path.c_str()); return false; }
    if (!header.empty()) fwrite(header.data(), 1, header.size(), f);
    const bool ok = fwrite(data, 1, n, f) == n;
    fclose(f);
    return ok;
}

bool write_image(const std::string& prefix, int frame, const TargetDesc& t, const std::vector<uint8_t>& px, uint32_t w, uint32_t h)
{
    char name[32];
    snprintf(name, sizeof(name), "_%04d", frame);
    if (t.kind == 1) {   // BGRA8 (Bgra8UnormSrgb, src/window/app.rs:59-75) -> binary PPM, RGB
        std::vector<uint8_t> rgb((size_t)w * h * 3);
        for (size_t i = 0; i < (size_t)w * h; i++) { rgb[3 * i] = px[4 * i + 2]; rgb[3 * i + 1] = px[4 * i + 1]; rgb[3 * i + 2] = px[4 * i]; }
        return write_file(prefix + name + ".ppm", rgb.data(), rgb.size(), "P6\n" + std::to_string(w) + " " + std::to_string(h) + "\n255\n");
    }
    if (t.kind == 0) {   // RGBA16F -> PFM (float32 RGB, little endian, rows bottom to top)
        std::vector<float> rgb((size_t)w * h * 3);
        const uint16_t* hp = reinterpret_cast<const uint16_t*>(px.data());
        for (uint32_t y = 0; y < h; y++)
            for (uint32_t x = 0; x < w; x++)
                for (int c = 0; c < 3; c++) rgb[((size_t)(h - 1 - y) * w + x) * 3 + c] = half_to_float(hp[((size_t)y * w + x) * 4 + c]);
        return write_file(prefix + name + ".pfm", rgb.data(), rgb.size() * 4, "PF\n" + std::to_string(w) + " " + std::to_string(h) + "\n-1.0\n");
    }
    fprintf(stderr, "rc_headless: --out-image supports the colour targets only (use --out-raw for %s)\n", t.name);
    return false;
}

// SURVEY 8d orbit: radius 0.75*diag about the bbox centre, height +0.25*diag, `n` equal azimuth steps
// (double arithmetic rounded to f32 once, like radiancecascade_b200/scenes.py orbit_camera)
void orbit(const rc_scene_info& si, int frame, int n, rc::Vec3& pos, rc::Vec3& tgt, float& znear, float& zfar)
{
    const double PI = 3.14159265358979323846;
    double c[3], d2 = 0.0;
    for (int i = 0; i < 3; i++) {
        c[i] = 0.5 * ((double)si.bbox_min[i] + (double)si.bbox_max[i]);
        const double e = (double)si.bbox_max[i] - (double)si.bbox_min[i];
        d2 += e * e;
    }
    const double diag = std::sqrt(d2), az = 2.0 * PI * (double)(((frame % n) + n) % n) / (double)n;
    pos = {(float)(c[0] + 0.75 * diag * std::cos(az)), (float)(c[1] + 0.25 * diag), (float)(c[2] + 0.75 * diag * std::sin(az))};
    tgt = {(float)c[0], (float)c[1], (float)c[2]};
    znear = 0.1f;
    zfar = (float)(4.0 * diag);
}

}  // namespace

int main(int argc, char** argv)
{
    Options o;
    if (!parse(argc, argv, o)) { usage(stderr); return 2; }
    const TargetDesc* td = nullptr;
    for (const TargetDesc& t : kTargets) if (o.target == t.name) td = &t;
    if (!td) { fprintf(stderr, "rc_headless: unknown --read target %s\n", o.target.c_str()); return 2; }

    try {
        rc::AppState state;                                       // AppState::new (src/app.rs:24-37)
        state.projection = rc::Projection(o.width, o.height, o.fovy, 0.1f, 100.0f);
        state.enable_normal_map = o.normal_map;
        // DefaultRenderer::new through the reference's plugin seam (src/app.rs:3-7)
        std::unique_ptr<rc::RenderStage<rc::AppState>> stage =
            std::make_unique<rc::DefaultRenderer>(o.device, o.width, o.height, state, o.scene, o.cascade);
        rc::DefaultRenderer& renderer = static_cast<rc::DefaultRenderer&>(*stage);
        const rc_scene_info si = renderer.scene_info();

        // lights: AppInternal::update overwrites the light buffer from AppState every frame (src/window/app.rs:124-130)
        const double cy = 0.5 * ((double)si.bbox_min[1] + (double)si.bbox_max[1]), hy = (double)si.bbox_max[1] - (double)si.bbox_min[1];
        if (o.light == "bench") {
            state.light_position = {(float)(0.5 * ((double)si.bbox_min[0] + si.bbox_max[0])), (float)(cy + 0.4 * hy),
                                    (float)(0.5 * ((double)si.bbox_min[2] + si.bbox_max[2]))};
        } else if (o.light == "origin") {
            state.light_position = {0.f, 0.f, 0.f};
        } else if (o.light == "room") {
            std::vector<rc::Vec3> pts;
            for (double fx : {0.25, 0.75})
                for (double fz : {0.25, 0.75})
                    pts.push_back({(float)(si.bbox_min[0] + fx * ((double)si.bbox_max[0] - si.bbox_min[0])), (float)(si.bbox_min[1] + 0.75 * hy),
                                   (float)(si.bbox_min[2] + fz * ((double)si.bbox_max[2] - si.bbox_min[2]))});
            state.light_position = pts[0];
            state.extra_lights.assign(pts.begin() + 1, pts.end());
        } else if (!parse_vec3(o.light.c_str(), state.light_position)) {
            fprintf(stderr, "rc_headless: bad --light %s\n", o.light.c_str());
            return 2;
        }

        if (o.camera == "default") {   // scripted stand-in for the keyboard / mouse handlers (src/camera.rs:115-168)
            using K = rc::CameraController::Key;
            rc::CameraController& cc = state.camera_controller;
            cc.process_keyboard(K::Forward, o.walk[0] > 0.f); cc.process_keyboard(K::Backward, o.walk[0] < 0.f);
            cc.process_keyboard(K::Right, o.walk[1] > 0.f); cc.process_keyboard(K::Left, o.walk[1] < 0.f);
            cc.process_keyboard(K::Up, o.walk[2] > 0.f); cc.process_keyboard(K::Down, o.walk[2] < 0.f);
        }

        std::vector<uint8_t> px(renderer.target_bytes(td->id));
        double sum_ms[RC_STAGE_COUNT] = {0}, sum_wall = 0.0;
        static const char* kStage[RC_STAGE_COUNT] = {"gbuffer", "probes", "march", "merge", "gather", "frame"};
        for (int it = -o.warmup; it < o.frames; it++) {
            const int frame = o.first_frame + (it < 0 ? 0 : it);
            const auto t0 = std::chrono::steady_clock::now();
            // App::handle_redraw: update (controller -> camera uniform, light uniform) -> render -> present
            if (o.camera == "default") {
                state.camera_controller.process_mouse(o.walk[3], o.walk[4]);
                state.camera_controller.update_camera(state.camera, o.dt);   // also clamps the pitch (src/camera.rs:194-198)
                state.uniform_camera.reset();
            } else {
                rc::Vec3 pos = o.eye, tgt = o.look;
                float zn = 0.1f, zf = 100.0f;
                if (o.camera == "orbit") {
                    orbit(si, frame, o.orbit_frames, pos, tgt, zn, zf);
                } else {   // explicit camera: keep the orbit's near / far planes (far = 4 x the scene diagonal)
                    rc::Vec3 unused_pos, unused_tgt;
                    orbit(si, 0, o.orbit_frames, unused_pos, unused_tgt, zn, zf);
                }
                state.projection = rc::Projection(o.width, o.height, o.fovy, zn, zf);
                state.uniform_camera = rc::UniformCamera::look_at(pos, tgt, state.projection);
            }
            stage->update(state);
            stage->render(state, nullptr);
            renderer.read_target(td->id, px.data(), px.size());     // waits for the frame
            const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (it < 0) continue;
            const auto ms = renderer.stage_times();
            for (int s = 0; s < RC_STAGE_COUNT; s++) sum_ms[s] += ms[s];
            sum_wall += wall;
            if (!o.quiet) {
                printf("{\"frame\": %d, \"wall_ms\": %.4f, \"stage_ms\": {", frame, wall);
                for (int s = 0; s < RC_STAGE_COUNT; s++) printf("%s\"%s\": %.5f", s ? ", " : "", kStage[s], ms[s]);
                printf("}, \"eye\": [%.9g, %.9g, %.9g]}\n", state.uniform_camera ? state.uniform_camera->eye[0] : state.camera.position[0],
                       state.uniform_camera ? state.uniform_camera->eye[1] : state.camera.position[1],
                       state.uniform_camera ? state.uniform_camera->eye[2] : state.camera.position[2]);
            }
            char name[32];
            snprintf(name, sizeof(name), "_%04d.bin", frame);
            if (!o.out_raw.empty() && !write_file(o.out_raw + name, px.data(), px.size())) return 4;
            if (!o.out_img.empty()) {
                const auto tile = renderer.tile();
                if (!write_image(o.out_img, frame, *td, px, tile[2], tile[3])) return 4;
            }
        }
        uint64_t rays = 0;
        for (const rc_level_info& l : renderer.levels()) rays += l.texel_count;
        const double fms = sum_ms[RC_STAGE_FRAME] / o.frames;
        printf("{\"summary\": true, \"scene\": \"%s\", \"width\": %u, \"height\": %u, \"frames\": %d, \"target\": \"%s\", \"models\": %u, "
               "\"triangles\": %u, \"rays_per_frame\": %llu, \"kernel_launches\": %u, \"mean_wall_ms\": %.4f, \"mean_stage_ms\": {",
               o.scene.c_str(), o.width, o.height, o.frames, td->name, si.num_models, si.num_triangles, (unsigned long long)rays,
               renderer.launch_count(), sum_wall / o.frames);
        for (int s = 0; s < RC_STAGE_COUNT; s++) printf("%s\"%s\": %.5f", s ? ", " : "", kStage[s], sum_ms[s] / o.frames);
        printf("}, \"gray_samples_per_s\": %.4f}\n", fms > 0.0 ? (double)rays / (fms * 1e-3) / 1e9 : 0.0);
    } catch (const rc::Error& e) {
        fprintf(stderr, "rc_headless: %s (rc_status %d)\n", e.what(), (int)e.status);
        return 3;
    }
    return 0;
}
