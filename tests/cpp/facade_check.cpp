// facade_check.cpp — exercises include/rc_b200.hpp (the C++ host façade with the reference's type names)
// and prints its results as one JSON object; tests/test_cpp_facade.py compiles and runs it and compares
// them with the committed golden vectors and with the Python mirror.   usage: facade_check CUBE.obj
#include <cstdio>
#include <string>

#include "rc_b200.hpp"

static std::string hex(const void* p, size_t n)
{
    static const char* d = "0123456789abcdef";
    std::string s;
    for (size_t i = 0; i < n; i++) { const unsigned char c = ((const unsigned char*)p)[i]; s += d[c >> 4]; s += d[c & 15]; }
    return s;
}

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    printf("{");
    // AppState::new + the clamp update_camera applies before the first frame
    rc::AppState st;
    st.camera_controller.update_camera(st.camera, 1.0f / 60.0f);
    printf("\"default_pitch\": %.9g, ", st.camera.pitch);
    st.projection.resize(1360, 1360);
    rc::UniformCamera u = rc::UniformCamera::from_camera_project(st.camera, st.projection);
    printf("\"default_1360x1360\": \"%s\", ", hex(&u, 80).c_str());
    st.projection.resize(1360, 768);
    u = rc::UniformCamera::from_camera_project(st.camera, st.projection);
    printf("\"default_1360x768\": \"%s\", ", hex(&u, 80).c_str());
    // look-at path, golden orbit frame 5 of the cube
    const rc::Projection proj(1920, 1080, 45.0f, 0.1f, 13.856406211853027f);
    u = rc::UniformCamera::look_at({2.2912986278533936f, 0.8660253882408142f, 1.2247246503829956f}, {0.f, 0.f, 0.f}, proj);
    printf("\"orbit5_cube_1920x1080\": \"%s\", ", hex(&u, 80).c_str());
    const rc::Mat4 view = rc::Camera({1.f, 2.f, 3.f}, 0.3f, -0.2f).calc_matrix(), pm = proj.calc_matrix();
    printf("\"view\": \"%s\", \"proj\": \"%s\", ", hex(view.data(), 64).c_str(), hex(pm.data(), 64).c_str());
    // scripted CameraController: W + D held, mouse moved, one scroll line, three frames of 1/60 s
    rc::Camera cam({0.5f, 1.0f, -2.0f}, 0.7f, 0.1f);
    rc::CameraController cc(4.0f, 0.4f);
    cc.process_keyboard(rc::CameraController::Key::Forward, true);
    cc.process_keyboard(rc::CameraController::Key::Right, true);
    cc.process_keyboard(rc::CameraController::Key::Up, true);
    cc.process_scroll_lines(1.0f);
    for (int i = 0; i < 3; i++) { cc.process_mouse(12.0, -7.0); cc.update_camera(cam, 1.0f / 60.0f); }
    printf("\"walk\": [%.9g, %.9g, %.9g, %.9g, %.9g], ", cam.position[0], cam.position[1], cam.position[2], cam.yaw, cam.pitch);
    rc::Camera up({0.f, 0.f, 0.f}, 0.f, 1.5f);
    cc.process_mouse(0.0, -1000.0);
    cc.update_camera(up, 1.0f);
    printf("\"clamped_pitch\": %.9g, ", up.pitch);
    printf("\"light\": \"%s\", ", hex(&static_cast<const rc_light&>(rc::UniformLight({1.f, 2.f, 3.f})), 16).c_str());
    // device-free ingest
    rc::ObjScene sc = rc::ObjScene::load(argv[1]);
    const rc_scene_info si = sc.info();
    auto ms = sc.model_stream(0);
    const rc::ObjScene::Material mat = sc.model_material(0);
    printf("\"models\": %u, \"vertices\": %u, \"triangles\": %u, \"stream_floats\": %zu, \"indices\": %zu, \"enable_bit\": %u, \"name\": \"%s\", ",
           si.num_models, si.num_vertices, si.num_triangles, ms.first.size(), ms.second.size(), mat.enable_bit, sc.model_name(0).c_str());
    bool threw = false;
    try { rc::ObjScene::load("/nonexistent/nothing.obj"); } catch (const rc::Error& e) { threw = e.status == RC_ERR_SCENE_LOAD; }
    printf("\"missing_scene_throws\": %s, ", threw ? "true" : "false");
    // the renderer needs a GPU: report what happened, the test knows which outcome to expect
    int status = 0;
    unsigned launches = 0;
    try {
        rc::AppState s2;
        rc::DefaultRenderer r(0, 64, 64, s2, argv[1]);
        s2.uniform_camera = u;
        r.update(s2);
        r.render(s2);
        r.synchronize();
        launches = r.launch_count();
    } catch (const rc::Error& e) { status = (int)e.status; }
    printf("\"renderer_status\": %d, \"launches\": %u}\n", status, launches);
    return 0;
}
