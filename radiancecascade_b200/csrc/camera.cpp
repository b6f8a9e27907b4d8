// camera.cpp — host façade for the reference's Camera / Projection / UniformCamera
// (src/camera.rs:9-80) with glam 0.29.2 conventions (column-major, right-handed,
// depth 0..1).  Plain f32 arithmetic, no fused multiply-add.
#include <cmath>
#include <cstring>

#include "../../include/rc_b200.h"

namespace {

struct V3 { float x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline V3 normalize(V3 a) { float r = 1.0f / std::sqrt(dot(a, a)); return {a.x * r, a.y * r, a.z * r}; }

// glam Mat4::look_to_rh(eye, dir, up)
void look_to_rh(V3 eye, V3 dir, V3 up, float m[16])
{
    V3 f = normalize(dir);
    V3 s = normalize(cross(f, up));
    V3 u = cross(s, f);
    const float cols[16] = {s.x, u.x, -f.x, 0.f, s.y, u.y, -f.y, 0.f, s.z, u.z, -f.z, 0.f,
                            -dot(eye, s), -dot(eye, u), dot(eye, f), 1.f};
    memcpy(m, cols, sizeof(cols));
}

// glam Mat4 * Mat4 (column j = ((A0*b0 + A1*b1) + A2*b2) + A3*b3)
void mat_mul(const float a[16], const float b[16], float out[16])
{
    for (int j = 0; j < 4; j++)
        for (int r = 0; r < 4; r++) {
            float acc = a[0 * 4 + r] * b[j * 4 + 0];
            acc = acc + a[1 * 4 + r] * b[j * 4 + 1];
            acc = acc + a[2 * 4 + r] * b[j * 4 + 2];
            acc = acc + a[3 * 4 + r] * b[j * 4 + 3];
            out[j * 4 + r] = acc;
        }
}

V3 camera_dir(float yaw, float pitch)
{
    float sp = std::sin(pitch), cp = std::cos(pitch), sy = std::sin(yaw), cy = std::cos(yaw);
    return normalize(V3{cp * cy, sp, cp * sy});  // src/camera.rs:44-49
}

}  // namespace

extern "C" {

void rc_camera_view_matrix(const float position[3], float yaw, float pitch, float out16[16])
{
    look_to_rh(V3{position[0], position[1], position[2]}, camera_dir(yaw, pitch), V3{0.f, 1.f, 0.f}, out16);
}

void rc_projection_matrix(float fovy, float aspect, float znear, float zfar, float out16[16])
{
    // glam Mat4::perspective_rh
    float half = 0.5f * fovy;
    float sin_fov = std::sin(half), cos_fov = std::cos(half);
    float h = cos_fov / sin_fov;
    float w = h / aspect;
    float r = zfar / (znear - zfar);
    const float cols[16] = {w, 0.f, 0.f, 0.f, 0.f, h, 0.f, 0.f, 0.f, 0.f, r, -1.f, 0.f, 0.f, r * znear, 0.f};
    memcpy(out16, cols, sizeof(cols));
}

void rc_uniform_camera(const float position[3], float yaw, float pitch, float fovy, float aspect, float znear,
                       float zfar, rc_camera* out)
{
    float view[16], proj[16];
    rc_camera_view_matrix(position, yaw, pitch, view);
    rc_projection_matrix(fovy, aspect, znear, zfar, proj);
    mat_mul(proj, view, out->view_proj);  // src/camera.rs:20
    out->eye[0] = position[0]; out->eye[1] = position[1]; out->eye[2] = position[2]; out->eye[3] = 1.0f;
}

void rc_uniform_camera_look_at(const float position[3], const float target[3], float fovy, float aspect,
                               float znear, float zfar, rc_camera* out)
{
    float view[16], proj[16];
    V3 eye{position[0], position[1], position[2]};
    V3 dir = normalize(sub(V3{target[0], target[1], target[2]}, eye));
    look_to_rh(eye, dir, V3{0.f, 1.f, 0.f}, view);
    rc_projection_matrix(fovy, aspect, znear, zfar, proj);
    mat_mul(proj, view, out->view_proj);
    out->eye[0] = position[0]; out->eye[1] = position[1]; out->eye[2] = position[2]; out->eye[3] = 1.0f;
}

}  // extern "C"
