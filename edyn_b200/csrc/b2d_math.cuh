// Device FP32 vector/quaternion/matrix helpers for the Edyn hot path.
// Expression order follows the reference's math headers (cited per function, paths relative to
// /root/reference) because several downstream decisions are threshold tests on these values
// (support_feature_tolerance, contact_breaking_threshold, ...).  The library is built with
// -fmad=false so that products and sums round exactly as the CPU path does.
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>

namespace b2d {

#define B2D_HD __host__ __device__ __forceinline__
#define B2D_D __device__ __forceinline__

constexpr float EPS = FLT_EPSILON;          // EDYN_EPSILON   include/edyn/math/scalar.hpp:16
constexpr float SCALAR_MAX = FLT_MAX;       // EDYN_SCALAR_MAX                         :18
constexpr float LARGE = 1e18f;              // large_scalar   include/edyn/math/constants.hpp:16
constexpr float HALF_SQRT2 = 0.7071067811865475244f;

struct v3 { float x, y, z; };
struct v2 { float x, y; };
struct q4 { float x, y, z, w; };
struct m3 { v3 r0, r1, r2; };               // row-major, include/edyn/math/matrix3x3.hpp:13

B2D_HD v3 mk3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
B2D_HD v3 mk3(float4 f) { return mk3(f.x, f.y, f.z); }
B2D_HD q4 mkq(float4 f) { q4 q; q.x = f.x; q.y = f.y; q.z = f.z; q.w = f.w; return q; }
B2D_HD float4 f4(v3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
B2D_HD float4 f4(q4 q) { return make_float4(q.x, q.y, q.z, q.w); }
B2D_HD float get(v3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
B2D_HD void set(v3 &v, int i, float s) { if (i == 0) v.x = s; else if (i == 1) v.y = s; else v.z = s; }

// include/edyn/math/vector3.hpp:58-140
B2D_HD v3 operator+(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
B2D_HD v3 operator-(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
B2D_HD v3 operator-(v3 a) { return mk3(-a.x, -a.y, -a.z); }
B2D_HD v3 operator*(v3 a, v3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
B2D_HD v3 operator*(v3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
B2D_HD v3 operator*(float s, v3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
B2D_HD v3 operator/(v3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }              // :104 true division
B2D_HD v3 &operator+=(v3 &a, v3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
B2D_HD v3 &operator-=(v3 &a, v3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
B2D_HD v3 &operator*=(v3 &a, float s) { a.x *= s; a.y *= s; a.z *= s; return a; }
B2D_HD v3 &operator/=(v3 &a, float s) { float z = 1.0f / s; a.x *= z; a.y *= z; a.z *= z; return a; }   // :119 reciprocal
B2D_HD float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
B2D_HD v3 cross(v3 v, v3 w) { return mk3(v.y * w.z - v.z * w.y, v.z * w.x - v.x * w.z, v.x * w.y - v.y * w.x); }
B2D_HD float length_sqr(v3 a) { return dot(a, a); }
B2D_HD float length(v3 a) { return sqrtf(length_sqr(a)); }
B2D_HD float distance_sqr(v3 a, v3 b) { return length_sqr(a - b); }
B2D_HD v3 vmin(v3 a, v3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
B2D_HD v3 vmax(v3 a, v3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
B2D_HD bool try_normalize(v3 &v) {                      // vector3.hpp:239-248 (compares against a double literal)
    float lsqr = length_sqr(v);
    if ((double)lsqr > 1e-18) { v /= sqrtf(lsqr); return true; }
    return false;
}
B2D_HD v3 project_plane(v3 p, v3 q, v3 n) { return p - n * dot(p - q, n); }   // vector3.hpp:255
B2D_HD int max_index_abs(v3 v) {                        // vector3.hpp:292-312
    float ax = fabsf(v.x), ay = fabsf(v.y), az = fabsf(v.z);
    float mv = ax; int mi = 0;
    if (ay > mv) { mv = ay; mi = 1; }
    if (az > mv) { mi = 2; }
    return mi;
}
B2D_HD float clamp_unit(float s) { return s < 0.0f ? 0.0f : (1.0f < s ? 1.0f : s); }       // math.hpp:46 (std::clamp)
B2D_HD v3 lerp(v3 a, v3 b, float s) { return a * (1.0f - s) + b * s; }                       // math.hpp:69
B2D_HD float square(float a) { return a * a; }
B2D_HD v2 operator-(v2 a, v2 b) { v2 r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
B2D_HD v2 neg(v2 a) { v2 r; r.x = -a.x; r.y = -a.y; return r; }

// include/edyn/math/quaternion.hpp:64-71, :117-149, :257
B2D_HD q4 operator*(q4 q, q4 r) {
    q4 o;
    o.x = q.w * r.x + q.x * r.w + q.y * r.z - q.z * r.y;
    o.y = q.w * r.y + q.y * r.w + q.z * r.x - q.x * r.z;
    o.z = q.w * r.z + q.z * r.w + q.x * r.y - q.y * r.x;
    o.w = q.w * r.w - q.x * r.x - q.y * r.y - q.z * r.z;
    return o;
}
B2D_HD q4 operator*(q4 q, float s) { q4 o; o.x = q.x * s; o.y = q.y * s; o.z = q.z * s; o.w = q.w * s; return o; }
B2D_HD q4 operator/(q4 q, float s) { q4 o; o.x = q.x / s; o.y = q.y / s; o.z = q.z / s; o.w = q.w / s; return o; }
B2D_HD q4 operator+(q4 a, q4 b) { q4 o; o.x = a.x + b.x; o.y = a.y + b.y; o.z = a.z + b.z; o.w = a.w + b.w; return o; }
B2D_HD float length_sqr(q4 q) { return q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w; }
B2D_HD q4 normalize(q4 q) { return q / sqrtf(length_sqr(q)); }
B2D_HD q4 conjugate(q4 q) { q4 o; o.x = -q.x; o.y = -q.y; o.z = -q.z; o.w = q.w; return o; }
B2D_HD v3 rotate(q4 q, v3 v) {
    v3 r = mk3(q.x, q.y, q.z);
    return v + cross(2.0f * r, cross(r, v) + q.w * v);
}
B2D_HD v3 quat_x(q4 q) { return rotate(q, mk3(1, 0, 0)); }
B2D_HD v3 quat_y(q4 q) { return rotate(q, mk3(0, 1, 0)); }
B2D_HD v3 quat_z(q4 q) { return rotate(q, mk3(0, 0, 1)); }
B2D_HD q4 quat_derivative(q4 q, v3 w) { q4 a; a.x = w.x; a.y = w.y; a.z = w.z; a.w = 0; return (a * q) * 0.5f; }

// src/edyn/math/quaternion.cpp:7-22 -- exponential map with a Taylor branch below 1e-3 rad/s.
B2D_D q4 integrate(q4 q, v3 w, float dt) {
    const float ws = length(w);
    float t;
    if (ws < 0.001f) {
        const float k = 1.0f / 48.0f;
        t = 0.5f * dt - dt * dt * dt * k * ws * ws;
    } else {
        t = sinf(0.5f * ws * dt) / ws;
    }
    q4 r; r.x = w.x * t; r.y = w.y * t; r.z = w.z * t; r.w = cosf(0.5f * ws * dt);
    return normalize(r * q);
}

// include/edyn/math/matrix3x3.hpp
B2D_HD v3 col(const m3 &m, int i) { return mk3(get(m.r0, i), get(m.r1, i), get(m.r2, i)); }
B2D_HD float column_dot(const m3 &m, int i, v3 v) { return get(m.r0, i) * v.x + get(m.r1, i) * v.y + get(m.r2, i) * v.z; }
B2D_HD v3 operator*(const m3 &m, v3 v) { return mk3(dot(m.r0, v), dot(m.r1, v), dot(m.r2, v)); }            // :62
B2D_HD v3 mulT(v3 v, const m3 &m) { return mk3(column_dot(m, 0, v), column_dot(m, 1, v), column_dot(m, 2, v)); }   // :66 (v * m)
B2D_HD m3 operator*(const m3 &m, const m3 &n) {                                                              // :54-60
    m3 o;
    o.r0 = mk3(column_dot(n, 0, m.r0), column_dot(n, 1, m.r0), column_dot(n, 2, m.r0));
    o.r1 = mk3(column_dot(n, 0, m.r1), column_dot(n, 1, m.r1), column_dot(n, 2, m.r1));
    o.r2 = mk3(column_dot(n, 0, m.r2), column_dot(n, 1, m.r2), column_dot(n, 2, m.r2));
    return o;
}
B2D_HD m3 transpose(const m3 &m) { m3 o; o.r0 = col(m, 0); o.r1 = col(m, 1); o.r2 = col(m, 2); return o; }
B2D_HD m3 m3_columns(v3 a, v3 b, v3 c) { m3 o; o.r0 = mk3(a.x, b.x, c.x); o.r1 = mk3(a.y, b.y, c.y); o.r2 = mk3(a.z, b.z, c.z); return o; }
B2D_HD m3 m3_zero() { m3 o; o.r0 = o.r1 = o.r2 = mk3(0, 0, 0); return o; }
B2D_HD m3 to_m3(q4 q) {                                 // :252-265
    float d = length_sqr(q);
    float s = 2.0f / d;
    float xs = q.x * s, ys = q.y * s, zs = q.z * s;
    float wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    float xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
    float yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    m3 o;
    o.r0 = mk3(1.0f - (yy + zz), xy - wz, xz + wy);
    o.r1 = mk3(xy + wz, 1.0f - (xx + zz), yz - wx);
    o.r2 = mk3(xz - wy, yz + wx, 1.0f - (xx + yy));
    return o;
}
// inertia_world_inv = R inv_I R^T   (sys/update_inertias.cpp:12-16, util/rigidbody.cpp:75-77)
B2D_HD m3 world_inertia(q4 orn, const m3 &inv_I) { m3 b = to_m3(orn); return b * inv_I * transpose(b); }

// include/edyn/math/transform.hpp:10-35
B2D_HD v3 to_world(v3 p, v3 pos, q4 orn) { return pos + rotate(orn, p); }
B2D_HD v3 to_object(v3 p, v3 pos, q4 orn) { return rotate(conjugate(orn), p - pos); }
B2D_HD v3 to_world(v3 p, v3 pos, const m3 &b) { return pos + b * p; }
B2D_HD v3 to_object(v3 p, v3 pos, const m3 &b) { return mulT(p - pos, b); }

struct box3 { v3 mn, mx; };                              // comp/aabb.hpp:12-43
B2D_HD box3 inset(const box3 &b, float v) { box3 o; o.mn = b.mn + mk3(v, v, v); o.mx = b.mx - mk3(v, v, v); return o; }
B2D_HD bool intersect(const box3 &a, const box3 &b) {   // src/edyn/math/geom.cpp:762-770 (closed intervals)
    return (a.mn.x <= b.mx.x) && (a.mx.x >= b.mn.x) && (a.mn.y <= b.mx.y) && (a.mx.y >= b.mn.y) &&
           (a.mn.z <= b.mx.z) && (a.mx.z >= b.mn.z);
}

} // namespace b2d
