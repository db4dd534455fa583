// TEST INFRASTRUCTURE -- CPU oracle for the Edyn per-step hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use anything under oracle/.  The product (edyn_b200/, include/) never does.
//
// FP32 math restated from the reference headers (paths relative to /root/reference):
//   include/edyn/math/vector3.hpp:12-330, quaternion.hpp:10-260, matrix3x3.hpp:13-300,
//   transform.hpp:10-35, src/edyn/math/quaternion.cpp:7-22.
// Operation order is kept expression-by-expression so that, compiled without FMA
// contraction, results are bit-identical to the reference's own translation units
// (checked in tests/test_oracle_vs_ref.py against oracle/_ref).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <algorithm>

namespace ora {

using scalar = float;
constexpr scalar EPS = FLT_EPSILON;        // EDYN_EPSILON, math/scalar.hpp:16
constexpr scalar SCALAR_MAX = FLT_MAX;     // EDYN_SCALAR_MAX, math/scalar.hpp:18
constexpr scalar LARGE = scalar(1e18);     // large_scalar, math/constants.hpp:16
constexpr scalar HALF_SQRT2 = scalar(0.7071067811865475244008443621048490);
constexpr scalar PI = scalar(3.1415926535897932384626433832795029);

struct vec3 {
    scalar x, y, z;
    scalar &operator[](size_t i) { return (&x)[i]; }
    scalar operator[](size_t i) const { return (&x)[i]; }
};
struct vec2 { scalar x, y; };

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator*(vec3 a, scalar s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(scalar s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
// vector3.hpp:104 divides component-wise ...
inline vec3 operator/(vec3 a, scalar s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3 &operator+=(vec3 &a, vec3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
inline vec3 &operator-=(vec3 &a, vec3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
inline vec3 &operator*=(vec3 &a, scalar s) { a.x *= s; a.y *= s; a.z *= s; return a; }
// ... while vector3.hpp:119-125 (operator/=) multiplies by the reciprocal.
inline vec3 &operator/=(vec3 &a, scalar s) { scalar z = scalar(1) / s; a.x *= z; a.y *= z; a.z *= z; return a; }
inline bool operator==(vec3 a, vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline scalar dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 v, vec3 w) {
    return {v.y * w.z - v.z * w.y, v.z * w.x - v.x * w.z, v.x * w.y - v.y * w.x};
}
inline scalar length_sqr(vec3 a) { return dot(a, a); }
inline scalar length(vec3 a) { return std::sqrt(length_sqr(a)); }
inline scalar distance_sqr(vec3 a, vec3 b) { return length_sqr(a - b); }
inline vec3 vmin(vec3 a, vec3 b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
inline vec3 vmax(vec3 a, vec3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }
inline bool try_normalize(vec3 &v) {            // vector3.hpp:239-248 (double literal on purpose)
    scalar lsqr = length_sqr(v);
    if (lsqr > 1e-18) { v /= std::sqrt(lsqr); return true; }
    return false;
}
inline vec3 project_plane(vec3 p, vec3 q, vec3 n) { return p - n * dot(p - q, n); }   // vector3.hpp:255
inline size_t max_index_abs(vec3 v) {           // vector3.hpp:292-312
    vec3 a{std::abs(v.x), std::abs(v.y), std::abs(v.z)};
    scalar mv = a.x; size_t mi = 0;
    if (a.y > mv) { mv = a.y; mi = 1; }
    if (a.z > mv) { mi = 2; }
    return mi;
}
inline scalar clamp_unit(scalar s) { return std::clamp(s, scalar(0), scalar(1)); }    // math.hpp:46
inline vec3 lerp(vec3 a, vec3 b, scalar s) { return a * (scalar(1) - s) + b * s; }     // math.hpp:69
inline scalar square(scalar a) { return a * a; }

inline vec2 operator-(vec2 a, vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline vec2 operator-(vec2 a) { return {-a.x, -a.y}; }

struct quat { scalar x, y, z, w; };
inline quat operator*(quat q, quat r) {          // quaternion.hpp:64-71
    return {q.w * r.x + q.x * r.w + q.y * r.z - q.z * r.y,
            q.w * r.y + q.y * r.w + q.z * r.x - q.x * r.z,
            q.w * r.z + q.z * r.w + q.x * r.y - q.y * r.x,
            q.w * r.w - q.x * r.x - q.y * r.y - q.z * r.z};
}
inline quat operator*(quat q, scalar s) { return {q.x * s, q.y * s, q.z * s, q.w * s}; }
inline quat operator/(quat q, scalar s) { return {q.x / s, q.y / s, q.z / s, q.w / s}; }
inline quat operator+(quat a, quat b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline scalar length_sqr(quat q) { return q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w; }
inline quat normalize(quat q) { return q / std::sqrt(length_sqr(q)); }
inline quat conjugate(quat q) { return {-q.x, -q.y, -q.z, q.w}; }
inline vec3 rotate(quat q, vec3 v) {             // quaternion.hpp:145-148
    vec3 r{q.x, q.y, q.z};
    return v + cross(scalar(2) * r, cross(r, v) + q.w * v);
}
inline vec3 quat_x(quat q) { return rotate(q, vec3{1, 0, 0}); }
inline vec3 quat_y(quat q) { return rotate(q, vec3{0, 1, 0}); }
inline vec3 quat_z(quat q) { return rotate(q, vec3{0, 0, 1}); }
inline quat quat_derivative(quat q, vec3 w) {    // quaternion.hpp:257
    return (quat{w.x, w.y, w.z, 0} * q) * scalar(0.5);
}
// src/edyn/math/quaternion.cpp:7-22 (exponential map; Taylor branch below 1e-3 rad/s)
inline quat integrate(quat q, vec3 w, scalar dt) {
    const scalar ws = length(w);
    const scalar min_ws = scalar(0.001);
    const scalar half = scalar(0.5);
    scalar t;
    if (ws < min_ws) {
        const scalar k = scalar(1) / scalar(48);
        t = half * dt - dt * dt * dt * k * ws * ws;
    } else {
        t = std::sin(half * ws * dt) / ws;
    }
    quat r{w.x * t, w.y * t, w.z * t, std::cos(half * ws * dt)};
    return normalize(r * q);
}

struct mat3 {
    vec3 row[3];
    vec3 column(size_t i) const { return {row[0][i], row[1][i], row[2][i]}; }
    scalar column_dot(size_t i, vec3 v) const { return row[0][i] * v.x + row[1][i] * v.y + row[2][i] * v.z; }
};
inline mat3 mat3_zero() { return {{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}}; }
inline vec3 operator*(const mat3 &m, vec3 v) { return {dot(m.row[0], v), dot(m.row[1], v), dot(m.row[2], v)}; }
inline vec3 operator*(vec3 v, const mat3 &m) { return {m.column_dot(0, v), m.column_dot(1, v), m.column_dot(2, v)}; }
inline mat3 operator*(const mat3 &m, const mat3 &n) {      // matrix3x3.hpp:54-60
    return {{{n.column_dot(0, m.row[0]), n.column_dot(1, m.row[0]), n.column_dot(2, m.row[0])},
             {n.column_dot(0, m.row[1]), n.column_dot(1, m.row[1]), n.column_dot(2, m.row[1])},
             {n.column_dot(0, m.row[2]), n.column_dot(1, m.row[2]), n.column_dot(2, m.row[2])}}};
}
inline mat3 transpose(const mat3 &m) { return {{m.column(0), m.column(1), m.column(2)}}; }
inline mat3 mat3_columns(vec3 a, vec3 b, vec3 c) { return {{{a.x, b.x, c.x}, {a.y, b.y, c.y}, {a.z, b.z, c.z}}}; }
inline mat3 to_mat3(quat q) {                    // matrix3x3.hpp:252-265
    scalar d = length_sqr(q);
    scalar s = 2 / d;
    scalar xs = q.x * s, ys = q.y * s, zs = q.z * s;
    scalar wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    scalar xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
    scalar yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    return {{{1 - (yy + zz), xy - wz, xz + wy},
             {xy + wz, 1 - (xx + zz), yz - wx},
             {xz - wy, yz + wx, 1 - (xx + yy)}}};
}
inline mat3 inverse_symmetric(const mat3 &m) {   // matrix3x3.hpp:177-204
    scalar det = dot(m.row[0], cross(m.row[1], m.row[2]));
    scalar det_inv = scalar(1) / det;
    scalar a11 = m.row[0][0], a12 = m.row[0][1], a13 = m.row[0][2];
    scalar a22 = m.row[1][1], a23 = m.row[1][2];
    scalar a33 = m.row[2][2];
    mat3 r{};
    r.row[0][0] = det_inv * (a22 * a33 - a23 * a23);
    r.row[0][1] = det_inv * (a13 * a23 - a12 * a33);
    r.row[0][2] = det_inv * (a12 * a23 - a13 * a22);
    r.row[1][0] = r.row[0][1];
    r.row[1][1] = det_inv * (a11 * a33 - a13 * a13);
    r.row[1][2] = det_inv * (a12 * a13 - a11 * a23);
    r.row[2][0] = r.row[0][2];
    r.row[2][1] = r.row[1][2];
    r.row[2][2] = det_inv * (a11 * a22 - a12 * a12);
    return r;
}

inline vec3 to_world(vec3 p, vec3 pos, quat orn) { return pos + rotate(orn, p); }              // transform.hpp:31
inline vec3 to_object(vec3 p, vec3 pos, quat orn) { return rotate(conjugate(orn), p - pos); }   // transform.hpp:26
inline vec3 to_world(vec3 p, vec3 pos, const mat3 &b) { return pos + b * p; }                  // transform.hpp:21
inline vec3 to_object(vec3 p, vec3 pos, const mat3 &b) { return (p - pos) * b; }               // transform.hpp:13

struct aabb { vec3 min, max; };
inline aabb inset(const aabb &b, vec3 v) { return {b.min + v, b.max - v}; }                    // comp/aabb.hpp:16
inline bool intersect(const aabb &a, const aabb &b) {                                          // geom.cpp:762-770
    return (a.min.x <= b.max.x) && (a.max.x >= b.min.x) &&
           (a.min.y <= b.max.y) && (a.max.y >= b.min.y) &&
           (a.min.z <= b.max.z) && (a.max.z >= b.min.z);
}

} // namespace ora
