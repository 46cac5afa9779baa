// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// CPU restatement (oracle) of pbrt-rust's PathIntegrator hot path.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// build, load or call anything under oracle/.  The product (pbrt-rust_b200/) never
// includes or links it.
//
// oracle_math.hpp: scalar helpers, vectors, bounds, transforms.
// All arithmetic is IEEE f32 without contraction (compile with
// -ffp-contract=off -fno-fast-math) except where the
// reference itself widens to f64 (cross products, triangle edge fallback).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>

namespace orc {

typedef float Float;

// src/core/pbrt.rs:23-34
static const Float PI = 3.14159265358979323846f;
static const Float PI_OVER2 = 1.57079632679489661923f;
static const Float PI_OVER4 = 0.78539816339744830961f;
static const Float INV_PI = 0.31830988618379067154f;
static const Float INV2_PI = 0.15915494309189533577f;
static const Float INV4_PI = 0.07957747154594766788f;
static const Float INFINITY_F = std::numeric_limits<float>::infinity();
static const Float SHADOW_EPSILON = 0.0001f;
static const Float MACHINE_EPSILON = std::numeric_limits<float>::epsilon() * 0.5f;
// src/core/rng.rs:4  hexf32!("0x1.fffffep-1")
static const Float ONE_MINUS_EPSILON = 0x1.fffffep-1f;

// src/core/pbrt.rs:60-76
inline uint32_t float_to_bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float bits_to_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// src/core/pbrt.rs:80-95
inline float next_float_up(float v) {
    if (std::isinf(v) && v > 0.0f) return v;
    float i = v;
    if (i == -0.0f) i = 0.0f;
    uint32_t ui = float_to_bits(i);
    if (i >= 0.0f) ui += 1; else ui -= 1;
    return bits_to_float(ui);
}
// src/core/pbrt.rs:97-112
inline float next_float_down(float v) {
    if (std::isinf(v) && v < 0.0f) return v;
    float i = v;
    if (i == 0.0f) i = -0.0f;
    uint32_t ui = float_to_bits(i);
    if (i > 0.0f) ui -= 1; else ui += 1;
    return bits_to_float(ui);
}

// src/core/pbrt.rs:206-208
inline Float gamma(int n) { return ((Float)n * MACHINE_EPSILON) / (1.0f - (Float)n * MACHINE_EPSILON); }

// src/core/pbrt.rs:172-182 (compare based; NaN passes through)
template <typename T> inline T clamp(T val, T low, T high) {
    if (val < low) return low; else if (val > high) return high; else return val;
}
// src/core/pbrt.rs:136-145
inline Float lerp(Float t, Float x, Float y) { return x * (1.0f - t) + y * t; }
// src/core/pbrt.rs:167-169
inline Float radians(Float deg) { return (PI / 180.0f) * deg; }

// src/core/pbrt.rs:184-204
template <typename P> inline int find_interval(int size, P pred) {
    int first = 0, len = size;
    while (len > 0) {
        int half = len >> 1, middle = first + half;
        if (pred(middle)) { first = middle + 1; len -= half + 1; } else { len = half; }
    }
    return clamp(first - 1, 0, size - 2);
}

// src/core/pbrt.rs:147-165
inline bool quadratic(Float a, Float b, Float c, Float* t0, Float* t1) {
    double discrim = (double)b * (double)b - 4.0 * (double)a * (double)c;
    if (discrim < 0.0) return false;
    double root = std::sqrt(discrim);
    double q = (b < 0.0f) ? -0.5 * ((double)b - root) : -0.5 * ((double)b + root);
    *t0 = (Float)(q / (double)a);
    *t1 = (Float)((double)c / q);
    if (*t0 > *t1) std::swap(*t0, *t1);
    return true;
}

// Rust `as usize` / `as isize` / `as i32`: saturating, NaN -> 0 (Appendix A.5).
inline int64_t f2i_sat(float f) {
    if (f != f) return 0;
    if (f >= 9.2233720368547758e18f) return INT64_MAX;
    if (f <= -9.2233720368547758e18f) return INT64_MIN;
    return (int64_t)f;
}
inline uint64_t f2u_sat(float f) {
    if (f != f || f <= 0.0f) return 0;
    if (f >= 1.8446744073709552e19f) return UINT64_MAX;
    return (uint64_t)f;
}

// ---- Vector3f / Point3f / Normal3f (src/core/geometry/{vector,point,normal}.rs).
// One struct serves all three; the reference's distinct ops are spelled out below.
struct V3 {
    Float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(Float a, Float b, Float c) : x(a), y(b), z(c) {}
    Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    Float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, Float s) { return V3(a.x * s, a.y * s, a.z * s); }
// vector.rs:486-495, point.rs:564-569, normal.rs:255-264: reciprocal then multiply
inline V3 operator/(V3 a, Float s) { Float d = 1.0f / s; return V3(a.x * d, a.y * d, a.z * d); }
inline V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
// vector.rs:259-266
inline Float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Float abs_dot(V3 a, V3 b) { return std::fabs(dot(a, b)); }
inline Float length_squared(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline Float length(V3 a) { return std::sqrt(length_squared(a)); }
// vector.rs:335-337: *self / self.length()
inline V3 normalize(V3 a) { return a / length(a); }
inline V3 vabs(V3 a) { return V3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
// vector.rs:339-353 (f64 products, narrowed)
inline V3 cross(V3 a, V3 b) {
    double v1x = a.x, v1y = a.y, v1z = a.z, v2x = b.x, v2y = b.y, v2z = b.z;
    return V3((Float)((v1y * v2z) - (v1z * v2y)), (Float)((v1z * v2x) - (v1x * v2z)),
              (Float)((v1x * v2y) - (v1y * v2x)));
}
// vector.rs:290-300: f32::max is maxNum
inline Float max_component(V3 a) { return std::fmax(a.x, std::fmax(a.y, a.z)); }
// vector.rs:302-316
inline int max_dimension(V3 a) { return (a.x > a.y) ? ((a.x > a.z) ? 0 : 2) : ((a.y > a.z) ? 1 : 2); }
inline V3 permute(V3 a, int x, int y, int z) { return V3(a[x], a[y], a[z]); }
// vector.rs:318-332 / normal.rs:95-117
inline V3 face_forward(V3 n, V3 v) { return (dot(n, v) < 0.0f) ? -n : n; }
inline Float distance_squared(V3 a, V3 b) { return length_squared(a - b); }
inline Float distance(V3 a, V3 b) { return length(a - b); }
// vector.rs:589-600
inline void coordinate_system(V3 v1, V3* v2, V3* v3) {
    if (std::fabs(v1.x) > std::fabs(v1.y))
        *v2 = V3(-v1.z, 0.0f, v1.x) / std::sqrt(v1.x * v1.x + v1.z * v1.z);
    else
        *v2 = V3(0.0f, v1.z, -v1.y) / std::sqrt(v1.y * v1.y + v1.z * v1.z);
    *v3 = cross(v1, *v2);
}

struct P2 { Float x, y; P2() : x(0), y(0) {} P2(Float a, Float b) : x(a), y(b) {} };

// ---- Bounds3f (src/core/geometry/bounds.rs)
struct Bounds3 {
    V3 p_min, p_max;
    // bounds.rs:459-471: inverted box
    Bounds3() : p_min(std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()),
                p_max(std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest()) {}
    Bounds3(V3 a, V3 b) : p_min(a), p_max(b) {}
    const V3& operator[](int i) const { return i == 0 ? p_min : p_max; }
};
inline V3 vmin(V3 a, V3 b) { return V3(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)); }
inline V3 vmax(V3 a, V3 b) { return V3(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)); }
// bounds.rs:536-541
inline Bounds3 bounds_from_points(V3 a, V3 b) { return Bounds3(vmin(a, b), vmax(a, b)); }
// bounds.rs:393-410
inline Bounds3 union_point(const Bounds3& b, V3 p) { return Bounds3(vmin(b.p_min, p), vmax(b.p_max, p)); }
inline Bounds3 union_bounds(const Bounds3& a, const Bounds3& b) { return Bounds3(vmin(a.p_min, b.p_min), vmax(a.p_max, b.p_max)); }
inline V3 diagonal(const Bounds3& b) { return b.p_max - b.p_min; }
// bounds.rs:507-513
inline Float surface_area(const Bounds3& b) { V3 d = diagonal(b); return (d.x * d.y + d.x * d.z + d.y * d.z) * 2.0f; }
// bounds.rs:343-356
inline int maximum_extent(const Bounds3& b) {
    V3 d = diagonal(b);
    if (d.x > d.y && d.x > d.z) return 0; else if (d.y > d.z) return 1; else return 2;
}
// bounds.rs:371-391
inline V3 bounds_offset(const Bounds3& b, V3 p) {
    V3 o = p - b.p_min;
    if (b.p_max.x > b.p_min.x) o.x /= b.p_max.x - b.p_min.x;
    if (b.p_max.y > b.p_min.y) o.y /= b.p_max.y - b.p_min.y;
    if (b.p_max.z > b.p_min.z) o.z /= b.p_max.z - b.p_min.z;
    return o;
}
inline bool bounds_inside(const Bounds3& b, V3 p) {
    return p.x >= b.p_min.x && p.x <= b.p_max.x && p.y >= b.p_min.y && p.y <= b.p_max.y && p.z >= b.p_min.z && p.z <= b.p_max.z;
}
// bounds.rs:515-523
inline void bounding_sphere(const Bounds3& b, V3* c, Float* rad) {
    *c = (b.p_min + b.p_max) / 2.0f;
    *rad = bounds_inside(b, *c) ? distance(b.p_max, *c) : 0.0f;
}

// ---- Ray (src/core/geometry/ray.rs:9-16)
struct Ray {
    V3 o, d;
    Float t_max, time;
    Ray() : t_max(INFINITY_F), time(0) {}
    Ray(V3 o_, V3 d_, Float tm, Float ti) : o(o_), d(d_), t_max(tm), time(ti) {}
    // RayDifferential (ray.rs:18-42): `diff` is Some; has_differentials is true whenever it is
    bool has_diff = false;
    V3 rxo, rxd, ryo, ryd;
};

// bounds.rs:559-580
inline bool bounds_intersect_p2(const Bounds3& b, const Ray& ray, V3 inv_dir, const int dir_isneg[3]) {
    Float tmin = (b[dir_isneg[0]].x - ray.o.x) * inv_dir.x;
    Float tmax = (b[1 - dir_isneg[0]].x - ray.o.x) * inv_dir.x;
    Float tymin = (b[dir_isneg[1]].y - ray.o.y) * inv_dir.y;
    Float tymax = (b[1 - dir_isneg[1]].y - ray.o.y) * inv_dir.y;
    tmax *= 1.0f + 2.0f * gamma(3);
    tymax *= 1.0f + 2.0f * gamma(3);
    if (tmin > tymax || tymin > tmax) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    Float tzmin = (b[dir_isneg[2]].z - ray.o.z) * inv_dir.z;
    Float tzmax = (b[1 - dir_isneg[2]].z - ray.o.z) * inv_dir.z;
    tzmax *= 1.0f + 2.0f * gamma(3);
    if (tmin > tzmax || tzmin > tmax) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    return (tmin < ray.t_max) && (tmax > 0.0f);
}

// src/core/geometry/geometry.rs:6-24
inline V3 offset_ray_origin(V3 p, V3 p_error, V3 n, V3 w) {
    Float d = dot(vabs(n), p_error);
    V3 offset = n * d;
    if (dot(w, n) < 0.0f) offset = -offset;
    V3 po = p + offset;
    for (int i = 0; i < 3; ++i) {
        if (offset[i] > 0.0f) po[i] = next_float_up(po[i]);
        else if (offset[i] < 0.0f) po[i] = next_float_down(po[i]);
    }
    return po;
}

// ---- Matrix4x4 / Transform (src/core/transform.rs)
struct M4 {
    Float m[4][4];
    M4() { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m[i][j] = (i == j) ? 1.0f : 0.0f; }
};
// transform.rs:147-159
inline M4 m4_mul(const M4& a, const M4& b) {
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
    return r;
}
// transform.rs:78-145 (Gauss-Jordan with full pivoting)
inline M4 m4_inverse(const M4& m) {
    int indxc[4] = {0, 0, 0, 0}, indxr[4] = {0, 0, 0, 0}, ipiv[4] = {0, 0, 0, 0};
    M4 minv = m;
    for (int i = 0; i < 4; ++i) {
        int irow = 0, icol = 0;
        Float big = 0.0f;
        for (int j = 0; j < 4; ++j) {
            if (ipiv[j] != 1) {
                for (int k = 0; k < 4; ++k) {
                    if (ipiv[k] == 0) {
                        Float a = std::fabs(minv.m[j][k]);
                        if (a >= big) { big = a; irow = j; icol = k; }
                    }
                }
            }
        }
        ipiv[icol] += 1;
        if (irow != icol) for (int k = 0; k < 4; ++k) std::swap(minv.m[irow][k], minv.m[icol][k]);
        indxr[i] = irow; indxc[i] = icol;
        Float pivinv = 1.0f / minv.m[icol][icol];
        minv.m[icol][icol] = 1.0f;
        for (int j = 0; j < 4; ++j) minv.m[icol][j] *= pivinv;
        for (int j = 0; j < 4; ++j) {
            if (j != icol) {
                Float save = minv.m[j][icol];
                minv.m[j][icol] = 0.0f;
                for (int k = 0; k < 4; ++k) minv.m[j][k] -= minv.m[icol][k] * save;
            }
        }
    }
    for (int i = 0; i < 4; ++i) {
        int j = 3 - i;
        if (indxr[j] != indxc[j]) for (int k = 0; k < 4; ++k) std::swap(minv.m[k][indxr[j]], minv.m[k][indxc[j]]);
    }
    return minv;
}

struct Transform {
    M4 m, m_inv;
    Transform() {}
    Transform(const M4& a, const M4& b) : m(a), m_inv(b) {}
};
// transform.rs:647-656
inline Transform operator*(const Transform& a, const Transform& b) { return Transform(m4_mul(a.m, b.m), m4_mul(b.m_inv, a.m_inv)); }
inline Transform t_inverse(const Transform& t) { return Transform(t.m_inv, t.m); }
inline Transform t_from_matrix(const M4& m) { return Transform(m, m4_inverse(m)); }
// transform.rs:255-271
inline Transform t_translate(V3 d) {
    M4 m, mi;
    m.m[0][3] = d.x; m.m[1][3] = d.y; m.m[2][3] = d.z;
    mi.m[0][3] = -d.x; mi.m[1][3] = -d.y; mi.m[2][3] = -d.z;
    return Transform(m, mi);
}
// transform.rs:273-289
inline Transform t_scale(Float x, Float y, Float z) {
    M4 m, mi;
    m.m[0][0] = x; m.m[1][1] = y; m.m[2][2] = z;
    mi.m[0][0] = 1.0f / x; mi.m[1][1] = 1.0f / y; mi.m[2][2] = 1.0f / z;
    return Transform(m, mi);
}
// transform.rs:399-411
inline Transform t_perspective(Float fov, Float n, Float f) {
    M4 persp;
    persp.m[2][2] = f / (f - n); persp.m[2][3] = -f * n / (f - n);
    persp.m[3][2] = 1.0f; persp.m[3][3] = 0.0f;
    Float inv_tan_ang = 1.0f / std::tan(radians(fov) / 2.0f);
    return t_scale(inv_tan_ang, inv_tan_ang, 1.0f) * t_from_matrix(persp);
}
// transform.rs:357-393
inline Transform t_look_at(V3 pos, V3 look, V3 up) {
    M4 c2w;
    c2w.m[0][3] = pos.x; c2w.m[1][3] = pos.y; c2w.m[2][3] = pos.z; c2w.m[3][3] = 1.0f;
    V3 dir = normalize(look - pos);
    if (length(cross(normalize(up), dir)) == 0.0f) return Transform();
    V3 right = normalize(cross(normalize(up), dir));
    V3 new_up = cross(dir, right);
    c2w.m[0][0] = right.x; c2w.m[1][0] = right.y; c2w.m[2][0] = right.z; c2w.m[3][0] = 0.0f;
    c2w.m[0][1] = new_up.x; c2w.m[1][1] = new_up.y; c2w.m[2][1] = new_up.z; c2w.m[3][1] = 0.0f;
    c2w.m[0][2] = dir.x; c2w.m[1][2] = dir.y; c2w.m[2][2] = dir.z; c2w.m[3][2] = 0.0f;
    return Transform(m4_inverse(c2w), c2w);
}
// transform.rs:413-431
inline V3 m4_point(const M4& M, V3 p) {
    Float x = p.x, y = p.y, z = p.z;
    Float xp = x * M.m[0][0] + y * M.m[0][1] + z * M.m[0][2] + M.m[0][3];
    Float yp = x * M.m[1][0] + y * M.m[1][1] + z * M.m[1][2] + M.m[1][3];
    Float zp = x * M.m[2][0] + y * M.m[2][1] + z * M.m[2][2] + M.m[2][3];
    Float wp = x * M.m[3][0] + y * M.m[3][1] + z * M.m[3][2] + M.m[3][3];
    if (wp == 1.0f) return V3(xp, yp, zp);
    return V3(xp, yp, zp) / wp;
}
// transform.rs:433-456
inline V3 m4_point_error(const M4& M, V3 p, V3* p_error) {
    Float x = p.x, y = p.y, z = p.z;
    Float xs = std::fabs(x * M.m[0][0]) + std::fabs(y * M.m[0][1]) + std::fabs(z * M.m[0][2]) + std::fabs(M.m[0][3]);
    Float ys = std::fabs(x * M.m[1][0]) + std::fabs(y * M.m[1][1]) + std::fabs(z * M.m[1][2]) + std::fabs(M.m[1][3]);
    Float zs = std::fabs(x * M.m[2][0]) + std::fabs(y * M.m[2][1]) + std::fabs(z * M.m[2][2]) + std::fabs(M.m[2][3]);
    *p_error = V3(xs, ys, zs) * gamma(3);
    return m4_point(M, p);
}
// transform.rs:458-494
inline V3 m4_point_abs_error(const M4& M, V3 p, V3 pe, V3* abs_error) {
    Float x = p.x, y = p.y, z = p.z;
    abs_error->x = (gamma(3) + 1.0f) * (std::fabs(M.m[0][0]) * pe.x + std::fabs(M.m[0][1]) * pe.y + std::fabs(M.m[0][2]) * pe.z) +
                   gamma(3) * (std::fabs(M.m[0][0] * x) + std::fabs(M.m[0][1] * y) + std::fabs(M.m[0][2] * z) + std::fabs(M.m[0][3]));
    abs_error->y = (gamma(3) + 1.0f) * (std::fabs(M.m[1][0]) * pe.x + std::fabs(M.m[1][1]) * pe.y + std::fabs(M.m[1][2]) * pe.z) +
                   gamma(3) * (std::fabs(M.m[1][0] * x) + std::fabs(M.m[1][1] * y) + std::fabs(M.m[1][2] * z) + std::fabs(M.m[1][3]));
    abs_error->z = (gamma(3) + 1.0f) * (std::fabs(M.m[2][0]) * pe.x + std::fabs(M.m[2][1]) * pe.y + std::fabs(M.m[2][2]) * pe.z) +
                   gamma(3) * (std::fabs(M.m[2][0] * x) + std::fabs(M.m[2][1] * y) + std::fabs(M.m[2][2] * z) + std::fabs(M.m[2][3]));
    return m4_point(M, p);
}
// transform.rs:496-508
inline V3 m4_vector(const M4& M, V3 v) {
    Float x = v.x, y = v.y, z = v.z;
    return V3(x * M.m[0][0] + y * M.m[0][1] + z * M.m[0][2], x * M.m[1][0] + y * M.m[1][1] + z * M.m[1][2],
              x * M.m[2][0] + y * M.m[2][1] + z * M.m[2][2]);
}
// transform.rs:510-527
inline V3 m4_vector_error(const M4& M, V3 v, V3* abs_error) {
    Float x = v.x, y = v.y, z = v.z, g = gamma(3);
    abs_error->x = g * (std::fabs(x * M.m[0][0]) + std::fabs(y * M.m[0][1]) + std::fabs(z * M.m[0][2]));
    abs_error->y = g * (std::fabs(x * M.m[1][0]) + std::fabs(y * M.m[1][1]) + std::fabs(z * M.m[1][2]));
    abs_error->z = g * (std::fabs(x * M.m[2][0]) + std::fabs(y * M.m[2][1]) + std::fabs(z * M.m[2][2]));
    return m4_vector(M, v);
}
// transform.rs:529-541: uses the inverse's transpose; pass m_inv
inline V3 m4_normal(const M4& Minv, V3 n) {
    Float x = n.x, y = n.y, z = n.z;
    return V3(x * Minv.m[0][0] + y * Minv.m[1][0] + z * Minv.m[2][0], x * Minv.m[0][1] + y * Minv.m[1][1] + z * Minv.m[2][1],
              x * Minv.m[0][2] + y * Minv.m[1][2] + z * Minv.m[2][2]);
}
// transform.rs:543-577 (differentials dropped: constant textures only)
inline Ray m4_ray(const M4& M, const Ray& r) {
    V3 o_error;
    V3 o = m4_point_error(M, r.o, &o_error);
    V3 d = m4_vector(M, r.d);
    Float l2 = length_squared(d);
    Float t_max = r.t_max;
    if (l2 > 0.0f) {
        Float dt = dot(vabs(d), o_error) / l2;
        o += d * dt;
        t_max -= dt;
    }
    return Ray(o, d, t_max, r.time);
}
// transform.rs:579-591
inline Ray m4_ray_error(const M4& M, const Ray& r, V3* o_error, V3* d_error) {
    V3 o = m4_point_error(M, r.o, o_error);
    V3 d = m4_vector_error(M, r.d, d_error);
    Float l2 = length_squared(d);
    if (l2 > 0.0f) {
        Float dt = dot(vabs(d), *o_error) / l2;
        o += d * dt;
    }
    return Ray(o, d, r.t_max, r.time);
}
// transform.rs:638-644
inline bool m4_swaps_handedness(const M4& M) {
    Float det = M.m[0][0] * (M.m[1][1] * M.m[2][2] - M.m[1][2] * M.m[2][1]) - M.m[0][1] * (M.m[1][0] * M.m[2][2] - M.m[1][2] * M.m[2][0]) +
                M.m[0][2] * (M.m[1][0] * M.m[2][1] - M.m[1][1] * M.m[2][0]);
    return det < 0.0f;
}

}  // namespace orc
