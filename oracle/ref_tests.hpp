// TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see oracle_math.hpp header).
// ref_tests.hpp: the reference's OWN integration tests for the hot path (pbrt-rust `tests/*.rs`),
// re-run against the CPU oracle.  This is how the oracle is pinned: each
// function below follows one `#[test]` of the reference, cited by file:line, and returns the
// number of violated assertions (0 = the reference's test passes on the oracle).
// The loops stay in C++ because the reference's tests run 10^7..10^8 intersection tests.
#pragma once
#include <atomic>
#include <thread>
#include <vector>
#include "oracle_render.hpp"

namespace orc {
namespace reftest {

template <typename F> inline void par_seeds(uint64_t n, int nthreads, F f) {
    if (nthreads <= 1) { for (uint64_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<uint64_t> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&]() { for (;;) { uint64_t i = next.fetch_add(1); if (i >= n) break; f(i); } });
    for (auto& x : th) x.join();
}

// tests/shapes.rs:24-32
inline Float pexp(RNG& rng, Float e) { Float logu = lerp(rng.uniform_float(), -e, e); return std::pow(10.0f, logu); }
inline Float punif(RNG& rng, Float range) { return lerp(rng.uniform_float(), -range, range); }
// core/sampling.rs:212-218
inline V3 uniform_sample_sphere(P2 u) {
    Float z = 1.0f - 2.0f * u.x;
    Float r = std::sqrt(std::fmax(1.0f - z * z, 0.0f));
    Float phi = 2.0f * PI * u.y;
    return V3(r * std::cos(phi), r * std::sin(phi), z);
}
// core/geometry/geometry.rs:27-33
inline V3 spherical_direction(Float st, Float ct, Float phi) { return V3(st * std::cos(phi), st * std::sin(phi), ct); }

// A free-standing triangle mesh on the oracle's flat tables (create_trianglemesh with identity transforms).
struct Mesh {
    std::vector<float> p;
    std::vector<uint32_t> idx;
    std::vector<pbrt_b200_prim> prims;
    pbrt_b200_scene_desc d;
    RenderScene s;
    void finish() {
        size_t nt = idx.size() / 3;
        prims.resize(nt);
        for (size_t i = 0; i < nt; ++i) {
            std::memset(&prims[i], 0, sizeof prims[i]);
            prims[i].shape_kind = PBRT_B200_SHAPE_TRIANGLE; prims[i].shape_index = (uint32_t)i; prims[i].material = -1; prims[i].area_light = -1;
            prims[i].creation_index = (uint32_t)i;
        }
        std::memset(&d, 0, sizeof d);
        d.vertex_p = p.data(); d.n_vertices = p.size() / 3; d.tri_indices = idx.data(); d.n_triangles = nt; d.prims = prims.data(); d.n_prims = nt;
        s.init(d);
    }
    V3 P(size_t i) const { return V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
};

// Triangle::area, triangle.rs:550-554
inline Float tri_area(V3 p0, V3 p1, V3 p2) { return 0.5f * length(cross(p1 - p0, p2 - p0)); }

// tests/shapes.rs:148-171 get_random_trianlge
template <typename F> inline bool random_triangle(F value, Mesh& m) {
    V3 v[3];
    for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) v[j][k] = value();
    if (length_squared(cross(v[1] - v[0], v[2] - v[0])) < 1.0e-20f) return false;
    m.p.clear(); m.idx = {0, 1, 2};
    for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) m.p.push_back(v[j][k]);
    m.finish();
    return true;
}

inline pbrt_b200_light tri_light(const Mesh& m) {
    pbrt_b200_light l; std::memset(&l, 0, sizeof l);
    l.type = PBRT_B200_LIGHT_DIFFUSE; l.shape_kind = PBRT_B200_SHAPE_TRIANGLE; l.shape_index = 0; l.shape_flags = 0;
    l.area = tri_area(m.P(0), m.P(1), m.P(2));
    return l;
}

// Triangle::intersect / intersect_p on slot `slot` (closest = Triangle::intersect incl. :236-263)
inline bool tri_isect(const Mesh& m, uint32_t slot, const Ray& r, Hit* h) { Ray rr = r; return prim_intersect(m.s, slot, rr, h, nullptr); }
inline bool tri_isect_p(const Mesh& m, uint32_t slot, const Ray& r) { return prim_intersect_p(m.s, slot, r, nullptr); }

// ---- tests/shapes.rs:35-146 triangle_watertight (commented out upstream with `//#[test]`, still a valid property)
inline uint64_t triangle_watertight(uint64_t nseeds, int nthreads) {
    RNG rng(12111);
    const size_t ntheta = 16, nphi = 16, nvertices = ntheta * nphi;
    std::vector<V3> vertices;
    for (size_t t = 0; t < ntheta; ++t) {
        Float theta = PI * (Float)t / (Float)(ntheta - 1);
        Float ct = std::cos(theta), st = std::sin(theta);
        for (size_t p = 0; p < nphi; ++p) {
            Float phi = 2.0f * PI * (Float)p / (Float)(nphi - 1);
            Float radius = 1.0f;
            if (t == 0) vertices.push_back(V3(0, 0, radius));
            else if (t == ntheta - 1) vertices.push_back(V3(0, 0, -radius));
            else if (p == nphi - 1) vertices.push_back(vertices[vertices.size() - (nphi - 1)]);  // upstream writes `t == nphi - 1` (a typo: pbrt-v3 has p); rows close exactly either way only with p
            else { radius += 5.0f * rng.uniform_float(); vertices.push_back(V3(0, 0, 0) + spherical_direction(st, ct, phi) * radius); }
        }
    }
    Mesh m;
    for (V3 v : vertices) { m.p.push_back(v.x); m.p.push_back(v.y); m.p.push_back(v.z); }
    auto offset = [&](size_t t, size_t p) { return (uint32_t)(t * nphi + p); };
    for (size_t p = 0; p < nphi - 1; ++p) { m.idx.push_back(offset(0, 0)); m.idx.push_back(offset(1, p)); m.idx.push_back(offset(1, p + 1)); }
    for (size_t t = 1; t < ntheta - 2; ++t)
        for (size_t p = 0; p < nphi - 1; ++p) {
            m.idx.push_back(offset(t, p)); m.idx.push_back(offset(t + 1, p)); m.idx.push_back(offset(t + 1, p + 1));
            m.idx.push_back(offset(t, p)); m.idx.push_back(offset(t + 1, p + 1)); m.idx.push_back(offset(t, p + 1));
        }
    for (size_t p = 0; p < nphi - 1; ++p) { m.idx.push_back(offset(ntheta - 1, 0)); m.idx.push_back(offset(ntheta - 2, p)); m.idx.push_back(offset(ntheta - 2, p + 1)); }
    m.finish();
    (void)nvertices;
    const uint32_t ntris = (uint32_t)(m.idx.size() / 3);
    std::atomic<uint64_t> fails(0);
    par_seeds(nseeds, nthreads, [&](uint64_t i) {
        RNG r(i);
        P2 u; u.x = r.uniform_float(); u.y = r.uniform_float();
        V3 p = V3(0, 0, 0) + uniform_sample_sphere(u) * 0.5f;
        u.x = r.uniform_float(); u.y = r.uniform_float();
        Ray ray(p, uniform_sample_sphere(u), INFINITY_F, 0.0f);
        int nhits = 0;
        for (uint32_t k = 0; k < ntris; ++k) { Hit h; if (tri_isect(m, k, ray, &h)) ++nhits; }
        if (nhits < 1) fails++;
        V3 pv = vertices[r.uniform_int32_2((uint32_t)vertices.size())];
        ray.d = pv - ray.o;
        nhits = 0;
        for (uint32_t k = 0; k < ntris; ++k) { Hit h; if (tri_isect(m, k, ray, &h)) ++nhits; }
        if (nhits < 1) fails++;
    });
    return fails.load();
}

// ---- tests/shapes.rs:173-224 triangle_reintersect.  out2 = {triangles that were hit, spawned rays checked}
inline uint64_t triangle_reintersect(uint64_t ntri, uint64_t nrays, int nthreads, uint64_t* out2) {
    std::atomic<uint64_t> fails(0), hit_tris(0), checked(0);
    par_seeds(ntri, nthreads, [&](uint64_t i) {
        RNG rng(i);
        Mesh m;
        if (!random_triangle([&]() { return pexp(rng, 8.0f); }, m)) return;
        P2 u; u.x = rng.uniform_float(); u.y = rng.uniform_float();
        // Triangle::sample, triangle.rs:556-584 (position only)
        P2 b = uniform_sample_triangle(u);
        V3 ptri = m.P(0) * b.x + m.P(1) * b.y + m.P(2) * (1.0f - b.x - b.y);
        V3 o; for (int j = 0; j < 3; ++j) o[j] = pexp(rng, 8.0f);
        Ray r(o, ptri - o, INFINITY_F, 0.0f);
        Hit h;
        if (!tri_isect(m, 0, r, &h)) return;
        hit_tris++;
        SurfaceInteraction isect = make_interaction(m.s, r, h);
        uint64_t local_fail = 0;
        for (uint64_t j = 0; j < nrays; ++j) {
            u.x = rng.uniform_float(); u.y = rng.uniform_float();
            V3 w = uniform_sample_sphere(u);
            Ray rout = spawn_ray(isect.p, isect.p_error, isect.n, w, isect.time);
            if (tri_isect_p(m, 0, rout)) local_fail++;
            Hit h2;
            if (tri_isect(m, 0, rout, &h2)) local_fail++;
            V3 p2; for (int k = 0; k < 3; ++k) p2[k] = pexp(rng, 8.0f);
            // Interaction::spawn_rayto_point, interaction.rs:38-44
            V3 origin = offset_ray_origin(isect.p, isect.p_error, isect.n, p2 - isect.p);
            Ray rto(origin, p2 - isect.p, 1.0f - SHADOW_EPSILON, isect.time);
            if (tri_isect_p(m, 0, rto)) local_fail++;
            if (tri_isect(m, 0, rto, &h2)) local_fail++;
        }
        checked += 2 * nrays;
        fails += local_fail;
    });
    if (out2) { out2[0] = hit_tris.load(); out2[1] = checked.load(); }
    return fails.load();
}

// Triangle::solid_angle, triangle.rs:586-624 (closed form; not on the hot path, restated here only as the
// independent yardstick the reference's test uses for Shape::sample_interaction's pdf)
inline Float tri_solid_angle(const Mesh& m, V3 p) {
    V3 ps[3] = {normalize(m.P(0) - p), normalize(m.P(1) - p), normalize(m.P(2) - p)};
    V3 c01 = cross(ps[0], ps[1]), c12 = cross(ps[1], ps[2]), c20 = cross(ps[2], ps[0]);
    if (length_squared(c01) > 0.0f) c01 = normalize(c01);
    if (length_squared(c12) > 0.0f) c12 = normalize(c12);
    if (length_squared(c20) > 0.0f) c20 = normalize(c20);
    return std::fabs(std::acos(clamp(dot(c01, -c12), -1.0f, 1.0f)) + std::acos(clamp(dot(c12, -c20), -1.0f, 1.0f)) +
                     std::acos(clamp(dot(c20, -c01), -1.0f, 1.0f)) - PI);
}

inline Float sa_error(Float a, Float b) {  // tests/shapes.rs:276-282
    if (std::fabs(a) < 1.0e-4f || std::fabs(b) < 1.0e-4f) return std::fabs(a - b);
    return std::fabs((a - b) / b);
}

inline V3 far_reference_point(RNG& rng, Float range) {  // tests/shapes.rs:239-247
    V3 pc(punif(rng, range), punif(rng, range), punif(rng, range));
    uint32_t idx = rng.uniform_int32() % 3;
    pc[idx] = (rng.uniform_float() > 0.5f) ? (-range - 3.0f) : (range + 3.0f);
    return pc;
}

// ---- tests/shapes.rs:226-299 triangle_sampling.  out2 = {triangles compared, pdf<=0 count}
inline uint64_t triangle_sampling(uint64_t count, int nthreads, uint64_t* out2) {
    std::atomic<uint64_t> fails(0), compared(0), badpdf(0);
    par_seeds(30, nthreads, [&](uint64_t i) {
        const Float range = 10.0f;
        RNG rng(i);
        Mesh m;
        if (!random_triangle([&]() { return punif(rng, range); }, m)) return;
        V3 pc = far_reference_point(rng, range);
        uint64_t hits = 0;
        for (uint64_t j = 0; j < count; ++j) {
            P2 u(radical_inverse(0, j), radical_inverse(1, j));
            Ray ray(pc, uniform_sample_sphere(u), INFINITY_F, 0.0f);
            if (tri_isect_p(m, 0, ray)) ++hits;
        }
        double unif = (double)hits / ((double)count * (double)INV4_PI);
        pbrt_b200_light l = tri_light(m);
        InteractionData ref; ref.p = pc;
        double est = 0.0;
        for (uint64_t j = 0; j < count; ++j) {
            P2 u(radical_inverse(0, j), radical_inverse(1, j));
            Float pdf = 0.0f;
            triangle_sample_interaction(m.s, l, ref, u, &pdf);
            if (!(pdf > 0.0f)) { badpdf++; continue; }
            est += 1.0 / ((double)count * (double)pdf);
        }
        if (est > 1.0e-3) {
            compared++;
            if (!(sa_error((Float)est, (Float)unif) < 0.1f)) fails++;
        }
    });
    if (out2) { out2[0] = compared.load(); out2[1] = badpdf.load(); }
    return fails.load() + badpdf.load();
}

// ---- tests/shapes.rs:301-352 triangle_solid_angle
inline uint64_t triangle_solid_angle(int nthreads, uint64_t* out2) {
    std::atomic<uint64_t> fails(0), compared(0), badpdf(0);
    par_seeds(50, nthreads, [&](uint64_t i) {
        const Float range = 10.0f;
        RNG rng(100 + i);
        Mesh m;
        if (!random_triangle([&]() { return punif(rng, range); }, m)) return;
        V3 pc = far_reference_point(rng, range);
        const uint64_t count = 64 * 1024;
        pbrt_b200_light l = tri_light(m);
        InteractionData ref; ref.p = pc;
        double est = 0.0;
        for (uint64_t j = 0; j < count; ++j) {
            P2 u(radical_inverse(0, j), radical_inverse(1, j));
            Float pdf = 0.0f;
            triangle_sample_interaction(m.s, l, ref, u, &pdf);
            if (!(pdf > 0.0f)) { badpdf++; continue; }
            est += 1.0 / ((double)count * (double)pdf);
        }
        compared++;
        if (!(sa_error(tri_solid_angle(m, pc), (Float)est) < 0.015f)) fails++;
    });
    if (out2) { out2[0] = compared.load(); out2[1] = badpdf.load(); }
    return fails.load() + badpdf.load();
}

// ---- tests/shapes.rs:586-607 triangle_badcases: exact known answer, must be `false`
inline uint64_t triangle_badcases() {
    Mesh m;
    m.p = {-1113.45459f, -79.049614f, -56.2431908f, -1113.45459f, -87.0922699f, -56.2431908f, -1113.45459f, -79.2090149f, -56.2431908f};
    m.idx = {0, 1, 2};
    m.finish();
    Ray ray(V3(-1081.47925f, 99.9999542f, 87.7701111f), V3(-32.1072998f, -183.355865f, -144.607635f), 0.9999f, 0.0f);
    Hit h;
    return tri_isect(m, 0, ray, &h) ? 1 : 0;
}

inline pbrt_b200_sphere make_sphere(const Transform& o2w, Float radius) {
    pbrt_b200_sphere sp; std::memset(&sp, 0, sizeof sp);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { sp.object_to_world[4 * i + j] = o2w.m.m[i][j]; sp.world_to_object[4 * i + j] = o2w.m_inv.m[i][j]; }
    sp.radius = radius;
    return sp;
}

// Sphere::intersect_p = the root selection of sphere_test (full spheres)
inline bool sphere_isect_p(const pbrt_b200_sphere& sp, const Ray& r) { Float t; return sphere_test(sp, r, &t, nullptr); }

// ---- tests/shapes.rs:354-389 sphere_solid_angle.  The reference's second yardstick (Shape::solid_angle via
// Sphere::sample_interaction, sphere.rs:313-380) is not on the hot path (no sphere area lights); the closed form
// 2*pi*(1 - cos(theta_max)) that cone sampling integrates to replaces it.
inline uint64_t sphere_solid_angle() {
    // Transform::translate(1, .5, -.8) * Transform::rotate_x(30)
    Float th = radians(30.0f), s = std::sin(th), c = std::cos(th);
    M4 rx; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) rx.m[i][j] = (i == j) ? 1.0f : 0.0f;
    rx.m[1][1] = c; rx.m[1][2] = -s; rx.m[2][1] = s; rx.m[2][2] = c;
    M4 rxt; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) rxt.m[i][j] = rx.m[j][i];
    Transform tr = t_translate(V3(1.0f, 0.5f, -0.8f)) * Transform(rx, rxt);
    pbrt_b200_sphere sp = make_sphere(tr, 1.0f);
    const uint64_t n = 128 * 1024;
    auto mc = [&](V3 p) {
        uint64_t nh = 0;
        for (uint64_t i = 0; i < n; ++i) {
            P2 u(radical_inverse(0, i), radical_inverse(1, i));
            Ray ray(p, uniform_sample_sphere(u), INFINITY_F, 0.0f);
            if (sphere_isect_p(sp, ray)) ++nh;
        }
        return (Float)nh / (INV4_PI * (Float)n);
    };
    uint64_t fails = 0;
    if (!(std::fabs(mc(V3(1.0f, 0.9f, -0.8f)) - 4.0f * PI) < 0.01f)) fails++;
    V3 p(-1.25f, -1.0f, 0.8f);
    Float d2 = distance_squared(p, V3(1.0f, 0.5f, -0.8f));
    Float cos_max = std::sqrt(std::fmax(0.0f, 1.0f - 1.0f / d2));
    Float closed = 2.0f * PI * (1.0f - cos_max);
    if (!(std::fabs(mc(p) - closed) < 0.001f)) fails++;
    return fails;
}

// ---- tests/shapes.rs:412-487 test_reintersect_convex + full_sphere_reintersect.  out2 = {spheres hit, rays checked}
inline uint64_t full_sphere_reintersect(uint64_t nspheres, uint64_t nrays, int nthreads, uint64_t* out2) {
    std::atomic<uint64_t> fails(0), nhit(0), checked(0);
    par_seeds(nspheres, nthreads, [&](uint64_t i) {
        RNG rng(i);
        Float radius = pexp(rng, 4.0f);
        Transform iden;
        pbrt_b200_sphere sp = make_sphere(iden, radius);
        V3 o; for (int k = 0; k < 3; ++k) o[k] = pexp(rng, 8.0f);
        // Sphere::object_bound (sphere.rs:43-47) under the identity; Bounds3::lerp (bounds.rs:395-401)
        V3 t; for (int k = 0; k < 3; ++k) t[k] = rng.uniform_float();
        V3 p2(lerp(t.x, -radius, radius), lerp(t.y, -radius, radius), lerp(t.z, -radius, radius));
        Ray r(o, p2 - o, INFINITY_F, 0.0f);
        if (rng.uniform_float() < 0.5f) r.d = normalize(r.d);
        Float thit;
        if (!sphere_test(sp, r, &thit, nullptr)) return;
        nhit++;
        SurfaceInteraction isect = sphere_interaction(sp, r, thit);
        uint64_t lf = 0;
        for (uint64_t j = 0; j < nrays; ++j) {
            P2 u; u.x = rng.uniform_float(); u.y = rng.uniform_float();
            V3 w = uniform_sample_sphere(u);
            w = face_forward(w, isect.n);
            Ray rout = spawn_ray(isect.p, isect.p_error, isect.n, w, isect.time);
            if (sphere_isect_p(sp, rout)) lf++;
            V3 p3; for (int k = 0; k < 3; ++k) p3[k] = pexp(rng, 8.0f);
            w = p3 - isect.p;
            w = face_forward(w, isect.n);
            p3 = isect.p + w;
            V3 origin = offset_ray_origin(isect.p, isect.p_error, isect.n, p3 - isect.p);
            Ray rto(origin, p3 - isect.p, 1.0f - SHADOW_EPSILON, isect.time);
            if (sphere_isect_p(sp, rto)) lf++;
        }
        checked += 2 * nrays;
        fails += lf;
    });
    if (out2) { out2[0] = nhit.load(); out2[1] = checked.load(); }
    return fails.load();
}

// ---- tests/sampling.rs ----------------------------------------------------------------------
inline uint32_t multiply_generator(const uint32_t* C, uint32_t a) {  // lowdiscrepancy.rs:428-440
    uint32_t v = 0;
    for (int i = 0; a != 0; ++i, a >>= 1) if (a & 1) v ^= C[i];
    return v;
}
inline Float sample_generator_matrix(const uint32_t* C, uint32_t a, uint32_t scramble) {  // lowdiscrepancy.rs:465-467
    return std::fmin((Float)(multiply_generator(C, a) ^ scramble) * 0x1.0p-32f, ONE_MINUS_EPSILON);
}

// tests/sampling.rs:15-21
inline uint64_t radical_inverse_test() {
    uint64_t f = 0;
    for (uint32_t a = 0; a < 1024; ++a) if ((Float)reverse_bits32(a) * 2.3283064365386963e-10f != radical_inverse(0, a)) f++;
    return f;
}
// tests/sampling.rs:23-52 (upstream's expected-value loop is inert -- `relative_eq!` is never asserted and `n *= inv_base as u32`
// zeroes n -- so the expectation is evaluated digit by digit in f64, which is what pbrt-v3's original test intends)
inline uint64_t scrambled_radical_inverse_test() {
    uint64_t f = 0;
    const HaltonTables& T = HaltonTables::get();
    for (size_t dim = 0; dim < 128; ++dim) {
        RNG rng(dim);
        const uint32_t base = T.primes[dim];
        std::vector<uint16_t> perm(base);
        for (uint32_t i = 0; i < base; ++i) perm[i] = (uint16_t)(base - 1 - i);
        shuffle(perm.data(), perm.size(), 1, rng);
        const uint32_t idxs[7] = {0u, 1u, 2u, 1151u, 32351u, 4363211u, 681122u};
        for (uint32_t index : idxs) {
            double val = 0.0, inv_base = 1.0 / (double)base, inv_bi = inv_base;
            uint32_t n = index;
            while (n > 0) { uint32_t di = perm[n % base]; val += di * inv_bi; n /= base; inv_bi *= inv_base; }
            val += perm[0] * (double)base / ((double)base - 1.0) * inv_bi;
            double got = scrambled_radical_inverse(dim, index, perm.data());
            if (!(std::fabs(got - val) <= 1.0e-5 * std::fmax(std::fabs(val), std::fabs(got)) + 1e-7)) f++;
        }
    }
    return f;
}
// tests/sampling.rs:54-83
inline uint64_t generator_matrix() {
    uint64_t f = 0;
    uint32_t c[32], crev[32];
    for (int i = 0; i < 32; ++i) { c[i] = 1u << i; crev[i] = reverse_bits32(c[i]); }
    for (uint32_t a = 0; a < 128; ++a) {
        if (a != multiply_generator(c, a)) f++;
        if (radical_inverse(0, a) != (Float)reverse_bits32(multiply_generator(c, a)) * 2.3283064365386963e-10f) f++;
        if (radical_inverse(0, a) != sample_generator_matrix(crev, a, 0)) f++;
    }
    RNG rng;
    for (int i = 0; i < 32; ++i) { c[i] = rng.uniform_int32(); crev[i] = reverse_bits32(c[i]); }
    for (uint32_t a = 0; a < 1024; ++a) if (reverse_bits32(multiply_generator(c, a)) != multiply_generator(crev, a)) f++;
    return f;
}
// tests/sampling.rs:85-97
inline uint64_t gray_code_sample_test() {
    uint32_t c[32];
    for (int i = 0; i < 32; ++i) c[i] = 1u << i;
    std::vector<Float> v(64, 0.0f);
    gray_code_sample1d(c, (uint32_t)v.size(), 0, v.data());
    uint64_t f = 0;
    for (size_t a = 0; a < v.size(); ++a) {
        Float u = (Float)multiply_generator(c, (uint32_t)a) * 2.3283064365386963e-10f;
        bool found = false;
        for (Float x : v) if (x == u) found = true;
        if (!found) f++;
    }
    return f;
}
// tests/sampling.rs:99-106
inline uint64_t sobol_test(const uint32_t* sobol32) {
    SamplerTables T; T.sobol32 = sobol32;
    uint64_t f = 0;
    for (uint32_t i = 0; i < 8192; ++i) if (sobol_sample_float(T, i, 0, 0) != (Float)reverse_bits32(i) * 2.3283064365386963e-10f) f++;
    return f;
}
// tests/sampling.rs:108-157 check_sampler + elementary_intervals (only ZeroTwoSequenceSampler is enabled upstream, logsamples 2;
// `max_log` extends the same property to more sample counts)
inline uint64_t elementary_intervals(int max_log) {
    uint64_t f = 0;
    for (int logsamples = 2; logsamples <= max_log; ++logsamples) {
        ZeroTwoSequenceSampler s((uint64_t)1 << logsamples, 2);
        s.start_pixel(0, 0);
        std::vector<P2> samples;
        do { samples.push_back(s.get_2d()); } while (s.start_next_sample());
        for (int i = 0; i <= logsamples; ++i) {
            int64_t nx = (int64_t)1 << i, ny = (int64_t)1 << (logsamples - i);
            std::vector<int> count((size_t)1 << logsamples, 0);
            for (P2 p : samples) {
                Float x = (Float)nx * p.x, y = (Float)ny * p.y;
                if (!(x >= 0.0f && x < (Float)nx && y >= 0.0f && y < (Float)ny)) { f++; continue; }
                int64_t index = (int64_t)std::floor(y) * nx + (int64_t)std::floor(x);
                if (index < 0 || index >= (int64_t)count.size()) { f++; continue; }
                if (count[(size_t)index] != 0) f++;
                count[(size_t)index] += 1;
            }
        }
    }
    return f;
}
// Distribution1D::sample_discrete with u_remapped, sampling.rs:65-85
inline size_t sample_discrete_remap(const Distribution1D& d, Float u, Float* pdf, Float* uremapped) {
    size_t off = d.sample_discrete(u, pdf);
    if (uremapped) *uremapped = (u - d.cdf[off]) / (d.cdf[off + 1] - d.cdf[off]);
    return off;
}
// tests/sampling.rs:202-256
inline uint64_t distribution1d_discrete() {
    uint64_t f = 0;
    Distribution1D dist(std::vector<Float>{0.0f, 1.0f, 0.0f, 3.0f});
    if (dist.count() != 4) f++;
    if (dist.discrete_pdf(0) != 0.0f || dist.discrete_pdf(1) != 0.25f || dist.discrete_pdf(2) != 0.0f || dist.discrete_pdf(3) != 0.75f) f++;
    Float pdf = 0.0f, ur = 0.0f;
    if (dist.sample_discrete(0.0f, &pdf) != 1 || pdf != 0.25f) f++;
    if (sample_discrete_remap(dist, 0.125f, &pdf, &ur) != 1 || pdf != 0.25f || ur != 0.5f) f++;
    if (dist.sample_discrete(0.24999f, &pdf) != 1 || pdf != 0.25f) f++;
    if (dist.sample_discrete(0.250001f, &pdf) != 3 || pdf != 0.75f) f++;
    if (sample_discrete_remap(dist, 0.625f, &pdf, &ur) != 3 || pdf != 0.75f || ur != 0.5f) f++;
    if (dist.sample_discrete(ONE_MINUS_EPSILON, &pdf) != 3 || pdf != 0.75f) f++;
    if (dist.sample_discrete(1.0f, &pdf) != 3 || pdf != 0.75f) f++;
    Float u = 0.25f, umax = 0.25f;
    for (int i = 0; i < 20; ++i) { u = next_float_down(u); umax = next_float_up(umax); }
    while (u < umax) {
        size_t interval = dist.sample_discrete(u, nullptr);
        if (interval == 3) break;
        if (interval != 1) f++;
        u = next_float_up(u);
    }
    if (!(u < umax)) f++;
    while (u <= umax) {
        if (dist.sample_discrete(u, nullptr) != 3) f++;
        u = next_float_up(u);
    }
    return f;
}
// tests/sampling.rs:258-280 (the `relative_eq!` lines are asserted here, as pbrt-v3's EXPECT_FLOAT_EQ does)
inline uint64_t distribution1d_continuous() {
    uint64_t f = 0;
    Distribution1D dist(std::vector<Float>{1.0f, 1.0f, 2.0f, 4.0f, 8.0f});
    auto near = [](Float a, Float b) { return std::fabs(a - b) <= 1.0e-5f * std::fmax(std::fabs(a), std::fabs(b)) + 1e-12f; };
    if (dist.count() != 5) f++;
    Float pdf = 0.0f; size_t off = 0;
    if (dist.sample_continuous(0.0f, &pdf, &off) != 0.0f) f++;
    if (!near((Float)dist.count() * 1.0f / 16.0f, pdf)) f++;  // pbrt-v3: count * 1/16
    if (off != 0) f++;
    if (!near(0.8f, dist.sample_continuous(0.5f, &pdf, &off))) f++;
    if (!near(0.9f, dist.sample_continuous(0.75f, &pdf, &off))) f++;
    if (!near((Float)dist.count() * 8.0f / 16.0f, pdf)) f++;
    if (off != 4) f++;
    if (!near(0.0f, dist.sample_continuous(0.0f, &pdf, nullptr))) f++;
    if (!near(1.0f, dist.sample_continuous(1.0f, &pdf, nullptr))) f++;
    return f;
}

// ---- tests/fp.rs -----------------------------------------------------------------------------
inline Float get_float(RNG& rng) { Float f; do { f = bits_to_float(rng.uniform_int32()); } while (std::isnan(f)); return f; }
// tests/fp.rs:23-44
inline uint64_t next_float_up_down() {
    uint64_t f = 0;
    if (!(next_float_up(-0.0f) > 0.0f)) f++;
    if (!(next_float_down(0.0f) < 0.0f)) f++;
    if (next_float_up(INFINITY_F) != INFINITY_F) f++;
    if (!(next_float_down(INFINITY_F) < INFINITY_F)) f++;
    if (next_float_down(-INFINITY_F) != -INFINITY_F) f++;
    if (!(next_float_up(-INFINITY_F) > -INFINITY_F)) f++;
    RNG rng;
    for (int i = 0; i < 100000; ++i) {
        Float v = get_float(rng);
        if (std::isinf(v)) continue;
        if (std::nextafterf(v, INFINITY_F) != next_float_up(v)) f++;
        if (std::nextafterf(v, -INFINITY_F) != next_float_down(v)) f++;
    }
    return f;
}
// tests/fp.rs:46-57
inline uint64_t float_bits() {
    uint64_t f = 0;
    RNG rng(1);
    for (int i = 0; i < 100000; ++i) {
        uint32_t ui = rng.uniform_int32();
        Float v = bits_to_float(ui);
        if (std::isnan(v)) continue;
        if (ui != float_to_bits(v)) f++;
    }
    return f;
}
// tests/fp.rs:73-120
inline EFloat get_efloat(RNG& rng, Float min_exp = -6.0f, Float max_exp = 6.0f) {
    Float logu = lerp(rng.uniform_float(), min_exp, max_exp);
    Float val = std::pow(10.0f, logu);
    Float err = 0.0f;
    switch (rng.uniform_int32_2(4)) {
        case 1: { uint32_t ulp = rng.uniform_int32_2(1024); Float off = bits_to_float(float_to_bits(val) + ulp); err = std::fabs(off - val); break; }
        case 2: { uint32_t ulp = rng.uniform_int32_2(1024 * 1024); Float off = bits_to_float(float_to_bits(val) + ulp); err = std::fabs(off - val); break; }
        case 3: err = (4.0f * rng.uniform_float()) * std::fabs(val); break;
        default: break;
    }
    Float sign = rng.uniform_float() < 0.5f ? -1.0f : 1.0f;
    return EFloat(sign * val, err);
}
inline double get_precise(const EFloat& ef, RNG& rng) {
    switch (rng.uniform_int32_2(3)) {
        case 0: return (double)ef.low;
        case 1: return (double)ef.high;
        case 2: {
            Float t = rng.uniform_float();
            double p = (1.0 - (double)t) * (double)ef.low + (double)t * (double)ef.high;
            if (p > (double)ef.high) p = (double)ef.high;
            if (p < (double)ef.low) p = (double)ef.low;
            return p;
        }
        default: return (double)ef.v;
    }
}
// tests/fp.rs:160-226: op 0 add, 1 sub, 2 mul, 3 div (efloat_abs / efloat_sqrt: EFloat::abs/sqrt are not used by Sphere::intersect)
inline uint64_t efloat_arith(int op, uint64_t iters, int nthreads) {
    std::atomic<uint64_t> fails(0);
    const uint64_t chunk = 4096;
    par_seeds((iters + chunk - 1) / chunk, nthreads, [&](uint64_t c) {
        uint64_t lf = 0;
        for (uint64_t trial = c * chunk; trial < std::min(iters, (c + 1) * chunk); ++trial) {
            RNG rng(trial);
            EFloat ef[2]; ef[0] = get_efloat(rng); ef[1] = get_efloat(rng);
            double pr[2]; pr[0] = get_precise(ef[0], rng); pr[1] = get_precise(ef[1], rng);
            EFloat r; double p;
            if (op == 3) {
                Float abs_err = std::fmax(std::fabs(ef[1].high - ef[1].v), std::fabs(ef[1].v - ef[1].low));  // get_absolute_error, efloat.rs:74-79
                if (ef[1].low * ef[1].high < 0.0f || abs_err > 0.25f * std::fabs(ef[1].low)) continue;
            }
            switch (op) { case 0: r = ef[0] + ef[1]; p = pr[0] + pr[1]; break; case 1: r = ef[0] - ef[1]; p = pr[0] - pr[1]; break;
                          case 2: r = ef[0] * ef[1]; p = pr[0] * pr[1]; break; default: r = ef[0] / ef[1]; p = pr[0] / pr[1]; }
            if (!(p >= (double)r.low) || !(p <= (double)r.high)) lf++;
        }
        fails += lf;
    });
    return fails.load();
}

// ---- tests/bounds.rs:23-34 bounds3_union
inline uint64_t bounds3_union() {
    uint64_t f = 0;
    auto eq = [](const Bounds3& a, const Bounds3& b) { return a.p_min.x == b.p_min.x && a.p_min.y == b.p_min.y && a.p_min.z == b.p_min.z &&
                                                              a.p_max.x == b.p_max.x && a.p_max.y == b.p_max.y && a.p_max.z == b.p_max.z; };
    Bounds3 a = bounds_from_points(V3(-10, -10, 5), V3(0, 20, 10));
    Bounds3 b;
    if (!eq(a, union_bounds(a, b))) f++;
    if (!eq(b, union_bounds(b, b))) f++;
    Bounds3 d(V3(-15, 10, 30), V3(-15, 10, 30));
    if (!eq(bounds_from_points(V3(-15, -10, 5), V3(0, 20, 30)), union_bounds(a, d))) f++;
    return f;
}

// ---- tests/bitops.rs (log2_int / round_up_pow2 as the samplers use them)
inline uint64_t bitops() {
    uint64_t f = 0;
    for (int i = 0; i < 31; ++i) { int32_t v = (int32_t)(1u << i); if (log2_int(v) != i) f++; }
    for (int i = 1; i < 31; ++i) { int32_t v = (int32_t)(1u << i); if (log2_int(v + 1) != i) f++; }
    if (round_up_pow2_32(7) != 8) f++;
    for (int32_t i = 1; i < (1 << 24); ++i) {
        bool p2 = (i & (i - 1)) == 0;
        if (p2) { if (round_up_pow2_32(i) != i) f++; } else if (round_up_pow2_32(i) != (1 << (log2_int(i) + 1))) f++;
        if (p2) { if (round_up_pow2_64(i) != i) f++; } else if (round_up_pow2_64(i) != ((int64_t)1 << (log2_int(i) + 1))) f++;
    }
    for (int i = 0; i < 30; ++i) {
        int32_t v = 1 << i;
        if (round_up_pow2_32(v) != v) f++;
        if (v > 2 && round_up_pow2_32(v - 1) != v) f++;
        if (round_up_pow2_32(v + 1) != 2 * v) f++;
    }
    return f;
}

// ---- tests/find_interval.rs
inline uint64_t find_interval_test() {
    uint64_t f = 0;
    const Float a[10] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9};
    const int n = 10;
    if (find_interval(n, [&](int i) { return a[i] <= -1.0f; }) != 0) f++;
    if (find_interval(n, [&](int i) { return a[i] <= 100.0f; }) != n - 2) f++;
    for (int i = 0; i < n - 1; ++i) {
        if (find_interval(n, [&](int j) { return a[j] <= (Float)i; }) != i) f++;
        if (find_interval(n, [&](int j) { return a[j] <= (Float)i + 0.5f; }) != i) f++;
        if (i > 0 && find_interval(n, [&](int j) { return a[j] <= (Float)i - 0.5f; }) != i - 1) f++;
    }
    return f;
}

}  // namespace reftest
}  // namespace orc
