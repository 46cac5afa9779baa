// TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see oracle_math.hpp header).
// C entry points of the CPU oracle, loaded with ctypes by tests/, smoke() and
// bench.py's CPU-baseline legs only.
#include <thread>
#include <vector>
#include <atomic>
#include <cstdio>
#include "oracle_accel.hpp"
#if __has_include("oracle_render.hpp")
#include "oracle_render.hpp"
#define ORC_HAVE_RENDER 1
#endif

using namespace orc;

namespace {
template <typename F> void parallel_for(uint64_t n, int nthreads, F f) {
    if (nthreads <= 1 || n < 1024) { f(0, n, 0); return; }
    std::vector<std::thread> th;
    std::atomic<uint64_t> next(0);
    const uint64_t chunk = 4096;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t]() {
            for (;;) {
                uint64_t b = next.fetch_add(chunk);
                if (b >= n) break;
                f(b, std::min(n, b + chunk), t);
            }
        });
    for (auto& x : th) x.join();
}
}  // namespace

extern "C" {

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// BVHAccel::new (bvh.rs:145-198)
int orc_bvh_build(const float* prim_bounds, uint64_t n, int max_prims, int split_method, pbrt_b200_bvh_node* nodes_out, uint32_t* ordered_out,
                  uint64_t* n_nodes_out) {
    std::vector<pbrt_b200_bvh_node> nodes;
    std::vector<uint32_t> ordered;
    bvh_build(prim_bounds, n, (size_t)max_prims, split_method, nodes, ordered);
    std::memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(pbrt_b200_bvh_node));
    std::memcpy(ordered_out, ordered.data(), ordered.size() * sizeof(uint32_t));
    *n_nodes_out = nodes.size();
    return 0;
}

// Scene::intersect over a batch.  counters_out (optional): [nodes_tested, prims_tested, rays].
// per_ray_counts (optional): 2 x uint32 per ray (nodes, prims).
int orc_intersect(const pbrt_b200_scene_desc* desc, const pbrt_b200_ray* rays, uint64_t n, pbrt_b200_hit* hits, int nthreads,
                  uint64_t* counters_out, uint32_t* per_ray_counts) {
    SceneView s; s.init(*desc);
    int nt = nthreads <= 0 ? orc_hardware_threads() : nthreads;
    std::vector<Counters> cs(nt);
    parallel_for(n, nt, [&](uint64_t b, uint64_t e, int t) {
        for (uint64_t i = b; i < e; ++i) {
            Ray r(V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].t_max, rays[i].time);
            Hit h;
            Counters c;
            bool found = scene_intersect(s, r, &h, &c);
            cs[t].nodes_tested += c.nodes_tested; cs[t].tris_tested += c.tris_tested; cs[t].rays += 1;
            if (per_ray_counts) { per_ray_counts[2 * i] = (uint32_t)c.nodes_tested; per_ray_counts[2 * i + 1] = (uint32_t)c.tris_tested; }
            if (found) { hits[i].prim = s.d.prims[h.slot].creation_index; hits[i].t = h.t; hits[i].b0 = h.b0; hits[i].b1 = h.b1; }
            else { hits[i].prim = PBRT_B200_NO_HIT; hits[i].t = r.t_max; hits[i].b0 = hits[i].b1 = 0.0f; }
        }
    });
    if (counters_out) {
        counters_out[0] = counters_out[1] = counters_out[2] = 0;
        for (auto& c : cs) { counters_out[0] += c.nodes_tested; counters_out[1] += c.tris_tested; counters_out[2] += c.rays; }
    }
    return 0;
}

// Scene::intersect_p over a batch.
int orc_intersect_p(const pbrt_b200_scene_desc* desc, const pbrt_b200_ray* rays, uint64_t n, uint8_t* occluded, int nthreads, uint64_t* counters_out) {
    SceneView s; s.init(*desc);
    int nt = nthreads <= 0 ? orc_hardware_threads() : nthreads;
    std::vector<Counters> cs(nt);
    parallel_for(n, nt, [&](uint64_t b, uint64_t e, int t) {
        for (uint64_t i = b; i < e; ++i) {
            Ray r(V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].t_max, rays[i].time);
            occluded[i] = scene_intersect_p(s, r, &cs[t]) ? 1 : 0;
        }
    });
    if (counters_out) {
        counters_out[0] = counters_out[1] = counters_out[2] = 0;
        for (auto& c : cs) { counters_out[0] += c.nodes_tested; counters_out[1] += c.tris_tested; counters_out[2] += c.rays; }
    }
    return 0;
}

// Triangle::intersect on a free-standing triangle (tests/shapes.rs fixtures).
// out = {t, b0, b1, b2, p.xyz, p_error.xyz, n.xyz}; returns 1 on hit.
int orc_triangle_intersect(const float* p9, const float* ray_o, const float* ray_d, float t_max, int reverse_orientation, float* out13) {
    float verts[9]; std::memcpy(verts, p9, sizeof verts);
    uint32_t idx[3] = {0, 1, 2};
    pbrt_b200_prim pr; std::memset(&pr, 0, sizeof pr);
    pr.flags = reverse_orientation ? PBRT_B200_PRIM_REVERSE_ORIENTATION : 0;
    pbrt_b200_scene_desc d; std::memset(&d, 0, sizeof d);
    d.vertex_p = verts; d.n_vertices = 3; d.tri_indices = idx; d.n_triangles = 1; d.prims = &pr; d.n_prims = 1;
    SceneView s; s.init(d);
    Ray r(V3(ray_o[0], ray_o[1], ray_o[2]), V3(ray_d[0], ray_d[1], ray_d[2]), t_max, 0.0f);
    Hit h;
    if (!prim_intersect(s, 0, r, &h, nullptr)) return 0;
    Ray r0(V3(ray_o[0], ray_o[1], ray_o[2]), V3(ray_d[0], ray_d[1], ray_d[2]), t_max, 0.0f);
    SurfaceInteraction si = make_interaction(s, r0, h);
    float o[13] = {h.t, h.b0, h.b1, h.b2, si.p.x, si.p.y, si.p.z, si.p_error.x, si.p_error.y, si.p_error.z, si.n.x, si.n.y, si.n.z};
    std::memcpy(out13, o, sizeof o);
    return 1;
}

// Sphere::intersect for a sphere given by its transforms.  out as above.
int orc_sphere_intersect(const pbrt_b200_sphere* sp, const float* ray_o, const float* ray_d, float t_max, float* out13) {
    Ray r(V3(ray_o[0], ray_o[1], ray_o[2]), V3(ray_d[0], ray_d[1], ray_d[2]), t_max, 0.0f);
    Float t;
    if (!sphere_test(*sp, r, &t, nullptr)) return 0;
    SurfaceInteraction si = sphere_interaction(*sp, r, t);
    float o[13] = {t, 0, 0, 0, si.p.x, si.p.y, si.p.z, si.p_error.x, si.p_error.y, si.p_error.z, si.n.x, si.n.y, si.n.z};
    std::memcpy(out13, o, sizeof o);
    return 1;
}

// Interaction::spawn_ray (interaction.rs:32-36): origin only.
void orc_spawn_ray_origin(const float* p, const float* p_error, const float* n, const float* d, float* o_out) {
    V3 o = offset_ray_origin(V3(p[0], p[1], p[2]), V3(p_error[0], p_error[1], p_error[2]), V3(n[0], n[1], n[2]), V3(d[0], d[1], d[2]));
    o_out[0] = o.x; o_out[1] = o.y; o_out[2] = o.z;
}

float orc_next_float_up(float v) { return next_float_up(v); }
float orc_next_float_down(float v) { return next_float_down(v); }
float orc_gamma(int n) { return gamma(n); }
int orc_find_interval(const float* a, int size, float x) { return find_interval(size, [&](int i) { return a[i] <= x; }); }

// EFloat ops for tests/fp.rs: op 0 add,1 sub,2 mul,3 div.  in: (v,err) x2, out: v,low,high
void orc_efloat_op(int op, float av, float aerr, float bv, float berr, float* out3) {
    EFloat a(av, aerr), b(bv, berr), r;
    switch (op) { case 0: r = a + b; break; case 1: r = a - b; break; case 2: r = a * b; break; default: r = a / b; }
    out3[0] = r.v; out3[1] = r.low; out3[2] = r.high;
}

}  // extern "C"
