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

#ifdef ORC_HAVE_RENDER
extern "C" {

// SamplerIntegrator::render.  stats_out[12] = camera_rays, intersection_tests, shadow_tests, zero_radiance,
// direct_den, closest{nodes,prims,rays}, any{nodes,prims,rays}, reserved
int orc_render(const pbrt_b200_scene_desc* sdesc, const pbrt_b200_render_desc* rd, float* rgbw, int nthreads, uint64_t* stats_out) {
    RenderJob job;
    setup_job(job, *sdesc, *rd);
    RenderCounters rc;
    render(job, rgbw, nthreads <= 0 ? orc_hardware_threads() : nthreads, &rc);
    if (stats_out) {
        uint64_t v[12] = {rc.camera_rays, rc.intersection_tests, rc.shadow_tests, rc.zero_radiance, rc.direct_den,
                          rc.trav_closest.nodes_tested, rc.trav_closest.tris_tested, rc.trav_closest.rays,
                          rc.trav_any.nodes_tested, rc.trav_any.tris_tested, rc.trav_any.rays, 0};
        std::memcpy(stats_out, v, sizeof v);
    }
    return 0;
}

// SpatialLightDistribution::lookup (lightdistrib.rs:231-340) at n points: voxel[3*n] (integer voxel coordinates) and
// func[n * n_lights] (the per-voxel light_contrib the Distribution1D is built from).  nvoxels_out[3] = grid resolution.
int orc_spatial_lookup(const pbrt_b200_scene_desc* sdesc, const float* points, uint64_t n, int32_t* voxel, float* func, int32_t* nvoxels_out) {
    RenderScene scene;
    scene.init_render(*sdesc);
    SpatialLightDistribution sp(&scene, 64);
    const size_t nl = sdesc->n_lights;
    for (int k = 0; k < 3; ++k) nvoxels_out[k] = (int32_t)sp.nvoxels[k];
    for (uint64_t i = 0; i < n; ++i) {
        V3 p(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
        int64_t pi[3];
        sp.voxel_of(p, pi);
        for (int k = 0; k < 3; ++k) voxel[3 * i + k] = (int32_t)pi[k];
        const Distribution1D& d = sp.lookup(p);
        for (size_t j = 0; j < nl; ++j) func[i * nl + j] = d.func[j];
    }
    return 0;
}

// Film::write_image arithmetic (film.rs:217-264) on an {r,g,b,w} buffer
void orc_film_resolve(const float* rgbw, uint64_t npix, float scale, float* rgb) {
    for (uint64_t i = 0; i < npix; ++i) {
        const float* p = rgbw + 4 * i;
        float xyz[3] = {0.412453f * p[0] + 0.357580f * p[1] + 0.180423f * p[2], 0.212671f * p[0] + 0.715160f * p[1] + 0.072169f * p[2],
                        0.019334f * p[0] + 0.119193f * p[1] + 0.950227f * p[2]};
        float c[3] = {3.240479f * xyz[0] - 1.537150f * xyz[1] - 0.498535f * xyz[2], -0.969256f * xyz[0] + 1.875991f * xyz[1] + 0.041556f * xyz[2],
                      0.055648f * xyz[0] - 0.204043f * xyz[1] + 1.057311f * xyz[2]};
        if (p[3] != 0.0f) { float inv = 1.0f / p[3]; for (int k = 0; k < 3; ++k) c[k] = std::fmax(c[k] * inv, 0.0f); }
        for (int k = 0; k < 3; ++k) rgb[3 * i + k] = c[k] * scale;
    }
}

// Sampler stream for one pixel: for sample s in [0, nsamples): camera sample (5 values) followed by
// `n1d2d` repetitions of get_1d + get_2d (3 values).  out has nsamples * (5 + 3*n1d2d) floats.
// The tile's sampler is cloned with `seed` and start_pixel()ed on every pixel in `prefix_pixels`
// (x fastest) before (px,py), as the tile loop would (integrator.rs:302-322).
int orc_sampler_stream(const pbrt_b200_sampler* sd, int64_t seed, const int* prefix_pixels, int n_prefix, int px, int py, int nsamples, int n1d2d, float* out) {
    SamplerTables t; t.sobol32 = sd->sobol_matrices32; t.vdc = sd->vdc_matrices; t.vdc_inv = sd->vdc_matrices_inv;
    std::unique_ptr<Sampler> base = make_sampler(*sd, t);
    std::unique_ptr<Sampler> s = base->clone(seed);
    for (int i = 0; i < n_prefix; ++i) s->start_pixel(prefix_pixels[2 * i], prefix_pixels[2 * i + 1]);
    s->start_pixel(px, py);
    size_t k = 0;
    for (int i = 0; i < nsamples; ++i) {
        CameraSample cs = s->get_camera_sample(px, py);
        out[k++] = cs.pfilm.x; out[k++] = cs.pfilm.y; out[k++] = cs.time; out[k++] = cs.plens.x; out[k++] = cs.plens.y;
        for (int j = 0; j < n1d2d; ++j) { out[k++] = s->get_1d(); P2 p = s->get_2d(); out[k++] = p.x; out[k++] = p.y; }
        if (!s->start_next_sample()) break;
    }
    return (int)k;
}

float orc_sobol_sample_float(const uint32_t* sobol32, uint64_t a, int dim, uint32_t scramble) { SamplerTables t; t.sobol32 = sobol32; return sobol_sample_float(t, a, dim, scramble); }
uint64_t orc_sobol_interval_to_index(const uint64_t* vdc, const uint64_t* vdc_inv, uint32_t m, uint64_t frame, int px, int py) {
    SamplerTables t; t.vdc = vdc; t.vdc_inv = vdc_inv; return sobol_interval_to_index(t, m, frame, px, py);
}
float orc_radical_inverse(int base_index, uint64_t n) { return radical_inverse(base_index, n); }
float orc_scrambled_radical_inverse(int base_index, uint64_t n) { const HaltonTables& T = HaltonTables::get(); return scrambled_radical_inverse(base_index, n, &T.perms[T.prime_sums[base_index]]); }
uint64_t orc_inverse_radical_inverse(uint64_t base, uint64_t inverse, uint64_t ndigits) { return inverse_radical_inverse(base, inverse, ndigits); }
void orc_halton_tables(uint32_t* primes1000, uint32_t* sums1000, uint16_t* perms, uint64_t* nperms) {
    const HaltonTables& T = HaltonTables::get();
    if (primes1000) std::memcpy(primes1000, T.primes.data(), 4000);
    if (sums1000) std::memcpy(sums1000, T.prime_sums.data(), 4000);
    if (perms) std::memcpy(perms, T.perms.data(), T.perms.size() * 2);
    if (nperms) *nperms = T.perms.size();
}
void orc_zerotwo_matrices(uint32_t* vdc32, uint32_t* sobol32) { std::memcpy(vdc32, ZeroTwoMatrices::get().vdc, 128); std::memcpy(sobol32, ZeroTwoMatrices::get().sobol1, 128); }
uint32_t orc_rng_u32(uint64_t seq, int use_seq, int skip) { RNG r; if (use_seq) r.set_sequence(seq); uint32_t v = 0; for (int i = 0; i <= skip; ++i) v = r.uniform_int32(); return v; }

// Distribution1D (tests/sampling.rs:202-281): mode 0 discrete -> out {index, pdf}; mode 1 continuous -> out {x, pdf, offset}
void orc_distribution1d(const float* func, int n, float u, int mode, float* out3) {
    Distribution1D d(std::vector<Float>(func, func + n));
    if (mode == 0) { Float pdf; size_t i = d.sample_discrete(u, &pdf); out3[0] = (float)i; out3[1] = pdf; out3[2] = d.func_int; }
    else { Float pdf; size_t off; Float x = d.sample_continuous(u, &pdf, &off); out3[0] = x; out3[1] = pdf; out3[2] = (float)off; }
}

// BSDF of material `m` on a canonical frame (n = +z, dpdu = +x): f, pdf and sample_f.
// in: wo[3], wi[3], u[2].  out: f[3], pdf, sampled f[3], sampled wi[3], sampled pdf, sampled flags, ncomp(non-specular)
// `x` (may be null): the row's pbrt_b200_material_ext with constant parameters only -- how uber and substrate are reached
void orc_bsdf_eval_ext(const pbrt_b200_material* m, const pbrt_b200_material_ext* x, const float* wo, const float* wi, const float* u, int flags, float* out13) {
    SurfaceInteraction si;
    si.n = si.sh_n = V3(0, 0, 1); si.dpdu = si.sh_dpdu = V3(1, 0, 0); si.dpdv = si.sh_dpdv = V3(0, 1, 0);
    BSDF b;
    if (x) {
        SceneView sv;
        std::memset(&sv.d, 0, sizeof sv.d);
        sv.d.materials = m; sv.d.n_materials = 1; sv.d.material_ext = x;
        compute_scattering_functions_ext(sv, 0, si, &b, true);
    } else compute_scattering_functions(*m, si, &b);
    for (int i = 0; i < 13; ++i) out13[i] = 0.0f;
    if (!b.valid) { out13[12] = -1.0f; return; }
    V3 WO(wo[0], wo[1], wo[2]), WI(wi[0], wi[1], wi[2]);
    Spectrum f = b.f(WO, WI, flags);
    out13[0] = f.c[0]; out13[1] = f.c[1]; out13[2] = f.c[2]; out13[3] = b.pdf(WO, WI, flags);
    V3 swi; Float spdf = 0.0f; int st = 0;
    Spectrum sf = b.sample_f(WO, &swi, P2(u[0], u[1]), &spdf, flags, &st);
    out13[4] = sf.c[0]; out13[5] = sf.c[1]; out13[6] = sf.c[2]; out13[7] = swi.x; out13[8] = swi.y; out13[9] = swi.z; out13[10] = spdf; out13[11] = (float)st;
    out13[12] = (float)b.num_components(BSDF_ALL & ~BSDF_SPECULAR);
}

// n evaluations of the same material / wo: wi[3n], u[2n] -> out[13n] (orc_bsdf_eval's layout); for the Monte-Carlo property
// tests of the shading half (white furnace, pdf normalisation, reciprocity: tests/test_oracle_shading_properties.py)
void orc_bsdf_eval(const pbrt_b200_material* m, const float* wo, const float* wi, const float* u, int flags, float* out13) {
    orc_bsdf_eval_ext(m, nullptr, wo, wi, u, flags, out13);
}
void orc_bsdf_eval_batch(const pbrt_b200_material* m, const float* wo, const float* wi, const float* u, int flags, uint64_t n, float* out) {
    for (uint64_t i = 0; i < n; ++i) orc_bsdf_eval(m, wo, wi + 3 * i, u + 2 * i, flags, out + 13 * i);
}
void orc_bsdf_eval_batch_ext(const pbrt_b200_material* m, const pbrt_b200_material_ext* x, const float* wo, const float* wi, const float* u, int flags, uint64_t n,
                             float* out) {
    for (uint64_t i = 0; i < n; ++i) orc_bsdf_eval_ext(m, x, wo, wi + 3 * i, u + 2 * i, flags, out + 13 * i);
}
// PerspectiveCamera::generate_ray_differential + scale_differential(1 / sqrt(spp)) (spp = 0: unscaled) -> o d rxo rxd ryo ryd
void orc_generate_ray_differential(const pbrt_b200_camera* c, const float* cs5, uint32_t spp, float* out18) {
    CameraSample cs; cs.pfilm = P2(cs5[0], cs5[1]); cs.time = cs5[2]; cs.plens = P2(cs5[3], cs5[4]);
    Ray r = generate_ray(*c, cs, spp);
    const V3 v[6] = {r.o, r.d, r.rxo, r.rxd, r.ryo, r.ryd};
    for (int k = 0; k < 6; ++k) { out18[3 * k] = v[k].x; out18[3 * k + 1] = v[k].y; out18[3 * k + 2] = v[k].z; }
}
// Light::sample_li for n sample points u[2n] from the reference point (p, n): out[8n] = {Li.rgb, wi.xyz, pdf, Light::pdf_li(wi)}
int orc_light_sample_batch(const pbrt_b200_scene_desc* sdesc, int light, const float* ref_p, const float* ref_n, const float* u, uint64_t n, float* out) {
    RenderScene scene;
    scene.init_render(*sdesc);
    if (light < 0 || (uint64_t)light >= sdesc->n_lights) return 1;
    InteractionData ref;
    ref.p = V3(ref_p[0], ref_p[1], ref_p[2]); ref.n = V3(ref_n[0], ref_n[1], ref_n[2]); ref.p_error = V3(0, 0, 0);
    for (uint64_t i = 0; i < n; ++i) {
        LightSample ls = light_sample_li(scene, light, ref, P2(u[2 * i], u[2 * i + 1]));
        float* o = out + 8 * i;
        o[0] = ls.Li.c[0]; o[1] = ls.Li.c[1]; o[2] = ls.Li.c[2]; o[3] = ls.wi.x; o[4] = ls.wi.y; o[5] = ls.wi.z; o[6] = ls.pdf;
        o[7] = (ls.pdf > 0.0f) ? light_pdf_li(scene, light, ref, ls.wi) : 0.0f;
    }
    return 0;
}

// Texture::evaluate of the program textures[first, first + count) at n interaction points.  in[16 * i]: p(3) uv(2) dpdx(3) dpdy(3)
// dudx dvdx dudy dvdy, one pad -> out[3 * i] (float textures: the value in every channel)
void orc_texture_eval(const pbrt_b200_scene_desc* sdesc, uint32_t first, uint32_t count, uint64_t n, const float* in, float* out) {
    SceneView sv;
    sv.init(*sdesc);
    pbrt_b200_texref ref{first, count};
    for (uint64_t i = 0; i < n; ++i) {
        const float* q = in + 16 * i;
        SurfaceInteraction si;
        si.p = V3(q[0], q[1], q[2]); si.uv = P2(q[3], q[4]); si.dpdx = V3(q[5], q[6], q[7]); si.dpdy = V3(q[8], q[9], q[10]);
        si.dudx = q[11]; si.dvdx = q[12]; si.dudy = q[13]; si.dvdy = q[14];
        Spectrum v = tex_eval(sv, ref, si);
        out[3 * i] = v.c[0]; out[3 * i + 1] = v.c[1]; out[3 * i + 2] = v.c[2];
    }
}

// PerspectiveCamera::generate_ray for one camera sample -> o[3], d[3]
void orc_generate_ray(const pbrt_b200_camera* c, const float* cs5, float* out6) {
    CameraSample cs; cs.pfilm = P2(cs5[0], cs5[1]); cs.time = cs5[2]; cs.plens = P2(cs5[3], cs5[4]);
    Ray r = generate_ray(*c, cs);
    out6[0] = r.o.x; out6[1] = r.o.y; out6[2] = r.o.z; out6[3] = r.d.x; out6[4] = r.d.y; out6[5] = r.d.z;
}

}  // extern "C"
#endif

// ---- the reference's own tests (tests/*.rs) re-run on the oracle: oracle/ref_tests.hpp ----
#ifdef ORC_HAVE_RENDER
#include <string>
#include "ref_tests.hpp"
extern "C" {
// which: name of the reference #[test]; a,b: size parameters (0 = the reference's own counts); out2: optional info.
// Returns the number of violated assertions, or (uint64_t)-1 for an unknown name.
uint64_t orc_reftest(const char* which, uint64_t a, uint64_t b, int nthreads, const void* table, uint64_t* out2) {
    using namespace orc::reftest;
    std::string w(which);
    int nt = nthreads <= 0 ? orc_hardware_threads() : nthreads;
    uint64_t dummy[2] = {0, 0};
    if (!out2) out2 = dummy;
    if (w == "triangle_watertight") return triangle_watertight(a ? a : 100000, nt);
    if (w == "triangle_reintersect") return triangle_reintersect(a ? a : 1000, b ? b : 10000, nt, out2);
    if (w == "triangle_sampling") return triangle_sampling(a ? a : 512 * 1024, nt, out2);
    if (w == "triangle_solid_angle") return triangle_solid_angle(nt, out2);
    if (w == "triangle_badcases") return triangle_badcases();
    if (w == "sphere_solid_angle") return sphere_solid_angle();
    if (w == "full_sphere_reintersect") return full_sphere_reintersect(a ? a : 100, b ? b : 10000, nt, out2);
    if (w == "radical_inverse_test") return radical_inverse_test();
    if (w == "scrambled_radical_inverse_test") return scrambled_radical_inverse_test();
    if (w == "generator_matrix") return generator_matrix();
    if (w == "gray_code_sample_test") return gray_code_sample_test();
    if (w == "sobol") return sobol_test((const uint32_t*)table);
    if (w == "elementary_intervals") return elementary_intervals(a ? (int)a : 2);
    if (w == "distribution1d_discrete") return distribution1d_discrete();
    if (w == "distribution1d_continuous") return distribution1d_continuous();
    if (w == "next_float_up_down") return next_float_up_down();
    if (w == "float_bits") return float_bits();
    if (w == "efloat_add") return efloat_arith(0, a ? a : 1000000, nt);
    if (w == "efloat_sub") return efloat_arith(1, a ? a : 1000000, nt);
    if (w == "efloat_mul") return efloat_arith(2, a ? a : 1000000, nt);
    if (w == "efloat_div") return efloat_arith(3, a ? a : 1000000, nt);
    if (w == "bounds3_union") return bounds3_union();
    if (w == "bitops") return bitops();
    if (w == "find_interval_test") return find_interval_test();
    return (uint64_t)-1;
}
}
#endif
