// TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see oracle_math.hpp header).
// oracle_render.hpp: lights, light distributions, estimate_direct, PathIntegrator::li,
// PerspectiveCamera, Film and SamplerIntegrator::render.
#pragma once
#include <thread>
#include <atomic>
#include <mutex>
#include "oracle_reflection.hpp"
#include "oracle_texture.hpp"
#include "oracle_sampling.hpp"

namespace orc {

// InteractionData (src/core/interaction.rs:62-88): p, p_error, n (wo/time unused here)
struct InteractionData {
    V3 p, p_error, n;
    Float time = 0;
};

// Interaction::spawn_ray, interaction.rs:32-36
inline Ray spawn_ray(V3 p, V3 p_error, V3 n, V3 d, Float time) { return Ray(offset_ray_origin(p, p_error, n, d), d, INFINITY_F, time); }
// Interaction::spawn_rayto_interaction, interaction.rs:46-52
inline Ray spawn_ray_to(const InteractionData& a, const InteractionData& b) {
    V3 o = offset_ray_origin(a.p, a.p_error, a.n, b.p - a.p);
    V3 t = offset_ray_origin(b.p, b.p_error, b.n, o - b.p);
    V3 d = t - o;
    return Ray(o, d, 1.0f - SHADOW_EPSILON, a.time);
}

// Distribution2D, src/core/sampling.rs:93-143
struct Distribution2D {
    std::vector<Distribution1D> pconditional_v;
    Distribution1D pmarginal;
    Distribution2D() {}
    Distribution2D(const std::vector<Float>& func, size_t nu, size_t nv) {
        for (size_t v = 0; v < nv; ++v) pconditional_v.emplace_back(std::vector<Float>(func.begin() + v * nu, func.begin() + (v + 1) * nu));
        std::vector<Float> mf;
        for (size_t v = 0; v < nv; ++v) mf.push_back(pconditional_v[v].func_int);
        pmarginal = Distribution1D(mf);
    }
    P2 sample_continuous(P2 u, Float* pdf) const {
        Float pdfs[2]; size_t v = 0;
        Float d1 = pmarginal.sample_continuous(u.y, &pdfs[1], &v);
        Float d0 = pconditional_v[v].sample_continuous(u.x, &pdfs[0], nullptr);
        *pdf = pdfs[0] * pdfs[1];
        return P2(d0, d1);
    }
    Float pdf(P2 p) const {
        size_t nu = pconditional_v[0].count(), nv = pmarginal.count();
        size_t iu = clamp<size_t>((size_t)f2u_sat(p.x * (Float)nu), 0, nu - 1);
        size_t iv = clamp<size_t>((size_t)f2u_sat(p.y * (Float)nv), 0, nv - 1);
        return pconditional_v[iv].func[iu] / pmarginal.func_int;
    }
};


struct RenderScene : SceneView {
    V3 world_center; Float world_radius = 0;   // Light::preprocess (distant.rs:53-60, infinite.rs:104-111)
    std::vector<Distribution2D> inf_distrib;   // per infinite light
    std::vector<int> infinite_lights;          // Scene.infinite_lights, scene.rs:41-50
    void init_render(const pbrt_b200_scene_desc& desc) {
        init(desc);
        if (d.n_nodes) bounding_sphere(wb, &world_center, &world_radius);
        inf_distrib.resize(d.n_lights);
        for (uint64_t i = 0; i < d.n_lights; ++i)
            if (d.lights[i].type == PBRT_B200_LIGHT_INFINITE) {
                infinite_lights.push_back((int)i);
                // InfiniteAreaLight::new with the 1x1 constant map (infinite.rs:35-100): 2x2 importance image
                Spectrum L = spec3(d.lights[i].L);
                std::vector<Float> img(4);
                for (int k = 0; k < 4; ++k) {
                    int v = k / 2;
                    Float sin_theta = std::sin(PI * ((Float)v + 0.5f) / 2.0f);
                    img[k] = L.y() * sin_theta;
                }
                inf_distrib[i] = Distribution2D(img, 2, 2);
            }
    }
};

// ---- Light::power, for PowerLightDistribution (integrator.rs:239-247)
inline Spectrum light_power(const RenderScene& s, const pbrt_b200_light& l) {
    Spectrum L = spec3(l.L);
    switch (l.type) {
        case PBRT_B200_LIGHT_POINT: return L * 4.0f * PI;                                     // point.rs:45-47
        case PBRT_B200_LIGHT_DISTANT: return L * PI * s.world_radius * s.world_radius;        // distant.rs:48-51
        case PBRT_B200_LIGHT_SPOT: return L * 2.0f * PI * (1.0f - 0.5f * (l.cos_falloff_start + l.cos_total_width));  // spot.rs:62-64
        case PBRT_B200_LIGHT_DIFFUSE: return L * l.area * PI;                                 // diffuse.rs:79-81
        case PBRT_B200_LIGHT_INFINITE: return L * s.world_radius * s.world_radius * PI;       // infinite.rs:97-102
    }
    return Spectrum(0.0f);
}

// Triangle::sample + Shape::sample_interaction (triangle.rs:556-584, shape.rs:40-58)
inline InteractionData triangle_sample_interaction(const RenderScene& s, const pbrt_b200_light& l, const InteractionData& ref, P2 u, Float* pdf) {
    P2 b = uniform_sample_triangle(u);
    const uint32_t* idx = s.d.tri_indices + 3 * (size_t)l.shape_index;
    V3 p0 = s.P(idx[0]), p1 = s.P(idx[1]), p2 = s.P(idx[2]);
    InteractionData it;
    it.p = p0 * b.x + p1 * b.y + p2 * (1.0f - b.x - b.y);
    it.n = normalize(cross(p1 - p0, p2 - p0));
    if ((l.shape_flags & PBRT_B200_PRIM_HAS_N) && s.d.vertex_n) {
        V3 ns = s.N(idx[0]) * b.x + s.N(idx[1]) * b.y + s.N(idx[2]) * (1.0f - b.x - b.y);
        it.n = face_forward(it.n, ns);
    } else if (((l.shape_flags & PBRT_B200_PRIM_REVERSE_ORIENTATION) != 0) ^ ((l.shape_flags & PBRT_B200_PRIM_SWAPS_HANDEDNESS) != 0)) {
        it.n = it.n * -1.0f;
    }
    V3 pabs = vabs(p0 * b.x) + vabs(p1 * b.y) + vabs(p2 * (1.0f - b.x - b.y));
    it.p_error = pabs * gamma(6);
    *pdf = 1.0f / l.area;
    V3 wi = it.p - ref.p;
    if (length_squared(wi) == 0.0f) *pdf = 0.0f;
    else {
        wi = normalize(wi);
        *pdf *= distance_squared(ref.p, it.p) / abs_dot(it.n, -wi);
        if (std::isinf(*pdf)) *pdf = 0.0f;
    }
    return it;
}

// Shape::pdf_wi for a triangle light, shape.rs:63-82: re-intersects the single shape with s = None
inline Float triangle_pdf_wi(const RenderScene& s, const pbrt_b200_light& l, const InteractionData& ref, V3 wi) {
    Ray ray = spawn_ray(ref.p, ref.p_error, ref.n, wi, ref.time);
    pbrt_b200_prim pr; std::memset(&pr, 0, sizeof pr);
    pr.shape_kind = PBRT_B200_SHAPE_TRIANGLE; pr.shape_index = l.shape_index; pr.flags = l.shape_flags;
    uint32_t vi[3]; V3 p[3]; P2 uv[3];
    triangle_fetch(s, pr, vi, p, uv);
    Float t, b0, b1, b2;
    if (!triangle_test(ray, p[0], p[1], p[2], uv, true, &t, &b0, &b1, &b2)) return 0.0f;
    SurfaceInteraction isect = triangle_interaction(s, pr, ray, b0, b1, b2, false);
    Float pdf = distance_squared(ref.p, isect.p) / (dot(isect.n, -wi) * l.area);  // signed dot (quirk a-Q2)
    if (std::isinf(pdf)) pdf = 0.0f;
    return pdf;
}

struct LightSample {
    Spectrum Li;
    V3 wi;
    Float pdf = 0;
    InteractionData p1;  // VisibilityTester.p1
};

// ---- Sphere as an area-light shape: Sphere::sample / sample_interaction / pdf_wi (sphere.rs:295-395, full spheres)
inline V3 uniform_sample_sphere_(P2 u) {  // sampling.rs:212-218
    Float z = 1.0f - 2.0f * u.x;
    Float r = std::sqrt(std::fmax(1.0f - z * z, 0.0f));
    Float phi = 2.0f * PI * u.y;
    return V3(r * std::cos(phi), r * std::sin(phi), z);
}
inline InteractionData sphere_sample(const pbrt_b200_sphere& sp, const pbrt_b200_light& l, P2 u, Float* pdf) {  // sphere.rs:295-311
    M4 o2w = m4_from(sp.object_to_world), w2o = m4_from(sp.world_to_object);
    V3 pobj = V3(0, 0, 0) + uniform_sample_sphere_(u) * sp.radius;
    InteractionData it;
    it.n = normalize(m4_normal(w2o, pobj));
    if (l.shape_flags & PBRT_B200_PRIM_REVERSE_ORIENTATION) it.n = it.n * -1.0f;
    pobj = pobj * (sp.radius / distance(pobj, V3(0, 0, 0)));
    V3 pobj_error = vabs(pobj) * gamma(5);
    it.p = m4_point_abs_error(o2w, pobj, pobj_error, &it.p_error);
    *pdf = 1.0f / l.area;
    return it;
}
inline InteractionData sphere_sample_interaction(const RenderScene& s, const pbrt_b200_light& l, const InteractionData& ref, P2 u, Float* pdf) {  // sphere.rs:313-380
    const pbrt_b200_sphere& sp = s.d.spheres[l.shape_index];
    V3 pcenter = m4_point(m4_from(sp.object_to_world), V3(0, 0, 0));
    V3 porigin = offset_ray_origin(ref.p, ref.p_error, ref.n, pcenter - ref.p);
    if (distance_squared(porigin, pcenter) <= sp.radius * sp.radius) {  // inside: uniform over the sphere, converted to solid angle
        InteractionData intr = sphere_sample(sp, l, u, pdf);
        V3 wi = intr.p - ref.p;
        if (length_squared(wi) == 0.0f) *pdf = 0.0f;
        else {
            wi = normalize(wi);
            *pdf *= distance_squared(ref.p, intr.p) / abs_dot(intr.n, -wi);
        }
        if (std::isinf(*pdf)) *pdf = 0.0f;
        return intr;
    }
    // uniform in the subtended cone
    Float dc = distance(ref.p, pcenter);
    Float invdc = 1.0f / dc;
    V3 wc = (pcenter - ref.p) * invdc, wcx, wcy;
    coordinate_system(wc, &wcx, &wcy);
    Float sin_thetamax = sp.radius * invdc;
    Float sin_thetamax2 = sin_thetamax * sin_thetamax;
    Float inv_sin_thetamax = 1.0f / sin_thetamax;
    Float cos_thetamax = std::sqrt(std::fmax(1.0f - sin_thetamax2, 0.0f));
    Float cos_theta = (cos_thetamax - 1.0f) * u.x + 1.0f;
    Float sin_theta2 = 1.0f - cos_theta * cos_theta;
    if (sin_thetamax2 < 0.00068523f) {  // Taylor expansion for small angles
        sin_theta2 = sin_thetamax2 * u.x;
        cos_theta = std::sqrt(1.0f - sin_theta2);
    }
    Float cos_alpha = sin_theta2 * inv_sin_thetamax + cos_theta * std::sqrt(std::fmax(1.0f - sin_theta2 * inv_sin_thetamax * inv_sin_thetamax, 0.0f));
    Float sin_alpha = std::sqrt(std::fmax(1.0f - cos_alpha * cos_alpha, 0.0f));
    Float phi = u.y * 2.0f * PI;
    // spherical_direction_basis(sin_alpha, cos_alpha, phi, -wcx, -wcy, -wc), geometry.rs:36-38
    V3 nworld = (wcx * -1.0f) * sin_alpha * std::cos(phi) + (wcy * -1.0f) * sin_alpha * std::sin(phi) + (wc * -1.0f) * cos_alpha;
    V3 pworld = pcenter + nworld * sp.radius;
    InteractionData it;  // it.n stays (0,0,0): sphere.rs:371-374 never sets it (quirk a-Q8) => only two-sided sphere lights emit through light sampling
    it.p = pworld;
    it.p_error = vabs(pworld) * gamma(5);
    *pdf = 1.0f / (2.0f * PI * (1.0f - cos_thetamax));
    return it;
}
inline Float sphere_pdf_wi(const RenderScene& s, const pbrt_b200_light& l, const InteractionData& ref, V3 wi) {  // sphere.rs:382-395
    const pbrt_b200_sphere& sp = s.d.spheres[l.shape_index];
    V3 pcenter = m4_point(m4_from(sp.object_to_world), V3(0, 0, 0));
    V3 porigin = offset_ray_origin(ref.p, ref.p_error, ref.n, pcenter - ref.p);
    if (distance_squared(porigin, pcenter) <= sp.radius * sp.radius) {  // shape_pdfwi, shape.rs:63-82
        Ray ray = spawn_ray(ref.p, ref.p_error, ref.n, wi, ref.time);
        Float t;
        if (!sphere_test(sp, ray, &t, nullptr)) return 0.0f;
        SurfaceInteraction isect = sphere_interaction(sp, ray, t);
        Float pdf = distance_squared(ref.p, isect.p) / (dot(isect.n, -wi) * l.area);  // signed dot (quirk a-Q2)
        if (std::isinf(pdf)) pdf = 0.0f;
        return pdf;
    }
    Float sin_thetamax2 = sp.radius * sp.radius / distance_squared(ref.p, pcenter);
    Float cos_thetamax = std::sqrt(std::fmax(1.0f - sin_thetamax2, 0.0f));
    return 1.0f / (2.0f * PI * (1.0f - cos_thetamax));  // uniform_cone_pdf
}

// Light::sample_li
inline LightSample light_sample_li(const RenderScene& s, int li, const InteractionData& ref, P2 u) {
    const pbrt_b200_light& l = s.d.lights[li];
    LightSample r;
    Spectrum L = spec3(l.L);
    switch (l.type) {
        case PBRT_B200_LIGHT_POINT: {  // point.rs:53-69
            V3 pl(l.pos[0], l.pos[1], l.pos[2]);
            r.wi = normalize(pl - ref.p); r.pdf = 1.0f;
            r.p1.p = pl; r.p1.time = ref.time;
            r.Li = L / distance_squared(pl, ref.p);
            break;
        }
        case PBRT_B200_LIGHT_SPOT: {  // spot.rs:46-59,70-85
            V3 pl(l.pos[0], l.pos[1], l.pos[2]);
            r.wi = normalize(pl - ref.p); r.pdf = 1.0f;
            r.p1.p = pl; r.p1.time = ref.time;
            V3 wl = normalize(m4_vector(m4_from(l.world_to_light), -r.wi));
            Float ct = wl.z, fall;
            if (ct < l.cos_total_width) fall = 0.0f;
            else if (ct >= l.cos_falloff_start) fall = 1.0f;
            else { Float delta = (ct - l.cos_total_width) / (l.cos_falloff_start - l.cos_total_width); fall = (delta * delta) * (delta * delta); }
            r.Li = L * fall / distance_squared(pl, ref.p);
            break;
        }
        case PBRT_B200_LIGHT_DISTANT: {  // distant.rs:66-83
            V3 w(l.dir[0], l.dir[1], l.dir[2]);
            r.wi = w; r.pdf = 1.0f;
            r.p1.p = ref.p + w * (2.0f * s.world_radius); r.p1.time = ref.time;
            r.Li = L;
            break;
        }
        case PBRT_B200_LIGHT_DIFFUSE: {  // diffuse.rs:91-106
            Float pdf;
            InteractionData ps = l.shape_kind == PBRT_B200_SHAPE_SPHERE ? sphere_sample_interaction(s, l, ref, u, &pdf) : triangle_sample_interaction(s, l, ref, u, &pdf);
            if (pdf == 0.0f || length_squared(ps.p - ref.p) == 0.0f) { r.pdf = 0.0f; r.Li = Spectrum(0.0f); return r; }
            r.pdf = pdf;
            r.wi = normalize(ps.p - ref.p);
            r.Li = (l.two_sided || dot(ps.n, -r.wi) > 0.0f) ? L : Spectrum(0.0f);  // AreaLight::l, diffuse.rs:68-75
            r.p1 = ps; r.p1.time = ref.time;
            break;
        }
        case PBRT_B200_LIGHT_INFINITE: {  // infinite.rs:141-170 (light_to_world = identity)
            Float map_pdf = 0.0f;
            P2 uv = s.inf_distrib[li].sample_continuous(u, &map_pdf);
            if (map_pdf == 0.0f) { r.Li = Spectrum(0.0f); r.pdf = 0.0f; return r; }
            Float theta = uv.y * PI, phi = uv.x * 2.0f * PI;
            Float cos_theta = std::cos(theta), sin_theta = std::sin(theta), sin_phi = std::sin(phi), cos_phi = std::cos(phi);
            r.wi = V3(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta);
            r.pdf = map_pdf / (2.0f * PI * PI * sin_theta);
            if (sin_theta == 0.0f) r.pdf = 0.0f;
            r.p1.p = ref.p + r.wi * (2.0f * s.world_radius); r.p1.time = ref.time;
            r.Li = L;
            break;
        }
    }
    return r;
}

// Light::pdf_li
inline Float light_pdf_li(const RenderScene& s, int li, const InteractionData& ref, V3 wi) {
    const pbrt_b200_light& l = s.d.lights[li];
    if (l.type == PBRT_B200_LIGHT_DIFFUSE) return l.shape_kind == PBRT_B200_SHAPE_SPHERE ? sphere_pdf_wi(s, l, ref, wi) : triangle_pdf_wi(s, l, ref, wi);
    if (l.type == PBRT_B200_LIGHT_INFINITE) {  // infinite.rs:131-139
        Float theta = spherical_theta(wi), phi = spherical_phi(wi);
        Float sin_theta = std::sin(theta);
        if (sin_theta == 0.0f) return 0.0f;
        return s.inf_distrib[li].pdf(P2(phi * INV2_PI, theta * INV_PI)) / (2.0f * PI * PI * sin_theta);
    }
    return 0.0f;
}
// Light::le (only infinite lights return non-zero; infinite.rs:120-129, constant map)
inline Spectrum light_le(const RenderScene& s, int li) {
    const pbrt_b200_light& l = s.d.lights[li];
    return l.type == PBRT_B200_LIGHT_INFINITE ? spec3(l.L) : Spectrum(0.0f);
}
inline bool is_delta_light(const pbrt_b200_light& l) { return l.type == PBRT_B200_LIGHT_POINT || l.type == PBRT_B200_LIGHT_DISTANT || l.type == PBRT_B200_LIGHT_SPOT; }

// SurfaceInteraction::le, interaction.rs:344-349 + AreaLight::l
inline Spectrum surface_le(const RenderScene& s, const SurfaceInteraction& si, V3 w) {
    int al = s.d.prims[si.slot].area_light;
    if (al < 0) return Spectrum(0.0f);
    const pbrt_b200_light& l = s.d.lights[al];
    return (l.two_sided || dot(si.n, w) > 0.0f) ? spec3(l.L) : Spectrum(0.0f);
}

struct RenderCounters {
    uint64_t camera_rays = 0, intersection_tests = 0, shadow_tests = 0, zero_radiance = 0, direct_den = 0;
    Counters trav_closest, trav_any;
};

// estimate_direct, src/core/integrator.rs:109-237 (handle_media = false, specular = false)
inline Spectrum estimate_direct(const RenderScene& s, const SurfaceInteraction& it, const BSDF& bsdf, P2 uscatt, int li, P2 ulight, RenderCounters& rc) {
    const pbrt_b200_light& light = s.d.lights[li];
    const int flags = BSDF_ALL & ~BSDF_SPECULAR;
    Spectrum Ld(0.0f);
    InteractionData ref; ref.p = it.p; ref.p_error = it.p_error; ref.n = it.n; ref.time = it.time;
    LightSample ls = light_sample_li(s, li, ref, ulight);
    Float lightpdf = ls.pdf, scattpdf = 0.0f;
    V3 wi = ls.wi;
    Spectrum Li = ls.Li;
    if (lightpdf > 0.0f && !Li.is_black()) {
        Spectrum f = bsdf.f(it.wo, wi, flags) * abs_dot(wi, it.sh_n);
        scattpdf = bsdf.pdf(it.wo, wi, flags);
        if (!f.is_black()) {
            Ray r = spawn_ray_to(ref, ls.p1);
            rc.shadow_tests++;
            if (scene_intersect_p(s, r, &rc.trav_any)) Li = Spectrum(0.0f);
            if (!Li.is_black()) {
                if (is_delta_light(light)) Ld += f * Li / lightpdf;
                else { Float weight = power_heuristic(1, lightpdf, 1, scattpdf); Ld += f * Li * weight / lightpdf; }
            }
        }
    }
    if (!is_delta_light(light)) {
        int sampled_type = 0;
        Spectrum f = bsdf.sample_f(it.wo, &wi, uscatt, &scattpdf, flags, &sampled_type);
        f = f * abs_dot(wi, it.sh_n);
        bool sampled_specular = (sampled_type & BSDF_SPECULAR) != 0;
        if (!f.is_black() && scattpdf > 0.0f) {
            Float weight = 1.0f;
            if (!sampled_specular) {
                lightpdf = light_pdf_li(s, li, ref, wi);
                if (lightpdf == 0.0f) return Ld;
                weight = power_heuristic(1, scattpdf, 1, lightpdf);
            }
            Ray ray = spawn_ray(it.p, it.p_error, it.n, wi, it.time);
            Hit h;
            rc.intersection_tests++;
            bool found = scene_intersect(s, ray, &h, &rc.trav_closest);
            Spectrum li_(0.0f);
            if (found) {
                if (s.d.prims[h.slot].area_light == li) {  // Arc::ptr_eq, integrator.rs:221-228
                    Ray r0 = spawn_ray(it.p, it.p_error, it.n, wi, it.time);
                    SurfaceInteraction lsi = make_interaction(s, r0, h);
                    li_ = surface_le(s, lsi, -wi);
                }
            } else li_ = light_le(s, li);
            if (!li_.is_black()) Ld += f * li_ * weight / scattpdf;
        }
    }
    return Ld;
}

// SpatialLightDistribution, src/core/lightdistrib.rs:105-340.  The reference keeps the per-voxel distributions in a
// lock-free hash table filled on first touch; a voxel's distribution is a pure function of (scene, voxel), so a dense
// lazily-filled table gives the same answers.
struct SpatialLightDistribution {
    const RenderScene* scene = nullptr;
    size_t nvoxels[3] = {1, 1, 1};
    mutable std::vector<std::atomic<const Distribution1D*>> table;
    mutable std::atomic<uint64_t> ncreated{0};
    SpatialLightDistribution(const RenderScene* s, size_t max_voxels) : scene(s) {  // lightdistrib.rs:113-150
        const Bounds3& b = s->wb;
        V3 diag = b.p_max - b.p_min;
        int me = (diag.x > diag.y && diag.x > diag.z) ? 0 : (diag.y > diag.z ? 1 : 2);  // bounds.rs:346-360
        Float bmax = diag[me];
        for (int i = 0; i < 3; ++i) nvoxels[i] = std::max<size_t>(1, (size_t)f2u_sat(std::round(diag[i] / bmax * (Float)max_voxels)));
        table = std::vector<std::atomic<const Distribution1D*>>(nvoxels[0] * nvoxels[1] * nvoxels[2]);
        for (auto& e : table) e.store(nullptr);
    }
    ~SpatialLightDistribution() { for (auto& e : table) delete e.load(); }
    // compute_dsitribution, lightdistrib.rs:152-228
    Distribution1D compute(const int64_t pi[3]) const {
        const Bounds3& wb = scene->wb;
        V3 p0((Float)pi[0] / (Float)nvoxels[0], (Float)pi[1] / (Float)nvoxels[1], (Float)pi[2] / (Float)nvoxels[2]);
        V3 p1((Float)(pi[0] + 1) / (Float)nvoxels[0], (Float)(pi[1] + 1) / (Float)nvoxels[1], (Float)(pi[2] + 1) / (Float)nvoxels[2]);
        auto blerp = [](const Bounds3& b, V3 t) { return V3(lerp(t.x, b.p_min.x, b.p_max.x), lerp(t.y, b.p_min.y, b.p_max.y), lerp(t.z, b.p_min.z, b.p_max.z)); };
        V3 a = blerp(wb, p0), c = blerp(wb, p1);
        Bounds3 vb(V3(std::fmin(a.x, c.x), std::fmin(a.y, c.y), std::fmin(a.z, c.z)), V3(std::fmax(a.x, c.x), std::fmax(a.y, c.y), std::fmax(a.z, c.z)));
        const size_t nl = scene->d.n_lights;
        const int nsamples = 128;
        std::vector<Float> contrib(nl, 0.0f);
        for (int i = 0; i < nsamples; ++i) {
            V3 po = blerp(vb, V3(radical_inverse(0, (uint64_t)i), radical_inverse(1, (uint64_t)i), radical_inverse(2, (uint64_t)i)));
            InteractionData intr; intr.p = po; intr.p_error = V3(0, 0, 0); intr.n = V3(0, 0, 0); intr.time = 0.0f;
            P2 u(radical_inverse(3, (uint64_t)i), radical_inverse(4, (uint64_t)i));
            for (size_t j = 0; j < nl; ++j) {
                LightSample ls = light_sample_li(*scene, (int)j, intr, u);
                if (ls.pdf > 0.0f) contrib[j] += ls.Li.y() / ls.pdf;
            }
        }
        Float sum = 0.0f;
        for (Float v : contrib) sum += v;
        Float avg = sum / ((Float)nsamples * (Float)nl);
        Float min_contrib = avg > 0.0f ? 0.001f * avg : 1.0f;
        for (Float& v : contrib) v = std::fmax(v, min_contrib);
        return Distribution1D(contrib);
    }
    void voxel_of(V3 p, int64_t pi[3]) const {  // lightdistrib.rs:236-246
        const Bounds3& wb = scene->wb;
        V3 o = p - wb.p_min;  // Bounds3::offset, bounds.rs:372-390
        if (wb.p_max.x > wb.p_min.x) o.x /= wb.p_max.x - wb.p_min.x;
        if (wb.p_max.y > wb.p_min.y) o.y /= wb.p_max.y - wb.p_min.y;
        if (wb.p_max.z > wb.p_min.z) o.z /= wb.p_max.z - wb.p_min.z;
        for (int i = 0; i < 3; ++i) pi[i] = clamp<int64_t>(f2i_sat(o[i] * (Float)nvoxels[i]), 0, (int64_t)nvoxels[i] - 1);
    }
    const Distribution1D& lookup(V3 p) const {
        int64_t pi[3];
        voxel_of(p, pi);
        size_t idx = ((size_t)pi[2] * nvoxels[1] + (size_t)pi[1]) * nvoxels[0] + (size_t)pi[0];
        const Distribution1D* d = table[idx].load(std::memory_order_acquire);
        if (!d) {
            Distribution1D* fresh = new Distribution1D(compute(pi));
            const Distribution1D* expected = nullptr;
            if (table[idx].compare_exchange_strong(expected, fresh, std::memory_order_acq_rel)) { d = fresh; ncreated++; }
            else { delete fresh; d = expected; }
        }
        return *d;
    }
};

struct IntegratorParams {
    int kind = PBRT_B200_INTEGRATOR_PATH;
    std::vector<int> nlight_samples;  // DirectLightingIntegrator::preprocess, directlighting.rs:61-76
    int max_depth = 5;
    Float rr_threshold = 1.0f;
    int pixel_bounds[4];
    Distribution1D light_distrib;                       // uniform or power
    std::shared_ptr<SpatialLightDistribution> spatial;  // "spatial" with more than one light
    const Distribution1D& lookup(V3 p) const { return spatial ? spatial->lookup(p) : light_distrib; }  // LightDistribution::lookup
};

// uniform_sample_onelight, src/core/integrator.rs:81-106
inline Spectrum uniform_sample_onelight(const RenderScene& s, const SurfaceInteraction& it, const BSDF& bsdf, Sampler& sampler, const Distribution1D& distrib,
                                        RenderCounters& rc) {
    size_t nlights = s.d.n_lights;
    if (nlights == 0) return Spectrum(0.0f);
    Float lightpdf = 0.0f;
    size_t lightnum = distrib.sample_discrete(sampler.get_1d(), &lightpdf);
    if (lightpdf == 0.0f) return Spectrum(0.0f);
    P2 ulight = sampler.get_2d();
    P2 uscattering = sampler.get_2d();
    return estimate_direct(s, it, bsdf, uscattering, (int)lightnum, ulight, rc) / lightpdf;
}

// PathIntegrator::li, src/integrators/path.rs:79-222
inline Spectrum path_li(const RenderScene& s, const IntegratorParams& ip, Ray ray, Sampler& sampler, RenderCounters& rc) {
    Spectrum L(0.0f), beta(1.0f);
    bool specular_bounce = false;
    int bounces = 0;
    Float etascale = 1.0f;
    for (;;) {
        Hit h;
        Ray r0 = ray;
        rc.intersection_tests++;
        bool found = scene_intersect(s, ray, &h, &rc.trav_closest);
        SurfaceInteraction isect;
        if (found) isect = make_interaction(s, r0, h);
        if (bounces == 0 || specular_bounce) {
            if (found) L += surface_le(s, isect, -ray.d) * beta;
            else for (int li : s.infinite_lights) L += light_le(s, li) * beta;
        }
        if (!found || bounces >= ip.max_depth) break;
        BSDF bsdf;
        int mat = s.d.prims[h.slot].material;
        (void)mat;
        scattering_functions(s, ray, h.slot, isect, &bsdf, true);  // isect.compute_scattering_functions(&ray, ..), path.rs:123
        if (!bsdf.valid) {  // path.rs:124-129
            ray = spawn_ray(isect.p, isect.p_error, isect.n, ray.d, isect.time);
            continue;
        }
        const Distribution1D& distrib = ip.lookup(isect.p);  // path.rs:132
        if (bsdf.num_components(BSDF_ALL & ~BSDF_SPECULAR) > 0) {
            rc.direct_den++;
            Spectrum Ld = beta * uniform_sample_onelight(s, isect, bsdf, sampler, distrib, rc);
            if (Ld.is_black()) rc.zero_radiance++;
            L += Ld;
        }
        V3 wo = -ray.d, wi;
        Float pdf = 0.0f;
        int flags = 0;
        Spectrum f = bsdf.sample_f(wo, &wi, sampler.get_2d(), &pdf, BSDF_ALL, &flags);
        if (f.is_black() || pdf == 0.0f) break;
        beta *= f * abs_dot(wi, isect.sh_n) / pdf;
        specular_bounce = (flags & BSDF_SPECULAR) != 0;
        if ((flags & BSDF_SPECULAR) && (flags & BSDF_TRANSMISSION)) {
            Float eta = bsdf.eta;
            etascale *= (dot(wo, isect.n) > 0.0f) ? eta * eta : 1.0f / (eta * eta);
        }
        ray = spawn_ray(isect.p, isect.p_error, isect.n, wi, isect.time);
        Spectrum rrbeta = beta * etascale;
        if (rrbeta.max_component_value() < ip.rr_threshold && bounces > 3) {
            Float q = std::fmax(1.0f - rrbeta.max_component_value(), 0.05f);
            if (sampler.get_1d() < q) break;
            beta = beta / (1.0f - q);
        }
        bounces += 1;
    }
    return L;
}

// PerspectiveCamera::generate_ray_differential, src/cameras/perspective.rs:120-179; dx_camera / dy_camera :64-70.
// `spp` > 0 applies Ray::scale_differential(1 / sqrt(spp)) the way the render loop does (integrator.rs:341, ray.rs:34-41).
inline Ray generate_ray(const pbrt_b200_camera& c, const CameraSample& cs, uint32_t spp = 0) {
    M4 r2c = m4_from(c.raster_to_camera), c2w = m4_from(c.camera_to_world);
    V3 pcamera = m4_point(r2c, V3(cs.pfilm.x, cs.pfilm.y, 0.0f));
    Ray r(V3(0, 0, 0), normalize(pcamera), INFINITY_F, 0.0f);
    V3 p2t = m4_point(r2c, V3(0, 0, 0));
    V3 dx_camera = m4_point(r2c, V3(1, 0, 0)) - p2t, dy_camera = m4_point(r2c, V3(0, 1, 0)) - p2t;
    if (c.lens_radius > 0.0f) {
        P2 pl = concentric_sample_disk(cs.plens);
        pl = P2(pl.x * c.lens_radius, pl.y * c.lens_radius);
        Float ft = c.focal_distance / r.d.z;
        V3 pfocus = r.o + r.d * ft;
        r.o = V3(pl.x, pl.y, 0.0f);
        r.d = normalize(pfocus - r.o);
        V3 dx = normalize(pcamera + dx_camera);
        ft = c.focal_distance / dx.z;
        pfocus = V3(0, 0, 0) + dx * ft;
        r.rxo = V3(pl.x, pl.y, 0.0f);
        r.rxd = normalize(pfocus - r.rxo);
        V3 dy = normalize(pcamera + dy_camera);
        ft = c.focal_distance / dy.z;
        pfocus = V3(0, 0, 0) + dy * ft;
        r.ryo = V3(pl.x, pl.y, 0.0f);
        r.ryd = normalize(pfocus - r.ryo);
    } else {
        r.rxo = r.o; r.ryo = r.o;
        r.rxd = normalize(pcamera + dx_camera);
        r.ryd = normalize(pcamera + dy_camera);
    }
    r.time = lerp(cs.time, c.shutter_open, c.shutter_close);
    Ray w = m4_ray(c2w, r);  // Transform::transform_ray carries the differentials over (transform.rs:562-572)
    w.has_diff = true;
    w.rxo = m4_point(c2w, r.rxo); w.ryo = m4_point(c2w, r.ryo);
    w.rxd = m4_vector(c2w, r.rxd); w.ryd = m4_vector(c2w, r.ryd);
    if (spp > 0) {
        Float sc = 1.0f / std::sqrt((Float)spp);
        w.rxo = w.o + (w.rxo - w.o) * sc; w.ryo = w.o + (w.ryo - w.o) * sc;
        w.rxd = w.d + (w.rxd - w.d) * sc; w.ryd = w.d + (w.ryd - w.d) * sc;
    }
    return w;
}

// FilmTile::add_sample, src/core/film.rs:292-331, accumulating straight into the film-sized
// {r,g,b,w} buffer (clipping to the tile's pixel bounds == clipping to the crop window because
// a tile's film bounds cover every pixel its samples can reach, film.rs:126-140).
struct FilmAccum {
    const pbrt_b200_film* film;
    int x0, y0, x1, y1, width;
    std::vector<double> sum;  // oracle keeps f64 sums so that thread/tile order cannot matter
    void init(const pbrt_b200_film* f) {
        film = f; x0 = f->cropped_pixel_bounds[0]; y0 = f->cropped_pixel_bounds[1]; x1 = f->cropped_pixel_bounds[2]; y1 = f->cropped_pixel_bounds[3];
        width = x1 - x0;
        sum.assign((size_t)4 * width * (y1 - y0), 0.0);
    }
};
template <typename AddFn>
inline void film_add_sample(const pbrt_b200_film& film, P2 pfilm, Spectrum L, Float sample_weight, AddFn add) {
    if (L.y() > film.max_sample_luminance) L *= Spectrum(film.max_sample_luminance / L.y());
    Float rx = film.filter_radius[0], ry = film.filter_radius[1];
    Float irx = 1.0f / rx, iry = 1.0f / ry;
    Float dx = pfilm.x - 0.5f, dy = pfilm.y - 0.5f;
    int64_t p0x = f2i_sat(std::ceil(dx - rx)), p0y = f2i_sat(std::ceil(dy - ry));
    int64_t p1x = f2i_sat(std::floor(dx + rx)) + 1, p1y = f2i_sat(std::floor(dy + ry)) + 1;
    p0x = std::max<int64_t>(p0x, film.cropped_pixel_bounds[0]); p0y = std::max<int64_t>(p0y, film.cropped_pixel_bounds[1]);
    p1x = std::min<int64_t>(p1x, film.cropped_pixel_bounds[2]); p1y = std::min<int64_t>(p1y, film.cropped_pixel_bounds[3]);
    const int TW = 16;
    for (int64_t y = p0y; y < p1y; ++y) {
        Float fy = std::fabs(((Float)y - dy) * iry * (Float)TW);
        int iy = (int)std::min<int64_t>(f2u_sat(std::floor(fy)), TW - 1);
        for (int64_t x = p0x; x < p1x; ++x) {
            Float fx = std::fabs(((Float)x - dx) * irx * (Float)TW);
            int ix = (int)std::min<int64_t>(f2u_sat(std::floor(fx)), TW - 1);
            Float fw = film.filter_table[iy * TW + ix];
            Spectrum c = L * Spectrum(sample_weight) * Spectrum(fw);
            add((int)x, (int)y, c, fw);
        }
    }
}

// uniform_sample_all_lights, src/core/integrator.rs:40-79 (handle_media = false); nlight_samples[j] =
// round_count(light.nsamples()) with nsamples() = pbrt_b200_light.n_samples.
inline Spectrum uniform_sample_all_lights(const RenderScene& s, const SurfaceInteraction& it, const BSDF& bsdf, Sampler& sampler,
                                          const std::vector<int>& nlight_samples, RenderCounters& rc) {
    Spectrum L(0.0f);
    std::vector<P2> larray, sarray;
    for (size_t j = 0; j < s.d.n_lights; ++j) {
        int nsamples = nlight_samples[j];
        bool have_l = sampler.get_2d_array(nsamples, &larray);
        bool have_s = sampler.get_2d_array(nsamples, &sarray);
        if (!have_l || !have_s) {
            P2 ulight = sampler.get_2d();
            P2 uscattering = sampler.get_2d();
            L += estimate_direct(s, it, bsdf, uscattering, (int)j, ulight, rc);
        } else {
            Spectrum Ld(0.0f);
            for (int k = 0; k < nsamples; ++k) Ld += estimate_direct(s, it, bsdf, sarray[k], (int)j, larray[k], rc);
            L += Ld / (Float)nsamples;
        }
    }
    return L;
}

// DirectLightingIntegrator::li (directlighting.rs:78-119) and WhittedIntegrator::li (whitted.rs:52-105) with
// SamplerIntegrator::specular_reflect / specular_transmit (integrator.rs:409-520; the ray differentials they build only
// feed texture filtering, which this path does not have).
inline Spectrum recursive_li(const RenderScene& s, const IntegratorParams& ip, Ray ray, Sampler& sampler, RenderCounters& rc, int depth) {
    Spectrum L(0.0f);
    Hit h;
    Ray r0 = ray;
    rc.intersection_tests++;
    if (!scene_intersect(s, ray, &h, &rc.trav_closest)) {
        for (size_t li = 0; li < s.d.n_lights; ++li) L += light_le(s, (int)li);  // every light's le(): non-zero for infinite lights only
        return L;
    }
    SurfaceInteraction isect = make_interaction(s, r0, h);
    BSDF bsdf;
    int mat = s.d.prims[h.slot].material;
    (void)mat;
    scattering_functions(s, ray, h.slot, isect, &bsdf, false);
    if (!bsdf.valid) return recursive_li(s, ip, spawn_ray(isect.p, isect.p_error, isect.n, ray.d, isect.time), sampler, rc, depth);
    V3 wo = isect.wo;
    L += surface_le(s, isect, wo);
    if (ip.kind == PBRT_B200_INTEGRATOR_WHITTED) {
        InteractionData ref; ref.p = isect.p; ref.p_error = isect.p_error; ref.n = isect.n; ref.time = isect.time;
        for (size_t li = 0; li < s.d.n_lights; ++li) {  // whitted.rs:85-97
            LightSample ls = light_sample_li(s, (int)li, ref, sampler.get_2d());
            if (ls.Li.is_black() || ls.pdf == 0.0f) continue;
            Spectrum f = bsdf.f(wo, ls.wi, BSDF_ALL);
            if (!f.is_black()) {
                rc.shadow_tests++;
                Ray sh = spawn_ray_to(ref, ls.p1);
                if (!scene_intersect_p(s, sh, &rc.trav_any)) L += f * ls.Li * abs_dot(ls.wi, isect.n) / ls.pdf;
            }
        }
    } else if (s.d.n_lights > 0) {
        if (ip.kind == PBRT_B200_INTEGRATOR_DIRECT_ALL) L += uniform_sample_all_lights(s, isect, bsdf, sampler, ip.nlight_samples, rc);
        else {  // uniform_sample_onelight with no distribution: light_num = min(u * n, n - 1), pdf 1/n (integrator.rs:88-97)
            size_t nl = s.d.n_lights;
            Float u = sampler.get_1d();
            size_t ln = std::min((size_t)f2u_sat(u * (Float)nl), nl - 1);
            Float lightpdf = 1.0f / (Float)nl;
            P2 ulight = sampler.get_2d();
            P2 uscattering = sampler.get_2d();
            L += estimate_direct(s, isect, bsdf, uscattering, (int)ln, ulight, rc) / lightpdf;
        }
    }
    if (depth + 1 < ip.max_depth) {
        for (int pass = 0; pass < 2; ++pass) {  // specular_reflect, then specular_transmit
            V3 wi;
            Float pdf = 0.0f;
            int st = 0;
            Spectrum f = bsdf.sample_f(wo, &wi, sampler.get_2d(), &pdf, (pass == 0 ? BSDF_REFLECTION : BSDF_TRANSMISSION) | BSDF_SPECULAR, &st);
            if (pdf > 0.0f && !f.is_black() && abs_dot(wi, isect.sh_n) != 0.0f) {
                Ray rd = spawn_ray(isect.p, isect.p_error, isect.n, wi, isect.time);
                if (ray.has_diff) {  // `if let Some(ref diff) = r.diff`, integrator.rs:427-452 / 476-513
                    V3 ns = isect.sh_n;
                    rd.has_diff = true;
                    rd.rxo = isect.p + isect.dpdx; rd.ryo = isect.p + isect.dpdy;
                    V3 dndx = isect.sh_dndu * isect.dudx + isect.sh_dndv * isect.dvdx;
                    V3 dndy = isect.sh_dndu * isect.dudy + isect.sh_dndv * isect.dvdy;
                    V3 dwodx = -ray.rxd - wo, dwody = -ray.ryd - wo;
                    Float ddndx = dot(dwodx, ns) + dot(wo, dndx), ddndy = dot(dwody, ns) + dot(wo, dndy);
                    if (pass == 0) {
                        rd.rxd = wi - dwodx + (dndx * dot(wo, ns) + ns * ddndx) * 2.0f;
                        rd.ryd = wi - dwody + (dndy * dot(wo, ns) + ns * ddndy) * 2.0f;
                    } else {
                        Float eta = bsdf.eta;
                        V3 w = -wo;
                        if (dot(wo, ns) < 0.0f) eta = 1.0f / eta;  // (the reference inverts eta but does not negate ns, :493-497)
                        Float mu = eta * dot(w, ns) - dot(wi, ns);
                        Float dmudx = (eta - (eta * eta * dot(w, ns)) / dot(wi, ns)) * ddndx;
                        Float dmudy = (eta - (eta * eta * dot(w, ns)) / dot(wi, ns)) * ddndy;
                        rd.rxd = wi + dwodx * eta - (dndx * mu + ns * dmudx);
                        rd.ryd = wi + dwody * eta - (dndy * mu + ns * dmudy);
                    }
                }
                Spectrum Li = recursive_li(s, ip, rd, sampler, rc, depth + 1);
                if (pass == 0) L += f * Li * abs_dot(wi, isect.sh_n) / pdf;
                else L += f * Li * (abs_dot(wi, isect.sh_n) / pdf);
            }
        }
    }
    return L;
}

struct RenderJob {
    RenderScene scene;
    pbrt_b200_render_desc rd;
    SamplerTables tables;
    IntegratorParams ip;
};

std::unique_ptr<Sampler> make_sampler(const pbrt_b200_sampler& sd, const SamplerTables& t);  // oracle_sampling_extra.hpp

inline void setup_job(RenderJob& job, const pbrt_b200_scene_desc& sdesc, const pbrt_b200_render_desc& rd) {
    job.scene.init_render(sdesc);
    job.rd = rd;
    job.tables.sobol32 = rd.sampler.sobol_matrices32; job.tables.vdc = rd.sampler.vdc_matrices; job.tables.vdc_inv = rd.sampler.vdc_matrices_inv;
    job.ip.max_depth = rd.integrator.max_depth; job.ip.rr_threshold = rd.integrator.rr_threshold;
    job.ip.kind = (int)rd.integrator.kind;
    for (int i = 0; i < 4; ++i) job.ip.pixel_bounds[i] = rd.integrator.pixel_bounds[i];
    // create_light_sample_distribution, lightdistrib.rs:20-31
    size_t nl = sdesc.n_lights;
    std::vector<Float> f(nl, 1.0f);
    const bool uniform = rd.integrator.light_sample_strategy == PBRT_B200_LIGHTS_UNIFORM || nl == 1;
    if (!uniform && rd.integrator.light_sample_strategy == PBRT_B200_LIGHTS_POWER)
        for (size_t i = 0; i < nl; ++i) f[i] = light_power(job.scene, sdesc.lights[i]).y();
    job.ip.light_distrib = Distribution1D(f);
    job.ip.spatial.reset();
    if (!uniform && rd.integrator.light_sample_strategy != PBRT_B200_LIGHTS_POWER && nl > 0)
        job.ip.spatial = std::make_shared<SpatialLightDistribution>(&job.scene, 64);
}

}  // namespace orc
#include "oracle_volpath.hpp"
namespace orc {

// SamplerIntegrator::render, src/core/integrator.rs:263-403.  rgbw is ADDED to.
inline void render(const RenderJob& job, float* rgbw, int nthreads, RenderCounters* total) {
    const pbrt_b200_render_desc& rd = job.rd;
    const int* sb = rd.sampler.sample_bounds;
    const int tilesize = 16;
    int ntx = (sb[2] - sb[0] + tilesize - 1) / tilesize, nty = (sb[3] - sb[1] + tilesize - 1) / tilesize;
    // pbrt_b200_render_desc.tile_order (include/pbrt_b200.h): 0 = the reference's row-major tile numbering; S = S x S-tile super-tiles
    const uint32_t S = rd.tile_order, nstx = S ? (uint32_t)((ntx + (int)S - 1) / (int)S) : 0u, nsty = S ? (uint32_t)((nty + (int)S - 1) / (int)S) : 0u;
    const uint32_t n_positions = S ? nstx * nsty * S * S : (uint32_t)(ntx * nty);
    uint32_t tile_begin = rd.tile_begin, tile_end = rd.tile_end ? rd.tile_end : n_positions;
    uint32_t s_begin = rd.sample_begin, s_end = rd.sample_end ? rd.sample_end : 0xffffffffu;  // 0 => until start_next_sample() says stop
    FilmAccum acc; acc.init(&rd.film);
    std::mutex mu;
    std::atomic<uint32_t> next(tile_begin);
    std::vector<RenderCounters> rcs(nthreads);
    const bool trace_samples = getenv("ORC_TRACE_SAMPLES") != nullptr;  // debugging aid: one line per camera sample on stderr
    auto worker = [&](int tid) {
        RenderCounters& rc = rcs[tid];
        std::unique_ptr<Sampler> base = make_sampler(rd.sampler, job.tables);
        IntegratorParams ipl = job.ip;
        if (ipl.kind == PBRT_B200_INTEGRATOR_DIRECT_ALL) {  // DirectLightingIntegrator::preprocess, directlighting.rs:61-76
            for (size_t j = 0; j < job.scene.d.n_lights; ++j) ipl.nlight_samples.push_back(base->round_count((int)std::max<uint32_t>(job.scene.d.lights[j].n_samples, 1u)));
            for (int i = 0; i < ipl.max_depth; ++i)
                for (size_t j = 0; j < job.scene.d.n_lights; ++j) { base->request_2d_array(ipl.nlight_samples[j]); base->request_2d_array(ipl.nlight_samples[j]); }
        }
        std::vector<double> local;
        for (;;) {
            uint32_t tile = next.fetch_add(1);
            if (tile >= tile_end) break;
            {   // interleaved tile ownership (pbrt_b200_render_desc.tile_group/mod/rem)
                uint32_t g = rd.tile_group ? rd.tile_group : 1u, m = rd.tile_mod ? rd.tile_mod : 1u;
                if (((tile - tile_begin) / g) % m != rd.tile_rem) continue;
            }
            int tx = tile % ntx, ty = tile / ntx;
            if (S) {
                const uint32_t st = tile / (S * S), w = tile - st * (S * S);
                tx = (int)((st % nstx) * S + w % S); ty = (int)((st / nstx) * S + w / S);
                if (tx >= ntx || ty >= nty) continue;  // position past the image edge
            }
            std::unique_ptr<Sampler> ts = base->clone((int64_t)ty * ntx + tx);
            int x0 = sb[0] + tx * tilesize, x1 = std::min(x0 + tilesize, sb[2]);
            int y0 = sb[1] + ty * tilesize, y1 = std::min(y0 + tilesize, sb[3]);
            struct Contrib { int x, y; Spectrum c; Float w; };
            std::vector<Contrib> contribs;
            for (int y = y0; y < y1; ++y)
                for (int x = x0; x < x1; ++x) {
                    ts->start_pixel(x, y);
                    const int* pb = job.ip.pixel_bounds;
                    if (!(x >= pb[0] && x < pb[2] && y >= pb[1] && y < pb[3])) continue;
                    bool more = true;
                    if (s_begin > 0) more = ts->set_sample_number(s_begin);
                    while (more) {
                        if (ts->current_pixel_sample_index >= s_end) break;
                        CameraSample cs = ts->get_camera_sample(x, y);
                        Ray ray = generate_ray(rd.camera, cs, rd.sampler.samples_per_pixel);
                        rc.camera_rays++;
                        Spectrum L = ipl.kind == PBRT_B200_INTEGRATOR_PATH ? path_li(job.scene, ipl, ray, *ts, rc)
                                     : ipl.kind == PBRT_B200_INTEGRATOR_VOLPATH ? volpath_li(job.scene, ipl, ray, rd.integrator.camera_medium, *ts, rc)
                                                                                : recursive_li(job.scene, ipl, ray, *ts, rc, 0);
                        if (trace_samples) fprintf(stderr, "orc sample %d %d %llu L %.9g %.9g %.9g ntests %llu\n", x, y, (unsigned long long)ts->current_pixel_sample_index,
                                                   L.c[0], L.c[1], L.c[2], (unsigned long long)rc.intersection_tests + rc.shadow_tests);
                        if (L.has_nans()) L = Spectrum(0.0f);                 // integrator.rs:350-368
                        else if (L.y() < -1.0e-5f) L = Spectrum(0.0f);
                        else if (std::isinf(L.y())) L = Spectrum(0.0f);
                        film_add_sample(rd.film, cs.pfilm, L, 1.0f, [&](int px, int py, Spectrum c, Float fw) { contribs.push_back({px, py, c, fw}); });
                        more = ts->start_next_sample();
                    }
                }
            std::lock_guard<std::mutex> g(mu);
            for (auto& c : contribs) {
                size_t o = 4 * ((size_t)(c.y - acc.y0) * acc.width + (c.x - acc.x0));
                acc.sum[o] += c.c.c[0]; acc.sum[o + 1] += c.c.c[1]; acc.sum[o + 2] += c.c.c[2]; acc.sum[o + 3] += c.w;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
    for (size_t i = 0; i < acc.sum.size(); ++i) rgbw[i] += (float)acc.sum[i];
    if (total)
        for (auto& rc : rcs) {
            total->camera_rays += rc.camera_rays; total->intersection_tests += rc.intersection_tests; total->shadow_tests += rc.shadow_tests;
            total->zero_radiance += rc.zero_radiance; total->direct_den += rc.direct_den;
            total->trav_closest.nodes_tested += rc.trav_closest.nodes_tested; total->trav_closest.tris_tested += rc.trav_closest.tris_tested;
            total->trav_closest.rays += rc.trav_closest.rays;
            total->trav_any.nodes_tested += rc.trav_any.nodes_tested; total->trav_any.tris_tested += rc.trav_any.tris_tested; total->trav_any.rays += rc.trav_any.rays;
        }
}

}  // namespace orc
