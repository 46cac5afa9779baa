// TEST INFRASTRUCTURE (see oracle.py): CPU restatement of pbrt-rust's VolPathIntegrator with homogeneous media.
//
//   VolPathIntegrator::li                      src/integrators/volpath.rs:82-221
//   HomogeneousMedium::{tr, sample}            src/media/homogeneous.rs:29-73
//   HenyeyGreenstein::{p, sample_p}, phase_hg  src/core/medium.rs:162-221
//   MediumInterface / GeometricPrimitive       src/core/medium.rs:131-160, src/core/primitive.rs:126-141
//   Interaction::get_medium_vec / spawn_ray    src/core/interaction.rs:32-66
//   VisibilityTester::tr                       src/core/light.rs:125-150
//   Scene::intersect_tr                        src/core/scene.rs:68-87
//   estimate_direct (handle_media = true)      src/core/integrator.rs:109-237
//
// Deviations of the reference from pbrt-v3 that are reproduced on purpose:
//   * volpath.rs:131-135: a surface without a BSDF (a medium boundary) executes `bounces -= 1; continue`, and `continue` skips the
//     `bounces += 1` at the end of the `loop` body (pbrt-v3's `for (...; ++bounces)` does not): every boundary crossing LOWERS the
//     bounce count by one, and at bounce 0 the `usize` wraps (release build: no overflow check), which ends the path at the next
//     `bounces >= max_depth` test.  A medium seen directly through its boundary therefore contributes nothing; media are visible when the
//     camera sits inside one (Camera.medium) or behind at least one real bounce.
//   * MediumInteraction.wo = -ray.d is not normalised (homogeneous.rs:52-54), as in pbrt-v3.
#pragma once
// Included by oracle_render.hpp (between the light / estimate_direct code and `render`); not a stand-alone header.

namespace orc {

struct MediumRef { int idx = -1; };  // index into scene_desc.media, -1 = vacuum

inline Spectrum medium_sigma_t(const pbrt_b200_medium& m) { return spec3(m.sigma_a) + spec3(m.sigma_s); }  // HomogeneousMedium::new

inline Spectrum spec_neg(Spectrum s) { return Spectrum(-s.c[0], -s.c[1], -s.c[2]); }
inline Spectrum spec_exp(Spectrum s) { return Spectrum(std::exp(s.c[0]), std::exp(s.c[1]), std::exp(s.c[2])); }

// HomogeneousMedium::tr, homogeneous.rs:30-33
inline Spectrum medium_tr(const pbrt_b200_medium& m, const Ray& ray) {
    Float dist = std::fmin(ray.t_max * length(ray.d), std::numeric_limits<Float>::max());
    return spec_exp(spec_neg(medium_sigma_t(m)) * dist);
}

// medium.rs:162-166
inline Float phase_hg(Float cos_theta, Float g) {
    Float denom = 1.0f + g * g + 2.0f * g * cos_theta;
    return INV4_PI * (1.0f - g * g) / (denom * std::sqrt(denom));
}
// HenyeyGreenstein::sample_p, medium.rs:184-221
inline Float hg_sample_p(Float g, V3 wo, V3* wi, P2 u) {
    Float cos_theta;
    if (std::fabs(g) < 1.0e-3f) cos_theta = 1.0f - 2.0f * u.x;
    else {
        Float sqr = (1.0f - g * g) / (1.0f + g - 2.0f * g * u.x);
        cos_theta = -(1.0f + g * g - sqr * sqr) / (2.0f * g);
    }
    Float sin_theta = std::sqrt(std::fmax(0.0f, 1.0f - cos_theta * cos_theta));
    Float phi = 2.0f * PI * u.y;
    V3 v1, v2;
    coordinate_system(wo, &v1, &v2);
    *wi = v1 * sin_theta * std::cos(phi) + v2 * sin_theta * std::sin(phi) + wo * cos_theta;  // spherical_direction_basis, geometry.rs:36-38
    return phase_hg(cos_theta, g);
}

// The medium interface an intersection carries (primitive.rs:134-140) and Interaction::get_medium_vec (interaction.rs:54-66)
struct VolInterface { int inside = -1, outside = -1; };
inline VolInterface hit_interface(const RenderScene& s, uint32_t slot, int ray_medium) {
    VolInterface mi;
    if (s.d.prim_media) {
        mi.inside = s.d.prim_media[slot].inside; mi.outside = s.d.prim_media[slot].outside;
        if (mi.inside != mi.outside) return mi;  // is_medium_transition
    }
    mi.inside = mi.outside = ray_medium;  // MediumInterface::new(ray.medium)
    return mi;
}
inline int medium_for(const VolInterface& mi, V3 n, V3 w) { return dot(w, n) > 0.0f ? mi.outside : mi.inside; }

struct VolRay { Ray r; int medium = -1; };

// VisibilityTester::tr, light.rs:125-150: p0 -> p1 through every material-less surface on the way
inline Spectrum visibility_tr(const RenderScene& s, const InteractionData& p0, const VolInterface& mi0, const InteractionData& p1, RenderCounters& rc) {
    VolRay vr;
    vr.r = spawn_ray_to(p0, p1);
    vr.medium = medium_for(mi0, p0.n, vr.r.d);
    Spectrum Tr(1.0f);
    for (;;) {
        Hit h;
        Ray r0 = vr.r;
        rc.intersection_tests++;  // Scene::intersect, not intersect_p
        bool hit = scene_intersect(s, vr.r, &h, &rc.trav_closest);
        if (hit && s.d.prims[h.slot].material >= 0) return Spectrum(0.0f);
        if (vr.medium >= 0) Tr *= medium_tr(s.d.media[vr.medium], vr.r);
        if (!hit) break;
        SurfaceInteraction isect = make_interaction(s, r0, h);
        InteractionData a; a.p = isect.p; a.p_error = isect.p_error; a.n = isect.n; a.time = isect.time;
        VolInterface mi = hit_interface(s, h.slot, vr.medium);
        vr.r = spawn_ray_to(a, p1);
        vr.medium = medium_for(mi, isect.n, vr.r.d);
    }
    return Tr;
}

// Scene::intersect_tr, scene.rs:68-87
inline bool intersect_tr(const RenderScene& s, VolRay vr, Hit* h_out, Ray* r_hit, Spectrum* Tr, RenderCounters& rc) {
    *Tr = Spectrum(1.0f);
    for (;;) {
        Hit h;
        Ray r0 = vr.r;
        rc.intersection_tests++;
        bool hit = scene_intersect(s, vr.r, &h, &rc.trav_closest);
        if (vr.medium >= 0) *Tr *= medium_tr(s.d.media[vr.medium], vr.r);
        if (!hit) return false;
        if (s.d.prims[h.slot].material >= 0) { *h_out = h; *r_hit = r0; return true; }
        SurfaceInteraction isect = make_interaction(s, r0, h);
        VolInterface mi = hit_interface(s, h.slot, vr.medium);
        vr.r = spawn_ray(isect.p, isect.p_error, isect.n, vr.r.d, isect.time);
        vr.medium = medium_for(mi, isect.n, vr.r.d);
    }
}

// What estimate_direct needs to know about the scattering point: a surface with its BSDF, or a point in a medium with its phase function
struct VolVertex {
    bool surface = true;
    InteractionData ref;      // p, p_error, n (n = 0 in a medium), time
    V3 wo, sh_n;
    const BSDF* bsdf = nullptr;
    Float g = 0;              // HenyeyGreenstein
    VolInterface mi;
};

// estimate_direct with handle_media = true, integrator.rs:109-237
inline Spectrum vol_estimate_direct(const RenderScene& s, const VolVertex& it, P2 uscatt, int li, P2 ulight, RenderCounters& rc) {
    const pbrt_b200_light& light = s.d.lights[li];
    const int flags = BSDF_ALL & ~BSDF_SPECULAR;
    Spectrum Ld(0.0f);
    LightSample ls = light_sample_li(s, li, it.ref, ulight);
    Float lightpdf = ls.pdf, scattpdf = 0.0f;
    V3 wi = ls.wi;
    Spectrum Li = ls.Li;
    if (lightpdf > 0.0f && !Li.is_black()) {
        Spectrum f;
        if (it.surface) { f = it.bsdf->f(it.wo, wi, flags) * abs_dot(wi, it.sh_n); scattpdf = it.bsdf->pdf(it.wo, wi, flags); }
        else { Float p = phase_hg(dot(it.wo, wi), it.g); f = Spectrum(p); scattpdf = p; }
        if (!f.is_black()) {
            Li *= visibility_tr(s, it.ref, it.mi, ls.p1, rc);
            if (!Li.is_black()) {
                if (is_delta_light(light)) Ld += f * Li / lightpdf;
                else { Float weight = power_heuristic(1, lightpdf, 1, scattpdf); Ld += f * Li * weight / lightpdf; }
            }
        }
    }
    if (!is_delta_light(light)) {
        Spectrum f;
        bool sampled_specular = false;
        if (it.surface) {
            int sampled_type = 0;
            f = it.bsdf->sample_f(it.wo, &wi, uscatt, &scattpdf, flags, &sampled_type);
            f = f * abs_dot(wi, it.sh_n);
            sampled_specular = (sampled_type & BSDF_SPECULAR) != 0;
        } else {
            Float p = hg_sample_p(it.g, it.wo, &wi, uscatt);
            f = Spectrum(p); scattpdf = p;
        }
        if (!f.is_black() && scattpdf > 0.0f) {
            Float weight = 1.0f;
            if (!sampled_specular) {
                lightpdf = light_pdf_li(s, li, it.ref, wi);
                if (lightpdf == 0.0f) return Ld;
                weight = power_heuristic(1, scattpdf, 1, lightpdf);
            }
            VolRay vr;
            vr.r = spawn_ray(it.ref.p, it.ref.p_error, it.ref.n, wi, it.ref.time);
            vr.medium = medium_for(it.mi, it.ref.n, wi);
            Hit h; Ray r_hit; Spectrum Tr;
            bool found = intersect_tr(s, vr, &h, &r_hit, &Tr, rc);
            Spectrum li_(0.0f);
            if (found) {
                if (s.d.prims[h.slot].area_light == li) {
                    SurfaceInteraction lsi = make_interaction(s, r_hit, h);
                    li_ = surface_le(s, lsi, -wi);
                }
            } else li_ = light_le(s, li);
            if (!li_.is_black()) Ld += f * li_ * Tr * weight / scattpdf;
        }
    }
    return Ld;
}

// uniform_sample_onelight(handle_media = true), integrator.rs:81-106
inline Spectrum vol_uniform_sample_onelight(const RenderScene& s, const VolVertex& it, Sampler& sampler, const Distribution1D& distrib, RenderCounters& rc) {
    size_t nlights = s.d.n_lights;
    if (nlights == 0) return Spectrum(0.0f);
    Float lightpdf = 0.0f;
    size_t lightnum = distrib.sample_discrete(sampler.get_1d(), &lightpdf);
    if (lightpdf == 0.0f) return Spectrum(0.0f);
    P2 ulight = sampler.get_2d();
    P2 uscattering = sampler.get_2d();
    return vol_estimate_direct(s, it, uscattering, (int)lightnum, ulight, rc) / lightpdf;
}

// VolPathIntegrator::li, volpath.rs:82-221 (no BSSRDF: none of the five materials has one)
inline Spectrum volpath_li(const RenderScene& s, const IntegratorParams& ip, Ray ray0, int camera_medium, Sampler& sampler, RenderCounters& rc) {
    Spectrum L(0.0f), beta(1.0f);
    VolRay ray; ray.r = ray0; ray.medium = camera_medium;
    bool specular_bounce = false;
    uint64_t bounces = 0;  // usize: `bounces -= 1` at 0 wraps in a release build (see the header note)
    const uint64_t max_depth = (uint64_t)std::max(ip.max_depth, 0);
    Float etascale = 1.0f;
    for (;;) {
        Hit h;
        Ray r0 = ray.r;
        rc.intersection_tests++;
        bool found = scene_intersect(s, ray.r, &h, &rc.trav_closest);
        SurfaceInteraction isect;
        if (found) isect = make_interaction(s, r0, h);
        // HomogeneousMedium::sample, homogeneous.rs:35-72
        bool mi_valid = false;
        V3 mi_p;
        if (ray.medium >= 0) {
            const pbrt_b200_medium& m = s.d.media[ray.medium];
            Spectrum sigma_t = medium_sigma_t(m), sigma_s = spec3(m.sigma_s);
            int channel = (int)std::min<uint64_t>(f2u_sat(sampler.get_1d() * 3.0f), 2);
            Float dist = -std::log(1.0f - sampler.get_1d()) / sigma_t.c[channel];
            Float dlen = length(ray.r.d);
            Float t = std::fmin(dist / dlen, ray.r.t_max);
            bool sampled_medium = t < ray.r.t_max;
            if (sampled_medium) { mi_valid = true; mi_p = ray.r.o + ray.r.d * t; }
            Spectrum Tr = spec_exp(spec_neg(sigma_t) * std::fmin(t, std::numeric_limits<Float>::max()) * dlen);
            Spectrum density = sampled_medium ? sigma_t * Tr : Tr;
            Float pdf = 0.0f;
            for (int i = 0; i < 3; ++i) pdf += density.c[i];
            pdf *= 1.0f / 3.0f;
            if (pdf == 0.0f) pdf = 1.0f;
            beta *= sampled_medium ? Tr * sigma_s / pdf : Tr / pdf;
        }
        if (beta.is_black()) break;
        if (mi_valid) {
            if (bounces >= max_depth) break;
            const Distribution1D& distrib = ip.lookup(mi_p);
            VolVertex v;
            v.surface = false;
            v.ref.p = mi_p; v.ref.p_error = V3(0, 0, 0); v.ref.n = V3(0, 0, 0); v.ref.time = ray.r.time;
            v.wo = -ray.r.d; v.g = s.d.media[ray.medium].g;
            v.mi.inside = v.mi.outside = ray.medium;  // MediumInterface::new(medium)
            Spectrum Ld = beta * vol_uniform_sample_onelight(s, v, sampler, distrib, rc);
            if (Ld.is_black()) rc.zero_radiance++;
            L += Ld;
            V3 wi;
            hg_sample_p(v.g, v.wo, &wi, sampler.get_2d());
            int med = ray.medium;
            ray.r = spawn_ray(mi_p, V3(0, 0, 0), V3(0, 0, 0), wi, ray.r.time);
            ray.medium = med;
            specular_bounce = false;
        } else {
            if (bounces == 0 || specular_bounce) {
                if (found) L += surface_le(s, isect, -ray.r.d) * beta;
                else for (int li : s.infinite_lights) L += light_le(s, li) * beta;
            }
            if (!found || bounces >= max_depth) break;
            BSDF bsdf;
            int mat = s.d.prims[h.slot].material;
            (void)mat;
            scattering_functions(s, ray.r, h.slot, isect, &bsdf, true);
            VolInterface mif = hit_interface(s, h.slot, ray.medium);
            if (!bsdf.valid) {  // volpath.rs:131-135
                V3 d = ray.r.d;
                ray.r = spawn_ray(isect.p, isect.p_error, isect.n, d, isect.time);
                ray.medium = medium_for(mif, isect.n, d);
                bounces -= 1;   // wraps at 0, and the `continue` skips the increment below
                continue;
            }
            const Distribution1D& distrib = ip.lookup(isect.p);
            VolVertex v;
            v.surface = true;
            v.ref.p = isect.p; v.ref.p_error = isect.p_error; v.ref.n = isect.n; v.ref.time = isect.time;
            v.wo = isect.wo; v.sh_n = isect.sh_n; v.bsdf = &bsdf; v.mi = mif;
            Spectrum Ld = beta * vol_uniform_sample_onelight(s, v, sampler, distrib, rc);
            if (Ld.is_black()) rc.zero_radiance++;
            L += Ld;
            V3 wo = -ray.r.d, wi;
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
            ray.r = spawn_ray(isect.p, isect.p_error, isect.n, wi, isect.time);
            ray.medium = medium_for(mif, isect.n, wi);
        }
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

}  // namespace orc
