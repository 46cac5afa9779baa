// Device shading: hit-frame reconstruction, the lobes reachable from
// matte/plastic/mirror/glass/metal, lobe selection/evaluation (BSDF), light sampling.
// Reference arithmetic cited per function (paths relative to pbrt-rust/src).  Shading is
// tolerance-parity only (libm differs from CUDA's), the op order still follows the reference.
#pragma once
#include "scene.cuh"
#include "trace.cuh"
#include "vecmath.cuh"

namespace pb {

// RGBSpectrum, core/spectrum.rs:78-168
struct rgb {
    float r, g, b;
    PB_D rgb() {}
    PB_D explicit rgb(float v) : r(v), g(v), b(v) {}
    PB_D rgb(float r_, float g_, float b_) : r(r_), g(g_), b(b_) {}
};
PB_D rgb operator+(rgb a, rgb b) { return rgb(a.r + b.r, a.g + b.g, a.b + b.b); }
PB_D rgb operator-(rgb a, rgb b) { return rgb(a.r - b.r, a.g - b.g, a.b - b.b); }
PB_D rgb operator*(rgb a, rgb b) { return rgb(a.r * b.r, a.g * b.g, a.b * b.b); }
PB_D rgb operator/(rgb a, rgb b) { return rgb(a.r / b.r, a.g / b.g, a.b / b.b); }
PB_D rgb operator*(rgb a, float s) { return rgb(a.r * s, a.g * s, a.b * s); }
PB_D rgb operator/(rgb a, float s) { return rgb(a.r / s, a.g / s, a.b / s); }
PB_D rgb operator+(rgb a, float s) { return rgb(a.r + s, a.g + s, a.b + s); }
PB_D rgb operator-(rgb a, float s) { return rgb(a.r - s, a.g - s, a.b - s); }
PB_D bool is_black(rgb a) { return a.r == 0.0f && a.g == 0.0f && a.b == 0.0f; }
PB_D float lum(rgb a) { return 0.212671f * a.r + 0.715160f * a.g + 0.072169f * a.b; }  // spectrum.rs:123-127
PB_D float max_comp(rgb a) { return fmaxf(fmaxf(a.r, a.g), a.b); }
PB_D rgb rgb_sqrt(rgb a) { return rgb(sqrtf(a.r), sqrtf(a.g), sqrtf(a.b)); }
PB_D rgb rgb_clamp0(rgb a) {  // clamps(0, INFINITY)
    return rgb(clampf(a.r, 0.0f, PB_INF), clampf(a.g, 0.0f, PB_INF), clampf(a.b, 0.0f, PB_INF));
}
PB_D rgb rgb3(const float* p) { return rgb(p[0], p[1], p[2]); }

// ---- surface interaction (core/interaction.rs:149-249), only what path shading reads
struct Surf {
    f3 p, p_error, n, wo;
    f3 sh_n, sh_dpdu;
};

// Triangle::intersect tail (shapes/triangle.rs:236-392) for the accepted hit.
// shape_some mirrors the `s: Option<Arc<Shapes>>` argument (None from Shape::pdf_wi).
// slot: the primitive's BVH slot when known (uvs pre-gathered, scene.cuh); nrm3: its three normals pre-gathered ({n.xyz, -} x 3:
// DevScene::slot_n or ::light_tris) or nullptr to gather them through the index buffer.
PB_D Surf triangle_surface(const DevScene& s, f3 p0, f3 p1, f3 p2, uint32_t flags, uint32_t shape_index, f3 ray_d, float b0, float b1, float b2,
                           bool shape_some, uint32_t slot = PB_NO_SLOT, const float4* nrm3 = nullptr) {
    float2 uv0, uv1, uv2;
    fetch_uv(s, flags, shape_index, &uv0, &uv1, &uv2, slot);
    float2 duv02 = make_float2(uv0.x - uv2.x, uv0.y - uv2.y), duv12 = make_float2(uv1.x - uv2.x, uv1.y - uv2.y);
    f3 dp02 = p0 - p2, dp12 = p1 - p2;
    float determinant = duv02.x * duv12.y - duv02.y * duv12.x;
    bool degenerateuv = fabsf(determinant) < 1.0e-8f;
    f3 dpdu(0.f, 0.f, 0.f), dpdv(0.f, 0.f, 0.f);
    if (!degenerateuv) {
        float inv = 1.0f / determinant;
        dpdu = (dp02 * duv12.y - dp12 * duv02.y) * inv;
        dpdv = (dp02 * -duv12.x + dp12 * duv02.x) * inv;
    }
    if (degenerateuv || len2(cross(dpdu, dpdv)) == 0.0f) {
        f3 ng = cross(p2 - p0, p1 - p0);
        coordinate_system(normalize(ng), &dpdu, &dpdv);
    }
    Surf si;
    float xs = fabsf(b0 * p0.x) + fabsf(b1 * p1.x) + fabsf(b2 * p2.x);
    float ys = fabsf(b0 * p0.y) + fabsf(b1 * p1.y) + fabsf(b2 * p2.y);
    float zs = fabsf(b0 * p0.z) + fabsf(b1 * p1.z) + fabsf(b2 * p2.z);
    si.p_error = f3(xs, ys, zs) * gamma_n(7);
    si.p = p0 * b0 + p1 * b1 + p2 * b2;
    bool ro = flags & PBRT_B200_PRIM_REVERSE_ORIENTATION, sh = flags & PBRT_B200_PRIM_SWAPS_HANDEDNESS;
    bool flip = ro != sh;
    f3 nn = normalize(cross(dp02, dp12));
    si.n = flip ? -nn : nn;
    si.sh_n = si.n;
    si.sh_dpdu = dpdu;
    si.wo = -ray_d;  // triangle.rs:296: not normalised
    bool has_n = (flags & PBRT_B200_PRIM_HAS_N) && s.vertex_n, has_s = (flags & PBRT_B200_PRIM_HAS_S) && s.vertex_s;
    if (has_n || has_s) {
        uint32_t i0 = 0, i1 = 0, i2 = 0;
        if (has_s || (has_n && !nrm3)) {
            const uint32_t* idx = s.tri_indices + 3ull * shape_index;
            i0 = idx[0]; i1 = idx[1]; i2 = idx[2];
        }
        f3 ns;
        if (has_n && nrm3) {
            const float4 n0 = __ldg(nrm3), n1 = __ldg(nrm3 + 1), n2 = __ldg(nrm3 + 2);
            ns = f3(n0.x, n0.y, n0.z) * b0 + f3(n1.x, n1.y, n1.z) * b1 + f3(n2.x, n2.y, n2.z) * b2;
            ns = (len2(ns) > 0.0f) ? normalize(ns) : si.n;
        } else if (has_n) {
            const float* N = s.vertex_n;
            ns = f3(N[3 * i0], N[3 * i0 + 1], N[3 * i0 + 2]) * b0 + f3(N[3 * i1], N[3 * i1 + 1], N[3 * i1 + 2]) * b1 +
                 f3(N[3 * i2], N[3 * i2 + 1], N[3 * i2 + 2]) * b2;
            ns = (len2(ns) > 0.0f) ? normalize(ns) : si.n;
        } else ns = si.n;
        f3 ss;
        if (has_s) {
            const float* S = s.vertex_s;
            ss = f3(S[3 * i0], S[3 * i0 + 1], S[3 * i0 + 2]) * b0 + f3(S[3 * i1], S[3 * i1 + 1], S[3 * i1 + 2]) * b1 +
                 f3(S[3 * i2], S[3 * i2 + 1], S[3 * i2 + 2]) * b2;
            ss = (len2(ss) > 0.0f) ? normalize(ss) : normalize(dpdu);
        } else ss = normalize(dpdu);
        f3 ts = cross(ss, ns);
        if (len2(ts) > 0.0f) { ts = normalize(ts); ss = cross(ts, ns); }
        else coordinate_system(ns, &ss, &ts);
        if (ro) ts = -ts;
        // set_shading_geometry(.., orientation_is_authoritative = true), interaction.rs:234-255
        f3 shn = normalize(cross(ss, ts));
        if (shape_some) {
            if (flip) shn = -shn;
            si.n = face_forward(si.n, shn);
        }
        si.sh_n = shn;
        si.sh_dpdu = ss;
    }
    return si;
}

// Sphere::intersect tail (shapes/sphere.rs:100-196, full sphere) + Transform::transform_surface_interaction
// (core/transform.rs:607-636).  The shape is passed as None there => no orientation flip.
static __device__ __noinline__ Surf sphere_surface(const pbrt_b200_sphere* spp, f3 ray_o, f3 ray_d, float t) {
    const pbrt_b200_sphere& sp = *spp;
    f3 oo, od, oe, de;
    sphere_object_ray(sp, ray_o, ray_d, &oo, &od, &oe, &de);
    const float phi_max = (PB_PI / 180.0f) * 360.0f;  // radians(360)
    const float theta_min = acosf(-1.0f), theta_max = acosf(1.0f);
    f3 ph = oo + od * t;
    ph = ph * (sp.radius / len(ph));  // p_hit.distance(origin)
    if (ph.x == 0.0f && ph.y == 0.0f) ph.x = 1.0e-5f * sp.radius;
    float theta = acosf(clampf(ph.z / sp.radius, -1.0f, 1.0f));
    float zradius = sqrtf(ph.x * ph.x + ph.y * ph.y);
    float inv_radius = 1.0f / zradius;
    float cos_phi = ph.x * inv_radius, sin_phi = ph.y * inv_radius;
    f3 dpdu(-phi_max * ph.y, phi_max * ph.x, 0.0f);
    f3 dpdv = f3(ph.z * cos_phi, ph.z * sin_phi, -sp.radius * sinf(theta)) * (theta_max - theta_min);
    f3 pe = vabs(ph) * gamma_n(5);
    f3 n = normalize(cross(dpdu, dpdv));  // SurfaceInteraction::new, shape None
    f3 wo = normalize(-od);
    Surf r;
    r.p = xf_point_abs_err(sp.object_to_world, ph, pe, &r.p_error);
    r.n = normalize(xf_normal(sp.world_to_object, n));
    r.wo = normalize(xf_vector(sp.object_to_world, wo));
    r.sh_n = normalize(xf_normal(sp.world_to_object, n));
    r.sh_dpdu = xf_vector(sp.object_to_world, dpdu);
    r.sh_n = face_forward(r.sh_n, r.n);
    return r;
}

// Surface at a closest-hit record.
PB_D Surf surface_at(const DevScene& s, uint32_t slot, f3 ray_o, f3 ray_d, float t, float b0, float b1, float b2, uint32_t* flags_out) {
    const float4* tp = s.tris + 3ull * slot;
    float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
    uint32_t fl = __float_as_uint(v1.w);
    *flags_out = fl;
    if (fl & PB_TRI_SPHERE) return sphere_surface(s.spheres + __float_as_uint(v2.w), ray_o, ray_d, t);
    return triangle_surface(s, f3(v0.x, v0.y, v0.z), f3(v1.x, v1.y, v1.z), f3(v2.x, v2.y, v2.z), fl, __float_as_uint(v2.w), ray_d, b0, b1, b2, true, slot,
                            s.slot_n ? s.slot_n + 3ull * slot : nullptr);
}

// Hit inside an instanced object: TransformedPrimitive::intersect (primitive.rs:58-80) -- the object's interaction is
// computed with the ray in object space, then Transform::transform_surface_interaction (transform.rs:607-636) takes it
// to world space unless prim_to_world is the identity.
static __device__ __noinline__ Surf surface_at_instance(const DevScene* sp, uint32_t inst, uint32_t slot, f3 ray_o, f3 ray_d, float t, float b0, float b1, float b2,
                                                        uint32_t* flags_out) {
    const DevScene& s = *sp;
    const DevInstance& in = s.instances[inst];
    f3 o2, d2;
    float tm2;
    xf_ray(in.world_to_prim, ray_o, ray_d, PB_INF, &o2, &d2, &tm2);
    Surf si = surface_at(s, slot, o2, d2, t, b0, b1, b2, flags_out);
    if (in.flags & PB_INST_IDENTITY) return si;
    Surf r;
    r.p = xf_point_abs_err(in.prim_to_world, si.p, si.p_error, &r.p_error);
    r.n = normalize(xf_normal(in.world_to_prim, si.n));
    r.wo = normalize(xf_vector(in.prim_to_world, si.wo));
    r.sh_n = normalize(xf_normal(in.world_to_prim, si.sh_n));
    r.sh_dpdu = xf_vector(in.prim_to_world, si.sh_dpdu);
    r.sh_n = face_forward(r.sh_n, r.n);
    return r;
}
template <bool INST>
PB_D Surf surface_at_hit(const DevScene& s, uint32_t inst, uint32_t slot, f3 ray_o, f3 ray_d, float t, float b0, float b1, float b2, uint32_t* flags_out) {
    if (INST && inst != PBRT_B200_NO_HIT) return surface_at_instance(s.self_dev, inst, slot, ray_o, ray_d, t, b0, b1, b2, flags_out);
    return surface_at(s, slot, ray_o, ray_d, t, b0, b1, b2, flags_out);
}

// ---- BxDF local-frame helpers, core/reflection.rs:78-176
PB_D float cos2_theta(f3 w) { return w.z * w.z; }
PB_D float sin2_theta(f3 w) { return fmaxf(1.0f - cos2_theta(w), 0.0f); }
PB_D float sin_theta(f3 w) { return sqrtf(sin2_theta(w)); }
PB_D float tan_theta(f3 w) { return sin_theta(w) / w.z; }
PB_D float tan2_theta(f3 w) { return sin2_theta(w) / cos2_theta(w); }
PB_D float cos_phi(f3 w) { float st = sin_theta(w); return st == 0.0f ? 1.0f : clampf(w.x / st, -1.0f, 1.0f); }
PB_D float sin_phi(f3 w) { float st = sin_theta(w); return st == 0.0f ? 0.0f : clampf(w.y / st, -1.0f, 1.0f); }
PB_D bool same_hemi(f3 a, f3 b) { return a.z * b.z > 0.0f; }
PB_D f3 reflect_about(f3 wo, f3 n) { return -wo + n * 2.0f * dot(wo, n); }
PB_D bool refract_dir(f3 wi, f3 n, float eta, f3* wt) {  // reflection.rs:160-174
    float ci = dot(n, wi);
    float s2i = fmaxf(1.0f - ci * ci, 0.0f);
    float s2t = eta * eta * s2i;
    if (s2t >= 1.0f) return false;
    float ct = sqrtf(1.0f - s2t);
    *wt = n * (eta * ci - ct) + (-wi) * eta;
    return true;
}
// reflection.rs:29-52
PB_D float fr_dielectric(float ci, float etai, float etat) {
    ci = clampf(ci, -1.0f, 1.0f);
    if (!(ci > 0.0f)) { float tmp = etai; etai = etat; etat = tmp; ci = fabsf(ci); }
    float si = sqrtf(fmaxf(1.0f - ci * ci, 0.0f));
    float st = etai / etat * si;
    if (st >= 1.0f) return 1.0f;
    float ct = sqrtf(fmaxf(1.0f - st * st, 0.0f));
    float rparl = ((etat * ci) - (etai * ct)) / ((etat * ci) + (etai * ct));
    float rperp = ((etai * ci) - (etat * ct)) / ((etai * ci) + (etat * ct));
    return (rparl * rparl + rperp * rperp) / 2.0f;
}
// reflection.rs:54-76 with etai = 1
PB_D rgb fr_conductor(float ci, rgb etat, rgb k) {
    ci = clampf(ci, -1.0f, 1.0f);
    rgb eta = etat / rgb(1.0f), etak = k / rgb(1.0f);
    float c2 = ci * ci, s2 = 1.0f - c2;
    rgb eta2 = eta * eta, etak2 = etak * etak;
    rgb t0 = eta2 - etak2 - rgb(s2);
    rgb a2b2 = rgb_sqrt(t0 * t0 + eta2 * etak2 * 4.0f);
    rgb t1 = a2b2 + rgb(c2);
    rgb a = rgb_sqrt((a2b2 + t0) * 0.5f);
    rgb t2 = a * ci * 2.0f;
    rgb Rs = (t1 - t2) / (t1 + t2);
    rgb t3 = a2b2 * c2 + rgb(s2 * s2);
    rgb t4 = t2 * s2;
    rgb Rp = Rs * (t3 - t4) / (t3 + t4);
    return (Rp + Rs) * 0.5f;
}

// ---- sampling, core/sampling.rs
PB_D float2 concentric_disk(float2 u) {  // :154-176
    float ox = u.x * 2.0f - 1.0f, oy = u.y * 2.0f - 1.0f;
    if (ox == 0.0f && oy == 0.0f) return make_float2(0.f, 0.f);
    float th, r;
    if (fabsf(ox) > fabsf(oy)) { r = ox; th = PB_PI_OVER4 * (oy / ox); }
    else { r = oy; th = PB_PI_OVER2 - PB_PI_OVER4 * (ox / oy); }
    float sn, cs;
    sincosf(th, &sn, &cs);  // one argument reduction for both (th is in [-pi/4, 3pi/4])
    return make_float2(cs * r, sn * r);
}
PB_D f3 cosine_hemisphere(float2 u) {  // :188-193
    float2 d = concentric_disk(u);
    return f3(d.x, d.y, sqrtf(fmaxf(0.0f, 1.0f - d.x * d.x - d.y * d.y)));
}
PB_D float power_heuristic(float fpdf, float gpdf) { return (fpdf * fpdf) / (fpdf * fpdf + gpdf * gpdf); }  // :328-333, nf = ng = 1

// ---- Trowbridge-Reitz, core/microfacet.rs:249-406 (samplevis = true)
PB_D float roughness_to_alpha(float rough) {
    rough = fmaxf(rough, 1.0e-3f);
    float x = logf(rough);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
struct TRDist { float ax, ay; };
PB_D TRDist tr_make(float ax, float ay) { TRDist d; d.ax = fmaxf(ax, 0.001f); d.ay = fmaxf(ay, 0.001f); return d; }
PB_D float tr_d(TRDist t, f3 wh) {
    float t2 = tan2_theta(wh);
    if (isinf(t2)) return 0.0f;
    float c4 = cos2_theta(wh) * cos2_theta(wh);
    float cp = cos_phi(wh), sp = sin_phi(wh);
    float e = ((cp * cp) / (t.ax * t.ax) + (sp * sp) / (t.ay * t.ay)) * t2;
    return 1.0f / (PB_PI * t.ax * t.ay * c4 * (1.0f + e) * (1.0f + e));
}
PB_D float tr_lambda(TRDist t, f3 w) {
    float att = fabsf(tan_theta(w));
    if (isinf(att)) return 0.0f;
    float cp = cos_phi(w), sp = sin_phi(w);
    float alpha = sqrtf((cp * cp) * t.ax * t.ax + (sp * sp) * t.ay * t.ay);
    float a2t2 = (alpha * att) * (alpha * att);
    return (-1.0f + sqrtf(1.0f + a2t2)) / 2.0f;
}
PB_D float tr_g1(TRDist t, f3 w) { return 1.0f / (1.0f + tr_lambda(t, w)); }
PB_D float tr_g(TRDist t, f3 wo, f3 wi) { return 1.0f / (1.0f + tr_lambda(t, wo) + tr_lambda(t, wi)); }
PB_D float tr_pdf(TRDist t, f3 wo, f3 wh) { return tr_d(t, wh) * tr_g1(t, wo) * absdot(wo, wh) / fabsf(wo.z); }
PB_D void tr_sample11(float ct, float u1, float u2, float* sx, float* sy) {
    if (ct > 0.9999f) {
        float r = sqrtf(u1 / (1.0f - u1));
        float phi = 6.28318530718f * u2;
        *sx = r * cosf(phi); *sy = r * sinf(phi);
        return;
    }
    float st = sqrtf(fmaxf(1.0f - ct * ct, 0.0f));
    float tt = st / ct;
    float a = 1.0f / tt;
    float G1 = 2.0f / (1.0f + sqrtf(1.0f + 1.0f / (a * a)));
    float A = 2.0f * u1 / G1 - 1.0f;
    float tmp = 1.0f / (A * A - 1.0f);
    if (tmp > 1.0e10f) tmp = 1.0e10f;
    float B = tt;
    float D = sqrtf(fmaxf(B * B * tmp * tmp - (A * A - B * B) * tmp, 0.0f));
    float s1 = B * tmp - D, s2 = B * tmp + D;
    *sx = (A < 0.0f || s2 > 1.0f / tt) ? s1 : s2;
    float sg;
    if (u2 > 0.5f) { sg = 1.0f; u2 = 2.0f * (u2 - 0.5f); } else { sg = -1.0f; u2 = 2.0f * (0.5f - u2); }
    float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) / (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
    *sy = sg * z * sqrtf(1.0f + *sx * *sx);
}
PB_D f3 tr_sample_wh(TRDist t, f3 wo, float2 u) {
    bool flip = wo.z < 0.0f;
    f3 wi = flip ? -wo : wo;
    f3 ws = normalize(f3(t.ax * wi.x, t.ay * wi.y, wi.z));
    float sx, sy;
    tr_sample11(ws.z, u.x, u.y, &sx, &sy);
    float cp = cos_phi(ws), sp = sin_phi(ws);
    float tmp = cp * sx - sp * sy;
    sy = sp * sx + cp * sy;
    sx = tmp;
    sx = t.ax * sx; sy = t.ay * sy;
    f3 wh = normalize(f3(-sx, -sy, 1.0f));
    return flip ? -wh : wh;
}

// ---- lobes (closed set; BxDFType bits as in reflection.rs:181-190)
enum { BX_REFLECTION = 1, BX_TRANSMISSION = 2, BX_DIFFUSE = 4, BX_GLOSSY = 8, BX_SPECULAR = 16, BX_ALL = 31 };
enum LobeKind { LOBE_LAMBERT = 0, LOBE_OREN_NAYAR = 1, LOBE_MIRROR = 2, LOBE_FRESNEL_SPECULAR = 3, LOBE_MICRO_REFL_DIEL = 4, LOBE_MICRO_REFL_COND = 5, LOBE_MICRO_TRANS = 6,
                LOBE_SPEC_REFL_DIEL = 7, LOBE_SPEC_TRANS = 8,  // glass without allow_multiple_lobes (whitted / directlighting), glass.rs:69-84; uber
                LOBE_FRESNEL_BLEND = 9 };  // substrate (reflection.rs:1141-1222): c0 = Rd, c1 = Rs

struct Lobe {
    int kind, type;
    rgb c0, c1, c2;  // R | (R,T) | (R, eta, k)
    float p0, p1;    // OrenNayar A,B | dielectric etai, etat | etaa, etab
    TRDist tr;
};

PB_D bool lobe_matches(const Lobe& l, int flags) { return (l.type & flags) == l.type; }

// KM = compile-time mask of the lobe kinds a material can produce (1 << LobeKind): the shade kernel of one
// material bin carries only that material's BxDF code (smaller kernels, fewer registers); KM_ALL = generic.
#define KM_ALL 0x1ff
#define KM_TEX 0x3ff /* the texture-parameterised materials add uber's lobes and substrate's FresnelBlend (texture.cuh) */
#define KM_HAS(k) ((KM & (1 << (k))) != 0)
enum { KM_MATTE = (1 << LOBE_LAMBERT) | (1 << LOBE_OREN_NAYAR), KM_PLASTIC = (1 << LOBE_LAMBERT) | (1 << LOBE_MICRO_REFL_DIEL), KM_MIRROR = 1 << LOBE_MIRROR,
       KM_GLASS = (1 << LOBE_FRESNEL_SPECULAR) | (1 << LOBE_MICRO_REFL_DIEL) | (1 << LOBE_MICRO_TRANS), KM_METAL = 1 << LOBE_MICRO_REFL_COND };

template <int KM = KM_ALL>
PB_D rgb lobe_f(const Lobe& l, f3 wo, f3 wi) {
    switch (l.kind) {
        case LOBE_LAMBERT: if (!KM_HAS(LOBE_LAMBERT)) break; return l.c0 * PB_INV_PI;  // reflection.rs:823-825
        case LOBE_OREN_NAYAR: {                        // reflection.rs:925-952
            if (!KM_HAS(LOBE_OREN_NAYAR)) break;
            float sti = sin_theta(wi), sto = sin_theta(wo);
            float max_cos = 0.0f;
            if (sti > 1e-4f && sto > 1e-4f) {
                float dcos = cos_phi(wi) * cos_phi(wo) + sin_phi(wi) * sin_phi(wo);
                max_cos = fmaxf(dcos, 0.0f);
            }
            float sa, tb;
            if (fabsf(wi.z) > fabsf(wo.z)) { sa = sto; tb = sti / fabsf(wi.z); } else { sa = sti; tb = sto / fabsf(wo.z); }
            return l.c0 * PB_INV_PI * (l.p0 + l.p1 * max_cos * sa * tb);
        }
        case LOBE_MIRROR: case LOBE_SPEC_REFL_DIEL: case LOBE_SPEC_TRANS: return rgb(0.0f);  // reflection.rs:630-632, 682-684
        case LOBE_FRESNEL_SPECULAR: return rgb(1.0f);  // reflection.rs:745-747 (reference quirk)
        case LOBE_MICRO_REFL_DIEL:
        case LOBE_MICRO_REFL_COND: {  // reflection.rs:985-1003
            if (!KM_HAS(LOBE_MICRO_REFL_DIEL) && !KM_HAS(LOBE_MICRO_REFL_COND)) break;
            float cto = fabsf(wo.z), cti = fabsf(wi.z);
            f3 wh = wi + wo;
            if (cti == 0.0f || cto == 0.0f) return rgb(0.0f);
            if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return rgb(0.0f);
            wh = normalize(wh);
            float cih = dot(wi, wh);
            const bool diel = !KM_HAS(LOBE_MICRO_REFL_COND) || (KM_HAS(LOBE_MICRO_REFL_DIEL) && l.kind == LOBE_MICRO_REFL_DIEL);
            rgb F = diel ? rgb(fr_dielectric(cih, l.p0, l.p1)) : fr_conductor(fabsf(cih), l.c1, l.c2);
            float d = tr_d(l.tr, wh), g = tr_g(l.tr, wo, wi);
            return l.c0 * d * g * F / (4.0f * cti * cto);
        }
        case LOBE_MICRO_TRANS: {  // reflection.rs:1064-1095
            if (!KM_HAS(LOBE_MICRO_TRANS)) break;
            if (same_hemi(wo, wi)) return rgb(0.0f);
            float cto = wo.z, cti = wi.z;
            if (cti == 0.0f || cto == 0.0f) return rgb(0.0f);
            float eta = wo.z > 0.0f ? l.p1 / l.p0 : l.p0 / l.p1;
            f3 wh = normalize(wo + wi * eta);
            if (wh.z < 0.0f) wh = -wh;
            if (dot(wo, wh) * dot(wi, wh) > 0.0f) return rgb(0.0f);
            rgb F(fr_dielectric(dot(wo, wh), l.p0, l.p1));
            float sd = dot(wo, wh) + eta * dot(wi, wh);
            float factor = 1.0f / eta;
            return (rgb(1.0f) - F) * l.c0 *
                   fabsf(tr_d(l.tr, wh) * tr_g(l.tr, wo, wi) * eta * eta * absdot(wi, wh) * absdot(wo, wh) * factor * factor / (cti * cto * sd * sd));
        }
        case LOBE_FRESNEL_BLEND: {  // reflection.rs:1167-1185
            if (!KM_HAS(LOBE_FRESNEL_BLEND)) break;
            float a = 1.0f - 0.5f * fabsf(wi.z), b = 1.0f - 0.5f * fabsf(wo.z);
            rgb diffuse = l.c0 * (rgb(1.0f) - l.c1) * (28.0f / (23.0f * PB_PI)) * (1.0f - (a * a) * (a * a) * a) * (1.0f - (b * b) * (b * b) * b);
            f3 wh = wi + wo;
            if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return rgb(0.0f);
            wh = normalize(wh);
            float c = 1.0f - dot(wi, wh);
            rgb schlick = l.c1 + (rgb(1.0f) - l.c1) * ((c * c) * (c * c) * c);
            return diffuse + schlick * (tr_d(l.tr, wh) / (4.0f * absdot(wi, wh) * fmaxf(fabsf(wi.z), fabsf(wo.z))));
        }
    }
    return rgb(0.0f);
}

template <int KM = KM_ALL>
PB_D float lobe_pdf(const Lobe& l, f3 wo, f3 wi) {
    switch (l.kind) {
        case LOBE_LAMBERT: case LOBE_OREN_NAYAR: case LOBE_FRESNEL_SPECULAR:
            return same_hemi(wo, wi) ? fabsf(wi.z) * PB_INV_PI : 0.0f;  // reflection.rs:438-445, 788-794
        case LOBE_MIRROR: case LOBE_SPEC_REFL_DIEL: case LOBE_SPEC_TRANS: return 0.0f;
        case LOBE_MICRO_REFL_DIEL: case LOBE_MICRO_REFL_COND: {  // reflection.rs:1021-1027
            if (!KM_HAS(LOBE_MICRO_REFL_DIEL) && !KM_HAS(LOBE_MICRO_REFL_COND)) break;
            if (!same_hemi(wo, wi)) return 0.0f;
            f3 wh = normalize(wo + wi);
            return tr_pdf(l.tr, wo, wh) / (4.0f * dot(wo, wh));
        }
        case LOBE_MICRO_TRANS: {  // reflection.rs:1115-1129
            if (!KM_HAS(LOBE_MICRO_TRANS)) break;
            if (same_hemi(wo, wi)) return 0.0f;
            float eta = wo.z > 0.0f ? l.p0 / l.p1 : l.p1 / l.p0;
            f3 wh = normalize(wo + wi * eta);
            if (dot(wo, wh) * dot(wi, wh) > 0.0f) return 0.0f;
            float sd = dot(wo, wh) + eta * dot(wi, wh);
            float dwh = fabsf(eta * eta * dot(wi, wh)) / (sd * sd);
            return tr_pdf(l.tr, wo, wh) * dwh;
        }
        case LOBE_FRESNEL_BLEND: {  // reflection.rs:1212-1218
            if (!KM_HAS(LOBE_FRESNEL_BLEND)) break;
            if (!same_hemi(wo, wi)) return 0.0f;
            f3 wh = normalize(wo + wi);
            return 0.5f * (fabsf(wi.z) * PB_INV_PI + tr_pdf(l.tr, wo, wh) / (4.0f * dot(wo, wh)));
        }
    }
    return 0.0f;
}

// BxDF::sample_f; *pdf and *stype keep their incoming values on early returns, like the
// reference's &mut parameters.
template <int KM = KM_ALL>
PB_D rgb lobe_sample(const Lobe& l, f3 wo, f3* wi, float2 u, float* pdf, int* stype) {
    switch (l.kind) {
        case LOBE_LAMBERT: case LOBE_OREN_NAYAR: {  // reflection.rs:392-405
            if (!KM_HAS(LOBE_LAMBERT) && !KM_HAS(LOBE_OREN_NAYAR)) break;
            *wi = cosine_hemisphere(u);
            if (wo.z < 0.0f) wi->z *= -1.0f;
            *pdf = lobe_pdf<KM>(l, wo, *wi);
            return lobe_f<KM>(l, wo, *wi);
        }
        case LOBE_MIRROR: {  // reflection.rs:634-640, FresnelNoOp
            if (!KM_HAS(LOBE_MIRROR)) break;
            *wi = f3(-wo.x, -wo.y, wo.z);
            *pdf = 1.0f;
            return rgb(1.0f) * l.c0 / fabsf(wi->z);
        }
        case LOBE_SPEC_REFL_DIEL: {  // reflection.rs:634-640 with FresnelDielectric(1, eta)
            if (!KM_HAS(LOBE_SPEC_REFL_DIEL)) break;
            *wi = f3(-wo.x, -wo.y, wo.z);
            *pdf = 1.0f;
            return rgb(fr_dielectric(wi->z, l.p0, l.p1)) * l.c0 / fabsf(wi->z);
        }
        case LOBE_SPEC_TRANS: {  // reflection.rs:686-708, TransportMode::Radiance
            if (!KM_HAS(LOBE_SPEC_TRANS)) break;
            float etai = wo.z > 0.0f ? l.p0 : l.p1, etat = wo.z > 0.0f ? l.p1 : l.p0;
            f3 nf = (wo.z < 0.0f) ? f3(-0.0f, -0.0f, -1.0f) : f3(0.0f, 0.0f, 1.0f);  // face_foward_vec
            if (!refract_dir(wo, nf, etai / etat, wi)) return rgb(0.0f);
            *pdf = 1.0f;
            rgb ft = l.c0 * (rgb(1.0f) - rgb(fr_dielectric(wi->z, l.p0, l.p1)));
            ft = ft * ((etai * etai) / (etat * etat));
            return ft / fabsf(wi->z);
        }
        case LOBE_FRESNEL_SPECULAR: {  // reflection.rs:749-786
            if (!KM_HAS(LOBE_FRESNEL_SPECULAR)) break;
            float F = fr_dielectric(wo.z, l.p0, l.p1);
            if (u.x < F) {
                *wi = f3(-wo.x, -wo.y, wo.z);
                *stype = BX_SPECULAR | BX_REFLECTION;
                *pdf = F;
                return l.c0 / fabsf(wi->z) * F;
            }
            float etai = wo.z > 0.0f ? l.p0 : l.p1, etat = wo.z > 0.0f ? l.p1 : l.p0;
            f3 nf = (wo.z < 0.0f) ? f3(-0.0f, -0.0f, -1.0f) : f3(0.0f, 0.0f, 1.0f);  // face_foward_vec
            if (!refract_dir(wo, nf, etai / etat, wi)) return rgb(0.0f);
            rgb ft = l.c1 * (1.0f - F);
            ft = ft * ((etai * etai) / (etat * etat));  // TransportMode::Radiance
            *stype = BX_SPECULAR | BX_TRANSMISSION;
            *pdf = 1.0f - F;
            return ft / fabsf(wi->z);
        }
        case LOBE_MICRO_REFL_DIEL: case LOBE_MICRO_REFL_COND: {  // reflection.rs:1005-1019
            if (!KM_HAS(LOBE_MICRO_REFL_DIEL) && !KM_HAS(LOBE_MICRO_REFL_COND)) break;
            if (wo.z == 0.0f) return rgb(0.0f);
            f3 wh = tr_sample_wh(l.tr, wo, u);
            if (dot(wo, wh) < 0.0f) return rgb(0.0f);
            *wi = reflect_about(wo, wh);
            if (!same_hemi(wo, *wi)) return rgb(0.0f);
            *pdf = tr_pdf(l.tr, wo, wh) / (4.0f * dot(wo, wh));
            return lobe_f<KM>(l, wo, *wi);
        }
        case LOBE_MICRO_TRANS: {  // reflection.rs:1097-1113
            if (!KM_HAS(LOBE_MICRO_TRANS)) break;
            if (wo.z == 0.0f) return rgb(0.0f);
            f3 wh = tr_sample_wh(l.tr, wo, u);
            if (dot(wo, wh) < 0.0f) return rgb(0.0f);
            float eta = wo.z > 0.0f ? l.p0 / l.p1 : l.p1 / l.p0;
            if (!refract_dir(wo, wh, eta, wi)) return rgb(0.0f);
            *pdf = lobe_pdf<KM>(l, wo, *wi);
            return lobe_f<KM>(l, wo, *wi);
        }
        case LOBE_FRESNEL_BLEND: {  // reflection.rs:1187-1210
            if (!KM_HAS(LOBE_FRESNEL_BLEND)) break;
            if (u.x < 0.5f) {
                u.x = fminf(2.0f * u.x, PB_ONE_MINUS_EPSILON);
                *wi = cosine_hemisphere(u);
                if (wo.z < 0.0f) wi->z *= -1.0f;
            } else {
                u.x = fminf(2.0f * (u.x - 0.5f), PB_ONE_MINUS_EPSILON);
                f3 wh = tr_sample_wh(l.tr, wo, u);
                *wi = reflect_about(wo, wh);
                if (!same_hemi(wo, *wi)) return rgb(0.0f);
            }
            *pdf = lobe_pdf<KM>(l, wo, *wi);
            return lobe_f<KM>(l, wo, *wi);
        }
    }
    return rgb(0.0f);
}

// ---- BSDF, core/reflection.rs:1496-1689
template <int NL>
struct BsdfN {
    float eta;
    f3 ns, ng, ss, ts;
    int n;
    Lobe lobe[NL];
    bool valid;  // si.bsdf is Some
};
using Bsdf = BsdfN<2>;   // the five hot materials add at most two lobes
using BsdfX = BsdfN<5>;  // uber adds up to five (uber.rs:41-112)
template <class B>
PB_D void bsdf_init(B& b, const Surf& si, float eta) {
    b.eta = eta; b.ns = si.sh_n; b.ss = normalize(si.sh_dpdu); b.ng = si.n; b.ts = cross(b.ns, b.ss); b.n = 0; b.valid = true;
}
template <class B>
PB_D f3 to_local(const B& b, f3 v) { return f3(dot(v, b.ss), dot(v, b.ts), dot(v, b.ns)); }
template <class B>
PB_D f3 to_world(const B& b, f3 v) {
    return f3(b.ss.x * v.x + b.ts.x * v.y + b.ns.x * v.z, b.ss.y * v.x + b.ts.y * v.y + b.ns.y * v.z, b.ss.z * v.x + b.ts.z * v.y + b.ns.z * v.z);
}
template <class B>
PB_D int bsdf_count(const B& b, int flags) { int c = 0; for (int i = 0; i < b.n; ++i) c += lobe_matches(b.lobe[i], flags) ? 1 : 0; return c; }
template <int KM = KM_ALL, class B>
PB_D rgb bsdf_f(const B& b, f3 wow, f3 wiw, int flags) {
    f3 wi = to_local(b, wiw), wo = to_local(b, wow);
    if (wo.z == 0.0f) return rgb(0.0f);
    bool refl = dot(wiw, b.ng) * dot(wow, b.ng) > 0.0f;
    rgb res(0.0f);
    for (int i = 0; i < b.n; ++i) {
        const Lobe& l = b.lobe[i];
        if (lobe_matches(l, flags) && ((refl && (l.type & BX_REFLECTION)) || (!refl && (l.type & BX_TRANSMISSION)))) res = res + lobe_f<KM>(l, wo, wi);
    }
    return res;
}
template <int KM = KM_ALL, class B>
PB_D float bsdf_pdf(const B& b, f3 wow, f3 wiw, int flags) {
    if (b.n == 0) return 0.0f;
    f3 wo = to_local(b, wow), wi = to_local(b, wiw);
    if (wo.z == 0.0f) return 0.0f;
    float p = 0.0f;
    int m = 0;
    for (int i = 0; i < b.n; ++i)
        if (lobe_matches(b.lobe[i], flags)) { m += 1; p += lobe_pdf<KM>(b.lobe[i], wo, wi); }
    return m > 0 ? p / (float)m : 0.0f;
}
// *pdf / *stype in-out as in the reference (path.rs initialises pdf = 0, flags = 0).
template <int KM = KM_ALL, class B>
PB_D rgb bsdf_sample(const B& b, f3 wow, f3* wiw, float2 u, float* pdf, int flags, int* stype) {
    int m = bsdf_count(b, flags);
    if (m == 0) { *pdf = 0.0f; *stype = 0; return rgb(0.0f); }
    float fm = (float)m;
    float fl = floorf(u.x * fm);
    int comp = (fl != fl || fl <= 0.0f) ? 0 : (fl >= 2147483647.0f ? 2147483647 : (int)fl);  // `as usize` saturates
    comp = min(comp, m - 1);
    int idx = 0, count = comp;
    for (int i = 0; i < b.n; ++i) {
        bool mt = lobe_matches(b.lobe[i], flags);
        if (mt && count == 0) { idx = i; break; }
        else if (mt) count -= 1;
    }
    const Lobe& l = b.lobe[idx];
    float2 ur = make_float2(fminf(u.x * fm - (float)comp, PB_ONE_MINUS_EPSILON), u.y);
    f3 wo = to_local(b, wow), wi(0.f, 0.f, 0.f);
    if (wo.z == 0.0f) return rgb(0.0f);
    *pdf = 0.0f;
    *stype = l.type;
    rgb f = lobe_sample<KM>(l, wo, &wi, ur, pdf, stype);
    if (*pdf == 0.0f) { *stype = 0; return rgb(0.0f); }
    *wiw = to_world(b, wi);
    if ((l.type & BX_SPECULAR) == 0 && m > 1)
        for (int i = 0; i < b.n; ++i)
            if (i != idx && lobe_matches(b.lobe[i], flags)) *pdf += lobe_pdf<KM>(b.lobe[i], wo, wi);
    if (m > 1) *pdf /= fm;
    if ((l.type & BX_SPECULAR) == 0) {
        bool refl = dot(*wiw, b.ng) * dot(wow, b.ng) > 0.0f;
        f = rgb(0.0f);
        for (int i = 0; i < b.n; ++i) {
            const Lobe& q = b.lobe[i];
            if (lobe_matches(q, flags) && ((refl && (q.type & BX_REFLECTION)) || (!refl && (q.type & BX_TRANSMISSION)))) f = f + lobe_f<KM>(q, wo, wi);
        }
    }
    return f;
}

// Material::compute_scattering_functions for the five hot materials (constant textures, no
// bump, mode = Radiance).  MULTI = allow_multiple_lobes: true from path.rs:123, false from whitted.rs:75 and
// directlighting.rs:90 (only glass looks at it).
// MAT >= 0: the material type is known at compile time (the shade kernel of that bin).
template <int MAT = -1, bool MULTI = true, class BS>
PB_D void material_bsdf(const pbrt_b200_material& m, const Surf& si, BS& b) {
    b.valid = false; b.n = 0;
    rgb A = rgb_clamp0(rgb3(m.a)), B = rgb_clamp0(rgb3(m.b));
    switch (MAT >= 0 ? (uint32_t)MAT : m.type) {
        case PBRT_B200_MAT_MATTE: {  // materials/matte.rs:28-52
            bsdf_init(b, si, 1.0f);
            float sig = clampf(m.f0, 0.0f, 90.0f);
            if (!is_black(A)) {
                Lobe& l = b.lobe[b.n++];
                l.type = BX_REFLECTION | BX_DIFFUSE; l.c0 = A;
                if (sig == 0.0f) l.kind = LOBE_LAMBERT;
                else {  // OrenNayar::new, reflection.rs:908-921
                    l.kind = LOBE_OREN_NAYAR;
                    float sg = (PB_PI / 180.0f) * sig, s2 = sg * sg;
                    l.p0 = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
                    l.p1 = 0.45f * s2 / (s2 + 0.09f);
                }
            }
            break;
        }
        case PBRT_B200_MAT_PLASTIC: {  // materials/plastic.rs:34-69
            bsdf_init(b, si, 1.0f);
            if (!is_black(A)) { Lobe& l = b.lobe[b.n++]; l.kind = LOBE_LAMBERT; l.type = BX_REFLECTION | BX_DIFFUSE; l.c0 = A; }
            if (!is_black(B)) {
                Lobe& l = b.lobe[b.n++];
                l.kind = LOBE_MICRO_REFL_DIEL; l.type = BX_REFLECTION | BX_GLOSSY; l.c0 = B; l.p0 = 1.5f; l.p1 = 1.0f;
                float rough = m.f0;
                if (m.remap_roughness) rough = roughness_to_alpha(rough);
                l.tr = tr_make(rough, rough);
            }
            break;
        }
        case PBRT_B200_MAT_MIRROR: {  // materials/mirror.rs:23-41
            bsdf_init(b, si, 1.0f);
            if (!is_black(A)) { Lobe& l = b.lobe[b.n++]; l.kind = LOBE_MIRROR; l.type = BX_REFLECTION | BX_SPECULAR; l.c0 = A; }
            break;
        }
        case PBRT_B200_MAT_GLASS: {  // materials/glass.rs:35-92
            float eta = m.f2, ur = m.f0, vr = m.f1;
            if (is_black(A) && is_black(B)) return;  // si.bsdf stays None: pass-through surface
            bsdf_init(b, si, eta);
            if (!MULTI && ur == 0.0f && vr == 0.0f) {
                if (!is_black(A)) { Lobe& l = b.lobe[b.n++]; l.kind = LOBE_SPEC_REFL_DIEL; l.type = BX_REFLECTION | BX_SPECULAR; l.c0 = A; l.p0 = 1.0f; l.p1 = eta; }
                if (!is_black(B)) { Lobe& l = b.lobe[b.n++]; l.kind = LOBE_SPEC_TRANS; l.type = BX_TRANSMISSION | BX_SPECULAR; l.c0 = B; l.p0 = 1.0f; l.p1 = eta; }
            } else if (ur == 0.0f && vr == 0.0f) {
                Lobe& l = b.lobe[b.n++];
                l.kind = LOBE_FRESNEL_SPECULAR; l.type = BX_REFLECTION | BX_TRANSMISSION | BX_SPECULAR; l.c0 = A; l.c1 = B; l.p0 = 1.0f; l.p1 = eta;
            } else {
                if (m.remap_roughness) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
                TRDist tr = tr_make(ur, vr);
                if (!is_black(A)) { Lobe& l = b.lobe[b.n++]; l.kind = LOBE_MICRO_REFL_DIEL; l.type = BX_REFLECTION | BX_GLOSSY; l.c0 = A; l.p0 = 1.0f; l.p1 = eta; l.tr = tr; }
                if (!is_black(B)) { Lobe& l = b.lobe[b.n++]; l.kind = LOBE_MICRO_TRANS; l.type = BX_TRANSMISSION | BX_GLOSSY; l.c0 = B; l.p0 = 1.0f; l.p1 = eta; l.tr = tr; }
            }
            break;
        }
        case PBRT_B200_MAT_METAL: {  // materials/metal.rs:78-112; textures are not clamped there
            bsdf_init(b, si, 1.0f);
            float ur = m.f0, vr = m.f1;
            if (m.remap_roughness) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
            Lobe& l = b.lobe[b.n++];
            l.kind = LOBE_MICRO_REFL_COND; l.type = BX_REFLECTION | BX_GLOSSY; l.c0 = rgb(1.0f);
            l.c1 = rgb3(m.a);  // eta
            l.c2 = rgb3(m.b);  // k
            l.tr = tr_make(ur, vr);
            break;
        }
        default: break;
    }
}

}  // namespace pb
