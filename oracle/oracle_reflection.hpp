// TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see oracle_math.hpp header).
// oracle_reflection.hpp: RGBSpectrum, Fresnel, Trowbridge-Reitz, the BxDFs reachable from
// matte/plastic/mirror/glass/metal, BSDF, and the five materials.
#pragma once
#include "oracle_accel.hpp"
#include "oracle_sampling.hpp"

namespace orc {

// ---- RGBSpectrum: src/core/spectrum.rs:78-168, ops from pbrt_macros/src/lib.rs:113-668
struct Spectrum {
    Float c[3];
    Spectrum() { c[0] = c[1] = c[2] = 0; }
    explicit Spectrum(Float v) { c[0] = c[1] = c[2] = v; }
    Spectrum(Float r, Float g, Float b) { c[0] = r; c[1] = g; c[2] = b; }
    bool is_black() const { return c[0] == 0.0f && c[1] == 0.0f && c[2] == 0.0f; }
    Float y() const { return 0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2]; }  // spectrum.rs:123-127
    Float max_component_value() const { Float m = c[0]; for (int i = 1; i < 3; ++i) m = std::fmax(m, c[i]); return m; }
    bool has_nans() const { return c[0] != c[0] || c[1] != c[1] || c[2] != c[2]; }
};
#define ORC_SPEC_OP(op)                                                                                                      \
    inline Spectrum operator op(Spectrum a, Spectrum b) { return Spectrum(a.c[0] op b.c[0], a.c[1] op b.c[1], a.c[2] op b.c[2]); } \
    inline Spectrum operator op(Spectrum a, Float b) { return Spectrum(a.c[0] op b, a.c[1] op b, a.c[2] op b); }
ORC_SPEC_OP(+) ORC_SPEC_OP(-) ORC_SPEC_OP(*) ORC_SPEC_OP(/)
#undef ORC_SPEC_OP
inline Spectrum& operator+=(Spectrum& a, Spectrum b) { a = a + b; return a; }
inline Spectrum& operator*=(Spectrum& a, Spectrum b) { a = a * b; return a; }
inline Spectrum spec_sqrt(Spectrum a) { return Spectrum(std::sqrt(a.c[0]), std::sqrt(a.c[1]), std::sqrt(a.c[2])); }
inline Spectrum spec_clamp(Spectrum a, Float lo, Float hi) { return Spectrum(clamp(a.c[0], lo, hi), clamp(a.c[1], lo, hi), clamp(a.c[2], lo, hi)); }
inline Spectrum spec3(const float* p) { return Spectrum(p[0], p[1], p[2]); }

// ---- BxDF geometry helpers, src/core/reflection.rs:78-176
inline Float cos_theta(V3 w) { return w.z; }
inline Float cos2_theta(V3 w) { return w.z * w.z; }
inline Float abs_cos_theta(V3 w) { return std::fabs(w.z); }
inline Float sin2_theta(V3 w) { return std::fmax(1.0f - cos2_theta(w), 0.0f); }
inline Float sin_theta(V3 w) { return std::sqrt(sin2_theta(w)); }
inline Float tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }
inline Float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
inline Float cos_phi(V3 w) { Float s = sin_theta(w); return s == 0.0f ? 1.0f : clamp(w.x / s, -1.0f, 1.0f); }
inline Float sin_phi(V3 w) { Float s = sin_theta(w); return s == 0.0f ? 0.0f : clamp(w.y / s, -1.0f, 1.0f); }
inline Float cos2_phi(V3 w) { return cos_phi(w) * cos_phi(w); }
inline Float sin2_phi(V3 w) { return sin_phi(w) * sin_phi(w); }
inline V3 reflect(V3 wo, V3 n) { return -wo + n * 2.0f * dot(wo, n); }
inline bool refract(V3 wi, V3 n, Float eta, V3* wt) {
    Float cos_thetai = dot(n, wi);
    Float sin2_thetai = std::fmax(1.0f - cos_thetai * cos_thetai, 0.0f);
    Float sin2_thetat = eta * eta * sin2_thetai;
    if (sin2_thetat >= 1.0f) return false;
    Float cos_thetat = std::sqrt(1.0f - sin2_thetat);
    *wt = n * (eta * cos_thetai - cos_thetat) + (-wi) * eta;
    return true;
}
inline bool same_hemisphere(V3 w, V3 wp) { return w.z * wp.z > 0.0f; }
inline Float spherical_theta(V3 v) { return std::acos(clamp(v.z, -1.0f, 1.0f)); }                        // geometry.rs:40-43
inline Float spherical_phi(V3 v) { Float p = std::atan2(v.y, v.x); return p < 0.0f ? p + 2.0f * PI : p; }  // geometry.rs:45-54

// reflection.rs:29-52
inline Float fr_dielectric(Float cos_thetai, Float etai, Float etat) {
    cos_thetai = clamp(cos_thetai, -1.0f, 1.0f);
    bool entering = cos_thetai > 0.0f;
    if (!entering) { std::swap(etai, etat); cos_thetai = std::fabs(cos_thetai); }
    Float sin_thetai = std::sqrt(std::fmax(1.0f - cos_thetai * cos_thetai, 0.0f));
    Float sin_thetat = etai / etat * sin_thetai;
    if (sin_thetat >= 1.0f) return 1.0f;
    Float cos_thetat = std::sqrt(std::fmax(1.0f - sin_thetat * sin_thetat, 0.0f));
    Float r_parl = ((etat * cos_thetai) - (etai * cos_thetat)) / ((etat * cos_thetai) + (etai * cos_thetat));
    Float r_perp = ((etai * cos_thetai) - (etat * cos_thetat)) / ((etai * cos_thetai) + (etat * cos_thetat));
    return (r_parl * r_parl + r_perp * r_perp) / 2.0f;
}
// reflection.rs:54-76
inline Spectrum fr_conductor(Float cos_thetai, Spectrum etai, Spectrum etat, Spectrum k) {
    cos_thetai = clamp(cos_thetai, -1.0f, 1.0f);
    Spectrum eta = etat / etai, etak = k / etai;
    Float cos_thetai2 = cos_thetai * cos_thetai, sin_thetai2 = 1.0f - cos_thetai2;
    Spectrum eta2 = eta * eta, etak2 = etak * etak;
    Spectrum t0 = eta2 - etak2 - Spectrum(sin_thetai2);
    Spectrum a2plusb2 = spec_sqrt(t0 * t0 + eta2 * etak2 * 4.0f);
    Spectrum t1 = a2plusb2 + Spectrum(cos_thetai2);
    Spectrum a = spec_sqrt((a2plusb2 + t0) * 0.5f);
    Spectrum t2 = a * cos_thetai * 2.0f;
    Spectrum Rs = (t1 - t2) / (t1 + t2);
    Spectrum t3 = a2plusb2 * cos_thetai2 + Spectrum(sin_thetai2 * sin_thetai2);
    Spectrum t4 = t2 * sin_thetai2;
    Spectrum Rp = Rs * (t3 - t4) / (t3 + t4);
    return (Rp + Rs) * 0.5f;
}

enum { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16, BSDF_ALL = 31 };

// ---- TrowbridgeReitzDistribution (samplevis = true), src/core/microfacet.rs:249-406
inline Float roughness_to_alpha(Float roughness) {
    roughness = std::fmax(roughness, 1.0e-3f);
    Float x = std::log(roughness);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
struct TrowbridgeReitz {
    Float alphax, alphay;
    TrowbridgeReitz() : alphax(0), alphay(0) {}
    TrowbridgeReitz(Float ax, Float ay) : alphax(std::fmax(ax, 0.001f)), alphay(std::fmax(ay, 0.001f)) {}
    Float d(V3 wh) const {
        Float t2 = tan2_theta(wh);
        if (std::isinf(t2)) return 0.0f;
        Float cos4 = cos2_theta(wh) * cos2_theta(wh);
        Float e = (cos2_phi(wh) / (alphax * alphax) + sin2_phi(wh) / (alphay * alphay)) * t2;
        return 1.0f / (PI * alphax * alphay * cos4 * (1.0f + e) * (1.0f + e));
    }
    Float lambda(V3 w) const {
        Float att = std::fabs(tan_theta(w));
        if (std::isinf(att)) return 0.0f;
        Float alpha = std::sqrt(cos2_phi(w) * alphax * alphax + sin2_phi(w) * alphay * alphay);
        Float a2t2 = (alpha * att) * (alpha * att);
        return (-1.0f + std::sqrt(1.0f + a2t2)) / 2.0f;
    }
    Float g1(V3 w) const { return 1.0f / (1.0f + lambda(w)); }
    Float g(V3 wo, V3 wi) const { return 1.0f / (1.0f + lambda(wo) + lambda(wi)); }
    Float pdf(V3 wo, V3 wh) const { return d(wh) * g1(wo) * abs_dot(wo, wh) / abs_cos_theta(wo); }
    static void sample11(Float cos_theta, Float u1, Float u2, Float* slopex, Float* slopey) {  // microfacet.rs:249-292
        if (cos_theta > 0.9999f) {
            Float r = std::sqrt(u1 / (1.0f - u1));
            Float phi = 6.28318530718f * u2;
            *slopex = r * std::cos(phi); *slopey = r * std::sin(phi);
            return;
        }
        Float sin_theta = std::sqrt(std::fmax(1.0f - cos_theta * cos_theta, 0.0f));
        Float tan_theta = sin_theta / cos_theta;
        Float a = 1.0f / tan_theta;
        Float G1 = 2.0f / (1.0f + std::sqrt(1.0f + 1.0f / (a * a)));
        Float A = 2.0f * u1 / G1 - 1.0f;
        Float tmp = 1.0f / (A * A - 1.0f);
        if (tmp > 1.0e10f) tmp = 1.0e10f;
        Float B = tan_theta;
        Float D = std::sqrt(std::fmax(B * B * tmp * tmp - (A * A - B * B) * tmp, 0.0f));
        Float slopex1 = B * tmp - D, slopex2 = B * tmp + D;
        *slopex = (A < 0.0f || slopex2 > 1.0f / tan_theta) ? slopex1 : slopex2;
        Float s;
        if (u2 > 0.5f) { s = 1.0f; u2 = 2.0f * (u2 - 0.5f); } else { s = -1.0f; u2 = 2.0f * (0.5f - u2); }
        Float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) / (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
        *slopey = s * z * std::sqrt(1.0f + *slopex * *slopex);
    }
    static V3 sample(V3 wi, Float ax, Float ay, Float u1, Float u2) {  // microfacet.rs:294-318
        V3 ws = normalize(V3(ax * wi.x, ay * wi.y, wi.z));
        Float sx, sy;
        sample11(cos_theta(ws), u1, u2, &sx, &sy);
        Float tmp = cos_phi(ws) * sx - sin_phi(ws) * sy;
        sy = sin_phi(ws) * sx + cos_phi(ws) * sy;
        sx = tmp;
        sx = ax * sx; sy = ay * sy;
        return normalize(V3(-sx, -sy, 1.0f));
    }
    V3 sample_wh(V3 wo, P2 u) const {  // microfacet.rs:366-403, samplevis branch
        bool flip = wo.z < 0.0f;
        V3 wh = sample(flip ? -wo : wo, alphax, alphay, u.x, u.y);
        if (flip) wh = -wh;
        return wh;
    }
};

// ---- BxDFs (closed set reachable from the five hot materials)
enum BxKind { BX_LAMBERT, BX_OREN_NAYAR, BX_SPEC_REFL_NOOP, BX_FRESNEL_SPECULAR, BX_MICRO_REFL, BX_MICRO_TRANS, BX_SPEC_REFL, BX_SPEC_TRANS,
              BX_FRESNEL_BLEND /* r = Rd, t = Rs; substrate.rs */ };
enum FresnelKind { FR_DIELECTRIC, FR_CONDUCTOR };

struct BxDF {
    BxKind kind;
    int type;           // BxDFType bits
    Spectrum r, t;      // R / T (or Kd...)
    Float A = 0, B = 0; // OrenNayar
    Float etaa = 1, etab = 1;
    TrowbridgeReitz distrib;
    FresnelKind fresnel = FR_DIELECTRIC;
    Float fr_etai = 1, fr_etat = 1;  // FresnelDielectric
    Spectrum fr_eta, fr_k;           // FresnelConductor (etai = 1)

    bool matches_flags(int t_) const { return (type & t_) == type; }  // reflection.rs:450-454
    Spectrum fresnel_eval(Float cosi) const {
        if (fresnel == FR_DIELECTRIC) return Spectrum(fr_dielectric(cosi, fr_etai, fr_etat));
        return fr_conductor(std::fabs(cosi), Spectrum(1.0f), fr_eta, fr_k);  // reflection.rs:566-570
    }

    Spectrum f(V3 wo, V3 wi) const {
        switch (kind) {
            case BX_LAMBERT: return r * INV_PI;  // reflection.rs:823-825
            case BX_OREN_NAYAR: {                // reflection.rs:925-952
                Float sin_thetai = sin_theta(wi), sin_thetao = sin_theta(wo);
                Float max_cos = 0.0f;
                if (sin_thetai > 1e-4f && sin_thetao > 1e-4f) {
                    Float dcos = cos_phi(wi) * cos_phi(wo) + sin_phi(wi) * sin_phi(wo);
                    max_cos = std::fmax(dcos, 0.0f);
                }
                Float sin_alpha, tan_beta;
                if (abs_cos_theta(wi) > abs_cos_theta(wo)) { sin_alpha = sin_thetao; tan_beta = sin_thetai / abs_cos_theta(wi); }
                else { sin_alpha = sin_thetai; tan_beta = sin_thetao / abs_cos_theta(wo); }
                return r * INV_PI * (A + B * max_cos * sin_alpha * tan_beta);
            }
            case BX_SPEC_REFL_NOOP: case BX_SPEC_REFL: case BX_SPEC_TRANS: return Spectrum(0.0f);  // reflection.rs:630-632, 682-684
            case BX_FRESNEL_SPECULAR: return Spectrum(1.0f);   // reflection.rs:745-747 (quirk a-Q3)
            case BX_MICRO_REFL: {                               // reflection.rs:985-1003
                Float cos_thetao = abs_cos_theta(wo), cos_thetai = abs_cos_theta(wi);
                V3 wh = wi + wo;
                if (cos_thetai == 0.0f || cos_thetao == 0.0f) return Spectrum(0.0f);
                if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return Spectrum(0.0f);
                wh = normalize(wh);
                Spectrum F = fresnel_eval(dot(wi, wh));
                Float d = distrib.d(wh), g = distrib.g(wo, wi);
                return r * d * g * F / (4.0f * cos_thetai * cos_thetao);
            }
            case BX_MICRO_TRANS: {  // reflection.rs:1064-1095
                if (same_hemisphere(wo, wi)) return Spectrum(0.0f);
                Float cos_thetao = cos_theta(wo), cos_thetai = cos_theta(wi);
                if (cos_thetai == 0.0f || cos_thetao == 0.0f) return Spectrum(0.0f);
                Float eta = cos_theta(wo) > 0.0f ? etab / etaa : etaa / etab;
                V3 wh = normalize(wo + wi * eta);
                if (wh.z < 0.0f) wh = -wh;
                if (dot(wo, wh) * dot(wi, wh) > 0.0f) return Spectrum(0.0f);
                Spectrum F(fr_dielectric(dot(wo, wh), etaa, etab));
                Float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
                Float factor = 1.0f / eta;  // TransportMode::Radiance
                return (Spectrum(1.0f) - F) * t *
                       std::fabs(distrib.d(wh) * distrib.g(wo, wi) * eta * eta * abs_dot(wi, wh) * abs_dot(wo, wh) * factor * factor /
                                 (cos_thetai * cos_thetao * sqrt_denom * sqrt_denom));
            }
            case BX_FRESNEL_BLEND: {  // reflection.rs:1167-1185
                auto pow5 = [](Float v) { return (v * v) * (v * v) * v; };
                Spectrum diffuse = r * (Spectrum(1.0f) - t) * (28.0f / (23.0f * PI)) * (1.0f - pow5(1.0f - 0.5f * abs_cos_theta(wi))) *
                                   (1.0f - pow5(1.0f - 0.5f * abs_cos_theta(wo)));
                V3 wh = wi + wo;
                if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return Spectrum(0.0f);
                wh = normalize(wh);
                Spectrum schlick = t + (Spectrum(1.0f) - t) * pow5(1.0f - dot(wi, wh));  // schlick_fresnel, :1157-1161
                Spectrum specular = schlick * (distrib.d(wh) / (4.0f * abs_dot(wi, wh) * std::fmax(abs_cos_theta(wi), abs_cos_theta(wo))));
                return diffuse + specular;
            }
        }
        return Spectrum(0.0f);
    }

    Float pdf(V3 wo, V3 wi) const {
        switch (kind) {
            case BX_LAMBERT: case BX_OREN_NAYAR: case BX_FRESNEL_SPECULAR:  // reflection.rs:438-445, 788-794
                return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * INV_PI : 0.0f;
            case BX_SPEC_REFL_NOOP: case BX_SPEC_REFL: case BX_SPEC_TRANS: return 0.0f;  // reflection.rs:642-644, 710-712
            case BX_MICRO_REFL: {  // reflection.rs:1021-1027
                if (!same_hemisphere(wo, wi)) return 0.0f;
                V3 wh = normalize(wo + wi);
                return distrib.pdf(wo, wh) / (4.0f * dot(wo, wh));
            }
            case BX_MICRO_TRANS: {  // reflection.rs:1115-1129
                if (same_hemisphere(wo, wi)) return 0.0f;
                Float eta = cos_theta(wo) > 0.0f ? etaa / etab : etab / etaa;
                V3 wh = normalize(wo + wi * eta);
                if (dot(wo, wh) * dot(wi, wh) > 0.0f) return 0.0f;
                Float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
                Float dwh_dwi = std::fabs(eta * eta * dot(wi, wh)) / (sqrt_denom * sqrt_denom);
                return distrib.pdf(wo, wh) * dwh_dwi;
            }
            case BX_FRESNEL_BLEND: {  // reflection.rs:1212-1218
                if (!same_hemisphere(wo, wi)) return 0.0f;
                V3 wh = normalize(wo + wi);
                Float pdf_wh = distrib.pdf(wo, wh);
                return 0.5f * (abs_cos_theta(wi) * INV_PI + pdf_wh / (4.0f * dot(wo, wh)));
            }
        }
        return 0.0f;
    }

    // `pdf` and `sampled_type` are in/out exactly like the reference's &mut parameters:
    // early returns leave them untouched.
    Spectrum sample_f(V3 wo, V3* wi, P2 u, Float* pdf_, int* sampled_type) const {
        switch (kind) {
            case BX_LAMBERT: case BX_OREN_NAYAR: {  // reflection.rs:392-405
                *wi = cosine_sample_hemisphere(u);
                if (wo.z < 0.0f) wi->z *= -1.0f;
                *pdf_ = pdf(wo, *wi);
                return f(wo, *wi);
            }
            case BX_SPEC_REFL_NOOP: {  // reflection.rs:634-640 with FresnelNoOp
                *wi = V3(-wo.x, -wo.y, wo.z);
                *pdf_ = 1.0f;
                return Spectrum(1.0f) * r / abs_cos_theta(*wi);
            }
            case BX_SPEC_REFL: {  // reflection.rs:634-640 with FresnelDielectric (glass without allow_multiple_lobes)
                *wi = V3(-wo.x, -wo.y, wo.z);
                *pdf_ = 1.0f;
                return fresnel_eval(cos_theta(*wi)) * r / abs_cos_theta(*wi);
            }
            case BX_SPEC_TRANS: {  // reflection.rs:686-708, mode = Radiance
                Float etai, etat;
                if (cos_theta(wo) > 0.0f) { etai = etaa; etat = etab; } else { etai = etab; etat = etaa; }
                if (!refract(wo, face_forward(V3(0, 0, 1), wo), etai / etat, wi)) return Spectrum(0.0f);
                *pdf_ = 1.0f;
                Spectrum ft = t * (Spectrum(1.0f) - Spectrum(fr_dielectric(cos_theta(*wi), etaa, etab)));
                ft = ft * ((etai * etai) / (etat * etat));
                return ft / abs_cos_theta(*wi);
            }
            case BX_FRESNEL_SPECULAR: {  // reflection.rs:749-786
                Float F = fr_dielectric(cos_theta(wo), etaa, etab);
                if (u.x < F) {
                    *wi = V3(-wo.x, -wo.y, wo.z);
                    *sampled_type = BSDF_SPECULAR | BSDF_REFLECTION;
                    *pdf_ = F;
                    return r / abs_cos_theta(*wi) * F;
                }
                Float etai, etat;
                if (cos_theta(wo) > 0.0f) { etai = etaa; etat = etab; } else { etai = etab; etat = etaa; }
                if (!refract(wo, face_forward(V3(0, 0, 1), wo), etai / etat, wi)) return Spectrum(0.0f);
                Spectrum ft = t * (1.0f - F);
                ft = ft * ((etai * etai) / (etat * etat));
                *sampled_type = BSDF_SPECULAR | BSDF_TRANSMISSION;
                *pdf_ = 1.0f - F;
                return ft / abs_cos_theta(*wi);
            }
            case BX_FRESNEL_BLEND: {  // reflection.rs:1187-1210
                if (u.x < 0.5f) {
                    u.x = std::fmin(2.0f * u.x, ONE_MINUS_EPSILON);
                    *wi = cosine_sample_hemisphere(u);
                    if (wo.z < 0.0f) wi->z *= -1.0f;
                } else {
                    u.x = std::fmin(2.0f * (u.x - 0.5f), ONE_MINUS_EPSILON);
                    V3 wh = distrib.sample_wh(wo, u);
                    *wi = reflect(wo, wh);
                    if (!same_hemisphere(wo, *wi)) return Spectrum(0.0f);
                }
                *pdf_ = pdf(wo, *wi);
                return f(wo, *wi);
            }
            case BX_MICRO_REFL: {  // reflection.rs:1005-1019
                if (wo.z == 0.0f) return Spectrum(0.0f);
                V3 wh = distrib.sample_wh(wo, u);
                if (dot(wo, wh) < 0.0f) return Spectrum(0.0f);
                *wi = reflect(wo, wh);
                if (!same_hemisphere(wo, *wi)) return Spectrum(0.0f);
                *pdf_ = distrib.pdf(wo, wh) / (4.0f * dot(wo, wh));
                return f(wo, *wi);
            }
            case BX_MICRO_TRANS: {  // reflection.rs:1097-1113
                if (wo.z == 0.0f) return Spectrum(0.0f);
                V3 wh = distrib.sample_wh(wo, u);
                if (dot(wo, wh) < 0.0f) return Spectrum(0.0f);
                Float eta = cos_theta(wo) > 0.0f ? etaa / etab : etab / etaa;
                if (!refract(wo, wh, eta, wi)) return Spectrum(0.0f);
                *pdf_ = pdf(wo, *wi);
                return f(wo, *wi);
            }
        }
        return Spectrum(0.0f);
    }
};

// ---- BSDF, src/core/reflection.rs:1496-1689
struct BSDF {
    Float eta = 1;
    V3 ns, ng, ss, ts;
    int n_bxdfs = 0;
    union { BxDF bxdfs[5]; };  // uber adds up to five (uber.rs:41-112); a union member is not default-constructed: only added lobes are ever written
    bool valid = false;  // si.bsdf is Some
    BSDF() {}

    void init(const SurfaceInteraction& si, Float eta_) {  // BSDF::new
        eta = eta_; ns = si.sh_n; ss = normalize(si.sh_dpdu); ng = si.n; ts = cross(ns, ss); n_bxdfs = 0; valid = true;
    }
    void add(const BxDF& b) { bxdfs[n_bxdfs++] = b; }
    int num_components(int flags) const { int n = 0; for (int i = 0; i < n_bxdfs; ++i) n += bxdfs[i].matches_flags(flags); return n; }
    V3 world_to_local(V3 v) const { return V3(dot(v, ss), dot(v, ts), dot(v, ns)); }
    V3 local_to_world(V3 v) const {
        return V3(ss.x * v.x + ts.x * v.y + ns.x * v.z, ss.y * v.x + ts.y * v.y + ns.y * v.z, ss.z * v.x + ts.z * v.y + ns.z * v.z);
    }
    Spectrum f(V3 wow, V3 wiw, int flags) const {
        V3 wi = world_to_local(wiw), wo = world_to_local(wow);
        if (wo.z == 0.0f) return Spectrum(0.0f);
        bool refl = dot(wiw, ng) * dot(wow, ng) > 0.0f;
        Spectrum res(0.0f);
        for (int i = 0; i < n_bxdfs; ++i) {
            const BxDF& b = bxdfs[i];
            if (b.matches_flags(flags) && ((refl && (b.type & BSDF_REFLECTION)) || (!refl && (b.type & BSDF_TRANSMISSION)))) res += b.f(wo, wi);
        }
        return res;
    }
    Spectrum sample_f(V3 wow, V3* wiw, P2 u, Float* pdf, int ty, int* sampled_type) const {
        int matching = num_components(ty);
        if (matching == 0) { *pdf = 0.0f; *sampled_type = 0; return Spectrum(0.0f); }
        int comp = (int)std::min<int64_t>(f2u_sat(std::floor(u.x * (Float)matching)), matching - 1);
        const BxDF* bx = nullptr;
        int count = comp, idx = 0;
        for (int i = 0; i < n_bxdfs; ++i) {
            bool m = bxdfs[i].matches_flags(ty);
            if (m && count == 0) { bx = &bxdfs[i]; idx = i; break; }
            else if (m) count -= 1;
        }
        P2 ur(std::fmin(u.x * (Float)matching - (Float)comp, ONE_MINUS_EPSILON), u.y);
        V3 wo = world_to_local(wow), wi;
        if (wo.z == 0.0f) return Spectrum(0.0f);
        *pdf = 0.0f;
        *sampled_type = bx->type;
        Spectrum f_ = bx->sample_f(wo, &wi, ur, pdf, sampled_type);
        if (*pdf == 0.0f) { *sampled_type = 0; return Spectrum(0.0f); }
        *wiw = local_to_world(wi);
        if ((bx->type & BSDF_SPECULAR) == 0 && matching > 1)
            for (int i = 0; i < n_bxdfs; ++i)
                if (idx != i && bxdfs[i].matches_flags(ty)) *pdf += bxdfs[i].pdf(wo, wi);
        if (matching > 1) *pdf /= (Float)matching;
        if ((bx->type & BSDF_SPECULAR) == 0) {
            bool refl = dot(*wiw, ng) * dot(wow, ng) > 0.0f;
            f_ = Spectrum(0.0f);
            for (int i = 0; i < n_bxdfs; ++i)
                if (bxdfs[i].matches_flags(ty) && ((refl && (bxdfs[i].type & BSDF_REFLECTION)) || (!refl && (bxdfs[i].type & BSDF_TRANSMISSION))))
                    f_ += bxdfs[i].f(wo, wi);
        }
        return f_;
    }
    Float pdf(V3 wow, V3 wiw, int flags) const {
        if (n_bxdfs == 0) return 0.0f;
        V3 wo = world_to_local(wow), wi = world_to_local(wiw);
        if (wo.z == 0.0f) return 0.0f;
        Float p = 0.0f;
        int matching = 0;
        for (int i = 0; i < n_bxdfs; ++i)
            if (bxdfs[i].matches_flags(flags)) { matching += 1; p += bxdfs[i].pdf(wo, wi); }
        return matching > 0 ? p / (Float)matching : 0.0f;
    }
};

// ---- Material::compute_scattering_functions for the five hot materials (constant textures,
// no bump map, mode = Radiance; allow_multiple_lobes = true as passed by path.rs:123, false from whitted.rs:75 and
// directlighting.rs:90 -- only glass looks at it).
inline void compute_scattering_functions(const pbrt_b200_material& m, const SurfaceInteraction& si, BSDF* bsdf, bool allow_multiple_lobes = true) {
    bsdf->valid = false;
    switch (m.type) {
        case PBRT_B200_MAT_MATTE: {  // matte.rs:28-52
            bsdf->init(si, 1.0f);
            Spectrum r = spec_clamp(spec3(m.a), 0.0f, INFINITY_F);
            Float sig = clamp(m.f0, 0.0f, 90.0f);
            if (!r.is_black()) {
                BxDF b;
                b.type = BSDF_REFLECTION | BSDF_DIFFUSE; b.r = r;
                if (sig == 0.0f) b.kind = BX_LAMBERT;
                else {  // OrenNayar::new, reflection.rs:908-921
                    b.kind = BX_OREN_NAYAR;
                    Float sigma = radians(sig), sigma2 = sigma * sigma;
                    b.A = 1.0f - (sigma2 / (2.0f * (sigma2 + 0.33f)));
                    b.B = 0.45f * sigma2 / (sigma2 + 0.09f);
                }
                bsdf->add(b);
            }
            break;
        }
        case PBRT_B200_MAT_PLASTIC: {  // plastic.rs:34-69
            bsdf->init(si, 1.0f);
            Spectrum kd = spec_clamp(spec3(m.a), 0.0f, INFINITY_F);
            if (!kd.is_black()) { BxDF b; b.kind = BX_LAMBERT; b.type = BSDF_REFLECTION | BSDF_DIFFUSE; b.r = kd; bsdf->add(b); }
            Spectrum ks = spec_clamp(spec3(m.b), 0.0f, INFINITY_F);
            if (!ks.is_black()) {
                BxDF b; b.kind = BX_MICRO_REFL; b.type = BSDF_REFLECTION | BSDF_GLOSSY; b.r = ks;
                b.fresnel = FR_DIELECTRIC; b.fr_etai = 1.5f; b.fr_etat = 1.0f;
                Float rough = m.f0;
                if (m.remap_roughness) rough = roughness_to_alpha(rough);
                b.distrib = TrowbridgeReitz(rough, rough);
                bsdf->add(b);
            }
            break;
        }
        case PBRT_B200_MAT_MIRROR: {  // mirror.rs:23-41
            bsdf->init(si, 1.0f);
            Spectrum R = spec_clamp(spec3(m.a), 0.0f, INFINITY_F);
            if (!R.is_black()) { BxDF b; b.kind = BX_SPEC_REFL_NOOP; b.type = BSDF_REFLECTION | BSDF_SPECULAR; b.r = R; bsdf->add(b); }
            break;
        }
        case PBRT_B200_MAT_GLASS: {  // glass.rs:35-92
            Float eta = m.f2, urough = m.f0, vrough = m.f1;
            Spectrum R = spec_clamp(spec3(m.a), 0.0f, INFINITY_F), T = spec_clamp(spec3(m.b), 0.0f, INFINITY_F);
            if (R.is_black() && T.is_black()) return;  // si.bsdf stays None (quirk a-Q4)
            bsdf->init(si, eta);
            bool is_specular = urough == 0.0f && vrough == 0.0f;
            if (is_specular && allow_multiple_lobes) {
                BxDF b; b.kind = BX_FRESNEL_SPECULAR; b.type = BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR;
                b.r = R; b.t = T; b.etaa = 1.0f; b.etab = eta;
                bsdf->add(b);
            } else {
                if (m.remap_roughness) { urough = roughness_to_alpha(urough); vrough = roughness_to_alpha(vrough); }
                TrowbridgeReitz distrib(urough, vrough);
                if (!R.is_black()) {
                    BxDF b; b.kind = BX_MICRO_REFL; b.type = BSDF_REFLECTION | BSDF_GLOSSY; b.r = R; b.distrib = distrib;
                    b.fresnel = FR_DIELECTRIC; b.fr_etai = 1.0f; b.fr_etat = eta;
                    if (is_specular) { b.kind = BX_SPEC_REFL; b.type = BSDF_REFLECTION | BSDF_SPECULAR; }  // glass.rs:69-74
                    bsdf->add(b);
                }
                if (!T.is_black()) {
                    BxDF b; b.kind = BX_MICRO_TRANS; b.type = BSDF_TRANSMISSION | BSDF_GLOSSY; b.t = T; b.distrib = distrib; b.etaa = 1.0f; b.etab = eta;
                    if (is_specular) { b.kind = BX_SPEC_TRANS; b.type = BSDF_TRANSMISSION | BSDF_SPECULAR; }  // glass.rs:79-84
                    bsdf->add(b);
                }
            }
            break;
        }
        case PBRT_B200_MAT_METAL: {  // metal.rs:78-112
            bsdf->init(si, 1.0f);
            Float urough = m.f0, vrough = m.f1;
            if (m.remap_roughness) { urough = roughness_to_alpha(urough); vrough = roughness_to_alpha(vrough); }
            BxDF b; b.kind = BX_MICRO_REFL; b.type = BSDF_REFLECTION | BSDF_GLOSSY; b.r = Spectrum(1.0f);
            b.fresnel = FR_CONDUCTOR; b.fr_eta = spec3(m.a); b.fr_k = spec3(m.b);
            b.distrib = TrowbridgeReitz(urough, vrough);
            bsdf->add(b);
            break;
        }
        default: break;
    }
}

}  // namespace orc
