// Device textures: full surface interactions, ray differentials, the texture interpreter, MIPMap lookups, bump mapping
// and the texture-parameterised materials (the five hot ones plus uber and substrate).  Only the Q_TEX shade kernels reach this code: a
// scene without a `textured` material row compiles and runs exactly what it did before.
// Reference arithmetic cited per function (paths relative to pbrt-rust/src); shading is tolerance-parity (see shading.cuh).
#pragma once
#include "shading.cuh"

namespace pb {

#define PB_TEX_STACK 8 /* operands of one postfix texture program (host.py: TEX_STACK_LIMIT) */
#define PB_EWA_MAX_SPAN 2048 /* texels per axis one EWA lookup may scan (mip_ewa) */

// RayDifferential (core/geometry/ray.rs:18-42) of the ray that hit the surface
struct RayDiff {
    bool has;
    f3 rxo, rxd, ryo, ryd;
};

// The rest of SurfaceInteraction (core/interaction.rs:149-183) next to `Surf`
struct SurfX {
    float2 uv;
    f3 dpdu, dpdv, dndu, dndv;        // geometric
    f3 sh_dpdv, sh_dndu, sh_dndv;     // Shading (sh_n and sh_dpdu live in Surf)
    f3 dpdx, dpdy;
    float dudx, dvdx, dudy, dvdy;
    bool shape_some, shape_flip;      // `shape` is Some / reverse_orientation ^ transform_swapshandedness
};

// SurfaceInteraction::set_shading_geometry, interaction.rs:234-255
PB_D void set_shading_geometry(Surf& si, SurfX& sx, f3 dpdus, f3 dpdvs, f3 dndus, f3 dndvs, bool authoritative) {
    si.sh_n = normalize(cross(dpdus, dpdvs));
    if (sx.shape_some) {
        if (sx.shape_flip) si.sh_n = -si.sh_n;
        if (authoritative) si.n = face_forward(si.n, si.sh_n);
        else si.sh_n = face_forward(si.sh_n, si.n);
    }
    si.sh_dpdu = dpdus; sx.sh_dpdv = dpdvs; sx.sh_dndu = dndus; sx.sh_dndv = dndvs;
}

// Triangle::intersect tail (shapes/triangle.rs:236-392) with every field a texture can read
PB_D void triangle_surface_full(const DevScene& s, uint32_t slot, f3 ray_d, float b0, float b1, float b2, Surf& si, SurfX& sx, uint32_t* flags_out) {
    const float4* tp = s.tris + 3ull * slot;
    const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
    const uint32_t flags = __float_as_uint(v1.w), shape_index = __float_as_uint(v2.w);
    *flags_out = flags;
    const f3 p0(v0.x, v0.y, v0.z), p1(v1.x, v1.y, v1.z), p2(v2.x, v2.y, v2.z);
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
    float xs = fabsf(b0 * p0.x) + fabsf(b1 * p1.x) + fabsf(b2 * p2.x);
    float ys = fabsf(b0 * p0.y) + fabsf(b1 * p1.y) + fabsf(b2 * p2.y);
    float zs = fabsf(b0 * p0.z) + fabsf(b1 * p1.z) + fabsf(b2 * p2.z);
    si.p_error = f3(xs, ys, zs) * gamma_n(7);
    si.p = p0 * b0 + p1 * b1 + p2 * b2;
    sx.uv = make_float2(uv0.x * b0 + uv1.x * b1 + uv2.x * b2, uv0.y * b0 + uv1.y * b1 + uv2.y * b2);
    const bool ro = flags & PBRT_B200_PRIM_REVERSE_ORIENTATION, sh = flags & PBRT_B200_PRIM_SWAPS_HANDEDNESS;
    const bool flip = ro != sh;
    sx.shape_some = true; sx.shape_flip = flip;
    f3 nn = normalize(cross(dp02, dp12));
    si.n = flip ? -nn : nn;
    si.sh_n = si.n;
    si.sh_dpdu = dpdu;
    si.wo = -ray_d;
    sx.dpdu = dpdu; sx.dpdv = dpdv; sx.sh_dpdv = dpdv;
    sx.dndu = sx.dndv = sx.sh_dndu = sx.sh_dndv = f3(0.f, 0.f, 0.f);
    const bool has_n = (flags & PBRT_B200_PRIM_HAS_N) && s.vertex_n, has_s = (flags & PBRT_B200_PRIM_HAS_S) && s.vertex_s;
    if (has_n || has_s) {
        const uint32_t* idx = s.tri_indices + 3ull * shape_index;
        const uint32_t i0 = idx[0], i1 = idx[1], i2 = idx[2];
        f3 n0(0.f, 0.f, 0.f), n1(0.f, 0.f, 0.f), n2(0.f, 0.f, 0.f), ns;
        if (has_n) {
            const float* N = s.vertex_n;
            n0 = f3(N[3 * i0], N[3 * i0 + 1], N[3 * i0 + 2]); n1 = f3(N[3 * i1], N[3 * i1 + 1], N[3 * i1 + 2]); n2 = f3(N[3 * i2], N[3 * i2 + 1], N[3 * i2 + 2]);
            ns = n0 * b0 + n1 * b1 + n2 * b2;
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
        f3 dndu(0.f, 0.f, 0.f), dndv(0.f, 0.f, 0.f);  // triangle.rs:339-377
        if (has_n) {
            f3 dn1 = n0 - n2, dn2 = n1 - n2;
            if (fabsf(determinant) < 1.0e-8f) {
                f3 dn = cross(n2 - n0, n1 - n0);
                if (len2(dn) != 0.0f) coordinate_system(dn, &dndu, &dndv);
            } else {
                float invdet = 1.0f / determinant;
                dndu = (dn1 * duv12.y - dn2 * duv02.y) * invdet;
                dndv = (dn1 * -duv12.x + dn2 * duv02.x) * invdet;
            }
        }
        if (ro) ts = -ts;
        set_shading_geometry(si, sx, ss, ts, dndu, dndv, true);
    }
}

// Sphere::intersect tail (shapes/sphere.rs:150-196) + transform_surface_interaction (transform.rs:607-636)
PB_D void sphere_surface_full(const pbrt_b200_sphere& sp, f3 ray_o, f3 ray_d, float t, Surf& r, SurfX& sx) {
    f3 oo, od, oe, de;
    sphere_object_ray(sp, ray_o, ray_d, &oo, &od, &oe, &de);
    const float phi_max = (PB_PI / 180.0f) * 360.0f;
    const float theta_min = acosf(-1.0f), theta_max = acosf(1.0f);
    f3 ph = oo + od * t;
    ph = ph * (sp.radius / len(ph));
    if (ph.x == 0.0f && ph.y == 0.0f) ph.x = 1.0e-5f * sp.radius;
    float phi = atan2f(ph.y, ph.x);
    if (phi < 0.0f) phi += 2.0f * PB_PI;
    float u = phi / phi_max;
    float theta = acosf(clampf(ph.z / sp.radius, -1.0f, 1.0f));
    float v = (theta - theta_min) / (theta_max - theta_min);
    float zradius = sqrtf(ph.x * ph.x + ph.y * ph.y);
    float inv_radius = 1.0f / zradius;
    float cos_phi = ph.x * inv_radius, sin_phi = ph.y * inv_radius;
    f3 dpdu(-phi_max * ph.y, phi_max * ph.x, 0.0f);
    f3 dpdv = f3(ph.z * cos_phi, ph.z * sin_phi, -sp.radius * sinf(theta)) * (theta_max - theta_min);
    f3 d2pduu = f3(ph.x, ph.y, 0.0f) * -phi_max * phi_max;
    f3 d2pduv = f3(-sin_phi, cos_phi, 0.0f) * (theta_max - theta_min) * ph.z * phi_max;
    f3 d2pdvv = f3(ph.x, ph.y, ph.z) * -(theta_max - theta_min) * (theta_max - theta_min);
    float E = dot(dpdu, dpdu), F = dot(dpdu, dpdv), G = dot(dpdv, dpdv);
    f3 Nn = normalize(cross(dpdu, dpdv));
    float e = dot(Nn, d2pduu), f = dot(Nn, d2pduv), g = dot(Nn, d2pdvv);
    float inv_EGF2 = 1.0f / (E * G - F * F);
    f3 dndu = dpdu * (f * F - e * G) * inv_EGF2 + dpdv * (e * F - f * E) * inv_EGF2;
    f3 dndv = dpdu * (g * F - f * G) * inv_EGF2 + dpdv * (f * F - g * E) * inv_EGF2;
    f3 pe = vabs(ph) * gamma_n(5);
    f3 n = Nn;
    f3 wo = normalize(-od);
    r.p = xf_point_abs_err(sp.object_to_world, ph, pe, &r.p_error);
    r.n = normalize(xf_normal(sp.world_to_object, n));
    r.wo = normalize(xf_vector(sp.object_to_world, wo));
    r.sh_n = normalize(xf_normal(sp.world_to_object, n));
    r.sh_dpdu = xf_vector(sp.object_to_world, dpdu);
    r.sh_n = face_forward(r.sh_n, r.n);
    sx.uv = make_float2(u, v);
    sx.dpdu = r.sh_dpdu; sx.dpdv = sx.sh_dpdv = xf_vector(sp.object_to_world, dpdv);
    sx.dndu = sx.sh_dndu = xf_normal(sp.world_to_object, dndu);
    sx.dndv = sx.sh_dndv = xf_normal(sp.world_to_object, dndv);
    sx.shape_some = false; sx.shape_flip = false;  // the sphere passes `None` as the interaction's shape (sphere.rs:188-191)
}

// Closest-hit record -> full interaction, through an instance when the hit has one (primitive.rs:58-80)
static __device__ __noinline__ void surface_full(const DevScene* sp, uint32_t inst, uint32_t slot, f3 ray_o, f3 ray_d, float t, float b0, float b1, float b2,
                                                 Surf* si_out, SurfX* sx_out, uint32_t* flags_out) {
    const DevScene& s = *sp;
    Surf si;
    SurfX sx;
    const DevInstance* in = (inst != PBRT_B200_NO_HIT) ? s.instances + inst : nullptr;
    f3 o = ray_o, d = ray_d;
    if (in) { float tm; xf_ray(in->world_to_prim, ray_o, ray_d, PB_INF, &o, &d, &tm); }
    const uint32_t fl = __float_as_uint(__ldg(&s.tris[3ull * slot + 1].w));
    if (fl & PB_TRI_SPHERE) {
        *flags_out = fl;
        sphere_surface_full(s.spheres[__float_as_uint(__ldg(&s.tris[3ull * slot + 2].w))], o, d, t, si, sx);
    } else triangle_surface_full(s, slot, d, b0, b1, b2, si, sx, flags_out);
    if (in && !(in->flags & PB_INST_IDENTITY)) {  // transform_surface_interaction, transform.rs:607-636
        Surf r;
        r.p = xf_point_abs_err(in->prim_to_world, si.p, si.p_error, &r.p_error);
        r.n = normalize(xf_normal(in->world_to_prim, si.n));
        r.wo = normalize(xf_vector(in->prim_to_world, si.wo));
        r.sh_n = normalize(xf_normal(in->world_to_prim, si.sh_n));
        r.sh_dpdu = xf_vector(in->prim_to_world, si.sh_dpdu);
        r.sh_n = face_forward(r.sh_n, r.n);
        sx.dpdu = xf_vector(in->prim_to_world, sx.dpdu); sx.dpdv = xf_vector(in->prim_to_world, sx.dpdv);
        sx.sh_dpdv = xf_vector(in->prim_to_world, sx.sh_dpdv);
        sx.dndu = xf_normal(in->world_to_prim, sx.dndu); sx.dndv = xf_normal(in->world_to_prim, sx.dndv);
        sx.sh_dndu = xf_normal(in->world_to_prim, sx.sh_dndu); sx.sh_dndv = xf_normal(in->world_to_prim, sx.sh_dndv);
        si = r;
    }
    sx.dpdx = sx.dpdy = f3(0.f, 0.f, 0.f);
    sx.dudx = sx.dvdx = sx.dudy = sx.dvdy = 0.0f;
    *si_out = si; *sx_out = sx;
}

// solve_linearsystem_2x2, core/transform.rs:174-186
PB_D bool solve_2x2(float a00, float a01, float a10, float a11, float b0, float b1, float* x0, float* x1) {
    float det = a00 * a11 - a01 * a10;
    if (fabsf(det) < 1.0e-10f) return false;
    *x0 = (a11 * b0 - a01 * b1) / det;
    *x1 = (a00 * b1 - a10 * b0) / det;
    return !(isnan(*x0) || isnan(*x1));
}
// SurfaceInteraction::compute_differentials, interaction.rs:269-342
PB_D void compute_differentials(const Surf& si, SurfX& sx, const RayDiff& r) {
    sx.dudx = sx.dvdx = sx.dudy = sx.dvdy = 0.0f;
    sx.dpdx = sx.dpdy = f3(0.f, 0.f, 0.f);
    if (!r.has) return;
    float d = dot(si.n, si.p);
    float tx = -(dot(si.n, r.rxo) - d) / dot(si.n, r.rxd);
    if (isinf(tx) || isnan(tx)) return;
    f3 px = r.rxo + r.rxd * tx;
    float ty = -(dot(si.n, r.ryo) - d) / dot(si.n, r.ryd);
    if (isinf(ty) || isnan(ty)) return;
    f3 py = r.ryo + r.ryd * ty;
    sx.dpdx = px - si.p;
    sx.dpdy = py - si.p;
    int d0, d1;
    if (fabsf(si.n.x) > fabsf(si.n.y) && fabsf(si.n.x) > fabsf(si.n.z)) { d0 = 1; d1 = 2; }
    else if (fabsf(si.n.y) > fabsf(si.n.z)) { d0 = 0; d1 = 2; }
    else { d0 = 0; d1 = 1; }
    float a00 = comp(sx.dpdu, d0), a01 = comp(sx.dpdv, d0), a10 = comp(sx.dpdu, d1), a11 = comp(sx.dpdv, d1);
    float bx0 = comp(px, d0) - comp(si.p, d0), bx1 = comp(px, d1) - comp(si.p, d1);
    float by0 = comp(py, d0) - comp(si.p, d0), by1 = comp(py, d1) - comp(si.p, d1);
    if (!solve_2x2(a00, a01, a10, a11, bx0, bx1, &sx.dudx, &sx.dvdx)) sx.dudx = sx.dvdx = 0.0f;
    if (!solve_2x2(a00, a01, a10, a11, by0, by1, &sx.dudy, &sx.dvdy)) sx.dudy = sx.dvdy = 0.0f;
}

// ---- Perlin noise, core/texture.rs:25-67,330-431 -------------------------------------------------------------------------------
static __device__ const unsigned char c_noise_perm[512] = {
    151, 160, 137, 91, 90, 15, 131, 13, 201, 95, 96, 53, 194, 233, 7, 225, 140, 36, 103, 30, 69, 142, 8, 99, 37, 240, 21, 10, 23, 190, 6, 148, 247, 120,
    234, 75, 0, 26, 197, 62, 94, 252, 219, 203, 117, 35, 11, 32, 57, 177, 33, 88, 237, 149, 56, 87, 174, 20, 125, 136, 171, 168, 68, 175, 74, 165, 71,
    134, 139, 48, 27, 166, 77, 146, 158, 231, 83, 111, 229, 122, 60, 211, 133, 230, 220, 105, 92, 41, 55, 46, 245, 40, 244, 102, 143, 54, 65, 25, 63,
    161, 1, 216, 80, 73, 209, 76, 132, 187, 208, 89, 18, 169, 200, 196, 135, 130, 116, 188, 159, 86, 164, 100, 109, 198, 173, 186, 3, 64, 52, 217, 226,
    250, 124, 123, 5, 202, 38, 147, 118, 126, 255, 82, 85, 212, 207, 206, 59, 227, 47, 16, 58, 17, 182, 189, 28, 42, 223, 183, 170, 213, 119, 248, 152,
    2, 44, 154, 163, 70, 221, 153, 101, 155, 167, 43, 172, 9, 129, 22, 39, 253, 19, 98, 108, 110, 79, 113, 224, 232, 178, 185, 112, 104, 218, 246, 97,
    228, 251, 34, 242, 193, 238, 210, 144, 12, 191, 179, 162, 241, 81, 51, 145, 235, 249, 14, 239, 107, 49, 192, 214, 31, 181, 199, 106, 157, 184, 84,
    204, 176, 115, 121, 50, 45, 127, 4, 150, 254, 138, 236, 205, 93, 222, 114, 67, 29, 24, 72, 243, 141, 128, 195, 78, 66, 215, 61, 156, 180,
    151, 160, 137, 91, 90, 15, 131, 13, 201, 95, 96, 53, 194, 233, 7, 225, 140, 36, 103, 30, 69, 142, 8, 99, 37, 240, 21, 10, 23, 190, 6, 148, 247, 120,
    234, 75, 0, 26, 197, 62, 94, 252, 219, 203, 117, 35, 11, 32, 57, 177, 33, 88, 237, 149, 56, 87, 174, 20, 125, 136, 171, 168, 68, 175, 74, 165, 71,
    134, 139, 48, 27, 166, 77, 146, 158, 231, 83, 111, 229, 122, 60, 211, 133, 230, 220, 105, 92, 41, 55, 46, 245, 40, 244, 102, 143, 54, 65, 25, 63,
    161, 1, 216, 80, 73, 209, 76, 132, 187, 208, 89, 18, 169, 200, 196, 135, 130, 116, 188, 159, 86, 164, 100, 109, 198, 173, 186, 3, 64, 52, 217, 226,
    250, 124, 123, 5, 202, 38, 147, 118, 126, 255, 82, 85, 212, 207, 206, 59, 227, 47, 16, 58, 17, 182, 189, 28, 42, 223, 183, 170, 213, 119, 248, 152,
    2, 44, 154, 163, 70, 221, 153, 101, 155, 167, 43, 172, 9, 129, 22, 39, 253, 19, 98, 108, 110, 79, 113, 224, 232, 178, 185, 112, 104, 218, 246, 97,
    228, 251, 34, 242, 193, 238, 210, 144, 12, 191, 179, 162, 241, 81, 51, 145, 235, 249, 14, 239, 107, 49, 192, 214, 31, 181, 199, 106, 157, 184, 84,
    204, 176, 115, 121, 50, 45, 127, 4, 150, 254, 138, 236, 205, 93, 222, 114, 67, 29, 24, 72, 243, 141, 128, 195, 78, 66, 215, 61, 156, 180};

// The noise functions are written with __fmul_rn / __fadd_rn: this file is compiled with --use_fast_math (shade.o), which contracts a * b + c,
// and a bump map differentiates its texture numerically over a pixel footprint (material.rs:46-87: (displace(p + du dpdu) - displace(p)) / du
// with du ~ 1e-4 at 1080p).  Rounding that differs from the CPU's per evaluation is noise of ~1e-7 / du in the slope -- enough to turn the
// shading normal by 1e-4 rad and flip a grazing same_hemisphere test once in 10^4 samples (T1 at 1920x1080: relMSE 1.4e-3 from eight
// pixels).  With separately rounded operations the three evaluations are the SAME function of their (slightly different) arguments and the
// differences cancel to first order: relMSE back to the 1e-6 range (tests/test_gpu_textures.py::test_textured_scene_at_bench_resolution).
PB_D float mul_rn(float a, float b) { return __fmul_rn(a, b); }
PB_D float add_rn(float a, float b) { return __fadd_rn(a, b); }
PB_D float noise_grad(int x, int y, int z, float dx, float dy, float dz) {
    int h = c_noise_perm[c_noise_perm[c_noise_perm[x] + y] + z] & 15;
    float u = (h < 8 || h == 12 || h == 13) ? dx : dy;
    float v = (h < 4 || h == 12 || h == 13) ? dy : dz;
    return ((h & 1) ? -u : u) + ((h & 2) ? -v : v);
}
PB_D float noise_weight(float t) {
    const float t3 = mul_rn(mul_rn(t, t), t), t4 = mul_rn(t3, t);
    return add_rn(add_rn(mul_rn(mul_rn(6.0f, t4), t), -mul_rn(15.0f, t4)), mul_rn(10.0f, t3));
}
PB_D float lerpf(float t, float a, float b) { return add_rn(mul_rn(a, add_rn(1.0f, -t)), mul_rn(b, t)); }
// `x.floor() as usize`: NaN / negative -> 0, saturating above (what Rust's cast does); returns the float the reference subtracts
PB_D unsigned long long f2u_sat(float f) { return (f != f || f <= 0.0f) ? 0ull : (f >= 1.8446744e19f ? 0xffffffffffffffffull : (unsigned long long)f); }
// cell = `floor(v) as usize`: (the float the reference subtracts, cell & 255).  Below 2^31 the cast is exact in 32 bits; above, a float is
// a multiple of 256 (low bits 0) until the cast saturates at 2^64 (all ones: 255)
PB_D float noise_cell(float f, int* cell) {
    if (!(f > 0.0f)) { *cell = 0; return 0.0f; }                  // NaN / negative saturate to 0
    if (f < 2147483648.0f) { const int i = (int)f; *cell = i & 255; return (float)i; }
    if (f >= 1.8446744e19f) { *cell = 255; return 1.8446744e19f; }
    *cell = 0;
    return f;
}
static __device__ __noinline__ float noise3(float x, float y, float z) {
    int ix, iy, iz;
    const float fx = noise_cell(floorf(x), &ix), fy = noise_cell(floorf(y), &iy), fz = noise_cell(floorf(z), &iz);
    const float dx = add_rn(x, -fx), dy = add_rn(y, -fy), dz = add_rn(z, -fz);
    float w000 = noise_grad(ix, iy, iz, dx, dy, dz), w100 = noise_grad(ix + 1, iy, iz, dx - 1.0f, dy, dz);
    float w010 = noise_grad(ix, iy + 1, iz, dx, dy - 1.0f, dz), w110 = noise_grad(ix + 1, iy + 1, iz, dx - 1.0f, dy - 1.0f, dz);
    float w001 = noise_grad(ix, iy, iz + 1, dx, dy, dz - 1.0f), w101 = noise_grad(ix + 1, iy, iz + 1, dx - 1.0f, dy, dz - 1.0f);
    float w011 = noise_grad(ix, iy + 1, iz + 1, dx, dy - 1.0f, dz - 1.0f), w111 = noise_grad(ix + 1, iy + 1, iz + 1, dx - 1.0f, dy - 1.0f, dz - 1.0f);
    float wx = noise_weight(dx), wy = noise_weight(dy), wz = noise_weight(dz);
    float x00 = lerpf(wx, w000, w100), x10 = lerpf(wx, w010, w110), x01 = lerpf(wx, w001, w101), x11 = lerpf(wx, w011, w111);
    return lerpf(wz, lerpf(wy, x00, x10), lerpf(wy, x01, x11));
}
PB_D float smooth_step(float mn, float mx, float v) {
    const float t = clampf(__fdiv_rn(add_rn(v, -mn), add_rn(mx, -mn)), 0.0f, 1.0f);
    return mul_rn(mul_rn(t, t), add_rn(mul_rn(-2.0f, t), 3.0f));
}
PB_D float fbm(f3 p, f3 dpdx, f3 dpdy, float omega, int max_octaves) {  // texture.rs:384-405
    float l2 = fmaxf(len2(dpdx), len2(dpdy));
    float n = clampf(-1.0f - 0.5f * (logf(l2) * 1.442695040888963387004650940071f), 0.0f, (float)max_octaves);
    int nint = (int)f2u_sat(floorf(n));
    float sum = 0.0f, lambda = 1.0f, o = 1.0f;
    for (int i = 0; i < nint; ++i) {
        sum = add_rn(sum, mul_rn(o, noise3(mul_rn(p.x, lambda), mul_rn(p.y, lambda), mul_rn(p.z, lambda))));
        lambda = mul_rn(lambda, 1.99f); o = mul_rn(o, omega);
    }
    float npartial = n - (float)nint;
    sum = add_rn(sum, mul_rn(mul_rn(o, smooth_step(0.3f, 0.7f, npartial)), noise3(mul_rn(p.x, lambda), mul_rn(p.y, lambda), mul_rn(p.z, lambda))));
    return sum;
}
PB_D float turbulence(f3 p, f3 dpdx, f3 dpdy, float omega, int max_octaves) {  // texture.rs:407-437 (`o + |noise|` as written there)
    float l2 = fmaxf(len2(dpdx), len2(dpdy));
    float n = clampf(-1.0f - 0.5f * log2f(l2), 0.0f, (float)max_octaves);
    int nint = (int)f2u_sat(floorf(n));
    float sum = 0.0f, lambda = 1.0f, o = 1.0f;
    for (int i = 0; i < nint; ++i) {
        sum = add_rn(sum, add_rn(o, fabsf(noise3(mul_rn(p.x, lambda), mul_rn(p.y, lambda), mul_rn(p.z, lambda)))));
        lambda = mul_rn(lambda, 1.99f); o = mul_rn(o, omega);
    }
    float npartial = n - (float)nint;
    sum = add_rn(sum, add_rn(o, lerpf(smooth_step(0.3f, 0.7f, npartial), 0.2f, fabsf(noise3(mul_rn(p.x, lambda), mul_rn(p.y, lambda), mul_rn(p.z, lambda))))));
    for (int i = nint; i < max_octaves; ++i) { sum = add_rn(sum, mul_rn(o, 0.2f)); o = mul_rn(o, omega); }
    return sum;
}

// ---- MIPMap, core/mipmap.rs:202-391 -----------------------------------------------------------------------------------------------
PB_D int mip_res(uint32_t r, int l) { return max(1, (int)(r >> l)); }
// One level of a pyramid: resolution (powers of two), float offset of its first texel
struct MipLevel { int u, v; uint32_t off; };
PB_D MipLevel mip_level(const pbrt_b200_mipmap& m, int level) {
    MipLevel L;
    uint32_t off = 0;
    for (int i = 0; i < level; ++i) off += (uint32_t)mip_res(m.width, i) * (uint32_t)mip_res(m.height, i);
    L.u = mip_res(m.width, level); L.v = mip_res(m.height, level); L.off = off;
    return L;
}
// float -> texel coordinate (the reference's `as isize`), kept in 32 bits: coordinates beyond +-2^30 texels -- a texture coordinate of
// ten thousand image widths at the finest level -- are clamped there instead of wrapping exactly
PB_D int mip_coord(float f) { return (int)fminf(fmaxf(f, -1073741824.0f), 1073741824.0f); }
PB_D rgb mip_texel(const pbrt_b200_mipmap& m, const MipLevel& L, int s, int t) {  // :301-321
    if (m.wrap == PBRT_B200_WRAP_REPEAT) { s &= L.u - 1; t &= L.v - 1; }  // mod_ of a power of two (two's complement: also for negatives)
    else if (m.wrap == PBRT_B200_WRAP_CLAMP) { s = min(max(s, 0), L.u - 1); t = min(max(t, 0), L.v - 1); }
    else if (s < 0 || s >= L.u || t < 0 || t >= L.v) return rgb(0.0f);
    const float* p = m.texels + (size_t)(L.off + (uint32_t)t * (uint32_t)L.u + (uint32_t)s) * m.channels;
    return m.channels == 1 ? rgb(__ldg(p)) : rgb(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}
PB_D rgb mip_triangle(const pbrt_b200_mipmap& m, int level, float2 st) {  // :323-335
    level = min(max(level, 0), (int)m.n_levels - 1);
    const MipLevel L = mip_level(m, level);
    const float s = add_rn(mul_rn(st.x, (float)L.u), -0.5f), t = add_rn(mul_rn(st.y, (float)L.v), -0.5f);
    const float fs = floorf(s), ft = floorf(t);
    int s0 = mip_coord(fs), t0 = mip_coord(ft);
    const float ds = add_rn(s, -fs), dt = add_rn(t, -ft);
    // separately rounded, like the noise functions: image maps are bump maps too
    const rgb a = mip_texel(m, L, s0 + 1, t0 + 1), b = mip_texel(m, L, s0 + 1, t0), c = mip_texel(m, L, s0, t0 + 1), d = mip_texel(m, L, s0, t0);
    const float w1 = mul_rn(ds, dt), w2 = mul_rn(ds, add_rn(1.0f, -dt)), w3 = mul_rn(add_rn(1.0f, -ds), dt), w4 = mul_rn(add_rn(1.0f, -ds), add_rn(1.0f, -dt));
    return rgb(add_rn(add_rn(add_rn(mul_rn(d.r, w4), mul_rn(c.r, w3)), mul_rn(b.r, w2)), mul_rn(a.r, w1)),
               add_rn(add_rn(add_rn(mul_rn(d.g, w4), mul_rn(c.g, w3)), mul_rn(b.g, w2)), mul_rn(a.g, w1)),
               add_rn(add_rn(add_rn(mul_rn(d.b, w4), mul_rn(c.b, w3)), mul_rn(b.b, w2)), mul_rn(a.b, w1)));
}
PB_D rgb mip_ewa(const pbrt_b200_mipmap& m, int level, float2 st, float2 d0, float2 d1) {  // :337-391
    if (level >= (int)m.n_levels) return mip_texel(m, mip_level(m, (int)m.n_levels - 1), 0, 0);
    const MipLevel L = mip_level(m, level);
    const float ur = (float)L.u, vr = (float)L.v;
    st.x = st.x * ur - 0.5f; st.y = st.y * vr - 0.5f;
    d0.x *= ur; d0.y *= vr; d1.x *= ur; d1.y *= vr;
    float A = d0.y * d0.y + d1.y * d1.y + 1.0f;
    float B = -2.0f * (d0.x * d0.y + d1.x * d1.y);
    float C = d0.x * d0.x + d1.x * d1.x + 1.0f;
    float invf = 1.0f / (A * C - B * B * 0.25f);
    A *= invf; B *= invf; C *= invf;
    float det = -B * B + 4.0f * A * C;
    float idet = 1.0f / det;
    float usq = sqrtf(det * C), vsq = sqrtf(det * A);
    int s0 = mip_coord(ceilf(st.x - 2.0f * idet * usq)), s1 = mip_coord(floorf(st.x + 2.0f * idet * usq));
    int t0 = mip_coord(ceilf(st.y - 2.0f * idet * vsq)), t1 = mip_coord(floorf(st.y + 2.0f * idet * vsq));
    // The level is chosen so that the minor axis is about a texel and the major axis at most max_anisotropy times that: tens of texels.
    // Non-finite differentials (overflowed derivatives) would make the reference scan 2^64 texels; a kernel must not: PB_EWA_MAX_SPAN
    // texels per axis, centred on the lookup point, is the most one lookup scans
    if (s1 - s0 > PB_EWA_MAX_SPAN) { const int c = mip_coord(st.x); s0 = c - PB_EWA_MAX_SPAN / 2; s1 = c + PB_EWA_MAX_SPAN / 2; }
    if (t1 - t0 > PB_EWA_MAX_SPAN) { const int c = mip_coord(st.y); t0 = c - PB_EWA_MAX_SPAN / 2; t1 = c + PB_EWA_MAX_SPAN / 2; }
    rgb sum(0.0f);
    float sum_w = 0.0f;
    for (int it = t0; it <= t1; ++it) {
        float tt = (float)it - st.y;
        for (int is = s0; is <= s1; ++is) {
            float ss = (float)is - st.x;
            float r2 = A * ss * ss + B * ss * tt + C * tt * tt;
            if (r2 < 1.0f) {
                int index = min((int)f2u_sat(r2 * 128.0f), 127);
                float wgt = expf(-2.0f * ((float)index / 127.0f)) - expf(-2.0f);  // WEIGHT_LUT, :40-50
                sum = sum + mip_texel(m, L, is, it) * wgt;
                sum_w += wgt;
            }
        }
    }
    return sum / sum_w;
}
static __device__ __noinline__ rgb mip_lookup(const pbrt_b200_mipmap* mp, float2 st, float2 d0, float2 d1) {  // lookup2 :228-269, lookup :202-226
    const pbrt_b200_mipmap& m = *mp;
    const int L = (int)m.n_levels;
    if (m.do_trilinear) {
        float width = fmaxf(fmaxf(fabsf(d0.x), fabsf(d0.y)), fmaxf(fabsf(d1.x), fabsf(d1.y)));
        float level = (float)(L - 1) + log2f(fmaxf(width, 1.0e-8f));
        if (level < 0.0f) return mip_triangle(m, 0, st);
        if (level >= (float)(L - 1)) return mip_texel(m, mip_level(m, L - 1), 0, 0);
        float il = floorf(level), delta = level - il;
        return mip_triangle(m, (int)il, st) * (1.0f - delta) + mip_triangle(m, (int)il + 1, st) * delta;
    }
    if (d0.x * d0.x + d0.y * d0.y < d1.x * d1.x + d1.y * d1.y) { float2 t = d0; d0 = d1; d1 = t; }
    float majorl = sqrtf(d0.x * d0.x + d0.y * d0.y), minorl = sqrtf(d1.x * d1.x + d1.y * d1.y);
    if (minorl * m.max_anisotropy < majorl && minorl > 0.0f) {
        float scale = majorl / (minorl * m.max_anisotropy);
        d1.x *= scale; d1.y *= scale;
        minorl *= scale;
    }
    if (minorl == 0.0f) return mip_triangle(m, 0, st);
    float lod = fmaxf((float)L - 1.0f + log2f(minorl), 0.0f);
    int ilod = (int)f2u_sat(floorf(lod));
    float dl = lod - (float)ilod;
    return mip_ewa(m, ilod, st, d0, d1) * (1.0f - dl) + mip_ewa(m, ilod + 1, st, d0, d1) * dl;
}

// ---- mappings, core/texture.rs:122-317 ---------------------------------------------------------------------------------------
struct TexCtx {  // what Texture::evaluate reads of the SurfaceInteraction
    f3 p, dpdx, dpdy;
    float2 uv;
    float dudx, dvdx, dudy, dvdy;
};
PB_D float2 map_sphere(const float* w2t, f3 p) {
    f3 v = normalize(xf_point(w2t, p));
    float theta = acosf(clampf(v.z, -1.0f, 1.0f));
    float phi = atan2f(v.y, v.x);
    if (phi < 0.0f) phi += 2.0f * PB_PI;
    return make_float2(theta * PB_INV_PI, phi * 0.15915494309189533577f);
}
PB_D float2 map_cylinder(const float* w2t, f3 p) {
    f3 v = normalize(xf_point(w2t, p));
    return make_float2(PB_PI + atan2f(v.y, v.x) * 0.15915494309189533577f, v.z);
}
PB_D void fix_seam(float2& d) {
    if (d.y > 0.5f) d.y = 1.0f - d.y;
    else if (d.y < -0.5f) d.y = -(d.y + 1.0f);
}
PB_D float2 map2d(const pbrt_b200_texnode& n, const TexCtx& c, float2* dstdx, float2* dstdy) {
    if (n.mapping == PBRT_B200_MAP_UV) {
        const float su = n.m[0], sv = n.m[1];
        *dstdx = make_float2(su * c.dudx, sv * c.dvdx);
        *dstdy = make_float2(su * c.dudy, sv * c.dvdy);
        return make_float2(su * c.uv.x + n.m[2], sv * c.uv.y + n.m[3]);
    }
    if (n.mapping == PBRT_B200_MAP_PLANAR) {
        const f3 vs(n.m[0], n.m[1], n.m[2]), vt(n.m[3], n.m[4], n.m[5]);
        *dstdx = make_float2(dot(c.dpdx, vs), dot(c.dpdx, vt));
        *dstdy = make_float2(dot(c.dpdy, vs), dot(c.dpdy, vt));
        return make_float2(n.m[6] + dot(c.p, vs), n.m[7] + dot(c.p, vt));
    }
    const bool sph = n.mapping == PBRT_B200_MAP_SPHERICAL;
    const float delta = 0.1f, inv = 1.0f / delta;
    const float2 st = sph ? map_sphere(n.m, c.p) : map_cylinder(n.m, c.p);
    const f3 px = c.p + c.dpdx * delta, py = c.p + c.dpdy * delta;
    const float2 sx = sph ? map_sphere(n.m, px) : map_cylinder(n.m, px), sy = sph ? map_sphere(n.m, py) : map_cylinder(n.m, py);
    *dstdx = make_float2((sx.x - st.x) * inv, (sx.y - st.y) * inv);
    *dstdy = make_float2((sy.x - st.x) * inv, (sy.y - st.y) * inv);
    fix_seam(*dstdx);
    fix_seam(*dstdy);
    return st;
}
PB_D bool checker_even(long long a) { return a % 2 == 0; }

// Texture::evaluate over a postfix program (include/pbrt_b200.h)
static __device__ __noinline__ rgb tex_eval(const DevScene* sp, pbrt_b200_texref ref, const TexCtx* cp) {
    const DevScene& s = *sp;
    const TexCtx& c = *cp;
    rgb stack[PB_TEX_STACK];
    int top = 0;
    for (uint32_t k = 0; k < ref.count; ++k) {
        const pbrt_b200_texnode& n = s.textures[ref.first + k];
        float2 dstdx, dstdy;
        rgb out(0.0f);
        switch (n.kind) {
            case PBRT_B200_TEX_CONSTANT: out = rgb(n.v[0], n.v[1], n.v[2]); break;
            case PBRT_B200_TEX_SCALE: { rgb b = stack[--top], a = stack[--top]; out = a * b; break; }
            case PBRT_B200_TEX_MIX: {
                float amt = stack[--top].r;
                rgb t2 = stack[--top], t1 = stack[--top];
                out = t1 * (1.0f - amt) + t2 * amt;
                break;
            }
            case PBRT_B200_TEX_BILERP: {
                float2 st = map2d(n, c, &dstdx, &dstdy);
                out = rgb3(n.v) * (1.0f - st.y) * (1.0f - st.x) + rgb3(n.v + 3) * (1.0f - st.x) * st.y + rgb3(n.v + 6) * (1.0f - st.y) * st.x +
                      rgb3(n.v + 9) * st.y * st.x;
                break;
            }
            case PBRT_B200_TEX_IMAGEMAP: {
                float2 st = map2d(n, c, &dstdx, &dstdy);
                out = mip_lookup(s.mipmaps + n.image, st, dstdx, dstdy);
                break;
            }
            case PBRT_B200_TEX_UV: {
                float2 st = map2d(n, c, &dstdx, &dstdy);
                out = rgb(st.x - floorf(st.x), st.y - floorf(st.y), 0.0f);
                break;
            }
            case PBRT_B200_TEX_CHECKERBOARD2D: {
                rgb t2 = stack[--top], t1 = stack[--top];
                float2 st = map2d(n, c, &dstdx, &dstdy);
                const bool first = checker_even((long long)floorf(st.x) + (long long)floorf(st.y));
                out = first ? t1 : t2;
                if (n.flags & PBRT_B200_TEX_AA_CLOSEDFORM) {
                    float ds = fmaxf(fabsf(dstdx.x), fabsf(dstdy.x)), dt = fmaxf(fabsf(dstdx.y), fabsf(dstdy.y));
                    float s0 = st.x - ds, s1 = st.x + ds, t0 = st.y - dt, t1v = st.y + dt;
                    if (!(floorf(s0) == floorf(s1) && floorf(t0) == floorf(t1v))) {
#define PB_BUMPINT(x) (floorf((x) / 2.0f) + 2.0f * fmaxf((x) / 2.0f - floorf((x) / 2.0f) - 0.5f, 0.0f))
                        float sint = (PB_BUMPINT(s1) - PB_BUMPINT(s0)) / (2.0f * ds), tint = (PB_BUMPINT(t1v) - PB_BUMPINT(t0)) / (2.0f * dt);
#undef PB_BUMPINT
                        float area2 = sint * tint - 2.0f * sint * tint;  // checkerboard.rs:62 as written
                        if (ds > 1.0f || dt > 1.0f) area2 = 0.5f;
                        out = t1 * (1.0f - area2) + t2 * area2;
                    }
                }
                break;
            }
            case PBRT_B200_TEX_CHECKERBOARD3D: {
                rgb t2 = stack[--top], t1 = stack[--top];
                f3 p = xf_point(n.m, c.p);
                out = checker_even((long long)floorf(p.x) + (long long)floorf(p.y) + (long long)floorf(p.z)) ? t1 : t2;
                break;
            }
            case PBRT_B200_TEX_DOTS: {
                rgb inside = stack[--top], outside = stack[--top];
                float2 st = map2d(n, c, &dstdx, &dstdy);
                float scell = (float)f2u_sat(floorf(st.x + 0.5f)), tcell = (float)f2u_sat(floorf(st.y + 0.5f));
                out = outside;
                if (noise3(scell + 0.5f, tcell + 0.5f, 0.5f) > 0.0f) {
                    const float radius = 0.35f, max_shift = 0.5f - radius;
                    float scenter = scell + max_shift * noise3(scell + 1.5f, tcell + 2.8f, 0.5f);
                    float tcenter = tcell + max_shift * noise3(scell + 4.5f, tcell + 9.8f, 0.5f);
                    float dx = st.x - scenter, dy = st.y - tcenter;
                    if (dx * dx + dy * dy < radius * radius) out = inside;
                }
                break;
            }
            case PBRT_B200_TEX_FBM: case PBRT_B200_TEX_WRINKLED: {
                f3 p = xf_point(n.m, c.p), dx = xf_vector(n.m, c.dpdx), dy = xf_vector(n.m, c.dpdy);
                out = rgb(n.kind == PBRT_B200_TEX_FBM ? fbm(p, dx, dy, n.v[0], (int)n.v[1]) : turbulence(p, dx, dy, n.v[0], (int)n.v[1]));
                break;
            }
            case PBRT_B200_TEX_MARBLE: {  // marble.rs:40-76
                const float C[9][3] = {{0.58f, 0.58f, 0.6f}, {0.58f, 0.58f, 0.6f}, {0.58f, 0.58f, 0.6f}, {0.5f, 0.5f, 0.5f}, {0.6f, 0.59f, 0.58f},
                                       {0.58f, 0.58f, 0.6f}, {0.58f, 0.58f, 0.6f}, {0.2f, 0.2f, 0.33f}, {0.58f, 0.58f, 0.6f}};
                f3 p = xf_point(n.m, c.p), dx = xf_vector(n.m, c.dpdx), dy = xf_vector(n.m, c.dpdy);
                const float scale = n.v[2], variation = n.v[3];
                p = p * scale;
                float marble = p.y + variation * fbm(p, dx * scale, dy * scale, n.v[0], (int)n.v[1]);
                float t = 0.5f + 0.5f * sinf(marble);
                int first = (int)min(5ull, f2u_sat(floorf(t * 6.0f)));
                rgb c0 = rgb3(C[first]), c1 = rgb3(C[first + 1]), c2 = rgb3(C[first + 2]), c3 = rgb3(C[first + 3]);
                rgb s0 = c0 * (1.0f - t) + c1 * t, s1 = c1 * (1.0f - t) + c2 * t, s2 = c2 * (1.0f - t) + c3 * t;
                s0 = s0 * (1.0f - t) + s1 * t;
                s1 = s1 * (1.0f - t) + s2 * t;
                out = (s0 * (1.0f - t) + s1 * t) * 1.5f;
                break;
            }
            case PBRT_B200_TEX_WINDY: {
                f3 p = xf_point(n.m, c.p), dx = xf_vector(n.m, c.dpdx), dy = xf_vector(n.m, c.dpdy);
                float wstrength = fbm(p * 0.1f, dx * 0.1f, dy * 0.1f, 0.5f, 3);
                float wheight = fbm(p, dx, dy, 0.5f, 6);
                out = rgb(fabsf(wstrength) * wheight);
                break;
            }
            default: break;
        }
        stack[top++] = out;
    }
    return top > 0 ? stack[top - 1] : rgb(0.0f);
}

PB_D TexCtx tex_ctx(const Surf& si, const SurfX& sx) {
    TexCtx c;
    c.p = si.p; c.dpdx = sx.dpdx; c.dpdy = sx.dpdy; c.uv = sx.uv;
    c.dudx = sx.dudx; c.dvdx = sx.dvdx; c.dudy = sx.dudy; c.dvdy = sx.dvdy;
    return c;
}

// Material::bump, core/material.rs:46-87
PB_D void bump_map(const DevScene* sp, pbrt_b200_texref d, Surf& si, SurfX& sx) {
    TexCtx c = tex_ctx(si, sx);
    float du = 0.5f * (fabsf(sx.dudx) + fabsf(sx.dudy));
    if (du == 0.0f) du = 0.0005f;
    TexCtx e = c;
    e.p = si.p + si.sh_dpdu * du;
    e.uv = make_float2(sx.uv.x + du, sx.uv.y + 0.0f);
    const float udisplace = tex_eval(sp, d, &e).r;
    float dv = 0.5f * (fabsf(sx.dvdx) + fabsf(sx.dvdy));
    if (dv == 0.0f) dv = 0.0005f;
    e.p = si.p + sx.sh_dpdv * dv;
    e.uv = make_float2(sx.uv.x + 0.0f, sx.uv.y + dv);
    const float vdisplace = tex_eval(sp, d, &e).r;
    const float displace = tex_eval(sp, d, &c).r;
    f3 dpdu = si.sh_dpdu + si.sh_n * ((udisplace - displace) / du) + sx.sh_dndu * displace;
    f3 dpdv = sx.sh_dpdv + si.sh_n * ((vdisplace - displace) / dv) + sx.sh_dndv * displace;
    set_shading_geometry(si, sx, dpdu, dpdv, sx.sh_dndu, sx.sh_dndv, false);
}

// compute_scattering_functions of a `textured` material row: differentials, bump map, parameter evaluation, then the lobes
// (materials/{matte,plastic,mirror,glass,metal}.rs through material_bsdf; uber.rs:41-112; substrate.rs:34-62)
template <bool MULTI, class B>
static __device__ __noinline__ void material_bsdf_tex(const DevScene* sp, int mat, Surf* sip, SurfX* sxp, const RayDiff* rdp, B* bp) {
    const DevScene& s = *sp;
    Surf& si = *sip;
    SurfX& sx = *sxp;
    B& b = *bp;
    compute_differentials(si, sx, *rdp);
    const pbrt_b200_material m = s.materials[mat];
    const pbrt_b200_material_ext& x = s.material_ext[mat];
    if (x.bump.count) bump_map(sp, x.bump, si, sx);
    const TexCtx c = tex_ctx(si, sx);
    rgb S[5];
    float F[3];
    for (int k = 0; k < 5; ++k) S[k] = x.s_tex[k].count ? tex_eval(sp, x.s_tex[k], &c) : rgb3(x.s_const[k]);
    for (int k = 0; k < 3; ++k) F[k] = x.f_tex[k].count ? tex_eval(sp, x.f_tex[k], &c).r : x.f_const[k];
    b.valid = false; b.n = 0;
    if (m.type <= PBRT_B200_MAT_METAL) {
        pbrt_b200_material cm = m;
        cm.a[0] = S[0].r; cm.a[1] = S[0].g; cm.a[2] = S[0].b; cm.b[0] = S[1].r; cm.b[1] = S[1].g; cm.b[2] = S[1].b;
        cm.f0 = F[0]; cm.f1 = F[1]; cm.f2 = F[2];
        material_bsdf<-1, MULTI>(cm, si, b);
        return;
    }
    if (m.type == PBRT_B200_MAT_UBER) {
        const float e = F[2];
        const rgb op = rgb_clamp0(S[4]);
        const rgb t = rgb_clamp0(rgb(1.0f) - op);  // (-op + 1).clamps(0, inf)
        if (!is_black(t)) {
            bsdf_init(b, si, 1.0f);
            Lobe& l = b.lobe[b.n++]; l.kind = LOBE_SPEC_TRANS; l.type = BX_TRANSMISSION | BX_SPECULAR; l.c0 = t; l.p0 = 1.0f; l.p1 = 1.0f;
        } else bsdf_init(b, si, e);
        const rgb kd = op * rgb_clamp0(S[0]);
        if (!is_black(kd)) { Lobe& l = b.lobe[b.n++]; l.kind = LOBE_LAMBERT; l.type = BX_REFLECTION | BX_DIFFUSE; l.c0 = kd; }
        const rgb ks = op * rgb_clamp0(S[1]);
        if (!is_black(ks)) {
            float ru = F[0], rv = F[1];
            if (m.remap_roughness) { ru = roughness_to_alpha(ru); rv = roughness_to_alpha(rv); }
            Lobe& l = b.lobe[b.n++]; l.kind = LOBE_MICRO_REFL_DIEL; l.type = BX_REFLECTION | BX_GLOSSY; l.c0 = ks; l.p0 = 1.0f; l.p1 = e; l.tr = tr_make(ru, rv);
        }
        const rgb kr = op * rgb_clamp0(S[2]);
        if (!is_black(kr)) { Lobe& l = b.lobe[b.n++]; l.kind = LOBE_SPEC_REFL_DIEL; l.type = BX_REFLECTION | BX_SPECULAR; l.c0 = kr; l.p0 = 1.0f; l.p1 = e; }
        const rgb kt = op * rgb_clamp0(S[3]);
        if (!is_black(kt)) { Lobe& l = b.lobe[b.n++]; l.kind = LOBE_SPEC_TRANS; l.type = BX_TRANSMISSION | BX_SPECULAR; l.c0 = kt; l.p0 = 1.0f; l.p1 = e; }
    } else if (m.type == PBRT_B200_MAT_SUBSTRATE) {
        const rgb d = rgb_clamp0(S[0]), sp_ = rgb_clamp0(S[1]);
        if (!is_black(d) || !is_black(sp_)) {
            bsdf_init(b, si, 1.0f);
            float ru = F[0], rv = F[1];
            if (m.remap_roughness) { ru = roughness_to_alpha(ru); rv = roughness_to_alpha(rv); }
            Lobe& l = b.lobe[b.n++]; l.kind = LOBE_FRESNEL_BLEND; l.type = BX_REFLECTION | BX_GLOSSY; l.c0 = d; l.c1 = sp_; l.tr = tr_make(ru, rv);
        }
    }
}

// SamplerIntegrator::specular_reflect / specular_transmit's ray differentials (core/integrator.rs:427-452, 476-513)
PB_D RayDiff specular_differentials(const Surf& si, const SurfX& sx, const RayDiff& in, f3 wo, f3 wi, float bsdf_eta, bool transmit) {
    RayDiff o;
    o.has = in.has;
    if (!in.has) return o;
    const f3 ns = si.sh_n;
    o.rxo = si.p + sx.dpdx; o.ryo = si.p + sx.dpdy;
    const f3 dndx = sx.sh_dndu * sx.dudx + sx.sh_dndv * sx.dvdx, dndy = sx.sh_dndu * sx.dudy + sx.sh_dndv * sx.dvdy;
    const f3 dwodx = -in.rxd - wo, dwody = -in.ryd - wo;
    const float ddndx = dot(dwodx, ns) + dot(wo, dndx), ddndy = dot(dwody, ns) + dot(wo, dndy);
    if (!transmit) {
        o.rxd = wi - dwodx + (dndx * dot(wo, ns) + ns * ddndx) * 2.0f;
        o.ryd = wi - dwody + (dndy * dot(wo, ns) + ns * ddndy) * 2.0f;
    } else {
        float eta = bsdf_eta;
        const f3 w = -wo;
        if (dot(wo, ns) < 0.0f) eta = 1.0f / eta;  // eta inverted, ns not negated: as the reference has it (:493-497)
        const float mu = eta * dot(w, ns) - dot(wi, ns);
        const float dmudx = (eta - (eta * eta * dot(w, ns)) / dot(wi, ns)) * ddndx;
        const float dmudy = (eta - (eta * eta * dot(w, ns)) / dot(wi, ns)) * ddndy;
        o.rxd = wi + dwodx * eta - (dndx * mu + ns * dmudx);
        o.ryd = wi + dwody * eta - (dndy * mu + ns * dmudy);
    }
    return o;
}

// three float4 per ray: {rxo, rxd.x} {rxd.yz, ryo.xy} {ryo.z, ryd}
PB_D void store_diff(float4* a, size_t i, const RayDiff& d) {
    a[3 * i] = make_float4(d.rxo.x, d.rxo.y, d.rxo.z, d.rxd.x);
    a[3 * i + 1] = make_float4(d.rxd.y, d.rxd.z, d.ryo.x, d.ryo.y);
    a[3 * i + 2] = make_float4(d.ryo.z, d.ryd.x, d.ryd.y, d.ryd.z);
}
PB_D RayDiff load_diff(const float4* a, size_t i, bool has) {
    RayDiff d;
    d.has = has;
    if (has) {
        const float4 q0 = a[3 * i], q1 = a[3 * i + 1], q2 = a[3 * i + 2];
        d.rxo = f3(q0.x, q0.y, q0.z); d.rxd = f3(q0.w, q1.x, q1.y); d.ryo = f3(q1.z, q1.w, q2.x); d.ryd = f3(q2.y, q2.z, q2.w);
    }
    return d;
}

}  // namespace pb
