// TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see oracle_math.hpp header).
// oracle_texture.hpp: textures, MIPMap lookups, ray differentials, bump mapping and the texture-parameterised materials (incl. uber and
// substrate) of SURVEY §8 f3, restated from the reference:
//   src/core/texture.rs (mappings :122-317, noise / fbm / turbulence :330-431), src/textures/*.rs, src/core/mipmap.rs:202-374,
//   src/core/interaction.rs:269-342 (compute_differentials), src/core/material.rs:46-87 (bump),
//   src/materials/{matte,plastic,mirror,glass,metal,uber,substrate}.rs.
// Parity unpinned by the reference's tests (it holds none for textures); pinned by tests/test_oracle_textures.py through closed forms
// (checkerboard parity, MIP level of a constant image, bilinear interpolation of a ramp, noise range / lattice zeros).
#pragma once
#include "oracle_reflection.hpp"

namespace orc {

// ---- Perlin noise, texture.rs:25-67,330-383 -----------------------------------------------------------------------------
static const int NOISE_PERM[512] = {
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

inline Float noise_grad(int x, int y, int z, Float dx, Float dy, Float dz) {  // texture.rs:368-375
    int h = NOISE_PERM[NOISE_PERM[NOISE_PERM[x] + y] + z];
    h &= 15;
    Float u = (h < 8 || h == 12 || h == 13) ? dx : dy;
    Float v = (h < 4 || h == 12 || h == 13) ? dy : dz;
    return ((h & 1) ? -u : u) + ((h & 2) ? -v : v);
}
inline Float noise_weight(Float t) {  // :377-382
    Float t3 = t * t * t, t4 = t3 * t;
    return 6.0f * t4 * t - 15.0f * t4 + 10.0f * t3;
}
// texture.rs:330-362.  `x.floor() as usize` saturates: a negative coordinate lands in cell 0 with a negative offset (kept).
inline Float noise(Float x, Float y, Float z) {
    uint64_t ixu = f2u_sat(std::floor(x)), iyu = f2u_sat(std::floor(y)), izu = f2u_sat(std::floor(z));
    Float dx = x - (Float)ixu, dy = y - (Float)iyu, dz = z - (Float)izu;
    int ix = (int)(ixu & 255u), iy = (int)(iyu & 255u), iz = (int)(izu & 255u);
    Float w000 = noise_grad(ix, iy, iz, dx, dy, dz);
    Float w100 = noise_grad(ix + 1, iy, iz, dx - 1.0f, dy, dz);
    Float w010 = noise_grad(ix, iy + 1, iz, dx, dy - 1.0f, dz);
    Float w110 = noise_grad(ix + 1, iy + 1, iz, dx - 1.0f, dy - 1.0f, dz);
    Float w001 = noise_grad(ix, iy, iz + 1, dx, dy, dz - 1.0f);
    Float w101 = noise_grad(ix + 1, iy, iz + 1, dx - 1.0f, dy, dz - 1.0f);
    Float w011 = noise_grad(ix, iy + 1, iz + 1, dx, dy - 1.0f, dz - 1.0f);
    Float w111 = noise_grad(ix + 1, iy + 1, iz + 1, dx - 1.0f, dy - 1.0f, dz - 1.0f);
    Float wx = noise_weight(dx), wy = noise_weight(dy), wz = noise_weight(dz);
    Float x00 = lerp(wx, w000, w100), x10 = lerp(wx, w010, w110), x01 = lerp(wx, w001, w101), x11 = lerp(wx, w011, w111);
    Float y0 = lerp(wy, x00, x10), y1 = lerp(wy, x01, x11);
    return lerp(wz, y0, y1);
}
inline Float noisep(V3 p) { return noise(p.x, p.y, p.z); }
inline Float smooth_step(Float mn, Float mx, Float value) {  // :427-431
    Float v = clamp((value - mn) / (mx - mn), 0.0f, 1.0f);
    return v * v * (-2.0f * v + 3.0f);
}
inline Float fbm(V3 p, V3 dpdx, V3 dpdy, Float omega, int max_octaves) {  // :384-405 (log2 = ln * inv_log2, pbrt.rs:114-118)
    Float len2 = std::fmax(length_squared(dpdx), length_squared(dpdy));
    Float n = clamp(-1.0f - 0.5f * (std::log(len2) * 1.442695040888963387004650940071f), 0.0f, (Float)max_octaves);
    int nint = (int)f2u_sat(std::floor(n));
    Float sum = 0.0f, lambda = 1.0f, o = 1.0f;
    for (int i = 0; i < nint; ++i) { sum += o * noisep(p * lambda); lambda *= 1.99f; o *= omega; }
    Float npartial = n - (Float)nint;
    sum += o * smooth_step(0.3f, 0.7f, npartial) * noisep(p * lambda);
    return sum;
}
inline Float turbulence(V3 p, V3 dpdx, V3 dpdy, Float omega, int max_octaves) {  // :407-437 (`sum += o + |noise|`: the reference's own slip, kept)
    Float len2 = std::fmax(length_squared(dpdx), length_squared(dpdy));
    Float n = clamp(-1.0f - 0.5f * std::log2(len2), 0.0f, (Float)max_octaves);
    int nint = (int)f2u_sat(std::floor(n));
    Float sum = 0.0f, lambda = 1.0f, o = 1.0f;
    for (int i = 0; i < nint; ++i) { sum += o + std::fabs(noisep(p * lambda)); lambda *= 1.99f; o *= omega; }
    Float npartial = n - (Float)nint;
    sum += o + lerp(smooth_step(0.3f, 0.7f, npartial), 0.2f, std::fabs(noisep(p * lambda)));
    for (int i = nint; i < max_octaves; ++i) { sum += o * 0.2f; o *= omega; }
    return sum;
}

// ---- MIPMap, mipmap.rs:202-374 ---------------------------------------------------------------------------------------------
struct MipView {
    const pbrt_b200_mipmap& m;
    explicit MipView(const pbrt_b200_mipmap& mm) : m(mm) {}
    int levels() const { return (int)m.n_levels; }
    int ures(int l) const { return std::max(1, (int)(m.width >> l)); }
    int vres(int l) const { return std::max(1, (int)(m.height >> l)); }
    const float* level_ptr(int l) const {
        size_t off = 0;
        for (int i = 0; i < l; ++i) off += (size_t)ures(i) * vres(i) * m.channels;
        return m.texels + off;
    }
    Spectrum texel(int level, int64_t s, int64_t t) const {  // :301-321
        int64_t u = ures(level), v = vres(level);
        if (m.wrap == PBRT_B200_WRAP_REPEAT) { s = ((s % u) + u) % u; t = ((t % v) + v) % v; }
        else if (m.wrap == PBRT_B200_WRAP_CLAMP) { s = clamp<int64_t>(s, 0, u - 1); t = clamp<int64_t>(t, 0, v - 1); }  // reference: clamp(s, 0, u), past the row
        else if (s < 0 || s >= u || t < 0 || t >= v) return Spectrum(0.0f);
        const float* p = level_ptr(level) + ((size_t)t * u + s) * m.channels;
        return m.channels == 1 ? Spectrum(p[0]) : Spectrum(p[0], p[1], p[2]);
    }
    Spectrum triangle(int level, P2 st) const {  // :323-335
        level = clamp(level, 0, levels() - 1);
        Float s = st.x * (Float)ures(level) - 0.5f, t = st.y * (Float)vres(level) - 0.5f;
        int64_t s0 = (int64_t)std::floor(s), t0 = (int64_t)std::floor(t);
        Float ds = s - (Float)s0, dt = t - (Float)t0;
        Spectrum tmp1 = texel(level, s0 + 1, t0 + 1) * (ds * dt);
        Spectrum tmp2 = texel(level, s0 + 1, t0) * (ds * (1.0f - dt));
        Spectrum tmp3 = texel(level, s0, t0 + 1) * ((1.0f - ds) * dt);
        Spectrum tmp4 = texel(level, s0, t0) * ((1.0f - ds) * (1.0f - dt));
        return tmp4 + tmp3 + tmp2 + tmp1;
    }
    Spectrum lookup(P2 st, Float width) const {  // :202-226
        Float level = (Float)(levels() - 1) + std::log2(std::fmax(width, 1.0e-8f));
        if (level < 0.0f) return triangle(0, st);
        if (level >= (Float)(levels() - 1)) return texel(levels() - 1, 0, 0);
        Float ilevel = std::floor(level), delta = level - ilevel;
        return triangle((int)ilevel, st) * (1.0f - delta) + triangle((int)ilevel + 1, st) * delta;
    }
    static Float weight_lut(int i) {  // :40-50
        Float r2 = (Float)i / 127.0f;
        return std::exp(-2.0f * r2) - std::exp(-2.0f);
    }
    Spectrum ewa(int level, P2 st, P2 dst0, P2 dst1) const {  // :337-391
        if (level >= levels()) return texel(levels() - 1, 0, 0);
        st.x = st.x * (Float)ures(level) - 0.5f; st.y = st.y * (Float)vres(level) - 0.5f;
        dst0.x *= (Float)ures(level); dst0.y *= (Float)vres(level);
        dst1.x *= (Float)ures(level); dst1.y *= (Float)vres(level);
        Float A = dst0.y * dst0.y + dst1.y * dst1.y + 1.0f;
        Float B = -2.0f * (dst0.x * dst0.y + dst1.x * dst1.y);
        Float C = dst0.x * dst0.x + dst1.x * dst1.x + 1.0f;
        Float invf = 1.0f / (A * C - B * B * 0.25f);
        A *= invf; B *= invf; C *= invf;
        Float det = -B * B + 4.0f * A * C;
        Float idet = 1.0f / det;
        Float usqrt = std::sqrt(det * C), vsqrt = std::sqrt(det * A);
        int64_t s0 = (int64_t)std::ceil(st.x - 2.0f * idet * usqrt), s1 = (int64_t)std::floor(st.x + 2.0f * idet * usqrt);
        int64_t t0 = (int64_t)std::ceil(st.y - 2.0f * idet * vsqrt), t1 = (int64_t)std::floor(st.y + 2.0f * idet * vsqrt);
        Spectrum sum(0.0f);
        Float sum_wts = 0.0f;
        for (int64_t it = t0; it <= t1; ++it) {
            Float tt = (Float)it - st.y;
            for (int64_t is = s0; is <= s1; ++is) {
                Float ss = (Float)is - st.x;
                Float r2 = A * ss * ss + B * ss * tt + C * tt * tt;
                if (r2 < 1.0f) {
                    int index = std::min((int)f2u_sat(r2 * 128.0f), 127);
                    Float weight = weight_lut(index);
                    sum += texel(level, is, it) * weight;
                    sum_wts += weight;
                }
            }
        }
        return sum / sum_wts;
    }
    Spectrum lookup2(P2 st, P2 dst0, P2 dst1) const {  // :228-269
        if (m.do_trilinear) {
            Float x = std::fmax(std::fabs(dst0.x), std::fabs(dst0.y)), y = std::fmax(std::fabs(dst1.x), std::fabs(dst1.y));
            return lookup(st, std::fmax(x, y));
        }
        if (dst0.x * dst0.x + dst0.y * dst0.y < dst1.x * dst1.x + dst1.y * dst1.y) std::swap(dst0, dst1);
        Float majorl = std::sqrt(dst0.x * dst0.x + dst0.y * dst0.y), minorl = std::sqrt(dst1.x * dst1.x + dst1.y * dst1.y);
        if (minorl * m.max_anisotropy < majorl && minorl > 0.0f) {
            Float scale = majorl / (minorl * m.max_anisotropy);
            dst1.x *= scale; dst1.y *= scale;
            minorl *= scale;
        }
        if (minorl == 0.0f) return triangle(0, st);
        Float lod = std::fmax((Float)levels() - 1.0f + std::log2(minorl), 0.0f);
        int ilod = (int)f2u_sat(std::floor(lod));
        Float d = lod - (Float)ilod;
        return ewa(ilod, st, dst0, dst1) * (1.0f - d) + ewa(ilod + 1, st, dst0, dst1) * d;
    }
};

// ---- texture mappings, texture.rs:122-317 ------------------------------------------------------------------------------------
inline M4 m4_from16(const float* m) { return m4_from(m); }
inline P2 map_sphere(const M4& w2t, V3 p) {  // :177-187
    V3 vec = normalize(m4_point(w2t, p) - V3(0, 0, 0));
    return P2(spherical_theta(vec) * INV_PI, spherical_phi(vec) * INV2_PI);
}
inline P2 map_cylinder(const M4& w2t, V3 p) {  // :222-229
    V3 vec = normalize(m4_point(w2t, p) - V3(0, 0, 0));
    return P2(PI + std::atan2(vec.y, vec.x) * INV2_PI, vec.z);
}
inline void fix_seam(P2& d) {  // :198-204
    if (d.y > 0.5f) d.y = 1.0f - d.y;
    else if (d.y < -0.5f) d.y = -(d.y + 1.0f);
}
inline P2 map2d(const pbrt_b200_texnode& n, const SurfaceInteraction& si, P2* dstdx, P2* dstdy) {
    switch (n.mapping) {
        case PBRT_B200_MAP_UV: {
            Float su = n.m[0], sv = n.m[1], du = n.m[2], dv = n.m[3];
            *dstdx = P2(su * si.dudx, sv * si.dvdx);
            *dstdy = P2(su * si.dudy, sv * si.dvdy);
            return P2(su * si.uv.x + du, sv * si.uv.y + dv);
        }
        case PBRT_B200_MAP_SPHERICAL:
        case PBRT_B200_MAP_CYLINDRICAL: {
            M4 w2t = m4_from16(n.m);
            bool sph = n.mapping == PBRT_B200_MAP_SPHERICAL;
            auto f = [&](V3 p) { return sph ? map_sphere(w2t, p) : map_cylinder(w2t, p); };
            P2 st = f(si.p);
            const Float delta = 0.1f, inv = 1.0f / delta;  // Vector2 / Float multiplies by the reciprocal (vector.rs:162-170)
            P2 sx = f(si.p + si.dpdx * delta), sy = f(si.p + si.dpdy * delta);
            *dstdx = P2((sx.x - st.x) * inv, (sx.y - st.y) * inv);
            *dstdy = P2((sy.x - st.x) * inv, (sy.y - st.y) * inv);
            fix_seam(*dstdx);
            fix_seam(*dstdy);
            return st;
        }
        default: {  // planar, :266-283
            V3 vs(n.m[0], n.m[1], n.m[2]), vt(n.m[3], n.m[4], n.m[5]);
            *dstdx = P2(dot(si.dpdx, vs), dot(si.dpdx, vt));
            *dstdy = P2(dot(si.dpdy, vs), dot(si.dpdy, vt));
            return P2(n.m[6] + dot(si.p, vs), n.m[7] + dot(si.p, vt));
        }
    }
}
inline V3 map3d(const pbrt_b200_texnode& n, const SurfaceInteraction& si, V3* dpdx, V3* dpdy) {  // IdentityMapping3D, :299-317
    M4 w2t = m4_from16(n.m);
    *dpdx = m4_vector(w2t, si.dpdx);
    *dpdy = m4_vector(w2t, si.dpdy);
    return m4_point(w2t, si.p);
}

inline bool checker_even(int64_t a) { return a % 2 == 0; }  // `% 2 == 0` on isize: true for even negatives too

// Texture::evaluate over a postfix program (include/pbrt_b200.h)
inline Spectrum tex_eval(const SceneView& s, pbrt_b200_texref ref, const SurfaceInteraction& si) {
    Spectrum stack[16];
    int sp = 0;
    for (uint32_t k = 0; k < ref.count; ++k) {
        const pbrt_b200_texnode& n = s.d.textures[ref.first + k];
        P2 dstdx, dstdy;
        V3 dpdx, dpdy;
        switch (n.kind) {
            case PBRT_B200_TEX_CONSTANT: stack[sp++] = Spectrum(n.v[0], n.v[1], n.v[2]); break;
            case PBRT_B200_TEX_SCALE: { Spectrum b = stack[--sp], a = stack[--sp]; stack[sp++] = a * b; break; }
            case PBRT_B200_TEX_MIX: {
                Float amt = stack[--sp].c[0];
                Spectrum t2 = stack[--sp], t1 = stack[--sp];
                stack[sp++] = t1 * (1.0f - amt) + t2 * amt;
                break;
            }
            case PBRT_B200_TEX_BILERP: {  // biler.rs:27-36
                P2 st = map2d(n, si, &dstdx, &dstdy);
                Spectrum v00 = spec3(n.v), v01 = spec3(n.v + 3), v10 = spec3(n.v + 6), v11 = spec3(n.v + 9);
                stack[sp++] = v00 * (1.0f - st.y) * (1.0f - st.x) + v01 * (1.0f - st.x) * st.y + v10 * (1.0f - st.y) * st.x + v11 * st.y * st.x;
                break;
            }
            case PBRT_B200_TEX_IMAGEMAP: {
                P2 st = map2d(n, si, &dstdx, &dstdy);
                stack[sp++] = MipView(s.d.mipmaps[n.image]).lookup2(st, dstdx, dstdy);
                break;
            }
            case PBRT_B200_TEX_UV: {
                P2 st = map2d(n, si, &dstdx, &dstdy);
                stack[sp++] = Spectrum(st.x - std::floor(st.x), st.y - std::floor(st.y), 0.0f);
                break;
            }
            case PBRT_B200_TEX_CHECKERBOARD2D: {  // checkerboard.rs:28-73
                Spectrum t2 = stack[--sp], t1 = stack[--sp];
                P2 st = map2d(n, si, &dstdx, &dstdy);
                bool first = checker_even((int64_t)std::floor(st.x) + (int64_t)std::floor(st.y));
                if (!(n.flags & PBRT_B200_TEX_AA_CLOSEDFORM)) { stack[sp++] = first ? t1 : t2; break; }
                Float ds = std::fmax(std::fabs(dstdx.x), std::fabs(dstdy.x)), dt = std::fmax(std::fabs(dstdx.y), std::fabs(dstdy.y));
                Float s0 = st.x - ds, s1 = st.x + ds, t0 = st.y - dt, t1_ = st.y + dt;
                if (std::floor(s0) == std::floor(s1) && std::floor(t0) == std::floor(t1_)) { stack[sp++] = first ? t1 : t2; break; }
                auto bumpint = [](Float x) { return std::floor(x / 2.0f) + 2.0f * std::fmax(x / 2.0f - std::floor(x / 2.0f) - 0.5f, 0.0f); };
                Float sint = (bumpint(s1) - bumpint(s0)) / (2.0f * ds), tint = (bumpint(t1_) - bumpint(t0)) / (2.0f * dt);
                Float area2 = sint * tint - 2.0f * sint * tint;  // the reference's expression (pbrt-v3 has sint + tint - 2 sint tint)
                if (ds > 1.0f || dt > 1.0f) area2 = 0.5f;
                stack[sp++] = t1 * (1.0f - area2) + t2 * area2;
                break;
            }
            case PBRT_B200_TEX_CHECKERBOARD3D: {
                Spectrum t2 = stack[--sp], t1 = stack[--sp];
                V3 p = map3d(n, si, &dpdx, &dpdy);
                stack[sp++] = checker_even((int64_t)std::floor(p.x) + (int64_t)std::floor(p.y) + (int64_t)std::floor(p.z)) ? t1 : t2;
                break;
            }
            case PBRT_B200_TEX_DOTS: {  // dots.rs:28-57 (`as usize` saturates negative cells to 0)
                Spectrum inside = stack[--sp], outside = stack[--sp];
                P2 st = map2d(n, si, &dstdx, &dstdy);
                Float scell = (Float)f2u_sat(std::floor(st.x + 0.5f)), tcell = (Float)f2u_sat(std::floor(st.y + 0.5f));
                Spectrum r = outside;
                if (noise(scell + 0.5f, tcell + 0.5f, 0.5f) > 0.0f) {
                    const Float radius = 0.35f, max_shift = 0.5f - radius;
                    Float scenter = scell + max_shift * noise(scell + 1.5f, tcell + 2.8f, 0.5f);
                    Float tcenter = tcell + max_shift * noise(scell + 4.5f, tcell + 9.8f, 0.5f);
                    Float dx = st.x - scenter, dy = st.y - tcenter;
                    if (dx * dx + dy * dy < radius * radius) r = inside;
                }
                stack[sp++] = r;
                break;
            }
            case PBRT_B200_TEX_FBM: {
                V3 p = map3d(n, si, &dpdx, &dpdy);
                stack[sp++] = Spectrum(fbm(p, dpdx, dpdy, n.v[0], (int)n.v[1]));
                break;
            }
            case PBRT_B200_TEX_WRINKLED: {
                V3 p = map3d(n, si, &dpdx, &dpdy);
                stack[sp++] = Spectrum(turbulence(p, dpdx, dpdy, n.v[0], (int)n.v[1]));
                break;
            }
            case PBRT_B200_TEX_MARBLE: {  // marble.rs:40-76
                static const Float C[9][3] = {{0.58f, 0.58f, 0.6f}, {0.58f, 0.58f, 0.6f}, {0.58f, 0.58f, 0.6f}, {0.5f, 0.5f, 0.5f}, {0.6f, 0.59f, 0.58f},
                                              {0.58f, 0.58f, 0.6f}, {0.58f, 0.58f, 0.6f}, {0.2f, 0.2f, 0.33f}, {0.58f, 0.58f, 0.6f}};
                V3 p = map3d(n, si, &dpdx, &dpdy);
                Float scale = n.v[2], variation = n.v[3];
                p = p * scale;
                Float marble = p.y + variation * fbm(p, dpdx * scale, dpdy * scale, n.v[0], (int)n.v[1]);
                Float t = 0.5f + 0.5f * std::sin(marble);
                int first = (int)std::min<uint64_t>(5, f2u_sat(std::floor(t * 6.0f)));
                Spectrum c0 = spec3(C[first]), c1 = spec3(C[first + 1]), c2 = spec3(C[first + 2]), c3 = spec3(C[first + 3]);
                Spectrum s0 = c0 * (1.0f - t) + c1 * t, s1 = c1 * (1.0f - t) + c2 * t, s2 = c2 * (1.0f - t) + c3 * t;
                s0 = s0 * (1.0f - t) + s1 * t;
                s1 = s1 * (1.0f - t) + s2 * t;
                stack[sp++] = (s0 * (1.0f - t) + s1 * t) * 1.5f;
                break;
            }
            case PBRT_B200_TEX_WINDY: {  // windy.rs:18-27
                V3 p = map3d(n, si, &dpdx, &dpdy);
                Float wstrength = fbm(p * 0.1f, dpdx * 0.1f, dpdy * 0.1f, 0.5f, 3);
                Float wheight = fbm(p, dpdx, dpdy, 0.5f, 6);
                stack[sp++] = Spectrum(std::fabs(wstrength) * wheight);
                break;
            }
            default: stack[sp++] = Spectrum(0.0f); break;
        }
    }
    return sp > 0 ? stack[sp - 1] : Spectrum(0.0f);
}

// ---- SurfaceInteraction::compute_differentials, interaction.rs:269-342 ----------------------------------------------------------
inline bool solve_linear_2x2(const Float A[2][2], const Float B[2], Float* x0, Float* x1) {  // pbrt.rs solve_linearsystem_2x2
    Float det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    if (std::fabs(det) < 1.0e-10f) return false;
    *x0 = (A[1][1] * B[0] - A[0][1] * B[1]) / det;
    *x1 = (A[0][0] * B[1] - A[1][0] * B[0]) / det;
    if (std::isnan(*x0) || std::isnan(*x1)) return false;
    return true;
}
inline void compute_differentials(SurfaceInteraction& si, const Ray& r) {
    si.dudx = si.dvdx = si.dudy = si.dvdy = 0.0f;
    si.dpdx = si.dpdy = V3();
    if (!r.has_diff) return;
    Float d = dot(si.n, si.p);
    Float tx = -(dot(si.n, r.rxo) - d) / dot(si.n, r.rxd);
    if (std::isinf(tx) || std::isnan(tx)) return;
    V3 px = r.rxo + r.rxd * tx;
    Float ty = -(dot(si.n, r.ryo) - d) / dot(si.n, r.ryd);
    if (std::isinf(ty) || std::isnan(ty)) return;
    V3 py = r.ryo + r.ryd * ty;
    si.dpdx = px - si.p;
    si.dpdy = py - si.p;
    int dim[2];
    if (std::fabs(si.n.x) > std::fabs(si.n.y) && std::fabs(si.n.x) > std::fabs(si.n.z)) { dim[0] = 1; dim[1] = 2; }
    else if (std::fabs(si.n.y) > std::fabs(si.n.z)) { dim[0] = 0; dim[1] = 2; }
    else { dim[0] = 0; dim[1] = 1; }
    Float A[2][2] = {{si.dpdu[dim[0]], si.dpdv[dim[0]]}, {si.dpdu[dim[1]], si.dpdv[dim[1]]}};
    Float Bx[2] = {px[dim[0]] - si.p[dim[0]], px[dim[1]] - si.p[dim[1]]};
    Float By[2] = {py[dim[0]] - si.p[dim[0]], py[dim[1]] - si.p[dim[1]]};
    if (!solve_linear_2x2(A, Bx, &si.dudx, &si.dvdx)) si.dudx = si.dvdx = 0.0f;
    if (!solve_linear_2x2(A, By, &si.dudy, &si.dvdy)) si.dudy = si.dvdy = 0.0f;
}

// ---- Material::bump, material.rs:46-87 ---------------------------------------------------------------------------------------
inline void bump(const SceneView& s, pbrt_b200_texref d, SurfaceInteraction& si) {
    SurfaceInteraction ev = si;
    Float du = 0.5f * (std::fabs(si.dudx) + std::fabs(si.dudy));
    if (du == 0.0f) du = 0.0005f;
    ev.p = si.p + si.sh_dpdu * du;
    ev.uv = P2(si.uv.x + du, si.uv.y + 0.0f);
    ev.n = normalize(cross(si.sh_dpdu, si.sh_dpdv)) + si.dndu * du;
    Float udisplace = tex_eval(s, d, ev).c[0];
    Float dv = 0.5f * (std::fabs(si.dvdx) + std::fabs(si.dvdy));
    if (dv == 0.0f) dv = 0.0005f;
    ev.p = si.p + si.sh_dpdv * dv;
    ev.uv = P2(si.uv.x + 0.0f, si.uv.y + dv);
    ev.n = normalize(cross(si.sh_dpdu, si.sh_dpdv)) + si.dndv * dv;
    Float vdisplace = tex_eval(s, d, ev).c[0];
    Float displace = tex_eval(s, d, si).c[0];
    V3 dpdu = si.sh_dpdu + si.sh_n * ((udisplace - displace) / du) + si.sh_dndu * displace;
    V3 dpdv = si.sh_dpdv + si.sh_n * ((vdisplace - displace) / dv) + si.sh_dndv * displace;
    set_shading_geometry(si, dpdu, dpdv, si.sh_dndu, si.sh_dndv, false);
}

// ---- materials with texture-valued parameters (pbrt_b200_material_ext) ---------------------------------------------------------
struct MatParams {
    Spectrum s[5];
    Float f[3];
};
inline void compute_scattering_functions_ext(const SceneView& sv, int mat, SurfaceInteraction& si, BSDF* bsdf, bool allow_multiple_lobes) {
    const pbrt_b200_material& m = sv.d.materials[mat];
    const pbrt_b200_material_ext& x = sv.d.material_ext[mat];
    if (x.bump.count) bump(sv, x.bump, si);
    MatParams P;
    for (int k = 0; k < 5; ++k) P.s[k] = x.s_tex[k].count ? tex_eval(sv, x.s_tex[k], si) : spec3(x.s_const[k]);
    for (int k = 0; k < 3; ++k) P.f[k] = x.f_tex[k].count ? tex_eval(sv, x.f_tex[k], si).c[0] : x.f_const[k];
    if (m.type <= PBRT_B200_MAT_METAL) {  // the five hot materials: the constant-parameter code with the evaluated values
        pbrt_b200_material c = m;
        for (int k = 0; k < 3; ++k) { c.a[k] = P.s[0].c[k]; c.b[k] = P.s[1].c[k]; }
        c.f0 = P.f[0]; c.f1 = P.f[1]; c.f2 = P.f[2];
        compute_scattering_functions(c, si, bsdf, allow_multiple_lobes);
        return;
    }
    bsdf->valid = false;
    if (m.type == PBRT_B200_MAT_UBER) {  // uber.rs:41-112
        Float e = P.f[2];
        Spectrum op = spec_clamp(P.s[4], 0.0f, INFINITY_F);
        Spectrum t = spec_clamp(op * -1.0f + Spectrum(1.0f), 0.0f, INFINITY_F);
        if (!t.is_black()) {
            bsdf->init(si, 1.0f);
            BxDF b; b.kind = BX_SPEC_TRANS; b.type = BSDF_TRANSMISSION | BSDF_SPECULAR; b.t = t; b.etaa = 1.0f; b.etab = 1.0f;
            bsdf->add(b);
        } else bsdf->init(si, e);
        Spectrum kd = op * spec_clamp(P.s[0], 0.0f, INFINITY_F);
        if (!kd.is_black()) { BxDF b; b.kind = BX_LAMBERT; b.type = BSDF_REFLECTION | BSDF_DIFFUSE; b.r = kd; bsdf->add(b); }
        Spectrum ks = op * spec_clamp(P.s[1], 0.0f, INFINITY_F);
        if (!ks.is_black()) {
            Float ru = P.f[0], rv = P.f[1];
            if (m.remap_roughness) { ru = roughness_to_alpha(ru); rv = roughness_to_alpha(rv); }
            BxDF b; b.kind = BX_MICRO_REFL; b.type = BSDF_REFLECTION | BSDF_GLOSSY; b.r = ks;
            b.fresnel = FR_DIELECTRIC; b.fr_etai = 1.0f; b.fr_etat = e; b.distrib = TrowbridgeReitz(ru, rv);
            bsdf->add(b);
        }
        Spectrum kr = op * spec_clamp(P.s[2], 0.0f, INFINITY_F);
        if (!kr.is_black()) {
            BxDF b; b.kind = BX_SPEC_REFL; b.type = BSDF_REFLECTION | BSDF_SPECULAR; b.r = kr; b.fresnel = FR_DIELECTRIC; b.fr_etai = 1.0f; b.fr_etat = e;
            bsdf->add(b);
        }
        Spectrum kt = op * spec_clamp(P.s[3], 0.0f, INFINITY_F);
        if (!kt.is_black()) {
            BxDF b; b.kind = BX_SPEC_TRANS; b.type = BSDF_TRANSMISSION | BSDF_SPECULAR; b.t = kt; b.etaa = 1.0f; b.etab = e;
            bsdf->add(b);
        }
    } else if (m.type == PBRT_B200_MAT_SUBSTRATE) {  // substrate.rs:34-62
        bsdf->init(si, 1.0f);
        bsdf->valid = false;  // si.bsdf is only set inside the branch below
        Spectrum d = spec_clamp(P.s[0], 0.0f, INFINITY_F), sp = spec_clamp(P.s[1], 0.0f, INFINITY_F);
        Float ru = P.f[0], rv = P.f[1];
        if (!d.is_black() || !sp.is_black()) {
            if (m.remap_roughness) { ru = roughness_to_alpha(ru); rv = roughness_to_alpha(rv); }
            BxDF b; b.kind = BX_FRESNEL_BLEND; b.type = BSDF_REFLECTION | BSDF_GLOSSY; b.r = d; b.t = sp; b.distrib = TrowbridgeReitz(ru, rv);
            bsdf->add(b);
            bsdf->valid = true;
        }
    }
}

// Primitive::compute_scattering_functions as the integrators call it (interaction.rs:258-267): differentials first, then the material.
inline void scattering_functions(const SceneView& sv, const Ray& ray, int slot, SurfaceInteraction& si, BSDF* bsdf, bool allow_multiple_lobes) {
    bsdf->valid = false;
    int mat = sv.d.prims[slot].material;
    if (mat < 0) return;
    const pbrt_b200_material& m = sv.d.materials[mat];
    // (a scene without textures never reads the differentials: skipped there, which keeps the CPU baseline's timing what it was)
    if (sv.d.material_ext) compute_differentials(si, ray);
    if (m.textured && sv.d.material_ext) compute_scattering_functions_ext(sv, mat, si, bsdf, allow_multiple_lobes);
    else compute_scattering_functions(m, si, bsdf, allow_multiple_lobes);
}

}  // namespace orc
