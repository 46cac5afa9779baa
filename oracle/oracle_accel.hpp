// TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see oracle_math.hpp header).
// oracle_accel.hpp: BVHAccel build + traversal, Triangle, Sphere, SurfaceInteraction.
#pragma once
#include <vector>
#include <cassert>
#include "oracle_math.hpp"
#include "../include/pbrt_b200.h"  // flat scene tables (layout only; no product code)

namespace orc {

// ------------------------------------------------------------------------
// BVH build: src/accelerators/bvh.rs:145-375, 662-693 (SAH; Middle; EqualCounts)
// ------------------------------------------------------------------------
struct BVHPrimitiveInfo {  // bvh.rs:53-68
    size_t primitive_number;
    Bounds3 bounds;
    V3 centroid;
};
struct BVHBuildNode {  // bvh.rs:97-105
    Bounds3 bounds;
    int left, right;  // indices into the arena, -1 = none
    int splitaxis;
    size_t first_prim_offset, n_primitives;
};
static const int NBUCKETS = 12;  // bvh.rs:35

// core::iter::Iterator::partition_in_place (used at bvh.rs:288,361): repeatedly find
// the first element failing the predicate and swap it with the last element passing it.
template <typename P>
inline size_t partition_in_place(BVHPrimitiveInfo* a, size_t n, P pred) {
    size_t true_count = 0;
    size_t head = 0, tail = n;
    for (;;) {
        // find(is_false): advance head, counting trues
        bool found_head = false;
        while (head < tail) {
            bool p = pred(a[head]);
            true_count += p ? 1 : 0;
            if (!p) { found_head = true; break; }
            ++head;
        }
        if (!found_head) break;
        // rfind(is_true)
        bool found_tail = false;
        while (tail > head + 1) {
            --tail;
            if (pred(a[tail])) { found_tail = true; break; }
        }
        if (!found_tail) break;
        std::swap(a[head], a[tail]);
        true_count += 1;
        ++head;
    }
    return true_count;
}

struct BVHBuilder {
    size_t max_prims;
    int split_method;
    std::vector<BVHBuildNode> arena;
    std::vector<uint32_t> ordered;
    size_t total_nodes = 0;

    void init_leaf(BVHBuildNode& node, std::vector<BVHPrimitiveInfo>& info, size_t start, size_t end, const Bounds3& bounds) {
        size_t offset = ordered.size();
        for (size_t i = start; i < end; ++i) ordered.push_back((uint32_t)info[i].primitive_number);
        node.first_prim_offset = offset; node.n_primitives = end - start; node.bounds = bounds;
        node.left = node.right = -1;
    }

    // bvh.rs:301-375; returns (mid, create_leaf)
    std::pair<size_t, bool> split_sah(const Bounds3& bounds, const Bounds3& cb, int dim, size_t nprims, size_t start, size_t end,
                                      std::vector<BVHPrimitiveInfo>& info) {
        if (nprims <= 2) {
            size_t mid = (start + end) / 2;
            if (start != end - 1 && info[end - 1].centroid[dim] < info[start].centroid[dim]) std::swap(info[start], info[end - 1]);
            return {mid, false};
        }
        struct Bucket { size_t count = 0; Bounds3 bounds; } buckets[NBUCKETS];
        auto bucket_of = [&](const BVHPrimitiveInfo& p) {
            size_t b = (size_t)f2u_sat((Float)NBUCKETS * bounds_offset(cb, p.centroid)[dim]);
            if (b == (size_t)NBUCKETS) b = NBUCKETS - 1;
            assert(b < (size_t)NBUCKETS);
            return b;
        };
        for (size_t i = start; i < end; ++i) {
            size_t b = bucket_of(info[i]);
            buckets[b].count += 1;
            buckets[b].bounds = union_bounds(buckets[b].bounds, info[i].bounds);
        }
        Float cost[NBUCKETS - 1];
        for (int i = 0; i < NBUCKETS - 1; ++i) {
            Bounds3 b0, b1;
            size_t c0 = 0, c1 = 0;
            for (int j = 0; j <= i; ++j) { b0 = union_bounds(b0, buckets[j].bounds); c0 += buckets[j].count; }
            for (int j = i + 1; j < NBUCKETS; ++j) { b1 = union_bounds(b1, buckets[j].bounds); c1 += buckets[j].count; }
            cost[i] = 1.0f + ((Float)c0 * surface_area(b0) + (Float)c1 * surface_area(b1)) / surface_area(bounds);
        }
        Float min_cost = cost[0];
        size_t min_bucket = 0;
        for (int i = 1; i < NBUCKETS - 1; ++i)
            if (cost[i] < min_cost) { min_cost = cost[i]; min_bucket = i; }
        Float leaf_cost = (Float)nprims;
        if (nprims > max_prims || min_cost < leaf_cost) {
            size_t pmid = partition_in_place(&info[start], end - start, [&](const BVHPrimitiveInfo& pi) { return bucket_of(pi) <= min_bucket; }) + start;
            return {pmid, false};
        }
        return {0, true};
    }

    // bvh.rs:200-283
    int recursive_build(std::vector<BVHPrimitiveInfo>& info, size_t start, size_t end) {
        int ni = (int)arena.size();
        arena.push_back(BVHBuildNode());
        total_nodes += 1;
        Bounds3 bounds;
        for (size_t i = start; i < end; ++i) bounds = union_bounds(bounds, info[i].bounds);
        size_t nprims = end - start;
        if (nprims == 1) { init_leaf(arena[ni], info, start, end, bounds); return ni; }
        Bounds3 cb;
        for (size_t i = start; i < end; ++i) cb = union_point(cb, info[i].centroid);
        int dim = maximum_extent(cb);
        size_t mid = (start + end) / 2;
        if (cb.p_max[dim] == cb.p_min[dim]) { init_leaf(arena[ni], info, start, end, bounds); return ni; }
        if (split_method == PBRT_B200_SPLIT_MIDDLE) {
            Float pmid = (cb.p_min[dim] + cb.p_max[dim]) / 2.0f;  // bvh.rs:285-289
            mid = start + partition_in_place(&info[start], end - start, [&](const BVHPrimitiveInfo& p) { return p.centroid[dim] < pmid; });
        }
        if ((split_method == PBRT_B200_SPLIT_MIDDLE && (mid == start || mid == end)) || split_method == PBRT_B200_SPLIT_EQUAL) {
            // bvh.rs:291-299 select_nth_unstable_by: element order inside the halves is
            // implementation-defined in Rust; only SAH is pinned.
            mid = (start + end) / 2;
            std::nth_element(info.begin() + start, info.begin() + mid, info.begin() + end,
                             [dim](const BVHPrimitiveInfo& a, const BVHPrimitiveInfo& b) { return a.centroid[dim] < b.centroid[dim]; });
        } else {
            // bvh.rs:253-270: the `_` arm also catches Middle whose partition succeeded, so
            // "middle" is followed by a SAH split of the same range (reference quirk).
            auto r = split_sah(bounds, cb, dim, nprims, start, end, info);
            if (r.second) { init_leaf(arena[ni], info, start, end, bounds); return ni; }
            mid = r.first;
        }
        // bvh.rs:275-276: right subtree first
        int right = recursive_build(info, mid, end);
        int left = recursive_build(info, start, mid);
        BVHBuildNode& node = arena[ni];
        node.left = left; node.right = right;
        node.bounds = union_bounds(arena[left].bounds, arena[right].bounds);
        node.splitaxis = dim; node.n_primitives = 0;
        return ni;
    }

    // bvh.rs:662-693
    size_t flatten(std::vector<pbrt_b200_bvh_node>& nodes, int n, size_t* offset) {
        size_t my = (*offset)++;
        const BVHBuildNode& node = arena[n];
        pbrt_b200_bvh_node ln;
        std::memset(&ln, 0, sizeof ln);
        ln.bounds[0] = node.bounds.p_min.x; ln.bounds[1] = node.bounds.p_min.y; ln.bounds[2] = node.bounds.p_min.z;
        ln.bounds[3] = node.bounds.p_max.x; ln.bounds[4] = node.bounds.p_max.y; ln.bounds[5] = node.bounds.p_max.z;
        if (node.n_primitives > 0) {
            ln.n_prims = (uint16_t)node.n_primitives; ln.offset = (uint32_t)node.first_prim_offset;
        } else {
            flatten(nodes, node.left, offset);
            ln.offset = (uint32_t)flatten(nodes, node.right, offset);
            ln.axis = (uint8_t)node.splitaxis;
        }
        nodes[my] = ln;
        return my;
    }
};

// BVHAccel::new, bvh.rs:145-198
inline void bvh_build(const float* prim_bounds, size_t n, size_t max_prims, int split_method, std::vector<pbrt_b200_bvh_node>& nodes,
                      std::vector<uint32_t>& ordered) {
    nodes.clear(); ordered.clear();
    if (n == 0) return;
    BVHBuilder b;
    b.max_prims = std::min<size_t>(255, max_prims);
    b.split_method = split_method;
    std::vector<BVHPrimitiveInfo> info(n);
    for (size_t i = 0; i < n; ++i) {
        const float* pb = prim_bounds + 6 * i;
        info[i].primitive_number = i;
        info[i].bounds = Bounds3(V3(pb[0], pb[1], pb[2]), V3(pb[3], pb[4], pb[5]));
        info[i].centroid = info[i].bounds.p_min * 0.5f + info[i].bounds.p_max * 0.5f;  // bvh.rs:66
    }
    b.arena.reserve(2 * n);
    b.ordered.reserve(n);
    int root = b.recursive_build(info, 0, n);
    nodes.resize(b.total_nodes);
    size_t offset = 0;
    b.flatten(nodes, root, &offset);
    assert(offset == b.total_nodes);
    ordered.swap(b.ordered);
}

// ------------------------------------------------------------------------
// Scene view over the flat tables
// ------------------------------------------------------------------------
struct Counters {
    uint64_t nodes_tested = 0;   // Bounds3f::intersect_p2 calls
    uint64_t tris_tested = 0;    // Shape::intersect / intersect_p calls
    uint64_t rays = 0;
};

struct Hit {
    uint32_t slot = PBRT_B200_NO_HIT;  // index into prims[] (BVH order); for an instanced hit the row of the object's primitive
    Float t = 0, b0 = 0, b1 = 0, b2 = 0;
    uint32_t inst = PBRT_B200_NO_HIT;  // instances[] index when the hit went through a TransformedPrimitive
};

struct SceneView {
    pbrt_b200_scene_desc d;
    Bounds3 wb;  // Scene.wb, scene.rs:33
    V3 P(uint32_t vi) const { return V3(d.vertex_p[3 * vi], d.vertex_p[3 * vi + 1], d.vertex_p[3 * vi + 2]); }
    V3 N(uint32_t vi) const { return V3(d.vertex_n[3 * vi], d.vertex_n[3 * vi + 1], d.vertex_n[3 * vi + 2]); }
    V3 S(uint32_t vi) const { return V3(d.vertex_s[3 * vi], d.vertex_s[3 * vi + 1], d.vertex_s[3 * vi + 2]); }
    P2 UV(uint32_t vi) const { return P2(d.vertex_uv[2 * vi], d.vertex_uv[2 * vi + 1]); }
    uint64_t top_nodes() const { return d.n_objects ? d.n_top_nodes : d.n_nodes; }
    void init(const pbrt_b200_scene_desc& desc) {
        d = desc;
        if (d.n_nodes) {
            const float* b = d.nodes[0].bounds;
            wb = Bounds3(V3(b[0], b[1], b[2]), V3(b[3], b[4], b[5]));
        }
    }
};

// ------------------------------------------------------------------------
// Triangle: src/shapes/triangle.rs
// ------------------------------------------------------------------------
// triangle.rs:136-233 plus the degenerate-parameterisation rejection :236-263
// (shared with intersect_p :400-495 when `closest` is false, which skips :236-263).
inline bool triangle_test(const Ray& r, V3 p0, V3 p1, V3 p2, const P2 uv[3], bool closest, Float* t_out, Float* b0_out, Float* b1_out, Float* b2_out) {
    V3 p0t = p0 - r.o, p1t = p1 - r.o, p2t = p2 - r.o;
    int kz = max_dimension(vabs(r.d));
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    V3 d = permute(r.d, kx, ky, kz);
    p0t = permute(p0t, kx, ky, kz); p1t = permute(p1t, kx, ky, kz); p2t = permute(p2t, kx, ky, kz);
    Float Sx = -d.x / d.z, Sy = -d.y / d.z, Sz = 1.0f / d.z;
    p0t.x += Sx * p0t.z; p0t.y += Sy * p0t.z;
    p1t.x += Sx * p1t.z; p1t.y += Sy * p1t.z;
    p2t.x += Sx * p2t.z; p2t.y += Sy * p2t.z;
    Float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
    Float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
    Float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
        double p2txp1ty = (double)p2t.x * (double)p1t.y, p2typ1tx = (double)p2t.y * (double)p1t.x;
        e0 = (float)(p2typ1tx - p2txp1ty);
        double p0txp2ty = (double)p0t.x * (double)p2t.y, p0typ2tx = (double)p0t.y * (double)p2t.x;
        e1 = (float)(p0typ2tx - p0txp2ty);
        double p1txp0ty = (double)p1t.x * (double)p0t.y, p1typ0tx = (double)p1t.y * (double)p0t.x;
        e2 = (float)(p1typ0tx - p1txp0ty);
    }
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
    Float det = e0 + e1 + e2;
    if (det == 0.0f) return false;
    p0t.z *= Sz; p1t.z *= Sz; p2t.z *= Sz;
    Float tscaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if (det < 0.0f && (tscaled >= 0.0f || tscaled < r.t_max * det)) return false;
    else if (det > 0.0f && (tscaled <= 0.0f || tscaled >= r.t_max * det)) return false;
    Float invdet = 1.0f / det;
    Float b0 = e0 * invdet, b1 = e1 * invdet, b2 = e2 * invdet;
    Float t = tscaled * invdet;
    Float maxzt = max_component(vabs(V3(p0t.z, p1t.z, p2t.z)));
    Float deltaz = gamma(3) * maxzt;
    Float maxxt = max_component(vabs(V3(p0t.x, p1t.x, p2t.x)));
    Float maxyt = max_component(vabs(V3(p0t.y, p1t.y, p2t.y)));
    Float deltax = gamma(5) * (maxxt + maxzt);
    Float deltay = gamma(5) * (maxyt + maxzt);
    Float deltae = 2.0f * (gamma(2) * maxxt * maxyt + deltay * maxxt + deltax * maxyt);
    Float maxe = max_component(vabs(V3(e0, e1, e2)));
    Float deltat = 3.0f * (gamma(3) * maxe * maxzt + deltae * maxzt + deltaz * maxe) * std::fabs(invdet);
    if (t <= deltat) return false;
    if (closest) {
        // triangle.rs:236-263: a hit on a triangle whose dpdu x dpdv AND geometric normal both
        // vanish is rejected ("the intersection is bogus").
        P2 duv02(uv[0].x - uv[2].x, uv[0].y - uv[2].y), duv12(uv[1].x - uv[2].x, uv[1].y - uv[2].y);
        V3 dp02 = p0 - p2, dp12 = p1 - p2;
        Float determinant = duv02.x * duv12.y - duv02.y * duv12.x;
        bool degenerateuv = std::fabs(determinant) < 1.0e-8f;
        V3 dpdu, dpdv;
        if (!degenerateuv) {
            Float inv = 1.0f / determinant;
            dpdu = (dp02 * duv12.y - dp12 * duv02.y) * inv;
            dpdv = (dp02 * -duv12.x + dp12 * duv02.x) * inv;
        }
        if (degenerateuv || length_squared(cross(dpdu, dpdv)) == 0.0f) {
            V3 ng = cross(p2 - p0, p1 - p0);
            if (length_squared(ng) == 0.0f) return false;
        }
    }
    *t_out = t; *b0_out = b0; *b1_out = b1; *b2_out = b2;
    return true;
}

inline void triangle_fetch(const SceneView& s, const pbrt_b200_prim& pr, uint32_t vi[3], V3 p[3], P2 uv[3]) {
    const uint32_t* idx = s.d.tri_indices + 3 * (size_t)pr.shape_index;
    for (int k = 0; k < 3; ++k) { vi[k] = idx[k]; p[k] = s.P(idx[k]); }
    if ((pr.flags & PBRT_B200_PRIM_HAS_UV) && s.d.vertex_uv) {
        for (int k = 0; k < 3; ++k) uv[k] = s.UV(idx[k]);
    } else {  // triangle.rs:109-115
        uv[0] = P2(0, 0); uv[1] = P2(1, 0); uv[2] = P2(1, 1);
    }
}

// ------------------------------------------------------------------------
// EFloat: src/core/efloat.rs
// ------------------------------------------------------------------------
struct EFloat {
    Float v, low, high;
    EFloat() : v(0), low(0), high(0) {}
    EFloat(Float v_, Float err = 0.0f) : v(v_) {  // efloat.rs:18-32
        if (err == 0.0f) { low = v; high = v; } else { low = next_float_down(v - err); high = next_float_up(v + err); }
    }
};
inline EFloat operator+(EFloat a, EFloat b) { EFloat r; r.v = a.v + b.v; r.low = next_float_down(a.low + b.low); r.high = next_float_up(a.high + b.high); return r; }
inline EFloat operator-(EFloat a, EFloat b) { EFloat r; r.v = a.v - b.v; r.low = next_float_down(a.low - b.high); r.high = next_float_up(a.high - b.low); return r; }
inline EFloat operator*(EFloat a, EFloat b) {
    EFloat r; r.v = a.v * b.v;
    Float p[4] = {a.low * b.low, a.high * b.low, a.low * b.high, a.high * b.high};
    r.low = next_float_down(std::fmin(std::fmin(p[0], p[1]), std::fmin(p[2], p[3])));
    r.high = next_float_up(std::fmax(std::fmax(p[0], p[1]), std::fmax(p[2], p[3])));
    return r;
}
inline EFloat operator/(EFloat a, EFloat b) {  // efloat.rs:134-156 (tests the NUMERATOR for straddling zero)
    EFloat r; r.v = a.v / b.v;
    if (a.low < 0.0f && a.high > 0.0f) { r.low = -INFINITY_F; r.high = INFINITY_F; }
    else {
        Float q[4] = {a.low / b.low, a.high / b.low, a.low / b.high, a.high / b.high};
        r.low = next_float_down(std::fmin(std::fmin(q[0], q[1]), std::fmin(q[2], q[3])));
        r.high = next_float_up(std::fmax(std::fmax(q[0], q[1]), std::fmax(q[2], q[3])));
    }
    return r;
}
// efloat.rs:211-231
inline bool ef_quadratic(EFloat a, EFloat b, EFloat c, EFloat* t0, EFloat* t1) {
    double discrim = (double)b.v * (double)b.v - 4.0 * (double)a.v * (double)c.v;
    if (discrim < 0.0) return false;
    double root = std::sqrt(discrim);
    EFloat frd((Float)root, (Float)((double)MACHINE_EPSILON * root));
    EFloat q = (b.v < 0.0f) ? EFloat(-0.5f) * (b - frd) : EFloat(-0.5f) * (b + frd);  // Mul<Float>: EFloat::from(f) * self
    *t0 = q / a;
    *t1 = c / q;
    if (t0->v > t1->v) std::swap(*t0, *t1);
    return true;
}

inline M4 m4_from(const float* m) { M4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = m[4 * i + j]; return r; }

// Sphere::intersect root selection, src/shapes/sphere.rs:59-108 (full spheres: the
// clipping branch :111-141 is unreachable when zmin=-r, zmax=r, phimax=360).
// Returns t_shape_hit.v and the object-space ray.
inline bool sphere_test(const pbrt_b200_sphere& sp, const Ray& r, Float* t_out, Ray* obj_ray) {
    V3 o_err, d_err;
    Ray ray = m4_ray_error(m4_from(sp.world_to_object), r, &o_err, &d_err);
    EFloat ox(ray.o.x, o_err.x), oy(ray.o.y, o_err.y), oz(ray.o.z, o_err.z);
    EFloat dx(ray.d.x, d_err.x), dy(ray.d.y, d_err.y), dz(ray.d.z, d_err.z);
    EFloat a = dx * dx + dy * dy + dz * dz;
    EFloat b = EFloat(2.0f) * (dx * ox + dy * oy + dz * oz);
    EFloat c = ox * ox + oy * oy + oz * oz - EFloat(sp.radius) * EFloat(sp.radius);
    EFloat t0, t1;
    if (!ef_quadratic(a, b, c, &t0, &t1)) return false;
    if (t0.high > ray.t_max || t1.low <= 0.0f) return false;
    EFloat ts = t0;
    if (ts.low <= 0.0f) {
        ts = t1;
        if (ts.high > ray.t_max) return false;
    }
    *t_out = ts.v;
    if (obj_ray) *obj_ray = ray;
    return true;
}

// ------------------------------------------------------------------------
// BVHAccel::intersect / intersect_p: src/accelerators/bvh.rs:705-814
// GeometricPrimitive::intersect: src/core/primitive.rs:126-147
// ------------------------------------------------------------------------
template <bool ANY> inline bool bvh_traverse_range(const SceneView& s, uint64_t node_base, uint64_t n_nodes, uint64_t prim_base, Ray& r, Hit* hit, Counters* cnt);
inline bool prim_intersect(const SceneView& s, uint32_t slot, Ray& r, Hit* hit, Counters* cnt);
inline bool prim_intersect_p(const SceneView& s, uint32_t slot, const Ray& r, Counters* cnt);

// TransformedPrimitive::intersect / intersect_p, src/core/primitive.rs:58-89 (static transform: interpolate returns the start
// transform, transform.rs:1493-1497; Transform::inverse swaps m and m_inv, :240-242)
template <bool ANY>
inline bool instance_intersect(const SceneView& s, uint32_t inst, Ray& r, Hit* hit, Counters* cnt) {
    const pbrt_b200_instance& in = s.d.instances[inst];
    const pbrt_b200_object& ob = s.d.objects[in.object];
    Ray ray = m4_ray(m4_from(in.world_to_prim), r);
    bool found;
    if (ob.n_nodes) found = bvh_traverse_range<ANY>(s, ob.node_offset, ob.n_nodes, ob.prim_offset, ray, hit, cnt);
    else found = ANY ? prim_intersect_p(s, (uint32_t)ob.prim_offset, ray, cnt) : prim_intersect(s, (uint32_t)ob.prim_offset, ray, hit, cnt);
    if (!found) return false;
    if (!ANY) { r.t_max = ray.t_max; hit->inst = inst; }  // primitive.rs:72
    return true;
}

inline bool prim_intersect(const SceneView& s, uint32_t slot, Ray& r, Hit* hit, Counters* cnt) {
    const pbrt_b200_prim& pr = s.d.prims[slot];
    if (pr.shape_kind == PBRT_B200_SHAPE_INSTANCE) return instance_intersect<false>(s, pr.shape_index, r, hit, cnt);
    if (cnt) cnt->tris_tested++;
    if (pr.shape_kind == PBRT_B200_SHAPE_TRIANGLE) {
        uint32_t vi[3]; V3 p[3]; P2 uv[3];
        triangle_fetch(s, pr, vi, p, uv);
        Float t, b0, b1, b2;
        if (!triangle_test(r, p[0], p[1], p[2], uv, true, &t, &b0, &b1, &b2)) return false;
        r.t_max = t;  // primitive.rs:137
        hit->slot = slot; hit->t = t; hit->b0 = b0; hit->b1 = b1; hit->b2 = b2; hit->inst = PBRT_B200_NO_HIT;
        return true;
    } else {
        Float t;
        if (!sphere_test(s.d.spheres[pr.shape_index], r, &t, nullptr)) return false;
        r.t_max = t;
        hit->slot = slot; hit->t = t; hit->b0 = hit->b1 = hit->b2 = 0.0f; hit->inst = PBRT_B200_NO_HIT;
        return true;
    }
}
inline bool prim_intersect_p(const SceneView& s, uint32_t slot, const Ray& r, Counters* cnt) {
    const pbrt_b200_prim& pr = s.d.prims[slot];
    if (pr.shape_kind == PBRT_B200_SHAPE_INSTANCE) { Ray rr = r; Hit h; return instance_intersect<true>(s, pr.shape_index, rr, &h, cnt); }
    if (cnt) cnt->tris_tested++;
    if (pr.shape_kind == PBRT_B200_SHAPE_TRIANGLE) {
        uint32_t vi[3]; V3 p[3]; P2 uv[3];
        triangle_fetch(s, pr, vi, p, uv);
        Float t, b0, b1, b2;
        return triangle_test(r, p[0], p[1], p[2], uv, false, &t, &b0, &b1, &b2);
    } else {
        Float t;
        return sphere_test(s.d.spheres[pr.shape_index], r, &t, nullptr);
    }
}

inline Bounds3 node_bounds(const pbrt_b200_bvh_node& n) { return Bounds3(V3(n.bounds[0], n.bounds[1], n.bounds[2]), V3(n.bounds[3], n.bounds[4], n.bounds[5])); }

// node_base / prim_base: where this accelerator's nodes and primitive rows start in the flat arrays (0 for Scene.aggregate;
// an object's BVH stores offsets relative to its own base, include/pbrt_b200.h pbrt_b200_object)
template <bool ANY>
inline bool bvh_traverse_range(const SceneView& s, uint64_t node_base, uint64_t n_nodes, uint64_t prim_base, Ray& r, Hit* hit, Counters* cnt) {
    if (n_nodes == 0) return false;
    const pbrt_b200_bvh_node* nodes = s.d.nodes + node_base;
    bool found = false;
    V3 inv_dir(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    int dir_isneg[3] = {inv_dir.x < 0.0f, inv_dir.y < 0.0f, inv_dir.z < 0.0f};
    size_t to_visit = 0, current = 0;
    size_t stack[64];
    for (;;) {
        const pbrt_b200_bvh_node& node = nodes[current];
        if (cnt) cnt->nodes_tested++;
        if (bounds_intersect_p2(node_bounds(node), r, inv_dir, dir_isneg)) {
            if (node.n_prims > 0) {
                for (uint32_t i = 0; i < node.n_prims; ++i) {
                    const uint32_t slot = (uint32_t)(prim_base + node.offset + i);
                    if (ANY) { if (prim_intersect_p(s, slot, r, cnt)) return true; }
                    else if (prim_intersect(s, slot, r, hit, cnt)) found = true;
                }
                if (to_visit == 0) break;
                current = stack[--to_visit];
            } else {
                if (dir_isneg[node.axis]) { stack[to_visit++] = current + 1; current = node.offset; }
                else { stack[to_visit++] = node.offset; current = current + 1; }
            }
        } else {
            if (to_visit == 0) break;
            current = stack[--to_visit];
        }
    }
    return found;
}
template <bool ANY>
inline bool bvh_traverse(const SceneView& s, Ray& r, Hit* hit, Counters* cnt) {
    if (s.d.n_nodes == 0) return false;
    if (cnt) cnt->rays++;
    return bvh_traverse_range<ANY>(s, 0, s.top_nodes(), 0, r, hit, cnt);
}
// Scene::intersect / intersect_p, src/core/scene.rs:54-66
inline bool scene_intersect(const SceneView& s, Ray& r, Hit* hit, Counters* cnt) { return bvh_traverse<false>(s, r, hit, cnt); }
inline bool scene_intersect_p(const SceneView& s, Ray& r, Counters* cnt) { Hit h; return bvh_traverse<true>(s, r, &h, cnt); }

// ------------------------------------------------------------------------
// SurfaceInteraction: src/core/interaction.rs:149-249
// ------------------------------------------------------------------------
struct SurfaceInteraction {
    V3 p, p_error, wo, n;
    P2 uv;
    V3 dpdu, dpdv;
    V3 sh_n, sh_dpdu, sh_dpdv;  // Shading{n,dpdu,dpdv}
    Float time = 0;
    uint32_t slot = PBRT_B200_NO_HIT;
    // what textures, bump mapping and the specular ray differentials read
    V3 dndu, dndv, sh_dndu, sh_dndv;
    V3 dpdx, dpdy;                              // compute_differentials, interaction.rs:269-342
    Float dudx = 0, dvdx = 0, dudy = 0, dvdy = 0;
    bool shape_some = false, shape_flip = false;  // `shape` is Some / reverse_orientation ^ transform_swapshandedness
};

// SurfaceInteraction::new, interaction.rs:186-232.  `flip` = shape is Some and
// reverse_orientation ^ transform_swapshandedness.
inline SurfaceInteraction si_new(V3 p, V3 p_error, P2 uv, V3 wo, V3 dpdu, V3 dpdv, Float time, bool flip) {
    SurfaceInteraction si;
    V3 n = normalize(cross(dpdu, dpdv));
    si.sh_n = n;
    if (flip) { n = n * -1.0f; si.sh_n = si.sh_n * -1.0f; }
    si.n = n; si.time = time; si.p_error = p_error; si.wo = normalize(wo); si.p = p; si.uv = uv;
    si.dpdu = dpdu; si.dpdv = dpdv; si.sh_dpdu = dpdu; si.sh_dpdv = dpdv;
    return si;
}

// SurfaceInteraction::set_shading_geometry, interaction.rs:234-255
inline void set_shading_geometry(SurfaceInteraction& si, V3 dpdus, V3 dpdvs, V3 dndus, V3 dndvs, bool orientation_is_authoritative) {
    si.sh_n = normalize(cross(dpdus, dpdvs));
    if (si.shape_some) {
        if (si.shape_flip) si.sh_n = -si.sh_n;
        if (orientation_is_authoritative) si.n = face_forward(si.n, si.sh_n);
        else si.sh_n = face_forward(si.sh_n, si.n);
    }
    si.sh_dpdu = dpdus; si.sh_dpdv = dpdvs; si.sh_dndu = dndus; si.sh_dndv = dndvs;
}

// Triangle::intersect tail, triangle.rs:236-392, for the accepted hit.
// `shape_some` mirrors the `s: Option<Arc<Shapes>>` argument (None from Shape::pdf_wi).
inline SurfaceInteraction triangle_interaction(const SceneView& s, const pbrt_b200_prim& pr, const Ray& r, Float b0, Float b1, Float b2, bool shape_some) {
    uint32_t vi[3]; V3 p[3]; P2 uv[3];
    triangle_fetch(s, pr, vi, p, uv);
    V3 p0 = p[0], p1 = p[1], p2 = p[2];
    P2 duv02(uv[0].x - uv[2].x, uv[0].y - uv[2].y), duv12(uv[1].x - uv[2].x, uv[1].y - uv[2].y);
    V3 dp02 = p0 - p2, dp12 = p1 - p2;
    Float determinant = duv02.x * duv12.y - duv02.y * duv12.x;
    bool degenerateuv = std::fabs(determinant) < 1.0e-8f;
    V3 dpdu, dpdv;
    if (!degenerateuv) {
        Float inv = 1.0f / determinant;
        dpdu = (dp02 * duv12.y - dp12 * duv02.y) * inv;
        dpdv = (dp02 * -duv12.x + dp12 * duv02.x) * inv;
    }
    if (degenerateuv || length_squared(cross(dpdu, dpdv)) == 0.0f) {
        V3 ng = cross(p2 - p0, p1 - p0);
        coordinate_system(normalize(ng), &dpdu, &dpdv);
    }
    Float xs = std::fabs(b0 * p0.x) + std::fabs(b1 * p1.x) + std::fabs(b2 * p2.x);
    Float ys = std::fabs(b0 * p0.y) + std::fabs(b1 * p1.y) + std::fabs(b2 * p2.y);
    Float zs = std::fabs(b0 * p0.z) + std::fabs(b1 * p1.z) + std::fabs(b2 * p2.z);
    V3 perror = V3(xs, ys, zs) * gamma(7);
    V3 phit = p0 * b0 + p1 * b1 + p2 * b2;
    P2 uvhit(uv[0].x * b0 + uv[1].x * b1 + uv[2].x * b2, uv[0].y * b0 + uv[1].y * b1 + uv[2].y * b2);
    bool ro = pr.flags & PBRT_B200_PRIM_REVERSE_ORIENTATION, sh = pr.flags & PBRT_B200_PRIM_SWAPS_HANDEDNESS;
    bool flip = ro ^ sh;
    SurfaceInteraction isect = si_new(phit, perror, uvhit, -r.d, dpdu, dpdv, r.time, shape_some && flip);
    isect.shape_some = shape_some; isect.shape_flip = flip;
    V3 nn = normalize(cross(dp02, dp12));
    isect.n = nn; isect.sh_n = nn;
    isect.wo = -r.d;  // triangle.rs:296 (NOT normalised)
    if (flip) { isect.n = -nn; isect.sh_n = -nn; }
    bool has_n = (pr.flags & PBRT_B200_PRIM_HAS_N) && s.d.vertex_n, has_s = (pr.flags & PBRT_B200_PRIM_HAS_S) && s.d.vertex_s;
    if (has_n || has_s) {
        V3 ns;
        if (has_n) {
            ns = s.N(vi[0]) * b0 + s.N(vi[1]) * b1 + s.N(vi[2]) * b2;
            if (length_squared(ns) > 0.0f) ns = normalize(ns); else ns = isect.n;
        } else ns = isect.n;
        V3 ss;
        if (has_s) {
            ss = s.S(vi[0]) * b0 + s.S(vi[1]) * b1 + s.S(vi[2]) * b2;
            if (length_squared(ss) > 0.0f) ss = normalize(ss); else ss = normalize(isect.dpdu);
        } else ss = normalize(isect.dpdu);
        V3 ts = cross(ss, ns);
        if (length_squared(ts) > 0.0f) { ts = normalize(ts); ss = cross(ts, ns); }
        else coordinate_system(ns, &ss, &ts);
        // dndu / dndv of the interpolated normal, triangle.rs:339-377
        V3 dndu, dndv;
        if (has_n) {
            V3 dn1 = s.N(vi[0]) - s.N(vi[2]), dn2 = s.N(vi[1]) - s.N(vi[2]);
            Float det = duv02.x * duv12.y - duv02.y * duv12.x;
            if (std::fabs(det) < 1.0e-8f) {
                V3 dn = cross(s.N(vi[2]) - s.N(vi[0]), s.N(vi[1]) - s.N(vi[0]));
                if (length_squared(dn) != 0.0f) coordinate_system(dn, &dndu, &dndv);
            } else {
                Float invdet = 1.0f / det;
                dndu = (dn1 * duv12.y - dn2 * duv02.y) * invdet;
                dndv = (dn1 * -duv12.x + dn2 * duv02.x) * invdet;
            }
        }
        if (ro) ts = -ts;
        set_shading_geometry(isect, ss, ts, dndu, dndv, true);  // interaction.rs:234-255
    }
    return isect;
}

// Sphere::intersect tail, sphere.rs:100-196 (full sphere) incl. the world transform
// (transform.rs:607-636).  Shape is passed as None => no orientation flip (quirk a-Q8).
inline SurfaceInteraction sphere_interaction(const pbrt_b200_sphere& sp, const Ray& world_ray, Float t) {
    V3 o_err, d_err;
    M4 w2o = m4_from(sp.world_to_object), o2w = m4_from(sp.object_to_world);
    Ray ray = m4_ray_error(w2o, world_ray, &o_err, &d_err);
    const Float phi_max = radians(360.0f), theta_min = std::acos(-1.0f), theta_max = std::acos(1.0f);
    V3 p_hit = ray.o + ray.d * t;
    p_hit = p_hit * (sp.radius / distance(p_hit, V3(0, 0, 0)));
    if (p_hit.x == 0.0f && p_hit.y == 0.0f) p_hit.x = 1.0e-5f * sp.radius;
    Float phi = std::atan2(p_hit.y, p_hit.x);
    if (phi < 0.0f) phi += 2.0f * PI;
    Float u = phi / phi_max;
    Float theta = std::acos(clamp(p_hit.z / sp.radius, -1.0f, 1.0f));
    Float v = (theta - theta_min) / (theta_max - theta_min);
    Float zradius = std::sqrt(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
    Float inv_radius = 1.0f / zradius;
    Float cos_phi = p_hit.x * inv_radius, sin_phi = p_hit.y * inv_radius;
    V3 dpdu(-phi_max * p_hit.y, phi_max * p_hit.x, 0.0f);
    V3 dpdv = V3(p_hit.z * cos_phi, p_hit.z * sin_phi, -sp.radius * std::sin(theta)) * (theta_max - theta_min);
    // dndu / dndv from the fundamental forms, sphere.rs:166-184
    V3 d2pduu = V3(p_hit.x, p_hit.y, 0.0f) * -phi_max * phi_max;
    V3 d2pduv = V3(-sin_phi, cos_phi, 0.0f) * (theta_max - theta_min) * p_hit.z * phi_max;
    V3 d2pdvv = V3(p_hit.x, p_hit.y, p_hit.z) * -(theta_max - theta_min) * (theta_max - theta_min);
    Float E = dot(dpdu, dpdu), F = dot(dpdu, dpdv), G = dot(dpdv, dpdv);
    V3 Nn = normalize(cross(dpdu, dpdv));
    Float e = dot(Nn, d2pduu), f = dot(Nn, d2pduv), g = dot(Nn, d2pdvv);
    Float inv_EGF2 = 1.0f / (E * G - F * F);
    V3 dndu = dpdu * (f * F - e * G) * inv_EGF2 + dpdv * (e * F - f * E) * inv_EGF2;
    V3 dndv = dpdu * (g * F - f * G) * inv_EGF2 + dpdv * (f * F - g * E) * inv_EGF2;
    V3 p_error = vabs(p_hit) * gamma(5);
    SurfaceInteraction s = si_new(p_hit, p_error, P2(u, v), -ray.d, dpdu, dpdv, ray.time, false);
    // transform_surface_interaction
    SurfaceInteraction ret;
    ret.dndu = ret.sh_dndu = m4_normal(w2o, dndu); ret.dndv = ret.sh_dndv = m4_normal(w2o, dndv);
    ret.p = m4_point_abs_error(o2w, s.p, s.p_error, &ret.p_error);
    ret.n = normalize(m4_normal(w2o, s.n));
    ret.wo = normalize(m4_vector(o2w, s.wo));
    ret.time = s.time; ret.uv = s.uv;
    ret.dpdu = m4_vector(o2w, s.dpdu); ret.dpdv = m4_vector(o2w, s.dpdv);
    ret.sh_n = normalize(m4_normal(w2o, s.sh_n));
    ret.sh_dpdu = m4_vector(o2w, s.sh_dpdu); ret.sh_dpdv = m4_vector(o2w, s.sh_dpdv);
    ret.sh_n = face_forward(ret.sh_n, ret.n);
    return ret;
}

// Transform::transform_surface_interaction, transform.rs:607-636 (the fields the path integrator reads)
inline SurfaceInteraction m4_surface_interaction(const M4& m, const M4& m_inv, const SurfaceInteraction& s) {
    SurfaceInteraction ret;
    ret.p = m4_point_abs_error(m, s.p, s.p_error, &ret.p_error);
    ret.n = normalize(m4_normal(m_inv, s.n));
    ret.wo = normalize(m4_vector(m, s.wo));
    ret.time = s.time; ret.uv = s.uv;
    ret.dpdu = m4_vector(m, s.dpdu); ret.dpdv = m4_vector(m, s.dpdv);
    ret.sh_n = normalize(m4_normal(m_inv, s.sh_n));
    ret.sh_dpdu = m4_vector(m, s.sh_dpdu); ret.sh_dpdv = m4_vector(m, s.sh_dpdv);
    ret.sh_n = face_forward(ret.sh_n, ret.n);
    ret.dndu = m4_normal(m_inv, s.dndu); ret.dndv = m4_normal(m_inv, s.dndv);
    ret.sh_dndu = m4_normal(m_inv, s.sh_dndu); ret.sh_dndv = m4_normal(m_inv, s.sh_dndv);
    ret.shape_some = s.shape_some; ret.shape_flip = s.shape_flip;
    ret.slot = s.slot;
    return ret;
}
inline bool m4_is_identity(const M4& m) {  // transform.rs:229-238
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) if (m.m[i][j] != (i == j ? 1.0f : 0.0f)) return false;
    return true;
}

inline SurfaceInteraction make_interaction(const SceneView& s, const Ray& r, const Hit& h) {
    const pbrt_b200_prim& pr = s.d.prims[h.slot];
    if (h.inst != PBRT_B200_NO_HIT) {  // TransformedPrimitive::intersect, primitive.rs:58-80
        const pbrt_b200_instance& in = s.d.instances[h.inst];
        M4 p2w = m4_from(in.prim_to_world), w2p = m4_from(in.world_to_prim);
        Ray ray = m4_ray(w2p, r);
        SurfaceInteraction si = (pr.shape_kind == PBRT_B200_SHAPE_TRIANGLE) ? triangle_interaction(s, pr, ray, h.b0, h.b1, h.b2, true)
                                                                            : sphere_interaction(s.d.spheres[pr.shape_index], ray, h.t);
        si.slot = h.slot;
        if (!m4_is_identity(p2w)) si = m4_surface_interaction(p2w, w2p, si);
        return si;
    }
    SurfaceInteraction si = (pr.shape_kind == PBRT_B200_SHAPE_TRIANGLE) ? triangle_interaction(s, pr, r, h.b0, h.b1, h.b2, true)
                                                                        : sphere_interaction(s.d.spheres[pr.shape_index], r, h.t);
    si.slot = h.slot;
    return si;
}

}  // namespace orc
