// C-ABI entry points: scene upload and the batch intersect API (include/pbrt_b200.h).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "error.h"
#include "scene.cuh"
#include "trace.cuh"
#include "util.cuh"

using namespace pb;

namespace pb {
void render_release_scene_state(pbrt_b200_scene* scene);  // render.cu
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
// Persistent warps pull rays from the batch (trace.cuh: trace_queue); rays are 32-byte records
// read as two float4 (128-bit loads), hits are written as one uint4.
struct BatchClosestJob {
    const float4* rays; uint4* hits; const float4* tris;
    PB_D bool load(uint32_t i, f3* o, f3* d, float* t_max) const {
        float4 a = __ldg(rays + 2ull * i), b = __ldg(rays + 2ull * i + 1);
        *o = f3(a.x, a.y, a.z); *d = f3(b.x, b.y, b.z); *t_max = a.w;
        return true;
    }
    PB_D void store(uint32_t i, const TravRay& r) const {
        uint4 out;
        if (r.found) {
            out.x = __float_as_uint(__ldg(tris + 3ull * r.hit.slot).w);  // creation_index
            out.y = __float_as_uint(r.hit.t); out.z = __float_as_uint(r.hit.b0); out.w = __float_as_uint(r.hit.b1);
        } else {
            out.x = PBRT_B200_NO_HIT; out.y = __float_as_uint(r.hit.t); out.z = 0u; out.w = 0u;
        }
        hits[i] = out;
    }
};
struct BatchAnyJob {
    const float4* rays; uint8_t* occluded;
    PB_D bool load(uint32_t i, f3* o, f3* d, float* t_max) const {
        float4 a = __ldg(rays + 2ull * i), b = __ldg(rays + 2ull * i + 1);
        *o = f3(a.x, a.y, a.z); *d = f3(b.x, b.y, b.z); *t_max = a.w;
        return true;
    }
    PB_D void store(uint32_t i, const TravRay& r) const { occluded[i] = r.found ? 1 : 0; }
};

__global__ void __launch_bounds__(PB_TRACE_BLOCK) k_intersect_batch(DevScene s, const float4* __restrict__ rays, uint32_t n, uint4* __restrict__ hits,
                                                                   uint32_t* fetch, TraceTune tune) {
    BatchClosestJob job{rays, hits, s.tris};
    trace_queue<false>(s, job, n, fetch, tune);
}

// A/B variants for tuning (tools/trace_ab.py): one ray per thread, while-while or if-if
template <int VARIANT>
__global__ void __launch_bounds__(PB_TRACE_BLOCK) k_intersect_batch_1rpt(DevScene s, const float4* __restrict__ rays, uint32_t n, uint4* __restrict__ hits) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    BatchClosestJob job{rays, hits, s.tris};
    f3 o, d; float t_max;
    job.load(i, &o, &d, &t_max);
    TravRay r;
    r.found = VARIANT == 0 ? traverse<false>(s, o, d, t_max, &r.hit) : traverse_ifif<false>(s, o, d, t_max, &r.hit);
    job.store(i, r);
}

__global__ void __launch_bounds__(PB_TRACE_BLOCK) k_intersect_p_batch(DevScene s, const float4* __restrict__ rays, uint32_t n, uint8_t* __restrict__ occluded,
                                                                     uint32_t* fetch) {
    BatchAnyJob job{rays, occluded};
    trace_queue<true>(s, job, n, fetch);
}

// ---------------------------------------------------------------------------
// scene
// ---------------------------------------------------------------------------
namespace {

template <typename T>
int upload(pbrt_b200_scene* sc, const T* host, uint64_t count, const T** dev_out) {
    *dev_out = nullptr;
    if (count == 0 || host == nullptr) return PBRT_B200_OK;
    void* d = nullptr;
    PB_CUDA_TRY(cudaMalloc(&d, count * sizeof(T)));
    sc->allocs[sc->n_allocs++] = d;
    sc->device_bytes += count * sizeof(T);
    PB_CUDA_TRY(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *dev_out = reinterpret_cast<const T*>(d);
    return PBRT_B200_OK;
}

int validate(const pbrt_b200_scene_desc* d) {
    if (d->abi_version != PBRT_B200_ABI_VERSION) return fail(PBRT_B200_ERR_INVALID, "scene_create: abi_version mismatch");
    if (d->n_prims && !d->prims) return fail(PBRT_B200_ERR_INVALID, "scene_create: prims is null");
    if (d->n_nodes && !d->nodes) return fail(PBRT_B200_ERR_INVALID, "scene_create: nodes is null");
    if (d->n_prims && !d->n_nodes) return fail(PBRT_B200_ERR_INVALID, "scene_create: primitives without a BVH (Accelerator \"bvh\" is required)");
    if (d->n_prims > 0x7ffffff0ull) return fail(PBRT_B200_ERR_INVALID, "scene_create: too many primitives");
    for (uint64_t i = 0; i < d->n_prims; ++i) {
        const pbrt_b200_prim& p = d->prims[i];
        if (p.shape_kind == PBRT_B200_SHAPE_TRIANGLE) {
            if (p.shape_index >= d->n_triangles) return fail(PBRT_B200_ERR_INVALID, "scene_create: triangle index out of range");
            const uint32_t* ix = d->tri_indices + 3ull * p.shape_index;
            if (ix[0] >= d->n_vertices || ix[1] >= d->n_vertices || ix[2] >= d->n_vertices)
                return fail(PBRT_B200_ERR_INVALID, "scene_create: vertex index out of range");
        } else if (p.shape_kind == PBRT_B200_SHAPE_SPHERE) {
            if (p.shape_index >= d->n_spheres) return fail(PBRT_B200_ERR_INVALID, "scene_create: sphere index out of range");
        } else {
            return fail(PBRT_B200_ERR_UNSUPPORTED, "scene_create: shape kind outside the hot path (triangle, sphere)");
        }
        if (p.material >= (int64_t)d->n_materials) return fail(PBRT_B200_ERR_INVALID, "scene_create: material index out of range");
        if (p.area_light >= (int64_t)d->n_lights) return fail(PBRT_B200_ERR_INVALID, "scene_create: area light index out of range");
    }
    for (uint64_t i = 0; i < d->n_materials; ++i)
        if (d->materials[i].type > PBRT_B200_MAT_METAL) return fail(PBRT_B200_ERR_UNSUPPORTED, "scene_create: material outside the hot path");
    for (uint64_t i = 0; i < d->n_lights; ++i) {
        const pbrt_b200_light& l = d->lights[i];
        if (l.type > PBRT_B200_LIGHT_INFINITE) return fail(PBRT_B200_ERR_UNSUPPORTED, "scene_create: light type outside the hot path");
        if (l.type == PBRT_B200_LIGHT_DIFFUSE && l.shape_kind != PBRT_B200_SHAPE_TRIANGLE)
            return fail(PBRT_B200_ERR_UNSUPPORTED, "scene_create: only triangle area lights are on the hot path");
    }
    return PBRT_B200_OK;
}

// LinearBVHNode[] -> fat nodes (see scene.cuh).  Also checks what the reference would
// panic on (stack deeper than 64, bvh.rs:722).
int build_fat_nodes(const pbrt_b200_scene_desc* d, std::vector<float4>& fat, std::vector<uint8_t>& last_flag, uint32_t* root_ref) {
    const uint64_t nn = d->n_nodes;
    *root_ref = PB_REF_NONE;
    if (nn == 0) return PBRT_B200_OK;
    std::vector<uint32_t> fat_index(nn, 0xffffffffu);
    uint32_t nfat = 0;
    for (uint64_t i = 0; i < nn; ++i)
        if (d->nodes[i].n_prims == 0) fat_index[i] = nfat++;
    fat.assign(4ull * nfat, make_float4(0, 0, 0, 0));
    auto ref_of = [&](uint64_t c, uint32_t* out) -> bool {
        if (c >= nn) return false;
        const pbrt_b200_bvh_node& n = d->nodes[c];
        if (n.n_prims > 0) {
            if ((uint64_t)n.offset + n.n_prims > d->n_prims) return false;
            last_flag[n.offset + n.n_prims - 1] = 1;
            *out = PB_LEAF_BIT | n.offset;
        } else {
            *out = fat_index[c];
        }
        return true;
    };
    if (!ref_of(0, root_ref)) return fail(PBRT_B200_ERR_INVALID, "scene_create: root node refers past the primitive table");
    for (uint64_t i = 0; i < nn; ++i) {
        const pbrt_b200_bvh_node& n = d->nodes[i];
        if (n.n_prims != 0) continue;
        uint64_t c0 = i + 1, c1 = n.offset;
        uint32_t r0, r1;
        if (n.axis > 2 || c1 <= i || !ref_of(c0, &r0) || !ref_of(c1, &r1))
            return fail(PBRT_B200_ERR_INVALID, "scene_create: malformed LinearBVHNode array");
        const float* a = d->nodes[c0].bounds;
        const float* b = d->nodes[c1].bounds;
        float4* q = &fat[4ull * fat_index[i]];
        q[0] = make_float4(a[0], a[1], a[2], a[3]);
        q[1] = make_float4(a[4], a[5], b[0], b[1]);
        q[2] = make_float4(b[2], b[3], b[4], b[5]);
        uint32_t ax = n.axis, z = 0;
        float4 m;
        std::memcpy(&m.x, &r0, 4); std::memcpy(&m.y, &r1, 4); std::memcpy(&m.z, &ax, 4); std::memcpy(&m.w, &z, 4);
        q[3] = m;
    }
    // depth check (explicit stack; the pending-entry count equals the tree depth)
    {
        std::vector<std::pair<uint64_t, int>> st;
        st.push_back({0, 0});
        int maxd = 0;
        while (!st.empty()) {
            auto e = st.back(); st.pop_back();
            maxd = e.second > maxd ? e.second : maxd;
            const pbrt_b200_bvh_node& n = d->nodes[e.first];
            if (n.n_prims == 0) { st.push_back({e.first + 1, e.second + 1}); st.push_back({n.offset, e.second + 1}); }
        }
        if (maxd >= PB_STACK_DEPTH) return fail(PBRT_B200_ERR_INVALID, "scene_create: BVH deeper than the reference's 64-entry traversal stack");
    }
    return PBRT_B200_OK;
}

}  // namespace

extern "C" int pbrt_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" void pbrt_b200_scene_destroy(pbrt_b200_scene* sc) {
    if (!sc) return;
    cudaSetDevice(sc->device);
    render_release_scene_state(sc);
    for (int i = 0; i < sc->n_allocs; ++i) cudaFree(sc->allocs[i]);
    if (sc->scratch) cudaFree(sc->scratch);
    if (sc->fetch_counter) cudaFree(sc->fetch_counter);
    delete sc;
}

extern "C" int pbrt_b200_scene_create(const pbrt_b200_scene_desc* d, int device, pbrt_b200_scene** out) {
    if (!d || !out) return fail(PBRT_B200_ERR_INVALID, "scene_create: null argument");
    *out = nullptr;
    int rc = validate(d);
    if (rc) return rc;
    int ndev = pbrt_b200_device_count();
    if (ndev <= 0) return fail(PBRT_B200_ERR_NO_DEVICE, "scene_create: no CUDA device visible; this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(PBRT_B200_ERR_INVALID, "scene_create: device ordinal out of range");
    PB_CUDA_TRY(cudaSetDevice(device));

    std::vector<float4> fat;
    std::vector<uint8_t> last(d->n_prims, 0);
    uint32_t root_ref;
    rc = build_fat_nodes(d, fat, last, &root_ref);
    if (rc) return rc;

    // leaf records: vertices gathered into BVH slot order
    std::vector<float4> tris(3ull * d->n_prims);
    for (uint64_t s = 0; s < d->n_prims; ++s) {
        const pbrt_b200_prim& p = d->prims[s];
        uint32_t fl = (p.flags & PB_TRI_FLAGS_MASK) | (last[s] ? PB_TRI_LAST : 0u);
        float4 v[3] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
        if (p.shape_kind == PBRT_B200_SHAPE_TRIANGLE) {
            const uint32_t* ix = d->tri_indices + 3ull * p.shape_index;
            for (int k = 0; k < 3; ++k) {
                const float* vp = d->vertex_p + 3ull * ix[k];
                v[k].x = vp[0]; v[k].y = vp[1]; v[k].z = vp[2];
            }
        } else {
            fl |= PB_TRI_SPHERE;
        }
        std::memcpy(&v[0].w, &p.creation_index, 4);
        std::memcpy(&v[1].w, &fl, 4);
        std::memcpy(&v[2].w, &p.shape_index, 4);
        tris[3 * s] = v[0]; tris[3 * s + 1] = v[1]; tris[3 * s + 2] = v[2];
    }

    pbrt_b200_scene* sc = new pbrt_b200_scene();
    sc->device = device;
    sc->n_prims = d->n_prims; sc->n_nodes = d->n_nodes;
    std::memset(&sc->dev, 0, sizeof sc->dev);
    DevScene& ds = sc->dev;
    ds.root_ref = root_ref;
    ds.n_fat = (uint32_t)(fat.size() / 4);
    ds.n_slots = (uint32_t)d->n_prims;
    ds.n_lights = (uint32_t)d->n_lights;
    ds.n_materials = (uint32_t)d->n_materials;
    if (d->n_nodes) {
        std::memcpy(ds.root_box, d->nodes[0].bounds, sizeof ds.root_box);
        // Bounds3f::bounding_sphere (bounds.rs:515-523) on Scene.wb, used by DistantLight::preprocess
        float c[3];
        for (int k = 0; k < 3; ++k) c[k] = (ds.root_box[k] + ds.root_box[3 + k]) * (1.0f / 2.0f);
        bool inside = true;
        for (int k = 0; k < 3; ++k) inside = inside && c[k] >= ds.root_box[k] && c[k] <= ds.root_box[3 + k];
        float dx = ds.root_box[3] - c[0], dy = ds.root_box[4] - c[1], dz = ds.root_box[5] - c[2];
        ds.world_center[0] = c[0]; ds.world_center[1] = c[1]; ds.world_center[2] = c[2];
        ds.world_radius = inside ? sqrtf(dx * dx + dy * dy + dz * dz) : 0.0f;
    }
#define PB_UP(expr) do { rc = (expr); if (rc) { pbrt_b200_scene_destroy(sc); return rc; } } while (0)
    PB_UP(upload(sc, fat.data(), fat.size(), &ds.nodes));
    PB_UP(upload(sc, tris.data(), tris.size(), &ds.tris));
    PB_UP(upload(sc, d->prims, d->n_prims, &ds.prims));
    PB_UP(upload(sc, d->vertex_p, 3 * d->n_vertices, &ds.vertex_p));
    PB_UP(upload(sc, d->vertex_n, d->vertex_n ? 3 * d->n_vertices : 0, &ds.vertex_n));
    PB_UP(upload(sc, d->vertex_s, d->vertex_s ? 3 * d->n_vertices : 0, &ds.vertex_s));
    PB_UP(upload(sc, d->vertex_uv, d->vertex_uv ? 2 * d->n_vertices : 0, &ds.vertex_uv));
    PB_UP(upload(sc, d->tri_indices, 3 * d->n_triangles, &ds.tri_indices));
    PB_UP(upload(sc, d->spheres, d->n_spheres, &ds.spheres));
    PB_UP(upload(sc, d->materials, d->n_materials, &ds.materials));
    PB_UP(upload(sc, d->lights, d->n_lights, &ds.lights));
#undef PB_UP
    *out = sc;
    return PBRT_B200_OK;
}

extern "C" int pbrt_b200_scene_world_bound(const pbrt_b200_scene* sc, float* b) {
    if (!sc || !b) return fail(PBRT_B200_ERR_INVALID, "scene_world_bound: null argument");
    std::memcpy(b, sc->dev.root_box, 6 * sizeof(float));
    return PBRT_B200_OK;
}

// ---------------------------------------------------------------------------
// batch intersect
// ---------------------------------------------------------------------------
// tunables for A/B measurement of the batch closest-hit kernel: [variant, refill_below, chunk, grid]
static int g_tune[8] = {0, PB_REFILL_BELOW, PB_FETCH_CHUNK, 0, PB_INTERIOR_MIN, 0, 0, 0};
extern "C" void pbrt_b200_debug_tune(int key, int value) { if (key >= 0 && key < 8) g_tune[key] = value; }

namespace {
int ensure_fetch_counter(pbrt_b200_scene* sc) {
    if (sc->fetch_counter) return PBRT_B200_OK;
    PB_CUDA_TRY(cudaMalloc((void**)&sc->fetch_counter, 64));
    int sm = 148, per_sm = 8;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, sc->device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_intersect_batch, PB_TRACE_BLOCK, 0);
    sc->trace_grid = sm * (per_sm > 0 ? per_sm : 1);  // persistent: one resident wave of CTAs
    return PBRT_B200_OK;
}
}  // namespace

extern "C" int pbrt_b200_intersect_dev(pbrt_b200_scene* sc, const pbrt_b200_ray* rays, uint64_t n, pbrt_b200_hit* hits, void* stream) {
    if (!sc || (n && (!rays || !hits))) return fail(PBRT_B200_ERR_INVALID, "intersect_dev: null argument");
    if (n == 0) return PBRT_B200_OK;
    PB_CUDA_TRY(cudaSetDevice(sc->device));
    if (n > 0xfffffff0ull) return fail(PBRT_B200_ERR_INVALID, "intersect_dev: batch too large (split it)");
    int rc = ensure_fetch_counter(sc);
    if (rc) return rc;
    PB_CUDA_TRY(cudaMemsetAsync(sc->fetch_counter, 0, sizeof(uint32_t), (cudaStream_t)stream));
    if (g_tune[0] == 1 || g_tune[0] == 2) {
        unsigned blocks = (unsigned)((n + PB_TRACE_BLOCK - 1) / PB_TRACE_BLOCK);
        if (g_tune[0] == 1) k_intersect_batch_1rpt<0><<<blocks, PB_TRACE_BLOCK, 0, (cudaStream_t)stream>>>(sc->dev, reinterpret_cast<const float4*>(rays), (uint32_t)n, reinterpret_cast<uint4*>(hits));
        else k_intersect_batch_1rpt<1><<<blocks, PB_TRACE_BLOCK, 0, (cudaStream_t)stream>>>(sc->dev, reinterpret_cast<const float4*>(rays), (uint32_t)n, reinterpret_cast<uint4*>(hits));
    } else {
        TraceTune tune{g_tune[1], g_tune[2], g_tune[4]};
        int grid = g_tune[3] > 0 ? g_tune[3] : sc->trace_grid;
        k_intersect_batch<<<grid, PB_TRACE_BLOCK, 0, (cudaStream_t)stream>>>(sc->dev, reinterpret_cast<const float4*>(rays), (uint32_t)n,
                                                                             reinterpret_cast<uint4*>(hits), sc->fetch_counter, tune);
    }
    PB_CUDA_TRY(cudaGetLastError());
    return PBRT_B200_OK;
}

extern "C" int pbrt_b200_intersect_p_dev(pbrt_b200_scene* sc, const pbrt_b200_ray* rays, uint64_t n, uint8_t* occluded, void* stream) {
    if (!sc || (n && (!rays || !occluded))) return fail(PBRT_B200_ERR_INVALID, "intersect_p_dev: null argument");
    if (n == 0) return PBRT_B200_OK;
    PB_CUDA_TRY(cudaSetDevice(sc->device));
    if (n > 0xfffffff0ull) return fail(PBRT_B200_ERR_INVALID, "intersect_p_dev: batch too large (split it)");
    int rc = ensure_fetch_counter(sc);
    if (rc) return rc;
    PB_CUDA_TRY(cudaMemsetAsync(sc->fetch_counter, 0, sizeof(uint32_t), (cudaStream_t)stream));
    k_intersect_p_batch<<<sc->trace_grid, PB_TRACE_BLOCK, 0, (cudaStream_t)stream>>>(sc->dev, reinterpret_cast<const float4*>(rays), (uint32_t)n, occluded,
                                                                                    sc->fetch_counter);
    PB_CUDA_TRY(cudaGetLastError());
    return PBRT_B200_OK;
}

namespace {
int ensure_scratch(pbrt_b200_scene* sc, uint64_t bytes) {
    if (sc->scratch_bytes >= bytes) return PBRT_B200_OK;
    if (sc->scratch) cudaFree(sc->scratch);
    sc->scratch = nullptr; sc->scratch_bytes = 0;
    PB_CUDA_TRY(cudaMalloc(&sc->scratch, bytes));
    sc->scratch_bytes = bytes;
    return PBRT_B200_OK;
}
}  // namespace

extern "C" int pbrt_b200_intersect(pbrt_b200_scene* sc, const pbrt_b200_ray* rays, uint64_t n, pbrt_b200_hit* hits) {
    if (!sc || (n && (!rays || !hits))) return fail(PBRT_B200_ERR_INVALID, "intersect: null argument");
    if (n == 0) return PBRT_B200_OK;
    PB_CUDA_TRY(cudaSetDevice(sc->device));
    int rc = ensure_scratch(sc, n * (sizeof(pbrt_b200_ray) + sizeof(pbrt_b200_hit)));
    if (rc) return rc;
    pbrt_b200_ray* dr = reinterpret_cast<pbrt_b200_ray*>(sc->scratch);
    pbrt_b200_hit* dh = reinterpret_cast<pbrt_b200_hit*>(reinterpret_cast<char*>(sc->scratch) + n * sizeof(pbrt_b200_ray));
    PB_CUDA_TRY(cudaMemcpyAsync(dr, rays, n * sizeof(pbrt_b200_ray), cudaMemcpyHostToDevice, 0));
    rc = pbrt_b200_intersect_dev(sc, dr, n, dh, nullptr);
    if (rc) return rc;
    PB_CUDA_TRY(cudaMemcpyAsync(hits, dh, n * sizeof(pbrt_b200_hit), cudaMemcpyDeviceToHost, 0));
    PB_CUDA_TRY(cudaStreamSynchronize(0));
    return PBRT_B200_OK;
}

extern "C" int pbrt_b200_intersect_p(pbrt_b200_scene* sc, const pbrt_b200_ray* rays, uint64_t n, uint8_t* occluded) {
    if (!sc || (n && (!rays || !occluded))) return fail(PBRT_B200_ERR_INVALID, "intersect_p: null argument");
    if (n == 0) return PBRT_B200_OK;
    PB_CUDA_TRY(cudaSetDevice(sc->device));
    int rc = ensure_scratch(sc, n * (sizeof(pbrt_b200_ray) + 1));
    if (rc) return rc;
    pbrt_b200_ray* dr = reinterpret_cast<pbrt_b200_ray*>(sc->scratch);
    uint8_t* dz = reinterpret_cast<uint8_t*>(sc->scratch) + n * sizeof(pbrt_b200_ray);
    PB_CUDA_TRY(cudaMemcpyAsync(dr, rays, n * sizeof(pbrt_b200_ray), cudaMemcpyHostToDevice, 0));
    rc = pbrt_b200_intersect_p_dev(sc, dr, n, dz, nullptr);
    if (rc) return rc;
    PB_CUDA_TRY(cudaMemcpyAsync(occluded, dz, n, cudaMemcpyDeviceToHost, 0));
    PB_CUDA_TRY(cudaStreamSynchronize(0));
    return PBRT_B200_OK;
}

// Film::write_image arithmetic, core/film.rs:217-264 (host helper; no device work)
extern "C" int pbrt_b200_film_resolve(const float* rgbw, uint64_t npixels, float scale, float* rgb_out) {
    if (npixels && (!rgbw || !rgb_out)) return fail(PBRT_B200_ERR_INVALID, "film_resolve: null argument");
    for (uint64_t i = 0; i < npixels; ++i) {
        const float* p = rgbw + 4 * i;
        // merge_film_tile: tile RGB -> XYZ (spectrum.rs:495-504)
        float X = 0.412453f * p[0] + 0.357580f * p[1] + 0.180423f * p[2];
        float Y = 0.212671f * p[0] + 0.715160f * p[1] + 0.072169f * p[2];
        float Z = 0.019334f * p[0] + 0.119193f * p[1] + 0.950227f * p[2];
        // write_image: XYZ -> RGB (spectrum.rs:484-493)
        float r = 3.240479f * X - 1.537150f * Y - 0.498535f * Z;
        float g = -0.969256f * X + 1.875991f * Y + 0.041556f * Z;
        float b = 0.055648f * X - 0.204043f * Y + 1.057311f * Z;
        float w = p[3];
        if (w != 0.0f) {
            float inv = 1.0f / w;
            r = fmaxf(r * inv, 0.0f); g = fmaxf(g * inv, 0.0f); b = fmaxf(b * inv, 0.0f);
        }
        rgb_out[3 * i] = r * scale; rgb_out[3 * i + 1] = g * scale; rgb_out[3 * i + 2] = b * scale;
    }
    return PBRT_B200_OK;
}
