// TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see oracle_math.hpp header).
// oracle_sampling.hpp: RNG, Distribution1D, sampling routines, Sobol/Halton/(0,2) samplers.
#pragma once
#include <vector>
#include <memory>
#include <stdexcept>
#include "oracle_math.hpp"
#include "../include/pbrt_b200.h"

namespace orc {

// ---- PCG32, src/core/rng.rs:25-76
struct RNG {
    uint64_t state = 0x853c49e6748fea9bULL, inc = 0xda3e39cb94b95bdbULL;
    RNG() {}
    explicit RNG(uint64_t seq) { set_sequence(seq); }
    uint32_t uniform_int32() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18) ^ old) >> 27);
        uint32_t rot = (uint32_t)(old >> 59);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    uint32_t uniform_int32_2(uint32_t b) {  // rng.rs:52-62
        uint32_t threshold = (~b + 1u) % b;
        for (;;) { uint32_t r = uniform_int32(); if (r >= threshold) return r % b; }
    }
    Float uniform_float() { return std::fmin(ONE_MINUS_EPSILON, (Float)uniform_int32() * 0x1.0p-32f); }
    void set_sequence(uint64_t initseq) {
        state = 0; inc = (initseq << 1) | 1;
        uniform_int32();
        state += 0x853c49e6748fea9bULL;
        uniform_int32();
    }
};

// src/core/sampling.rs:178-186
template <typename T> inline void shuffle(T* samp, size_t count, size_t ndim, RNG& rng) {
    for (size_t i = 0; i < count; ++i) {
        size_t other = i + rng.uniform_int32_2((uint32_t)(count - i));
        for (size_t j = 0; j < ndim; ++j) std::swap(samp[ndim * i + j], samp[ndim * other + j]);
    }
}

// ---- Distribution1D, src/core/sampling.rs:6-92
struct Distribution1D {
    std::vector<Float> func, cdf;
    Float func_int = 0;
    Distribution1D() {}
    explicit Distribution1D(const std::vector<Float>& f) : func(f) {
        size_t n = f.size();
        cdf.assign(n + 1, 0.0f);
        for (size_t i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + func[i - 1] / (Float)n;
        func_int = cdf[n];
        if (func_int == 0.0f) { for (size_t i = 1; i < n + 1; ++i) cdf[i] = (Float)i / (Float)n; }
        else { for (size_t i = 1; i < n + 1; ++i) cdf[i] /= func_int; }
    }
    size_t count() const { return func.size(); }
    size_t sample_discrete(Float u, Float* pdf) const {
        size_t offset = (size_t)find_interval((int)cdf.size(), [&](int i) { return cdf[i] <= u; });
        if (pdf) *pdf = func_int > 0.0f ? func[offset] / (func_int * (Float)count()) : 0.0f;
        return offset;
    }
    Float sample_continuous(Float u, Float* pdf, size_t* off) const {
        size_t offset = (size_t)find_interval((int)cdf.size(), [&](int i) { return cdf[i] <= u; });
        if (off) *off = offset;
        Float du = u - cdf[offset];
        Float diff = cdf[offset + 1] - cdf[offset];
        if (diff > 0.0f) du /= diff;
        if (pdf) *pdf = func_int > 0.0f ? func[offset] / func_int : 0.0f;
        return ((Float)offset + du) / (Float)count();
    }
    Float discrete_pdf(size_t index) const { return func[index] / (func_int * (Float)count()); }
};

// src/core/sampling.rs:154-176
inline P2 concentric_sample_disk(P2 u) {
    Float ox = u.x * 2.0f - 1.0f, oy = u.y * 2.0f - 1.0f;
    if (ox == 0.0f && oy == 0.0f) return P2(0, 0);
    Float theta, r;
    if (std::fabs(ox) > std::fabs(oy)) { r = ox; theta = PI_OVER4 * (oy / ox); }
    else { r = oy; theta = PI_OVER2 - PI_OVER4 * (ox / oy); }
    return P2(std::cos(theta) * r, std::sin(theta) * r);
}
// src/core/sampling.rs:188-193
inline V3 cosine_sample_hemisphere(P2 u) {
    P2 d = concentric_sample_disk(u);
    Float z = std::sqrt(std::fmax(0.0f, 1.0f - d.x * d.x - d.y * d.y));
    return V3(d.x, d.y, z);
}
// src/core/sampling.rs:244-248
inline P2 uniform_sample_triangle(P2 u) { Float su0 = std::sqrt(u.x); return P2(1.0f - su0, u.y * su0); }
// src/core/sampling.rs:328-333
inline Float power_heuristic(int nf, Float fpdf, int ng, Float gpdf) {
    Float f = (Float)nf * fpdf, g = (Float)ng * gpdf;
    return (f * f) / (f * f + g * g);
}

struct CameraSample { P2 pfilm, plens; Float time; };

struct SamplerTables {
    const uint32_t* sobol32 = nullptr;
    const uint64_t* vdc = nullptr;
    const uint64_t* vdc_inv = nullptr;
};

// ---- Sampler trait, src/core/sampler.rs:24-42 (2D sample arrays: DirectLightingIntegrator "all", directlighting.rs:61-76)
struct Sampler {
    uint64_t samples_per_pixel = 1;
    int px = 0, py = 0;  // current_pixel
    uint64_t current_pixel_sample_index = 0;
    virtual ~Sampler() {}
    virtual int round_count(int n) const { return n; }                                   // sampler.rs:34
    virtual void request_2d_array(int) { throw std::runtime_error("oracle: sample arrays are only restated for the global samplers (sobol, halton)"); }
    virtual bool get_2d_array(int, std::vector<P2>*) { return false; }                  // None: no arrays were requested
    virtual void start_pixel(int x, int y) = 0;
    virtual Float get_1d() = 0;
    virtual P2 get_2d() = 0;
    virtual bool start_next_sample() = 0;
    virtual bool set_sample_number(uint64_t n) = 0;
    virtual std::unique_ptr<Sampler> clone(int64_t seed) const = 0;
    // sampler.rs:170-180
    CameraSample get_camera_sample(int x, int y) {
        CameraSample cs;
        P2 u = get_2d();
        cs.pfilm = P2((Float)x + u.x, (Float)y + u.y);
        cs.time = get_1d();
        cs.plens = get_2d();
        return cs;
    }
};

// ---- GlobalSampler behaviour, src/core/sampler.rs:255-355 (ARRAY_START_DIM = array_end_dim = 5)
struct GlobalSampler : Sampler {
    size_t dimension = 0;
    uint64_t interval_sample_index = 0;
    static const size_t ARRAY_START_DIM = 5;
    size_t array_end_dim = 5;
    std::vector<int> samples_2d_array_sizes;  // request_2d_array, sampler.rs:116-124
    size_t array_2d_offset = 0;
    virtual uint64_t get_index_for_sample(uint64_t sample_num) = 0;
    virtual Float sample_dimension(uint64_t index, size_t dim) const = 0;
    // Dimensions the sampler's tables hold (1024 Sobol' matrices, 1000 primes).  The reference indexes past them and panics; the oracle
    // and the CUDA path both hand out 0.5 from there on, so that a very deep path ends the same way on both sides.
    virtual size_t dim_limit() const = 0;
    Float dim_value(uint64_t index, size_t dim) const { return dim < dim_limit() ? sample_dimension(index, dim) : 0.5f; }
    void request_2d_array(int n) override { samples_2d_array_sizes.push_back(n); }
    // get_2d_array, sampler.rs:149-166.  global_start_pixel! (sampler.rs:268-303) fills every array for every sample of the
    // pixel up front: element j of array i is sample_dimension(get_index_for_sample(j), 5 + 2i [+1]) -- a pure function of
    // (pixel, j, i), so it is evaluated here when it is asked for; the values are the same.
    bool get_2d_array(int n, std::vector<P2>* out) override {
        if (array_2d_offset == samples_2d_array_sizes.size()) return false;
        if (samples_2d_array_sizes[array_2d_offset] != n) throw std::runtime_error("oracle: get_2d_array size mismatch (the reference asserts)");
        size_t dim = ARRAY_START_DIM + 2 * array_2d_offset;
        out->resize(n);
        for (int k = 0; k < n; ++k) {
            uint64_t idx = get_index_for_sample(current_pixel_sample_index * (uint64_t)n + (uint64_t)k);
            Float y = sample_dimension(idx, dim + 1);
            Float x = sample_dimension(idx, dim);
            (*out)[k] = P2(x, y);
        }
        array_2d_offset += 1;
        return true;
    }
    void start_pixel(int x, int y) override {
        px = x; py = y; current_pixel_sample_index = 0;
        array_2d_offset = 0;
        dimension = 0;
        interval_sample_index = get_index_for_sample(0);
        array_end_dim = ARRAY_START_DIM + 2 * samples_2d_array_sizes.size();
    }
    bool start_next_sample() override {
        dimension = 0;
        interval_sample_index = get_index_for_sample(current_pixel_sample_index + 1);
        array_2d_offset = 0;
        current_pixel_sample_index += 1;
        return current_pixel_sample_index < samples_per_pixel;
    }
    bool set_sample_number(uint64_t n) override {
        dimension = 0;
        interval_sample_index = get_index_for_sample(n);
        array_2d_offset = 0;
        current_pixel_sample_index = n;
        return current_pixel_sample_index < samples_per_pixel;
    }
    Float get_1d() override {
        if (dimension >= ARRAY_START_DIM && dimension < array_end_dim) dimension = array_end_dim;
        Float r = dim_value(interval_sample_index, dimension);
        dimension += 1;
        return r;
    }
    P2 get_2d() override {
        if (dimension + 1 >= ARRAY_START_DIM && dimension < array_end_dim) dimension = array_end_dim;
        Float y = dim_value(interval_sample_index, dimension + 1);
        Float x = dim_value(interval_sample_index, dimension);
        dimension += 2;
        return P2(x, y);
    }
};

// src/core/pbrt.rs round_up_pow2_32 / log2_int
inline int32_t round_up_pow2_32(int32_t v) { v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; return v + 1; }
inline int log2_int(int32_t v) { return 31 - __builtin_clz((uint32_t)v); }

// src/core/lowdiscrepancy.rs:512-543
inline uint64_t sobol_interval_to_index(const SamplerTables& T, uint32_t m, uint64_t frame, int px, int py) {
    if (m == 0) return 0;
    uint32_t m2 = m << 1;
    uint64_t index = frame << m2;
    uint64_t delta = 0;
    for (int c = 0; frame != 0; frame >>= 1, ++c)
        if (frame & 1) delta ^= T.vdc[(m - 1) * 52 + c];
    uint64_t b = ((uint64_t)((uint32_t)px << m) | (uint64_t)(int64_t)py) ^ delta;
    for (int c = 0; b != 0; b >>= 1, ++c)
        if (b & 1) index ^= T.vdc_inv[(m - 1) * 52 + c];
    return index;
}
// src/core/lowdiscrepancy.rs:549-569
inline Float sobol_sample_float(const SamplerTables& T, uint64_t a, size_t dimension, uint32_t scramble) {
    uint32_t v = scramble;
    for (size_t i = dimension * 52; a != 0; a >>= 1, ++i)
        if (a & 1) v ^= T.sobol32[i];
    return std::fmin((Float)v * 0x1.0p-32f, ONE_MINUS_EPSILON);
}

// ---- SobolSampler, src/samplers/sobol.rs:34-117
struct SobolSampler : GlobalSampler {
    SamplerTables T;
    int sb[4];
    int32_t resolution, log2_resolution;
    SobolSampler(uint64_t spp, const int sample_bounds[4], const SamplerTables& t) : T(t) {
        samples_per_pixel = spp;  // NOT rounded up (quirk a-Q7)
        for (int i = 0; i < 4; ++i) sb[i] = sample_bounds[i];
        int dx = sb[2] - sb[0], dy = sb[3] - sb[1];
        resolution = round_up_pow2_32(std::max(dx, dy));
        log2_resolution = log2_int(resolution);
    }
    uint64_t get_index_for_sample(uint64_t n) override { return sobol_interval_to_index(T, (uint32_t)log2_resolution, n, px - sb[0], py - sb[1]); }
    size_t dim_limit() const override { return 1024; }  // NUM_SOBOL_DIMENSIONS
    Float sample_dimension(uint64_t index, size_t dim) const override {
        Float s = sobol_sample_float(T, index, dim, 0);
        if (dim == 0 || dim == 1) {
            s = s * (Float)resolution + (Float)sb[dim];
            s = clamp(s - (Float)(dim == 0 ? px : py), 0.0f, ONE_MINUS_EPSILON);
        }
        return s;
    }
    std::unique_ptr<Sampler> clone(int64_t) const override { return std::unique_ptr<Sampler>(new SobolSampler(*this)); }
};

}  // namespace orc
#include "oracle_sampling_extra.hpp"
