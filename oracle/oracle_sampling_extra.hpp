// TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see oracle_math.hpp header).
// oracle_sampling_extra.hpp: HaltonSampler, ZeroTwoSequenceSampler, make_sampler.
#pragma once
#include "oracle_sampling.hpp"

namespace orc {

// ---- primes / permutations: src/core/lowdiscrepancy.rs:9-192 (PRIMES, PRIME_SUMS), :359-378
struct HaltonTables {
    std::vector<uint32_t> primes, prime_sums;
    std::vector<uint16_t> perms;
    HaltonTables() {
        const int N = 1000;  // PRIME_TABLE_SIZE
        for (uint32_t c = 2; (int)primes.size() < N; ++c) {
            bool ok = true;
            for (uint32_t p : primes) { if (p * p > c) break; if (c % p == 0) { ok = false; break; } }
            if (ok) primes.push_back(c);
        }
        uint32_t sum = 0;
        for (uint32_t p : primes) { prime_sums.push_back(sum); sum += p; }
        // compute_radical_inverse_permutations(&mut RNG::default()), halton.rs:13-15
        RNG rng;
        perms.resize(sum);
        size_t off = 0;
        for (int i = 0; i < N; ++i) {
            for (uint32_t j = 0; j < primes[i]; ++j) perms[off + j] = (uint16_t)j;
            shuffle(&perms[off], primes[i], 1, rng);
            off += primes[i];
        }
    }
    static const HaltonTables& get() { static HaltonTables t; return t; }
};

inline uint32_t reverse_bits32(uint32_t n) {  // lowdiscrepancy.rs:381-389
    n = (n << 16) | (n >> 16);
    n = ((n & 0x00ff00ff) << 8) | ((n & 0xff00ff00) >> 8);
    n = ((n & 0x0f0f0f0f) << 4) | ((n & 0xf0f0f0f0) >> 4);
    n = ((n & 0x33333333) << 2) | ((n & 0xcccccccc) >> 2);
    n = ((n & 0x55555555) << 1) | ((n & 0xaaaaaaaa) >> 1);
    return n;
}
inline uint64_t reverse_bits64(uint64_t n) { return ((uint64_t)reverse_bits32((uint32_t)n) << 32) | (uint64_t)reverse_bits32((uint32_t)(n >> 32)); }

// radical_inverse(base_index, n): pbrt_macros/src/lib.rs:92-110, lowdiscrepancy.rs:398-414
inline Float radical_inverse(size_t base_index, uint64_t n) {
    if (base_index == 0) return (Float)reverse_bits64(n) * 0x1.0p-64f;  // no ONE_MINUS_EPSILON clamp in base 2
    const uint64_t base = HaltonTables::get().primes[base_index];
    Float inv_base = 1.0f / (Float)base, inv_basen = 1.0f;
    uint64_t rev = 0;
    while (n != 0) {
        uint64_t next = n / base, digit = n - next * base;
        rev = rev * base + digit;
        inv_basen *= inv_base;
        n = next;
    }
    return std::fmin((Float)rev * inv_basen, ONE_MINUS_EPSILON);
}
// lowdiscrepancy.rs:416-426
inline uint64_t inverse_radical_inverse(uint64_t base, uint64_t inverse, uint64_t ndigits) {
    uint64_t index = 0;
    for (uint64_t i = 0; i < ndigits; ++i) { uint64_t digit = inverse % base; inverse /= base; index = index * base + digit; }
    return index;
}
// lowdiscrepancy.rs:468-484
inline Float scrambled_radical_inverse(size_t base_index, uint64_t a, const uint16_t* perm) {
    const uint64_t base = HaltonTables::get().primes[base_index];
    Float inv_base = 1.0f / (Float)base, inv_basen = 1.0f;
    uint64_t rev = 0;
    while (a != 0) {
        uint64_t next = a / base, digit = a - next * base;
        rev = rev * base + perm[digit];
        inv_basen *= inv_base;
        a = next;
    }
    Float res = inv_basen * ((Float)rev + inv_base * (Float)perm[0] / (1.0f - inv_base));
    return std::fmin(res, ONE_MINUS_EPSILON);
}
inline int64_t mod_i64(int64_t a, int64_t b) { int64_t r = a - (a / b) * b; return r < 0 ? r + b : r; }  // pbrt.rs:226-236
inline void extended_gcd(int64_t a, int64_t b, int64_t* x, int64_t* y) {  // halton.rs:19-27
    if (b == 0) { *x = 1; *y = 0; return; }
    int64_t d = a / b, r1, r2;
    extended_gcd(b, a % b, &r1, &r2);
    *x = r2; *y = r1 - (d * r2);
}
inline int64_t multiplicative_inverse(int64_t a, int64_t n) { int64_t x, y; extended_gcd(a, n, &x, &y); return mod_i64(x, n); }

// ---- HaltonSampler, src/samplers/halton.rs:63-166
struct HaltonSampler : GlobalSampler {
    static const int K_MAX_RESOLUTION = 128;
    int64_t base_scales[2], base_exponents[2];
    uint64_t sample_stride;
    int64_t mult_inverse[2];
    uint64_t offset_for_current_pixel = 0;
    int64_t pixel_for_offset[2] = {0, 0};  // AtomicIsize::default() (NOT i32::MAX as in pbrt-v3)
    HaltonSampler(uint64_t spp, const int sb[4]) {
        samples_per_pixel = spp;
        int64_t res[2] = {sb[2] - sb[0], sb[3] - sb[1]};
        for (int i = 0; i < 2; ++i) {
            int64_t base = i == 0 ? 2 : 3, scale = 1, exp = 0;
            while (scale < std::min<int64_t>(res[i], K_MAX_RESOLUTION)) { scale *= base; exp += 1; }
            base_scales[i] = scale; base_exponents[i] = exp;
        }
        sample_stride = (uint64_t)(base_scales[0] * base_scales[1]);
        mult_inverse[0] = multiplicative_inverse(base_scales[1], base_scales[0]);
        mult_inverse[1] = multiplicative_inverse(base_scales[0], base_scales[1]);
    }
    uint64_t get_index_for_sample(uint64_t n) override {
        if (px != pixel_for_offset[0] || py != pixel_for_offset[1]) {
            offset_for_current_pixel = 0;
            if (sample_stride > 1) {
                int64_t pm[2] = {mod_i64(px, K_MAX_RESOLUTION), mod_i64(py, K_MAX_RESOLUTION)};
                for (int i = 0; i < 2; ++i) {
                    uint64_t dimoffset = inverse_radical_inverse(i == 0 ? 2 : 3, (uint64_t)pm[i], (uint64_t)base_exponents[i]);
                    offset_for_current_pixel += dimoffset * (sample_stride / (uint64_t)base_scales[i]) * (uint64_t)mult_inverse[i];
                }
                offset_for_current_pixel %= sample_stride;
            }
            pixel_for_offset[0] = px; pixel_for_offset[1] = py;
        }
        return offset_for_current_pixel + n * sample_stride;
    }
    size_t dim_limit() const override { return 1000; }  // PRIME_TABLE_SIZE, lowdiscrepancy.rs:9
    Float sample_dimension(uint64_t index, size_t dim) const override {
        const HaltonTables& T = HaltonTables::get();
        if (dim == 0) return radical_inverse(0, index >> (uint64_t)base_exponents[0]);
        if (dim == 1) return radical_inverse(1, index / (uint64_t)base_scales[1]);
        return scrambled_radical_inverse(dim, index, &T.perms[T.prime_sums[dim]]);
    }
    std::unique_ptr<Sampler> clone(int64_t) const override { return std::unique_ptr<Sampler>(new HaltonSampler(*this)); }
};

// ---- (0,2)-sequence: lowdiscrepancy.rs:428-510, samplers/zerotwosequence.rs:32-105
inline void gray_code_sample1d(const uint32_t* C, uint32_t n, uint32_t scramble, Float* p) {
    uint32_t v = scramble;
    for (uint32_t i = 0; i < n; ++i) {
        p[i] = std::fmin((Float)v * 0x1.0p-32f, ONE_MINUS_EPSILON);
        v ^= C[__builtin_ctz(i + 1)];
    }
}
inline void gray_code_sample2d(const uint32_t* C0, const uint32_t* C1, uint32_t n, uint32_t s0, uint32_t s1, P2* p) {
    uint32_t v0 = s0, v1 = s1;
    for (uint32_t i = 0; i < n; ++i) {
        p[i].x = std::fmin((Float)v0 * 0x1.0p-32f, ONE_MINUS_EPSILON);
        p[i].y = std::fmin((Float)v1 * 0x1.0p-32f, ONE_MINUS_EPSILON);
        v0 ^= C0[__builtin_ctz(i + 1)];
        v1 ^= C1[__builtin_ctz(i + 1)];
    }
}
struct ZeroTwoMatrices {
    uint32_t vdc[32], sobol1[32];
    ZeroTwoMatrices() {
        // CVAN_DER_CORPUT (lowdiscrepancy.rs:194-201) = CSOBOL[0] (:203-209): identity, MSB first.
        for (int i = 0; i < 32; ++i) vdc[i] = 0x80000000u >> i;
        // CSOBOL[1] (:210-217): v[0] = 2^31, v[i] = v[i-1] ^ (v[i-1] >> 1) (checked against the table in tests)
        sobol1[0] = 0x80000000u;
        for (int i = 1; i < 32; ++i) sobol1[i] = sobol1[i - 1] ^ (sobol1[i - 1] >> 1);
    }
    static const ZeroTwoMatrices& get() { static ZeroTwoMatrices m; return m; }
};
inline void vander_corput(size_t nspps, size_t npix, Float* samples, RNG& rng) {
    uint32_t scramble = rng.uniform_int32();
    gray_code_sample1d(ZeroTwoMatrices::get().vdc, (uint32_t)(nspps * npix), scramble, samples);
    for (size_t i = 0; i < npix; ++i) shuffle(samples + i * nspps, nspps, 1, rng);
    shuffle(samples, npix, nspps, rng);
}
inline void sobol_2d(size_t nspps, size_t npix, P2* samples, RNG& rng) {
    uint32_t s0 = rng.uniform_int32(), s1 = rng.uniform_int32();
    const ZeroTwoMatrices& M = ZeroTwoMatrices::get();
    gray_code_sample2d(M.vdc, M.sobol1, (uint32_t)(nspps * npix), s0, s1, samples);
    for (size_t i = 0; i < npix; ++i) shuffle(samples + i * nspps, nspps, 1, rng);
    shuffle(samples, npix, nspps, rng);
}
inline int64_t round_up_pow2_64(int64_t v) { v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; v |= v >> 32; return v + 1; }

struct ZeroTwoSequenceSampler : Sampler {
    std::vector<std::vector<Float>> samples_1d;
    std::vector<std::vector<P2>> samples_2d;
    size_t current_1d_dimension = 0, current_2d_dimension = 0;
    std::vector<int> samples_2d_array_sizes;        // request_2d_array_default!, sampler.rs:116-124
    std::vector<std::vector<P2>> sample_array_2d;
    size_t array_2d_offset = 0;
    RNG rng;
    int round_count(int n) const override { return (int)round_up_pow2_64(n); }  // zerotwosequence.rs:86-88
    void request_2d_array(int n) override { samples_2d_array_sizes.push_back(n); sample_array_2d.emplace_back((size_t)n * samples_per_pixel); }
    bool get_2d_array(int n, std::vector<P2>* out) override {  // get_2d_array_default!, sampler.rs:149-166
        if (array_2d_offset == sample_array_2d.size()) return false;
        if (samples_2d_array_sizes[array_2d_offset] != n) throw std::runtime_error("oracle: get_2d_array size mismatch (the reference asserts)");
        const P2* a = sample_array_2d[array_2d_offset].data() + current_pixel_sample_index * (uint64_t)n;
        out->assign(a, a + n);
        array_2d_offset += 1;
        return true;
    }
    ZeroTwoSequenceSampler(uint64_t spp, size_t ndims) {
        samples_per_pixel = (uint64_t)round_up_pow2_64((int64_t)spp);
        for (size_t i = 0; i < ndims; ++i) { samples_1d.emplace_back(samples_per_pixel, 0.0f); samples_2d.emplace_back(samples_per_pixel); }
    }
    void start_pixel(int x, int y) override {
        for (auto& s : samples_1d) vander_corput(1, samples_per_pixel, s.data(), rng);
        for (auto& s : samples_2d) sobol_2d(1, samples_per_pixel, s.data(), rng);
        for (size_t i = 0; i < sample_array_2d.size(); ++i)  // zerotwosequence.rs:67-71 (no 1D arrays are requested on this path)
            sobol_2d((size_t)samples_2d_array_sizes[i], samples_per_pixel, sample_array_2d[i].data(), rng);
        px = x; py = y; current_pixel_sample_index = 0;
        array_2d_offset = 0;
    }
    Float get_1d() override {
        if (current_1d_dimension < samples_1d.size()) return samples_1d[current_1d_dimension++][current_pixel_sample_index];
        return rng.uniform_float();
    }
    P2 get_2d() override {
        if (current_2d_dimension < samples_2d.size()) return samples_2d[current_2d_dimension++][current_pixel_sample_index];
        Float y = rng.uniform_float();  // sampler.rs:247-249: y drawn first
        Float x = rng.uniform_float();
        return P2(x, y);
    }
    bool start_next_sample() override {
        current_1d_dimension = current_2d_dimension = 0;
        array_2d_offset = 0;
        current_pixel_sample_index += 1;
        return current_pixel_sample_index < samples_per_pixel;
    }
    bool set_sample_number(uint64_t n) override {
        current_1d_dimension = current_2d_dimension = 0;
        array_2d_offset = 0;
        current_pixel_sample_index = n;
        return current_pixel_sample_index < samples_per_pixel;
    }
    std::unique_ptr<Sampler> clone(int64_t seed) const override {
        ZeroTwoSequenceSampler* s = new ZeroTwoSequenceSampler(*this);
        s->rng.set_sequence((uint64_t)seed);
        return std::unique_ptr<Sampler>(s);
    }
};

inline std::unique_ptr<Sampler> make_sampler(const pbrt_b200_sampler& sd, const SamplerTables& t) {
    switch (sd.kind) {
        case PBRT_B200_SAMPLER_SOBOL: return std::unique_ptr<Sampler>(new SobolSampler(sd.samples_per_pixel, sd.sample_bounds, t));
        case PBRT_B200_SAMPLER_HALTON: return std::unique_ptr<Sampler>(new HaltonSampler(sd.samples_per_pixel, sd.sample_bounds));
        default: return std::unique_ptr<Sampler>(new ZeroTwoSequenceSampler(sd.samples_per_pixel, sd.n_sampled_dimensions));
    }
}

}  // namespace orc
