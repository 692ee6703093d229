// projectultra_b200/csrc/pu_rng.cuh — counter-based random numbers of the channel simulator.
//
// The reference's WattersonChannel draws from std::mt19937 + std::normal_distribution<float>
// (src/sim/hf_channel.hpp:67-70,133-150,261-263), a sequential, implementation-defined stream.  The batched
// simulator replaces it by a counter-based generator so that every sample of every frame is a pure function of
// (seed, sample index) and any frame can be regenerated anywhere (BASELINE.json north_star).  SPECIFICATION
// (restated independently in oracle/pu_oracle_channel.c, the CPU twin):
//   * Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11), key = 64-bit frame seed
//     (lo, hi), counter = (index, 0, stream, 0); stream 1 = fading, stream 2 = additive noise;
//   * uniform: u = ((w >> 9) + 0.5) * 2^-23, exactly representable, in (0, 1);
//   * Gaussian pair by Box-Muller: r = sqrt(-2 ln u1), (r cos 2 pi u2, r sin 2 pi u2), with ln, sin, cos evaluated
//     by the fixed polynomials below using only IEEE add/mul/fma/sqrt, so host and device agree bit for bit;
//   * fading innovations of sample n: Philox(index = n >> 1, stream 1) -> words w0..w3 = (tap1.re, tap1.im, tap2.re, tap2.im); sample n
//     takes the 16-bit half (n & 1) of each word: z = (u16 - 32767.5) * sqrt(12) / 65536, a UNIFORM variate of variance 1 - 2^-32.
//     The reference draws Gaussians here (hf_channel.hpp:261-263), but a fading tap is the one-pole sum of ~1/(2a) >= 385 (flutter)
//     ... 38 000 (good) innovations, so the tap process is Gaussian whatever the innovation's shape (excess kurtosis <= 1.2 * 2a =
//     3e-3); Box-Muller on 4 variates per sample was 60 % of the channel kernel's instructions (v03: 6.6 ms per 53 248 frames);
//   * noise normal of sample n: Philox(index = n >> 2, stream 2), pair (n & 3) >> 1, cos branch if n even else sin.
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace pu {
namespace rng {

#if defined(__CUDA_ARCH__)
#define PU_RNG __device__ __forceinline__
PU_RNG float r_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
PU_RNG float r_mul(float a, float b) { return __fmul_rn(a, b); }
PU_RNG float r_add(float a, float b) { return __fadd_rn(a, b); }
PU_RNG float r_sub(float a, float b) { return __fsub_rn(a, b); }
PU_RNG float r_sqrt(float a) { return __fsqrt_rn(a); }
PU_RNG uint32_t r_bits(float x) { return __float_as_uint(x); }
PU_RNG float r_float(uint32_t u) { return __uint_as_float(u); }
PU_RNG uint32_t r_mulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
#else
#define PU_RNG __host__ __device__ inline
PU_RNG float r_fma(float a, float b, float c) { return std::fmaf(a, b, c); }
PU_RNG float r_mul(float a, float b) { volatile float r = a * b; return r; }
PU_RNG float r_add(float a, float b) { volatile float r = a + b; return r; }
PU_RNG float r_sub(float a, float b) { volatile float r = a - b; return r; }
PU_RNG float r_sqrt(float a) { return std::sqrt(a); }
PU_RNG uint32_t r_bits(float x) { uint32_t u; std::memcpy(&u, &x, 4); return u; }
PU_RNG float r_float(uint32_t u) { float x; std::memcpy(&x, &u, 4); return x; }
PU_RNG uint32_t r_mulhi(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }
#endif

constexpr uint32_t kStreamFading = 1, kStreamNoise = 2;

struct U4 { uint32_t x, y, z, w; };

PU_RNG U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = r_mulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = r_mulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0;
        k1 += W1;
    }
    return U4{c0, c1, c2, c3};
}

PU_RNG float uniform23(uint32_t w) { return r_mul(r_add(static_cast<float>(w >> 9), 0.5f), 1.1920928955078125e-07f); }

// natural log for u in (0, 1]: exponent/mantissa split at sqrt(2), degree-9 polynomial (Cephes logf coefficients)
PU_RNG float log_poly(float u) {
    const uint32_t b = r_bits(u);
    int e = static_cast<int>(b >> 23) - 127;
    float m = r_float((b & 0x007fffffu) | 0x3f800000u);
    if (m > 1.41421356f) { m = r_mul(m, 0.5f); e += 1; }
    const float f = r_sub(m, 1.0f);
    const float z = r_mul(f, f);
    float p = 7.0376836292e-2f;
    p = r_fma(p, f, -1.1514610310e-1f);
    p = r_fma(p, f, 1.1676998740e-1f);
    p = r_fma(p, f, -1.2420140846e-1f);
    p = r_fma(p, f, 1.4249322787e-1f);
    p = r_fma(p, f, -1.6668057665e-1f);
    p = r_fma(p, f, 2.0000714765e-1f);
    p = r_fma(p, f, -2.4999993993e-1f);
    p = r_fma(p, f, 3.3333331174e-1f);
    float y = r_mul(r_mul(f, z), p);
    y = r_fma(-0.5f, z, y);
    const float fe = static_cast<float>(e);
    float r = r_add(f, y);
    r = r_fma(fe, -2.12194440e-4f, r);
    r = r_fma(fe, 0.693359375f, r);
    return r;
}

// (sin, cos)(2 pi u) for u in (0, 1): quadrant split, then Cephes sinf/cosf polynomials on [-pi/4, pi/4]
PU_RNG void sincos_2pi(float u, float* s_out, float* c_out) {
    const float t = r_mul(u, 4.0f);
    const int k = static_cast<int>(r_add(t, 0.5f));
    const float x = r_mul(r_sub(t, static_cast<float>(k)), 1.57079632679489662f);
    const float x2 = r_mul(x, x);
    float ps = -1.9515295891e-4f;
    ps = r_fma(ps, x2, 8.3321608736e-3f);
    ps = r_fma(ps, x2, -1.6666654611e-1f);
    const float s = r_fma(r_mul(x, x2), ps, x);
    float pc = 2.443315711809948e-5f;
    pc = r_fma(pc, x2, -1.388731625493765e-3f);
    pc = r_fma(pc, x2, 4.166664568298827e-2f);
    const float c = r_fma(r_mul(x2, x2), pc, r_fma(-0.5f, x2, 1.0f));
    switch (k & 3) {
        case 0: *s_out = s; *c_out = c; break;
        case 1: *s_out = c; *c_out = -s; break;
        case 2: *s_out = -s; *c_out = -c; break;
        default: *s_out = -c; *c_out = s; break;
    }
}

PU_RNG void box_muller(uint32_t w0, uint32_t w1, float* z_cos, float* z_sin) {
    const float r = r_sqrt(r_mul(-2.0f, log_poly(uniform23(w0))));
    float s, c;
    sincos_2pi(uniform23(w1), &s, &c);
    *z_cos = r_mul(r, c);
    *z_sin = r_mul(r, s);
}

// additive-noise normal of sample n
PU_RNG float noise_normal(uint32_t k0, uint32_t k1, uint32_t n) {
    const U4 w = philox4x32_10(n >> 2, 0u, kStreamNoise, 0u, k0, k1);
    const bool second = (n & 2) != 0;
    float zc, zs;
    box_muller(second ? w.z : w.x, second ? w.w : w.y, &zc, &zs);
    return (n & 1) ? zs : zc;
}

// unit-variance uniform innovation from 16 random bits
PU_RNG float innovation16(uint32_t u16) { return r_mul(r_sub(static_cast<float>(u16), 32767.5f), 5.2857806906e-05f); }

// the four fading innovations of samples 2 i and 2 i + 1: z[h][c], h = n & 1, c = (tap1.re, tap1.im, tap2.re, tap2.im)
PU_RNG void fading_innovations2(uint32_t k0, uint32_t k1, uint32_t i, float (&z)[2][4]) {
    const U4 w = philox4x32_10(i, 0u, kStreamFading, 0u, k0, k1);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        z[0][c] = innovation16(ws[c] & 0xffffu);
        z[1][c] = innovation16(ws[c] >> 16);
    }
}

}  // namespace rng
}  // namespace pu
