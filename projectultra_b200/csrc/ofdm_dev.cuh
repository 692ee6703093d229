// projectultra_b200/csrc/ofdm_dev.cuh — device-side types and helpers shared by the OFDM receive kernels
// (ofdm_demod.cu: general presynced path; ofdm_diff.cu: warp-FFT path for differential no-pilot modes).
#pragma once
#include <cuda_runtime.h>

#include "pu/pu_capi.h"
#include "ref_math.cuh"

namespace pu {

constexpr int kMaxCarr = 64;
constexpr int kDbgScalars = 10;

struct OfdmDev {
    int nfft, log2n, cp, sym_len, n_data, n_pilot, bps, mod;
    float ce_margin, sample_rate;
    const float2* twiddle;   // [nfft/2]
    const float2* nco;       // [max_symbols * sym_len] (cos, sin)
    int nco_len;
    const int* data_bin;     // [n_data]
    const int* pilot_bin;    // [n_pilot]
    const float2* zc;        // [n_data] known LTS symbol on data carrier i
    const float* pilot_sign; // [n_pilot]
    const int* interp_lo;    // [n_data]
    const int* interp_hi;
    const float* interp_alpha;
    const int* llr_perm;     // optional [perm_len]: output position of LLR index i (fused deinterleave), or NULL
    int perm_len;
};

// ---- std::complex<float> arithmetic as GCC lowers it (no FMA, naive formulas) ----
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                       __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
__device__ __forceinline__ float2 cscale(float s, float2 a) { return make_float2(__fmul_rn(a.x, s), __fmul_rn(a.y, s)); }
__device__ __forceinline__ float2 cdivs(float2 a, float s) { return make_float2(__fdiv_rn(a.x, s), __fdiv_rn(a.y, s)); }
__device__ __forceinline__ float cnorm(float2 a) { return __fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)); }
// complex / complex: libgcc __divsc3 evaluates the textbook formula in double when the hardware has doubles
__device__ __forceinline__ float2 cdiv(float2 a, float2 b) {
    const double aa = a.x, bb = a.y, cc = b.x, dd = b.y;
    const double den = __dadd_rn(__dmul_rn(cc, cc), __dmul_rn(dd, dd));
    const double x = __ddiv_rn(__dadd_rn(__dmul_rn(aa, cc), __dmul_rn(bb, dd)), den);
    const double y = __ddiv_rn(__dsub_rn(__dmul_rn(bb, cc), __dmul_rn(aa, dd)), den);
    return make_float2(static_cast<float>(x), static_cast<float>(y));
}
// std::abs(complex<float>) = hypotf: glibc evaluates sqrt(x*x + y*y) in double
__device__ __forceinline__ float cabs_ref(float2 a) {
    const double x = a.x, y = a.y;
    return static_cast<float>(__dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y))));
}
__device__ __forceinline__ float clampf(float lo, float hi, float v) { return fmaxf(lo, fminf(hi, v)); }

// soft_demap::clipLLR, soft_demap.hpp:22-29
__device__ __forceinline__ float clip_llr(float llr) {
    float c = clampf(-10.0f, 10.0f, llr);
    if (fabsf(c) < 0.5f) c = (c >= 0.0f) ? 0.5f : -0.5f;
    return c;
}

}  // namespace pu
