// projectultra_b200/csrc/ref_math.cuh — device restatements of the host libm routines the reference's receive
// path calls (through std::sin/std::cos/std::atan2/std::abs/std::exp on float), so that LLRs can match the
// reference to the last bit instead of "within a few ulp".
//
// The reference links glibc 2.39 (Ubuntu 24.04, the image both the build container and the GPU box run).  These
// are third-party algorithms that are NOT part of /root/reference; they are restated from their published
// descriptions and pinned by tests that compare them with the host libm on millions of inputs
// (tests/test_ref_math.py runs the same expressions on the CPU; tests/test_ofdm_gpu.py on the GPU):
//   atan2f / atanf : fdlibm (Sun Microsystems) single-precision e_atan2f.c / s_atanf.c as shipped by glibc 2.39
//                    -- breakpoints 7/16, 11/16, 19/16, 39/16, an 11-term odd polynomial, all arithmetic fp32.
//   sinf / cosf    : the Arm Optimized Routines single-precision algorithm adopted by glibc 2.28+
//                    (sysdeps/ieee754/flt-32/s_sincosf.h): reduce by pi/2 with a 2^24-scaled inverse, then a
//                    degree-7/8 polynomial evaluated in double and rounded once.  x86-64 hosts with FMA run the
//                    *_fma ifunc variants, whose contractions are reproduced here with explicit fma().
//   hypotf         : sqrt((double)x*x + (double)y*y) rounded to float (glibc >= 2.35).
// Valid for |x| < 120 (sin/cos) -- phases on this path are within a few multiples of pi; larger arguments fall
// back to CUDA's sinf/cosf.
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace pu {
namespace refmath {

// The same source runs on the device (explicit round-to-nearest intrinsics, never contracted) and on the host
// (plain operators; host objects are built for baseline x86-64, which has no FMA to contract into), so the CPU
// test can compare these restatements with the host libm directly.
#if defined(__CUDA_ARCH__)
#define PU_RM __device__ __forceinline__
PU_RM float rm_fadd(float a, float b) { return __fadd_rn(a, b); }
PU_RM float rm_fsub(float a, float b) { return __fsub_rn(a, b); }
PU_RM float rm_fmul(float a, float b) { return __fmul_rn(a, b); }
PU_RM float rm_fdiv(float a, float b) { return __fdiv_rn(a, b); }
PU_RM double rm_dadd(double a, double b) { return __dadd_rn(a, b); }
PU_RM double rm_dmul(double a, double b) { return __dmul_rn(a, b); }
PU_RM double rm_dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
PU_RM double rm_dsqrt(double a) { return __dsqrt_rn(a); }
PU_RM int32_t rm_f2i(float x) { return __float_as_int(x); }
#else
#define PU_RM __host__ __device__ inline
PU_RM float rm_fadd(float a, float b) { volatile float r = a + b; return r; }
PU_RM float rm_fsub(float a, float b) { volatile float r = a - b; return r; }
PU_RM float rm_fmul(float a, float b) { volatile float r = a * b; return r; }
PU_RM float rm_fdiv(float a, float b) { volatile float r = a / b; return r; }
PU_RM double rm_dadd(double a, double b) { volatile double r = a + b; return r; }
PU_RM double rm_dmul(double a, double b) { volatile double r = a * b; return r; }
PU_RM double rm_dfma(double a, double b, double c) { return std::fma(a, b, c); }
PU_RM double rm_dsqrt(double a) { return std::sqrt(a); }
PU_RM int32_t rm_f2i(float x) { int32_t i; std::memcpy(&i, &x, 4); return i; }
#endif

PU_RM float atanf_ref(float x) {
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float a0 = 3.3333334327e-01f, a1 = -2.0000000298e-01f, a2 = 1.4285714924e-01f, a3 = -1.1111110449e-01f,
                a4 = 9.0908870101e-02f, a5 = -7.6918758452e-02f, a6 = 6.6610731184e-02f, a7 = -5.8335702866e-02f,
                a8 = 4.9768779427e-02f, a9 = -3.6531571299e-02f, a10 = 1.6285819933e-02f;
    const int32_t hx = rm_f2i(x);
    const int32_t ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c800000) {                       // |x| >= 2^26
        if (ix > 0x7f800000) return rm_fadd(x, x);
        const float r = rm_fadd(atanhi[3], atanlo[3]);
        return hx > 0 ? r : -r;
    }
    if (ix < 0x3ee00000) {                        // |x| < 0.4375
        if (ix < 0x31000000) return x;            // |x| < 2^-29
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {                    // |x| < 1.1875
            if (ix < 0x3f300000) { id = 0; x = rm_fdiv(rm_fsub(rm_fmul(2.0f, x), 1.0f), rm_fadd(2.0f, x)); }
            else { id = 1; x = rm_fdiv(rm_fsub(x, 1.0f), rm_fadd(x, 1.0f)); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = rm_fdiv(rm_fsub(x, 1.5f), rm_fadd(1.0f, rm_fmul(1.5f, x))); }
            else { id = 3; x = rm_fdiv(-1.0f, x); }
        }
    }
    const float z = rm_fmul(x, x), w = rm_fmul(z, z);
#define PU_H(c, acc) rm_fadd((c), rm_fmul(w, (acc)))
    const float s1 = rm_fmul(z, PU_H(a0, PU_H(a2, PU_H(a4, PU_H(a6, PU_H(a8, a10))))));
    const float s2 = rm_fmul(w, PU_H(a1, PU_H(a3, PU_H(a5, PU_H(a7, a9)))));
#undef PU_H
    const float xs = rm_fmul(x, rm_fadd(s1, s2));
    if (id < 0) return rm_fsub(x, xs);
    const float r = rm_fsub(atanhi[id], rm_fsub(rm_fsub(xs, atanlo[id]), x));
    return hx < 0 ? -r : r;
}

PU_RM float atan2f_ref(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
                pi_lo = -8.7422776573e-08f;
    const int32_t hx = rm_f2i(x), hy = rm_f2i(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return rm_fadd(x, y);
    if (hx == 0x3f800000) return atanf_ref(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        switch (m) {
            case 0: case 1: return y;
            case 2: return rm_fadd(pi, tiny);
            default: return rm_fsub(-pi, tiny);
        }
    }
    if (ix == 0) return hy < 0 ? rm_fsub(-pi_o_2, tiny) : rm_fadd(pi_o_2, tiny);
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
                case 0: return rm_fadd(pi_o_4, tiny);
                case 1: return rm_fsub(-pi_o_4, tiny);
                case 2: return rm_fadd(rm_fmul(3.0f, pi_o_4), tiny);
                default: return rm_fsub(rm_fmul(-3.0f, pi_o_4), tiny);
            }
        }
        switch (m) {
            case 0: return 0.0f;
            case 1: return -0.0f;
            case 2: return rm_fadd(pi, tiny);
            default: return rm_fsub(-pi, tiny);
        }
    }
    if (iy == 0x7f800000) return hy < 0 ? rm_fsub(-pi_o_2, tiny) : rm_fadd(pi_o_2, tiny);
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 24) z = rm_fadd(pi_o_2, rm_fmul(0.5f, pi_lo));
    else if (hx < 0 && k < -26) z = 0.0f;
    else z = atanf_ref(fabsf(rm_fdiv(y, x)));
    switch (m) {
        case 0: return z;
        case 1: return -z;
        case 2: return rm_fsub(pi, rm_fsub(z, pi_lo));
        default: return rm_fsub(rm_fsub(z, pi_lo), pi);
    }
}

// ---- sinf / cosf ----
struct SinCosPoly { double c0, c1, c2, c3, c4, s1, s2, s3; };

PU_RM uint32_t abstop12(float x) { return (static_cast<uint32_t>(rm_f2i(x)) >> 20) & 0x7ff; }

// n even: sin polynomial, n odd: cos polynomial; `neg` selects the negated-cosine coefficient set
PU_RM float sincos_poly(double x, double x2, int n, bool neg) {
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = rm_dmul(x, x2);
        const double s1 = rm_dfma(x2, S3, S2);
        const double x7 = rm_dmul(x3, x2);
        const double s = rm_dfma(x3, S1, x);
        return static_cast<float>(rm_dfma(x7, s1, s));
    }
    const double sg = neg ? -1.0 : 1.0;
    const double C0 = sg * 0x1p0, C1 = sg * -0x1.ffffffd0c621cp-2, C2 = sg * 0x1.55553e1068f19p-5,
                 C3 = sg * -0x1.6c087e89a359dp-10, C4 = sg * 0x1.99343027bf8c3p-16;
    const double x4 = rm_dmul(x2, x2);
    const double c2 = rm_dfma(x2, C4, C3);
    const double c1 = rm_dfma(x2, C1, C0);
    const double x6 = rm_dmul(x4, x2);
    const double c = rm_dfma(x4, C2, c1);
    return static_cast<float>(rm_dfma(x6, c2, c));
}

PU_RM double reduce_fast(double x, int* np) {
    const double r = rm_dmul(x, 0x1.45F306DC9C883p+23);
    const int n = (static_cast<int32_t>(r) + 0x800000) >> 24;
    *np = n;
    return rm_dfma(-static_cast<double>(n), 0x1.921FB54442D18p0, x);
}

PU_RM float sinf_ref(float y) {
    const double x = y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        if (abstop12(y) < abstop12(0x1p-12f)) return y;
        return sincos_poly(x, rm_dmul(x, x), 0, false);
    }
    if (abstop12(y) < abstop12(120.0f)) {
        int n;
        const double xr = reduce_fast(x, &n);
        const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
        return sincos_poly(rm_dmul(xr, s), rm_dmul(xr, xr), n, (n & 2) != 0);
    }
    return ::sinf(y);
}

PU_RM float cosf_ref(float y) {
    const double x = y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
        return sincos_poly(x, rm_dmul(x, x), 1, false);
    }
    if (abstop12(y) < abstop12(120.0f)) {
        int n;
        const double xr = reduce_fast(x, &n);
        const int q = (n + 1) & 3;
        const double s = (q == 1 || q == 2) ? -1.0 : 1.0;
        return sincos_poly(rm_dmul(xr, s), rm_dmul(xr, xr), n ^ 1, ((n + 1) & 2) != 0);
    }
    return ::cosf(y);
}

PU_RM void sincosf_ref(float y, float* s, float* c) {
    *s = sinf_ref(y);
    *c = cosf_ref(y);
}

// std::abs(std::complex<float>) -> hypotf
PU_RM float hypotf_ref(float x, float y) {
    const double a = x, b = y;
    return static_cast<float>(rm_dsqrt(rm_dadd(rm_dmul(a, a), rm_dmul(b, b))));
}

}  // namespace refmath
}  // namespace pu
