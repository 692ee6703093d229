// projectultra_b200/csrc/ofdm_diff_demap.cuh — soft demapping of one differential carrier, shared by the warp-FFT
// kernels (ofdm_diff.cu, ofdm_diff512.cu).
#pragma once
#include "ofdm_dev.cuh"

namespace pu {

// ---- soft demapping of one differential carrier -------------------------------------------------------------------
// Exact path: soft_demap.hpp:173-237 with the host libm restatements of ref_math.cuh.
static __device__ __noinline__ void demap_exact(int mod, float2 sym, float2 prev, bool first, float nv, float (&l)[3]) {
    l[0] = l[1] = l[2] = 0.0f;
    const float2 df = cmul(sym, cconj(prev));
    const float sp = __fmul_rn(cabs_ref(sym), first ? 1.0f : cabs_ref(prev));   // |(1,0)| == 1 exactly
    if (sp < 1e-6f) return;                                                     // weak-signal gate (:178,199,224)
    const float phase = refmath::atan2f_ref(df.y, df.x);
    if (mod == PU_MOD_DBPSK) {              // :173-187
        l[0] = clip_llr(__fdiv_rn(__fmul_rn(__fmul_rn(2.0f, sp), refmath::cosf_ref(phase)), nv));
    } else if (mod == PU_MOD_DQPSK) {       // :192-213
        const float scale = __fdiv_rn(__fmul_rn(2.0f, sp), nv);
        const float pi = 3.14159265358979f;
        l[0] = clip_llr(__fmul_rn(scale, refmath::sinf_ref(__fadd_rn(phase, pi / 4))));
        l[1] = clip_llr(__fmul_rn(scale, refmath::cosf_ref(__fmul_rn(2.0f, phase))));
    } else {                                // D8PSK :217-237
        const float conf = __fdiv_rn(sp, nv);
        l[0] = clip_llr(__fmul_rn(conf, refmath::sinf_ref(phase)));
        l[1] = clip_llr(__fmul_rn(conf, refmath::sinf_ref(__fmul_rn(2.0f, phase))));
        l[2] = clip_llr(__fmul_rn(conf, refmath::sinf_ref(__fmul_rn(4.0f, phase))));
    }
}

// Saturation filter.  With d = sym * conj(prev), |d| = |sym||prev| = sp, so the reference's LLRs are, in exact
// arithmetic,  DBPSK 2 dx / nv;  DQPSK sqrt2 (dx+dy) / nv  and  2 (dx^2-dy^2) / (nv sp);  D8PSK dy / nv,
// 2 dx dy / (nv sp), 4 dx dy (dx^2-dy^2) / (nv sp^3): no atan2/sin/cos needed to know them approximately.  The
// reference evaluates scale * trig(k * atan2f(dy, dx) [+ pi/4]) in fp32; its absolute error is below
// scale * (k * 5e-7 + 5e-7) (<= 1 ulp atan2f at |phase| <= pi, one rounding of the angle, <= 1 ulp sinf/cosf, k <= 4),
// and the fp32 evaluation below is within 1e-6 relative.  So when every LLR of the carrier satisfies
// |approx| >= 10.01 + 1e-5 * scale the reference's clipLLR returns exactly +-10 with the sign of the approximation,
// and (sp_approx > 2e-6) rules out the weak-signal gate.  >= 99 % of carriers end here; the others
// take the exact path.  Returns false when the exact path is required.
__device__ __forceinline__ bool demap_saturated(int mod, float2 df, float nv, float (&l)[3]) {
    const float dx = df.x, dy = df.y;
    const float r2 = fmaf(dx, dx, dy * dy);
    const float sp = sqrtf(r2);                       // approximate |sym||prev|
    const float inv_nv = __frcp_rn(nv);
    const float inv_sp = __frcp_rn(sp);
    float a0, a1 = 1e30f, a2 = 1e30f, scale;
    if (mod == PU_MOD_DBPSK) {
        scale = 2.0f * sp * inv_nv;
        a0 = 2.0f * dx * inv_nv;
    } else if (mod == PU_MOD_DQPSK) {
        scale = 2.0f * sp * inv_nv;
        a0 = 1.41421356f * (dx + dy) * inv_nv;
        a1 = 2.0f * (dx - dy) * (dx + dy) * inv_nv * inv_sp;
    } else {
        scale = sp * inv_nv;
        const float sc = dx * dy * inv_nv * inv_sp;                  // conf * sin(phase) cos(phase)
        a0 = dy * inv_nv;
        a1 = 2.0f * sc;
        a2 = 4.0f * sc * (dx - dy) * (dx + dy) * inv_sp * inv_sp;
    }
    const float thr = fmaf(scale, 1e-5f, 10.01f);
    const bool ok = (sp > 2e-6f) && (fabsf(a0) >= thr) && (fabsf(a1) >= thr) && (fabsf(a2) >= thr) && (scale < 1e30f);
    l[0] = copysignf(10.0f, a0);
    l[1] = copysignf(10.0f, a1);
    l[2] = copysignf(10.0f, a2);
    return ok;
}

__device__ __forceinline__ void store_llrs(float* __restrict__ out, int base, int bps, const float (&l)[3], int llr_limit,
                                           const int* __restrict__ perm, int perm_len) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        if (b < bps) {
            const int pos = base + b;
            if (pos < llr_limit) {
                const int dst = (perm && pos < perm_len) ? perm[pos] : pos;
                out[dst] = l[b];
            }
        }
    }
}

// Same filter with the per-carrier 1/nv hoisted by the caller and MUFU.RSQ instead of IEEE sqrt + reciprocal: the extra
// relative error (<= 2^-22 of the rsqrt, a few roundings) stays far inside the 1e-5 * scale + 0.01 margin derived above.
// The weak-signal gate is tested on r2 = sp^2 (> 4e-12) so that a flushed rsqrt input cannot pass it.
// rel_margin: callers that feed an approximation of sym * conj(prev) (ofdm_diff512.cu forms it from the raw FFT bins,
// bin * conj(previous bin) / |h|^2: <= 10 ulp away from the reference's fp32 value relative to |d|, which moves the
// LLR laws by <= 8 * 1.2e-6 * scale for the k = 4 term) widen the relative part of the margin accordingly.
__device__ __forceinline__ bool demap_saturated_fast(int mod, float2 df, float inv_nv, float (&l)[3], float rel_margin = 1e-5f) {
    const float dx = df.x, dy = df.y;
    const float r2 = fmaf(dx, dx, dy * dy);
    float inv_sp;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv_sp) : "f"(r2));
    const float sp = r2 * inv_sp;
    float a0, a1 = 1e30f, a2 = 1e30f, scale;
    if (mod == PU_MOD_DBPSK) {
        scale = 2.0f * sp * inv_nv;
        a0 = 2.0f * dx * inv_nv;
    } else if (mod == PU_MOD_DQPSK) {
        scale = 2.0f * sp * inv_nv;
        a0 = 1.41421356f * (dx + dy) * inv_nv;
        a1 = 2.0f * (dx - dy) * (dx + dy) * inv_nv * inv_sp;
    } else {
        scale = sp * inv_nv;
        const float sc = dx * dy * inv_nv * inv_sp;
        a0 = dy * inv_nv;
        a1 = 2.0f * sc;
        a2 = 4.0f * sc * (dx - dy) * (dx + dy) * inv_sp * inv_sp;
    }
    const float thr = fmaf(scale, rel_margin, 10.01f);
    const bool ok = (r2 > 4e-12f) && (r2 < 1e30f) && (fabsf(a0) >= thr) && (fabsf(a1) >= thr) && (fabsf(a2) >= thr) && (scale < 1e30f);
    l[0] = copysignf(10.0f, a0);
    l[1] = copysignf(10.0f, a1);
    l[2] = copysignf(10.0f, a2);
    return ok;
}

}  // namespace pu
