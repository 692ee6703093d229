// projectultra_b200/csrc/fft_smem.cuh — the reference's iterative radix-2 FFT on a shared-memory buffer, CTA-wide
// (used by the acquisition kernel's analytic signal and by the transmitter's inverse FFT).
#pragma once
#include "ofdm_dev.cuh"

namespace pu {

// fft_impl (src/dsp/fft.cpp:89-121) on buf[NFFT] in shared memory: bit-reversal permutation, log2 N butterfly stages,
// 1/N scaling for the inverse.  Entered and left with the CTA synchronised.
template <int NFFT>
__device__ void fft_smem(float2* buf, const float2* __restrict__ tw, bool inverse) {
    constexpr int LOG2N = (NFFT == 512) ? 9 : 10;
    const int tid = threadIdx.x, T = blockDim.x;
    for (int i = tid; i < NFFT; i += T) {
        const int j = static_cast<int>(__brev(static_cast<unsigned>(i)) >> (32 - LOG2N));
        if (i < j) { const float2 a = buf[i]; buf[i] = buf[j]; buf[j] = a; }
    }
    __syncthreads();
    for (int s = 1; s <= LOG2N; ++s) {
        const int half = 1 << (s - 1);
        for (int b = tid; b < NFFT / 2; b += T) {
            const int k = b & (half - 1);
            const int i0 = ((b >> (s - 1)) << s) | k;
            float2 w = __ldg(&tw[k << (LOG2N - s)]);
            if (inverse) w.y = -w.y;                       // std::conj(w)
            const float2 t = cmul(w, buf[i0 + half]);      // Complex t = w * data[i + k + half]
            const float2 a = buf[i0];
            buf[i0 + half] = csub(a, t);
            buf[i0] = cadd(a, t);
        }
        __syncthreads();
    }
    if (inverse) {
        const float scale = __fdiv_rn(1.0f, static_cast<float>(NFFT));
        for (int i = tid; i < NFFT; i += T) buf[i] = make_float2(__fmul_rn(buf[i].x, scale), __fmul_rn(buf[i].y, scale));
        __syncthreads();
    }
}

}  // namespace pu
