// projectultra_b200/csrc/chirp_sync.cu — batched dual-chirp synchronisation (chirp half), one frame per CTA.
//
// Reference behaviour: sync::ChirpSync::detectDualChirp (src/sync/chirp_sync.hpp:349-506) with detectChirpTemplate (:560-629)
// and computeComplexTemplateCorrelation (:639-662), configured as OFDMChirpWaveform does (src/waveform/ofdm_chirp_waveform.cpp:
// 39-49), followed by the receive glue of OFDMChirpWaveform::detectSync / process (:129-199): the training symbols start at
// down_chirp_start + chirp + gap, the CFO rotator starts from the phase accumulated since sample 0.
// Every correlation is three ordered 24 000-tap sums exactly as the reference accumulates them (one search position per
// thread); what runs in parallel are the search positions, which the reference evaluates independently.  The first-maximum
// rule of the reference's ascending scans is kept by the reductions (larger value wins, equal value -> smaller position).
#include <cfloat>
#include <cmath>
#include <vector>

#include "pu_internal.h"

namespace pu {

constexpr int kChirpThreads = 256;
struct ChirpShared {
    float red_c[kChirpThreads / 32];
    int red_p[kChirpThreads / 32];
    float best_c;
    int best_p;
    float c0, c2;
};

// computeComplexTemplateCorrelation(samples, off) over a window of Lw samples (:639-662)
__device__ float chirp_corr(const float* __restrict__ x, int Lw, int off, const float* __restrict__ ts, const float* __restrict__ tc, int n, float te) {
    if (off + n > Lw) return 0.0f;
    const float* w = x + off;
    float ci = 0.0f, cq = 0.0f, se = 0.0f;
    for (int i = 0; i < n; ++i) {
        const float s = w[i];
        ci = __fadd_rn(ci, __fmul_rn(s, __ldg(&tc[i])));
        cq = __fadd_rn(cq, __fmul_rn(s, __ldg(&ts[i])));
        se = __fadd_rn(se, __fmul_rn(s, s));
    }
    const float denom = __fsqrt_rn(__fmul_rn(se, te));
    if (denom < 1e-10f) return 0.0f;
    return __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(ci, ci), __fmul_rn(cq, cq))), denom);
}

// Coarse search of detectChirpTemplate (:575-582) for the 32 consecutive coarse positions p0, p0 + 48, ... of one warp, staged through
// shared memory.  Lanes sit 48 samples apart, so reading x[p0 + 48 l + i] with all lanes at the same tap i would put 16 lanes on each
// of two banks (and, from global memory, 32 cache lines per instruction: what the first version of this kernel did).  The sums stay
// ordered per lane, but nothing requires the lanes to be at the SAME tap: lane l runs s_l = 15 l mod 32 taps behind, which moves its
// sample read to bank (l + t) mod 32 and its template reads to bank (t - 15 l) mod 32 -- all different across the warp.
constexpr int kChirpTile = 512;                                      // taps per staged tile
constexpr int kChirpXs = 48 * 31 + kChirpTile + 32;                  // samples per tile: all 32 windows, skew included
struct ChirpWarpBuf { float xs[kChirpXs]; float tc[kChirpTile + 32]; float ts[kChirpTile + 32]; };

__device__ float chirp_corr_coarse_warp(ChirpWarpBuf& W, const float* __restrict__ x, int Lw, int p0, const float* __restrict__ ts,
                                        const float* __restrict__ tc, int n, float te) {
    const int lane = threadIdx.x & 31;
    const int skew = (15 * lane) & 31;
    float ci = 0.0f, cq = 0.0f, se = 0.0f;
    for (int t0 = 0; t0 < n + 32; t0 += kChirpTile) {
        __syncwarp();
        const int xbase = p0 + t0 - 32;                              // tile = x[xbase .. xbase + kChirpXs), taps [t0 - 32, t0 + kChirpTile)
        for (int j = lane; j < kChirpXs; j += 32) {
            const int idx = xbase + j;
            W.xs[j] = (idx >= 0 && idx < Lw) ? x[idx] : 0.0f;
        }
        for (int j = lane; j < kChirpTile + 32; j += 32) {
            const int i = t0 - 32 + j;
            const bool in = i >= 0 && i < n;
            W.tc[j] = in ? __ldg(&tc[i]) : 0.0f;
            W.ts[j] = in ? __ldg(&ts[i]) : 0.0f;
        }
        __syncwarp();
        const int xo = 48 * lane + 32 - skew - t0, to = 32 - skew - t0;
        // this lane's steps inside the tile: tap i = t - skew must lie in [0, n)
        const int t_lo = max(t0, skew), t_hi = min(t0 + kChirpTile, n + skew);
        const float* px = W.xs + xo;
        const float* pc = W.tc + to;
        const float* ps = W.ts + to;
#pragma unroll 4
        for (int t = t_lo; t < t_hi; ++t) {
            const float sv = px[t];
            ci = __fadd_rn(ci, __fmul_rn(sv, pc[t]));
            cq = __fadd_rn(cq, __fmul_rn(sv, ps[t]));
            se = __fadd_rn(se, __fmul_rn(sv, sv));
        }
    }
    const float denom = __fsqrt_rn(__fmul_rn(se, te));
    if (denom < 1e-10f) return 0.0f;
    return __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(ci, ci), __fmul_rn(cq, cq))), denom);
}

// CTA-wide "first maximum": every thread holds its best (c, p) over an ascending subset (c > running best, strict); the result is
// the maximum value at its smallest position, seeded with (seed_c, seed_p).
__device__ void chirp_reduce(ChirpShared& S, float c, int p, float seed_c, int seed_p) {
    const int tid = threadIdx.x;
    for (int o = 16; o > 0; o >>= 1) {
        const float oc = __shfl_xor_sync(0xffffffffu, c, o);
        const int op = __shfl_xor_sync(0xffffffffu, p, o);
        if (oc > c || (oc == c && op >= 0 && (p < 0 || op < p))) { c = oc; p = op; }
    }
    if ((tid & 31) == 0) { S.red_c[tid >> 5] = c; S.red_p[tid >> 5] = p; }
    __syncthreads();
    if (tid == 0) {
        float bc = seed_c;
        int bp = seed_p;
        // candidates beat the seed only when strictly larger (the scan's `corr > best_corr`); among themselves ties go to the
        // smaller position, which the ascending scan would have met first
        float cc = -1.0f;
        int cp = -1;
        for (int w = 0; w < kChirpThreads / 32; ++w) {
            const float oc = S.red_c[w];
            const int op = S.red_p[w];
            if (op >= 0 && (oc > cc || (oc == cc && op < cp))) { cc = oc; cp = op; }
        }
        if (cp >= 0 && cc > bc) { bc = cc; bp = cp; }
        S.best_c = bc;
        S.best_p = bp;
    }
    __syncthreads();
}

// detectChirpTemplate (:560-629) on the window x[0 .. Lw): returns the position or -1, *corr_out = best correlation seen
__device__ int chirp_detect_template(ChirpShared& S, ChirpWarpBuf* WB, const float* __restrict__ x, int Lw, const float* ts, const float* tc, int n, float te,
                                     float threshold, float* corr_out) {
    const int tid = threadIdx.x;
    *corr_out = 0.0f;
    if (Lw < n) return -1;
    const int search_len = Lw - n;
    // coarse search, step 48
    float bc = 0.0f;
    int bp = -1;
    {
        const int n_pos = (search_len + 47) / 48;                    // positions 0, 48, ... < search_len
        const int warp = tid >> 5, lane = tid & 31;
        for (int g = warp; g * 32 < n_pos; g += kChirpThreads / 32) {   // 32 consecutive positions per warp and round, ascending
            const float c = chirp_corr_coarse_warp(WB[warp], x, Lw, g * 32 * 48, ts, tc, n, te);
            const int m = g * 32 + lane;
            if (m < n_pos && c > bc) { bc = c; bp = m * 48; }
        }
    }
    chirp_reduce(S, bp >= 0 ? bc : -1.0f, bp, 0.0f, -1);
    float best = S.best_c;
    int best_pos = S.best_p;
    __syncthreads();
    *corr_out = best;
    if (best_pos < 0 || best < __fmul_rn(threshold, 0.3f)) return -1;
    // fine search, step 1
    const int fine_start = max(0, best_pos - 48), fine_end = min(search_len, best_pos + 48);
    bc = -1.0f;
    bp = -1;
    for (int pos = fine_start + tid; pos <= fine_end; pos += kChirpThreads) {
        const float c = chirp_corr(x, Lw, pos, ts, tc, n, te);
        if (c > bc) { bc = c; bp = pos; }
    }
    chirp_reduce(S, bc, bp, best, best_pos);
    best = S.best_c;
    best_pos = S.best_p;
    __syncthreads();
    // parabolic interpolation
    if (best_pos > 0 && best_pos < search_len - 1) {
        if (tid == 0) S.c0 = chirp_corr(x, Lw, best_pos - 1, ts, tc, n, te);
        if (tid == 32) S.c2 = chirp_corr(x, Lw, best_pos + 1, ts, tc, n, te);
        __syncthreads();
        const float c0 = S.c0, c1 = best, c2 = S.c2;
        const float denom = __fmul_rn(2.0f, __fadd_rn(__fsub_rn(c0, __fmul_rn(2.0f, c1)), c2));
        if (fabsf(denom) > 1e-10f) {
            float delta = __fdiv_rn(__fsub_rn(c0, c2), denom);
            delta = fmaxf(-1.0f, fminf(1.0f, delta));
            best_pos = static_cast<int>(roundf(__fadd_rn(static_cast<float>(best_pos), delta)));
        }
        __syncthreads();
    }
    *corr_out = best;
    return best >= threshold ? best_pos : -1;
}

// out_info[b] = {success, up_chirp_start, down_chirp_start, start_sample (training start) or -1}; out_f[b] = {cfo_hz, up corr, down corr,
// initial rotator phase}.  frame_start / frame_nsym (optional) = the window handed to the presynced kernel (0 symbols when not found).
__global__ void __launch_bounds__(kChirpThreads) chirp_detect_kernel(ChirpDev c, const float* __restrict__ samples, size_t frame_stride, int L,
                                                                     float threshold, int sym_len, int4* __restrict__ out_info,
                                                                     float4* __restrict__ out_f, int* __restrict__ frame_start,
                                                                     int* __restrict__ frame_nsym, float* __restrict__ cfo_out,
                                                                     float* __restrict__ phase_out, int* __restrict__ n_llr,
                                                                     int llr_per_symbol, int llr_stride) {
    __shared__ ChirpShared S;
    extern __shared__ __align__(16) unsigned char chirp_smem[];
    ChirpWarpBuf* WB = reinterpret_cast<ChirpWarpBuf*>(chirp_smem);   // one staging buffer per warp
    const int tid = threadIdx.x;
    const float* x = samples + static_cast<size_t>(blockIdx.x) * frame_stride;
    int success = 0, up_start = -1, down_start = -1, start = -1;   // DualChirpResult defaults (chirp_sync.hpp:317-324)
    float cfo = 0.0f, up_corr = 0.0f, dn_corr = 0.0f, phase = 0.0f;
    do {
        if (L < 2 * c.n + c.gap) break;                                                       // :368-372
        const int up_pos = chirp_detect_template(S, WB, x, L, c.up_s, c.up_c, c.n, c.up_e, threshold, &up_corr);
        if (up_pos < 0) break;
        const int ds = up_pos + c.n / 2, expected = up_pos + c.n + c.gap, margin = 2 * c.n;      // :423-435
        int de = min(L, expected + margin);
        if (ds >= L) break;
        if (de <= ds + c.n) de = min(L, ds + 2 * c.n);
        float dc;
        const int rel = chirp_detect_template(S, WB, x + ds, de - ds, c.dn_s, c.dn_c, c.n, c.dn_e, threshold, &dc);
        if (rel < 0) break;
        const int down_pos = rel + ds;
        dn_corr = dc;
        const int expected_gap = c.n + c.gap, actual_gap = down_pos - up_pos;                    // :451-468
        const float gap_error = static_cast<float>(actual_gap - expected_gap);
        cfo = __fdiv_rn(gap_error, __fmul_rn(2.0f, c.cfo_to_samples));
        if (fabsf(cfo) > 100.0f) break;                                                         // :478-482
        const float up_correction = __fmul_rn(cfo, c.cfo_to_samples), down_correction = __fmul_rn(-cfo, c.cfo_to_samples);
        up_start = static_cast<int>(roundf(__fadd_rn(static_cast<float>(up_pos), up_correction)));       // :496-497
        down_start = static_cast<int>(roundf(__fadd_rn(static_cast<float>(down_pos), down_correction)));
        success = 1;
        // OFDMChirpWaveform::detectSync (:156-163): training starts after the down chirp and its gap; process (:177-181): the CFO
        // rotator starts from the phase accumulated since sample 0, wrapped to [-pi, pi]
        start = down_start + c.n + c.gap;
        const double pi = 3.14159265358979323846;
        float ph = static_cast<float>(__ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(-2.0f, pi), static_cast<double>(cfo)), static_cast<double>(start)),
                                                static_cast<double>(c.fs)));
        while (static_cast<double>(ph) > pi) ph = static_cast<float>(__dsub_rn(static_cast<double>(ph), __dmul_rn(2.0f, pi)));
        while (static_cast<double>(ph) < -pi) ph = static_cast<float>(__dadd_rn(static_cast<double>(ph), __dmul_rn(2.0f, pi)));
        phase = ph;
    } while (false);
    if (tid == 0) {
        out_info[blockIdx.x] = make_int4(success, up_start, down_start, start);
        out_f[blockIdx.x] = make_float4(cfo, up_corr, dn_corr, phase);
        const bool usable = success && start >= 0 && start < L;
        if (frame_start) frame_start[blockIdx.x] = usable ? start : 0;
        const int nsym = usable ? (L - start) / sym_len : 0;
        if (frame_nsym) frame_nsym[blockIdx.x] = nsym;
        if (cfo_out) cfo_out[blockIdx.x] = cfo;
        if (phase_out) phase_out[blockIdx.x] = phase;
        if (n_llr) {   // process() hands out soft bits only when processPresynced reports a codeword (ofdm_chirp_waveform.cpp:183-196)
            const int n = max(0, nsym - 2) * llr_per_symbol;
            n_llr[blockIdx.x] = n >= 648 ? min(n, llr_stride) : 0;
        }
    }
}

cudaError_t chirp_detect_launch(const ChirpDev& c, const float* samples, size_t B, size_t frame_stride, int L, float threshold, int sym_len,
                                int4* out_info, float4* out_f, int* frame_start, int* frame_nsym, float* cfo_out, float* phase_out,
                                int* n_llr, int llr_per_symbol, int llr_stride, cudaStream_t st) {
    if (B == 0) return cudaSuccess;
    const size_t smem = sizeof(ChirpWarpBuf) * (kChirpThreads / 32);
    cudaFuncSetAttribute(chirp_detect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    chirp_detect_kernel<<<static_cast<unsigned>(B), kChirpThreads, smem, st>>>(c, samples, frame_stride, L, threshold, sym_len, out_info, out_f,
                                                                           frame_start, frame_nsym, cfo_out, phase_out, n_llr,
                                                                           llr_per_symbol, llr_stride);
    return cudaGetLastError();
}

void chirp_templates_host(float fs, std::vector<float>& t, ChirpDev& c) {
    const float f_start = 300.0f, f_end = 2700.0f, duration_ms = 500.0f, gap_ms = 100.0f;
    const size_t n = static_cast<size_t>(fs * duration_ms / 1000.0f);
    const float T = duration_ms / 1000.0f, k = (f_end - f_start) / T;
    t.assign(4 * n, 0.0f);
    float ue = 0.0f, de = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        const float tt = static_cast<float>(i) / fs;
        const float phase = static_cast<float>(2.0f * 3.14159265358979323846 * (f_start * tt + 0.5f * k * tt * tt));
        t[i] = std::sin(phase);
        t[n + i] = std::cos(phase);
        ue += t[i] * t[i];
    }
    for (size_t i = 0; i < n; ++i) {
        const float tt = static_cast<float>(i) / fs;
        const float phase = static_cast<float>(2.0f * 3.14159265358979323846 * (f_end * tt - 0.5f * k * tt * tt));
        t[2 * n + i] = std::sin(phase);
        t[3 * n + i] = std::cos(phase);
        de += t[2 * n + i] * t[2 * n + i];
    }
    c.n = static_cast<int>(n);
    c.gap = static_cast<int>(static_cast<size_t>(fs * gap_ms / 1000.0f));
    c.fs = fs;
    c.cfo_to_samples = fs / ((f_end - f_start) / T);
    c.up_e = ue; c.dn_e = de;
    c.up_s = c.up_c = c.dn_s = c.dn_c = nullptr;
}

}  // namespace pu
