// projectultra_b200/csrc/ofdm_acquire.cu — batched Schmidl-Cox acquisition of OFDM frames: the
// SEARCHING state of OFDMDemodulator::process, run for thousands of frames at once, one frame per CTA.
//
// Reference behaviour (every step restated in the reference's order of floating-point operations, so the sync offset,
// the coarse CFO and the consumed-sample count are the reference's own):
//   OFDMDemodulator::process, SEARCHING state            src/ofdm/demodulator.cpp:474-600
//   Impl::hasMinimumEnergy (stateful noise floor)        src/ofdm/ofdm_sync.cpp:20-50
//   Impl::toAnalytic (Hilbert via FFT / inverse FFT)     :56-84, src/dsp/fft.cpp:89-121
//   Impl::measureSchmidlCoxCorrelation                   :118-163
//   Impl::estimateCoarseCFO                              :230-261
//   Impl::refineLTSTiming                                :386-461
// The caller's frame is fed to process() in `chunk`-sample pieces (tools/test_mode_snr.cpp:65-70 uses 960): the search
// runs at every call once MIN_SEARCH_SAMPLES are buffered, always from offset 0, and what it can see depends on how
// much has arrived -- the kernel replays that sequence of calls.  Frames longer than 2 * OVERLAP_SAMPLES (40 000) would
// make the reference trim its buffer between calls; they are rejected by the host wrapper.
//
// Machine mapping: the state machine is executed redundantly (uniformly) by all threads of the CTA; the analytic
// signal's two FFTs are radix-2 stages over a shared-memory buffer with all threads (any schedule of the reference's
// butterflies gives the reference's bits); the ordered sums of a probe (DC, P, R1, R2) are serial chains and run on
// one thread each, in different warps so that they overlap; the 1961 candidates of the LTS matched filter are
// independent 560-tap ordered sums, one candidate per thread.
#include <cfloat>

#include "fft_smem.cuh"
#include "ofdm_dev.cuh"
#include "pu_internal.h"
#include "ref_math.cuh"

namespace pu {

struct AcqDev {
    int nfft, log2n, cp, sym_len;      // sym_len here = nfft + cp + guard (data symbols); preamble symbols are nfft + cp
    float sample_rate, sync_threshold;
    const float2* twiddle;             // [nfft / 2]
    const float* lts_i;                // [nfft + cp] LTS passband templates (generateSequences, demodulator.cpp:99-132)
    const float* lts_q;
    float lts_energy_ref;              // 0.5 * sum(I^2, Q^2 interleaved), refineLTSTiming :405-411
    float lts_threshold;               // 0.35 (512-FFT) / 0.05 (>= 1024), :451
};

constexpr int kAcqThreads = 256;
constexpr int kMinSearchSamples = 4000;       // demodulator_constants.hpp:41
constexpr int kSearchStep = 8;                // :50
constexpr float kPlateauThreshold = 0.90f;    // :51
constexpr int kPlateauWindow = 300;           // :52
constexpr int kMinPlateau = 15;               // :53
constexpr int kEnergyStep = 16;               // :127
constexpr float kNoiseFloorAlpha = 0.01f;     // :46
constexpr float kEnergyRatio = 4.0f;          // :58

struct AcqShared {
    float dc, R1, R2, corr, nf;
    float2 P;
    int flag;
    float best_corr[kAcqThreads / 32];
    int best_off[kAcqThreads / 32];
};

// toAnalytic of the NFFT samples x[0..NFFT) minus dc (ofdm_sync.cpp:56-84): result in buf.
template <int NFFT>
__device__ void analytic_smem(float2* buf, const float* __restrict__ x, float dc, bool remove_dc, const float2* __restrict__ tw) {
    const int tid = threadIdx.x, T = blockDim.x;
    for (int i = tid; i < NFFT; i += T) buf[i] = make_float2(remove_dc ? __fsub_rn(x[i], dc) : x[i], 0.0f);
    __syncthreads();
    fft_smem<NFFT>(buf, tw, false);
    for (int i = tid; i < NFFT; i += T) {
        if (i >= 1 && i < NFFT / 2) buf[i] = make_float2(__fmul_rn(buf[i].x, 2.0f), __fmul_rn(buf[i].y, 2.0f));   // freq[i] *= 2.0f
        else if (i > NFFT / 2) buf[i] = make_float2(0.0f, 0.0f);
    }
    __syncthreads();
    fft_smem<NFFT>(buf, tw, true);
}

// measureSchmidlCoxCorrelation(offset) over a buffer of `size` samples (ofdm_sync.cpp:118-163); uniform result.
// The value depends on the samples only (the buffer size enters through the bound check), and every call of process()
// searches again from offset 0: probes are memoised per offset (all probe offsets are multiples of 8) for the life of
// the frame, which removes ~3/4 of the FFTs of a frame that synchronises late or never.
constexpr int kAcqMaxProbes = 40000 / 8 + 1;
template <int NFFT>
__device__ float sc_correlation(const AcqDev& a, float2* buf, AcqShared& S, float* __restrict__ memo, const float* __restrict__ x, int size,
                                int offset) {
    if (offset + a.cp + NFFT > size) return 0.0f;
    const float cached = memo[offset >> 3];
    if (cached >= 0.0f) return cached;              // uniform: written below by thread 0 between two barriers
    const float* w = x + offset + a.cp;
    const int tid = threadIdx.x;
    if (tid == 0) {                                   // ordered DC sum (:131-135)
        float sum = 0.0f;
        for (int i = 0; i < NFFT; ++i) sum = __fadd_rn(sum, w[i]);
        S.dc = __fdiv_rn(sum, static_cast<float>(NFFT));
    }
    __syncthreads();
    analytic_smem<NFFT>(buf, w, S.dc, true, a.twiddle);
    constexpr int H = NFFT / 2;
    if (tid == 0) {                                   // P += conj(a[i]) * a[i + half] (:147-149)
        float2 P = make_float2(0.0f, 0.0f);
        for (int i = 0; i < H; ++i) P = cadd(P, cmul(cconj(buf[i]), buf[i + H]));
        S.P = P;
    } else if (tid == 32) {
        float r = 0.0f;
        for (int i = 0; i < H; ++i) r = __fadd_rn(r, cnorm(buf[i]));
        S.R1 = r;
    } else if (tid == 64) {
        float r = 0.0f;
        for (int i = 0; i < H; ++i) r = __fadd_rn(r, cnorm(buf[i + H]));
        S.R2 = r;
    }
    __syncthreads();
    const float norm = __fsqrt_rn(__fmul_rn(S.R1, S.R2));
    const float corr = (norm < 1e-10f) ? 0.0f : __fdiv_rn(cabs_ref(S.P), norm);
    if (tid == 0) memo[offset >> 3] = corr;
    __syncthreads();                                  // S.* may be rewritten by the next probe; the memo entry is visible
    return corr;
}

// out_int[b] = {found, sync_offset, data_start (samples consumed), n_calls}; out_cfo[b] = coarse CFO (Hz)
template <int NFFT>
__global__ void __launch_bounds__(kAcqThreads) ofdm_acquire_kernel(AcqDev a, const float* __restrict__ samples, size_t frame_stride,
                                                                   int L, int chunk, int4* __restrict__ out_int, float* __restrict__ out_cfo) {
    __shared__ float2 buf[NFFT];
    __shared__ AcqShared S;
    __shared__ float memo[kAcqMaxProbes];               // Schmidl-Cox correlation per probe offset / 8, < 0 = not computed yet
    const int tid = threadIdx.x;
    for (int i = tid; i < kAcqMaxProbes; i += blockDim.x) memo[i] = -1.0f;
    __syncthreads();
    const float* x = samples + static_cast<size_t>(blockIdx.x) * frame_stride;
    const int P = NFFT + a.cp;                          // preamble_symbol_len
    const int total = 6 * P, window = 2 * P;            // preamble_total_len, correlation_window (demodulator.cpp:466-468)
    float nf = 0.0f;                                    // noise_floor_energy (demodulator_impl.hpp:62), uniform
    int found = 0, sync_offset = 0, data_start = 0, calls = 0;
    float cfo = 0.0f;

    for (int size = min(chunk, L);; size = min(size + chunk, L)) {
        ++calls;
        if (size >= kMinSearchSamples) {
            const int search_end = (size > total + window) ? size - total - window : 0;
            bool hit = false;
            int peak_pos = 0;
            for (int i = 0; i < search_end; i += kSearchStep) {
                // ---- hasMinimumEnergy(i, window) (ofdm_sync.cpp:20-50)
                bool enough = false;
                if (i + window <= size) {
                    if (tid == 0) {
                        float sum_sq = 0.0f;
                        int count = 0;
                        for (int k = 0; k < window; k += kEnergyStep) {
                            const float s = x[i + k];
                            sum_sq = __fadd_rn(sum_sq, __fmul_rn(s, s));
                            ++count;
                        }
                        const float energy = __fdiv_rn(sum_sq, static_cast<float>(count));
                        float f = nf;
                        if (f < 1e-20f) f = __fmul_rn(energy, 0.1f);
                        if (energy < f) f = energy;
                        else if (energy < __fmul_rn(f, 3.0f))
                            f = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, kNoiseFloorAlpha), f), __fmul_rn(kNoiseFloorAlpha, energy));
                        S.nf = f;
                        S.flag = energy >= __fmul_rn(f, kEnergyRatio);
                    }
                    __syncthreads();
                    nf = S.nf;
                    enough = S.flag != 0;
                    __syncthreads();
                }
                if (!enough) { i += window / 2 - kSearchStep; continue; }
                const float corr = sc_correlation<NFFT>(a, buf, S, memo, x, size, i);
                if (corr > a.sync_threshold) {
                    // ---- plateau search (demodulator.cpp:504-531)
                    int plateau = 0;
                    float peak = corr;
                    peak_pos = i;
                    for (int j = 0; j <= kPlateauWindow && i + j + total < size; j += 8) {
                        const float r = sc_correlation<NFFT>(a, buf, S, memo, x, size, i + j);
                        if (r >= kPlateauThreshold) ++plateau;
                        if (r > peak) { peak = r; peak_pos = i + j; }
                    }
                    if (plateau >= kMinPlateau) { hit = true; break; }
                }
            }
            if (hit) {
                // ---- estimateCoarseCFO(sync_offset) (ofdm_sync.cpp:230-261)
                float coarse = 0.0f;
                if (peak_pos + a.cp + NFFT <= size) {
                    analytic_smem<NFFT>(buf, x + peak_pos + a.cp, 0.0f, false, a.twiddle);
                    if (tid == 0) {
                        float2 Pc = make_float2(0.0f, 0.0f);
                        for (int i = 0; i < NFFT / 2; ++i) Pc = cadd(Pc, cmul(cconj(buf[i]), buf[i + NFFT / 2]));
                        S.P = Pc;
                    }
                    __syncthreads();
                    const float phase = refmath::atan2f_ref(S.P.y, S.P.x);
                    // phase * sample_rate / (M_PI * fft_len): float product, division in double, rounded to float
                    coarse = static_cast<float>(__ddiv_rn(static_cast<double>(__fmul_rn(phase, a.sample_rate)),
                                                          __dmul_rn(3.14159265358979323846, static_cast<double>(NFFT))));
                    const float max_cfo = static_cast<float>(static_cast<unsigned>(a.sample_rate) / static_cast<unsigned>(NFFT));
                    coarse = fmaxf(-max_cfo, fminf(max_cfo, coarse));
                    __syncthreads();
                }
                // ---- refineLTSTiming(sync_offset) (ofdm_sync.cpp:386-461)
                const int coarse_lts = peak_pos + 4 * P;
                const int back = 3 * P, fwd = P / 2;
                long refined;
                if (coarse_lts < back || coarse_lts + fwd + P > size) {
                    refined = coarse_lts;                                  // "not enough data, using coarse timing"
                } else {
                    float bc = 0.0f;
                    int bo = coarse_lts;
                    for (int delta = -back + tid; delta <= fwd; delta += blockDim.x) {
                        const float* r = x + coarse_lts + delta;
                        float ci = 0.0f, cq = 0.0f, er = 0.0f;
                        for (int k = 0; k < P; ++k) {
                            const float v = r[k];
                            ci = __fadd_rn(ci, __fmul_rn(v, __ldg(&a.lts_i[k])));
                            cq = __fadd_rn(cq, __fmul_rn(v, __ldg(&a.lts_q[k])));
                            er = __fadd_rn(er, __fmul_rn(v, v));
                        }
                        const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(ci, ci), __fmul_rn(cq, cq)));
                        const float norm = __fsqrt_rn(__fmul_rn(er, a.lts_energy_ref));
                        const float c = (norm > 1e-6f) ? __fdiv_rn(mag, norm) : 0.0f;
                        if (c > bc) { bc = c; bo = coarse_lts + delta; }   // ascending delta per thread: first maximum kept
                    }
                    // first maximum over all candidates: larger correlation wins, equal correlation -> smaller offset
                    for (int o = 16; o > 0; o >>= 1) {
                        const float oc = __shfl_xor_sync(0xffffffffu, bc, o);
                        const int oo = __shfl_xor_sync(0xffffffffu, bo, o);
                        if (oc > bc || (oc == bc && oc > 0.0f && oo < bo)) { bc = oc; bo = oo; }
                    }
                    if ((tid & 31) == 0) { S.best_corr[tid >> 5] = bc; S.best_off[tid >> 5] = bo; }
                    __syncthreads();
                    bc = S.best_corr[0];
                    bo = S.best_off[0];
                    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) {
                        const float oc = S.best_corr[w];
                        const int oo = S.best_off[w];
                        if (oc > bc || (oc == bc && oc > 0.0f && oo < bo)) { bc = oc; bo = oo; }
                    }
                    __syncthreads();
                    refined = (bc < a.lts_threshold) ? -1 : bo;            // SIZE_MAX: Schmidl-Cox false positive
                }
                if (refined >= 0) {
                    found = 1;
                    sync_offset = peak_pos;
                    cfo = coarse;
                    data_start = static_cast<int>(refined) + 2 * P;         // consume (manual_timing_offset == 0)
                    break;
                }
                // false positive: nothing is trimmed (size <= 2 * OVERLAP_SAMPLES); the next call searches again from 0
            }
        }
        if (size >= L) break;
    }
    if (tid == 0) {
        out_int[blockIdx.x] = make_int4(found, sync_offset, data_start, calls);
        out_cfo[blockIdx.x] = cfo;
    }
}

cudaError_t ofdm_acquire_launch(const AcqDev& a, const float* samples, size_t B, size_t frame_stride, int L, int chunk, int4* out_int,
                                float* out_cfo, cudaStream_t st) {
    if (B == 0) return cudaSuccess;
    if (a.nfft == 512)
        ofdm_acquire_kernel<512><<<static_cast<unsigned>(B), kAcqThreads, 0, st>>>(a, samples, frame_stride, L, chunk, out_int, out_cfo);
    else
        ofdm_acquire_kernel<1024><<<static_cast<unsigned>(B), kAcqThreads, 0, st>>>(a, samples, frame_stride, L, chunk, out_int, out_cfo);
    return cudaGetLastError();
}

}  // namespace pu
