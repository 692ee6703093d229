// projectultra_b200/csrc/psk_demod.cu — batched single-carrier and multi-carrier DPSK soft demodulators for sm_100a and
// the pu_dpsk_* / pu_mcdpsk_* entry points of the C ABI (externally timed frames: the Monte-Carlo sim knows where the
// data starts; Barker / chirp acquisition is SURVEY §8f next-2).
//
// Reference behaviour
//   DPSKDemodulator (src/psk/dpsk.hpp): carrier tables :311-323, correlateSymbol :777-787 (I += x cos, Q -= x sin,
//   sequential fp32 sums, / N), demodulateSoft :827-879 (diff = cur conj(prev), |diff|, atan2, optional CFO / initial
//   phase compensation :857-865, confidence = min(10 |diff|, 5)), phaseToBits :1002-1052, setReferenceSymbol :889-892.
//   MultiCarrierDPSKDemodulator (src/psk/multi_carrier_dpsk.hpp): demodulateOneSymbol :663-678 (per-carrier DFT bin
//   with std::polar(1, -phase), phase accumulated in fp32), processTraining :390-422, setReference :424-435,
//   demodulateSoft :437-472 (normalised differential, confidence = |cur| N_c 4, clamp +-10).
//
// Numerics: the correlations are SEQUENTIAL fp32 sums in the reference, so each (symbol[, carrier]) sum is one thread
// walking its samples in order with unfused multiply/add -- bit-identical; parallelism comes from the thousands of
// symbols in a batch.  Carrier / mixer tables are computed on the host with the host libm exactly as the reference
// computes them (MC-DPSK mixer phases reach 167 rad).  atan2f / sinf / cosf / hypotf are the restatements of
// ref_math.cuh.  Compiled with -fmad=false.
//
// Kernel shape: correlation = a warp owns 32 (symbol[, carrier]) sums; samples are staged through a padded shared
// tile with coalesced 128-byte row loads, 32 samples per row per round, so HBM sees each sample once (bound: HBM,
// 4 bytes and 4 flops per sample).  A second, tiny kernel turns neighbouring correlations into LLRs.
#include <cmath>
#include <complex>
#include <memory>
#include <new>
#include <vector>

#include "ofdm_dev.cuh"
#include "pu_internal.h"

namespace pu {

std::vector<float> mcdpsk_carrier_freqs(const pu_mcdpsk_config& c);   // psk_tx.cpp

constexpr int kPskThreads = 128;
constexpr int kPskChunk = 32;

// ---------------------------------------------------------------------------------------------- single carrier
// corr[frame][s] for s in [0, n_corr): symbol s starts at sample first_start + s * sps.
__global__ void __launch_bounds__(kPskThreads) dpsk_correlate_kernel(const float* __restrict__ samples, size_t frame_stride,
                                                                     long first_start, int sps, int n_corr,
                                                                     const float* __restrict__ ccos, const float* __restrict__ csin,
                                                                     float2* __restrict__ corr) {
    extern __shared__ float sm[];
    float* tcos = sm;                       // [sps]
    float* tsin = sm + sps;                 // [sps]
    float* tile = sm + 2 * sps;             // [warps][32][33]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < sps; i += blockDim.x) {
        tcos[i] = __ldg(&ccos[i]);
        tsin[i] = __ldg(&csin[i]);
    }
    __syncthreads();
    const size_t frame = blockIdx.y;
    const float* x = samples + frame * frame_stride + first_start;
    const int s0 = (blockIdx.x * (kPskThreads / 32) + warp) * 32;   // first symbol of this warp
    if (s0 >= n_corr) return;
    float* wt = tile + warp * 32 * 33;
    const int s = s0 + lane;
    float I = 0.0f, Q = 0.0f;
    for (int c0 = 0; c0 < sps; c0 += kPskChunk) {
        const int cn = min(kPskChunk, sps - c0);
        __syncwarp();
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {                              // coalesced: row r = symbol s0 + r, 32 consecutive samples
            float v = 0.0f;
            if (s0 + r < n_corr && lane < cn) v = __ldg(&x[static_cast<size_t>(s0 + r) * sps + c0 + lane]);
            wt[r * 33 + lane] = v;
        }
        __syncwarp();
        for (int i = 0; i < cn; ++i) {                              // dpsk.hpp:781-784, in order, unfused
            const float xv = wt[lane * 33 + i];
            I = __fadd_rn(I, __fmul_rn(xv, tcos[c0 + i]));
            Q = __fsub_rn(Q, __fmul_rn(xv, tsin[c0 + i]));
        }
    }
    if (s < n_corr) corr[frame * n_corr + s] = make_float2(__fdiv_rn(I, static_cast<float>(sps)), __fdiv_rn(Q, static_cast<float>(sps)));
}

__device__ __forceinline__ float wrap_0_2pi(float phase) {   // while (phase < 0) phase += 2 pi; while (phase >= 2 pi) phase -= 2 pi
    const double two_pi = 2.0f * 3.14159265358979323846;
    while (phase < 0.0f) phase = static_cast<float>(__dadd_rn(static_cast<double>(phase), two_pi));
    while (static_cast<double>(phase) >= two_pi) phase = static_cast<float>(__dsub_rn(static_cast<double>(phase), two_pi));
    return phase;
}

__global__ void dpsk_llr_kernel(const float2* __restrict__ corr, int n_corr, int has_ref, int n_sym, int mod, int sps, float sample_rate,
                                const float* __restrict__ est_cfo, const float* __restrict__ phase_off, size_t B,
                                float* __restrict__ llr, size_t llr_stride) {
    const size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (g >= B * static_cast<size_t>(n_sym)) return;
    const size_t frame = g / n_sym;
    const int s = static_cast<int>(g - frame * n_sym);
    const float2* c = corr + frame * n_corr + has_ref;
    const float2 cur = c[s];
    const float2 prev = (s > 0 || has_ref) ? c[s - 1] : make_float2(1.0f, 0.0f);   // prev_symbol_ (:840, :889-892)
    const float2 df = cmul(cur, cconj(prev));                                       // :848
    const float magnitude = cabs_ref(df);                                           // :851
    float phase = refmath::atan2f_ref(df.y, df.x);                                  // :854
    const float cfo = est_cfo ? est_cfo[frame] : 0.0f, poff = phase_off ? phase_off[frame] : 0.0f;
    if (fabsf(cfo) > 0.5f || fabsf(poff) > 0.01f) {                                 // :857-865
        const double two_pi = 2.0f * 3.14159265358979323846, pi = 3.14159265358979323846;
        const float cfo_phase = static_cast<float>(__ddiv_rn(__dmul_rn(__dmul_rn(two_pi, static_cast<double>(cfo)), static_cast<double>(sps)),
                                                             static_cast<double>(sample_rate)));
        phase = __fsub_rn(phase, cfo_phase);
        phase = __fsub_rn(phase, poff);
        while (static_cast<double>(phase) > pi) phase = static_cast<float>(__dsub_rn(static_cast<double>(phase), two_pi));
        while (static_cast<double>(phase) < -pi) phase = static_cast<float>(__dadd_rn(static_cast<double>(phase), two_pi));
    }
    const float confidence = fminf(__fmul_rn(magnitude, 10.0f), 5.0f);              // :868
    phase = wrap_0_2pi(phase);                                                      // phaseToBits :1006-1007
    const int bps = mod + 1;
    float l[3];
    if (mod == 0) {
        l[0] = __fmul_rn(confidence, refmath::cosf_ref(phase));
    } else {
        l[0] = __fmul_rn(confidence, refmath::sinf_ref(phase));
        l[1] = __fmul_rn(confidence, refmath::sinf_ref(__fmul_rn(2.0f, phase)));
        if (mod == 2) l[2] = __fmul_rn(confidence, refmath::sinf_ref(__fmul_rn(4.0f, phase)));
    }
    float* out = llr + frame * llr_stride;
    for (int b = 0; b < bps; ++b) {
        const size_t pos = static_cast<size_t>(s) * bps + b;
        if (pos < llr_stride) out[pos] = l[b];
    }
}

// ---------------------------------------------------------------------------------------------- multi carrier
// corr[frame][sym][c]: pair index t = sym * nc + c; a CTA owns kPskThreads consecutive pairs of one frame.
__global__ void __launch_bounds__(kPskThreads) mcdpsk_correlate_kernel(const float* __restrict__ samples, size_t frame_stride, int sps,
                                                                       int nc, int n_sym, const float2* __restrict__ mixer,
                                                                       float2* __restrict__ corr) {
    extern __shared__ float sm[];            // tile [rows][33]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t frame = blockIdx.y;
    const float* x = samples + frame * frame_stride;
    const int n_pairs = n_sym * nc;
    const int t0 = blockIdx.x * kPskThreads;
    const int t = t0 + tid;
    const bool live = t < n_pairs;
    const int sym_lo = t0 / nc;
    const int sym_hi = min(n_sym - 1, (min(t0 + kPskThreads, n_pairs) - 1) / nc);
    const int rows = sym_hi - sym_lo + 1;
    const int sym = live ? t / nc : sym_lo, c = live ? t - sym * nc : 0;
    const float2* mx = mixer + static_cast<size_t>(c) * sps;
    const int row = sym - sym_lo;
    float sr = 0.0f, si = 0.0f;
    for (int c0 = 0; c0 < sps; c0 += kPskChunk) {
        const int cn = min(kPskChunk, sps - c0);
        __syncthreads();
        for (int r = warp; r < rows; r += kPskThreads / 32)
            sm[r * 33 + lane] = (lane < cn) ? __ldg(&x[static_cast<size_t>(sym_lo + r) * sps + c0 + lane]) : 0.0f;
        __syncthreads();
        for (int i = 0; i < cn; ++i) {                               // multi_carrier_dpsk.hpp:671-675, in order, unfused
            const float xv = sm[row * 33 + i];
            const float2 m = __ldg(&mx[c0 + i]);
            sr = __fadd_rn(sr, __fmul_rn(m.x, xv));
            si = __fadd_rn(si, __fmul_rn(m.y, xv));
        }
    }
    if (live) corr[frame * n_pairs + t] = make_float2(__fdiv_rn(sr, static_cast<float>(sps)), __fdiv_rn(si, static_cast<float>(sps)));
}

__global__ void mcdpsk_llr_kernel(const float2* __restrict__ corr, int nc, int n_sym, int training, int bits, int sps, float sample_rate,
                                  const float2* __restrict__ expected_diff, size_t B, float* __restrict__ llr, size_t llr_stride,
                                  float* __restrict__ residual_cfo) {
    const int n_data = n_sym - training - 1;
    const size_t per_frame = static_cast<size_t>(n_data > 0 ? n_data : 0) * nc;
    const size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (residual_cfo && g < B) {                                     // processTraining (:390-422), one thread per frame
        const float2* c0 = corr + g * static_cast<size_t>(n_sym) * nc;
        float cfo = 0.0f;
        if (training >= 2) {
            float sum = 0.0f;
            for (int c = 0; c < nc; ++c) {
                const float2 actual = cmul(c0[nc + c], cconj(c0[c]));
                const float2 err = cmul(actual, cconj(expected_diff[c]));
                sum = __fadd_rn(sum, refmath::atan2f_ref(err.y, err.x));
            }
            const float avg = __fdiv_rn(sum, static_cast<float>(nc));
            const float symbol_duration = __fdiv_rn(static_cast<float>(sps), sample_rate);
            cfo = static_cast<float>(__ddiv_rn(static_cast<double>(avg), __dmul_rn(2.0f * 3.14159265358979323846, static_cast<double>(symbol_duration))));
            cfo = fmaxf(-50.0f, fminf(50.0f, __fadd_rn(0.0f, cfo)));   // cfo_hz_ += residual, clamped (:420-421)
        }
        residual_cfo[g] = cfo;
    }
    if (g >= B * per_frame) return;
    const size_t frame = g / per_frame;
    const int t = static_cast<int>(g - frame * per_frame);
    const int sd = t / nc, c = t - sd * nc;
    const float2* cf = corr + frame * static_cast<size_t>(n_sym) * nc;
    const float2 cur = cf[(training + 1 + sd) * nc + c];
    const float2 pv = cf[(training + sd) * nc + c];                  // previous data symbol, or the reference symbol for sd == 0
    const float pmag = cabs_ref(pv);
    const float pthr = sd == 0 ? 0.001f : 0.0001f;                   // setReference (:430) vs demodulateSoft (:448)
    const float2 prev = pmag > pthr ? cdivs(pv, pmag) : make_float2(1.0f, 0.0f);
    const float mag = cabs_ref(cur);
    const float2 nrm = mag > 0.0001f ? cdivs(cur, mag) : make_float2(1.0f, 0.0f);
    const float2 df = cmul(nrm, cconj(prev));
    float phase = refmath::atan2f_ref(df.y, df.x);
    const float confidence = __fmul_rn(__fmul_rn(mag, static_cast<float>(nc)), 4.0f);
    phase = wrap_0_2pi(phase);
    float* out = llr + frame * llr_stride;
    const size_t base = static_cast<size_t>(t) * bits;
    if (bits == 2) {
        const float sb0 = __fmul_rn(confidence, refmath::sinf_ref(phase));
        const float sb1 = __fmul_rn(confidence, refmath::sinf_ref(__fmul_rn(2.0f, phase)));
        if (base < llr_stride) out[base] = fmaxf(-10.0f, fminf(10.0f, sb0));
        if (base + 1 < llr_stride) out[base + 1] = fmaxf(-10.0f, fminf(10.0f, sb1));
    } else {
        const float sb = __fmul_rn(confidence, refmath::cosf_ref(phase));
        if (base < llr_stride) out[base] = fmaxf(-10.0f, fminf(10.0f, sb));
    }
}

struct PskDevMem {
    void* p = nullptr;
    ~PskDevMem() { if (p) cudaFree(p); }
    template <class T>
    pu_status upload(const T* src, size_t n) {
        if (p) { cudaFree(p); p = nullptr; }
        PU_CUDA_TRY(cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 16)));
        if (n) PU_CUDA_TRY(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
        return PU_OK;
    }
};

}  // namespace pu

struct pu_dpsk {
    pu_ctx* ctx = nullptr;
    int device = 0;
    pu_dpsk_config cfg{};
    pu::PskDevMem d_cos, d_sin;
    pu::Buffer corr;
};

struct pu_mcdpsk {
    pu_ctx* ctx = nullptr;
    int device = 0;
    pu_mcdpsk_config cfg{};
    pu::PskDevMem d_mixer, d_expected;
    pu::Buffer corr;
};

// Shared host-staging helper: samples (+ up to two per-frame float arrays) up, `out_floats` per frame (+ one per-frame
// float array) back.
template <class Launch>
static pu_status psk_host_call(pu_ctx* ctx, cudaStream_t st, const float* samples, size_t B, size_t L, const float* a0, const float* a1,
                               float* out, size_t out_stride, float* aux_out, Launch&& launch) {
    const size_t slab = std::max<size_t>(1, std::min<size_t>(B, (64u << 20) / std::max<size_t>(L * sizeof(float), 1)));
    pu_status s;
    const size_t in_floats = slab * (L + 2), out_floats = slab * (out_stride + 1);
    if ((s = ctx->d_in.reserve(in_floats * sizeof(float))) != PU_OK) return s;
    if ((s = ctx->h_in.reserve(in_floats * sizeof(float))) != PU_OK) return s;
    if ((s = ctx->d_out.reserve(out_floats * sizeof(float))) != PU_OK) return s;
    if ((s = ctx->h_out.reserve(out_floats * sizeof(float))) != PU_OK) return s;
    for (size_t off = 0; off < B; off += slab) {
        const size_t nb = std::min(slab, B - off);
        float* hin = static_cast<float*>(ctx->h_in.ptr);
        std::memcpy(hin, samples + off * L, nb * L * sizeof(float));
        for (size_t b = 0; b < nb; ++b) {
            hin[slab * L + b] = a0 ? a0[off + b] : 0.0f;
            hin[slab * L + slab + b] = a1 ? a1[off + b] : 0.0f;
        }
        float* din = static_cast<float*>(ctx->d_in.ptr);
        float* dout = static_cast<float*>(ctx->d_out.ptr);
        PU_CUDA_TRY(cudaMemcpyAsync(din, hin, in_floats * sizeof(float), cudaMemcpyHostToDevice, st));
        PU_CUDA_TRY(cudaMemsetAsync(dout, 0, out_floats * sizeof(float), st));
        if ((s = launch(din, nb, a0 ? din + slab * L : nullptr, a1 ? din + slab * L + slab : nullptr, dout, dout + slab * out_stride)) != PU_OK) return s;
        float* hout = static_cast<float*>(ctx->h_out.ptr);
        PU_CUDA_TRY(cudaMemcpyAsync(hout, dout, out_floats * sizeof(float), cudaMemcpyDeviceToHost, st));
        PU_CUDA_TRY(cudaStreamSynchronize(st));
        std::memcpy(out + off * out_stride, hout, nb * out_stride * sizeof(float));
        if (aux_out) std::memcpy(aux_out + off, hout + slab * out_stride, nb * sizeof(float));
    }
    return PU_OK;
}

extern "C" {

pu_status pu_dpsk_create(pu_ctx* ctx, const pu_dpsk_config* cfg, pu_dpsk** out) {
    PU_REQUIRE(ctx && cfg && out, "pu_dpsk_create: NULL argument");
    *out = nullptr;
    if (cfg->samples_per_symbol == 0 || cfg->samples_per_symbol > 8192 || cfg->modulation > 2 || !(cfg->sample_rate > 0)) {
        pu::set_error("pu_dpsk_create: samples_per_symbol must be in [1, 8192], modulation 0..2 (DBPSK/DQPSK/D8PSK)");
        return PU_ERR_UNSUPPORTED;
    }
    PU_CUDA_TRY(cudaSetDevice(ctx->device));
    std::unique_ptr<pu_dpsk> h(new (std::nothrow) pu_dpsk());
    if (!h) return PU_ERR_NOMEM;
    h->ctx = ctx;
    h->device = ctx->device;
    h->cfg = *cfg;
    const int n = static_cast<int>(cfg->samples_per_symbol);
    std::vector<float> cs(n), sn(n);
    const float inc = static_cast<float>(2.0f * 3.14159265358979323846 * cfg->carrier_freq / cfg->sample_rate);   // dpsk.hpp:315
    for (int i = 0; i < n; ++i) {
        const float phase = inc * i;
        sn[i] = std::sin(phase);
        cs[i] = std::cos(phase);
    }
    pu_status s;
    if ((s = h->d_cos.upload(cs.data(), cs.size())) != PU_OK) return s;
    if ((s = h->d_sin.upload(sn.data(), sn.size())) != PU_OK) return s;
    *out = h.release();
    return PU_OK;
}

void pu_dpsk_destroy(pu_dpsk* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    h->corr.release();
    delete h;
}

int pu_dpsk_bits_per_symbol(const pu_dpsk* h) { return h ? static_cast<int>(h->cfg.modulation) + 1 : -1; }

static pu_status dpsk_launch(pu_dpsk* h, const float* d_samples, size_t B, size_t L, size_t data_start, int ref_mode,
                             const float* d_cfo, const float* d_poff, float* d_llr, size_t llr_stride, cudaStream_t st) {
    const int sps = static_cast<int>(h->cfg.samples_per_symbol);
    const int has_ref = (ref_mode == 1 && data_start >= static_cast<size_t>(sps)) ? 1 : 0;
    const int n_sym = static_cast<int>((L - data_start) / sps);
    if (n_sym == 0) return PU_OK;
    const int n_corr = n_sym + has_ref;
    pu_status s;
    if ((s = h->corr.reserve(B * static_cast<size_t>(n_corr) * sizeof(float2))) != PU_OK) return s;
    float2* corr = static_cast<float2*>(h->corr.ptr);
    const size_t smem = sizeof(float) * (2 * static_cast<size_t>(sps) + (pu::kPskThreads / 32) * 32 * 33);
    (void)cudaGetLastError();
    if (smem > 48 * 1024) cudaFuncSetAttribute(pu::dpsk_correlate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    const size_t kMaxY = 65535;
    for (size_t off = 0; off < B; off += kMaxY) {
        const size_t nb = std::min(kMaxY, B - off);
        const dim3 grid(static_cast<unsigned>((n_corr + pu::kPskThreads - 1) / pu::kPskThreads), static_cast<unsigned>(nb));
        pu::dpsk_correlate_kernel<<<grid, pu::kPskThreads, smem, st>>>(d_samples + off * L, L, static_cast<long>(data_start) - has_ref * sps, sps,
                                                                     n_corr, static_cast<const float*>(h->d_cos.p),
                                                                     static_cast<const float*>(h->d_sin.p), corr + off * n_corr);
        h->ctx->launches.fetch_add(1);
    }
    const size_t total = B * static_cast<size_t>(n_sym);
    pu::dpsk_llr_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, st>>>(corr, n_corr, has_ref, n_sym, static_cast<int>(h->cfg.modulation),
                                                                                  sps, h->cfg.sample_rate, d_cfo, d_poff, B, d_llr, llr_stride);
    h->ctx->launches.fetch_add(1);
    PU_CUDA_TRY(cudaGetLastError());
    return PU_OK;
}

pu_status pu_dpsk_demod_soft_batch(pu_dpsk* h, const float* samples, size_t B, size_t L, size_t data_start, int ref_mode,
                                   const float* est_cfo_hz, const float* phase_offset, float* llr_out, size_t llr_stride,
                                   pu_memspace space, void* stream) {
    PU_REQUIRE(h, "pu_dpsk_demod_soft_batch: NULL handle");
    if (B == 0) return PU_OK;
    PU_REQUIRE(samples && llr_out, "pu_dpsk_demod_soft_batch: NULL data pointer");
    PU_REQUIRE(data_start <= L, "pu_dpsk_demod_soft_batch: data_start beyond the frame");
    PU_REQUIRE(ref_mode == 0 || ref_mode == 1, "pu_dpsk_demod_soft_batch: ref_mode must be 0 or 1");
    PU_REQUIRE(llr_stride > 0, "pu_dpsk_demod_soft_batch: llr_stride is zero");
    pu_ctx* ctx = h->ctx;
    PU_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = pu::pick_stream(ctx, stream, space);
    if (space == PU_MEM_DEVICE) return dpsk_launch(h, samples, B, L, data_start, ref_mode, est_cfo_hz, phase_offset, llr_out, llr_stride, st);
    return psk_host_call(ctx, st, samples, B, L, est_cfo_hz, phase_offset, llr_out, llr_stride, nullptr,
                         [&](const float* din, size_t nb, const float* c, const float* p, float* dout, float*) {
                             return dpsk_launch(h, din, nb, L, data_start, ref_mode, c, p, dout, llr_stride, st);
                         });
}

pu_status pu_mcdpsk_create(pu_ctx* ctx, const pu_mcdpsk_config* cfg, pu_mcdpsk** out) {
    PU_REQUIRE(ctx && cfg && out, "pu_mcdpsk_create: NULL argument");
    *out = nullptr;
    if (cfg->num_carriers < 1 || cfg->num_carriers > 64 || cfg->samples_per_symbol == 0 || cfg->samples_per_symbol > 8192 ||
        (cfg->bits_per_symbol != 1 && cfg->bits_per_symbol != 2) || !(cfg->sample_rate > 0)) {
        pu::set_error("pu_mcdpsk_create: num_carriers in [1, 64], samples_per_symbol in [1, 8192], bits_per_symbol 1 or 2");
        return PU_ERR_UNSUPPORTED;
    }
    PU_CUDA_TRY(cudaSetDevice(ctx->device));
    std::unique_ptr<pu_mcdpsk> h(new (std::nothrow) pu_mcdpsk());
    if (!h) return PU_ERR_NOMEM;
    h->ctx = ctx;
    h->device = ctx->device;
    h->cfg = *cfg;
    const int nc = static_cast<int>(cfg->num_carriers), sps = static_cast<int>(cfg->samples_per_symbol);
    const std::vector<float> freqs = pu::mcdpsk_carrier_freqs(*cfg);
    std::vector<std::complex<float>> mixer(static_cast<size_t>(nc) * sps), expected(nc);
    for (int c = 0; c < nc; ++c) {
        const float inc = static_cast<float>(2.0f * 3.14159265358979323846 * freqs[c] / cfg->sample_rate);   // :667
        float phase = 0.0f;
        for (int i = 0; i < sps; ++i) {
            mixer[static_cast<size_t>(c) * sps + i] = std::polar(1.0f, -phase);                              // :672
            phase += inc;
        }
        const float expected_phase = static_cast<float>((c * 1 - c * 0) * 3.14159265358979323846 / 2.0f);    // :404
        expected[c] = std::polar(1.0f, expected_phase);
    }
    pu_status s;
    if ((s = h->d_mixer.upload(mixer.data(), mixer.size())) != PU_OK) return s;
    if ((s = h->d_expected.upload(expected.data(), expected.size())) != PU_OK) return s;
    *out = h.release();
    return PU_OK;
}

void pu_mcdpsk_destroy(pu_mcdpsk* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    h->corr.release();
    delete h;
}

static pu_status mcdpsk_launch(pu_mcdpsk* h, const float* d_samples, size_t B, size_t L, float* d_llr, size_t llr_stride, float* d_cfo,
                               cudaStream_t st) {
    const int sps = static_cast<int>(h->cfg.samples_per_symbol), nc = static_cast<int>(h->cfg.num_carriers);
    const int training = static_cast<int>(h->cfg.training_symbols);
    const int n_sym = static_cast<int>(L / sps);
    const int n_pairs = n_sym * nc;
    pu_status s;
    if ((s = h->corr.reserve(B * static_cast<size_t>(n_pairs) * sizeof(float2))) != PU_OK) return s;
    float2* corr = static_cast<float2*>(h->corr.ptr);
    const int rows_max = pu::kPskThreads / nc + 2;
    const size_t smem = sizeof(float) * static_cast<size_t>(rows_max) * 33;
    (void)cudaGetLastError();
    const size_t kMaxY = 65535;
    for (size_t off = 0; off < B; off += kMaxY) {
        const size_t nb = std::min(kMaxY, B - off);
        const dim3 grid(static_cast<unsigned>((n_pairs + pu::kPskThreads - 1) / pu::kPskThreads), static_cast<unsigned>(nb));
        pu::mcdpsk_correlate_kernel<<<grid, pu::kPskThreads, smem, st>>>(d_samples + off * L, L, sps, nc, n_sym,
                                                                       static_cast<const float2*>(h->d_mixer.p), corr + off * n_pairs);
        h->ctx->launches.fetch_add(1);
    }
    const int n_data = n_sym - training - 1;
    const size_t total = std::max(B * static_cast<size_t>(std::max(n_data, 0)) * nc, B);
    pu::mcdpsk_llr_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, st>>>(
        corr, nc, n_sym, training, static_cast<int>(h->cfg.bits_per_symbol), sps, h->cfg.sample_rate,
        static_cast<const float2*>(h->d_expected.p), B, d_llr, llr_stride, d_cfo);
    h->ctx->launches.fetch_add(1);
    PU_CUDA_TRY(cudaGetLastError());
    return PU_OK;
}

pu_status pu_mcdpsk_demod_soft_batch(pu_mcdpsk* h, const float* samples, size_t B, size_t L, float* llr_out, size_t llr_stride,
                                     float* residual_cfo_hz, pu_memspace space, void* stream) {
    PU_REQUIRE(h, "pu_mcdpsk_demod_soft_batch: NULL handle");
    if (B == 0) return PU_OK;
    PU_REQUIRE(samples && llr_out, "pu_mcdpsk_demod_soft_batch: NULL data pointer");
    PU_REQUIRE(llr_stride > 0, "pu_mcdpsk_demod_soft_batch: llr_stride is zero");
    PU_REQUIRE(L >= static_cast<size_t>(h->cfg.training_symbols + 1) * h->cfg.samples_per_symbol,
               "pu_mcdpsk_demod_soft_batch: frame shorter than training + reference symbols");
    pu_ctx* ctx = h->ctx;
    PU_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = pu::pick_stream(ctx, stream, space);
    if (space == PU_MEM_DEVICE) return mcdpsk_launch(h, samples, B, L, llr_out, llr_stride, residual_cfo_hz, st);
    return psk_host_call(ctx, st, samples, B, L, nullptr, nullptr, llr_out, llr_stride, residual_cfo_hz,
                         [&](const float* din, size_t nb, const float*, const float*, float* dout, float* daux) {
                             return mcdpsk_launch(h, din, nb, L, dout, llr_stride, residual_cfo_hz ? daux : nullptr, st);
                         });
}

}  // extern "C"
