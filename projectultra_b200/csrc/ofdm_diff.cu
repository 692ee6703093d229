// projectultra_b200/csrc/ofdm_diff.cu — warp-FFT receive kernel for the differential no-pilot OFDM modes
// (DBPSK / DQPSK / D8PSK, use_pilots = false) with zero CFO: the Monte-Carlo headline path
// (512-FFT DQPSK R1/2 of BASELINE.json, 1024-FFT NVIS DQPSK/D8PSK).
//
// Reference behaviour: OFDMDemodulator::processPresynced (src/ofdm/demodulator.cpp:854-985) after reset() +
// setFrequencyOffset(0) on a fresh object, differential branch of SURVEY App. E:
//   toBaseband (channel_equalizer.cpp:19-57, CFO rotator off because |cfo| <= 0.01) -> extractSymbol + radix-2 FFT
//   (:59-71, src/dsp/fft.cpp:89-121) -> H = F(last LTS)/zc (:179-185) -> ZF equalise (:747-770) -> demapD*PSK
//   (src/ofdm/soft_demap.hpp:173-237).  In these modes no tracker changes state, so every symbol
//   of a frame is independent up to its differential reference: the symbols of a frame are spread over the warps of
//   a CTA instead of being walked in order.
//
// Numerics: identical to ofdm_demod.cu -- every butterfly is the reference's unfused (w*b, a+t, a-t) in the
// reference's stage order, so bins, H, equalised symbols and LLRs are bit-identical to the general kernel and the
// oracle.  Only the SCHEDULE differs:
//   * one warp computes one symbol's FFT.  Pass A: each lane loads N/32 samples (coalesced 128-byte rows in
//     bit-reversed row order), mixes them with the NCO table and runs the first log2(N/32) radix-2 stages in
//     registers.  One padded shared-memory transpose.  Pass B: the remaining stages in registers, PRUNED to the
//     butterflies that feed the used carriers (|carrier| < N/32 around DC): stage q needs 2^(LB-q) of its N/32/2
//     butterflies per lane, each producing one output.  For N = 512 the last stage pairs lanes l and l^16 by shuffle.
//   * equalisation / demapping run over (symbol, carrier) pairs on all threads of the CTA.
// HBM traffic: every sample of the symbols that are used is read exactly once, LLRs are written once.
#include <cfloat>

#include "ofdm_dev.cuh"
#include "ofdm_diff_demap.cuh"
#include "pu_async.cuh"
#include "pu_internal.h"

namespace pu {

constexpr int kDiffWarps = 4;          // warps (= symbols in flight) per CTA; one frame per CTA
constexpr int kDiffMaxSym = 40;        // symbols per frame supported by this kernel (M1 DBPSK needs 24)

struct DiffTw { float2 a[16]; };       // pass-A twiddles tw[32*m] (compile-time register indices, constant bank)

__device__ __forceinline__ void bfly(float2& a, float2& b, float2 w) {
    const float2 t = cmul(w, b);       // Complex t = w * data[i + k + half]   (fft.cpp:108)
    b = csub(a, t);
    a = cadd(a, t);
}
__device__ __forceinline__ float2 bfly_lo(float2 a, float2 b, float2 w) { return cadd(a, cmul(w, b)); }
__device__ __forceinline__ float2 bfly_hi(float2 a, float2 b, float2 w) { return csub(a, cmul(w, b)); }

template <int NFFT>
struct DiffGeom {
    static constexpr int LOG2N = (NFFT == 512) ? 9 : 10;
    static constexpr int EPL = NFFT / 32;                    // elements per lane
    static constexpr int LA = (NFFT == 512) ? 4 : 5;         // stages of pass A
    static constexpr int LB = LA;                            // in-lane stages of pass B (J = EPL elements)
    static constexpr int CW = (NFFT == 512) ? 16 : 32;       // p = c + CW*j (+256*b8 for N = 512)
    static constexpr int PADSH = (NFFT == 512) ? 4 : 5;      // one float2 of padding per 2^PADSH
    static constexpr int BUF = NFFT + (NFFT >> PADSH);       // float2 per warp
};

template <int NFFT>
__global__ void __launch_bounds__(kDiffWarps * 32) ofdm_diff_kernel(
    OfdmDev d, DiffTw twa, const float* __restrict__ samples, size_t frame_stride, int n_symbols, int training,
    float* __restrict__ llr_out, size_t llr_stride, int llr_limit, float* __restrict__ snr_db_out,
    float* __restrict__ final_cfo_out) {
    using G = DiffGeom<NFFT>;
    constexpr int LOG2N = G::LOG2N, EPL = G::EPL, LA = G::LA, LB = G::LB, CW = G::CW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* fftbuf = reinterpret_cast<float2*>(smem_raw);                 // [kDiffWarps][BUF]
    float2* F = fftbuf + kDiffWarps * G::BUF;                             // [n_symbols][nd] bins, then equalised symbols
    const int nd = d.n_data;
    float2* Hs = F + n_symbols * nd;                                      // [nd] channel estimate
    float* eabs = reinterpret_cast<float*>(Hs + kMaxCarr);                // [n_symbols][nd] scratch: list of carriers for the exact demapper
    float* hp_s = eabs + n_symbols * nd;                                  // [nd] |H|^2
    float* nv_s = hp_s + kMaxCarr;                                        // [nd] carrier noise variance
    float* habs = nv_s + kMaxCarr;                                        // [nd] |H| (SNR report only)
    int* slow_count = reinterpret_cast<int*>(habs + kMaxCarr);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t frame = blockIdx.x;
    const float* x = samples + frame * frame_stride;
    float2* buf = fftbuf + warp * G::BUF;
    const int c = lane & (CW - 1);
    const int b8 = (NFFT == 512) ? (lane >> 4) : 0;
    const int nlo = nd / 2;               // carriers -nlo..-1 are data carriers 0..nlo-1; +1..+(nd-nlo) follow (setupCarriers)
    const int nhi = nd - nlo;

    // ---- per-lane twiddles of pass B (loop invariant): stage q pairs (j, j + 2^q); low outputs use k = c,
    //      high outputs k = c + CW*(2^q - 1); table index k << (LOG2N - (LA + 1 + q))
    float2 wl[LB], wh[LB];
#pragma unroll
    for (int q = 0; q < LB; ++q) {
        const int sh = LOG2N - (LA + 1 + q);
        wl[q] = __ldg(&d.twiddle[c << sh]);
        wh[q] = __ldg(&d.twiddle[(c + CW * ((1 << q) - 1)) << sh]);
    }
    float2 wlast = make_float2(0.0f, 0.0f);
    if (NFFT == 512) wlast = __ldg(&d.twiddle[b8 ? (c + 240) : c]);   // stage 9: k = c (bin c) or c + 16*15 (bin 496 + c)

    const int first = training > 0 ? training - 1 : 0;   // data H uses the LAST training symbol only (:179-185)
    const int rlane = static_cast<int>(__brev(static_cast<unsigned>(lane)) >> 27);   // brev5(lane)
    for (int s = first + warp; s < n_symbols; s += kDiffWarps) {
        const float* xs = x + static_cast<size_t>(s) * d.sym_len + d.cp;
        const float2* nco = d.nco + static_cast<size_t>(s) * d.sym_len + d.cp;
        // ---- pass A: lane g owns bit-reversed positions EPL*g .. EPL*g + EPL-1 = samples brev5(g) + 32*brev_LA(q)
        float2 v[EPL];
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
            const int brq = static_cast<int>(__brev(static_cast<unsigned>(q)) >> (32 - LA));
            const int n = rlane + 32 * brq;
            const float xv = __ldg(&xs[n]);
            const float2 o = __ldg(&nco[n]);
            v[q] = make_float2(__fmul_rn(o.x, xv), __fmul_rn(-o.y, xv));   // samples[i] * conj(osc) (channel_equalizer.cpp:36)
        }
#pragma unroll
        for (int t = 1; t <= LA; ++t) {
            const int half = 1 << (t - 1);
#pragma unroll
            for (int pr = 0; pr < EPL / 2; ++pr) {
                const int kq = pr & (half - 1);
                const int a = ((pr >> (t - 1)) << t) | kq;
                bfly(v[a], v[a + half], twa.a[kq << (LA - t)]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
            const int p = EPL * lane + q;
            buf[p + (p >> G::PADSH)] = v[q];
        }
        __syncwarp();
        // ---- pass B: lane (b8, c) owns p = c + CW*j (+ 256*b8), j < EPL
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            const int p = c + CW * j + 256 * b8;
            v[j] = buf[p + (p >> G::PADSH)];
        }
        // stage q = 0: all butterflies (j, j+1), k = c
#pragma unroll
        for (int j = 0; j < EPL; j += 2) bfly(v[j], v[j + 1], wl[0]);
        // stages q >= 1: only the outputs that feed j = 0 (carriers > 0) and j = EPL-1 (carriers < 0)
#pragma unroll
        for (int q = 1; q < LB; ++q) {
            const int step = 1 << (q + 1), h = 1 << q;
#pragma unroll
            for (int j = 0; j < EPL; j += step) {
                v[j] = bfly_lo(v[j], v[j + h], wl[q]);                            // pair (j, j+h), low output at j
                v[j + step - 1] = bfly_hi(v[j + step - 1 - h], v[j + step - 1], wh[q]);   // pair (j+step-1-h, j+step-1), high output
            }
        }
        float2 pos_bin, neg_bin;      // carrier +c and carrier c - CW (valid for c >= 1)
        if (NFFT == 512) {
            // stage 9 pairs lane (0,c) [A] with lane (1,c) [B]: bin c = A0 + w B0 on lane (0,c); bin 496+c = A15 - w B15 on lane (1,c)
            const float2 send = b8 ? v[0] : v[EPL - 1];
            float2 recv;
            recv.x = __shfl_xor_sync(0xffffffffu, send.x, 16);
            recv.y = __shfl_xor_sync(0xffffffffu, send.y, 16);
            pos_bin = bfly_lo(v[0], recv, wlast);           // meaningful on b8 == 0
            neg_bin = bfly_hi(recv, v[EPL - 1], wlast);     // meaningful on b8 == 1
        } else {
            pos_bin = v[0];
            neg_bin = v[EPL - 1];
        }
        float2* Fs = F + s * nd;
        if (NFFT == 512) {
            if (b8 == 0) { if (c >= 1 && c <= nhi) Fs[nlo + c - 1] = pos_bin; }
            else { const int cc = CW - c; if (c >= 1 && cc <= nlo) Fs[nlo - cc] = neg_bin; }
        } else {
            if (c >= 1 && c <= nhi) Fs[nlo + c - 1] = pos_bin;
            const int cc = CW - c;
            if (c >= 1 && cc <= nlo) Fs[nlo - cc] = neg_bin;
        }
    }
    __syncthreads();

    // ---- estimateChannelFromLTS for data carriers (channel_equalizer.cpp:141,179-185) and the per-carrier constants of equalize
    for (int i = tid; i < nd; i += blockDim.x) {
        const float2 h = training > 0 ? cdiv(F[(training - 1) * nd + i], d.zc[i]) : make_float2(1.0f, 0.0f);
        const float hp = cnorm(h);
        Hs[i] = h;
        hp_s[i] = hp;
        nv_s[i] = (hp > 1e-6f) ? clampf(1e-6f, 100.0f, __fdiv_rn(0.1f, hp)) : 100.0f;   // noise_variance stays 0.1 (:762-768)
        habs[i] = cabs_ref(h);
    }
    __syncthreads();

    // ---- equalize (:747-770): ZF with pilot_phase_correction == (1,0) and timing_offset == 0
    const int nds = n_symbols - training;
    const int items = nds > 0 ? nds * nd : 0;
    int* slow_list = reinterpret_cast<int*>(eabs);       // [items] carriers that need the exact demapper
    if (tid == 0) *slow_count = 0;
    for (int it = tid; it < items; it += blockDim.x) {
        const int sd = it / nd, i = it - sd * nd;
        float2* slot = F + (training + sd) * nd + i;
        const float2 rx = *slot, h = Hs[i];
        const float hp = hp_s[i];
        const float2 one = make_float2(1.0f, 0.0f);
        float2 e;
        if (hp > 1e-6f) e = cmul(cmul(cdivs(cmul(rx, cconj(h)), hp), one), one);   // :761
        else e = cmul(cmul(rx, one), one);
        *slot = e;
    }
    __syncthreads();

    // ---- demodulateSymbol (demodulator.cpp:279-316) + soft_demap.hpp: saturation filter first, exact path for the rest
    float* out = llr_out + frame * llr_stride;
    const int bps = d.bps, mod = d.mod;
    for (int it = tid; it < items; it += blockDim.x) {
        const int sd = it / nd, i = it - sd * nd;
        const int s = training + sd;
        const float2 sym = F[s * nd + i];
        const float2 prev = sd > 0 ? F[(s - 1) * nd + i] : make_float2(1.0f, 0.0f);   // differential reference (1,0) (:251-255)
        const float nv = __fmul_rn(nv_s[i], d.ce_margin);
        float l[3];
        if (demap_saturated(mod, cmul(sym, cconj(prev)), nv, l)) store_llrs(out, it * bps, bps, l, llr_limit, d.llr_perm, d.perm_len);
        else slow_list[atomicAdd(slow_count, 1)] = it;
    }
    __syncthreads();
    const int n_slow = *slow_count;
    for (int k = tid; k < n_slow; k += blockDim.x) {     // dense: the first n_slow threads of the CTA
        const int it = slow_list[k];
        const int sd = it / nd, i = it - sd * nd;
        const int s = training + sd;
        const float2 sym = F[s * nd + i];
        const float2 prev = sd > 0 ? F[(s - 1) * nd + i] : make_float2(1.0f, 0.0f);
        float l[3];
        demap_exact(mod, sym, prev, sd == 0, __fmul_rn(nv_s[i], d.ce_margin), l);
        store_llrs(out, it * bps, bps, l, llr_limit, d.llr_perm, d.perm_len);
    }
    if (tid == 0) {
        if (snr_db_out) {   // reporting-only SNR estimate of estimateChannelFromLTS (:208-225), getEstimatedSNR (demodulator.cpp:797-799)
            float snr_lin = 1.0f;
            if (training > 0) {
                float sum = 0.0f;
                for (int i = 0; i < nd; ++i) sum = __fadd_rn(sum, habs[i]);
                const float avg = __fdiv_rn(sum, static_cast<float>(nd));
                if (avg > 1e-6f) snr_lin = clampf(0.1f, 10000.0f, __fdiv_rn(__fmul_rn(avg, avg), 0.1f));
            }
            snr_db_out[frame] = 10.0f * log10f(snr_lin);
        }
        if (final_cfo_out) final_cfo_out[frame] = 0.0f;
    }
}

// Returns true when the configuration / call is one this kernel covers.
bool ofdm_diff_supported(const OfdmDev& d, int n_symbols, int training) {
    const bool differential = d.mod == PU_MOD_DBPSK || d.mod == PU_MOD_DQPSK || d.mod == PU_MOD_D8PSK;
    if (!differential || d.n_pilot != 0) return false;
    if (n_symbols > kDiffMaxSym || n_symbols < 1 || training < 0) return false;
    const int cw = d.nfft == 512 ? 16 : 32;
    const int nlo = d.n_data / 2, nhi = d.n_data - nlo;
    return nlo < cw && nhi < cw;   // every used carrier is one of the bins c, N - (cw - c) with 1 <= c < cw
}

size_t ofdm_diff_smem(const OfdmDev& d, int n_symbols) {
    const size_t buf = d.nfft == 512 ? DiffGeom<512>::BUF : DiffGeom<1024>::BUF;
    return sizeof(float2) * (kDiffWarps * buf + static_cast<size_t>(n_symbols) * d.n_data + kMaxCarr) +
           sizeof(float) * (static_cast<size_t>(n_symbols) * d.n_data + 3 * kMaxCarr + 4);
}

cudaError_t ofdm_diff_launch(const OfdmDev& d, const float2* host_twiddle, const float* samples, size_t B, size_t frame_stride,
                             int n_symbols, int training, float* llr, size_t llr_stride, int llr_limit, float* snr_db,
                             float* final_cfo, cudaStream_t st) {
    DiffTw twa;
    for (int m = 0; m < 16; ++m) twa.a[m] = (32 * m < d.nfft / 2) ? host_twiddle[32 * m] : make_float2(0.0f, 0.0f);
    const size_t smem = ofdm_diff_smem(d, n_symbols);
    const unsigned grid = static_cast<unsigned>(B);
    if (d.nfft == 512) {
        static std::atomic<uint64_t> attr{0};
        if (const cudaError_t e = smem_optin(attr, ofdm_diff_kernel<512>, 100 * 1024); e != cudaSuccess) return e;
        ofdm_diff_kernel<512><<<grid, kDiffWarps * 32, smem, st>>>(d, twa, samples, frame_stride, n_symbols, training, llr, llr_stride,
                                                                  llr_limit, snr_db, final_cfo);
    } else {
        static std::atomic<uint64_t> attr{0};
        if (const cudaError_t e = smem_optin(attr, ofdm_diff_kernel<1024>, 100 * 1024); e != cudaSuccess) return e;
        ofdm_diff_kernel<1024><<<grid, kDiffWarps * 32, smem, st>>>(d, twa, samples, frame_stride, n_symbols, training, llr, llr_stride,
                                                                   llr_limit, snr_db, final_cfo);
    }
    return cudaGetLastError();
}

}  // namespace pu
