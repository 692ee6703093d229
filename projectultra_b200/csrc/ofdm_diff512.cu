// projectultra_b200/csrc/ofdm_diff512.cu — persistent, TMA-staged, packed-fp32 receive kernel for the 512-FFT
// differential no-pilot OFDM modes with zero CFO (the BASELINE.json headline: M1 512-FFT DQPSK R1/2).
//
// Same reference behaviour and the same arithmetic DAG as ofdm_diff.cu (OFDMDemodulator::processPresynced,
// src/ofdm/demodulator.cpp:854-985: toBaseband channel_equalizer.cpp:19-57 -> extractSymbol + radix-2 FFT :59-71,
// src/dsp/fft.cpp:89-121 -> H from the last LTS symbol :179-185 -> ZF equalise :747-770 -> demapD*PSK
// src/ofdm/soft_demap.hpp:173-237); every butterfly is still the reference's unfused (t = w*b, a+t, a-t) in the
// reference's stage order, so bins / H / equalised symbols / LLRs stay bit-identical.  What changes is the machine
// mapping, chosen from the r03 profile (ofdm_diff_kernel: 78 % issue-slot utilisation, 56 % of all instructions are
// scalar FMUL/FADD of the butterflies, 28 % of stall samples wait on the first LDG of a symbol):
//   * TWO SYMBOLS PER WARP, PACKED: each 64-bit register pair holds the same quantity of symbols s and s+1, and the
//     butterflies are issued as sm_100 packed-fp32 instructions (FFMA2 / FADD2: two IEEE-rn fp32 operations per
//     issue slot; same FP-pipe throughput as scalar, half the issue slots).  ptxas contracts `mul.rn.f32x2` +
//     `add.rn.f32x2` into one FFMA2 even with --fmad=false (measured: tools/ubench/f32x2_bench.cu), which would
//     break bit-exactness, so every product is written as fma(a, b, Z) with Z = (-0, -0) passed as a KERNEL ARGUMENT:
//     a*b + (-0) is exactly RN(a*b) including the sign of a zero product, and ptxas cannot fold an unknown addend.
//   * butterflies whose twiddle is tw[0] = (1, -0) skip the multiplication: (1, -0) * b == b for every finite b up
//     to the sign of a zero component, and a zero whose sign could differ can only reach a bin when all 512 samples
//     of the symbol are zero (any other zero is produced by a cancellation x - x = +0 in both evaluations); then the
//     carrier is below the demapper's weak-signal gate (soft_demap.hpp:178,199,224) and its LLRs are 0 either way.
//   * PERSISTENT CTAs (2 per SM) loop over frames; the samples of the next frame are staged into shared memory by
//     cp.async.bulk (TMA, one copy per symbol, mbarrier complete_tx) while the current frame is computed, so no warp
//     ever waits on HBM latency and sample loads are conflict-free LDS.
//   * one 128-bit shared-memory transpose per element pair (re_s, re_s1, im_s, im_s1) between the two FFT passes.
// HBM traffic: every sample of the symbols that are used is read exactly once, LLRs are written once.
#include <cfloat>

#include "ofdm_dev.cuh"
#include "ofdm_diff_demap.cuh"
#include "pu_internal.h"

namespace pu {

typedef unsigned long long u64;

constexpr int kP512MaxWarps = 6;        // symbol pairs in flight per CTA (M1 DQPSK: 12 symbols = 6 pairs)
constexpr int kP512MaxSym = 40;
constexpr int kP512Buf = 512 + 32;      // float4 per warp: element p lives at p + (p >> 4)
constexpr size_t kP512SmemMax = 227 * 1024;

struct P512Tw { u64 re[8], im[8]; };    // pass-A twiddles tw[32 m] as broadcast pairs (w.x, w.x), (w.y, w.y)

struct C2 { u64 re, im; };              // one complex quantity of symbols (s, s+1): re = (re_s, re_s1), im likewise

__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// t = w * b as GCC lowers std::complex<float> multiplication (ofdm_dev.cuh: cmul), for two symbols at once
__device__ __forceinline__ C2 cmul2(u64 wre, u64 wim, C2 b, u64 Z) {
    C2 t;
    t.re = sub2(fma2(wre, b.re, Z), fma2(wim, b.im, Z));
    t.im = add2(fma2(wre, b.im, Z), fma2(wim, b.re, Z));
    return t;
}
__device__ __forceinline__ void bfly2(C2& a, C2& b, u64 wre, u64 wim, u64 Z) {   // fft.cpp:108-110
    const C2 t = cmul2(wre, wim, b, Z);
    b.re = sub2(a.re, t.re); b.im = sub2(a.im, t.im);
    a.re = add2(a.re, t.re); a.im = add2(a.im, t.im);
}
__device__ __forceinline__ void bfly2_w0(C2& a, C2& b) {                         // twiddle tw[0] = (1, -0): t == b
    const C2 t = b;
    b.re = sub2(a.re, t.re); b.im = sub2(a.im, t.im);
    a.re = add2(a.re, t.re); a.im = add2(a.im, t.im);
}
__device__ __forceinline__ C2 bfly2_lo(C2 a, C2 b, u64 wre, u64 wim, u64 Z) {
    const C2 t = cmul2(wre, wim, b, Z);
    C2 r; r.re = add2(a.re, t.re); r.im = add2(a.im, t.im); return r;
}
__device__ __forceinline__ C2 bfly2_hi(C2 a, C2 b, u64 wre, u64 wim, u64 Z) {
    const C2 t = cmul2(wre, wim, b, Z);
    C2 r; r.re = sub2(a.re, t.re); r.im = sub2(a.im, t.im); return r;
}

// ---- mbarrier / bulk-copy wrappers (PTX ISA 8.x, sm_90+) ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(u64* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, uint32_t bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int kP512Carr = 32;           // data carriers per frame are < 32 (ofdm_diff512_supported)
constexpr int kP512Groups = 2;          // frames in flight per CTA (each on its own W warps, named barrier and mbarriers)

struct P512Smem {     // offsets (bytes) into dynamic shared memory, computed identically on host and device
    size_t nco, tb, grp0, grp_stride;                       // CTA-wide: NCO slices, transpose buffers; then one block per group
    size_t xs, F, Hs, slow, hp, nv, habs, misc;             // offsets inside a group block
    size_t total;
};
__host__ __device__ inline P512Smem p512_layout(int n_symbols, int first, int sym_len, int nd, int warps) {
    P512Smem L;
    size_t o = 0;
    auto take = [&o](size_t bytes) { const size_t at = o; o += (bytes + 15) & ~size_t(15); return at; };
    const int npairs = (n_symbols - first + 1) / 2;
    L.nco = take(static_cast<size_t>(npairs) * 512 * sizeof(float4));
    L.tb = take(static_cast<size_t>(kP512Groups) * warps * kP512Buf * sizeof(float4));
    L.grp0 = o;
    o = 0;
    L.xs = take(static_cast<size_t>(n_symbols - first) * 512 * sizeof(float));             // samples after the CP only
    L.F = take(2 * static_cast<size_t>(n_symbols) * nd * sizeof(float2));     // [frame parity][symbol][carrier]
    L.Hs = take(2 * kP512Carr * sizeof(float2));                               // [frame parity][carrier]
    L.slow = take(static_cast<size_t>(n_symbols) * nd * sizeof(int));          // items (symbol, carrier) for the exact demapper
    L.hp = take(2 * kP512Carr * sizeof(float));
    L.nv = take(2 * kP512Carr * sizeof(float));
    L.habs = take(2 * kP512Carr * sizeof(float));
    L.misc = take(64);    // [0] full mbarrier, [8] empty mbarrier, [16] slow_count[2]
    L.grp_stride = (o + 127) & ~size_t(127);
    L.total = L.grp0 + kP512Groups * L.grp_stride;
    return L;
}

__global__ void __launch_bounds__(kP512Groups * kP512MaxWarps * 32, 1) ofdm_diff512_kernel(
    OfdmDev d, P512Tw twa, u64 Z, const float* __restrict__ samples, size_t frame_stride, size_t B, int n_symbols, int training,
    float* __restrict__ llr_out, size_t llr_stride, int llr_limit, float* __restrict__ snr_db_out, float* __restrict__ final_cfo_out) {
    constexpr int EPL = 16;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = (blockDim.x >> 5) / kP512Groups;               // warps per group
    const int grp = (threadIdx.x >> 5) / W;                      // which of the CTA's frames-in-flight this thread works on
    const int tid = threadIdx.x - grp * W * 32, lane = tid & 31, warp = tid >> 5, T = W * 32;
    const int nd = d.n_data;
    const int first = training > 0 ? training - 1 : 0;   // data H uses the LAST training symbol only (:179-185)
    const int n_proc = n_symbols - first;
    const int npairs = (n_proc + 1) >> 1;
    const P512Smem L = p512_layout(n_symbols, first, d.sym_len, nd, W);
    float4* nco_s = reinterpret_cast<float4*>(smem_raw + L.nco);                       // [npairs][16][32]
    float4* tb = reinterpret_cast<float4*>(smem_raw + L.tb) + (grp * W + warp) * kP512Buf;
    unsigned char* gb = smem_raw + L.grp0 + grp * L.grp_stride;
    float* xs = reinterpret_cast<float*>(gb + L.xs);
    float2* F = reinterpret_cast<float2*>(gb + L.F);
    float2* Hs = reinterpret_cast<float2*>(gb + L.Hs);
    int* slow_item = reinterpret_cast<int*>(gb + L.slow);
    float* hp_s = reinterpret_cast<float*>(gb + L.hp);
    float* nv_s = reinterpret_cast<float*>(gb + L.nv);
    float* habs = reinterpret_cast<float*>(gb + L.habs);
    u64* full_bar = reinterpret_cast<u64*>(gb + L.misc);
    u64* empty_bar = full_bar + 1;
    int* slow_count = reinterpret_cast<int*>(full_bar + 2);
    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(T) : "memory"); };

    const int c = lane & 15, b8 = lane >> 4;
    const int nlo = nd / 2, nhi = nd - nlo;
    const uint32_t sym_bytes = 512 * sizeof(float);     // the samples after the cyclic prefix are all that is staged

    if (tid == 0) {
        mbar_init(full_bar, 1);
        mbar_init(empty_bar, W);
        slow_count[0] = slow_count[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // NCO slices of the symbol pairs (loop invariant: resident in shared memory for the life of the CTA)
    for (int i = threadIdx.x; i < npairs * 512; i += blockDim.x) nco_s[i] = __ldg(&d.nco2[static_cast<size_t>(first + 2 * (i >> 9)) * 512 + (i & 511)]);
    __syncthreads();
    const size_t frame0 = static_cast<size_t>(blockIdx.x) * kP512Groups + grp, frame_step = static_cast<size_t>(gridDim.x) * kP512Groups;
    // stage the first frame of this group
    if (warp == 0 && frame0 < B) {
        if (lane == 0) mbar_expect_tx(full_bar, sym_bytes * n_proc);
        __syncwarp();
        const float* src = samples + frame0 * frame_stride + static_cast<size_t>(first) * d.sym_len + d.cp;
        for (int sy = lane; sy < n_proc; sy += 32) bulk_g2s(xs + sy * 512, src + static_cast<size_t>(sy) * d.sym_len, sym_bytes, full_bar);
    }

    // ---- per-lane twiddles of pass B as broadcast pairs (loop invariant).  Stage q pairs (j, j + 2^q); low outputs
    //      use k = c, high outputs k = c + 16 (2^q - 1); table index k << (4 - q).  Stage 9: k = c or c + 240.
    u64 wl_re[4], wl_im[4], wh_re[4], wh_im[4], wlast_re, wlast_im;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int sh = 4 - q;
        const float2 a = __ldg(&d.twiddle[c << sh]);
        const float2 b = __ldg(&d.twiddle[(c + 16 * ((1 << q) - 1)) << sh]);
        wl_re[q] = pk(a.x, a.x); wl_im[q] = pk(a.y, a.y);
        wh_re[q] = pk(b.x, b.x); wh_im[q] = pk(b.y, b.y);
    }
    {
        const float2 a = __ldg(&d.twiddle[b8 ? (c + 240) : c]);
        wlast_re = pk(a.x, a.x); wlast_im = pk(a.y, a.y);
    }
    const int rlane = static_cast<int>(__brev(static_cast<unsigned>(lane)) >> 27);   // brev5(lane)

    uint32_t it = 0;
    for (size_t frame = frame0; frame < B; frame += frame_step, ++it) {
        mbar_wait(full_bar, it & 1);
        for (int pr = warp; pr < npairs; pr += W) {
            const int s = first + 2 * pr;
            const bool have2 = (s + 1) < n_symbols;
            const float* x0 = xs + (s - first) * 512;
            const float* x1 = have2 ? x0 + 512 : x0;            // odd tail: the upper half recomputes symbol s (never stored)
            const float4* nc = nco_s + pr * 512 + lane;                   // [q][lane]: entry of sample brev5(lane) + 32 brev4(q)
            // ---- pass A: lane g owns bit-reversed positions 16 g .. 16 g + 15 = samples brev5(g) + 32 brev4(q)
            C2 v[EPL];
#pragma unroll
            for (int q = 0; q < EPL; ++q) {
                const int brq = static_cast<int>(__brev(static_cast<unsigned>(q)) >> 28);
                const int n = rlane + 32 * brq;
                const u64 xv = pk(x0[n], x1[n]);
                const float4 o = nc[32 * q];                          // (cos_s, cos_s1, -sin_s, -sin_s1)
                v[q].re = fma2(pk(o.x, o.y), xv, Z);                  // samples[i] * conj(osc) (channel_equalizer.cpp:36)
                v[q].im = fma2(pk(o.z, o.w), xv, Z);
            }
            if (pr + W >= npairs) {       // this warp has taken its last samples of the frame out of shared memory
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_bar);
                const size_t next = frame + frame_step;
                if (warp == 0 && next < B) {                          // refill the staging buffer for the next frame
                    mbar_wait(empty_bar, it & 1);
                    if (lane == 0) mbar_expect_tx(full_bar, sym_bytes * n_proc);
                    __syncwarp();
                    const float* src = samples + next * frame_stride + static_cast<size_t>(first) * d.sym_len + d.cp;
                    for (int sy = lane; sy < n_proc; sy += 32)
                        bulk_g2s(xs + sy * 512, src + static_cast<size_t>(sy) * d.sym_len, sym_bytes, full_bar);
                }
            }
#pragma unroll
            for (int t = 1; t <= 4; ++t) {
                const int half = 1 << (t - 1);
#pragma unroll
                for (int p2 = 0; p2 < EPL / 2; ++p2) {
                    const int kq = p2 & (half - 1);
                    const int a = ((p2 >> (t - 1)) << t) | kq;
                    const int m = kq << (4 - t);
                    if (m == 0) bfly2_w0(v[a], v[a + half]);
                    else bfly2(v[a], v[a + half], twa.re[m], twa.im[m], Z);
                }
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < EPL; ++q) {
                float a0, a1, b0, b1;
                upk(v[q].re, a0, a1); upk(v[q].im, b0, b1);
                tb[17 * lane + q] = make_float4(a0, a1, b0, b1);       // p = 16 lane + q at p + (p >> 4)
            }
            __syncwarp();
            // ---- pass B: lane (b8, c) owns p = c + 16 j + 256 b8
#pragma unroll
            for (int j = 0; j < EPL; ++j) {
                const float4 e = tb[c + 17 * j + 272 * b8];
                v[j].re = pk(e.x, e.y); v[j].im = pk(e.z, e.w);
            }
#pragma unroll
            for (int j = 0; j < EPL; j += 2) bfly2(v[j], v[j + 1], wl_re[0], wl_im[0], Z);
#pragma unroll
            for (int q = 1; q < 4; ++q) {
                const int step = 1 << (q + 1), h = 1 << q;
#pragma unroll
                for (int j = 0; j < EPL; j += step) {
                    v[j] = bfly2_lo(v[j], v[j + h], wl_re[q], wl_im[q], Z);
                    v[j + step - 1] = bfly2_hi(v[j + step - 1 - h], v[j + step - 1], wh_re[q], wh_im[q], Z);
                }
            }
            // stage 9 pairs lane (0,c) [A] with lane (1,c) [B]: bin c = A0 + w B0 on lane (0,c); bin 496+c = A15 - w B15 on lane (1,c)
            const C2 send = b8 ? v[0] : v[EPL - 1];
            C2 recv;
            recv.re = __shfl_xor_sync(0xffffffffu, send.re, 16);
            recv.im = __shfl_xor_sync(0xffffffffu, send.im, 16);
            const C2 bin = b8 ? bfly2_hi(recv, v[EPL - 1], wlast_re, wlast_im, Z) : bfly2_lo(v[0], recv, wlast_re, wlast_im, Z);
            int idx = -1;
            if (b8 == 0) { if (c >= 1 && c <= nhi) idx = nlo + c - 1; }
            else { const int cc = 16 - c; if (c >= 1 && cc <= nlo) idx = nlo - cc; }
            float2* Fb = F + (it & 1) * n_symbols * nd;                 // bins of this frame (double-buffered by frame parity)
            if (idx >= 0) {
                float r0, r1, i0, i1;
                upk(bin.re, r0, r1); upk(bin.im, i0, i1);
                Fb[s * nd + idx] = make_float2(r0, i0);
                if (have2) Fb[(s + 1) * nd + idx] = make_float2(r1, i1);
                if (pr == 0) {
                    // estimateChannelFromLTS for data carriers (channel_equalizer.cpp:141,179-185) + the per-carrier constants of
                    // equalize, straight from the registers of the warp that transformed the last LTS symbol (s == training - 1)
                    const float2 h = training > 0 ? cdiv(make_float2(r0, i0), __ldg(&d.zc[idx])) : make_float2(1.0f, 0.0f);
                    const float hp = cnorm(h);
                    const int o = (it & 1) * kP512Carr + idx;
                    Hs[o] = h;
                    hp_s[o] = hp;
                    nv_s[o] = (hp > 1e-6f) ? clampf(1e-6f, 100.0f, __fdiv_rn(0.1f, hp)) : 100.0f;   // noise_variance stays 0.1 (:762-768)
                    if (snr_db_out) habs[o] = cabs_ref(h);
                }
            }
        }
        group_sync();                                                  // B1: bins of all symbols + H are in shared memory
        if (tid == 0) slow_count[(it + 1) & 1] = 0;                    // every reader of that counter (frame it - 1) is past B1

        // ---- equalize (:747-770, ZF with pilot_phase_correction == (1,0) and timing_offset == 0) + demodulateSymbol
        //      (demodulator.cpp:279-316) + soft_demap.hpp, fused: lane = carrier, each warp walks a contiguous run of data symbols
        //      and keeps the previous equalised symbol in registers (the run's first predecessor is re-equalised, not exchanged).
        const float2* Fb = F + (it & 1) * n_symbols * nd;
        const int nds = n_symbols - training;
        float* out = llr_out + frame * llr_stride;
        const int bps = d.bps, mod = d.mod;
        int* scount = slow_count + (it & 1);
        if (lane < nd && nds > 0) {
            const int o = (it & 1) * kP512Carr + lane;
            const float2 h = Hs[o];
            const float hp = hp_s[o];
            const float nv = __fmul_rn(nv_s[o], d.ce_margin);
            const float inv_nv = __frcp_rn(nv);
            const float2 one = make_float2(1.0f, 0.0f);
            auto equalize = [&](int sd) {
                const float2 rx = Fb[(training + sd) * nd + lane];
                if (hp > 1e-6f) return cmul(cmul(cdivs(cmul(rx, cconj(h)), hp), one), one);   // :761
                return cmul(cmul(rx, one), one);
            };
            const int ch = (nds + W - 1) / W;
            const int sd0 = warp * ch, sd1 = min(nds, sd0 + ch);
            float2 prev = make_float2(1.0f, 0.0f);                     // differential reference (1,0) (:251-255)
            if (sd0 > 0 && sd0 < sd1) prev = equalize(sd0 - 1);
            for (int sd = sd0; sd < sd1; ++sd) {
                const float2 sym = equalize(sd);
                const int item = sd * nd + lane;
                float l[3];
                if (demap_saturated_fast(mod, cmul(sym, cconj(prev)), inv_nv, l)) {
                    store_llrs(out, item * bps, bps, l, llr_limit, d.llr_perm, d.perm_len);
                } else {
                    slow_item[atomicAdd(scount, 1)] = item;
                }
                prev = sym;
            }
        }
        group_sync();                                                  // B2: the list of carriers that need the exact demapper is complete
        const int n_slow = *scount;
        for (int k = tid; k < n_slow; k += T) {     // dense: the first n_slow threads of the group; re-equalises its two symbols
            const int item = slow_item[k];
            const int sd = item / nd, i = item - sd * nd;
            const int o = (it & 1) * kP512Carr + i;
            const float2 h = Hs[o];
            const float hp = hp_s[o];
            const float2 one = make_float2(1.0f, 0.0f);
            auto equalize = [&](int sdx) {
                const float2 rx = Fb[(training + sdx) * nd + i];
                if (hp > 1e-6f) return cmul(cmul(cdivs(cmul(rx, cconj(h)), hp), one), one);
                return cmul(cmul(rx, one), one);
            };
            const float2 sym = equalize(sd);
            const float2 prev = sd > 0 ? equalize(sd - 1) : make_float2(1.0f, 0.0f);
            float l[3];
            demap_exact(mod, sym, prev, sd == 0, __fmul_rn(nv_s[o], d.ce_margin), l);
            store_llrs(out, item * bps, bps, l, llr_limit, d.llr_perm, d.perm_len);
        }
        if (tid == 0) {
            if (snr_db_out) {   // reporting-only SNR estimate of estimateChannelFromLTS (:208-225), getEstimatedSNR (demodulator.cpp:797-799)
                float snr_lin = 1.0f;
                if (training > 0) {
                    float sum = 0.0f;
                    for (int i = 0; i < nd; ++i) sum = __fadd_rn(sum, habs[(it & 1) * kP512Carr + i]);
                    const float avg = __fdiv_rn(sum, static_cast<float>(nd));
                    if (avg > 1e-6f) snr_lin = clampf(0.1f, 10000.0f, __fdiv_rn(__fmul_rn(avg, avg), 0.1f));
                }
                snr_db_out[frame] = 10.0f * log10f(snr_lin);
            }
            if (final_cfo_out) final_cfo_out[frame] = 0.0f;
        }
        // no barrier here: F / H are double-buffered by frame parity, the slow list is only appended to after the next B1
    }
}

static int p512_warps(int n_symbols, int training) {
    const int first = training > 0 ? training - 1 : 0;
    const int npairs = (n_symbols - first + 1) / 2;
    return npairs < kP512MaxWarps ? npairs : kP512MaxWarps;
}

// Returns true when the configuration / call is one this kernel covers (a superset of ofdm_diff_supported's
// conditions: 512-FFT, 16-byte aligned rows so that the bulk copies are legal, the packed NCO table present).
bool ofdm_diff512_supported(const OfdmDev& d, int n_symbols, int training, const float* samples, size_t frame_stride) {
    const bool differential = d.mod == PU_MOD_DBPSK || d.mod == PU_MOD_DQPSK || d.mod == PU_MOD_D8PSK;
    if (!differential || d.n_pilot != 0 || d.nfft != 512 || !d.nco2) return false;
    if (n_symbols > kP512MaxSym || n_symbols < 1 || training < 0 || training > n_symbols) return false;
    if ((d.sym_len & 3) || (d.cp & 3) || (frame_stride & 3) || (reinterpret_cast<uintptr_t>(samples) & 15)) return false;
    const int first = training > 0 ? training - 1 : 0;
    if (n_symbols - first < 1 || n_symbols - first > 32) return false;
    const int nlo = d.n_data / 2, nhi = d.n_data - nlo;
    if (!(nlo < 16 && nhi < 16)) return false;
    const P512Smem L = p512_layout(n_symbols, first, d.sym_len, d.n_data, p512_warps(n_symbols, training));
    return L.total <= kP512SmemMax;
}

cudaError_t ofdm_diff512_launch(const OfdmDev& d, const float2* host_twiddle, const float* samples, size_t B, size_t frame_stride,
                                int n_symbols, int training, float* llr, size_t llr_stride, int llr_limit, float* snr_db,
                                float* final_cfo, int sm_count, cudaStream_t st) {
    P512Tw twa;
    for (int m = 0; m < 8; ++m) {
        const float2 w = host_twiddle[32 * m];
        uint32_t xr, xi;
        memcpy(&xr, &w.x, 4); memcpy(&xi, &w.y, 4);
        twa.re[m] = (static_cast<u64>(xr) << 32) | xr;
        twa.im[m] = (static_cast<u64>(xi) << 32) | xi;
    }
    const int first = training > 0 ? training - 1 : 0;
    const int warps = p512_warps(n_symbols, training);
    const P512Smem L = p512_layout(n_symbols, first, d.sym_len, d.n_data, warps);
    static size_t attr = 0;
    if (L.total > attr) {
        const cudaError_t e = cudaFuncSetAttribute(ofdm_diff512_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kP512SmemMax));
        if (e != cudaSuccess) return e;
        attr = kP512SmemMax;
    }
    const size_t max_ctas = static_cast<size_t>(sm_count > 0 ? sm_count : 148);      // persistent: one CTA per SM
    const size_t want_ctas = (B + kP512Groups - 1) / kP512Groups;
    const unsigned grid = static_cast<unsigned>(want_ctas < max_ctas ? want_ctas : max_ctas);
    const u64 Z = 0x8000000080000000ull;    // (-0.0f, -0.0f): see the header comment
    ofdm_diff512_kernel<<<grid, kP512Groups * warps * 32, L.total, st>>>(d, twa, Z, samples, frame_stride, B, n_symbols, training, llr, llr_stride, llr_limit,
                                                          snr_db, final_cfo);
    return cudaGetLastError();
}

}  // namespace pu
