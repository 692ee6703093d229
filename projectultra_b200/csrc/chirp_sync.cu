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
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pu_async.cuh"
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

// ---------------------------------------------------------------------------------------------------------------------------------
// Two-tier form of detectChirpTemplate.  What the reference computes at every one of the ~870 coarse positions -- three ordered
// 24 000-tap fp32 sums -- decides only (a) which position holds the first maximum and (b) the value there.  So the positions are
// RANKED with a cheap estimate and only the leaders are evaluated the reference's way:
//   tier 1  the window is low-passed (47 taps, 4 kHz) and decimated 6:1 -- the template occupies 300..2700 Hz, so
//           sum_i x[p+i] t[i] ~ 6 sum_j xf[p+6j] t[6j] -- and every coarse position (48 samples = 8 decimated ones apart) gets a
//           4 000-tap correlation from shared memory on the tensor cores (tf32, diagonal sums of a small GEMM); energies come from
//           48-sample partial sums.
//           Measured against the ordered sums: rms error 1-2 % of the correlation floor of a noise-only window.
//   tier 2  the 16 best-ranked positions are evaluated exactly (lane = position, its three sums three independent chains), the result is the first maximum among them, and the search ends when every unverified position
//           is out of reach:  estimate + 4 x (largest |exact - estimate| seen) < best exact value.  Otherwise the next 16 are
//           verified, down to all of them -- the result is then the brute-force one by construction.
// The fine search (+-48 positions, parabolic neighbours included) stays exact: 99 positions in four warps over one linear tile.
// tests/test_chirp_sync_gpu.py runs both forms on the same frames (PU_CHIRP_SEARCH=exact selects the brute-force kernel).
constexpr int kRankD = 6, kRankNT = 47, kRankC = 23;   // decimation, low-pass taps, centre tap
constexpr int kC2Threads = 256, kC2Warps = kC2Threads / 32;
constexpr int kC2N = 24000, kC2Nd = kC2N / kRankD;     // the 48 kHz chirp: 24 000 taps, 4 000 decimated
constexpr int kC2MaxPos = 3000;                         // coarse positions per window the shared-memory budget of one SM allows
constexpr int kC2TileIn = kC2Threads * kRankD + kRankNT - 1;
constexpr int kC2TilePad = (kC2TileIn + 2 + 3) & ~3;     // one input tile buffer (two of them: double-buffered)
// Fine stage: how far the correlation magnitude can rise between a grid point and a position <= 1.5 samples away (grid every 3 samples).
// The complex correlation is band-limited to the template's 300..2700 Hz, i.e. +-1200 Hz around 1500 Hz, so by Bernstein's inequality its
// magnitude changes by at most 2 pi 1200 / 48000 = 0.157 of its supremum per sample: a maximum M within 1.5 samples of a grid point implies
// an estimate of at least (1 - 0.236) M there.  (A 6-sample grid with the single-path main lobe's 1.15 missed 2 of 16 384 two-path frames:
// the paths' interference pattern has a 32-sample period.  tools/chirp_ab.py is that experiment.)
constexpr float kFineReach = 1.0f / (1.0f - 1.5f * 0.157f);
constexpr int kC2Cand = 16, kC2VTaps = 128;             // exact coarse stage: leaders per round, taps per staged tile
constexpr int kC2Row = kC2VTaps + 1;                    // odd row stride: lane = candidate reads conflict-free
constexpr int kC2VBuf = kC2Cand * kC2Row + 2 * kC2VTaps;   // floats per staging buffer: sample rows + template tile (cos, sin)
// Shared memory of one frame, carved from the dynamic allocation for the window's coarse-position budget `maxpos` and the number of
// positions ranked per pass `ppart` (host: chirp2_layout): the decimated window is the large item, so long windows are ranked in
// several passes over a buffer that holds one part (+ the 4 000-tap overhang) -- 65 KB at 66 000 samples in one pass, 73 KB at
// 86 600 in two: three frames per SM either way.
struct Chirp2Smem {
    float* xd;                        // tier 1: decimated part of the window, whole tiles.  tier 2: rows[nbuf][16][kC2Row] + template tiles.  fine: lin[6][272] + template tiles
    float* tile;                      // low-pass input tile
    float (*acc)[2];
    float* a;                         // estimate of the normalised correlation
    unsigned short* order;            // order[rank] = coarse index, best estimate first
    float* seg;                       // 48-sample energies -> exclusive prefix
    float* lp;                        // [48] low-pass taps
    float* exq; float* exe; float* ex;   // [128] each
    int* cand;                        // [32]
    float* best_c; int* best_p; int* go; float* best_se; int* nranked;
};
__host__ __device__ inline int chirp2_tiles(int ppart) {       // never less than the exact phases need: three staging buffers
    // (+ 288: the last band of the tensor-core estimates reads up to 8 (ppart + 533) + 7, feeding diagonals past the last position)
    const int t = (8 * ppart + kC2Nd + 288 + kC2Threads - 1) / kC2Threads, floor_t = (3 * kC2VBuf + kC2Threads - 1) / kC2Threads;
    return t > floor_t ? t : floor_t;
}
__host__ __device__ inline int chirp2_rank_floats(int maxpos) { return (2 * maxpos + maxpos + (maxpos + 1) / 2 + 3) & ~3; }
__host__ __device__ inline int chirp2_seg_floats(int maxpos) { return (maxpos + kC2N / 48 + 2 * (kC2Threads / 8) + 8 + 3) & ~3; }
__host__ __device__ inline size_t chirp2_smem_floats(int maxpos, int ppart) {
    return static_cast<size_t>(chirp2_tiles(ppart)) * kC2Threads + 2 * kC2TilePad + chirp2_rank_floats(maxpos) + chirp2_seg_floats(maxpos) +
           48 + 3 * 128 + 32 + 8;
}
__device__ inline Chirp2Smem chirp2_carve(unsigned char* base, int maxpos, int ppart) {
    Chirp2Smem S;
    float* p = reinterpret_cast<float*>(base);
    S.xd = p; p += static_cast<size_t>(chirp2_tiles(ppart)) * kC2Threads;
    S.tile = p; p += 2 * kC2TilePad;
    S.acc = reinterpret_cast<float (*)[2]>(p);
    S.a = p + 2 * maxpos;
    S.order = reinterpret_cast<unsigned short*>(p + 3 * maxpos);
    p += chirp2_rank_floats(maxpos);
    S.seg = p; p += chirp2_seg_floats(maxpos);
    S.lp = p; p += 48;
    S.exq = p; p += 128;
    S.exe = p; p += 128;
    S.ex = p; p += 128;
    S.cand = reinterpret_cast<int*>(p); p += 32;
    S.best_c = p; S.best_p = reinterpret_cast<int*>(p + 1); S.go = reinterpret_cast<int*>(p + 2); S.best_se = p + 3;
    S.nranked = reinterpret_cast<int*>(p + 4);
    return S;
}
// positions per ranking pass (a multiple of 8): the fewest passes that keep three frames on an SM, else the smallest footprint of <= 4 passes
inline int chirp2_layout(int maxpos, size_t* bytes) {
    int best = 0;
    size_t best_b = 0;
    for (int parts = 1; parts <= 4; ++parts) {
        const int ppart = ((maxpos + parts - 1) / parts + 7) & ~7;
        const size_t b = chirp2_smem_floats(maxpos, ppart) * sizeof(float);
        if (!best || b < best_b) { best = ppart; best_b = b; }
        if (b <= 75 * 1024) break;
    }
    *bytes = best_b;
    return best;
}

__device__ unsigned long long g_chirp2_clk[8];            // cycles of thread 0 per phase (diagnostics): ranking 1b, 1c, select, coarse exact, fine ranking, fine exact
__device__ unsigned long long g_chirp2_stats[3];          // {searches, coarse verification rounds, fine runs}: pu_chirp_search_stats

// the reference's closing arithmetic (:655-661)
__device__ __forceinline__ float chirp_norm(float ci, float cq, float se, float te) {
    const float denom = __fsqrt_rn(__fmul_rn(se, te));
    if (denom < 1e-10f) return 0.0f;
    return __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(ci, ci), __fmul_rn(cq, cq))), denom);
}

// D (16 x 8, fp32) += A (16 x 8, tf32, row) x B (8 x 8, tf32, col): the warp-level tensor-core instruction.  fp32 bit patterns are
// passed as they are (the unit reads the upper 19 bits: truncation to tf32), fine for a ranking estimate.
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async4_zfill(uint32_t dst, const void* src, bool valid) {
    const int bytes = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

// x = frame, L = its length, w0 = window start, Lw = window length: detectChirpTemplate(x + w0, Lw) (:560-629)
__device__ int chirp_detect_template2(const Chirp2Smem& S, const float* __restrict__ x, int L, int w0, int Lw, const float* __restrict__ ts,
                                      const float* __restrict__ tc, const float* __restrict__ tds, const float* __restrict__ tdc, float te,
                                      float threshold, int ppart, float guard, float* corr_out) {
    constexpr int n = kC2N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    *corr_out = 0.0f;
    if (Lw < n) return -1;
    const int search_len = Lw - n;
    const int n_pos = (search_len + 47) / 48;
    if (n_pos == 0) return -1;
    const float* xw = x + w0;
    __syncthreads();                                         // the previous call's shared state is dead
    // low-pass + 6:1 decimation: xd[i] = xf[base + 6 i] (window index) for ntile x 256 outputs; seg_base >= 0: 48-sample energies too
    auto decimate = [&](int base, int ntile, int seg_base) {
        // input tiles are double-buffered with cp.async (zero fill outside the frame): the synchronous form spent ~7 cycles per issued
        // instruction in this phase, waiting for each tile's loads with nothing else to do
        auto load = [&](int tl, int buf) {
            const int q0 = base + tl * kC2Threads * kRankD - kRankC;        // window index of tile[0]
            float* dst = S.tile + buf * kC2TilePad;
            for (int j = tid; j < kC2TileIn; j += kC2Threads) {
                const int g = w0 + q0 + j;
                const bool ok = g >= 0 && g < L;
                cp_async4_zfill(smem_u32(dst + j), ok ? x + g : x, ok);
            }
            cp_async_commit();
        };
        load(0, 0);
        for (int tl = 0; tl < ntile; ++tl) {
            if (tl + 1 < ntile) load(tl + 1, (tl + 1) & 1); else cp_async_commit();
            cp_async_wait_but_one();
            __syncthreads();
            {
                const float* t = S.tile + (tl & 1) * kC2TilePad + kRankD * tid;
                float y = S.lp[kRankC] * t[kRankC];            // symmetric taps: lp[k] == lp[kRankNT - 1 - k], half the multiplies
#pragma unroll
                for (int k = 0; k < kRankC; ++k) y = fmaf(S.lp[k], t[k] + t[kRankNT - 1 - k], y);
                S.xd[tl * kC2Threads + tid] = y;
                if (seg_base >= 0) {
                    float e = 0.0f;                             // 6 samples per thread, 8 threads per 48-sample segment
#pragma unroll
                    for (int k = 0; k < kRankD; ++k) { const float v = t[kRankC + k]; e = fmaf(v, v, e); }
                    e += __shfl_xor_sync(0xffffffffu, e, 1);
                    e += __shfl_xor_sync(0xffffffffu, e, 2);
                    e += __shfl_xor_sync(0xffffffffu, e, 4);
                    if ((tid & 7) == 0) S.seg[seg_base + tl * (kC2Threads / 8) + (tid >> 3)] = e;   // (parts overlap: same values)
                }
            }
            __syncthreads();                                    // this buffer is loaded again two tiles on
        }
    };
    long long tk = clock64();
    auto mark = [&](int phase) { if (tid == 0) { const long long t = clock64(); atomicAdd(&g_chirp2_clk[phase], static_cast<unsigned long long>(t - tk)); tk = t; } };
    for (int j = tid; j < n_pos; j += kC2Threads) { S.acc[j][0] = 0.0f; S.acc[j][1] = 0.0f; }
    int nseg = 0;
    for (int p0 = 0; p0 < n_pos; p0 += ppart) {                // ranking passes over [p0, p0 + np)
        const int np = min(ppart, n_pos - p0);
        // ---------------- tier 1b: the part's decimated samples (decimated index 8 p0 + i -> xd[i]) and 48-sample energies
        const int nxd = 8 * (np - 1) + kC2Nd;                  // decimated samples the part's positions touch
        const int ntile = (nxd + kC2Threads - 1) / kC2Threads;
        // (the band walk below reads a little past the last tile, against zero template rows: keep that finite)
        for (int i = ntile * kC2Threads + tid; i < chirp2_tiles(ppart) * kC2Threads; i += kC2Threads) S.xd[i] = 0.0f;
        decimate(48 * p0, ntile, p0);
        mark(0);
        nseg = p0 + ntile * (kC2Threads / 8);
        // ---------------- tier 1c: correlation estimates on the tensor cores.  With X[i][s] = xd[8 i + s] (one row per coarse step) and
        // T_q[s] = t[8 q + s] (the decimated template cut into 500 rows of 8), the estimate of position m is the sum over q of
        // <X[m + q], T_q>: the sums along the diagonals of G = X T^T.  A warp walks one band of diagonals: step s multiplies the 16 rows
        // X[16 b + 8 s ..] by the 8 template rows T[8 s ..] (one m16n8k8 per template; element (r, n) belongs to diagonal 16 b + r - n at
        // EVERY step, so the band's sums accumulate in the same registers), and the rows 8..15 of a step are the rows 0..7 of the next:
        // two shared loads, four template loads and two MMAs per 2 048 multiply-adds, against 74 instructions per 64 with FMAs.
        {
            const int g = lane >> 2, t4 = lane & 3;
            const int nband = (np + 6) / 16 + 1;               // position m takes part in bands (m .. m + 7) / 16
            for (int bnd = warp; bnd < nband; bnd += kC2Warps) {
                const int i0 = 16 * bnd;
                float dc[4] = {0.0f, 0.0f, 0.0f, 0.0f}, ds[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                const float* xa = S.xd + 8 * (i0 + g) + t4;
                uint32_t af[4];
                af[0] = __float_as_uint(xa[0]);
                af[2] = __float_as_uint(xa[4]);
#pragma unroll 3
                for (int st = 0; st < (kC2Nd / 8 + 7) / 8; ++st) {
                    af[1] = __float_as_uint(xa[64 * st + 64]);
                    af[3] = __float_as_uint(xa[64 * st + 68]);
                    const int q = 8 * st + g;                   // template row of this lane's B column
                    uint32_t c0 = 0, c1 = 0, s0 = 0, s1 = 0;
                    if (q < kC2Nd / 8) {
                        c0 = __float_as_uint(__ldg(&tdc[8 * q + t4]));
                        c1 = __float_as_uint(__ldg(&tdc[8 * q + t4 + 4]));
                        s0 = __float_as_uint(__ldg(&tds[8 * q + t4]));
                        s1 = __float_as_uint(__ldg(&tds[8 * q + t4 + 4]));
                    }
                    mma_tf32_16x8x8(dc, af, c0, c1);
                    mma_tf32_16x8x8(ds, af, s0, s1);
                    af[0] = af[1];
                    af[2] = af[3];
                }
#pragma unroll
                for (int v = 0; v < 4; ++v) {                   // fragment element v of lane (g, t4): row g + 8 (v >> 1), column 2 t4 + (v & 1)
                    const int ml = i0 + g + 8 * (v >> 1) - (2 * t4 + (v & 1));
                    if (ml >= 0 && ml < np) { atomicAdd(&S.acc[p0 + ml][0], dc[v]); atomicAdd(&S.acc[p0 + ml][1], ds[v]); }
                }
            }
        }
        __syncthreads();                                       // the next part overwrites xd
        mark(1);
    }
    // exclusive prefix of the segment energies (one warp)
    if (warp == 0) {
        float carry = 0.0f;
        for (int b = 0; b < nseg + 1; b += 32) {
            const float v = (b + lane < nseg) ? S.seg[b + lane] : 0.0f;
            float incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            if (b + lane <= nseg) S.seg[b + lane] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    for (int m = tid; m < n_pos; m += kC2Threads) {
        const float se = S.seg[m + n / 48] - S.seg[m];
        const float denom = sqrtf(fmaxf(se, 0.0f) * te);
        S.a[m] = denom < 1e-10f ? 0.0f : sqrtf(S.acc[m][0] * S.acc[m][0] + S.acc[m][1] * S.acc[m][1]) / denom;
    }
    __syncthreads();
    // Ranking: order[r] = the position with the r-th largest estimate (ties by position).  Only the leaders are ever needed unless the stop
    // rule fails, so the full rank-by-counting (n_pos^2 comparisons: 9 % of the kernel's instructions) is replaced by a selection: a
    // 4 096-bin histogram of the estimates' bit patterns (they are >= 0, so the patterns order like the values) finds the bin that holds
    // the 17th largest, the positions from that bin upwards are compacted and ranked among themselves.  A later round that reaches
    // beyond them ranks everything (rank_all).
    auto rank_all = [&]() {
        for (int m = tid; m < n_pos; m += kC2Threads) {
            const float am = S.a[m];
            int rk = 0;
            for (int o = 0; o < n_pos; ++o) {
                const float ao = S.a[o];
                rk += (ao > am || (ao == am && o < m)) ? 1 : 0;
            }
            S.order[rk] = static_cast<unsigned short>(m);
        }
        __syncthreads();
        if (tid == 0) *S.nranked = n_pos;
        __syncthreads();
    };
    {
        unsigned* hist = reinterpret_cast<unsigned*>(S.xd);                           // the decimated window is dead
        unsigned short* list = reinterpret_cast<unsigned short*>(S.xd + 4096);        // (xd holds >= 6 656 floats)
        for (int i = tid; i < 4096; i += kC2Threads) hist[i] = 0;
        if (tid == 0) *S.nranked = 0;
        __syncthreads();
        for (int m = tid; m < n_pos; m += kC2Threads) atomicAdd(&hist[__float_as_uint(S.a[m]) >> 19], 1u);
        __syncthreads();
        if (warp == 0) {                                       // lane l owns bins [128 l, 128 l + 128)
            unsigned own = 0;
            for (int b = 0; b < 128; ++b) own += hist[128 * lane + b];
            unsigned suf = own;                                // positions in this lane's bins and above
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned u = __shfl_down_sync(0xffffffffu, suf, o);
                if (lane + o < 32) suf += u;
            }
            const unsigned target = static_cast<unsigned>(min(kC2Cand + 1, n_pos));   // + 1: the stop rule looks at the first position not verified
            const unsigned has = __ballot_sync(0xffffffffu, suf >= target);
            const int top = 31 - __clz(static_cast<int>(has));             // the highest lane whose suffix still reaches the target
            if (lane == top) {
                unsigned run = suf - own;
                int b = 127;
                for (; b > 0; --b) {
                    run += hist[128 * lane + b];
                    if (run >= target) break;
                }
                *S.go = 128 * lane + b;
            }
        }
        __syncthreads();
        const unsigned bstar = static_cast<unsigned>(*S.go);
        for (int m = tid; m < n_pos; m += kC2Threads)
            if ((__float_as_uint(S.a[m]) >> 19) >= bstar) list[atomicAdd(S.nranked, 1)] = static_cast<unsigned short>(m);
        __syncthreads();
        const int cntl = *S.nranked;
        for (int i = tid; i < cntl; i += kC2Threads) {
            const int m = list[i];
            const float am = S.a[m];
            int rk = 0;
            for (int j = 0; j < cntl; ++j) {
                const int o = list[j];
                const float ao = S.a[o];
                rk += (ao > am || (ao == am && o < m)) ? 1 : 0;
            }
            S.order[rk] = static_cast<unsigned short>(m);
        }
        __syncthreads();
    }
    if (tid == 0) { (*S.best_c) = 0.0f; (*S.best_p) = -1; }
    __syncthreads();
    mark(2);
    // ---------------- tier 2: exact evaluation of the leaders, 16 per round.
    // Lane = candidate: warp 0 runs the ci chains in lanes 0..15 and the cq chains of the same candidates in lanes 16..31 (the two halves
    // read the same sample word: a broadcast), warp 1 the energy chains; every warp stages.  Tiles of 128 taps: 16 sample rows (odd stride)
    // + the template tile travel through a cp.async ring of as many buffers as the decimated window's storage holds (3..6); a group is
    // committed every iteration, empty ones past the end, so "all groups but the newest nbuf - 2" always means "tile t has landed".
    // (v48 breakdown per 64-tap tile of 32 candidates: ~250 cycles barrier + wait, ~730 staging -- 64 four-byte LDGSTS, the rows of scattered
    // candidates have no common alignment --, ~820 summing under contention: half the rows and twice the taps per tile halve the first two.)
    const int nbuf = min(6, (chirp2_tiles(ppart) * kC2Threads) / kC2VBuf);
    float (*rows)[kC2Cand][kC2Row] = reinterpret_cast<float (*)[kC2Cand][kC2Row]>(S.xd);
    float (*tpl)[2][kC2VTaps] = reinterpret_cast<float (*)[2][kC2VTaps]>(S.xd + ((nbuf * kC2Cand * kC2Row + 3) & ~3));   // 16-byte aligned
    constexpr int ntv = (n + kC2VTaps - 1) / kC2VTaps;
    float errmax = 0.0f;                                     // thread 0
    for (int nv = 0; nv < n_pos; nv += kC2Cand) {
        const int cnt = min(kC2Cand, n_pos - nv);
        if (nv + cnt > *S.nranked) rank_all();                 // (uniform: shared value)
        if (tid < kC2Cand) S.cand[tid] = tid < cnt ? 48 * static_cast<int>(S.order[nv + tid]) : 0;
        __syncthreads();
        auto stage = [&](int tile, int buf) {
            for (int r = warp; r < kC2Cand; r += kC2Warps) {
                const float* src = xw + S.cand[r] + kC2VTaps * tile;
                const uint32_t dst = smem_u32(&rows[buf][r][0]);
#pragma unroll
                for (int q = 0; q < kC2VTaps / 32; ++q)
                    if (kC2VTaps * tile + 32 * q + lane < n) cp_async4(dst + 4 * (32 * q + lane), src + 32 * q + lane);
            }
            if (warp >= kC2Warps - 2 && kC2VTaps * tile + 4 * lane < n)     // the template tile travels with the samples (192 KB per chirp: L2, not L1)
                cp_async16(smem_u32(&tpl[buf][warp - (kC2Warps - 2)][4 * lane]), (warp == kC2Warps - 2 ? tc : ts) + kC2VTaps * tile + 4 * lane);
            cp_async_commit();
        };
        float sum = 0.0f;
        float* xch = &S.acc[0][0];                              // [16] cq (the ranking accumulators are dead)
        for (int i = 0; i < nbuf - 1; ++i) stage(i, i);
        for (int tile = 0, buf = 0; tile < ntv; ++tile, buf = buf + 1 == nbuf ? 0 : buf + 1) {
            cp_async_wait_but(nbuf - 2);
            __syncthreads();
            if (tile + nbuf - 1 < ntv) stage(tile + nbuf - 1, buf == 0 ? nbuf - 1 : buf - 1); else cp_async_commit();
            const int tn = min(kC2VTaps, n - kC2VTaps * tile);
            if (warp == 0) {
                const float* row = rows[buf][lane & (kC2Cand - 1)];
                const float* tv = tpl[buf][lane >> 4];
#pragma unroll 4
                for (int t = 0; t < tn; t += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(tv + t);
                    sum = __fadd_rn(sum, __fmul_rn(row[t], v.x));
                    sum = __fadd_rn(sum, __fmul_rn(row[t + 1], v.y));
                    sum = __fadd_rn(sum, __fmul_rn(row[t + 2], v.z));
                    sum = __fadd_rn(sum, __fmul_rn(row[t + 3], v.w));
                }
            } else if (warp == 1) {
                const float* row = rows[buf][lane & (kC2Cand - 1)];
#pragma unroll 16
                for (int t = 0; t < tn; ++t) { const float v = row[t]; sum = __fadd_rn(sum, __fmul_rn(v, v)); }
            }
        }
        if (warp == 0 && lane >= kC2Cand) xch[lane - kC2Cand] = sum;
        if (warp == 1 && lane < kC2Cand) S.exe[lane] = sum;
        __syncthreads();
        if (warp == 0 && lane < kC2Cand) S.ex[lane] = chirp_norm(sum, xch[lane], S.exe[lane], te);
        __syncthreads();
        if (tid == 0) {
            float bc = (*S.best_c);
            int bp = (*S.best_p);
            for (int r = 0; r < cnt; ++r) {
                const float c = S.ex[r];
                const int pos = S.cand[r];
                errmax = fmaxf(errmax, fabsf(c - S.a[S.order[nv + r]]));
                if (c > bc || (c == bc && bp >= 0 && pos < bp)) { bc = c; bp = pos; (*S.best_se) = S.exe[r]; }   // the ascending scan's first maximum
            }
            (*S.best_c) = bc;
            (*S.best_p) = bp;
            int go = 0;
            if (nv + kC2Cand < n_pos) {
                // the largest unverified estimate (or, past the ranked leaders, an upper bound of it: the smallest ranked one)
                const float next = S.a[S.order[min(nv + kC2Cand, *S.nranked - 1)]];
                go = !(next == 0.0f || next + guard * errmax + 1e-6f < bc);
            }
            (*S.go) = go;
        }
        __syncthreads();
        if (!(*S.go)) {
            if (tid == 0) { atomicAdd(&g_chirp2_stats[0], 1ull); atomicAdd(&g_chirp2_stats[1], static_cast<unsigned long long>(nv / kC2Cand + 1)); }
            break;
        }
    }
    float best = (*S.best_c);
    int best_pos = (*S.best_p);
    __syncthreads();
    mark(3);
    *corr_out = best;
    if (best_pos < 0 || best < __fmul_rn(threshold, 0.3f)) return -1;
    // ---------------- fine search (:600-612) and the parabola's neighbours (:615-625), two-tier as well.  The correlation magnitude is
    // band-limited (2.4 kHz: main lobe 40 samples null to null), so estimates every 3 samples -- the decimated correlation again, two phases,
    // on the grid gbase + 3k -- locate its maximum; a run of 16 consecutive positions around it is evaluated exactly (lane = position over
    // one linear tile: consecutive banks), and further runs follow while (a) a neighbour of the current first maximum is not exact yet
    // or (b) an unverified grid point is within reach: kFineReach x estimate (the most the magnitude can rise next to a grid point)
    // + 4 x the largest |exact - estimate| seen.  S.ex[pos - q0] = exact value or -1 (not evaluated: cannot win, cannot be needed).
    const int fine_start = max(0, best_pos - 48), fine_end = min(search_len, best_pos + 48);
    const int q0 = max(0, fine_start - 1), q1 = min(search_len, fine_end + 1);
    {
        // grid of estimates every 3 samples: two decimation phases (offsets 0 and 3) of the 6:1 low-passed window
        const int klo = -((best_pos - fine_start) / 6);
        const int gbase = best_pos + 6 * klo;                  // window index of grid point 0 (>= fine_start)
        const int nk = (fine_end - gbase) / 3 + 1;             // grid point kk sits at gbase + 3 kk (<= 33 points)
        float* est = S.exq;
        const float dn = sqrtf(fmaxf(*S.best_se, 0.0f) * te);  // the window energy moves by < 0.5 % over +-48 samples
        for (int ph = 0; ph < 2; ++ph) {
            const int nph = (nk - ph + 1) / 2;                 // points of this phase: kk = ph, ph + 2, ...
            if (nph <= 0) break;
            decimate(gbase + 3 * ph, (nph - 1 + kC2Nd + kC2Threads - 1) / kC2Threads, -1);
            for (int k = warp; k < nph; k += kC2Warps) {
                float ac = 0.0f, as = 0.0f;
                for (int j = lane; j < kC2Nd; j += 32) {
                    const float v = S.xd[k + j];
                    ac = fmaf(v, __ldg(&tdc[j]), ac);
                    as = fmaf(v, __ldg(&tds[j]), as);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { ac += __shfl_xor_sync(0xffffffffu, ac, o); as += __shfl_xor_sync(0xffffffffu, as, o); }
                if (lane == 0) est[2 * k + ph] = dn < 1e-10f ? 0.0f : sqrtf(ac * ac + as * as) / dn;
            }
            __syncthreads();                                   // the next phase overwrites xd
        }
        for (int i = tid; i < 128; i += kC2Threads) S.ex[i] = -1.0f;
        __syncthreads();
        mark(4);
        if (tid == 0) S.ex[best_pos - q0] = best;              // exact from the coarse stage
        constexpr int NB = 6, FR = 16, FT = 256;               // staging ring (see the coarse stage); positions per exact run; taps per tile
        float (*lin)[FT + FR] = reinterpret_cast<float (*)[FT + FR]>(S.xd);
        float (*ftpl)[2][FT] = reinterpret_cast<float (*)[2][FT]>(S.xd + NB * (FT + FR));   // [NB][cos, sin][FT], 16-byte aligned
        float ferr = 0.0f;                                     // thread 0
        for (int it = 0;; ++it) {
            if (tid == 0) {
                int centre = -1;
                if (it == 0) {
                    int kb = 0;
                    for (int k = 1; k < nk; ++k) if (est[k] > est[kb]) kb = k;
                    centre = gbase + 3 * kb;
                } else {
                    for (int k = 0; k < nk; ++k) {
                        const float v = S.ex[gbase + 3 * k - q0];
                        if (v >= 0.0f) ferr = fmaxf(ferr, fabsf(v - est[k]));
                    }
                    float fb = best;
                    int fp = best_pos;
                    for (int pos = fine_start; pos <= fine_end; ++pos) {
                        const float c = S.ex[pos - q0];
                        if (c > fb) { fb = c; fp = pos; }
                    }
                    if (fp > 0 && fp < search_len - 1 && (S.ex[fp - 1 - q0] < 0.0f || S.ex[fp + 1 - q0] < 0.0f)) centre = fp;
                    for (int k = 0; k < nk && centre < 0; ++k)
                        if (S.ex[gbase + 3 * k - q0] < 0.0f && est[k] * kFineReach + guard * ferr + 1e-6f >= fb) centre = gbase + 3 * k;
                }
                S.cand[0] = centre < 0 ? -1 : min(max(centre - FR / 2, q0), max(q0, q1 - (FR - 1)));
                if (centre >= 0) atomicAdd(&g_chirp2_stats[2], 1ull);
            }
            __syncthreads();
            const int r0 = S.cand[0];
            if (r0 < 0) break;
            auto stage = [&](int tile, int buf) {
                for (int j = tid; j < FT + FR; j += kC2Threads) {
                    const int wi = r0 + FT * tile + j;        // window index; lanes past q1 read on into the frame or zeros
                    const bool ok = w0 + wi < L;
                    cp_async4_zfill(smem_u32(&lin[buf][j]), ok ? xw + wi : x, ok);
                }
                if (tid < FT / 2 && FT * tile + 4 * (tid & 63) < n)
                    cp_async16(smem_u32(&ftpl[buf][tid >> 6][4 * (tid & 63)]), (tid < 64 ? tc : ts) + FT * tile + 4 * (tid & 63));
                cp_async_commit();
            };
            // warp 0: ci of position r0 + lane in lanes 0..15, cq of the same positions in lanes 16..31; warp 1: their energies
            float sum = 0.0f;
            float* xch = &S.acc[0][0];                         // [32] cq, se
            constexpr int nft = (n + FT - 1) / FT;
            for (int i = 0; i < NB - 1; ++i) stage(i, i);
            for (int tile = 0; tile < nft; ++tile) {
                cp_async_wait_but(NB - 2);
                __syncthreads();
                if (tile + NB - 1 < nft) stage(tile + NB - 1, (tile + NB - 1) % NB); else cp_async_commit();
                const float* row = lin[tile % NB] + (lane & (FR - 1));
                const int tn = min(FT, n - FT * tile);
                if (warp == 0) {
                    const float* tv = ftpl[tile % NB][lane >> 4];
#pragma unroll 4
                    for (int t = 0; t < tn; t += 4) {
                        const float4 v = *reinterpret_cast<const float4*>(tv + t);
                        sum = __fadd_rn(sum, __fmul_rn(row[t], v.x));
                        sum = __fadd_rn(sum, __fmul_rn(row[t + 1], v.y));
                        sum = __fadd_rn(sum, __fmul_rn(row[t + 2], v.z));
                        sum = __fadd_rn(sum, __fmul_rn(row[t + 3], v.w));
                    }
                } else if (warp == 1) {
#pragma unroll 16
                    for (int t = 0; t < tn; ++t) { const float v = row[t]; sum = __fadd_rn(sum, __fmul_rn(v, v)); }
                }
            }
            if (warp == 0 && lane >= FR) xch[lane - FR] = sum;
            if (warp == 1 && lane < FR) xch[FR + lane] = sum;
            __syncthreads();
            if (warp == 0 && lane < FR && r0 + lane <= q1) S.ex[r0 + lane - q0] = chirp_norm(sum, xch[lane], xch[FR + lane], te);
            __syncthreads();
        }
    }
    mark(5);
    if (tid == 0) {
        for (int pos = fine_start; pos <= fine_end; ++pos) {
            const float c = S.ex[pos - q0];
            if (c > best) { best = c; best_pos = pos; }
        }
        if (best_pos > 0 && best_pos < search_len - 1) {
            const float c0 = S.ex[best_pos - 1 - q0], c1 = best, c2 = S.ex[best_pos + 1 - q0];
            const float denom = __fmul_rn(2.0f, __fadd_rn(__fsub_rn(c0, __fmul_rn(2.0f, c1)), c2));
            if (fabsf(denom) > 1e-10f) {
                float delta = __fdiv_rn(__fsub_rn(c0, c2), denom);
                delta = fmaxf(-1.0f, fminf(1.0f, delta));
                best_pos = static_cast<int>(roundf(__fadd_rn(static_cast<float>(best_pos), delta)));
            }
        }
        (*S.best_c) = best;
        (*S.best_p) = best_pos;
    }
    __syncthreads();
    best = (*S.best_c);
    best_pos = (*S.best_p);
    *corr_out = best;
    return best >= threshold ? best_pos : -1;
}

// out_info[b] = {success, up_chirp_start, down_chirp_start, start_sample (training start) or -1}; out_f[b] = {cfo_hz, up corr, down corr,
// initial rotator phase}.  frame_start / frame_nsym (optional) = the window handed to the presynced kernel (0 symbols when not found).
template <bool TWO_TIER>
__global__ void __launch_bounds__(TWO_TIER ? kC2Threads : kChirpThreads, TWO_TIER ? 3 : 1) chirp_detect_kernel(ChirpDev c, const float* __restrict__ samples, size_t frame_stride, int L,
                                                                     float threshold, int sym_len, int4* __restrict__ out_info,
                                                                     float4* __restrict__ out_f, int* __restrict__ frame_start,
                                                                     int* __restrict__ frame_nsym, float* __restrict__ cfo_out,
                                                                     float* __restrict__ phase_out, int* __restrict__ n_llr,
                                                                     int llr_per_symbol, int llr_stride, int maxpos, int ppart, float guard) {
    __shared__ ChirpShared S;
    extern __shared__ __align__(16) unsigned char chirp_smem[];
    ChirpWarpBuf* WB = reinterpret_cast<ChirpWarpBuf*>(chirp_smem);   // brute-force form: one staging buffer per warp
    const Chirp2Smem S2 = chirp2_carve(chirp_smem, maxpos, ppart);           // two-tier form
    const int tid = threadIdx.x;
    const float* x = samples + static_cast<size_t>(blockIdx.x) * frame_stride;
    if constexpr (TWO_TIER) {
        if (tid < 48) S2.lp[tid] = tid < kRankNT ? __ldg(&c.lp[tid]) : 0.0f;
    }
    // detectChirpTemplate on the window [w0, w0 + Lw) of the frame with the up (0) or down (1) template
    auto detect = [&](int w0, int Lw, int down, float* corr) {
        const float* ts = down ? c.dn_s : c.up_s;
        const float* tc = down ? c.dn_c : c.up_c;
        const float te = down ? c.dn_e : c.up_e;
        if constexpr (TWO_TIER) {
            const float* dec = c.dec + (down ? 2 : 0) * static_cast<size_t>(c.nd);
            return chirp_detect_template2(S2, x, L, w0, Lw, ts, tc, dec, dec + c.nd, te, threshold, ppart, guard, corr);
        } else {
            return chirp_detect_template(S, WB, x + w0, Lw, ts, tc, c.n, te, threshold, corr);
        }
    };
    int success = 0, up_start = -1, down_start = -1, start = -1;   // DualChirpResult defaults (chirp_sync.hpp:317-324)
    float cfo = 0.0f, up_corr = 0.0f, dn_corr = 0.0f, phase = 0.0f;
    do {
        if (L < 2 * c.n + c.gap) break;                                                       // :368-372
        const int up_pos = detect(0, L, 0, &up_corr);
        if (up_pos < 0) break;
        const int ds = up_pos + c.n / 2, expected = up_pos + c.n + c.gap, margin = 2 * c.n;      // :423-435
        int de = min(L, expected + margin);
        if (ds >= L) break;
        if (de <= ds + c.n) de = min(L, ds + 2 * c.n);
        float dc;
        const int rel = detect(ds, de - ds, 1, &dc);
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
    // two-tier search for the 48 kHz chirp (24 000 taps) and windows of at most 3 000 coarse positions (168 000-sample buffers); PU_CHIRP_SEARCH=exact keeps the
    // brute-force kernel (the A/B reference of tests/test_chirp_sync_gpu.py), which also serves every other geometry
    const char* env = std::getenv("PU_CHIRP_SEARCH");   // read per call: the A/B test flips it
    const bool force_exact = env && std::strcmp(env, "exact") == 0;
    const int maxpos = (std::max(L - c.n, 0) + 47) / 48 + 1;
    const bool two_tier = !force_exact && c.n == kC2N && c.nd == kC2Nd && maxpos <= kC2MaxPos;
    cudaError_t e;
    if (two_tier) {
        size_t smem = 0;
        const int ppart = chirp2_layout(maxpos, &smem);
        // stop rule of the exact stages: estimate + guard x (largest |exact - estimate| seen) must stay below the best exact value.
        // PU_CHIRP_GUARD (tests): a huge value makes every search verify ALL positions round by round -- the multi-round paths,
        // which the ranking makes rare, then run on every frame and must reproduce the brute-force search.
        const char* genv = std::getenv("PU_CHIRP_GUARD");
        const float guard = genv ? static_cast<float>(std::atof(genv)) : 4.0f;
        static std::atomic<uint64_t> done{0};
        if ((e = smem_optin(done, chirp_detect_kernel<true>, 232448)) != cudaSuccess) return e;
        chirp_detect_kernel<true><<<static_cast<unsigned>(B), kC2Threads, smem, st>>>(c, samples, frame_stride, L, threshold, sym_len, out_info, out_f,
                                                                                     frame_start, frame_nsym, cfo_out, phase_out, n_llr,
                                                                                     llr_per_symbol, llr_stride, maxpos, ppart, guard);
    } else {
        const size_t smem = sizeof(ChirpWarpBuf) * (kChirpThreads / 32);
        static std::atomic<uint64_t> done{0};
        if ((e = smem_optin(done, chirp_detect_kernel<false>, static_cast<int>(smem))) != cudaSuccess) return e;
        chirp_detect_kernel<false><<<static_cast<unsigned>(B), kChirpThreads, smem, st>>>(c, samples, frame_stride, L, threshold, sym_len, out_info,
                                                                                         out_f, frame_start, frame_nsym, cfo_out, phase_out,
                                                                                         n_llr, llr_per_symbol, llr_stride, 0, 0, 0.0f);
    }
    return cudaGetLastError();
}

cudaError_t chirp_phase_cycles(unsigned long long* out /* [8] */) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return e;
    if ((e = cudaMemcpyFromSymbol(out, g_chirp2_clk, sizeof(g_chirp2_clk))) != cudaSuccess) return e;
    const unsigned long long zero[8] = {};
    return cudaMemcpyToSymbol(g_chirp2_clk, zero, sizeof(zero));
}

cudaError_t chirp_search_stats(unsigned long long* out /* [3] */) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return e;
    if ((e = cudaMemcpyFromSymbol(out, g_chirp2_stats, sizeof(g_chirp2_stats))) != cudaSuccess) return e;
    const unsigned long long zero[3] = {0, 0, 0};
    return cudaMemcpyToSymbol(g_chirp2_stats, zero, sizeof(zero));
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
    // ranking tables of the two-tier search: every 6th template sample times 6, then the Kaiser-windowed (beta 4.5) 4 kHz low-pass
    const size_t nd = (n + kRankD - 1) / kRankD;
    c.nd = static_cast<int>(nd);
    c.dec = c.lp = nullptr;
    t.resize(4 * n + 4 * nd + 64, 0.0f);
    for (size_t q = 0; q < 4; ++q)
        for (size_t j = 0; j < nd; ++j) t[4 * n + q * nd + j] = static_cast<float>(kRankD) * t[q * n + kRankD * j];
    auto bessel_i0 = [](double v) { double s = 1.0, term = 1.0; for (int k = 1; k < 40; ++k) { term *= (v / (2.0 * k)) * (v / (2.0 * k)); s += term; } return s; };
    const double fc = 4000.0 / static_cast<double>(fs), beta = 4.5;
    double h[kRankNT], hs = 0.0;
    for (int k = 0; k < kRankNT; ++k) {
        const double m = k - kRankC, r = m / kRankC;
        const double sinc = m == 0.0 ? 2.0 * fc : std::sin(2.0 * 3.14159265358979323846 * fc * m) / (3.14159265358979323846 * m);
        h[k] = sinc * bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / bessel_i0(beta);
        hs += h[k];
    }
    for (int k = 0; k < kRankNT; ++k) t[4 * n + 4 * nd + k] = static_cast<float>(h[k] / hs);
}

}  // namespace pu
