// projectultra_b200/csrc/channel.cu — batched Watterson two-tap Rayleigh fading + AWGN channel for sm_100a and
// the pu_channel_* entry points of the C ABI.
//
// Reference behaviour: sim::WattersonChannel (src/sim/hf_channel.hpp:67-168,258-275) -- per sample
//   fading_k <- (1-a) fading_k + a * sqrt(1/a) * (N(0,1) + j N(0,1)),  k = 1, 2, both starting from (1, 0)  (:91-92, :258-275)
//   out = x[n] g1 |fading_1| + x[n-d-1] g2 |fading_2| + sigma N(0,1)                                    (:125-153)
// with d = floor(delay_ms * fs / 1000) (:72-78; the d+1-entry zero-initialised deque gives an effective delay of d+1
// samples), a = 1 - exp(-2 pi f_d / fs) (:87-88) and sigma = rms(input) * 10^(-snr/20) (:110-119).
// What is NOT reproduced is the reference's random stream (mt19937 + libstdc++ normal_distribution, implementation
// defined): it is replaced by the counter-based generator of pu_rng.cuh, and the one-pole recurrence is evaluated as
// a 32-wide inclusive scan with fused multiply-adds instead of sample by sample.  Both are a SPECIFICATION of this
// simulator, restated independently by the CPU twin oracle/pu_oracle_channel.c (bit-identical output); against the
// reference the channel is checked statistically (tests/test_channel_gpu.py).  The optional CFO injector of the
// reference (applyCFO, :173-232) is out of scope: the Monte-Carlo configs are CFO-free.
//
// Kernel shape: one warp per frame; lane l owns sample 32 g + l, so global loads/stores are fully coalesced and
// the recurrence across the 32 samples of a group is a Kogge-Stone scan over warp shuffles (5 fma per component).
#include <cmath>
#include <memory>
#include <new>
#include <vector>

#include "pu_internal.h"
#include "pu_rng.cuh"

namespace pu {

struct ChannelParams {
    int fading, multipath, noise;
    int delay;               // d: second path reads x[n - d - 1]
    float g1, g2;
    float alpha, noise_scale;       // a, sqrt(1/a)
    float apow2[5];          // (1-a)^(2^s)
    float apl[32];           // (1-a)^(l+1)
};

__global__ void __launch_bounds__(256) channel_kernel(ChannelParams p, const float* __restrict__ tx_pool, size_t pool_stride,
                                                      const uint32_t* __restrict__ tx_index, const float* __restrict__ noise_std,
                                                      const uint64_t* __restrict__ seed, size_t B, int L,
                                                      float* __restrict__ rx) {
    const int lane = threadIdx.x & 31;
    const size_t frame = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
    if (frame >= B) return;
    const float* x = tx_pool + static_cast<size_t>(tx_index ? tx_index[frame] : 0) * pool_stride;
    float* y = rx + frame * static_cast<size_t>(L);
    const uint64_t sd = seed[frame];
    const uint32_t k0 = static_cast<uint32_t>(sd), k1 = static_cast<uint32_t>(sd >> 32);
    const float sigma = noise_std[frame];
    const float apl = p.apl[lane];
    float carry[4] = {1.0f, 0.0f, 1.0f, 0.0f};   // fading1_, fading2_ start at (1, 0)
    for (int base = 0; base < L; base += 32) {
        const int n = base + lane;
        const bool in = n < L;
        const float xv = in ? __ldg(&x[n]) : 0.0f;
        float m1 = 1.0f, m2 = 1.0f;
        if (p.fading) {
            float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (in) rng::fading_normals(k0, k1, static_cast<uint32_t>(n), z);
            float f[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v = rng::r_mul(p.alpha, rng::r_mul(p.noise_scale, z[c]));
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    const float up = __shfl_up_sync(0xffffffffu, v, 1 << s);
                    if (lane >= (1 << s)) v = rng::r_fma(p.apow2[s], up, v);
                }
                f[c] = rng::r_fma(apl, carry[c], v);
                carry[c] = __shfl_sync(0xffffffffu, f[c], 31);
            }
            m1 = rng::r_sqrt(rng::r_fma(f[0], f[0], rng::r_mul(f[1], f[1])));
            m2 = rng::r_sqrt(rng::r_fma(f[2], f[2], rng::r_mul(f[3], f[3])));
        }
        if (!in) continue;
        float out;
        if (p.multipath && p.delay > 0) {
            const int nd = n - p.delay - 1;
            const float xd = nd >= 0 ? __ldg(&x[nd]) : 0.0f;
            out = rng::r_fma(rng::r_mul(xd, p.g2), m2, rng::r_mul(rng::r_mul(xv, p.g1), m1));
        } else {
            out = rng::r_mul(xv, m1);
        }
        if (p.noise) out = rng::r_fma(sigma, rng::noise_normal(k0, k1, static_cast<uint32_t>(n)), out);
        y[n] = out;
    }
}

// AWGN-only fast path: no recurrence, each lane produces 4 consecutive samples from one Philox call.
__global__ void __launch_bounds__(256) awgn_kernel(const float* __restrict__ tx_pool, size_t pool_stride,
                                                   const uint32_t* __restrict__ tx_index, const float* __restrict__ noise_std,
                                                   const uint64_t* __restrict__ seed, size_t B, int L, float* __restrict__ rx) {
    const int lane = threadIdx.x & 31;
    const size_t frame = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
    if (frame >= B) return;
    const float* x = tx_pool + static_cast<size_t>(tx_index ? tx_index[frame] : 0) * pool_stride;
    float* y = rx + frame * static_cast<size_t>(L);
    const uint64_t sd = seed[frame];
    const uint32_t k0 = static_cast<uint32_t>(sd), k1 = static_cast<uint32_t>(sd >> 32);
    const float sigma = noise_std[frame];
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    for (int q = lane; 4 * q < L; q += 32) {
        const rng::U4 w = rng::philox4x32_10(static_cast<uint32_t>(q), 0u, rng::kStreamNoise, 0u, k0, k1);
        float z[4];
        rng::box_muller(w.x, w.y, &z[0], &z[1]);
        rng::box_muller(w.z, w.w, &z[2], &z[3]);
        const int n = 4 * q;
        if (vec && n + 3 < L) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + n));
            float4 o;
            o.x = rng::r_fma(sigma, z[0], xv.x);
            o.y = rng::r_fma(sigma, z[1], xv.y);
            o.z = rng::r_fma(sigma, z[2], xv.z);
            o.w = rng::r_fma(sigma, z[3], xv.w);
            *reinterpret_cast<float4*>(y + n) = o;
        } else {
            for (int j = 0; j < 4 && n + j < L; ++j) y[n + j] = rng::r_fma(sigma, z[j], __ldg(&x[n + j]));
        }
    }
}

static ChannelParams make_params(const pu_channel_config& c) {
    ChannelParams p{};
    p.fading = c.fading_enabled != 0;
    p.multipath = c.multipath_enabled != 0;
    p.noise = c.noise_enabled != 0;
    p.delay = static_cast<int>(static_cast<size_t>(c.delay_spread_ms * static_cast<float>(c.sample_rate) / 1000.0f));   // hf_channel.hpp:72-75
    p.g1 = c.path1_gain;
    p.g2 = c.path2_gain;
    const float norm_doppler = c.doppler_spread_hz / static_cast<float>(c.sample_rate);
    p.alpha = static_cast<float>(1.0f - std::exp(-2.0f * 3.14159265358979323846 * norm_doppler));                      // :87-88
    p.noise_scale = p.alpha > 0.0f ? std::sqrt(1.0f / p.alpha) : 0.0f;                                                  // :264
    const float a = 1.0f - p.alpha;
    float v = a;
    for (int s = 0; s < 5; ++s) { p.apow2[s] = v; v = v * v; }
    v = a;
    for (int l = 0; l < 32; ++l) { p.apl[l] = v; v = v * a; }
    return p;
}

}  // namespace pu

namespace pu {
// pu_channel_noise_std for every row of tx[B][stride] (device): the ordered fp32 power sum of the reference's tools
// (tools/test_mode_snr.cpp:58-61) / of WattersonChannel::process (hf_channel.hpp:110-119), one thread per frame.  The
// SNR-dependent factor -- powf(10, snr/10) (convention 1) or powf(10, -snr/20) (convention 0) -- is evaluated on the host
// per SNR point and gathered per frame, so no device pow enters the noise level.
__global__ void noise_std_kernel(const float* __restrict__ tx, size_t stride, int L, size_t B, const float* __restrict__ factor, int convention,
                                 float* __restrict__ out) {
    const size_t b = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (b >= B) return;
    const float* w = tx + b * stride;
    float acc = 0.0f;
    for (int i = 0; i < L; ++i) acc = __fadd_rn(acc, __fmul_rn(w[i], w[i]));
    const float mean = __fdiv_rn(acc, static_cast<float>(L));
    out[b] = convention == 0 ? __fmul_rn(__fsqrt_rn(mean), factor[b]) : __fsqrt_rn(__fdiv_rn(mean, factor[b]));
}
}  // namespace pu

extern "C" {

pu_status pu_channel_params(const pu_channel_config* cfg, int32_t* delay_samples, float* alpha, float* noise_scale,
                            float* apow2_5, float* apl_32) {
    PU_REQUIRE(cfg, "pu_channel_params: NULL config");
    const pu::ChannelParams p = pu::make_params(*cfg);
    if (delay_samples) *delay_samples = p.delay;
    if (alpha) *alpha = p.alpha;
    if (noise_scale) *noise_scale = p.noise_scale;
    if (apow2_5) std::memcpy(apow2_5, p.apow2, sizeof(p.apow2));
    if (apl_32) std::memcpy(apl_32, p.apl, sizeof(p.apl));
    return PU_OK;
}

pu_status pu_channel_noise_std_batch(pu_ctx* ctx, const float* tx, size_t tx_stride, size_t L, size_t B, const float* snr_factor,
                                     int convention, float* noise_std, void* stream) {
    PU_REQUIRE(ctx, "pu_channel_noise_std_batch: NULL context");
    if (B == 0) return PU_OK;
    PU_REQUIRE(tx && snr_factor && noise_std && tx_stride >= L && L > 0, "pu_channel_noise_std_batch: bad argument");
    PU_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    (void)cudaGetLastError();
    pu::noise_std_kernel<<<static_cast<unsigned>((B + 63) / 64), 64, 0, st>>>(tx, tx_stride, static_cast<int>(L), B, snr_factor, convention, noise_std);
    ctx->launches.fetch_add(1);
    PU_CUDA_TRY(cudaGetLastError());
    return PU_OK;
}

float pu_channel_noise_std(const float* tx, size_t L, float snr_db, int convention) {
    if (!tx || L == 0) return 0.0f;
    float acc = 0.0f;
    for (size_t i = 0; i < L; ++i) acc += tx[i] * tx[i];
    if (convention == 0) {          // WattersonChannel::process, hf_channel.hpp:110-119
        const float rms = std::sqrt(acc / L);
        return rms * std::pow(10.0f, -snr_db / 20.0f);
    }
    const float sp = acc / L;       // tools/test_mode_snr.cpp:58-61 (mean frame power, AWGN tools)
    return std::sqrt(sp / std::pow(10.0f, snr_db / 10.0f));
}

pu_status pu_channel_apply_batch(pu_ctx* ctx, const pu_channel_config* cfg, const float* tx_pool, size_t pool_stride,
                                 size_t pool_count, const uint32_t* tx_index, const float* noise_std,
                                 const uint64_t* seed, size_t B, size_t L, float* rx, pu_memspace space, void* stream) {
    PU_REQUIRE(ctx && cfg, "pu_channel_apply_batch: NULL argument");
    if (B == 0 || L == 0) return PU_OK;
    PU_REQUIRE(tx_pool && noise_std && seed && rx, "pu_channel_apply_batch: NULL data pointer");
    PU_REQUIRE(pool_stride >= L && pool_count >= 1, "pu_channel_apply_batch: pool_stride < L or empty pool");
    PU_REQUIRE(L < (1u << 30), "pu_channel_apply_batch: frame too long");
    PU_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = pu::pick_stream(ctx, stream, space);
    const pu::ChannelParams p = pu::make_params(*cfg);
    const bool awgn_only = !p.fading && !(p.multipath && p.delay > 0) && p.noise;
    auto launch = [&](const float* d_pool, const uint32_t* d_idx, const float* d_std, const uint64_t* d_seed, size_t nb,
                      float* d_rx) -> pu_status {
        const unsigned blocks = static_cast<unsigned>((nb * 32 + 255) / 256);
        (void)cudaGetLastError();
        if (awgn_only) pu::awgn_kernel<<<blocks, 256, 0, st>>>(d_pool, pool_stride, d_idx, d_std, d_seed, nb, static_cast<int>(L), d_rx);
        else pu::channel_kernel<<<blocks, 256, 0, st>>>(p, d_pool, pool_stride, d_idx, d_std, d_seed, nb, static_cast<int>(L), d_rx);
        ctx->launches.fetch_add(1);
        PU_CUDA_TRY(cudaGetLastError());
        return PU_OK;
    };
    if (space == PU_MEM_DEVICE) return launch(tx_pool, tx_index, noise_std, seed, B, rx);

    // host buffers (tests, small batches): plain staging
    float *d_pool = nullptr, *d_std = nullptr, *d_rx = nullptr;
    uint32_t* d_idx = nullptr;
    uint64_t* d_seed = nullptr;
    std::vector<uint32_t> idx(B, 0);
    if (tx_index) for (size_t b = 0; b < B; ++b) { idx[b] = tx_index[b]; PU_REQUIRE(idx[b] < pool_count, "pu_channel_apply_batch: tx_index out of range"); }
    PU_CUDA_TRY(cudaMalloc(&d_pool, pool_count * pool_stride * sizeof(float)));
    PU_CUDA_TRY(cudaMalloc(&d_std, B * sizeof(float)));
    PU_CUDA_TRY(cudaMalloc(&d_rx, B * L * sizeof(float)));
    PU_CUDA_TRY(cudaMalloc(&d_idx, B * sizeof(uint32_t)));
    PU_CUDA_TRY(cudaMalloc(&d_seed, B * sizeof(uint64_t)));
    PU_CUDA_TRY(cudaMemcpyAsync(d_pool, tx_pool, pool_count * pool_stride * sizeof(float), cudaMemcpyHostToDevice, st));
    PU_CUDA_TRY(cudaMemcpyAsync(d_std, noise_std, B * sizeof(float), cudaMemcpyHostToDevice, st));
    PU_CUDA_TRY(cudaMemcpyAsync(d_idx, idx.data(), B * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PU_CUDA_TRY(cudaMemcpyAsync(d_seed, seed, B * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    pu_status s = launch(d_pool, d_idx, d_std, d_seed, B, d_rx);
    if (s == PU_OK) {
        cudaError_t e = cudaMemcpyAsync(rx, d_rx, B * L * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { pu::set_error("pu_channel_apply_batch: %s", cudaGetErrorString(e)); s = PU_ERR_CUDA; }
    }
    cudaFree(d_pool); cudaFree(d_std); cudaFree(d_rx); cudaFree(d_idx); cudaFree(d_seed);
    return s;
}

}  // extern "C"
