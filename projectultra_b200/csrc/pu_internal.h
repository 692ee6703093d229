// projectultra_b200/csrc/pu_internal.h — shared internals of libpu_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "pu/pu_capi.h"

namespace pu {

void set_error(const char* fmt, ...);

#define PU_CUDA_TRY(expr)                                                                       \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            ::pu::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PU_ERR_CUDA;                                                                 \
        }                                                                                       \
    } while (0)

#define PU_REQUIRE(cond, msg)                     \
    do {                                          \
        if (!(cond)) {                            \
            ::pu::set_error("%s", msg);           \
            return PU_ERR_INVALID;                \
        }                                         \
    } while (0)

// Grow-only device / pinned-host scratch buffer.
struct Buffer {
    void* ptr = nullptr;
    size_t cap = 0;
    bool pinned_host = false;
    pu_status reserve(size_t bytes);
    void release();
};

}  // namespace pu

namespace pu {
// One lane of the host-buffer pipeline (pu_receive_decode_batch with PU_MEM_HOST): its own stream and buffers, so
// that the H2D copy of slab i+1 overlaps the kernels and the D2H copy of slab i.
struct PipeSlot {
    cudaStream_t stream = nullptr;
    Buffer d_in, d_llr, d_out, h_in, h_out;
    size_t off = 0, nb = 0;   // slab in flight (nb == 0: idle)
};
}  // namespace pu

struct pu_ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    cudaStream_t stream = nullptr;  // context-owned stream used when the caller passes NULL
    std::atomic<uint64_t> launches{0};
    std::atomic<uint64_t> h2d_bytes{0}, d2h_bytes{0};   // bytes moved by the host-buffer pipeline (pu_receive_decode_batch, PU_MEM_HOST)
    // staging for PU_MEM_HOST calls
    pu::Buffer d_in, d_out, d_aux, h_in, h_out;
    pu::Buffer f_llr, f_bytes, f_ok, f_iters, f_out;   // scratch of pu_frame_decode_batch (grow-only, reused across calls)
    pu::PipeSlot pipe[2];
    cudaEvent_t pipe_ev = nullptr;
    // scratch of pu_linksim_run (sweep.cu), grow-only and reused across calls: [0..6] batch buffers (rx, llr, info, ok, iters, n_llr, sync),
    // [7 + 4 s .. 10 + 4 s] slot s = {pinned descriptors, device descriptors, device counters, pinned counters}
    pu::Buffer cfo_mix;            // pu_channel_apply_cfo_batch: (cos, sin)(2 pi 1500 i / fs) per sample index, host libm
    size_t cfo_mix_len = 0;
    uint32_t cfo_mix_fs = 0;
    pu::Buffer sweep[18];
    cudaEvent_t sweep_ev[2] = {nullptr, nullptr};
};

namespace pu {
// PU_MEM_DEVICE: `stream` is the caller's stream, NULL meaning the CUDA default stream (as in the runtime API);
// PU_MEM_HOST: NULL selects the context's own non-blocking stream for the staged copies and the kernel.
inline cudaStream_t pick_stream(pu_ctx* ctx, void* stream, pu_memspace space = PU_MEM_HOST) {
    if (stream || space == PU_MEM_DEVICE) return reinterpret_cast<cudaStream_t>(stream);
    return ctx->stream;
}
}  // namespace pu

// chirp_sync.cu: dual-chirp synchronisation shared by the OFDM_CHIRP and MC-DPSK receive entry points
namespace pu {
struct ChirpDev {
    int n, gap;                        // chirp samples (24 000), gap samples (4 800)
    float fs, cfo_to_samples;          // sample rate; sample_rate / chirp_rate
    const float* up_s; const float* up_c; const float* dn_s; const float* dn_c;   // templates (generateTemplate, chirp_sync.hpp:706-735)
    float up_e, dn_e;                  // template energies
    // ranking pass of the two-tier coarse search (chirp_sync.cu): the templates taken every 6th sample and scaled by 6,
    // {up sin, up cos, down sin, down cos} x nd floats, and the 47-tap low-pass in front of the 6:1 decimation of the samples
    const float* dec; const float* lp; int nd;
};
// templates of the 300 -> 2700 Hz, 500 ms chirp pair with 100 ms gaps (OFDMChirpWaveform::getChirpConfig, ofdm_chirp_waveform.cpp:39-49 ==
// MultiCarrierDPSKConfig::getChirpConfig, multi_carrier_dpsk.hpp:78-88): host table of 4 n floats {up sin, up cos, down sin, down cos}
// followed by the ranking tables (4 nd decimated template floats, 64 low-pass taps);
// the caller uploads it and chirp_dev_bind points the descriptor at the device copy
void chirp_templates_host(float fs, std::vector<float>& table, ChirpDev& c);
inline void chirp_dev_bind(ChirpDev& c, const float* dev_table) {
    c.up_s = dev_table; c.up_c = dev_table + c.n; c.dn_s = dev_table + 2 * static_cast<size_t>(c.n); c.dn_c = dev_table + 3 * static_cast<size_t>(c.n);
    c.dec = dev_table + 4 * static_cast<size_t>(c.n);
    c.lp = c.dec + 4 * static_cast<size_t>(c.nd);
}
void tools_apply_cfo(float* samples, size_t n, float cfo_hz, float sample_rate);     // tools_cfo.cpp
float channel_power_sum(const float* tx, size_t L);                                   // channel.cu: the two halves of pu_channel_noise_std
float channel_noise_std_from_sum(float acc, size_t L, float snr_db, int convention);
cudaError_t chirp_search_stats(unsigned long long* out);
cudaError_t chirp_phase_cycles(unsigned long long* out);   // [8] thread-0 cycles per phase of the two-tier search, then cleared   // {searches, rounds} since the last call, then cleared
cudaError_t chirp_detect_launch(const ChirpDev& c, const float* samples, size_t B, size_t frame_stride, int L, float threshold, int sym_len,
                                int4* out_info, float4* out_f, int* frame_start, int* frame_nsym, float* cfo_out, float* phase_out,
                                int* n_llr, int llr_per_symbol, int llr_stride, cudaStream_t st);
}  // namespace pu
