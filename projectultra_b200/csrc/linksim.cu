// projectultra_b200/csrc/linksim.cu — the batched Monte-Carlo link: (channel ->) demodulate -> LDPC decode ->
// count frame / bit errors, chained on one stream with all intermediates resident in HBM.
//
// Reference behaviour: the trial loop of the reference's Monte-Carlo tools, e.g. tools/test_mode_snr.cpp:40-105 and
// tools/test_ofdm_chirp_pilots.cpp:166-260: per trial  channel.process -> demodulator.processPresynced ->
// getSoftBits (first 648) -> LDPCDecoder::decodeSoft -> success iff lastDecodeSuccess() AND decoded bytes == payload
// (test_mode_snr.cpp:98-104).  Trials are independent, so here they are the batch dimension.
#include <memory>
#include <new>

#include "pu_internal.h"

// implemented in ofdm_demod.cu / ldpc_decode.cu
extern "C" pu_status pu_ofdm_presynced_batch(pu_ofdm*, const float*, size_t, size_t, int, const float*, const float*,
                                             float*, size_t, float*, float*, pu_memspace, void*);
extern "C" pu_status pu_ldpc_decode_batch(pu_ldpc*, const float*, size_t, size_t, uint8_t*, size_t, uint8_t*, int32_t*,
                                          pu_memspace, void*);
extern "C" int pu_ldpc_info_bits(const pu_ldpc*);
pu_ctx* pu_ofdm_context(pu_ofdm* h);   // ofdm_demod.cu

namespace pu {

// counters[bin][6] = {frames, frame_errors, bit_errors, bits, decode_failures, iteration_sum}
__global__ void count_errors_kernel(const uint8_t* __restrict__ info, size_t info_stride, const uint8_t* __restrict__ ok,
                                    const int32_t* __restrict__ iters, const uint8_t* __restrict__ payload_pool,
                                    size_t payload_stride, const uint32_t* __restrict__ tx_index,
                                    const uint32_t* __restrict__ bin, int payload_bytes, size_t B,
                                    unsigned long long* __restrict__ counters) {
    const size_t b = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (b >= B) return;
    const uint8_t* got = info + b * info_stride;
    const uint8_t* want = payload_pool + static_cast<size_t>(tx_index ? tx_index[b] : 0) * payload_stride;
    int bit_err = 0;
    for (int i = 0; i < payload_bytes; ++i) bit_err += __popc(static_cast<unsigned>(got[i] ^ want[i]));
    const int success = ok[b] && bit_err == 0;   // tools/test_mode_snr.cpp:98-104
    unsigned long long* c = counters + static_cast<size_t>(bin ? bin[b] : 0) * 6;
    atomicAdd(&c[0], 1ull);
    if (!success) atomicAdd(&c[1], 1ull);
    if (bit_err) atomicAdd(&c[2], static_cast<unsigned long long>(bit_err));
    atomicAdd(&c[3], static_cast<unsigned long long>(payload_bytes) * 8ull);
    if (!ok[b]) atomicAdd(&c[4], 1ull);
    atomicAdd(&c[5], static_cast<unsigned long long>(iters ? iters[b] : 0));
}

}  // namespace pu

extern "C" {

pu_status pu_count_errors(pu_ctx* ctx, const uint8_t* info_bytes, size_t info_stride, const uint8_t* ok,
                          const int32_t* iters, const uint8_t* payload_pool, size_t payload_stride,
                          const uint32_t* tx_index, const uint32_t* bin, size_t payload_bytes, size_t B,
                          uint64_t* counters, void* stream) {
    PU_REQUIRE(ctx && info_bytes && ok && payload_pool && counters, "pu_count_errors: NULL argument");
    PU_REQUIRE(payload_bytes <= info_stride && payload_bytes <= payload_stride, "pu_count_errors: payload longer than its buffers");
    if (B == 0) return PU_OK;
    PU_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    (void)cudaGetLastError();
    pu::count_errors_kernel<<<static_cast<unsigned>((B + 255) / 256), 256, 0, st>>>(
        info_bytes, info_stride, ok, iters, payload_pool, payload_stride, tx_index, bin, static_cast<int>(payload_bytes), B,
        reinterpret_cast<unsigned long long*>(counters));
    ctx->launches.fetch_add(1);
    PU_CUDA_TRY(cudaGetLastError());
    return PU_OK;
}

pu_status pu_receive_decode_batch(pu_ofdm* ofdm, pu_ldpc* ldpc, const float* samples, size_t B, size_t L,
                                  int training_symbols, const float* cfo_hz, const float* cfo_phase,
                                  uint8_t* info_bytes, size_t info_stride, uint8_t* ok, int32_t* iters,
                                  pu_memspace space, void* stream) {
    PU_REQUIRE(ofdm && ldpc, "pu_receive_decode_batch: NULL handle");
    if (B == 0) return PU_OK;
    PU_REQUIRE(samples && info_bytes, "pu_receive_decode_batch: NULL data pointer");
    pu_ctx* ctx = pu_ofdm_context(ofdm);
    PU_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = pu::pick_stream(ctx, stream, space);
    const size_t kb = static_cast<size_t>((pu_ldpc_info_bits(ldpc) + 7) / 8);
    PU_REQUIRE(info_stride >= kb, "pu_receive_decode_batch: info_stride too small");
    pu_status s;
    if (space == PU_MEM_DEVICE) {
        if ((s = ctx->d_aux.reserve(B * PU_LDPC_N * sizeof(float))) != PU_OK) return s;
        float* d_llr = static_cast<float*>(ctx->d_aux.ptr);
        PU_CUDA_TRY(cudaMemsetAsync(d_llr, 0, B * PU_LDPC_N * sizeof(float), st));   // frames shorter than a codeword: erasures
        if ((s = pu_ofdm_presynced_batch(ofdm, samples, B, L, training_symbols, cfo_hz, cfo_phase, d_llr, PU_LDPC_N, nullptr,
                                         nullptr, PU_MEM_DEVICE, st)) != PU_OK) return s;
        return pu_ldpc_decode_batch(ldpc, d_llr, PU_LDPC_N, B, info_bytes, info_stride, ok, iters, PU_MEM_DEVICE, st);
    }
    // host buffers: pinned staging in slabs; samples go up, only info bytes / flags come back
    const size_t slab = std::min<size_t>(B, 16384);
    const size_t in_floats = slab * (L + 2);
    const size_t out_row = kb + 1 + sizeof(int32_t);
    if ((s = ctx->d_in.reserve(in_floats * sizeof(float))) != PU_OK) return s;
    if ((s = ctx->h_in.reserve(in_floats * sizeof(float))) != PU_OK) return s;
    if ((s = ctx->d_aux.reserve(slab * PU_LDPC_N * sizeof(float))) != PU_OK) return s;
    if ((s = ctx->d_out.reserve(slab * out_row + 64)) != PU_OK) return s;
    if ((s = ctx->h_out.reserve(slab * out_row + 64)) != PU_OK) return s;
    for (size_t off = 0; off < B; off += slab) {
        const size_t nb = std::min(slab, B - off);
        float* hin = static_cast<float*>(ctx->h_in.ptr);
        std::memcpy(hin, samples + off * L, nb * L * sizeof(float));
        float* hcfo = hin + slab * L;
        float* hph = hcfo + slab;
        for (size_t b = 0; b < nb; ++b) {
            hcfo[b] = cfo_hz ? cfo_hz[off + b] : 0.0f;
            hph[b] = cfo_phase ? cfo_phase[off + b] : 0.0f;
        }
        float* din = static_cast<float*>(ctx->d_in.ptr);
        PU_CUDA_TRY(cudaMemcpyAsync(din, hin, nb * L * sizeof(float), cudaMemcpyHostToDevice, st));
        PU_CUDA_TRY(cudaMemcpyAsync(din + slab * L, hcfo, 2 * slab * sizeof(float), cudaMemcpyHostToDevice, st));
        float* d_llr = static_cast<float*>(ctx->d_aux.ptr);
        PU_CUDA_TRY(cudaMemsetAsync(d_llr, 0, nb * PU_LDPC_N * sizeof(float), st));
        if ((s = pu_ofdm_presynced_batch(ofdm, din, nb, L, training_symbols, din + slab * L, din + slab * L + slab, d_llr,
                                         PU_LDPC_N, nullptr, nullptr, PU_MEM_DEVICE, st)) != PU_OK) return s;
        uint8_t* d_info = static_cast<uint8_t*>(ctx->d_out.ptr);
        int32_t* d_iters = reinterpret_cast<int32_t*>(d_info + ((slab * kb + 15) / 16) * 16);
        uint8_t* d_ok = reinterpret_cast<uint8_t*>(d_iters + slab);
        if ((s = pu_ldpc_decode_batch(ldpc, d_llr, PU_LDPC_N, nb, d_info, kb, d_ok, d_iters, PU_MEM_DEVICE, st)) != PU_OK) return s;
        uint8_t* hout = static_cast<uint8_t*>(ctx->h_out.ptr);
        const size_t total = static_cast<size_t>(reinterpret_cast<uint8_t*>(d_ok + slab) - d_info);
        PU_CUDA_TRY(cudaMemcpyAsync(hout, d_info, total, cudaMemcpyDeviceToHost, st));
        PU_CUDA_TRY(cudaStreamSynchronize(st));
        const int32_t* h_iters = reinterpret_cast<const int32_t*>(hout + ((slab * kb + 15) / 16) * 16);
        const uint8_t* h_ok = reinterpret_cast<const uint8_t*>(h_iters + slab);
        for (size_t b = 0; b < nb; ++b) {
            std::memcpy(info_bytes + (off + b) * info_stride, hout + b * kb, kb);
            if (ok) ok[off + b] = h_ok[b];
            if (iters) iters[off + b] = h_iters[b];
        }
    }
    return PU_OK;
}

}  // extern "C"
