// projectultra_b200/csrc/frame_v2.cu — protocol-v2 multi-codeword frames behind the batched decoder.
//
// Reference behaviour: RxPipeline::decodeFrame (src/gui/modem/rx_pipeline.cpp:348-445) over v2::decodeSingleCodeword
// (src/protocol/frame_v2.cpp:1134-1156), v2::parseHeader (:1175-1230, CRC-16/CCITT of :111-124), CodewordStatus::reassemble /
// reassembleCodewords (:952-982,1023-1044), and v2::encodeFrameWithLDPC (:1079-1127) on the transmit side.
// All codewords of all frames go through ONE launch of the LDPC kernel (the reference decodes CW1+ only once CW0 parses; decoding them
// anyway changes no result, the counters below follow the reference's control flow); a second kernel — one warp per frame — parses
// the header, applies the "enough codewords / all decoded" rules and reassembles the frame bytes.  Byte work, HBM-trivial.
#include <algorithm>
#include <cstring>
#include <vector>

#include "pu_internal.h"

namespace pu {

__host__ __device__ inline uint16_t frame_crc16(const uint8_t* d, int len) {   // ControlFrame::calculateCRC, frame_v2.cpp:111-124
    uint16_t crc = 0xFFFF;
    for (int i = 0; i < len; ++i) {
        crc ^= static_cast<uint16_t>(static_cast<uint16_t>(d[i]) << 8);
        for (int j = 0; j < 8; ++j) crc = (crc & 0x8000) ? static_cast<uint16_t>((crc << 1) ^ 0x1021) : static_cast<uint16_t>(crc << 1);
    }
    return crc;
}

inline int frame_bytes_per_codeword(int rate) {   // getBytesPerCodeword, frame_v2.hpp:551-566
    switch (rate) {
        case PU_RATE_1_4: return 162 / 8;
        case PU_RATE_1_3: return 216 / 8;
        case PU_RATE_1_2: return 324 / 8;
        case PU_RATE_2_3: return 432 / 8;
        case PU_RATE_3_4: return 486 / 8;
        case PU_RATE_5_6: return 540 / 8;
        default: return 162 / 8;
    }
}

// info[b] = {success, frame_type, codewords_ok, codewords_failed, expected codewords}; frame_len[b] = bytes written to frame_out[b]
__global__ void __launch_bounds__(128) frame_assemble_kernel(const uint8_t* __restrict__ cw_bytes, size_t cw_stride, const uint8_t* __restrict__ cw_ok,
                                                             int B, int ncw, int bpc, int kbytes, uint8_t* __restrict__ frame_out,
                                                             size_t frame_cap, int32_t* __restrict__ frame_len, int32_t* __restrict__ info) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const uint8_t* cw = cw_bytes + static_cast<size_t>(b) * ncw * cw_stride;
    const uint8_t* ok = cw_ok + static_cast<size_t>(b) * ncw;
    int success = 0, type = 0, n_ok = 0, n_fail = 0, expected = 0, out_len = 0;
    do {
        if (ncw < 1) break;                                              // soft_bits.size() < LDPC_BLOCK (:352-354)
        if (!ok[0] || kbytes < bpc) { n_fail = 1; break; }               // CW0 failed (:370-375)
        n_ok = 1;
        // parseHeader on lane 0, broadcast
        int valid = 0, is_control = 0, payload_len = 0;
        if (lane == 0) {
            const uint8_t* d = cw;
            if (((d[0] << 8) | d[1]) == 0x554C) {
                type = d[2];
                is_control = type == 0x10 || type == 0x11 || type == 0x16 || type == 0x17 || type == 0x20 || type == 0x21 || type == 0x40;
                if (is_control) {
                    valid = ((d[18] << 8) | d[19]) == frame_crc16(d, 18);
                    expected = 1;
                } else {
                    expected = d[12];
                    payload_len = (d[13] << 8) | d[14];
                    valid = ((d[15] << 8) | d[16]) == frame_crc16(d, 15);
                }
            }
        }
        valid = __shfl_sync(0xffffffffu, valid, 0);
        if (!valid) { type = 0; expected = 0; break; }                   // invalid header (:380-384)
        type = __shfl_sync(0xffffffffu, type, 0);
        is_control = __shfl_sync(0xffffffffu, is_control, 0);
        expected = __shfl_sync(0xffffffffu, expected, 0);
        payload_len = __shfl_sync(0xffffffffu, payload_len, 0);
        if (ncw < expected) break;                                       // waiting for more codewords (:394-401)
        for (int i0 = 1; i0 < expected; i0 += 32) {                      // CW1+ (:410-426)
            const int i = i0 + lane;
            const unsigned good = __ballot_sync(0xffffffffu, i < expected && ok[i]);
            const unsigned bad = __ballot_sync(0xffffffffu, i < expected && !ok[i]);
            n_ok += __popc(good);
            n_fail += __popc(bad);
        }
        if (n_fail) break;                                               // allSuccess (:429)
        success = 1;
        if (expected == 0) break;                                        // (undefined in the reference: it writes decoded[0] of an empty vector)
        // CodewordStatus::reassemble -> reassembleCodewords (frame_v2.cpp:952-982,1023-1044)
        const int expected_size = is_control ? 20 : 17 + payload_len + 2;
        uint8_t* out = frame_out + static_cast<size_t>(b) * frame_cap;
        int pos = 0;
        for (int i = 0; i < expected && pos < expected_size; ++i) {
            const uint8_t* src = cw + static_cast<size_t>(i) * cw_stride;
            int avail = bpc;
            if (i > 0 && src[0] == 0xD5) { src += 2; avail = bpc - 2; }  // marker + index skipped; otherwise the legacy fallback copies all
            const int take = min(expected_size - pos, avail);
            for (int j = lane; j < take; j += 32)
                if (static_cast<size_t>(pos + j) < frame_cap) out[pos + j] = src[j];
            pos += take;
        }
        out_len = pos;
    } while (false);
    if (lane == 0) {
        frame_len[b] = out_len;
        int32_t* o = info + static_cast<size_t>(b) * 5;
        o[0] = success; o[1] = type; o[2] = n_ok; o[3] = n_fail; o[4] = expected;
    }
}

}  // namespace pu

extern "C" {

pu_status pu_frame_decode_batch(pu_ctx* ctx, pu_ldpc* dec, const float* llr, size_t B, size_t num_codewords, uint8_t* frame_out,
                                size_t frame_cap, int32_t* frame_len, int32_t* info, pu_memspace space, void* stream) {
    PU_REQUIRE(ctx && dec, "pu_frame_decode_batch: NULL handle");
    if (B == 0) return PU_OK;
    PU_REQUIRE(llr && frame_out && frame_len && info, "pu_frame_decode_batch: NULL data pointer");
    PU_REQUIRE(num_codewords >= 1 && num_codewords <= 255 && frame_cap > 0 && B * num_codewords < (1u << 30), "pu_frame_decode_batch: bad size");
    PU_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = pu::pick_stream(ctx, stream, space);
    const int rate = pu_ldpc_rate(dec), kbytes = (pu_ldpc_info_bits(dec) + 7) / 8, bpc = pu::frame_bytes_per_codeword(rate);
    const size_t ncw = B * num_codewords;
    // grow-only scratch on the context (the calling convention is one thread per context, as for the reference's objects)
    pu_status s;
    const float* d_llr = llr;
    uint8_t* d_out = frame_out;
    int32_t* d_len = frame_len;
    int32_t* d_info = info;
    const size_t out_off_len = ((B * frame_cap + 15) / 16) * 16, out_off_info = out_off_len + B * sizeof(int32_t);
    if (space == PU_MEM_HOST) {
        if ((s = ctx->f_llr.reserve(ncw * PU_LDPC_N * sizeof(float))) != PU_OK) return s;
        if ((s = ctx->f_out.reserve(out_off_info + B * 5 * sizeof(int32_t))) != PU_OK) return s;
        PU_CUDA_TRY(cudaMemcpyAsync(ctx->f_llr.ptr, llr, ncw * PU_LDPC_N * sizeof(float), cudaMemcpyHostToDevice, st));
        PU_CUDA_TRY(cudaMemsetAsync(ctx->f_out.ptr, 0, B * frame_cap, st));
        d_llr = static_cast<const float*>(ctx->f_llr.ptr);
        d_out = static_cast<uint8_t*>(ctx->f_out.ptr);
        d_len = reinterpret_cast<int32_t*>(d_out + out_off_len);
        d_info = reinterpret_cast<int32_t*>(d_out + out_off_info);
    }
    if ((s = ctx->f_bytes.reserve(ncw * static_cast<size_t>(kbytes))) != PU_OK) return s;
    if ((s = ctx->f_ok.reserve(ncw)) != PU_OK) return s;
    if ((s = ctx->f_iters.reserve(ncw * sizeof(int32_t))) != PU_OK) return s;
    uint8_t* d_bytes = static_cast<uint8_t*>(ctx->f_bytes.ptr);
    uint8_t* d_ok = static_cast<uint8_t*>(ctx->f_ok.ptr);
    s = pu_ldpc_decode_batch(dec, d_llr, PU_LDPC_N, ncw, d_bytes, static_cast<size_t>(kbytes), d_ok, static_cast<int32_t*>(ctx->f_iters.ptr),
                             PU_MEM_DEVICE, st);
    if (s != PU_OK) return s;
    const int warps = 4;
    (void)cudaGetLastError();
    pu::frame_assemble_kernel<<<static_cast<unsigned>((B + warps - 1) / warps), warps * 32, 0, st>>>(
        d_bytes, static_cast<size_t>(kbytes), d_ok, static_cast<int>(B),
        static_cast<int>(num_codewords), bpc, kbytes, d_out, frame_cap, d_len, d_info);
    ctx->launches.fetch_add(1);
    PU_CUDA_TRY(cudaGetLastError());
    if (space == PU_MEM_HOST) {
        PU_CUDA_TRY(cudaMemcpyAsync(frame_out, d_out, B * frame_cap, cudaMemcpyDeviceToHost, st));
        PU_CUDA_TRY(cudaMemcpyAsync(frame_len, d_len, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        PU_CUDA_TRY(cudaMemcpyAsync(info, d_info, B * 5 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    }
    PU_CUDA_TRY(cudaStreamSynchronize(st));           // results are complete on return (the scratch is reused by the next call)
    return PU_OK;
}

// v2::encodeFrameWithLDPC (frame_v2.cpp:1079-1127), host side: CW0 = the first bytes_per_cw frame bytes, CW1+ = {0xD5, index, payload},
// zero-padded, each LDPC-encoded to 81 bytes.  out == NULL queries the codeword count.
pu_status pu_frame_encode(int code_rate, const uint8_t* frame, size_t n_bytes, uint8_t* out, size_t out_cap, size_t* n_codewords) {
    PU_REQUIRE(n_codewords && (frame || n_bytes == 0), "pu_frame_encode: NULL argument");
    const size_t bpc = static_cast<size_t>(pu::frame_bytes_per_codeword(code_rate)), pay = bpc - 2;
    const size_t ncw = n_bytes <= bpc ? 1 : 1 + (n_bytes - bpc + pay - 1) / pay;
    *n_codewords = ncw;
    if (!out) return PU_OK;
    PU_REQUIRE(out_cap >= ncw * 81, "pu_frame_encode: output buffer too small");
    PU_REQUIRE(ncw <= 255, "pu_frame_encode: frame needs more than 255 codewords");
    std::vector<uint8_t> chunk(bpc);
    size_t offset = 0;
    for (size_t i = 0; i < ncw; ++i) {
        std::fill(chunk.begin(), chunk.end(), 0);
        if (i == 0) {
            std::memcpy(chunk.data(), frame, std::min(bpc, n_bytes));
            offset = bpc;
        } else {
            chunk[0] = 0xD5;
            chunk[1] = static_cast<uint8_t>(i);
            std::memcpy(chunk.data() + 2, frame + offset, std::min(pay, n_bytes - offset));
            offset += pay;
        }
        size_t len = 0;
        const pu_status s = pu_ldpc_encode(code_rate, chunk.data(), bpc, out + i * 81, 81, &len);
        if (s != PU_OK) return s;
    }
    return PU_OK;
}

}  // extern "C"
