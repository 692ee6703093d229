// oracle/ref_build/ref_frame.cpp — TEST INFRASTRUCTURE, not product code.
//
// extern "C" shim over the UNMODIFIED reference's protocol-v2 codeword framing (src/protocol/frame_v2.cpp, compiled where it lies):
// DataFrame / ControlFrame serialisation, encodeFrameWithLDPC, parseHeader, decodeSingleCodeword, CodewordStatus::reassemble, driven
// with the glue of RxPipeline::decodeFrame (src/gui/modem/rx_pipeline.cpp:348-445; the class itself pulls the GUI modem in).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <fcntl.h>
#include <unistd.h>

#include "ultra/types.hpp"
#include "ultra/fec.hpp"
#include "protocol/frame_v2.hpp"

using namespace ultra;
namespace v2 = ultra::protocol::v2;

extern "C" {

// DataFrame::makeData(src, dst, seq, payload, rate).serialize(): header (17) + payload + frame CRC (2)
long ref_data_frame_serialize(int rate, const char* src, const char* dst, int seq, const uint8_t* payload, size_t len, uint8_t* out, size_t cap) {
    v2::DataFrame f = v2::DataFrame::makeData(src, dst, static_cast<uint16_t>(seq), Bytes(payload, payload + len), static_cast<CodeRate>(rate));
    Bytes b = f.serialize();
    if (b.size() > cap) return -static_cast<long>(b.size());
    std::memcpy(out, b.data(), b.size());
    return static_cast<long>(b.size());
}

// ControlFrame::makeProbe(src, dst).serialize(): 20 bytes
long ref_control_frame_serialize(const char* src, const char* dst, uint8_t* out, size_t cap) {
    Bytes b = v2::ControlFrame::makeProbe(src, dst).serialize();
    if (b.size() > cap) return -static_cast<long>(b.size());
    std::memcpy(out, b.data(), b.size());
    return static_cast<long>(b.size());
}

// encodeFrameWithLDPC(frame, rate): returns the codeword count; out = 81 bytes per codeword
long ref_frame_encode(int rate, const uint8_t* frame, size_t len, uint8_t* out, size_t cap) {
    std::vector<Bytes> cws = v2::encodeFrameWithLDPC(Bytes(frame, frame + len), static_cast<CodeRate>(rate));
    size_t off = 0;
    for (const Bytes& c : cws) {
        if (off + c.size() > cap) return -1;
        std::memcpy(out + off, c.data(), c.size());
        off += c.size();
    }
    return static_cast<long>(cws.size());
}

// RxPipeline::decodeFrame(soft_bits, num_codewords) at frame rate `rate`:
// info[5] = {success, frame_type, codewords_ok, codewords_failed, expected codewords (0 = header not reached)}; returns frame_data size
long ref_frame_decode(int rate, const float* soft, size_t n_soft, int num_codewords, uint8_t* out, size_t cap, int32_t* info) {
    constexpr size_t LDPC_BLOCK = v2::LDPC_CODEWORD_BITS;
    info[0] = info[1] = info[2] = info[3] = info[4] = 0;
    if (n_soft < LDPC_BLOCK) return 0;
    const CodeRate frame_rate = static_cast<CodeRate>(rate);
    std::vector<float> cw0_bits(soft, soft + LDPC_BLOCK);
    auto [ok0, cw0_data] = v2::decodeSingleCodeword(cw0_bits, frame_rate);     // == RxPipeline::decodeSingleCodeword (:499-511)
    if (!ok0) { info[3]++; return 0; }
    info[2]++;
    v2::HeaderInfo header;                                                       // RxPipeline::parseHeader (:513-527)
    if (v2::identifyCodeword(cw0_data).type == v2::CodewordType::HEADER) header = v2::parseHeader(cw0_data);
    if (!header.valid) return 0;
    info[1] = static_cast<int>(header.type);
    const int expected = header.total_cw;
    info[4] = expected;
    if (num_codewords < expected) return 0;                                      // waiting for more codewords
    v2::CodewordStatus st;
    st.decoded.resize(expected, false);
    st.data.resize(expected);
    st.decoded[0] = true;
    st.data[0] = cw0_data;
    for (int i = 1; i < expected; i++) {
        std::vector<float> cw_bits(soft + i * LDPC_BLOCK, soft + (i + 1) * LDPC_BLOCK);
        auto [ok, data] = v2::decodeSingleCodeword(cw_bits, frame_rate);
        if (ok) { st.decoded[i] = true; st.data[i] = data; info[2]++; }
        else info[3]++;
    }
    if (!st.allSuccess()) return 0;
    info[0] = 1;
    Bytes frame = st.reassemble();
    if (frame.size() > cap) return -static_cast<long>(frame.size());
    std::memcpy(out, frame.data(), frame.size());
    return static_cast<long>(frame.size());
}

}  // extern "C"
