// oracle/ref_build/ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
//
// Thin extern "C" shim over the UNMODIFIED reference (secup/ProjectUltra) compiled from
// the sources where they lie under /root/reference (see Makefile in this directory).
// Output: oracle/_ref/libpu_ref.so.  Used only by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs, as the checker / CPU baseline.
//
// No reference source is copied here: this file only *calls* the reference's public
// classes (ultra::LDPCEncoder/LDPCDecoder/ChannelInterleaver/Interleaver, include/ultra/fec.hpp;
// ultra::OFDMModulator/OFDMDemodulator, include/ultra/ofdm.hpp; ultra::FFT/NCO,
// include/ultra/dsp.hpp; ultra::sim::WattersonChannel, src/sim/hf_channel.hpp;
// ultra::soft_demap::*, src/ofdm/soft_demap.hpp; DPSK classes, src/psk/dpsk.hpp).
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <random>
#include <span>
#include <string>
#include <vector>
#include <fcntl.h>
#include <unistd.h>

#include "ultra/types.hpp"
#include "ultra/fec.hpp"
#include "ultra/dsp.hpp"
#include "ultra/logging.hpp"
#include "ultra/ofdm.hpp"
#include "ofdm/soft_demap.hpp"
#include "sim/hf_channel.hpp"
#include "psk/dpsk.hpp"
#include "sync/chirp_sync.hpp"

using namespace ultra;

extern "C" {

// POD mirror of ultra::ModemConfig (include/ultra/types.hpp:139-234); identical layout to
// pu_modem_config in include/pu/pu_capi.h so tests can pass one ctypes struct to both.
struct ref_modem_config {
    uint32_t sample_rate;
    uint32_t center_freq;
    uint32_t fft_size;
    uint32_t num_carriers;
    uint32_t cp_mode;        // 0 SHORT, 1 MEDIUM, 2 LONG
    uint32_t symbol_guard;
    uint32_t pilot_spacing;
    uint32_t use_pilots;
    uint32_t modulation;     // ultra::Modulation value
    uint32_t code_rate;      // ultra::CodeRate value
    float output_scale;
    float tx_cfo_hz;
};

}  // extern "C"

namespace {

ModemConfig to_cfg(const ref_modem_config* c) {
    ModemConfig m;
    m.sample_rate = c->sample_rate;
    m.center_freq = c->center_freq;
    m.fft_size = c->fft_size;
    m.num_carriers = c->num_carriers;
    m.cp_mode = static_cast<CyclicPrefixMode>(c->cp_mode);
    m.symbol_guard = c->symbol_guard;
    m.pilot_spacing = c->pilot_spacing;
    m.use_pilots = c->use_pilots != 0;
    m.modulation = static_cast<Modulation>(c->modulation);
    m.code_rate = static_cast<CodeRate>(c->code_rate);
    m.output_scale = c->output_scale;
    m.tx_cfo_hz = c->tx_cfo_hz;
    return m;
}

// The reference prints unconditionally to stderr on the hot path
// (channel_equalizer.cpp:27-32,96-100,...); silence fd 2 around calls.
struct StderrSilencer {
    int saved = -1;
    StderrSilencer() {
        fflush(stderr);
        saved = dup(2);
        int nul = open("/dev/null", O_WRONLY);
        if (nul >= 0) { dup2(nul, 2); close(nul); }
    }
    ~StderrSilencer() {
        fflush(stderr);
        if (saved >= 0) { dup2(saved, 2); close(saved); }
    }
};

std::once_flag g_init;
void init_once() {
    std::call_once(g_init, [] { setLogLevel(LogLevel::ERROR); });
}

}  // namespace

extern "C" {

int ref_abi_version() { return 1; }

// ---------------------------------------------------------------- LDPC
// LDPCEncoder::encode (src/fec/ldpc_encoder.cpp:193-257)
long ref_ldpc_encode(int rate, const uint8_t* data, size_t len, uint8_t* out, size_t cap) {
    init_once();
    LDPCEncoder enc(static_cast<CodeRate>(rate));
    Bytes r = enc.encode(ByteSpan(data, len));
    if (r.size() > cap) return -1;
    std::memcpy(out, r.data(), r.size());
    return (long)r.size();
}

// LDPCDecoder::decodeSoft (src/fec/ldpc_decoder.cpp:283-428)
long ref_ldpc_decode_soft(int rate, int max_iter, const float* llr, size_t n,
                          uint8_t* out, size_t cap, int* ok, int* iters) {
    init_once();
    LDPCDecoder dec(static_cast<CodeRate>(rate));
    if (max_iter >= 0) dec.setMaxIterations(max_iter);
    Bytes r = dec.decodeSoft(std::span<const float>(llr, n));
    if (ok) *ok = dec.lastDecodeSuccess() ? 1 : 0;
    if (iters) *iters = dec.lastIterations();
    if (r.size() > cap) return -1;
    if (!r.empty()) std::memcpy(out, r.data(), r.size());
    return (long)r.size();
}

// LDPCDecoder::decode (hard bits, src/fec/ldpc_decoder.cpp:267-281)
long ref_ldpc_decode_hard(int rate, const uint8_t* coded, size_t len,
                          uint8_t* out, size_t cap, int* ok, int* iters) {
    init_once();
    LDPCDecoder dec(static_cast<CodeRate>(rate));
    Bytes r = dec.decode(ByteSpan(coded, len));
    if (ok) *ok = dec.lastDecodeSuccess() ? 1 : 0;
    if (iters) *iters = dec.lastIterations();
    if (r.size() > cap) return -1;
    if (!r.empty()) std::memcpy(out, r.data(), r.size());
    return (long)r.size();
}

// B independent 648-LLR codewords through one decoder object (what the Monte-Carlo tools do,
// tools/test_mode_snr.cpp:35,87).  out is [B][out_stride].
int ref_ldpc_decode_batch(int rate, int max_iter, const float* llr, size_t B,
                          uint8_t* out, size_t out_stride, uint8_t* ok, int32_t* iters) {
    init_once();
    LDPCDecoder dec(static_cast<CodeRate>(rate));
    if (max_iter >= 0) dec.setMaxIterations(max_iter);
    for (size_t b = 0; b < B; ++b) {
        Bytes r = dec.decodeSoft(std::span<const float>(llr + b * 648, 648));
        if (r.size() > out_stride) return -1;
        std::memcpy(out + b * out_stride, r.data(), r.size());
        ok[b] = dec.lastDecodeSuccess() ? 1 : 0;
        iters[b] = dec.lastIterations();
    }
    return 0;
}

// ChannelInterleaver (src/fec/ldpc_decoder.cpp:574-620)
int ref_channel_interleave(size_t bps, size_t total, const float* in, size_t n, float* out, int inverse) {
    ChannelInterleaver ci(bps, total);
    std::vector<float> r = inverse ? ci.deinterleave(std::span<const float>(in, n))
                                   : ci.interleave(std::span<const float>(in, n));
    std::memcpy(out, r.data(), r.size() * sizeof(float));
    return (int)r.size();
}

int ref_channel_interleave_bytes(size_t bps, size_t total, const uint8_t* in, size_t n, uint8_t* out, int inverse) {
    ChannelInterleaver ci(bps, total);
    Bytes r = inverse ? ci.deinterleave(ByteSpan(in, n)) : ci.interleave(ByteSpan(in, n));
    std::memcpy(out, r.data(), r.size());
    return (int)r.size();
}

// Interleaver (src/fec/ldpc_decoder.cpp:454-540)
int ref_block_interleave(size_t rows, size_t cols, const float* in, size_t n, float* out, int inverse) {
    Interleaver il(rows, cols);
    std::vector<float> r = inverse ? il.deinterleave(std::span<const float>(in, n))
                                   : il.interleave(std::span<const float>(in, n));
    std::memcpy(out, r.data(), r.size() * sizeof(float));
    return (int)r.size();
}

int ref_block_interleave_bytes(size_t rows, size_t cols, const uint8_t* in, size_t n, uint8_t* out, int inverse) {
    Interleaver il(rows, cols);
    Bytes r = inverse ? il.deinterleave(ByteSpan(in, n)) : il.interleave(ByteSpan(in, n));
    std::memcpy(out, r.data(), r.size());
    return (int)r.size();
}

// ---------------------------------------------------------------- DSP primitives
// FFT::forward / inverse (src/dsp/fft.cpp:124-168), interleaved re/im
int ref_fft(size_t n, const float* in, float* out, int inverse) {
    FFT f(n);
    if (inverse) f.inverse(reinterpret_cast<const Complex*>(in), reinterpret_cast<Complex*>(out));
    else f.forward(reinterpret_cast<const Complex*>(in), reinterpret_cast<Complex*>(out));
    return 0;
}

// NCO::next (src/dsp/filters.cpp:232-238), interleaved (cos, sin)
int ref_nco(float freq, float fs, size_t n, float* out) {
    NCO nco(freq, fs);
    for (size_t i = 0; i < n; ++i) {
        Complex c = nco.next();
        out[2 * i] = c.real();
        out[2 * i + 1] = c.imag();
    }
    return 0;
}

// soft_demap::* (src/ofdm/soft_demap.hpp) for one symbol; returns number of LLRs.
int ref_soft_demap(int mod, float re, float im, float pre, float pim, float nv, float* out) {
    Complex s(re, im), p(pre, pim);
    switch (static_cast<Modulation>(mod)) {
        case Modulation::DBPSK: out[0] = soft_demap::demapDBPSK(s, p, nv); return 1;
        case Modulation::BPSK: out[0] = soft_demap::demapBPSK(s, nv); return 1;
        case Modulation::DQPSK: { auto r = soft_demap::demapDQPSK(s, p, nv); out[0] = r[0]; out[1] = r[1]; return 2; }
        case Modulation::QPSK: { auto r = soft_demap::demapQPSK(s, nv); out[0] = r[0]; out[1] = r[1]; return 2; }
        case Modulation::D8PSK: { auto r = soft_demap::demapD8PSK(s, p, nv); for (int i = 0; i < 3; ++i) out[i] = r[i]; return 3; }
        case Modulation::QAM16: { auto r = soft_demap::demapQAM16(s, nv); for (int i = 0; i < 4; ++i) out[i] = r[i]; return 4; }
        case Modulation::QAM32: { auto r = soft_demap::demapQAM32(s, nv); for (int i = 0; i < 5; ++i) out[i] = r[i]; return 5; }
        case Modulation::QAM64: { auto r = soft_demap::demapQAM64(s, nv); for (int i = 0; i < 6; ++i) out[i] = r[i]; return 6; }
        case Modulation::QAM256: { auto r = soft_demap::demapQAM256(s, nv); for (int i = 0; i < 8; ++i) out[i] = r[i]; return 8; }
        default: return -1;
    }
}

// ---------------------------------------------------------------- OFDM TX
// layout 0: generateTrainingSymbols(2) + modulate  (the presynced layout, tools/test_ofdm_chirp_pilots.cpp:183-191)
// layout 1: generatePreamble() + modulate          (the Schmidl-Cox layout, tools/test_mode_snr.cpp:47-52)
long ref_ofdm_tx(const ref_modem_config* c, int layout, const uint8_t* data, size_t len,
                 float* out, size_t cap) {
    init_once();
    StderrSilencer quiet;
    ModemConfig cfg = to_cfg(c);
    OFDMModulator mod(cfg);
    Samples head = layout == 0 ? mod.generateTrainingSymbols(2) : mod.generatePreamble();
    Samples body = mod.modulate(ByteSpan(data, len), cfg.modulation);
    size_t total = head.size() + body.size();
    if (total > cap) return -(long)total;
    std::memcpy(out, head.data(), head.size() * sizeof(float));
    std::memcpy(out + head.size(), body.data(), body.size() * sizeof(float));
    return (long)total;
}

// ---------------------------------------------------------------- OFDM RX (presynced path)
// Oracle recipe, SURVEY §8(c) == OFDMChirpWaveform::process (src/waveform/ofdm_chirp_waveform.cpp:185-199):
//   reset(); setFrequencyOffset[WithPhase](); processPresynced(span, training); drain getSoftBits().
// cfo_mode: 0 = do not set (reference then runs estimateCFOFromTraining), 1 = setFrequencyOffset(cfo),
//           2 = setFrequencyOffsetWithPhase(cfo, phase).
// Returns the number of soft bits (all of them, in order), or -needed if cap is too small.
long ref_ofdm_presynced(const ref_modem_config* c, const float* samples, size_t L, int training,
                        int cfo_mode, float cfo_hz, float cfo_phase,
                        float* llr_out, size_t cap, float* snr_db, float* final_cfo) {
    init_once();
    StderrSilencer quiet;
    ModemConfig cfg = to_cfg(c);
    OFDMDemodulator d(cfg);
    d.reset();
    if (cfo_mode == 1) d.setFrequencyOffset(cfo_hz);
    else if (cfo_mode == 2) d.setFrequencyOffsetWithPhase(cfo_hz, cfo_phase);
    d.processPresynced(SampleSpan(samples, L), training);
    if (snr_db) *snr_db = d.getEstimatedSNR();
    if (final_cfo) *final_cfo = d.getFrequencyOffset();
    std::vector<float> all;
    for (;;) {
        std::vector<float> chunk = d.getSoftBits();
        if (chunk.empty()) break;
        all.insert(all.end(), chunk.begin(), chunk.end());
    }
    if (all.size() > cap) return -(long)all.size();
    std::memcpy(llr_out, all.data(), all.size() * sizeof(float));
    return (long)all.size();
}

// Many frames back to back (B frames of L samples); llr_out is [B][n_llr_stride]; counts[b] = soft bits produced.
int ref_ofdm_presynced_batch(const ref_modem_config* c, const float* samples, size_t B, size_t L, int training,
                             int cfo_mode, const float* cfo_hz, const float* cfo_phase,
                             float* llr_out, size_t stride, int32_t* counts) {
    init_once();
    StderrSilencer quiet;
    ModemConfig cfg = to_cfg(c);
    for (size_t b = 0; b < B; ++b) {
        OFDMDemodulator d(cfg);   // fresh object per frame, like tools/test_ofdm_chirp_pilots.cpp:169
        d.reset();
        float f = cfo_hz ? cfo_hz[b] : 0.0f, p = cfo_phase ? cfo_phase[b] : 0.0f;
        if (cfo_mode == 1) d.setFrequencyOffset(f);
        else if (cfo_mode == 2) d.setFrequencyOffsetWithPhase(f, p);
        d.processPresynced(SampleSpan(samples + b * L, L), training);
        size_t n = 0;
        for (;;) {
            std::vector<float> chunk = d.getSoftBits();
            if (chunk.empty()) break;
            size_t take = std::min(chunk.size(), stride - n);
            std::memcpy(llr_out + b * stride + n, chunk.data(), take * sizeof(float));
            n += take;
            if (n >= stride) break;
        }
        counts[b] = (int32_t)n;
    }
    return 0;
}

// Schmidl-Cox path exactly as tools/test_mode_snr.cpp:65-70: feed `chunk`-sample pieces to process(),
// then one getSoftBits() (<= 648).  Returns number of soft bits.
long ref_ofdm_process(const ref_modem_config* c, const float* samples, size_t L, size_t chunk,
                      float* llr_out, size_t cap, int* synced, float* snr_db) {
    init_once();
    StderrSilencer quiet;
    ModemConfig cfg = to_cfg(c);
    OFDMDemodulator d(cfg);
    for (size_t i = 0; i < L; i += chunk) {
        size_t len = std::min(chunk, L - i);
        d.process(SampleSpan(samples + i, len));
    }
    if (snr_db) *snr_db = d.getEstimatedSNR();
    std::vector<float> soft = d.getSoftBits();
    if (synced) *synced = d.isSynced() ? 1 : 0;
    if (soft.size() > cap) return -(long)soft.size();
    std::memcpy(llr_out, soft.data(), soft.size() * sizeof(float));
    return (long)soft.size();
}

// Same call sequence, also reporting what acquisition decided: getLastSyncOffset() and getFrequencyOffset() (the
// Schmidl-Cox coarse CFO; no tracker changes it in the no-pilot differential modes) -- for the GPU acquisition tests.
long ref_ofdm_process_info(const ref_modem_config* c, const float* samples, size_t L, size_t chunk,
                           float* llr_out, size_t cap, int* synced, long* sync_offset, float* cfo_hz) {
    init_once();
    StderrSilencer quiet;
    ModemConfig cfg = to_cfg(c);
    OFDMDemodulator d(cfg);
    bool ever = false;
    long off = -1;
    float cfo = 0.0f;
    for (size_t i = 0; i < L; i += chunk) {
        size_t len = std::min(chunk, L - i);
        d.process(SampleSpan(samples + i, len));
        if (!ever && d.isSynced()) {
            ever = true;
            off = (long)d.getLastSyncOffset();
            cfo = d.getFrequencyOffset();
        }
    }
    std::vector<float> soft = d.getSoftBits();
    if (synced) *synced = ever ? 1 : 0;
    if (sync_offset) *sync_offset = off;
    if (cfo_hz) *cfo_hz = cfo;
    if (soft.size() > cap) return -(long)soft.size();
    std::memcpy(llr_out, soft.data(), soft.size() * sizeof(float));
    return (long)soft.size();
}

// ---------------------------------------------------------------- dual-chirp synchronisation (src/sync/chirp_sync.hpp)
static sync::ChirpConfig chirp_cfg(float fs, float tx_cfo) {   // OFDMChirpWaveform::getChirpConfig, ofdm_chirp_waveform.cpp:39-49
    sync::ChirpConfig c;
    c.sample_rate = fs;
    c.f_start = 300.0f;
    c.f_end = 2700.0f;
    c.duration_ms = 500.0f;
    c.gap_ms = 100.0f;
    c.use_dual_chirp = true;
    c.tx_cfo_hz = tx_cfo;
    return c;
}
// ChirpSync::generate(): [up chirp][gap][down chirp][gap]
long ref_chirp_generate(float fs, float tx_cfo, float* out, size_t cap) {
    init_once();
    StderrSilencer quiet;
    sync::ChirpSync cs(chirp_cfg(fs, tx_cfo));
    Samples s = cs.generate();
    if (s.size() > cap) return -(long)s.size();
    std::memcpy(out, s.data(), s.size() * sizeof(float));
    return (long)s.size();
}
// ChirpSync::detectDualChirp: info = {success, up_chirp_start, down_chirp_start}, f = {cfo_hz, up_correlation, down_correlation}
int ref_chirp_detect_dual(float fs, const float* x, size_t L, float threshold, int32_t* info, float* f) {
    init_once();
    StderrSilencer quiet;
    fflush(stdout);
    int saved = dup(1), nul = open("/dev/null", O_WRONLY);   // detectDualChirp printf()s on stdout
    if (nul >= 0) { dup2(nul, 1); close(nul); }
    sync::ChirpSync cs(chirp_cfg(fs, 0.0f));
    auto r = cs.detectDualChirp(SampleSpan(x, L), threshold);
    fflush(stdout);
    if (saved >= 0) { dup2(saved, 1); close(saved); }
    info[0] = r.success ? 1 : 0; info[1] = r.up_chirp_start; info[2] = r.down_chirp_start;
    f[0] = r.cfo_hz; f[1] = r.up_correlation; f[2] = r.down_correlation;
    return 0;
}
// The receive sequence of tools/test_iwaveform.cpp:127-160 on an OFDM_CHIRP frame, with the glue of
// OFDMChirpWaveform::detectSync / process (ofdm_chirp_waveform.cpp:129-199) applied to the reference's own ChirpSync and
// OFDMDemodulator objects (the waveform class itself pulls the protocol layer into the build):
//   detectDualChirp -> start_sample = down_chirp_start + chirp + gap -> setFrequencyOffsetWithPhase(cfo, accumulated phase)
//   -> processPresynced(span from start_sample, 2) -> all soft bits.
long ref_ofdm_chirp_receive(const ref_modem_config* c, const float* x, size_t L, float threshold, int32_t* info, float* cfo_out,
                            float* llr_out, size_t cap) {
    init_once();
    StderrSilencer quiet;
    fflush(stdout);
    int saved = dup(1), nul = open("/dev/null", O_WRONLY);
    if (nul >= 0) { dup2(nul, 1); close(nul); }
    ModemConfig cfg = to_cfg(c);
    sync::ChirpSync cs(chirp_cfg((float)cfg.sample_rate, 0.0f));
    auto r = cs.detectDualChirp(SampleSpan(x, L), threshold);
    fflush(stdout);
    if (saved >= 0) { dup2(saved, 1); close(saved); }
    info[0] = r.success ? 1 : 0; info[1] = r.up_chirp_start; info[2] = r.down_chirp_start; info[3] = -1;
    *cfo_out = r.cfo_hz;
    if (!r.success) return 0;
    size_t chirp_samples = cs.getChirpSamples();
    size_t gap_samples = static_cast<size_t>(cfg.sample_rate * 100.0f / 1000.0f);
    int start_sample = r.down_chirp_start + chirp_samples + gap_samples;
    info[3] = start_sample;
    if (start_sample < 0 || (size_t)start_sample >= L) return 0;            // tools/test_iwaveform.cpp:145-148
    size_t training_start_sample = (size_t)start_sample;
    float cfo_hz = r.cfo_hz;                                                  // waveform.setFrequencyOffset(sync_result.cfo_hz)
    float initial_phase_rad = -2.0f * M_PI * cfo_hz * training_start_sample / cfg.sample_rate;
    while (initial_phase_rad > M_PI) initial_phase_rad -= 2.0f * M_PI;
    while (initial_phase_rad < -M_PI) initial_phase_rad += 2.0f * M_PI;
    OFDMDemodulator d(cfg);
    d.setFrequencyOffsetWithPhase(cfo_hz, initial_phase_rad);
    bool ready = d.processPresynced(SampleSpan(x + start_sample, L - start_sample), 2);
    size_t n = 0;
    if (ready) {
        while (d.hasPendingData()) {
            auto chunk = d.getSoftBits();
            if (chunk.empty()) break;
            for (float v : chunk) { if (n < cap) llr_out[n] = v; ++n; }
        }
    }
    return (long)n;
}

// ---------------------------------------------------------------- Watterson channel
// sim::WattersonChannel(cfg, seed).process (src/sim/hf_channel.hpp:67-168).  Used for statistical
// comparison only (its mt19937+normal_distribution stream is replaced by a counter RNG in the product).
int ref_watterson(float snr_db, float delay_ms, float doppler_hz, float g1, float g2,
                  int fading, int multipath, int noise, uint32_t seed,
                  const float* in, size_t n, float* out) {
    sim::WattersonChannel::Config cc;
    cc.snr_db = snr_db;
    cc.delay_spread_ms = delay_ms;
    cc.doppler_spread_hz = doppler_hz;
    cc.path1_gain = g1;
    cc.path2_gain = g2;
    cc.fading_enabled = fading != 0;
    cc.multipath_enabled = multipath != 0;
    cc.noise_enabled = noise != 0;
    cc.cfo_enabled = false;
    sim::WattersonChannel ch(cc, seed);
    Samples r = ch.process(SampleSpan(in, n));
    std::memcpy(out, r.data(), n * sizeof(float));
    return 0;
}

// WattersonChannel with ONLY the CFO injector active (applyCFO, src/sim/hf_channel.hpp:173-232): no fading, multipath or noise, so the
// output is deterministic and pins the twin / the CUDA kernel bit for bit.
int ref_watterson_cfo(float cfo_hz, const float* in, size_t n, float* out) {
    sim::WattersonChannel::Config cc;
    cc.snr_db = 100.0f;
    cc.delay_spread_ms = 0.0f;
    cc.doppler_spread_hz = 0.0f;
    cc.cfo_hz = cfo_hz;
    cc.fading_enabled = false;
    cc.multipath_enabled = false;
    cc.noise_enabled = false;
    cc.cfo_enabled = true;
    sim::WattersonChannel ch(cc, 1);
    Samples r = ch.process(SampleSpan(in, n));
    std::memcpy(out, r.data(), n * sizeof(float));
    return 0;
}

// The tools' CFO injector (tools/test_iwaveform.cpp:67-118) lives in a program, not in the library: this is its loop around the
// reference's own FFT class (the arithmetic that matters), so that the product's restatement of both (csrc/tools_cfo.cpp) is pinned to
// compiled reference code.
int ref_tools_apply_cfo(float* x, size_t n, float cfo_hz, float fs) {
    if (n < 128 || std::abs(cfo_hz) < 0.001f) return 0;                    // the tool's early return
    size_t m = 1;
    while (m < n) m *= 2;                                                   // next power of two
    FFT engine(m);                                                          // the reference's FFT class does both transforms
    std::vector<Complex> t(m, Complex(0, 0)), f(m), z(m);
    for (size_t i = 0; i < n; ++i) t[i] = Complex(x[i], 0);
    engine.forward(t.data(), f.data());
    for (size_t k = 1; k < m / 2; ++k) f[k] *= 2.0f;                        // one-sided spectrum: analytic signal
    std::fill(f.begin() + m / 2 + 1, f.end(), Complex(0, 0));
    engine.inverse(f.data(), z.data());
    const float step = 2.0f * static_cast<float>(M_PI) * cfo_hz / fs;
    float ph = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        x[i] = std::real(z[i] * Complex(std::cos(ph), std::sin(ph)));
        ph += step;
        if (ph > M_PI) ph -= 2.0f * M_PI;                                   // float phase, double constants: as the tool wraps it
        else if (ph < -M_PI) ph += 2.0f * M_PI;
    }
    return 0;
}

// ---------------------------------------------------------------- single-carrier DPSK
// DPSKModulator / DPSKDemodulator (src/psk/dpsk.hpp).  mod: 0 DBPSK(2), 1 DQPSK(4), 2 D8PSK(8)
static DPSKConfig dpsk_cfg(int mod_order, int samples_per_symbol) {
    DPSKConfig c;
    c.modulation = mod_order == 2 ? DPSKModulation::DBPSK : mod_order == 4 ? DPSKModulation::DQPSK : DPSKModulation::D8PSK;
    c.samples_per_symbol = samples_per_symbol;
    return c;
}

long ref_dpsk_modulate(int mod_order, int sps, int with_preamble, const uint8_t* data, size_t len,
                       float* out, size_t cap) {
    init_once();
    StderrSilencer quiet;
    DPSKModulator m(dpsk_cfg(mod_order, sps));
    Samples pre;
    if (with_preamble) pre = m.generatePreamble();
    Samples body = m.modulate(ByteSpan(data, len));
    size_t total = pre.size() + body.size();
    if (total > cap) return -(long)total;
    std::memcpy(out, pre.data(), pre.size() * sizeof(float));
    std::memcpy(out + pre.size(), body.data(), body.size() * sizeof(float));
    return (long)total;
}

// findPreamble + demodulateSoft, as tools/test_dpsk_snr.cpp does.  data_start = -1: run findPreamble;
// otherwise the caller gives the genie data offset and the reference symbol is taken as the demodulator's
// default state after reset().
long ref_dpsk_demod_soft(int mod_order, int sps, const float* samples, size_t L, long data_start,
                         float* llr_out, size_t cap, long* found_at) {
    init_once();
    StderrSilencer quiet;
    DPSKDemodulator d(dpsk_cfg(mod_order, sps));
    long start = data_start;
    if (data_start < 0) {
        int off = d.findPreamble(SampleSpan(samples, L));
        if (found_at) *found_at = off;
        if (off < 0) return 0;
        start = off;
    }
    std::vector<float> soft = d.demodulateSoft(SampleSpan(samples + start, L - start));
    if (soft.size() > cap) return -(long)soft.size();
    std::memcpy(llr_out, soft.data(), soft.size() * sizeof(float));
    return (long)soft.size();
}

// ---------------------------------------------------------------- CPU baseline loops (timed by bench.py)
// Genie-timed oracle recipe + LDPC decode over B frames; returns wall seconds, fills per-frame ok flags.
double ref_time_presynced_decode(const ref_modem_config* c, const float* samples, size_t B, size_t L,
                                 int rate, uint8_t* info_out, size_t info_stride, uint8_t* ok) {
    init_once();
    StderrSilencer quiet;
    ModemConfig cfg = to_cfg(c);
    LDPCDecoder dec(static_cast<CodeRate>(rate));
    auto t0 = std::chrono::steady_clock::now();
    for (size_t b = 0; b < B; ++b) {
        OFDMDemodulator d(cfg);
        d.reset();
        d.setFrequencyOffset(0.0f);
        d.processPresynced(SampleSpan(samples + b * L, L), 2);
        std::vector<float> soft = d.getSoftBits();
        if (soft.size() >= 648) {
            Bytes r = dec.decodeSoft(std::span<const float>(soft.data(), 648));
            ok[b] = dec.lastDecodeSuccess() ? 1 : 0;
            std::memcpy(info_out + b * info_stride, r.data(), std::min(r.size(), info_stride));
        } else {
            ok[b] = 0;
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

double ref_time_ldpc_decode(int rate, int max_iter, const float* llr, size_t B,
                            uint8_t* out, size_t out_stride, uint8_t* ok, int32_t* iters) {
    init_once();
    auto t0 = std::chrono::steady_clock::now();
    ref_ldpc_decode_batch(rate, max_iter, llr, B, out, out_stride, ok, iters);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
