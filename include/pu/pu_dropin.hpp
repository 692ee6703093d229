// include/pu/pu_dropin.hpp — C++20 host classes that present ProjectUltra's own call surface over the C ABI of
// libpu_b200.so (include/pu/pu_capi.h).  Header-only; link with -lpu_b200.
//
//   pu::LDPCEncoder / pu::LDPCDecoder        <->  ultra::LDPCEncoder / ultra::LDPCDecoder   include/ultra/fec.hpp:20-77
//   pu::Interleaver / pu::ChannelInterleaver <->  ultra::Interleaver / ChannelInterleaver   include/ultra/fec.hpp:85-142
//   pu::OFDMDemodulator                      <->  ultra::OFDMDemodulator (presynced path)   include/ultra/ofdm.hpp:58-127
//   pu::OFDMModulator                        <->  ultra::OFDMModulator (host TX stimulus)   include/ultra/ofdm.hpp:24-52
//   pu::OfdmChirpWaveform                    <->  ultra::OFDMChirpWaveform                  src/waveform/ofdm_chirp_waveform.cpp:129-230
//   pu::McDpskWaveform                       <->  ultra::MCDPSKWaveform                     src/waveform/mc_dpsk_waveform.cpp:33-265
//
// Same method names, argument meaning, ownership (borrowed spans in, values out) and error behaviour as the
// reference: no exceptions on the decode path, failure = lastDecodeSuccess()==false with best-effort
// bytes, empty input -> {} and failure.  What differs: construction throws std::runtime_error when no sm_100 GPU /
// library is usable (there is NO CPU fallback).  Acquisition (Schmidl-Cox in OFDMDemodulator::process, the dual chirp in
// IWaveform::detectSync) runs on the GPU as well.
//
// Define PU_DROPIN_WITH_ULTRA before including this header (with the reference's include/ and src/ on the include
// path) to use the reference's own types (ultra::ModemConfig, Modulation, CodeRate, Bytes, ...) and to make
// pu::OfdmChirpWaveform derive from ultra::IWaveform, so reference drivers compile against either implementation.
#pragma once
#include <cmath>
#include <cstdio>
#include <complex>
#include <cstdint>
#include <cstring>
#include <memory>
#include <span>
#include <stdexcept>
#include <string>
#include <vector>

#include "pu/pu_capi.h"

#ifdef PU_DROPIN_WITH_ULTRA
#include "ultra/types.hpp"
#include "waveform/waveform_interface.hpp"
#endif

namespace pu {

#ifdef PU_DROPIN_WITH_ULTRA
using ultra::ByteSpan;
using ultra::Bytes;
using ultra::CodeRate;
using ultra::Complex;
using ultra::ModemConfig;
using ultra::Modulation;
using ultra::SampleSpan;
using ultra::Samples;
using ultra::Symbol;
#else
// stand-alone mirrors of include/ultra/types.hpp:13-23,27-39,92-101,139-234 (receive-path fields only)
using Complex = std::complex<float>;
using Symbol = std::vector<Complex>;
using Samples = std::vector<float>;
using Bytes = std::vector<uint8_t>;
using SampleSpan = std::span<const float>;
using ByteSpan = std::span<const uint8_t>;
enum class Modulation : uint8_t { DBPSK = 0, BPSK = 1, DQPSK = 2, QPSK = 3, D8PSK = 4, QAM8 = 5, QAM16 = 6, QAM32 = 7,
                                  QAM64 = 8, QAM128 = 9, QAM256 = 10, AUTO = 255 };
enum class CodeRate : uint8_t { R1_4 = 0, R1_3 = 1, R1_2 = 2, R2_3 = 3, R3_4 = 4, R5_6 = 5, R7_8 = 6, AUTO = 255 };
enum class CyclicPrefixMode : uint8_t { SHORT = 0, MEDIUM = 1, LONG = 2 };
struct ModemConfig {
    uint32_t sample_rate = 48000, center_freq = 1500, fft_size = 512, num_carriers = 30;
    CyclicPrefixMode cp_mode = CyclicPrefixMode::MEDIUM;
    uint32_t symbol_guard = 4, pilot_spacing = 2;
    bool use_pilots = true;
    Modulation modulation = Modulation::QPSK;
    CodeRate code_rate = CodeRate::R1_2;
    bool adaptive_eq_enabled = false;
    float output_scale = 40.0f, tx_cfo_hz = 0.0f;
    uint32_t getCyclicPrefix() const {
        const uint32_t base = cp_mode == CyclicPrefixMode::SHORT ? 32u : cp_mode == CyclicPrefixMode::LONG ? 64u : 48u;
        return base * (fft_size / 512);
    }
    uint32_t getSymbolDuration() const { return fft_size + getCyclicPrefix() + symbol_guard; }
};
#endif

namespace detail {

[[noreturn]] inline void fail(const char* what, pu_status s) {
    throw std::runtime_error(std::string(what) + ": " + pu_status_string(s) + ": " + pu_last_error());
}

// One context per process and device, created on first use and kept for the life of the process (the reference's
// objects need no context; this keeps their constructors' signatures).
inline pu_ctx* shared_context(int device = -1) {
    static pu_ctx* ctx = nullptr;
    static int dev = 0;
    if (!ctx) {
        if (device >= 0) dev = device;
        const pu_status s = pu_init(dev, &ctx);
        if (s != PU_OK) fail("pu_init", s);
    }
    return ctx;
}

inline pu_modem_config to_pod(const ModemConfig& c) {
    pu_modem_config p{};
    p.sample_rate = c.sample_rate;
    p.center_freq = c.center_freq;
    p.fft_size = c.fft_size;
    p.num_carriers = c.num_carriers;
    p.cp_mode = static_cast<uint32_t>(c.cp_mode);
    p.symbol_guard = c.symbol_guard;
    p.pilot_spacing = c.pilot_spacing;
    p.use_pilots = c.use_pilots ? 1u : 0u;
    p.modulation = static_cast<uint32_t>(c.modulation);
    p.code_rate = static_cast<uint32_t>(c.code_rate);
    p.output_scale = c.output_scale;
    p.tx_cfo_hz = c.tx_cfo_hz;
    return p;
}

}  // namespace detail

// Selects the CUDA device used by every drop-in object of this process; call before constructing the first one.
inline void use_device(int device) { (void)detail::shared_context(device); }

// ------------------------------------------------------------------------------------------------ FEC
class LDPCEncoder {   // ultra::LDPCEncoder, include/ultra/fec.hpp:20-41
public:
    explicit LDPCEncoder(CodeRate rate) : rate_(rate) {}
    Bytes encode(ByteSpan data) {   // src/fec/ldpc_encoder.cpp:193-257
        size_t n = 0;
        Bytes out(getCodedSize(data.size()) + 81);
        const pu_status s = pu_ldpc_encode(static_cast<int>(rate_), data.data(), data.size(), out.data(), out.size(), &n);
        if (s != PU_OK) return {};
        out.resize(n);
        return out;
    }
    size_t getCodedSize(size_t input_size) const {   // ldpc_encoder.cpp:259-264: whole 648-bit blocks
        static const int k_of[] = {162, 324, 324, 432, 486, 540, 324};
        const size_t k = static_cast<size_t>(k_of[static_cast<unsigned>(rate_) < 7 ? static_cast<unsigned>(rate_) : 2]);
        const size_t blocks = (input_size * 8 + k - 1) / k;
        return (blocks * PU_LDPC_N + 7) / 8;
    }
    CodeRate getRate() const { return rate_; }
    void setRate(CodeRate rate) { rate_ = rate; }

private:
    CodeRate rate_;
};

class LDPCDecoder {   // ultra::LDPCDecoder, include/ultra/fec.hpp:48-77
public:
    explicit LDPCDecoder(CodeRate rate) {
        const pu_status s = pu_ldpc_create(detail::shared_context(), static_cast<int>(rate), -1, &h_);
        if (s != PU_OK) detail::fail("pu_ldpc_create", s);
    }
    ~LDPCDecoder() { pu_ldpc_destroy(h_); }
    LDPCDecoder(const LDPCDecoder&) = delete;
    LDPCDecoder& operator=(const LDPCDecoder&) = delete;

    Bytes decode(ByteSpan coded_data) {   // src/fec/ldpc_decoder.cpp:267-281
        Bytes out(coded_data.size() + 128);
        size_t n = 0;
        if (pu_ldpc_decode_hard(h_, coded_data.data(), coded_data.size(), out.data(), out.size(), &n, &ok_, &iters_) != PU_OK) {
            ok_ = 0;
            return {};
        }
        out.resize(n);
        return out;
    }
    Bytes decodeSoft(std::span<const float> llrs) {   // ldpc_decoder.cpp:283-428
        Bytes out(llrs.size() / 8 + 128);
        size_t n = 0;
        if (pu_ldpc_decode_soft(h_, llrs.data(), llrs.size(), out.data(), out.size(), &n, &ok_, &iters_) != PU_OK) {
            ok_ = 0;
            return {};
        }
        out.resize(n);
        return out;
    }
    bool lastDecodeSuccess() const { return ok_ != 0; }
    int lastIterations() const { return iters_; }
    void setRate(CodeRate rate) { (void)pu_ldpc_set_rate(h_, static_cast<int>(rate)); }
    CodeRate getRate() const { return static_cast<CodeRate>(pu_ldpc_rate(h_)); }
    void setMaxIterations(int max_iter) { (void)pu_ldpc_set_max_iterations(h_, max_iter); }
    pu_ldpc* handle() const { return h_; }   // for the batched entry points

private:
    pu_ldpc* h_ = nullptr;
    int ok_ = 0, iters_ = 0;
};

namespace detail {
template <class T>
inline std::vector<T> permute_fwd(const std::vector<uint32_t>& perm, std::span<const T> in) {
    // out[perm[i]] = in[i]; inputs shorter than the table are zero-padded (ldpc_decoder.cpp:466-481, :587-600)
    std::vector<T> out(perm.size(), T(0));
    for (size_t i = 0; i < perm.size() && i < in.size(); ++i) out[perm[i]] = in[i];
    return out;
}
inline Bytes permute_bits(const std::vector<uint32_t>& perm, ByteSpan data) {
    // bit i of the input (MSB-first) goes to bit perm[i] of the output (ldpc_decoder.cpp:483-510, :622-672)
    const size_t total = perm.size();
    Bytes out((total + 7) / 8, 0);
    for (size_t i = 0; i < total && i < data.size() * 8; ++i)
        if ((data[i >> 3] >> (7 - (i & 7))) & 1) out[perm[i] >> 3] |= static_cast<uint8_t>(1u << (7 - (perm[i] & 7)));
    return out;
}
}  // namespace detail

class Interleaver {   // ultra::Interleaver, include/ultra/fec.hpp:85-107; ldpc_decoder.cpp:454-540
public:
    Interleaver(size_t rows, size_t cols) : rows_(rows), cols_(cols), perm_(rows * cols), inv_(rows * cols) {
        pu_block_interleaver_perm(rows, cols, perm_.data());
        for (size_t i = 0; i < perm_.size(); ++i) inv_[perm_[i]] = static_cast<uint32_t>(i);
    }
    Bytes interleave(ByteSpan data) { return detail::permute_bits(perm_, data); }
    Bytes deinterleave(ByteSpan data) { return detail::permute_bits(inv_, data); }
    std::vector<float> interleave(std::span<const float> s) { return detail::permute_fwd<float>(perm_, s); }
    std::vector<float> deinterleave(std::span<const float> s) { return detail::permute_fwd<float>(inv_, s); }
    size_t getPermutation(size_t i) const { return i < perm_.size() ? perm_[i] : 0; }
    size_t getRows() const { return rows_; }
    size_t getCols() const { return cols_; }

private:
    size_t rows_, cols_;
    std::vector<uint32_t> perm_, inv_;
};

class ChannelInterleaver {   // ultra::ChannelInterleaver, include/ultra/fec.hpp:120-142; ldpc_decoder.cpp:547-672
public:
    explicit ChannelInterleaver(size_t bits_per_symbol, size_t total_bits = 648)
        : bps_(bits_per_symbol), perm_(total_bits), inv_(total_bits) {
        pu_channel_interleaver_perm(bits_per_symbol, total_bits, perm_.data(), inv_.data(), &step_);
    }
    std::vector<float> interleave(std::span<const float> s) { return detail::permute_fwd<float>(perm_, s); }
    std::vector<float> deinterleave(std::span<const float> s) { return detail::permute_fwd<float>(inv_, s); }
    Bytes interleave(ByteSpan d) { return detail::permute_bits(perm_, d); }
    Bytes deinterleave(ByteSpan d) { return detail::permute_bits(inv_, d); }
    size_t getSymbolSeparation() const { return bps_ ? step_ / bps_ : 0; }   // ldpc_decoder.cpp:583
    size_t getStep() const { return step_; }

private:
    size_t bps_, step_ = 0;
    std::vector<uint32_t> perm_, inv_;
};

// ------------------------------------------------------------------------------------------------ OFDM
class OFDMModulator {   // ultra::OFDMModulator, include/ultra/ofdm.hpp:24-52 (host side: TX stimulus only)
public:
    explicit OFDMModulator(const ModemConfig& config) : cfg_(config) {}
    // generateTrainingSymbols(2) followed by modulate(data, mod): the presynced frame (layout 0), or
    // generatePreamble() followed by modulate (layout 1); modulator.cpp:348-580
    Samples frame(ByteSpan data, Modulation mod, int layout = 0) {
        pu_modem_config p = detail::to_pod(cfg_);
        p.modulation = static_cast<uint32_t>(mod);
        size_t n = 0;
        pu_ofdm_tx(&p, layout, data.data(), data.size(), nullptr, 0, &n);
        Samples out(n);
        if (pu_ofdm_tx(&p, layout, data.data(), data.size(), out.data(), out.size(), &n) != PU_OK) return {};
        return out;
    }
    // The reference's modulator is stateful: generatePreamble() / generateTrainingSymbols() restart the TX mixer and the
    // differential state (modulator.cpp:479-488,534-546), modulate() continues from where they left it.  The mixer has advanced
    // 2 x (fft + cp) samples behind the Schmidl-Cox preamble (one STS and one LTS waveform, each repeated) but 2 x (fft + cp + guard)
    // behind two training symbols, so the data symbols differ: modulate() follows whichever prefix was generated last.
    Samples generatePreamble() { layout_ = 1; return frame(ByteSpan{}, cfg_.modulation, 1); }   // Schmidl-Cox preamble (:479-532)
    Samples generateTrainingSymbols(int count = 2) {   // only the reference's default count is on the path
        layout_ = 0;
        Samples f = frame(ByteSpan{}, cfg_.modulation, 0);
        f.resize(std::min<size_t>(f.size(), static_cast<size_t>(count) * samplesPerSymbol()));
        return f;
    }
    Samples modulate(ByteSpan data, Modulation mod) {   // data symbols that follow the prefix generated last (:348-477)
        Samples f = frame(data, mod, layout_);
        const size_t skip = std::min<size_t>(f.size(), frame(ByteSpan{}, mod, layout_).size());
        return Samples(f.begin() + static_cast<std::ptrdiff_t>(skip), f.end());
    }
    size_t samplesPerSymbol() const { return cfg_.getSymbolDuration(); }

private:
    ModemConfig cfg_;
    int layout_ = 0;     // 0: behind generateTrainingSymbols (also before any prefix), 1: behind generatePreamble
};

class OFDMDemodulator {   // ultra::OFDMDemodulator, include/ultra/ofdm.hpp:58-127
public:
    explicit OFDMDemodulator(const ModemConfig& config) : cfg_(config) {
        const pu_modem_config p = detail::to_pod(config);
        const pu_status s = pu_ofdm_create(detail::shared_context(), &p, &h_);
        if (s != PU_OK) detail::fail("pu_ofdm_create", s);
        sym_len_ = static_cast<size_t>(pu_ofdm_symbol_samples(h_));
        bits_per_symbol_ = static_cast<size_t>(pu_ofdm_bits_per_symbol(h_));
    }
    ~OFDMDemodulator() { pu_ofdm_destroy(h_); }
    OFDMDemodulator(const OFDMDemodulator&) = delete;
    OFDMDemodulator& operator=(const OFDMDemodulator&) = delete;

    // demodulator.cpp:459-743.  Samples accumulate on the host; every call hands the whole buffer to
    // pu_ofdm_process_batch, which replays the reference's sequence of process() calls for the chunk size of the FIRST call
    // (the reference's outcome depends on the chunking; callers feed equal chunks, tools/test_mode_snr.cpp:65-70), so the
    // synchronisation decision, the coarse CFO and the soft bits are the reference's.  Returns true once a codeword's
    // worth of soft bits (648) is available, like the reference.  Not replayed: the mid-frame preamble re-check and the
    // idle / timeout resets of the SYNCED state (:604-662,692-735), which need idle calls.
    bool process(SampleSpan samples) {
        if (rx_.empty()) chunk_ = std::max<size_t>(samples.size(), 1);
        const bool idle_call = samples.empty();
        rx_.insert(rx_.end(), samples.begin(), samples.end());
        if (!idle_call || !synced_) {
            const size_t cap = (rx_.size() / sym_len_ + 1) * bits_per_symbol_;
            std::vector<float> llr(std::max<size_t>(cap, 1), 0.0f);
            int32_t n = 0, info[4] = {0, 0, 0, 0};
            float cfo = 0.0f, snr_db = 0.0f;
            const pu_status s = pu_ofdm_process_batch(h_, rx_.data(), 1, rx_.size(), chunk_, 0.0f, llr.data(), llr.size(), &n, info, &cfo,
                                                      &snr_db, PU_MEM_HOST, nullptr);
            if (s != PU_OK) {
                fprintf(stderr, "pu::OFDMDemodulator::process: %s: %s\n", pu_status_string(s), pu_last_error());
                return false;
            }
            synced_ = info[0] != 0;
            if (!synced_) {
                if (rx_.size() > 40000 && !warned_) {
                    // beyond 2 * OVERLAP_SAMPLES without a preamble the reference trims its buffer (:593-597); this class keeps the
                    // stream and searches its first 40 000 samples only -- say so instead of silently never synchronising
                    warned_ = true;
                    fprintf(stderr, "pu::OFDMDemodulator::process: %zu samples buffered without a preamble in the first 40000; call "
                                    "reset() between transmissions (the reference would have trimmed its buffer here)\n", rx_.size());
                }
                return false;
            }
            last_sync_offset_ = static_cast<size_t>(info[1]);
            cfo_hz_ = cfo;
            snr_db_ = snr_db;
            // the kernel demodulates the whole stream again: hand out only the soft bits behind the ones already drained
            const size_t total = static_cast<size_t>(std::max(n, 0));
            const size_t from = std::min(consumed_, total);
            soft_.assign(llr.begin() + static_cast<std::ptrdiff_t>(from), llr.begin() + static_cast<std::ptrdiff_t>(total));
        }
        const bool has_codeword = soft_.size() >= PU_LDPC_N;
        if (!has_codeword && synced_ && idle_call) {   // "frame complete": leftover bits dropped, back to SEARCHING (:725-735)
            soft_.clear();
            rx_.clear();
            consumed_ = 0;
            synced_ = false;
        }
        return has_codeword;
    }

    // demodulator.cpp:854-985.  The frame is demodulated by the CUDA kernel from a fresh tracker state with the CFO
    // and phase given to setFrequencyOffset[WithPhase]; like the reference, the soft-bit FIFO is replaced.
    bool processPresynced(SampleSpan samples, int training_symbols = 2) {
        if (samples.size() < sym_len_) return false;               // :865-867
        soft_.clear();
        if (!cfo_set_ && training_symbols >= 2 && std::fabs(cfo_hz_) < 0.1f) {
            // no setFrequencyOffset since reset(): the reference estimates the CFO from the training symbols (:918-925,
            // estimateCFOFromTraining, ofdm_sync.cpp:278-380) and keeps the correction phase reset() left behind
            float est = 0.0f;
            if (pu_ofdm_training_cfo_batch(h_, samples.data(), 1, samples.size(), training_symbols, &est, PU_MEM_HOST, nullptr) != PU_OK) return false;
            cfo_hz_ = est;
        }
        const size_t n_sym = samples.size() / sym_len_;
        const size_t n_data = n_sym > static_cast<size_t>(training_symbols) ? n_sym - static_cast<size_t>(training_symbols) : 0;
        const size_t n_llr = n_data * bits_per_symbol_;
        std::vector<float> llr(std::max<size_t>(n_llr, 1), 0.0f);
        float snr_db = 0.0f, fcfo = cfo_hz_;
        const pu_status s = pu_ofdm_presynced_batch(h_, samples.data(), 1, samples.size(), training_symbols, &cfo_hz_, &cfo_phase_,
                                                    llr.data(), llr.size(), &snr_db, &fcfo, PU_MEM_HOST, nullptr);
        if (s != PU_OK) return false;
        llr.resize(n_llr);
        soft_ = std::move(llr);
        snr_db_ = snr_db;
        cfo_hz_ = fcfo;            // getFrequencyOffset() reports the tracked value (:801-803)
        synced_ = true;
        return soft_.size() >= PU_LDPC_N;
    }
    std::vector<float> getSoftBits() {   // drains at most 648 per call (:766-791)
        if (soft_.size() <= PU_LDPC_N) {
            std::vector<float> out = std::move(soft_);
            soft_.clear();
            consumed_ += out.size();
            return out;
        }
        std::vector<float> out(soft_.begin(), soft_.begin() + PU_LDPC_N);
        soft_.erase(soft_.begin(), soft_.begin() + PU_LDPC_N);
        consumed_ += out.size();
        return out;
    }
    Bytes getData() {   // hard decisions of the FIFO, bit = (llr > 0) as the reference has it (:745-764)
        Bytes data;
        uint8_t byte = 0;
        int cnt = 0;
        for (float l : soft_) {
            byte = static_cast<uint8_t>((byte << 1) | (l > 0 ? 1 : 0));
            if (++cnt == 8) { data.push_back(byte); byte = 0; cnt = 0; }
        }
        consumed_ += soft_.size();
        soft_.clear();
        return data;
    }
    float getEstimatedSNR() const { return snr_db_; }
    float getFrequencyOffset() const { return cfo_hz_; }
    void setFrequencyOffset(float cfo_hz) { cfo_hz_ = cfo_hz; cfo_phase_ = 0.0f; cfo_set_ = true; }                    // :805-814
    void setFrequencyOffsetWithPhase(float cfo_hz, float phase) { cfo_hz_ = cfo_hz; cfo_phase_ = phase; cfo_set_ = true; }   // :816-825
    Symbol getConstellationSymbols() const { return {}; }   // GUI scatter plot: not produced by the batch kernels
    bool isSynced() const { return synced_; }
    bool hasPendingData() const { return !soft_.empty(); }
    size_t getLastSyncOffset() const { return last_sync_offset_; }
    void setTimingOffset(int) {}
    void reset() {   // :987-1017
        soft_.clear();
        rx_.clear();
        consumed_ = 0;
        warned_ = false;
        last_sync_offset_ = 0;
        cfo_hz_ = 0.0f;
        cfo_phase_ = 0.0f;
        cfo_set_ = false;
        synced_ = false;
        snr_db_ = 0.0f;
    }
    pu_ofdm* handle() const { return h_; }

private:
    ModemConfig cfg_;
    pu_ofdm* h_ = nullptr;
    size_t sym_len_ = 0, bits_per_symbol_ = 0;
    std::vector<float> soft_;
    std::vector<float> rx_;        // samples handed to process() so far
    size_t chunk_ = 960, last_sync_offset_ = 0;
    size_t consumed_ = 0;          // soft bits of the current stream already handed out by getSoftBits() / getData()
    bool warned_ = false;
    float cfo_hz_ = 0.0f, cfo_phase_ = 0.0f, snr_db_ = 0.0f;
    bool cfo_set_ = false, synced_ = false;
};

// ------------------------------------------------------------------------------------------------ waveform plugin
#define PU_IWAVEFORM_BASE : public IWaveform
#define PU_OVERRIDE override
#ifdef PU_DROPIN_WITH_ULTRA
using ultra::IWaveform;
using ultra::SyncResult;
using ultra::WaveformCapabilities;
namespace protocol = ultra::protocol;
#else
namespace protocol {
enum class WaveformMode : uint8_t { OFDM_COX = 0x00, OTFS_EQ = 0x01, OTFS_RAW = 0x02, MFSK = 0x03, MC_DPSK = 0x04, OFDM_CHIRP = 0x05,
                                    AUTO = 0xFF };   // src/protocol/frame_v2.hpp:28-36
}
struct SyncResult {   // src/waveform/waveform_interface.hpp:37-44
    bool detected = false;
    int start_sample = -1;
    float correlation = 0.0f, cfo_hz = 0.0f, snr_estimate = 0.0f;
    bool has_training = false;
};
struct WaveformCapabilities {   // :25-34
    bool supports_cfo_correction = false, supports_doppler_correction = false, requires_pilots = false;
    bool supports_differential = true;
    float min_snr_db = 0.0f, max_snr_db = 30.0f, max_throughput_bps = 1000.0f, preamble_duration_ms = 500.0f;
};
class IWaveform {   // src/waveform/waveform_interface.hpp:47-157, the same pure-virtual surface
public:
    virtual ~IWaveform() = default;
    virtual std::string getName() const = 0;
    virtual protocol::WaveformMode getMode() const = 0;
    virtual WaveformCapabilities getCapabilities() const = 0;
    virtual void configure(Modulation mod, CodeRate rate) = 0;
    virtual void setFrequencyOffset(float cfo_hz) = 0;
    virtual void setTxFrequencyOffset(float cfo_hz) = 0;
    virtual Modulation getModulation() const = 0;
    virtual CodeRate getCodeRate() const = 0;
    virtual float getFrequencyOffset() const = 0;
    virtual Samples generatePreamble() = 0;
    virtual Samples modulate(const Bytes& encoded_data) = 0;
    virtual bool detectSync(SampleSpan samples, SyncResult& result, float threshold = 0.3f) = 0;
    virtual bool process(SampleSpan samples) = 0;
    virtual std::vector<float> getSoftBits() = 0;
    virtual void reset() = 0;
    virtual bool isSynced() const = 0;
    virtual bool hasData() const = 0;
    virtual float estimatedSNR() const = 0;
    virtual float estimatedCFO() const = 0;
    virtual std::vector<std::complex<float>> getConstellationSymbols() const = 0;
    virtual std::string getStatusString() const = 0;
    virtual int getCarrierCount() const = 0;
    virtual float getThroughput(CodeRate rate) const = 0;
    virtual int getSamplesPerSymbol() const = 0;
    virtual int getPreambleSamples() const = 0;
    virtual int getMinSamplesForFrame() const = 0;
};
#endif
using WaveformPtr = std::unique_ptr<IWaveform>;

// The RX data path of ultra::OFDMChirpWaveform (src/waveform/ofdm_chirp_waveform.cpp): configure ->
// detectSync (dual-chirp detector, chirp_sync.hpp:349-506) -> setFrequencyOffset -> process(span starting at the first training
// symbol) -> getSoftBits; callers with external (genie) timing call process() directly, as tools/test_ofdm_chirp_pilots.cpp does.
class OfdmChirpWaveform PU_IWAVEFORM_BASE {
public:
    OfdmChirpWaveform() {   // ofdm_chirp_waveform.cpp:10-18
        cfg_.fft_size = 512;
        cfg_.num_carriers = 30;
        cfg_.modulation = Modulation::DQPSK;
        cfg_.code_rate = CodeRate::R1_2;
        cfg_.use_pilots = false;
        rebuild();
    }
    explicit OfdmChirpWaveform(const ModemConfig& cfg) : cfg_(cfg) {   // :20-31: chirp mode is differential, without pilots
        if (!differential(cfg_.modulation)) cfg_.modulation = Modulation::DQPSK;
        cfg_.use_pilots = false;
        rebuild();
    }

    std::string getName() const PU_OVERRIDE { return "OFDM-CHIRP"; }
    protocol::WaveformMode getMode() const PU_OVERRIDE { return protocol::WaveformMode::OFDM_CHIRP; }
    WaveformCapabilities getCapabilities() const PU_OVERRIDE {   // ofdm_chirp_waveform.cpp:60-72
        WaveformCapabilities c;
        c.supports_cfo_correction = true;
        c.supports_doppler_correction = true;
        c.requires_pilots = false;
        c.supports_differential = true;
        c.min_snr_db = 10.0f;
        c.max_snr_db = 20.0f;
        c.max_throughput_bps = getThroughput(CodeRate::R2_3);
        c.preamble_duration_ms = chirp_total() * 1000.0f / cfg_.sample_rate;
        return c;
    }
    void configure(Modulation mod, CodeRate rate) PU_OVERRIDE {   // :68-84: only differential modulations, never pilots
        if (!differential(mod)) mod = Modulation::DQPSK;
        cfg_.modulation = mod;
        cfg_.code_rate = rate;
        cfg_.use_pilots = false;
        rebuild();
    }
    void setFrequencyOffset(float cfo_hz) PU_OVERRIDE { cfo_hz_ = cfo_hz; demod_->setFrequencyOffset(cfo_hz); }   // :87-93
    void setTxFrequencyOffset(float cfo_hz) PU_OVERRIDE { cfg_.tx_cfo_hz = cfo_hz; mod_ = std::make_unique<OFDMModulator>(cfg_); }
    Modulation getModulation() const PU_OVERRIDE { return cfg_.modulation; }
    CodeRate getCodeRate() const PU_OVERRIDE { return cfg_.code_rate; }
    float getFrequencyOffset() const PU_OVERRIDE { return cfo_hz_; }

    Samples generatePreamble() PU_OVERRIDE {   // [CHIRP][TRAINING_SYMBOLS] (:104-118)
        size_t n = 0;
        pu_chirp_generate(static_cast<float>(cfg_.sample_rate), cfg_.tx_cfo_hz, nullptr, 0, &n);
        Samples pre(n);
        pu_chirp_generate(static_cast<float>(cfg_.sample_rate), cfg_.tx_cfo_hz, pre.data(), pre.size(), &n);
        const Samples training = mod_->generateTrainingSymbols(2);
        pre.insert(pre.end(), training.begin(), training.end());
        return pre;
    }
    Samples modulate(const Bytes& encoded) PU_OVERRIDE { return mod_->modulate(ByteSpan(encoded.data(), encoded.size()), cfg_.modulation); }

    bool detectSync(SampleSpan samples, SyncResult& result, float threshold = 0.3f) PU_OVERRIDE {   // :129-172: dual-chirp detection
        result = SyncResult{};
        int32_t info[4] = {0, -1, -1, -1};
        float val[4] = {0, 0, 0, 0};
        if (samples.empty() ||
            pu_ofdm_chirp_receive_batch(demod_->handle(), samples.data(), 1, samples.size(), threshold, nullptr, 0, nullptr, info, val, nullptr,
                                        PU_MEM_HOST, nullptr) != PU_OK)
            return false;
        result.detected = info[0] != 0;
        result.correlation = std::max(val[1], val[2]);
        result.cfo_hz = val[0];
        result.has_training = true;
        if (result.detected) {
            synced_ = true;
            last_cfo_ = val[0];
            result.start_sample = info[3];
            training_start_sample_ = static_cast<size_t>(info[3]);
        }
        return result.detected;
    }
    bool process(SampleSpan samples) PU_OVERRIDE {   // :174-219: processPresynced on the span from the training start
        // CFO rotator phase accumulated since sample 0 of the audio, wrapped to [-pi, pi] (:177-181)
        constexpr double kPi = 3.14159265358979323846;
        float initial_phase_rad = static_cast<float>(-2.0f * kPi * cfo_hz_ * training_start_sample_ / cfg_.sample_rate);
        while (initial_phase_rad > kPi) initial_phase_rad = static_cast<float>(initial_phase_rad - 2.0f * kPi);
        while (initial_phase_rad < -kPi) initial_phase_rad = static_cast<float>(initial_phase_rad + 2.0f * kPi);
        demod_->setFrequencyOffsetWithPhase(cfo_hz_, initial_phase_rad);
        const bool ready = demod_->processPresynced(samples, 2);
        if (ready) {   // drain everything (:202-212); an incomplete frame leaves the previous soft bits alone
            soft_.clear();
            while (demod_->hasPendingData()) {
                std::vector<float> c = demod_->getSoftBits();
                if (c.empty()) break;
                soft_.insert(soft_.end(), c.begin(), c.end());
            }
        }
        return ready;
    }
    std::vector<float> getSoftBits() PU_OVERRIDE { return std::move(soft_); }
    void reset() PU_OVERRIDE {   // preserves the CFO, as the code does (:221-230)
        soft_.clear();
        synced_ = false;
        demod_->reset();
    }
    bool isSynced() const PU_OVERRIDE { return synced_ || demod_->isSynced(); }                     // :232-234
    bool hasData() const PU_OVERRIDE { return !soft_.empty() || demod_->hasPendingData(); }
    float estimatedSNR() const PU_OVERRIDE { return demod_->getEstimatedSNR(); }
    float estimatedCFO() const PU_OVERRIDE { return std::fabs(last_cfo_) > 0.1f ? last_cfo_ : demod_->getFrequencyOffset(); }   // :247-255
    std::vector<std::complex<float>> getConstellationSymbols() const PU_OVERRIDE { return {}; }
    std::string getStatusString() const PU_OVERRIDE { return "OFDM-Chirp " + std::to_string(cfg_.num_carriers) + " carriers"; }
    int getCarrierCount() const PU_OVERRIDE { return static_cast<int>(cfg_.num_carriers); }
    float getThroughput(CodeRate rate) const PU_OVERRIDE {   // :277-306: every carrier carries data
        const int bits_per_carrier = cfg_.modulation == Modulation::DBPSK ? 1 : cfg_.modulation == Modulation::D8PSK ? 3 : 2;
        const float raw_bps = static_cast<float>(cfg_.sample_rate) / getSamplesPerSymbol() * static_cast<int>(cfg_.num_carriers) * bits_per_carrier;
        float code_ratio = 0.5f;
        switch (rate) {
            case CodeRate::R1_4: code_ratio = 0.25f; break;
            case CodeRate::R1_3: code_ratio = 0.333f; break;
            case CodeRate::R1_2: code_ratio = 0.5f; break;
            case CodeRate::R2_3: code_ratio = 0.667f; break;
            case CodeRate::R3_4: code_ratio = 0.75f; break;
            case CodeRate::R5_6: code_ratio = 0.833f; break;
            default: break;
        }
        return raw_bps * code_ratio;
    }
    int getSamplesPerSymbol() const PU_OVERRIDE { return static_cast<int>(cfg_.getSymbolDuration()); }
    int getPreambleSamples() const PU_OVERRIDE { return chirp_total() + 2 * getSamplesPerSymbol(); }   // :322-327
    int getMinSamplesForFrame() const PU_OVERRIDE {   // 2 training symbols + the data symbols of one codeword
        const int bps = pu_ofdm_bits_per_symbol(demod_->handle());
        const int nsym = bps > 0 ? (PU_LDPC_N + bps - 1) / bps : 0;
        return (2 + nsym) * getSamplesPerSymbol();
    }

private:
    static bool differential(Modulation m) { return m == Modulation::DBPSK || m == Modulation::DQPSK || m == Modulation::D8PSK; }
    int chirp_total() const {   // ChirpSync::getTotalSamples for getChirpConfig() (:39-49): two 500 ms chirps, two 100 ms gaps
        const size_t chirp = static_cast<size_t>(cfg_.sample_rate * 500.0f / 1000.0f), gap = static_cast<size_t>(cfg_.sample_rate * 100.0f / 1000.0f);
        return static_cast<int>(2 * chirp + 2 * gap);
    }
    void rebuild() {
        demod_ = std::make_unique<OFDMDemodulator>(cfg_);
        mod_ = std::make_unique<OFDMModulator>(cfg_);
    }
    ModemConfig cfg_;
    std::unique_ptr<OFDMDemodulator> demod_;
    std::unique_ptr<OFDMModulator> mod_;
    std::vector<float> soft_;
    float cfo_hz_ = 0.0f, last_cfo_ = 0.0f;
    size_t training_start_sample_ = 0;   // SyncResult::start_sample of the last detection (0 before any: phase 0)
    bool synced_ = false;
};

// ultra::MCDPSKWaveform (src/waveform/mc_dpsk_waveform.cpp): [chirp pair][training][reference][data].  detectSync = dual-chirp
// detection with the training start two chirps and two gaps behind the up chirp (:100-142); process = setChirpDetected(cfo) ->
// MultiCarrierDPSKDemodulator::process -> getSoftBits (:144-170), i.e. Hilbert-FIR CFO correction + processGotChirp.
struct McDpskConfig {   // receive-path fields of ultra::MultiCarrierDPSKConfig (src/psk/multi_carrier_dpsk.hpp:26-51)
    float sample_rate = 48000.0f;
    int num_carriers = 8;
    float freq_low = 500.0f, freq_high = 2500.0f;
    int samples_per_symbol = 512, bits_per_symbol = 2, training_symbols = 8;
    float tx_cfo_hz = 0.0f;
};
class McDpskWaveform PU_IWAVEFORM_BASE {
public:
    explicit McDpskWaveform(const McDpskConfig& cfg = McDpskConfig{}) : cfg_(cfg), ctx_(detail::shared_context()) { rebuild(); }
    ~McDpskWaveform() {
        if (h_) pu_mcdpsk_destroy(h_);
    }
    McDpskWaveform(const McDpskWaveform&) = delete;
    McDpskWaveform& operator=(const McDpskWaveform&) = delete;

    std::string getName() const PU_OVERRIDE { return "MC-DPSK"; }
    protocol::WaveformMode getMode() const PU_OVERRIDE { return protocol::WaveformMode::MC_DPSK; }
    WaveformCapabilities getCapabilities() const PU_OVERRIDE {   // mc_dpsk_waveform.cpp:33-44
        WaveformCapabilities c;
        c.supports_cfo_correction = true;
        c.supports_doppler_correction = true;
        c.requires_pilots = false;
        c.supports_differential = true;
        c.min_snr_db = -3.0f;
        c.max_snr_db = 15.0f;
        c.max_throughput_bps = getThroughput(CodeRate::R1_4);
        c.preamble_duration_ms = 500.0f * 2 + 100.0f * 2;
        return c;
    }
    void configure(Modulation mod, CodeRate rate) PU_OVERRIDE {   // :46-69
        modulation_ = mod;
        code_rate_ = rate;
        if (mod != Modulation::DQPSK && mod != Modulation::DBPSK && mod != Modulation::D8PSK) modulation_ = Modulation::DQPSK;
        cfg_.bits_per_symbol = mod == Modulation::DBPSK ? 1 : mod == Modulation::D8PSK ? 3 : 2;
        rebuild();
    }
    void setFrequencyOffset(float cfo_hz) PU_OVERRIDE { cfo_hz_ = cfo_hz; }                                   // :71-76
    void setTxFrequencyOffset(float cfo_hz) PU_OVERRIDE { cfg_.tx_cfo_hz = cfo_hz; }                          // :78-84
    Modulation getModulation() const PU_OVERRIDE { return modulation_; }
    CodeRate getCodeRate() const PU_OVERRIDE { return code_rate_; }
    float getFrequencyOffset() const PU_OVERRIDE { return cfo_hz_; }

    Samples generatePreamble() PU_OVERRIDE {   // chirp pair + training + reference (multi_carrier_dpsk.hpp:104-115)
        size_t n = 0;
        pu_chirp_generate(cfg_.sample_rate, cfg_.tx_cfo_hz, nullptr, 0, &n);
        Samples pre(n);
        pu_chirp_generate(cfg_.sample_rate, cfg_.tx_cfo_hz, pre.data(), pre.size(), &n);
        const Samples head = tx(Bytes{});
        pre.insert(pre.end(), head.begin(), head.end());
        return pre;
    }
    Samples modulate(const Bytes& encoded) PU_OVERRIDE {   // data symbols behind the reference symbol (:176-243)
        Samples all = tx(encoded);
        const size_t pre = static_cast<size_t>(cfg_.training_symbols + 1) * cfg_.samples_per_symbol;
        return all.size() > pre ? Samples(all.begin() + pre, all.end()) : Samples{};
    }

    bool detectSync(SampleSpan samples, SyncResult& result, float threshold = 0.3f) PU_OVERRIDE {   // :100-142
        result = SyncResult{};
        int32_t info[4] = {0, -1, -1, -1};
        float val[4] = {0, 0, 0, 0};
        if (!h_ || samples.empty() ||
            pu_mcdpsk_chirp_receive_batch(h_, samples.data(), 1, samples.size(), threshold, nullptr, 0, nullptr, info, val, nullptr, PU_MEM_HOST,
                                          nullptr) != PU_OK)
            return false;
        result.detected = info[0] != 0;
        result.start_sample = info[1];
        result.correlation = std::max(val[1], val[2]);
        result.cfo_hz = val[0];
        result.has_training = true;
        if (result.detected) {
            synced_ = true;
            last_cfo_ = val[0];
            result.start_sample = info[3];
        }
        return result.detected;
    }
    bool process(SampleSpan samples) PU_OVERRIDE {   // :144-170
        soft_.clear();
        if (!h_ || samples.empty()) return false;
        const size_t nsym = samples.size() / static_cast<size_t>(cfg_.samples_per_symbol);
        const size_t stride = std::max<size_t>(1, nsym * static_cast<size_t>(cfg_.num_carriers * cfg_.bits_per_symbol));
        std::vector<float> llr(stride);
        int32_t n = 0;
        float after = cfo_hz_;
        if (pu_mcdpsk_got_chirp_batch(h_, samples.data(), 1, samples.size(), &cfo_hz_, llr.data(), stride, &n, &after, PU_MEM_HOST, nullptr) != PU_OK)
            return false;
        demod_cfo_ = after;
        if (n <= 0) return false;
        llr.resize(static_cast<size_t>(n));
        soft_ = std::move(llr);
        synced_ = true;
        return true;
    }
    std::vector<float> getSoftBits() PU_OVERRIDE { return std::move(soft_); }
    void reset() PU_OVERRIDE {   // keeps the CFO (:176-184)
        soft_.clear();
        synced_ = false;
    }
    bool isSynced() const PU_OVERRIDE { return synced_; }
    bool hasData() const PU_OVERRIDE { return !soft_.empty(); }
    float estimatedSNR() const PU_OVERRIDE { return 0.0f; }
    float estimatedCFO() const PU_OVERRIDE { return demod_cfo_; }
    std::vector<std::complex<float>> getConstellationSymbols() const PU_OVERRIDE { return {}; }
    std::string getStatusString() const PU_OVERRIDE {
        return "MC-DPSK " + std::to_string(cfg_.num_carriers) + " carriers @ " + std::to_string(static_cast<int>(getThroughput(code_rate_))) + " bps";
    }
    int getCarrierCount() const PU_OVERRIDE { return cfg_.num_carriers; }
    float getThroughput(CodeRate rate) const PU_OVERRIDE {   // :220-236
        const float raw_bps = cfg_.sample_rate / cfg_.samples_per_symbol * cfg_.num_carriers * cfg_.bits_per_symbol;
        float code_ratio = 0.25f;
        switch (rate) {
            case CodeRate::R1_4: code_ratio = 0.25f; break;
            case CodeRate::R1_3: code_ratio = 0.333f; break;
            case CodeRate::R1_2: code_ratio = 0.5f; break;
            case CodeRate::R2_3: code_ratio = 0.667f; break;
            case CodeRate::R3_4: code_ratio = 0.75f; break;
            case CodeRate::R5_6: code_ratio = 0.833f; break;
            default: break;
        }
        return raw_bps * code_ratio;
    }
    int getSamplesPerSymbol() const PU_OVERRIDE { return cfg_.samples_per_symbol; }
    int getPreambleSamples() const PU_OVERRIDE {   // ChirpSync::getTotalSamples (:238-247)
        const size_t chirp = static_cast<size_t>(cfg_.sample_rate * 500.0f / 1000.0f), gap = static_cast<size_t>(cfg_.sample_rate * 100.0f / 1000.0f);
        return static_cast<int>(2 * chirp + 2 * gap);
    }
    int getMinSamplesForFrame() const PU_OVERRIDE {   // :254-265
        const int bits = cfg_.num_carriers * cfg_.bits_per_symbol;
        return (cfg_.training_symbols + 1 + (PU_LDPC_N + bits - 1) / bits) * cfg_.samples_per_symbol;
    }

private:
    pu_mcdpsk_config pod() const {
        pu_mcdpsk_config c{};
        c.sample_rate = cfg_.sample_rate;
        c.freq_low = cfg_.freq_low; c.freq_high = cfg_.freq_high;
        c.num_carriers = static_cast<uint32_t>(cfg_.num_carriers);
        c.samples_per_symbol = static_cast<uint32_t>(cfg_.samples_per_symbol);
        c.bits_per_symbol = static_cast<uint32_t>(cfg_.bits_per_symbol);
        c.training_symbols = static_cast<uint32_t>(cfg_.training_symbols);
        return c;
    }
    void rebuild() {   // D8PSK over MC-DPSK (3 bits per carrier) is not on the batched path: the waveform then reports "no data"
        if (h_) { pu_mcdpsk_destroy(h_); h_ = nullptr; }
        const pu_mcdpsk_config c = pod();
        pu_mcdpsk_create(ctx_, &c, &h_);
    }
    Samples tx(const Bytes& data) const {
        const pu_mcdpsk_config c = pod();
        size_t n = 0;
        (void)pu_mcdpsk_tx(&c, data.data(), data.size(), nullptr, 0, &n);   // length query
        Samples out(n);
        if (pu_mcdpsk_tx(&c, data.data(), data.size(), out.data(), out.size(), &n) != PU_OK) return {};
        return out;
    }
    McDpskConfig cfg_;
    pu_ctx* ctx_;
    pu_mcdpsk* h_ = nullptr;
    Modulation modulation_ = Modulation::DQPSK;
    CodeRate code_rate_ = CodeRate::R1_4;
    std::vector<float> soft_;
    float cfo_hz_ = 0.0f, last_cfo_ = 0.0f, demod_cfo_ = 0.0f;
    bool synced_ = false;
};

// ultra::OFDMNvisWaveform ("OFDM_COX", src/waveform/ofdm_cox_waveform.cpp): the Schmidl-Cox OFDM waveform.  TX = OFDMModulator::
// generatePreamble / modulate; RX = OFDMDemodulator::process on the stream (acquisition + SYNCED state: pu_ofdm_process_batch), with
// detectSync reporting what process() found (:98-121).
class OFDMNvisWaveform PU_IWAVEFORM_BASE {
public:
    OFDMNvisWaveform() {   // :9-17
        cfg_.fft_size = 512;
        cfg_.num_carriers = 30;
        cfg_.modulation = Modulation::QPSK;
        cfg_.code_rate = CodeRate::R1_2;
        cfg_.use_pilots = true;
        rebuild();
    }
    explicit OFDMNvisWaveform(const ModemConfig& cfg) : cfg_(cfg) { rebuild(); }

    std::string getName() const PU_OVERRIDE { return "OFDM-COX"; }
    protocol::WaveformMode getMode() const PU_OVERRIDE { return protocol::WaveformMode::OFDM_COX; }
    WaveformCapabilities getCapabilities() const PU_OVERRIDE {   // :30-49
        WaveformCapabilities c;
        c.supports_cfo_correction = true;
        c.supports_doppler_correction = true;
        c.requires_pilots = cfg_.use_pilots;
        c.supports_differential = true;
        c.min_snr_db = differential(cfg_.modulation) ? 12.0f : 17.0f;
        c.max_snr_db = 35.0f;
        c.max_throughput_bps = getThroughput(CodeRate::R3_4);
        c.preamble_duration_ms = 2.0f * getSamplesPerSymbol() * 1000.0f / cfg_.sample_rate;
        return c;
    }
    void configure(Modulation mod, CodeRate rate) PU_OVERRIDE {   // :51-66: pilots follow the modulation family
        cfg_.modulation = mod;
        cfg_.code_rate = rate;
        cfg_.use_pilots = !differential(mod);
        rebuild();
    }
    void setFrequencyOffset(float cfo_hz) PU_OVERRIDE { cfo_hz_ = cfo_hz; demod_->setFrequencyOffset(cfo_hz); }   // :68-73
    void setTxFrequencyOffset(float cfo_hz) PU_OVERRIDE { cfg_.tx_cfo_hz = cfo_hz; rebuild(); }                   // :75-81
    Modulation getModulation() const PU_OVERRIDE { return cfg_.modulation; }
    CodeRate getCodeRate() const PU_OVERRIDE { return cfg_.code_rate; }
    float getFrequencyOffset() const PU_OVERRIDE { return cfo_hz_; }
    Samples generatePreamble() PU_OVERRIDE { return mod_->generatePreamble(); }                                   // :83-88
    Samples modulate(const Bytes& encoded) PU_OVERRIDE { return mod_->modulate(ByteSpan(encoded.data(), encoded.size()), cfg_.modulation); }
    bool detectSync(SampleSpan samples, SyncResult& result, float /*threshold*/ = 0.3f) PU_OVERRIDE {   // :98-121
        demod_->process(samples);
        if (!demod_->isSynced()) return false;
        result.detected = true;
        result.start_sample = static_cast<int>(demod_->getLastSyncOffset());
        result.cfo_hz = demod_->getFrequencyOffset();
        result.snr_estimate = demod_->getEstimatedSNR();
        result.has_training = true;
        return true;
    }
    bool process(SampleSpan samples) PU_OVERRIDE {   // :123-135
        const bool ready = demod_->process(samples);
        if (ready) soft_ = demod_->getSoftBits();
        return ready;
    }
    std::vector<float> getSoftBits() PU_OVERRIDE { return std::move(soft_); }
    void reset() PU_OVERRIDE { demod_->reset(); soft_.clear(); }   // :141-148
    bool isSynced() const PU_OVERRIDE { return demod_->isSynced(); }
    bool hasData() const PU_OVERRIDE { return !soft_.empty() || demod_->hasPendingData(); }
    float estimatedSNR() const PU_OVERRIDE { return demod_->getEstimatedSNR(); }
    float estimatedCFO() const PU_OVERRIDE { return demod_->getFrequencyOffset(); }
    std::vector<std::complex<float>> getConstellationSymbols() const PU_OVERRIDE { return demod_->getConstellationSymbols(); }
    std::string getStatusString() const PU_OVERRIDE {
        return "OFDM-COX " + std::to_string(cfg_.num_carriers) + " carriers" + (cfg_.use_pilots ? " (pilots)" : "");
    }
    int getCarrierCount() const PU_OVERRIDE { return static_cast<int>(cfg_.num_carriers); }
    float getThroughput(CodeRate rate) const PU_OVERRIDE {   // :190-232
        const float symbol_rate = static_cast<float>(cfg_.sample_rate) / getSamplesPerSymbol();
        const float raw_bps = symbol_rate * data_carriers() * bits_per_carrier();
        float code_ratio = 0.5f;
        switch (rate) {
            case CodeRate::R1_4: code_ratio = 0.25f; break;
            case CodeRate::R1_3: code_ratio = 0.333f; break;
            case CodeRate::R1_2: code_ratio = 0.5f; break;
            case CodeRate::R2_3: code_ratio = 0.667f; break;
            case CodeRate::R3_4: code_ratio = 0.75f; break;
            case CodeRate::R5_6: code_ratio = 0.833f; break;
            default: break;
        }
        return raw_bps * code_ratio;
    }
    int getSamplesPerSymbol() const PU_OVERRIDE { return static_cast<int>(cfg_.getSymbolDuration()); }   // OFDMModulator::samplesPerSymbol
    int getPreambleSamples() const PU_OVERRIDE { return 2 * getSamplesPerSymbol(); }                   // :247-250
    int getMinSamplesForFrame() const PU_OVERRIDE {                                                    // :252-279
        const int bits_per_symbol = data_carriers() * bits_per_carrier();
        return 2 * getSamplesPerSymbol() + ((PU_LDPC_N + bits_per_symbol - 1) / bits_per_symbol) * getSamplesPerSymbol();
    }
    void setUsePilots(bool use_pilots) { cfg_.use_pilots = use_pilots; rebuild(); }                    // :281-284

private:
    static bool differential(Modulation m) { return m == Modulation::DBPSK || m == Modulation::DQPSK || m == Modulation::D8PSK; }
    int bits_per_carrier() const {
        switch (cfg_.modulation) {
            case Modulation::DBPSK: case Modulation::BPSK: return 1;
            case Modulation::D8PSK: case Modulation::QAM8: return 3;
            case Modulation::QAM16: return 4;
            case Modulation::QAM32: return 5;
            case Modulation::QAM64: return 6;
            default: return 2;
        }
    }
    int data_carriers() const {
        int n = static_cast<int>(cfg_.num_carriers);
        if (cfg_.use_pilots && cfg_.pilot_spacing > 0) n -= static_cast<int>(cfg_.num_carriers / cfg_.pilot_spacing);
        return n;
    }
    void rebuild() {
        demod_ = std::make_unique<OFDMDemodulator>(cfg_);
        mod_ = std::make_unique<OFDMModulator>(cfg_);
    }
    ModemConfig cfg_;
    std::unique_ptr<OFDMDemodulator> demod_;
    std::unique_ptr<OFDMModulator> mod_;
    std::vector<float> soft_;
    float cfo_hz_ = 0.0f;
};

// ultra::WaveformFactory (src/waveform/waveform_factory.cpp:11-61): the same mode -> implementation mapping, including the fallbacks
// (AUTO and the deprecated MFSK -> MC-DPSK, the unwrapped OTFS modes -> OFDM_COX).
class WaveformFactory {
public:
    static WaveformPtr create(protocol::WaveformMode mode) {
        switch (mode) {
            case protocol::WaveformMode::MC_DPSK: case protocol::WaveformMode::AUTO: case protocol::WaveformMode::MFSK:
                return std::make_unique<McDpskWaveform>();
            case protocol::WaveformMode::OFDM_COX: case protocol::WaveformMode::OTFS_EQ: case protocol::WaveformMode::OTFS_RAW:
                return std::make_unique<OFDMNvisWaveform>();
            case protocol::WaveformMode::OFDM_CHIRP:
                return std::make_unique<OfdmChirpWaveform>();
            default:
                return nullptr;
        }
    }
    static WaveformPtr create(protocol::WaveformMode mode, const ModemConfig& config) {
        switch (mode) {
            case protocol::WaveformMode::MC_DPSK: {
                McDpskConfig c;
                c.sample_rate = static_cast<float>(config.sample_rate);
                return std::make_unique<McDpskWaveform>(c);
            }
            case protocol::WaveformMode::OFDM_COX: return std::make_unique<OFDMNvisWaveform>(config);
            case protocol::WaveformMode::OFDM_CHIRP: return std::make_unique<OfdmChirpWaveform>(config);
            default: return create(mode);
        }
    }
    static WaveformPtr createMCDPSK(int num_carriers) {
        McDpskConfig c;
        c.num_carriers = num_carriers;
        return std::make_unique<McDpskWaveform>(c);
    }
    static std::vector<protocol::WaveformMode> getAvailableModes() {
        return {protocol::WaveformMode::MC_DPSK, protocol::WaveformMode::OFDM_CHIRP, protocol::WaveformMode::OFDM_COX};
    }
    static bool isSupported(protocol::WaveformMode mode) {
        return mode == protocol::WaveformMode::MC_DPSK || mode == protocol::WaveformMode::OFDM_COX || mode == protocol::WaveformMode::OFDM_CHIRP ||
               mode == protocol::WaveformMode::AUTO;
    }
};

// ------------------------------------------------------------------------------------------------ PSK demodulator classes
// ultra::DPSKDemodulator (src/psk/dpsk.hpp:309-1059), receive side: findPreamble (Barker-13 x 3 acquisition, :338-481) leaves the
// reference symbol, the CFO estimate and the initial phase offset behind; demodulateSoft (:827-879) continues from them.
struct DPSKConfig {   // ultra::DPSKConfig, dpsk.hpp:42-50
    float sample_rate = 48000.0f, carrier_freq = 1500.0f;
    int samples_per_symbol = 1536;
    int modulation = 1;      // DPSKModulation: 0 DBPSK, 1 DQPSK, 2 D8PSK
    int bits_per_symbol() const { return modulation == 0 ? 1 : modulation == 2 ? 3 : 2; }
};
class DPSKDemodulator {
public:
    explicit DPSKDemodulator(const DPSKConfig& cfg) : cfg_(cfg) {
        const pu_dpsk_config c{cfg.sample_rate, cfg.carrier_freq, static_cast<uint32_t>(cfg.samples_per_symbol), static_cast<uint32_t>(cfg.modulation)};
        const pu_status s = pu_dpsk_create(detail::shared_context(), &c, &h_);
        if (s != PU_OK) detail::fail("pu_dpsk_create", s);
    }
    ~DPSKDemodulator() { pu_dpsk_destroy(h_); }
    DPSKDemodulator(const DPSKDemodulator&) = delete;
    DPSKDemodulator& operator=(const DPSKDemodulator&) = delete;

    // Returns the offset of the first data sample, or -1 (:338-481).  The reference symbol (the last preamble symbol, :470-478) is kept.
    int findPreamble(SampleSpan samples, int /*num_symbols*/ = 32) {
        int32_t start = -1;
        float cfo = 0.0f, ph = 0.0f;
        if (samples.empty() ||
            pu_dpsk_receive_batch(h_, samples.data(), 1, samples.size(), nullptr, 0, nullptr, &start, &cfo, &ph, PU_MEM_HOST, nullptr) != PU_OK)
            return -1;
        if (start < 0) return -1;
        estimated_cfo_ = cfo;
        initial_phase_offset_ = ph;
        const size_t sps = static_cast<size_t>(cfg_.samples_per_symbol);
        if (static_cast<size_t>(start) >= sps) ref_.assign(samples.begin() + (start - static_cast<int>(sps)), samples.begin() + start);
        return start;
    }
    float getEstimatedCFO() const { return estimated_cfo_; }
    std::vector<float> demodulateSoft(SampleSpan samples) {   // :827-879, from the reference symbol set by findPreamble / setReferenceSymbol
        const size_t sps = static_cast<size_t>(cfg_.samples_per_symbol), nsym = samples.size() / sps;
        if (nsym == 0) return {};
        const bool have_ref = ref_.size() == sps;
        std::vector<float> buf;
        buf.reserve((have_ref ? sps : 0) + nsym * sps);
        if (have_ref) buf.insert(buf.end(), ref_.begin(), ref_.end());
        buf.insert(buf.end(), samples.begin(), samples.begin() + nsym * sps);
        std::vector<float> llr(nsym * static_cast<size_t>(cfg_.bits_per_symbol()));
        if (pu_dpsk_demod_soft_batch(h_, buf.data(), 1, buf.size(), have_ref ? sps : 0, have_ref ? 1 : 0, &estimated_cfo_, &initial_phase_offset_,
                                     llr.data(), llr.size(), PU_MEM_HOST, nullptr) != PU_OK)
            return {};
        ref_.assign(samples.begin() + (nsym - 1) * sps, samples.begin() + nsym * sps);   // prev_symbol_ = the last symbol (:877)
        return llr;
    }
    Bytes demodulate(SampleSpan samples) {   // :805-821: negative LLR = bit 1, whole bytes only
        const std::vector<float> soft = demodulateSoft(samples);
        Bytes out;
        for (size_t i = 0; i + 8 <= soft.size(); i += 8) {
            uint8_t byte = 0;
            for (int b = 0; b < 8; ++b)
                if (soft[i + b] < 0) byte = static_cast<uint8_t>(byte | (1 << (7 - b)));
            out.push_back(byte);
        }
        return out;
    }
    void reset() { ref_.clear(); estimated_cfo_ = 0.0f; initial_phase_offset_ = 0.0f; }   // :881-886
    void setReferenceSymbol(SampleSpan ref_samples) {                                       // :889-892
        const size_t sps = static_cast<size_t>(cfg_.samples_per_symbol);
        if (ref_samples.size() >= sps) ref_.assign(ref_samples.begin(), ref_samples.begin() + sps);
    }
    const DPSKConfig& config() const { return cfg_; }

private:
    DPSKConfig cfg_;
    pu_dpsk* h_ = nullptr;
    std::vector<float> ref_;       // samples of the symbol whose correlation is prev_symbol_
    float estimated_cfo_ = 0.0f, initial_phase_offset_ = 0.0f;
};

// ultra::MultiCarrierDPSKDemodulator (src/psk/multi_carrier_dpsk.hpp:258-701) behind an externally detected chirp:
// setChirpDetected(cfo) -> process(training + reference + data) -> getSoftBits, i.e. processGotChirp (:533-627).
class MultiCarrierDPSKDemodulator {
public:
    explicit MultiCarrierDPSKDemodulator(const McDpskConfig& cfg) : cfg_(cfg) {
        pu_mcdpsk_config c{};
        c.sample_rate = cfg.sample_rate; c.freq_low = cfg.freq_low; c.freq_high = cfg.freq_high;
        c.num_carriers = static_cast<uint32_t>(cfg.num_carriers); c.samples_per_symbol = static_cast<uint32_t>(cfg.samples_per_symbol);
        c.bits_per_symbol = static_cast<uint32_t>(cfg.bits_per_symbol); c.training_symbols = static_cast<uint32_t>(cfg.training_symbols);
        const pu_status s = pu_mcdpsk_create(detail::shared_context(), &c, &h_);
        if (s != PU_OK) detail::fail("pu_mcdpsk_create", s);
    }
    ~MultiCarrierDPSKDemodulator() { pu_mcdpsk_destroy(h_); }
    MultiCarrierDPSKDemodulator(const MultiCarrierDPSKDemodulator&) = delete;
    MultiCarrierDPSKDemodulator& operator=(const MultiCarrierDPSKDemodulator&) = delete;

    void setChirpDetected(float cfo_hz = 0.0f) { cfo_hz_ = cfo_hz; got_chirp_ = true; buffer_.clear(); }   // :356-361
    void setCFO(float cfo_hz) { cfo_hz_ = cfo_hz; }                                                        // :333
    // Appends to the internal buffer (:285-303).  Without setChirpDetected the demodulator would search for the chirp itself
    // (processIdle): that path belongs to the waveform's detectSync (pu::McDpskWaveform) and is not replayed here.
    bool process(SampleSpan samples) {
        if (frame_ready_) return true;
        if (!got_chirp_) return false;
        buffer_.insert(buffer_.end(), samples.begin(), samples.end());
        const size_t nsym = buffer_.size() / static_cast<size_t>(cfg_.samples_per_symbol);
        const size_t stride = std::max<size_t>(1, nsym * static_cast<size_t>(cfg_.num_carriers * cfg_.bits_per_symbol));
        std::vector<float> llr(stride);
        int32_t n = 0;
        float after = cfo_hz_;
        if (pu_mcdpsk_got_chirp_batch(h_, buffer_.data(), 1, buffer_.size(), &cfo_hz_, llr.data(), stride, &n, &after, PU_MEM_HOST, nullptr) != PU_OK)
            return false;
        if (n <= 0) return false;          // not enough samples yet (or rejected by the 5 Hz rule): keep buffering
        llr.resize(static_cast<size_t>(n));
        soft_ = std::move(llr);
        cfo_hz_ = after;
        frame_ready_ = true;
        buffer_.clear();
        return true;
    }
    std::vector<float> getSoftBits() {   // :305-312
        std::vector<float> out = std::move(soft_);
        soft_.clear();
        if (frame_ready_) { frame_ready_ = false; got_chirp_ = false; }
        return out;
    }
    bool isSynced() const { return got_chirp_ || frame_ready_; }
    bool isFrameReady() const { return frame_ready_; }
    bool hasPendingData() const { return !buffer_.empty() || got_chirp_ || frame_ready_; }
    float getEstimatedCFO() const { return cfo_hz_; }
    void reset() { buffer_.clear(); soft_.clear(); cfo_hz_ = 0.0f; got_chirp_ = frame_ready_ = false; }   // :372-383
    const McDpskConfig& getConfig() const { return cfg_; }

private:
    McDpskConfig cfg_;
    pu_mcdpsk* h_ = nullptr;
    std::vector<float> buffer_, soft_;
    float cfo_hz_ = 0.0f;
    bool got_chirp_ = false, frame_ready_ = false;
};

#undef PU_IWAVEFORM_BASE
#undef PU_OVERRIDE

}  // namespace pu
