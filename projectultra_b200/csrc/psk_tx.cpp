// projectultra_b200/csrc/psk_tx.cpp — host-side single-carrier and multi-carrier DPSK transmitters used to build the
// TX waveform pool of the link simulation (TX on the GPU is SURVEY §8(f) next-3).
//
// Reference behaviour: DPSKModulator (src/psk/dpsk.hpp:102-307): generatePreamble (:118-153, Barker-13 x 3 in DBPSK),
// generateReferenceSymbol (:158-173), modulate / modulateSymbol (:212-279), buildPulseShape (:289-300);
// MultiCarrierDPSKModulator (src/psk/multi_carrier_dpsk.hpp:91-257): generateTrainingSequence (:118-150),
// generateReferenceSymbol (:153-173), modulate (:176-243).  float/double promotion and operation order follow the
// reference expression by expression so that the waveforms are bit-identical (tests/test_psk_tx.py); the libm calls
// are the host's, as in the reference.
#include <cmath>
#include <complex>
#include <vector>

#include "pu_internal.h"

namespace {

using cfloat = std::complex<float>;
constexpr double kPi = 3.14159265358979323846;   // M_PI

struct ScTx {
    pu_dpsk_config cfg;
    float carrier_phase = 0.0f, symbol_phase = 0.0f;
    std::vector<float> pulse;
    explicit ScTx(const pu_dpsk_config& c) : cfg(c) {
        const int n = static_cast<int>(c.samples_per_symbol);
        pulse.resize(n);
        for (int i = 0; i < n; ++i) {                                   // buildPulseShape, :289-300
            const float t = static_cast<float>(i) / n;
            pulse[i] = static_cast<float>(0.5f * (1.0f - std::cos(2.0f * kPi * t)));
        }
    }
    float carrier_inc() const { return static_cast<float>(2.0f * kPi * cfg.carrier_freq / cfg.sample_rate); }
    int bits_per_symbol() const { return cfg.modulation == 0 ? 1 : cfg.modulation == 1 ? 2 : 3; }
    float phase_increment(int v) const {                                // DPSKConfig::phase_increment, :72-86
        switch (cfg.modulation) {
            case 0: return v ? static_cast<float>(kPi) : 0.0f;
            case 1: return static_cast<float>((v * 2 + 1) * kPi / 4.0f);
            default: return static_cast<float>((v & 7) * kPi / 4.0f + kPi / 8.0f);
        }
    }
    void preamble(std::vector<float>& out) {                            // generatePreamble, :118-153
        static const int barker[13] = {1, 1, 1, 1, 1, -1, -1, 1, 1, -1, 1, -1, 1};
        const float inc = carrier_inc();
        float phase = 0.0f, sym_phase = 0.0f;
        for (int rep = 0; rep < 3; ++rep)
            for (int s = 0; s < 13; ++s) {
                if (barker[s] < 0) sym_phase = static_cast<float>(sym_phase + kPi);
                for (uint32_t i = 0; i < cfg.samples_per_symbol; ++i) {
                    out.push_back(std::cos(phase + sym_phase));
                    phase += inc;
                    if (phase > 2.0f * kPi) phase = static_cast<float>(phase - 2.0f * kPi);
                }
            }
        carrier_phase = phase;
        symbol_phase = sym_phase;
    }
    void reference_symbol(std::vector<float>& out) {                    // generateReferenceSymbol, :158-173
        const float inc = carrier_inc();
        float phase = 0.0f;
        for (uint32_t i = 0; i < cfg.samples_per_symbol; ++i) {
            out.push_back(std::cos(phase));
            phase += inc;
        }
        carrier_phase = phase;
        symbol_phase = 0.0f;
    }
    void symbol(int value, std::vector<float>& out) {                   // modulateSymbol, :249-279 (pulse shaping on, the default)
        symbol_phase += phase_increment(value);
        while (symbol_phase >= 2.0f * kPi) symbol_phase = static_cast<float>(symbol_phase - 2.0f * kPi);
        const float inc = carrier_inc();
        for (uint32_t i = 0; i < cfg.samples_per_symbol; ++i) {
            out.push_back(pulse[i] * std::cos(carrier_phase + symbol_phase));
            carrier_phase += inc;
        }
        while (carrier_phase >= 2.0f * kPi) carrier_phase = static_cast<float>(carrier_phase - 2.0f * kPi);
    }
    void modulate(const uint8_t* data, size_t n, std::vector<float>& out) {   // :212-247
        const int bps = bits_per_symbol();
        std::vector<int> bits;
        for (size_t i = 0; i < n; ++i)
            for (int b = 7; b >= 0; --b) bits.push_back((data[i] >> b) & 1);
        while (bits.size() % bps != 0) bits.push_back(0);
        for (size_t i = 0; i < bits.size(); i += bps) {
            int v = 0;
            for (int b = 0; b < bps; ++b)
                if (bits[i + b]) v |= 1 << (bps - 1 - b);
            symbol(v, out);
        }
    }
};

std::vector<float> mc_freqs(const pu_mcdpsk_config& c) {                // getCarrierFreqs, multi_carrier_dpsk.hpp:56-67
    std::vector<float> f(c.num_carriers);
    if (c.num_carriers == 1) {
        f[0] = (c.freq_low + c.freq_high) / 2.0f;
    } else {
        const float spacing = (c.freq_high - c.freq_low) / static_cast<float>(c.num_carriers - 1);
        for (uint32_t i = 0; i < c.num_carriers; ++i) f[i] = c.freq_low + static_cast<float>(i) * spacing;
    }
    return f;
}

}  // namespace

namespace pu {
std::vector<float> mcdpsk_carrier_freqs(const pu_mcdpsk_config& c) { return mc_freqs(c); }
}  // namespace pu

extern "C" {

pu_status pu_dpsk_tx(const pu_dpsk_config* cfg, int layout, const uint8_t* data, size_t n_bytes, float* out, size_t out_cap,
                     size_t* out_len) {
    PU_REQUIRE(cfg && out_len, "pu_dpsk_tx: NULL argument");
    PU_REQUIRE(cfg->samples_per_symbol > 0 && cfg->modulation <= 2 && cfg->sample_rate > 0, "pu_dpsk_tx: bad configuration");
    PU_REQUIRE(layout >= 0 && layout <= 2, "pu_dpsk_tx: layout must be 0, 1 or 2");
    ScTx tx(*cfg);
    std::vector<float> w;
    if (layout == 0) tx.preamble(w);
    else if (layout == 1) tx.reference_symbol(w);
    if (n_bytes) tx.modulate(data, n_bytes, w);
    *out_len = w.size();
    if (out_cap < w.size() || !out) {
        pu::set_error("pu_dpsk_tx: output buffer too small (%zu samples needed)", w.size());
        return PU_ERR_INVALID;
    }
    std::memcpy(out, w.data(), w.size() * sizeof(float));
    return PU_OK;
}

pu_status pu_mcdpsk_tx(const pu_mcdpsk_config* cfg, const uint8_t* data, size_t n_bytes, float* out, size_t out_cap,
                       size_t* out_len) {
    PU_REQUIRE(cfg && out_len, "pu_mcdpsk_tx: NULL argument");
    PU_REQUIRE(cfg->num_carriers >= 1 && cfg->num_carriers <= 64 && cfg->samples_per_symbol > 0 &&
                   (cfg->bits_per_symbol == 1 || cfg->bits_per_symbol == 2), "pu_mcdpsk_tx: bad configuration");
    const int nc = static_cast<int>(cfg->num_carriers), sps = static_cast<int>(cfg->samples_per_symbol);
    const int bits_c = static_cast<int>(cfg->bits_per_symbol), ntr = static_cast<int>(cfg->training_symbols);
    const std::vector<float> freqs = mc_freqs(*cfg);
    std::vector<cfloat> prev(nc, cfloat(1.0f, 0.0f));
    std::vector<float> w(static_cast<size_t>(ntr + 1) * sps, 0.0f);
    for (int sym = 0; sym < ntr; ++sym)                                  // generateTrainingSequence, :118-150
        for (int c = 0; c < nc; ++c) {
            const float phase_offset = static_cast<float>((c * sym) * kPi / 2.0f);
            const cfloat tsym = std::polar(1.0f, phase_offset);
            const float inc = static_cast<float>(2.0f * kPi * freqs[c] / cfg->sample_rate);
            for (int i = 0; i < sps; ++i) {
                const float t = i * inc;
                const cfloat m = tsym * std::polar(1.0f, t);
                w[static_cast<size_t>(sym) * sps + i] += m.real() / nc;
            }
        }
    for (int c = 0; c < nc; ++c) {                                       // generateReferenceSymbol, :153-173
        const float inc = static_cast<float>(2.0f * kPi * freqs[c] / cfg->sample_rate);
        const cfloat ref(1.0f, 0.0f);
        prev[c] = ref;
        for (int i = 0; i < sps; ++i) {
            const float t = i * inc;
            const cfloat m = ref * std::polar(1.0f, t);
            w[static_cast<size_t>(ntr) * sps + i] += m.real() / nc;
        }
    }
    std::vector<int> bits;                                               // modulate, :176-243
    for (size_t i = 0; i < n_bytes; ++i)
        for (int b = 7; b >= 0; --b) bits.push_back((data[i] >> b) & 1);
    const int per_sym = nc * bits_c;
    const int nsym = static_cast<int>((bits.size() + per_sym - 1) / per_sym);
    bits.resize(static_cast<size_t>(nsym) * per_sym, 0);
    const size_t base = w.size();
    w.resize(base + static_cast<size_t>(nsym) * sps, 0.0f);
    static const float dqpsk_phases[] = {static_cast<float>(kPi / 4), static_cast<float>(3 * kPi / 4),
                                         static_cast<float>(-3 * kPi / 4), static_cast<float>(-kPi / 4)};
    size_t bit_idx = 0;
    for (int sym = 0; sym < nsym; ++sym)
        for (int c = 0; c < nc; ++c) {
            int v = 0;
            for (int b = 0; b < bits_c; ++b) v = (v << 1) | bits[bit_idx++];
            const float change = bits_c == 2 ? dqpsk_phases[v] : (v ? static_cast<float>(kPi) : 0.0f);
            cfloat cur = prev[c] * std::polar(1.0f, change);
            cur /= std::abs(cur);
            prev[c] = cur;
            const float inc = static_cast<float>(2.0f * kPi * freqs[c] / cfg->sample_rate);
            for (int i = 0; i < sps; ++i) {
                const float t = i * inc;
                const cfloat m = cur * std::polar(1.0f, t);
                w[base + static_cast<size_t>(sym) * sps + i] += m.real() / nc;
            }
        }
    *out_len = w.size();
    if (out_cap < w.size() || !out) {
        pu::set_error("pu_mcdpsk_tx: output buffer too small (%zu samples needed)", w.size());
        return PU_ERR_INVALID;
    }
    std::memcpy(out, w.data(), w.size() * sizeof(float));
    return PU_OK;
}

}  // extern "C"
