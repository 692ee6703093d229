// projectultra_b200/csrc/ofdm_tx.cpp — host-side OFDM transmitter used to build the TX waveform pool of the
// link simulation (the GPU path consumes received samples; TX on the GPU is SURVEY §8(f) next-3).
//
// Reference behaviour: OFDMModulator (src/ofdm/modulator.cpp): generateTrainingSymbols (:534-580),
// generatePreamble (:479-532), modulate (:348-477), createOFDMSymbol (:217-270), complexToReal (:272-283),
// mapBits (:76-106), with FFT::inverse = radix-2 DIT with conjugated twiddles and 1/N scaling
// (src/dsp/fft.cpp:89-121).  std::complex<float> and the same operation order are used so the waveform is
// bit-identical to the reference's (checked against the oracle in tests/test_host_tx.py).
#include <cmath>
#include <complex>
#include <vector>

#include "ofdm_plan.h"
#include "pu_internal.h"

namespace pu {

namespace {

void inverse_fft(std::vector<cfloat>& a, const std::vector<cfloat>& tw) {
    const size_t n = a.size();
    for (size_t i = 0, rev = 0; i + 1 < n; ++i) {          // bit-reversal permutation
        if (i < rev) std::swap(a[i], a[rev]);
        size_t bit = n >> 1;
        while (bit <= rev) { rev -= bit; bit >>= 1; }
        rev += bit;
    }
    for (size_t span = 2; span <= n; span <<= 1) {
        const size_t half = span >> 1, stride = n / span;
        for (size_t blk = 0; blk < n; blk += span)
            for (size_t k = 0; k < half; ++k) {
                const cfloat t = std::conj(tw[k * stride]) * a[blk + k + half];
                a[blk + k + half] = a[blk + k] - t;
                a[blk + k] = a[blk + k] + t;
            }
    }
    const float scale = 1.0f / static_cast<float>(n);
    for (auto& v : a) v *= scale;
}

cfloat constellation_point(uint32_t bits, uint32_t mod) {
    switch (mod) {
        case PU_MOD_BPSK: return (bits & 1) ? cfloat(1, 0) : cfloat(-1, 0);
        case PU_MOD_QAM16: {
            static const float lv[4] = {-3, -1, 3, 1};
            const float s = 0.3162277660168379f;
            return cfloat(lv[(bits >> 2) & 3] * s, lv[bits & 3] * s);
        }
        case PU_MOD_QAM32: {   // 4 (I) x 8 (Q) rectangular, Gray coded per axis, scale 1/sqrt(26)
            static const float il[4] = {-3, -1, 1, 3}, ql[8] = {-7, -5, -3, -1, 1, 3, 5, 7};
            const float s = 0.1961161351381840f;
            const uint32_t qb = (bits >> 2) & 7, ib = bits & 3;
            uint32_t qi = 0, ii = 0;
            for (uint32_t i = 0; i < 4; ++i) if ((i ^ (i >> 1)) == ib) { ii = i; break; }
            for (uint32_t i = 0; i < 8; ++i) if ((i ^ (i >> 1)) == qb) { qi = i; break; }
            return cfloat(il[ii] * s, ql[qi] * s);
        }
        case PU_MOD_QAM64: {
            static const float lv[8] = {-7, -5, -1, -3, 7, 5, 1, 3};
            const float s = 0.1543033499620919f;
            return cfloat(lv[(bits >> 3) & 7] * s, lv[bits & 7] * s);
        }
        case PU_MOD_QAM256: {
            static const float lv[16] = {-15, -13, -9, -11, -1, -3, -7, -5, 15, 13, 9, 11, 1, 3, 7, 5};
            const float s = 0.0645497224367903f;
            return cfloat(lv[(bits >> 4) & 15] * s, lv[bits & 15] * s);
        }
        default: {   // QPSK
            const float q = 0.7071067811865476f;
            return cfloat((bits & 2) ? q : -q, (bits & 1) ? q : -q);
        }
    }
}

struct Transmitter {
    const OfdmPlan& p;
    std::vector<cfloat> osc;   // TX mixer samples (center_freq + tx_cfo), consumed in order
    size_t osc_pos = 0;
    std::vector<cfloat> diff_state;
    std::vector<float> out;

    explicit Transmitter(const OfdmPlan& plan, size_t max_samples)
        : p(plan), osc(plan.nco(static_cast<float>(plan.cfg.center_freq) + plan.cfg.tx_cfo_hz, max_samples)),
          diff_state(plan.n_data, cfloat(1, 0)) {}

    // one OFDM symbol: carriers -> IFFT -> cyclic prefix -> mix up, real part, scale
    void emit_symbol(const std::vector<cfloat>& data_syms, bool pilots, std::vector<float>* dst) {
        std::vector<cfloat> fd(p.nfft, cfloat(0, 0));
        for (int i = 0; i < p.n_data && i < static_cast<int>(data_syms.size()); ++i) fd[p.data_bin[i]] = data_syms[i];
        if (pilots)
            for (int i = 0; i < p.n_pilot; ++i) fd[p.pilot_bin[i]] = cfloat(p.pilot_sign[i], 0);
        inverse_fft(fd, p.twiddle);
        const float scale = p.cfg.output_scale;
        auto put = [&](const cfloat& v) {
            const cfloat mixed = v * osc[osc_pos++];
            dst->push_back(mixed.real() * scale);
        };
        for (int i = p.nfft - p.cp; i < p.nfft; ++i) put(fd[i]);
        for (int i = 0; i < p.nfft; ++i) put(fd[i]);
    }
    void emit_guard(std::vector<float>* dst) {
        for (uint32_t g = 0; g < p.cfg.symbol_guard; ++g) { dst->push_back(0.0f); ++osc_pos; }
    }
};

}  // namespace

cfloat ofdm_constellation_point(uint32_t bits, uint32_t mod) { return constellation_point(bits, mod); }   // tables of ofdm_tx_gpu.cu

std::vector<float> ofdm_modulate_frame(const OfdmPlan& p, int layout, const uint8_t* data, size_t n_bytes) {
    const uint32_t mod = p.cfg.modulation;
    const int bpc = p.bps;
    const size_t per_sym = static_cast<size_t>(p.n_data) * bpc;
    const size_t n_sym = (n_bytes * 8 + per_sym - 1) / per_sym;
    Transmitter tx(p, (n_sym + 8) * static_cast<size_t>(p.sym_len));
    std::vector<cfloat> lts(p.n_data);
    for (int i = 0; i < p.n_data; ++i) lts[i] = p.sync_seq[i % p.cfg.num_carriers];

    if (layout == 0) {             // 2 x (LTS + guard): the chirp-synced / presynced layout
        for (int s = 0; s < 2; ++s) {
            tx.emit_symbol(lts, true, &tx.out);
            tx.emit_guard(&tx.out);
        }
    } else {                       // Schmidl-Cox: silence, 4 x STS (one waveform repeated), 2 x LTS (one waveform repeated)
        tx.out.assign(static_cast<size_t>(p.nfft + p.cp), 0.0f);
        std::vector<cfloat> sts(p.n_data);
        for (int i = 0; i < p.n_data; ++i) sts[i] = (p.data_bin[i] % 2 == 0) ? lts[i] : cfloat(0, 0);
        std::vector<float> one;
        tx.emit_symbol(sts, false, &one);
        for (int r = 0; r < 4; ++r) tx.out.insert(tx.out.end(), one.begin(), one.end());
        one.clear();
        tx.emit_symbol(lts, true, &one);
        for (int r = 0; r < 2; ++r) tx.out.insert(tx.out.end(), one.begin(), one.end());
    }

    size_t byte = 0, bit = 0;
    while (byte < n_bytes) {
        std::vector<cfloat> syms;
        syms.reserve(p.n_data);
        for (int c = 0; c < p.n_data && byte < n_bytes; ++c) {
            uint32_t v = 0;
            for (int b = 0; b < bpc; ++b) {          // MSB-first, zero padded when the data runs out mid-carrier
                v <<= 1;
                if (byte < n_bytes) {
                    v |= (data[byte] >> (7 - bit)) & 1u;
                    if (++bit == 8) { bit = 0; ++byte; }
                }
            }
            cfloat s;
            if (mod == PU_MOD_DBPSK) {
                s = tx.diff_state[c] * ((v & 1) ? cfloat(-1, 0) : cfloat(1, 0));
                tx.diff_state[c] = s;
            } else if (mod == PU_MOD_DQPSK) {        // 00 -> 0, 01 -> +90, 10 -> 180, 11 -> 270 degrees
                static const cfloat step[4] = {cfloat(1, 0), cfloat(0, 1), cfloat(-1, 0), cfloat(0, -1)};
                s = tx.diff_state[c] * step[v & 3];
                tx.diff_state[c] = s;
            } else if (mod == PU_MOD_D8PSK) {        // 45-degree steps with a 22.5-degree offset
                const float pi = 3.14159265358979f;
                const float ang = static_cast<float>(v & 7) * (pi / 4.0f) + pi / 8.0f;
                s = tx.diff_state[c] * cfloat(std::cos(ang), std::sin(ang));
                tx.diff_state[c] = s;
            } else {
                s = constellation_point(v, mod);
            }
            syms.push_back(s);
        }
        syms.resize(p.n_data, cfloat(0, 0));         // unused carriers of the last symbol carry nothing
        tx.emit_symbol(syms, true, &tx.out);
        tx.emit_guard(&tx.out);
    }
    return tx.out;
}

}  // namespace pu

extern "C" pu_status pu_ofdm_tx(const pu_modem_config* cfg, int layout, const uint8_t* data, size_t n_bytes,
                                float* out, size_t out_cap, size_t* out_len) {
    PU_REQUIRE(cfg && out_len, "pu_ofdm_tx: NULL argument");
    PU_REQUIRE(layout == 0 || layout == 1, "pu_ofdm_tx: layout must be 0 (training) or 1 (Schmidl-Cox preamble)");
    pu::OfdmPlan plan;
    const char* why = "";
    if (!pu::make_ofdm_plan(*cfg, &plan, &why)) {
        pu::set_error("pu_ofdm_tx: %s", why);
        return PU_ERR_UNSUPPORTED;
    }
    std::vector<float> w = pu::ofdm_modulate_frame(plan, layout, data, n_bytes);
    *out_len = w.size();
    PU_REQUIRE(out && out_cap >= w.size(), "pu_ofdm_tx: output buffer too small");
    std::memcpy(out, w.data(), w.size() * sizeof(float));
    return PU_OK;
}
