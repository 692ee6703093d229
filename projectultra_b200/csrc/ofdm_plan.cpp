// projectultra_b200/csrc/ofdm_plan.cpp — see ofdm_plan.h.
#include "ofdm_plan.h"

#include <cmath>
#include <random>

namespace pu {

static constexpr double kPi = 3.14159265358979323846;   // M_PI

std::vector<cfloat> OfdmPlan::nco(float freq_hz, size_t n) const {
    // phase_inc_ = 2*pi*f/fs evaluated in double and stored as float; the wrap compares the float phase with the
    // double 2*pi (filters.cpp:228-238).  The float accumulation drifts, so the sequence is tabulated, not recomputed.
    std::vector<cfloat> out(n);
    const float inc = static_cast<float>(2.0f * kPi * freq_hz / static_cast<float>(cfg.sample_rate));
    float phase = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        out[i] = cfloat(std::cos(phase), std::sin(phase));
        phase += inc;
        if (phase > 2.0f * kPi) phase = static_cast<float>(phase - 2.0f * kPi);
        if (phase < 0) phase = static_cast<float>(phase + 2.0f * kPi);
    }
    return out;
}

bool make_ofdm_plan(const pu_modem_config& cfg, OfdmPlan* p, const char** why) {
    static const char* dummy;
    if (!why) why = &dummy;
    *p = OfdmPlan{};
    p->cfg = cfg;
    if (cfg.fft_size != 512 && cfg.fft_size != 1024) { *why = "fft_size must be 512 or 1024"; return false; }
    if (cfg.num_carriers < 2 || cfg.num_carriers > 64) { *why = "num_carriers must be in [2, 64]"; return false; }
    if (cfg.use_pilots && cfg.pilot_spacing < 2) { *why = "pilot_spacing must be >= 2 when pilots are used"; return false; }
    if (cfg.sample_rate == 0) { *why = "sample_rate is zero"; return false; }
    switch (cfg.modulation) {
        case PU_MOD_DBPSK: case PU_MOD_BPSK: case PU_MOD_DQPSK: case PU_MOD_QPSK: case PU_MOD_D8PSK:
        case PU_MOD_QAM16: case PU_MOD_QAM32: case PU_MOD_QAM64: case PU_MOD_QAM256: break;
        default: *why = "modulation has no demapper in the reference (QAM8/AUTO)"; return false;
    }
    p->nfft = static_cast<int>(cfg.fft_size);
    p->log2n = p->nfft == 512 ? 9 : 10;
    const int base_cp = cfg.cp_mode == 0 ? 32 : cfg.cp_mode == 2 ? 64 : 48;   // getCyclicPrefix, types.hpp:197-208
    p->cp = base_cp * (p->nfft / 512);
    p->sym_len = p->nfft + p->cp + static_cast<int>(cfg.symbol_guard);
    p->bps = bits_per_symbol(cfg.modulation);
    switch (cfg.modulation) {
        case PU_MOD_D8PSK: p->ce_margin = 1.1f; break;
        case PU_MOD_QAM16: p->ce_margin = 1.2f; break;
        case PU_MOD_QAM32: p->ce_margin = 1.5f; break;
        case PU_MOD_QAM64: p->ce_margin = 1.8f; break;
        case PU_MOD_QAM256: p->ce_margin = 2.5f; break;
        default: p->ce_margin = 1.0f;
    }

    // carriers -floor(Nc/2) .. +ceil(Nc/2) without DC; every pilot_spacing-th one (counting from the lowest) is a pilot
    struct Slot { int bin; bool pilot_by_spacing; };
    std::vector<Slot> slots;
    const int nc = static_cast<int>(cfg.num_carriers);
    for (int c = -(nc / 2), ord = 0; c <= (nc + 1) / 2; ++c) {
        if (c == 0) continue;
        const int bin = (c + p->nfft) % p->nfft;
        const bool spaced = cfg.pilot_spacing ? (ord % static_cast<int>(cfg.pilot_spacing) == 0) : false;
        slots.push_back({bin, spaced});
        if (cfg.use_pilots && spaced) p->pilot_bin.push_back(bin);
        else p->data_bin.push_back(bin);
        ++ord;
    }
    p->n_data = static_cast<int>(p->data_bin.size());
    p->n_pilot = static_cast<int>(p->pilot_bin.size());

    p->sync_seq.resize(nc);
    for (int n = 0; n < nc; ++n) {
        const float ph = static_cast<float>(-kPi * 1.0 * static_cast<double>(n) * static_cast<double>(n + 1) / static_cast<double>(nc));
        p->sync_seq[n] = cfloat(std::cos(ph), std::sin(ph));
    }
    std::mt19937 gen(0x50494C54u);
    p->pilot_sign.resize(p->n_pilot);
    for (auto& s : p->pilot_sign) s = (gen() & 1) ? 1.0f : -1.0f;

    if (p->n_pilot > 0) {
        std::vector<int> pilot_pos(slots.size(), -1);
        for (size_t s = 0, k = 0; s < slots.size(); ++s)
            if (slots[s].pilot_by_spacing) pilot_pos[s] = static_cast<int>(k++);
        for (size_t s = 0; s < slots.size(); ++s) {
            if (slots[s].pilot_by_spacing) continue;
            int below = -1, above = -1;
            for (int j = static_cast<int>(s) - 1; j >= 0 && below < 0; --j) if (slots[j].pilot_by_spacing) below = j;
            for (size_t j = s + 1; j < slots.size() && above < 0; ++j) if (slots[j].pilot_by_spacing) above = static_cast<int>(j);
            float a = 0.5f;
            if (below >= 0 && above >= 0) a = static_cast<float>(static_cast<int>(s) - below) / static_cast<float>(above - below);
            p->interp_lo.push_back(below >= 0 ? pilot_pos[below] : -1);
            p->interp_hi.push_back(above >= 0 ? pilot_pos[above] : -1);
            p->interp_alpha.push_back(a);
        }
    }

    p->twiddle.resize(p->nfft / 2);
    for (int k = 0; k < p->nfft / 2; ++k) {
        const float ang = static_cast<float>(-2.0f * kPi * static_cast<double>(k) / static_cast<double>(p->nfft));
        p->twiddle[k] = cfloat(std::cos(ang), std::sin(ang));
    }
    return true;
}

}  // namespace pu
