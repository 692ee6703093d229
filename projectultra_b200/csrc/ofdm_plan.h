// projectultra_b200/csrc/ofdm_plan.h — host-side tables of one OFDM mode (carrier map, known sequences,
// interpolation table, FFT twiddles, NCO samples), mirroring what OFDMDemodulator::Impl builds at
// construction (src/ofdm/demodulator.cpp:26-193) and what the modulator builds (src/ofdm/modulator.cpp:129-215).
#pragma once
#include <complex>
#include <cstdint>
#include <vector>

#include "pu/pu_capi.h"

namespace pu {

using cfloat = std::complex<float>;

inline int bits_per_symbol(uint32_t mod) {   // getBitsPerSymbol, include/ultra/types.hpp:42-56
    switch (mod) {
        case PU_MOD_DBPSK: case PU_MOD_BPSK: return 1;
        case PU_MOD_DQPSK: case PU_MOD_QPSK: return 2;
        case PU_MOD_D8PSK: case PU_MOD_QAM8: return 3;
        case PU_MOD_QAM16: return 4;
        case PU_MOD_QAM32: return 5;
        case PU_MOD_QAM64: return 6;
        case PU_MOD_QAM256: return 8;
        default: return 1;
    }
}
inline bool is_differential(uint32_t mod) { return mod == PU_MOD_DBPSK || mod == PU_MOD_DQPSK || mod == PU_MOD_D8PSK; }

struct OfdmPlan {
    pu_modem_config cfg{};
    int nfft = 0, log2n = 0, cp = 0, sym_len = 0;
    int n_data = 0, n_pilot = 0, bps = 0;
    float ce_margin = 1.0f;                 // getCEErrorMargin, src/ofdm/soft_demap.hpp:243-264
    std::vector<int> data_bin, pilot_bin;   // FFT bin of each data / pilot carrier (setupCarriers, demodulator.cpp:45-67)
    std::vector<cfloat> sync_seq;           // Zadoff-Chu u=1 over num_carriers (generateSequences, :69-78)
    std::vector<float> pilot_sign;          // +-1 from mt19937(0x50494C54) & 1 (:80-85)
    // interpolation of each data carrier between the bracketing pilots (buildInterpTable, :137-193);
    // lo/hi are POSITIONS in pilot_bin (or -1)
    std::vector<int> interp_lo, interp_hi;
    std::vector<float> interp_alpha;
    std::vector<cfloat> twiddle;            // exp(-2 pi i k / N), k < N/2, float angle (src/dsp/fft.cpp:76-80)
    // NCO::next() samples for `n` consecutive samples from phase 0 (src/dsp/filters.cpp:228-238)
    std::vector<cfloat> nco(float freq_hz, size_t n) const;
};

// Returns false (and fills *why) for configurations outside the implemented path.
bool make_ofdm_plan(const pu_modem_config& cfg, OfdmPlan* plan, const char** why);

}  // namespace pu
