// oracle/ref_build/ref_stages.cpp — TEST INFRASTRUCTURE, not product code.
//
// Per-stage dumps of the reference's presynced OFDM receive path.  The reference is not patched:
// this TU pre-includes the standard headers, then opens OFDMDemodulator's pimpl with
// `#define private public` and drives OFDMDemodulator::Impl's own member functions in the same
// order as OFDMDemodulator::processPresynced (src/ofdm/demodulator.cpp:854-985) so that every
// intermediate (FFT bins, channel estimate, CFO/noise/timing trackers, equalised symbols) can be
// compared with the CUDA path stage by stage.
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <random>
#include <span>
#include <string>
#include <vector>
#include <fcntl.h>
#include <unistd.h>

#define private public
#include "ultra/ofdm.hpp"
#undef private
#include "ultra/dsp.hpp"
#include "ultra/logging.hpp"
#include "ofdm/demodulator_impl.hpp"

using namespace ultra;

extern "C" {

struct ref_modem_config {
    uint32_t sample_rate, center_freq, fft_size, num_carriers, cp_mode, symbol_guard,
        pilot_spacing, use_pilots, modulation, code_rate;
    float output_scale, tx_cfo_hz;
};

// Per data symbol record (floats): layout documented in tests/refapi.py
//   [0] freq_offset_hz used to mix THIS symbol   [1] freq_offset_hz after tracking
//   [2] noise_variance after update              [3] timing_offset_samples after update
//   [4] estimated_snr_linear                     [5] pilot_phase_correction.re [6] .im
//   [7] carrier_phase_correction.re [8] .im      [9] snr_symbol_count after
#define REF_STAGE_SCALARS 10

// Returns number of data symbols processed.  Arrays (caller allocated, may be NULL):
//   carriers_out [n_used]            : fft bin index of each used carrier, data carriers first then pilots
//   lts_bins     [training][n_used][2]
//   h_lts        [n_used][2]         : channel_estimate after estimateChannelFromLTS
//   bins         [n_sym][n_used][2]  : FFT output at used carriers
//   h            [n_sym][n_used][2]  : channel_estimate after updateChannelEstimate (or LTS H if no pilots)
//   eq           [n_sym][n_data][2]
//   nv           [n_sym][n_data]
//   scalars      [n_sym][REF_STAGE_SCALARS]
//   llr          [cap]               : all soft bits in order; *n_llr = count
long ref_ofdm_presynced_stages(const ref_modem_config* c, const float* samples, size_t L, int training,
                               int cfo_mode, float cfo_hz, float cfo_phase, int max_sym,
                               int32_t* carriers_out, int32_t* n_data_out, int32_t* n_pilot_out,
                               float* lts_bins, float* h_lts, float* bins, float* h, float* eq, float* nv,
                               float* scalars, float* llr, size_t cap, long* n_llr) {
    setLogLevel(LogLevel::ERROR);
    fflush(stderr);
    int saved = dup(2);
    int nul = open("/dev/null", O_WRONLY);
    if (nul >= 0) { dup2(nul, 2); close(nul); }

    ModemConfig cfg;
    cfg.sample_rate = c->sample_rate;
    cfg.center_freq = c->center_freq;
    cfg.fft_size = c->fft_size;
    cfg.num_carriers = c->num_carriers;
    cfg.cp_mode = static_cast<CyclicPrefixMode>(c->cp_mode);
    cfg.symbol_guard = c->symbol_guard;
    cfg.pilot_spacing = c->pilot_spacing;
    cfg.use_pilots = c->use_pilots != 0;
    cfg.modulation = static_cast<Modulation>(c->modulation);
    cfg.code_rate = static_cast<CodeRate>(c->code_rate);
    cfg.output_scale = c->output_scale;
    cfg.tx_cfo_hz = c->tx_cfo_hz;

    long n_sym = 0;
    {
        OFDMDemodulator d(cfg);
        d.reset();
        if (cfo_mode == 1) d.setFrequencyOffset(cfo_hz);
        else if (cfo_mode == 2) d.setFrequencyOffsetWithPhase(cfo_hz, cfo_phase);
        auto& im = *d.impl_;

        const size_t nd = im.data_carrier_indices.size();
        const size_t np = im.pilot_carrier_indices.size();
        const size_t nu = nd + np;
        if (n_data_out) *n_data_out = (int32_t)nd;
        if (n_pilot_out) *n_pilot_out = (int32_t)np;
        std::vector<int> used(im.data_carrier_indices);
        used.insert(used.end(), im.pilot_carrier_indices.begin(), im.pilot_carrier_indices.end());
        if (carriers_out) for (size_t i = 0; i < nu; ++i) carriers_out[i] = used[i];

        // LTS bins first, with a scratch demodulator in the same state (toBaseband is stateful).
        if (lts_bins && training > 0) {
            OFDMDemodulator d2(cfg);
            d2.reset();
            if (cfo_mode == 1) d2.setFrequencyOffset(cfo_hz);
            else if (cfo_mode == 2) d2.setFrequencyOffsetWithPhase(cfo_hz, cfo_phase);
            auto& i2 = *d2.impl_;
            i2.mixer.reset();
            for (int s = 0; s < training; ++s) {
                auto bb = i2.toBaseband(SampleSpan(samples + s * i2.symbol_samples, i2.symbol_samples));
                auto fd = i2.extractSymbol(bb, 0);
                for (size_t i = 0; i < nu; ++i) {
                    lts_bins[(s * nu + i) * 2] = fd[used[i]].real();
                    lts_bins[(s * nu + i) * 2 + 1] = fd[used[i]].imag();
                }
            }
        }

        // Same resets as processPresynced (demodulator.cpp:869-905); a fresh object already has them,
        // they are repeated for clarity of the contract.
        im.soft_bits.clear();
        im.rx_buffer.clear();
        im.mixer.reset();
        std::fill(im.channel_estimate.begin(), im.channel_estimate.end(), Complex(1, 0));
        im.snr_symbol_count = 0;
        im.estimated_snr_linear = 1.0f;
        im.noise_variance = 0.1f;
        im.symbols_since_sync = 0;
        im.prev_pilot_phases.clear();
        im.pilot_phase_correction = Complex(1, 0);
        im.dbpsk_prev_equalized.clear();
        im.carrier_phase_initialized = false;
        im.carrier_phase_correction = Complex(1, 0);
        im.lts_phase_offset = Complex(1, 0);
        im.state.store(OFDMDemodulator::Impl::State::SYNCED);

        const float* ptr = samples;
        size_t remaining = L;
        if (cfo_mode == 0 && training >= 2 && std::abs(im.freq_offset_hz) < 0.1f) {
            float cfo = im.estimateCFOFromTraining(ptr, training, 0.0f);
            im.freq_offset_hz = cfo;
            im.freq_offset_filtered = cfo;
        }
        if (training > 0) {
            im.estimateChannelFromLTS(ptr, training);
            ptr += training * im.symbol_samples;
            remaining -= training * im.symbol_samples;
        }
        if (h_lts) for (size_t i = 0; i < nu; ++i) {
            h_lts[2 * i] = im.channel_estimate[used[i]].real();
            h_lts[2 * i + 1] = im.channel_estimate[used[i]].imag();
        }
        im.dbpsk_prev_equalized.clear();

        while (remaining >= im.symbol_samples && n_sym < max_sym) {
            float cfo_used = im.freq_offset_hz;
            auto bb = im.toBaseband(SampleSpan(ptr, im.symbol_samples));
            auto fd = im.extractSymbol(bb, 0);
            if (!im.pilot_carrier_indices.empty()) im.updateChannelEstimate(fd);
            auto e = im.equalize(fd);
            im.demodulateSymbol(e, im.config.modulation);
            if (bins) for (size_t i = 0; i < nu; ++i) {
                bins[(n_sym * nu + i) * 2] = fd[used[i]].real();
                bins[(n_sym * nu + i) * 2 + 1] = fd[used[i]].imag();
            }
            if (h) for (size_t i = 0; i < nu; ++i) {
                h[(n_sym * nu + i) * 2] = im.channel_estimate[used[i]].real();
                h[(n_sym * nu + i) * 2 + 1] = im.channel_estimate[used[i]].imag();
            }
            if (eq) for (size_t i = 0; i < nd; ++i) {
                eq[(n_sym * nd + i) * 2] = e[i].real();
                eq[(n_sym * nd + i) * 2 + 1] = e[i].imag();
            }
            if (nv) for (size_t i = 0; i < nd; ++i) nv[n_sym * nd + i] = im.carrier_noise_var[i];
            if (scalars) {
                float* sc = scalars + n_sym * REF_STAGE_SCALARS;
                sc[0] = cfo_used;
                sc[1] = im.freq_offset_hz;
                sc[2] = im.noise_variance;
                sc[3] = im.timing_offset_samples;
                sc[4] = im.estimated_snr_linear;
                sc[5] = im.pilot_phase_correction.real();
                sc[6] = im.pilot_phase_correction.imag();
                sc[7] = im.carrier_phase_correction.real();
                sc[8] = im.carrier_phase_correction.imag();
                sc[9] = (float)im.snr_symbol_count;
            }
            ptr += im.symbol_samples;
            remaining -= im.symbol_samples;
            ++n_sym;
        }
        size_t n = std::min(cap, im.soft_bits.size());
        if (llr) std::memcpy(llr, im.soft_bits.data(), n * sizeof(float));
        if (n_llr) *n_llr = (long)im.soft_bits.size();
    }

    fflush(stderr);
    if (saved >= 0) { dup2(saved, 2); close(saved); }
    return n_sym;
}

}  // extern "C"
