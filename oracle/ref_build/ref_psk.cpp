// oracle/ref_build/ref_psk.cpp — TEST INFRASTRUCTURE, not product code.
//
// extern "C" shim over the UNMODIFIED reference's DPSK classes (src/psk/dpsk.hpp, src/psk/multi_carrier_dpsk.hpp),
// compiled from the sources where they lie.  The private members that findPreamble / setReferenceWithTraining leave
// behind (estimated_cfo_, initial_phase_offset_) are set through `#define private public` so the externally-timed
// demodulation can be pinned with arbitrary values; the reference is not patched.
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
#include "ultra/dsp.hpp"
#include "ultra/logging.hpp"
#define private public
#include "psk/dpsk.hpp"
#include "psk/multi_carrier_dpsk.hpp"
#undef private

using namespace ultra;

namespace {
struct Quiet {
    int saved = -1;
    Quiet() {
        fflush(stderr);
        saved = dup(2);
        int nul = open("/dev/null", O_WRONLY);
        if (nul >= 0) { dup2(nul, 2); close(nul); }
    }
    ~Quiet() {
        fflush(stderr);
        if (saved >= 0) { dup2(saved, 2); close(saved); }
    }
};
DPSKConfig sc_cfg(int mod, int sps, float fc, float fs) {
    DPSKConfig c;
    c.sample_rate = fs;
    c.carrier_freq = fc;
    c.samples_per_symbol = sps;
    c.modulation = mod == 0 ? DPSKModulation::DBPSK : mod == 1 ? DPSKModulation::DQPSK : DPSKModulation::D8PSK;
    return c;
}
MultiCarrierDPSKConfig mc_cfg(int nc, int sps, int bits, float f_lo, float f_hi, float fs, int training) {
    MultiCarrierDPSKConfig c;
    c.sample_rate = fs;
    c.num_carriers = nc;
    c.freq_low = f_lo;
    c.freq_high = f_hi;
    c.samples_per_symbol = sps;
    c.bits_per_symbol = bits;
    c.training_symbols = training;
    return c;
}
}  // namespace

extern "C" {

// DPSKModulator: layout 0 = generatePreamble() (Barker-13 x 3) + modulate(data), as tools/test_dpsk_snr.cpp:47-52;
// layout 1 = generateReferenceSymbol() + modulate(data); layout 2 = modulate(data) only.
long ref_dpsk_tx(int mod, int sps, float fc, float fs, int layout, const uint8_t* data, size_t len, float* out, size_t cap) {
    Quiet q;
    DPSKModulator m(sc_cfg(mod, sps, fc, fs));
    Samples pre;
    if (layout == 0) pre = m.generatePreamble();
    else if (layout == 1) pre = m.generateReferenceSymbol();
    Samples body = m.modulate(ByteSpan(data, len));
    const size_t total = pre.size() + body.size();
    if (total > cap) return -static_cast<long>(total);
    std::memcpy(out, pre.data(), pre.size() * sizeof(float));
    std::memcpy(out + pre.size(), body.data(), body.size() * sizeof(float));
    return static_cast<long>(total);
}

// Externally timed DPSKDemodulator::demodulateSoft.  ref_mode 0: fresh object; 1: setReferenceSymbol on the symbol
// before data_start.  est_cfo / phase_off are written into the (private) members findPreamble would set.
long ref_dpsk_demod_soft_ex(int mod, int sps, float fc, float fs, const float* x, size_t L, long data_start, int ref_mode,
                            float est_cfo, float phase_off, float* llr, size_t cap) {
    Quiet q;
    DPSKDemodulator d(sc_cfg(mod, sps, fc, fs));
    if (ref_mode == 1 && data_start >= sps) d.setReferenceSymbol(SampleSpan(x + data_start - sps, sps));
    d.estimated_cfo_ = est_cfo;
    d.initial_phase_offset_ = phase_off;
    std::vector<float> soft = d.demodulateSoft(SampleSpan(x + data_start, L - data_start));
    for (size_t i = 0; i < soft.size() && i < cap; ++i) llr[i] = soft[i];
    return static_cast<long>(soft.size());
}

// The receive sequence of tools/test_dpsk_snr.cpp:66-73: findPreamble on the whole frame, then demodulateSoft on the span
// that starts at the returned data start.  Returns the number of soft bits (0 when no preamble was found).
long ref_dpsk_receive(int mod, int sps, float fc, float fs, const float* x, size_t L, long* data_start, float* est_cfo,
                      float* phase_off, float* llr, size_t cap) {
    Quiet q;
    DPSKDemodulator d(sc_cfg(mod, sps, fc, fs));
    const int ds = d.findPreamble(SampleSpan(x, L));
    *data_start = ds;
    *est_cfo = d.estimated_cfo_;
    *phase_off = d.initial_phase_offset_;
    if (!(ds > 0 && ds < (int)L)) return 0;
    std::vector<float> soft = d.demodulateSoft(SampleSpan(x + ds, L - ds));
    for (size_t i = 0; i < soft.size() && i < cap; ++i) llr[i] = soft[i];
    return static_cast<long>(soft.size());
}

// MultiCarrierDPSKModulator: generateTrainingSequence() + generateReferenceSymbol() + modulate(data) (no chirp: the
// frame as processGotChirp sees it after an external chirp detection).
long ref_mcdpsk_tx(int nc, int sps, int bits, float f_lo, float f_hi, float fs, int training, const uint8_t* data, size_t len,
                   float* out, size_t cap) {
    Quiet q;
    MultiCarrierDPSKModulator m(mc_cfg(nc, sps, bits, f_lo, f_hi, fs, training));
    Samples tr = m.generateTrainingSequence();
    Samples rf = m.generateReferenceSymbol();
    Samples body = m.modulate(Bytes(data, data + len));
    const size_t total = tr.size() + rf.size() + body.size();
    if (total > cap) return -static_cast<long>(total);
    std::memcpy(out, tr.data(), tr.size() * sizeof(float));
    std::memcpy(out + tr.size(), rf.data(), rf.size() * sizeof(float));
    std::memcpy(out + tr.size() + rf.size(), body.data(), body.size() * sizeof(float));
    return static_cast<long>(total);
}

// Legacy API of MultiCarrierDPSKDemodulator (multi_carrier_dpsk.hpp:390-472): processTraining -> setReference ->
// demodulateSoft on a frame [training][reference][data].  *residual_cfo = cfo_hz_ after processTraining (starts at 0).
long ref_mcdpsk_demod_soft(int nc, int sps, int bits, float f_lo, float f_hi, float fs, int training, const float* x, size_t L,
                           float* llr, size_t cap, float* residual_cfo) {
    Quiet q;
    MultiCarrierDPSKDemodulator d(mc_cfg(nc, sps, bits, f_lo, f_hi, fs, training));
    const size_t tr = static_cast<size_t>(training) * sps;
    if (L < tr + sps) return -1;
    d.processTraining(SampleSpan(x, tr));
    if (residual_cfo) *residual_cfo = d.cfo_hz_;
    d.setReference(SampleSpan(x + tr, sps));
    std::vector<float> soft = d.demodulateSoft(SampleSpan(x + tr + sps, L - tr - sps));
    for (size_t i = 0; i < soft.size() && i < cap; ++i) llr[i] = soft[i];
    return static_cast<long>(soft.size());
}

// MultiCarrierDPSKDemodulator behind an externally detected chirp, as MCDPSKWaveform::process drives it
// (src/waveform/mc_dpsk_waveform.cpp:144-170): setChirpDetected(cfo) -> process(training + ref + data) -> getSoftBits().
// processGotChirp (multi_carrier_dpsk.hpp:533-627) applies the Hilbert-FIR CFO correction when |cfo| > 0.1 Hz.
long ref_mcdpsk_got_chirp(int nc, int sps, int bits, float f_lo, float f_hi, float fs, int training, const float* x, size_t L,
                          float chirp_cfo, float* llr, size_t cap, int* ready_out, float* cfo_after) {
    Quiet q;
    MultiCarrierDPSKDemodulator d(mc_cfg(nc, sps, bits, f_lo, f_hi, fs, training));
    d.setChirpDetected(chirp_cfo);
    const bool ready = d.process(SampleSpan(x, L));
    if (ready_out) *ready_out = ready ? 1 : 0;
    if (cfo_after) *cfo_after = d.getEstimatedCFO();
    if (!ready) return 0;
    std::vector<float> soft = d.getSoftBits();
    for (size_t i = 0; i < soft.size() && i < cap; ++i) llr[i] = soft[i];
    return static_cast<long>(soft.size());
}

// The IWaveform receive sequence (tools/test_iwaveform.cpp:127-160) on an MC-DPSK frame [chirp pair][training][ref][data], with the
// glue of MCDPSKWaveform::detectSync / setFrequencyOffset / process (src/waveform/mc_dpsk_waveform.cpp:100-170) applied to the
// reference's own ChirpSync and MultiCarrierDPSKDemodulator objects (the waveform class itself pulls the protocol layer in):
//   detectDualChirp -> start_sample = up_chirp_start + 2 chirps + 2 gaps -> setChirpDetected(cfo) -> process(span) -> getSoftBits.
// info = {success, up_chirp_start, down_chirp_start, start_sample or -1}, f = {cfo_hz, up correlation, down correlation}.
long ref_mcdpsk_chirp_receive(int nc, int sps, int bits, float f_lo, float f_hi, float fs, int training, const float* x, size_t L,
                              float threshold, int32_t* info, float* f, float* llr, size_t cap, float* cfo_after) {
    Quiet q;
    fflush(stdout);
    int saved = dup(1), nul = open("/dev/null", O_WRONLY);   // detectDualChirp printf()s on stdout
    if (nul >= 0) { dup2(nul, 1); close(nul); }
    MultiCarrierDPSKConfig cfg = mc_cfg(nc, sps, bits, f_lo, f_hi, fs, training);
    sync::ChirpSync cs(cfg.getChirpConfig());
    auto r = cs.detectDualChirp(SampleSpan(x, L), threshold);
    fflush(stdout);
    if (saved >= 0) { dup2(saved, 1); close(saved); }
    info[0] = r.success ? 1 : 0; info[1] = r.up_chirp_start; info[2] = r.down_chirp_start; info[3] = -1;
    f[0] = r.cfo_hz; f[1] = r.up_correlation; f[2] = r.down_correlation;
    if (cfo_after) *cfo_after = r.cfo_hz;
    if (!r.success) return 0;
    size_t chirp_samples = cs.getChirpSamples();
    size_t gap_samples = static_cast<size_t>(cfg.sample_rate * cfg.getChirpConfig().gap_ms / 1000.0f);
    int start_sample = r.up_chirp_start + 2 * chirp_samples + 2 * gap_samples;      // use_dual_chirp
    info[3] = start_sample;
    if (start_sample >= L) return 0;                                                 // test_iwaveform.cpp:143 (int vs size_t, as there)
    MultiCarrierDPSKDemodulator d(cfg);
    d.setCFO(r.cfo_hz);                                                              // MCDPSKWaveform::setFrequencyOffset
    d.setChirpDetected(r.cfo_hz);
    const bool ready = d.process(SampleSpan(x + start_sample, L - start_sample));
    if (cfo_after) *cfo_after = d.getEstimatedCFO();
    if (!ready) return 0;
    std::vector<float> soft = d.getSoftBits();
    for (size_t i = 0; i < soft.size() && i < cap; ++i) llr[i] = soft[i];
    return static_cast<long>(soft.size());
}

}  // extern "C"
