/* oracle/pu_oracle_psk.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's single-carrier and multi-carrier DPSK demodulators with external timing:
 *   DPSKDemodulator            src/psk/dpsk.hpp:309-323 (carrier tables), :777-787 (correlateSymbol),
 *                              :827-879 (demodulateSoft), :889-892 (setReferenceSymbol), :1002-1052 (phaseToBits)
 *   MultiCarrierDPSKDemodulator src/psk/multi_carrier_dpsk.hpp:390-422 (processTraining), :424-435 (setReference),
 *                              :437-472 (demodulateSoft), :663-678 (demodulateOneSymbol)
 * Parity status: PINNED by tests/test_oracle_psk.py against the unmodified reference (oracle/_ref, ref_psk.cpp) and
 * against golden vectors under tests/golden/ generated from it.  fp32 with the reference's operation order; the libm
 * calls (sinf, cosf, atan2f, hypotf) are the host's, as in the reference.  Compile with -ffp-contract=off. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "pu_oracle.h"

#define ORC_PI 3.14159265358979323846 /* M_PI */

/* GCC lowers std::complex<float> * std::complex<float> to the naive formula (plus a NaN fix-up that never triggers
 * on finite data) */
static void cmulf(float ar, float ai, float br, float bi, float* re, float* im) {
    *re = ar * br - ai * bi;
    *im = ar * bi + ai * br;
}

/* ------------------------------------------------------------------ single-carrier DPSK */
static void dpsk_correlate(const float* x, int n, const float* ccos, const float* csin, float* re, float* im) {
    float I = 0.0f, Q = 0.0f;                         /* dpsk.hpp:777-787 */
    for (int i = 0; i < n; ++i) {
        I += x[i] * ccos[i];
        Q -= x[i] * csin[i];
    }
    *re = I / (float)n;
    *im = Q / (float)n;
}

static int dpsk_phase_to_bits(int mod, float phase, float confidence, float* out) {   /* dpsk.hpp:1002-1052 */
    while (phase < 0) phase = (float)(phase + 2.0f * ORC_PI);
    while (phase >= 2.0f * ORC_PI) phase = (float)(phase - 2.0f * ORC_PI);
    if (mod == 0) {
        out[0] = confidence * cosf(phase);
        return 1;
    }
    out[0] = confidence * sinf(phase);
    out[1] = confidence * sinf(2.0f * phase);
    if (mod == 1) return 2;
    out[2] = confidence * sinf(4.0f * phase);
    return 3;
}

/* mod: 0 DBPSK, 1 DQPSK, 2 D8PSK (enum DPSKModulation, dpsk.hpp:31-35).
 * ref_mode 0: prev_symbol_ = (1,0) (fresh object / reset(), :881-886); 1: setReferenceSymbol on the symbol that
 * precedes data_start (:889-892, what findPreamble does at :470-478).  est_cfo / phase_off are the members
 * estimated_cfo_ / initial_phase_offset_ that findPreamble or setReferenceWithTraining leave behind (:858-865). */
long orc_dpsk_demod_soft(int mod, int sps, float fc, float fs, const float* x, size_t L, long data_start, int ref_mode,
                         float est_cfo, float phase_off, float* llr, size_t cap) {
    if (sps <= 0 || data_start < 0 || (size_t)data_start > L) return -1;
    float* ccos = (float*)malloc(sizeof(float) * (size_t)sps);
    float* csin = (float*)malloc(sizeof(float) * (size_t)sps);
    const float carrier_inc = (float)(2.0f * ORC_PI * fc / fs);        /* :315 */
    for (int i = 0; i < sps; ++i) {
        const float phase = carrier_inc * (float)i;
        csin[i] = sinf(phase);
        ccos[i] = cosf(phase);
    }
    float pr = 1.0f, pi = 0.0f;
    if (ref_mode == 1 && data_start >= sps) dpsk_correlate(x + data_start - sps, sps, ccos, csin, &pr, &pi);
    const int bps = mod == 0 ? 1 : mod == 1 ? 2 : 3;
    const size_t nsym = (L - (size_t)data_start) / (size_t)sps;
    long n = 0;
    for (size_t s = 0; s < nsym; ++s) {
        float cr, ci, dr, di;
        dpsk_correlate(x + (size_t)data_start + s * (size_t)sps, sps, ccos, csin, &cr, &ci);
        cmulf(cr, ci, pr, -pi, &dr, &di);                              /* current * conj(prev), :848 */
        const float magnitude = hypotf(dr, di);                        /* std::abs, :851 */
        float phase = atan2f(di, dr);                                  /* :854 */
        if (fabsf(est_cfo) > 0.5f || fabsf(phase_off) > 0.01f) {       /* :857-865 */
            const float cfo_phase = (float)(2.0f * ORC_PI * est_cfo * sps / fs);
            phase -= cfo_phase;
            phase -= phase_off;
            while (phase > ORC_PI) phase = (float)(phase - 2.0f * ORC_PI);
            while (phase < -ORC_PI) phase = (float)(phase + 2.0f * ORC_PI);
        }
        const float confidence = fminf(magnitude * 10.0f, 5.0f);       /* :868 */
        float bits[3];
        const int nb = dpsk_phase_to_bits(mod, phase, confidence, bits);
        for (int b = 0; b < nb; ++b) {
            if ((size_t)n < cap) llr[n] = bits[b];
            ++n;
        }
        pr = cr;
        pi = ci;
        (void)bps;
    }
    free(ccos);
    free(csin);
    return n;
}

/* ------------------------------------------------------------------ multi-carrier DPSK */
static float mc_carrier_freq(int c, int nc, float f_lo, float f_hi) {   /* getCarrierFreqs, multi_carrier_dpsk.hpp:56-67 */
    if (nc == 1) return (f_lo + f_hi) / 2.0f;
    const float spacing = (f_hi - f_lo) / (float)(nc - 1);
    return f_lo + (float)c * spacing;
}

static void mc_demod_one(const float* x, int sps, float freq, float fs, float* re, float* im) {   /* :663-678 */
    const float phase_inc = (float)(2.0f * ORC_PI * freq / fs);
    float sr = 0.0f, si = 0.0f, phase = 0.0f;
    for (int i = 0; i < sps; ++i) {
        const float mr = 1.0f * cosf(-phase), mi = 1.0f * sinf(-phase);   /* std::polar(1.0f, -phase) */
        sr += x[i] * mr;
        si += x[i] * mi;
        phase += phase_inc;
    }
    *re = sr / (float)sps;
    *im = si / (float)sps;
}

/* Frame = [training_symbols][1 reference symbol][data symbols], starting at x[0] (the layout processGotChirp sees with
 * an externally detected chirp, :533-627).  Returns the number of soft bits; *residual_cfo receives the value
 * processTraining would add to cfo_hz_ (the caller applies the |cfo| > 5 Hz rejection rule of :591-598). */
long orc_mcdpsk_demod_soft(int nc, int sps, int bits_per_symbol, float f_lo, float f_hi, float fs, const float* x, size_t L,
                           int training_symbols, float* llr, size_t cap, float* residual_cfo) {
    if (nc < 1 || nc > 64 || sps <= 0) return -1;
    const size_t pre = (size_t)(training_symbols + 1) * (size_t)sps;
    if (L < pre) return -1;
    float pr[64], pi[64];
    if (residual_cfo) {                               /* processTraining, :390-422 */
        *residual_cfo = 0.0f;
        if (training_symbols >= 2) {
            float sum = 0.0f;
            for (int c = 0; c < nc; ++c) {
                float a0r, a0i, a1r, a1i, dr, di, er, ei;
                const float f = mc_carrier_freq(c, nc, f_lo, f_hi);
                mc_demod_one(x, sps, f, fs, &a0r, &a0i);
                mc_demod_one(x + sps, sps, f, fs, &a1r, &a1i);
                const float expected_phase = (float)((c * 1 - c * 0) * ORC_PI / 2.0f);
                const float xr = 1.0f * cosf(expected_phase), xi = 1.0f * sinf(expected_phase);
                cmulf(a1r, a1i, a0r, -a0i, &dr, &di);
                cmulf(dr, di, xr, -xi, &er, &ei);
                sum += atan2f(ei, er);
            }
            const float avg = sum / (float)nc;
            const float symbol_duration = (float)sps / fs;
            *residual_cfo = (float)(avg / (2.0f * ORC_PI * symbol_duration));
            *residual_cfo = fmaxf(-50.0f, fminf(50.0f, 0.0f + *residual_cfo));   /* cfo_hz_ += residual, clamped (:420-421) */
        }
    }
    const float* ref = x + (size_t)training_symbols * (size_t)sps;     /* setReference, :424-435 */
    for (int c = 0; c < nc; ++c) {
        float r, i;
        mc_demod_one(ref, sps, mc_carrier_freq(c, nc, f_lo, f_hi), fs, &r, &i);
        const float a = hypotf(r, i);
        if (a > 0.001f) { pr[c] = r / a; pi[c] = i / a; }
        else { pr[c] = 1.0f; pi[c] = 0.0f; }
    }
    const float* data = x + pre;
    const size_t nsym = (L - pre) / (size_t)sps;
    long n = 0;
    for (size_t s = 0; s < nsym; ++s) {               /* demodulateSoft, :437-472 */
        for (int c = 0; c < nc; ++c) {
            float r, i, nr, ni, dr, di;
            mc_demod_one(data + s * (size_t)sps, sps, mc_carrier_freq(c, nc, f_lo, f_hi), fs, &r, &i);
            const float mag = hypotf(r, i);
            if (mag > 0.0001f) { nr = r / mag; ni = i / mag; }
            else { nr = 1.0f; ni = 0.0f; }
            cmulf(nr, ni, pr[c], -pi[c], &dr, &di);
            pr[c] = nr;
            pi[c] = ni;
            float phase = atan2f(di, dr);
            const float confidence = mag * (float)nc * 4.0f;
            while (phase < 0) phase = (float)(phase + 2.0f * ORC_PI);
            while (phase >= 2.0f * ORC_PI) phase = (float)(phase - 2.0f * ORC_PI);
            if (bits_per_symbol == 2) {
                const float sb0 = confidence * sinf(phase), sb1 = confidence * sinf(2.0f * phase);
                if ((size_t)n < cap) llr[n] = fmaxf(-10.0f, fminf(10.0f, sb0));
                ++n;
                if ((size_t)n < cap) llr[n] = fmaxf(-10.0f, fminf(10.0f, sb1));
                ++n;
            } else {
                const float sb = confidence * cosf(phase);
                if ((size_t)n < cap) llr[n] = fmaxf(-10.0f, fminf(10.0f, sb));
                ++n;
            }
        }
    }
    return n;
}
