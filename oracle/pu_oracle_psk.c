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

/* ------------------------------------------------------------------ Barker-13 acquisition of the single-carrier DPSK waveform
 * TEST INFRASTRUCTURE (DPSK half).  DPSKDemodulator::findPreamble (src/psk/dpsk.hpp:338-481) with
 * computeDifferentialScore (:489-546), estimateCFOTolerant (:550-593), estimateInitialPhaseOffset (:659-704) and
 * refineTimingWithMatchedFilter (:708-771), in the reference's order of floating-point operations. */
#define BK_SYMS 39   /* BARKER_LEN * REPEATS */
#define BK_DIFFS 38
static const int BARKER13[13] = {1, 1, 1, 1, 1, -1, -1, 1, 1, -1, 1, -1, 1};

typedef struct { const float* x; size_t L; int sps; const float* ccos; const float* csin; float fc, fs; int pattern[BK_DIFFS]; } bk_t;

static float bk_score(const bk_t* b, int start, float min_energy) {   /* computeDifferentialScore, :489-546 */
    float sr[BK_SYMS], si[BK_SYMS], total_energy = 0;
    for (int s = 0; s < BK_SYMS; ++s) {
        dpsk_correlate(b->x + start + s * b->sps, b->sps, b->ccos, b->csin, &sr[s], &si[s]);
        total_energy += sr[s] * sr[s] + si[s] * si[s];
    }
    if (total_energy < min_energy * BK_SYMS) return 0;
    float cr = 0, ci = 0, magnitude_sum = 0;
    for (int i = 0; i < BK_DIFFS; ++i) {
        float dr, di;
        cmulf(sr[i + 1], si[i + 1], sr[i], -si[i], &dr, &di);
        const float magnitude = hypotf(dr, di);
        if (magnitude < 1e-10f) continue;
        const float nr = dr / magnitude, ni = di / magnitude;
        float er, ei;
        cmulf(nr, ni, (float)b->pattern[i], -0.0f, &er, &ei);   /* diff_norm * std::conj(Complex(expected, 0)) */
        cr += er;
        ci += ei;
        magnitude_sum += magnitude;
    }
    if (magnitude_sum < 1e-10f) return 0;
    return hypotf(cr, ci) / BK_DIFFS;
}

static float bk_cfo_tolerant(const bk_t* b, int start) {   /* estimateCFOTolerant, :550-593 */
    float dr[BK_DIFFS], di[BK_DIFFS];
    int nd = 0;
    float pr = 0, pi = 0;
    for (int s = 0; s <= BK_DIFFS; ++s) {
        float cr, ci;
        dpsk_correlate(b->x + start + s * b->sps, b->sps, b->ccos, b->csin, &cr, &ci);
        if (s > 0 && hypotf(pr, pi) > 0.01f && hypotf(cr, ci) > 0.01f) {
            float xr, xi;
            cmulf(cr, ci, pr, -pi, &xr, &xi);
            const float m = hypotf(xr, xi);
            dr[nd] = xr / m;
            di[nd] = xi / m;
            ++nd;
        }
        pr = cr;
        pi = ci;
    }
    if (nd < 10) return 0;
    float cr = 0, ci = 0;
    for (int i = 0; i < nd && i < BK_DIFFS; ++i) {
        float er, ei;
        cmulf(dr[i], di[i], (float)b->pattern[i], -0.0f, &er, &ei);
        cr += er;
        ci += ei;
    }
    const float phase_offset = atan2f(ci, cr);
    const float symbol_duration = (float)b->sps / b->fs;
    const float cfo_hz = (float)(phase_offset / (2.0f * ORC_PI * symbol_duration));
    return -cfo_hz;
}

static int bk_refine_mf(const bk_t* b, int coarse) {   /* refineTimingWithMatchedFilter(cfo_hz = 0), :708-771 */
    const int n = 6 * b->sps;
    float* tmpl = (float*)malloc(sizeof(float) * (size_t)n);
    const float adj = b->fc + 0.0f;
    const float carrier_inc = (float)(2.0f * ORC_PI * adj / b->fs);
    float phase = 0.0f, symbol_phase = 0.0f;
    int k = 0;
    for (int s = 0; s < 6; ++s) {
        if (BARKER13[s] < 0) symbol_phase = (float)(symbol_phase + ORC_PI);
        for (int i = 0; i < b->sps; ++i) {
            tmpl[k++] = cosf(phase + symbol_phase);
            phase += carrier_inc;
            if (phase > 2.0f * ORC_PI) phase = (float)(phase - 2.0f * ORC_PI);
        }
    }
    float template_energy = 0;
    for (int j = 0; j < n; ++j) template_energy += tmpl[j] * tmpl[j];
    int fine_start = coarse - b->sps > 0 ? coarse - b->sps : 0;
    int fine_end = (int)b->L - n < coarse + b->sps ? (int)b->L - n : coarse + b->sps;
    float best_corr = -1;
    int best = coarse;
    for (int i = fine_start; i <= fine_end; ++i) {
        float corr = 0, sig = 0;
        for (int j = 0; j < n; ++j) {
            corr += b->x[i + j] * tmpl[j];
            sig += b->x[i + j] * b->x[i + j];
        }
        const float norm = sqrtf(sig * template_energy);
        if (norm < 1e-10f) continue;
        const float nc = fabsf(corr) / norm;
        if (nc > best_corr) { best_corr = nc; best = i; }
    }
    free(tmpl);
    return best;
}

static float bk_initial_phase(const bk_t* b, int start, float est_cfo) {   /* estimateInitialPhaseOffset, :659-704 */
    float errs[16];
    int ne = 0;
    float pr = 0, pi = 0;
    for (int s = 0; s <= 10 && s < BK_DIFFS; ++s) {
        const int offset = start + s * b->sps;
        if (offset + b->sps > (int)b->L) break;
        float cr, ci;
        dpsk_correlate(b->x + offset, b->sps, b->ccos, b->csin, &cr, &ci);
        if (s > 0 && hypotf(pr, pi) > 0.01f && hypotf(cr, ci) > 0.01f) {
            float dr, di;
            cmulf(cr, ci, pr, -pi, &dr, &di);
            float measured = atan2f(di, dr);
            const float expected = b->pattern[s - 1] > 0 ? 0.0f : (float)ORC_PI;
            const float cfo_phase = (float)(2.0f * ORC_PI * est_cfo * b->sps / b->fs);
            measured -= cfo_phase;
            float error = measured - expected;
            while (error > ORC_PI) error = (float)(error - 2.0f * ORC_PI);
            while (error < -ORC_PI) error = (float)(error + 2.0f * ORC_PI);
            errs[ne++] = error;
        }
        pr = cr;
        pi = ci;
    }
    if (ne == 0) return 0.0f;
    float sum = 0;
    for (int i = 0; i < ne; ++i) sum += errs[i];
    return sum / (float)ne;
}

/* Returns data_start (> 0) or -1; *est_cfo / *phase_off = the members findPreamble leaves behind (0 when it fails early). */
long orc_dpsk_find_preamble(int sps, float fc, float fs, const float* x, size_t L, float* est_cfo, float* phase_off) {
    *est_cfo = 0.0f;
    *phase_off = 0.0f;
    if (sps <= 0) return -1;
    const int preamble_samples = BK_SYMS * sps;
    if ((int)L < preamble_samples + preamble_samples / 2) return -1;
    float energy = 0;
    const size_t check = L < (size_t)(preamble_samples * 2) ? L : (size_t)(preamble_samples * 2);
    for (size_t i = 0; i < check; ++i) energy += x[i] * x[i];
    const float rms = sqrtf(energy / (float)check);
    if (rms < 0.01f) return -1;
    float* ccos = (float*)malloc(sizeof(float) * (size_t)sps * 2);
    float* csin = ccos + sps;
    const float carrier_inc = (float)(2.0f * ORC_PI * fc / fs);        /* :315 */
    for (int i = 0; i < sps; ++i) {
        const float phase = carrier_inc * (float)i;
        csin[i] = sinf(phase);
        ccos[i] = cosf(phase);
    }
    bk_t b = {x, L, sps, ccos, csin, fc, fs, {0}};
    for (int s = 1; s < BK_SYMS; ++s) b.pattern[s - 1] = BARKER13[s % 13];
    const int limit = preamble_samples * 4;
    const int max_search = (int)L - preamble_samples < limit ? (int)L - preamble_samples : limit;
    float best_score = 0, sum_scores = 0;
    int best_offset = -1, num_scores = 0;
    for (int start = 0; start < max_search; start += sps) {
        const float score = bk_score(&b, start, 0.001f);
        sum_scores += score;
        num_scores++;
        if (score > best_score) { best_score = score; best_offset = start; }
    }
    const float global_avg = num_scores > 0 ? sum_scores / (float)num_scores : 0;
    if (best_offset >= 0 && best_score > 0.80f * 0.7f) {
        const int fine_start = best_offset - sps > 0 ? best_offset - sps : 0;
        const int fine_end = max_search < best_offset + sps ? max_search : best_offset + sps;
        for (int start = fine_start; start < fine_end; ++start) {
            const float score = bk_score(&b, start, 0.001f);
            if (score > best_score) { best_score = score; best_offset = start; }
        }
    }
    long result = -1;
    if (!(best_score < 0.80f) && !(global_avg > 0 && best_score < global_avg * 1.3f)) {
        const float cfo = bk_cfo_tolerant(&b, best_offset);
        *est_cfo = cfo;
        if (fabsf(cfo) < 0.5f) best_offset = bk_refine_mf(&b, best_offset);
        *phase_off = bk_initial_phase(&b, best_offset, cfo);
        result = (long)best_offset + preamble_samples;
    }
    free(ccos);
    return result;
}

/* ------------------------------------------------------------------ MC-DPSK behind an externally detected chirp
 * TEST INFRASTRUCTURE.  MultiCarrierDPSKDemodulator::processGotChirp with external_chirp_detected_ (multi_carrier_dpsk.hpp:
 * 533-627) as MCDPSKWaveform::process drives it (src/waveform/mc_dpsk_waveform.cpp:144-170): the whole buffer (training +
 * reference + data) is frequency-shifted through the 127-tap Hilbert FIR when |cfo| > 0.1 Hz (applyCFOCorrection :633-658,
 * HilbertTransform src/dsp/filters.cpp:266-317: Blackman-windowed taps, real path delayed by 63 samples), then
 * processTraining / the |cfo| > 5 Hz rejection (:576-598) / setReference / demodulateSoft. */
static void mc_cfo_correct(float* x, size_t L, float cfo_hz, float fs) {   /* applyCFOCorrection, :633-658 */
    if (fabsf(cfo_hz) < 0.01f || L < 128) return;
    enum { TAPS = 127 };
    float coeffs[TAPS], delay[TAPS];
    const int M = (TAPS - 1) / 2;
    for (int n = 0; n < TAPS; ++n) {                                     /* HilbertTransform ctor, filters.cpp:266-291 */
        const int k = n - M;
        if (k == 0) coeffs[n] = 0;
        else if (k % 2 != 0) coeffs[n] = (float)(2.0f / (ORC_PI * k));
        else coeffs[n] = 0;
        const float w = (float)(2.0f * ORC_PI * n / (TAPS - 1));
        coeffs[n] *= 0.42f - 0.5f * cosf(w) + 0.08f * cosf(2.0f * w);
        delay[n] = 0;
    }
    size_t idx = 0;
    const float phase_inc = (float)(-2.0f * ORC_PI * cfo_hz / fs);
    float phase = 0.0f;                                                  /* cfo_initial_phase_ of a fresh object */
    for (size_t i = 0; i < L; ++i) {
        delay[idx] = x[i];                                               /* HilbertTransform::process, filters.cpp:293-317 */
        float q = 0;
        size_t j = idx;
        for (size_t k = 0; k < TAPS; ++k) {
            q += coeffs[k] * delay[j];
            if (j == 0) j = TAPS;
            --j;
        }
        const float real = delay[(idx + TAPS - (size_t)M) % TAPS];
        idx = (idx + 1) % TAPS;
        const float rr = cosf(phase), ri = sinf(phase);
        float sr, si;
        cmulf(real, q, rr, ri, &sr, &si);                                /* analytic[i] * rotation */
        x[i] = sr;
        phase += phase_inc;
        if (phase > ORC_PI) phase = (float)(phase - 2.0f * ORC_PI);
        if (phase < -ORC_PI) phase = (float)(phase + 2.0f * ORC_PI);
    }
}

/* Returns the number of soft bits (0 when the frame is rejected or too short); *cfo_after = cfo_hz_ after processGotChirp */
long orc_mcdpsk_got_chirp(int nc, int sps, int bits_per_symbol, float f_lo, float f_hi, float fs, int training_symbols, const float* x,
                          size_t L, float chirp_cfo, float* llr, size_t cap, float* cfo_after) {
    const size_t pre = (size_t)(training_symbols + 1) * (size_t)sps;
    if (cfo_after) *cfo_after = chirp_cfo;
    /* processGotChirp waits for full_preamble + at least one codeword's worth of symbols when the buffer is not longer than the preamble */
    if (L <= pre) return 0;
    float* buf = (float*)malloc(sizeof(float) * L);
    memcpy(buf, x, sizeof(float) * L);
    float cfo = chirp_cfo;
    if (fabsf(cfo) > 0.1f) { mc_cfo_correct(buf, L, cfo, fs); cfo = 0.0f; }   /* :565-569; applyCFOCorrection resets cfo_hz_ */
    const float dual_chirp_cfo = cfo;                                           /* :572 (taken after the correction) */
    float residual = 0.0f;
    const long n = orc_mcdpsk_demod_soft(nc, sps, bits_per_symbol, f_lo, f_hi, fs, buf, L, training_symbols, llr, cap, &residual);
    free(buf);
    /* processTraining: cfo_hz_ += residual, clamped (:419-421); orc_mcdpsk_demod_soft reports clamp(0 + residual) */
    float after = cfo;
    if (training_symbols >= 2) {
        /* recover the unclamped residual only matters beyond +-50 Hz, far outside the 5 Hz rule below */
        after = fmaxf(-50.0f, fminf(50.0f, cfo + residual));
    }
    const int has_chirp_cfo = fabsf(dual_chirp_cfo) > 0.1f;                     /* :587 */
    if (has_chirp_cfo) after = cfo;                                             /* :592-596 (saved_cfo, taken after the correction) */
    if (cfo_after) *cfo_after = after;
    if (fabsf(dual_chirp_cfo) < 0.1f && fabsf(after) > 5.0f) return 0;          /* false positive: back to IDLE (:599-607) */
    return n;
}

/* The IWaveform receive sequence (tools/test_iwaveform.cpp:127-160) on an MC-DPSK frame: MCDPSKWaveform::detectSync
 * (src/waveform/mc_dpsk_waveform.cpp:100-142: detectDualChirp, start_sample = up_chirp_start + 2 chirps + 2 gaps),
 * setFrequencyOffset(cfo) (:71-76) and process (:144-170: setChirpDetected(cfo) -> process(span) -> getSoftBits).
 * info[4] = {success, up_chirp_start, down_chirp_start, start_sample or -1}; f[3] = {cfo_hz, up corr, down corr}. */
long orc_mcdpsk_chirp_receive(int nc, int sps, int bits_per_symbol, float f_lo, float f_hi, float fs, int training_symbols, const float* x,
                              size_t L, float threshold, int32_t* info, float* f, float* llr, size_t cap, float* cfo_after) {
    orc_chirp_detect_dual(fs, x, L, threshold, info, f);     /* MultiCarrierDPSKConfig::getChirpConfig (:78-88) == the OFDM_CHIRP chirp */
    info[3] = -1;
    if (cfo_after) *cfo_after = f[0];
    if (!info[0]) return 0;
    const size_t chirp_samples = (size_t)(fs * 500.0f / 1000.0f);
    const size_t gap_samples = (size_t)(fs * 100.0f / 1000.0f);
    const int start = (int)((size_t)info[1] + 2 * chirp_samples + 2 * gap_samples);
    info[3] = start;
    if ((size_t)start >= L) return 0;                        /* test_iwaveform.cpp:143: int compared as size_t */
    return orc_mcdpsk_got_chirp(nc, sps, bits_per_symbol, f_lo, f_hi, fs, training_symbols, x + start, L - (size_t)start, f[0], llr, cap,
                                cfo_after);
}
