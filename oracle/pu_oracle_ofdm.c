/* oracle/pu_oracle_ofdm.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see pu_oracle.h).
 * CPU restatement of the reference OFDM layer for the presynced receive path and the matching
 * transmitter: radix-2 FFT, NCO, constellation maps, soft demappers, LTS channel estimate, pilot
 * tracking, interpolation, equaliser, symbol demodulation.  References cited per function:
 * src/dsp/fft.cpp, src/dsp/filters.cpp, src/ofdm/{modulator,demodulator,channel_equalizer}.cpp,
 * src/ofdm/soft_demap.hpp, src/ofdm/demodulator_constants.hpp.
 *
 * C99 `float _Complex` is used so that complex * and / lower to the same libgcc routines
 * (__mulsc3 / __divsc3) that std::complex<float> uses in the reference build. */
#include "pu_oracle.h"
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#undef I /* <complex.h> macro; the demappers use I/Q as plain variable names */
typedef float _Complex cf;
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define MKC(re, im) CMPLXF((re), (im))
static inline float cnorm(cf z) { return crealf(z) * crealf(z) + cimagf(z) * cimagf(z); } /* std::norm */
static inline cf cscale(float s, cf z) { return MKC(crealf(z) * s, cimagf(z) * s); }       /* T * complex<T> */
static inline cf cdivs(cf z, float s) { return MKC(crealf(z) / s, cimagf(z) / s); }        /* complex<T> / T */

enum { M_DBPSK = 0, M_BPSK = 1, M_DQPSK = 2, M_QPSK = 3, M_D8PSK = 4, M_QAM8 = 5, M_QAM16 = 6, M_QAM32 = 7,
       M_QAM64 = 8, M_QAM256 = 10 };

static int bits_per_sym(int mod) { /* getBitsPerSymbol, types.hpp:42-56 */
    switch (mod) {
        case M_DBPSK: case M_BPSK: return 1;
        case M_DQPSK: case M_QPSK: return 2;
        case M_D8PSK: case M_QAM8: return 3;
        case M_QAM16: return 4;
        case M_QAM32: return 5;
        case M_QAM64: return 6;
        case M_QAM256: return 8;
        default: return 1;
    }
}

/* ------------------------------------------------------------------ FFT
 * twiddles: fft.cpp:76-80 (angle computed in double, stored as float, cosf/sinf);
 * fft_impl: fft.cpp:89-121 (bit reversal, in-place radix-2 DIT, inverse scales by 1/N). */
static void fft_twiddles(size_t n, cf* tw) {
    for (size_t k = 0; k < n / 2; ++k) {
        float angle = (float)(-2.0f * M_PI * (double)k / (double)n);
        tw[k] = MKC(cosf(angle), sinf(angle));
    }
}

static void fft_inplace(cf* d, size_t n, const cf* tw, int inverse) {
    size_t j = 0;
    for (size_t i = 0; i + 1 < n; ++i) {
        if (i < j) { cf t = d[i]; d[i] = d[j]; d[j] = t; }
        size_t k = n / 2;
        while (k <= j) { j -= k; k /= 2; }
        j += k;
    }
    for (size_t len = 2; len <= n; len *= 2) {
        size_t half = len / 2, step = n / len;
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < half; ++k) {
                cf w = tw[k * step];
                if (inverse) w = conjf(w);
                cf t = w * d[i + k + half];
                d[i + k + half] = d[i + k] - t;
                d[i + k] = d[i + k] + t;
            }
    }
    if (inverse) {
        float scale = 1.0f / (float)n;
        for (size_t i = 0; i < n; ++i) d[i] = cscale(scale, d[i]);
    }
}

int orc_fft(size_t n, const float* in, float* out, int inverse) {
    cf* tw = (cf*)malloc(sizeof(cf) * n);
    fft_twiddles(n, tw);
    memcpy(out, in, sizeof(float) * 2 * n);
    fft_inplace((cf*)out, n, tw, inverse);
    free(tw);
    return 0;
}

/* ------------------------------------------------------------------ the tools' CFO injector
 * tools/test_iwaveform.cpp:67-118 (applyCFO): zero-padded FFT over the next power of two -> positive frequencies doubled,
 * negative ones zeroed -> inverse FFT = analytic signal; rotation by the float phase recurrence (phase += 2 pi cfo / fs,
 * wrapped against the double pi) and the real part.  Applied by the tools to the clean TX audio before the channel (:501-506). */
int orc_tools_apply_cfo(float* x, size_t n, float cfo_hz, float fs) {
    if (n < 128 || fabsf(cfo_hz) < 0.001f) return 0;                       /* :68 */
    size_t m = 1;
    while (m < n) m *= 2;                                                   /* :74-75 */
    cf* tw = (cf*)malloc(sizeof(cf) * m);
    cf* z = (cf*)calloc(m, sizeof(cf));
    fft_twiddles(m, tw);
    for (size_t i = 0; i < n; ++i) z[i] = MKC(x[i], 0.0f);                  /* :81-84 */
    fft_inplace(z, m, tw, 0);
    for (size_t i = 1; i < m / 2; ++i) z[i] = cscale(2.0f, z[i]);           /* :93-95 */
    for (size_t i = m / 2 + 1; i < m; ++i) z[i] = MKC(0.0f, 0.0f);          /* :96-98 */
    fft_inplace(z, m, tw, 1);
    float phase = 0.0f;
    const float inc = 2.0f * (float)M_PI * cfo_hz / fs;                     /* :106 */
    for (size_t i = 0; i < n; ++i) {
        const cf rot = MKC(cosf(phase), sinf(phase));
        x[i] = crealf(z[i] * rot);                                          /* :109-110 */
        phase += inc;
        if (phase > M_PI) phase = (float)((double)phase - 2.0f * M_PI);     /* :112-113: float -= double */
        else if (phase < -M_PI) phase = (float)((double)phase + 2.0f * M_PI);
    }
    free(z);
    free(tw);
    return 0;
}

/* ------------------------------------------------------------------ NCO
 * filters.cpp:228-238: phase_inc = 2*pi*f/fs (double expr -> float); out = (cos,sin)(phase);
 * phase += inc; wrap compares against the double 2*pi. */
typedef struct { float phase, inc; } nco_t;
static void nco_init(nco_t* o, float freq, float fs) {
    o->phase = 0.0f;
    o->inc = (float)(2.0f * M_PI * freq / fs);
}
static cf nco_next(nco_t* o) {
    cf out = MKC(cosf(o->phase), sinf(o->phase));
    o->phase += o->inc;
    if (o->phase > 2.0f * M_PI) o->phase = (float)(o->phase - 2.0f * M_PI);
    if (o->phase < 0) o->phase = (float)(o->phase + 2.0f * M_PI);
    return out;
}

int orc_nco(float freq, float fs, size_t n, float* out) {
    nco_t o;
    nco_init(&o, freq, fs);
    for (size_t i = 0; i < n; ++i) {
        cf c = nco_next(&o);
        out[2 * i] = crealf(c);
        out[2 * i + 1] = cimagf(c);
    }
    return 0;
}

/* ------------------------------------------------------------------ soft demappers (soft_demap.hpp)
 * constants: demodulator_constants.hpp:22-26,88-107 */
#define MAX_LLR 10.0f
#define MIN_LLR_MAG 0.5f
static float clip_llr(float llr) { /* soft_demap.hpp:22-29 */
    float c = fmaxf(-MAX_LLR, fminf(MAX_LLR, llr));
    if (fabsf(c) < MIN_LLR_MAG) c = (c >= 0) ? MIN_LLR_MAG : -MIN_LLR_MAG;
    return c;
}
static float ce_margin(int mod) { /* getCEErrorMargin, soft_demap.hpp:243-264 */
    switch (mod) {
        case M_D8PSK: case M_QAM8: return 1.1f;
        case M_QAM16: return 1.2f;
        case M_QAM32: return 1.5f;
        case M_QAM64: return 1.8f;
        case M_QAM256: return 2.5f;
        default: return 1.0f;
    }
}
#define QPSK_SCALE 0.7071067811865476f
#define QAM16_THRESHOLD 0.6324555320336759f
#define QAM32_SCALE 0.1961161351381840f
#define QAM64_D2 0.3086067f
#define QAM64_D4 0.6172134f
#define QAM256_D2 0.1290994f
#define QAM256_D4 0.2581989f
#define QAM256_D8 0.5163978f

static int demap_coherent(int mod, cf sym, float nv, float* o) {
    float I = crealf(sym), Q = cimagf(sym);
    switch (mod) {
        case M_BPSK: /* :37-39 */
            o[0] = clip_llr(-2.0f * I / nv);
            return 1;
        case M_QAM16: { /* :49-64 */
            float s = 2.0f / nv;
            o[0] = clip_llr(-s * I);
            o[1] = clip_llr(s * (fabsf(I) - QAM16_THRESHOLD));
            o[2] = clip_llr(-s * Q);
            o[3] = clip_llr(s * (fabsf(Q) - QAM16_THRESHOLD));
            return 4;
        }
        case M_QAM32: { /* :68-121, brute-force max-log over the 4(I) x 8(Q) grid */
            static const float IL[4] = {-3, -1, 1, 3};
            static const int IG[4] = {0, 1, 3, 2};
            static const float QL[8] = {-7, -5, -3, -1, 1, 3, 5, 7};
            static const int QG[8] = {0, 1, 3, 2, 6, 7, 5, 4};
            float s = 2.0f / nv;
            for (int b = 0; b < 5; ++b) {
                int mask = 1 << (4 - b);
                float d0 = 1e10f, d1 = 1e10f;
                for (int qi = 0; qi < 8; ++qi)
                    for (int ii = 0; ii < 4; ++ii) {
                        float pr = IL[ii] * QAM32_SCALE, pi = QL[qi] * QAM32_SCALE;
                        int bits = (QG[qi] << 2) | IG[ii];
                        float dr = I - pr, di = Q - pi;
                        float dist = dr * dr + di * di;
                        if (bits & mask) { if (dist < d1) d1 = dist; }
                        else { if (dist < d0) d0 = dist; }
                    }
                o[b] = clip_llr(s * (d1 - d0));
            }
            return 5;
        }
        case M_QAM64: { /* :124-141 */
            float s = 2.0f / nv;
            o[0] = clip_llr(-s * I);
            o[1] = clip_llr(s * (fabsf(I) - QAM64_D4));
            o[2] = clip_llr(s * (fabsf(fabsf(I) - QAM64_D4) - QAM64_D2));
            o[3] = clip_llr(-s * Q);
            o[4] = clip_llr(s * (fabsf(Q) - QAM64_D4));
            o[5] = clip_llr(s * (fabsf(fabsf(Q) - QAM64_D4) - QAM64_D2));
            return 6;
        }
        case M_QAM256: { /* :144-163 */
            float s = 2.0f / nv;
            o[0] = clip_llr(-s * I);
            o[1] = clip_llr(s * (fabsf(I) - QAM256_D8));
            o[2] = clip_llr(s * (fabsf(fabsf(I) - QAM256_D8) - QAM256_D4));
            o[3] = clip_llr(s * (fabsf(fabsf(fabsf(I) - QAM256_D8) - QAM256_D4) - QAM256_D2));
            o[4] = clip_llr(-s * Q);
            o[5] = clip_llr(s * (fabsf(Q) - QAM256_D8));
            o[6] = clip_llr(s * (fabsf(fabsf(Q) - QAM256_D8) - QAM256_D4));
            o[7] = clip_llr(s * (fabsf(fabsf(fabsf(Q) - QAM256_D8) - QAM256_D4) - QAM256_D2));
            return 8;
        }
        case M_QPSK:
        default: { /* :42-45; demodulateSymbol's default branch also uses demapQPSK (demodulator.cpp:350-354) */
            float s = -2.0f * QPSK_SCALE / nv;
            o[0] = clip_llr(I * s);
            o[1] = clip_llr(Q * s);
            return 2;
        }
    }
}

static int demap_diff(int mod, cf sym, cf prev, float nv, float* o) {
    cf diff = sym * conjf(prev);
    float phase = atan2f(cimagf(diff), crealf(diff));
    float sp = cabsf(sym) * cabsf(prev);
    if (mod == M_DBPSK) { /* :173-187 */
        if (sp < 1e-6f) { o[0] = 0.0f; return 1; }
        float c = cosf(phase);
        o[0] = clip_llr(2.0f * sp * c / nv);
        return 1;
    }
    if (mod == M_DQPSK) { /* :192-213 */
        o[0] = o[1] = 0.0f;
        if (sp < 1e-6f) return 2;
        float scale = 2.0f * sp / nv;
        const float pi = 3.14159265358979f;
        o[0] = clip_llr(scale * sinf(phase + pi / 4));
        o[1] = clip_llr(scale * cosf(2 * phase));
        return 2;
    }
    /* D8PSK :217-237 */
    o[0] = o[1] = o[2] = 0.0f;
    if (sp < 1e-6f) return 3;
    float conf = sp / nv;
    o[0] = clip_llr(conf * sinf(phase));
    o[1] = clip_llr(conf * sinf(2.0f * phase));
    o[2] = clip_llr(conf * sinf(4.0f * phase));
    return 3;
}

int orc_soft_demap(int mod, float re, float im, float pre, float pim, float nv, float* out) {
    if (mod == M_DBPSK || mod == M_DQPSK || mod == M_D8PSK) return demap_diff(mod, MKC(re, im), MKC(pre, pim), nv, out);
    if (mod == M_QAM8) return -1;
    return demap_coherent(mod, MKC(re, im), nv, out);
}

/* ------------------------------------------------------------------ shared modem tables */
#define MAX_CARR 128
#define MAX_FFT 4096
typedef struct {
    orc_modem_config c;
    int nfft, cp, sym_len;
    int n_data, n_pilot;
    int data_idx[MAX_CARR], pilot_idx[MAX_CARR];
    cf sync_seq[MAX_CARR];  /* Zadoff-Chu, length num_carriers */
    cf pilot_seq[MAX_CARR];
    /* interpolation table, demodulator.cpp:137-193 */
    int it_idx[MAX_CARR], it_lo[MAX_CARR], it_hi[MAX_CARR];
    float it_alpha[MAX_CARR];
    int n_interp;
    cf* tw;
} modem_t;

static int cyclic_prefix(const orc_modem_config* c) { /* types.hpp:197-208 */
    int base = c->cp_mode == 0 ? 32 : c->cp_mode == 2 ? 64 : 48;
    return base * (int)(c->fft_size / 512);
}

static int modem_init(modem_t* m, const orc_modem_config* c) {
    memset(m, 0, sizeof(*m));
    m->c = *c;
    m->nfft = (int)c->fft_size;
    if (m->nfft > MAX_FFT || (m->nfft & (m->nfft - 1)) || c->num_carriers > MAX_CARR - 1) return -1;
    m->cp = cyclic_prefix(c);
    m->sym_len = m->nfft + m->cp + (int)c->symbol_guard;
    /* setupCarriers, demodulator.cpp:45-67 == modulator.cpp:143-181 */
    int neg = (int)c->num_carriers / 2, pos = ((int)c->num_carriers + 1) / 2, count = 0;
    int is_pilot_any[MAX_CARR * 2], order_idx[MAX_CARR * 2], n_all = 0;
    for (int i = -neg; i <= pos; ++i) {
        if (i == 0) continue;
        int idx = (i + m->nfft) % m->nfft;
        if (!c->use_pilots) m->data_idx[m->n_data++] = idx;
        else if (count % (int)c->pilot_spacing == 0) m->pilot_idx[m->n_pilot++] = idx;
        else m->data_idx[m->n_data++] = idx;
        /* buildInterpTable ignores use_pilots (demodulator.cpp:151) */
        is_pilot_any[n_all] = (count % (int)c->pilot_spacing == 0);
        order_idx[n_all++] = idx;
        ++count;
    }
    /* generateSequences, demodulator.cpp:69-85 */
    size_t N = c->num_carriers;
    for (size_t n = 0; n < N; ++n) {
        float phase = (float)(-M_PI * 1.0 * (double)n * (double)(n + 1) / (double)N);
        m->sync_seq[n] = MKC(cosf(phase), sinf(phase));
    }
    orc_mt19937 rng;
    orc_mt_seed(&rng, 0x50494C54u);
    for (int i = 0; i < m->n_pilot; ++i) m->pilot_seq[i] = (orc_mt_next(&rng) & 1) ? MKC(1, 0) : MKC(-1, 0);
    /* buildInterpTable, demodulator.cpp:137-193 */
    for (int ci = 0; ci < n_all; ++ci) {
        if (is_pilot_any[ci]) continue;
        int lo = -1, hi = -1, lo_ci = -1, hi_ci = -1;
        for (int j = ci - 1; j >= 0; --j) if (is_pilot_any[j]) { lo = order_idx[j]; lo_ci = j; break; }
        for (int j = ci + 1; j < n_all; ++j) if (is_pilot_any[j]) { hi = order_idx[j]; hi_ci = j; break; }
        float alpha = 0.5f;
        if (lo_ci >= 0 && hi_ci >= 0) {
            float dist = (float)(hi_ci - lo_ci);
            alpha = (dist > 0) ? (float)(ci - lo_ci) / dist : 0.5f;
        }
        m->it_idx[m->n_interp] = order_idx[ci];
        m->it_lo[m->n_interp] = lo;
        m->it_hi[m->n_interp] = hi;
        m->it_alpha[m->n_interp] = alpha;
        m->n_interp++;
    }
    m->tw = (cf*)malloc(sizeof(cf) * (size_t)m->nfft);
    fft_twiddles((size_t)m->nfft, m->tw);
    return 0;
}
static void modem_free(modem_t* m) { free(m->tw); m->tw = NULL; }

/* ------------------------------------------------------------------ transmitter (modulator.cpp) */
static cf map_bits(uint32_t bits, int mod) { /* mapBits, modulator.cpp:76-106 */
    switch (mod) {
        case M_BPSK: return (bits & 1) ? MKC(1, 0) : MKC(-1, 0);
        case M_QAM16: {
            static const float lv[] = {-3, -1, 3, 1};
            const float s = 0.3162277660168379f;
            return MKC(lv[(bits >> 2) & 3] * s, lv[bits & 3] * s);
        }
        case M_QAM32: { /* qam32_point, modulator.cpp:53-73 */
            static const float IL[4] = {-3, -1, 1, 3};
            static const int IG[4] = {0, 1, 3, 2};
            static const float QL[8] = {-7, -5, -3, -1, 1, 3, 5, 7};
            static const int QG[8] = {0, 1, 3, 2, 6, 7, 5, 4};
            int qb = (bits >> 2) & 7, ib = bits & 3, qi = 0, ii = 0;
            for (int i = 0; i < 4; ++i) if (IG[i] == ib) { ii = i; break; }
            for (int i = 0; i < 8; ++i) if (QG[i] == qb) { qi = i; break; }
            return MKC(IL[ii] * QAM32_SCALE, QL[qi] * QAM32_SCALE);
        }
        case M_QAM64: {
            static const float lv[] = {-7, -5, -1, -3, 7, 5, 1, 3};
            const float s = 0.1543033499620919f;
            return MKC(lv[(bits >> 3) & 7] * s, lv[bits & 7] * s);
        }
        case M_QAM256: {
            static const float lv[] = {-15, -13, -9, -11, -1, -3, -7, -5, 15, 13, 9, 11, 1, 3, 7, 5};
            const float s = 0.0645497224367903f;
            return MKC(lv[(bits >> 4) & 15] * s, lv[bits & 15] * s);
        }
        case M_QPSK:
        default: {
            const float q = QPSK_SCALE;
            switch (bits & 3) {
                case 0: return MKC(-q, -q);
                case 1: return MKC(-q, q);
                case 2: return MKC(q, -q);
                default: return MKC(q, q);
            }
        }
    }
}

typedef struct { modem_t* m; nco_t mixer; cf prev[MAX_CARR]; } tx_t;

/* createOFDMSymbol (modulator.cpp:217-270) + complexToReal (:272-283); appends nfft+cp samples */
static size_t tx_symbol(tx_t* t, const cf* data_syms, int n_syms, int with_pilots, float* out) {
    modem_t* m = t->m;
    static _Thread_local cf fd[MAX_FFT];
    for (int i = 0; i < m->nfft; ++i) fd[i] = MKC(0, 0);
    for (int i = 0; i < m->n_data && i < n_syms; ++i) fd[m->data_idx[i]] = data_syms[i];
    if (with_pilots) for (int i = 0; i < m->n_pilot; ++i) fd[m->pilot_idx[i]] = m->pilot_seq[i];
    fft_inplace(fd, (size_t)m->nfft, m->tw, 1);
    size_t o = 0;
    float scale = m->c.output_scale;
    for (int i = m->nfft - m->cp; i < m->nfft; ++i) { cf x = fd[i] * nco_next(&t->mixer); out[o++] = crealf(x) * scale; }
    for (int i = 0; i < m->nfft; ++i) { cf x = fd[i] * nco_next(&t->mixer); out[o++] = crealf(x) * scale; }
    return o;
}

long orc_ofdm_tx(const orc_modem_config* c, int layout, const uint8_t* data, size_t len, float* out, size_t cap) {
    modem_t m;
    if (modem_init(&m, c)) return -1;
    tx_t t;
    t.m = &m;
    nco_init(&t.mixer, (float)c->center_freq + c->tx_cfo_hz, (float)c->sample_rate); /* modulator.cpp:132 */
    for (int i = 0; i < m.n_data; ++i) t.prev[i] = MKC(1, 0);                       /* :488 / :546 */
    size_t o = 0;
    int mod = (int)c->modulation, bpc = bits_per_sym(mod);
    size_t need_syms = (len * 8 + (size_t)(m.n_data * bpc) - 1) / (size_t)(m.n_data * bpc);
    size_t head = layout == 0 ? (size_t)(2 * m.sym_len) : (size_t)(7 * (m.nfft + m.cp));
    if (head + need_syms * (size_t)m.sym_len > cap) { modem_free(&m); return -(long)(head + need_syms * (size_t)m.sym_len); }
    cf lts[MAX_CARR];
    for (int i = 0; i < m.n_data; ++i) lts[i] = m.sync_seq[(size_t)i % c->num_carriers];
    if (layout == 0) { /* generateTrainingSymbols(2), modulator.cpp:534-580 */
        for (int s = 0; s < 2; ++s) {
            o += tx_symbol(&t, lts, m.n_data, 1, out + o);
            for (uint32_t g = 0; g < c->symbol_guard; ++g) { out[o++] = 0.0f; nco_next(&t.mixer); }
        }
    } else { /* generatePreamble, modulator.cpp:479-532: guard + 4 x STS (one STS mixed once, repeated) + 2 x LTS (one LTS repeated) */
        size_t g = (size_t)(m.nfft + m.cp);
        for (size_t i = 0; i < g; ++i) out[o++] = 0.0f;
        cf sts[MAX_CARR];
        for (int i = 0; i < m.n_data; ++i) /* createSchmidlCoxSTS :298-330: even bins only, seq index advances for all */
            sts[i] = (m.data_idx[i] % 2 == 0) ? m.sync_seq[(size_t)i % c->num_carriers] : MKC(0, 0);
        size_t n1 = tx_symbol(&t, sts, m.n_data, 0, out + o);
        for (int r = 1; r < 4; ++r) memcpy(out + o + (size_t)r * n1, out + o, n1 * sizeof(float));
        o += 4 * n1;
        size_t n2 = tx_symbol(&t, lts, m.n_data, 1, out + o);
        memcpy(out + o + n2, out + o, n2 * sizeof(float));
        o += 2 * n2;
    }
    /* modulate, modulator.cpp:348-477 */
    size_t data_idx = 0, bit_idx = 0;
    while (data_idx < len) {
        cf syms[MAX_CARR];
        int ns = 0;
        for (int cidx = 0; cidx < m.n_data && data_idx < len; ++cidx) {
            uint32_t bits = 0;
            for (int b = 0; b < bpc; ++b) {
                bits <<= 1;
                if (data_idx < len) {
                    bits |= (uint32_t)((data[data_idx] >> (7 - bit_idx)) & 1);
                    if (++bit_idx >= 8) { bit_idx = 0; ++data_idx; }
                }
            }
            cf s;
            if (mod == M_DBPSK) {
                s = t.prev[cidx] * ((bits & 1) ? MKC(-1, 0) : MKC(1, 0));
                t.prev[cidx] = s;
            } else if (mod == M_DQPSK) {
                static const float pr[4] = {1, 0, -1, 0}, pi_[4] = {0, 1, 0, -1};
                s = t.prev[cidx] * MKC(pr[bits & 3], pi_[bits & 3]);
                t.prev[cidx] = s;
            } else if (mod == M_D8PSK) {
                const float pi = 3.14159265358979f;
                float angle = (float)(bits & 7) * (pi / 4.0f) + pi / 8.0f;
                s = t.prev[cidx] * MKC(cosf(angle), sinf(angle));
                t.prev[cidx] = s;
            } else {
                s = map_bits(bits, mod);
            }
            syms[ns++] = s;
        }
        while (ns < m.n_data) syms[ns++] = MKC(0, 0);
        o += tx_symbol(&t, syms, m.n_data, 1, out + o);
        for (uint32_t g = 0; g < c->symbol_guard; ++g) { out[o++] = 0.0f; nco_next(&t.mixer); }
    }
    modem_free(&m);
    return (long)o;
}

/* ------------------------------------------------------------------ receiver state (demodulator_impl.hpp) */
typedef struct {
    modem_t* m;
    nco_t mixer;
    cf H[MAX_FFT];               /* channel_estimate */
    float noise_variance, est_snr_lin;
    int snr_symbol_count, symbols_since_sync;
    float freq_offset_hz, freq_offset_filtered, freq_corr_phase;
    cf prev_pilots[MAX_CARR];
    int have_prev;
    cf ppc;                      /* pilot_phase_correction */
    cf cpc;                      /* carrier_phase_correction */
    int cpc_init;
    float timing;                /* timing_offset_samples */
    cf prev_eq[MAX_CARR];        /* dbpsk_prev_equalized */
    int have_prev_eq;
    float cnv[MAX_CARR];         /* carrier_noise_var */
} rx_t;

/* toBaseband (channel_equalizer.cpp:19-57) + extractSymbol (:59-71): mixes ALL sym_len samples (the NCO and the
 * CFO rotator advance over CP and guard too), FFTs samples [cp, cp+nfft). */
static void rx_symbol_fft(rx_t* r, const float* x, cf* fd) {
    modem_t* m = r->m;
    static _Thread_local cf bb[MAX_FFT + 1024];
    float inc = (float)(-2.0f * M_PI * r->freq_offset_hz / (float)m->c.sample_rate);
    int rot = fabsf(r->freq_offset_hz) > 0.01f;
    for (int i = 0; i < m->sym_len; ++i) {
        cf osc = nco_next(&r->mixer);
        cf mixed = cscale(x[i], conjf(osc));
        if (rot) {
            cf corr = MKC(cosf(r->freq_corr_phase), sinf(r->freq_corr_phase));
            mixed = mixed * corr;
            r->freq_corr_phase += inc;
            if (r->freq_corr_phase > M_PI) r->freq_corr_phase = (float)(r->freq_corr_phase - 2.0f * M_PI);
            else if (r->freq_corr_phase < -M_PI) r->freq_corr_phase = (float)(r->freq_corr_phase + 2.0f * M_PI);
        }
        bb[i] = mixed;
    }
    for (int i = 0; i < m->nfft; ++i) fd[i] = bb[m->cp + i];
    fft_inplace(fd, (size_t)m->nfft, m->tw, 0);
}

/* interpolateChannel, channel_equalizer.cpp:601-631 */
static void rx_interpolate(rx_t* r) {
    modem_t* m = r->m;
    for (int d = 0; d < m->n_interp; ++d) {
        int lo = m->it_lo[d], hi = m->it_hi[d], idx = m->it_idx[d];
        if (lo >= 0 && hi >= 0) {
            cf H1 = r->H[lo], H2 = r->H[hi];
            cf pd = H2 * conjf(H1);
            float ph = fabsf(atan2f(cimagf(pd), crealf(pd)));
            if (ph > 1.5708f) r->H[idx] = (m->it_alpha[d] < 0.5f) ? H1 : H2;
            else r->H[idx] = cscale(1.0f - m->it_alpha[d], H1) + cscale(m->it_alpha[d], H2);
        } else if (lo >= 0) r->H[idx] = r->H[lo];
        else if (hi >= 0) r->H[idx] = r->H[hi];
    }
}

/* updateChannelEstimate, channel_equalizer.cpp:330-595 */
static void rx_update_channel(rx_t* r, const cf* fd) {
    modem_t* m = r->m;
    int np = m->n_pilot, mod = (int)m->c.modulation;
    float alpha = (r->snr_symbol_count == 0) ? 1.0f : 0.9f;
    cf h[MAX_CARR];
    cf h_sum = MKC(0, 0);
    for (int i = 0; i < np; ++i) {
        h[i] = fd[m->pilot_idx[i]] / m->pilot_seq[i];
        h_sum = h_sum + h[i];
    }
    if (!r->cpc_init && np > 0) { /* :348-357 */
        cf h_avg = cdivs(h_sum, (float)np);
        float mag = cabsf(h_avg);
        if (mag > 0.01f) {
            r->cpc = cdivs(conjf(h_avg), mag);
            r->cpc_init = 1;
        }
    }
    for (int i = 0; i < np; ++i) h[i] = h[i] * r->cpc;
    h_sum = h_sum * r->cpc;
    float sp_sum = 0.0f;
    for (int i = 0; i < np; ++i) sp_sum += cnorm(h[i]);
    float signal_power = sp_sum / (float)np;

    float noise_sum = 0.0f;
    size_t noise_count = 0;
    for (int i = 0; i < np; ++i) { /* :395-412 */
        int idx = m->pilot_idx[i];
        if (r->have_prev) {
            cf ph = r->prev_pilots[i], ch = h[i];
            if (cnorm(ph) > 1e-6f && cnorm(ch) > 1e-6f) {
                cf d = ch - ph;
                noise_sum += cnorm(d);
                noise_count++;
            }
        }
        cf h_old = r->H[idx];
        r->H[idx] = cscale(alpha, h[i]) + cscale(1.0f - alpha, h_old);
    }
    if (noise_count == 0) { /* :415-418 */
        noise_sum = signal_power / 31.6f;
        noise_count = 1;
    }
    if (r->have_prev) { /* :421-470 */
        cf pd_sum = MKC(0, 0);
        int valid = 0;
        for (int i = 0; i < np; ++i) {
            cf d = h[i] * conjf(r->prev_pilots[i]);
            if (cnorm(r->prev_pilots[i]) > 1e-6f && cnorm(h[i]) > 1e-6f) {
                float mag = cabsf(d);
                if (mag > 1e-6f) {
                    pd_sum = pd_sum + cdivs(d, mag);
                    valid++;
                }
            }
        }
        if (valid > 0) {
            cf avg = cdivs(pd_sum, (float)valid);
            float apd = atan2f(cimagf(avg), crealf(avg));
            r->ppc = MKC(cosf(-apd), sinf(-apd));
            float sym_dur = (float)m->sym_len / (float)m->c.sample_rate;
            float residual = (float)(apd / (2.0f * M_PI * sym_dur));
            float total = r->freq_offset_hz + residual;
            float a = 0.3f;
            if (r->symbols_since_sync < 10) {
                float progress = (float)r->symbols_since_sync / 10;
                a = 0.9f * (1.0f - progress) + 0.3f * progress;
            }
            if (fabsf(residual) > 10.0f) a = fmaxf(a, 0.9f);
            r->symbols_since_sync++;
            r->freq_offset_filtered = a * total + (1.0f - a) * r->freq_offset_filtered;
            r->freq_offset_hz = fmaxf(-90.0f, fminf(90.0f, r->freq_offset_filtered));
        }
    } else {
        r->ppc = MKC(1, 0);
    }
    if (r->snr_symbol_count >= 3) { /* timing slope, :473-509 */
        float sk = 0, sk2 = 0, sph = 0, skp = 0;
        int cnt = 0;
        for (int i = 0; i < np; ++i) {
            if (cnorm(h[i]) < 1e-6f) continue;
            int k = m->pilot_idx[i];
            if (k > m->nfft / 2) k -= m->nfft;
            float ph = cargf(h[i]);
            sk += (float)k;
            sk2 += (float)(k * k);
            sph += ph;
            skp += (float)k * ph;
            cnt++;
        }
        if (cnt >= 3) {
            float n = (float)cnt;
            float denom = n * sk2 - sk * sk;
            if (fabsf(denom) > 1e-6f) {
                float slope = (n * skp - sk * sph) / denom;
                float inst = (float)(slope * (float)m->nfft / (2.0f * M_PI));
                r->timing = 0.3f * inst + (1.0f - 0.3f) * r->timing;
                float maxt = 50.0f * ((float)m->nfft / 512.0f);
                r->timing = fmaxf(-maxt, fminf(maxt, r->timing));
            }
        }
    }
    for (int i = 0; i < np; ++i) r->prev_pilots[i] = h[i];
    r->have_prev = 1;

    int coherent = (mod != M_DBPSK && mod != M_DQPSK && mod != M_D8PSK);
    int fix = coherent && fabsf(r->timing) > 0.1f;
    if (fix) /* :525-546 */
        for (int i = 0; i < np; ++i) {
            int idx = m->pilot_idx[i], k = idx;
            if (k > m->nfft / 2) k -= m->nfft;
            float tp = (float)(2.0f * M_PI * (double)k * r->timing / (float)m->nfft);
            r->H[idx] = r->H[idx] * cexpf(MKC(0, -tp));
        }
    rx_interpolate(r);
    if (fix) { /* :552-567 */
        for (int i = 0; i < np; ++i) {
            int idx = m->pilot_idx[i], k = idx;
            if (k > m->nfft / 2) k -= m->nfft;
            float tp = (float)(2.0f * M_PI * (double)k * r->timing / (float)m->nfft);
            r->H[idx] = r->H[idx] * cexpf(MKC(0, tp));
        }
        for (int i = 0; i < m->n_data; ++i) {
            int idx = m->data_idx[i], k = idx;
            if (k > m->nfft / 2) k -= m->nfft;
            float tp = (float)(2.0f * M_PI * (double)k * r->timing / (float)m->nfft);
            r->H[idx] = r->H[idx] * cexpf(MKC(0, tp));
        }
    }
    if (noise_count > 1 && noise_sum > 0.0f) { /* :584-592 */
        r->noise_variance = noise_sum / (float)(noise_count - 1);
        if (r->noise_variance < 1e-6f) r->noise_variance = 1e-6f;
        float inst = signal_power / r->noise_variance;
        inst = fmaxf(0.1f, fminf(10000.0f, inst));
        r->est_snr_lin = 0.3f * inst + (1.0f - 0.3f) * r->est_snr_lin;
    }
    r->snr_symbol_count++;
}

/* equalize, channel_equalizer.cpp:728-840 (adaptive LMS/RLS branch is off by default, types.hpp:170; not restated) */
static void rx_equalize(rx_t* r, const cf* fd, cf* eq) {
    modem_t* m = r->m;
    int mod = (int)m->c.modulation, nd = m->n_data;
    if (mod == M_DBPSK || mod == M_DQPSK || mod == M_D8PSK) { /* :736-771 */
        for (int i = 0; i < nd; ++i) {
            int idx = m->data_idx[i];
            cf rx = fd[idx], h = r->H[idx];
            float hp = cnorm(h);
            int k = idx;
            if (k > m->nfft / 2) k -= m->nfft;
            float tp = (float)(2.0f * M_PI * (double)k * r->timing / (float)m->nfft);
            cf tc = cexpf(MKC(0, tp));
            if (hp > 1e-6f) {
                eq[i] = cdivs(rx * conjf(h), hp) * r->ppc * tc;
                r->cnv[i] = r->noise_variance / hp;
            } else {
                eq[i] = rx * r->ppc * tc;
                r->cnv[i] = 100.0f;
            }
            r->cnv[i] = fmaxf(1e-6f, fminf(100.0f, r->cnv[i]));
        }
        return;
    }
    for (int i = 0; i < nd; ++i) { /* :806-818 */
        int idx = m->data_idx[i];
        cf rx = fd[idx], h = r->H[idx];
        float hp = cnorm(h);
        float den = hp + r->noise_variance;
        if (den < 1e-10f) {
            eq[i] = MKC(0, 0);
            r->cnv[i] = 100.0f;
        } else {
            eq[i] = cdivs(conjf(h) * rx, den);
            r->cnv[i] = r->noise_variance / (hp + 1e-6f);
            r->cnv[i] = fmaxf(1e-6f, fminf(100.0f, r->cnv[i]));
        }
    }
    float avg = 0.0f; /* fade erasure :823-837 */
    for (int i = 0; i < nd; ++i) avg += cnorm(r->H[m->data_idx[i]]);
    avg /= (float)nd;
    float thr = 0.1f * avg;
    for (int i = 0; i < nd; ++i)
        if (cnorm(r->H[m->data_idx[i]]) < thr) r->cnv[i] = 100.0f;
}

/* demodulateSymbol, demodulator.cpp:199-435, including the literal decision-directed tracker block (:362-434) */
static size_t rx_demod_symbol(rx_t* r, const cf* eq, float* out) {
    modem_t* m = r->m;
    int mod = (int)m->c.modulation, nd = m->n_data;
    float margin = ce_margin(mod);
    size_t o = 0;
    if ((mod == M_DQPSK || mod == M_D8PSK || mod == M_DBPSK) && !r->have_prev_eq) { /* :248-277, :286-288; lts_carrier_phases == (1,0) (channel_equalizer.cpp:300) */
        for (int i = 0; i < nd; ++i) r->prev_eq[i] = MKC(1, 0);
        r->have_prev_eq = 1;
    }
    for (int i = 0; i < nd; ++i) {
        float nv = r->cnv[i] * margin;
        if (mod == M_DBPSK || mod == M_DQPSK || mod == M_D8PSK) {
            o += (size_t)demap_diff(mod, eq[i], r->prev_eq[i], nv, out + o);
            r->prev_eq[i] = eq[i];
        } else {
            int n = demap_coherent(mod, eq[i], nv, out + o);
            for (int b = 0; b < n; ++b) out[o + (size_t)b] *= 1.0f; /* llr_sign == +1 (:225,233) */
            o += (size_t)n;
        }
    }
    if ((mod == M_DQPSK || mod == M_D8PSK) && r->have_prev_eq && r->snr_symbol_count >= 1) { /* :362-434 */
        cf pe_sum = MKC(0, 0);
        int valid = 0;
        float dd_alpha = (r->snr_symbol_count < 3) ? 0.3f : 0.15f;
        for (int i = 0; i < nd; ++i) {
            int idx = m->data_idx[i];
            cf prev = r->prev_eq[i]; /* already overwritten with eq[i] above (:307,:314) */
            float sp = cabsf(eq[i]) * cabsf(prev);
            if (sp > 0.1f) {
                cf d = eq[i] * conjf(prev);
                float ph = atan2f(cimagf(d), crealf(d));
                float expected;
                if (mod == M_DQPSK) {
                    int q = (int)round(ph * 2.0f / M_PI);
                    q = ((q % 4) + 4) % 4;
                    expected = (float)(q * M_PI / 2.0f);
                } else {
                    int q = (int)round(ph * 4.0f / M_PI);
                    q = ((q % 8) + 8) % 8;
                    expected = (float)(q * M_PI / 4.0f);
                }
                float pe = ph - expected;
                while (pe > M_PI) pe = (float)(pe - 2 * M_PI);
                while (pe < -M_PI) pe = (float)(pe + 2 * M_PI);
                float maxe = (mod == M_DQPSK) ? 0.7f : 0.35f;
                if (fabsf(pe) < maxe) r->H[idx] = r->H[idx] * MKC(cosf(-pe * dd_alpha), sinf(-pe * dd_alpha));
                pe_sum = pe_sum + cscale(sp, MKC(cosf(pe), sinf(pe)));
                valid++;
            }
        }
        if (valid >= 5) {
            float ape = atan2f(cimagf(pe_sum), crealf(pe_sum));
            cf corr = MKC(cosf(-ape), sinf(-ape));
            float a = (r->snr_symbol_count < 5) ? 0.5f : 0.2f;
            r->ppc = cscale(powf(cabsf(corr), a), r->ppc) * MKC(cosf(a * cargf(corr)), sinf(a * cargf(corr)));
            float mag = cabsf(r->ppc);
            if (mag > 0.01f) r->ppc = cdivs(r->ppc, mag);
        }
    }
    return o;
}

/* estimateCFOFromTraining(samples, num_symbols, coarse_cfo_hz = 0), src/ofdm/ofdm_sync.cpp:278-380: correlation of the FFT parts of the
 * first two training symbols after a LOCAL mixer (NCO restarted at phase 0) has brought them to baseband. */
float orc_ofdm_training_cfo(const orc_modem_config* c, const float* samples, size_t L, int num_symbols) {
    modem_t m;
    if (num_symbols < 2 || modem_init(&m, c)) return 0.0f;
    size_t fft_len = (size_t)m.nfft, cp = (size_t)m.cp, sym_len = (size_t)m.sym_len, total = 2 * sym_len;
    if (L < total) { modem_free(&m); return 0.0f; }
    nco_t mix;
    nco_init(&mix, (float)c->center_freq, (float)c->sample_rate);
    cf* bb = (cf*)malloc(total * sizeof(cf));
    for (size_t i = 0; i < total; ++i) bb[i] = cscale(samples[i], conjf(nco_next(&mix)));      /* samples[i] * std::conj(osc) */
    cf P = MKC(0, 0);
    float E1 = 0.0f, E2 = 0.0f;
    for (size_t i = 0; i < fft_len; ++i) {
        cf z1 = bb[cp + i], z2 = bb[sym_len + cp + i];
        P = P + conjf(z1) * z2;
        E1 += cnorm(z1);
        E2 += cnorm(z2);
    }
    free(bb);
    float corr_mag = cabsf(P) / sqrtf(E1 * E2 + 1e-10f);
    float cfo = 0.0f;
    if (corr_mag >= 0.3f) {
        float phase = atan2f(cimagf(P), crealf(P));
        cfo = (float)((double)(phase * (float)c->sample_rate) / ((double)2.0f * M_PI * (double)sym_len));
        float max_cfo = (float)c->sample_rate / (2.0f * (float)sym_len);
        cfo = fmaxf(-max_cfo, fminf(max_cfo, cfo));
    }
    modem_free(&m);
    return cfo;
}

/* processPresynced, demodulator.cpp:854-985 with the oracle recipe's state at entry */
long orc_ofdm_presynced(const orc_modem_config* c, const float* samples, size_t L, int training,
                        int cfo_mode, float cfo_hz, float cfo_phase, float* llr_out, size_t cap,
                        float* snr_db, float* final_cfo, orc_stage_dump* dump) {
    if (cfo_mode < 0 || cfo_mode > 2) return -2;
    if (cfo_mode == 0) {   /* after reset(): chirp_cfo_estimated is false and freq_offset_hz is 0 -> estimateCFOFromTraining (:920-925) */
        cfo_hz = (training >= 2) ? orc_ofdm_training_cfo(c, samples, L, training) : 0.0f;
        cfo_phase = 0.0f;
    }
    modem_t m;
    if (modem_init(&m, c)) return -1;
    if (L < (size_t)m.sym_len) { modem_free(&m); return 0; } /* :864-866 */
    rx_t* r = (rx_t*)calloc(1, sizeof(rx_t));
    r->m = &m;
    nco_init(&r->mixer, (float)c->center_freq, (float)c->sample_rate);
    for (int i = 0; i < m.nfft; ++i) r->H[i] = MKC(1, 0);
    r->noise_variance = 0.1f;
    r->est_snr_lin = 1.0f;
    r->freq_offset_hz = r->freq_offset_filtered = cfo_hz;     /* :805-825 */
    r->freq_corr_phase = (cfo_mode == 2) ? cfo_phase : 0.0f;
    r->ppc = MKC(1, 0);
    r->cpc = MKC(1, 0);
    int nd = m.n_data, np = m.n_pilot, nu = nd + np;
    static _Thread_local cf fd[MAX_FFT];
    if (dump && dump->carriers) {
        for (int i = 0; i < nd; ++i) dump->carriers[i] = m.data_idx[i];
        for (int i = 0; i < np; ++i) dump->carriers[nd + i] = m.pilot_idx[i];
    }
    const float* ptr = samples;
    size_t remaining = L;
    if (training > 0) { /* estimateChannelFromLTS, channel_equalizer.cpp:77-328 */
        cf h_last[MAX_CARR], h_sum_p[MAX_CARR];
        for (int i = 0; i < np; ++i) h_sum_p[i] = MKC(0, 0);
        for (int s = 0; s < training; ++s) {
            rx_symbol_fft(r, ptr + (size_t)s * (size_t)m.sym_len, fd);
            if (dump && dump->lts_bins) {
                for (int i = 0; i < nd; ++i) { dump->lts_bins[(s * nu + i) * 2] = crealf(fd[m.data_idx[i]]); dump->lts_bins[(s * nu + i) * 2 + 1] = cimagf(fd[m.data_idx[i]]); }
                for (int i = 0; i < np; ++i) { dump->lts_bins[(s * nu + nd + i) * 2] = crealf(fd[m.pilot_idx[i]]); dump->lts_bins[(s * nu + nd + i) * 2 + 1] = cimagf(fd[m.pilot_idx[i]]); }
            }
            for (int i = 0; i < nd; ++i) {
                cf tx = m.sync_seq[(size_t)i % c->num_carriers];
                h_last[i] = MKC(0, 0);
                if (cabsf(tx) > 0.01f) h_last[i] = fd[m.data_idx[i]] / tx;
            }
            for (int i = 0; i < np; ++i) {
                cf tx = m.pilot_seq[i];
                if (cabsf(tx) > 0.01f) h_sum_p[i] = h_sum_p[i] + fd[m.pilot_idx[i]] / tx;
            }
        }
        for (int i = 0; i < nd; ++i) r->H[m.data_idx[i]] = h_last[i];               /* :179-185 */
        float inv = 1.0f / (float)training;
        for (int i = 0; i < np; ++i) r->H[m.pilot_idx[i]] = cscale(inv, h_sum_p[i]); /* :188-194 (h_sum * inv_count) */
        float mag_sum = 0.0f;                                                       /* :208-225 */
        for (int i = 0; i < nd; ++i) mag_sum += cabsf(r->H[m.data_idx[i]]);
        float mag_avg = mag_sum / (float)nd;
        if (mag_avg > 1e-6f && r->noise_variance > 1e-10f) {
            float sp = mag_avg * mag_avg;
            r->est_snr_lin = fmaxf(0.1f, fminf(10000.0f, sp / r->noise_variance));
        }
        r->snr_symbol_count = training;                                             /* :327 */
        ptr += (size_t)training * (size_t)m.sym_len;
        remaining -= (size_t)training * (size_t)m.sym_len;
    }
    if (dump && dump->h_lts) {
        for (int i = 0; i < nd; ++i) { dump->h_lts[2 * i] = crealf(r->H[m.data_idx[i]]); dump->h_lts[2 * i + 1] = cimagf(r->H[m.data_idx[i]]); }
        for (int i = 0; i < np; ++i) { dump->h_lts[2 * (nd + i)] = crealf(r->H[m.pilot_idx[i]]); dump->h_lts[2 * (nd + i) + 1] = cimagf(r->H[m.pilot_idx[i]]); }
    }
    size_t o = 0;
    long ns = 0;
    float tmp[MAX_CARR * 8];
    cf eq[MAX_CARR];
    while (remaining >= (size_t)m.sym_len) { /* :958-976 */
        float cfo_used = r->freq_offset_hz;
        rx_symbol_fft(r, ptr, fd);
        if (np > 0) rx_update_channel(r, fd);
        rx_equalize(r, fd, eq);
        size_t n = rx_demod_symbol(r, eq, tmp);
        for (size_t i = 0; i < n; ++i) {
            if (o < cap) llr_out[o] = tmp[i];
            ++o;
        }
        if (dump && ns < dump->max_sym) {
            if (dump->bins) {
                for (int i = 0; i < nd; ++i) { dump->bins[(ns * nu + i) * 2] = crealf(fd[m.data_idx[i]]); dump->bins[(ns * nu + i) * 2 + 1] = cimagf(fd[m.data_idx[i]]); }
                for (int i = 0; i < np; ++i) { dump->bins[(ns * nu + nd + i) * 2] = crealf(fd[m.pilot_idx[i]]); dump->bins[(ns * nu + nd + i) * 2 + 1] = cimagf(fd[m.pilot_idx[i]]); }
            }
            if (dump->h) {
                for (int i = 0; i < nd; ++i) { dump->h[(ns * nu + i) * 2] = crealf(r->H[m.data_idx[i]]); dump->h[(ns * nu + i) * 2 + 1] = cimagf(r->H[m.data_idx[i]]); }
                for (int i = 0; i < np; ++i) { dump->h[(ns * nu + nd + i) * 2] = crealf(r->H[m.pilot_idx[i]]); dump->h[(ns * nu + nd + i) * 2 + 1] = cimagf(r->H[m.pilot_idx[i]]); }
            }
            if (dump->eq) for (int i = 0; i < nd; ++i) { dump->eq[(ns * nd + i) * 2] = crealf(eq[i]); dump->eq[(ns * nd + i) * 2 + 1] = cimagf(eq[i]); }
            if (dump->nv) for (int i = 0; i < nd; ++i) dump->nv[ns * nd + i] = r->cnv[i];
            if (dump->scalars) {
                float* sc = dump->scalars + ns * ORC_STAGE_SCALARS;
                sc[0] = cfo_used; sc[1] = r->freq_offset_hz; sc[2] = r->noise_variance; sc[3] = r->timing;
                sc[4] = r->est_snr_lin; sc[5] = crealf(r->ppc); sc[6] = cimagf(r->ppc);
                sc[7] = crealf(r->cpc); sc[8] = cimagf(r->cpc); sc[9] = (float)r->snr_symbol_count;
            }
        }
        ptr += m.sym_len;
        remaining -= (size_t)m.sym_len;
        ++ns;
    }
    if (snr_db) *snr_db = 10.0f * log10f(r->est_snr_lin);
    if (final_cfo) *final_cfo = r->freq_offset_hz;
    free(r);
    modem_free(&m);
    return o > cap ? -(long)o : (long)o;
}

int orc_ofdm_presynced_batch(const orc_modem_config* c, const float* samples, size_t B, size_t L, int training,
                             int cfo_mode, const float* cfo_hz, const float* cfo_phase,
                             float* llr_out, size_t stride, int32_t* counts) {
    float* tmp = (float*)malloc(sizeof(float) * 16384);
    for (size_t b = 0; b < B; ++b) {
        long n = orc_ofdm_presynced(c, samples + b * L, L, training, cfo_mode, cfo_hz ? cfo_hz[b] : 0.0f,
                                    cfo_phase ? cfo_phase[b] : 0.0f, tmp, 16384, NULL, NULL, NULL);
        if (n < 0) { free(tmp); return (int)n; }
        size_t take = (size_t)n < stride ? (size_t)n : stride;
        memcpy(llr_out + b * stride, tmp, take * sizeof(float));
        counts[b] = (int32_t)take;
    }
    free(tmp);
    return 0;
}

double orc_time_presynced_decode(const orc_modem_config* c, const float* samples, size_t B, size_t L, int rate,
                                 uint8_t* info_out, size_t info_stride, uint8_t* ok) {
    struct timespec a, b2;
    float* llr = (float*)malloc(sizeof(float) * 16384);
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (size_t b = 0; b < B; ++b) {
        long n = orc_ofdm_presynced(c, samples + b * L, L, 2, 1, 0.0f, 0.0f, llr, 16384, NULL, NULL, NULL);
        ok[b] = 0;
        if (n >= 648) {
            int s = 0, it = 0;
            uint8_t tmp[128];
            long nb = orc_ldpc_decode_soft(rate, 50, llr, 648, tmp, sizeof(tmp), &s, &it);
            ok[b] = (uint8_t)s;
            memcpy(info_out + b * info_stride, tmp, (size_t)nb < info_stride ? (size_t)nb : info_stride);
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &b2);
    free(llr);
    return (double)(b2.tv_sec - a.tv_sec) + 1e-9 * (double)(b2.tv_nsec - a.tv_nsec);
}

/* ------------------------------------------------------------------ Schmidl-Cox acquisition
 * TEST INFRASTRUCTURE.  OFDMDemodulator::process fed in `chunk`-sample pieces, then getSoftBits():
 *   SEARCHING state            src/ofdm/demodulator.cpp:474-600
 *   hasMinimumEnergy           src/ofdm/ofdm_sync.cpp:20-50
 *   toAnalytic                 :56-84
 *   measureSchmidlCoxCorrelation :118-163
 *   estimateCoarseCFO          :230-261
 *   refineLTSTiming            :386-461
 *   LTS passband templates     src/ofdm/demodulator.cpp:99-132
 *   SYNCED state               src/ofdm/demodulator.cpp:665-690 (= the presynced symbol loop without training symbols)
 * Restricted to L <= 2 * OVERLAP_SAMPLES (40000): beyond it the reference trims its buffer between calls.
 * Pinned against the compiled reference by tests/test_oracle_ofdm.py::test_process_path_matches_reference. */
typedef struct {
    modem_t* m;
    const float* x; /* the frame */
    size_t size;    /* rx_buffer.size() at the current call */
    float noise_floor;
    float* lts_i; float* lts_q; int lts_len;
} acq_t;

static void acq_analytic(const modem_t* m, const float* s, size_t len, cf* out) { /* toAnalytic, ofdm_sync.cpp:56-84 (len == fft size) */
    for (size_t i = 0; i < len; ++i) out[i] = MKC(s[i], 0);
    fft_inplace(out, len, m->tw, 0);
    for (size_t i = 1; i < len / 2; ++i) out[i] = cscale(2.0f, out[i]);
    for (size_t i = len / 2 + 1; i < len; ++i) out[i] = MKC(0, 0);
    fft_inplace(out, len, m->tw, 1);
}

static int acq_min_energy(acq_t* a, size_t offset, size_t window) { /* ofdm_sync.cpp:20-50 */
    if (offset + window > a->size) return 0;
    float sum_sq = 0;
    size_t count = 0;
    for (size_t i = 0; i < window; i += 16) { float s = a->x[offset + i]; sum_sq += s * s; ++count; }
    float energy = sum_sq / (float)count;
    if (a->noise_floor < 1e-20f) a->noise_floor = energy * 0.1f;
    if (energy < a->noise_floor) a->noise_floor = energy;
    else if (energy < a->noise_floor * 3.0f) a->noise_floor = (1.0f - 0.01f) * a->noise_floor + 0.01f * energy;
    return energy >= a->noise_floor * 4.0f;
}

static float acq_sc_corr(acq_t* a, size_t offset) { /* measureSchmidlCoxCorrelation, ofdm_sync.cpp:118-163 */
    const modem_t* m = a->m;
    size_t fft_len = (size_t)m->nfft, half = fft_len / 2, cp = (size_t)m->cp;
    if (offset + cp + fft_len > a->size) return 0.0f;
    const float* w = a->x + offset + cp;
    float dc_sum = 0.0f;
    for (size_t i = 0; i < fft_len; ++i) dc_sum += w[i];
    float dc = dc_sum / (float)fft_len;
    static _Thread_local float tmp[MAX_FFT];
    static _Thread_local cf an[MAX_FFT];
    for (size_t i = 0; i < fft_len; ++i) tmp[i] = w[i] - dc;
    acq_analytic(m, tmp, fft_len, an);
    cf P = MKC(0, 0);
    float R1 = 0.0f, R2 = 0.0f;
    for (size_t i = 0; i < half; ++i) {
        P = P + conjf(an[i]) * an[i + half];
        R1 += cnorm(an[i]);
        R2 += cnorm(an[i + half]);
    }
    float normalization = sqrtf(R1 * R2);
    if (normalization < 1e-10f) return 0.0f;
    return cabsf(P) / normalization;
}

static float acq_coarse_cfo(acq_t* a, size_t sync_offset) { /* estimateCoarseCFO, ofdm_sync.cpp:230-261 */
    const modem_t* m = a->m;
    size_t fft_len = (size_t)m->nfft, half = fft_len / 2, start = sync_offset + (size_t)m->cp;
    if (start + fft_len > a->size) return 0.0f;
    static _Thread_local cf an[MAX_FFT];
    acq_analytic(m, a->x + start, fft_len, an);
    cf P = MKC(0, 0);
    for (size_t i = 0; i < half; ++i) P = P + conjf(an[i]) * an[i + half];
    float phase = atan2f(cimagf(P), crealf(P));
    float cfo = (float)((double)(phase * (float)m->c.sample_rate) / (M_PI * (double)fft_len));
    float max_cfo = (float)(m->c.sample_rate / (uint32_t)fft_len);
    if (cfo > max_cfo) cfo = max_cfo;
    if (cfo < -max_cfo) cfo = -max_cfo;
    return cfo;
}

static long acq_refine_lts(acq_t* a, size_t coarse_sts) { /* refineLTSTiming, ofdm_sync.cpp:386-461; -1 == SIZE_MAX */
    const modem_t* m = a->m;
    size_t P = (size_t)(m->nfft + m->cp), coarse_lts = coarse_sts + 4 * P;
    int back = (int)(3 * P), fwd = (int)(P / 2);
    if (coarse_lts < (size_t)back || coarse_lts + (size_t)fwd + (size_t)a->lts_len > a->size) return (long)coarse_lts;
    float energy_ref = 0.0f;
    for (int i = 0; i < a->lts_len; ++i) { energy_ref += a->lts_i[i] * a->lts_i[i]; energy_ref += a->lts_q[i] * a->lts_q[i]; }
    energy_ref *= 0.5f;
    float best = 0.0f;
    size_t best_off = coarse_lts;
    for (int delta = -back; delta <= fwd; ++delta) {
        size_t off = coarse_lts + (size_t)(long)delta;
        float ci = 0.0f, cq = 0.0f, er = 0.0f;
        for (int i = 0; i < a->lts_len; ++i) {
            float r = a->x[off + (size_t)i];
            ci += r * a->lts_i[i];
            cq += r * a->lts_q[i];
            er += r * r;
        }
        float mag = sqrtf(ci * ci + cq * cq), norm = sqrtf(er * energy_ref);
        float corr = (norm > 1e-6f) ? mag / norm : 0.0f;
        if (corr > best) { best = corr; best_off = off; }
    }
    float thr = (m->nfft >= 1024) ? 0.05f : 0.35f;
    return best < thr ? -1 : (long)best_off;
}

/* info[4] = {synchronised, last_sync_offset, samples consumed before the first data symbol, process() calls until sync} */
long orc_ofdm_process(const orc_modem_config* c, const float* samples, size_t L, size_t chunk, float sync_threshold,
                      float* llr_out, size_t cap, int32_t* info, float* coarse_cfo) {
    modem_t m;
    if (modem_init(&m, c)) return -1;
    if (L > 40000 || chunk == 0) { modem_free(&m); return -3; }
    if (sync_threshold <= 0.0f) sync_threshold = 0.80f; /* ModemConfig::sync_threshold, types.hpp:188 */
    /* generateSequences, demodulator.cpp:99-132 */
    int P = m.nfft + m.cp;
    static _Thread_local cf lf[MAX_FFT];
    for (int i = 0; i < m.nfft; ++i) lf[i] = MKC(0, 0);
    for (int i = 0; i < m.n_data; ++i) lf[m.data_idx[i]] = m.sync_seq[(size_t)i % c->num_carriers];
    for (int i = 0; i < m.n_pilot; ++i) lf[m.pilot_idx[i]] = m.pilot_seq[i];
    fft_inplace(lf, (size_t)m.nfft, m.tw, 1);
    float* li = (float*)malloc(sizeof(float) * (size_t)P * 2);
    float* lq = li + P;
    nco_t osc;
    nco_init(&osc, (float)c->center_freq, (float)c->sample_rate);
    for (int i = 0; i < P; ++i) {
        cf bb = i < m.cp ? lf[m.nfft - m.cp + i] : lf[i - m.cp];
        cf mixed = bb * nco_next(&osc);
        li[i] = crealf(mixed);
        lq[i] = cimagf(mixed);
    }
    acq_t a = {&m, samples, 0, 0.0f, li, lq, P};
    size_t total = 6 * (size_t)P, window = 2 * (size_t)P;
    int found = 0, calls = 0;
    size_t sync_offset = 0, data_start = 0;
    float cfo = 0.0f;
    for (size_t size = chunk < L ? chunk : L;; size = size + chunk < L ? size + chunk : L) {
        ++calls;
        a.size = size;
        if (size >= 4000) { /* MIN_SEARCH_SAMPLES */
            size_t search_end = size > total + window ? size - total - window : 0, peak_pos = 0;
            int hit = 0;
            for (size_t i = 0; i < search_end; i += 8) {
                if (!acq_min_energy(&a, i, window)) { i += window / 2 - 8; continue; }
                float corr = acq_sc_corr(&a, i);
                if (corr > sync_threshold) {
                    size_t plateau = 0;
                    float peak = corr;
                    peak_pos = i;
                    for (size_t j = 0; j <= 300 && i + j + total < size; j += 8) {
                        float r = acq_sc_corr(&a, i + j);
                        if (r >= 0.90f) plateau++;
                        if (r > peak) { peak = r; peak_pos = i + j; }
                    }
                    if (plateau >= 15) { hit = 1; break; }
                }
            }
            if (hit) {
                float coarse = acq_coarse_cfo(&a, peak_pos);
                long refined = acq_refine_lts(&a, peak_pos);
                if (refined >= 0) {
                    found = 1;
                    sync_offset = peak_pos;
                    cfo = coarse;
                    data_start = (size_t)refined + 2 * (size_t)P;
                    break;
                }
            }
        }
        if (size >= L) break;
    }
    free(li);
    modem_free(&m);
    if (info) { info[0] = found; info[1] = (int32_t)sync_offset; info[2] = (int32_t)data_start; info[3] = calls; }
    if (coarse_cfo) *coarse_cfo = cfo;
    if (!found || data_start >= L) return 0;
    return orc_ofdm_presynced(c, samples + data_start, L - data_start, 0, 1, cfo, 0.0f, llr_out, cap, NULL, NULL, NULL);
}

/* ------------------------------------------------------------------ dual-chirp synchronisation (chirp half)
 * TEST INFRASTRUCTURE.  sync::ChirpSync (src/sync/chirp_sync.hpp) with the configuration OFDMChirpWaveform gives it
 * (src/waveform/ofdm_chirp_waveform.cpp:39-49: 300 -> 2700 Hz, 500 ms, 100 ms gaps, dual chirp):
 *   generate                          :58-108     generateTemplate                  :706-735
 *   detectDualChirp                   :349-506    detectChirpTemplate               :560-629
 *   computeComplexTemplateCorrelation :639-662
 * and the receive glue of OFDMChirpWaveform::detectSync / process (ofdm_chirp_waveform.cpp:129-199) as driven by
 * tools/test_iwaveform.cpp:127-160. */
typedef struct { float fs, f_start, f_end, duration_ms, gap_ms, amplitude; size_t n, gap; float *up_s, *up_c, *dn_s, *dn_c; float up_e, dn_e; } chirp_t;

static void chirp_init(chirp_t* c, float fs) {
    c->fs = fs; c->f_start = 300.0f; c->f_end = 2700.0f; c->duration_ms = 500.0f; c->gap_ms = 100.0f; c->amplitude = 0.5f;
    c->n = (size_t)(c->fs * c->duration_ms / 1000.0f);
    c->gap = (size_t)(c->fs * c->gap_ms / 1000.0f);
    c->up_s = (float*)malloc(sizeof(float) * c->n * 4);
    c->up_c = c->up_s + c->n; c->dn_s = c->up_c + c->n; c->dn_c = c->dn_s + c->n;
    const float T = c->duration_ms / 1000.0f, k = (c->f_end - c->f_start) / T;
    c->up_e = 0.0f;
    for (size_t i = 0; i < c->n; ++i) {   /* generateTemplate, :706-735 */
        const float t = (float)i / c->fs;
        const float phase = (float)(2.0f * M_PI * (c->f_start * t + 0.5f * k * t * t));
        c->up_s[i] = sinf(phase);
        c->up_c[i] = cosf(phase);
        c->up_e += c->up_s[i] * c->up_s[i];
    }
    c->dn_e = 0.0f;
    for (size_t i = 0; i < c->n; ++i) {
        const float t = (float)i / c->fs;
        const float phase = (float)(2.0f * M_PI * (c->f_end * t - 0.5f * k * t * t));
        c->dn_s[i] = sinf(phase);
        c->dn_c[i] = cosf(phase);
        c->dn_e += c->dn_s[i] * c->dn_s[i];
    }
}
static void chirp_free(chirp_t* c) { free(c->up_s); }

long orc_chirp_generate(float fs, float tx_cfo, float* out, size_t cap) {   /* ChirpSync::generate, :58-108 */
    chirp_t c;
    chirp_init(&c, fs);
    const size_t total = 2 * c.n + 2 * c.gap;
    if (total > cap) { chirp_free(&c); return -(long)total; }
    memset(out, 0, sizeof(float) * total);
    const float T = c.duration_ms / 1000.0f, k = (c.f_end - c.f_start) / T;
    const float fu = c.f_start + tx_cfo, fd = c.f_end + tx_cfo;
    for (size_t i = 0; i < c.n; ++i) {
        const float t = (float)i / c.fs;
        const float phase = (float)(2.0f * M_PI * (fu * t + 0.5f * k * t * t));
        out[i] = c.amplitude * sinf(phase);
    }
    for (size_t i = 0; i < c.n; ++i) {
        const float t = (float)i / c.fs;
        const float phase = (float)(2.0f * M_PI * (fd * t - 0.5f * k * t * t));
        out[c.n + c.gap + i] = c.amplitude * sinf(phase);
    }
    chirp_free(&c);
    return (long)total;
}

static float chirp_corr(const float* x, size_t L, size_t off, const float* ts, const float* tc, size_t n, float te) {   /* :639-662 */
    if (off + n > L) return 0.0f;
    float ci = 0.0f, cq = 0.0f, se = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        const float s = x[off + i];
        ci += s * tc[i];
        cq += s * ts[i];
        se += s * s;
    }
    const float denom = sqrtf(se * te);
    if (denom < 1e-10f) return 0.0f;
    return sqrtf(ci * ci + cq * cq) / denom;
}

static int chirp_detect_template(const float* x, size_t L, const float* ts, const float* tc, size_t n, float te, float threshold,
                                 float* corr_out) {   /* detectChirpTemplate, :560-629 */
    *corr_out = 0.0f;
    if (L < n) return -1;
    const size_t search_len = L - n;
    float best = 0.0f;
    int pos_best = -1;
    for (size_t pos = 0; pos < search_len; pos += 48) {
        const float c = chirp_corr(x, L, pos, ts, tc, n, te);
        if (c > best) { best = c; pos_best = (int)pos; }
    }
    *corr_out = best;
    if (pos_best < 0 || best < threshold * 0.3f) return -1;
    const int fine_start = pos_best - 48 > 0 ? pos_best - 48 : 0;
    const int fine_end = (int)search_len < pos_best + 48 ? (int)search_len : pos_best + 48;
    for (int pos = fine_start; pos <= fine_end; ++pos) {
        const float c = chirp_corr(x, L, (size_t)pos, ts, tc, n, te);
        if (c > best) { best = c; pos_best = pos; }
    }
    if (pos_best > 0 && pos_best < (int)search_len - 1) {
        const float c0 = chirp_corr(x, L, (size_t)(pos_best - 1), ts, tc, n, te), c1 = best;
        const float c2 = chirp_corr(x, L, (size_t)(pos_best + 1), ts, tc, n, te);
        const float denom = 2.0f * (c0 - 2.0f * c1 + c2);
        if (fabsf(denom) > 1e-10f) {
            float delta = (c0 - c2) / denom;
            delta = fmaxf(-1.0f, fminf(1.0f, delta));
            pos_best = (int)roundf((float)pos_best + delta);
        }
    }
    *corr_out = best;
    return best >= threshold ? pos_best : -1;
}

/* info[3] = {success, up_chirp_start, down_chirp_start}; f[3] = {cfo_hz, up_correlation, down_correlation} */
int orc_chirp_detect_dual(float fs, const float* x, size_t L, float threshold, int32_t* info, float* f) {   /* :349-506 */
    chirp_t c;
    chirp_init(&c, fs);
    info[0] = 0; info[1] = -1; info[2] = -1;   /* DualChirpResult defaults, chirp_sync.hpp:317-324 */
    f[0] = 0.0f; f[1] = 0.0f; f[2] = 0.0f;
    do {
        if (L < 2 * c.n + c.gap) break;
        float up_corr;
        const int up_pos = chirp_detect_template(x, L, c.up_s, c.up_c, c.n, c.up_e, threshold, &up_corr);
        f[1] = up_corr;
        if (up_pos < 0) break;
        const size_t ds = (size_t)up_pos + c.n / 2, expected = (size_t)up_pos + c.n + c.gap, margin = 2 * c.n;
        size_t de = L < expected + margin ? L : expected + margin;
        if (ds >= L) break;
        if (de <= ds + c.n) de = L < ds + 2 * c.n ? L : ds + 2 * c.n;
        float dn_corr;
        const int rel = chirp_detect_template(x + ds, de - ds, c.dn_s, c.dn_c, c.n, c.dn_e, threshold, &dn_corr);
        if (rel < 0) break;
        const int down_pos = rel + (int)ds;
        f[2] = dn_corr;
        const float T = c.duration_ms / 1000.0f, chirp_rate = (c.f_end - c.f_start) / T, cfo_to_samples = c.fs / chirp_rate;
        const int expected_gap = (int)(c.n + c.gap), actual_gap = down_pos - up_pos;
        const float gap_error = (float)(actual_gap - expected_gap);
        f[0] = gap_error / (2.0f * cfo_to_samples);
        if (fabsf(f[0]) > 100.0f) break;
        const float up_corr_s = f[0] * cfo_to_samples, dn_corr_s = -f[0] * cfo_to_samples;
        info[1] = (int)roundf((float)up_pos + up_corr_s);
        info[2] = (int)roundf((float)down_pos + dn_corr_s);
        info[0] = 1;
    } while (0);
    chirp_free(&c);
    return 0;
}

/* info[4] = {success, up_chirp_start, down_chirp_start, start_sample (training start) or -1} */
long orc_ofdm_chirp_receive(const orc_modem_config* c, const float* x, size_t L, float threshold, int32_t* info, float* cfo_out,
                            float* llr_out, size_t cap) {
    float f[3];
    orc_chirp_detect_dual((float)c->sample_rate, x, L, threshold, info, f);
    info[3] = -1;
    *cfo_out = f[0];
    if (!info[0]) return 0;
    const size_t chirp_samples = (size_t)((float)c->sample_rate * 500.0f / 1000.0f);
    const size_t gap_samples = (size_t)((float)c->sample_rate * 100.0f / 1000.0f);   /* config_.sample_rate * 100.0f / 1000.0f */
    const int start = (int)((size_t)info[2] + chirp_samples + gap_samples);
    info[3] = start;
    if (start < 0 || (size_t)start >= L) return 0;
    const float cfo = f[0];
    float ph = (float)(-2.0f * M_PI * cfo * (float)(size_t)start / (float)c->sample_rate);
    while (ph > M_PI) ph = (float)(ph - 2.0f * M_PI);
    while (ph < -M_PI) ph = (float)(ph + 2.0f * M_PI);
    const long n = orc_ofdm_presynced(c, x + start, L - (size_t)start, 2, 2, cfo, ph, llr_out, cap, NULL, NULL, NULL);
    return n >= 648 ? n : 0;   /* process() only collects soft bits when processPresynced reports a codeword */
}
