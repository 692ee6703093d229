/* oracle/pu_oracle_channel.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see pu_oracle.h).
 *
 * CPU twin of the batched channel simulator.  Two things are restated here:
 *  (1) the REFERENCE's channel algorithm, sim::WattersonChannel (src/sim/hf_channel.hpp:67-168, 258-275):
 *      d = floor(delay_ms*fs/1000) with an effective tap delay of d+1 samples (zero-filled deque of d+1 entries,
 *      :73-78,142-146), a = 1 - exp(-2 pi f_d/fs) (:87-88), both fading taps start at (1,0) (:91-92), fading is
 *      updated BEFORE use each sample (:125-127), only its magnitude multiplies the real signal (:135-136),
 *      out = x g1 |f1| + x[n-d-1] g2 |f2| + sigma N(0,1) (:139-153);
 *  (2) the simulator's own SPECIFICATION of the random stream and of the recurrence evaluation order, which
 *      replaces the reference's implementation-defined mt19937 + std::normal_distribution stream (north_star:
 *      "counter-based per-frame RNG, so any frame can be regenerated on the CPU").  The spec is written out in
 *      DESIGN.md ("Channel RNG") and is implemented independently of projectultra_b200/csrc/pu_rng.cuh:
 *        Philox4x32-10, key = seed (lo, hi), counter = (index, 0, stream, 0), stream 1 fading / 2 noise;
 *        u = ((w >> 9) + 0.5) 2^-23; Box-Muller with the fixed ln / sin / cos polynomials below (IEEE ops only);
 *        fading innovations of sample n: Philox index n>>1, 16-bit half (n&1) of word c, z = (u16 - 32767.5) sqrt(12)/65536
 *        (uniform, unit variance: the tap is a sum of >= 385 of them, Gaussian by the central limit theorem);
 *        noise normal of sample n from index n>>2, word pair (n&3)>>1, cosine branch for even n;
 *        the one-pole recurrence evaluated per 128-sample group: serial over the 4 samples of a lane, Kogge-Stone over lanes.
 * Output is bit-identical to the CUDA kernels (tests/test_channel_gpu.py); against the reference the channel is
 * compared statistically. */
#include "pu_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1,
                 n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    memcpy(out, ctr, 16);
    philox(out, key[0], key[1]);
}

static float uni(uint32_t w) { return ((float)(w >> 9) + 0.5f) * 1.1920928955078125e-07f; }

static float ln_poly(float u) {
    uint32_t b;
    memcpy(&b, &u, 4);
    int e = (int)(b >> 23) - 127;
    uint32_t mb = (b & 0x007fffffu) | 0x3f800000u;
    float m;
    memcpy(&m, &mb, 4);
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    float f = m - 1.0f, z = f * f;
    static const float co[9] = {7.0376836292e-2f, -1.1514610310e-1f, 1.1676998740e-1f, -1.2420140846e-1f,
                                1.4249322787e-1f, -1.6668057665e-1f, 2.0000714765e-1f, -2.4999993993e-1f,
                                3.3333331174e-1f};
    float p = co[0];
    for (int i = 1; i < 9; ++i) p = fmaf(p, f, co[i]);
    float y = (f * z) * p;
    y = fmaf(-0.5f, z, y);
    float fe = (float)e, r = f + y;
    r = fmaf(fe, -2.12194440e-4f, r);
    r = fmaf(fe, 0.693359375f, r);
    return r;
}

static void sincos2pi(float u, float* s_out, float* c_out) {
    float t = u * 4.0f;
    int k = (int)(t + 0.5f);
    float x = (t - (float)k) * 1.57079632679489662f, x2 = x * x;
    float ps = fmaf(fmaf(-1.9515295891e-4f, x2, 8.3321608736e-3f), x2, -1.6666654611e-1f);
    float s = fmaf(x * x2, ps, x);
    float pc = fmaf(fmaf(2.443315711809948e-5f, x2, -1.388731625493765e-3f), x2, 4.166664568298827e-2f);
    float c = fmaf(x2 * x2, pc, fmaf(-0.5f, x2, 1.0f));
    switch (k & 3) {
        case 0: *s_out = s; *c_out = c; break;
        case 1: *s_out = c; *c_out = -s; break;
        case 2: *s_out = -s; *c_out = -c; break;
        default: *s_out = -c; *c_out = s; break;
    }
}

static void bm(uint32_t w0, uint32_t w1, float* zc, float* zs) {
    float r = sqrtf(-2.0f * ln_poly(uni(w0)));
    float s, c;
    sincos2pi(uni(w1), &s, &c);
    *zc = r * c;
    *zs = r * s;
}

/* Gaussian primitives exposed for distribution tests */
float orc_noise_normal(uint64_t seed, uint32_t n) {
    uint32_t c[4] = {n >> 2, 0, 2, 0};
    philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    float zc, zs;
    if (n & 2) bm(c[2], c[3], &zc, &zs);
    else bm(c[0], c[1], &zc, &zs);
    return (n & 1) ? zs : zc;
}

/* the four unit-variance fading innovations of sample n (the name is historical: they are uniform, not normal) */
void orc_fading_normals(uint64_t seed, uint32_t n, float z[4]) {
    uint32_t c[4] = {n >> 1, 0, 1, 0};
    philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    for (int k = 0; k < 4; ++k) {
        uint32_t u = (n & 1) ? (c[k] >> 16) : (c[k] & 0xffffu);
        z[k] = ((float)u - 32767.5f) * 5.2857806906e-05f;
    }
}

/* one frame; returns 0 */
int orc_channel_apply(float delay_ms, float doppler_hz, float g1, float g2, uint32_t fs, int fading, int multipath,
                      int noise, const float* x, size_t L, float noise_std, uint64_t seed, float* y) {
    size_t d = (size_t)(delay_ms * (float)fs / 1000.0f);
    float nd = doppler_hz / (float)fs;
    float alpha = (float)(1.0f - exp(-2.0f * 3.14159265358979323846 * nd));
    float ns = alpha > 0.0f ? sqrtf(1.0f / alpha) : 0.0f;
    float a = 1.0f - alpha, qj[4], qs[5], ql[32];
    float v = a;
    for (int j = 0; j < 4; ++j) { qj[j] = v; v = v * a; }
    v = qj[3];
    for (int s = 0; s < 5; ++s) { qs[s] = v; v = v * v; }
    v = 1.0f;
    for (int l = 0; l < 32; ++l) { ql[l] = v; v = v * qj[3]; }
    float carry[4] = {1.0f, 0.0f, 1.0f, 0.0f};
    for (size_t base = 0; base < L; base += 128) {
        float f[4][128];
        if (fading) {
            for (int c = 0; c < 4; ++c) {
                float s[32][4], A[32];
                for (int l = 0; l < 32; ++l) {
                    for (int j = 0; j < 4; ++j) {
                        size_t n = base + 4 * (size_t)l + (size_t)j;
                        float z[4] = {0, 0, 0, 0};
                        if (n < L) orc_fading_normals(seed, (uint32_t)n, z);
                        float e = alpha * (ns * z[c]);
                        s[l][j] = j == 0 ? e : fmaf(a, s[l][j - 1], e);
                    }
                    A[l] = s[l][3];
                }
                for (int t = 0; t < 5; ++t)
                    for (int l = 31; l >= (1 << t); --l) A[l] = fmaf(qs[t], A[l - (1 << t)], A[l]);
                for (int l = 0; l < 32; ++l) {
                    float e = l ? A[l - 1] : 0.0f;
                    float pl = fmaf(ql[l], carry[c], e);
                    for (int j = 0; j < 4; ++j) f[c][4 * l + j] = fmaf(qj[j], pl, s[l][j]);
                }
                carry[c] = f[c][127];
            }
        }
        for (int l = 0; l < 128 && base + (size_t)l < L; ++l) {
            size_t n = base + (size_t)l;
            float m1 = 1.0f, m2 = 1.0f;
            if (fading) {
                m1 = sqrtf(fmaf(f[0][l], f[0][l], f[1][l] * f[1][l]));
                m2 = sqrtf(fmaf(f[2][l], f[2][l], f[3][l] * f[3][l]));
            }
            float out;
            if (multipath && d > 0) {
                float xd = n >= d + 1 ? x[n - d - 1] : 0.0f;
                out = fmaf(xd * g2, m2, (x[n] * g1) * m1);
            } else {
                out = x[n] * m1;
            }
            if (noise) out = fmaf(noise_std, orc_noise_normal(seed, (uint32_t)n), out);
            y[n] = out;
        }
    }
    return 0;
}

/* WattersonChannel::applyCFO (src/sim/hf_channel.hpp:173-232) with cfo_phase_ = 0 at entry: mix to baseband around 1500 Hz, 48-tap
 * running-sum lowpass, rotate by the CFO phase recurrence, mix back up.  Frames shorter than 256 samples are returned unchanged. */
int orc_channel_apply_cfo(const float* x, size_t L, float cfo_hz, uint32_t sample_rate, float* y) {
    memcpy(y, x, L * sizeof(float));
    if (L < 256 || !(fabsf(cfo_hz) > 0.001f)) return 0;      /* :163 (process) and :174 */
    const float fc = 1500.0f, fs = (float)sample_rate;
    float* I_bb = (float*)malloc(L * sizeof(float));
    float* Q_bb = (float*)malloc(L * sizeof(float));
    float* I_f = (float*)malloc(L * sizeof(float));
    float* Q_f = (float*)malloc(L * sizeof(float));
    for (size_t i = 0; i < L; ++i) {
        float t = (float)i / fs;
        float mix_phase = (float)(2.0f * M_PI * fc * t);
        I_bb[i] = x[i] * cosf(mix_phase);
        Q_bb[i] = x[i] * sinf(mix_phase);
    }
    const size_t win = 48;
    float I_sum = 0, Q_sum = 0;
    for (size_t i = 0; i < L; ++i) {
        I_sum += I_bb[i];
        Q_sum += Q_bb[i];
        if (i >= win) { I_sum -= I_bb[i - win]; Q_sum -= Q_bb[i - win]; }
        size_t n = i + 1 < win ? i + 1 : win;
        I_f[i] = I_sum / (float)n;
        Q_f[i] = Q_sum / (float)n;
    }
    float phase = 0.0f, phase_inc = (float)(2.0f * M_PI * cfo_hz / (double)sample_rate);
    for (size_t i = 0; i < L; ++i) {
        float t = (float)i / fs;
        float mix_phase = (float)(2.0f * M_PI * fc * t);
        float cc = cosf(phase), cs = sinf(phase);
        float I_c = I_f[i] * cc - Q_f[i] * cs, Q_c = I_f[i] * cs + Q_f[i] * cc;
        y[i] = 2.0f * (I_c * cosf(mix_phase) - Q_c * sinf(mix_phase));
        phase += phase_inc;
        if (phase > 2.0f * M_PI) phase = (float)(phase - 2.0f * M_PI);
    }
    free(I_bb); free(Q_bb); free(I_f); free(Q_f);
    return 0;
}

/* noise sigma conventions: 0 = WattersonChannel::process (hf_channel.hpp:110-119), 1 = AWGN tools (test_mode_snr.cpp:58-61) */
float orc_channel_noise_std(const float* tx, size_t L, float snr_db, int convention) {
    float acc = 0.0f;
    for (size_t i = 0; i < L; ++i) acc += tx[i] * tx[i];
    if (convention == 0) return sqrtf(acc / (float)L) * powf(10.0f, -snr_db / 20.0f);
    return sqrtf((acc / (float)L) / powf(10.0f, snr_db / 10.0f));
}
