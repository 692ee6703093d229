/* oracle/pu_oracle_fec.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see pu_oracle.h).
 * CPU restatement of the reference FEC layer: H construction, systematic encoder, flooding scaled
 * min-sum decoder, block and channel interleavers.  Reference: src/fec/ldpc_decoder.cpp,
 * src/fec/ldpc_encoder.cpp (file:line cited per function). */
#include "pu_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ std::mt19937
 * ISO C++ [rand.eng.mers] parameters (w=32,n=624,m=397,r=31,a=0x9908b0df,u=11,d=0xffffffff,
 * s=7,b=0x9d2c5680,t=15,c=0xefc60000,l=18,f=1812433253); the reference seeds it at
 * ldpc_decoder.cpp:72 and demodulator.cpp:82.  Pinned by tests/test_rng.cpp:38-39. */
void orc_mt_seed(orc_mt19937* g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

uint32_t orc_mt_next(orc_mt19937* g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            uint32_t v = g->mt[(i + 397) % 624] ^ (y >> 1);
            if (y & 1u) v ^= 0x9908b0dfu;
            g->mt[i] = v;
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* ------------------------------------------------------------------ code parameters
 * getCodeParams, ldpc_decoder.cpp:21-36: R1_3 / R7_8 / anything else fall back to R1/2 sizes. */
int orc_ldpc_params(int rate, int* k, int* m) {
    switch (rate) {
        case 0: *k = 162; *m = 486; break;  /* R1_4 */
        case 2: *k = 324; *m = 324; break;  /* R1_2 */
        case 3: *k = 432; *m = 216; break;  /* R2_3 */
        case 4: *k = 486; *m = 162; break;  /* R3_4 */
        case 5: *k = 540; *m = 108; break;  /* R5_6 */
        default: *k = 324; *m = 324; break;
    }
    return 0;
}

/* ------------------------------------------------------------------ H = [H_data | I]
 * buildMatrix, ldpc_decoder.cpp:64-137 (mirrored in ldpc_encoder.cpp:70-129):
 *   rng = mt19937(0x12345678 + rate); per info bit j: list the checks whose degree is still < 6,
 *   Fisher-Yates shuffle it with rng() % i (i = size..2, swap [i-1] <-> [rng()%i]), connect j to the first
 *   target_var_degree of them; afterwards every empty check row draws one info bit rng() % k; finally the
 *   identity column k+i is appended to row i. */
int orc_ldpc_build(int rate, orc_ldpc_code* code) {
    int k, m;
    orc_ldpc_params(rate, &k, &m);
    memset(code, 0, sizeof(*code));
    code->rate = rate;
    code->k = k;
    code->m = m;

    orc_mt19937 rng;
    orc_mt_seed(&rng, (uint32_t)(0x12345678 + rate));

    int target_check_degree = 4;
    int target_var_degree = (target_check_degree * m) / k;
    if (target_var_degree < 3) target_var_degree = 3;
    if (target_var_degree > m / 2) target_var_degree = m / 2;
    int max_check_degree = target_check_degree + 2;

    int check_deg[ORC_LDPC_MAX_M];
    int avail[ORC_LDPC_MAX_M];
    memset(check_deg, 0, sizeof(check_deg));

    for (int j = 0; j < k; ++j) {
        int na = 0;
        for (int i = 0; i < m; ++i)
            if (check_deg[i] < max_check_degree) avail[na++] = i;
        for (int i = na; i > 1; --i) {
            int r = (int)(orc_mt_next(&rng) % (uint32_t)i);
            int t = avail[i - 1];
            avail[i - 1] = avail[r];
            avail[r] = t;
        }
        int conn = target_var_degree < na ? target_var_degree : na;
        for (int d = 0; d < conn; ++d) {
            int chk = avail[d];
            code->row[chk][code->row_deg[chk]++] = j;
            check_deg[chk]++;
        }
    }
    for (int i = 0; i < m; ++i) {
        if (code->row_deg[i] == 0) {
            int j = (int)(orc_mt_next(&rng) % (uint32_t)k);
            code->row[i][code->row_deg[i]++] = j;
        }
    }
    code->n_edges = 0;
    for (int i = 0; i < m; ++i) {
        code->row[i][code->row_deg[i]++] = k + i;
        code->n_edges += code->row_deg[i];
    }
    return 0;
}

static const orc_ldpc_code* cached_code(int rate) {
    static orc_ldpc_code cache[8];
    static int built[8];
    int slot = (rate >= 0 && rate < 8) ? rate : 7;
    if (!built[slot]) {
        orc_ldpc_build(rate, &cache[slot]);
        built[slot] = 1;
    }
    return &cache[slot];
}

/* ------------------------------------------------------------------ encoder
 * LDPCEncoder::encode, ldpc_encoder.cpp:193-257: whole input to bits MSB-first, k bits per block
 * (zero padded at the end), parity_i = XOR of the info bits on row i, codeword [info | parity] packed
 * MSB-first into 81 bytes per block. */
long orc_ldpc_encode(int rate, const uint8_t* data, size_t len, uint8_t* out, size_t cap) {
    const orc_ldpc_code* c = cached_code(rate);
    int k = c->k, m = c->m, n = k + m;
    size_t total_bits = len * 8, bit_off = 0, o = 0;
    uint8_t info[ORC_LDPC_N], cw[ORC_LDPC_N];
    while (bit_off < total_bits) {
        for (int j = 0; j < k; ++j) {
            size_t b = bit_off + (size_t)j;
            info[j] = b < total_bits ? (uint8_t)((data[b >> 3] >> (7 - (b & 7))) & 1) : 0;
        }
        memcpy(cw, info, (size_t)k);
        for (int i = 0; i < m; ++i) {
            uint8_t s = 0;
            for (int e = 0; e < c->row_deg[i] - 1; ++e) s ^= info[c->row[i][e]];
            cw[k + i] = s;
        }
        uint8_t byte = 0;
        int cnt = 0;
        for (int j = 0; j < n; ++j) {
            byte = (uint8_t)((byte << 1) | cw[j]);
            if (++cnt == 8) {
                if (o >= cap) return -1;
                out[o++] = byte;
                byte = 0;
                cnt = 0;
            }
        }
        if (cnt > 0) {
            if (o >= cap) return -1;
            out[o++] = (uint8_t)(byte << (8 - cnt));
        }
        bit_off += (size_t)k;
    }
    return (long)o;
}

/* ------------------------------------------------------------------ decoder core
 * decodeBP, ldpc_decoder.cpp:153-259 (the multi-block loop at :307-393 is the same arithmetic):
 *   v2c <- channel LLR (missing LLRs are 0 = erasure, :160-166); c2v <- 0
 *   for it in 0..max_iter-1:
 *     c2v[i][e] = (prod over e2!=e of sign(v2c), "msg < 0" flips)  * (min over e2!=e of |v2c|) * 0.75   (:181-202)
 *     total[j]  = llr_in[j] + sum of c2v in ascending check order                                        (:206-213)
 *     v2c[i][e] = clamp(total[j] - c2v[i][e], -50, +50)                                                  (:216-224)
 *     hard = total < 0 ; stop with last_iters = it when every row XORs to zero                           (:227-235)
 *   last_iters = max_iter on failure.  Info bits packed MSB-first, last byte left-justified (:242-256). */
static int decode_core(const orc_ldpc_code* c, int max_iter, const float* llr, size_t n_llr,
                       float* total, int* iters_out) {
    int k = c->k, m = c->m, n = k + m;
    static _Thread_local float llr_in[ORC_LDPC_N];
    static _Thread_local float v2c[ORC_LDPC_MAX_M][ORC_LDPC_MAX_ROW];
    static _Thread_local float c2v[ORC_LDPC_MAX_M][ORC_LDPC_MAX_ROW];
    for (int j = 0; j < n; ++j) {
        llr_in[j] = (size_t)j < n_llr ? llr[j] : 0.0f;
        total[j] = llr_in[j];
    }
    for (int i = 0; i < m; ++i)
        for (int e = 0; e < c->row_deg[i]; ++e) {
            v2c[i][e] = llr_in[c->row[i][e]];
            c2v[i][e] = 0.0f;
        }
    int ok = 0, it;
    for (it = 0; it < max_iter; ++it) {
        for (int i = 0; i < m; ++i) {
            int deg = c->row_deg[i];
            for (int e = 0; e < deg; ++e) {
                float sign = 1.0f, min_abs = FLT_MAX;
                for (int e2 = 0; e2 < deg; ++e2) {
                    if (e2 == e) continue;
                    float msg = v2c[i][e2];
                    if (msg < 0) sign = -sign;
                    float a = fabsf(msg);
                    if (a < min_abs) min_abs = a;
                }
                c2v[i][e] = sign * min_abs * 0.75f;
            }
        }
        for (int j = 0; j < n; ++j) total[j] = llr_in[j];
        for (int i = 0; i < m; ++i)
            for (int e = 0; e < c->row_deg[i]; ++e) total[c->row[i][e]] += c2v[i][e];
        for (int i = 0; i < m; ++i)
            for (int e = 0; e < c->row_deg[i]; ++e) {
                float v = total[c->row[i][e]] - c2v[i][e];
                v = fmaxf(-50.0f, fminf(50.0f, v));  /* std::max(-50, std::min(50, v)) */
                v2c[i][e] = v;
            }
        int all_zero = 1;
        for (int i = 0; i < m && all_zero; ++i) {
            int s = 0;
            for (int e = 0; e < c->row_deg[i]; ++e) s ^= (total[c->row[i][e]] < 0) ? 1 : 0;
            if (s) all_zero = 0;
        }
        if (all_zero) {
            ok = 1;
            break;
        }
    }
    *iters_out = it;
    return ok;
}

static size_t pack_bits(const uint8_t* bits, size_t nbits, uint8_t* out) {
    size_t o = 0;
    uint8_t byte = 0;
    int cnt = 0;
    for (size_t j = 0; j < nbits; ++j) {
        byte = (uint8_t)((byte << 1) | bits[j]);
        if (++cnt == 8) {
            out[o++] = byte;
            byte = 0;
            cnt = 0;
        }
    }
    if (cnt > 0) out[o++] = (uint8_t)(byte << (8 - cnt));
    return o;
}

int orc_ldpc_decode_block(const orc_ldpc_code* code, int max_iter, const float* llr, size_t n_llr,
                          uint8_t* info_bytes, int* ok, int* iters, float* llr_total_out) {
    float total[ORC_LDPC_N];
    uint8_t bits[ORC_LDPC_N];
    int it = 0;
    int s = decode_core(code, max_iter, llr, n_llr, total, &it);
    for (int j = 0; j < code->k; ++j) bits[j] = total[j] < 0 ? 1 : 0;
    pack_bits(bits, (size_t)code->k, info_bytes);
    if (ok) *ok = s;
    if (iters) *iters = it;
    if (llr_total_out) memcpy(llr_total_out, total, sizeof(float) * (size_t)(code->k + code->m));
    return 0;
}

/* LDPCDecoder::decodeSoft, ldpc_decoder.cpp:283-428: empty input -> {} and failure (:285-288);
 * <= n LLRs -> one block; otherwise full blocks concatenated at BIT level (:386-390), a trailing partial
 * block is zero-padded (:396-407); last_success = AND over full blocks, then overwritten by the trailing
 * partial block's decodeBP (:400 sets last_success inside decodeBP); last_iters = last block decoded. */
long orc_ldpc_decode_soft(int rate, int max_iter, const float* llr, size_t n_in, uint8_t* out, size_t cap,
                          int* ok, int* iters) {
    const orc_ldpc_code* c = cached_code(rate);
    int k = c->k, n = c->k + c->m;
    if (max_iter < 0) max_iter = 50;
    if (n_in == 0) {
        if (ok) *ok = 0;
        if (iters) *iters = 0;
        return 0;
    }
    float total[ORC_LDPC_N];
    int it = 0;
    if (n_in <= (size_t)n) {
        uint8_t bits[ORC_LDPC_N];
        int s = decode_core(c, max_iter, llr, n_in, total, &it);
        for (int j = 0; j < k; ++j) bits[j] = total[j] < 0 ? 1 : 0;
        if ((size_t)((k + 7) / 8) > cap) return -1;
        if (ok) *ok = s;
        if (iters) *iters = it;
        return (long)pack_bits(bits, (size_t)k, out);
    }
    size_t nblk = (n_in + (size_t)n - 1) / (size_t)n;
    uint8_t* all = (uint8_t*)malloc(nblk * (size_t)k);
    size_t nb = 0, off = 0;
    int success = 1;
    while (off + (size_t)n <= n_in) {
        int s = decode_core(c, max_iter, llr + off, (size_t)n, total, &it);
        if (!s) success = 0;
        for (int j = 0; j < k; ++j) all[nb++] = total[j] < 0 ? 1 : 0;
        off += (size_t)n;
    }
    if (off < n_in) {
        int s = decode_core(c, max_iter, llr + off, n_in - off, total, &it);
        success = s; /* decodeBP assigns last_success (:178,233) */
        for (int j = 0; j < k; ++j) all[nb++] = total[j] < 0 ? 1 : 0;
    }
    if ((nb + 7) / 8 > cap) {
        free(all);
        return -1;
    }
    long r = (long)pack_bits(all, nb, out);
    free(all);
    if (ok) *ok = success;
    if (iters) *iters = it;
    return r;
}

int orc_ldpc_decode_batch(int rate, int max_iter, const float* llr, size_t B, uint8_t* out, size_t out_stride,
                          uint8_t* ok, int32_t* iters) {
    const orc_ldpc_code* c = cached_code(rate);
    if (max_iter < 0) max_iter = 50;
    if ((size_t)((c->k + 7) / 8) > out_stride) return -1;
    for (size_t b = 0; b < B; ++b) {
        int s, it;
        orc_ldpc_decode_block(c, max_iter, llr + b * ORC_LDPC_N, ORC_LDPC_N, out + b * out_stride, &s, &it, NULL);
        ok[b] = (uint8_t)s;
        iters[b] = it;
    }
    return 0;
}

double orc_time_ldpc_decode(int rate, int max_iter, const float* llr, size_t B, uint8_t* out, size_t out_stride,
                            uint8_t* ok, int32_t* iters) {
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    orc_ldpc_decode_batch(rate, max_iter, llr, B, out, out_stride, ok, iters);
    clock_gettime(CLOCK_MONOTONIC, &b);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}

/* ------------------------------------------------------------------ ChannelInterleaver
 * findCoprimeStep, ldpc_decoder.cpp:547-572; permutation dest = (i*step) % total (:595-599);
 * interleave out[perm[i]] = in[i] (:602-610); deinterleave out[inv[i]] = in[i] (:612-620). */
static size_t gcd_sz(size_t a, size_t b) {
    while (b) {
        size_t t = b;
        b = a % b;
        a = t;
    }
    return a;
}

size_t orc_channel_interleaver_step(size_t n, size_t total) {
    size_t target = n * 3;
    if (target >= total) target = total / 2;
    for (size_t s = target; s < total; ++s)
        if (gcd_sz(s, total) == 1) return s;
    for (size_t s = n + 1; s < total; ++s)
        if (gcd_sz(s, total) == 1) return s;
    return n + 1;
}

int orc_channel_interleave(size_t bps, size_t total, const float* in, size_t n, float* out, int inverse) {
    size_t step = orc_channel_interleaver_step(bps, total);
    size_t* perm = (size_t*)malloc(sizeof(size_t) * total * 2);
    size_t* inv = perm + total;
    for (size_t i = 0; i < total; ++i) {
        size_t d = (i * step) % total;
        perm[i] = d;
        inv[d] = i;
    }
    for (size_t i = 0; i < total; ++i) out[i] = 0.0f;
    size_t lim = n < total ? n : total;
    for (size_t i = 0; i < lim; ++i) out[inverse ? inv[i] : perm[i]] = in[i];
    free(perm);
    return (int)total;
}

/* Interleaver (rows x cols transpose), ldpc_decoder.cpp:454-464, soft versions :524-540 */
int orc_block_interleave(size_t rows, size_t cols, const float* in, size_t n, float* out, int inverse) {
    size_t tot = rows * cols;
    for (size_t i = 0; i < n; ++i) out[i] = 0.0f;
    for (size_t i = 0; i < n && i < tot; ++i) {
        size_t p = (i % cols) * rows + (i / cols);
        if (inverse) out[i] = in[p];
        else out[p] = in[i];
    }
    return (int)n;
}

/* ------------------------------------------------------------------ protocol-v2 codeword framing
 * src/protocol/frame_v2.hpp:139-155,212-217,551-566 (constants, isControlFrame, bytes per codeword), frame_v2.cpp:111-124 (CRC-16/CCITT,
 * init 0xFFFF, poly 0x1021), :952-982 (reassembleCodewords), :1023-1044 (CodewordStatus::reassemble), :1079-1127 (encodeFrameWithLDPC),
 * :1134-1156 (decodeSingleCodeword), :1175-1230 (parseHeader), and RxPipeline::decodeFrame (src/gui/modem/rx_pipeline.cpp:348-445). */
uint16_t orc_frame_crc16(const uint8_t* data, size_t len) {
    uint16_t crc = 0xFFFF;
    for (size_t i = 0; i < len; i++) {
        crc ^= (uint16_t)((uint16_t)data[i] << 8);
        for (int j = 0; j < 8; j++) crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x1021) : (uint16_t)(crc << 1);
    }
    return crc;
}

size_t orc_frame_bytes_per_codeword(int rate) {   /* getInfoBitsForRate / 8, rounded down */
    switch (rate) {
        case 0: return 162 / 8;
        case 1: return 216 / 8;
        case 2: return 324 / 8;
        case 3: return 432 / 8;
        case 4: return 486 / 8;
        case 5: return 540 / 8;
        default: return 162 / 8;
    }
}

/* encodeFrameWithLDPC: CW0 = the first bytes_per_cw frame bytes, CW1+ = {0xD5, index, payload}, zero-padded; returns the codeword count */
long orc_frame_encode(int rate, const uint8_t* frame, size_t len, uint8_t* out, size_t cap) {
    const size_t bpc = orc_frame_bytes_per_codeword(rate), pay = bpc - 2;
    uint8_t chunk[80];
    size_t off = 0, ncw = 0, offset = bpc;
    memset(chunk, 0, sizeof chunk);
    memcpy(chunk, frame, len < bpc ? len : bpc);
    for (;;) {
        if (off + 81 > cap) return -1;
        if (orc_ldpc_encode(rate, chunk, bpc, out + off, 81) != 81) return -1;
        off += 81;
        ncw++;
        if (offset >= len) break;
        memset(chunk, 0, sizeof chunk);
        chunk[0] = 0xD5;
        chunk[1] = (uint8_t)ncw;
        const size_t remaining = len - offset;
        memcpy(chunk + 2, frame + offset, remaining < pay ? remaining : pay);
        offset += pay;
    }
    return (long)ncw;
}

/* parseHeader: returns 1 when valid; total_cw / payload_len / is_control / type as HeaderInfo */
static int frame_parse_header(const uint8_t* d, int* type, int* is_control, int* total_cw, int* payload_len) {
    if (((d[0] << 8) | d[1]) != 0x554C) return 0;
    *type = d[2];
    *is_control = d[2] == 0x10 || d[2] == 0x11 || d[2] == 0x16 || d[2] == 0x17 || d[2] == 0x20 || d[2] == 0x21 || d[2] == 0x40;
    if (*is_control) {
        if (((d[18] << 8) | d[19]) != orc_frame_crc16(d, 18)) return 0;
        *total_cw = 1;
        *payload_len = 0;
    } else {
        *total_cw = d[12];
        *payload_len = (d[13] << 8) | d[14];
        if (((d[15] << 8) | d[16]) != orc_frame_crc16(d, 15)) return 0;
    }
    return 1;
}

/* RxPipeline::decodeFrame: info[5] = {success, frame_type, codewords_ok, codewords_failed, expected}; returns the frame_data size */
long orc_frame_decode(int rate, const float* soft, size_t n_soft, int num_codewords, uint8_t* out, size_t cap, int32_t* info) {
    const size_t bpc = orc_frame_bytes_per_codeword(rate);
    const orc_ldpc_code* c = cached_code(rate);
    uint8_t cw[256][72];
    uint8_t buf[96];
    int ok, it;
    info[0] = info[1] = info[2] = info[3] = info[4] = 0;
    if (n_soft < ORC_LDPC_N) return 0;
    orc_ldpc_decode_block(c, 50, soft, ORC_LDPC_N, buf, &ok, &it, NULL);
    if (!ok || (size_t)((c->k + 7) / 8) < bpc) { info[3]++; return 0; }     /* decodeSingleCodeword */
    memcpy(cw[0], buf, bpc);
    info[2]++;
    int type = 0, is_control = 0, expected = 0, payload_len = 0;
    if (!frame_parse_header(cw[0], &type, &is_control, &expected, &payload_len)) return 0;
    info[1] = type;
    info[4] = expected;
    if (num_codewords < expected) return 0;
    int all = 1;
    for (int i = 1; i < expected; i++) {
        orc_ldpc_decode_block(c, 50, soft + (size_t)i * ORC_LDPC_N, ORC_LDPC_N, buf, &ok, &it, NULL);
        if (ok) { memcpy(cw[i], buf, bpc); info[2]++; }
        else { info[3]++; all = 0; }
    }
    if (!all) return 0;      /* allSuccess over `expected` entries: expected == 0 leaves CW0 only... see below */
    info[0] = 1;
    /* CodewordStatus::reassemble: decoded.empty() (expected == 0) -> {}; otherwise header + payload + CRC bytes */
    if (expected == 0) return 0;
    const size_t expected_size = is_control ? 20 : (size_t)(17 + payload_len + 2);
    size_t n = 0;
    for (int i = 0; i < expected && n < expected_size; i++) {
        const size_t remaining = expected_size - n;
        const uint8_t* src = cw[i];
        size_t avail = bpc;
        if (i > 0 && cw[i][0] == 0xD5) { src = cw[i] + 2; avail = bpc - 2; }   /* marker + index skipped; else legacy fallback */
        const size_t take = remaining < avail ? remaining : avail;
        if (n + take > cap) return -1;
        memcpy(out + n, src, take);
        n += take;
    }
    return (long)n;
}
