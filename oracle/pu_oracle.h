/* oracle/pu_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's (secup/ProjectUltra, /root/reference) algorithms for the
 * receive-chain hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this; the product (projectultra_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  Every function here is checked in tests/test_oracle_*.py against
 *   (1) the reference's own structural KATs (tests/test_rng.cpp:38-39 Fisher-Yates order,
 *       tests/test_multiblock_ldpc.cpp encode->decode identities, SURVEY App. C fingerprints), and
 *   (2) outputs of the unmodified reference compiled here (oracle/_ref/libpu_ref.so) on the same inputs,
 *       plus golden vectors generated from it and committed under tests/golden/.
 * All arithmetic is fp32 with the reference's operation order; compile with -ffp-contract=off.
 */
#ifndef PU_ORACLE_H
#define PU_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* POD mirror of ultra::ModemConfig (include/ultra/types.hpp:139-234) */
typedef struct {
    uint32_t sample_rate, center_freq, fft_size, num_carriers, cp_mode, symbol_guard, pilot_spacing, use_pilots,
        modulation, code_rate;
    float output_scale, tx_cfo_hz;
} orc_modem_config;

/* std::mt19937 (used by ldpc_decoder.cpp:72, demodulator.cpp:82, tests/test_rng.cpp) */
typedef struct { uint32_t mt[624]; int idx; } orc_mt19937;
void orc_mt_seed(orc_mt19937* g, uint32_t seed);
uint32_t orc_mt_next(orc_mt19937* g);

/* ---- LDPC (src/fec/ldpc_decoder.cpp, src/fec/ldpc_encoder.cpp) ---- */
#define ORC_LDPC_N 648
#define ORC_LDPC_MAX_M 486
#define ORC_LDPC_MAX_ROW 16
typedef struct {
    int rate, k, m;
    int row_deg[ORC_LDPC_MAX_M];                   /* including the identity column */
    int row[ORC_LDPC_MAX_M][ORC_LDPC_MAX_ROW];     /* variable indices in stored order */
    int n_edges;
} orc_ldpc_code;
int orc_ldpc_params(int rate, int* k, int* m);                   /* getCodeParams, ldpc_decoder.cpp:21-36 */
int orc_ldpc_build(int rate, orc_ldpc_code* code);               /* buildMatrix, ldpc_decoder.cpp:64-137 */
long orc_ldpc_encode(int rate, const uint8_t* data, size_t len, uint8_t* out, size_t cap); /* ldpc_encoder.cpp:193-257 */
/* decodeBP on one <=648-LLR block (ldpc_decoder.cpp:153-259); hard_out[648] optional */
int orc_ldpc_decode_block(const orc_ldpc_code* code, int max_iter, const float* llr, size_t n_llr,
                          uint8_t* info_bytes, int* ok, int* iters, float* llr_total_out);
/* decodeSoft incl. multi-block bit concatenation (ldpc_decoder.cpp:283-428) */
long orc_ldpc_decode_soft(int rate, int max_iter, const float* llr, size_t n, uint8_t* out, size_t cap, int* ok, int* iters);
int orc_ldpc_decode_batch(int rate, int max_iter, const float* llr, size_t B, uint8_t* out, size_t out_stride,
                          uint8_t* ok, int32_t* iters);

/* ---- protocol-v2 codeword framing (src/protocol/frame_v2.cpp; RxPipeline::decodeFrame, src/gui/modem/rx_pipeline.cpp:348-445) ---- */
uint16_t orc_frame_crc16(const uint8_t* data, size_t len);
size_t orc_frame_bytes_per_codeword(int rate);
long orc_frame_encode(int rate, const uint8_t* frame, size_t len, uint8_t* out, size_t cap);
long orc_frame_decode(int rate, const float* soft, size_t n_soft, int num_codewords, uint8_t* out, size_t cap, int32_t* info);

/* ---- interleavers (ldpc_decoder.cpp:454-620) ---- */
size_t orc_channel_interleaver_step(size_t bits_per_symbol, size_t total);
int orc_channel_interleave(size_t bps, size_t total, const float* in, size_t n, float* out, int inverse);
int orc_block_interleave(size_t rows, size_t cols, const float* in, size_t n, float* out, int inverse);

/* ---- DSP primitives ---- */
int orc_fft(size_t n, const float* in_ri, float* out_ri, int inverse);  /* fft.cpp:76-121 */
int orc_nco(float freq, float fs, size_t n, float* out_ri);             /* filters.cpp:228-238 */
int orc_soft_demap(int mod, float re, float im, float pre, float pim, float nv, float* out); /* soft_demap.hpp */

/* ---- OFDM TX (src/ofdm/modulator.cpp) layout 0: training(2)+data, layout 1: S-C preamble+data ---- */
long orc_ofdm_tx(const orc_modem_config* c, int layout, const uint8_t* data, size_t len, float* out, size_t cap);

/* ---- OFDM RX presynced path (demodulator.cpp:854-985 and callees) ---- */
#define ORC_STAGE_SCALARS 10
typedef struct {   /* optional per-stage dumps, same layout as ref_ofdm_presynced_stages */
    int32_t* carriers; float* lts_bins; float* h_lts; float* bins; float* h; float* eq; float* nv; float* scalars;
    int max_sym;
} orc_stage_dump;
float orc_ofdm_training_cfo(const orc_modem_config* c, const float* samples, size_t L, int num_symbols);
long orc_ofdm_presynced(const orc_modem_config* c, const float* samples, size_t L, int training,
                        int cfo_mode, float cfo_hz, float cfo_phase, float* llr_out, size_t cap,
                        float* snr_db, float* final_cfo, orc_stage_dump* dump);
/* OFDMDemodulator::process fed in chunk-sample pieces + the soft bits of every complete data symbol (Schmidl-Cox path,
 * demodulator.cpp:459-760, ofdm_sync.cpp).  info[4] = {synchronised, sync offset, samples consumed, calls}. */
long orc_ofdm_process(const orc_modem_config* c, const float* samples, size_t L, size_t chunk, float sync_threshold,
                      float* llr_out, size_t cap, int32_t* info, float* coarse_cfo);
/* sync::ChirpSync (src/sync/chirp_sync.hpp) as OFDMChirpWaveform configures it, and the waveform's receive glue */
long orc_chirp_generate(float fs, float tx_cfo, float* out, size_t cap);
int orc_chirp_detect_dual(float fs, const float* x, size_t L, float threshold, int32_t* info, float* f);
long orc_ofdm_chirp_receive(const orc_modem_config* c, const float* x, size_t L, float threshold, int32_t* info, float* cfo_out,
                            float* llr_out, size_t cap);
int orc_ofdm_presynced_batch(const orc_modem_config* c, const float* samples, size_t B, size_t L, int training,
                             int cfo_mode, const float* cfo_hz, const float* cfo_phase,
                             float* llr_out, size_t stride, int32_t* counts);

/* timed loops for bench.py cpu_baseline kind="port" */
double orc_time_presynced_decode(const orc_modem_config* c, const float* samples, size_t B, size_t L, int rate,
                                 uint8_t* info_out, size_t info_stride, uint8_t* ok);
double orc_time_ldpc_decode(int rate, int max_iter, const float* llr, size_t B, uint8_t* out, size_t out_stride,
                            uint8_t* ok, int32_t* iters);

#ifdef __cplusplus
}
#endif
#endif

/* ---- channel twin (oracle/pu_oracle_channel.c); declared late to keep the header append-only ---- */
#ifdef __cplusplus
extern "C" {
#endif
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float orc_noise_normal(uint64_t seed, uint32_t n);
void orc_fading_normals(uint64_t seed, uint32_t n, float z[4]);
int orc_channel_apply(float delay_ms, float doppler_hz, float g1, float g2, uint32_t fs, int fading, int multipath,
                      int noise, const float* x, size_t L, float noise_std, uint64_t seed, float* y);
int orc_channel_apply_cfo(const float* x, size_t L, float cfo_hz, uint32_t sample_rate, float* y);
float orc_channel_noise_std(const float* tx, size_t L, float snr_db, int convention);
#ifdef __cplusplus
}
#endif

/* ---- single- and multi-carrier DPSK demodulators (oracle/pu_oracle_psk.c) ---- */
#ifdef __cplusplus
extern "C" {
#endif
/* DPSKDemodulator::findPreamble (dpsk.hpp:338-481): data_start (> 0) or -1, and the members it leaves behind */
long orc_dpsk_find_preamble(int sps, float fc, float fs, const float* x, size_t L, float* est_cfo, float* phase_off);
long orc_mcdpsk_got_chirp(int nc, int sps, int bits_per_symbol, float f_lo, float f_hi, float fs, int training_symbols, const float* x,
                          size_t L, float chirp_cfo, float* llr, size_t cap, float* cfo_after);
long orc_mcdpsk_chirp_receive(int nc, int sps, int bits_per_symbol, float f_lo, float f_hi, float fs, int training_symbols, const float* x,
                              size_t L, float threshold, int32_t* info, float* f, float* llr, size_t cap, float* cfo_after);
long orc_dpsk_demod_soft(int mod, int sps, float fc, float fs, const float* x, size_t L, long data_start, int ref_mode,
                         float est_cfo, float phase_off, float* llr, size_t cap);
long orc_mcdpsk_demod_soft(int nc, int sps, int bits_per_symbol, float f_lo, float f_hi, float fs, const float* x, size_t L,
                           int training_symbols, float* llr, size_t cap, float* residual_cfo);
#ifdef __cplusplus
}
#endif
