"""ctypes binding for oracle/libpu_oracle.so (plain-C CPU restatement of the reference algorithms).
TEST INFRASTRUCTURE: imported only by tests/, smoke() and bench.py's cpu_baseline leg."""
import ctypes as C
import os
import subprocess
import numpy as np

from refapi import ModemConfig, RATE_K, STAGE_SCALARS  # same POD layout

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libpu_oracle.so")

_lib = None


class LdpcCode(C.Structure):
    _fields_ = [("rate", C.c_int), ("k", C.c_int), ("m", C.c_int), ("row_deg", C.c_int * 486),
                ("row", (C.c_int * 16) * 486), ("n_edges", C.c_int)]


class StageDump(C.Structure):
    _fields_ = [("carriers", C.POINTER(C.c_int32)), ("lts_bins", C.POINTER(C.c_float)),
                ("h_lts", C.POINTER(C.c_float)), ("bins", C.POINTER(C.c_float)), ("h", C.POINTER(C.c_float)),
                ("eq", C.POINTER(C.c_float)), ("nv", C.POINTER(C.c_float)), ("scalars", C.POINTER(C.c_float)),
                ("max_sym", C.c_int)]


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        for f in ("orc_ldpc_encode", "orc_ldpc_decode_soft", "orc_ofdm_tx", "orc_ofdm_presynced", "orc_ofdm_process"):
            getattr(L, f).restype = C.c_long
        L.orc_channel_interleaver_step.restype = C.c_size_t
        L.orc_time_presynced_decode.restype = C.c_double
        L.orc_time_ldpc_decode.restype = C.c_double
        L.orc_mt_next.restype = C.c_uint32
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a):
    return np.ascontiguousarray(np.frombuffer(bytes(a), dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else a,
                                dtype=np.uint8)


def mt19937(seed, n):
    st = (C.c_uint32 * 625)()
    lib().orc_mt_seed(st, C.c_uint32(seed))
    return np.array([lib().orc_mt_next(st) for _ in range(n)], dtype=np.uint64)


def ldpc_build(rate):
    code = LdpcCode()
    lib().orc_ldpc_build(rate, C.byref(code))
    rows = [[code.row[i][e] for e in range(code.row_deg[i])] for i in range(code.m)]
    return code.k, code.m, rows


def ldpc_encode(rate, data):
    d = _u8(data)
    out = np.zeros(81 * (len(d) * 8 // RATE_K[rate] + 2), np.uint8)
    n = lib().orc_ldpc_encode(rate, _p(d, C.c_uint8), C.c_size_t(len(d)), _p(out, C.c_uint8), C.c_size_t(len(out)))
    assert n >= 0
    return out[:n].copy()


def ldpc_decode_soft(rate, llr, max_iter=-1):
    x = _f32(llr)
    out = np.zeros(128 * (len(x) // 648 + 2), np.uint8)
    ok, it = C.c_int(0), C.c_int(0)
    n = lib().orc_ldpc_decode_soft(rate, max_iter, _p(x, C.c_float), C.c_size_t(len(x)), _p(out, C.c_uint8),
                                   C.c_size_t(len(out)), C.byref(ok), C.byref(it))
    assert n >= 0
    return out[:n].copy(), bool(ok.value), it.value


def ldpc_decode_batch(rate, llr, max_iter=-1):
    x = _f32(llr).reshape(-1, 648)
    B = x.shape[0]
    kb = (RATE_K.get(rate, 324) + 7) // 8
    out = np.zeros((B, kb), np.uint8)
    ok = np.zeros(B, np.uint8)
    it = np.zeros(B, np.int32)
    r = lib().orc_ldpc_decode_batch(rate, max_iter, _p(x, C.c_float), C.c_size_t(B), _p(out, C.c_uint8),
                                    C.c_size_t(kb), _p(ok, C.c_uint8), _p(it, C.c_int32))
    assert r == 0
    return out, ok, it


def channel_interleaver_step(bps, total=648):
    return lib().orc_channel_interleaver_step(C.c_size_t(bps), C.c_size_t(total))


def channel_interleave(bps, x, inverse=False, total=648):
    x = _f32(x)
    out = np.zeros(total, np.float32)
    lib().orc_channel_interleave(C.c_size_t(bps), C.c_size_t(total), _p(x, C.c_float), C.c_size_t(len(x)),
                                 _p(out, C.c_float), int(inverse))
    return out


def block_interleave(rows, cols, x, inverse=False):
    x = _f32(x)
    out = np.zeros(len(x), np.float32)
    lib().orc_block_interleave(C.c_size_t(rows), C.c_size_t(cols), _p(x, C.c_float), C.c_size_t(len(x)),
                               _p(out, C.c_float), int(inverse))
    return out


def tools_apply_cfo(x, cfo_hz, sample_rate=48000.0):
    """tools/test_iwaveform.cpp:67-118 (oracle/pu_oracle_ofdm.c: orc_tools_apply_cfo) on a copy of x."""
    out = np.ascontiguousarray(x, np.float32).copy()
    lib().orc_tools_apply_cfo(_p(out, C.c_float), C.c_size_t(len(out)), C.c_float(cfo_hz), C.c_float(sample_rate))
    return out


def fft(x, inverse=False):
    z = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.zeros_like(z)
    lib().orc_fft(C.c_size_t(len(z)), _p(z.view(np.float32), C.c_float), _p(out.view(np.float32), C.c_float),
                  int(inverse))
    return out


def nco(freq, fs, n):
    out = np.zeros(n, np.complex64)
    lib().orc_nco(C.c_float(freq), C.c_float(fs), C.c_size_t(n), _p(out.view(np.float32), C.c_float))
    return out


def soft_demap(mod, sym, prev=1 + 0j, nv=0.1):
    out = np.zeros(8, np.float32)
    n = lib().orc_soft_demap(mod, C.c_float(sym.real), C.c_float(sym.imag), C.c_float(prev.real),
                             C.c_float(prev.imag), C.c_float(nv), _p(out, C.c_float))
    return out[:n].copy()


def ofdm_tx(cfg, data, layout=0):
    d = _u8(data)
    cap = 400000
    out = np.zeros(cap, np.float32)
    n = lib().orc_ofdm_tx(C.byref(cfg), layout, _p(d, C.c_uint8), C.c_size_t(len(d)), _p(out, C.c_float),
                          C.c_size_t(cap))
    assert n >= 0, n
    return out[:n].copy()


def ofdm_presynced(cfg, samples, training=2, cfo_mode=1, cfo_hz=0.0, cfo_phase=0.0):
    x = _f32(samples)
    cap = 16384
    out = np.zeros(cap, np.float32)
    snr, fc = C.c_float(0), C.c_float(0)
    n = lib().orc_ofdm_presynced(C.byref(cfg), _p(x, C.c_float), C.c_size_t(len(x)), training, cfo_mode,
                                 C.c_float(cfo_hz), C.c_float(cfo_phase), _p(out, C.c_float), C.c_size_t(cap),
                                 C.byref(snr), C.byref(fc), None)
    assert n >= 0, n
    return out[:n].copy(), snr.value, fc.value


def ofdm_training_cfo(cfg, samples, num_symbols=2):
    """estimateCFOFromTraining(samples, num_symbols, 0) (src/ofdm/ofdm_sync.cpp:278-380)."""
    x = _f32(samples)
    f = lib().orc_ofdm_training_cfo
    f.restype = C.c_float
    return float(f(C.byref(cfg), _p(x, C.c_float), C.c_size_t(len(x)), int(num_symbols)))


def ofdm_process(cfg, samples, chunk=960, sync_threshold=0.0):
    """OFDMDemodulator::process in chunk-sample pieces + getSoftBits() (first <= 648 soft bits):
    (llr, synced, sync_offset, coarse_cfo, data_start, calls)."""
    x = _f32(samples)
    cap = 16384
    out = np.zeros(cap, np.float32)
    info = np.zeros(4, np.int32)
    cfo = C.c_float(0)
    n = lib().orc_ofdm_process(C.byref(cfg), _p(x, C.c_float), C.c_size_t(len(x)), C.c_size_t(chunk), C.c_float(sync_threshold),
                               _p(out, C.c_float), C.c_size_t(cap), _p(info, C.c_int32), C.byref(cfo))
    assert n >= 0, n
    return out[:min(n, 648)].copy(), bool(info[0]), int(info[1]), float(cfo.value), int(info[2]), int(info[3])


def chirp_generate(fs=48000.0, tx_cfo=0.0):
    out = np.zeros(80000, np.float32)
    L = lib()
    L.orc_chirp_generate.restype = C.c_long
    n = L.orc_chirp_generate(C.c_float(fs), C.c_float(tx_cfo), _p(out, C.c_float), C.c_size_t(len(out)))
    assert n >= 0
    return out[:n].copy()


def ofdm_chirp_receive(cfg, samples, threshold=0.15):
    """(llr, info[4] = {success, up start, down start, training start}, cfo) -- see orc_ofdm_chirp_receive."""
    x = _f32(samples)
    out = np.zeros(8192, np.float32)
    info = np.zeros(4, np.int32)
    cfo = C.c_float(0)
    L = lib()
    L.orc_ofdm_chirp_receive.restype = C.c_long
    n = L.orc_ofdm_chirp_receive(C.byref(cfg), _p(x, C.c_float), C.c_size_t(len(x)), C.c_float(threshold), _p(info, C.c_int32),
                                 C.byref(cfo), _p(out, C.c_float), C.c_size_t(len(out)))
    assert 0 <= n <= len(out), n
    return out[:n].copy(), info, float(cfo.value)


def ofdm_presynced_stages(cfg, samples, training=2, cfo_mode=1, cfo_hz=0.0, cfo_phase=0.0, max_sym=64):
    x = _f32(samples)
    nd, npil = cfg.n_data, cfg.n_pilots
    nu = nd + npil
    carriers = np.zeros(nu, np.int32)
    lts = np.zeros((training, nu), np.complex64)
    h_lts = np.zeros(nu, np.complex64)
    bins = np.zeros((max_sym, nu), np.complex64)
    h = np.zeros((max_sym, nu), np.complex64)
    eq = np.zeros((max_sym, nd), np.complex64)
    nv = np.zeros((max_sym, nd), np.float32)
    sc = np.zeros((max_sym, STAGE_SCALARS), np.float32)
    fp = lambda a: _p(a.view(np.float32), C.c_float)
    d = StageDump(_p(carriers, C.c_int32), fp(lts), fp(h_lts), fp(bins), fp(h), fp(eq), _p(nv, C.c_float),
                  _p(sc, C.c_float), max_sym)
    cap = 16384
    llr = np.zeros(cap, np.float32)
    snr, fc = C.c_float(0), C.c_float(0)
    n = lib().orc_ofdm_presynced(C.byref(cfg), _p(x, C.c_float), C.c_size_t(len(x)), training, cfo_mode,
                                 C.c_float(cfo_hz), C.c_float(cfo_phase), _p(llr, C.c_float), C.c_size_t(cap),
                                 C.byref(snr), C.byref(fc), C.byref(d))
    assert n >= 0, n
    ns = (len(x) - training * cfg.symbol_samples) // cfg.symbol_samples
    ns = min(ns, max_sym)
    return dict(n_sym=ns, carriers=carriers, lts_bins=lts, h_lts=h_lts, bins=bins[:ns], h=h[:ns], eq=eq[:ns],
                nv=nv[:ns], scalars=sc[:ns], llr=llr[:n].copy(), snr_db=snr.value, final_cfo=fc.value)


def ofdm_presynced_batch(cfg, samples, n_llr, training=2, cfo_mode=1, cfo_hz=None, cfo_phase=None):
    x = _f32(samples)
    B, L = x.shape
    out = np.zeros((B, n_llr), np.float32)
    counts = np.zeros(B, np.int32)
    f = _f32(cfo_hz) if cfo_hz is not None else None
    p = _f32(cfo_phase) if cfo_phase is not None else None
    r = lib().orc_ofdm_presynced_batch(C.byref(cfg), _p(x, C.c_float), C.c_size_t(B), C.c_size_t(L), training,
                                       cfo_mode, _p(f, C.c_float) if f is not None else None,
                                       _p(p, C.c_float) if p is not None else None,
                                       _p(out, C.c_float), C.c_size_t(n_llr), _p(counts, C.c_int32))
    assert r == 0, r
    return out, counts


def time_presynced_decode(cfg, samples, rate):
    x = _f32(samples)
    B, L = x.shape
    kb = (RATE_K.get(rate, 324) + 7) // 8
    info = np.zeros((B, kb), np.uint8)
    ok = np.zeros(B, np.uint8)
    t = lib().orc_time_presynced_decode(C.byref(cfg), _p(x, C.c_float), C.c_size_t(B), C.c_size_t(L), rate,
                                        _p(info, C.c_uint8), C.c_size_t(kb), _p(ok, C.c_uint8))
    return t, info, ok


def time_ldpc_decode(rate, llr, max_iter=-1):
    x = _f32(llr).reshape(-1, 648)
    B = x.shape[0]
    kb = (RATE_K.get(rate, 324) + 7) // 8
    out = np.zeros((B, kb), np.uint8)
    ok = np.zeros(B, np.uint8)
    it = np.zeros(B, np.int32)
    t = lib().orc_time_ldpc_decode(rate, max_iter, _p(x, C.c_float), C.c_size_t(B), _p(out, C.c_uint8),
                                   C.c_size_t(kb), _p(ok, C.c_uint8), _p(it, C.c_int32))
    return t, out, ok, it


# ---------------------------------------------------------------- DPSK (oracle/pu_oracle_psk.c)
def dpsk_demod_soft(mod, sps, samples, data_start, ref_mode=0, est_cfo=0.0, phase_off=0.0, fc=1500.0, fs=48000.0):
    x = _f32(samples)
    out = np.zeros(4096, np.float32)
    L = lib()
    L.orc_dpsk_demod_soft.restype = C.c_long
    n = L.orc_dpsk_demod_soft(mod, sps, C.c_float(fc), C.c_float(fs), _p(x, C.c_float), C.c_size_t(len(x)), C.c_long(data_start),
                              ref_mode, C.c_float(est_cfo), C.c_float(phase_off), _p(out, C.c_float), C.c_size_t(len(out)))
    assert 0 <= n <= len(out), n
    return out[:n].copy()


def mcdpsk_got_chirp(nc, samples, chirp_cfo, sps=512, bits=2, f_lo=500.0, f_hi=2500.0, fs=48000.0, training=8):
    """processGotChirp behind an external chirp: (llr, cfo_after)."""
    x = _f32(samples)
    out = np.zeros(8192, np.float32)
    cfo = C.c_float(0)
    L = lib()
    L.orc_mcdpsk_got_chirp.restype = C.c_long
    n = L.orc_mcdpsk_got_chirp(nc, sps, bits, C.c_float(f_lo), C.c_float(f_hi), C.c_float(fs), training, _p(x, C.c_float),
                               C.c_size_t(len(x)), C.c_float(chirp_cfo), _p(out, C.c_float), C.c_size_t(len(out)), C.byref(cfo))
    assert 0 <= n <= len(out), n
    return out[:n].copy(), float(cfo.value)


def mcdpsk_chirp_receive(nc, samples, threshold=0.15, sps=512, bits=2, f_lo=500.0, f_hi=2500.0, fs=48000.0, training=8):
    """MCDPSKWaveform detectSync -> setFrequencyOffset -> process -> getSoftBits: (llr, info[4], f[3] = {cfo, up corr, down corr}, cfo_after)."""
    x = _f32(samples)
    out = np.zeros(8192, np.float32)
    info = np.zeros(4, np.int32)
    f = np.zeros(3, np.float32)
    after = C.c_float(0)
    L = lib()
    L.orc_mcdpsk_chirp_receive.restype = C.c_long
    n = L.orc_mcdpsk_chirp_receive(nc, sps, bits, C.c_float(f_lo), C.c_float(f_hi), C.c_float(fs), training, _p(x, C.c_float),
                                   C.c_size_t(len(x)), C.c_float(threshold), _p(info, C.c_int32), _p(f, C.c_float), _p(out, C.c_float),
                                   C.c_size_t(len(out)), C.byref(after))
    assert 0 <= n <= len(out), n
    return out[:n].copy(), info, f, float(after.value)


def dpsk_find_preamble(sps, samples, fc=1500.0, fs=48000.0):
    """DPSKDemodulator::findPreamble -> (data_start or -1, est_cfo, phase_offset)."""
    x = _f32(samples)
    L = lib()
    L.orc_dpsk_find_preamble.restype = C.c_long
    cfo, ph = C.c_float(0), C.c_float(0)
    ds = L.orc_dpsk_find_preamble(sps, C.c_float(fc), C.c_float(fs), _p(x, C.c_float), C.c_size_t(len(x)), C.byref(cfo), C.byref(ph))
    return int(ds), float(cfo.value), float(ph.value)


def dpsk_receive(mod, sps, samples, fc=1500.0, fs=48000.0):
    """findPreamble + demodulateSoft (tools/test_dpsk_snr.cpp:66-73) -> (llr, data_start, est_cfo, phase_offset)."""
    ds, cfo, ph = dpsk_find_preamble(sps, samples, fc, fs)
    if not (0 < ds < len(samples)):
        return np.zeros(0, np.float32), ds, cfo, ph
    return dpsk_demod_soft(mod, sps, samples, ds, 1, cfo, ph, fc, fs), ds, cfo, ph


def mcdpsk_demod_soft(nc, samples, sps=512, bits=2, f_lo=500.0, f_hi=2500.0, fs=48000.0, training=8):
    x = _f32(samples)
    out = np.zeros(8192, np.float32)
    cfo = C.c_float(0)
    L = lib()
    L.orc_mcdpsk_demod_soft.restype = C.c_long
    n = L.orc_mcdpsk_demod_soft(nc, sps, bits, C.c_float(f_lo), C.c_float(f_hi), C.c_float(fs), _p(x, C.c_float), C.c_size_t(len(x)),
                                training, _p(out, C.c_float), C.c_size_t(len(out)), C.byref(cfo))
    assert 0 <= n <= len(out), n
    return out[:n].copy(), cfo.value


def frame_encode(rate, frame):
    """v2::encodeFrameWithLDPC(frame, rate) -> uint8 [ncw, 81]."""
    f = np.ascontiguousarray(frame, np.uint8)
    out = np.zeros(256 * 81, np.uint8)
    L = lib()
    L.orc_frame_encode.restype = C.c_long
    n = L.orc_frame_encode(int(rate), _p(f, C.c_uint8), C.c_size_t(len(f)), _p(out, C.c_uint8), C.c_size_t(len(out)))
    assert n > 0, n
    return out[:n * 81].reshape(n, 81).copy()


def frame_decode(rate, soft, num_codewords=None):
    """RxPipeline::decodeFrame(soft_bits, num_codewords) -> (frame bytes, info[5] = {success, type, cw ok, cw failed, expected})."""
    x = _f32(soft).reshape(-1)
    if num_codewords is None:
        num_codewords = len(x) // 648
    assert len(x) >= num_codewords * 648
    out = np.zeros(8192, np.uint8)
    info = np.zeros(5, np.int32)
    L = lib()
    L.orc_frame_decode.restype = C.c_long
    n = L.orc_frame_decode(int(rate), _p(x, C.c_float), C.c_size_t(len(x)), int(num_codewords), _p(out, C.c_uint8), C.c_size_t(len(out)),
                             _p(info, C.c_int32))
    assert 0 <= n <= len(out), n
    return out[:n].copy(), info
