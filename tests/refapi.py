"""ctypes binding for oracle/_ref/libpu_ref.so (the UNMODIFIED reference, compiled by
oracle/ref_build/Makefile).  TEST INFRASTRUCTURE: imported only by tests/, smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never on the product path."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(os.path.dirname(_HERE), "oracle", "_ref", "libpu_ref.so")

# ultra::Modulation (include/ultra/types.hpp:27-39)
DBPSK, BPSK, DQPSK, QPSK, D8PSK, QAM8, QAM16, QAM32, QAM64, QAM256 = 0, 1, 2, 3, 4, 5, 6, 7, 8, 10
BITS_PER_SYM = {DBPSK: 1, BPSK: 1, DQPSK: 2, QPSK: 2, D8PSK: 3, QAM8: 3, QAM16: 4, QAM32: 5, QAM64: 6, QAM256: 8}
# ultra::CodeRate (include/ultra/types.hpp:92-101)
R1_4, R1_3, R1_2, R2_3, R3_4, R5_6, R7_8 = range(7)
RATE_K = {R1_4: 162, R1_2: 324, R2_3: 432, R3_4: 486, R5_6: 540}
STAGE_SCALARS = 10


class ModemConfig(C.Structure):
    """POD mirror of ultra::ModemConfig; same layout as pu_modem_config (include/pu/pu_capi.h)."""
    _fields_ = [("sample_rate", C.c_uint32), ("center_freq", C.c_uint32), ("fft_size", C.c_uint32),
                ("num_carriers", C.c_uint32), ("cp_mode", C.c_uint32), ("symbol_guard", C.c_uint32),
                ("pilot_spacing", C.c_uint32), ("use_pilots", C.c_uint32), ("modulation", C.c_uint32),
                ("code_rate", C.c_uint32), ("output_scale", C.c_float), ("tx_cfo_hz", C.c_float)]

    @property
    def cp(self):
        return {0: 32, 1: 48, 2: 64}[self.cp_mode] * (self.fft_size // 512)

    @property
    def symbol_samples(self):
        return self.fft_size + self.cp + self.symbol_guard

    @property
    def n_pilots(self):
        if not self.use_pilots:
            return 0
        return (self.num_carriers + self.pilot_spacing - 1) // self.pilot_spacing

    @property
    def n_data(self):
        return self.num_carriers - self.n_pilots

    @property
    def bits_per_symbol(self):
        return self.n_data * BITS_PER_SYM[self.modulation]


def config_m1(mod=DQPSK, rate=R1_2, use_pilots=None, pilot_spacing=2):
    """M1: ModemConfig defaults (types.hpp:139-195): 512-FFT, 30 carriers, CP 48, guard 4."""
    if use_pilots is None:
        use_pilots = mod not in (DBPSK, DQPSK, D8PSK)  # tools/test_mode_snr.cpp:30
    return ModemConfig(48000, 1500, 512, 30, 1, 4, pilot_spacing, int(use_pilots), mod, rate, 40.0, 0.0)


def config_m3(mod=QAM32, rate=R3_4, use_pilots=None, pilot_spacing=4):
    """M3: presets::nvis_mode() (types.hpp:342-355) 1024-FFT, 59 carriers, CP 96, guard 0;
    coherent modes get pilots/4 (tools/test_nvis_mode.cpp:209-212)."""
    if use_pilots is None:
        use_pilots = mod not in (DBPSK, DQPSK, D8PSK)
    return ModemConfig(48000, 1500, 1024, 59, 1, 0, pilot_spacing if use_pilots else 2, int(use_pilots),
                       mod, rate, 40.0, 0.0)


_lib = None


def available():
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_SO)
        L.ref_ldpc_encode.restype = C.c_long
        L.ref_ldpc_decode_soft.restype = C.c_long
        L.ref_ldpc_decode_hard.restype = C.c_long
        L.ref_ofdm_tx.restype = C.c_long
        L.ref_ofdm_presynced.restype = C.c_long
        L.ref_ofdm_process.restype = C.c_long
        L.ref_ofdm_process_info.restype = C.c_long
        L.ref_ofdm_presynced_stages.restype = C.c_long
        L.ref_dpsk_modulate.restype = C.c_long
        L.ref_dpsk_demod_soft.restype = C.c_long
        L.ref_time_presynced_decode.restype = C.c_double
        L.ref_time_ldpc_decode.restype = C.c_double
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a):
    return np.ascontiguousarray(np.frombuffer(bytes(a), dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else a,
                                dtype=np.uint8)


def ldpc_encode(rate, data):
    d = _u8(data)
    out = np.zeros(81 * (len(d) * 8 // RATE_K[rate] + 2), np.uint8)
    n = lib().ref_ldpc_encode(rate, _p(d, C.c_uint8), C.c_size_t(len(d)), _p(out, C.c_uint8), C.c_size_t(len(out)))
    assert n >= 0
    return out[:n].copy()


def ldpc_decode_soft(rate, llr, max_iter=-1):
    x = _f32(llr)
    out = np.zeros(128 * (len(x) // 648 + 2), np.uint8)
    ok, it = C.c_int(0), C.c_int(0)
    n = lib().ref_ldpc_decode_soft(rate, max_iter, _p(x, C.c_float), C.c_size_t(len(x)), _p(out, C.c_uint8),
                                   C.c_size_t(len(out)), C.byref(ok), C.byref(it))
    assert n >= 0
    return out[:n].copy(), bool(ok.value), it.value


def ldpc_decode_hard(rate, coded):
    d = _u8(coded)
    out = np.zeros(128 * (len(d) // 81 + 2), np.uint8)
    ok, it = C.c_int(0), C.c_int(0)
    n = lib().ref_ldpc_decode_hard(rate, _p(d, C.c_uint8), C.c_size_t(len(d)), _p(out, C.c_uint8),
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
    r = lib().ref_ldpc_decode_batch(rate, max_iter, _p(x, C.c_float), C.c_size_t(B), _p(out, C.c_uint8),
                                    C.c_size_t(kb), _p(ok, C.c_uint8), _p(it, C.c_int32))
    assert r == 0
    return out, ok, it


def channel_interleave(bps, x, inverse=False, total=648):
    x = _f32(x)
    out = np.zeros(total, np.float32)
    lib().ref_channel_interleave(C.c_size_t(bps), C.c_size_t(total), _p(x, C.c_float), C.c_size_t(len(x)),
                                 _p(out, C.c_float), int(inverse))
    return out


def channel_interleave_bytes(bps, data, inverse=False, total=648):
    d = _u8(data)
    out = np.zeros((total + 7) // 8, np.uint8)
    lib().ref_channel_interleave_bytes(C.c_size_t(bps), C.c_size_t(total), _p(d, C.c_uint8), C.c_size_t(len(d)),
                                       _p(out, C.c_uint8), int(inverse))
    return out


def block_interleave(rows, cols, x, inverse=False):
    x = _f32(x)
    out = np.zeros(len(x), np.float32)
    lib().ref_block_interleave(C.c_size_t(rows), C.c_size_t(cols), _p(x, C.c_float), C.c_size_t(len(x)),
                               _p(out, C.c_float), int(inverse))
    return out


def block_interleave_bytes(rows, cols, data, inverse=False):
    d = _u8(data)
    out = np.zeros((rows * cols + 7) // 8, np.uint8)
    lib().ref_block_interleave_bytes(C.c_size_t(rows), C.c_size_t(cols), _p(d, C.c_uint8), C.c_size_t(len(d)),
                                     _p(out, C.c_uint8), int(inverse))
    return out


def fft(x, inverse=False):
    z = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.zeros_like(z)
    lib().ref_fft(C.c_size_t(len(z)), _p(z.view(np.float32), C.c_float), _p(out.view(np.float32), C.c_float),
                  int(inverse))
    return out


def nco(freq, fs, n):
    out = np.zeros(n, np.complex64)
    lib().ref_nco(C.c_float(freq), C.c_float(fs), C.c_size_t(n), _p(out.view(np.float32), C.c_float))
    return out


def soft_demap(mod, sym, prev=1 + 0j, nv=0.1):
    out = np.zeros(8, np.float32)
    n = lib().ref_soft_demap(mod, C.c_float(sym.real), C.c_float(sym.imag), C.c_float(prev.real),
                             C.c_float(prev.imag), C.c_float(nv), _p(out, C.c_float))
    return out[:n].copy()


def ofdm_tx(cfg, data, layout=0):
    d = _u8(data)
    cap = 400000
    out = np.zeros(cap, np.float32)
    n = lib().ref_ofdm_tx(C.byref(cfg), layout, _p(d, C.c_uint8), C.c_size_t(len(d)), _p(out, C.c_float),
                          C.c_size_t(cap))
    assert n >= 0, n
    return out[:n].copy()


def ofdm_presynced(cfg, samples, training=2, cfo_mode=1, cfo_hz=0.0, cfo_phase=0.0):
    x = _f32(samples)
    cap = 8192
    out = np.zeros(cap, np.float32)
    snr, fc = C.c_float(0), C.c_float(0)
    n = lib().ref_ofdm_presynced(C.byref(cfg), _p(x, C.c_float), C.c_size_t(len(x)), training, cfo_mode,
                                 C.c_float(cfo_hz), C.c_float(cfo_phase), _p(out, C.c_float), C.c_size_t(cap),
                                 C.byref(snr), C.byref(fc))
    assert n >= 0, n
    return out[:n].copy(), snr.value, fc.value


def ofdm_presynced_batch(cfg, samples, n_llr, training=2, cfo_mode=1, cfo_hz=None, cfo_phase=None):
    x = _f32(samples)
    B, L = x.shape
    out = np.zeros((B, n_llr), np.float32)
    counts = np.zeros(B, np.int32)
    f = _f32(cfo_hz) if cfo_hz is not None else None
    p = _f32(cfo_phase) if cfo_phase is not None else None
    lib().ref_ofdm_presynced_batch(C.byref(cfg), _p(x, C.c_float), C.c_size_t(B), C.c_size_t(L), training, cfo_mode,
                                   _p(f, C.c_float) if f is not None else None,
                                   _p(p, C.c_float) if p is not None else None,
                                   _p(out, C.c_float), C.c_size_t(n_llr), _p(counts, C.c_int32))
    return out, counts


def ofdm_process(cfg, samples, chunk=960):
    x = _f32(samples)
    out = np.zeros(2048, np.float32)
    synced, snr = C.c_int(0), C.c_float(0)
    n = lib().ref_ofdm_process(C.byref(cfg), _p(x, C.c_float), C.c_size_t(len(x)), C.c_size_t(chunk),
                               _p(out, C.c_float), C.c_size_t(len(out)), C.byref(synced), C.byref(snr))
    assert n >= 0
    return out[:n].copy(), bool(synced.value), snr.value


def ofdm_process_info(cfg, samples, chunk=960):
    """OFDMDemodulator::process in `chunk`-sample pieces + getSoftBits(): (llr, synced, sync_offset, coarse_cfo)."""
    x = _f32(samples)
    out = np.zeros(2048, np.float32)
    synced, off, cfo = C.c_int(0), C.c_long(0), C.c_float(0)
    n = lib().ref_ofdm_process_info(C.byref(cfg), _p(x, C.c_float), C.c_size_t(len(x)), C.c_size_t(chunk),
                                    _p(out, C.c_float), C.c_size_t(len(out)), C.byref(synced), C.byref(off), C.byref(cfo))
    assert n >= 0
    return out[:n].copy(), bool(synced.value), int(off.value), float(cfo.value)


def ofdm_presynced_stages(cfg, samples, training=2, cfo_mode=1, cfo_hz=0.0, cfo_phase=0.0, max_sym=64):
    x = _f32(samples)
    nd, npil = cfg.n_data, cfg.n_pilots
    nu = nd + npil
    carriers = np.zeros(nu, np.int32)
    nd_o, np_o = C.c_int32(0), C.c_int32(0)
    lts = np.zeros((training, nu), np.complex64)
    h_lts = np.zeros(nu, np.complex64)
    bins = np.zeros((max_sym, nu), np.complex64)
    h = np.zeros((max_sym, nu), np.complex64)
    eq = np.zeros((max_sym, nd), np.complex64)
    nv = np.zeros((max_sym, nd), np.float32)
    sc = np.zeros((max_sym, STAGE_SCALARS), np.float32)
    cap = 16384
    llr = np.zeros(cap, np.float32)
    n_llr = C.c_long(0)
    fp = lambda a: _p(a.view(np.float32), C.c_float)
    ns = lib().ref_ofdm_presynced_stages(
        C.byref(cfg), _p(x, C.c_float), C.c_size_t(len(x)), training, cfo_mode, C.c_float(cfo_hz),
        C.c_float(cfo_phase), max_sym, _p(carriers, C.c_int32), C.byref(nd_o), C.byref(np_o),
        fp(lts), fp(h_lts), fp(bins), fp(h), fp(eq), _p(nv, C.c_float), _p(sc, C.c_float),
        _p(llr, C.c_float), C.c_size_t(cap), C.byref(n_llr))
    assert nd_o.value == nd and np_o.value == npil, (nd_o.value, np_o.value, nd, npil)
    return dict(n_sym=ns, carriers=carriers, lts_bins=lts, h_lts=h_lts, bins=bins[:ns], h=h[:ns], eq=eq[:ns],
                nv=nv[:ns], scalars=sc[:ns], llr=llr[:n_llr.value].copy())


def watterson(x, snr_db, delay_ms, doppler_hz, g1=0.707, g2=0.707, fading=True, multipath=True, noise=True, seed=42):
    x = _f32(x)
    out = np.zeros_like(x)
    lib().ref_watterson(C.c_float(snr_db), C.c_float(delay_ms), C.c_float(doppler_hz), C.c_float(g1), C.c_float(g2),
                        int(fading), int(multipath), int(noise), C.c_uint32(seed), _p(x, C.c_float),
                        C.c_size_t(len(x)), _p(out, C.c_float))
    return out


def watterson_cfo(x, cfo_hz):
    """WattersonChannel::process with only the CFO injector active (applyCFO, hf_channel.hpp:173-232)."""
    x = _f32(x)
    out = np.zeros_like(x)
    lib().ref_watterson_cfo(C.c_float(cfo_hz), _p(x, C.c_float), C.c_size_t(len(x)), _p(out, C.c_float))
    return out


def tools_apply_cfo(x, cfo_hz, sample_rate=48000.0):
    """tools/test_iwaveform.cpp:67-118 around the compiled reference FFT (ref_harness.cpp: ref_tools_apply_cfo)."""
    out = _f32(x).copy()
    lib().ref_tools_apply_cfo(_p(out, C.c_float), C.c_size_t(len(out)), C.c_float(cfo_hz), C.c_float(sample_rate))
    return out


def dpsk_modulate(mod_order, sps, data, with_preamble=True):
    d = _u8(data)
    cap = 2_000_000
    out = np.zeros(cap, np.float32)
    n = lib().ref_dpsk_modulate(mod_order, sps, int(with_preamble), _p(d, C.c_uint8), C.c_size_t(len(d)),
                                _p(out, C.c_float), C.c_size_t(cap))
    assert n >= 0
    return out[:n].copy()


def dpsk_demod_soft(mod_order, sps, samples, data_start=-1):
    x = _f32(samples)
    out = np.zeros(8192, np.float32)
    found = C.c_long(-2)
    n = lib().ref_dpsk_demod_soft(mod_order, sps, _p(x, C.c_float), C.c_size_t(len(x)), C.c_long(data_start),
                                  _p(out, C.c_float), C.c_size_t(len(out)), C.byref(found))
    assert n >= 0
    return out[:n].copy(), found.value


def dpsk_receive(mod, sps, samples, fc=1500.0, fs=48000.0):
    """tools/test_dpsk_snr.cpp:66-73: findPreamble + demodulateSoft -> (llr, data_start, est_cfo, phase_offset)."""
    x = _f32(samples)
    out = np.zeros(4096, np.float32)
    L = lib()
    L.ref_dpsk_receive.restype = C.c_long
    ds, cfo, ph = C.c_long(0), C.c_float(0), C.c_float(0)
    n = L.ref_dpsk_receive(mod, sps, C.c_float(fc), C.c_float(fs), _p(x, C.c_float), C.c_size_t(len(x)), C.byref(ds), C.byref(cfo),
                           C.byref(ph), _p(out, C.c_float), C.c_size_t(len(out)))
    assert 0 <= n <= len(out), n
    return out[:n].copy(), int(ds.value), float(cfo.value), float(ph.value)


def chirp_generate(fs=48000.0, tx_cfo=0.0):
    out = np.zeros(80000, np.float32)
    L = lib()
    L.ref_chirp_generate.restype = C.c_long
    n = L.ref_chirp_generate(C.c_float(fs), C.c_float(tx_cfo), _p(out, C.c_float), C.c_size_t(len(out)))
    assert n >= 0
    return out[:n].copy()


def ofdm_chirp_receive(cfg, samples, threshold=0.15):
    """tools/test_iwaveform.cpp:127-160 on an OFDM_CHIRP frame: (llr, info[4] = {success, up start, down start, training start}, cfo)."""
    x = _f32(samples)
    out = np.zeros(8192, np.float32)
    info = np.zeros(4, np.int32)
    cfo = C.c_float(0)
    L = lib()
    L.ref_ofdm_chirp_receive.restype = C.c_long
    n = L.ref_ofdm_chirp_receive(C.byref(cfg), _p(x, C.c_float), C.c_size_t(len(x)), C.c_float(threshold), _p(info, C.c_int32),
                                 C.byref(cfo), _p(out, C.c_float), C.c_size_t(len(out)))
    assert 0 <= n <= len(out), n
    return out[:n].copy(), info, float(cfo.value)


def mcdpsk_got_chirp(nc, samples, chirp_cfo, sps=512, bits=2, f_lo=500.0, f_hi=2500.0, fs=48000.0, training=8):
    """setChirpDetected(cfo) -> process(training + ref + data) -> getSoftBits(): (llr, ready, cfo_after)."""
    x = _f32(samples)
    out = np.zeros(8192, np.float32)
    ready, cfo = C.c_int(0), C.c_float(0)
    L = lib()
    L.ref_mcdpsk_got_chirp.restype = C.c_long
    n = L.ref_mcdpsk_got_chirp(nc, sps, bits, C.c_float(f_lo), C.c_float(f_hi), C.c_float(fs), training, _p(x, C.c_float),
                               C.c_size_t(len(x)), C.c_float(chirp_cfo), _p(out, C.c_float), C.c_size_t(len(out)), C.byref(ready), C.byref(cfo))
    assert 0 <= n <= len(out), n
    return out[:n].copy(), bool(ready.value), float(cfo.value)


def mcdpsk_chirp_receive(nc, samples, threshold=0.15, sps=512, bits=2, f_lo=500.0, f_hi=2500.0, fs=48000.0, training=8):
    """MCDPSKWaveform detectSync -> setFrequencyOffset -> process -> getSoftBits: (llr, info[4], f[3] = {cfo, up corr, down corr}, cfo_after)."""
    x = _f32(samples)
    out = np.zeros(8192, np.float32)
    info = np.zeros(4, np.int32)
    f = np.zeros(3, np.float32)
    after = C.c_float(0)
    L = lib()
    L.ref_mcdpsk_chirp_receive.restype = C.c_long
    n = L.ref_mcdpsk_chirp_receive(nc, sps, bits, C.c_float(f_lo), C.c_float(f_hi), C.c_float(fs), training, _p(x, C.c_float),
                                   C.c_size_t(len(x)), C.c_float(threshold), _p(info, C.c_int32), _p(f, C.c_float), _p(out, C.c_float),
                                   C.c_size_t(len(out)), C.byref(after))
    assert 0 <= n <= len(out), n
    return out[:n].copy(), info, f, float(after.value)


def time_presynced_decode(cfg, samples, rate):
    x = _f32(samples)
    B, L = x.shape
    kb = (RATE_K.get(rate, 324) + 7) // 8
    info = np.zeros((B, kb), np.uint8)
    ok = np.zeros(B, np.uint8)
    t = lib().ref_time_presynced_decode(C.byref(cfg), _p(x, C.c_float), C.c_size_t(B), C.c_size_t(L), rate,
                                        _p(info, C.c_uint8), C.c_size_t(kb), _p(ok, C.c_uint8))
    return t, info, ok


def time_ldpc_decode(rate, llr, max_iter=-1):
    x = _f32(llr).reshape(-1, 648)
    B = x.shape[0]
    kb = (RATE_K.get(rate, 324) + 7) // 8
    out = np.zeros((B, kb), np.uint8)
    ok = np.zeros(B, np.uint8)
    it = np.zeros(B, np.int32)
    t = lib().ref_time_ldpc_decode(rate, max_iter, _p(x, C.c_float), C.c_size_t(B), _p(out, C.c_uint8),
                                   C.c_size_t(kb), _p(ok, C.c_uint8), _p(it, C.c_int32))
    return t, out, ok, it


# ---------------------------------------------------------------- DPSK (oracle/ref_build/ref_psk.cpp)
def dpsk_tx(mod, sps, data, layout=0, fc=1500.0, fs=48000.0):
    """DPSKModulator: layout 0 Barker preamble + data, 1 reference symbol + data, 2 data only.  mod 0/1/2 = DBPSK/DQPSK/D8PSK."""
    d = _u8(data)
    cap = 400000
    out = np.zeros(cap, np.float32)
    L = lib()
    L.ref_dpsk_tx.restype = C.c_long
    n = L.ref_dpsk_tx(mod, sps, C.c_float(fc), C.c_float(fs), layout, _p(d, C.c_uint8), C.c_size_t(len(d)), _p(out, C.c_float),
                      C.c_size_t(cap))
    assert n >= 0, n
    return out[:n].copy()


def dpsk_demod_soft_ex(mod, sps, samples, data_start, ref_mode=0, est_cfo=0.0, phase_off=0.0, fc=1500.0, fs=48000.0):
    x = _f32(samples)
    out = np.zeros(4096, np.float32)
    L = lib()
    L.ref_dpsk_demod_soft_ex.restype = C.c_long
    n = L.ref_dpsk_demod_soft_ex(mod, sps, C.c_float(fc), C.c_float(fs), _p(x, C.c_float), C.c_size_t(len(x)), C.c_long(data_start),
                                 ref_mode, C.c_float(est_cfo), C.c_float(phase_off), _p(out, C.c_float), C.c_size_t(len(out)))
    assert 0 <= n <= len(out), n
    return out[:n].copy()


def mcdpsk_tx(nc, data, sps=512, bits=2, f_lo=500.0, f_hi=2500.0, fs=48000.0, training=8):
    d = _u8(data)
    cap = 400000
    out = np.zeros(cap, np.float32)
    L = lib()
    L.ref_mcdpsk_tx.restype = C.c_long
    n = L.ref_mcdpsk_tx(nc, sps, bits, C.c_float(f_lo), C.c_float(f_hi), C.c_float(fs), training, _p(d, C.c_uint8), C.c_size_t(len(d)),
                        _p(out, C.c_float), C.c_size_t(cap))
    assert n >= 0, n
    return out[:n].copy()


def mcdpsk_demod_soft(nc, samples, sps=512, bits=2, f_lo=500.0, f_hi=2500.0, fs=48000.0, training=8):
    x = _f32(samples)
    out = np.zeros(8192, np.float32)
    cfo = C.c_float(0)
    L = lib()
    L.ref_mcdpsk_demod_soft.restype = C.c_long
    n = L.ref_mcdpsk_demod_soft(nc, sps, bits, C.c_float(f_lo), C.c_float(f_hi), C.c_float(fs), training, _p(x, C.c_float),
                                C.c_size_t(len(x)), _p(out, C.c_float), C.c_size_t(len(out)), C.byref(cfo))
    assert 0 <= n <= len(out), n
    return out[:n].copy(), cfo.value


def frame_encode(rate, frame):
    """v2::encodeFrameWithLDPC(frame, rate) -> uint8 [ncw, 81]."""
    f = np.ascontiguousarray(frame, np.uint8)
    out = np.zeros(256 * 81, np.uint8)
    L = lib()
    L.ref_frame_encode.restype = C.c_long
    n = L.ref_frame_encode(int(rate), _p(f, C.c_uint8), C.c_size_t(len(f)), _p(out, C.c_uint8), C.c_size_t(len(out)))
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
    L.ref_frame_decode.restype = C.c_long
    n = L.ref_frame_decode(int(rate), _p(x, C.c_float), C.c_size_t(len(x)), int(num_codewords), _p(out, C.c_uint8), C.c_size_t(len(out)),
                             _p(info, C.c_int32))
    assert 0 <= n <= len(out), n
    return out[:n].copy(), info


def data_frame_serialize(rate, payload, src="N0CALL", dst="W1AW", seq=7):
    """DataFrame::makeData(src, dst, seq, payload, rate).serialize()."""
    p = np.ascontiguousarray(payload, np.uint8)
    out = np.zeros(70000, np.uint8)
    L = lib()
    L.ref_data_frame_serialize.restype = C.c_long
    n = L.ref_data_frame_serialize(int(rate), src.encode(), dst.encode(), int(seq), _p(p, C.c_uint8), C.c_size_t(len(p)), _p(out, C.c_uint8),
                                   C.c_size_t(len(out)))
    assert n > 0, n
    return out[:n].copy()


def control_frame_serialize(src="N0CALL", dst="W1AW"):
    out = np.zeros(64, np.uint8)
    L = lib()
    L.ref_control_frame_serialize.restype = C.c_long
    n = L.ref_control_frame_serialize(src.encode(), dst.encode(), _p(out, C.c_uint8), C.c_size_t(len(out)))
    assert n == 20, n
    return out[:n].copy()
