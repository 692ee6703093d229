"""ctypes helpers for the oracle's channel twin (oracle/pu_oracle_channel.c).  TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np

import oracleapi as O


def _lib():
    L = O.lib()
    L.orc_noise_normal.restype = C.c_float
    L.orc_channel_noise_std.restype = C.c_float
    return L


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    _lib().orc_philox4x32_10(c, k, out)
    return [int(v) for v in out]


def noise_normals(seed, n):
    L = _lib()
    return np.array([L.orc_noise_normal(C.c_uint64(seed), C.c_uint32(i)) for i in range(n)], dtype=np.float32)


def fading_normals(seed, n):
    L = _lib()
    z = (C.c_float * 4)()
    out = np.zeros((n, 4), np.float32)
    for i in range(n):
        L.orc_fading_normals(C.c_uint64(seed), C.c_uint32(i), z)
        out[i] = z[:]
    return out


def channel_apply(ch, x, noise_std, seed):
    """ch: projectultra_b200.linksim.ChannelConfig-like (same field names)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros_like(x)
    _lib().orc_channel_apply(C.c_float(ch.delay_spread_ms), C.c_float(ch.doppler_spread_hz), C.c_float(ch.path1_gain),
                             C.c_float(ch.path2_gain), C.c_uint32(ch.sample_rate), int(ch.fading_enabled),
                             int(ch.multipath_enabled), int(ch.noise_enabled), x.ctypes.data_as(C.POINTER(C.c_float)),
                             C.c_size_t(len(x)), C.c_float(noise_std), C.c_uint64(int(seed)),
                             y.ctypes.data_as(C.POINTER(C.c_float)))
    return y


def noise_std(x, snr_db, convention=0):
    x = np.ascontiguousarray(x, dtype=np.float32)
    return float(_lib().orc_channel_noise_std(x.ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(len(x)),
                                              C.c_float(snr_db), int(convention)))


def channel_apply_cfo(x, cfo_hz, sample_rate=48000):
    """orc_channel_apply_cfo: WattersonChannel::applyCFO (hf_channel.hpp:173-232), CFO phase 0 at entry."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros_like(x)
    _lib().orc_channel_apply_cfo(x.ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(len(x)), C.c_float(cfo_hz), C.c_uint32(sample_rate),
                                 y.ctypes.data_as(C.POINTER(C.c_float)))
    return y
