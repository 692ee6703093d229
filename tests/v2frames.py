"""Test helper: protocol-v2 frame bytes as the reference serialises them (DataFrame::serialize, src/protocol/frame_v2.cpp:502-555;
calculateCodewords(payload, rate) :439-460; CRC-16/CCITT :111-124) -- checked against the compiled reference in tests/test_oracle_fec.py,
used by the GPU tests, which cannot call the reference."""
import numpy as np

BYTES_PER_CW = {0: 20, 1: 27, 2: 40, 3: 54, 4: 60, 5: 67}


def crc16(data):
    crc = 0xFFFF
    for b in bytes(data):
        crc ^= b << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc


def codewords_for(payload_len, rate):
    total, bpc = 17 + payload_len + 2, BYTES_PER_CW[rate]
    return 1 if total <= bpc else 1 + -(-(total - bpc) // (bpc - 2))


def data_frame(payload, rate, seq=7, src_hash=0x123456, dst_hash=0xABCDEF, ftype=0x30, flags=0x01, total_cw=None):
    p = bytes(payload) if isinstance(payload, (bytes, bytearray)) else bytes(np.ascontiguousarray(payload, np.uint8))
    tcw = codewords_for(len(p), rate) if total_cw is None else total_cw
    h = bytes([0x55, 0x4C, ftype, flags, seq >> 8, seq & 0xFF, src_hash >> 16, (src_hash >> 8) & 0xFF, src_hash & 0xFF,
               dst_hash >> 16, (dst_hash >> 8) & 0xFF, dst_hash & 0xFF, tcw, len(p) >> 8, len(p) & 0xFF])
    c = crc16(h)
    body = h + bytes([c >> 8, c & 0xFF]) + p
    f = crc16(body)
    return np.frombuffer(body + bytes([f >> 8, f & 0xFF]), np.uint8).copy()


def control_frame(ftype=0x10, seq=1, src_hash=0x123456, dst_hash=0xABCDEF, payload=b"\x00" * 6, flags=0x01):
    h = bytes([0x55, 0x4C, ftype, flags, seq >> 8, seq & 0xFF, src_hash >> 16, (src_hash >> 8) & 0xFF, src_hash & 0xFF,
               dst_hash >> 16, (dst_hash >> 8) & 0xFF, dst_hash & 0xFF]) + bytes(payload)[:6].ljust(6, b"\x00")
    c = crc16(h)
    return np.frombuffer(h + bytes([c >> 8, c & 0xFF]), np.uint8).copy()


def codeword_llrs(codewords, rng, flip=0.0, mag=6.0, sigma=0.0):
    """[ncw, 81] coded bytes -> float32 [ncw * 648] LLRs (+ => bit 0), optional i.i.d. sign flips / Gaussian jitter."""
    bits = np.unpackbits(np.ascontiguousarray(codewords, np.uint8).reshape(-1)).astype(np.float32)
    l = (1.0 - 2.0 * bits) * mag
    if flip:
        l = l * np.where(rng.random(len(l)) < flip, -1.0, 1.0)
    if sigma:
        l = l + rng.normal(0.0, sigma, len(l))
    return l.astype(np.float32)
