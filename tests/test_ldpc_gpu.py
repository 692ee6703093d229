"""GPU parity tests of the CUDA LDPC decoder (through the C ABI) against the oracle: bit-exact info bytes,
success flag and iteration count, on golden vectors, random codewords at easy / waterfall / stress points,
the demod-realistic all-ties distribution, edge cases, and size-independent properties at 1M codewords."""
import numpy as np
import pytest

import oracleapi as O
import refapi as R

pytestmark = pytest.mark.gpu
RATES = [R.R1_4, R.R1_2, R.R2_3, R.R3_4, R.R5_6]
SIGMAS = {R.R1_4: (0.7, 1.1, 1.2, 1.5), R.R1_2: (0.5, 0.68, 0.74, 0.95), R.R2_3: (0.45, 0.58, 0.64, 0.8),
          R.R3_4: (0.4, 0.55, 0.6, 0.8), R.R5_6: (0.4, 0.55, 0.62, 0.8)}


@pytest.fixture(scope="module")
def ctx():
    from projectultra_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def make_llrs(rate, sigma, B, rng, clip=True):
    k = R.RATE_K[rate]
    data = rng.integers(0, 256, (B, (k + 7) // 8), dtype=np.uint8)
    llr = np.empty((B, 648), np.float32)
    cws = [O.ldpc_encode(rate, d[:k // 8]) for d in data[:16]]
    for b in range(B):
        bits = np.unpackbits(cws[b % 16])[:648].astype(np.float32)
        y = (1 - 2 * bits) + sigma * rng.standard_normal(648).astype(np.float32)
        llr[b] = 2 * y / sigma ** 2
    if clip:
        llr = np.clip(llr, -10, 10)
    return llr.astype(np.float32)


def assert_same(gpu, cpu):
    for name, a, b in zip(("info", "ok", "iters"), gpu, cpu):
        a = np.asarray(a)
        bad = np.nonzero((a != b).reshape(len(a), -1).any(axis=1))[0]
        assert len(bad) == 0, f"{name} differs on codewords {bad[:8]} ({len(bad)} total)"


@pytest.mark.parametrize("rate", RATES)
def test_golden(ctx, golden, rate):
    from projectultra_b200 import capi
    g = golden["ldpc"]
    dec = capi.LdpcDecoder(ctx, rate)
    info, ok, it = dec.decode_batch(g[f"r{rate}_llr"])
    assert_same((info, ok, it), (g[f"r{rate}_info"], g[f"r{rate}_ok"], g[f"r{rate}_iters"]))
    o, okm, itm = dec.decode_soft(g[f"r{rate}_mb_llr"])
    assert (o == g[f"r{rate}_mb_out"]).all() and [int(okm), itm] == list(g[f"r{rate}_mb_ok"])


@pytest.mark.parametrize("rate", RATES)
def test_bitexact_random(ctx, rate):
    from projectultra_b200 import capi
    dec = capi.LdpcDecoder(ctx, rate)
    rng = np.random.default_rng(500 + rate)
    n_conv = 0
    for sigma in SIGMAS[rate]:
        llr = make_llrs(rate, sigma, 400 if rate != R.R1_4 else 250, rng)
        cpu = O.ldpc_decode_batch(rate, llr)
        assert_same(dec.decode_batch(llr), cpu)
        n_conv += int(cpu[1].sum())
    assert n_conv > 0
    # unclipped LLRs (the channel LLR itself is never clamped) and exact zeros (erasures)
    llr = make_llrs(rate, 0.6, 200, rng, clip=False) * 6.0
    llr[:, ::5] = 0.0
    assert_same(dec.decode_batch(llr), O.ldpc_decode_batch(rate, llr))


@pytest.mark.parametrize("rate", RATES)
def test_all_ties_distribution(ctx, rate):
    # demod-realistic LLRs: exactly +-10 with iid flips -> every check-node minimum is a tie
    from projectultra_b200 import capi
    dec = capi.LdpcDecoder(ctx, rate)
    rng = np.random.default_rng(900 + rate)
    k = R.RATE_K[rate]
    cw = O.ldpc_encode(rate, rng.integers(0, 256, k // 8, dtype=np.uint8))
    bits = np.unpackbits(cw)[:648]
    for p in (0.02, 0.06, 0.13):
        flip = rng.random((300, 648)) < p
        llr = np.where((bits[None, :] > 0) ^ flip, -10.0, 10.0).astype(np.float32)
        assert_same(dec.decode_batch(llr), O.ldpc_decode_batch(rate, llr))


def test_max_iterations_and_rate_switch(ctx):
    from projectultra_b200 import capi
    dec = capi.LdpcDecoder(ctx, R.R1_2)
    rng = np.random.default_rng(4)
    llr = make_llrs(R.R1_2, 0.72, 200, rng)
    for mi in (0, 1, 2, 7, 50, 80):
        dec.set_max_iterations(mi)
        assert_same(dec.decode_batch(llr), O.ldpc_decode_batch(R.R1_2, llr, max_iter=mi))
    dec.set_max_iterations(50)
    for rate in (R.R5_6, R.R1_4, R.R1_3, R.R7_8):      # R1/3, R7/8 silently use R1/2 dimensions
        dec.set_rate(rate)
        l2 = make_llrs(rate if rate in R.RATE_K else R.R1_2, 0.7, 64, rng)
        assert_same(dec.decode_batch(l2), O.ldpc_decode_batch(rate, l2))


@pytest.mark.parametrize("rate", RATES)
@pytest.mark.parametrize("blocks", [1, 2, 5])
def test_multiblock_identity(ctx, rate, blocks):
    # tests/test_multiblock_ldpc.cpp:104-230 through the drop-in decodeSoft / decode entry points
    from projectultra_b200 import capi
    dec = capi.LdpcDecoder(ctx, rate)
    rng = np.random.default_rng(rate * 10 + blocks)
    k = R.RATE_K[rate]
    nbytes = (k // 8) * blocks
    data = rng.integers(0, 256, nbytes, dtype=np.uint8)
    cw = capi.ldpc_encode(rate, data)
    llr = np.where(np.unpackbits(cw) == 1, -6.0, 6.0).astype(np.float32)
    out, ok, it = dec.decode_soft(llr)
    assert ok and it == 0 and (out[:nbytes] == data).all()
    o2, ok2, it2 = O.ldpc_decode_soft(rate, llr)
    assert (out == o2).all() and (ok, it) == (ok2, it2)
    out, ok, it = dec.decode_hard(cw)
    assert ok and (out[:nbytes] == data).all()
    # frame sizes 24 / 46 / 279 bytes (:441-488) with noise and a partial trailing block
    for n in (24, 46, 279):
        d = rng.integers(0, 256, n, dtype=np.uint8)
        l = np.where(np.unpackbits(capi.ldpc_encode(rate, d)) == 1, -4.0, 4.0).astype(np.float32)
        l = (l + 1.5 * rng.standard_normal(len(l))).astype(np.float32)[: len(l) - 37]
        a, b = dec.decode_soft(l), O.ldpc_decode_soft(rate, l)
        assert (a[0] == b[0]).all() and a[1:] == b[1:]


def test_edge_cases(ctx):
    from projectultra_b200 import capi
    dec = capi.LdpcDecoder(ctx, R.R1_2)
    out, ok, it = dec.decode_soft(np.zeros(0, np.float32))
    assert len(out) == 0 and not ok                                    # ldpc_decoder.cpp:285-288
    out, ok, it = dec.decode_soft(np.zeros(648, np.float32))
    assert ok and it == 0 and not out.any()
    short = np.full(100, -3.0, np.float32)                              # short input is zero padded (:160-166)
    a, b = dec.decode_soft(short), O.ldpc_decode_soft(R.R1_2, short)
    assert (a[0] == b[0]).all() and a[1:] == b[1:]
    info, ok, it = dec.decode_batch(np.zeros((0, 648), np.float32))
    assert len(info) == 0


@pytest.mark.parametrize("rate", RATES)
def test_device_memory_path_and_properties_at_full_size(ctx, rate):
    """Config 2 size (1M codewords) on device memory: a pool of valid codewords at +-10 must decode to its data
    in 0 iterations everywhere; a noisy batch must give the same answer per pool entry regardless of position
    (determinism), and every codeword flagged ok must satisfy H.c = 0 (checked on a sample with the oracle's H)."""
    import torch
    from projectultra_b200 import capi
    dec = capi.LdpcDecoder(ctx, rate)
    rng = np.random.default_rng(77 + rate)
    k = R.RATE_K[rate]
    P = 256
    data = rng.integers(0, 256, (P, k // 8), dtype=np.uint8)
    cws = np.stack([np.unpackbits(O.ldpc_encode(rate, d))[:648] for d in data])
    pool = np.where(cws == 1, -10.0, 10.0).astype(np.float32)
    B = 1 << 20
    idx = torch.arange(B, device="cuda") % P
    llr = torch.from_numpy(pool).cuda()[idx].contiguous()
    info, ok, it = dec.decode_batch(llr)
    torch.cuda.synchronize()
    assert bool(ok.all()) and int(it.max()) == 0
    want = torch.from_numpy(data).cuda()[idx]
    assert bool((info[:, : k // 8] == want).all())
    # noisy, deterministic per pool entry
    noise = torch.from_numpy((SIGMAS[rate][1] * 4 * rng.standard_normal((P, 648))).astype(np.float32)).cuda()
    noisy = (torch.from_numpy(pool).cuda() * 0.4 + noise)[idx].contiguous()
    info, ok, it = dec.decode_batch(noisy)
    torch.cuda.synchronize()
    assert bool((info.view(B // P, P, -1) == info[:P]).all()) and bool((it.view(-1, P) == it[:P]).all())
    cpu = O.ldpc_decode_batch(rate, noisy[:P].cpu().numpy())
    assert_same((info[:P].cpu().numpy(), ok[:P].cpu().numpy(), it[:P].cpu().numpy()), cpu)
    assert 0 < int(ok[:P].sum()) < P or rate in (R.R3_4, R.R5_6)


def oracle_decode_parallel(rate, llr):
    """The oracle on all host cores (ctypes releases the GIL; the oracle's scratch is thread-local)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    O.ldpc_decode_batch(rate, llr[:1])          # builds the cached code on this thread first
    n = max(1, min(len(os.sched_getaffinity(0)), 32))
    parts = np.array_split(np.arange(len(llr)), n)
    with ThreadPoolExecutor(n) as ex:
        res = list(ex.map(lambda idx: O.ldpc_decode_batch(rate, llr[idx]), parts))
    return tuple(np.concatenate([r[i] for r in res]) for i in range(3))


@pytest.mark.parametrize("rate", RATES)
def test_bitexact_large(ctx, rate):
    """SURVEY test T3 at scale: 36k noisy codewords per rate (easy / waterfall / stress, vectorised generation) --
    info bytes, ok flag and iteration count identical to the oracle, including the never-converging ones."""
    from projectultra_b200 import capi
    dec = capi.LdpcDecoder(ctx, rate)
    rng = np.random.default_rng(4200 + rate)
    k = R.RATE_K[rate]
    cws = np.stack([np.unpackbits(O.ldpc_encode(rate, rng.integers(0, 256, k // 8, dtype=np.uint8)))[:648] for _ in range(32)])
    per = 12000
    sig = SIGMAS[rate]
    seen_fail = seen_ok = 0
    for sigma in (sig[0], sig[1], sig[3]):
        bits = cws[rng.integers(0, 32, per)].astype(np.float32)
        y = (1 - 2 * bits) + sigma * rng.standard_normal((per, 648)).astype(np.float32)
        llr = np.clip(2 * y / sigma ** 2, -10, 10).astype(np.float32)
        cpu = oracle_decode_parallel(rate, llr)
        assert_same(dec.decode_batch(llr), cpu)
        seen_ok += int(cpu[1].sum())
        seen_fail += int((cpu[2] == 50).sum())
    assert seen_ok > 0 and seen_fail > 0
