"""GPU parity tests (run on the B200 with -m gpu): every kernel is called THROUGH THE C ABI and compared
bit for bit with the CPU oracle on the same int32 keys and ciphertexts."""
import itertools

import numpy as np
import pytest

import tfhe_jl_b200 as T
from tfhe_jl_b200 import _cabi
from conftest import PLAIN_GATES, random_torus
from oracle import oracle as O

pytestmark = pytest.mark.gpu
N = 1024


def make_ctx(keys, flags=_cabi.FLAG_SPLIT_FFT, parties=1):
    P = keys.params
    ctx = T.Context(n=P.n, N=P.N, k=P.k, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, parties=parties, flags=flags)
    ctx.load_bk(keys.bk)
    ctx.load_ksk(keys.ksk)
    return ctx


@pytest.fixture(scope="module")
def gctx80(keys80):
    return make_ctx(keys80)


@pytest.fixture(scope="module")
def gctx80_unsplit(keys80):
    return make_ctx(keys80, _cabi.FLAG_UNSPLIT_FFT)


@pytest.fixture(scope="module")
def gctx80_small(keys80_small):
    return make_ctx(keys80_small)


@pytest.fixture(scope="module")
def gctx128(keys128):
    return make_ctx(keys128)


# ---- K1: transformed_mul (polynomials.jl:142-144) vs the exact integer convolution ----
def test_polymul_exact_random_and_adversarial(gctx80_small):
    rng = np.random.default_rng(0)
    x = random_torus(rng, 64, N); y = random_torus(rng, 64, N)
    # adversarial rows: extreme magnitudes, constant signs (worst case for the rounding bound)
    x[0] = 2 ** 31 - 1; y[0] = 2 ** 31 - 1
    x[1] = -(2 ** 31); y[1] = -(2 ** 31)
    x[2] = -(2 ** 31); y[2] = 2 ** 31 - 1
    x[3] = np.where(rng.integers(0, 2, N) == 1, 2 ** 31 - 1, -(2 ** 31)); y[3] = np.where(rng.integers(0, 2, N) == 1, 2 ** 31 - 1, -(2 ** 31))
    x[4] = 0x7FFF8000; y[4] = 0x7FFF8000          # pieces at their largest magnitude
    x[5] = 0; x[6] = 0; x[6, 0] = 1               # zero and identity
    x[7] = rng.integers(0, 2, N); x[8] = rng.integers(-512, 512, N)   # key bits, 10-bit digits
    got = gctx80_small.polymul(x, y)
    for i in range(x.shape[0]):
        assert np.array_equal(got[i], O.polymul(x[i], y[i], O.ROUTE_EXACT)), f"row {i}"
    assert np.array_equal(got[6], y[6]) and not got[5].any()


def test_polymul_empty_batch(gctx80_small):
    assert gctx80_small.polymul(np.empty((0, N), np.int32), np.empty((0, N), np.int32)).shape == (0, N)


# ---- K2: tgsw_extern_mul (tgsw.jl:125-129) ----
@pytest.mark.parametrize("flags", [_cabi.FLAG_SPLIT_FFT, _cabi.FLAG_UNSPLIT_FFT], ids=["split", "unsplit"])
def test_extern_product_matches_exact_oracle(keys80_small, octx80_small, flags):
    ctx = make_ctx(keys80_small, flags)
    rng = np.random.default_rng(1)
    acc = random_torus(rng, 12, 2, N)
    acc[0] = 0; acc[1] = 2 ** 31 - 1; acc[2] = -(2 ** 31)     # zero / extreme accumulators
    idx = rng.integers(0, 24, 12).astype(np.int32)
    got = ctx.extern_product(acc, idx)
    for g in range(12):
        assert np.array_equal(got[g], octx80_small.extern_mul(int(idx[g]), acc[g], O.ROUTE_EXACT)), f"item {g}"


def test_extern_product_128bit_params(keys128, octx128, gctx128):
    rng = np.random.default_rng(2)
    acc = random_torus(rng, 4, 2, N)
    idx = np.array([0, 1, 300, 629], dtype=np.int32)
    got = gctx128.extern_product(acc, idx)
    for g in range(4):
        assert np.array_equal(got[g], octx128.extern_mul(int(idx[g]), acc[g], O.ROUTE_EXACT))


def test_extern_product_rejects_bad_index(gctx80_small):
    with pytest.raises(T.TFHEB200Error):
        gctx80_small.extern_product(np.zeros((1, 2, N), np.int32), np.array([24], np.int32))


# ---- K3: blind_rotate (bootstrap.jl:32-39), accumulator equality after every iteration count ----
def test_blind_rotate_prefixes_match_exact_oracle(octx80_small, gctx80_small):
    rng = np.random.default_rng(3)
    acc = random_torus(rng, 3, 2, N)
    bara = rng.integers(-N, N, (3, 24)).astype(np.int32)
    bara[0, 2] = 0; bara[1, :] = 0; bara[2, 0] = -N           # skipped iterations, all-zero row, extreme shift
    for n_iter in (0, 1, 2, 5):
        got = gctx80_small.blind_rotate(acc, bara, n_iter)
        for g in range(3):
            assert np.array_equal(got[g], octx80_small.blind_rotate(acc[g], bara[g], O.ROUTE_EXACT, n_iter)), (n_iter, g)


def test_blind_rotate_full_small_key(octx80_small, gctx80_small):
    rng = np.random.default_rng(4)
    acc = random_torus(rng, 5, 2, N)
    bara = rng.integers(-N, N, (5, 24)).astype(np.int32)
    got = gctx80_small.blind_rotate(acc, bara)
    for g in range(5):
        assert np.array_equal(got[g], octx80_small.blind_rotate(acc[g], bara[g], O.ROUTE_FFT))


# ---- bootstrap_wo_keyswitch / keyswitch / bootstrap at full 80-bit size ----
def test_bootstrap_wo_ks_and_keyswitch_80(keys80, octx80, gctx80):
    rng = O.Rng(5)
    bits = np.random.default_rng(5).integers(0, 2, 6).astype(bool)
    x = O.encrypt(rng, keys80, bits)
    u = gctx80.bootstrap_wo_ks(x)
    assert np.array_equal(u, octx80.bootstrap_wo_ks(x))
    assert np.array_equal(gctx80.keyswitch(u), octx80.keyswitch(u))
    assert np.array_equal(gctx80.bootstrap(x), octx80.bootstrap(x))
    # keyswitch on arbitrary (non-bootstrapped) inputs, incl. all-zero digits and extreme words
    v = random_torus(np.random.default_rng(6), 5, N + 1)
    v[0] = 0; v[1] = -(2 ** 31); v[2] = 2 ** 31 - 1
    assert np.array_equal(gctx80.keyswitch(v), octx80.keyswitch(v))


BINARY = [O.NAND, O.OR, O.AND, O.XOR, O.XNOR, O.NOR, O.ANDNY, O.ANDYN, O.ORNY, O.ORYN]


@pytest.mark.parametrize("op", BINARY + [O.NOT, O.CONSTANT, O.MUX], ids=lambda g: O.GATE_NAMES[g])
def test_gate_ciphertexts_equal_oracle_80(op, keys80, octx80, gctx80):
    """Every gate of gates.jl: ciphertext-identical to the oracle AND the truth table of test/runtests.jl:8-40."""
    nargs = 3 if op == O.MUX else (1 if op in (O.NOT, O.CONSTANT) else 2)
    bits = np.array(list(itertools.product([False, True], repeat=nargs)))
    rng = O.Rng(100 + op)
    if op == O.CONSTANT:
        flags = np.zeros((2, keys80.params.n + 1), dtype=np.int32); flags[1, 0] = 1
        got = gctx80.gate(op, flags)
        assert np.array_equal(got, octx80.gate(op, flags))
        assert O.decrypt(keys80, got).tolist() == [False, True]
        return
    cts = [O.encrypt(rng, keys80, bits[:, i]) for i in range(nargs)]
    got = gctx80.gate(op, *cts)
    assert np.array_equal(got, octx80.gate(op, *cts))
    plain = {O.NOT: lambda a: ~a, O.MUX: lambda a, b, c: np.where(a, b, c)}.get(op) or PLAIN_GATES[op]
    assert np.array_equal(O.decrypt(keys80, got), plain(*[bits[:, i] for i in range(nargs)]))


def test_unsplit_mode_equals_split_mode(keys80, gctx80, gctx80_unsplit):
    rng = O.Rng(7)
    bits = np.random.default_rng(7).integers(0, 2, (8, 2)).astype(bool)
    x, y = O.encrypt(rng, keys80, bits[:, 0]), O.encrypt(rng, keys80, bits[:, 1])
    assert np.array_equal(gctx80.gate(O.NAND, x, y), gctx80_unsplit.gate(O.NAND, x, y))


def test_nand_128(keys128, octx128, gctx128):
    """test/runtests.jl:43-57"""
    bits = np.array(list(itertools.product([False, True], repeat=2)))
    rng = O.Rng(8)
    x, y = O.encrypt(rng, keys128, bits[:, 0]), O.encrypt(rng, keys128, bits[:, 1])
    got = gctx128.gate(O.NAND, x, y)
    assert np.array_equal(got, octx128.gate(O.NAND, x, y))
    assert np.array_equal(O.decrypt(keys128, got), ~(bits[:, 0] & bits[:, 1]))


def test_ragged_and_empty_batches(keys80, octx80, gctx80):
    rng = O.Rng(9)
    for count in (0, 1, 3):      # odd counts exercise the partially filled last CTA
        bits = np.random.default_rng(count).integers(0, 2, (count, 2)).astype(bool)
        x, y = O.encrypt(rng, keys80, bits[:, 0]), O.encrypt(rng, keys80, bits[:, 1])
        got = gctx80.gate(O.AND, x, y, count=count)
        assert got.shape == (count, keys80.params.n + 1)
        if count:
            assert np.array_equal(got, octx80.gate(O.AND, x, y))


def test_missing_keys_is_an_error(keys80):
    P = keys80.params
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit)
    with pytest.raises(T.TFHEB200Error) as e:
        ctx.gate(O.NAND, np.zeros((1, P.n + 1), np.int32), np.zeros((1, P.n + 1), np.int32))
    assert e.value.code == _cabi.ENOKEY


def test_large_batch_properties(keys80, octx80, gctx80):
    """At a batch far beyond what the oracle can check gate by gate: every output decrypts to the plaintext
    NAND, phase error stays inside the 1/16 contract (gates.jl:1-6), a random sample is ciphertext-identical
    to the oracle, and the result does not depend on how the batch is chunked."""
    B = 4096
    prng = np.random.default_rng(10)
    bits = prng.integers(0, 2, (B, 2)).astype(bool)
    rng = O.Rng(10)
    x, y = O.encrypt(rng, keys80, bits[:, 0]), O.encrypt(rng, keys80, bits[:, 1])
    got = gctx80.gate(O.NAND, x, y)
    assert np.array_equal(O.decrypt(keys80, got), ~(bits[:, 0] & bits[:, 1]))
    ph = O.phase(keys80, got).astype(np.float64) / 2 ** 32
    assert np.abs(np.abs(ph) - 0.125).max() < 1 / 16
    pick = prng.choice(B, 16, replace=False)
    assert np.array_equal(got[pick], octx80.gate(O.NAND, x[pick], y[pick]))
    assert np.array_equal(gctx80.gate(O.NAND, x[1000:1037], y[1000:1037]), got[1000:1037])


def test_device_pointer_entry_points(keys80, gctx80):
    import torch
    rng = O.Rng(11)
    bits = np.random.default_rng(11).integers(0, 2, (5, 3)).astype(bool)
    cts = [O.encrypt(rng, keys80, bits[:, i]) for i in range(3)]
    d = [torch.from_numpy(c).cuda() for c in cts]
    out = torch.empty_like(d[0])
    s = torch.cuda.current_stream().cuda_stream
    for op, nargs in ((O.NAND, 2), (O.MUX, 3), (O.NOT, 1)):
        ptrs = [d[i].data_ptr() if i < nargs else 0 for i in range(3)]
        gctx80.gate_dev(op, ptrs[0], ptrs[1], ptrs[2], out.data_ptr(), 5, stream=s)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), gctx80.gate(op, *cts[:nargs]))


# ---- host mirror of the TFHE.jl API, GPU keygen included (test/runtests.jl through the mirrored names) ----
def test_api_mirror_truth_tables():
    rng = np.random.default_rng(123)
    sk, ck = T.make_key_pair(rng)
    gates = [(T.gate_nand, 2, lambda a, b: not (a and b)), (T.gate_xor, 2, lambda a, b: a != b),
             (T.gate_mux, 3, lambda a, b, c: b if a else c), (T.gate_not, 1, lambda a: not a),
             (T.gate_orny, 2, lambda a, b: (not a) or b)]
    for gate, nargs, ref in gates:
        for bits in itertools.product([False, True], repeat=nargs):
            ebits = [T.encrypt(rng, sk, b) for b in bits]
            assert T.decrypt(sk, gate(ck, *ebits)) == ref(*bits), (gate.__name__, bits)
    # batched variant: one call for the whole truth table
    tt = np.array(list(itertools.product([False, True], repeat=2)))
    out = T.gate_nand(ck, T.encrypt(rng, sk, tt[:, 0]), T.encrypt(rng, sk, tt[:, 1]))
    assert np.array_equal(T.decrypt(sk, out), ~(tt[:, 0] & tt[:, 1]))
    assert T.decrypt(sk, T.gate_constant(ck, True)) is True and T.decrypt(sk, T.gate_constant(ck, False)) is False
    # the GPU-generated bootstrap key is a valid key for the oracle too
    octx = O.Context(O.KeySet(O.PARAMS_80, sk.key, None, ck.bootstrap_key, ck.keyswitch_key))
    x, y = T.encrypt(rng, sk, tt[:, 0]), T.encrypt(rng, sk, tt[:, 1])
    assert np.array_equal(T.gate_and(ck, x, y).data, octx.gate(O.AND, x.data, y.data))


@pytest.mark.parametrize("flags", [_cabi.FLAG_SPLIT_FFT, _cabi.FLAG_UNSPLIT_FFT], ids=["split", "unsplit"])
@pytest.mark.parametrize("count", [3, 37, 38, 74, 75, 148, 149, 296, 297, 444, 445, 601, 700, 1024, 1185])
def test_every_batch_size_dispatch_path_equals_oracle(keys80_small, octx80_small, flags, count):
    """The library picks a kernel shape by batch size: up to one gate per two SMs (two-piece transform) a cluster of two
    CTAs per gate (37 MUX gates = 74 bootstraps is the last such batch, 74 NAND gates likewise), up to 2 gates per SM
    (1 with the one-piece transform) the latency kernel (one gate per CTA spread over 4 groups) + sliced key switch, above
    that up to four gates per CTA.  With a short LWE key (n = 24) the oracle can check EVERY ciphertext of every path; 445 / 601 leave the last
    CTA ragged; from 593 gates up the last wave of CTAs carries fewer gates per CTA (700: 148 x 4 + 108 x 1, 1 024:
    148 x 4 + 144 x 3, 1 185: two full waves + one gate; MUX doubles the bootstraps)."""
    P = keys80_small.params
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, flags=flags)
    ctx.load_bk(keys80_small.bk); ctx.load_ksk(keys80_small.ksk)
    bits = np.random.default_rng(count).integers(0, 2, (count, 3)).astype(bool)
    rng = O.Rng(count)
    x, y, z = (O.encrypt(rng, keys80_small, bits[:, i]) for i in range(3))
    assert np.array_equal(ctx.gate(O.NAND, x, y), octx80_small.gate(O.NAND, x, y))
    assert np.array_equal(ctx.gate(O.MUX, x, y, z), octx80_small.gate(O.MUX, x, y, z))


def test_cluster_kernel_can_be_disabled_and_agrees(keys80, octx80, gctx80, monkeypatch):
    """Full-size key (n = 500): the two-CTA cluster kernel (default for <= 74 bootstraps), the one-CTA latency kernel
    (TFHE_B200_CLUSTER=0) and the oracle give the same ciphertexts, for a plain gate and for MUX (two bootstraps per gate
    in one launch)."""
    rng = O.Rng(22)
    bits = np.random.default_rng(22).integers(0, 2, (20, 3)).astype(bool)
    x, y, z = (O.encrypt(rng, keys80, bits[:, i]) for i in range(3))
    got_nand, got_mux = gctx80.gate(O.NAND, x, y), gctx80.gate(O.MUX, x, y, z)
    monkeypatch.setenv("TFHE_B200_CLUSTER", "0")
    P = keys80.params
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit)
    ctx.load_bk(keys80.bk); ctx.load_ksk(keys80.ksk)
    assert np.array_equal(got_nand, ctx.gate(O.NAND, x, y))
    assert np.array_equal(got_mux, ctx.gate(O.MUX, x, y, z))
    assert np.array_equal(got_nand[:4], octx80.gate(O.NAND, x[:4], y[:4]))
    assert np.array_equal(got_mux[:2], octx80.gate(O.MUX, x[:2], y[:2], z[:2]))
    assert np.array_equal(O.decrypt(keys80, got_mux), np.where(bits[:, 0], bits[:, 1], bits[:, 2]))


def test_latency_path_can_be_disabled_and_agrees(keys80, gctx80, monkeypatch):
    """TFHE_B200_LOWLAT=0 routes small batches through the throughput kernels (1 and 2 gates per CTA) and the
    one-CTA-per-ciphertext key switch; same ciphertexts."""
    rng = O.Rng(21)
    bits = np.random.default_rng(21).integers(0, 2, (150, 2)).astype(bool)
    x, y = O.encrypt(rng, keys80, bits[:, 0]), O.encrypt(rng, keys80, bits[:, 1])
    want = gctx80.gate(O.XOR, x, y)
    monkeypatch.setenv("TFHE_B200_LOWLAT", "0")
    P = keys80.params
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit)
    ctx.load_bk(keys80.bk); ctx.load_ksk(keys80.ksk)
    assert np.array_equal(ctx.gate(O.XOR, x, y), want)            # 150 gates: 2 per CTA
    assert np.array_equal(ctx.gate(O.XOR, x[:3], y[:3]), want[:3])  # 3 gates: 1 per CTA


@pytest.mark.parametrize("count", [2560, 2601, 4096, 4161, 4737, 4800])
def test_tiled_keyswitch_equals_per_ciphertext_kernel(keys80, gctx80, monkeypatch, count):
    """Large batches use keyswitch_tile_kernel (table streamed through shared memory): tiles of 32 ciphertexts from 2 560
    up to one wave of them (4 736 on 148 SMs), tiles of 64 above; TFHE_B200_KS_TILE=0 selects the one-CTA-per-ciphertext
    kernel.  Random dimension-1024 inputs, ragged tiles."""
    rng = np.random.default_rng(count)
    u = rng.integers(-2 ** 31, 2 ** 31, (count, 1025), dtype=np.int64).astype(np.int32)
    got = gctx80.keyswitch(u)
    monkeypatch.setenv("TFHE_B200_KS_TILE", "0")
    P = keys80.params
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit)
    ctx.load_bk(keys80.bk); ctx.load_ksk(keys80.ksk)
    assert np.array_equal(got, ctx.keyswitch(u))
    octx = O.Context(keys80)
    assert np.array_equal(got[:8], octx.keyswitch(u[:8]))


def test_multi_device_context_shards_equal_single_device(keys80_small, octx80_small):
    """tfhe_b200_multi_*: the batch is cut into contiguous shards, one host thread per GPU, disjoint output slices.
    With one GPU the code path (threads, sharding, replicated key load) is the same; with two or more the result must
    still equal the single-device ciphertexts bit for bit."""
    ndev = T.device_count()
    P = keys80_small.params
    kw = dict(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit)
    single = T.Context(**kw)
    single.load_bk(keys80_small.bk); single.load_ksk(keys80_small.ksk)
    count = 1201
    bits = np.random.default_rng(9).integers(0, 2, (count, 3)).astype(bool)
    rng = O.Rng(9)
    x, y, z = (O.encrypt(rng, keys80_small, bits[:, i]) for i in range(3))
    want_nand, want_mux = single.gate(O.NAND, x, y), single.gate(O.MUX, x, y, z)
    assert np.array_equal(want_nand[:64], octx80_small.gate(O.NAND, x[:64], y[:64]))
    for devs in ([0], list(range(ndev)) if ndev > 1 else None):
        if devs is None and ndev == 1:
            devs = None   # "all visible devices"
        m = _cabi.MultiContext(devices=devs, **kw)
        assert m.devices == (len(devs) if devs else ndev)
        m.load_bk(keys80_small.bk); m.load_ksk(keys80_small.ksk)
        assert np.array_equal(m.gate(O.NAND, x, y), want_nand)
        assert np.array_equal(m.gate(O.MUX, x, y, z), want_mux)
        assert np.array_equal(m.gate(O.NAND, x[:3], y[:3]), want_nand[:3])      # fewer gates than devices x waves
        assert np.array_equal(m.bootstrap(x[:70]), single.bootstrap(x[:70]))
        assert m.kernel_launches > 0
        m.close()
    with pytest.raises(T.TFHEB200Error):
        _cabi.MultiContext(devices=[0, 0], **kw)
    with pytest.raises(T.TFHEB200Error):
        _cabi.MultiContext(devices=[ndev], **kw)


def test_cloud_key_over_all_devices_matches_reference_usage():
    """docs/src/manual.md:28-35 usage with CloudKey(devices="all"): one gate call, every GPU."""
    rng = np.random.default_rng(11)
    sk, ck = T.make_key_pair(rng, devices="all")
    bits = rng.integers(0, 2, (2, 700)).astype(bool)
    x, y = T.encrypt(rng, sk, bits[0]), T.encrypt(rng, sk, bits[1])
    assert np.array_equal(T.decrypt(sk, T.gate_xor(ck, x, y)), bits[0] ^ bits[1])
    assert ck.mctx.devices == T.device_count()


def test_pageable_and_unaligned_host_buffers(keys80, gctx80):
    """The C ABI accepts plain pageable host memory (a Julia Matrix{Int32}); a batch large enough for the chunked,
    double-buffered copy pipeline (4 chunks) must give the same ciphertexts as the device-resident path."""
    count = 8 * 4 * 148 + 5
    rng = O.Rng(12)
    bits = np.random.default_rng(12).integers(0, 2, (64, 2)).astype(bool)
    x = np.tile(O.encrypt(rng, keys80, bits[:, 0]), (count // 64 + 1, 1))[:count]
    y = np.tile(O.encrypt(rng, keys80, bits[:, 1]), (count // 64 + 1, 1))[:count]
    out = gctx80.gate(O.AND, x, y)
    assert np.array_equal(out, np.tile(out[:64], (count // 64 + 1, 1))[:count])
    assert np.array_equal(out[:8], O.Context(keys80).gate(O.AND, x[:8], y[:8]))


def test_device_resident_batch_is_walked_in_pieces(keys80_small, octx80_small, monkeypatch):
    """tfhe_b200_gate_batch_dev bounds its scratch (the extracted samples between blind rotation and key switch) by walking
    a large device-resident batch in pieces of whole CTA waves (2^20 gates by default).  With the piece size forced down
    to one wave, a ragged 1 500-gate NAND / MUX batch must give the ciphertexts of the one-pass run and of the oracle."""
    import torch
    P = keys80_small.params
    count = 1500
    bits = np.random.default_rng(31).integers(0, 2, (count, 3)).astype(bool)
    rng = O.Rng(31)
    cts = [O.encrypt(rng, keys80_small, bits[:, i]) for i in range(3)]
    d = [torch.from_numpy(c).cuda() for c in cts]
    s = torch.cuda.current_stream().cuda_stream
    outs = []
    for piece in (None, "1"):
        if piece:
            monkeypatch.setenv("TFHE_B200_DEV_PIECE", piece)     # clamped up to one wave of 4 x SM count gates
        ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit)
        ctx.load_bk(keys80_small.bk); ctx.load_ksk(keys80_small.ksk)
        res = []
        for op, nargs in ((O.NAND, 2), (O.MUX, 3)):
            out = torch.empty_like(d[0])
            ptrs = [d[i].data_ptr() if i < nargs else 0 for i in range(3)]
            ctx.gate_dev(op, ptrs[0], ptrs[1], ptrs[2], out.data_ptr(), count, stream=s)
            torch.cuda.synchronize()
            res.append(out.cpu().numpy())
        outs.append(res)
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[1][0], octx80_small.gate(O.NAND, cts[0], cts[1]))
    assert np.array_equal(outs[1][1][:200], octx80_small.gate(O.MUX, cts[0][:200], cts[1][:200], cts[2][:200]))


@pytest.fixture(scope="module")
def keys128_small():
    return O.keygen(O.small_params(O.PARAMS_128, 24), 654)


@pytest.mark.parametrize("flags", [_cabi.FLAG_SPLIT_FFT, _cabi.FLAG_UNSPLIT_FFT], ids=["split", "unsplit"])
@pytest.mark.parametrize("count", [3, 74, 75, 149, 444, 445, 601, 700])
def test_every_batch_size_dispatch_path_equals_oracle_128(keys128_small, flags, count):
    """The same walk through the dispatch for the 128-bit set (l = 3, Bg = 2^7, api.jl:55-69): cluster kernel with three
    groups per CTA, latency kernel with six groups (up to 3 gates per SM), the four-gates-per-CTA kernel with its
    output-stationary step for three digit polynomials per component, balanced last wave, and the key switch over rows padded
    to the same stride as n = 24 gives.  Every ciphertext against the oracle."""
    P = keys128_small.params
    octx = O.Context(keys128_small)
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, flags=flags)
    ctx.load_bk(keys128_small.bk); ctx.load_ksk(keys128_small.ksk)
    bits = np.random.default_rng(count).integers(0, 2, (count, 3)).astype(bool)
    rng = O.Rng(1000 + count)
    x, y, z = (O.encrypt(rng, keys128_small, bits[:, i]) for i in range(3))
    assert np.array_equal(ctx.gate(O.NAND, x, y), octx.gate(O.NAND, x, y))
    assert np.array_equal(ctx.gate(O.MUX, x, y, z), octx.gate(O.MUX, x, y, z))


@pytest.mark.parametrize("mu", [1 << 30, -(1 << 29), 123456789])
def test_bootstrap_with_arbitrary_test_vector_value(keys80_small, octx80_small, gctx80_small, mu):
    """bootstrap(bk, ks, mu, x) (bootstrap.jl:92-95) takes ANY mu; the gates only ever pass 1/8.  Both halves and the fused
    call against the oracle, also on a batch large enough for the four-gates-per-CTA kernel."""
    rng = O.Rng(77)
    x = O.encrypt(rng, keys80_small, np.random.default_rng(7).integers(0, 2, 600).astype(bool))
    u = gctx80_small.bootstrap_wo_ks(x, mu)
    assert np.array_equal(u, octx80_small.bootstrap_wo_ks(x, mu))
    assert np.array_equal(gctx80_small.bootstrap(x[:5], mu), octx80_small.bootstrap(x[:5], mu))
    assert np.array_equal(gctx80_small.bootstrap(x, mu), octx80_small.keyswitch(u))


def test_bad_gate_arguments_are_errors_not_crashes(keys80_small, gctx80_small):
    """Unknown opcode, a binary gate without its second operand, MUX without its third: TFHE_B200_EINVAL with a message."""
    n1 = keys80_small.params.n + 1
    x = np.zeros((2, n1), np.int32)
    for args in ((99, x, x, None), (O.NAND, x, None, None), (O.MUX, x, x, None), (O.NOT, None, None, None)):
        with pytest.raises(T.TFHEB200Error) as e:
            gctx80_small.gate(args[0], args[1], args[2], args[3], count=2)
        assert e.value.code == _cabi.EINVAL and str(e.value)
    # the context is still usable afterwards
    assert gctx80_small.gate(O.NOT, x).shape == (2, n1)
