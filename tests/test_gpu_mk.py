"""GPU parity tests of the multi-key path (mk_internals.jl) through the C ABI."""
import numpy as np
import pytest

import tfhe_jl_b200 as T
from tfhe_jl_b200 import _cabi
from conftest import random_torus
from oracle import oracle as O

pytestmark = pytest.mark.gpu
N = 1024


def make_mk_ctx(mk, flags=_cabi.FLAG_SPLIT_FFT):
    P = mk.params
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, parties=mk.parties, flags=flags)
    ctx.load_bk(mk.bk)
    ctx.load_ksk(mk.ksk)
    return ctx


@pytest.mark.parametrize("p", [2, 4, 8])
def test_mk_extern_product_matches_exact_oracle(p):
    mk = O.mk_keygen(O.small_params(O.MK_PARAMS[p], 4), p, 40 + p)
    octx = O.MKContext(mk)
    ctx = make_mk_ctx(mk)
    rng = np.random.default_rng(p)
    cnt = 2 * p
    acc = random_torus(rng, cnt, p + 1, N)
    acc[0] = 0; acc[1] = -(2 ** 31)
    party = (np.arange(cnt) % p).astype(np.int32)
    idx = rng.integers(0, 4, cnt).astype(np.int32)
    got = ctx.extern_product(acc, idx, party)
    for g in range(cnt):
        assert np.array_equal(got[g], octx.extern_mul(int(party[g]), int(idx[g]), acc[g], O.ROUTE_EXACT)), g


def test_mk_nand_2party_full(mkkeys2, mkctx2):
    """test/runtests.jl:60-100 on the GPU, ciphertext-identical to the oracle"""
    ctx = make_mk_ctx(mkkeys2)
    bits = np.random.default_rng(1).integers(0, 2, (10, 2)).astype(bool)
    rng = O.Rng(2)
    x, y = O.mk_encrypt(rng, mkkeys2, bits[:, 0]), O.mk_encrypt(rng, mkkeys2, bits[:, 1])
    got = ctx.mk_nand(x, y)
    assert np.array_equal(got, mkctx2.nand(x, y))
    assert np.array_equal(O.mk_decrypt(mkkeys2, got), ~(bits[:, 0] & bits[:, 1]))
    u = ctx.bootstrap_wo_ks(x[:2])
    assert np.array_equal(u, mkctx2.bootstrap_wo_ks(x[:2]))
    assert np.array_equal(ctx.keyswitch(u), mkctx2.keyswitch(u))


def test_mk_unsplit_equals_split(mkkeys2):
    a, b = make_mk_ctx(mkkeys2), make_mk_ctx(mkkeys2, _cabi.FLAG_UNSPLIT_FFT)
    rng = O.Rng(3)
    x, y = O.mk_encrypt(rng, mkkeys2, [True, False]), O.mk_encrypt(rng, mkkeys2, [True, True])
    assert np.array_equal(a.mk_nand(x, y), b.mk_nand(x, y))


@pytest.mark.parametrize("p", [4, 8])
def test_mk_nand_small_key_4_and_8_parties(p):
    """The 4- and 8-party sets (mk_api.jl:16-34) are never exercised by the reference's tests; a short
    LWE key keeps the oracle fast while covering every code path."""
    mk = O.mk_keygen(O.small_params(O.MK_PARAMS[p], 6), p, 50 + p)
    octx = O.MKContext(mk)
    ctx = make_mk_ctx(mk)
    rng = O.Rng(p)
    x, y = O.mk_encrypt(rng, mk, [True, False, True]), O.mk_encrypt(rng, mk, [True, True, False])
    assert np.array_equal(ctx.mk_nand(x, y), octx.nand(x, y))


def test_mk_api_mirror_2party():
    """examples/multikey.jl through the mirrored API names, keys generated with GPU polymuls."""
    rng = np.random.default_rng(5)
    params = T.mktfhe_parameters_2party
    sks = [T.SecretKey(rng, params) for _ in range(2)]
    shared = T.SharedKey(rng, params)
    parts = [T.CloudKeyPart(rng, sk, shared) for sk in sks]
    ck = T.MKCloudKey(parts)
    m1 = rng.integers(0, 2, 10).astype(bool); m2 = rng.integers(0, 2, 10).astype(bool)
    e1, e2 = T.mk_encrypt(rng, sks, m1), T.mk_encrypt(rng, sks, m2)
    assert np.array_equal(T.mk_decrypt(sks, e1), m1)
    out = T.mk_gate_nand(ck, e1, e2)
    assert np.array_equal(T.mk_decrypt(sks, out), ~(m1 & m2))


@pytest.mark.parametrize("p,count", [(2, 601), (4, 310), (8, 9)])
def test_mk_nand_large_batch_ring_configurations(p, count, monkeypatch):
    """Batches above 2 x SM count take the 4-gates-per-CTA ring kernel (6-stage ring for 2 parties, 3-stage
    for 4); ragged counts leave groups without a gate in the last CTA.  Every ciphertext must equal the
    oracle's, and the one-gate-per-CTA kernel (TFHE_B200_MK_RING=0) must agree too."""
    mk = O.mk_keygen(O.small_params(O.MK_PARAMS[p], 3), p, 70 + p)
    octx = O.MKContext(mk)
    bits = np.random.default_rng(p).integers(0, 2, (count, 2)).astype(bool)
    rng = O.Rng(p)
    x, y = O.mk_encrypt(rng, mk, bits[:, 0]), O.mk_encrypt(rng, mk, bits[:, 1])
    want = octx.nand(x, y)
    assert np.array_equal(make_mk_ctx(mk).mk_nand(x, y), want)
    assert np.array_equal(make_mk_ctx(mk, _cabi.FLAG_UNSPLIT_FFT).mk_nand(x, y), want)
    monkeypatch.setenv("TFHE_B200_MK_RING", "0")
    assert np.array_equal(make_mk_ctx(mk).mk_nand(x, y), want)


def test_mk_latency_kernel_agrees_with_ring_kernel(mkkeys2, mkctx2, monkeypatch):
    """Small 2-party batches take the latency kernel (one gate per CTA, 12 digit polynomials over 6 groups);
    with TFHE_B200_LOWLAT=0 the same batch goes through the ring kernel (2 gates per CTA).  Same ciphertexts,
    equal to the oracle's."""
    rng = O.Rng(31)
    bits = np.random.default_rng(31).integers(0, 2, (5, 2)).astype(bool)
    x, y = O.mk_encrypt(rng, mkkeys2, bits[:, 0]), O.mk_encrypt(rng, mkkeys2, bits[:, 1])
    want = mkctx2.nand(x, y)
    for flags in (_cabi.FLAG_SPLIT_FFT, _cabi.FLAG_UNSPLIT_FFT):
        assert np.array_equal(make_mk_ctx(mkkeys2, flags).mk_nand(x, y), want)
    monkeypatch.setenv("TFHE_B200_LOWLAT", "0")
    assert np.array_equal(make_mk_ctx(mkkeys2).mk_nand(x, y), want)


def random_mk_keys(p, n, seed):
    """Key-shaped random material (uniform torus words, binary LWE keys).  Ciphertext parity does not need a valid key:
    the arithmetic is the same, and every operand reaches full magnitude.  Saves the oracle's 53 s 8-party keygen."""
    P = O.small_params(O.MK_PARAMS[p], n)
    rng = np.random.default_rng(seed)
    bk = rng.integers(-2 ** 31, 2 ** 31, (p, P.n, P.l * (2 * p + 2), N), dtype=np.int64).astype(np.int32)
    ksk = rng.integers(-2 ** 31, 2 ** 31, (p,) + P.ksk_shape, dtype=np.int64).astype(np.int32)
    lwe = rng.integers(0, 2, (p, P.n)).astype(np.int32)
    return O.MKKeySet(P, p, lwe, bk, ksk, rng.integers(0, 2, (p, N)).astype(np.int32))


def sm_count():
    import torch
    return torch.cuda.get_device_properties(0).multi_processor_count


@pytest.mark.parametrize("count", [4100, 4800], ids=["tiles_of_32", "tiles_of_64"])
def test_mk_keyswitch_tiled_full_size_equals_per_ciphertext_kernel_and_oracle(monkeypatch, count):
    """mk_keyswitch (mk_internals.jl:397-411) at n = 500 (table stride 512), so that batches of 4 096 and more really
    take keyswitch_tile_kernel in its MK mode (one launch per party, non-zero in/out offsets, the joint b accumulated
    with integer atomics).  Compared with the one-CTA-per-ciphertext kernel (TFHE_B200_KS_TILE=0) on every row and
    with the oracle on the first and the last (ragged tile) rows."""
    p = 2
    mk = random_mk_keys(p, 500, 90)
    u = np.random.default_rng(5).integers(-2 ** 31, 2 ** 31, (count, p * N + 1), dtype=np.int64).astype(np.int32)
    got = make_mk_ctx(mk).keyswitch(u)
    monkeypatch.setenv("TFHE_B200_KS_TILE", "0")
    assert np.array_equal(got, make_mk_ctx(mk).keyswitch(u))
    rows = np.r_[0:4, count - 4:count]
    assert np.array_equal(got[rows], O.MKContext(mk).keyswitch(u[rows]))


def test_keyswitch_tile_stride_640(keys128, monkeypatch):
    """The 128-bit set (n = 630) pads its table rows to 640 words: the second instantiation of the tile kernel."""
    P = keys128.params
    count = 4099
    u = np.random.default_rng(6).integers(-2 ** 31, 2 ** 31, (count, N + 1), dtype=np.int64).astype(np.int32)
    def ctx():
        c = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit)
        c.load_ksk(keys128.ksk)
        return c
    got = ctx().keyswitch(u)
    monkeypatch.setenv("TFHE_B200_KS_TILE", "0")
    assert np.array_equal(got, ctx().keyswitch(u))
    rows = np.r_[0:4, count - 4:count]
    assert np.array_equal(got[rows], O.Context(keys128).keyswitch(u[rows]))


@pytest.mark.parametrize("p,count,nsample", [(2, None, 16), (4, None, 6), (8, 5, 5)])
def test_mk_nand_full_size_ring_kernels_equal_oracle(p, count, nsample, monkeypatch):
    """Full-size (n = 500) parity of the large-batch MK kernels (mk_internals.jl:348-515), judged on ciphertext
    equality with the oracle, never on decryption (the reference's own 2-party output noise, sigma ~ 0.05, makes an
    occasional gate decrypt wrongly in the reference too).

      2 parties: 4 x SMs + 1 gates -> mk_blind_rotate_ring_kernel with G = 4 (6-stage ring), last CTA ragged;
      4 parties: 2 x SMs + 2 gates -> G = 4 with the 3-stage ring;  8 parties: 5 gates (G = 2).
    A sample that includes the first and the ragged last gates is compared bit for bit with the oracle; ALL gates
    are compared with the one-gate-per-CTA kernel (TFHE_B200_MK_RING=0), so a wrong ciphertext anywhere shows.
    2 and 4 parties use real oracle keys (and report how many gates decrypt to NAND), 8 parties key-shaped random
    material (the oracle's 8-party keygen takes a minute)."""
    sms = sm_count()
    if count is None:
        count = 4 * sms + 1 if p == 2 else 2 * sms + 2
    real = p < 8
    mk = O.mk_keygen(O.MK_PARAMS[p], p, 300 + p) if real else random_mk_keys(p, 500, 300 + p)
    rng = np.random.default_rng(p)
    if real:
        bits = rng.integers(0, 2, (count, 2)).astype(bool)
        orng = O.Rng(p)
        x, y = O.mk_encrypt(orng, mk, bits[:, 0]), O.mk_encrypt(orng, mk, bits[:, 1])
    else:
        x = rng.integers(-2 ** 31, 2 ** 31, (count, p * 500 + 1), dtype=np.int64).astype(np.int32)
        y = rng.integers(-2 ** 31, 2 ** 31, (count, p * 500 + 1), dtype=np.int64).astype(np.int32)
    got = make_mk_ctx(mk).mk_nand(x, y)
    idx = np.unique(np.r_[0:nsample // 2, count - (nsample - nsample // 2):count])
    want = O.MKContext(mk).nand(x[idx], y[idx])
    assert np.array_equal(got[idx], want), f"{p}-party ring kernel differs from the oracle"
    monkeypatch.setenv("TFHE_B200_MK_RING", "0")
    monkeypatch.setenv("TFHE_B200_LOWLAT", "0")
    assert np.array_equal(got, make_mk_ctx(mk).mk_nand(x, y)), "ring kernel differs from the one-gate-per-CTA kernel"
    if real:
        ok = int(np.sum(O.mk_decrypt(mk, got) == ~(bits[:, 0] & bits[:, 1])))
        print(f"{p} parties: {ok}/{count} gates decrypt to NAND (oracle sample identical)")
        assert ok >= 0.9 * count


@pytest.mark.parametrize("mu", [1 << 29, 1 << 30, -(1 << 29), 123456789])
def test_mk_bootstrap_with_arbitrary_test_vector_value(mu):
    """mk_bootstrap(bk, ks, mu, x) (mk_internals.jl:512-515) = mk_keyswitch(mk_bootstrap_wo_keyswitch(mu, x)): the seam
    function takes ANY mu, mk_gate_nand only ever passes 1/8.  tfhe_b200_mk_bootstrap_batch and its two halves against the
    oracle, 2 parties, short LWE key."""
    mk = O.mk_keygen(O.small_params(O.MK_PARAMS[2], 6), 2, 61)
    octx = O.MKContext(mk)
    ctx = make_mk_ctx(mk)
    x = O.mk_encrypt(O.Rng(4), mk, [True, False, True])
    u = ctx.bootstrap_wo_ks(x, mu)
    assert np.array_equal(u, octx.bootstrap_wo_ks(x, mu))
    assert np.array_equal(ctx.mk_bootstrap(x, mu), octx.keyswitch(u))


def test_multi_device_context_mk_equals_single_device():
    """tfhe_b200_multi_mk_load_bk / _ksk / _mk_nand_batch / _bootstrap_batch on an MK context: the sharded result (one
    shard per visible GPU; with one GPU the same threads-and-slices code path) equals the single-device ciphertexts and
    the oracle's."""
    mk = O.mk_keygen(O.small_params(O.MK_PARAMS[2], 6), 2, 62)
    octx = O.MKContext(mk)
    P = mk.params
    count = 301
    bits = np.random.default_rng(62).integers(0, 2, (count, 2)).astype(bool)
    rng = O.Rng(62)
    x, y = O.mk_encrypt(rng, mk, bits[:, 0]), O.mk_encrypt(rng, mk, bits[:, 1])
    want = make_mk_ctx(mk).mk_nand(x, y)
    assert np.array_equal(want[:16], octx.nand(x[:16], y[:16]))
    m = _cabi.MultiContext(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, parties=2, devices=None)
    m.load_bk(mk.bk); m.load_ksk(mk.ksk)
    assert np.array_equal(m.mk_nand(x, y), want)
    assert np.array_equal(m.mk_nand(x[:1], y[:1]), want[:1])
    mu = 1 << 30
    assert np.array_equal(m.bootstrap(x[:9], mu), octx.keyswitch(octx.bootstrap_wo_ks(x[:9], mu)))
    m.close()


def test_mk_device_pointer_entry_points():
    """tfhe_b200_mk_nand_batch_dev / tfhe_b200_mk_bootstrap_batch_dev on device-resident ciphertexts and the caller's
    stream give the ciphertexts of the host-buffer calls."""
    import torch
    mk = O.mk_keygen(O.small_params(O.MK_PARAMS[2], 6), 2, 63)
    ctx = make_mk_ctx(mk)
    x, y = O.mk_encrypt(O.Rng(5), mk, [True, False, True, True, False]), O.mk_encrypt(O.Rng(6), mk, [True, True, False, False, False])
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    out = torch.empty_like(dx)
    s = torch.cuda.current_stream().cuda_stream
    ctx.mk_nand_dev(dx.data_ptr(), dy.data_ptr(), out.data_ptr(), 5, stream=s)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), ctx.mk_nand(x, y))
    mu = -(1 << 29)
    ctx.mk_bootstrap_dev(dx.data_ptr(), out.data_ptr(), 5, mu=mu, stream=s)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), ctx.mk_bootstrap(x, mu))


def test_mk_device_resident_batch_is_walked_in_pieces(monkeypatch):
    """tfhe_b200_mk_nand_batch_dev / _mk_bootstrap_batch_dev bound their scratch like the single-key call: with the piece
    size forced down to one wave of CTAs a ragged 700-gate batch gives the ciphertexts of the one-pass run."""
    import torch
    mk = O.mk_keygen(O.small_params(O.MK_PARAMS[2], 3), 2, 64)
    count = 700
    bits = np.random.default_rng(64).integers(0, 2, (count, 2)).astype(bool)
    x, y = O.mk_encrypt(O.Rng(7), mk, bits[:, 0]), O.mk_encrypt(O.Rng(8), mk, bits[:, 1])
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    s = torch.cuda.current_stream().cuda_stream
    res = []
    for piece in (None, "1"):
        if piece:
            monkeypatch.setenv("TFHE_B200_DEV_PIECE", piece)
        ctx = make_mk_ctx(mk)
        o1, o2 = torch.empty_like(dx), torch.empty_like(dx)
        ctx.mk_nand_dev(dx.data_ptr(), dy.data_ptr(), o1.data_ptr(), count, stream=s)
        ctx.mk_bootstrap_dev(dx.data_ptr(), o2.data_ptr(), count, mu=1 << 30, stream=s)
        torch.cuda.synchronize()
        res.append((o1.cpu().numpy(), o2.cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(res[1][0][:32], O.MKContext(mk).nand(x[:32], y[:32]))
