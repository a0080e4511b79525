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


def test_mk_keyswitch_tiled_equals_per_ciphertext_kernel(monkeypatch):
    """mk_keyswitch (mk_internals.jl:397-411) on a batch large enough for the tiled kernel (one launch per party,
    the joint b accumulated with integer atomics) against the one-CTA-per-ciphertext kernel and the oracle."""
    p = 2
    mk = O.mk_keygen(O.small_params(O.MK_PARAMS[p], 3), p, 90)
    count = 4100
    u = np.random.default_rng(5).integers(-2 ** 31, 2 ** 31, (count, p * N + 1), dtype=np.int64).astype(np.int32)
    got = make_mk_ctx(mk).keyswitch(u)
    monkeypatch.setenv("TFHE_B200_KS_TILE", "0")
    assert np.array_equal(got, make_mk_ctx(mk).keyswitch(u))
    assert np.array_equal(got[-4:], O.MKContext(mk).keyswitch(u[-4:]))
