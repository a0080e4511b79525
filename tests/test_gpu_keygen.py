"""GPU parity tests of the steps either side of the gate path: batched encrypt / decrypt
(api.jl:155-169, lwe.jl:38-59) and evaluation-key generation (bootstrap.jl:6-15, tgsw.jl:52-88, tlwe.jl:63-73,
keyswitch.jl:14-41) on the device.  The *_words entry points take their randomness from the caller, so the CPU side
(exact integer arithmetic, oracle/) is fed the SAME words and the results must agree bit for bit; the seeded entry
points are checked against the numpy restatement of the device generator and end to end through the gate path."""
import numpy as np
import pytest

import tfhe_jl_b200 as T
from conftest import random_torus
from oracle import oracle as O
from philox_ref import words

pytestmark = pytest.mark.gpu
N = 1024


def wrap(x):
    return ((np.asarray(x, dtype=np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int32)


def make_ctx(P, **kw):
    return T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, **kw)


def test_device_random_words_equal_numpy_philox():
    ctx = make_ctx(O.PARAMS_80)
    for seed, stream, count in ((0, 0, 9), (123, 1, 4099), (2 ** 63 + 5, 2 ** 40 + 7, 1000)):
        assert np.array_equal(ctx.random_words(seed, stream, count), words(seed, stream, count))


@pytest.mark.parametrize("key_len", [500, 1000, 37])
def test_lwe_encrypt_words_and_phase_equal_integer_arithmetic(key_len):
    """lwe.jl:38-59 with explicit randomness, single-key (n) and MK (p*n) key lengths, odd lengths, full-range keys."""
    ctx = make_ctx(O.PARAMS_80)
    rng = np.random.default_rng(key_len)
    count = 301
    key = random_torus(rng, key_len) if key_len == 37 else rng.integers(0, 2, key_len).astype(np.int32)
    a, mu, noise = random_torus(rng, count, key_len), random_torus(rng, count), random_torus(rng, count)
    ct = ctx.lwe_encrypt_words(key, mu, noise, a)
    dot = wrap((a.astype(np.int64) * key.astype(np.int64)).sum(axis=1))
    assert np.array_equal(ct[:, :-1], a)
    assert np.array_equal(ct[:, -1], wrap(mu.astype(np.int64) + noise + dot))
    assert np.array_equal(ctx.lwe_phase(key, ct), wrap(mu.astype(np.int64) + noise))
    if key_len == 500:   # the oracle's lwe_phase on the same ciphertexts
        ks = O.KeySet(O.PARAMS_80, key, None, None, None)
        assert np.array_equal(ctx.lwe_phase(key, ct), O.phase(ks, ct))
        assert np.array_equal(ctx.decrypt(key, ct), O.phase(ks, ct) > 0)


def test_seeded_encrypt_uses_the_documented_streams_and_decrypts(keys80):
    ctx = make_ctx(O.PARAMS_80)
    P = keys80.params
    bits = np.random.default_rng(3).integers(0, 2, 5000).astype(bool)
    ct = ctx.encrypt(keys80.lwe_key, bits, P.lwe_sigma, seed=99)
    assert np.array_equal(ct[:, :-1].reshape(-1), words(99, 1, bits.size * P.n))             # mask = stream 1 of the seed
    noise = O.phase(keys80, ct).astype(np.int64) - np.where(bits, 2 ** 29, -2 ** 29)          # api.jl:155-158
    assert abs(noise.mean()) < 4 * P.lwe_sigma * 2 ** 32 / np.sqrt(bits.size)
    assert 0.93 < noise.std() / (P.lwe_sigma * 2 ** 32) < 1.07
    assert np.array_equal(ctx.decrypt(keys80.lwe_key, ct), bits) and np.array_equal(O.decrypt(keys80, ct), bits)
    assert np.array_equal(ct, ctx.encrypt(keys80.lwe_key, bits, P.lwe_sigma, seed=99))        # same seed, same ciphertexts
    assert not np.array_equal(ct, ctx.encrypt(keys80.lwe_key, bits, P.lwe_sigma, seed=100))


def expected_bk(P, lwe_key, tlwe_key, a, noise):
    """tgsw_encrypt (tgsw.jl:84-88) with the exact integer negacyclic product as ground truth."""
    S = P.n * P.l * 2
    bk = np.empty((S, 2, N), dtype=np.int32)
    for s in range(S):
        i, r, j = s // (2 * P.l), (s // 2) % P.l, s % 2
        bk[s, 0] = a[s]
        bk[s, 1] = wrap(noise[s].astype(np.int64) + O.polymul(tlwe_key, a[s], O.ROUTE_EXACT))   # tlwe.jl:63-73
        bk[s, j, 0] = wrap(int(bk[s, j, 0]) + (int(lwe_key[i]) << (32 - (r + 1) * P.bgbit)))     # tgsw.jl:62-69
    return bk.reshape(P.n, P.l, 2, 2, N)


def expected_ksk(P, out_key, in_key, a, noise):
    base1 = (1 << P.basebit) - 1
    h = np.arange(1, base1 + 1, dtype=np.int64)[None, None, :]
    shift = (32 - np.arange(1, P.t + 1, dtype=np.int64) * P.basebit)[None, :, None]
    msg = wrap((in_key.astype(np.int64)[:, None, None] * h) << shift)                            # keyswitch.jl:35
    b = wrap(msg.astype(np.int64) + noise + wrap((a.astype(np.int64) * out_key.astype(np.int64)).sum(axis=-1)))
    return np.concatenate([a, b[..., None]], axis=-1)


@pytest.mark.parametrize("base", [O.PARAMS_80, O.PARAMS_128], ids=["80bit", "128bit"])
def test_keygen_words_equal_exact_arithmetic_and_the_loaded_key_evaluates_gates(base):
    P = O.small_params(base, 6)
    rng = np.random.default_rng(P.l)
    lwe_key = rng.integers(0, 2, P.n).astype(np.int32)
    tlwe_key = rng.integers(0, 2, N).astype(np.int32)
    S = P.n * P.l * 2
    a, noise = random_torus(rng, S, N), wrap(rng.normal(0, 2 ** 32 * base.bs_sigma, (S, N)).round())
    base1 = (1 << P.basebit) - 1
    ka, kn = random_torus(rng, N, P.t, base1, P.n), wrap(rng.normal(0, 2 ** 32 * base.ks_sigma, (N, P.t, base1)).round())
    ctx = make_ctx(P)
    bk = ctx.keygen_bk(lwe_key, tlwe_key, a=a, noise=noise)
    ksk = ctx.keygen_ksk(lwe_key, tlwe_key, a=ka, noise=kn)
    assert np.array_equal(bk, expected_bk(P, lwe_key, tlwe_key, a, noise))
    assert np.array_equal(ksk, expected_ksk(P, lwe_key, tlwe_key, ka, kn))
    # the key the device assembled, transformed and loaded itself is the key the oracle builds from the same words
    keys = O.KeySet(P, lwe_key, tlwe_key[None, :], bk, ksk)
    bits = np.array([[0, 0, 1], [0, 1, 0], [1, 0, 1], [1, 1, 0]], dtype=bool)
    orng = O.Rng(5)
    x, y, z = (O.encrypt(orng, keys, bits[:, i]) for i in range(3))
    octx = O.Context(keys)
    for op, args in ((O.NAND, (x, y)), (O.XOR, (x, y)), (O.MUX, (x, y, z))):
        assert np.array_equal(ctx.gate(op, *args), octx.gate(op, *args))
    assert np.array_equal(O.decrypt(keys, ctx.gate(O.NAND, x, y)), ~(bits[:, 0] & bits[:, 1]))


def test_seeded_keygen_full_size_truth_table_and_oracle_identity():
    """make_key_pair + encrypt + gates + decrypt with every random word generated on the device (api.jl:116-169):
    the key works, the oracle given the same key produces the same ciphertexts, and a seed reproduces the key."""
    P = O.PARAMS_80
    rng = np.random.default_rng(8)
    lwe_key = rng.integers(0, 2, P.n).astype(np.int32)
    tlwe_key = rng.integers(0, 2, N).astype(np.int32)
    ctx = make_ctx(P)
    bk = ctx.keygen_bk(lwe_key, tlwe_key, sigma=P.bs_sigma, seed=1234)
    ksk = ctx.keygen_ksk(lwe_key, tlwe_key, sigma=P.ks_sigma, seed=1234)
    # the mask polynomials are stream 1 of the seed, in sample order
    assert np.array_equal(bk[0, 0, 0, 0] - np.eye(1, N, 0, dtype=np.int64)[0] * (int(lwe_key[0]) << 22), words(1234, 1, N))
    tt = np.array([[a, b] for a in (0, 1) for b in (0, 1)] * 16, dtype=bool)
    x = ctx.encrypt(lwe_key, tt[:, 0], P.lwe_sigma, seed=1)
    y = ctx.encrypt(lwe_key, tt[:, 1], P.lwe_sigma, seed=2)
    out = ctx.gate(O.NAND, x, y)
    assert np.array_equal(ctx.decrypt(lwe_key, out), ~(tt[:, 0] & tt[:, 1]))
    keys = O.KeySet(P, lwe_key, tlwe_key[None, :], bk, ksk)
    assert np.array_equal(out[:8], O.Context(keys).gate(O.NAND, x[:8], y[:8]))
    noise = O.phase(keys, out).astype(np.int64) - np.where(~(tt[:, 0] & tt[:, 1]), 2 ** 29, -2 ** 29)
    assert np.abs(noise).max() < 2 ** 28                                                       # gates.jl:1-6: noise < 1/16
    ctx2 = make_ctx(P)
    assert np.array_equal(ctx2.keygen_bk(lwe_key, tlwe_key, sigma=P.bs_sigma, seed=1234), bk)
    assert np.array_equal(ctx2.keygen_ksk(lwe_key, tlwe_key, sigma=P.ks_sigma, seed=1234), ksk)
    # keyswitch.jl:28-29: the key-switching noises are centred
    kn = O.phase(keys, ksk.reshape(-1, P.n + 1)).astype(np.int64)
    base1 = (1 << P.basebit) - 1
    h = np.arange(1, base1 + 1, dtype=np.int64)[None, None, :]
    shift = (32 - np.arange(1, P.t + 1, dtype=np.int64) * P.basebit)[None, :, None]
    msg = wrap((tlwe_key.astype(np.int64)[:, None, None] * h) << shift).reshape(-1)
    kn = wrap(kn - msg).astype(np.int64)
    assert abs(kn.mean()) < 1.0 and 0.9 < kn.std() / (P.ks_sigma * 2 ** 32) < 1.1


def test_encrypt_on_device_feeds_the_gate_path_without_host_ciphertexts(keys80, ):
    import torch
    P = keys80.params
    ctx = make_ctx(P)
    ctx.load_bk(keys80.bk); ctx.load_ksk(keys80.ksk)
    count = 4 * 148 * 2 + 3
    bits = np.random.default_rng(4).integers(0, 2, (2, count)).astype(bool)
    dx = torch.empty((count, P.n + 1), dtype=torch.int32, device="cuda"); dy = torch.empty_like(dx); do = torch.empty_like(dx)
    ctx.encrypt_dev(keys80.lwe_key, bits[0], P.lwe_sigma, 11, dx.data_ptr())
    ctx.encrypt_dev(keys80.lwe_key, bits[1], P.lwe_sigma, 12, dy.data_ptr())
    ctx.gate_dev(O.NAND, dx.data_ptr(), dy.data_ptr(), 0, do.data_ptr(), count)
    torch.cuda.synchronize()
    assert np.array_equal(ctx.decrypt(keys80.lwe_key, do.cpu().numpy()), ~(bits[0] & bits[1]))
    assert np.array_equal(dx.cpu().numpy(), ctx.encrypt(keys80.lwe_key, bits[0], P.lwe_sigma, 11))


def uni_encryptions(rng, P, p):
    """Uni-encryption-shaped random material for p parties (the expansion is exact integer arithmetic on whatever it is
    given): uni_enc [p][6][n][l][N] in the order c0, c1, d0, d1, f0, f1, public_b [p][l][N]."""
    return random_torus(rng, p, 6, P.n, P.l, N), random_torus(rng, p, P.l, N)


def expand_reference(P, p, uni_enc, public_b):
    """RGSW.Expand (mk_internals.jl:304-345) with the exact O(N^2) negacyclic product as ground truth."""
    n, l = P.n, P.l
    bk = np.zeros((p, n, l * (2 * p + 2), N), dtype=np.int32)
    for i in range(p):
        c0, c1, d0, d1, f0, f1 = uni_enc[i]
        for j in range(n):
            for jj in range(l):
                for ii in range(p):
                    xi, yi = jj * p + ii, l * p + jj * p + ii
                    if ii == i:
                        bk[i, j, xi], bk[i, j, yi] = d0[j, jj], d1[j, jj]
                        continue
                    u = O.decompose(wrap(public_b[ii, jj].astype(np.int64) - public_b[i, jj]), l, P.bgbit)   # :321
                    x = d0[j, jj].astype(np.int64); y = np.zeros(N, dtype=np.int64)
                    for r in range(l):
                        x += O.polymul(u[r], f0[j, r], O.ROUTE_EXACT)                                        # :330
                        y += O.polymul(u[r], f1[j, r], O.ROUTE_EXACT)                                        # :338
                    bk[i, j, xi], bk[i, j, yi] = wrap(x), wrap(y)
            bk[i, j, 2 * l * p:2 * l * p + l] = c0[j]
            bk[i, j, 2 * l * p + l:] = c1[j]
    return bk


@pytest.mark.parametrize("p", [2, 4, 8])
def test_mk_key_expansion_on_device_equals_exact_arithmetic(p):
    P = O.small_params(O.MK_PARAMS[p], 2)
    rng = np.random.default_rng(p)
    uni_enc, public_b = uni_encryptions(rng, P, p)
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, parties=p)
    got = ctx.mk_expand_load_bk(uni_enc, public_b)
    assert np.array_equal(got, expand_reference(P, p, uni_enc, public_b))


def test_mk_cloud_key_device_expansion_equals_host_expansion_and_evaluates():
    """examples/multikey.jl: the key expanded on the device is the key the round-1 host path builds, and it works."""
    rng = np.random.default_rng(5)
    params = T.mktfhe_parameters_2party
    sks = [T.SecretKey(rng, params) for _ in range(2)]
    shared = T.SharedKey(rng, params)
    parts = [T.CloudKeyPart(rng, sk, shared) for sk in sks]
    ck = T.MKCloudKey(parts)
    ck_host = T.MKCloudKey(parts, host_expand=True)
    assert np.array_equal(ck.bootstrap_key, ck_host.bootstrap_key)
    m1 = rng.integers(0, 2, 12).astype(bool); m2 = rng.integers(0, 2, 12).astype(bool)
    e1, e2 = T.mk_encrypt(rng, sks, m1), T.mk_encrypt(rng, sks, m2)
    out = T.mk_gate_nand(ck, e1, e2)
    assert np.array_equal(out.data, T.mk_gate_nand(ck_host, e1, e2).data)
    # the reference's 2-party parameters leave sigma ~ 0.05 of output noise: an occasional gate decrypts
    # wrongly in TFHE.jl too, so require most, not all (ciphertext identity with the host-expanded key is asserted above)
    assert np.sum(T.mk_decrypt(sks, out) == ~(m1 & m2)) >= 10
