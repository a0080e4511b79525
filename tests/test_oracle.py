"""CPU tests that PIN THE ORACLE: against exact integer arithmetic, and against everything the reference's
own tests hold for this path — decrypted truth tables (test/runtests.jl:8-57), the MK NAND trials
(test/runtests.jl:60-100) and the tutorial's answer 42 (examples/tutorial.jl:77).  The reference ships no
golden ciphertexts, so ciphertext-level parity with TFHE.jl itself stays unpinned."""
import itertools

import numpy as np
import pytest

from conftest import PLAIN_GATES, random_torus
from oracle import oracle as O

N = 1024


def exact_negacyclic(x, y):
    """Independent ground truth in Python big integers (object arrays)."""
    full = np.convolve(x.astype(object), y.astype(object))
    out = [int(full[i]) - (int(full[i + N]) if i + N < len(full) else 0) for i in range(N)]
    return np.array([((v + 2 ** 31) % 2 ** 32) - 2 ** 31 for v in out], dtype=np.int64).astype(np.int32)


def test_polymul_exact_matches_bigint():
    rng = np.random.default_rng(0)
    x, y = random_torus(rng, N), random_torus(rng, N)
    assert np.array_equal(O.polymul(x, y, O.ROUTE_EXACT), exact_negacyclic(x, y))


@pytest.mark.parametrize("bits", [1, 7, 10, 11])
def test_reference_fft_route_equals_exact(bits):
    """polynomials.jl:138-140: integer operand up to 11 bits against a full 32-bit torus operand."""
    rng = np.random.default_rng(bits)
    for _ in range(8):
        x = rng.integers(-(1 << (bits - 1)), 1 << (bits - 1), N).astype(np.int32) if bits > 1 else rng.integers(0, 2, N).astype(np.int32)
        y = random_torus(rng, N)
        assert np.array_equal(O.polymul(x, y, O.ROUTE_FFT), O.polymul(x, y, O.ROUTE_EXACT))
    # adversarial: all operands at maximal magnitude with random signs
    x = (rng.integers(0, 2, N) * 2 - 1).astype(np.int32) * ((1 << (bits - 1)) if bits > 1 else 1)
    y = np.where(rng.integers(0, 2, N) == 1, 2 ** 31 - 1, -(2 ** 31)).astype(np.int32)
    assert np.array_equal(O.polymul(x, y, O.ROUTE_FFT), O.polymul(x, y, O.ROUTE_EXACT))


def test_transform_roundtrip_and_evaluation():
    rng = np.random.default_rng(1)
    p = random_torus(rng, N)
    Z = O.forward_transform(p)
    assert np.array_equal(O.inverse_transform(Z), p)
    # Z_k = p(exp(-i*pi*(4k+1)/N))   
    k = np.array([0, 1, 17, 511])
    roots = np.exp(-1j * np.pi * (4 * k + 1) / N)
    ref = np.array([np.polyval(p[::-1].astype(np.float64), r) for r in roots])
    assert np.allclose(Z[k], ref, rtol=1e-9)


def test_mul_by_monomial_and_reverse():
    rng = np.random.default_rng(2)
    p = random_torus(rng, N)
    assert np.array_equal(O.mul_by_monomial(p, 0), p)
    assert np.array_equal(O.mul_by_monomial(p, N), (-p.astype(np.int64)).astype(np.int32))
    assert np.array_equal(O.mul_by_monomial(p, 2 * N + 3), O.mul_by_monomial(p, 3))
    assert np.array_equal(O.mul_by_monomial(O.mul_by_monomial(p, 700), -700), p)
    mono = np.zeros(N, dtype=np.int32); mono[5] = 1
    assert np.array_equal(O.mul_by_monomial(p, 5), O.polymul(mono, p, O.ROUTE_EXACT))
    r = O.reverse_polynomial(p)   # tlwe.jl:55-59 via polynomials.jl:32-35
    assert r[0] == p[0] and np.array_equal(r[1:], (-p[:0:-1].astype(np.int64)).astype(np.int32))


@pytest.mark.parametrize("l,bgbit", [(2, 10), (3, 7), (4, 7), (5, 6), (8, 4)])
def test_decompose_reconstructs_floor(l, bgbit):
    rng = np.random.default_rng(l)
    p = random_torus(rng, N)
    d = O.decompose(p, l, bgbit).astype(np.int64)
    assert d.min() >= -(1 << (bgbit - 1)) and d.max() < (1 << (bgbit - 1))
    recon = sum(d[r] << (32 - (r + 1) * bgbit) for r in range(l))
    err = (p.astype(np.int64) - recon) % 2 ** 32
    assert err.max() < (1 << (32 - l * bgbit))
    assert not O.decompose(np.zeros(N, dtype=np.int32), l, bgbit).any()   # why skipping abar == 0 is exact


def test_modswitch_range():
    x = np.array([0, 2 ** 20 - 1, 2 ** 20, -(2 ** 20) - 1, 2 ** 31 - 1, -(2 ** 31)], dtype=np.int32)
    assert O.decode_message(x, 2048).tolist() == [0, 0, 1, -1, -1024, -1024]


def test_extern_product_fft_equals_exact(octx80_small):
    rng = np.random.default_rng(3)
    acc = random_torus(rng, 2, N)
    for i in (0, 5, 23):
        assert np.array_equal(octx80_small.extern_mul(i, acc, O.ROUTE_FFT), octx80_small.extern_mul(i, acc, O.ROUTE_EXACT))


def test_blind_rotate_fft_equals_exact(octx80_small):
    rng = np.random.default_rng(4)
    acc = random_torus(rng, 2, N)
    bara = rng.integers(-N, N, 24).astype(np.int32); bara[3] = 0
    a = octx80_small.blind_rotate(acc, bara, O.ROUTE_FFT, n_iter=6)
    b = octx80_small.blind_rotate(acc, bara, O.ROUTE_EXACT, n_iter=6)
    assert np.array_equal(a, b)


def test_keyswitch_preserves_phase(keys80, octx80):
    rng = np.random.default_rng(5)
    ext_key = keys80.tlwe_key.reshape(-1)
    u = random_torus(rng, 4, N + 1)
    out = octx80.keyswitch(u)
    before = O.phase(keys80, u, key=ext_key).astype(np.int64)
    after = O.phase(keys80, out).astype(np.int64)
    err = ((after - before + 2 ** 31) % 2 ** 32 - 2 ** 31) / 2 ** 32
    assert np.abs(err).max() < 2e-3


# ---- the reference's own test-suite, restated (test/runtests.jl) ----
BINARY = [O.NAND, O.OR, O.AND, O.XOR, O.XNOR, O.NOR, O.ANDNY, O.ANDYN, O.ORNY, O.ORYN]


@pytest.mark.parametrize("op", BINARY, ids=[O.GATE_NAMES[g] for g in BINARY])
def test_gate_truth_table_80(op, keys80, octx80):
    """test/runtests.jl:26-40 ("gate" testcase)"""
    bits = np.array(list(itertools.product([False, True], repeat=2)))
    rng = O.Rng(1000 + op)
    x, y = O.encrypt(rng, keys80, bits[:, 0]), O.encrypt(rng, keys80, bits[:, 1])
    out = octx80.gate(op, x, y)
    assert np.array_equal(O.decrypt(keys80, out), PLAIN_GATES[op](bits[:, 0], bits[:, 1]))
    ph = O.phase(keys80, out).astype(np.float64) / 2 ** 32
    assert np.abs(np.abs(ph) - 0.125).max() < 1 / 16          # gates.jl:1-6 noise contract


def test_gate_not_constant_mux_80(keys80, octx80):
    rng = O.Rng(5)
    bits = np.array(list(itertools.product([False, True], repeat=3)))
    x, y, z = (O.encrypt(rng, keys80, bits[:, i]) for i in range(3))
    assert np.array_equal(O.decrypt(keys80, octx80.gate(O.NOT, x)), ~bits[:, 0])
    assert np.array_equal(O.decrypt(keys80, octx80.gate(O.MUX, x, y, z)), np.where(bits[:, 0], bits[:, 1], bits[:, 2]))
    flags = np.zeros((2, keys80.params.n + 1), dtype=np.int32); flags[1, 0] = 1
    assert O.decrypt(keys80, octx80.gate(O.CONSTANT, flags)).tolist() == [False, True]


def test_nand_128(keys128, octx128):
    """test/runtests.jl:43-57 ("single party, custom parameters")"""
    bits = np.array(list(itertools.product([False, True], repeat=2)))
    rng = O.Rng(9)
    x, y = O.encrypt(rng, keys128, bits[:, 0]), O.encrypt(rng, keys128, bits[:, 1])
    assert np.array_equal(O.decrypt(keys128, octx128.gate(O.NAND, x, y)), ~(bits[:, 0] & bits[:, 1]))


def test_mk_nand_2party(mkkeys2, mkctx2):
    """test/runtests.jl:60-100 ("multikey NAND"): 10 random trials, encrypt/decrypt round trip + NAND"""
    prng = np.random.default_rng(11)
    bits = prng.integers(0, 2, (10, 2)).astype(bool)
    rng = O.Rng(12)
    x, y = O.mk_encrypt(rng, mkkeys2, bits[:, 0]), O.mk_encrypt(rng, mkkeys2, bits[:, 1])
    assert np.array_equal(O.mk_decrypt(mkkeys2, x), bits[:, 0])
    assert np.array_equal(O.mk_decrypt(mkkeys2, y), bits[:, 1])
    out = mkctx2.nand(x, y)
    assert np.array_equal(O.mk_decrypt(mkkeys2, out), ~(bits[:, 0] & bits[:, 1]))


def test_mk_extern_product_fft_equals_exact_and_semantics():
    import dataclasses
    P = O.small_params(O.MK_PARAMS[2], 8)
    mk = O.mk_keygen(P, 2, 5)
    ctx = O.MKContext(mk)
    rng = np.random.default_rng(6)
    acc = random_torus(rng, 3, N)

    def mkphase(a):
        ph = a[2].astype(np.int64)
        for i in range(2):
            ph = ph - O.polymul(mk.tlwe_keys[i], a[i], O.ROUTE_EXACT)
        return (ph + 2 ** 31) % 2 ** 32 - 2 ** 31

    for party in range(2):
        for j in (0, 7):
            e = ctx.extern_mul(party, j, acc, O.ROUTE_EXACT)
            assert np.array_equal(ctx.extern_mul(party, j, acc, O.ROUTE_FFT), e)
            # RGSW (x) RLWE semantics: phase(out) ~ s * phase(acc)
            s = int(mk.lwe_keys[party, j])
            err = ((mkphase(e) - s * mkphase(acc) + 2 ** 31) % 2 ** 32 - 2 ** 31) / 2 ** 32
            assert np.abs(err).max() < 0.02


def test_tutorial_minimum_is_42(keys80, octx80):
    """examples/tutorial.jl:19-83: 16-bit encrypted minimum of 2017 and 42, level by level."""
    rng = O.Rng(123)
    to_bits = lambda v: np.array([(v >> i) & 1 for i in range(16)], dtype=bool)
    a, b = O.encrypt(rng, keys80, to_bits(2017)), O.encrypt(rng, keys80, to_bits(42))
    flags = np.zeros((1, keys80.params.n + 1), dtype=np.int32)
    carry = octx80.gate(O.CONSTANT, flags)                      # tutorial.jl:53
    for i in range(16):                                         # tutorial.jl:55-57 -> :42-45
        tmp = octx80.gate(O.XNOR, a[i:i + 1], b[i:i + 1])
        carry = octx80.gate(O.MUX, tmp, carry, a[i:i + 1])
    sel = np.repeat(carry, 16, axis=0)
    res = octx80.gate(O.MUX, sel, b, a)                         # tutorial.jl:61
    bits = O.decrypt(keys80, res)
    assert sum(int(v) << i for i, v in enumerate(bits)) == 42   # tutorial.jl:77


# ---- tlwe_mask_size > 1 (api.jl:30,55): the oracle's loops over k+1 against an independent statement ----
def _with_k(base, k, n):
    return O.Params(n, base.lwe_sigma, base.N, k, base.l, base.bgbit, base.bs_sigma, base.t, base.basebit, base.ks_sigma, 1)


@pytest.mark.parametrize("k", [2, 3])
def test_extern_product_mask_size_gt1_matches_definition(k):
    """tgsw_extern_mul (tgsw.jl:125-129) restated from the definition in numpy + Python integers:
    out_c' = sum_{c, r} digit_r(acc_c) (*) BK[i][r][c][c'], digits from tgsw.jl:99-117."""
    P = _with_k(O.PARAMS_80, k, 3)
    keys = O.keygen(P, 11 + k)
    octx = O.Context(keys)
    rng = np.random.default_rng(k)
    acc = random_torus(rng, k + 1, N)
    l, bg = P.l, P.bgbit
    offset = sum(1 << (32 - r * bg) for r in range(1, l + 1)) * (1 << (bg - 1))
    want = np.zeros((k + 1, N), dtype=np.int64)
    for c in range(k + 1):
        u = (acc[c].astype(np.int64) + offset) & 0xFFFFFFFF
        for r in range(l):
            digit = (((u >> (32 - (r + 1) * bg)) & ((1 << bg) - 1)) - (1 << (bg - 1))).astype(np.int32)
            for c2 in range(k + 1):
                want[c2] += exact_negacyclic(digit, keys.bk[1, r, c, c2]).astype(np.int64)
    want = (((want + 2 ** 31) % 2 ** 32) - 2 ** 31).astype(np.int32)
    assert np.array_equal(octx.extern_mul(1, acc, O.ROUTE_EXACT), want)
    assert np.array_equal(octx.extern_mul(1, acc, O.ROUTE_FFT), want)


def test_gate_truth_tables_mask_size_2(keys80):
    """Full-size 80-bit set with two mask polynomials: NAND / XOR / MUX decrypt to their truth tables and the phases stay
    inside the 1/16 contract of gates.jl:1-6 (a wrong index order anywhere in the k+1 loops would not)."""
    keys = O.keygen(_with_k(O.PARAMS_80, 2, 500), 77)
    octx = O.Context(keys)
    bits = np.array(list(itertools.product([False, True], repeat=3)))
    rng = O.Rng(52)
    x, y, z = (O.encrypt(rng, keys, bits[:, i]) for i in range(3))
    nand = octx.gate(O.NAND, x, y)
    assert np.array_equal(O.decrypt(keys, nand), ~(bits[:, 0] & bits[:, 1]))
    assert np.array_equal(O.decrypt(keys, octx.gate(O.XOR, x, y)), bits[:, 0] ^ bits[:, 1])
    assert np.array_equal(O.decrypt(keys, octx.gate(O.MUX, x, y, z)), np.where(bits[:, 0], bits[:, 1], bits[:, 2]))
    ph = O.phase(keys, nand).astype(np.float64) / 2 ** 32
    assert np.abs(np.abs(ph) - 0.125).max() < 1 / 16
