"""GPU parity for TLWE mask sizes k = 2 and 3 (`tfhe_parameters_80(; tlwe_mask_size = k)`, api.jl:30,55): the k-generic
kernels of blind_rotate_wide.cuh, called through the C ABI, against the oracle (whose code is k-generic by construction:
tfhe_oracle.c follows tgsw.jl:99-129 / bootstrap.jl:19-95 with loops over k+1)."""
import itertools

import numpy as np
import pytest

import tfhe_jl_b200 as T
from tfhe_jl_b200 import _cabi
from conftest import PLAIN_GATES, random_torus
from oracle import oracle as O

pytestmark = pytest.mark.gpu
N = 1024


def with_k(base, k, n=None):
    return O.Params(base.n if n is None else n, base.lwe_sigma, base.N, k, base.l, base.bgbit, base.bs_sigma, base.t,
                    base.basebit, base.ks_sigma, 1)


def make_ctx(keys, flags=_cabi.FLAG_SPLIT_FFT):
    P = keys.params
    ctx = T.Context(n=P.n, N=P.N, k=P.k, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, flags=flags)
    ctx.load_bk(keys.bk)
    ctx.load_ksk(keys.ksk)
    return ctx


_cache = {}


def small(base_name, k):
    """(keys, oracle context) with a 12-element LWE key: the exact O(N^2) route stays fast."""
    key = (base_name, k)
    if key not in _cache:
        base = O.PARAMS_80 if base_name == "80" else O.PARAMS_128
        keys = O.keygen(with_k(base, k, 12), 900 + k)
        _cache[key] = (keys, O.Context(keys))
    return _cache[key]


CASES = [("80", 2), ("80", 3), ("128", 2), ("128", 3)]


@pytest.mark.parametrize("flags", [_cabi.FLAG_SPLIT_FFT, _cabi.FLAG_UNSPLIT_FFT], ids=["split", "unsplit"])
@pytest.mark.parametrize("base,k", CASES)
def test_extern_product_matches_exact_oracle(base, k, flags):
    """tgsw_extern_mul (tgsw.jl:125-129) with l(k+1) x (k+1) key elements, zero / extreme accumulators included."""
    keys, octx = small(base, k)
    ctx = make_ctx(keys, flags)
    rng = np.random.default_rng(10 * k)
    acc = random_torus(rng, 7, k + 1, N)
    acc[0] = 0; acc[1] = 2 ** 31 - 1; acc[2] = -(2 ** 31)
    idx = rng.integers(0, 12, 7).astype(np.int32)
    got = ctx.extern_product(acc, idx)
    for g in range(7):
        assert np.array_equal(got[g], octx.extern_mul(int(idx[g]), acc[g], O.ROUTE_EXACT)), f"item {g}"


@pytest.mark.parametrize("base,k", CASES)
def test_blind_rotate_prefixes_match_exact_oracle(base, k):
    """blind_rotate (bootstrap.jl:32-39) on k+1 accumulator polynomials after 0, 1, 2, 5 and all iterations."""
    keys, octx = small(base, k)
    ctx = make_ctx(keys)
    rng = np.random.default_rng(3 + k)
    acc = random_torus(rng, 3, k + 1, N)
    bara = rng.integers(-N, N, (3, 12)).astype(np.int32)
    bara[0, 2] = 0; bara[1, :] = 0; bara[2, 0] = -N
    for n_iter in (0, 1, 2, 5, 12):
        got = ctx.blind_rotate(acc, bara, n_iter)
        for g in range(3):
            assert np.array_equal(got[g], octx.blind_rotate(acc[g], bara[g], O.ROUTE_EXACT, n_iter)), (n_iter, g)


@pytest.mark.parametrize("base,k", [("80", 2), ("128", 3)])
def test_every_gate_small_key(base, k):
    """All 13 gates of gates.jl: the extracted sample has k*N mask words, the key switch gathers over all of them."""
    keys, octx = small(base, k)
    ctx = make_ctx(keys)
    rng = O.Rng(40 + k)
    for op in [O.NAND, O.OR, O.AND, O.XOR, O.XNOR, O.NOR, O.ANDNY, O.ANDYN, O.ORNY, O.ORYN, O.NOT, O.MUX]:
        nargs = 3 if op == O.MUX else (1 if op == O.NOT else 2)
        bits = np.array(list(itertools.product([False, True], repeat=nargs)))
        cts = [O.encrypt(rng, keys, bits[:, i]) for i in range(nargs)]
        got = ctx.gate(op, *cts)
        assert np.array_equal(got, octx.gate(op, *cts)), O.GATE_NAMES[op]
        plain = {O.NOT: lambda a: ~a, O.MUX: lambda a, b, c: np.where(a, b, c)}.get(op) or PLAIN_GATES[op]
        assert np.array_equal(O.decrypt(keys, got), plain(*[bits[:, i] for i in range(nargs)])), O.GATE_NAMES[op]
    # the two halves of bootstrap (bootstrap.jl:85-95) on their own: widths k*N + 1 and n + 1
    x = O.encrypt(rng, keys, [True, False, True])
    u = ctx.bootstrap_wo_ks(x)
    assert u.shape == (3, k * N + 1) and np.array_equal(u, octx.bootstrap_wo_ks(x))
    assert np.array_equal(ctx.keyswitch(u), octx.keyswitch(u))
    assert np.array_equal(ctx.bootstrap(x), octx.bootstrap(x))


@pytest.mark.parametrize("count", [1, 2, 3, 301])
def test_ragged_batches_small_key(count):
    """Two gates per CTA: odd counts leave the last CTA half empty; 301 is more than one wave of 148 CTAs."""
    keys, octx = small("80", 2)
    ctx = make_ctx(keys)
    bits = np.random.default_rng(count).integers(0, 2, (count, 3)).astype(bool)
    rng = O.Rng(count)
    x, y, z = (O.encrypt(rng, keys, bits[:, i]) for i in range(3))
    assert np.array_equal(ctx.gate(O.NAND, x, y), octx.gate(O.NAND, x, y))
    assert np.array_equal(ctx.gate(O.MUX, x, y, z), octx.gate(O.MUX, x, y, z))
    assert ctx.gate(O.NAND, x[:0], y[:0], count=0).shape == (0, keys.params.n + 1)


@pytest.mark.parametrize("k", [2, 3])
def test_full_size_80bit_truth_table(k):
    """n = 500: NAND / MUX truth tables ciphertext-identical to the oracle, phases inside the 1/16 contract; the
    unsplit transform gives the same bits."""
    keys = O.keygen(with_k(O.PARAMS_80, k), 77)
    octx = O.Context(keys)
    ctx = make_ctx(keys)
    bits = np.array(list(itertools.product([False, True], repeat=3)))
    rng = O.Rng(50 + k)
    x, y, z = (O.encrypt(rng, keys, bits[:, i]) for i in range(3))
    nand, mux = ctx.gate(O.NAND, x, y), ctx.gate(O.MUX, x, y, z)
    assert np.array_equal(nand, octx.gate(O.NAND, x, y))
    assert np.array_equal(mux, octx.gate(O.MUX, x, y, z))
    assert np.array_equal(O.decrypt(keys, nand), ~(bits[:, 0] & bits[:, 1]))
    assert np.array_equal(O.decrypt(keys, mux), np.where(bits[:, 0], bits[:, 1], bits[:, 2]))
    ph = O.phase(keys, nand).astype(np.float64) / 2 ** 32
    assert np.abs(np.abs(ph) - 0.125).max() < 1 / 16
    assert np.array_equal(make_ctx(keys, _cabi.FLAG_UNSPLIT_FFT).gate(O.NAND, x, y), nand)


def test_device_keygen_words_k2_equals_exact_arithmetic():
    """tfhe_b200_keygen_bk_words with k = 2: every TLWE sample has two mask polynomials, b = noise + S_1 (*) a_1 +
    S_2 (*) a_2 (tlwe.jl:63-73), gadget on the diagonal of the (k+1) x (k+1) block (tgsw.jl:62-69)."""
    P = with_k(O.PARAMS_80, 2, 6)
    k, l = 2, P.l
    rng = np.random.default_rng(8)
    lwe_key = rng.integers(0, 2, P.n).astype(np.int32)
    tlwe_key = rng.integers(0, 2, (k, N)).astype(np.int32)
    S = P.n * l * (k + 1)
    a = random_torus(rng, S, k, N)
    noise = rng.integers(-2 ** 20, 2 ** 20, (S, N)).astype(np.int32)
    ctx = T.Context(n=P.n, k=k, l=l, bgbit=P.bgbit, t=P.t, basebit=P.basebit)
    bk = ctx.keygen_bk(lwe_key, tlwe_key, a=a, noise=noise)
    assert bk.shape == (P.n, l, k + 1, k + 1, N)
    want = np.empty_like(bk).reshape(S, k + 1, N)
    for s in range(S):
        i, r, j = s // (k + 1) // l, s // (k + 1) % l, s % (k + 1)
        want[s, :k] = a[s]
        b = noise[s].astype(np.int64)
        for c in range(k):
            b = b + O.polymul(tlwe_key[c], a[s, c], O.ROUTE_EXACT).astype(np.int64)
        want[s, k] = ((b + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int32)
        g = (int(lwe_key[i]) << (32 - (r + 1) * P.bgbit))
        want[s, j, 0] = np.int32(((int(want[s, j, 0]) + g + 2 ** 31) % 2 ** 32) - 2 ** 31)
    assert np.array_equal(bk.reshape(S, k + 1, N), want)
    # the key the device holds is the one it returned: an external product with it equals the oracle's on the returned words
    ksk = np.zeros(P.ksk_shape, dtype=np.int32)
    octx = O.Context(O.KeySet(P, lwe_key, tlwe_key, bk, ksk))
    acc = random_torus(rng, 2, k + 1, N)
    got = ctx.extern_product(acc, np.array([0, 5], np.int32))
    for g, i in enumerate((0, 5)):
        assert np.array_equal(got[g], octx.extern_mul(i, acc[g], O.ROUTE_EXACT))


def test_api_mirror_mask_size_2():
    """`tfhe_parameters_80(tlwe_mask_size=2)` through the host mirror of api.jl, keys generated on the device."""
    rng = np.random.default_rng(5)
    params = T.tfhe_parameters_80(tlwe_mask_size=2)
    sk = T.SecretKey(rng, params)
    ck = T.CloudKey(rng, sk, device_keygen=True)
    assert ck.bootstrap_key.shape == (500, 2, 3, 3, N) and ck.keyswitch_key.shape[0] == 2 * N
    tt = np.array(list(itertools.product([False, True], repeat=2)))
    out = T.gate_nand(ck, T.encrypt(rng, sk, tt[:, 0]), T.encrypt(rng, sk, tt[:, 1]))
    assert np.array_equal(T.decrypt(sk, out), ~(tt[:, 0] & tt[:, 1]))


def test_mk_context_rejects_mask_size_2():
    with pytest.raises(T.TFHEB200Error):
        T.Context(n=500, k=2, l=4, bgbit=7, parties=2)
