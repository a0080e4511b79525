"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py with the exact-integer
route of the oracle).  CPU: the oracle — both routes — reproduces them.  GPU: the CUDA kernels, through the
C ABI, reproduce them bit for bit."""
import os

import numpy as np
import pytest

from oracle import oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N = 1024


def load(name):
    return np.load(os.path.join(G, name))


def tiny_keys():
    d = load("tiny_bootstrap.npz")
    keys = O.keygen(O.small_params(O.PARAMS_80, int(d["n"])), int(d["seed"]))
    assert [int(keys.bk.astype(np.int64).sum() % 2 ** 31), int(keys.ksk.astype(np.int64).sum() % 2 ** 31)] == d["key_checksum"].tolist()
    return d, keys


def tiny_mk_keys():
    d = load("tiny_mk.npz")
    mk = O.mk_keygen(O.small_params(O.MK_PARAMS[2], int(d["n"])), int(d["parties"]), int(d["seed"]))
    assert int(mk.bk.astype(np.int64).sum() % 2 ** 31) == int(d["key_checksum"][0])
    return d, mk


# ------------------------------------------------------------------ CPU: pins the oracle
def test_oracle_polymul_golden():
    d = load("polymul.npz")
    for i in range(d["x"].shape[0]):
        assert np.array_equal(O.polymul(d["x"][i], d["y"][i], O.ROUTE_EXACT), d["prod"][i])
    for i in (0, 1, 3):   # operands inside the reference FFT's stated precision range (polynomials.jl:138-140)
        assert np.array_equal(O.polymul(d["x"][i], d["y"][i], O.ROUTE_FFT), d["prod"][i])


def test_oracle_primitives_golden():
    d = load("primitives.npz")
    p = d["p"]
    for l, bg in ((2, 10), (3, 7), (4, 7), (8, 4)):
        assert np.array_equal(O.decompose(p, l, bg), d[f"dec_{l}_{bg}"])
    assert np.array_equal(O.decode_message(p, 2 * N), d["modswitch"])
    assert np.array_equal(O.mul_by_monomial(p, 5), d["rot_5"])
    assert np.array_equal(O.mul_by_monomial(p, -700), d["rot_m700"])
    assert np.array_equal(O.mul_by_monomial(p, 1500), d["rot_1500"])
    assert np.array_equal(O.reverse_polynomial(p), d["reverse"])


def test_oracle_tiny_bootstrap_golden():
    d, keys = tiny_keys()
    ctx = O.Context(keys)
    for route in (O.ROUTE_EXACT, O.ROUTE_FFT):
        for i in range(2):
            assert np.array_equal(ctx.extern_mul(i, d["acc"][i], route), d["ext"][i])
            assert np.array_equal(ctx.blind_rotate(d["acc"][i], d["bara"][i], route), d["blind_rotate"][i])
        assert np.array_equal(ctx.bootstrap_wo_ks(d["cts"][0], route=route), d["bootstrap_wo_ks"])
        cts = d["cts"]
        for name, op in (("NAND", O.NAND), ("XOR", O.XOR), ("ORNY", O.ORNY), ("MUX", O.MUX)):
            assert np.array_equal(ctx.gate(op, cts[0], cts[1], cts[2] if op == O.MUX else None, route=route), d["gate_" + name])
    assert np.array_equal(ctx.keyswitch(d["bootstrap_wo_ks"]), d["keyswitch"])


def test_oracle_tiny_mk_golden():
    d, mk = tiny_mk_keys()
    ctx = O.MKContext(mk)
    for route in (O.ROUTE_EXACT, O.ROUTE_FFT):
        for i in range(2):
            assert np.array_equal(ctx.extern_mul(i, 1, d["acc"][i], route), d["ext"][i])
        assert np.array_equal(ctx.nand(d["x"], d["y"], route=route), d["nand"])


# ------------------------------------------------------------------ GPU: the kernels against the same vectors
def _gpu_ctx(keys, parties=1, flags=0):
    import tfhe_jl_b200 as T
    P = keys.params
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, parties=parties, flags=flags)
    ctx.load_bk(keys.bk)
    ctx.load_ksk(keys.ksk)
    return ctx


@pytest.mark.gpu
def test_gpu_polymul_golden():
    d, keys = tiny_keys()
    p = load("polymul.npz")
    assert np.array_equal(_gpu_ctx(keys).polymul(p["x"], p["y"]), p["prod"])


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, 1], ids=["split", "unsplit"])
def test_gpu_tiny_bootstrap_golden(flags):
    d, keys = tiny_keys()
    ctx = _gpu_ctx(keys, flags=flags)
    assert np.array_equal(ctx.extern_product(d["acc"], d["ext_index"]), d["ext"])
    assert np.array_equal(ctx.blind_rotate(d["acc"], d["bara"]), d["blind_rotate"])
    assert np.array_equal(ctx.bootstrap_wo_ks(d["cts"][0]), d["bootstrap_wo_ks"])
    assert np.array_equal(ctx.keyswitch(d["bootstrap_wo_ks"]), d["keyswitch"])
    cts = d["cts"]
    for name, op in (("NAND", O.NAND), ("XOR", O.XOR), ("ORNY", O.ORNY), ("MUX", O.MUX)):
        assert np.array_equal(ctx.gate(op, cts[0], cts[1], cts[2] if op == O.MUX else None), d["gate_" + name]), name


@pytest.mark.gpu
def test_gpu_tiny_mk_golden():
    d, mk = tiny_mk_keys()
    ctx = _gpu_ctx(mk, parties=2)
    assert np.array_equal(ctx.extern_product(d["acc"], np.array([1, 1], np.int32), np.array([0, 1], np.int32)), d["ext"])
    assert np.array_equal(ctx.mk_nand(d["x"], d["y"]), d["nand"])
