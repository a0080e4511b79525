"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/tfhe_b200.h declares, refuses to compute without a GPU (no CPU fallback), and the host mirror's
integer helpers agree with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

import tfhe_jl_b200 as T
from tfhe_jl_b200 import _cabi, api
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tfhe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tfhe_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/tfhe_b200.h but not exported"
    assert set(names) == set(T.lib()._signatures), "ctypes binding and header disagree"


def test_no_cpu_fallback():
    if T.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(T.TFHEB200Error) as e:
        T.Context(n=500)
    assert e.value.code == _cabi.ENODEV


def test_create_rejects_unsupported_parameters():
    lib = T.lib()
    h = ctypes.c_void_p()
    for bad in (dict(N=2048), dict(k=4), dict(k=0), dict(k=2, parties=2, l=4, bgbit=7), dict(l=7, bgbit=3), dict(parties=3, l=2, bgbit=10)):
        kw = dict(n=500, N=1024, k=1, l=2, bgbit=10, t=8, basebit=2, parties=1); kw.update(bad)
        rc = lib.tfhe_b200_create(ctypes.byref(_cabi.CParams(*[kw[f] for f in ("n", "N", "k", "l", "bgbit", "t", "basebit", "parties")])), 0, 0, ctypes.byref(h))
        assert rc == _cabi.EINVAL and not h.value
        assert lib.tfhe_b200_last_error(None)


def test_host_mirror_scalars_match_oracle():
    rng = np.random.default_rng(0)
    x = rng.integers(-(2 ** 31), 2 ** 31, 1000, dtype=np.int64).astype(np.int32)
    assert np.array_equal(api.decode_message(x, 2048), O.decode_message(x, 2048))
    for mu, ms in ((1, 8), (-1, 8), (1, 4), (-1, 4)):
        assert api.encode_message(mu, ms) == O.encode_message(mu, ms)
    for l, bg in ((2, 10), (4, 7), (8, 4)):
        assert np.array_equal(api._decompose(x[None, :1000].repeat(1, 0)[0][:1000], l, bg)[:, :1000],
                              O.decompose(np.pad(x, (0, 24)), l, bg)[:, :1000])
    assert int(api.dtot32(0.25)) == 2 ** 30 and int(api.dtot32(-0.25)) == -(2 ** 30)
    with pytest.raises(OverflowError):
        api.dtot32(0.75)


def test_host_mirror_encrypt_decrypt_roundtrip():
    rng = np.random.default_rng(1)
    sk = api.SecretKey(rng, api.tfhe_parameters_80())
    bits = rng.integers(0, 2, 64).astype(bool)
    ct = api.encrypt(rng, sk, bits)
    assert ct.data.shape == (64, 501)
    assert np.array_equal(api.decrypt(sk, ct), bits)
    assert api.decrypt(sk, api.encrypt(rng, sk, True)) is True
    # the oracle reads the same ciphertext layout
    ks = O.KeySet(O.PARAMS_80, sk.key, None, None, None)
    assert np.array_equal(O.decrypt(ks, ct.data), bits)
    # lwe.jl:67-82 arithmetic
    nand_lin = api.lwe_noiseless_trivial(api.encode_message(1, 8), 500) - ct[0] - ct[1]
    assert np.array_equal(nand_lin.data, O.gate_prologue(O.NAND, ct.data[0], ct.data[1], 500))
    xor_lin = api.lwe_noiseless_trivial(api.encode_message(1, 4), 500) + (ct[0] + ct[1]) * 2
    assert np.array_equal(xor_lin.data, O.gate_prologue(O.XOR, ct.data[0], ct.data[1], 500))


def test_host_keyswitch_key_matches_reference_semantics():
    """keyswitch.jl:14-41: key[i][j][h-1] decrypts to (s'_i * h) << (32 - (j+1)*basebit) + small noise."""
    rng = np.random.default_rng(2)
    out_key = api.rand_uniform_bool(rng, 40)
    in_key = api.rand_uniform_bool(rng, 16)
    ksk = api._keyswitch_key(rng, 2.0 ** -15, 8, 2, out_key, in_key)
    assert ksk.shape == (16, 8, 3, 41)
    ph = api._wrap(ksk[..., -1].astype(np.int64) - api._dot(ksk[..., :-1], out_key)).astype(np.int64)
    h = np.arange(1, 4)[None, None, :]; j = np.arange(1, 9)[None, :, None]
    msg = api._wrap((in_key[:, None, None].astype(np.int64) * h) << (32 - 2 * j)).astype(np.int64)
    err = ((ph - msg + 2 ** 31) % 2 ** 32 - 2 ** 31) / 2 ** 32
    assert np.abs(err).max() < 2e-4


def test_secret_key_and_ciphertext_files_round_trip(tmp_path):
    """The .npz key / ciphertext files (host logic; the cloud-key file needs a GPU and is tested there)."""
    import numpy as np
    import tfhe_jl_b200 as T
    rng = np.random.default_rng(3)
    sk = T.SecretKey(rng, T.tfhe_parameters_128())
    T.save_secret_key(tmp_path / "sk.npz", sk)
    sk2 = T.load_secret_key(tmp_path / "sk.npz")
    assert sk2.params == sk.params and np.array_equal(sk2.key, sk.key)
    ct = T.encrypt(rng, sk, [True, False, True])
    T.save_ciphertext(tmp_path / "ct.npz", ct)
    ct2 = T.load_ciphertext(tmp_path / "ct.npz")
    assert np.array_equal(ct2.data, ct.data) and list(T.decrypt(sk2, ct2)) == [True, False, True]
    import pytest
    with pytest.raises(ValueError):
        T.load_secret_key(tmp_path / "ct.npz")
