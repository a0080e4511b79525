"""ctypes binding of the CPU ORACLE (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product package
(``tfhe.jl_b200/``) must never import this module.

Parity status: "parity unpinned" at ciphertext level (Julia absent, no golden ciphertexts in the
reference); pinned on the reference's truth-table tests and on exact integer arithmetic.  See
``tfhe_oracle.h`` and DESIGN.md.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

NAND, OR, AND, XOR, XNOR, NOT, CONSTANT, NOR, ANDNY, ANDYN, ORNY, ORYN, MUX = range(13)
GATE_NAMES = ["NAND", "OR", "AND", "XOR", "XNOR", "NOT", "CONSTANT", "NOR", "ANDNY", "ANDYN", "ORNY", "ORYN", "MUX"]
ROUTE_EXACT, ROUTE_FFT = 0, 1


class _Params(C.Structure):
    _fields_ = [(f, C.c_int32) for f in ("n", "N", "k", "l", "bgbit", "t", "basebit", "parties")] + [
        (f, C.c_double) for f in ("lwe_sigma", "bs_sigma", "ks_sigma")
    ]


@dataclass(frozen=True)
class Params:
    """Mirror of SchemeParameters (api.jl:4-21)."""

    n: int
    lwe_sigma: float
    N: int
    k: int
    l: int
    bgbit: int
    bs_sigma: float
    t: int
    basebit: int
    ks_sigma: float
    parties: int = 1

    def c(self) -> _Params:
        return _Params(self.n, self.N, self.k, self.l, self.bgbit, self.t, self.basebit, self.parties,
                       self.lwe_sigma, self.bs_sigma, self.ks_sigma)

    @property
    def bk_shape(self):
        return (self.n, self.l, self.k + 1, self.k + 1, self.N)

    @property
    def ksk_shape(self):
        return (self.N * self.k, self.t, (1 << self.basebit) - 1, self.n + 1)


_S2PI = float(np.sqrt(2.0 / np.pi))
PARAMS_80 = Params(500, 2.0 ** -15 * _S2PI, 1024, 1, 2, 10, 9e-9 * _S2PI, 8, 2, 2.0 ** -15 * _S2PI, 1)   # api.jl:30-45
PARAMS_128 = Params(630, 2.0 ** -15, 1024, 1, 3, 7, 2.0 ** -25, 8, 2, 2.0 ** -15, 1)                      # api.jl:55-69
MK_PARAMS = {                                                                                             # mk_api.jl:4-34
    2: Params(500, 0.012467, 1024, 1, 4, 7, 3.29e-10, 8, 2, 2.44e-5, 2),
    4: Params(500, 0.012467, 1024, 1, 5, 6, 3.29e-10, 8, 2, 2.44e-5, 4),
    8: Params(500, 0.012467, 1024, 1, 8, 4, 3.29e-10, 8, 2, 2.44e-5, 8),
}


def small_params(base: Params, n: int) -> Params:
    """Same parameter set with a shorter LWE key, for fast parity tests."""
    return Params(n, base.lwe_sigma, base.N, base.k, base.l, base.bgbit, base.bs_sigma, base.t, base.basebit,
                  base.ks_sigma, base.parties)


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("tfhe_oracle.c", "mk_oracle.c", "tfhe_oracle.h", "tfhe_oracle_internal.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        flags = ""
        try:
            cpu = open("/proc/cpuinfo").read()
            if not (" avx2" in cpu and " fma" in cpu):
                flags = "ARCHFLAGS="
        except OSError:
            pass
        cmd = ["make", "-C", _HERE, "-B", "liboracle.so"] + ([flags] if flags else [])
        subprocess.run(cmd, check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        i32p, f64p, vp = C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_void_p
        sig = {
            "orc_rng_create": (vp, [C.c_uint64]), "orc_rng_destroy": (None, [vp]),
            "orc_encode_message": (C.c_int32, [C.c_int32, C.c_int32]),
            "orc_decode_message": (C.c_int32, [C.c_int32, C.c_int32]),
            "orc_dtot32": (C.c_int32, [C.c_double]),
            "orc_mul_by_monomial": (None, [i32p, C.c_int64, i32p, C.c_int]),
            "orc_reverse_polynomial": (None, [i32p, i32p, C.c_int]),
            "orc_polymul_exact": (None, [i32p, i32p, i32p, C.c_int]),
            "orc_polymul_fft": (None, [i32p, i32p, i32p, C.c_int]),
            "orc_forward_transform": (None, [i32p, f64p, C.c_int]),
            "orc_inverse_transform": (None, [f64p, i32p, C.c_int]),
            "orc_decomp_offset": (C.c_int32, [C.c_int, C.c_int]),
            "orc_decompose": (None, [i32p, C.c_int, C.c_int, C.c_int, i32p]),
            "orc_keygen": (None, [C.POINTER(_Params), C.c_uint64, i32p, i32p, i32p, i32p]),
            "orc_lwe_encrypt": (None, [vp, C.c_int32, C.c_double, i32p, C.c_int, i32p]),
            "orc_lwe_phase": (C.c_int32, [i32p, i32p, C.c_int]),
            "orc_create": (vp, [C.POINTER(_Params), i32p, i32p]), "orc_destroy": (None, [vp]),
            "orc_extern_mul": (None, [vp, C.c_int, i32p, i32p, C.c_int]),
            "orc_blind_rotate": (None, [vp, i32p, i32p, C.c_int, C.c_int]),
            "orc_tlwe_extract": (None, [i32p, C.c_int, C.c_int, i32p]),
            "orc_bootstrap_wo_ks": (None, [vp, C.c_int32, i32p, i32p, C.c_int]),
            "orc_keyswitch": (None, [vp, i32p, i32p]),
            "orc_bootstrap": (None, [vp, C.c_int32, i32p, i32p, C.c_int]),
            "orc_gate_batch": (None, [vp, C.c_int, i32p, i32p, i32p, i32p, C.c_size_t, C.c_int, C.c_int]),
            "orc_gate_prologue": (None, [C.c_int, i32p, i32p, C.c_int, i32p]),
            "orc_mk_keygen": (None, [C.POINTER(_Params), C.c_int, C.c_uint64, i32p, i32p, i32p, i32p]),
            "orc_mk_bk_words": (C.c_size_t, [C.POINTER(_Params), C.c_int]),
            "orc_mk_encrypt": (None, [vp, C.POINTER(_Params), C.c_int, i32p, C.c_int, i32p]),
            "orc_mk_phase": (C.c_int32, [C.POINTER(_Params), C.c_int, i32p, i32p]),
            "orc_mk_create": (vp, [C.POINTER(_Params), C.c_int, i32p, i32p]), "orc_mk_destroy": (None, [vp]),
            "orc_mk_extern_mul": (None, [vp, C.c_int, C.c_int, i32p, i32p, C.c_int]),
            "orc_mk_bootstrap_wo_ks": (None, [vp, C.c_int32, i32p, i32p, C.c_int]),
            "orc_mk_keyswitch": (None, [vp, i32p, i32p]),
            "orc_mk_nand_batch": (None, [vp, i32p, i32p, i32p, C.c_size_t, C.c_int, C.c_int]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ---------------------------------------------------------------- primitives
def encode_message(mu: int, ms: int) -> int:
    return lib().orc_encode_message(mu, ms)


def decode_message(phase, ms: int):
    phase = np.asarray(phase, dtype=np.int32)
    return np.array([lib().orc_decode_message(int(v), ms) for v in phase.ravel()], dtype=np.int32).reshape(phase.shape)


def mul_by_monomial(p, s: int):
    p = _i32(p); out = np.empty_like(p)
    lib().orc_mul_by_monomial(_p(p), int(s), _p(out), p.size)
    return out


def reverse_polynomial(p):
    p = _i32(p); out = np.empty_like(p)
    lib().orc_reverse_polynomial(_p(p), _p(out), p.size)
    return out


def polymul(x, y, route=ROUTE_EXACT):
    x, y = _i32(x), _i32(y); out = np.empty_like(x)
    (lib().orc_polymul_exact if route == ROUTE_EXACT else lib().orc_polymul_fft)(_p(x), _p(y), _p(out), x.size)
    return out


def forward_transform(p):
    p = _i32(p); out = np.empty(p.size, dtype=np.float64)
    lib().orc_forward_transform(_p(p), out.ctypes.data_as(C.POINTER(C.c_double)), p.size)
    return out.view(np.complex128)


def inverse_transform(z):
    z = np.ascontiguousarray(z, dtype=np.complex128); out = np.empty(z.size * 2, dtype=np.int32)
    lib().orc_inverse_transform(z.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)), _p(out), out.size)
    return out


def decompose(p, l: int, bgbit: int):
    p = _i32(p); out = np.empty((l, p.size), dtype=np.int32)
    lib().orc_decompose(_p(p), p.size, l, bgbit, _p(out))
    return out


def gate_prologue(op, x, y, n):
    out = np.empty(n + 1, dtype=np.int32)
    lib().orc_gate_prologue(op, _p(_i32(x)), _p(_i32(y)), n, _p(out))
    return out


# ---------------------------------------------------------------- keys and contexts
class Rng:
    def __init__(self, seed: int):
        self.h = lib().orc_rng_create(seed)

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().orc_rng_destroy(self.h); self.h = None


@dataclass
class KeySet:
    params: Params
    lwe_key: np.ndarray
    tlwe_key: np.ndarray
    bk: np.ndarray
    ksk: np.ndarray


def keygen(params: Params, seed: int = 123) -> KeySet:
    """make_key_pair (api.jl:139-146) with the oracle's own RNG."""
    P = params
    lwe_key = np.empty(P.n, dtype=np.int32)
    tlwe_key = np.empty((P.k, P.N), dtype=np.int32)
    bk = np.empty(P.bk_shape, dtype=np.int32)
    ksk = np.empty(P.ksk_shape, dtype=np.int32)
    cp = P.c()
    lib().orc_keygen(C.byref(cp), seed, _p(lwe_key), _p(tlwe_key), _p(bk), _p(ksk))
    return KeySet(P, lwe_key, tlwe_key, bk, ksk)


def encrypt(rng: Rng, keys: KeySet, bits) -> np.ndarray:
    """encrypt (api.jl:155-158), batched: returns [count][n+1]."""
    bits = np.atleast_1d(np.asarray(bits, dtype=bool))
    n = keys.params.n
    out = np.empty((bits.size, n + 1), dtype=np.int32)
    for i, b in enumerate(bits.ravel()):
        lib().orc_lwe_encrypt(rng.h, encode_message(1 if b else -1, 8), keys.params.lwe_sigma, _p(keys.lwe_key), n,
                              _p(out[i]))
    return out


def phase(keys: KeySet, cts, key=None) -> np.ndarray:
    cts = np.atleast_2d(_i32(cts))
    key = keys.lwe_key if key is None else _i32(key)
    return np.array([lib().orc_lwe_phase(_p(ct), _p(key), key.size) for ct in cts], dtype=np.int32)


def decrypt(keys: KeySet, cts) -> np.ndarray:
    """decrypt (api.jl:167-169)."""
    return phase(keys, cts) > 0


class Context:
    """The cloud-key side: BK (+ its transform, bootstrap.jl:12) and KSK."""

    def __init__(self, keys: KeySet):
        self.keys = keys
        self.P = keys.params
        cp = self.P.c()
        self.h = lib().orc_create(C.byref(cp), _p(keys.bk), _p(keys.ksk))

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().orc_destroy(self.h); self.h = None

    def extern_mul(self, i, acc, route=ROUTE_EXACT):
        acc = _i32(acc); out = np.empty_like(acc)
        lib().orc_extern_mul(self.h, i, _p(acc), _p(out), route)
        return out

    def blind_rotate(self, acc, bara, route=ROUTE_FFT, n_iter=None):
        acc = _i32(acc).copy(); bara = _i32(bara)
        lib().orc_blind_rotate(self.h, _p(acc), _p(bara), route, bara.size if n_iter is None else n_iter)
        return acc

    def bootstrap_wo_ks(self, x, mu=None, route=ROUTE_FFT):
        x = np.atleast_2d(_i32(x)); mu = encode_message(1, 8) if mu is None else mu
        out = np.empty((x.shape[0], self.P.N * self.P.k + 1), dtype=np.int32)
        for g in range(x.shape[0]):
            lib().orc_bootstrap_wo_ks(self.h, mu, _p(x[g]), _p(out[g]), route)
        return out

    def keyswitch(self, u):
        u = np.atleast_2d(_i32(u)); out = np.empty((u.shape[0], self.P.n + 1), dtype=np.int32)
        for g in range(u.shape[0]):
            lib().orc_keyswitch(self.h, _p(u[g]), _p(out[g]))
        return out

    def bootstrap(self, x, mu=None, route=ROUTE_FFT):
        return self.keyswitch(self.bootstrap_wo_ks(x, mu, route))

    def gate(self, op, x=None, y=None, z=None, route=ROUTE_FFT, nthreads=None, count=None):
        arrs = [None if a is None else np.atleast_2d(_i32(a)) for a in (x, y, z)]
        if count is None:
            count = next(a.shape[0] for a in arrs if a is not None)
        out = np.empty((count, self.P.n + 1), dtype=np.int32)
        nthreads = nthreads or min(os.cpu_count() or 1, max(count, 1))
        lib().orc_gate_batch(self.h, op, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(out), count, route, nthreads)
        return out


# ---------------------------------------------------------------- multi-key
@dataclass
class MKKeySet:
    params: Params
    parties: int
    lwe_keys: np.ndarray   # [p][n]
    bk: np.ndarray         # [p][n][l*(2p+2)][N]
    ksk: np.ndarray        # [p][N*k][t][base-1][n+1]
    tlwe_keys: np.ndarray  # [p][N] (test hook: lets tests check TLWE phases)


def mk_keygen(params: Params, parties: int, seed: int = 123) -> MKKeySet:
    P = params
    cp = P.c()
    lwe_keys = np.empty((parties, P.n), dtype=np.int32)
    bk = np.empty((parties, P.n, P.l * (2 * parties + 2), P.N), dtype=np.int32)
    assert bk.size == lib().orc_mk_bk_words(C.byref(cp), parties)
    ksk = np.empty((parties,) + P.ksk_shape, dtype=np.int32)
    tlwe_keys = np.empty((parties, P.N), dtype=np.int32)
    lib().orc_mk_keygen(C.byref(cp), parties, seed, _p(lwe_keys), _p(bk), _p(ksk), _p(tlwe_keys))
    return MKKeySet(P, parties, lwe_keys, bk, ksk, tlwe_keys)


def mk_encrypt(rng: Rng, keys: MKKeySet, bits) -> np.ndarray:
    bits = np.atleast_1d(np.asarray(bits, dtype=bool))
    P, p = keys.params, keys.parties
    cp = P.c()
    out = np.empty((bits.size, p * P.n + 1), dtype=np.int32)
    for i, b in enumerate(bits.ravel()):
        lib().orc_mk_encrypt(rng.h, C.byref(cp), p, _p(keys.lwe_keys), int(b), _p(out[i]))
    return out


def mk_phase(keys: MKKeySet, cts) -> np.ndarray:
    cts = np.atleast_2d(_i32(cts)); cp = keys.params.c()
    return np.array([lib().orc_mk_phase(C.byref(cp), keys.parties, _p(keys.lwe_keys), _p(ct)) for ct in cts],
                    dtype=np.int32)


def mk_decrypt(keys: MKKeySet, cts) -> np.ndarray:
    return mk_phase(keys, cts) > 0


class MKContext:
    def __init__(self, keys: MKKeySet):
        self.keys = keys; self.P = keys.params; self.p = keys.parties
        cp = self.P.c()
        self.h = lib().orc_mk_create(C.byref(cp), self.p, _p(keys.bk), _p(keys.ksk))

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().orc_mk_destroy(self.h); self.h = None

    def extern_mul(self, party, j, acc, route=ROUTE_EXACT):
        acc = _i32(acc); out = np.empty_like(acc)
        lib().orc_mk_extern_mul(self.h, party, j, _p(acc), _p(out), route)
        return out

    def bootstrap_wo_ks(self, x, mu=None, route=ROUTE_FFT):
        x = np.atleast_2d(_i32(x)); mu = encode_message(1, 8) if mu is None else mu
        out = np.empty((x.shape[0], self.p * self.P.N + 1), dtype=np.int32)
        for g in range(x.shape[0]):
            lib().orc_mk_bootstrap_wo_ks(self.h, mu, _p(x[g]), _p(out[g]), route)
        return out

    def keyswitch(self, u):
        u = np.atleast_2d(_i32(u)); out = np.empty((u.shape[0], self.p * self.P.n + 1), dtype=np.int32)
        for g in range(u.shape[0]):
            lib().orc_mk_keyswitch(self.h, _p(u[g]), _p(out[g]))
        return out

    def nand(self, x, y, route=ROUTE_FFT, nthreads=None):
        x = np.atleast_2d(_i32(x)); y = np.atleast_2d(_i32(y))
        out = np.empty_like(x)
        nthreads = nthreads or min(os.cpu_count() or 1, x.shape[0])
        lib().orc_mk_nand_batch(self.h, _p(x), _p(y), _p(out), x.shape[0], route, nthreads)
        return out
