"""Host-side mirror of the TFHE.jl API surface (src/TFHE.jl:24-61) over the B200 engine.

Same names, argument meaning and error behaviour as the reference's exported functions; ciphertexts
additionally come in *batched* form (an ``LweSample`` may hold ``[count][n+1]`` words), which is how
the gate functions reach the GPU: one C-ABI call per batch, no per-gate launch.

Key generation, encryption and decryption are the reference's "host-keep" side: they run
on the host in numpy, except that every torus-polynomial product of key generation
(``transformed_mul``, polynomials.jl:142-144) is computed by the GPU kernel K1 through the C ABI —
there is no CPU polynomial multiplier in this package and nothing here imports ``oracle/``.

``rng`` is a ``numpy.random.Generator`` (the reference takes a Julia ``AbstractRNG``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _cabi
from ._cabi import Context

Torus32 = np.int32   # numeric-functions.jl:1


# ------------------------------------------------------------------ numeric-functions.jl
def rand_uniform_bool(rng, *dims):            # numeric-functions.jl:4-6
    return rng.integers(0, 2, size=dims, dtype=np.int32)


def rand_uniform_torus32(rng, *dims):         # numeric-functions.jl:9-11
    return rng.integers(-(2 ** 31), 2 ** 31, size=dims, dtype=np.int64).astype(np.int32)


def rand_gaussian_float(rng, sigma, *dims):   # numeric-functions.jl:14-16
    return rng.standard_normal(size=dims) * sigma


def dtot32(d):                                # numeric-functions.jl:51-53 (trunc; InexactError -> OverflowError)
    v = np.trunc(np.asarray(d, dtype=np.float64) * 2.0 ** 32)
    if np.any(v >= 2.0 ** 31) or np.any(v < -(2.0 ** 31)):
        raise OverflowError("dtot32: value outside [-0.5, 0.5)")
    return v.astype(np.int64).astype(np.int32)


def rand_gaussian_torus32(rng, message, sigma, *dims):   # numeric-functions.jl:20-23
    return _wrap(np.int64(message) + dtot32(rng.standard_normal(size=dims) * sigma).astype(np.int64))


def encode_message(mu: int, message_space: int) -> int:  # numeric-functions.jl:42-45
    log2_ms = int(message_space).bit_length() - 1
    return int(_wrap(np.int64(mu) << (32 - log2_ms)))


def decode_message(phase, message_space: int):           # numeric-functions.jl:31-34
    log2_ms = int(message_space).bit_length() - 1
    p = _wrap(np.asarray(phase, dtype=np.int64) + (1 << (32 - log2_ms - 1)))
    return (p >> (32 - log2_ms)).astype(np.int32)


def _wrap(x):
    """Reduce int64 values to two's-complement int32 (Julia Int32 wrap-around)."""
    return ((np.asarray(x, dtype=np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int32)


def _dot(a, key):
    return _wrap((a.astype(np.int64) * key.astype(np.int64)).sum(axis=-1))


# ------------------------------------------------------------------ api.jl: parameters
@dataclass(frozen=True)
class SchemeParameters:                       # api.jl:4-21
    lwe_size: int
    lwe_noise_stddev: float
    tlwe_polynomial_degree: int
    tlwe_mask_size: int
    bs_decomp_length: int
    bs_log2_base: int
    bs_noise_stddev: float
    ks_decomp_length: int
    ks_log2_base: int
    ks_noise_stddev: float
    max_parties: int


_S2PI = float(np.sqrt(2.0 / np.pi))


def tfhe_parameters_80(tlwe_mask_size: int = 1) -> SchemeParameters:    # api.jl:30-45
    return SchemeParameters(500, 1 / 2 ** 15 * _S2PI, 1024, tlwe_mask_size, 2, 10, 9e-9 * _S2PI, 8, 2, 1 / 2 ** 15 * _S2PI, 1)


def tfhe_parameters_128(tlwe_mask_size: int = 1) -> SchemeParameters:   # api.jl:55-69
    return SchemeParameters(630, 1 / 2 ** 15, 1024, tlwe_mask_size, 3, 7, 1 / 2 ** 25, 8, 2, 1 / 2 ** 15, 1)


mktfhe_parameters_2party = SchemeParameters(500, 0.012467, 1024, 1, 4, 7, 3.29e-10, 8, 2, 2.44e-5, 2)   # mk_api.jl:4-10
mktfhe_parameters_4party = SchemeParameters(500, 0.012467, 1024, 1, 5, 6, 3.29e-10, 8, 2, 2.44e-5, 4)   # mk_api.jl:16-22
mktfhe_parameters_8party = SchemeParameters(500, 0.012467, 1024, 1, 8, 4, 3.29e-10, 8, 2, 2.44e-5, 8)   # mk_api.jl:28-34


def _context(params: SchemeParameters, parties: int, device: int, flags: int) -> Context:
    return Context(n=params.lwe_size, N=params.tlwe_polynomial_degree, k=params.tlwe_mask_size,
                   l=params.bs_decomp_length, bgbit=params.bs_log2_base, t=params.ks_decomp_length,
                   basebit=params.ks_log2_base, parties=parties, device=device, flags=flags)


# ------------------------------------------------------------------ lwe.jl
@dataclass
class LweSample:
    """lwe.jl:21-29.  ``data`` is ``[n+1]`` (one ciphertext) or ``[count][n+1]`` (a batch): a then b."""
    data: np.ndarray
    current_variance: float = 0.0

    @property
    def a(self):
        return self.data[..., :-1]

    @property
    def b(self):
        return self.data[..., -1]

    def __len__(self):
        return 1 if self.data.ndim == 1 else self.data.shape[0]

    def __getitem__(self, i):
        return LweSample(self.data[i], self.current_variance)

    # lwe.jl:67-82
    def __add__(self, o): return LweSample(_wrap(self.data.astype(np.int64) + o.data), self.current_variance + o.current_variance)
    def __sub__(self, o): return LweSample(_wrap(self.data.astype(np.int64) - o.data), self.current_variance + o.current_variance)
    def __neg__(self): return LweSample(_wrap(-self.data.astype(np.int64)), self.current_variance)
    def __mul__(self, y: int): return LweSample(_wrap(self.data.astype(np.int64) * int(y)), self.current_variance * y ** 2)
    __rmul__ = __mul__


def lwe_noiseless_trivial(mu: int, size: int) -> LweSample:            # lwe.jl:63-64
    d = np.zeros(size + 1, dtype=np.int32); d[-1] = mu
    return LweSample(d, 0.0)


def lwe_encrypt(rng, message, alpha: float, key: np.ndarray) -> LweSample:   # lwe.jl:38-43 (batched over `message`)
    message = np.asarray(message, dtype=np.int64)
    a = rand_uniform_torus32(rng, *message.shape, key.size)
    b = _wrap(message + dtot32(rng.standard_normal(size=message.shape) * alpha).astype(np.int64) + _dot(a, key))
    return LweSample(np.concatenate([a, np.asarray(b)[..., None]], axis=-1), alpha ** 2)


def lwe_phase(x: LweSample, key: np.ndarray):                          # lwe.jl:59
    return _wrap(x.b.astype(np.int64) - _dot(x.a, key))


# ------------------------------------------------------------------ key generation helpers (tlwe.jl, tgsw.jl, keyswitch.jl)
def _tlwe_encrypt_zero(rng, ctx: Context, alpha: float, tlwe_key: np.ndarray, count: int) -> np.ndarray:
    """tlwe.jl:63-73, `count` samples at once: returns [count][k+1][N]."""
    k, N = tlwe_key.shape
    a = rand_uniform_torus32(rng, count, k, N)
    b = dtot32(rng.standard_normal(size=(count, N)) * alpha).astype(np.int64)
    prod = ctx.polymul(np.broadcast_to(tlwe_key, (count, k, N)).reshape(-1, N), a.reshape(-1, N)).reshape(count, k, N)
    b = _wrap(b + prod.astype(np.int64).sum(axis=1))
    return np.concatenate([a, b[:, None, :]], axis=1)


def _bootstrap_key(rng, ctx: Context, alpha: float, lwe_key: np.ndarray, tlwe_key: np.ndarray, l: int, bgbit: int) -> np.ndarray:
    """bootstrap.jl:6-15 / tgsw.jl:52-88 in coefficient form: [n][l][k+1][k+1][N]."""
    n = lwe_key.size
    k, N = tlwe_key.shape
    bk = _tlwe_encrypt_zero(rng, ctx, alpha, tlwe_key, n * l * (k + 1)).reshape(n, l, k + 1, k + 1, N)
    gadget = np.array([1 << (32 - (r + 1) * bgbit) for r in range(l)], dtype=np.int64)    # tgsw.jl:14
    for j in range(k + 1):                                                               # tgsw.jl:62-69
        bk[:, :, j, j, 0] = _wrap(bk[:, :, j, j, 0].astype(np.int64) + lwe_key[:, None].astype(np.int64) * gadget[None, :])
    return bk


def _keyswitch_key(rng, alpha: float, t: int, basebit: int, out_key: np.ndarray, in_key: np.ndarray) -> np.ndarray:
    """keyswitch.jl:14-41: [N*k][t][base-1][n+1]."""
    base = 1 << basebit
    Nk, n = in_key.size, out_key.size
    noise = rand_gaussian_float(rng, alpha, Nk, t, base - 1)
    noise -= noise.sum() / noise.size                                                     # keyswitch.jl:29
    h = np.arange(1, base, dtype=np.int64)[None, None, :]
    shift = (32 - (np.arange(1, t + 1, dtype=np.int64) * basebit))[None, :, None]
    message = _wrap((in_key.astype(np.int64)[:, None, None] * h) << shift)                 # keyswitch.jl:35
    a = rand_uniform_torus32(rng, Nk, t, base - 1, n)
    b = _wrap(message.astype(np.int64) + dtot32(noise).astype(np.int64) + _dot(a, out_key))   # lwe.jl:49-55
    return np.concatenate([a, b[..., None]], axis=-1)


# ------------------------------------------------------------------ api.jl: keys, encrypt, decrypt
class SecretKey:                              # api.jl:92-100
    def __init__(self, rng, params: SchemeParameters):
        self.params = params
        self.key = rand_uniform_bool(rng, params.lwe_size)   # LweKey, lwe.jl:10-12


class CloudKey:                               # api.jl:111-127
    """Evaluation key.  Key material is generated on the host and loaded onto `device`; the int32
    coefficient form of the bootstrap key is kept (the reference keeps only its transform)."""

    def __init__(self, rng, secret_key: SecretKey, device: int = 0, flags: int = _cabi.FLAG_SPLIT_FFT, devices=None,
                 device_keygen: bool = False):
        """``devices``: a list of GPU ordinals or ``"all"`` — the key is then replicated on every listed GPU and each
        gate batch is sharded across them (one ``gate_nand(ck, x, y)`` uses every GPU).
        ``device_keygen``: every random word of the key (24 576 LWE and 2 000 TLWE encryptions at the 80-bit set) is
        generated on the GPU from one seed drawn from ``rng`` (tfhe_b200_keygen_bk / _ksk) instead of in numpy."""
        p = secret_key.params
        self.params = p
        self.mctx = None
        if device_keygen and devices is None:
            self.ctx = _context(p, 1, device, flags)
            tlwe_key = rand_uniform_bool(rng, p.tlwe_mask_size, p.tlwe_polynomial_degree)
            seed = int(rng.integers(0, 2 ** 63))
            self.bootstrap_key = self.ctx.keygen_bk(secret_key.key, tlwe_key, sigma=p.bs_noise_stddev, seed=seed)
            self.keyswitch_key = self.ctx.keygen_ksk(secret_key.key, tlwe_key.reshape(-1), sigma=p.ks_noise_stddev, seed=seed)
            return
        if devices is not None:
            self.mctx = _cabi.MultiContext(n=p.lwe_size, N=p.tlwe_polynomial_degree, k=p.tlwe_mask_size, l=p.bs_decomp_length,
                                           bgbit=p.bs_log2_base, t=p.ks_decomp_length, basebit=p.ks_log2_base, parties=1,
                                           devices=None if devices == "all" else list(devices), flags=flags)
            self.ctx = self.mctx.context(0)
        else:
            self.ctx = _context(p, 1, device, flags)
        tlwe_key = rand_uniform_bool(rng, p.tlwe_mask_size, p.tlwe_polynomial_degree)     # TLweKey, tlwe.jl:15-20
        self.bootstrap_key = _bootstrap_key(rng, self.ctx, p.bs_noise_stddev, secret_key.key, tlwe_key,
                                            p.bs_decomp_length, p.bs_log2_base)
        self.keyswitch_key = _keyswitch_key(rng, p.ks_noise_stddev, p.ks_decomp_length, p.ks_log2_base, secret_key.key,
                                            tlwe_key.reshape(-1))                         # extract_lwe_key, tlwe.jl:25-31
        holder = self.mctx if self.mctx is not None else self.ctx
        holder.load_bk(self.bootstrap_key)
        holder.load_ksk(self.keyswitch_key)


def make_key_pair(rng, params: Optional[SchemeParameters] = None, device: int = 0, flags: int = _cabi.FLAG_SPLIT_FFT,
                  devices=None):
    """api.jl:139-146"""
    if params is None:
        params = tfhe_parameters_80()
    secret_key = SecretKey(rng, params)
    return secret_key, CloudKey(rng, secret_key, device=device, flags=flags, devices=devices)


def encrypt(rng, key: SecretKey, message, ctx: Optional[Context] = None) -> LweSample:
    """api.jl:155-158; `message` may be a bool or an array of bools (batched).  With a device context (`ctx`, e.g.
    ``cloud_key.ctx``) the mask and the noise are generated on the GPU from one seed drawn from `rng`."""
    m = np.asarray(message, dtype=bool)
    if ctx is not None:
        out = ctx.encrypt(key.key, m.reshape(-1), key.params.lwe_noise_stddev, int(rng.integers(0, 2 ** 63)))
        return LweSample(out.reshape(m.shape + (key.key.size + 1,)), key.params.lwe_noise_stddev ** 2)
    mu = np.where(m, encode_message(1, 8), encode_message(-1, 8))
    return lwe_encrypt(rng, mu, key.params.lwe_noise_stddev, key.key)


def decrypt(key: SecretKey, sample: LweSample, ctx: Optional[Context] = None):
    """api.jl:167-169 (on the GPU when a device context is given)"""
    if ctx is not None:
        r = ctx.decrypt(key.key, sample.data.reshape(-1, key.key.size + 1)).reshape(sample.data.shape[:-1])
        return bool(r) if np.ndim(r) == 0 else r
    r = lwe_phase(sample, key.key) > 0
    return bool(r) if np.ndim(r) == 0 else r


# ------------------------------------------------------------------ gates.jl
def _gate(ck: CloudKey, op: int, *xs: LweSample) -> LweSample:
    single = xs[0].data.ndim == 1
    out = (ck.mctx if getattr(ck, "mctx", None) is not None else ck.ctx).gate(op, *[np.atleast_2d(x.data) for x in xs])
    return LweSample(out[0] if single else out, 0.0)


def gate_nand(ck, x, y): return _gate(ck, _cabi.NAND, x, y)        # gates.jl:15-18
def gate_or(ck, x, y): return _gate(ck, _cabi.OR, x, y)            # gates.jl:27-30
def gate_and(ck, x, y): return _gate(ck, _cabi.AND, x, y)          # gates.jl:39-42
def gate_xor(ck, x, y): return _gate(ck, _cabi.XOR, x, y)          # gates.jl:51-54
def gate_xnor(ck, x, y): return _gate(ck, _cabi.XNOR, x, y)        # gates.jl:63-66
def gate_not(ck, x): return _gate(ck, _cabi.NOT, x)                # gates.jl:76-79
def gate_nor(ck, x, y): return _gate(ck, _cabi.NOR, x, y)          # gates.jl:102-105
def gate_andny(ck, x, y): return _gate(ck, _cabi.ANDNY, x, y)      # gates.jl:114-117
def gate_andyn(ck, x, y): return _gate(ck, _cabi.ANDYN, x, y)      # gates.jl:126-129
def gate_orny(ck, x, y): return _gate(ck, _cabi.ORNY, x, y)        # gates.jl:138-141
def gate_oryn(ck, x, y): return _gate(ck, _cabi.ORYN, x, y)        # gates.jl:150-153
def gate_mux(ck, x, y, z): return _gate(ck, _cabi.MUX, x, y, z)    # gates.jl:163-177


def gate_constant(ck: CloudKey, value) -> LweSample:               # gates.jl:91-93
    v = np.asarray(value, dtype=bool)
    flags = np.zeros(v.shape + (ck.params.lwe_size + 1,), dtype=np.int32)
    flags[..., 0] = v
    out = ck.ctx.gate(_cabi.CONSTANT, np.atleast_2d(flags))
    return LweSample(out[0] if v.ndim == 0 else out, 0.0)


# ------------------------------------------------------------------ mk_internals.jl / mk_api.jl / mk_gates.jl
@dataclass
class MKLweSample:
    """mk_internals.jl:6-18.  ``data`` is ``[p*n+1]`` or ``[count][p*n+1]``: a[party][n] then b."""
    data: np.ndarray
    parties: int
    current_variance: float = 0.0

    @property
    def b(self):
        return self.data[..., -1]

    def __getitem__(self, i):
        return MKLweSample(self.data[i], self.parties, self.current_variance)


class SharedKey:                              # mk_internals.jl:101-112, mk_api.jl:44-50
    def __init__(self, rng, params: SchemeParameters):
        self.params = params
        self.a = rand_uniform_torus32(rng, params.bs_decomp_length, params.tlwe_polynomial_degree)


_keygen_ctx = {}


def _mk_ctx(params: SchemeParameters, device: int) -> Context:
    key = (params, device)
    if key not in _keygen_ctx:
        _keygen_ctx[key] = _context(params, params.max_parties, device, _cabi.FLAG_SPLIT_FFT)
    return _keygen_ctx[key]


class CloudKeyPart:                           # mk_api.jl:61-77
    def __init__(self, rng, secret_key: SecretKey, shared_key: SharedKey, device: int = 0):
        p = secret_key.params
        self.params = p
        ctx = _mk_ctx(p, device)
        l, N, alpha = p.bs_decomp_length, p.tlwe_polynomial_degree, p.bs_noise_stddev
        n = p.lwe_size
        gauss = lambda *dims: dtot32(rng.standard_normal(size=dims) * alpha).astype(np.int64)
        S = rand_uniform_bool(rng, N)                                                     # TLweKey (mask_size 1)
        mul = lambda x, y: ctx.polymul(np.broadcast_to(x, y.shape).reshape(-1, N), y.reshape(-1, N)).reshape(y.shape).astype(np.int64)
        # PublicKey (mk_internals.jl:115-139): b_i = S (*) a_i + e
        self.public_b = _wrap(mul(S, shared_key.a) + gauss(l, N))
        # BootstrapKeyPart (mk_internals.jl:419-439): n uni-encryptions (RGSW.UniEnc, :185-227)
        msg = secret_key.key.astype(np.int64)                                             # [n]
        gadget = np.array([1 << (32 - (r + 1) * p.bs_log2_base) for r in range(l)], dtype=np.int64)
        mg = msg[:, None] * gadget[None, :]                                               # [n][l]
        r = rand_uniform_bool(rng, n, 1, N)                                               # :195
        c1 = rand_uniform_torus32(rng, n, l, N)                                           # :198
        c0 = mul(S, c1) + gauss(n, l, N); c0[:, :, 0] += mg                               # :200-204
        d1 = mul(np.broadcast_to(r, (n, l, N)), np.broadcast_to(shared_key.a, (n, l, N)).copy()) + gauss(n, l, N)
        d1[:, :, 0] += mg                                                                 # :207-211
        d0 = mul(np.broadcast_to(r, (n, l, N)), np.broadcast_to(self.public_b, (n, l, N)).copy()) + gauss(n, l, N)   # :212-215
        f1 = rand_uniform_torus32(rng, n, l, N)                                           # :218
        f0 = mul(S, f1) + gauss(n, l, N) + r.astype(np.int64) * gadget[None, :, None]     # :220-224
        self.uni_enc = {"c0": _wrap(c0), "c1": c1, "d0": _wrap(d0), "d1": _wrap(d1), "f0": _wrap(f0), "f1": f1}
        self.ks = _keyswitch_key(rng, p.ks_noise_stddev, p.ks_decomp_length, p.ks_log2_base, secret_key.key, S)
        self.device = device


def _decompose(x: np.ndarray, l: int, bgbit: int) -> np.ndarray:
    """tgsw.jl:99-117 on the host (key expansion only): [...][N] -> [l][...][N]."""
    offset = sum(1 << (32 - r * bgbit) for r in range(1, l + 1)) * (1 << (bgbit - 1))
    v = (x.astype(np.int64) + offset) & 0xFFFFFFFF
    return np.stack([((v >> (32 - r * bgbit)) & ((1 << bgbit) - 1)) - (1 << (bgbit - 1)) for r in range(1, l + 1)]).astype(np.int32)


class MKCloudKey:                             # mk_api.jl:85-101
    def __init__(self, ck_parts: Sequence[CloudKeyPart], device: Optional[int] = None, flags: int = _cabi.FLAG_SPLIT_FFT,
                 host_expand: bool = False):
        """The key expansion (RGSW.Expand, mk_internals.jl:304-345: n*p*(p-1)*2*l^2 polynomial products), the transform
        and the load run on the device in one call (tfhe_b200_mk_expand_load_bk); ``host_expand=True`` keeps the
        round-1 path (numpy loops around batched GPU products) for comparison."""
        params = ck_parts[0].params
        parties = len(ck_parts)
        assert parties <= params.max_parties                                              # mk_api.jl:94
        device = ck_parts[0].device if device is None else device
        self.parties, self.params = parties, params
        if not host_expand:
            names = ("c0", "c1", "d0", "d1", "f0", "f1")
            uni_enc = np.stack([np.stack([part.uni_enc[k] for k in names]) for part in ck_parts])
            public_b = np.stack([part.public_b for part in ck_parts])
            self.keyswitch_key = np.stack([part.ks for part in ck_parts])
            self.ctx = _context(params, parties, device, flags)
            self.bootstrap_key = self.ctx.mk_expand_load_bk(uni_enc, public_b)
            self.ctx.load_ksk(self.keyswitch_key)
            return
        kctx = _mk_ctx(params, device)
        l, N, n, bgbit = params.bs_decomp_length, params.tlwe_polynomial_degree, params.lwe_size, params.bs_log2_base
        p = parties
        # MKBootstrapKey (mk_internals.jl:442-461): RGSW.Expand of every uni-encryption (:304-345)
        bk = np.empty((p, n, l * (2 * p + 2), N), dtype=np.int32)
        for i, part in enumerate(ck_parts):
            ue = part.uni_enc
            x = np.empty((n, l, p, N), dtype=np.int64); y = np.empty((n, l, p, N), dtype=np.int64)
            for ii, other in enumerate(ck_parts):
                x[:, :, ii] = ue["d0"]                                                    # :327
                if ii == i:
                    y[:, :, ii] = ue["d1"]                                                # :336
                    continue
                u = _decompose(_wrap(other.public_b.astype(np.int64) - part.public_b), l, bgbit)   # [r][jj][N]  :321
                u = np.transpose(u, (1, 0, 2))                                            # [jj][r][N]
                ub = np.broadcast_to(u[None], (n, l, l, N)).reshape(-1, N)
                for name, dst in (("f0", x), ("f1", y)):                                  # :330, :338
                    f = np.broadcast_to(ue[name][:, None], (n, l, l, N)).reshape(-1, N)
                    prod = kctx.polymul(ub, f).reshape(n, l, l, N).astype(np.int64).sum(axis=2)
                    if name == "f0": dst[:, :, ii] += prod
                    else: dst[:, :, ii] = prod
            bk[i, :, : l * p] = _wrap(x).reshape(n, l * p, N)
            bk[i, :, l * p: 2 * l * p] = _wrap(y).reshape(n, l * p, N)
            bk[i, :, 2 * l * p: 2 * l * p + l] = ue["c0"]
            bk[i, :, 2 * l * p + l:] = ue["c1"]
        self.bootstrap_key = bk
        self.keyswitch_key = np.stack([part.ks for part in ck_parts])
        self.ctx = _context(params, parties, device, flags)
        self.ctx.load_bk(bk)
        self.ctx.load_ksk(self.keyswitch_key)


def mk_encrypt(rng, secret_keys: Sequence[SecretKey], message) -> MKLweSample:           # mk_api.jl:110-126
    m = np.asarray(message, dtype=bool)
    params = secret_keys[0].params
    mu = np.where(m, encode_message(1, 8), encode_message(-1, 8)).astype(np.int64)
    keys = np.concatenate([sk.key for sk in secret_keys])
    a = rand_uniform_torus32(rng, *m.shape, keys.size)
    b = _wrap(mu + dtot32(rng.standard_normal(size=m.shape) * params.lwe_noise_stddev).astype(np.int64) + _dot(a, keys))
    return MKLweSample(np.concatenate([a, np.asarray(b)[..., None]], axis=-1), len(secret_keys), params.lwe_noise_stddev ** 2)


def mk_decrypt(secret_keys: Sequence[SecretKey], sample: MKLweSample):                   # mk_api.jl:135-138
    keys = np.concatenate([sk.key for sk in secret_keys])
    r = _wrap(sample.b.astype(np.int64) - _dot(sample.data[..., :-1], keys)) > 0          # mk_internals.jl:29-35
    return bool(r) if np.ndim(r) == 0 else r


def mk_gate_nand(ck: MKCloudKey, x: MKLweSample, y: MKLweSample) -> MKLweSample:         # mk_gates.jl:7-12
    single = x.data.ndim == 1
    out = ck.ctx.mk_nand(np.atleast_2d(x.data), np.atleast_2d(y.data))
    return MKLweSample(out[0] if single else out, ck.parties, 0.0)


# ------------------------------------------------------------------ device-resident batches (levelised circuits)
class DeviceLweBatch:
    """A batch of LWE ciphertexts that STAYS in HBM between gates: dependent circuits such as
    examples/tutorial.jl or a ripple-carry adder run level by level without a host round trip per gate.
    `tensor` is a CUDA int32 torch tensor of shape [count][n+1] (PyTorch is only the device-memory plumbing)."""

    def __init__(self, tensor):
        self.tensor = tensor

    @classmethod
    def from_host(cls, sample: LweSample):
        import torch
        return cls(torch.from_numpy(np.ascontiguousarray(np.atleast_2d(sample.data))).cuda())

    def to_host(self) -> LweSample:
        return LweSample(self.tensor.cpu().numpy(), 0.0)

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, idx):
        t = self.tensor[idx]
        return DeviceLweBatch(t if t.dim() == 2 else t[None, :])

    def repeat(self, count: int):
        return DeviceLweBatch(self.tensor.repeat(count, 1).contiguous())


def gate_dev(ck: CloudKey, op: int, *xs: DeviceLweBatch) -> DeviceLweBatch:
    """Any gate of gates.jl on device-resident operands: one asynchronous C-ABI call on torch's current stream."""
    import torch
    ts = [x.tensor.contiguous() for x in xs]
    out = torch.empty_like(ts[0])
    ptrs = [t.data_ptr() for t in ts] + [0] * (3 - len(ts))
    ck.ctx.gate_dev(op, ptrs[0], ptrs[1], ptrs[2], out.data_ptr(), ts[0].shape[0], stream=torch.cuda.current_stream().cuda_stream)
    return DeviceLweBatch(out)


def constant_dev(ck: CloudKey, values) -> DeviceLweBatch:
    """gate_constant (gates.jl:91-93) producing a device-resident batch."""
    import torch
    v = np.atleast_1d(np.asarray(values, dtype=bool))
    flags = torch.zeros((v.size, ck.params.lwe_size + 1), dtype=torch.int32)
    flags[:, 0] = torch.from_numpy(v.astype(np.int32))
    return gate_dev(ck, _cabi.CONSTANT, DeviceLweBatch(flags.cuda()))


# ------------------------------------------------------------------ binary key / ciphertext files
# The reference has no wire format.  These are numpy `.npz` archives holding the int32 arrays in exactly the
# layouts of the C ABI (include/tfhe_b200.h), plus the parameter tuple, so a key generated once — here, or exported
# from a real TFHE.jl install by julia/crosscheck.jl's conversion — can be reloaded onto any GPU without keygen.
_FORMAT = "tfhe-b200/1"


def _params_array(p: SchemeParameters) -> np.ndarray:
    return np.array([p.lwe_size, p.lwe_noise_stddev, p.tlwe_polynomial_degree, p.tlwe_mask_size, p.bs_decomp_length,
                     p.bs_log2_base, p.bs_noise_stddev, p.ks_decomp_length, p.ks_log2_base, p.ks_noise_stddev,
                     p.max_parties], dtype=np.float64)


def _params_from(a: np.ndarray) -> SchemeParameters:
    f = [float(v) for v in a]
    return SchemeParameters(int(f[0]), f[1], int(f[2]), int(f[3]), int(f[4]), int(f[5]), f[6], int(f[7]), int(f[8]), f[9], int(f[10]))


def _check_format(z, kind: str):
    if str(z["format"]) != _FORMAT or str(z["kind"]) != kind:
        raise ValueError(f"not a {kind} file of format {_FORMAT}")


def save_secret_key(path, key: SecretKey):
    np.savez(path, format=_FORMAT, kind="secret_key", params=_params_array(key.params), key=key.key.astype(np.int32))


def load_secret_key(path) -> SecretKey:
    with np.load(path) as z:
        _check_format(z, "secret_key")
        sk = SecretKey.__new__(SecretKey)
        sk.params, sk.key = _params_from(z["params"]), z["key"].astype(np.int32)
        return sk


def save_cloud_key(path, ck: CloudKey):
    np.savez(path, format=_FORMAT, kind="cloud_key", params=_params_array(ck.params),
             bootstrap_key=ck.bootstrap_key, keyswitch_key=ck.keyswitch_key)


def load_cloud_key(path, device: int = 0, flags: int = _cabi.FLAG_SPLIT_FFT) -> CloudKey:
    """Creates the GPU context and loads (transforms) the key on `device`; no key generation."""
    with np.load(path) as z:
        _check_format(z, "cloud_key")
        ck = CloudKey.__new__(CloudKey)
        ck.params = _params_from(z["params"])
        ck.bootstrap_key = np.ascontiguousarray(z["bootstrap_key"], dtype=np.int32)
        ck.keyswitch_key = np.ascontiguousarray(z["keyswitch_key"], dtype=np.int32)
    ck.mctx = None
    ck.ctx = _context(ck.params, 1, device, flags)
    ck.ctx.load_bk(ck.bootstrap_key)
    ck.ctx.load_ksk(ck.keyswitch_key)
    return ck


def save_ciphertext(path, sample: LweSample):
    np.savez(path, format=_FORMAT, kind="lwe", data=sample.data.astype(np.int32), variance=np.float64(sample.current_variance))


def load_ciphertext(path) -> LweSample:
    with np.load(path) as z:
        _check_format(z, "lwe")
        return LweSample(z["data"].astype(np.int32), float(z["variance"]))
