"""ctypes binding of libtfhe_b200.so — the same symbols a Julia ``ccall`` wrapper binds
(include/tfhe_b200.h).  There is no fallback: if the CUDA library is missing or
no GPU is present every compute call raises ``TFHEB200Error``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtfhe_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "tfhe_b200.h")

OK, EINVAL, ENODEV, ECUDA, ENOKEY, ENOMEM = 0, -1, -2, -3, -4, -5
FLAG_SPLIT_FFT, FLAG_UNSPLIT_FFT = 0, 1

NAND, OR, AND, XOR, XNOR, NOT, CONSTANT, NOR, ANDNY, ANDYN, ORNY, ORYN, MUX = range(13)


class TFHEB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"tfhe_b200 error {code}: {msg}")
        self.code = code


class CParams(C.Structure):
    _fields_ = [(f, C.c_int32) for f in ("n", "N", "k", "l", "bgbit", "t", "basebit", "parties")]


def build(force: bool = False) -> str:
    """Compile the library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".inc"))] + [HEADER_PATH]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        r = subprocess.run(["make", "-C", csrc] + (["-B"] if force else []), capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libtfhe_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


_lib = None


def lib():
    """Load the C-ABI library (never builds implicitly: a missing library is an error)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TFHEB200Error(ENODEV, f"{LIB_PATH} not built (run __graft_entry__.build()); no CPU fallback exists")
        L = C.CDLL(LIB_PATH)
        i32p, vp, sz = C.c_void_p, C.c_void_p, C.c_size_t   # int32 pointers are passed as raw addresses
        sig = {
            "tfhe_b200_device_count": (C.c_int, []),
            "tfhe_b200_create": (C.c_int, [C.POINTER(CParams), C.c_int, C.c_uint32, C.POINTER(vp)]),
            "tfhe_b200_destroy": (None, [vp]),
            "tfhe_b200_last_error": (C.c_char_p, [vp]),
            "tfhe_b200_kernel_launches": (C.c_uint64, [vp]),
            "tfhe_b200_synchronize": (C.c_int, [vp]),
            "tfhe_b200_measure_fp64_tflops": (C.c_int, [vp, C.POINTER(C.c_double)]),
            "tfhe_b200_measure_lds_gbps": (C.c_int, [vp, C.POINTER(C.c_double)]),
            "tfhe_b200_load_bk": (C.c_int, [vp, i32p]),
            "tfhe_b200_load_ksk": (C.c_int, [vp, i32p]),
            "tfhe_b200_gate_batch": (C.c_int, [vp, C.c_int, i32p, i32p, i32p, i32p, sz]),
            "tfhe_b200_bootstrap_batch": (C.c_int, [vp, C.c_int32, i32p, i32p, sz]),
            "tfhe_b200_bootstrap_wo_ks_batch": (C.c_int, [vp, C.c_int32, i32p, i32p, sz]),
            "tfhe_b200_keyswitch_batch": (C.c_int, [vp, i32p, i32p, sz]),
            "tfhe_b200_extern_product_batch": (C.c_int, [vp, i32p, i32p, i32p, sz]),
            "tfhe_b200_blind_rotate_batch": (C.c_int, [vp, i32p, i32p, C.c_int32, i32p, sz]),
            "tfhe_b200_polymul_batch": (C.c_int, [vp, i32p, i32p, i32p, sz]),
            "tfhe_b200_gate_batch_dev": (C.c_int, [vp, C.c_int, i32p, i32p, i32p, i32p, sz, vp]),
            "tfhe_b200_bootstrap_wo_ks_batch_dev": (C.c_int, [vp, C.c_int32, i32p, i32p, sz, vp]),
            "tfhe_b200_keyswitch_batch_dev": (C.c_int, [vp, i32p, i32p, sz, vp]),
            "tfhe_b200_mk_load_bk": (C.c_int, [vp, i32p]),
            "tfhe_b200_mk_load_ksk": (C.c_int, [vp, i32p]),
            "tfhe_b200_mk_nand_batch": (C.c_int, [vp, i32p, i32p, i32p, sz]),
            "tfhe_b200_mk_nand_batch_dev": (C.c_int, [vp, i32p, i32p, i32p, sz, vp]),
            "tfhe_b200_mk_bootstrap_wo_ks_batch": (C.c_int, [vp, C.c_int32, i32p, i32p, sz]),
            "tfhe_b200_mk_keyswitch_batch": (C.c_int, [vp, i32p, i32p, sz]),
            "tfhe_b200_mk_extern_product_batch": (C.c_int, [vp, i32p, i32p, i32p, i32p, sz]),
            "tfhe_b200_mk_bootstrap_batch": (C.c_int, [vp, C.c_int32, i32p, i32p, sz]),
            "tfhe_b200_mk_bootstrap_batch_dev": (C.c_int, [vp, C.c_int32, i32p, i32p, sz, vp]),
            "tfhe_b200_random_words": (C.c_int, [vp, C.c_uint64, C.c_uint64, i32p, sz]),
            "tfhe_b200_lwe_encrypt_words_batch": (C.c_int, [vp, i32p, C.c_int32, i32p, i32p, i32p, i32p, sz]),
            "tfhe_b200_encrypt_batch": (C.c_int, [vp, i32p, C.c_int32, vp, C.c_double, C.c_uint64, i32p, sz]),
            "tfhe_b200_encrypt_batch_dev": (C.c_int, [vp, i32p, C.c_int32, vp, C.c_double, C.c_uint64, i32p, sz, vp]),
            "tfhe_b200_lwe_phase_batch": (C.c_int, [vp, i32p, C.c_int32, i32p, i32p, vp, sz]),
            "tfhe_b200_keygen_bk_words": (C.c_int, [vp, i32p, i32p, i32p, i32p, i32p]),
            "tfhe_b200_keygen_bk": (C.c_int, [vp, i32p, i32p, C.c_double, C.c_uint64, i32p]),
            "tfhe_b200_keygen_ksk_words": (C.c_int, [vp, i32p, i32p, i32p, i32p, i32p]),
            "tfhe_b200_keygen_ksk": (C.c_int, [vp, i32p, i32p, C.c_double, C.c_uint64, i32p]),
            "tfhe_b200_mk_expand_load_bk": (C.c_int, [vp, i32p, i32p, i32p]),
            "tfhe_b200_multi_create": (C.c_int, [C.POINTER(CParams), C.POINTER(C.c_int), C.c_int, C.c_uint32, C.POINTER(vp)]),
            "tfhe_b200_multi_destroy": (None, [vp]),
            "tfhe_b200_multi_last_error": (C.c_char_p, [vp]),
            "tfhe_b200_multi_devices": (C.c_int, [vp]),
            "tfhe_b200_multi_context": (vp, [vp, C.c_int]),
            "tfhe_b200_multi_kernel_launches": (C.c_uint64, [vp]),
            "tfhe_b200_multi_load_bk": (C.c_int, [vp, i32p]),
            "tfhe_b200_multi_load_ksk": (C.c_int, [vp, i32p]),
            "tfhe_b200_multi_mk_load_bk": (C.c_int, [vp, i32p]),
            "tfhe_b200_multi_mk_load_ksk": (C.c_int, [vp, i32p]),
            "tfhe_b200_multi_gate_batch": (C.c_int, [vp, C.c_int, i32p, i32p, i32p, i32p, sz]),
            "tfhe_b200_multi_bootstrap_batch": (C.c_int, [vp, C.c_int32, i32p, i32p, sz]),
            "tfhe_b200_multi_mk_nand_batch": (C.c_int, [vp, i32p, i32p, i32p, sz]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        L._signatures = sig
        _lib = L
    return _lib


def _host(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.int32)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


def _addr(a):
    return None if a is None else a.ctypes.data


class Context:
    """One parameter set + one evaluation-key set on one GPU (CloudKey / MKCloudKey device side)."""

    def __init__(self, n, N=1024, k=1, l=2, bgbit=10, t=8, basebit=2, parties=1, device=0, flags=FLAG_SPLIT_FFT):
        self.n, self.N, self.k, self.l, self.bgbit, self.t, self.basebit, self.parties = n, N, k, l, bgbit, t, basebit, parties
        self.device, self.flags = device, flags
        self._h = C.c_void_p()
        cp = CParams(n, N, k, l, bgbit, t, basebit, parties)
        rc = lib().tfhe_b200_create(C.byref(cp), device, flags, C.byref(self._h))
        if rc != 0:
            raise TFHEB200Error(rc, (lib().tfhe_b200_last_error(None) or b"").decode())

    @classmethod
    def _borrowed(cls, handle, owner, n, N, k, l, bgbit, t, basebit, parties, device, flags):
        """A view of a context owned by `owner` (a MultiContext): never destroyed from here."""
        self = cls.__new__(cls)
        self.n, self.N, self.k, self.l, self.bgbit, self.t, self.basebit, self.parties = n, N, k, l, bgbit, t, basebit, parties
        self.device, self.flags = device, flags
        self._h = C.c_void_p(handle)
        self._owner = owner
        return self

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            if getattr(self, "_owner", None) is None:
                lib().tfhe_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:   # interpreter shutdown: module globals may already be gone
            pass

    def _ck(self, rc):
        if rc != 0:
            raise TFHEB200Error(rc, (lib().tfhe_b200_last_error(self._h) or b"").decode())

    @property
    def kernel_launches(self) -> int:
        return int(lib().tfhe_b200_kernel_launches(self._h))

    def synchronize(self):
        self._ck(lib().tfhe_b200_synchronize(self._h))

    def measure_fp64_tflops(self) -> float:
        v = C.c_double()
        self._ck(lib().tfhe_b200_measure_fp64_tflops(self._h, C.byref(v)))
        return v.value

    def measure_lds_gbps(self) -> float:
        v = C.c_double()
        self._ck(lib().tfhe_b200_measure_lds_gbps(self._h, C.byref(v)))
        return v.value

    # widths
    @property
    def ct_words(self):
        return self.parties * self.n + 1

    @property
    def ext_words(self):
        return self.parties * self.N * self.k + 1

    # ---- keys
    def load_bk(self, bk):
        if self.parties == 1:
            bk = _host(bk).reshape(self.n, self.l, self.k + 1, self.k + 1, self.N)
            self._ck(lib().tfhe_b200_load_bk(self._h, _addr(bk)))
        else:
            p = self.parties
            bk = _host(bk).reshape(p, self.n, self.l * (2 * p + 2), self.N)
            self._ck(lib().tfhe_b200_mk_load_bk(self._h, _addr(bk)))

    def load_ksk(self, ksk):
        shape = (self.N * self.k, self.t, (1 << self.basebit) - 1, self.n + 1)
        if self.parties == 1:
            ksk = _host(ksk).reshape(shape)
            self._ck(lib().tfhe_b200_load_ksk(self._h, _addr(ksk)))
        else:
            ksk = _host(ksk).reshape((self.parties,) + shape)
            self._ck(lib().tfhe_b200_mk_load_ksk(self._h, _addr(ksk)))

    # ---- single-key, host buffers
    def gate(self, op, x=None, y=None, z=None, count=None):
        arrs = [None if a is None else np.atleast_2d(_host(a)) for a in (x, y, z)]
        if count is None:
            count = next(a.shape[0] for a in arrs if a is not None)
        for a in arrs:
            assert a is None or a.shape == (count, self.n + 1), a.shape
        out = np.empty((count, self.n + 1), dtype=np.int32)
        self._ck(lib().tfhe_b200_gate_batch(self._h, op, _addr(arrs[0]), _addr(arrs[1]), _addr(arrs[2]), _addr(out), count))
        return out

    def bootstrap(self, x, mu=1 << 29):
        x = np.atleast_2d(_host(x)); out = np.empty_like(x)
        self._ck(lib().tfhe_b200_bootstrap_batch(self._h, mu, _addr(x), _addr(out), x.shape[0]))
        return out

    def bootstrap_wo_ks(self, x, mu=1 << 29):
        x = np.atleast_2d(_host(x)); out = np.empty((x.shape[0], self.ext_words), dtype=np.int32)
        fn = lib().tfhe_b200_bootstrap_wo_ks_batch if self.parties == 1 else lib().tfhe_b200_mk_bootstrap_wo_ks_batch
        self._ck(fn(self._h, mu, _addr(x), _addr(out), x.shape[0]))
        return out

    def mk_bootstrap(self, x, mu=1 << 29):
        x = np.atleast_2d(_host(x)); out = np.empty_like(x)
        assert self.parties > 1 and x.shape[1] == self.ct_words
        self._ck(lib().tfhe_b200_mk_bootstrap_batch(self._h, mu, _addr(x), _addr(out), x.shape[0]))
        return out

    def keyswitch(self, u):
        u = np.atleast_2d(_host(u)); out = np.empty((u.shape[0], self.ct_words), dtype=np.int32)
        fn = lib().tfhe_b200_keyswitch_batch if self.parties == 1 else lib().tfhe_b200_mk_keyswitch_batch
        self._ck(fn(self._h, _addr(u), _addr(out), u.shape[0]))
        return out

    def extern_product(self, acc, bk_index, party=None):
        acc = _host(acc); count = acc.shape[0]; out = np.empty_like(acc)
        idx = _host(np.broadcast_to(np.asarray(bk_index, dtype=np.int32), (count,)))
        if self.parties == 1:
            assert acc.shape == (count, self.k + 1, self.N)
            self._ck(lib().tfhe_b200_extern_product_batch(self._h, _addr(acc), _addr(idx), _addr(out), count))
        else:
            assert acc.shape == (count, self.parties + 1, self.N)
            par = _host(np.broadcast_to(np.asarray(party, dtype=np.int32), (count,)))
            self._ck(lib().tfhe_b200_mk_extern_product_batch(self._h, _addr(acc), _addr(par), _addr(idx), _addr(out), count))
        return out

    def blind_rotate(self, acc, bara, n_iter=None):
        acc = _host(acc); count = acc.shape[0]; out = np.empty_like(acc)
        bara = _host(bara, (count, self.n))
        n_iter = self.n if n_iter is None else n_iter
        self._ck(lib().tfhe_b200_blind_rotate_batch(self._h, _addr(acc), _addr(bara), n_iter, _addr(out), count))
        return out

    def polymul(self, x, y):
        x = np.atleast_2d(_host(x)); y = np.atleast_2d(_host(y)); out = np.empty_like(x)
        assert x.shape == y.shape and x.shape[1] == self.N
        self._ck(lib().tfhe_b200_polymul_batch(self._h, _addr(x), _addr(y), _addr(out), x.shape[0]))
        return out

    def mk_nand(self, x, y):
        x = np.atleast_2d(_host(x)); y = np.atleast_2d(_host(y)); out = np.empty_like(x)
        assert x.shape == y.shape and x.shape[1] == self.ct_words
        self._ck(lib().tfhe_b200_mk_nand_batch(self._h, _addr(x), _addr(y), _addr(out), x.shape[0]))
        return out

    # ---- the steps either side of the path, on the device
    def random_words(self, seed, stream, count):
        out = np.empty(count, dtype=np.int32)
        self._ck(lib().tfhe_b200_random_words(self._h, seed, stream, _addr(out), count))
        return out

    def lwe_encrypt_words(self, key, mu, noise, a):
        key = _host(key); a = np.atleast_2d(_host(a)); mu = _host(mu).reshape(-1); noise = _host(noise).reshape(-1)
        count = a.shape[0]
        assert a.shape[1] == key.size and mu.size == count and noise.size == count
        out = np.empty((count, key.size + 1), dtype=np.int32)
        self._ck(lib().tfhe_b200_lwe_encrypt_words_batch(self._h, _addr(key), key.size, _addr(mu), _addr(noise), _addr(a), _addr(out), count))
        return out

    def encrypt(self, key, bits, sigma, seed):
        """encrypt (api.jl:155-158) of an array of bits; mask and noise are generated on the device from `seed`."""
        key = _host(key); bits = np.ascontiguousarray(np.asarray(bits, dtype=bool).reshape(-1), dtype=np.uint8)
        out = np.empty((bits.size, key.size + 1), dtype=np.int32)
        self._ck(lib().tfhe_b200_encrypt_batch(self._h, _addr(key), key.size, bits.ctypes.data, float(sigma), int(seed), _addr(out), bits.size))
        return out

    def encrypt_dev(self, key, bits, sigma, seed, out_ptr, stream=0):
        key = _host(key); bits = np.ascontiguousarray(np.asarray(bits, dtype=bool).reshape(-1), dtype=np.uint8)
        self._ck(lib().tfhe_b200_encrypt_batch_dev(self._h, _addr(key), key.size, bits.ctypes.data, float(sigma), int(seed), out_ptr,
                                                   bits.size, stream or None))

    def lwe_phase(self, key, ct):
        key = _host(key); ct = np.atleast_2d(_host(ct))
        assert ct.shape[1] == key.size + 1
        phase = np.empty(ct.shape[0], dtype=np.int32)
        self._ck(lib().tfhe_b200_lwe_phase_batch(self._h, _addr(key), key.size, _addr(ct), _addr(phase), None, ct.shape[0]))
        return phase

    def decrypt(self, key, ct):
        key = _host(key); ct = np.atleast_2d(_host(ct))
        assert ct.shape[1] == key.size + 1
        bits = np.empty(ct.shape[0], dtype=np.uint8)
        self._ck(lib().tfhe_b200_lwe_phase_batch(self._h, _addr(key), key.size, _addr(ct), None, bits.ctypes.data, ct.shape[0]))
        return bits.astype(bool)

    def keygen_bk(self, lwe_key, tlwe_key, sigma=None, seed=None, a=None, noise=None, keep=True):
        """Generates, transforms and loads the bootstrapping key on the device; returns its int32 coefficient form
        ([n][l][k+1][k+1][N]) when `keep`.  Either (sigma, seed) or explicit randomness (a: [n*l*(k+1)][k][N] mask words,
        noise: [n*l*(k+1)][N])."""
        lwe_key = _host(lwe_key, (self.n,)); tlwe_key = _host(tlwe_key).reshape(-1)
        assert tlwe_key.size == self.N * self.k and self.parties == 1
        k1 = self.k + 1
        out = np.empty((self.n, self.l, k1, k1, self.N), dtype=np.int32) if keep else None
        if a is not None:
            a = _host(np.reshape(a, (self.n * self.l * k1, self.k, self.N))); noise = _host(noise, (self.n * self.l * k1, self.N))
            self._ck(lib().tfhe_b200_keygen_bk_words(self._h, _addr(lwe_key), _addr(tlwe_key), _addr(a), _addr(noise), _addr(out)))
        else:
            self._ck(lib().tfhe_b200_keygen_bk(self._h, _addr(lwe_key), _addr(tlwe_key), float(sigma), int(seed), _addr(out)))
        return out

    def mk_expand_load_bk(self, uni_enc, public_b, keep=True):
        """RGSW.Expand + transform + load of the MK bootstrapping key on the device (mk_internals.jl:304-345, 442-461).
        uni_enc [p][6][n][l][N] (c0, c1, d0, d1, f0, f1), public_b [p][l][N]; returns [p][n][l*(2p+2)][N] when `keep`."""
        p, n, l, N = self.parties, self.n, self.l, self.N
        uni_enc = _host(uni_enc, (p, 6, n, l, N)); public_b = _host(public_b, (p, l, N))
        out = np.empty((p, n, l * (2 * p + 2), N), dtype=np.int32) if keep else None
        self._ck(lib().tfhe_b200_mk_expand_load_bk(self._h, _addr(uni_enc), _addr(public_b), _addr(out)))
        return out

    def keygen_ksk(self, out_key, in_key, sigma=None, seed=None, a=None, noise=None, keep=True):
        """Generates and loads the key-switching key on the device; returns [N*k][t][base-1][n+1] when `keep`."""
        out_key = _host(out_key, (self.n,)); in_key = _host(in_key).reshape(-1)
        base1 = (1 << self.basebit) - 1
        assert in_key.size == self.N * self.k
        out = np.empty((self.N * self.k, self.t, base1, self.n + 1), dtype=np.int32) if keep else None
        if a is not None:
            a = _host(a, (self.N * self.k, self.t, base1, self.n)); noise = _host(noise, (self.N * self.k, self.t, base1))
            self._ck(lib().tfhe_b200_keygen_ksk_words(self._h, _addr(out_key), _addr(in_key), _addr(a), _addr(noise), _addr(out)))
        else:
            self._ck(lib().tfhe_b200_keygen_ksk(self._h, _addr(out_key), _addr(in_key), float(sigma), int(seed), _addr(out)))
        return out

    # ---- device buffers (raw addresses, e.g. torch ``tensor.data_ptr()``); asynchronous on ``stream``
    def gate_dev(self, op, x_ptr, y_ptr, z_ptr, out_ptr, count, stream=0):
        self._ck(lib().tfhe_b200_gate_batch_dev(self._h, op, x_ptr or None, y_ptr or None, z_ptr or None, out_ptr, count,
                                                stream or None))

    def bootstrap_wo_ks_dev(self, x_ptr, out_ptr, count, mu=1 << 29, stream=0):
        self._ck(lib().tfhe_b200_bootstrap_wo_ks_batch_dev(self._h, mu, x_ptr, out_ptr, count, stream or None))

    def keyswitch_dev(self, in_ptr, out_ptr, count, stream=0):
        self._ck(lib().tfhe_b200_keyswitch_batch_dev(self._h, in_ptr, out_ptr, count, stream or None))

    def mk_nand_dev(self, x_ptr, y_ptr, out_ptr, count, stream=0):
        self._ck(lib().tfhe_b200_mk_nand_batch_dev(self._h, x_ptr, y_ptr, out_ptr, count, stream or None))

    def mk_bootstrap_dev(self, x_ptr, out_ptr, count, mu=1 << 29, stream=0):
        self._ck(lib().tfhe_b200_mk_bootstrap_batch_dev(self._h, mu, x_ptr, out_ptr, count, stream or None))


class MultiContext:
    """One logical evaluation context over several GPUs (tfhe_b200_multi_*): keys replicated at load, every batch cut
    into contiguous shards, one host thread per device, disjoint slices of the output.  ``devices=None`` = all."""

    def __init__(self, n, N=1024, k=1, l=2, bgbit=10, t=8, basebit=2, parties=1, devices=None, flags=FLAG_SPLIT_FFT):
        self.n, self.N, self.k, self.l, self.bgbit, self.t, self.basebit, self.parties = n, N, k, l, bgbit, t, basebit, parties
        self.flags = flags
        self._h = C.c_void_p()
        cp = CParams(n, N, k, l, bgbit, t, basebit, parties)
        ids = None if devices is None else (C.c_int * len(devices))(*devices)
        rc = lib().tfhe_b200_multi_create(C.byref(cp), ids, 0 if devices is None else len(devices), flags, C.byref(self._h))
        if rc != 0:
            raise TFHEB200Error(rc, (lib().tfhe_b200_multi_last_error(None) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().tfhe_b200_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise TFHEB200Error(rc, (lib().tfhe_b200_multi_last_error(self._h) or b"").decode())

    @property
    def devices(self) -> int:
        return int(lib().tfhe_b200_multi_devices(self._h))

    def context(self, i: int = 0) -> "Context":
        """The single-device context of the i-th listed device (device-resident / *_dev entry points)."""
        h = lib().tfhe_b200_multi_context(self._h, i)
        if not h:
            raise IndexError(i)
        return Context._borrowed(h, self, self.n, self.N, self.k, self.l, self.bgbit, self.t, self.basebit, self.parties, i, self.flags)

    @property
    def kernel_launches(self) -> int:
        return int(lib().tfhe_b200_multi_kernel_launches(self._h))

    @property
    def ct_words(self):
        return self.parties * self.n + 1

    def load_bk(self, bk):
        bk = _host(bk)
        self._ck((lib().tfhe_b200_multi_load_bk if self.parties == 1 else lib().tfhe_b200_multi_mk_load_bk)(self._h, _addr(bk)))

    def load_ksk(self, ksk):
        ksk = _host(ksk)
        self._ck((lib().tfhe_b200_multi_load_ksk if self.parties == 1 else lib().tfhe_b200_multi_mk_load_ksk)(self._h, _addr(ksk)))

    def gate(self, op, x=None, y=None, z=None, count=None, out=None):
        arrs = [None if a is None else np.atleast_2d(_host(a)) for a in (x, y, z)]
        if count is None:
            count = next(a.shape[0] for a in arrs if a is not None)
        if out is None:
            out = np.empty((count, self.ct_words), dtype=np.int32)
        self._ck(lib().tfhe_b200_multi_gate_batch(self._h, op, _addr(arrs[0]), _addr(arrs[1]), _addr(arrs[2]), _addr(out), count))
        return out

    def bootstrap(self, x, mu=1 << 29):
        x = np.atleast_2d(_host(x)); out = np.empty_like(x)
        self._ck(lib().tfhe_b200_multi_bootstrap_batch(self._h, mu, _addr(x), _addr(out), x.shape[0]))
        return out

    def mk_nand(self, x, y):
        x = np.atleast_2d(_host(x)); y = np.atleast_2d(_host(y)); out = np.empty_like(x)
        self._ck(lib().tfhe_b200_multi_mk_nand_batch(self._h, _addr(x), _addr(y), _addr(out), x.shape[0]))
        return out


def device_count() -> int:
    return int(lib().tfhe_b200_device_count())
