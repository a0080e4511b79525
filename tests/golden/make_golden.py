"""Generates tests/golden/*.npz with the CPU oracle (exact-integer route wherever one exists).

The reference ships no golden ciphertexts and Julia is not installed, so these vectors pin
the ORACLE's arithmetic: they are cross-checked at generation time against an independent big-integer
convolution, and tests/test_golden.py checks both the oracle (CPU) and the CUDA kernels (GPU) against them.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

N = 1024


def torus(rng, *shape):
    return rng.integers(-(2 ** 31), 2 ** 31, size=shape, dtype=np.int64).astype(np.int32)


def bigint_negacyclic(x, y):
    full = np.convolve(x.astype(object), y.astype(object))
    out = [int(full[i]) - (int(full[i + N]) if i + N < len(full) else 0) for i in range(N)]
    return np.array([((v + 2 ** 31) % 2 ** 32) - 2 ** 31 for v in out], dtype=np.int64).astype(np.int32)


def main():
    rng = np.random.default_rng(20261018)
    # 1. polynomial products (transformed_mul, polynomials.jl:142-144)
    x = torus(rng, 6, N); y = torus(rng, 6, N)
    x[0] = rng.integers(0, 2, N); x[1] = rng.integers(-512, 512, N); x[2] = 2 ** 31 - 1; y[2] = -(2 ** 31)
    x[3] = 0; x[3, 1] = 1                                           # multiplication by X
    prod = np.stack([O.polymul(x[i], y[i], O.ROUTE_EXACT) for i in range(6)])
    for i in range(6):
        assert np.array_equal(prod[i], bigint_negacyclic(x[i], y[i]))
    np.savez_compressed(os.path.join(HERE, "polymul.npz"), x=x, y=y, prod=prod)

    # 2. decompose (tgsw.jl:99-117), modswitch (numeric-functions.jl:31-34), monomial rotation, extraction
    p = torus(rng, N)
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), p=p,
                        dec_2_10=O.decompose(p, 2, 10), dec_3_7=O.decompose(p, 3, 7), dec_4_7=O.decompose(p, 4, 7),
                        dec_8_4=O.decompose(p, 8, 4), modswitch=O.decode_message(p, 2 * N),
                        rot_5=O.mul_by_monomial(p, 5), rot_m700=O.mul_by_monomial(p, -700), rot_1500=O.mul_by_monomial(p, 1500),
                        reverse=O.reverse_polynomial(p))

    # 3. a tiny key set (80-bit TGSW parameters, n = 6): external product, blind rotation, bootstrap, gates.
    #    Keys are regenerated from the seed by the oracle's deterministic keygen; a checksum pins them.
    P = O.small_params(O.PARAMS_80, 6)
    keys = O.keygen(P, 4242)
    ctx = O.Context(keys)
    acc = torus(rng, 2, 2, N)
    ext = np.stack([ctx.extern_mul(i, acc[i], O.ROUTE_EXACT) for i in range(2)])
    bara = rng.integers(-N, N, (2, 6)).astype(np.int32); bara[1, 2] = 0
    br = np.stack([ctx.blind_rotate(acc[i], bara[i], O.ROUTE_EXACT) for i in range(2)])
    bits = np.array([[0, 0, 1], [0, 1, 1], [1, 0, 0], [1, 1, 0]], dtype=bool)
    r = O.Rng(99)
    cts = np.stack([O.encrypt(r, keys, bits[:, i]) for i in range(3)])
    u = ctx.bootstrap_wo_ks(cts[0], route=O.ROUTE_EXACT)
    gates = {O.GATE_NAMES[op]: ctx.gate(op, cts[0], cts[1], cts[2] if op == O.MUX else None, route=O.ROUTE_EXACT, nthreads=4)
             for op in (O.NAND, O.XOR, O.ORNY, O.MUX)}
    np.savez_compressed(os.path.join(HERE, "tiny_bootstrap.npz"), seed=4242, n=6,
                        key_checksum=np.array([int(keys.bk.astype(np.int64).sum() % 2 ** 31), int(keys.ksk.astype(np.int64).sum() % 2 ** 31)]),
                        acc=acc, ext_index=np.arange(2, dtype=np.int32), ext=ext, bara=bara, blind_rotate=br,
                        bits=bits, cts=cts, bootstrap_wo_ks=u, keyswitch=ctx.keyswitch(u),
                        **{"gate_" + k: v for k, v in gates.items()})

    # 4. multi-key: one external product per party (mk_internals.jl:348-391), exact route, 2 parties, n = 3
    PM = O.small_params(O.MK_PARAMS[2], 3)
    mk = O.mk_keygen(PM, 2, 777)
    mctx = O.MKContext(mk)
    macc = torus(rng, 2, 3, N)
    mext = np.stack([mctx.extern_mul(i, 1, macc[i], O.ROUTE_EXACT) for i in range(2)])
    r = O.Rng(5)
    mx, my = O.mk_encrypt(r, mk, [True, False]), O.mk_encrypt(r, mk, [True, True])
    np.savez_compressed(os.path.join(HERE, "tiny_mk.npz"), seed=777, n=3, parties=2,
                        key_checksum=np.array([int(mk.bk.astype(np.int64).sum() % 2 ** 31)]),
                        acc=macc, ext=mext, x=mx, y=my, nand=mctx.nand(mx, my, route=O.ROUTE_EXACT))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
