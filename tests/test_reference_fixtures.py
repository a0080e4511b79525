"""Reference-pinned parity.  tfhe.jl_b200/julia/crosscheck.jl, run where Julia and TFHE.jl are installed, exports keys
(MersenneTwister(123), test/runtests.jl:27), input ciphertexts and TFHE.jl's own output ciphertexts of every gate as
tests/golden/tfhejl/.  With those files present the oracle (CPU) and the CUDA path (GPU) must reproduce TFHE.jl's
ciphertexts bit for bit.  The build image has no Julia, so the directory does not exist yet and these tests SKIP —
which is exactly the "parity unpinned at ciphertext level" status DESIGN.md 5 reports."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tfhejl")
pinned = pytest.mark.skipif(not os.path.exists(os.path.join(DIR, "manifest.json")),
                            reason="parity unpinned: no TFHE.jl fixtures (run julia/crosscheck.jl <count> tests/golden/tfhejl on a Julia box)")


def load():
    m = json.load(open(os.path.join(DIR, "manifest.json")))
    rd = lambda name, *shape: np.fromfile(os.path.join(DIR, name + ".bin"), dtype="<i4").reshape(shape)
    n, N, k, l, t, c = m["n"], m["N"], m["k"], m["l"], m["t"], m["count"]
    base1 = (1 << m["basebit"]) - 1
    P = O.Params(n, O.PARAMS_80.lwe_sigma, N, k, l, m["bgbit"], O.PARAMS_80.bs_sigma, t, m["basebit"], O.PARAMS_80.ks_sigma, 1)
    keys = O.KeySet(P, rd("lwe_key", n), None, rd("bk", n, l, k + 1, k + 1, N), rd("ksk", N * k, t, base1, n + 1))
    cts = {name: rd(name, c, n + 1) for name in ("x", "y", "z")}
    outs = {g: rd("out_" + g, c, n + 1) for g in m["gates"]}
    return keys, cts, outs


@pinned
def test_oracle_reproduces_tfhejl_ciphertexts():
    keys, cts, outs = load()
    octx = O.Context(keys)
    for g, want in outs.items():
        op = O.GATE_NAMES.index(g)
        got = octx.gate(op, cts["x"], cts["y"], cts["z"] if g == "MUX" else None)
        assert np.array_equal(got, want), g


@pinned
@pytest.mark.gpu
def test_cuda_path_reproduces_tfhejl_ciphertexts():
    import tfhe_jl_b200 as T
    keys, cts, outs = load()
    P = keys.params
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit)
    ctx.load_bk(keys.bk); ctx.load_ksk(keys.ksk)
    for g, want in outs.items():
        op = O.GATE_NAMES.index(g)
        got = ctx.gate(op, cts["x"], cts["y"], cts["z"] if g == "MUX" else None)
        assert np.array_equal(got, want), g
