#!/usr/bin/env python
"""BASELINE.json configs[0] and configs[3]: dependent circuits with every intermediate ciphertext kept in HBM.

  tutorial   examples/tutorial.jl — 16-bit encrypted minimum (16 sequential XNOR+MUX levels, then 16 MUXes at once)
  adder32    32-bit ripple-carry adder (32 sequential carry levels, the sum XORs batched per level)

Prints one JSON line with wall-clock latency per circuit (device-synchronised), the sequential gate depth and, for comparison, the oracle's (CPU restatement of TFHE.jl) time for the tutorial circuit.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import tfhe_jl_b200 as T  # noqa: E402
from tfhe_jl_b200 import _cabi  # noqa: E402


def bits_of(v, n):
    return np.array([(v >> i) & 1 for i in range(n)], dtype=bool)


def value_of(bits):
    return sum(int(b) << i for i, b in enumerate(bits))


def tutorial(ck, a, b):
    carry = T.constant_dev(ck, False)
    for i in range(16):
        carry = T.gate_dev(ck, _cabi.MUX, T.gate_dev(ck, _cabi.XNOR, a[i], b[i]), carry, a[i])
    return T.gate_dev(ck, _cabi.MUX, carry.repeat(16), b, a)


def adder(ck, a, b, nbits):
    # level 0 for all bits at once: p = a xor b, g = a and b; then the carry chain (2 dependent gates per bit)
    p = T.gate_dev(ck, _cabi.XOR, a, b)
    g = T.gate_dev(ck, _cabi.AND, a, b)
    carry = T.constant_dev(ck, False)
    carries = [carry]
    for i in range(nbits - 1):
        carry = T.gate_dev(ck, _cabi.OR, g[i], T.gate_dev(ck, _cabi.AND, p[i], carry))
        carries.append(carry)
    cin = T.DeviceLweBatch(torch.cat([c.tensor for c in carries]))
    return T.gate_dev(ck, _cabi.XOR, p, cin)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return out, best * 1e3


def main():
    flags = int(os.environ.get("FLAGS", "0"))
    rng = np.random.default_rng(7)
    sk, ck = T.make_key_pair(rng, flags=flags)
    res = {"flags": flags}

    a, b = T.DeviceLweBatch.from_host(T.encrypt(rng, sk, bits_of(2017, 16))), T.DeviceLweBatch.from_host(T.encrypt(rng, sk, bits_of(42, 16)))
    out, ms = timed(lambda: tutorial(ck, a, b))
    assert value_of(T.decrypt(sk, out.to_host())) == 42
    res["tutorial_min16_ms"] = ms
    res["tutorial_gate_depth"] = 16 * 2 + 1

    x, y = 0xDEADBEEF, 0x12345678
    a, b = T.DeviceLweBatch.from_host(T.encrypt(rng, sk, bits_of(x, 32))), T.DeviceLweBatch.from_host(T.encrypt(rng, sk, bits_of(y, 32)))
    out, ms = timed(lambda: adder(ck, a, b, 32))
    assert value_of(T.decrypt(sk, out.to_host())) == (x + y) & 0xFFFFFFFF
    res["adder32_ms"] = ms
    res["adder32_gate_depth"] = 1 + 31 * 2 + 1

    # the same circuits through the levelised driver (tfhe.jl_b200/circuit.py): all ready gates of a level in one launch
    from tfhe_jl_b200.circuit import adder_circuit, minimum_circuit
    cm, ca = minimum_circuit(16), adder_circuit(32)
    ia = {"a": T.DeviceLweBatch.from_host(T.encrypt(rng, sk, bits_of(2017, 16))), "b": T.DeviceLweBatch.from_host(T.encrypt(rng, sk, bits_of(42, 16)))}
    out, ms = timed(lambda: cm.run(ck, ia)["min"])
    assert value_of(T.decrypt(sk, out.to_host())) == 42
    res["tutorial_min16_levelised_ms"], res["tutorial_levels"] = ms, cm.depth
    ib = {"a": a, "b": b}
    out, ms = timed(lambda: ca.run(ck, ib)["sum"])
    assert value_of(T.decrypt(sk, out.to_host())) == (x + y) & 0xFFFFFFFF
    res["adder32_levelised_ms"], res["adder32_levels"] = ms, ca.depth

    # ... and captured as one CUDA graph each (Circuit.compile)
    gm, ga = cm.compile(ck), ca.compile(ck)
    out, ms = timed(lambda: gm.run(ia)["min"])
    assert value_of(T.decrypt(sk, out.to_host())) == 42
    res["tutorial_min16_graph_ms"] = ms
    out, ms = timed(lambda: ga.run(ib)["sum"])
    assert value_of(T.decrypt(sk, out.to_host())) == (x + y) & 0xFFFFFFFF
    res["adder32_graph_ms"] = ms

    if os.environ.get("CPU", "1") == "1":
        from oracle import oracle as O
        keys = O.keygen(O.PARAMS_80, 5)
        octx = O.Context(keys)
        r = O.Rng(9)
        ea, eb = O.encrypt(r, keys, bits_of(2017, 16)), O.encrypt(r, keys, bits_of(42, 16))
        t0 = time.perf_counter()
        carry = octx.gate(O.CONSTANT, np.zeros((1, keys.params.n + 1), np.int32))
        for i in range(16):
            t = octx.gate(O.XNOR, ea[i:i + 1], eb[i:i + 1])
            carry = octx.gate(O.MUX, t, carry, ea[i:i + 1])
        o = octx.gate(O.MUX, np.repeat(carry, 16, 0), eb, ea, nthreads=os.cpu_count())
        res["tutorial_min16_cpu_port_ms"] = (time.perf_counter() - t0) * 1e3
        assert value_of(O.decrypt(keys, o)) == 42
    print(json.dumps(res))


if __name__ == "__main__":
    main()
