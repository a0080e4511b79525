#!/usr/bin/env python
"""The workload of the reference's examples/tutorial.jl on the B200 engine: the minimum of two encrypted 16-bit
integers (2017 and 42), through the host mirror of the TFHE.jl API (`tfhe_jl_b200`).

Three ways to evaluate the same circuit, all giving the same ciphertexts:
  1. gate by gate, as the reference runs it (one library call per gate; tutorial.jl:42-62);
  2. levelised (`Circuit.run`): all gates of a level in one call, every intermediate ciphertext resident in HBM;
  3. the levelised evaluation captured as ONE CUDA graph (`Circuit.compile`).

    python examples/tutorial.py            # needs a B200 (the library has no CPU fallback)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfhe_jl_b200 as T  # noqa: E402


def int_to_bits(x, nbits=16):
    return np.array([(x >> i) & 1 for i in range(nbits)], dtype=bool)


def bits_to_int(bits):
    return sum(int(b) << i for i, b in enumerate(bits))


def encrypted_minimum(ck, a, b):
    """lsb-first comparator chain (carry = a < b so far), then 16 MUXes that pick the smaller operand."""
    carry = T.gate_constant(ck, False)
    for i in range(len(a)):
        carry = T.gate_mux(ck, T.gate_xnor(ck, a[i], b[i]), carry, a[i])
    select = T.LweSample(np.repeat(carry.data[None, :], len(a), axis=0))
    return T.gate_mux(ck, select, b, a)          # 16 independent MUXes: ONE call


def main():
    rng = np.random.default_rng(123)
    secret_key, cloud_key = T.make_key_pair(rng)
    ciphertext1 = T.encrypt(rng, secret_key, int_to_bits(2017))
    ciphertext2 = T.encrypt(rng, secret_key, int_to_bits(42))

    t0 = time.perf_counter()
    answer = encrypted_minimum(cloud_key, ciphertext1, ciphertext2)
    t1 = time.perf_counter()
    result = bits_to_int(T.decrypt(secret_key, answer))
    print(f"Answer: {result}   (gate by gate, {1e3 * (t1 - t0):.1f} ms)")
    assert result == 42

    circuit = T.minimum_circuit(16)
    inputs = {"a": ciphertext1, "b": ciphertext2}
    circuit.run(cloud_key, inputs)                # warm-up: scratch buffers reach their final size
    t0 = time.perf_counter()
    out = circuit.run(cloud_key, inputs)["min"].to_host()
    t1 = time.perf_counter()
    print(f"Answer: {bits_to_int(T.decrypt(secret_key, out))}   (levelised, {circuit.depth} levels, {1e3 * (t1 - t0):.1f} ms)")
    assert np.array_equal(out.data, answer.data)  # the same ciphertexts, not only the same plaintext

    compiled = circuit.compile(cloud_key)
    best = float("inf")
    for _ in range(4):                            # the first replays also upload the graph
        t0 = time.perf_counter()
        out = compiled.run(inputs)["min"].to_host()
        best = min(best, time.perf_counter() - t0)
    print(f"Answer: {bits_to_int(T.decrypt(secret_key, out))}   (one CUDA graph, {1e3 * best:.1f} ms)")
    assert np.array_equal(out.data, answer.data)
    return result


if __name__ == "__main__":
    main()
