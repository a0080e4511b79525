"""BASELINE.json configs #1 and #4 as GPU parity cases: dependent gate circuits driven level by level through
the mirrored API (each level is one batched C-ABI call)."""
import numpy as np
import pytest

import tfhe_jl_b200 as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def keypair():
    rng = np.random.default_rng(123)
    return rng, *T.make_key_pair(rng)


def bits_of(v, n):
    return np.array([(v >> i) & 1 for i in range(n)], dtype=bool)


def test_tutorial_minimum_circuit(keypair):
    """examples/tutorial.jl: encrypted 16-bit minimum of 2017 and 42 -> 42 (80 blind rotations, depth 33)."""
    rng, sk, ck = keypair
    a, b = T.encrypt(rng, sk, bits_of(2017, 16)), T.encrypt(rng, sk, bits_of(42, 16))
    carry = T.gate_constant(ck, False)                                  # tutorial.jl:53
    for i in range(16):                                                 # tutorial.jl:55-57, 42-45
        carry = T.gate_mux(ck, T.gate_xnor(ck, a[i], b[i]), carry, a[i])
    sel = T.LweSample(np.repeat(carry.data[None, :], 16, axis=0))
    res = T.gate_mux(ck, sel, b, a)                                     # tutorial.jl:61 (16 independent MUXes, one call)
    assert sum(int(v) << i for i, v in enumerate(T.decrypt(sk, res))) == 42


def test_ripple_carry_adder_32bit(keypair):
    """32-bit ripple-carry adder over encrypted bits (not in the reference; built from its gates):
    sum_i = a_i ^ b_i ^ c_i, c_{i+1} = MUX(a_i ^ b_i, c_i, a_i).  Several additions run side by side as a batch."""
    rng, sk, ck = keypair
    xs = np.array([0xDEADBEEF, 123456789, 0xFFFFFFFF, 0], dtype=np.uint64)
    ys = np.array([0x12345678, 987654321, 1, 0], dtype=np.uint64)
    A = [T.encrypt(rng, sk, np.array([(int(x) >> i) & 1 for x in xs], dtype=bool)) for i in range(32)]
    B = [T.encrypt(rng, sk, np.array([(int(y) >> i) & 1 for y in ys], dtype=bool)) for i in range(32)]
    carry = T.gate_constant(ck, np.zeros(len(xs), dtype=bool))
    out = np.zeros(len(xs), dtype=np.uint64)
    for i in range(32):
        axb = T.gate_xor(ck, A[i], B[i])
        s = T.gate_xor(ck, axb, carry)
        carry = T.gate_mux(ck, axb, carry, A[i])
        out |= T.decrypt(sk, s).astype(np.uint64) << np.uint64(i)
    assert np.array_equal(out, (xs + ys) & np.uint64(0xFFFFFFFF))


def test_tutorial_circuit_device_resident(keypair):
    """The same minimum circuit with every intermediate ciphertext kept in HBM (DeviceLweBatch): only the inputs
    go up and only the 16 result bits come back.  Must agree bit for bit with the host-driven run."""
    from tfhe_jl_b200 import _cabi
    rng, sk, ck = keypair
    ha, hb = T.encrypt(rng, sk, bits_of(2017, 16)), T.encrypt(rng, sk, bits_of(42, 16))
    a, b = T.DeviceLweBatch.from_host(ha), T.DeviceLweBatch.from_host(hb)
    carry = T.constant_dev(ck, False)
    hcarry = T.gate_constant(ck, False)
    for i in range(16):
        carry = T.gate_dev(ck, _cabi.MUX, T.gate_dev(ck, _cabi.XNOR, a[i], b[i]), carry, a[i])
        if i < 2:   # spot-check against the host-driven path
            hcarry = T.gate_mux(ck, T.gate_xnor(ck, ha[i], hb[i]), hcarry, ha[i])
            assert np.array_equal(carry.to_host().data[0], hcarry.data)
    res = T.gate_dev(ck, _cabi.MUX, carry.repeat(16), b, a).to_host()
    assert sum(int(v) << i for i, v in enumerate(T.decrypt(sk, res))) == 42


def test_levelised_minimum_circuit_equals_gate_by_gate(keypair):
    """circuit.py batches all gates of a level into one launch; the ciphertexts must be the ones a gate-by-gate
    evaluation (the reference's order) produces."""
    from tfhe_jl_b200 import _cabi
    from tfhe_jl_b200.circuit import minimum_circuit
    rng, sk, ck = keypair
    ha, hb = T.encrypt(rng, sk, bits_of(1234, 16)), T.encrypt(rng, sk, bits_of(4321, 16))
    out = minimum_circuit(16).run(ck, {"a": ha, "b": hb})["min"].to_host()
    assert sum(int(v) << i for i, v in enumerate(T.decrypt(sk, out))) == 1234
    a, b = T.DeviceLweBatch.from_host(ha), T.DeviceLweBatch.from_host(hb)
    lt = T.constant_dev(ck, False)
    for i in range(16):
        lt = T.gate_dev(ck, _cabi.MUX, T.gate_dev(ck, _cabi.XNOR, a[i], b[i]), lt, a[i])
    want = T.gate_dev(ck, _cabi.MUX, lt.repeat(16), b, a).to_host()
    assert np.array_equal(out.data, want.data)


def test_levelised_adder32(keypair):
    from tfhe_jl_b200.circuit import adder_circuit
    rng, sk, ck = keypair
    x, y = 0xDEADBEEF, 0x12345678
    out = adder_circuit(32).run(ck, {"a": T.encrypt(rng, sk, bits_of(x, 32)), "b": T.encrypt(rng, sk, bits_of(y, 32))})
    assert sum(int(v) << i for i, v in enumerate(T.decrypt(sk, out["sum"].to_host()))) == (x + y) & 0xFFFFFFFF


def test_cuda_graph_capture_of_a_circuit_replays_bit_identical(keypair):
    """Circuit.compile captures every launch of the levelised evaluation in one CUDA graph; a
    replay on new inputs must produce the ciphertexts of the eager evaluation, and the library must stay usable on other
    streams afterwards."""
    from tfhe_jl_b200.circuit import minimum_circuit
    rng, sk, ck = keypair
    cm = minimum_circuit(16)
    compiled = cm.compile(ck)
    for va, vb in ((2017, 42), (7, 40000)):
        ia = {"a": T.encrypt(rng, sk, bits_of(va, 16)), "b": T.encrypt(rng, sk, bits_of(vb, 16))}
        got = compiled.run(ia)["min"].to_host()
        want = cm.run(ck, ia)["min"].to_host()
        assert np.array_equal(got.data, want.data)
        assert sum(int(v) << i for i, v in enumerate(T.decrypt(sk, got))) == min(va, vb)
    x, y = T.encrypt(rng, sk, [True, False]), T.encrypt(rng, sk, [True, True])
    assert np.array_equal(T.decrypt(sk, T.gate_nand(ck, x, y)), [False, True])


def test_cloud_key_file_round_trip(keypair, tmp_path):
    """A saved cloud key reloaded into a fresh context evaluates to the same ciphertexts."""
    rng, sk, ck = keypair
    T.save_cloud_key(tmp_path / "ck.npz", ck)
    ck2 = T.load_cloud_key(tmp_path / "ck.npz")
    x, y = T.encrypt(rng, sk, [True, False, True, False]), T.encrypt(rng, sk, [True, True, False, False])
    assert np.array_equal(T.gate_nand(ck2, x, y).data, T.gate_nand(ck, x, y).data)
