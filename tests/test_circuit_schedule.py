"""Host logic of the levelised circuit driver (tfhe.jl_b200/circuit.py): no GPU needed — the schedule returned by
`levels()` is executed on plaintext bits and compared with integer arithmetic."""
import numpy as np
import pytest

from tfhe_jl_b200 import _cabi
from tfhe_jl_b200.circuit import Circuit, adder_circuit, minimum_circuit

PLAIN = {
    _cabi.NAND: lambda a, b: ~(a & b), _cabi.OR: lambda a, b: a | b, _cabi.AND: lambda a, b: a & b,
    _cabi.XOR: lambda a, b: a ^ b, _cabi.XNOR: lambda a, b: ~(a ^ b), _cabi.NOR: lambda a, b: ~(a | b),
    _cabi.NOT: lambda a: ~a, _cabi.MUX: lambda a, b, c: np.where(a, b, c),
}


def run_plain(c: Circuit, inputs):
    table = np.zeros(c._nwires, dtype=bool)
    done = np.zeros(c._nwires, dtype=bool)
    for name, first, nbits in c._inputs:
        table[first:first + nbits] = inputs[name]; done[first:first + nbits] = True
    for steps in c.levels():
        for op, ops, outw in steps:
            if op == _cabi.CONSTANT:
                table[outw] = [c._consts[int(w)] for w in outw]
            else:
                assert done[ops].all(), "schedule used a wire before it was produced"
                table[outw] = PLAIN[op](*[table[ops[a]] for a in range(ops.shape[0])])
            done[outw] = True
    return {name: table[w] for name, w in c._outputs}


def bits(v, n):
    return np.array([(v >> i) & 1 for i in range(n)], dtype=bool)


def value(b):
    return sum(int(x) << i for i, x in enumerate(b))


def test_minimum_circuit_schedule():
    c = minimum_circuit(16)
    assert c.depth == 18                       # 1 (all XNORs at once) + 16 (MUX chain) + 1 (select)
    assert c.bootstraps == 80                  # 80 blind rotations in examples/tutorial.jl
    assert c.launches == 18
    lv = c.levels()
    assert [op for op, _, _ in lv[1]] == [_cabi.XNOR] and lv[1][0][1].shape == (2, 16)
    rng = np.random.default_rng(0)
    for a, b in [(2017, 42), (42, 2017), (7, 7), (0, 65535)] + [tuple(rng.integers(0, 65536, 2)) for _ in range(20)]:
        assert value(run_plain(c, {"a": bits(int(a), 16), "b": bits(int(b), 16)})["min"]) == min(int(a), int(b))


def test_adder_circuit_schedule():
    c = adder_circuit(32)
    assert c.depth == 33
    assert c.launches == 33                    # 1 (all propagate XORs) + 31 (carry MUX chain) + 1 (all sum XORs)
    rng = np.random.default_rng(1)
    for a, b in [(0xDEADBEEF, 0x12345678), (0xFFFFFFFF, 1), (0, 0)] + [tuple(rng.integers(0, 2 ** 32, 2)) for _ in range(20)]:
        got = value(run_plain(c, {"a": bits(int(a), 32), "b": bits(int(b), 32)})["sum"])
        assert got == (int(a) + int(b)) & 0xFFFFFFFF


def test_free_gates_do_not_add_levels_and_keep_order():
    c = Circuit()
    (a,), (b,) = c.input("a", 1), c.input("b", 1)
    x = c.not_(c.not_(c.nand(a, b)))           # two free gates after one bootstrap
    y = c.and_(x, c.not_(a))
    c.output("y", [y])
    assert c.depth == 2 and c.bootstraps == 2
    for va in (False, True):
        for vb in (False, True):
            out = run_plain(c, {"a": np.array([va]), "b": np.array([vb])})["y"][0]
            assert out == ((not (va and vb)) and (not va))


def test_arity_is_checked():
    c = Circuit()
    (a,) = c.input("a", 1)
    with pytest.raises(ValueError):
        c.gate(_cabi.NAND, a)
    with pytest.raises(ValueError):
        c.gate(_cabi.CONSTANT)


def test_random_dags_schedule_is_valid_and_equivalent():
    """Random circuits with every gate type: the slack-aware schedule never reads a wire before it is produced
    (asserted inside run_plain) and computes what gate-by-gate evaluation in creation order computes."""
    rng = np.random.default_rng(7)
    binary = [_cabi.NAND, _cabi.OR, _cabi.AND, _cabi.XOR, _cabi.XNOR, _cabi.NOR]
    for trial in range(30):
        c = Circuit()
        wires = c.input("x", 6) + [c.constant(bool(trial & 1))]
        ref = {}                                   # wire index -> function of the input bits (evaluated eagerly below)
        xin = rng.integers(0, 2, 6).astype(bool)
        val = {w.index: bool(v) for w, v in zip(wires[:6], xin)}
        val[wires[6].index] = bool(trial & 1)
        for _ in range(40):
            kind = rng.integers(0, 10)
            if kind == 0:
                a = wires[rng.integers(len(wires))]
                w = c.not_(a); val[w.index] = not val[a.index]
            elif kind == 1:
                a, b, d = (wires[i] for i in rng.integers(0, len(wires), 3))
                w = c.mux(a, b, d); val[w.index] = val[b.index] if val[a.index] else val[d.index]
            else:
                op = binary[rng.integers(len(binary))]
                a, b = (wires[i] for i in rng.integers(0, len(wires), 2))
                w = c.gate(op, a, b)
                val[w.index] = bool(PLAIN[op](np.array(val[a.index]), np.array(val[b.index])))
            wires.append(w)
        outs = [wires[i] for i in rng.integers(6, len(wires), 8)]
        c.output("o", outs)
        got = run_plain(c, {"x": xin})["o"]
        assert list(got) == [val[w.index] for w in outs], trial
        assert c.launches <= sum(1 for op, _, _ in c._gates if op not in (_cabi.NOT, _cabi.CONSTANT))
