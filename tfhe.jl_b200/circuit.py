"""Levelised execution of gate circuits over device-resident ciphertexts.

The reference evaluates a circuit gate by gate (`examples/tutorial.jl:42-62`): every `gate_*` call is one
bootstrap.  A circuit is a DAG, so all gates whose operands are ready can be bootstrapped in ONE launch.  `Circuit`
records gates symbolically, assigns each its ASAP level (1 + the deepest operand), and `run` issues, per level,
one batched C-ABI call per opcode over operands gathered from a wire table that stays in HBM.  The 16-bit minimum
of the tutorial drops from 33 sequential bootstraps to 18 levels, a 32-bit ripple-carry adder with MUX carries
from 64 to 33.

The schedule (`levels()`) is plain host logic and is unit-tested without a GPU; `run` needs the CUDA library.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import _cabi

_ARITY = {_cabi.NOT: 1, _cabi.CONSTANT: 0, _cabi.MUX: 3}
_FREE = (_cabi.NOT, _cabi.CONSTANT)      # no bootstrap: they do not add a level (gates.jl:76-93)


@dataclass(frozen=True)
class Wire:
    """Handle of one encrypted bit inside a `Circuit`."""
    index: int


class Circuit:
    def __init__(self):
        self._inputs: List[Tuple[str, int, int]] = []       # (name, first wire, nbits)
        self._gates: List[Tuple[int, Tuple[int, ...], int]] = []   # (op, operand wires, output wire / constant bit)
        self._consts: Dict[int, bool] = {}
        self._outputs: List[Tuple[str, List[int]]] = []
        self._level: List[int] = []                          # per wire
        self._nwires = 0

    # ---- construction
    def _new_wire(self, level: int) -> int:
        self._level.append(level)
        self._nwires += 1
        return self._nwires - 1

    def input(self, name: str, nbits: int) -> List[Wire]:
        first = self._nwires
        for _ in range(nbits):
            self._new_wire(0)
        self._inputs.append((name, first, nbits))
        return [Wire(first + i) for i in range(nbits)]

    def constant(self, bit: bool) -> Wire:
        w = self._new_wire(0)
        self._consts[w] = bool(bit)
        self._gates.append((_cabi.CONSTANT, (), w))
        return Wire(w)

    def gate(self, op: int, *operands: Wire) -> Wire:
        arity = _ARITY.get(op, 2)
        if len(operands) != arity or op == _cabi.CONSTANT:
            raise ValueError(f"gate {op} takes {arity} operand(s); use constant() for constants")
        idx = tuple(o.index for o in operands)
        depth = max(self._level[i] for i in idx)
        w = self._new_wire(depth if op in _FREE else depth + 1)
        self._gates.append((op, idx, w))
        return Wire(w)

    # the reference's gate names (gates.jl)
    def nand(self, a, b): return self.gate(_cabi.NAND, a, b)
    def or_(self, a, b): return self.gate(_cabi.OR, a, b)
    def and_(self, a, b): return self.gate(_cabi.AND, a, b)
    def xor(self, a, b): return self.gate(_cabi.XOR, a, b)
    def xnor(self, a, b): return self.gate(_cabi.XNOR, a, b)
    def nor(self, a, b): return self.gate(_cabi.NOR, a, b)
    def not_(self, a): return self.gate(_cabi.NOT, a)
    def mux(self, a, b, c): return self.gate(_cabi.MUX, a, b, c)

    def output(self, name: str, wires: Sequence[Wire]):
        self._outputs.append((name, [w.index for w in wires]))

    # ---- schedule
    @property
    def depth(self) -> int:
        return max(self._level, default=0)

    @property
    def launches(self) -> int:
        """Bootstrapped (level, opcode) batches = sequential bootstrap latencies of one run."""
        return sum(1 for steps in self.levels() for op, _, _ in steps if op not in _FREE)

    @property
    def bootstraps(self) -> int:
        return sum(2 if op == _cabi.MUX else 1 for op, _, _ in self._gates if op not in _FREE)

    def _schedule(self) -> List[int]:
        """Final level of every wire.  A gate may sit anywhere between its ASAP level and the level its consumers
        allow; every (level, opcode) pair costs one launch, and a launch of a few gates costs a full bootstrap
        latency, so gates with slack are moved to the level inside their window that already holds the largest
        batch of the same opcode (ripple-carry adder: all propagate XORs at level 1, all sum XORs at the last
        level, instead of one XOR launch beside every carry MUX).  Gates that feed NOT/CONSTANT stay ASAP."""
        asap = list(self._level)
        depth = self.depth
        consumers: Dict[int, List[int]] = {}
        producer_gate: Dict[int, Tuple[int, Tuple[int, ...]]] = {}
        for op, idx, w in self._gates:
            producer_gate[w] = (op, idx)
            for i in idx:
                consumers.setdefault(i, []).append(w)
        # ALAP by a reverse pass (wires are numbered in creation order, so consumers have larger indices)
        alap = [depth] * self._nwires
        for w in range(self._nwires - 1, -1, -1):
            for c in consumers.get(w, []):
                cop = producer_gate[c][0]
                alap[w] = min(alap[w], alap[c] if cop in _FREE else alap[c] - 1)
        movable = [w for w, (op, _) in producer_gate.items()
                   if op not in _FREE and not any(producer_gate[c][0] in _FREE for c in consumers.get(w, []))]
        level = list(asap)
        placed = [True] * self._nwires
        for w in movable:
            placed[w] = alap[w] == asap[w]
        groups: Dict[Tuple[int, int], int] = {}
        for w, (op, _) in producer_gate.items():
            if placed[w] and op not in _FREE:
                groups[(level[w], op)] = groups.get((level[w], op), 0) + 1
        for w in sorted((w for w in movable if not placed[w]), key=lambda w: (alap[w] - asap[w], asap[w], w)):
            op, idx = producer_gate[w]
            lo = max([level[i] if placed[i] else asap[i] for i in idx], default=0) + 1
            hi = min([(level[c] if placed[c] else asap[c]) - 1 for c in consumers.get(w, [])], default=depth)
            best = max(range(lo, hi + 1), key=lambda L: (groups.get((L, op), 0), -L))
            level[w], placed[w] = best, True
            groups[(best, op)] = groups.get((best, op), 0) + 1
        return level

    def levels(self) -> List[List[Tuple[int, np.ndarray, np.ndarray]]]:
        """Per level, a list of (op, operand wire indices [arity][count], output wire indices [count]).
        Free gates (NOT, CONSTANT) are scheduled in the level of their operand, after that level's bootstraps
        whose outputs they may consume — hence two passes per level: bootstrapped gates first, then free gates
        in creation order (a NOT of a NOT stays ordered)."""
        by_level: Dict[int, Dict[int, List[Tuple[Tuple[int, ...], int]]]] = {}
        free_by_level: Dict[int, List[Tuple[int, Tuple[int, ...], int]]] = {}
        sched = self._schedule()
        for op, idx, w in self._gates:
            lvl = sched[w]
            if op in _FREE:
                free_by_level.setdefault(lvl, []).append((op, idx, w))
            else:
                by_level.setdefault(lvl, {}).setdefault(op, []).append((idx, w))
        out = []
        for lvl in range(0, self.depth + 1):
            steps = []
            for op, items in sorted(by_level.get(lvl, {}).items()):
                ops = np.array([i for i, _ in items], dtype=np.int64).T.reshape(_ARITY.get(op, 2), len(items))
                steps.append((op, ops, np.array([w for _, w in items], dtype=np.int64)))
            for op, idx, w in free_by_level.get(lvl, []):
                ops = np.array(idx, dtype=np.int64).reshape(len(idx), 1)
                steps.append((op, ops, np.array([w], dtype=np.int64)))
            out.append(steps)
        return out

    # ---- execution on the GPU
    def run(self, ck, inputs: Dict[str, "object"]) -> Dict[str, "object"]:
        """`inputs[name]` is a DeviceLweBatch (or an LweSample, uploaded once) of the declared width; returns a
        dict of DeviceLweBatch outputs.  Every intermediate ciphertext lives in one HBM wire table."""
        import torch
        from .api import DeviceLweBatch, LweSample
        width = ck.params.lwe_size + 1
        table = torch.empty((self._nwires, width), dtype=torch.int32, device="cuda")
        for name, first, nbits in self._inputs:
            x = inputs[name]
            if isinstance(x, LweSample):
                x = DeviceLweBatch.from_host(x)
            if x.tensor.shape != (nbits, width):
                raise ValueError(f"input {name!r}: expected {nbits} ciphertexts of {width} words")
            table[first:first + nbits] = x.tensor
        stream = torch.cuda.current_stream().cuda_stream
        for steps in self.levels():
            for op, ops, outw in steps:
                count = len(outw)
                dst = torch.empty((count, width), dtype=torch.int32, device="cuda")
                if op == _cabi.CONSTANT:
                    flags = torch.zeros((count, width), dtype=torch.int32)
                    flags[:, 0] = torch.tensor([int(self._consts[int(w)]) for w in outw], dtype=torch.int32)
                    src = [flags.cuda()]
                else:
                    src = [table.index_select(0, torch.from_numpy(ops[a]).cuda()) for a in range(ops.shape[0])]
                ptrs = [s.data_ptr() for s in src] + [0] * (3 - len(src))
                ck.ctx.gate_dev(op, ptrs[0], ptrs[1], ptrs[2], dst.data_ptr(), count, stream=stream)
                table.index_copy_(0, torch.from_numpy(outw).cuda(), dst)
        return {name: DeviceLweBatch(table.index_select(0, torch.tensor(w, device="cuda")))
                for name, w in self._outputs}


    def compile(self, ck) -> "CompiledCircuit":
        """Capture the whole levelised evaluation (every gather, gate launch and scatter of every level) in ONE CUDA
        graph over a static wire table: `compiled.run(inputs)` copies the inputs in and replays the graph — one
        submission per circuit evaluation instead of ~6 per level."""
        return CompiledCircuit(self, ck)


class CompiledCircuit:
    """A circuit bound to a cloud key and captured as a CUDA graph.  The graph replays the same
    kernels on the same buffers as `Circuit.run`, so the ciphertexts are bit-identical to the eager evaluation."""

    def __init__(self, circuit: Circuit, ck):
        import torch
        self.circuit, self.ck = circuit, ck
        self.width = ck.params.lwe_size + 1
        self.table = torch.zeros((circuit._nwires, self.width), dtype=torch.int32, device="cuda")
        self._steps = []
        for steps in circuit.levels():
            for op, ops, outw in steps:
                count = len(outw)
                if op == _cabi.CONSTANT:
                    flags = torch.zeros((count, self.width), dtype=torch.int32)
                    flags[:, 0] = torch.tensor([int(circuit._consts[int(w)]) for w in outw], dtype=torch.int32)
                    src = [flags.cuda()]; idx = None
                else:
                    src = None; idx = [torch.from_numpy(ops[a]).cuda() for a in range(ops.shape[0])]
                self._steps.append((op, idx, src, torch.from_numpy(outw).cuda(), count))
        self._out_idx = {name: torch.tensor(w, device="cuda") for name, w in circuit._outputs}
        self._outputs = None
        self._execute()                      # eager warm-up: every scratch buffer of the library reaches its final size
        torch.cuda.synchronize()
        ck.ctx.synchronize()                 # no library event from outside the capture is waited on inside it
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._execute()
        ck.ctx.synchronize()
        self.launches_per_run = len(self._steps)

    def _execute(self):
        import torch
        stream = torch.cuda.current_stream().cuda_stream
        for op, idx, src, outw, count in self._steps:
            dst = torch.empty((count, self.width), dtype=torch.int32, device="cuda")
            srcs = src if src is not None else [self.table.index_select(0, i) for i in idx]
            ptrs = [t.data_ptr() for t in srcs] + [0] * (3 - len(srcs))
            self.ck.ctx.gate_dev(op, ptrs[0], ptrs[1], ptrs[2], dst.data_ptr(), count, stream=stream)
            self.table.index_copy_(0, outw, dst)
        self._outputs = {name: self.table.index_select(0, i) for name, i in self._out_idx.items()}

    def run(self, inputs: Dict[str, "object"]) -> Dict[str, "object"]:
        from .api import DeviceLweBatch, LweSample
        for name, first, nbits in self.circuit._inputs:
            x = inputs[name]
            if isinstance(x, LweSample):
                x = DeviceLweBatch.from_host(x)
            if x.tensor.shape != (nbits, self.width):
                raise ValueError(f"input {name!r}: expected {nbits} ciphertexts of {self.width} words")
            self.table[first:first + nbits].copy_(x.tensor)
        self.graph.replay()
        return {name: DeviceLweBatch(t.clone()) for name, t in self._outputs.items()}


# ---- the two circuits BASELINE.json names ----------------------------------------------------------------
def minimum_circuit(nbits: int = 16) -> Circuit:
    """examples/tutorial.jl:42-62: min(a, b) by an LSB-to-MSB comparison, then a bitwise select."""
    c = Circuit()
    a, b = c.input("a", nbits), c.input("b", nbits)
    lt = c.constant(False)
    for i in range(nbits):                       # tutorial.jl:45-48 (encrypted_compare_bit)
        lt = c.mux(c.xnor(a[i], b[i]), lt, a[i])
    c.output("min", [c.mux(lt, b[i], a[i]) for i in range(nbits)])   # tutorial.jl:60
    return c


def adder_circuit(nbits: int = 32) -> Circuit:
    """Ripple-carry adder (BASELINE.json configs[3]): p = a xor b; carry' = p ? carry : a; sum = p xor carry."""
    c = Circuit()
    a, b = c.input("a", nbits), c.input("b", nbits)
    p = [c.xor(a[i], b[i]) for i in range(nbits)]
    carry = c.constant(False)
    carries = []
    for i in range(nbits):
        carries.append(carry)
        if i + 1 < nbits:
            carry = c.mux(p[i], carry, a[i])
    c.output("sum", [c.xor(p[i], carries[i]) for i in range(nbits)])
    return c
