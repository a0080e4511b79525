#!/usr/bin/env python
"""BASELINE.json configs[1]: batched NAND gate bootstrap sweep, 1K .. 1M gates on one B200.

For every batch size: device-resident throughput (tfhe_b200_gate_batch_dev, CUDA events) and host-buffer
end-to-end throughput (tfhe_b200_gate_batch, wall clock, copies inside), both FFT modes.  One JSON line.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import tfhe_jl_b200 as T  # noqa: E402
from oracle import oracle as O  # noqa: E402  (synthetic keys / inputs and the decryption check only)


def main():
    sizes = [int(s) for s in (sys.argv[1:] or ["1024", "4096", "16384", "65536", "262144", "1048576"])]
    pset = os.environ.get("PARAMS", "80")
    keys = O.keygen(O.PARAMS_128 if pset == "128" else O.PARAMS_80, 123)
    P = keys.params
    base = 4096
    bits = np.random.default_rng(0).integers(0, 2, (base, 2)).astype(bool)
    rng = O.Rng(1)
    bx, by = O.encrypt(rng, keys, bits[:, 0]), O.encrypt(rng, keys, bits[:, 1])
    want = ~(bits[:, 0] & bits[:, 1])
    res = {"workload": f"NAND, {pset}-bit parameters (n={P.n}, l={P.l}, Bg=2^{P.bgbit}), fresh ciphertexts tiled from a 4096-gate base batch", "rows": []}
    for flags, mode in ((0, "split"), (1, "unsplit")):
        ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, flags=flags)
        ctx.load_bk(keys.bk); ctx.load_ksk(keys.ksk)
        for B in sizes:
            reps = -(-B // base)
            hx = torch.from_numpy(np.tile(bx, (reps, 1))[:B]).pin_memory()
            hy = torch.from_numpy(np.tile(by, (reps, 1))[:B]).pin_memory()
            hout = torch.empty_like(hx).pin_memory()
            dx, dy = hx.cuda(), hy.cuda()
            dout = torch.empty_like(dx)
            s = torch.cuda.current_stream().cuda_stream
            run = lambda: ctx.gate_dev(O.NAND, dx.data_ptr(), dy.data_ptr(), 0, dout.data_ptr(), B, stream=s)
            run(); torch.cuda.synchronize()
            n = 3 if B <= 65536 else 1
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            if B <= 65536:   # warm the library's staging buffers (they grow on first use of a larger batch)
                T.lib().tfhe_b200_gate_batch(ctx._h, O.NAND, hx.data_ptr(), hy.data_ptr(), None, hout.data_ptr(), B)
            t0 = time.perf_counter()
            rc = T.lib().tfhe_b200_gate_batch(ctx._h, O.NAND, hx.data_ptr(), hy.data_ptr(), None, hout.data_ptr(), B)
            e2e_s = time.perf_counter() - t0
            assert rc == 0
            got = dout.cpu().numpy()
            assert np.array_equal(hout.numpy(), got)
            m = min(B, base)
            assert np.array_equal(O.decrypt(keys, got[:m]), want[:m])
            res["rows"].append({"mode": mode, "gates": B, "ms": ms, "gates_per_s": B / ms * 1e3, "e2e_gates_per_s": B / e2e_s})
            del hx, hy, hout, dx, dy, dout
    print(json.dumps(res))


if __name__ == "__main__":
    main()
