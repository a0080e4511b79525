"""Single-gate latency of the 128-bit parameter set (n = 630, l = 3, Bg = 2^7), device-resident operands."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tfhe_jl_b200 as T
from oracle import oracle as O

keys = O.keygen(O.PARAMS_128, 123); P = keys.params
res = {}
for flags in (0, 1):
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, flags=flags)
    ctx.load_bk(keys.bk); ctx.load_ksk(keys.ksk)
    rng = O.Rng(1)
    for B in (1, 16, 148):
        bits = np.random.default_rng(B).integers(0, 2, (B, 2)).astype(bool)
        x, y = O.encrypt(rng, keys, bits[:, 0]), O.encrypt(rng, keys, bits[:, 1])
        dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(); out = torch.empty_like(dx)
        s = torch.cuda.current_stream().cuda_stream
        fn = lambda: ctx.gate_dev(O.NAND, dx.data_ptr(), dy.data_ptr(), 0, out.data_ptr(), B, stream=s)
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        assert np.array_equal(O.decrypt(keys, out.cpu().numpy()), ~(bits[:, 0] & bits[:, 1]))
        res[f"flags{flags}_B{B}_ms"] = sorted(ts)[2]
print(json.dumps(res))
