"""Single-bootstrap latency and small-batch behaviour (development probe)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tfhe_jl_b200 as T
from oracle import oracle as O
keys = O.keygen(O.PARAMS_80, 123); P = keys.params
res = {}
for flags in (0, 1):
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, flags=flags); ctx.load_bk(keys.bk); ctx.load_ksk(keys.ksk)
    rng = O.Rng(1); bits = np.random.default_rng(0).integers(0, 2, (1024, 2)).astype(bool)
    dx = torch.from_numpy(O.encrypt(rng, keys, bits[:, 0])).cuda(); dy = torch.from_numpy(O.encrypt(rng, keys, bits[:, 1])).cuda()
    out = torch.empty_like(dx); s = torch.cuda.current_stream().cuda_stream
    for B in (1, 4, 16, 148, 592, 1024):
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.gate_dev(O.NAND, dx.data_ptr(), dy.data_ptr(), 0, out.data_ptr(), B, stream=s); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[f"flags{flags}_B{B}_ms"] = sorted(ts)[3]
print(json.dumps(res))
