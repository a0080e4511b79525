"""Balanced last wave of K3 (TFHE_B200_BALANCE, default 1) against full CTAs only (0): ms of one NAND batch of B gates,
device-resident, median of 7, both results compared bit for bit.  One JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tfhe_jl_b200 as T
from oracle import oracle as O
keys = O.keygen(O.PARAMS_80, 123); P = keys.params
sizes = (593, 600, 700, 900, 1024, 1500, 2048, 3000, 4096)
bits = np.random.default_rng(0).integers(0, 2, (max(sizes), 2)).astype(bool)
rng = O.Rng(1)
dx = torch.from_numpy(O.encrypt(rng, keys, bits[:, 0])).cuda(); dy = torch.from_numpy(O.encrypt(rng, keys, bits[:, 1])).cuda()
s = torch.cuda.current_stream().cuda_stream
res, outs = {}, {}
for bal in ("1", "0"):
    os.environ["TFHE_B200_BALANCE"] = bal
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit); ctx.load_bk(keys.bk); ctx.load_ksk(keys.ksk)
    for B in sizes:
        out = torch.empty((B, P.n + 1), dtype=torch.int32, device="cuda")
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.gate_dev(O.NAND, dx.data_ptr(), dy.data_ptr(), 0, out.data_ptr(), B, stream=s); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[f"balance{bal}_B{B}_ms"] = round(sorted(ts)[3], 3)
        if bal == "1": outs[B] = out.cpu()
        else: res[f"identical_B{B}"] = bool(torch.equal(outs[B], out.cpu()))
print(json.dumps(res))
