"""Small batches: the latency kernel (one gate per CTA, several waves) against K3 with a balanced wave.  ms per NAND batch."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tfhe_jl_b200 as T
from oracle import oracle as O
keys = O.keygen(O.PARAMS_80, 123); P = keys.params
sizes = (148, 149, 200, 296, 297, 370, 444, 445, 592)
bits = np.random.default_rng(0).integers(0, 2, (max(sizes), 2)).astype(bool)
rng = O.Rng(1)
dx = torch.from_numpy(O.encrypt(rng, keys, bits[:, 0])).cuda(); dy = torch.from_numpy(O.encrypt(rng, keys, bits[:, 1])).cuda()
s = torch.cuda.current_stream().cuda_stream
res, ref = {}, {}
for flags in (0, 1):
  for waves in ("3", "2", "1"):
    os.environ["TFHE_B200_LOWLAT_WAVES"] = waves
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, flags=flags); ctx.load_bk(keys.bk); ctx.load_ksk(keys.ksk)
    for B in sizes:
        out = torch.empty((B, P.n + 1), dtype=torch.int32, device="cuda")
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.gate_dev(O.NAND, dx.data_ptr(), dy.data_ptr(), 0, out.data_ptr(), B, stream=s); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[f"flags{flags}_waves{waves}_B{B}_ms"] = round(sorted(ts)[3], 3)
        if waves == "3" and flags == 0: ref[B] = out.cpu()
        elif flags == 0: assert torch.equal(ref[B], out.cpu()), (waves, B)
print(json.dumps(res))
