"""Key switch of B random dimension-1024 samples: tiles of 64 / tiles of 32 / one CTA per ciphertext.  ms (median of 7), bit-compared."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tfhe_jl_b200 as T
from oracle import oracle as O
keys = O.keygen(O.PARAMS_80, 123); P = keys.params
sizes = (1024, 2048, 3072, 4096, 6144, 8192, 9472, 16384)
u = torch.from_numpy(np.random.default_rng(0).integers(-2**31, 2**31, (max(sizes), 1025), dtype=np.int64).astype(np.int32)).cuda()
s = torch.cuda.current_stream().cuda_stream
res, ref = {}, {}
for name, env in (("tile64", {"TFHE_B200_KS_TILE32": "0", "TFHE_B200_KS_TILE_MIN": "1"}), ("tile32", {"TFHE_B200_KS_TILE32": "1", "TFHE_B200_KS_TILE_MIN": "1"}), ("per_ct", {"TFHE_B200_KS_TILE": "0"})):
    for k in ("TFHE_B200_KS_TILE32", "TFHE_B200_KS_TILE_MIN", "TFHE_B200_KS_TILE"): os.environ.pop(k, None)
    os.environ.update(env)
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit); ctx.load_bk(keys.bk); ctx.load_ksk(keys.ksk)
    for B in sizes:
        out = torch.empty((B, P.n + 1), dtype=torch.int32, device="cuda")
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.keyswitch_dev(u.data_ptr(), out.data_ptr(), B, stream=s); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[f"{name}_B{B}_ms"] = round(sorted(ts)[3], 3)
        if name == "tile64": ref[B] = out.cpu()
        else: res[f"{name}_B{B}_identical"] = bool(torch.equal(ref[B], out.cpu()))
print(json.dumps(res))
