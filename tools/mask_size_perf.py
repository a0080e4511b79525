"""NAND gates/s for tlwe_mask_size k = 2, 3 (blind_rotate_wide.cuh), device-resident inputs, CUDA events; every output is
checked by decryption and a sample of ciphertexts against the oracle.  One JSON line per (k, batch)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfhe_jl_b200 as T  # noqa: E402
from oracle import oracle as O  # noqa: E402

base = O.PARAMS_80
for k in (2, 3):
    P = O.Params(base.n, base.lwe_sigma, base.N, k, base.l, base.bgbit, base.bs_sigma, base.t, base.basebit, base.ks_sigma, 1)
    keys = O.keygen(P, 1)
    octx = O.Context(keys)
    ctx = T.Context(n=P.n, k=k, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit)
    ctx.load_bk(keys.bk); ctx.load_ksk(keys.ksk)
    for B in (1, 296, 4736):
        bits = np.random.default_rng(B).integers(0, 2, (B, 2)).astype(bool)
        hx, hy = O.encrypt(O.Rng(1), keys, bits[:, 0]), O.encrypt(O.Rng(2), keys, bits[:, 1])
        x, y = torch.from_numpy(hx).cuda(), torch.from_numpy(hy).cuda()
        out = torch.empty_like(x)
        s = torch.cuda.current_stream().cuda_stream
        ctx.gate_dev(O.NAND, x.data_ptr(), y.data_ptr(), 0, out.data_ptr(), B, stream=s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.gate_dev(O.NAND, x.data_ptr(), y.data_ptr(), 0, out.data_ptr(), B, stream=s)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        got = out.cpu().numpy()
        pick = np.random.default_rng(0).choice(B, min(B, 4), replace=False)
        print(json.dumps({"k": k, "gates": B, "ms": round(ms, 3), "gates_per_s": round(B / ms * 1e3, 1),
                          "decrypts": bool((O.decrypt(keys, got) == ~(bits[:, 0] & bits[:, 1])).all()),
                          "oracle_identical_sample": bool(np.array_equal(got[pick], octx.gate(O.NAND, hx[pick], hy[pick])))}), flush=True)
