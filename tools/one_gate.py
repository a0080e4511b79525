import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import tfhe_jl_b200 as T
from oracle import oracle as O
keys = O.keygen(O.PARAMS_80, 123); P = keys.params
ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit); ctx.load_bk(keys.bk); ctx.load_ksk(keys.ksk)
rng = O.Rng(1); bits = np.random.default_rng(0).integers(0, 2, (16, 2)).astype(bool)
dx = torch.from_numpy(O.encrypt(rng, keys, bits[:, 0])).cuda(); dy = torch.from_numpy(O.encrypt(rng, keys, bits[:, 1])).cuda()
out = torch.empty_like(dx); s = torch.cuda.current_stream().cuda_stream
for B in (1, 1, 16, 16):
    ctx.gate_dev(O.NAND, dx.data_ptr(), dy.data_ptr(), 0, out.data_ptr(), B, stream=s); torch.cuda.synchronize()
