"""MK-TFHE NAND throughput probe (BASELINE.json config #5): parties in {2,4,8}, device-resident inputs."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tfhe_jl_b200 as T
from oracle import oracle as O

def main():
    p = int(sys.argv[1]); B = int(sys.argv[2]); n = int(sys.argv[3]) if len(sys.argv) > 3 else 500
    flags = int(os.environ.get("FLAGS", "0"))
    P = O.small_params(O.MK_PARAMS[p], n)
    t0 = time.time(); mk = O.mk_keygen(P, p, 5); t_keygen = time.time() - t0
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, parties=p, flags=flags)
    ctx.load_bk(mk.bk); ctx.load_ksk(mk.ksk)
    base = min(B, 64)
    bits = np.random.default_rng(0).integers(0, 2, (base, 2)).astype(bool)
    rng = O.Rng(1)
    x = np.tile(O.mk_encrypt(rng, mk, bits[:, 0]), (B // base, 1)); y = np.tile(O.mk_encrypt(rng, mk, bits[:, 1]), (B // base, 1))
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(); out = torch.empty_like(dx)
    s = torch.cuda.current_stream().cuda_stream
    fn = lambda: ctx.mk_nand_dev(dx.data_ptr(), dy.data_ptr(), out.data_ptr(), B, stream=s)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    dec = O.mk_decrypt(mk, out.cpu().numpy()[:base])
    print(json.dumps({"parties": p, "n": n, "B": B, "flags": flags, "ms": ms, "gates_per_s": B / ms * 1e3,
                      "correct": bool(np.array_equal(dec, ~(bits[:, 0] & bits[:, 1]))), "oracle_keygen_s": t_keygen}))
main()
