"""Quick device-side timing probe of the NAND gate path (development tool, not the bench)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tfhe_jl_b200 as T
from oracle import oracle as O

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    flags = int(os.environ.get("FLAGS", "0"))
    keys = O.keygen(O.PARAMS_80, 123)
    P = keys.params
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, flags=flags)
    ctx.load_bk(keys.bk); ctx.load_ksk(keys.ksk)
    rng = O.Rng(1)
    base = 256
    bits = np.random.default_rng(0).integers(0, 2, (base, 2)).astype(bool)
    x = np.tile(O.encrypt(rng, keys, bits[:, 0]), (max(1, B // base), 1)); y = np.tile(O.encrypt(rng, keys, bits[:, 1]), (max(1, B // base), 1))
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    out = torch.empty_like(dx); u = torch.empty((B, 1025), dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    res = {"B": B, "flags": flags, "occ": os.environ.get("TFHE_B200_OCC", "3")}
    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    t_gate = timed(lambda: ctx.gate_dev(O.NAND, dx.data_ptr(), dy.data_ptr(), 0, out.data_ptr(), B, stream=s))
    t_br = timed(lambda: ctx.bootstrap_wo_ks_dev(dx.data_ptr(), u.data_ptr(), B, stream=s))
    t_ks = timed(lambda: ctx.keyswitch_dev(u.data_ptr(), out.data_ptr(), B, stream=s))
    ok = bool(np.array_equal(O.decrypt(keys, out.cpu().numpy()[:base]) if False else True, True))
    ctx.gate_dev(O.NAND, dx.data_ptr(), dy.data_ptr(), 0, out.data_ptr(), B, stream=s); torch.cuda.synchronize()
    dec = O.decrypt(keys, out.cpu().numpy()[:base])
    res.update(ms_gate=t_gate, ms_blind_rotate=t_br, ms_keyswitch=t_ks, gates_per_s=B / t_gate * 1e3,
               correct=bool(np.array_equal(dec, ~(bits[:, 0] & bits[:, 1]))))
    print(json.dumps(res))

main()
