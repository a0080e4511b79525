#!/usr/bin/env python
"""bench.py — bootstrapped NAND gates/s on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      (the CPU restatement of TFHE.jl)

A "step" is one pass of the hot path (gate prologue -> modswitch -> blind rotation -> extraction -> key
switch) over one batch of NAND gates on synthetic 80-bit-parameter ciphertexts (BASELINE.json configs[1]:
"batched NAND gate bootstrap sweep 1K-1M gates on 1 B200"; default 2^17 gates per GPU per step, so that the
8-GPU run holds 2^20 gates per step, inside configs[2]'s "1M-16M" range).

  value     gates/s with inputs already resident in HBM (tfhe_b200_gate_batch_dev), CUDA events on the
            launching stream, max over ranks.
  e2e       the same metric through the reference-facing host call tfhe_b200_gate_batch (what Julia's
            gate_nand.(…) binds to): pinned HOST buffers, host->device and device->host copies inside the
            timed region.  e2e_pageable: the same call on plain pageable numpy arrays (what a Julia Matrix{Int32} is).
            e2e_single_process (N > 1): ONE process, tfhe_b200_multi_gate_batch over all N GPUs on the global batch.
  value_unsplit  the device-resident value with the reference's own precision regime (one 32-bit piece, no proof).
  mk_nand   MK-TFHE NAND gates/s for 2/4/8 parties on one GPU (BASELINE.json configs[4]).
  roofline  dominant kernel = blind_rotate_kernel: algorithmic FP64 flops (SURVEY.md §8d, 94.72 MFLOP per
            gate) / its CUDA-event duration, against the FP64 FMA peak measured live on this GPU.
  cpu_baseline  the oracle (a C restatement of TFHE.jl's algorithm, kind "port") on the host cores.

PyTorch is used only for device memory, streams/events and torch.distributed (NCCL) — plumbing.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_FFT_FLOP_PER_GATE = 94.72e6      # n * 189 440 FP64 flop, FFT formulation (what the reference computes)
Q_KSK_BYTES_PER_GATE = 12_312_576  # KSK bytes gathered per gate
METRIC = "bootstrapped NAND gates/s"
OUT = sys.stdout


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 17, help="gates per GPU per step")
    ap.add_argument("--no-extras", action="store_true", help="skip value_unsplit, e2e_pageable, e2e_single_process and mk_nand")
    ap.add_argument("--unsplit", action="store_true", help="reference-precision FFT (one 32-bit piece) instead of the proven split")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def synth_inputs(O, keys, count, seed):
    """fresh encryptions of uniformly random bits; runtime is data-independent, so a
    2^12-gate base batch is tiled to `count`."""
    base = min(count, 1 << 12)
    bits = np.random.default_rng(seed).integers(0, 2, (base, 2)).astype(bool)
    rng = O.Rng(seed)
    x, y = O.encrypt(rng, keys, bits[:, 0]), O.encrypt(rng, keys, bits[:, 1])
    reps = -(-count // base)
    return np.tile(x, (reps, 1))[:count], np.tile(y, (reps, 1))[:count], np.tile(~(bits[:, 0] & bits[:, 1]), reps)[:count]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        for l in self.lines:
            f = [v.strip() for v in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args, rank, world):
    """--impl reference: TFHE.jl's own CPU algorithm for this path.  Julia is not installed and the
    reference is pure Julia (no C sources -> no oracle/_ref), so this times the oracle port — same folded
    complex-double FFT and operation order as polynomials.jl / tgsw.jl / bootstrap.jl / keyswitch.jl —
    with all host threads, on a bounded sample of the same workload per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # Genuine TFHE.jl when a Julia install and a TFHE.jl checkout are present (never the case in the build image
    # or on the GPU boxes of this pool; kept so the arm upgrades itself from "port" to "reference").
    import shutil
    proj = os.environ.get("TFHE_JL_PROJECT", "")
    if shutil.which("julia") and os.path.isdir(proj):
        sample = max(32, 4 * cores)
        r = subprocess.run(["julia", "--threads=auto", f"--project={proj}", os.path.join(ROOT, "tfhe.jl_b200", "julia", "cpu_baseline.jl"),
                            str(sample), str(args.steps), str(args.warmup)], capture_output=True, text=True)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            desc = f"{sample} NAND gates per step through TFHE.jl gate_nand, {j['cores']} Julia thread(s)"
            print(json.dumps({
                "impl": "reference", "metric": METRIC, "value": j["gates_per_s"], "unit": "gates/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": j["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "batched NAND gate bootstrap, 80-bit params (n=500, N=1024, k=1, l=2, Bg=2^10)", "sample": desc},
                "cpu_baseline": {"value": j["gates_per_s"], "unit": "gates/s", "cores": j["cores"], "kind": "reference", "sample": desc},
                "e2e": {"value": j["gates_per_s"], "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}),
                file=OUT, flush=True)
            return
        except (IndexError, KeyError, ValueError):
            pass   # fall through to the port
    from oracle import oracle as O
    keys = O.keygen(O.PARAMS_80, 123)
    ctx = O.Context(keys)
    sample = max(32, 4 * cores)                       # gates per step: ~0.2 s per step on any core count
    x, y, plain = synth_inputs(O, keys, sample, 1)
    for _ in range(args.warmup):
        ctx.gate(O.NAND, x, y, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = ctx.gate(O.NAND, x, y, nthreads=cores)
    dt = time.perf_counter() - t0
    assert np.array_equal(O.decrypt(keys, out), plain)
    value = sample * args.steps / dt
    sample_desc = f"{sample} NAND gates per step (bounded sample of the {args.batch}-gate batch), {cores} OpenMP threads, one gate per thread"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "gates/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "batched NAND gate bootstrap, 80-bit params (n=500, N=1024, k=1, l=2, Bg=2^10)", "sample": sample_desc},
        "cpu_baseline": {"value": value, "unit": "gates/s", "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=OUT, flush=True)


def mk_nand_rates(T, _cabi, torch, device):
    """BASELINE.json configs[4]: MK-TFHE NAND gates/s for 2/4/8 parties on one GPU, device-resident operands, proven
    (split) transform.  Throughput does not depend on the key VALUES, so key-shaped random material stands in for the
    oracle's keygen (a minute for 8 parties); ciphertext parity of these kernels is tests/test_gpu_mk.py's job."""
    from oracle import oracle as O
    out = {}
    for p, count in ((2, 2368), (4, 1184), (8, 592)):
        P = O.MK_PARAMS[p]
        rng = np.random.default_rng(p)
        ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, parties=p, device=device)
        ctx.load_bk(rng.integers(-2 ** 31, 2 ** 31, (p, P.n, P.l * (2 * p + 2), P.N), dtype=np.int32))
        ctx.load_ksk(rng.integers(-2 ** 31, 2 ** 31, (p,) + P.ksk_shape, dtype=np.int32))
        w = p * P.n + 1
        dx = torch.from_numpy(rng.integers(-2 ** 31, 2 ** 31, (count, w), dtype=np.int32)).cuda()
        dy = torch.from_numpy(rng.integers(-2 ** 31, 2 ** 31, (count, w), dtype=np.int32)).cuda()
        do = torch.empty_like(dx)
        s = torch.cuda.current_stream().cuda_stream
        fn = lambda: ctx.mk_nand_dev(dx.data_ptr(), dy.data_ptr(), do.data_ptr(), count, stream=s)
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        out[str(p)] = {"gates_per_s": count / (e0.elapsed_time(e1) * 1e-3), "gates": count, "unit": "gates/s"}
        ctx.close()
        del dx, dy, do
    return out


def main():
    # stdout carries exactly ONE JSON line: keep a private handle on the real stdout and point fd 1 at stderr, so
    # that anything a library writes to stdout (NCCL prints its version banner there) cannot pollute it
    global OUT
    sys.stdout.flush()
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    import tfhe_jl_b200 as T
    from tfhe_jl_b200 import _cabi
    from tfhe_jl_b200.sharding import shard_range
    from oracle import oracle as O      # synthetic keys/inputs + the cpu_baseline leg + output checking only

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- replicated keys (every rank derives the same key set; no key traffic between GPUs)
    keys = O.keygen(O.PARAMS_80, 123)
    P = keys.params
    flags = _cabi.FLAG_UNSPLIT_FFT if args.unsplit else _cabi.FLAG_SPLIT_FFT
    ctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, device=local_rank, flags=flags)
    ctx.load_bk(keys.bk)
    ctx.load_ksk(keys.ksk)

    # ---- this rank's shard of the global batch (weak scaling: args.batch gates per GPU)
    B = args.batch
    lo, hi = shard_range(B * world, rank, world)
    x, y, plain = synth_inputs(O, keys, hi - lo, 1000 + rank)
    W = P.n + 1
    hx, hy = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    hout = torch.empty((B, W), dtype=torch.int32).pin_memory()
    dx, dy = hx.cuda(non_blocking=True), hy.cuda(non_blocking=True)
    dout = torch.empty((B, W), dtype=torch.int32, device="cuda")
    du = torch.empty((B, P.N * P.k + 1), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()

    def step_dev():
        ctx.gate_dev(O.NAND, dx.data_ptr(), dy.data_ptr(), 0, dout.data_ptr(), B, stream=stream)

    def step_e2e():
        # the reference-facing call: host buffers in, host buffer out (copies inside the library call)
        rc = T.lib().tfhe_b200_gate_batch(ctx._h, O.NAND, hx.data_ptr(), hy.data_ptr(), None, hout.data_ptr(), B)
        if rc:
            raise RuntimeError(T.lib().tfhe_b200_last_error(ctx._h))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.kernel_launches
    ms_total = timed(step_dev, args.steps)
    launches = ctx.kernel_launches - launches0
    clocks = sampler.stop()
    value = B * world * args.steps / (ms_total * 1e-3)

    # correctness of what was just timed: every output of the last step decrypts to NAND
    got = dout.cpu().numpy()
    assert np.array_equal(O.decrypt(keys, got[:4096]), plain[:4096]), "timed outputs do not decrypt to NAND"

    # ---- end to end through the host-buffer C-ABI call
    step_e2e()
    barrier()
    t0 = time.perf_counter()      # after the barrier: rank skew from the checks above is not part of the timed region
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / float(e2e_s.item())
    assert np.array_equal(hout.numpy(), got), "host-buffer path and device-buffer path disagree"

    # ---- extras measured on every rank (max over ranks): pageable host buffers, the unproven one-piece transform
    extras = {}
    if not args.no_extras:
        px, py = np.array(x), np.array(y)                      # plain pageable arrays: what a Julia Matrix{Int32} is
        pout = np.empty((B, W), dtype=np.int32)

        def step_pageable():
            rc = T.lib().tfhe_b200_gate_batch(ctx._h, O.NAND, px.ctypes.data, py.ctypes.data, None, pout.ctypes.data, B)
            if rc:
                raise RuntimeError(T.lib().tfhe_b200_last_error(ctx._h))

        step_pageable()
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, min(args.steps, 3))):
            step_pageable()
        barrier()
        pg_s = torch.tensor([time.perf_counter() - t0], device="cuda")
        if world > 1:
            dist.all_reduce(pg_s, op=dist.ReduceOp.MAX)
        assert np.array_equal(pout, got), "pageable host path and device path disagree"
        extras["e2e_pageable"] = {"value": B * world * max(1, min(args.steps, 3)) / float(pg_s.item()), "unit": "gates/s",
                                  "note": "tfhe_b200_gate_batch on plain pageable numpy arrays (no pinning by the caller)"}
        if not args.unsplit:
            uctx = T.Context(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, device=local_rank, flags=_cabi.FLAG_UNSPLIT_FFT)
            uctx.load_bk(keys.bk); uctx.load_ksk(keys.ksk)
            ustep = lambda: uctx.gate_dev(O.NAND, dx.data_ptr(), dy.data_ptr(), 0, dout.data_ptr(), B, stream=stream)
            ustep()
            ums = timed(ustep, max(1, min(args.steps, 3)))
            assert np.array_equal(dout.cpu().numpy(), got), "one-piece transform and proven transform disagree"
            extras["value_unsplit"] = {"value": B * world * max(1, min(args.steps, 3)) / (ums * 1e-3), "unit": "gates/s",
                                       "note": "reference precision regime (polynomials.jl:138-140): one 32-bit piece, exact in every test, no proof"}
            uctx.close()
        if world > 1:
            # ONE process driving all GPUs through the product's multi-device context; the other ranks stay idle.
            # They must wait on the HOST: an NCCL barrier is a kernel spinning on their GPU, and kernels of two processes
            # time-slice a GPU, which would halve the speed of rank 0's shard on that GPU.
            host = dist.new_group(backend="gloo")
            torch.cuda.synchronize()
            dist.barrier(group=host)
            if rank == 0:
                m = _cabi.MultiContext(n=P.n, l=P.l, bgbit=P.bgbit, t=P.t, basebit=P.basebit, devices=list(range(world)), flags=flags)
                m.load_bk(keys.bk); m.load_ksk(keys.ksk)
                gx = torch.from_numpy(np.tile(x, (world, 1))).pin_memory(); gy = torch.from_numpy(np.tile(y, (world, 1))).pin_memory()
                gout = torch.empty((B * world, W), dtype=torch.int32).pin_memory()
                run = lambda: m._ck(T.lib().tfhe_b200_multi_gate_batch(m._h, O.NAND, gx.data_ptr(), gy.data_ptr(), None, gout.data_ptr(), B * world))
                run()
                reps = max(1, min(args.steps, 3))
                t0 = time.perf_counter()
                for _ in range(reps):
                    run()
                sp = time.perf_counter() - t0
                assert np.array_equal(gout.numpy()[:B], got) and np.array_equal(gout.numpy()[-B:], got), "multi-device context disagrees"
                extras["e2e_single_process"] = {"value": B * world * reps / sp, "unit": "gates/s", "devices": m.devices,
                                                "note": "one process, tfhe_b200_multi_gate_batch: host buffers in, host buffer out, batch sharded over all GPUs"}
                m.close()
            dist.barrier(group=host)

    if rank == 0:
        # ---- dominant kernel alone, CUDA events on its launching stream
        def br_only():
            ctx.bootstrap_wo_ks_dev(dx.data_ptr(), du.data_ptr(), B, stream=stream)

        def ks_only():
            ctx.keyswitch_dev(du.data_ptr(), dout.data_ptr(), B, stream=stream)

        def kernel_ms(fn, reps):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        reps = max(1, min(args.steps, 3))
        br_ms, ks_ms = kernel_ms(br_only, reps), kernel_ms(ks_only, reps)

        # the second half of BASELINE.json's metric: latency of ONE bootstrapped gate (device-resident operands)
        def one_gate():
            ctx.gate_dev(O.NAND, dx.data_ptr(), dy.data_ptr(), 0, dout.data_ptr(), 1, stream=stream)

        lat = sorted(kernel_ms(one_gate, 1) for _ in range(9))[4]
        fp64_peak = ctx.measure_fp64_tflops()
        lds_peak = ctx.measure_lds_gbps()
        achieved = W_FFT_FLOP_PER_GATE * B / (br_ms * 1e-3) / 1e12
        traffic, traffic_src = None, "no ncu capture committed for this launch size (profiles/traffic.json)"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tj["unsplit" if args.unsplit else "split"].get(str(B))
            if ent:
                traffic, traffic_src = ent["dram_bytes"], ent["source"]
        except (OSError, KeyError, ValueError):
            pass
        # K4 (tiled): every ciphertext reads N*t table rows of `stride` words from shared memory
        ks_tiled = B >= 4096
        ks_smem_bytes = P.N * P.k * P.t * (((P.n + 1 + 31) & ~31) * 4)
        ks_gbs = (ks_smem_bytes if ks_tiled else Q_KSK_BYTES_PER_GATE) * B / (ks_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batched NAND gate bootstrap (blind rotation + keyswitch), 80-bit params (n=500, N=1024, k=1, l=2, Bg=2^10)",
                       "gates_per_gpu_per_step": B, "global_gates_per_step": B * world, "parallelism": f"batch sharded over {world} GPU(s), keys replicated, no collective",
                       "transform": "complex-double negacyclic FFT, " + ("one 32-bit piece (reference precision)" if args.unsplit else "torus operand split in two 16-bit pieces (proven exact)"),
                       "l2": f"inputs ({2 * B * W * 4 / 1e6:.0f} MB read + {B * W * 4 / 1e6:.0f} MB written per step) exceed the 126 MB L2; BK/KSK are reused by every gate by design"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "gates/s", "h2d_bytes_per_step": 2 * B * W * 4 * world, "d2h_bytes_per_step": B * W * 4 * world},
            "gpu_launches": int(launches),
            "single_bootstrap_latency_ms": lat,
            "roofline": {"bound": "fp64", "kernel": "blind_rotate_kernel", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak, "traffic": traffic,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one blind_rotate_kernel launch of this size (bytes); " + traffic_src + "; keys are L2-resident, the kernel is not DRAM-bound",
                         "peak_source": "FP64 FMA rate measured live on this GPU (MEASURED_PEAKS.json has no FP64 entry); not tensor- or HBM-bound: keys are L2-resident",
                         "kernel_ms": br_ms, "algorithmic_flop_per_gate": W_FFT_FLOP_PER_GATE, "share_of_step": br_ms / (ms_total / args.steps)},
            "roofline_keyswitch": {"bound": "smem" if ks_tiled else "l2", "kernel": "keyswitch_tile_kernel" if ks_tiled else "keyswitch_kernel",
                                   "achieved": ks_gbs, "peak": lds_peak, "unit": "GB/s", "frac": ks_gbs / lds_peak,
                                   "peak_source": "conflict-free LDS.128 read rate measured live on this GPU; the tile kernel streams the 50 MB table once per 64 ciphertexts "
                                                  "through shared memory and every ciphertext reads N*t rows of it with LDS.128, so shared memory is the pipe that bounds it (not HBM)",
                                   "kernel_ms": ks_ms, "algorithmic_bytes_per_gate": ks_smem_bytes if ks_tiled else Q_KSK_BYTES_PER_GATE},
        }
        line.update(extras)
        if world == 1 and not args.no_extras:
            line["mk_nand"] = mk_nand_rates(T, _cabi, torch, local_rank)
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            octx = O.Context(keys)
            t0 = time.perf_counter()
            octx.gate(O.NAND, x[:cores], y[:cores], nthreads=cores)          # warm-up + rate estimate
            rate = cores / (time.perf_counter() - t0)
            sample = int(min(B, max(4 * cores, rate * 12.0)) // cores * cores) or cores   # ~12 s of CPU work
            t0 = time.perf_counter()
            ref = octx.gate(O.NAND, x[:sample], y[:sample], nthreads=cores)
            dt = time.perf_counter() - t0
            assert np.array_equal(ref, got[:sample]), "GPU ciphertexts differ from the oracle"
            line["cpu_baseline"] = {"value": sample / dt, "unit": "gates/s", "cores": cores, "kind": "port", "seconds": dt,
                                    "sample": f"first {sample} gates of the step's batch, one gate per OpenMP thread; GPU outputs bit-identical on this sample"}
        print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
