#!/usr/bin/env python
"""bench.py — bootstrapped NAND gates/s on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      (the CPU restatement of TFHE.jl)

A "step" is one pass of the hot path (gate prologue -> modswitch -> blind rotation -> extraction -> key
switch) over one batch of NAND gates on synthetic 80-bit-parameter ciphertexts (BASELINE.json configs[1]:
"batched NAND gate bootstrap sweep 1K-1M gates on 1 B200"; default 2^16 gates per GPU per step).

  value     gates/s with inputs already resident in HBM (tfhe_b200_gate_batch_dev), CUDA events on the
            launching stream, max over ranks.
  e2e       the same metric through the reference-facing host call tfhe_b200_gate_batch (what Julia's
            gate_nand.(…) binds to): pinned HOST buffers, host->device and device->host copies inside the
            timed region.
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
    ap.add_argument("--batch", type=int, default=1 << 16, help="gates per GPU per step")
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
        achieved = W_FFT_FLOP_PER_GATE * B / (br_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        traffic, traffic_src = None, "no ncu capture committed for this launch size (profiles/traffic.json)"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tj["unsplit" if args.unsplit else "split"].get(str(B))
            if ent:
                traffic, traffic_src = ent["dram_bytes"], ent["source"]
        except (OSError, KeyError, ValueError):
            pass
        hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        ks_gbs = Q_KSK_BYTES_PER_GATE * B / (ks_ms * 1e-3) / 1e9
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
            "roofline_keyswitch": {"bound": "hbm", "kernel": "keyswitch_tile_kernel" if B >= 4096 else "keyswitch_kernel", "achieved": ks_gbs, "peak": hbm_peak, "unit": "GB/s",
                                   "frac": ks_gbs / hbm_peak, "peak_source": hbm_src + "; algorithmic bytes = table rows gathered per gate; the tile kernel streams the 50 MB table once per 64 ciphertexts through shared memory, so frac > 1 is expected (bound: shared-memory pipe)",
                                   "kernel_ms": ks_ms, "algorithmic_bytes_per_gate": Q_KSK_BYTES_PER_GATE},
        }
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
