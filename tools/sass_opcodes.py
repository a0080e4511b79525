#!/usr/bin/env python
"""Static SASS opcode histogram of libtfhe_b200.so (cuobjdump -sass) -> profiles/r2/sass_opcodes.txt.

Shows what the kernels are made of: TMA bulk copies (UBLKCP), tensor-memory loads/stores (LDTM/STTM), mbarriers
(SYNCS), setmaxnreg (USETMAXREG), FP64 arithmetic, and that no tensor-core MMA is involved (the path is FP64 FFT work).
Runs on the build host, no GPU needed.
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "tfhe.jl_b200" / "libtfhe_b200.so"

PROOF = [
    ("UBLKCP", "cp.async.bulk (TMA bulk copy global->shared)"),
    ("LDTM", "tcgen05.ld (tensor memory -> registers)"),
    ("STTM", "tcgen05.st (registers -> tensor memory)"),
    ("SYNCS", "mbarrier operations"),
    ("USETMAXREG", "setmaxnreg (register rebalancing between warpgroups)"),
    ("UTCBAR", "tcgen05.commit"),
    ("UTCCP", "tcgen05.cp"),
    ("UTCHMMA", "tcgen05.mma"),
    ("UTCMMA", "tcgen05.mma"),
    ("HMMA", "mma.sync (none expected)"),
    ("DMMA", "FP64 mma (none expected)"),
    ("DFMA", "FP64 fused multiply-add"),
    ("DADD", "FP64 add"),
    ("DMUL", "FP64 multiply"),
    ("LDS", "shared-memory load"),
    ("STS", "shared-memory store"),
    ("BAR", "named / CTA barrier"),
    ("ATOMS", "shared-memory atomic"),
    ("SHFL", "warp shuffle"),
]

# the kernels the default dispatch launches (cabi.cu), by demangled-name prefix
DOMINANT = [
    "tfhe_b200::blind_rotate_kernel<2, 10, 2, 4, 5, 0, 3, 144>",   # K3, proven split, default
    "tfhe_b200::blind_rotate_kernel<2, 10, 1, 4, 6, 0, 0, 128>",   # K3, unsplit
    "tfhe_b200::blind_rotate_lowlat_kernel<2, 10, 2>",             # K3L
    "tfhe_b200::mk_blind_rotate_ring_kernel<4, 7, 2, 4, 6, 1>",    # MK, 2 parties
    "tfhe_b200::keyswitch_tile_kernel<512>",                       # K4
]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], check=True, capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), check=True,
                           capture_output=True, text=True).stdout.splitlines()
    per = []
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = collections.Counter()
            per.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    total = collections.Counter()
    for c in per:
        total.update(c)
    out = ["# SASS opcode histogram of tfhe.jl_b200/libtfhe_b200.so (cuobjdump -sass, sm_100a), round 2",
           "# static instruction counts; base mnemonic (modifiers stripped); regenerate with tools/sass_opcodes.py", "",
           f"kernels: {len(per)}   instructions: {sum(total.values())}", "",
           "## whole library: the instructions that prove what the kernels are made of",
           "| mnemonic | count | what it is |", "|---|---|---|"]
    out += [f"| {k} | {total.get(k, 0)} | {d} |" for k, d in PROOF]
    out += ["", "## the dominant kernels", ""]
    for want in DOMINANT:
        for name, c in zip(names, per):
            if want in name:
                out.append(f"### {name}")
                out.append(f"{sum(c.values())} instructions: " + ", ".join(f"{k} {v}" for k, v in c.most_common(24)))
                out.append("")
    dst = ROOT / "profiles" / "r2" / "sass_opcodes.txt"
    dst.write_text("\n".join(out))
    print(dst, len(per), "kernels")


if __name__ == "__main__":
    sys.exit(main())
