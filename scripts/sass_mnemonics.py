#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libtbcuda.so (cuobjdump -sass): the evidence that the hot kernels use the
instructions DESIGN.md says they use (VIADDMNMX(.S16x2) DPX, UBLKCP bulk TMA copies, SYNCS mbarriers, USETMAXREG,
uniform-datapath address arithmetic in the level-synchronous GEMM instance).  usage: sass_mnemonics.py [lib] > profiles/..."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tensorbranching.jl_b200", "libtbcuda.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["VIADDMNMX.S16x2", "VIADDMNMX", "VIMNMX", "FADD", "FMNMX", "DADD", "DSETP", "UBLKCP", "SYNCS", "USETMAXREG", "LDS", "STS", "LDG", "STG",
        "LD.E", "RED", "MEMBAR", "CCTL", "FENCE", "BAR", "UIADD3", "UISETP", "USHF", "BRA.U", "LDL", "STL", "CALL", "NANOSLEEP"]
print(f"# {os.path.basename(lib)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    ins = [re.sub(r"/\*.*?\*/", "", l).strip() for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,5}\*/", l)]
    ops = collections.Counter()
    for s in ins:
        s = re.sub(r"^@!?U?P\d+\s+", "", s)
        m = s.split()[0] if s else ""
        for k in KEYS:
            if m == k or m.startswith(k + "."):
                ops[k] += 1
    ops["VIADDMNMX"] -= ops["VIADDMNMX.S16x2"]
    print(f"\n{dem[:150]}\n  instructions {len(ins)}  " + "  ".join(f"{k}={v}" for k, v in ops.items() if v))
