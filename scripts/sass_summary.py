#!/usr/bin/env python
"""Static SASS summary of every kernel in libobm_b200.so (cuobjdump; no GPU needed): registers, stack, shared memory,
instruction count and the mnemonics that say how the kernel works (FP64 pipe, MUFU, shuffles, cp.async, reductions) —
and which tensor-core / TMA mnemonics are ABSENT (nothing on this path is a contraction; TMA was measured and removed).
usage: sass_summary.py > profiles/<tag>_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oceanbiome.jl_b200", "lib", "libobm_b200.so")
CU = "/usr/local/cuda/bin/cuobjdump"
GROUPS = [("FP64", ("DFMA", "DMUL", "DADD", "DSETP")), ("FP32", ("FFMA", "FMUL", "FADD", "FSETP")), ("MUFU", ("MUFU",)),
          ("SHFL", ("SHFL",)), ("LDG", ("LDG",)), ("STG", ("STG",)), ("RED", ("RED", "REDG", "ATOMG", "ATOMS")),
          ("LDGSTS", ("LDGSTS",)), ("LDS/STS", ("LDS", "STS")), ("LDC", ("LDC", "LDCU")), ("LDL/STL", ("LDL", "STL")),
          ("BAR", ("BAR",)), ("tcgen05/TMA", ("UTCMMA", "UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "DMMA"))]

res = subprocess.run([CU, "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
usage = {n: l for n, l in re.findall(r"Function (\S+):\n\s*(REG:.*)", res)}
sass = subprocess.run([CU, "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
print("# static SASS summary of oceanbiome.jl_b200/lib/libobm_b200.so (sm_100a), scripts/sass_summary.py")
print("# columns: instructions | " + " ".join(g for g, _ in GROUPS) + " | REG STACK SHARED")
for part in sass.split("\t\tFunction : ")[1:]:
    name = part.split("\n", 1)[0].strip()
    ops = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", part)
    c = collections.Counter(ops)
    counts = [sum(v for k, v in c.items() if k in names) for _, names in GROUPS]
    u = usage.get(name, "")
    r = {k: v for k, v in re.findall(r"(\w+)(?:\[0\])?:(\d+)", u)}
    print(f"{demangle(name)[:78]:78s} {len(ops):6d} | " + " ".join(f"{x:5d}" for x in counts)
          + f" | {r.get('REG', '?'):>3s} {r.get('STACK', '?'):>4s} {r.get('SHARED', '?'):>6s}")
