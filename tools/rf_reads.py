#!/usr/bin/env python
"""Register-operand reads of the pair kernel's evaluation loop, counted from the SASS (run here, no GPU):

    python tools/rf_reads.py [lib.so] [kernel-name-fragment]

The loop is the innermost backward branch of the kernel that encloses at least 40 packed FP32x2 instructions
(two tiles = four pair evaluations per lane and trip).  Per instruction: a 64-bit packed operand (``Rn.F32x2``) = 2 reads of
the register file, a 32-bit register = 1, uniform registers / immediates / constant-bank operands / operands
served by the reuse cache (``.reuse``) = 0.  tools/mixbench.cu measured the rate on B200: 2 reads per lane and
cycle (96 FFMA2 with three distinct operands: 3.0 cycles each; + 20 MUFU: + 0.5 each; + 56 LOP3: + 1.5 each)."""
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "cyclistsocialforce_b200", "libcsf_b200.so")
frag = sys.argv[2] if len(sys.argv) > 2 else "pair_tiled_kernelIfLb0ELi2E"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ins, on = [], False
for line in sass.splitlines():
    if "Function :" in line:
        on = frag in line
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if on and m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
best = None
for i, (addr, text) in enumerate(ins):
    m = re.search(r"BRA (0x[0-9a-f]+)", text)
    if m and int(m.group(1), 16) < addr:
        j = next(k for k, (a, _) in enumerate(ins) if a >= int(m.group(1), 16))
        packed = sum(1 for _, t in ins[j:i + 1] if re.search(r"\b(FFMA2|FMUL2|FADD2)\b", t))
        # the innermost such loop: at least 40 packed instructions, fewest instructions overall
        if packed >= 40 and (best is None or i - j < best[2] - best[1]):
            best = (packed, j, i)
packed, j, i = best
tot, per = 0, {}
for _, text in ins[j:i + 1]:
    text = re.sub(r"^@!?U?P\d+\s+", "", text)
    parts = text.split(None, 1)
    op = parts[0].split(".")[0]
    ops = [o.strip() for o in parts[1].split(",")] if len(parts) > 1 else []
    srcs = ops if op in ("STS", "BRA") else (ops[2:] if op in ("FSETP", "ISETP") else ops[1:])
    reads = 0
    for o in srcs:
        o2 = o.replace("-", "").replace("|", "").replace("~", "")
        if not re.match(r"^R\d+", o2) or ".reuse" in o:
            continue
        reads += 2 if "F32x2" in o else 1
    tot += reads
    n, r = per.get(op, (0, 0))
    per[op] = (n + 1, r + reads)
print(f"{frag}: evaluation loop = {i - j + 1} instructions, {packed} packed FP32x2, "
      f"{sum(n for o, (n, r) in per.items() if o == 'MUFU')} MUFU")
print(f"register operand reads per trip (4 pair evaluations per lane): {tot}  ->  {tot / 2:.0f} cycles per scheduler at 2 reads "
      f"per lane and cycle")
print(f"ceiling of the evaluation alone: {4 * 128 / (tot / 2) * 148 * 1.965e9 * 76 / 1e12:.1f} TFLOP/s of executed pair "
      f"arithmetic (76 flop per pair, 148 SMs, 1.965 GHz)")
for op, (n, r) in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print(f"  {op:8s} n = {n:3d}  reads = {r:4d}")
