#!/usr/bin/env python
"""SASS instruction summary per kernel of the in-tree library (run here, no GPU):

    python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt

Per kernel: instructions, packed FP32x2 (FFMA2/FMUL2/FADD2), scalar FP32 (FFMA/FMUL/FADD), FP64 (DFMA/DMUL/DADD),
MUFU, 1-D TMA bulk copies (UBLKCP), mbarrier operations (SYNCS), local-memory accesses (LDL/STL), registers."""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "cyclistsocialforce_b200", "libcsf_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*STACK:(\d+).*SHARED:(\d+)", line)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
GROUPS = [("FP32x2", r"\b(FFMA2|FMUL2|FADD2)\b"), ("FP32", r"\b(FFMA|FMUL|FADD)\b"), ("FP64", r"\b(DFMA|DMUL|DADD)\b"),
          ("MUFU", r"\bMUFU"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDL/STL", r"\b(LDL|STL)\b"),
          ("LDS/STS", r"\b(LDS|STS)\b"), ("LDG/STG", r"\b(LDG|STG)\b"), ("SHFL", r"\bSHFL"), ("ATOM", r"\b(ATOMS?|RED)\b")]
count = collections.defaultdict(lambda: collections.Counter())
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        count[cur]["n"] += 1
        for name, pat in GROUPS:
            if re.search(pat, line):
                count[cur][name] += 1
print("SASS summary of", os.path.basename(lib), "(sm_100a), kernels with >= 100 instructions\n")
hdr = ["instr"] + [g for g, _ in GROUPS] + ["regs", "stack"]
print(("%-92s" + " %8s" * len(hdr)) % tuple(["kernel"] + hdr))
for k in sorted(count, key=lambda k: demangle(k)):
    c = count[k]
    if c["n"] < 100:
        continue
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", demangle(k))
    name = re.sub(r"\((int|bool)\)", "", name)
    name = re.sub(r"\(.*", "", name)[:90]
    r = regs.get(k, (0, 0, 0))
    print(("%-92s" + " %8d" * len(hdr)) % tuple([name, c["n"]] + [c[g] for g, _ in GROUPS] + [r[0], r[1]]))
