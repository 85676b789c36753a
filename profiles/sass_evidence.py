"""SASS evidence for the claims in DESIGN.md, taken from the built library (no GPU needed):
    python profiles/sass_evidence.py > profiles/r02_c5_sass_excerpt.txt
Lists, for the dominant kernel k_fast<1, SC_BIGRAM|SC_HALF>, the TMA bulk copies (UBLKCP) and their mbarrier traffic (SYNCS),
the table updates (ATOMS.POPC.INC with an immediate shared-window address), the byte-SIMD instruction (VABSDIFF4), the
128-bit streaming stores, and the instruction mix of the kernel.
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "oxipng_b200", "liboxipng_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, ins = None, []
for ln in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    if cur and "k_fastILi1ELi12E" in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
print("kernel: k_fast<D=1, SC_BIGRAM|SC_HALF>  (", len(ins), "SASS instructions )")
pat = re.compile(r"UBLKCP|SYNCS|ATOMS|STG\.E\.EF\.128|BAR\.SYNC|REDUX")
seen = collections.Counter()
for a, t in ins:
    m = pat.search(t)
    if m:
        key = (t.split()[1] if t.startswith("@") else t.split()[0]) + ("imm" if "+0x" in t and "ATOMS" in t else "")
        seen[key] += 1
        if seen[key] <= 3:
            print(f"  /*{a:05x}*/  {t}")
mix = collections.Counter()
for a, t in ins:
    op = t.split()[1] if t.startswith("@") else t.split()[0]
    mix[op.split(".")[0]] += 1
print("instruction mix of the whole kernel:", ", ".join(f"{k} {v}" for k, v in mix.most_common(14)))
