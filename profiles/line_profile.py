"""Per-source-line instruction and stall-sample profile of one kernel of an ncu report.

    python profiles/line_profile.py <rep.ncu-rep> <lib.so> <kernel-mangled-substring> [N]

The ncu CSV source page is SASS-only, so the SASS offsets are joined with `nvdisasm -g` line info of the cubin
extracted from the shared library (built with -lineinfo).  Inlined code is attributed to the innermost line.
"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, lib, kern = sys.argv[1:4]
N = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
line_of = {}
for cub in os.listdir(tmp):
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    inside, cur = False, None
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(\S.*);", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the page lists every profiled launch one after another; keep the first
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
seen = set()
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    if base is None:
        base = a
    if a in seen:
        break
    seen.add(a)
    key = line_of.get(a - base)
    e = agg[key]
    e[0] += int(r[ci["Instructions Executed"]] or 0)
    e[1] += int(r[ci["# Samples"]] or 0)
    e[2] += int(r[ci["L1 Wavefronts Shared"]] or 0)
    for h in stalls:
        e[3][h] += int(r[ci[h]] or 0)
tot_i = sum(e[0] for e in agg.values()); tot_s = sum(e[1] for e in agg.values())
print(f"kernel {kern}: warp instructions {tot_i}, samples {tot_s}")
src = {}
for key, e in sorted(agg.items(), key=lambda t: -t[1][0])[:N]:
    text = ""
    if key:
        path = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", key[0])
        if path not in src and os.path.exists(path):
            src[path] = open(path).read().splitlines()
        if path in src and key[1] <= len(src[path]):
            text = src[path][key[1] - 1].strip()[:80]
    top = ", ".join(f"{h[6:]} {100 * v / max(e[1], 1):.0f}%" for h, v in e[3].most_common(2))
    print(f"{100 * e[0] / tot_i:5.1f}% inst {100 * e[1] / tot_s:5.1f}% smp  wf={e[2]:>10d}  {str(key):28s} {text:80s} [{top}]")
