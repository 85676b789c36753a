"""Top stall-sample SASS instructions of an ncu report (source page): python profiles/hot_sass.py rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
data = []
for k, r in enumerate(rows[2:]):
    try: v = int(r[ci["# Samples"]])
    except Exception: continue
    data.append((v, k, r))
tot = sum(v for v, _, _ in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ci[h]] or 0) for _, _, r in data) for h in stalls}
print("stall mix:", {k: f"{100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda t: -t[1]) if v * 50 > tot})
for v, k, r in sorted(data, key=lambda t: -t[0])[:N]:
    top = sorted(((int(r[ci[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(f"{v:7d} {100*v/tot:5.1f}%  #{k:5d} {r[ci['Source']].strip()[:70]:70s} exec={r[ci['Instructions Executed']]:>9s} wf={r[ci['L1 Wavefronts Shared']]:>9s} {top}")
