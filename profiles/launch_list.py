"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: python profiles/launch_list.py x.csv"""
import collections, csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[unit]
    a = agg.setdefault(r[ci["Kernel Name"]], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("launches   total ms   share  kernel")
for k, a in sorted(agg.items(), key=lambda t: -t[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print(f"{a[0]:8d} {a[1]:10.3f} {100 * a[1] / tot:6.1f}%  {k[:100]}")
