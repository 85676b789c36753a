"""Compares a dump of the real oxipng (tools/ref_dump.rs) with tests/golden/expected.json -- the SHA-256 pins of the
oracle that every CUDA parity test is held to.  All-equal turns this repo's parity from "bit-exact against the restated
reference" into "bit-exact against the reference".

    python tools/compare_ref_dump.py /tmp/ref_dump
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
exp = json.load(open(os.path.join(HERE, "..", "tests", "golden", "expected.json")))
dump = sys.argv[1]


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest() if os.path.exists(path) else None


bad = checked = 0
for name, e in sorted(exp.items()):
    got = sha(os.path.join(dump, f"{name}.data.bin"))
    if got is None:
        print(f"{name}: not in the dump (skipped)")
        continue
    checked += 1
    if got != e["data_sha"]:
        print(f"{name}: decoded PngImage.data differs")
        bad += 1
    for key, v in e["filters"].items():
        sname, alpha = key.split("/")
        got = sha(os.path.join(dump, f"{name}.filter.{sname}.{alpha}.bin"))
        checked += 1
        if got != v["sha"]:
            print(f"{name}: filter_image({sname}, optimize_alpha={alpha}) differs from the oracle")
            bad += 1
    for key, v in e.get("reductions", {}).items():
        if not isinstance(v, dict) and v is not None:
            continue
        p = os.path.join(dump, f"{name}.reduction.{key.replace('/', '.')}.bin")
        got = sha(p)
        want = None if v is None else v.get("data_sha", v.get("sha"))
        if key.startswith(("most_popular", "cooccurrence")):
            continue
        checked += 1
        if (got is None) != (want is None) or (want is not None and got != want):
            print(f"{name}: {key} differs (reference {'None' if got is None else got[:12]}, oracle {'None' if want is None else want[:12]})")
            bad += 1
print(f"{checked} comparisons, {bad} differences")
sys.exit(1 if bad else 0)
