"""Generates tests/golden/fixtures.npz + tests/golden/expected.json.

Run in the build container (needs /root/reference/tests/files, which does not exist on the GPU box):

    python tests/golden/make_golden.py

fixtures.npz holds a subset of the reference's own test images (tests/files/*.png, the inputs of
tests/filters.rs, tests/reduction.rs, tests/interlaced.rs, tests/regression.rs, ...) decoded into the reference's
in-memory layout (PngImage.data: unfiltered scanlines; see tests/pngload.py), so that GPU tests can run without
the reference tree.  expected.json pins, per fixture x strategy x optimize_alpha, the SHA-256 of the oracle's
filtered stream and the per-row filter choices, plus the outcome of every reduction -- a regression pin of the
oracle itself (the reference has no byte-level vectors for this path; see oracle/oxi_oracle.h).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from pngload import load_png  # noqa: E402

REF_FILES = "/root/reference/tests/files"

SUBSET = [
    # every colour type / depth the filter tests cover (tests/filters.rs)
    "filter_5_for_rgba_16", "filter_5_for_rgba_8", "filter_5_for_rgb_16", "filter_5_for_rgb_8",
    "filter_5_for_grayscale_alpha_16", "filter_5_for_grayscale_alpha_8", "filter_5_for_grayscale_16",
    "filter_5_for_grayscale_8", "filter_5_for_palette_4", "filter_5_for_palette_2", "filter_5_for_palette_1",
    "filter_4_for_rgba_8", "filter_3_for_rgb_8",
    # the bench image (benches/strategies.rs)
    "rgb_8_should_be_rgb_8",
    # interlaced (tests/interlaced.rs)
    "interlaced_rgba_8_should_be_rgba_8", "interlaced_rgb_16_should_be_rgb_16", "interlaced_rgba_16_should_be_rgba_16",
    "interlaced_palette_4_should_be_palette_4", "interlaced_palette_1_should_be_palette_1",
    "interlaced_grayscale_alpha_16_should_be_grayscale_alpha_16", "interlaced_palette_8_should_be_palette_8",
    "interlaced_small_files", "interlaced_vertical_filters", "interlaced_palette_2_should_be_palette_2",
    # alpha (tests/alpha.rs, *_reduce_alpha)
    "rgba_8_reduce_alpha", "rgba_16_reduce_alpha", "grayscale_alpha_8_reduce_alpha", "grayscale_alpha_16_reduce_alpha",
    "rgba_8_should_be_rgb_trns_8", "rgba_16_should_be_rgb_trns_16", "grayscale_alpha_8_should_be_grayscale_trns_8",
    # reductions (tests/reduction.rs)
    "rgba_8_should_be_palette_8", "rgba_8_should_be_grayscale_alpha_8", "rgb_16_should_be_rgb_8",
    "rgb_8_should_be_grayscale_8", "rgba_16_should_be_grayscale_8", "grayscale_8_should_be_grayscale_4",
    "grayscale_8_should_be_grayscale_2", "grayscale_8_should_be_grayscale_1", "grayscale_4_should_be_grayscale_2",
    "grayscale_2_should_be_grayscale_1", "grayscale_16_should_be_grayscale_8", "rgb_trns_8_should_be_palette_8",
    "grayscale_trns_8_should_be_grayscale_1", "palette_8_should_be_palette_4", "palette_8_should_be_palette_2",
    "palette_4_should_be_palette_2", "palette_should_be_reduced_with_dupes", "palette_should_be_reduced_with_unused",
    "palette_should_be_reduced_with_missing", "palette_should_be_reduced_with_both", "palette_8_should_be_rgba",
    "palette_8_should_be_grayscale_8", "rgb_8_should_be_palette_4", "rgba_8_should_be_rgb_8",
    # regressions / tiny images (tests/regression.rs)
    "issue-42", "issue-56", "issue-58", "issue-59", "issue-60", "issue-89", "issue-140", "issue-171", "issue-175",
    "issue-182", "small_files", "interlacing_0_to_1_small_files", "interlacing_1_to_0_small_files",
]

STRATEGIES = [("None", oracle.BASIC, 0), ("Sub", oracle.BASIC, 1), ("Up", oracle.BASIC, 2), ("Average", oracle.BASIC, 3),
              ("Paeth", oracle.BASIC, 4), ("MinSum", oracle.MINSUM, 0), ("Entropy", oracle.ENTROPY, 0),
              ("Bigrams", oracle.BIGRAMS, 0), ("BigEnt", oracle.BIGENT, 0)]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def reduction_outcomes(p) -> dict:
    """Outcome (None / summary) of every reduction on this fixture, with the reference's default-ish flags."""
    ih, d, pal, trns = p.ihdr, p.data, p.palette, p.trns
    out = {}

    def summ(r):
        if r is None:
            return None
        s = {"ihdr": [int(x) for x in r["ihdr"]], "sha": sha(r["data"]), "len": int(len(r["data"]))}
        if r.get("palette") is not None:
            s["palette"] = np.asarray(r["palette"]).reshape(-1).tolist()
        if r.get("trns") is not None:
            s["trns"] = r["trns"] if isinstance(r["trns"], int) else list(r["trns"])
        return s

    for oa in (False, True):
        out[f"reduced_alpha_channel/{int(oa)}"] = summ(oracle.reduced_alpha_channel(ih, d, oa))
    out["cleaned_alpha_channel"] = summ(oracle.cleaned_alpha_channel(ih, d))
    out["reduced_bit_depth_16_to_8/0"] = summ(oracle.reduced_bit_depth_16_to_8(ih, d, False))
    out["reduced_bit_depth_16_to_8/1"] = summ(oracle.reduced_bit_depth_16_to_8(ih, d, True))
    out["reduced_bit_depth_8_or_less"] = summ(oracle.reduced_bit_depth_8_or_less(ih, d, pal, trns if ih[2] == 0 else None))
    out["expanded_bit_depth_to_8"] = summ(oracle.expanded_bit_depth_to_8(ih, d, pal, trns if ih[2] == 0 else None))
    for ag in (False, True):
        out[f"reduced_to_indexed/{int(ag)}"] = summ(oracle.reduced_to_indexed(ih, d, ag, trns))
    out["reduced_rgb_to_grayscale"] = summ(oracle.reduced_rgb_to_grayscale(ih, d, trns if ih[2] == 2 else None))
    if ih[2] == 3 and ih[3] == 8:
        for oa in (False, True):
            out[f"reduced_palette/{int(oa)}"] = summ(oracle.reduced_palette(ih, d, pal, oa))
        out["sorted_palette"] = summ(oracle.sorted_palette(ih, d, pal))
        e = oracle.most_popular_edge_color(ih, len(pal), d)
        out["most_popular_edge_color"] = e
        if len(pal) > 2 and int(d.max(initial=0)) < len(pal):
            out["cooccurrence"] = sha(oracle.cooccurrence_build(ih, len(pal), d))
    return out


def main():
    arrays = {}
    expected = {}
    names = []
    for name in SUBSET:
        path = os.path.join(REF_FILES, name + ".png")
        p = load_png(path)
        names.append(name)
        arrays[name + "/ihdr"] = np.asarray(p.ihdr, dtype=np.int64)
        arrays[name + "/data"] = p.data
        if p.palette is not None:
            arrays[name + "/palette"] = p.palette
        if p.trns is not None:
            arrays[name + "/trns"] = np.asarray(p.trns, dtype=np.int64).reshape(-1)
        exp = {"data_sha": sha(p.data), "filters": {}, "reductions": reduction_outcomes(p)}
        for sname, kind, basic in STRATEGIES:
            for oa in (False, True):
                out, used, _ = oracle.filter_image(p.ihdr, p.data, kind, basic, None, oa)
                exp["filters"][f"{sname}/{int(oa)}"] = {"sha": sha(out), "used": bytes(used).hex()}
        expected[name] = exp
    arrays["names"] = np.asarray(names)
    np.savez_compressed(os.path.join(HERE, "fixtures.npz"), **arrays)
    with open(os.path.join(HERE, "expected.json"), "w") as f:
        json.dump(expected, f, separators=(",", ":"), sort_keys=True)
    print("fixtures:", len(names), "npz bytes:", os.path.getsize(os.path.join(HERE, "fixtures.npz")),
          "json bytes:", os.path.getsize(os.path.join(HERE, "expected.json")))


if __name__ == "__main__":
    main()
