"""Small cases that touch every kernel path; meant to be run under compute-sanitizer (memcheck / racecheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import helpers  # noqa: E402
import oracle  # noqa: E402
import oxipng_b200 as ox  # noqa: E402
from oxipng_b200 import reduction as R  # noqa: E402

names = [s[0] for s in helpers.STRATEGIES]
table = {s[0]: s for s in helpers.STRATEGIES}
cases = [(256, 12, 6, 8, False, "photo"), (250, 9, 2, 8, False, "random"), (128, 16, 6, 8, True, "photo"),
         (1024, 6, 6, 16, False, "random"), (300, 7, 0, 4, False, "random"), (2048, 5, 6, 8, False, "photo")]
for (w, h, ct, bd, il, kind) in cases:
    ih, d = helpers.synth_image(w, h, ct, bd, 3, kind, il)
    img = helpers.product_image(ih, d)
    for group in ([n] for n in names):
        outs, used = ox.filter_image_multi(img, [helpers.product_strategy(n) for n in group], False)
        want = oracle.filter_image(ih, d, table[group[0]][1], table[group[0]][2])
        assert np.array_equal(outs[0], want[0]) and np.array_equal(used[0], want[1]), (w, h, ct, bd, group)
    outs, used = ox.filter_image_multi(img, [helpers.product_strategy(n) for n in ("None", "Bigrams", "BigEnt")], False)
    if ct in (4, 6):
        ox.filter_image_multi(img, [helpers.product_strategy(n) for n in names], True)
    ox.unfilter_image(img.ihdr, outs[1])
    if not il:
        ox.deinterlace_image(ox.interlace_image(img))
ih, d = helpers.synth_image(64, 32, 6, 8, 9, "transparent")
img = helpers.product_image(ih, d)
R.cleaned_alpha_channel(img); R.reduced_alpha_channel(img, True); R.reduced_to_indexed(img, True); R.reduced_rgb_to_grayscale(img)
ih, d = helpers.synth_image(64, 32, 2, 16, 9, "random")
R.reduced_bit_depth_16_to_8(helpers.product_image(ih, d), True)
ih, d = helpers.synth_image(65, 33, 0, 8, 9, "flat")
g = helpers.product_image(ih, d)
R.reduced_bit_depth_8_or_less(g)
r = oracle.reduced_to_indexed(ih, d, True)
pi = helpers.product_image(r["ihdr"], r["data"], r["palette"])
R.reduced_palette(pi, False); R.sorted_palette(pi); R.CoOccurrenceMatrix.from_image(pi)
print("sanitizer cases ok")
