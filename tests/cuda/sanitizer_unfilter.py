"""k_unfilter under compute-sanitizer: tall images (tile pipeline, poller warp, mailbox ring), every pixel size, a batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import helpers  # noqa: E402
import oracle  # noqa: E402
import oxipng_b200 as ox  # noqa: E402

rng = np.random.default_rng(1)
for ct, bd, w, h in [(6, 8, 70, 200), (2, 8, 33, 150), (0, 8, 100, 130), (6, 16, 20, 140), (2, 16, 15, 100), (3, 4, 60, 129)]:
    ih, d = helpers.synth_image(w, h, ct, bd, 3, "photo")
    hdr = helpers.product_image(ih, d).ihdr
    for pre in (rng.integers(2, 5, h).astype(np.uint8), rng.integers(0, 5, h).astype(np.uint8)):
        filtered, _, _ = oracle.filter_image(ih, d, oracle.PREDEFINED, 0, pre)
        assert np.array_equal(ox.unfilter_image(hdr, filtered), d), (ct, bd, w, h)
streams = []
for i in range(40):
    ih, d = helpers.synth_image(40, 70, 6, 8, 50 + i, "photo")
    streams.append(oracle.filter_image(ih, d, oracle.MINSUM, 0)[0])
got = ox.unfilter_batch(helpers.product_image(ih, d).ihdr, np.stack(streams))
assert np.array_equal(got[-1], d)
print("sanitizer unfilter cases ok")
