"""C4: the reduction scans on 2048x2048 images with <= 256 colours (BASELINE configs[3]).

    python profiles/time_reductions.py            host-call times (H2D + kernels + D2H, pageable host buffers)
    ncu --metrics gpu__time_duration.sum --csv --log-file x.csv python profiles/time_reductions.py --once
        -> kernel-only times (python profiles/launch_list.py x.csv)
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402

import helpers  # noqa: E402
import oxipng_b200 as ox  # noqa: E402
from oxipng_b200 import reduction as R  # noqa: E402

once = "--once" in sys.argv
rng = np.random.default_rng(0xB2000004)
pal = rng.integers(0, 256, size=(256, 4), dtype=np.uint8)
pal[:, 3] = 255
blocks = rng.integers(0, 256, size=(256, 256))
idx = np.kron(blocks, np.ones((8, 8), dtype=np.int64))
salt = rng.random((2048, 2048)) < 0.02
idx[salt] = rng.integers(0, 256, size=int(salt.sum()))
rgba = np.ascontiguousarray(pal[idx]).reshape(-1)
img = helpers.product_image((2048, 2048, 6, 8, 0), rgba)
indexed = R.reduced_to_indexed(img, True)
rgb = R.reduced_alpha_channel(img, False)
gray = helpers.product_image((2048, 2048, 0, 8, 0), (idx % 16 * 17).astype(np.uint8).reshape(-1))
gray_rgb = helpers.product_image((2048, 2048, 2, 8, 0), np.repeat((idx % 256).astype(np.uint8).reshape(-1), 3))
d4 = R.reduced_bit_depth_8_or_less(gray)
rgba16 = helpers.product_image((2048, 2048, 6, 16, 0), np.repeat(rgba, 2))
N = rgba.size
cases = [
    ("reduced_to_indexed (RGBA8 -> 256 colours)", lambda: R.reduced_to_indexed(img, True), N + N // 4),
    ("reduced_alpha_channel (RGBA8 -> RGB8)", lambda: R.reduced_alpha_channel(img, False), N + N + 3 * N // 4),
    ("cleaned_alpha_channel (RGBA8)", lambda: R.cleaned_alpha_channel(img), 2 * N),
    ("reduced_rgb_to_grayscale (RGB8 gray -> G8)", lambda: R.reduced_rgb_to_grayscale(gray_rgb), gray_rgb.data.size * 4 // 3),
    ("reduced_bit_depth_16_to_8 (RGBA16 -> RGBA8)", lambda: R.reduced_bit_depth_16_to_8(rgba16, False), 3 * N),
    ("reduced_bit_depth_8_or_less (G8 -> G4)", lambda: R.reduced_bit_depth_8_or_less(gray), 2 * gray.data.size + gray.data.size // 2),
    ("expanded_bit_depth_to_8 (G4 -> G8)", lambda: R.expanded_bit_depth_to_8(d4), d4.data.size + gray.data.size),
    ("reduced_palette (indexed8)", lambda: R.reduced_palette(indexed, False), indexed.data.size),
    ("sorted_palette (indexed8)", lambda: R.sorted_palette(indexed), 2 * indexed.data.size),
    ("CoOccurrenceMatrix::from (indexed8)", lambda: R.CoOccurrenceMatrix.from_image(indexed), indexed.data.size),
]
for name, fn, alg in cases:
    fn()
    if once:
        continue
    ts = []
    for _ in range(5):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    ms = sorted(ts)[len(ts) // 2] * 1e3
    print(f"{name:48s} host call {ms:7.2f} ms   algorithmic {alg / 1e6:7.1f} MB")
