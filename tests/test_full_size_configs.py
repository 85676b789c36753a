"""BASELINE.json's configurations at FULL size on the GPU (inputs generated on the device, oxipng_b200/synth.py):
bit-exact against the CPU oracle where it finishes in seconds, and through size-independent properties elsewhere
(unfilter(filter(x)) == x through the GPU unfilter kernel, per-row filter ids in range, stream length, filter bytes)."""
import numpy as np
import pytest

import helpers
import oracle

pytestmark = pytest.mark.gpu


def _device_image(w, h, ch, bd, seed):
    import torch
    from oxipng_b200 import synth
    return synth.photo_image(w, h, ch, bd, seed, torch.device("cuda:0")).cpu().numpy()


def _props(ox, hdr, data, out, used, rows, row_len):
    assert out.size == data.size + rows and used.size == rows and int(used.max()) <= 4
    assert np.array_equal(out[::row_len + 1], used)  # every row starts with its filter id
    assert np.array_equal(ox.unfilter_image(hdr, out), data)  # round trip through the GPU unfilter kernel


def test_c1_minsum_rgba8_4096():
    import oxipng_b200 as ox
    w = h = 4096
    d = _device_image(w, h, 4, 8, 0xB2000001)
    d[:8 * w * 4] = 0  # rows 0-7 all zero: exercises png/mod.rs:398
    hdr = ox.IhdrData(w, h, ox.ColorType.RGBA(), ox.BitDepth.Eight)
    out, st = ox.PngImage(hdr, d).filter_image(ox.FilterStrategy.MinSum, False)
    used = np.array(st.lines, np.uint8)
    want, want_used, _ = oracle.filter_image((w, h, 6, 8, 0), d, oracle.MINSUM)
    assert np.array_equal(used, want_used) and np.array_equal(out, want)
    assert not used[:8].any()
    _props(ox, hdr, d, out, used, h, w * 4)


def test_c2_five_filters_and_entropy_rgb8_8192():
    import oxipng_b200 as ox
    w = h = 8192
    d = _device_image(w, h, 3, 8, 0xB2000002)
    hdr = ox.IhdrData(w, h, ox.ColorType.RGB(), ox.BitDepth.Eight)
    img = ox.PngImage(hdr, d)
    S = ox.FilterStrategy
    strat = [S.NONE, S.SUB, S.UP, S.AVERAGE, S.PAETH, S.Entropy]
    outs, used = ox.filter_image_multi(img, strat, False)
    for j, (kind, basic) in enumerate([(oracle.BASIC, f) for f in range(5)] + [(oracle.ENTROPY, 0)]):
        want, want_used, _ = oracle.filter_image((w, h, 2, 8, 0), d, kind, basic)
        assert np.array_equal(used[j], want_used), j
        assert np.array_equal(outs[j], want), j  # bit-exact filtered stream check of the named config
    _props(ox, hdr, d, outs[5], used[5], h, w * 3)


def test_c3a_bigrams_bigent_rgba16_4096():
    import oxipng_b200 as ox
    w = h = 4096
    d = _device_image(w, h, 4, 16, 0xB2000003)
    hdr = ox.IhdrData(w, h, ox.ColorType.RGBA(), ox.BitDepth.Sixteen)
    S = ox.FilterStrategy
    outs, used = ox.filter_image_multi(ox.PngImage(hdr, d), [S.Bigrams, S.BigEnt], False)
    # the oracle on 512 sampled rows of this 134 MB image (each needs only the previous raw row)
    L = w * 8
    rows = np.unique(np.concatenate([np.arange(0, 64), np.arange(64, h, 9)]))[:512]
    for j, kind in enumerate((oracle.BIGRAMS, oracle.BIGENT)):
        for r in rows[::4]:
            r = int(r)
            lo = max(r - 1, 0)
            sub = d[lo * L:(r + 1) * L]
            o, u, _ = oracle.filter_image((w, r + 1 - lo, 6, 16, 0), sub, kind)
            assert u[-1] == used[j][r], (j, r)
            assert np.array_equal(o[-(L + 1):], outs[j][r * (L + 1):(r + 1) * (L + 1)]), (j, r)
        _props(ox, hdr, d, outs[j], used[j], h, L)


def test_c3b_bigrams_bigent_adam7_rgba8_4096():
    import oxipng_b200 as ox
    w = h = 4096
    prog = _device_image(w, h, 4, 8, 0xB2000001)
    hdr_p = ox.IhdrData(w, h, ox.ColorType.RGBA(), ox.BitDepth.Eight)
    inter = ox.interlace_image(ox.PngImage(hdr_p, prog))      # GPU Adam7
    assert np.array_equal(inter.data, oracle.interlace_image((w, h, 6, 8, 0), prog))
    S = ox.FilterStrategy
    outs, used = ox.filter_image_multi(inter, [S.Bigrams, S.BigEnt], False)
    assert used[0].size == 7680  # 512+512+512+1024+1024+2048+2048 scanlines
    for j, kind in enumerate((oracle.BIGRAMS, oracle.BIGENT)):
        want, want_used, _ = oracle.filter_image((w, h, 6, 8, 1), inter.data, kind)
        assert np.array_equal(used[j], want_used) and np.array_equal(outs[j], want), j
        assert np.array_equal(ox.unfilter_image(inter.ihdr, outs[j]), inter.data)
    assert np.array_equal(ox.deinterlace_image(inter).data, prog)


def test_c4_reductions_2048():
    """reduction scans on 2048x2048 images with <= 256 colours (RGBA8 / RGB8 / indexed8), decisions and bytes vs the oracle"""
    from oxipng_b200 import reduction as R
    rng = np.random.default_rng(0xB2000004)
    pal = rng.integers(0, 256, size=(256, 4), dtype=np.uint8)
    pal[:, 3] = 255
    blocks = rng.integers(0, 256, size=(256, 256))
    idx = np.kron(blocks, np.ones((8, 8), dtype=np.int64))
    salt = rng.random((2048, 2048)) < 0.02
    idx[salt] = rng.integers(0, 256, size=int(salt.sum()))
    rgba = np.ascontiguousarray(pal[idx]).reshape(-1)
    ih = (2048, 2048, 6, 8, 0)
    img = helpers.product_image(ih, rgba)
    r = R.reduced_to_indexed(img, True)
    w = oracle.reduced_to_indexed(ih, rgba, True)
    assert np.array_equal(r.data, w["data"]) and np.array_equal(r.ihdr.color_type.palette_array(), w["palette"])
    a = R.reduced_alpha_channel(img, False)
    wa = oracle.reduced_alpha_channel(ih, rgba, False)
    assert a.ihdr.color_type.code == 2 and np.array_equal(a.data, wa["data"])
    assert R.reduced_rgb_to_grayscale(a) is None and oracle.reduced_rgb_to_grayscale(wa["ihdr"], wa["data"]) is None
    p = R.reduced_palette(r, False)
    wp = oracle.reduced_palette(w["ihdr"], w["data"], w["palette"], False)
    assert (p is None) == (wp is None)
    s = R.sorted_palette(r)
    ws = oracle.sorted_palette(w["ihdr"], w["data"], w["palette"])
    assert np.array_equal(s.data, ws["data"]) and np.array_equal(s.ihdr.color_type.palette_array(), ws["palette"])
    m = R.CoOccurrenceMatrix.from_image(r)
    assert np.array_equal(m.data, oracle.cooccurrence_build(w["ihdr"], 256, w["data"]))
    # 16 colours -> depth 4; 16-bit with hi == lo -> 8
    g = (idx % 16 * 17).astype(np.uint8).reshape(-1)
    gi = helpers.product_image((2048, 2048, 0, 8, 0), g)
    d4 = R.reduced_bit_depth_8_or_less(gi)
    w4 = oracle.reduced_bit_depth_8_or_less((2048, 2048, 0, 8, 0), g)
    assert int(d4.ihdr.bit_depth) == 4 and np.array_equal(d4.data, w4["data"])
    assert np.array_equal(R.expanded_bit_depth_to_8(d4).data, g)
    d16 = np.repeat(rgba, 2)
    r8 = R.reduced_bit_depth_16_to_8(helpers.product_image((2048, 2048, 6, 16, 0), d16), False)
    assert np.array_equal(r8.data, rgba)


def test_c5_whole_images_none_bigrams_bigent():
    """The headline workload itself: eight whole 1024x1024 RGBA8 images (two photographic, noisy, with all-zero rows, random)
    through the fused {None, Bigrams, BigEnt} launch (k_fast<1, SC_BIGRAM|SC_HALF>), every filtered stream and every
    per-row choice bit-exact against the oracle on all 8 x 1024 rows."""
    import concurrent.futures as cf
    import oxipng_b200 as ox
    w = h = 1024
    imgs = [_device_image(w, h, 4, 8, 0xB2000005 + i) for i in range(8)]
    rng = np.random.default_rng(55)
    for i, amp in ((2, 6), (3, 12), (4, 24)):  # extra noise on the colour channels: outlier hashes, flagged trials
        nz = rng.integers(-amp, amp + 1, size=imgs[i].size)
        nz[3::4] = 0
        imgs[i] = ((imgs[i].astype(np.int64) + nz) & 255).astype(np.uint8)
    imgs[5] = rng.integers(0, 256, size=imgs[5].size, dtype=np.uint8)  # every window an outlier, None competitive
    imgs[6][: 40 * w * 4] = 0                                          # all-zero rows (png/mod.rs:398-404)
    imgs[7] = np.repeat(rng.integers(0, 4, size=(h, w // 8, 4), dtype=np.uint8) * 60, 8, axis=1).reshape(-1)  # flat runs
    hdr = ox.IhdrData(w, h, ox.ColorType.RGBA(), ox.BitDepth.Eight)
    S = ox.FilterStrategy
    outs, used = ox.filter_batch(hdr, np.concatenate(imgs), 8, [S.NONE, S.Bigrams, S.BigEnt], False)
    per, rows = w * h * 4 + h, h
    kinds = [(oracle.BASIC, 0), (oracle.BIGRAMS, 0), (oracle.BIGENT, 0)]
    oracle.use_native()

    def check(t):
        i, j = t
        want, want_used, _ = oracle.filter_image((w, h, 6, 8, 0), imgs[i], *kinds[j])
        return (i, j, np.array_equal(used[j][i * rows:(i + 1) * rows], want_used),
                np.array_equal(outs[j][i * per:(i + 1) * per], want))

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        for i, j, ok_used, ok_out in ex.map(check, [(i, j) for i in range(8) for j in range(3)]):
            assert ok_used and ok_out, (i, j)
