"""GPU parity of the SURVEY 8f "next" rows: unfilter_image and Adam7 interlace / deinterlace (bit-exact vs the oracle)."""
import numpy as np
import pytest

import helpers
import oracle

pytestmark = pytest.mark.gpu

FORMATS = [(0, 1), (0, 2), (3, 4), (0, 8), (3, 8), (4, 8), (2, 8), (6, 8), (0, 16), (4, 16), (2, 16), (6, 16)]
SHAPES = [(1, 1), (1, 9), (2, 2), (3, 5), (5, 3), (8, 8), (9, 17), (33, 7), (40, 70), (257, 65)]


def test_interlace_and_deinterlace_match_oracle():
    import oxipng_b200 as ox
    for ct, bd in FORMATS:
        for (w, h) in SHAPES:
            ih, d = helpers.synth_image(w, h, ct, bd, 7 + w, "random")
            img = helpers.product_image(ih, d)
            want = oracle.interlace_image(ih, d)
            got = ox.interlace_image(img)
            assert got.ihdr.interlaced and np.array_equal(got.data, want), (ct, bd, w, h)
            back = ox.deinterlace_image(got)
            assert not back.ihdr.interlaced and np.array_equal(back.data, d), (ct, bd, w, h)
            assert np.array_equal(back.data, oracle.deinterlace_image((w, h, ct, bd, 1), want))
            assert ox.change_interlacing(img, False) is None and ox.change_interlacing(got, True) is None


def test_unfilter_inverts_every_strategy_and_matches_oracle():
    import oxipng_b200 as ox
    for ct, bd in FORMATS:
        for il in (False, True):
            for (w, h) in [(1, 1), (3, 5), (9, 17), (40, 70), (257, 65)]:
                ih, d = helpers.synth_image(w, h, ct, bd, 11 + h, "photo" if w > 8 else "random", il)
                hdr = helpers.product_image(ih, d).ihdr
                for name, kind, basic in helpers.STRATEGIES:
                    filtered, _, _ = oracle.filter_image(ih, d, kind, basic)
                    got = ox.unfilter_image(hdr, filtered)
                    assert np.array_equal(got, d), (ct, bd, il, w, h, name)
                    assert np.array_equal(got, oracle.unfilter_image(ih, filtered))


def test_unfilter_tile_pipeline_tall_images():
    """many 32-row tiles: long runs of rows that depend on the row above (tile-to-tile boundary words), runs broken by
    None/Sub rows (independent tiles), more tiles than pipeline warps, every pixel size"""
    import oxipng_b200 as ox
    rng = np.random.default_rng(5)
    cases = [(6, 8, 70, 700), (2, 8, 33, 1300), (0, 8, 300, 330), (4, 8, 64, 257), (6, 16, 45, 420), (2, 16, 31, 390),
             (3, 4, 100, 513), (4, 16, 9, 2100), (6, 8, 3, 25000)]
    for ct, bd, w, h in cases:
        ih, d = helpers.synth_image(w, h, ct, bd, 3 + w, "photo")
        hdr = helpers.product_image(ih, d).ihdr
        lists = [np.full(h, 4, np.uint8), np.full(h, 3, np.uint8), rng.integers(2, 5, h).astype(np.uint8),
                 rng.integers(0, 5, h).astype(np.uint8),
                 np.where(rng.random(h) < 0.03, 1, rng.integers(2, 5, h)).astype(np.uint8)]
        for i, pre in enumerate(lists):
            filtered, used, _ = oracle.filter_image(ih, d, oracle.PREDEFINED, 0, pre)
            assert np.array_equal(np.asarray(used), pre)
            got = ox.unfilter_image(hdr, filtered)
            assert np.array_equal(got, d), (ct, bd, w, h, i)
        il_ih, il_d = helpers.synth_image(w, min(h, 600), ct, bd, 5, "photo", True)
        filtered, _, _ = oracle.filter_image(il_ih, il_d, oracle.MINSUM, 0)
        assert np.array_equal(ox.unfilter_image(helpers.product_image(il_ih, il_d).ihdr, filtered), il_d), (ct, bd, w, h)


def test_unfilter_wide_rows_and_odd_sizes():
    """many chunks per row, rows that end inside a chunk / sub-chunk, one-pixel and one-row images, a stream with trailing
    bytes that do not form a scanline"""
    import oxipng_b200 as ox
    rng = np.random.default_rng(9)
    for ct, bd, w, h in [(6, 8, 5003, 37), (2, 8, 8191, 33), (0, 8, 20001, 5), (6, 16, 3001, 34), (2, 16, 2049, 40),
                         (4, 8, 4097, 3), (0, 1, 9999, 65), (3, 2, 777, 70), (6, 8, 1, 300), (2, 8, 2, 129), (6, 8, 31, 1),
                         (6, 8, 33, 32), (6, 8, 8, 33)]:
        ih, d = helpers.synth_image(w, h, ct, bd, 1 + h, "photo")
        hdr = helpers.product_image(ih, d).ihdr
        for pre in (rng.integers(0, 5, h).astype(np.uint8), rng.integers(2, 5, h).astype(np.uint8)):
            filtered, _, _ = oracle.filter_image(ih, d, oracle.PREDEFINED, 0, pre)
            assert np.array_equal(ox.unfilter_image(hdr, filtered), d), (ct, bd, w, h)
        # trailing bytes (less than one scanline): ignored like the reference's iterator
        extra = np.concatenate([filtered, np.array([3, 1, 2], np.uint8)[:min(3, max(1, d.size // h))]])
        want = oracle.unfilter_image(ih, extra)
        assert np.array_equal(ox.unfilter_image(hdr, extra), want), (ct, bd, w, h, "trailing")


def test_unfilter_concurrent_threads():
    """eight host threads decode different images at once: every call is a cooperative launch on its own stream (PngImage::new
    runs per file under rayon in the reference's directory mode)"""
    import concurrent.futures as cf
    import oxipng_b200 as ox
    shapes = [(300, 400, 6, 8), (64, 640, 2, 8), (1024, 160, 6, 8), (777, 90, 6, 16), (33, 700, 0, 8), (4000, 60, 2, 8),
              (128, 1280, 4, 8), (19, 2000, 6, 8)]
    cases = []
    for i, (w, h, ct, bd) in enumerate(shapes):
        ih, d = helpers.synth_image(w, h, ct, bd, 70 + i, "photo")
        filtered, _, _ = oracle.filter_image(ih, d, oracle.MINSUM if i % 2 else oracle.BASIC, 4)
        cases.append((helpers.product_image(ih, d).ihdr, filtered, d))

    def work(t):
        hdr, filtered, d = cases[t % len(cases)]
        return t, ox.unfilter_image(hdr, filtered)

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        for t, got in ex.map(work, range(64)):
            assert np.array_equal(got, cases[t % len(cases)][2]), t


def test_unfilter_batch_equals_single_images():
    """oxi_unfilter_batch: many streams of one IHDR in one launch (more items than CTA groups) == image by image"""
    import oxipng_b200 as ox
    for ct, bd, w, h, il, n in [(6, 8, 61, 70, False, 700), (2, 8, 33, 45, True, 90), (0, 16, 20, 33, False, 40), (6, 16, 17, 40, False, 25)]:
        streams, datas = [], []
        for i in range(n):
            ih, d = helpers.synth_image(w, h, ct, bd, 100 + i, "photo" if i % 3 else "random", il)
            kind = [oracle.MINSUM, oracle.ENTROPY, oracle.BIGRAMS][i % 3]
            filtered, _, _ = oracle.filter_image(ih, d, kind, 0)
            streams.append(filtered)
            datas.append(d)
        hdr = helpers.product_image(ih, d).ihdr
        got = ox.unfilter_batch(hdr, np.stack(streams))
        assert got.shape == (n, datas[0].size)
        for i in range(n):
            assert np.array_equal(got[i], datas[i]), (ct, bd, w, h, il, i)


def test_unfilter_golden_fixtures(golden):
    """the reference's own files: stream as stored in the PNG (whatever filters its encoder chose) -> pixels"""
    import oxipng_b200 as ox
    for c in golden:
        hdr = helpers.product_image(c.ihdr, c.data, c.palette, c.trns).ihdr
        filtered, _, _ = oracle.filter_image(c.ihdr, c.data, oracle.ENTROPY, 0)
        assert np.array_equal(ox.unfilter_image(hdr, filtered), c.data), c.name


def test_unfilter_rejects_bad_filter_byte():
    import oxipng_b200 as ox
    ih, d = helpers.synth_image(16, 4, 2, 8, 1, "random")
    filtered, _, _ = oracle.filter_image(ih, d, oracle.BASIC, 1)
    filtered = filtered.copy()
    filtered[49] = 9  # second row's filter byte
    with pytest.raises(ValueError):
        ox.unfilter_image(helpers.product_image(ih, d).ihdr, filtered)
    with pytest.raises(ValueError):
        oracle.unfilter_image(ih, filtered)
