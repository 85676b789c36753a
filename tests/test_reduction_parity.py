"""GPU parity of the reduction scans (src/reduction/*.rs) against the oracle: decisions (None vs Some), output bytes,
IHDR, palette and tRNS -- bit-exact."""
import numpy as np
import pytest

import helpers
import oracle

pytestmark = pytest.mark.gpu


def _same(got, want, label):
    if want is None:
        assert got is None, label
        return
    assert got is not None, label
    w, h, ct, bd, il = want["ihdr"]
    assert (got.ihdr.width, got.ihdr.height, got.ihdr.color_type.code, int(got.ihdr.bit_depth), int(got.ihdr.interlaced)) == \
        (w, h, ct, bd, il), label
    assert np.array_equal(got.data, want["data"]), (label, "data")
    if ct == 3:
        assert np.array_equal(got.ihdr.color_type.palette_array(), np.asarray(want["palette"]).reshape(-1, 4)), (label, "palette")
    if ct in (0, 2):
        wt = want.get("trns")
        gt = got.ihdr.color_type.transparent
        if wt is None:
            assert gt is None, (label, "trns")
        else:
            assert gt == (wt if ct == 0 else tuple(wt)), (label, "trns")


def _run_all(ihdr, data, palette, trns, label):
    from oxipng_b200 import reduction as R
    img = helpers.product_image(ihdr, data, palette, trns)
    ct = ihdr[2]
    for oa in (False, True):
        _same(R.reduced_alpha_channel(img, oa), oracle.reduced_alpha_channel(ihdr, data, oa), f"{label}/alpha/{oa}")
    _same(R.cleaned_alpha_channel(img), oracle.cleaned_alpha_channel(ihdr, data), f"{label}/clean")
    for fs in (False, True):
        want = oracle.reduced_bit_depth_16_to_8(ihdr, data, fs, trns=trns)
        _same(R.reduced_bit_depth_16_to_8(img, fs), want, f"{label}/16to8/{fs}")
    _same(R.reduced_bit_depth_8_or_less(img), oracle.reduced_bit_depth_8_or_less(ihdr, data, palette, trns if ct == 0 else None),
          f"{label}/8orless")
    _same(R.expanded_bit_depth_to_8(img), oracle.expanded_bit_depth_to_8(ihdr, data, palette, trns if ct == 0 else None),
          f"{label}/expand")
    for ag in (False, True):
        _same(R.reduced_to_indexed(img, ag), oracle.reduced_to_indexed(ihdr, data, ag, trns), f"{label}/indexed/{ag}")
    _same(R.reduced_rgb_to_grayscale(img), oracle.reduced_rgb_to_grayscale(ihdr, data, trns if ct == 2 else None),
          f"{label}/gray")
    if ct == 3 and ihdr[3] == 8:
        for oa in (False, True):
            _same(R.reduced_palette(img, oa), oracle.reduced_palette(ihdr, data, palette, oa), f"{label}/palette/{oa}")
        _same(R.sorted_palette(img), oracle.sorted_palette(ihdr, data, palette), f"{label}/sorted")
        assert R.most_popular_edge_color(len(palette), img) == oracle.most_popular_edge_color(ihdr, len(palette), data), label
        for ag, oa in ((False, False), (True, True)):
            _same(R.indexed_to_channels(img, ag, oa), oracle.indexed_to_channels(ihdr, data, palette, ag, oa),
                  f"{label}/to_channels")
        if len(palette) > 2 and int(data.max(initial=0)) < len(palette):
            m = R.CoOccurrenceMatrix.from_image(img)
            assert np.array_equal(m.data, oracle.cooccurrence_build(ihdr, len(palette), data)), (label, "cooccurrence")
        rm = np.arange(len(palette))[::-1].copy()
        if len(palette) > 1:
            _same(R.apply_palette_reorder(img, rm), oracle.apply_palette_reorder(ihdr, data, palette, rm), f"{label}/reorder")


def test_golden_fixture_reductions(golden):
    for c in golden:
        _run_all(c.ihdr, c.data, c.palette, c.trns, c.name)


def test_fixture_name_outcomes_on_gpu(golden):
    """What the reference's tests assert (tests/reduction.rs): output colour type / depth / palette length."""
    from oxipng_b200 import reduction as R
    by = {c.name: c for c in golden}
    def img(n):
        c = by[n]
        return helpers.product_image(c.ihdr, c.data, c.palette, c.trns)
    assert int(R.reduced_bit_depth_16_to_8(img("rgb_16_should_be_rgb_8"), False).ihdr.bit_depth) == 8
    assert R.reduced_rgb_to_grayscale(img("rgb_8_should_be_grayscale_8")).ihdr.color_type.code == 0
    assert R.reduced_alpha_channel(img("rgba_8_should_be_rgb_8"), False).ihdr.color_type.code == 2
    assert R.reduced_alpha_channel(img("rgba_8_should_be_rgb_trns_8"), True).ihdr.color_type.has_trns()
    for n, d in [("grayscale_8_should_be_grayscale_4", 4), ("grayscale_8_should_be_grayscale_2", 2),
                 ("grayscale_8_should_be_grayscale_1", 1)]:
        assert int(R.reduced_bit_depth_8_or_less(img(n)).ihdr.bit_depth) == d
    assert R.reduced_to_indexed(img("rgba_8_should_be_palette_8"), True).ihdr.color_type.code == 3
    for n, a, b in [("palette_should_be_reduced_with_dupes", 43, 35), ("palette_should_be_reduced_with_unused", 35, 33),
                    ("palette_should_be_reduced_with_both", 43, 33), ("palette_should_be_reduced_with_missing", 2, 3)]:
        i = img(n)
        assert len(i.ihdr.color_type.palette_array()) == a
        assert len(R.reduced_palette(i, False).ihdr.color_type.palette_array()) == b


def _palette_image(w, h, ncol, seed, channels, il=False, gray=False, alpha_mode="opaque"):
    """<= ncol colours drawn from a PRNG palette, chosen per 8x8 block with 2% salt noise."""
    rng = np.random.default_rng(seed)
    pal = rng.integers(0, 256, size=(ncol, 4), dtype=np.uint8)
    if gray:
        pal[:, 1] = pal[:, 0]
        pal[:, 2] = pal[:, 0]
    if alpha_mode == "opaque":
        pal[:, 3] = 255
    elif alpha_mode == "binary":
        pal[:, 3] = np.where(rng.random(ncol) < 0.3, 0, 255)
    blocks = rng.integers(0, ncol, size=((h + 7) // 8, (w + 7) // 8))
    idx = np.kron(blocks, np.ones((8, 8), dtype=np.int64))[:h, :w]
    salt = rng.random((h, w)) < 0.02
    idx[salt] = rng.integers(0, ncol, size=int(salt.sum()))
    px = pal[idx][..., :channels] if channels != 2 else pal[idx][..., [0, 3]]
    ct = {1: 0, 2: 4, 3: 2, 4: 6}[channels]
    ih = (w, h, ct, 8, 0)
    data = np.ascontiguousarray(px).reshape(-1)
    if il:
        data = oracle.interlace_image(ih, data)
        ih = (w, h, ct, 8, 1)
    return ih, data, pal, idx


@pytest.mark.parametrize("ncol", [1, 2, 5, 16, 200, 256, 257, 300])
def test_to_indexed_first_occurrence_order_and_overflow(ncol):
    for channels in (1, 2, 3, 4):
        ih, data, pal, idx = _palette_image(301, 157, ncol, 77 + ncol, channels)
        _run_all(ih, data, None, None, f"pal{ncol}/ch{channels}")


def test_c4_shapes_scaled():
    """config c4 (2048^2, <= 256 colours) at 512^2: alpha reducible, gray reducible, 16 colours, 16-bit hi==lo."""
    ih, data, pal, idx = _palette_image(512, 512, 256, 5, 4, alpha_mode="opaque")
    _run_all(ih, data, None, None, "c4/rgba-opaque")
    ih, data, pal, idx = _palette_image(512, 512, 256, 6, 4, alpha_mode="binary")
    _run_all(ih, data, None, None, "c4/rgba-binary")
    ih, data, pal, idx = _palette_image(512, 512, 200, 7, 3, gray=True)
    _run_all(ih, data, None, None, "c4/rgb-gray")
    ih, data, pal, idx = _palette_image(512, 512, 16, 8, 3, il=True)
    _run_all(ih, data, None, None, "c4/rgb16col-interlaced")
    # indexed image + its palette -> palette scans
    ih, data, pal, idx = _palette_image(512, 300, 16, 9, 4)
    r = oracle.reduced_to_indexed(ih, data, True)
    _run_all(r["ihdr"], r["data"], r["palette"], None, "c4/indexed16")
    for il in (False, True):
        ih2, d2, pal2, idx2 = _palette_image(333, 211, 61, 10, 4, il=il, alpha_mode="binary")
        r = oracle.reduced_to_indexed(ih2, d2, True)
        _run_all(r["ihdr"], r["data"], r["palette"], None, f"c4/indexed61/{il}")
    # 16-bit with hi == lo (16 -> 8 reducible) and a lossy one
    ih, d8 = helpers.synth_image(256, 128, 6, 8, 3, "photo")
    d16 = np.repeat(d8, 2)
    _run_all((256, 128, 6, 16, 0), d16, None, None, "c4/16-hi==lo")
    ih16, dd = helpers.synth_image(256, 128, 2, 16, 4, "random")
    _run_all(ih16, dd, None, None, "c4/16-random")


@pytest.mark.parametrize("bd", [1, 2, 4])
def test_sub_byte_expand_and_pack_roundtrip(bd):
    from oxipng_b200 import reduction as R
    for il in (False, True):
        for (w, h) in [(1, 1), (3, 5), (9, 4), (17, 17), (250, 33)]:
            for ct in (0, 3):
                ih, d = helpers.synth_image(w, h, ct, bd, 50 + w, "random", il)
                pal = np.arange(4 * (1 << bd), dtype=np.uint8).reshape(-1, 4) if ct == 3 else None
                _run_all(ih, d, pal, None, f"sub{bd}/{w}x{h}/{ct}/{il}")


def test_palette_scan_building_blocks():
    from oxipng_b200 import reduction as R
    rng = np.random.default_rng(1)
    d = rng.integers(0, 256, size=1_000_003, dtype=np.uint8)
    d[1000:200000] = 7
    assert np.array_equal(R.used_histogram256(d), np.bincount(d, minlength=256).astype(np.uint64))
    lut = rng.permutation(256).astype(np.uint8)
    assert np.array_equal(R.lut_remap(d, lut), lut[d])
