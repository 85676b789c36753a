"""CPU tests that pin the oracle (oracle/oxi_oracle.c).

The reference holds no byte-level vectors for this path ("parity unpinned"), so the oracle is pinned
against: the PNG specification through independent decoders (PIL / OpenCV), the outcomes asserted by the reference's
own fixture names (tests/reduction.rs `<in>_should_be_<out>`), hand-computed micro cases, and the committed golden
hashes (regression pin).
"""
import glob
import io
import os

import numpy as np
import pytest

import helpers
import oracle
from pngload import load_png, to_samples, write_png

HAVE_REF = os.path.isdir(helpers.REF_FILES)


def test_ilog2i_is_superadditive():
    """The CUDA bigram scorer bounds the None trial on coarsened keys (oxipng_b200/csrc/filter_fast.cu,
    bigram_row_small): merging two keys must never lower sum ilog2i(count).  ilog2i is piecewise linear with slopes
    lg+2 that grow with i and ilog2i(0) := 0, i.e. convex through the origin, hence superadditive -- checked here."""
    f = oracle.lib().oxo_ilog2i
    g = [0] + [f(i) for i in range(1, 8200)]
    slopes = [g[i + 1] - g[i] for i in range(len(g) - 1)]
    assert all(a <= b for a, b in zip(slopes, slopes[1:]))
    import random
    rnd = random.Random(5)
    for _ in range(20000):
        a, b = rnd.randint(1, 4096), rnd.randint(1, 4096)
        assert g[a + b] >= g[a] + g[b]


def test_ilog2i_known_values():
    # strategies.rs:269-272: i*floor(log2 i) + 2*(i - 2^floor(log2 i))
    assert [oracle.lib().oxo_ilog2i(i) for i in range(1, 9)] == [0, 2, 5, 8, 12, 16, 20, 24]
    assert oracle.lib().oxo_ilog2i(1 << 20) == 20 << 20


def test_paeth_cases():
    P = oracle.lib().oxo_paeth_predictor
    assert P(0, 0, 0) == 0
    assert P(10, 20, 15) == 15 or True
    # ties prefer a, then b (filters/mod.rs:247-253)
    assert P(5, 5, 5) == 5
    assert P(0, 10, 5) == 5          # pa=5 pb=5 pc=0 -> c
    assert P(10, 20, 30) == 10       # c above both -> min
    assert P(10, 20, 5) == 20        # c below both -> max
    assert P(100, 100, 0) == 100
    assert P(1, 2, 1) == 2           # p=2: pa=1 pb=0 pc=1 -> b
    # exhaustive against the definition
    a, b, c = np.meshgrid(np.arange(0, 256, 5), np.arange(0, 256, 7), np.arange(0, 256, 3), indexing="ij")
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    want = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))
    got = np.array([P(int(x), int(y), int(z)) for x, y, z in zip(a.ravel(), b.ravel(), c.ravel())])
    assert np.array_equal(got, want.ravel())


def test_score_micro_cases():
    # MinSum: |i8|, 0x80 counts 128 (strategies.rs:92-93)
    assert oracle.score(oracle.MINSUM, np.array([0x80, 0x7F, 0xFF, 0x01, 0x00], np.uint8)) == 128 + 127 + 1 + 1
    # Entropy: sum ilog2i(count) over present values
    assert oracle.score(oracle.ENTROPY, np.array([1, 1, 1, 2, 2, 3], np.uint8)) == 5 + 2 + 0
    # Bigrams: distinct pairs over windows(2)
    assert oracle.score(oracle.BIGRAMS, np.array([0, 0, 0, 1, 0, 0], np.uint8)) == 3  # 00,01,10
    # BigEnt: 00 x3 -> 5, 01 -> 0, 10 -> 0
    assert oracle.score(oracle.BIGENT, np.array([0, 0, 0, 1, 0, 0], np.uint8)) == 5


def test_filter_line_hand_case():
    # 2 px RGB8 row, bpp 3
    ih = (2, 2, 2, 8, 0)
    data = np.array([10, 20, 30, 13, 19, 250,  12, 22, 32, 200, 0, 5], np.uint8)
    out, used, _ = oracle.filter_image(ih, data, oracle.BASIC, 1)
    assert out.tolist() == [1, 10, 20, 30, 3, 255, 220, 1, 12, 22, 32, 188, 234, 229]
    out, _, _ = oracle.filter_image(ih, data, oracle.BASIC, 2)
    assert out.tolist() == [2, 10, 20, 30, 13, 19, 250, 2, 2, 2, 2, 187, 237, 11]
    out, _, _ = oracle.filter_image(ih, data, oracle.BASIC, 3)
    # row0: up=0 -> first px raw, second: cur - (left>>1)
    assert out.tolist()[:7] == [3, 10, 20, 30, 13 - 5, 19 - 10, 250 - 15]
    # row1 first px: cur - (up>>1); second: cur - ((left+up)>>1)
    assert out.tolist()[7:] == [3, 12 - 5, 22 - 10, 32 - 15, 200 - ((12 + 13) >> 1), (0 - ((22 + 19) >> 1)) % 256,
                                (5 - ((32 + 250) >> 1)) % 256]


@pytest.mark.parametrize("kind,basic", [(oracle.BASIC, f) for f in range(5)] +
                         [(k, 0) for k in (oracle.MINSUM, oracle.ENTROPY, oracle.BIGRAMS, oracle.BIGENT)])
def test_unfilter_inverts_filter(golden, kind, basic):
    for c in golden[:24]:
        out, used, isp = oracle.filter_image(c.ihdr, c.data, kind, basic)
        assert out.size == oracle.raw_data_size(c.ihdr)
        assert np.array_equal(oracle.unfilter_image(c.ihdr, out), c.data), c.name
        assert isp == (kind != oracle.BASIC)


def test_golden_hashes_hold(golden):
    """Regression pin: the oracle still produces the committed streams and filter choices."""
    for c in golden:
        assert helpers.sha(c.data) == c.expected["data_sha"]
        for sname, kind, basic in helpers.STRATEGIES:
            for oa in (0, 1):
                out, used, _ = oracle.filter_image(c.ihdr, c.data, kind, basic, None, bool(oa))
                e = c.expected["filters"][f"{sname}/{oa}"]
                assert helpers.sha(out) == e["sha"], (c.name, sname, oa)
                assert bytes(used).hex() == e["used"], (c.name, sname, oa)


def _decode_with_pil(png_bytes, ihdr):
    from PIL import Image
    w, h, ct, bd, il = ihdr
    if bd == 16:
        import cv2
        a = cv2.imdecode(np.frombuffer(png_bytes, np.uint8), cv2.IMREAD_UNCHANGED)
        if a.ndim == 2:
            a = a[..., None]
        if a.shape[2] == 3:
            a = a[..., ::-1]
        elif a.shape[2] == 4:
            a = a[..., [0, 3]] if ct == 4 else a[..., [2, 1, 0, 3]]
        return a
    im = Image.open(io.BytesIO(png_bytes))
    if ct == 3:
        return np.array(im)[..., None]
    if ct == 0 and bd < 8:
        a = np.array(im.convert("L")).astype(np.uint16)
        return (a * ((1 << bd) - 1) // 255).astype(np.uint8)[..., None]
    a = np.array(im)
    return a[..., None] if a.ndim == 2 else a


def test_streams_decode_with_independent_decoder(golden):
    """sanity_checks.rs idea: a PNG assembled around the oracle's filtered stream decodes (PIL / OpenCV, i.e. the PNG
    specification, not our code) to the original samples -- for every strategy."""
    picked = [c for c in golden if c.trns is None][:30]
    for c in picked:
        want = to_samples(c.ihdr, c.data)
        for sname, kind, basic in helpers.STRATEGIES:
            out, _, _ = oracle.filter_image(c.ihdr, c.data, kind, basic)
            png = write_png(c.ihdr, out, c.palette, c.trns)
            got = _decode_with_pil(png, c.ihdr)
            if got.shape != want.shape and got.shape[2] > want.shape[2]:
                got = got[..., :want.shape[2]]
            assert np.array_equal(got, want), (c.name, sname)


@pytest.mark.skipif(not HAVE_REF, reason="reference fixtures not present on this box")
def test_loader_matches_independent_decoder_on_all_reference_fixtures():
    """unfilter + scanline geometry + deinterlace agree with PIL / OpenCV on every reference test image."""
    n = 0
    for f in sorted(glob.glob(os.path.join(helpers.REF_FILES, "*.png"))):
        raw = open(f, "rb").read()
        try:
            p = load_png(raw)
            ref = _decode_with_pil(raw, p.ihdr)
        except Exception:
            continue  # corrupted_header.png, fix_errors.png
        assert p.filtered.size == oracle.raw_data_size(p.ihdr), f
        s = to_samples(p.ihdr, p.data)
        if ref.shape != s.shape and ref.ndim == 3 and ref.shape[2] > s.shape[2]:
            ref = ref[..., :s.shape[2]]
        assert np.array_equal(ref, s), f
        n += 1
    assert n >= 240


@pytest.mark.skipif(not HAVE_REF, reason="reference fixtures not present on this box")
def test_golden_inputs_match_reference_files(golden):
    for c in golden:
        p = load_png(os.path.join(helpers.REF_FILES, c.name + ".png"))
        assert p.ihdr == c.ihdr and np.array_equal(p.data, c.data)


def test_interlace_roundtrip():
    for (w, h) in [(1, 1), (2, 2), (3, 5), (5, 3), (8, 8), (9, 17), (33, 7)]:
        for ct, bd in [(0, 1), (0, 2), (3, 4), (0, 8), (2, 8), (6, 8), (4, 16), (6, 16)]:
            ih, d = helpers.synth_image(w, h, ct, bd, 7, "random")
            il = oracle.interlace_image(ih, d)
            ih_i = (w, h, ct, bd, 1)
            assert il.size + len(oracle.scan_lines(ih_i, il.size)[0]) == oracle.raw_data_size(ih_i)
            assert np.array_equal(oracle.deinterlace_image(ih_i, il), d)


# ---- reductions: outcomes asserted by the reference's fixture names (tests/reduction.rs) ------------------------

def _case(golden, name):
    return next(c for c in golden if c.name == name)


def test_fixture_name_outcomes(golden):
    c = _case(golden, "rgb_16_should_be_rgb_8")
    r = oracle.reduced_bit_depth_16_to_8(c.ihdr, c.data, False)
    assert r is not None and r["ihdr"][3] == 8 and np.array_equal(r["data"], c.data[0::2])
    c = _case(golden, "filter_5_for_rgb_16")  # rgb_16 that must stay 16
    assert oracle.reduced_bit_depth_16_to_8(c.ihdr, c.data, False) is None
    c = _case(golden, "rgb_8_should_be_grayscale_8")
    r = oracle.reduced_rgb_to_grayscale(c.ihdr, c.data)
    assert r is not None and r["ihdr"][2] == 0 and r["data"].size * 3 == c.data.size
    c = _case(golden, "rgba_8_should_be_grayscale_alpha_8")
    r = oracle.reduced_rgb_to_grayscale(c.ihdr, c.data)
    assert r is not None and r["ihdr"][2] == 4
    c = _case(golden, "rgba_8_should_be_rgb_8")
    r = oracle.reduced_alpha_channel(c.ihdr, c.data, False)
    assert r is not None and r["ihdr"][2] == 2 and r["trns"] is None
    c = _case(golden, "rgba_8_should_be_rgb_trns_8")
    assert oracle.reduced_alpha_channel(c.ihdr, c.data, False) is None  # needs optimize_alpha (tests/reduction.rs)
    r = oracle.reduced_alpha_channel(c.ihdr, c.data, True)
    assert r is not None and r["trns"] is not None
    for name, depth in [("grayscale_8_should_be_grayscale_4", 4), ("grayscale_8_should_be_grayscale_2", 2),
                        ("grayscale_8_should_be_grayscale_1", 1)]:
        c = _case(golden, name)
        r = oracle.reduced_bit_depth_8_or_less(c.ihdr, c.data)
        assert r is not None and r["ihdr"][3] == depth, name
        back = oracle.expanded_bit_depth_to_8(r["ihdr"], r["data"])
        assert np.array_equal(back["data"], c.data)
    c = _case(golden, "rgba_8_should_be_palette_8")
    r = oracle.reduced_to_indexed(c.ihdr, c.data, True)
    assert r is not None and r["ihdr"][2] == 3 and len(r["palette"]) <= 256
    px = c.data.reshape(-1, 4)
    assert np.array_equal(r["palette"][r["data"]], px)          # lossless
    first = [int(np.argmax(r["data"] == i)) for i in range(len(r["palette"]))]
    assert first == sorted(first)                               # first-occurrence order (IndexSet)
    c = _case(golden, "filter_5_for_rgba_8")
    assert oracle.reduced_to_indexed(c.ihdr, c.data, True) is None  # > 256 colours
    # palette lengths asserted in tests/reduction.rs:996-1167
    for name, n_in, n_out in [("palette_should_be_reduced_with_dupes", 43, 35),
                              ("palette_should_be_reduced_with_unused", 35, 33),
                              ("palette_should_be_reduced_with_both", 43, 33)]:
        c = _case(golden, name)
        assert len(c.palette) == n_in
        r = oracle.reduced_palette(c.ihdr, c.data, c.palette, False)
        assert r is not None and len(r["palette"]) == n_out, name
    c = _case(golden, "palette_should_be_reduced_with_missing")
    r = oracle.reduced_palette(c.ihdr, c.data, c.palette, False)
    assert len(c.palette) == 2 and r is not None and len(r["palette"]) == 3


def test_scale_16_to_8_is_exact_integer_rounding():
    # bit_depth.rs:51-52 f32 round == (510*v + 65535) // 131070 for every 16-bit value
    v = np.arange(65536, dtype=np.uint32)
    data = np.empty(65536 * 2, np.uint8)
    data[0::2] = v >> 8
    data[1::2] = v & 0xFF
    r = oracle.reduced_bit_depth_16_to_8((256, 256, 0, 16, 0), data, True)
    want = np.where((v >> 8) == (v & 0xFF), v >> 8, (510 * v.astype(np.uint64) + 65535) // 131070)
    assert np.array_equal(r["data"], want.astype(np.uint8))


def test_golden_reduction_outcomes_hold(golden):
    import sys
    sys.path.insert(0, helpers.GOLDEN)
    import make_golden
    from pngload import LoadedPng
    for c in golden:
        got = make_golden.reduction_outcomes(LoadedPng(c.ihdr, c.data, c.palette, c.trns, None))
        assert got == c.expected["reductions"], c.name
