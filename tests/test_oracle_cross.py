"""Two independently written restatements of the reference (oracle/oxi_oracle.c and oracle/oxi_oracle_py.py) must
agree on every small case: filtered stream, per-row choices, and reduction outcomes (CPU only)."""
import numpy as np
import pytest

import helpers
import oracle
from oracle import oxi_oracle_py as py

SHAPES = [(1, 1), (2, 3), (5, 4), (9, 7), (16, 5), (23, 9)]
FORMATS = [(0, 1), (0, 2), (3, 4), (0, 8), (3, 8), (4, 8), (2, 8), (6, 8), (0, 16), (4, 16), (2, 16), (6, 16)]


@pytest.mark.parametrize("ct,bd", FORMATS)
def test_filter_image_agrees(ct, bd):
    for il in (False, True):
        for (w, h) in SHAPES:
            kinds = ["photo", "random", "zeros"] + (["transparent"] if ct in (4, 6) else [])
            for kind in kinds:
                ih, d = helpers.synth_image(w, h, ct, bd, 17 * w + h, kind, il)
                data = bytes(d)
                for oa in ((False, True) if ct in (4, 6) else (False,)):
                    for sname, k, basic in helpers.STRATEGIES:
                        a_out, a_used, a_pre = oracle.filter_image(ih, d, k, basic, None, oa)
                        b_out, b_used, b_pre = py.filter_image(ih, data, k, basic, None, oa)
                        assert bytes(a_out) == b_out, (ih, kind, sname, oa)
                        assert a_used.tolist() == b_used and a_pre == b_pre, (ih, kind, sname, oa)
                    pre = [(3 * r + 1) % 5 for r in range(max(1, len(a_used) - 1))]
                    a_out, a_used, _ = oracle.filter_image(ih, d, oracle.PREDEFINED, 0, np.array(pre, np.uint8), oa)
                    b_out, b_used, _ = py.filter_image(ih, data, 5, 0, pre, oa)
                    assert bytes(a_out) == b_out and a_used.tolist() == b_used


def test_scan_lines_agree():
    for ct, bd in [(0, 1), (3, 2), (0, 8), (2, 8), (6, 16)]:
        for il in (0, 1):
            for w in range(1, 20):
                for h in range(1, 20):
                    ih = (w, h, ct, bd, il)
                    n = oracle.raw_data_size(ih)
                    ln, ps, px = oracle.scan_lines(ih, n, True)
                    got = py.scan_line_ranges(ih, n, True)
                    assert [(int(a), (int(b) or None), int(c)) for a, b, c in zip(ln, ps, px)] == got, ih


def test_reductions_agree():
    rng = np.random.default_rng(3)
    for ct, bd in [(4, 8), (6, 8), (4, 16), (6, 16)]:
        for kind in ("photo", "transparent", "zeros"):
            ih, d = helpers.synth_image(11, 6, ct, bd, 5, kind)
            r = oracle.cleaned_alpha_channel(ih, d)
            assert bytes(r["data"]) == py.cleaned_alpha_channel(ih, bytes(d))
            for oa in (False, True):
                a = oracle.reduced_alpha_channel(ih, d, oa)
                b = py.reduced_alpha_channel(ih, bytes(d), oa)
                assert (a is None) == (b is None)
                if a is not None:
                    assert tuple(a["ihdr"]) == b[0] and bytes(a["data"]) == b[1]
                    at = a["trns"] if not isinstance(a["trns"], tuple) else a["trns"][0]
                    assert at == b[2]
    for ct in (0, 2, 4, 6):
        ih, d = helpers.synth_image(9, 5, ct, 16, 2, "random")
        for fs in (False, True):
            a = oracle.reduced_bit_depth_16_to_8(ih, d, fs)
            b = py.reduced_bit_depth_16_to_8(ih, bytes(d), fs)
            assert (a is None) == (b is None) and (a is None or bytes(a["data"]) == b)
        dd = np.repeat(d[0::2], 2)
        assert bytes(oracle.reduced_bit_depth_16_to_8(ih, dd, False)["data"]) == py.reduced_bit_depth_16_to_8(ih, bytes(dd), False)
    for ct, bd in [(2, 8), (6, 8), (2, 16), (6, 16)]:
        ih, d = helpers.synth_image(7, 4, ct, bd, 9, "random")
        assert oracle.reduced_rgb_to_grayscale(ih, d) is None and py.reduced_rgb_to_grayscale(ih, bytes(d)) is None
        bdp = bd // 8
        px = d.reshape(-1, (3 if ct == 2 else 4), bdp).copy()
        px[:, 1] = px[:, 0]
        px[:, 2] = px[:, 0]
        g = px.reshape(-1)
        assert bytes(oracle.reduced_rgb_to_grayscale(ih, g)["data"]) == py.reduced_rgb_to_grayscale(ih, bytes(g))
    for ct in (0, 2, 4, 6):
        for ncol in (3, 40, 300):
            ch = py.CHANNELS[ct]
            pal = rng.integers(0, 256, size=(ncol, ch), dtype=np.uint8)
            d = pal[rng.integers(0, ncol, size=700)].reshape(-1)
            ih = (35, 20, ct, 8, 0)
            a = oracle.reduced_to_indexed(ih, d, True)
            b = py.reduced_to_indexed(ih, bytes(d), True)
            assert (a is None) == (b is None)
            if a is not None:
                assert bytes(a["data"]) == b[0]
                want = [tuple(int(v) for v in c) for c in b[1]]
                got = a["palette"].tolist()
                for g_, w_ in zip(got, want):
                    exp = ((w_[0],) * 3 + (255,) if ch == 1 else (w_[0],) * 3 + (w_[1],) if ch == 2
                           else tuple(w_) + (255,) if ch == 3 else tuple(w_))
                    assert tuple(g_) == exp
    for vals, want in [([0, 255], 1), ([0x55, 0xAA, 0], 2), ([0x11, 0xEE], 4), ([0x12], None), ([0x55, 0x33], 4)]:
        d = np.array(vals * 5, np.uint8)
        a = oracle.reduced_bit_depth_8_or_less((len(d), 1, 0, 8, 0), d)
        assert py.reduced_bit_depth_8_or_less_gray_bits(bytes(d)) == want
        assert (a is None) == (want is None) and (a is None or a["ihdr"][3] == want)
    for il in (0, 1):
        ih = (13, 9, 3, 8, il)
        n = oracle.raw_data_size(ih) - len(oracle.scan_lines(ih, oracle.raw_data_size(ih), True)[0])
        d = rng.integers(0, 7, size=n, dtype=np.uint8)
        assert oracle.cooccurrence_build(ih, 7, d).reshape(-1).tolist() == py.cooccurrence(ih, 7, bytes(d))
