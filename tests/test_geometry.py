"""Host geometry of the product (C ABI) against the oracle's restatement of the ScanLines state machine."""
import ctypes as C

import numpy as np
import pytest

import oracle
from oxipng_b200 import _lib


def _rows(ih, n):
    L = _lib.lib()
    h = _lib.CIhdr(*ih)
    cnt = L.oxi_num_scanlines(C.byref(h), n)
    ln = np.zeros(cnt, np.uint32)
    ps = np.zeros(cnt, np.uint8)
    px = np.zeros(cnt, np.uint32)
    L.oxi_scan_lines(C.byref(h), n, ln.ctypes.data, ps.ctypes.data, px.ctypes.data, cnt)
    return ln, ps, px


@pytest.mark.parametrize("ct,bd", [(0, 1), (0, 2), (3, 4), (0, 8), (4, 8), (2, 8), (6, 8), (0, 16), (6, 16)])
def test_scanlines_match_oracle_state_machine(ct, bd):
    L = _lib.lib()
    for il in (0, 1):
        for w in range(1, 40):
            for h in list(range(1, 20)) + [31, 32, 33]:
                ih = (w, h, ct, bd, il)
                hh = _lib.CIhdr(*ih)
                raw = oracle.raw_data_size(ih)
                assert L.oxi_raw_data_size(C.byref(hh)) == raw
                o_len, o_pass, o_px = oracle.scan_lines(ih, raw, True)
                n = int(raw - len(o_len))           # unfiltered byte count
                o_len, o_pass, o_px = oracle.scan_lines(ih, n, False)
                ln, ps, px = _rows(ih, n)
                assert np.array_equal(ln, o_len.astype(np.uint32)), ih
                assert np.array_equal(ps, o_pass), ih
                assert np.array_equal(px, o_px.astype(np.uint32)), ih
                assert int(ln.sum()) == n


def test_truncated_and_oversized_data_follow_the_iterator():
    for ih in [(9, 9, 2, 8, 1), (9, 9, 2, 8, 0), (5, 5, 0, 1, 1), (17, 3, 6, 16, 1)]:
        full = oracle.raw_data_size(ih) - len(oracle.scan_lines(ih, oracle.raw_data_size(ih), True)[0])
        for n in [0, 1, full // 3, full - 1, full]:
            o_len, o_pass, _ = oracle.scan_lines(ih, n, False)
            ln, ps, _ = _rows(ih, n)
            assert np.array_equal(ln, o_len.astype(np.uint32)), (ih, n)
            assert np.array_equal(ps, o_pass), (ih, n)
    # a non-interlaced image keeps yielding rows while bytes remain (the iterator never checks height)
    ih = (4, 2, 2, 8, 0)
    assert len(_rows(ih, 12 * 5)[0]) == 5 == len(oracle.scan_lines(ih, 60, False)[0])


def test_filter_bpp():
    L = _lib.lib()
    for ih, want in [((1, 1, 0, 1, 0), 1), ((1, 1, 3, 4, 0), 1), ((1, 1, 0, 16, 0), 2), ((1, 1, 2, 8, 0), 3),
                     ((1, 1, 6, 8, 0), 4), ((1, 1, 2, 16, 0), 6), ((1, 1, 6, 16, 0), 8), ((1, 1, 4, 8, 0), 2)]:
        assert L.oxi_filter_bpp(C.byref(_lib.CIhdr(*ih))) == want == oracle.filter_bpp(ih)
