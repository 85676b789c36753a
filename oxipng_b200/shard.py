"""Multi-GPU sharding of the filter-search path.  No data-path collective exists or is needed:

* by image: image i of a batch goes to one rank (contiguous, balanced shards);
* by row band, for one very large image: with optimize_alpha off a filtered row depends only on its own bytes and the
  previous *raw* row of the same pass (png/mod.rs:375-378,422), so a band needs a one-row halo and nothing else.  A band
  is handed to the unchanged C ABI as a small non-interlaced image whose first row is the halo; that row's output is
  dropped.  Outputs of the bands are disjoint, contiguous ranges of the filtered stream.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np


def image_shard(n_images: int, world: int, rank: int) -> tuple[int, int]:
    """-> (first image, count) of `rank`'s contiguous shard; counts differ by at most one."""
    base, extra = divmod(n_images, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


@dataclass(frozen=True)
class Band:
    pass_index: int      # index into the image's non-empty passes (0 for non-interlaced)
    width_px: int        # pixels per scanline of that pass
    row_len: int         # bytes per scanline
    first_row: int       # global scanline index of the band's first row
    n_rows: int
    has_halo: bool       # False when the band starts a pass (previous row = zeros)
    in_begin: int        # byte range of PngImage.data to ship, halo row included
    in_end: int
    out_begin: int       # byte range of the filtered stream this band produces
    out_end: int


def row_bands(scan_lines: Sequence[tuple], n_parts: int) -> List[Band]:
    """Split an image (given as PngImage.scan_lines(): [(len, pass, pixels)]) into about n_parts bands of nearly equal
    byte size that never cross a pass."""
    passes = []  # (pass id, len, pixels, first row, nrows, in_base)
    off = 0
    for r, (ln, ps, px) in enumerate(scan_lines):
        if not passes or passes[-1][0] != ps or passes[-1][1] != ln:
            passes.append([ps, ln, px, r, 0, off])
        passes[-1][4] += 1
        off += ln
    total = off
    bands: List[Band] = []
    for pi, (ps, ln, px, r0, nrows, in_base) in enumerate(passes):
        share = max(1, round(n_parts * (nrows * ln) / max(total, 1)))
        share = min(share, nrows)
        for b in range(share):
            k0, k1 = b * nrows // share, (b + 1) * nrows // share
            if k1 <= k0:
                continue
            halo = k0 > 0
            bands.append(Band(pi, px, ln, r0 + k0, k1 - k0, halo,
                              in_base + (k0 - (1 if halo else 0)) * ln, in_base + k1 * ln,
                              in_base + k0 * ln + r0 + k0, in_base + k1 * ln + r0 + k1))
    return bands


def filter_image_sharded(png, strategies, devices: Sequence[int] = (0,), bands_per_device: int = 1, flags: int = 0,
                         outs=None, used=None):
    """oxi_filter_image_sharded: one large image split into row bands (one-row halo), the bands filtered concurrently on
    `devices` -- one worker thread and stream per device inside the C library.  optimize_alpha is off by construction.
    -> ([filtered stream per strategy], [filters_used per strategy])"""
    import ctypes as C
    from ._lib import CStrategy, check, lib
    L = lib()
    L.oxi_filter_image_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(CStrategy), C.c_size_t, C.c_int,
                                           C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_void_p), C.c_uint]
    d = np.ascontiguousarray(png.data, dtype=np.uint8)
    h = png.ihdr._c()
    rows = L.oxi_num_scanlines(C.byref(h), d.size)
    k = len(strategies)
    cs = (CStrategy * k)()
    keep = []
    for j, s in enumerate(strategies):
        cs[j], ka = s._c()
        keep.append(ka)
    if outs is None:
        outs = [np.empty(d.size + rows, dtype=np.uint8) for _ in range(k)]
    if used is None:
        used = [np.empty(rows, dtype=np.uint8) for _ in range(k)]
    po = (C.c_void_p * k)(*[o.ctypes.data for o in outs])
    pu = (C.c_void_p * k)(*[u.ctypes.data for u in used])
    devs = (C.c_int * len(devices))(*devices)
    check(L.oxi_filter_image_sharded(C.byref(h), C.c_void_p(d.ctypes.data), d.size, cs, k, 0, devs, len(devices),
                                     int(bands_per_device), po, pu, flags), "oxi_filter_image_sharded")
    return outs, used


def filter_image_banded(png, strategy, n_parts: int, devices: Sequence[int] = (0,)):
    """filter_image of one large image split into about n_parts row bands over `devices` (concurrently; see
    filter_image_sharded).  -> (filtered stream, filters_used)"""
    per = max(1, -(-n_parts // max(len(devices), 1)))
    outs, used = filter_image_sharded(png, [strategy], devices, per)
    return outs[0], used[0]
