"""Host-side mirror of the reference's image model and filter API for the hot path.

Same names, argument meaning and return conventions as the Rust items they stand for, so that the parity tests
read like the reference's own tests/benches (`png.raw.filter_image(FilterStrategy::MinSum, false)`,
benches/strategies.rs:12-40):

  RowFilter        src/filters/mod.rs:7-15
  FilterStrategy   src/filters/strategies.rs:9-36   (Brute stays on the reference host code: not offered)
  BitDepth         src/colors.rs:96-107
  ColorType        src/colors.rs:9-90
  IhdrData         src/headers.rs:15-64
  PngImage         src/png/mod.rs:21-26, filter_image :355-431, scan_lines :329-332

All compute goes through the C ABI (include/oxipng_b200.h) to the CUDA kernels; nothing here filters on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from enum import IntEnum
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import CIhdr, CStrategy, check, lib


class RowFilter(IntEnum):
    None_ = 0
    Sub = 1
    Up = 2
    Average = 3
    Paeth = 4

    @classmethod
    def try_from(cls, v: int) -> "RowFilter":  # TryFrom<u8>, src/filters/mod.rs:17-26
        if v > 4 or v < 0:
            raise ValueError("invalid filter id")
        return cls(v)


RowFilter.ALL = [RowFilter.None_, RowFilter.Sub, RowFilter.Up, RowFilter.Average, RowFilter.Paeth]

_KINDS = {"Basic": 0, "MinSum": 1, "Entropy": 2, "Bigrams": 3, "BigEnt": 4, "Predefined": 5}


@dataclass(frozen=True)
class FilterStrategy:
    kind: str
    filter: Optional[RowFilter] = None
    lines: Optional[tuple] = None

    @staticmethod
    def Basic(f) -> "FilterStrategy":
        return FilterStrategy("Basic", RowFilter(int(f)))

    @staticmethod
    def Predefined(lines: Sequence[int]) -> "FilterStrategy":
        return FilterStrategy("Predefined", None, tuple(int(x) for x in lines))

    @staticmethod
    def Brute(num_lines: int, level: int) -> "FilterStrategy":
        raise NotImplementedError("Brute deflates on the host in the reference (strategies.rs:231-266); out of scope")

    def __str__(self) -> str:  # Display, strategies.rs:53-65
        return self.filter.name.rstrip("_") if self.kind == "Basic" else self.kind

    def _c(self):
        """-> (CStrategy, keepalive)"""
        keep = None
        s = CStrategy(_KINDS[self.kind], int(self.filter) if self.filter is not None else 0, None, 0)
        if self.kind == "Predefined":
            keep = np.asarray(self.lines, dtype=np.uint8)
            if keep.size:
                s.predefined = keep.ctypes.data
                s.n_predefined = keep.size
        return s, keep


FilterStrategy.NONE = FilterStrategy.Basic(RowFilter.None_)
FilterStrategy.SUB = FilterStrategy.Basic(RowFilter.Sub)
FilterStrategy.UP = FilterStrategy.Basic(RowFilter.Up)
FilterStrategy.AVERAGE = FilterStrategy.Basic(RowFilter.Average)
FilterStrategy.PAETH = FilterStrategy.Basic(RowFilter.Paeth)
FilterStrategy.MinSum = FilterStrategy("MinSum")
FilterStrategy.Entropy = FilterStrategy("Entropy")
FilterStrategy.Bigrams = FilterStrategy("Bigrams")
FilterStrategy.BigEnt = FilterStrategy("BigEnt")


class BitDepth(IntEnum):
    One = 1
    Two = 2
    Four = 4
    Eight = 8
    Sixteen = 16


@dataclass(frozen=True)
class ColorType:
    """code: PNG colour-type code (png_header_code). palette: (n,4) RGBA8 for Indexed.
    transparent: Grayscale.transparent_shade (int) / RGB.transparent_color ((r,g,b) of u16)."""
    code: int
    palette: Optional[tuple] = None
    transparent: object = None

    @staticmethod
    def Grayscale(transparent_shade: Optional[int] = None) -> "ColorType":
        return ColorType(0, None, transparent_shade)

    @staticmethod
    def RGB(transparent_color=None) -> "ColorType":
        return ColorType(2, None, tuple(transparent_color) if transparent_color is not None else None)

    @staticmethod
    def Indexed(palette) -> "ColorType":
        pal = np.asarray(palette, dtype=np.uint8).reshape(-1, 4)
        return ColorType(3, tuple(map(tuple, pal.tolist())), None)

    @staticmethod
    def GrayscaleAlpha() -> "ColorType":
        return ColorType(4)

    @staticmethod
    def RGBA() -> "ColorType":
        return ColorType(6)

    def png_header_code(self) -> int:
        return self.code

    def channels_per_pixel(self) -> int:
        return {0: 1, 3: 1, 4: 2, 2: 3, 6: 4}[self.code]

    def is_rgb(self) -> bool:
        return self.code in (2, 6)

    def is_gray(self) -> bool:
        return self.code in (0, 4)

    def has_alpha(self) -> bool:
        return self.code in (4, 6)

    def has_trns(self) -> bool:
        return self.code in (0, 2) and self.transparent is not None

    def palette_array(self) -> np.ndarray:
        return np.asarray(self.palette if self.palette else [], dtype=np.uint8).reshape(-1, 4)


@dataclass(frozen=True)
class IhdrData:
    width: int
    height: int
    color_type: ColorType
    bit_depth: BitDepth
    interlaced: bool = False

    def _c(self) -> CIhdr:
        return CIhdr(self.width, self.height, self.color_type.code, int(self.bit_depth), 1 if self.interlaced else 0)

    def bpp(self) -> int:  # headers.rs:32
        return int(self.bit_depth) * self.color_type.channels_per_pixel()

    def raw_data_size(self) -> int:  # headers.rs:38-64
        h = self._c()
        return lib().oxi_raw_data_size(C.byref(h))


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(a, dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


@dataclass
class PngImage:
    ihdr: IhdrData
    data: np.ndarray = field(repr=False)

    def __post_init__(self):
        self.data = _u8(self.data)

    def channels_per_pixel(self) -> int:
        return self.ihdr.color_type.channels_per_pixel()

    def bytes_per_channel(self) -> int:  # png/mod.rs:296-302
        return 2 if self.ihdr.bit_depth == BitDepth.Sixteen else 1

    def scan_lines(self, has_filter: bool = False):
        """[(len, pass or None, num_pixels)] of ScanLines (png/scan_lines.rs); has_filter=False only."""
        if has_filter:
            raise NotImplementedError("only the unfiltered layout is on this path")
        h = self.ihdr._c()
        n = lib().oxi_num_scanlines(C.byref(h), self.data.size)
        ln = np.zeros(n, dtype=np.uint32)
        ps = np.zeros(n, dtype=np.uint8)
        px = np.zeros(n, dtype=np.uint32)
        lib().oxi_scan_lines(C.byref(h), self.data.size, ln.ctypes.data, ps.ctypes.data, px.ctypes.data, n)
        return [(int(a), (int(b) if b else None), int(c)) for a, b, c in zip(ln, ps, px)]

    def filter_image(self, strategy: FilterStrategy, optimize_alpha: bool, flags: int = 0):
        """-> (filtered bytes, FilterStrategy): Predefined(filters_used) for heuristics, else `strategy`."""
        outs, used = filter_image_multi(self, [strategy], optimize_alpha, flags)
        if strategy.kind in ("Basic", "Predefined") or used[0].size == 0:  # png/mod.rs:426-430
            return outs[0], strategy
        return outs[0], FilterStrategy.Predefined(used[0].tolist())


def filter_image_multi(png: PngImage, strategies: Sequence[FilterStrategy], optimize_alpha: bool, flags: int = 0):
    """The evaluator fan-out (evaluate.rs:141-153) as one call: -> ([filtered...], [filters_used...])."""
    outs, used = filter_batch(png.ihdr, png.data, 1, strategies, optimize_alpha, flags)
    return outs, used


def filter_batch(ihdr: IhdrData, data, n_images: int, strategies: Sequence[FilterStrategy], optimize_alpha: bool,
                 flags: int = 0, outs=None, used=None):
    """n_images images of the same IHDR, contiguous in `data`; host buffers in and out."""
    L = lib()
    d = _u8(data)
    h = ihdr._c()
    per = d.size // max(n_images, 1)
    rows = L.oxi_num_scanlines(C.byref(h), per)
    k = len(strategies)
    cs = (CStrategy * k)()
    keep = []
    for j, s in enumerate(strategies):
        cs[j], ka = s._c()
        keep.append(ka)
    if outs is None:
        outs = [np.empty(n_images * (per + rows), dtype=np.uint8) for _ in range(k)]
    if used is None:
        used = [np.empty(n_images * rows, dtype=np.uint8) for _ in range(k)]
    po = (C.c_void_p * k)(*[o.ctypes.data for o in outs])
    pu = (C.c_void_p * k)(*[u.ctypes.data for u in used])
    check(L.oxi_filter_batch(C.byref(h), d.ctypes.data, per, n_images, cs, k, int(optimize_alpha), po, pu, flags),
          "oxi_filter_batch")
    # the reference emits data_len consumed + rows; trailing unconsumed bytes never appear
    consumed = sum(l for l, _, _ in PngImage(ihdr, d[:per]).scan_lines()) if rows else 0
    if consumed != per:
        outs = [o[:n_images * (consumed + rows)] for o in outs]
    return outs, used


def filter_batch_device(device: int, stream: int, ihdr: IhdrData, d_data: int, data_len: int, n_images: int,
                        strategies: Sequence[FilterStrategy], optimize_alpha: bool, d_outs: Sequence[int],
                        d_used: Sequence[int], flags: int = 0) -> None:
    """Device-resident variant: raw device pointers (e.g. torch tensor .data_ptr()) and a cudaStream_t handle.
    Asynchronous with respect to the host."""
    L = lib()
    h = ihdr._c()
    k = len(strategies)
    cs = (CStrategy * k)()
    keep = []
    for j, s in enumerate(strategies):
        cs[j], ka = s._c()
        keep.append(ka)
    po = (C.c_void_p * k)(*d_outs)
    pu = (C.c_void_p * k)(*d_used)
    check(L.oxi_filter_batch_device(device, C.c_void_p(stream), C.byref(h), C.c_void_p(d_data), data_len, n_images, cs,
                                    k, int(optimize_alpha), po, pu, flags), "oxi_filter_batch_device")


def filter_tier(png: PngImage, strategy: FilterStrategy, optimize_alpha: bool) -> str:
    h = png.ihdr._c()
    s, _keep = strategy._c()
    return lib().oxi_filter_tier(C.byref(h), png.data.size, C.byref(s), int(optimize_alpha)).decode()


# ---- the steps either side of the hot path -----------------------------------------------------------

def unfilter_image(ihdr: IhdrData, filtered) -> np.ndarray:
    """PngImage::unfilter_image (png/mod.rs:335-351): inflated IDAT stream -> PngImage.data.  Raises ValueError for an
    invalid filter-type byte (the reference's PngError::InvalidData)."""
    f = _u8(filtered)
    h = ihdr._c()
    out = np.empty(f.size + 16, dtype=np.uint8)
    n = C.c_size_t(0)
    rc = lib().oxi_unfilter_image(C.byref(h), C.c_void_p(f.ctypes.data), C.c_size_t(f.size), C.c_void_p(out.ctypes.data),
                                  C.byref(n))
    if rc == -5:
        raise ValueError("invalid filter type byte")
    check(rc, "oxi_unfilter_image")
    return out[:n.value]


def unfilter_image_device(ihdr: IhdrData, d_filtered: int, length: int, d_out: int, device: int = 0, stream: int = 0) -> int:
    """oxi_unfilter_image_device: the inflated stream and the result are device pointers (e.g. torch tensors'
    data_ptr()); returns the number of bytes written to d_out."""
    h = ihdr._c()
    n = C.c_size_t(0)
    L = lib()
    L.oxi_unfilter_image_device.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    rc = L.oxi_unfilter_image_device(device, C.c_void_p(stream), C.byref(h), C.c_void_p(d_filtered), C.c_size_t(length),
                                     C.c_void_p(d_out), C.byref(n))
    if rc == -5:
        raise ValueError("invalid filter type byte")
    check(rc, "oxi_unfilter_image_device")
    return n.value


def unfilter_batch(ihdr: IhdrData, filtered) -> np.ndarray:
    """oxi_unfilter_batch: filtered = (n_images, len) array of inflated streams of identical IHDR -> (n_images, data_len)."""
    f = np.ascontiguousarray(filtered, dtype=np.uint8)
    assert f.ndim == 2
    n_images, length = f.shape
    h = ihdr._c()
    out = np.empty(f.size + 16, dtype=np.uint8)
    n = C.c_size_t(0)
    L = lib()
    L.oxi_unfilter_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
    rc = L.oxi_unfilter_batch(C.byref(h), C.c_void_p(f.ctypes.data), length, n_images, C.c_void_p(out.ctypes.data), C.byref(n))
    if rc == -5:
        raise ValueError("invalid filter type byte")
    check(rc, "oxi_unfilter_batch")
    return out[:n.value * n_images].reshape(n_images, n.value)


def unfilter_batch_device(ihdr: IhdrData, d_filtered: int, length: int, n_images: int, d_out: int, device: int = 0,
                          stream: int = 0) -> int:
    """oxi_unfilter_batch_device: device pointers in and out; returns the bytes per image written to d_out."""
    h = ihdr._c()
    n = C.c_size_t(0)
    L = lib()
    L.oxi_unfilter_batch_device.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                            C.c_void_p]
    rc = L.oxi_unfilter_batch_device(device, C.c_void_p(stream), C.byref(h), C.c_void_p(d_filtered), length, n_images,
                                     C.c_void_p(d_out), C.byref(n))
    if rc == -5:
        raise ValueError("invalid filter type byte")
    check(rc, "oxi_unfilter_batch_device")
    return n.value


def interlace_image(png: PngImage) -> PngImage:
    """interlace_image (interlace.rs:6-60): progressive -> Adam7 layout."""
    h = png.ihdr._c()
    out = np.empty(png.data.size + 8 * png.ihdr.height + 64, dtype=np.uint8)
    n = C.c_size_t(0)
    check(lib().oxi_interlace_image(C.byref(h), C.c_void_p(png.data.ctypes.data), C.c_size_t(png.data.size),
                                    C.c_void_p(out.ctypes.data), C.byref(n)), "oxi_interlace_image")
    i = png.ihdr
    return PngImage(IhdrData(i.width, i.height, i.color_type, i.bit_depth, True), out[:n.value])


def deinterlace_image(png: PngImage) -> PngImage:
    """deinterlace_image (interlace.rs:62-150): Adam7 -> progressive layout."""
    h = png.ihdr._c()
    i = png.ihdr
    stride = (i.width * i.bpp() + 7) // 8
    out = np.empty(stride * i.height + 16, dtype=np.uint8)
    n = C.c_size_t(0)
    check(lib().oxi_deinterlace_image(C.byref(h), C.c_void_p(png.data.ctypes.data), C.c_size_t(png.data.size),
                                      C.c_void_p(out.ctypes.data), C.byref(n)), "oxi_deinterlace_image")
    return PngImage(IhdrData(i.width, i.height, i.color_type, i.bit_depth, False), out[:n.value])


def change_interlacing(png: PngImage, interlace: bool) -> Optional[PngImage]:
    """PngImage::change_interlacing (png/mod.rs:272-284): None when the image already has that layout."""
    if interlace == png.ihdr.interlaced:
        return None
    return interlace_image(png) if interlace else deinterlace_image(png)
