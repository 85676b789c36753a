"""Device-resident PngImage: the ctypes mirror of the oxi_image_* handle API.

A `DeviceImage` plays the role of the reference's `PngImage` while the pixels stay in HBM: the reduction methods have the
reference's names and `Option` convention (None = would not reduce) and return new DeviceImages on the same stream;
`filter_image_multi` is the evaluator's candidate fan-out (evaluate.rs:141-153) and returns only the filtered streams.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from ._lib import CAux, CIhdr, CStrategy, OXI_NONE, check, lib
from .png import FilterStrategy, IhdrData, PngImage
from .reduction import _aux_of, _color_type
from .png import BitDepth


def _setup(L):
    if getattr(L, "_resident_ready", False):
        return L
    H = C.c_void_p
    L.oxi_image_upload.argtypes = [C.c_int, C.c_void_p, C.POINTER(CIhdr), C.POINTER(CAux), C.c_void_p, C.c_size_t, C.POINTER(H)]
    L.oxi_image_wrap_device.argtypes = L.oxi_image_upload.argtypes
    L.oxi_image_from_filtered.argtypes = L.oxi_image_upload.argtypes
    L.oxi_image_free.argtypes = [H]
    L.oxi_image_free.restype = None
    L.oxi_image_info.argtypes = [H, C.POINTER(CIhdr), C.POINTER(CAux), C.POINTER(C.c_size_t)]
    L.oxi_image_device_ptr.argtypes = [H]
    L.oxi_image_device_ptr.restype = C.c_void_p
    L.oxi_image_download.argtypes = [H, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.oxi_image_sync.argtypes = [H]
    for name in ("cleaned_alpha_channel", "reduced_bit_depth_8_or_less", "expanded_bit_depth_to_8", "reduced_rgb_to_grayscale",
                 "sorted_palette", "interlace_image", "deinterlace_image"):
        getattr(L, f"oxi_{name}_device").argtypes = [H, C.POINTER(H)]
    for name in ("reduced_alpha_channel", "reduced_bit_depth_16_to_8", "reduced_to_indexed", "reduced_palette"):
        getattr(L, f"oxi_{name}_device").argtypes = [H, C.c_int, C.POINTER(H)]
    L.oxi_apply_palette_reorder_device.argtypes = [H, C.c_void_p, C.POINTER(H)]
    L.oxi_cooccurrence_build_device.argtypes = [H, C.c_uint32, C.c_void_p]
    L.oxi_used_histogram256_device.argtypes = [H, C.c_void_p]
    L.oxi_image_filter.argtypes = [H, C.POINTER(CStrategy), C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                   C.c_uint]
    L.oxi_image_filter_device.argtypes = L.oxi_image_filter.argtypes
    L._resident_ready = True
    return L


class DeviceImage:
    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)
        self._lib = _setup(lib())

    def __del__(self):
        try:
            if self._h:
                self._lib.oxi_image_free(self._h)
                self._h = None
        except Exception:
            pass

    # ---- construction / access -------------------------------------------------------------------------------------
    @staticmethod
    def upload(png: PngImage, device: int = 0, stream: int = 0) -> "DeviceImage":
        L = _setup(lib())
        h, a, out = png.ihdr._c(), _aux_of(png.ihdr.color_type), C.c_void_p()
        d = np.ascontiguousarray(png.data, dtype=np.uint8)
        check(L.oxi_image_upload(device, C.c_void_p(stream), C.byref(h), C.byref(a), C.c_void_p(d.ctypes.data), d.size,
                                 C.byref(out)), "oxi_image_upload")
        return DeviceImage(out.value)

    @staticmethod
    def from_filtered(ihdr: IhdrData, filtered: np.ndarray, device: int = 0, stream: int = 0) -> "DeviceImage":
        """PngImage::new's unfilter step (png/mod.rs:335-351) on the device."""
        L = _setup(lib())
        h, a, out = ihdr._c(), _aux_of(ihdr.color_type), C.c_void_p()
        d = np.ascontiguousarray(filtered, dtype=np.uint8)
        check(L.oxi_image_from_filtered(device, C.c_void_p(stream), C.byref(h), C.byref(a), C.c_void_p(d.ctypes.data), d.size,
                                        C.byref(out)), "oxi_image_from_filtered")
        return DeviceImage(out.value)

    def info(self):
        h, a, n = CIhdr(), CAux(), C.c_size_t()
        check(self._lib.oxi_image_info(self._h, C.byref(h), C.byref(a), C.byref(n)), "oxi_image_info")
        return h, a, n.value

    @property
    def ihdr(self) -> IhdrData:
        h, a, _ = self.info()
        return IhdrData(h.width, h.height, _color_type(h.color_type, a), BitDepth(h.bit_depth), bool(h.interlaced))

    @property
    def data_len(self) -> int:
        return self.info()[2]

    def device_ptr(self) -> int:
        return self._lib.oxi_image_device_ptr(self._h) or 0

    def download(self) -> PngImage:
        ih, n = self.ihdr, self.data_len
        out = np.empty(n, np.uint8)
        got = C.c_size_t()
        check(self._lib.oxi_image_download(self._h, C.c_void_p(out.ctypes.data), n, C.byref(got)), "oxi_image_download")
        return PngImage(ih, out[:got.value])

    def sync(self):
        check(self._lib.oxi_image_sync(self._h), "oxi_image_sync")

    # ---- reductions (names and Option convention of src/reduction/*) ---------------------------------------------------
    def _derive(self, name: str, *args) -> Optional["DeviceImage"]:
        out = C.c_void_p()
        rc = check(getattr(self._lib, name)(self._h, *args, C.byref(out)), name)
        return None if rc == OXI_NONE else DeviceImage(out.value)

    def cleaned_alpha_channel(self):
        return self._derive("oxi_cleaned_alpha_channel_device")

    def reduced_alpha_channel(self, optimize_alpha: bool):
        return self._derive("oxi_reduced_alpha_channel_device", int(optimize_alpha))

    def reduced_bit_depth_16_to_8(self, force_scale: bool):
        return self._derive("oxi_reduced_bit_depth_16_to_8_device", int(force_scale))

    def reduced_bit_depth_8_or_less(self):
        return self._derive("oxi_reduced_bit_depth_8_or_less_device")

    def expanded_bit_depth_to_8(self):
        return self._derive("oxi_expanded_bit_depth_to_8_device")

    def reduced_to_indexed(self, allow_grayscale: bool):
        return self._derive("oxi_reduced_to_indexed_device", int(allow_grayscale))

    def reduced_rgb_to_grayscale(self):
        return self._derive("oxi_reduced_rgb_to_grayscale_device")

    def reduced_palette(self, optimize_alpha: bool):
        return self._derive("oxi_reduced_palette_device", int(optimize_alpha))

    def sorted_palette(self):
        return self._derive("oxi_sorted_palette_device")

    def apply_palette_reorder(self, remapping):
        rm = np.ascontiguousarray(remapping, dtype=np.uint32)
        return self._derive("oxi_apply_palette_reorder_device", C.c_void_p(rm.ctypes.data))

    def interlace_image(self):
        return self._derive("oxi_interlace_image_device")

    def deinterlace_image(self):
        return self._derive("oxi_deinterlace_image_device")

    def cooccurrence_matrix(self, num_colors: int) -> np.ndarray:
        m = np.zeros((num_colors, num_colors), np.uint32)
        check(self._lib.oxi_cooccurrence_build_device(self._h, num_colors, C.c_void_p(m.ctypes.data)),
              "oxi_cooccurrence_build_device")
        return m

    def used_histogram256(self) -> np.ndarray:
        hist = np.zeros(256, np.uint64)
        check(self._lib.oxi_used_histogram256_device(self._h, C.c_void_p(hist.ctypes.data)), "oxi_used_histogram256_device")
        return hist

    # ---- the evaluator's candidate: k strategies, only the filtered streams come back --------------------------------
    def filter_image_multi(self, strategies: Sequence[FilterStrategy], optimize_alpha: bool, flags: int = 0):
        h, _, n = self.info()
        rows = int(self._lib.oxi_num_scanlines(C.byref(h), n))
        k = len(strategies)
        cs = (CStrategy * k)()
        keep = []
        for j, s in enumerate(strategies):
            cs[j], ka = s._c()
            keep.append(ka)
        outs = [np.empty(n + rows, np.uint8) for _ in range(k)]
        used = [np.empty(rows, np.uint8) for _ in range(k)]
        po = (C.c_void_p * k)(*[o.ctypes.data for o in outs])
        pu = (C.c_void_p * k)(*[u.ctypes.data for u in used])
        check(self._lib.oxi_image_filter(self._h, cs, k, int(optimize_alpha), po, pu, flags), "oxi_image_filter")
        return outs, used


class _LazyData:
    """`PngImage.data` of a resident image: the size is known, the bytes come to the host only if somebody looks."""

    def __init__(self, owner: "ResidentPng"):
        self._o = owner

    @property
    def size(self) -> int:
        return self._o.dev.data_len

    def __array__(self, dtype=None, copy=None):
        a = self._o.host().data
        return a if dtype is None else a.astype(dtype)

    def __eq__(self, other):
        return np.asarray(self) == np.asarray(other)

    def __ne__(self, other):
        return np.asarray(self) != np.asarray(other)

    def tobytes(self):
        return np.asarray(self).tobytes()


class ResidentPng:
    """A DeviceImage behind the attribute names the reduction flow reads from a PngImage (`ihdr`, `data`)."""

    def __init__(self, dev: DeviceImage):
        self.dev = dev
        self._ihdr = None
        self._host = None
        self.data = _LazyData(self)

    @property
    def ihdr(self) -> IhdrData:
        if self._ihdr is None:
            self._ihdr = self.dev.ihdr
        return self._ihdr

    def host(self) -> PngImage:
        if self._host is None:
            self._host = self.dev.download()
        return self._host

    def filter_image_multi(self, strategies, optimize_alpha, flags: int = 0):
        return self.dev.filter_image_multi(strategies, optimize_alpha, flags)


def _wrap(d: Optional[DeviceImage]) -> Optional[ResidentPng]:
    return None if d is None else ResidentPng(d)


class ResidentBackend:
    """The functions of `oxipng_b200.reduction` over ResidentPng: what a caller of perform_reductions
    (reduction/mod.rs:14-179) binds to when the image was uploaded once with DeviceImage.upload."""

    @staticmethod
    def upload(png: PngImage, device: int = 0, stream: int = 0) -> ResidentPng:
        return ResidentPng(DeviceImage.upload(png, device, stream))

    def cleaned_alpha_channel(self, p): return _wrap(p.dev.cleaned_alpha_channel())
    def reduced_alpha_channel(self, p, oa): return _wrap(p.dev.reduced_alpha_channel(oa))
    def reduced_bit_depth_16_to_8(self, p, fs): return _wrap(p.dev.reduced_bit_depth_16_to_8(fs))
    def reduced_bit_depth_8_or_less(self, p): return _wrap(p.dev.reduced_bit_depth_8_or_less())
    def expanded_bit_depth_to_8(self, p): return _wrap(p.dev.expanded_bit_depth_to_8())
    def reduced_to_indexed(self, p, ag): return _wrap(p.dev.reduced_to_indexed(ag))
    def reduced_rgb_to_grayscale(self, p): return _wrap(p.dev.reduced_rgb_to_grayscale())
    def reduced_palette(self, p, oa): return _wrap(p.dev.reduced_palette(oa))
    def sorted_palette(self, p): return _wrap(p.dev.sorted_palette())

    def indexed_to_channels(self, p, ag, oa):  # host, like the non-resident path (<= 20 000 bytes of growth)
        from . import reduction
        return reduction.indexed_to_channels(p.host(), ag, oa)

    @staticmethod
    def change_interlacing(p: ResidentPng, interlace: bool) -> Optional[ResidentPng]:  # png/mod.rs:239-254
        if bool(p.ihdr.interlaced) == bool(interlace):
            return None
        return _wrap(p.dev.interlace_image() if interlace else p.dev.deinterlace_image())
