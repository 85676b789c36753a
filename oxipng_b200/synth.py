"""Deterministic synthetic images for the benchmark configurations, generated on the device
with torch integer ops from a counter-based PRNG (splitmix64 of seed*0x100000001B3 + byte index), bit-identical to
the numpy generator the tests use (tests/helpers.py::synth_image, kind="photo").

photo: per channel base(x,y,c) = (3x + 5y + 64c) >> 4, plus uniform noise in [-3,3]; alpha (if any) = max.
16-bit samples: hi = photo value, lo = PRNG byte.  Output layout = PngImage.data (rows back to back, big-endian).
"""
from __future__ import annotations

import torch

_M64 = (1 << 64) - 1


def _i64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(z: torch.Tensor, n: int) -> torch.Tensor:
    return (z >> n) & ((1 << (64 - n)) - 1)


def _splitmix64(z: torch.Tensor) -> torch.Tensor:
    z = z + _i64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def _umod7(r: torch.Tensor) -> torch.Tensor:
    hi, lo = _lsr(r, 32), r & 0xFFFFFFFF  # 2^32 mod 7 == 4
    return ((hi % 7) * 4 + (lo % 7)) % 7


def photo_image(width: int, height: int, channels: int, bit_depth: int, seed: int, device) -> torch.Tensor:
    """-> uint8 tensor of height * width * channels * (bit_depth // 8) bytes (bit_depth 8 or 16)."""
    assert bit_depth in (8, 16)
    n = width * height * channels
    k = torch.arange(n, dtype=torch.int64, device=device)
    r = _splitmix64(k + _i64(seed * 0x100000001B3))
    c = k % channels
    x = (k // channels) % width
    y = k // (channels * width)
    v = (((3 * x + 5 * y + 64 * c) >> 4) + _umod7(r) - 3) & 0xFF
    has_alpha = channels in (2, 4)
    if has_alpha:
        v = torch.where(c == channels - 1, torch.full_like(v, 255), v)
    if bit_depth == 8:
        return v.to(torch.uint8)
    lo = _lsr(r, 20) & 0xFF
    if has_alpha:
        lo = torch.where(c == channels - 1, torch.full_like(lo, 255), lo)
    return torch.stack([v, lo], dim=1).reshape(-1).to(torch.uint8)


def photo_batch(n_images: int, width: int, height: int, channels: int, bit_depth: int, seed0: int, device) -> torch.Tensor:
    per = width * height * channels * (bit_depth // 8)
    out = torch.empty(n_images * per, dtype=torch.uint8, device=device)
    for i in range(n_images):
        out[i * per:(i + 1) * per] = photo_image(width, height, channels, bit_depth, seed0 + i, device)
    return out
