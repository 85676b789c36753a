"""Minimal PNG container reader for test fixtures (test infrastructure, not product).

Follows the shape of PngData::from_slice / PngImage::new in the reference (src/png/mod.rs:55-159,251-265):
walk the chunks, concatenate IDAT, inflate with zlib, undo the row filters (through the oracle's
restatement of unfilter_image).  Returns the image in the reference's in-memory layout: `data` =
unfiltered scanlines back to back (packed sub-byte samples, big-endian 16-bit, Adam7 passes in order).
"""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass

import numpy as np

import oracle

PNG_SIG = b"\x89PNG\r\n\x1a\n"


@dataclass
class LoadedPng:
    ihdr: tuple  # (width, height, color_type, bit_depth, interlaced)
    data: np.ndarray
    palette: np.ndarray | None  # (n,4) RGBA8
    trns: object  # None | int (gray) | (r,g,b)
    filtered: np.ndarray  # the inflated IDAT stream as stored in the file


def load_png(path_or_bytes) -> LoadedPng:
    raw = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    assert raw[:8] == PNG_SIG, "not a PNG"
    pos = 8
    ihdr = None
    idat = bytearray()
    plte = None
    trns_raw = None
    while pos + 8 <= len(raw):
        (ln,) = struct.unpack(">I", raw[pos:pos + 4])
        name = raw[pos + 4:pos + 8]
        body = raw[pos + 8:pos + 8 + ln]
        pos += 12 + ln
        if name == b"IHDR":
            w, h, bd, ct, _cm, _fm, il = struct.unpack(">IIBBBBB", body)
            ihdr = (w, h, ct, bd, il)
        elif name == b"PLTE":
            plte = np.frombuffer(body, dtype=np.uint8).reshape(-1, 3)
        elif name == b"tRNS":
            trns_raw = body
        elif name == b"IDAT":
            idat += body
        elif name == b"IEND":
            break
    assert ihdr is not None
    # tolerate truncated/odd zlib streams the way the reference's inflate does not need to: fixtures are valid
    filtered = np.frombuffer(zlib.decompress(bytes(idat)), dtype=np.uint8)
    data = oracle.unfilter_image(ihdr, filtered)
    palette = None
    trns = None
    ct = ihdr[2]
    if ct == 3:
        n = 0 if plte is None else len(plte)
        palette = np.full((n, 4), 255, dtype=np.uint8)
        if n:
            palette[:, :3] = plte
        if trns_raw is not None:
            k = min(len(trns_raw), n)
            palette[:k, 3] = np.frombuffer(trns_raw[:k], dtype=np.uint8)
    elif trns_raw is not None:
        if ct == 0 and len(trns_raw) >= 2:
            trns = struct.unpack(">H", trns_raw[:2])[0]
        elif ct == 2 and len(trns_raw) >= 6:
            trns = struct.unpack(">HHH", trns_raw[:6])
    return LoadedPng(ihdr, data, palette, trns, filtered)


def write_png(ihdr, filtered_stream, palette=None, trns=None) -> bytes:
    """Assemble a PNG file around an already-filtered stream (for decode-and-compare checks)."""
    w, h, ct, bd, il = ihdr

    def chunk(name, body):
        c = name + body
        return struct.pack(">I", len(body)) + c + struct.pack(">I", zlib.crc32(c) & 0xFFFFFFFF)

    out = bytearray(PNG_SIG)
    out += chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, bd, ct, 0, 0, 1 if il else 0))
    if ct == 3:
        pal = np.asarray(palette, dtype=np.uint8).reshape(-1, 4)
        out += chunk(b"PLTE", pal[:, :3].tobytes())
        if (pal[:, 3] != 255).any():
            out += chunk(b"tRNS", pal[:, 3].tobytes())
    elif trns is not None:
        if ct == 0:
            out += chunk(b"tRNS", struct.pack(">H", trns))
        elif ct == 2:
            out += chunk(b"tRNS", struct.pack(">HHH", *trns))
    out += chunk(b"IDAT", zlib.compress(bytes(filtered_stream), 1))
    out += chunk(b"IEND", b"")
    return bytes(out)


def to_samples(ihdr, data) -> np.ndarray:
    """Unpack the reference layout into an (H, W, channels) array of samples (uint8 or uint16),
    deinterlacing first when needed.  Independent of the product; used for decode-and-compare."""
    w, h, ct, bd, il = ihdr
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ct]
    d = np.asarray(data, dtype=np.uint8)
    if il:
        d = oracle.deinterlace_image(ihdr, d)
    stride = (w * ch * bd + 7) // 8
    rows = d[:stride * h].reshape(h, stride)
    if bd == 8:
        return rows[:, :w * ch].reshape(h, w, ch)
    if bd == 16:
        v = rows[:, :w * ch * 2].reshape(h, w * ch, 2).astype(np.uint16)
        return ((v[..., 0] << 8) | v[..., 1]).reshape(h, w, ch)
    bits = np.unpackbits(rows, axis=1)[:, :w * bd].reshape(h, w, bd)
    weights = (1 << np.arange(bd - 1, -1, -1)).astype(np.uint8)
    return (bits * weights).sum(axis=2).astype(np.uint8).reshape(h, w, 1)
