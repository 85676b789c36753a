"""Shared test helpers: golden fixtures, strategy tables, synthetic images, oracle<->product adapters."""
from __future__ import annotations

import hashlib
import json
import os
from dataclasses import dataclass

import numpy as np

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
REF_FILES = "/root/reference/tests/files"

# (name, oracle kind, basic filter)
STRATEGIES = [("None", oracle.BASIC, 0), ("Sub", oracle.BASIC, 1), ("Up", oracle.BASIC, 2),
              ("Average", oracle.BASIC, 3), ("Paeth", oracle.BASIC, 4), ("MinSum", oracle.MINSUM, 0),
              ("Entropy", oracle.ENTROPY, 0), ("Bigrams", oracle.BIGRAMS, 0), ("BigEnt", oracle.BIGENT, 0)]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@dataclass
class Case:
    name: str
    ihdr: tuple
    data: np.ndarray
    palette: object
    trns: object
    expected: dict


def load_golden() -> list[Case]:
    z = np.load(os.path.join(GOLDEN, "fixtures.npz"))
    exp = json.load(open(os.path.join(GOLDEN, "expected.json")))
    cases = []
    for name in z["names"].tolist():
        ih = tuple(int(x) for x in z[name + "/ihdr"])
        pal = z[name + "/palette"] if name + "/palette" in z else None
        trns = None
        if name + "/trns" in z:
            t = z[name + "/trns"].tolist()
            trns = int(t[0]) if len(t) == 1 else tuple(int(x) for x in t)
        cases.append(Case(name, ih, z[name + "/data"], pal, trns, exp[name]))
    return cases


def product_strategy(name: str, used=None):
    import oxipng_b200 as ox
    table = {"None": ox.FilterStrategy.NONE, "Sub": ox.FilterStrategy.SUB, "Up": ox.FilterStrategy.UP,
             "Average": ox.FilterStrategy.AVERAGE, "Paeth": ox.FilterStrategy.PAETH, "MinSum": ox.FilterStrategy.MinSum,
             "Entropy": ox.FilterStrategy.Entropy, "Bigrams": ox.FilterStrategy.Bigrams,
             "BigEnt": ox.FilterStrategy.BigEnt}
    if name == "Predefined":
        return ox.FilterStrategy.Predefined(used)
    return table[name]


def product_image(ihdr: tuple, data, palette=None, trns=None):
    import oxipng_b200 as ox
    w, h, ct, bd, il = ihdr
    if ct == 0:
        c = ox.ColorType.Grayscale(trns)
    elif ct == 2:
        c = ox.ColorType.RGB(trns)
    elif ct == 3:
        c = ox.ColorType.Indexed(palette if palette is not None else np.zeros((0, 4), np.uint8))
    elif ct == 4:
        c = ox.ColorType.GrayscaleAlpha()
    else:
        c = ox.ColorType.RGBA()
    return ox.PngImage(ox.IhdrData(w, h, c, ox.BitDepth(bd), bool(il)), data)


def row_bytes(ihdr: tuple) -> int:
    w, h, ct, bd, il = ihdr
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ct]
    return (w * ch * bd + 7) // 8


def splitmix64(z: np.ndarray) -> np.ndarray:
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synth_image(width: int, height: int, ct: int, bd: int, seed: int, kind: str = "photo", interlaced: bool = False):
    """Deterministic synthetic image in the reference layout (non-interlaced generation, optionally passed through
    the oracle's interlace).  kind: photo (ramp + noise), random, flat, zeros, transparent (alpha images: blocks of
    alpha=0 with random colour)."""
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ct]
    with np.errstate(over="ignore"):
        n = width * height * ch
        k = np.arange(n, dtype=np.uint64)
        r = splitmix64(np.uint64(seed) * np.uint64(0x100000001B3) + k)
    y, x, c = np.meshgrid(np.arange(height), np.arange(width), np.arange(ch), indexing="ij")
    noise = (r % np.uint64(7)).astype(np.int64).reshape(height, width, ch) - 3
    if kind == "photo" or kind == "transparent":
        base = ((3 * x + 5 * y + c * 64) >> 4)
        v = (base + noise) & 0xFF
    elif kind == "random":
        v = (r >> np.uint64(13)).astype(np.int64).reshape(height, width, ch) & 0xFF
    elif kind == "flat":
        v = np.full((height, width, ch), 77, dtype=np.int64) + (y // 64)
        v &= 0xFF
    else:
        v = np.zeros((height, width, ch), dtype=np.int64)
    if bd == 16:
        lo = (r >> np.uint64(20)).astype(np.int64).reshape(height, width, ch) & 0xFF
        samples = (v << 8) | lo
    else:
        samples = v & ((1 << bd) - 1)
    if ct in (4, 6) and kind != "zeros":
        a = np.full((height, width), (1 << bd) - 1, dtype=np.int64)
        if kind == "transparent":
            rr = (r.reshape(height, width, ch)[..., 0] >> np.uint64(40)).astype(np.int64)
            blk = ((x[..., 0] // 5) + (y[..., 0] // 3)) % 3 == 0
            a[blk] = 0
            a[(rr % 11) == 0] = 0
            a[0, :] = 0 if height > 2 else a[0, :]
            if height > 3:
                a[3, : width // 2] = 0
        samples[..., ch - 1] = a
    if bd == 16:
        by = np.empty((height, width * ch, 2), dtype=np.uint8)
        s = samples.reshape(height, width * ch)
        by[..., 0] = s >> 8
        by[..., 1] = s & 0xFF
        rows = by.reshape(height, -1)
    elif bd == 8:
        rows = samples.reshape(height, width * ch).astype(np.uint8)
    else:
        s = samples.reshape(height, width).astype(np.uint8)
        bits = ((s[..., None] >> np.arange(bd - 1, -1, -1)) & 1).astype(np.uint8).reshape(height, width * bd)
        pad = (-bits.shape[1]) % 8
        if pad:
            bits = np.concatenate([bits, np.zeros((height, pad), np.uint8)], axis=1)
        rows = np.packbits(bits, axis=1)
    data = np.ascontiguousarray(rows).reshape(-1)
    ihdr = (width, height, ct, bd, 0)
    if interlaced:
        data = oracle.interlace_image(ihdr, data)
        ihdr = (width, height, ct, bd, 1)
    return ihdr, data


# ---- oracle-backed adapters with the product's names (for end-to-end flow parity tests) ------------------------------

def _img_tuple(png):
    i = png.ihdr
    return (i.width, i.height, i.color_type.code, int(i.bit_depth), int(i.interlaced))


def _from_oracle(r):
    if r is None:
        return None
    return product_image(tuple(r["ihdr"]), r["data"], r.get("palette"), r.get("trns"))


class OracleReductions:
    """oxipng_b200.reduction's interface implemented with the CPU oracle (tests only)."""

    @staticmethod
    def _pt(png):
        ct = png.ihdr.color_type
        pal = ct.palette_array() if ct.code == 3 else None
        return _img_tuple(png), png.data, pal, ct.transparent

    def cleaned_alpha_channel(self, png):
        ih, d, _, _ = self._pt(png)
        r = oracle.cleaned_alpha_channel(ih, d)
        return None if r is None else product_image(ih, r["data"])

    def reduced_alpha_channel(self, png, oa):
        ih, d, _, _ = self._pt(png)
        return _from_oracle(oracle.reduced_alpha_channel(ih, d, oa))

    def reduced_bit_depth_16_to_8(self, png, fs):
        ih, d, pal, t = self._pt(png)
        return _from_oracle(oracle.reduced_bit_depth_16_to_8(ih, d, fs, trns=t))

    def reduced_rgb_to_grayscale(self, png):
        ih, d, _, t = self._pt(png)
        return _from_oracle(oracle.reduced_rgb_to_grayscale(ih, d, t if ih[2] == 2 else None))

    def expanded_bit_depth_to_8(self, png):
        ih, d, pal, t = self._pt(png)
        return _from_oracle(oracle.expanded_bit_depth_to_8(ih, d, pal, t if ih[2] == 0 else None))

    def reduced_palette(self, png, oa):
        ih, d, pal, _ = self._pt(png)
        if ih[2] != 3:
            return None
        return _from_oracle(oracle.reduced_palette(ih, d, pal, oa))

    def sorted_palette(self, png):
        ih, d, pal, _ = self._pt(png)
        if ih[2] != 3:
            return None
        return _from_oracle(oracle.sorted_palette(ih, d, pal))

    def reduced_to_indexed(self, png, ag):
        ih, d, _, t = self._pt(png)
        return _from_oracle(oracle.reduced_to_indexed(ih, d, ag, t))

    def reduced_bit_depth_8_or_less(self, png):
        ih, d, pal, t = self._pt(png)
        return _from_oracle(oracle.reduced_bit_depth_8_or_less(ih, d, pal, t if ih[2] == 0 else None))

    def indexed_to_channels(self, png, ag, oa):
        ih, d, pal, _ = self._pt(png)
        if ih[2] != 3:
            return None
        return _from_oracle(oracle.indexed_to_channels(ih, d, pal, ag, oa))


def oracle_change_interlacing(png, interlace):
    if interlace == png.ihdr.interlaced:
        return None
    ih = _img_tuple(png)
    ct = png.ihdr.color_type
    data = oracle.interlace_image(ih, png.data) if interlace else oracle.deinterlace_image(ih, png.data)
    return product_image((ih[0], ih[1], ih[2], ih[3], int(interlace)), data,
                         ct.palette_array() if ct.code == 3 else None, ct.transparent)


class RecordingEvaluator:
    """Stands in for Evaluator in flow tests: records what perform_reductions submits, in order."""

    def __init__(self):
        self.tried = []

    def try_image(self, png, description=""):
        self.tried.append((description, _img_tuple(png), sha(png.data),
                           sha(png.ihdr.color_type.palette_array()) if png.ihdr.color_type.code == 3 else None,
                           png.ihdr.color_type.transparent))
