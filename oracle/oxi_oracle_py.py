"""Second, independent restatement of the reference path in pure Python (small cases only).

TEST INFRASTRUCTURE ONLY.  Written separately from oracle/oxi_oracle.c, following the Rust control flow line by line
(iterators, Vec pushes, in-place mutation), so that a transcription slip in either restatement shows up as a
disagreement (tests/test_oracle_cross.py).  "parity unpinned" at byte level, like the C oracle (see oxi_oracle.h).
Citations are file:line relative to /root/reference.
"""
from __future__ import annotations

CHANNELS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}  # colors.rs:59-66


def bits_per_pixel(ih):  # headers.rs:32
    return ih[3] * CHANNELS[ih[2]]


def scan_line_ranges(ih, data_len, has_filter):
    """ScanLineRanges::next, png/scan_lines.rs:82-167 -> list of (len, pass or None, pixels)."""
    width, height, _ct, _bd, interlaced = ih
    bpp = bits_per_pixel(ih)
    left = data_len
    pas = [1, 0] if interlaced else None
    out = []
    while True:
        if left == 0:
            break
        if pas is not None:
            if width < 5 and pas[0] == 2:
                pas[0], pas[1] = 3, 4
            if height < 5 and pas[0] == 3:
                pas[0], pas[1] = 4, 0
            if width < 3 and pas[0] == 4:
                pas[0], pas[1] = 5, 2
            if height < 3 and pas[0] == 5:
                pas[0], pas[1] = 6, 0
            if width == 1 and pas[0] == 6:
                pas[0], pas[1] = 7, 1
            pixels_factor, y_steps = {1: (8, 8), 2: (8, 8), 3: (4, 8), 4: (4, 4), 5: (2, 4), 6: (2, 2), 7: (1, 2)}[pas[0]]
            pixels_per_line = width // pixels_factor
            gap = width % pixels_factor
            if pas[0] in (1, 3, 5) and gap > 0:
                pixels_per_line += 1
            elif pas[0] == 2 and gap >= 5:
                pixels_per_line += 1
            elif pas[0] == 4 and gap >= 3:
                pixels_per_line += 1
            elif pas[0] == 6 and gap >= 2:
                pixels_per_line += 1
            current_pass = pas[0]
            if pas[1] + y_steps >= height:
                pas[0] += 1
                pas[1] = {3: 4, 5: 2, 7: 1}.get(pas[0], 0)
            else:
                pas[1] += y_steps
        else:
            pixels_per_line, current_pass = width, None
        ln = -(-(pixels_per_line * bpp) // 8)
        if has_filter:
            ln += 1
        if left < ln:  # checked_sub(len)?
            break
        left -= ln
        out.append((ln, current_pass, pixels_per_line))
    return out


def paeth_predictor(a, b, c):  # filters/mod.rs:242-254
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    if pa <= pb and pa <= pc:
        return a
    if pb <= pc:
        return b
    return c


def optimize_alpha(f, bpp, data, prev_line, color_bytes):  # filters/mod.rs:114-186 (data: bytearray, mutated)
    if f == 0:
        return
    npx = len(data) // bpp

    def px(i):
        return data[i * bpp:(i + 1) * bpp]

    for i in range(npx):
        if all(b == 0 for b in px(i)[color_bytes:]):
            if i == 0:
                prev = next((q for q in range(npx) if any(b != 0 for b in px(q)[color_bytes:])), i)
            else:
                prev = i - 1
            if f == 1:
                if prev != i:
                    data[i * bpp:i * bpp + color_bytes] = data[prev * bpp:prev * bpp + color_bytes]
            elif f == 2:
                data[i * bpp:i * bpp + color_bytes] = prev_line[i * bpp:i * bpp + color_bytes]
            elif f == 3:
                for j in range(color_bytes):
                    if i == 0:
                        data[j] = prev_line[j] >> 1
                    else:
                        data[i * bpp + j] = (data[(i - 1) * bpp + j] + prev_line[i * bpp + j]) >> 1
            elif f == 4:
                for j in range(color_bytes):
                    if i == 0:
                        data[j] = min(data[prev * bpp + j], prev_line[j])
                    else:
                        data[i * bpp + j] = paeth_predictor(data[(i - 1) * bpp + j], prev_line[i * bpp + j],
                                                            prev_line[(i - 1) * bpp + j])


def filter_line(f, bpp, data, prev_line, buf, alpha_bytes):  # filters/mod.rs:46-111 (data mutated, buf extended)
    assert len(data) >= bpp and len(data) == len(prev_line)
    if alpha_bytes != 0:
        optimize_alpha(f, bpp, data, prev_line, bpp - alpha_bytes)
    buf.append(f)
    n = len(data)
    if f == 0:
        buf.extend(data)
    elif f == 1:
        buf.extend(data[0:bpp])
        buf.extend((data[i] - data[i - bpp]) & 255 for i in range(bpp, n))
    elif f == 2:
        buf.extend((data[i] - prev_line[i]) & 255 for i in range(n))
    elif f == 3:
        buf.extend((data[i] - (prev_line[i] >> 1)) & 255 for i in range(min(bpp, n)))
        buf.extend((data[i] - ((data[i - bpp] + prev_line[i]) >> 1)) & 255 for i in range(bpp, n))
    else:
        buf.extend((data[i] - prev_line[i]) & 255 for i in range(min(bpp, n)))
        buf.extend((data[i] - paeth_predictor(data[i - bpp], prev_line[i], prev_line[i - bpp])) & 255
                   for i in range(bpp, n))


def ilog2i(i):  # strategies.rs:269-272
    log = i.bit_length() - 1
    return (i * log + ((i - (1 << log)) << 1)) & 0xFFFFFFFF


def _as_i32(u):
    u &= 0xFFFFFFFF
    return u - (1 << 32) if u >= (1 << 31) else u


class MinSum:  # strategies.rs:76-101
    def reset(self):
        self.best = None

    def evaluate(self, line):
        size = sum((256 - b) if b >= 128 else b for b in line)
        if self.best is None or size < self.best:
            self.best = size
            return True
        return False


class Entropy:  # strategies.rs:105-149
    def reset(self):
        self.best = None

    def evaluate(self, line):
        hist = [0] * 256
        for b in line:
            hist[b] += 1
        size = _as_i32(sum(ilog2i(x) for x in hist if x))
        if self.best is None or size > self.best:
            self.best = size
            return True
        return False


class Bigrams:  # strategies.rs:153-192 (with the early exit)
    def reset(self):
        self.best = None

    def evaluate(self, line):
        seen = set()
        count = 0
        for i in range(len(line) - 1):
            bg = (line[i] << 8) | line[i + 1]
            if bg not in seen:
                count += 1
                if self.best is not None and count >= self.best:
                    return False
                seen.add(bg)
        self.best = count
        return True


class BigEnt:  # strategies.rs:195-227
    def reset(self):
        self.best = None

    def evaluate(self, line):
        counts = {}
        for i in range(len(line) - 1):
            bg = (line[i] << 8) | line[i + 1]
            counts[bg] = counts.get(bg, 0) + 1
        size = _as_i32(sum(ilog2i(x) for x in counts.values()))
        if self.best is None or size > self.best:
            self.best = size
            return True
        return False


EVALUATORS = {1: MinSum, 2: Entropy, 3: Bigrams, 4: BigEnt}


def filter_image(ih, data, kind, basic=0, predefined=None, optimize_alpha_flag=False):
    """PngImage::filter_image, png/mod.rs:355-431 -> (bytes, filters_used list, used_is_predefined)."""
    byte_depth = 2 if ih[3] == 16 else 1
    bpp = byte_depth * CHANNELS[ih[2]]
    alpha_bytes = byte_depth if (optimize_alpha_flag and ih[2] in (4, 6)) else 0
    output = bytearray()
    prev_line = bytearray()
    prev_pass = None
    filters_used = []
    applied = []
    evaluator = EVALUATORS[kind]() if kind in EVALUATORS else None
    off = 0
    for i, (ln, pas, _px) in enumerate(scan_line_ranges(ih, len(data), False)):
        line = data[off:off + ln]
        off += ln
        if prev_pass != pas or len(prev_line) == 0:
            prev_line = bytearray(ln)
            prev_pass = pas
        line_data = bytearray(line)
        if kind == 0 or kind == 5:
            f = basic if kind == 0 else (predefined[i] if predefined is not None and i < len(predefined) else 0)
            filter_line(f, bpp, line_data, prev_line, output, alpha_bytes)
            prev_line = line_data
            applied.append(f)
            continue
        best_filter = 0
        if all(x == 0 for x in line_data):
            filter_line(0, bpp, line_data, prev_line, output, alpha_bytes)
            prev_line = line_data
            filters_used.append(0)
            continue
        best_line = bytearray(ln + 1)
        best_line_raw = bytearray()
        evaluator.reset()
        for f in range(5):
            trial = bytearray()
            filter_line(f, bpp, line_data, prev_line, trial, alpha_bytes)
            if evaluator.evaluate(trial):
                best_line = bytearray(trial)
                best_line_raw = bytearray(line_data)
                best_filter = f
        output.extend(best_line)
        prev_line = best_line_raw
        filters_used.append(best_filter)
    if filters_used:
        return bytes(output), filters_used, True
    return bytes(output), applied, False


# ---- reductions (small cases) ------------------------------------------------------------------------------------

def cleaned_alpha_channel(ih, data):  # alpha.rs:11-32
    if ih[2] not in (4, 6):
        return None
    bd = 2 if ih[3] == 16 else 1
    bpp = CHANNELS[ih[2]] * bd
    cb = bpp - bd
    out = bytearray()
    for i in range(0, len(data) - bpp + 1, bpp):
        px = data[i:i + bpp]
        out.extend(bytes(bpp) if all(b == 0 for b in px[cb:]) else px)
    return bytes(out)


def reduced_alpha_channel(ih, data, optimize_alpha_flag):  # alpha.rs:35-108 -> (ihdr, data, trns) or None
    if ih[2] not in (4, 6):
        return None
    bd = 2 if ih[3] == 16 else 1
    bpp = CHANNELS[ih[2]] * bd
    cb = bpp - bd
    has_transparency = False
    used = [False] * 256
    pixels = [data[i:i + bpp] for i in range(0, len(data) - bpp + 1, bpp)]
    for px in pixels:
        if optimize_alpha_flag and all(b == 0 for b in px[cb:]):
            has_transparency = True
        elif any(b != 255 for b in px[cb:]):
            return None
        elif optimize_alpha_flag and all(b == px[0] for b in px[:cb]):
            used[px[0]] = True
    trns = None
    if has_transparency:
        if ih[2] == 4:
            trns = next((v for v in (0x00, 0xFF, 0x55, 0xAA) if not used[v]), None)
        if trns is None:
            trns = next((v for v in range(256) if not used[v]), None)
        if trns is None:
            return None
    raw = bytearray()
    for px in pixels:
        if trns is not None and all(b == 0 for b in px[cb:]):
            raw.extend([trns] * cb)
        else:
            raw.extend(px[:cb])
    t = None if trns is None else ((trns << 8) | trns if ih[3] == 16 else trns)
    return (ih[0], ih[1], 0 if ih[2] == 4 else 2, ih[3], ih[4]), bytes(raw), t


def reduced_bit_depth_16_to_8(ih, data, force_scale):  # bit_depth.rs:9-64 (scale via exact rational rounding)
    if ih[3] != 16:
        return None
    pairs = [(data[i], data[i + 1]) for i in range(0, len(data) - 1, 2)]
    if not force_scale:
        if any(h != l for h, l in pairs):
            return None
        return bytes(h for h, _ in pairs)
    out = bytearray()
    for h, l in pairs:
        if h == l:
            out.append(h)
        else:
            import struct
            v = struct.unpack("f", struct.pack("f", float((h << 8) | l)))[0]
            k = struct.unpack("f", struct.pack("f", 255.0 / 65535.0))[0]
            prod = struct.unpack("f", struct.pack("f", v * k))[0]  # f32 multiply
            out.append(int(prod + 0.5) & 255)  # round half away from zero for non-negative values
    return bytes(out)


def reduced_rgb_to_grayscale(ih, data):  # color.rs:100-137
    if ih[2] not in (2, 6):
        return None
    bd = 2 if ih[3] == 16 else 1
    bpp = CHANNELS[ih[2]] * bd
    out = bytearray()
    for i in range(0, len(data) - bpp + 1, bpp):
        px = data[i:i + bpp]
        if bd == 1:
            if px[0] != px[1] or px[1] != px[2]:
                return None
        elif px[0:2] != px[2:4] or px[2:4] != px[4:6]:
            return None
        out.extend(px[2 * bd:])
    return bytes(out)


def reduced_to_indexed(ih, data, allow_grayscale):  # color.rs:18-97 -> (index bytes, [pixel tuples in order]) or None
    if ih[3] != 8 or ih[2] == 3:
        return None
    if not allow_grayscale and ih[2] in (0, 4):
        return None
    ch = CHANNELS[ih[2]]
    palette = {}
    out = bytearray()
    for i in range(0, len(data) - ch + 1, ch):
        px = tuple(data[i:i + ch])
        idx = palette.setdefault(px, len(palette))
        if idx == 256:
            return None
        out.append(idx)
    return bytes(out), list(palette.keys())


def reduced_bit_depth_8_or_less_gray_bits(data):  # bit_depth.rs:82-112, the state machine itself
    minimum_bits, mask, divisions = 1, 1, 7

    def rotl(b, r):
        r &= 7
        return ((b << r) | (b >> (8 - r))) & 255 if r else b

    for b in data:
        if b == 0 or b == 255:
            continue
        while True:
            byte = rotl(b, minimum_bits)
            compare = byte & mask
            ok = True
            for _ in range(divisions):
                byte = rotl(byte, minimum_bits)
                if byte & mask != compare:
                    ok = False
                    break
            if ok:
                break
            minimum_bits <<= 1
            if minimum_bits == 8:
                return None
            mask = (1 << minimum_bits) - 1
            divisions = 8 // minimum_bits - 1
    return minimum_bits


def cooccurrence(ih, n, data):  # palette.rs:556-589
    m = [0] * (n * n)
    prev = None
    off = 0
    for ln, _p, _px in scan_line_ranges(ih, len(data), False):
        line = data[off:off + ln]
        off += ln
        prev_val = None
        for i, val in enumerate(line):
            if prev_val is not None:
                m[prev_val * n + val] += 1
            prev_val = val
            if prev is not None and i < len(prev):
                m[prev[i] * n + val] += 1
        prev = line
    for i in range(n):
        for j in range(i + 1):
            m[i * n + j] += m[j * n + i]
            m[j * n + i] = m[i * n + j]
    return m
