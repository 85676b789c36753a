/*
 * oxi_oracle.c -- CPU restatement (plain C) of oxipng 10.1.1's filter-search + reduction-scan path.
 *
 * TEST INFRASTRUCTURE ONLY (see oxi_oracle.h). "parity unpinned" at byte level: the Rust reference
 * cannot be built here and holds no byte-level golden vectors for this path; see the header for
 * what the oracle is pinned against instead.
 *
 * Citations are file:line relative to /root/reference.
 */
#include "oxi_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* geometry                                                                                   */
/* ------------------------------------------------------------------------------------------ */

/* src/colors.rs:59-66 */
static size_t channels_per_pixel(uint8_t color_type) {
    switch (color_type) {
    case 0: case 3: return 1;
    case 4: return 2;
    case 2: return 3;
    case 6: return 4;
    default: return 0;
    }
}
static int has_alpha(uint8_t ct) { return ct == 4 || ct == 6; } /* src/colors.rs:79 */

/* src/headers.rs:32 */
size_t oxo_bits_per_pixel(const oxo_ihdr *h) { return (size_t)h->bit_depth * channels_per_pixel(h->color_type); }

/* src/png/mod.rs:296-302: depths below 8 round up to one byte per channel */
static size_t bytes_per_channel(const oxo_ihdr *h) { return h->bit_depth == 16 ? 2 : 1; }

/* src/png/mod.rs:361 */
size_t oxo_filter_bpp(const oxo_ihdr *h) { return bytes_per_channel(h) * channels_per_pixel(h->color_type); }

static size_t bitmap_size(size_t bpp, size_t w, size_t h) { return ((w * bpp + 7) / 8) * h; } /* headers.rs:43-45 */

/* src/headers.rs:38-64 */
size_t oxo_raw_data_size(const oxo_ihdr *hd) {
    size_t w = hd->width, h = hd->height, bpp = oxo_bits_per_pixel(hd);
    if (!hd->interlaced) return bitmap_size(bpp, w, h) + h;
    size_t size = bitmap_size(bpp, (w + 7) >> 3, (h + 7) >> 3) + ((h + 7) >> 3);
    if (w > 4) size += bitmap_size(bpp, (w + 3) >> 3, (h + 7) >> 3) + ((h + 7) >> 3);
    size += bitmap_size(bpp, (w + 3) >> 2, (h + 3) >> 3) + ((h + 3) >> 3);
    if (w > 2) size += bitmap_size(bpp, (w + 1) >> 2, (h + 3) >> 2) + ((h + 3) >> 2);
    size += bitmap_size(bpp, (w + 1) >> 1, (h + 1) >> 2) + ((h + 1) >> 2);
    if (w > 1) size += bitmap_size(bpp, w >> 1, (h + 1) >> 1) + ((h + 1) >> 1);
    return size + bitmap_size(bpp, w, h >> 1) + (h >> 1);
}

/* ScanLineRanges iterator state, src/png/scan_lines.rs:55-80 */
typedef struct {
    int interlaced;
    uint32_t pass, row; /* pass.0, pass.1 */
    size_t bits_per_pixel;
    uint32_t width, height;
    size_t left;
    int has_filter;
} scan_iter;

static void scan_init(scan_iter *it, const oxo_ihdr *h, size_t data_len, int has_filter) {
    it->interlaced = h->interlaced != 0;
    it->pass = 1;
    it->row = 0;
    it->bits_per_pixel = oxo_bits_per_pixel(h);
    it->width = h->width;
    it->height = h->height;
    it->left = data_len;
    it->has_filter = has_filter;
}

/* src/png/scan_lines.rs:82-167. Returns 0 when exhausted. */
static int scan_next(scan_iter *it, size_t *len_out, uint8_t *pass_out, size_t *pixels_out) {
    if (it->left == 0) return 0;
    uint32_t pixels_per_line;
    uint8_t current_pass = 0;
    if (it->interlaced) {
        /* tiny-image pass skipping, applied one after another (:94-114) */
        if (it->width < 5 && it->pass == 2) { it->pass = 3; it->row = 4; }
        if (it->height < 5 && it->pass == 3) { it->pass = 4; it->row = 0; }
        if (it->width < 3 && it->pass == 4) { it->pass = 5; it->row = 2; }
        if (it->height < 3 && it->pass == 5) { it->pass = 6; it->row = 0; }
        if (it->width == 1 && it->pass == 6) { it->pass = 7; it->row = 1; }
        uint32_t pixels_factor, y_steps;
        switch (it->pass) { /* :115-123 */
        case 1: case 2: pixels_factor = 8; y_steps = 8; break;
        case 3: pixels_factor = 4; y_steps = 8; break;
        case 4: pixels_factor = 4; y_steps = 4; break;
        case 5: pixels_factor = 2; y_steps = 4; break;
        case 6: pixels_factor = 2; y_steps = 2; break;
        case 7: pixels_factor = 1; y_steps = 2; break;
        default: return 0; /* unreachable!() in the reference */
        }
        pixels_per_line = it->width / pixels_factor;
        uint32_t gap = it->width % pixels_factor; /* :126-141 */
        switch (it->pass) {
        case 1: case 3: case 5: if (gap > 0) pixels_per_line++; break;
        case 2: if (gap >= 5) pixels_per_line++; break;
        case 4: if (gap >= 3) pixels_per_line++; break;
        case 6: if (gap >= 2) pixels_per_line++; break;
        default: break;
        }
        current_pass = (uint8_t)it->pass;
        if ((uint64_t)it->row + y_steps >= it->height) { /* :143-153 */
            it->pass += 1;
            it->row = it->pass == 3 ? 4 : it->pass == 5 ? 2 : it->pass == 7 ? 1 : 0;
        } else {
            it->row += y_steps;
        }
    } else {
        pixels_per_line = it->width;
    }
    size_t bits_per_line = (size_t)pixels_per_line * it->bits_per_pixel;
    size_t len = (bits_per_line + 7) / 8;
    if (it->has_filter) len += 1;
    if (it->left < len) return 0; /* checked_sub(len)? */
    it->left -= len;
    *len_out = len;
    *pass_out = current_pass;
    *pixels_out = pixels_per_line;
    return 1;
}

size_t oxo_scan_lines(const oxo_ihdr *h, size_t data_len, int has_filter, size_t *len, uint8_t *pass, size_t *pixels,
                      size_t cap) {
    scan_iter it;
    scan_init(&it, h, data_len, has_filter);
    size_t n = 0, l, px;
    uint8_t p;
    while (scan_next(&it, &l, &p, &px)) {
        if (n < cap) {
            if (len) len[n] = l;
            if (pass) pass[n] = p;
            if (pixels) pixels[n] = px;
        }
        n++;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* filters                                                                                    */
/* ------------------------------------------------------------------------------------------ */

/* src/filters/mod.rs:242-254 */
uint8_t oxo_paeth_predictor(uint8_t a, uint8_t b, uint8_t c) {
    int p = (int)a + (int)b - (int)c;
    int pa = abs(p - (int)a), pb = abs(p - (int)b), pc = abs(p - (int)c);
    if (pa <= pb && pa <= pc) return a;
    if (pb <= pc) return b;
    return c;
}

static int px_alpha_all_zero(const uint8_t *px, size_t color_bytes, size_t bpp) {
    for (size_t k = color_bytes; k < bpp; k++)
        if (px[k] != 0) return 0;
    return 1;
}

/* src/filters/mod.rs:114-186: rewrite colour bytes of fully transparent pixels, left to right, in place */
static void optimize_alpha_line(int filter, size_t bpp, uint8_t *data, const uint8_t *prev_line, size_t len,
                                size_t color_bytes) {
    if (filter == 0) return; /* :115-118 */
    size_t npix = len / bpp; /* chunks_exact: a trailing partial pixel is ignored */
    for (size_t i = 0; i < npix; i++) {
        uint8_t *px = data + i * bpp;
        if (!px_alpha_all_zero(px, color_bytes, bpp)) continue; /* :123 */
        size_t prev; /* :126-132 */
        if (i == 0) {
            prev = 0;
            for (size_t q = 0; q < npix; q++)
                if (!px_alpha_all_zero(data + q * bpp, color_bytes, bpp)) { prev = q; break; }
        } else {
            prev = i - 1;
        }
        const uint8_t *up = prev_line + i * bpp;
        switch (filter) {
        case 1: /* Sub :140-156 */
            if (prev != i) memcpy(px, data + prev * bpp, color_bytes);
            break;
        case 2: /* Up :157-159 */
            memcpy(px, up, color_bytes);
            break;
        case 3: /* Average :160-170 */
            for (size_t j = 0; j < color_bytes; j++)
                px[j] = i == 0 ? (uint8_t)(up[j] >> 1) : (uint8_t)(((unsigned)px[j - bpp] + (unsigned)up[j]) >> 1);
            break;
        case 4: /* Paeth :171-181 */
            for (size_t j = 0; j < color_bytes; j++) {
                if (i == 0) {
                    uint8_t v = data[prev * bpp + j];
                    px[j] = v < up[j] ? v : up[j];
                } else {
                    px[j] = oxo_paeth_predictor(px[j - bpp], up[j], up[j - bpp]);
                }
            }
            break;
        default: break;
        }
    }
}

/* src/filters/mod.rs:46-111 */
void oxo_filter_line(int filter, size_t bpp, uint8_t *data, const uint8_t *prev, size_t len, uint8_t *out,
                     size_t alpha_bytes) {
    if (alpha_bytes != 0) optimize_alpha_line(filter, bpp, data, prev, len, bpp - alpha_bytes); /* :57-59 */
    out[0] = (uint8_t)filter; /* :62 */
    uint8_t *o = out + 1;
    switch (filter) {
    case 0:
        memcpy(o, data, len);
        break;
    case 1: /* :67-75 */
        for (size_t i = 0; i < len; i++) o[i] = i < bpp ? data[i] : (uint8_t)(data[i] - data[i - bpp]);
        break;
    case 2: /* :76-82 */
        for (size_t i = 0; i < len; i++) o[i] = (uint8_t)(data[i] - prev[i]);
        break;
    case 3: /* :83-95 */
        for (size_t i = 0; i < len; i++) {
            unsigned pred = i < bpp ? (unsigned)(prev[i] >> 1) : (((unsigned)data[i - bpp] + (unsigned)prev[i]) >> 1);
            o[i] = (uint8_t)(data[i] - (uint8_t)pred);
        }
        break;
    case 4: /* :96-109 */
        for (size_t i = 0; i < len; i++) {
            uint8_t pred = i < bpp ? prev[i] : oxo_paeth_predictor(data[i - bpp], prev[i], prev[i - bpp]);
            o[i] = (uint8_t)(data[i] - pred);
        }
        break;
    default: break;
    }
}

/* src/filters/mod.rs:188-239 */
void oxo_unfilter_line(int filter, size_t bpp, const uint8_t *data, const uint8_t *prev, size_t len, uint8_t *out) {
    switch (filter) {
    case 0:
        memcpy(out, data, len);
        break;
    case 1:
        for (size_t i = 0; i < len; i++) out[i] = i < bpp ? data[i] : (uint8_t)(data[i] + out[i - bpp]);
        break;
    case 2:
        for (size_t i = 0; i < len; i++) out[i] = (uint8_t)(data[i] + prev[i]);
        break;
    case 3:
        for (size_t i = 0; i < len; i++) {
            unsigned pred = i < bpp ? (unsigned)(prev[i] >> 1) : (((unsigned)out[i - bpp] + (unsigned)prev[i]) >> 1);
            out[i] = (uint8_t)(data[i] + (uint8_t)pred);
        }
        break;
    case 4:
        for (size_t i = 0; i < len; i++) {
            uint8_t pred = i < bpp ? prev[i] : oxo_paeth_predictor(out[i - bpp], prev[i], prev[i - bpp]);
            out[i] = (uint8_t)(data[i] + pred);
        }
        break;
    default: break;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* strategy evaluators                                                                        */
/* ------------------------------------------------------------------------------------------ */

/* src/filters/strategies.rs:269-272 */
uint32_t oxo_ilog2i(uint32_t i) {
    uint32_t log = 31u - (uint32_t)__builtin_clz(i);
    return i * log + ((i - (1u << log)) << 1);
}

/* strategies.rs:90-94: sum of |byte as i8|, 0x80 counts 128 */
uint64_t oxo_score_minsum(const uint8_t *line, size_t n) {
    uint64_t s = 0;
    for (size_t i = 0; i < n; i++) {
        int v = (int8_t)line[i];
        s += (uint64_t)(v < 0 ? -v : v);
    }
    return s;
}

/* strategies.rs:119-142 (the four sub-histograms are an ILP device; their sum is one histogram) */
int32_t oxo_score_entropy(const uint8_t *line, size_t n) {
    uint32_t hist[256];
    memset(hist, 0, sizeof hist);
    for (size_t i = 0; i < n; i++) hist[line[i]]++;
    uint32_t acc = 0; /* u32 fold, wrapping in release builds */
    for (int v = 0; v < 256; v++)
        if (hist[v]) acc += oxo_ilog2i(hist[v]);
    return (int32_t)acc;
}

/* strategies.rs:178-188 without the early exit: number of distinct bigrams over windows(2) */
uint64_t oxo_score_bigrams(const uint8_t *line, size_t n) {
    static _Thread_local uint8_t seen[0x10000 / 8];
    memset(seen, 0, sizeof seen);
    uint64_t count = 0;
    for (size_t i = 0; i + 1 < n; i++) {
        unsigned bg = ((unsigned)line[i] << 8) | line[i + 1];
        if (!(seen[bg >> 3] & (1u << (bg & 7)))) {
            seen[bg >> 3] |= (uint8_t)(1u << (bg & 7));
            count++;
        }
    }
    return count;
}

/* strategies.rs:211-220: sum over distinct bigrams of ilog2i(count); order independent */
int32_t oxo_score_bigent(const uint8_t *line, size_t n) {
    static _Thread_local uint32_t counts[0x10000];
    static _Thread_local uint16_t touched[0x10000];
    size_t nt = 0;
    for (size_t i = 0; i + 1 < n; i++) {
        unsigned bg = ((unsigned)line[i] << 8) | line[i + 1];
        if (counts[bg]++ == 0) touched[nt++] = (uint16_t)bg;
    }
    uint32_t acc = 0;
    for (size_t k = 0; k < nt; k++) {
        acc += oxo_ilog2i(counts[touched[k]]);
        counts[touched[k]] = 0;
    }
    return (int32_t)acc;
}

/* Evaluator state mirroring `best_size` of each evaluator (strategies.rs:76-227) */
typedef struct {
    int kind;
    uint64_t best_u; /* MinSum / Bigrams: usize::MAX */
    int32_t best_i;  /* Entropy / BigEnt: i32::MIN */
} evaluator;

static void eval_reset(evaluator *e) {
    e->best_u = UINT64_MAX;
    e->best_i = INT32_MIN;
}

/* returns 1 when this attempt is the best so far (strict comparisons: first filter wins ties) */
static int eval_evaluate(evaluator *e, const uint8_t *line, size_t n) {
    switch (e->kind) {
    case OXO_MINSUM: {
        uint64_t s = oxo_score_minsum(line, n);
        if (s < e->best_u) { e->best_u = s; return 1; } /* :95 */
        return 0;
    }
    case OXO_ENTROPY: {
        int32_t s = oxo_score_entropy(line, n);
        if (s > e->best_i) { e->best_i = s; return 1; } /* :143 */
        return 0;
    }
    case OXO_BIGRAMS: {
        uint64_t s = oxo_score_bigrams(line, n);
        if (s >= e->best_u) return 0; /* :182-184 early exit <=> count >= best */
        e->best_u = s;                /* :189 */
        return 1;
    }
    case OXO_BIGENT: {
        int32_t s = oxo_score_bigent(line, n);
        if (s > e->best_i) { e->best_i = s; return 1; } /* :221 */
        return 0;
    }
    default: return 0;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* filter_image / unfilter_image                                                              */
/* ------------------------------------------------------------------------------------------ */

/* src/png/mod.rs:355-431 */
int oxo_filter_image(const oxo_ihdr *h, const uint8_t *data, size_t data_len, int kind, int basic_filter,
                     const uint8_t *predefined, size_t n_predefined, int optimize_alpha, uint8_t *out,
                     size_t *out_len, uint8_t *filters_used, size_t *n_rows, int *used_is_predefined) {
    if (kind < OXO_BASIC || kind > OXO_PREDEFINED) return -1;
    if (kind == OXO_BASIC && (basic_filter < 0 || basic_filter > 4)) return -1;
    size_t bpp = oxo_filter_bpp(h);
    if (bpp == 0) return -1;
    size_t alpha_bytes = (optimize_alpha && has_alpha(h->color_type)) ? bytes_per_channel(h) : 0; /* :363-367 */

    scan_iter it;
    scan_init(&it, h, data_len, 0);
    size_t max_len = ((size_t)h->width * oxo_bits_per_pixel(h) + 7) / 8;
    uint8_t *prev_line = (uint8_t *)calloc(max_len + 1, 1);
    uint8_t *line_data = (uint8_t *)malloc(max_len + 1);
    uint8_t *best_line = (uint8_t *)malloc(max_len + 2);
    uint8_t *best_raw = (uint8_t *)malloc(max_len + 1);
    uint8_t *trial = (uint8_t *)malloc(max_len + 2);
    if (!prev_line || !line_data || !best_line || !best_raw || !trial) return -1;

    evaluator ev;
    ev.kind = kind;
    eval_reset(&ev);
    int have_prev = 0;
    uint8_t prev_pass = 0;
    size_t off = 0, in_off = 0, i = 0, len, px;
    uint8_t pass;
    while (scan_next(&it, &len, &pass, &px)) {
        if (!have_prev || prev_pass != pass) { /* :375-378 */
            memset(prev_line, 0, len);
            prev_pass = pass;
            have_prev = 1;
        }
        memcpy(line_data, data + in_off, len); /* :380 */
        in_off += len;
        if (kind == OXO_BASIC || kind == OXO_PREDEFINED) { /* :382-394 */
            int f = kind == OXO_BASIC ? basic_filter : (i < n_predefined ? predefined[i] : 0);
            if (f > 4) { free(prev_line); free(line_data); free(best_line); free(best_raw); free(trial); return -1; }
            oxo_filter_line(f, bpp, line_data, prev_line, len, out + off, alpha_bytes);
            off += len + 1;
            memcpy(prev_line, line_data, len);
            if (filters_used) filters_used[i] = (uint8_t)f;
            i++;
            continue;
        }
        int all_zero = 1; /* :398-404 */
        for (size_t k = 0; k < len; k++)
            if (line_data[k]) { all_zero = 0; break; }
        if (all_zero) {
            oxo_filter_line(0, bpp, line_data, prev_line, len, out + off, alpha_bytes);
            off += len + 1;
            memcpy(prev_line, line_data, len);
            if (filters_used) filters_used[i] = 0;
            i++;
            continue;
        }
        int best_filter = 0;
        eval_reset(&ev); /* :411 */
        memset(best_line, 0, len + 1);
        size_t best_raw_len = 0;
        for (int f = 0; f < 5; f++) { /* :412-420, RowFilter::ALL order */
            oxo_filter_line(f, bpp, line_data, prev_line, len, trial, alpha_bytes);
            if (eval_evaluate(&ev, trial, len + 1)) {
                memcpy(best_line, trial, len + 1);
                memcpy(best_raw, line_data, len);
                best_raw_len = len;
                best_filter = f;
            }
        }
        memcpy(out + off, best_line, len + 1); /* :421 */
        off += len + 1;
        /* :422 prev_line = best_line_raw (empty Vec if no trial won; cannot happen: None always wins first) */
        memcpy(prev_line, best_raw, best_raw_len);
        if (best_raw_len == 0) have_prev = 0;
        if (filters_used) filters_used[i] = (uint8_t)best_filter;
        i++;
    }
    if (out_len) *out_len = off;
    if (n_rows) *n_rows = i;
    if (used_is_predefined) *used_is_predefined = (kind != OXO_BASIC && kind != OXO_PREDEFINED && i > 0); /* :426-430 */
    free(prev_line); free(line_data); free(best_line); free(best_raw); free(trial);
    return 0;
}

/* src/png/mod.rs:335-351 */
int oxo_unfilter_image(const oxo_ihdr *h, const uint8_t *filtered, size_t flen, uint8_t *out, size_t *out_len) {
    size_t bpp = oxo_filter_bpp(h);
    scan_iter it;
    scan_init(&it, h, flen, 1);
    size_t max_len = ((size_t)h->width * oxo_bits_per_pixel(h) + 7) / 8;
    uint8_t *prev_line = (uint8_t *)calloc(max_len + 1, 1);
    if (!prev_line) return -1;
    int have_prev = 0;
    uint8_t prev_pass = 0, pass;
    size_t in_off = 0, off = 0, len, px;
    while (scan_next(&it, &len, &pass, &px)) {
        size_t dl = len - 1;
        if (!have_prev || prev_pass != pass) {
            memset(prev_line, 0, dl);
            prev_pass = pass;
            have_prev = 1;
        }
        uint8_t f = filtered[in_off];
        if (f > 4) { free(prev_line); return -1; }
        oxo_unfilter_line(f, bpp, filtered + in_off + 1, prev_line, dl, out + off);
        memcpy(prev_line, out + off, dl);
        in_off += len;
        off += dl;
    }
    if (out_len) *out_len = off;
    free(prev_line);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* interlacing, src/interlace.rs                                                              */
/* ------------------------------------------------------------------------------------------ */

static int get_bit(const uint8_t *p, size_t bit) { return (p[bit >> 3] >> (7 - (bit & 7))) & 1; }
static void put_bit(uint8_t *p, size_t bit, int v) {
    if (v) p[bit >> 3] |= (uint8_t)(0x80u >> (bit & 7));
    else p[bit >> 3] &= (uint8_t)~(0x80u >> (bit & 7));
}

/* interlace.rs:186-232 */
static const uint8_t A7_X0[8] = {0, 0, 4, 0, 2, 0, 1, 0};
static const uint8_t A7_Y0[8] = {0, 0, 0, 4, 0, 2, 0, 1};
static const uint8_t A7_DX[8] = {0, 8, 8, 4, 4, 2, 2, 1};
static const uint8_t A7_DY[8] = {0, 8, 8, 8, 4, 4, 2, 2};

/* interlace.rs:6-60. The reference routes each source pixel by (row % 8, pixel % 8) into one of seven
 * MSB-first bit vectors and pads every pass to a byte boundary after each source row; that is the
 * Adam7 pass decomposition, stated here pass by pass. */
int oxo_interlace_image(const oxo_ihdr *h, const uint8_t *data, size_t len, uint8_t *out, size_t *out_len) {
    size_t bpp = oxo_bits_per_pixel(h), w = h->width, ht = h->height;
    size_t src_stride = (w * bpp + 7) / 8;
    if (len < src_stride * ht) return -1;
    size_t off = 0;
    for (int p = 1; p <= 7; p++) {
        if (w <= A7_X0[p] || ht <= A7_Y0[p]) continue;
        size_t pw = (w - A7_X0[p] + A7_DX[p] - 1) / A7_DX[p];
        size_t ph = (ht - A7_Y0[p] + A7_DY[p] - 1) / A7_DY[p];
        size_t stride = (pw * bpp + 7) / 8;
        for (size_t r = 0; r < ph; r++) {
            const uint8_t *src = data + (A7_Y0[p] + r * A7_DY[p]) * src_stride;
            uint8_t *dst = out + off;
            memset(dst, 0, stride);
            for (size_t c = 0; c < pw; c++) {
                size_t sx = A7_X0[p] + c * A7_DX[p];
                if (bpp >= 8) {
                    memcpy(dst + c * (bpp / 8), src + sx * (bpp / 8), bpp / 8);
                } else {
                    for (size_t b = 0; b < bpp; b++) put_bit(dst, c * bpp + b, get_bit(src, sx * bpp + b));
                }
            }
            off += stride;
        }
    }
    if (out_len) *out_len = off;
    return 0;
}

/* interlace.rs:62-150 */
int oxo_deinterlace_image(const oxo_ihdr *h, const uint8_t *data, size_t len, uint8_t *out, size_t *out_len) {
    size_t bpp = oxo_bits_per_pixel(h), w = h->width, ht = h->height;
    size_t dst_stride = (w * bpp + 7) / 8;
    memset(out, 0, dst_stride * ht);
    size_t off = 0;
    for (int p = 1; p <= 7; p++) {
        if (w <= A7_X0[p] || ht <= A7_Y0[p]) continue;
        size_t pw = (w - A7_X0[p] + A7_DX[p] - 1) / A7_DX[p];
        size_t ph = (ht - A7_Y0[p] + A7_DY[p] - 1) / A7_DY[p];
        size_t stride = (pw * bpp + 7) / 8;
        for (size_t r = 0; r < ph; r++) {
            if (off + stride > len) return -1;
            const uint8_t *src = data + off;
            uint8_t *dst = out + (A7_Y0[p] + r * A7_DY[p]) * dst_stride;
            for (size_t c = 0; c < pw; c++) {
                size_t dx = A7_X0[p] + c * A7_DX[p];
                if (bpp >= 8) {
                    memcpy(dst + dx * (bpp / 8), src + c * (bpp / 8), bpp / 8);
                } else {
                    for (size_t b = 0; b < bpp; b++) put_bit(dst, dx * bpp + b, get_bit(src, c * bpp + b));
                }
            }
            off += stride;
        }
    }
    if (out_len) *out_len = dst_stride * ht;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* reductions                                                                                 */
/* ------------------------------------------------------------------------------------------ */

/* src/reduction/alpha.rs:11-32 */
int oxo_cleaned_alpha_channel(const oxo_ihdr *h, const uint8_t *data, size_t len, uint8_t *out) {
    if (!has_alpha(h->color_type)) return 1;
    size_t bd = bytes_per_channel(h), bpp = channels_per_pixel(h->color_type) * bd, cb = bpp - bd;
    size_t n = len / bpp;
    for (size_t i = 0; i < n; i++) {
        const uint8_t *px = data + i * bpp;
        if (px_alpha_all_zero(px, cb, bpp)) memset(out + i * bpp, 0, bpp);
        else memcpy(out + i * bpp, px, bpp);
    }
    return 0;
}

/* src/reduction/alpha.rs:35-108. out needs len bytes. */
int oxo_reduced_alpha_channel(const oxo_ihdr *h, const uint8_t *data, size_t len, int optimize_alpha, uint8_t *out,
                              size_t *out_len, oxo_ihdr *oh, oxo_aux *oaux) {
    if (!has_alpha(h->color_type)) return 1;
    size_t bd = bytes_per_channel(h), bpp = channels_per_pixel(h->color_type) * bd, cb = bpp - bd;
    size_t n = len / bpp;
    int has_transparency = 0;
    uint8_t used[256];
    memset(used, 0, sizeof used);
    for (size_t i = 0; i < n; i++) { /* :49-60 */
        const uint8_t *px = data + i * bpp;
        if (optimize_alpha && px_alpha_all_zero(px, cb, bpp)) {
            has_transparency = 1;
            continue;
        }
        int partial = 0;
        for (size_t k = cb; k < bpp; k++)
            if (px[k] != 255) partial = 1;
        if (partial) return 1;
        if (optimize_alpha) {
            int gray = 1;
            for (size_t k = 0; k < cb; k++)
                if (px[k] != px[0]) gray = 0;
            if (gray) used[px[0]] = 1;
        }
    }
    int trns = -1;
    if (has_transparency) { /* :62-75 */
        if (h->color_type == 4) {
            static const uint8_t tries[4] = {0x00, 0xFF, 0x55, 0xAA};
            for (int k = 0; k < 4 && trns < 0; k++)
                if (!used[tries[k]]) trns = tries[k];
        }
        for (int v = 0; v < 256 && trns < 0; v++)
            if (!used[v]) trns = v;
        if (trns < 0) return 1;
    }
    size_t o = 0;
    for (size_t i = 0; i < n; i++) { /* :77-85 */
        const uint8_t *px = data + i * bpp;
        if (trns >= 0 && px_alpha_all_zero(px, cb, bpp)) memset(out + o, trns, cb);
        else memcpy(out + o, px, cb);
        o += cb;
    }
    if (out_len) *out_len = o;
    if (oh) {
        *oh = *h;
        oh->color_type = h->color_type == 4 ? 0 : 2;
    }
    if (oaux) {
        memset(oaux, 0, sizeof *oaux);
        if (trns >= 0) { /* :88-91 */
            uint16_t t = h->bit_depth == 16 ? (uint16_t)((trns << 8) | trns) : (uint16_t)trns;
            oaux->has_trns = 1;
            oaux->trns[0] = oaux->trns[1] = oaux->trns[2] = t;
        }
    }
    return 0;
}

/* src/reduction/bit_depth.rs:9-64. out needs len/2 bytes. */
int oxo_reduced_bit_depth_16_to_8(const oxo_ihdr *h, const uint8_t *data, size_t len, int force_scale, uint8_t *out,
                                  size_t *out_len, oxo_ihdr *oh) {
    if (h->bit_depth != 16) return 1;
    size_t n = len / 2;
    if (!force_scale) {
        for (size_t i = 0; i < n; i++)
            if (data[2 * i] != data[2 * i + 1]) return 1; /* :19-22 */
        for (size_t i = 0; i < n; i++) out[i] = data[2 * i];
    } else {
        const float k = 255.0f / 65535.0f; /* :51-52, f32 arithmetic */
        for (size_t i = 0; i < n; i++) {
            uint8_t hi = data[2 * i], lo = data[2 * i + 1];
            if (hi == lo) { out[i] = hi; continue; }
            float val = (float)(uint16_t)((hi << 8) | lo);
            out[i] = (uint8_t)roundf(val * k);
        }
    }
    if (out_len) *out_len = n;
    if (oh) { *oh = *h; oh->bit_depth = 8; }
    return 0;
}

static uint8_t rotl8(uint8_t b, unsigned r) { r &= 7; return (uint8_t)((b << r) | (b >> ((8 - r) & 7))); }

/* src/reduction/bit_depth.rs:68-166. out needs len bytes. */
int oxo_reduced_bit_depth_8_or_less(const oxo_ihdr *h, const oxo_aux *aux, const uint8_t *data, size_t len,
                                    uint8_t *out, size_t *out_len, oxo_ihdr *oh, oxo_aux *oaux) {
    if (h->bit_depth != 8 || channels_per_pixel(h->color_type) != 1) return 1;
    unsigned minimum_bits;
    if (h->color_type == 3) { /* :73-80 */
        uint32_t n = aux ? aux->n_palette : 0;
        if (n <= 2) minimum_bits = 1;
        else if (n <= 4) minimum_bits = 2;
        else if (n <= 16) minimum_bits = 4;
        else return 1;
    } else { /* :82-112 */
        minimum_bits = 1;
        unsigned mask = 1, ndiv = 7; /* divisions = 1..8 */
        for (size_t i = 0; i < len; i++) {
            uint8_t b = data[i];
            if (b == 0 || b == 255) continue;
            for (;;) {
                uint8_t byte = rotl8(b, minimum_bits);
                unsigned compare = byte & mask;
                int ok = 1;
                for (unsigned d = 0; d < ndiv; d++) {
                    byte = rotl8(byte, minimum_bits);
                    if ((byte & mask) != compare) { ok = 0; break; }
                }
                if (ok) break;
                minimum_bits <<= 1;
                if (minimum_bits == 8) return 1;
                mask = (1u << minimum_bits) - 1;
                ndiv = 8 / minimum_bits - 1; /* 1..(8/minimum_bits) */
            }
        }
    }
    unsigned mask = (1u << minimum_bits) - 1, per = 8 / minimum_bits;
    scan_iter it;
    scan_init(&it, h, len, 0);
    size_t in_off = 0, o = 0, l, px;
    uint8_t pass;
    while (scan_next(&it, &l, &pass, &px)) { /* :117-131 */
        for (size_t c = 0; c < l; c += per) {
            unsigned nb = 0, shift = 8;
            for (size_t k = c; k < c + per && k < l; k++) {
                shift -= minimum_bits;
                nb |= (data[in_off + k] & mask) << shift;
            }
            out[o++] = (uint8_t)nb;
        }
        in_off += l;
    }
    if (out_len) *out_len = o;
    if (oh) { *oh = *h; oh->bit_depth = (uint8_t)minimum_bits; }
    if (oaux) {
        if (aux) *oaux = *aux; else memset(oaux, 0, sizeof *oaux);
        if (h->color_type == 0 && aux && aux->has_trns) { /* :134-153 */
            uint16_t trans = aux->trns[0];
            uint16_t reduced_trans = (uint16_t)((trans & 0xFF) >> (8 - minimum_bits));
            uint16_t check = reduced_trans;
            unsigned bits = minimum_bits;
            while (bits < 8) { check = (uint16_t)((check << bits) | check); bits <<= 1; }
            if (trans == check) { oaux->has_trns = 1; oaux->trns[0] = reduced_trans; }
            else { oaux->has_trns = 0; oaux->trns[0] = 0; }
        }
    }
    return 0;
}

/* src/reduction/bit_depth.rs:170-230. out needs width*height bytes (+7 slack per row). */
int oxo_expanded_bit_depth_to_8(const oxo_ihdr *h, const oxo_aux *aux, const uint8_t *data, size_t len, uint8_t *out,
                                size_t *out_len, oxo_ihdr *oh, oxo_aux *oaux) {
    unsigned depth = h->bit_depth;
    if (depth >= 8) return 1;
    unsigned ppb = 8 / depth, mask = (1u << depth) - 1;
    int is_gray = h->color_type == 0;
    scan_iter it;
    scan_init(&it, h, len, 0);
    size_t in_off = 0, o = 0, length = 0, l, px;
    uint8_t pass;
    while (scan_next(&it, &l, &pass, &px)) {
        for (size_t k = 0; k < l; k++) {
            uint8_t byte = data[in_off + k];
            for (unsigned q = 0; q < ppb; q++) {
                byte = rotl8(byte, depth);
                unsigned val = byte & mask;
                if (is_gray) {
                    unsigned bits = depth;
                    while (bits < 8) { val = ((val << bits) | val) & 0xFF; bits <<= 1; }
                }
                if (o < length + px) out[o++] = (uint8_t)val; /* push then truncate(length) :200-202 */
            }
        }
        length += px;
        o = length;
        in_off += l;
    }
    if (out_len) *out_len = o;
    if (oh) { *oh = *h; oh->bit_depth = 8; }
    if (oaux) {
        if (aux) *oaux = *aux; else memset(oaux, 0, sizeof *oaux);
        if (is_gray && aux && aux->has_trns) { /* :206-220 */
            uint16_t trans = aux->trns[0];
            unsigned bits = depth;
            while (bits < 8) { trans = (uint16_t)((trans << bits) | trans); bits <<= 1; }
            oaux->trns[0] = trans;
        }
    }
    return 0;
}

/* insertion-ordered set of <=257 keys of up to 4 bytes: IndexSet::insert_full (color.rs:25-33) */
typedef struct {
    uint32_t keys[257];
    uint32_t n;
    int32_t slots[1024]; /* open addressing; value = index+1 */
} ordered_set;

static void oset_init(ordered_set *s) {
    s->n = 0;
    memset(s->slots, 0, sizeof s->slots);
}
static uint32_t oset_insert_full(ordered_set *s, uint32_t key) {
    uint32_t hsh = (key * 2654435761u) >> 22;
    for (;;) {
        int32_t v = s->slots[hsh];
        if (v == 0) {
            s->keys[s->n] = key;
            s->slots[hsh] = (int32_t)(++s->n);
            return s->n - 1;
        }
        if (s->keys[v - 1] == key) return (uint32_t)(v - 1);
        hsh = (hsh + 1) & 1023;
    }
}

/* src/reduction/color.rs:18-97. out needs len/channels bytes. */
int oxo_reduced_to_indexed(const oxo_ihdr *h, const oxo_aux *aux, const uint8_t *data, size_t len,
                           int allow_grayscale, uint8_t *out, size_t *out_len, oxo_ihdr *oh, oxo_aux *oaux) {
    if (h->bit_depth != 8) return 1;
    if (h->color_type == 3) return 1;
    int is_gray = h->color_type == 0 || h->color_type == 4;
    if (!allow_grayscale && is_gray) return 1;
    size_t ch = channels_per_pixel(h->color_type);
    if (ch == 0) return 1;
    size_t n = len / ch;
    ordered_set *set = (ordered_set *)malloc(sizeof *set);
    if (!set) return -1;
    oset_init(set);
    for (size_t i = 0; i < n; i++) { /* build_palette :18-35 */
        uint32_t key = 0;
        for (size_t k = 0; k < ch; k++) key |= (uint32_t)data[i * ch + k] << (8 * k);
        uint32_t idx = oset_insert_full(set, key);
        if (idx == 256) { free(set); return 1; }
        out[i] = (uint8_t)idx;
    }
    if (out_len) *out_len = n;
    if (oh) { *oh = *h; oh->color_type = 3; }
    if (oaux) {
        memset(oaux, 0, sizeof *oaux);
        oaux->n_palette = set->n;
        for (uint32_t p = 0; p < set->n; p++) {
            uint32_t key = set->keys[p];
            uint8_t c0 = key & 0xFF, c1 = (key >> 8) & 0xFF, c2 = (key >> 16) & 0xFF, c3 = (key >> 24) & 0xFF;
            uint8_t *e = oaux->palette[p];
            switch (h->color_type) {
            case 0: { /* :51-62 */
                int is_t = aux && aux->has_trns && (uint8_t)aux->trns[0] == c0;
                e[0] = e[1] = e[2] = c0; e[3] = is_t ? 0 : 255;
                break;
            }
            case 2: { /* :64-77 */
                int is_t = aux && aux->has_trns && (uint8_t)aux->trns[0] == c0 && (uint8_t)aux->trns[1] == c1 &&
                           (uint8_t)aux->trns[2] == c2;
                e[0] = c0; e[1] = c1; e[2] = c2; e[3] = is_t ? 0 : 255;
                break;
            }
            case 4: e[0] = e[1] = e[2] = c0; e[3] = c1; break; /* :79-82 */
            default: e[0] = c0; e[1] = c1; e[2] = c2; e[3] = c3; break; /* :83-86 */
            }
        }
    }
    free(set);
    return 0;
}

/* src/reduction/color.rs:100-137. out needs len bytes. */
int oxo_reduced_rgb_to_grayscale(const oxo_ihdr *h, const oxo_aux *aux, const uint8_t *data, size_t len, uint8_t *out,
                                 size_t *out_len, oxo_ihdr *oh, oxo_aux *oaux) {
    if (h->color_type != 2 && h->color_type != 6) return 1;
    size_t bd = bytes_per_channel(h), bpp = channels_per_pixel(h->color_type) * bd, last = 2 * bd;
    size_t n = len / bpp, o = 0;
    for (size_t i = 0; i < n; i++) {
        const uint8_t *px = data + i * bpp;
        if (bd == 1) {
            if (px[0] != px[1] || px[1] != px[2]) return 1;
        } else if (px[0] != px[2] || px[1] != px[3] || px[2] != px[4] || px[3] != px[5]) {
            return 1;
        }
        memcpy(out + o, px + last, bpp - last);
        o += bpp - last;
    }
    if (out_len) *out_len = o;
    if (oh) { *oh = *h; oh->color_type = h->color_type == 2 ? 0 : 4; }
    if (oaux) {
        memset(oaux, 0, sizeof *oaux);
        if (h->color_type == 2 && aux && aux->has_trns && aux->trns[0] == aux->trns[1] &&
            aux->trns[1] == aux->trns[2]) { /* :120-128 */
            oaux->has_trns = 1;
            oaux->trns[0] = aux->trns[0];
        }
    }
    return 0;
}

/* src/reduction/color.rs:141-206. out needs 4*len bytes. */
int oxo_indexed_to_channels(const oxo_ihdr *h, const oxo_aux *aux, const uint8_t *data, size_t len,
                            int allow_grayscale, int optimize_alpha, uint8_t *out, size_t *out_len, oxo_ihdr *oh,
                            oxo_aux *oaux) {
    if (h->bit_depth != 8 || h->color_type != 3 || !aux) return 1;
    uint8_t pal[256][4];
    uint32_t np = aux->n_palette;
    memcpy(pal, aux->palette, sizeof pal);
    if (optimize_alpha)
        for (uint32_t i = 0; i < np; i++)
            if (pal[i][3] == 0) pal[i][0] = pal[i][1] = pal[i][2] = 0;
    int is_gray = 0, alpha = 0;
    if (allow_grayscale) {
        is_gray = 1;
        for (uint32_t i = 0; i < np; i++)
            if (pal[i][0] != pal[i][1] || pal[i][1] != pal[i][2]) is_gray = 0;
    }
    for (uint32_t i = 0; i < np; i++)
        if (pal[i][3] != 255) alpha = 1;
    uint8_t ct = is_gray ? (alpha ? 4 : 0) : (alpha ? 6 : 2);
    size_t out_size = channels_per_pixel(ct) * len;
    if (out_size - len > 20000) return 1; /* INDEXED_MAX_DIFF color.rs:16,184-187 */
    size_t s = is_gray ? 2 : 0, e = alpha ? 3 : 2, o = 0;
    static const uint8_t black[4] = {0, 0, 0, 255};
    for (size_t i = 0; i < len; i++) {
        const uint8_t *c = data[i] < np ? pal[data[i]] : black;
        for (size_t k = s; k <= e; k++) out[o++] = c[k];
    }
    if (out_len) *out_len = o;
    if (oh) { *oh = *h; oh->color_type = ct; }
    if (oaux) memset(oaux, 0, sizeof *oaux);
    return 0;
}

/* src/reduction/palette.rs:12-73. out needs len bytes. */
int oxo_reduced_palette(const oxo_ihdr *h, const oxo_aux *aux, const uint8_t *data, size_t len, int optimize_alpha,
                        uint8_t *out, oxo_aux *oaux) {
    if (h->bit_depth != 8 || h->color_type != 3 || !aux) return 1;
    uint8_t used[256];
    memset(used, 0, sizeof used);
    for (size_t i = 0; i < len; i++) used[data[i]] = 1; /* :20-23 */
    uint8_t condensed[256][4];
    uint32_t nc = 0;
    uint8_t byte_map[256];
    memset(byte_map, 0, sizeof byte_map);
    int did_change = 0;
    for (int i = 0; i < 256; i++) { /* :28-39 */
        if (!used[i]) continue;
        uint8_t c[4] = {0, 0, 0, 255};
        if ((uint32_t)i < aux->n_palette) memcpy(c, aux->palette[i], 4);
        if (optimize_alpha && c[3] == 0) c[0] = c[1] = c[2] = 0; /* add_color_to_set :64-73 */
        uint32_t idx = nc;
        for (uint32_t k = 0; k < nc; k++)
            if (!memcmp(condensed[k], c, 4)) { idx = k; break; }
        if (idx == nc) memcpy(condensed[nc++], c, 4);
        byte_map[i] = (uint8_t)idx;
        if (idx != (uint32_t)i) did_change = 1;
    }
    if (did_change) {
        for (size_t i = 0; i < len; i++) out[i] = byte_map[data[i]]; /* :41-43 */
    } else if (nc != aux->n_palette) {
        memcpy(out, data, len); /* :44-47 */
    } else {
        return 1;
    }
    if (oaux) {
        memset(oaux, 0, sizeof *oaux);
        oaux->n_palette = nc;
        memcpy(oaux->palette, condensed, (size_t)nc * 4);
    }
    return 0;
}

/* src/reduction/palette.rs:197-217 */
int oxo_most_popular_edge_color(const oxo_ihdr *h, size_t num_colors, const uint8_t *data, size_t len) {
    uint32_t counts[256];
    memset(counts, 0, sizeof counts);
    scan_iter it;
    scan_init(&it, h, len, 0);
    size_t off = 0, l, px;
    uint8_t pass;
    while (scan_next(&it, &l, &pass, &px)) {
        if (l >= 2) { /* &[first, .., last] needs two elements */
            counts[data[off]]++;
            counts[data[off + l - 1]]++;
        }
        off += l;
    }
    if (num_colors == 0) return -1;
    if (num_colors > 256) num_colors = 256;
    size_t best = 0; /* max_by_key returns the last maximum */
    for (size_t i = 0; i < num_colors; i++)
        if (counts[i] >= counts[best]) best = i;
    int equal = 0;
    for (int i = 0; i < 256; i++)
        if (counts[i] == counts[best]) equal++;
    return equal > 1 ? -1 : (int)best;
}

/* key of palette.rs:92-105 */
static int32_t luma_key(const uint8_t *c) {
    int32_t a = c[3];
    return ((a & 0xFE) << 18) + (a & 0x01) - (int32_t)c[0] * 299 - (int32_t)c[1] * 587 - (int32_t)c[2] * 114;
}

/* src/reduction/palette.rs:77-130. out needs len bytes. */
int oxo_sorted_palette(const oxo_ihdr *h, const oxo_aux *aux, const uint8_t *data, size_t len, uint8_t *out,
                       oxo_aux *oaux) {
    if (h->bit_depth != 8 || h->color_type != 3 || !aux || aux->n_palette <= 1) return 1;
    uint32_t n = aux->n_palette;
    uint32_t order[256], m = 0;
    int keep_first = oxo_most_popular_edge_color(h, n, data, len);
    for (uint32_t i = 0; i < n; i++)
        if ((int)i != keep_first) order[m++] = i;
    for (uint32_t i = 1; i < m; i++) { /* stable sort by key ascending */
        uint32_t v = order[i];
        int32_t kv = luma_key(aux->palette[v]);
        uint32_t j = i;
        while (j > 0 && luma_key(aux->palette[order[j - 1]]) > kv) { order[j] = order[j - 1]; j--; }
        order[j] = v;
    }
    uint32_t remap[256];
    uint32_t r = 0;
    if (keep_first >= 0) remap[r++] = (uint32_t)keep_first;
    for (uint32_t i = 0; i < m; i++) remap[r++] = order[i];
    int identity = 1;
    for (uint32_t i = 0; i < n; i++)
        if (remap[i] != i) identity = 0;
    if (identity) return 1;
    uint8_t byte_map[256];
    memset(byte_map, 0, sizeof byte_map);
    for (uint32_t i = 0; i < n; i++) byte_map[remap[i]] = (uint8_t)i;
    for (size_t i = 0; i < len; i++) out[i] = byte_map[data[i]];
    if (oaux) {
        memset(oaux, 0, sizeof *oaux);
        oaux->n_palette = n;
        for (uint32_t i = 0; i < n; i++) memcpy(oaux->palette[i], aux->palette[remap[i]], 4);
    }
    return 0;
}

/* src/reduction/palette.rs:556-589 */
int oxo_cooccurrence_build(const oxo_ihdr *h, uint32_t nc, const uint8_t *data, size_t len, uint32_t *m) {
    memset(m, 0, (size_t)nc * nc * sizeof(uint32_t));
    scan_iter it;
    scan_init(&it, h, len, 0);
    size_t off = 0, l, px, prev_off = 0, prev_len = 0;
    int have_prev = 0;
    uint8_t pass;
    while (scan_next(&it, &l, &pass, &px)) {
        for (size_t i = 0; i < l; i++) {
            size_t val = data[off + i];
            if (val >= nc) return -1; /* the reference would index out of bounds (panic) */
            if (i > 0) m[(size_t)data[off + i - 1] * nc + val]++;
            if (have_prev && i < prev_len) m[(size_t)data[prev_off + i] * nc + val]++;
        }
        prev_off = off;
        prev_len = l;
        have_prev = 1;
        off += l;
    }
    for (uint32_t i = 0; i < nc; i++) /* :580-587 */
        for (uint32_t j = 0; j <= i; j++) {
            size_t a = (size_t)i * nc + j, b = (size_t)j * nc + i;
            m[a] += m[b];
            m[b] = m[a];
        }
    return 0;
}

/* src/reduction/palette.rs:165-194. out needs len bytes. */
int oxo_apply_palette_reorder(const oxo_ihdr *h, const oxo_aux *aux, const uint32_t *remapping, const uint8_t *data,
                              size_t len, uint8_t *out, oxo_aux *oaux) {
    if (h->color_type != 3 || !aux) return 1;
    uint32_t n = aux->n_palette;
    int identity = 1;
    for (uint32_t i = 0; i < n; i++)
        if (remapping[i] != i) identity = 0;
    if (identity) return 1;
    uint8_t byte_map[256];
    memset(byte_map, 0, sizeof byte_map);
    if (oaux) { memset(oaux, 0, sizeof *oaux); oaux->n_palette = n; }
    for (uint32_t i = 0; i < n; i++) {
        if (remapping[i] >= n) return -1;
        if (oaux) memcpy(oaux->palette[i], aux->palette[remapping[i]], 4);
        byte_map[remapping[i]] = (uint8_t)i;
    }
    for (size_t i = 0; i < len; i++) out[i] = byte_map[data[i]];
    return 0;
}
