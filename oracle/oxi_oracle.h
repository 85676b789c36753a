/*
 * oxi_oracle.h -- CPU restatement of oxipng 10.1.1's filter-search and reduction-scan path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker (or the timed CPU baseline), never as the thing shipped.
 *
 * PARITY STATUS: "parity unpinned" at byte level.  The reference (Rust) cannot be built in this
 * image (no cargo/rustc) and its own tests hold no byte-level golden vectors for this path
 * (SURVEY.md section 8c).  The oracle is pinned instead against (1) the PNG specification via an
 * independent decoder (PIL) on streams it produces, (2) the outcomes the reference's fixture
 * names assert (tests/reduction.rs `<in>_should_be_<out>` wherever no deflate-size comparison is
 * involved), (3) hand-computed micro cases, (4) an independently written pure-Python
 * restatement (oracle/oxi_oracle_py.py).
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 */
#ifndef OXI_ORACLE_H
#define OXI_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/headers.rs:15-26 (IhdrData); colour-type codes src/colors.rs:48-56 */
typedef struct {
    uint32_t width, height;
    uint8_t color_type; /* 0 gray, 2 rgb, 3 indexed, 4 gray+alpha, 6 rgba */
    uint8_t bit_depth;  /* 1,2,4,8,16 */
    uint8_t interlaced; /* 0/1 */
} oxo_ihdr;

/* The parts of ColorType that carry data (src/colors.rs:9-29): palette or tRNS colour. */
typedef struct {
    uint32_t n_palette;      /* Indexed: palette.len() */
    uint8_t palette[256][4]; /* RGBA8 */
    int32_t has_trns;        /* Grayscale.transparent_shade / RGB.transparent_color is Some */
    uint16_t trns[3];        /* gray: trns[0]; rgb: r,g,b */
} oxo_aux;

/* FilterStrategy (src/filters/strategies.rs:9-29) without Brute */
enum { OXO_BASIC = 0, OXO_MINSUM = 1, OXO_ENTROPY = 2, OXO_BIGRAMS = 3, OXO_BIGENT = 4, OXO_PREDEFINED = 5 };

/* --- geometry ------------------------------------------------------------------------ */
size_t oxo_bits_per_pixel(const oxo_ihdr *h);  /* src/headers.rs:32 */
size_t oxo_raw_data_size(const oxo_ihdr *h);   /* src/headers.rs:38-64 */
size_t oxo_filter_bpp(const oxo_ihdr *h);      /* src/png/mod.rs:289-302,361 */
/* ScanLineRanges (src/png/scan_lines.rs:82-167). Writes up to cap entries; returns the count.
 * pass[i] is 0 for non-interlaced (None) else 1..7. Any output pointer may be NULL. */
size_t oxo_scan_lines(const oxo_ihdr *h, size_t data_len, int has_filter, size_t *len, uint8_t *pass,
                      size_t *pixels, size_t cap);

/* --- filters ------------------------------------------------------------------------- */
uint8_t oxo_paeth_predictor(uint8_t a, uint8_t b, uint8_t c); /* src/filters/mod.rs:242-254 */
/* src/filters/mod.rs:46-111 (+ optimize_alpha :114-186). data is mutated when alpha_bytes != 0.
 * Writes len+1 bytes to out. */
void oxo_filter_line(int filter, size_t bpp, uint8_t *data, const uint8_t *prev, size_t len, uint8_t *out,
                     size_t alpha_bytes);
/* src/filters/mod.rs:188-239. Writes len bytes. */
void oxo_unfilter_line(int filter, size_t bpp, const uint8_t *data, const uint8_t *prev, size_t len, uint8_t *out);
uint32_t oxo_ilog2i(uint32_t i); /* src/filters/strategies.rs:269-272 */
/* Scores exactly as the evaluators compute them over output[offset..] (filter byte included). */
uint64_t oxo_score_minsum(const uint8_t *line, size_t n);  /* strategies.rs:90-94 */
int32_t oxo_score_entropy(const uint8_t *line, size_t n);  /* strategies.rs:119-142 */
uint64_t oxo_score_bigrams(const uint8_t *line, size_t n); /* strategies.rs:178-188 (full count) */
int32_t oxo_score_bigent(const uint8_t *line, size_t n);   /* strategies.rs:211-220 */

/* PngImage::filter_image (src/png/mod.rs:355-431).
 * out must hold oxo_raw_data_size() bytes (more if data_len is inconsistent: data_len + rows).
 * filters_used receives one id per scanline (for Basic/Predefined: the filter applied).
 * *used_is_predefined = 1 when the reference would return Predefined(filters_used).
 * Returns 0, or -1 on bad arguments. */
int oxo_filter_image(const oxo_ihdr *h, const uint8_t *data, size_t data_len, int kind, int basic_filter,
                     const uint8_t *predefined, size_t n_predefined, int optimize_alpha, uint8_t *out,
                     size_t *out_len, uint8_t *filters_used, size_t *n_rows, int *used_is_predefined);
/* PngImage::unfilter_image (src/png/mod.rs:335-351). Returns 0, -1 on invalid filter byte. */
int oxo_unfilter_image(const oxo_ihdr *h, const uint8_t *filtered, size_t len, uint8_t *out, size_t *out_len);

/* --- interlacing (src/interlace.rs) ---------------------------------------------------- */
int oxo_interlace_image(const oxo_ihdr *h, const uint8_t *data, size_t len, uint8_t *out, size_t *out_len);
int oxo_deinterlace_image(const oxo_ihdr *h, const uint8_t *data, size_t len, uint8_t *out, size_t *out_len);

/* --- reductions: return 0 = Some(image), 1 = None, <0 = bad arguments -------------------- */
/* out buffers are caller-allocated with room for the worst case (see each function). */
int oxo_cleaned_alpha_channel(const oxo_ihdr *h, const uint8_t *data, size_t len, uint8_t *out); /* alpha.rs:11-32 */
int oxo_reduced_alpha_channel(const oxo_ihdr *h, const uint8_t *data, size_t len, int optimize_alpha, uint8_t *out,
                              size_t *out_len, oxo_ihdr *oh, oxo_aux *oaux); /* alpha.rs:35-108 */
int oxo_reduced_bit_depth_16_to_8(const oxo_ihdr *h, const uint8_t *data, size_t len, int force_scale, uint8_t *out,
                                  size_t *out_len, oxo_ihdr *oh); /* bit_depth.rs:9-64 */
int oxo_reduced_bit_depth_8_or_less(const oxo_ihdr *h, const oxo_aux *aux, const uint8_t *data, size_t len,
                                    uint8_t *out, size_t *out_len, oxo_ihdr *oh, oxo_aux *oaux); /* bit_depth.rs:68-166 */
int oxo_expanded_bit_depth_to_8(const oxo_ihdr *h, const oxo_aux *aux, const uint8_t *data, size_t len, uint8_t *out,
                                size_t *out_len, oxo_ihdr *oh, oxo_aux *oaux); /* bit_depth.rs:170-230 */
int oxo_reduced_to_indexed(const oxo_ihdr *h, const oxo_aux *aux, const uint8_t *data, size_t len,
                           int allow_grayscale, uint8_t *out, size_t *out_len, oxo_ihdr *oh,
                           oxo_aux *oaux); /* color.rs:18-97 */
int oxo_reduced_rgb_to_grayscale(const oxo_ihdr *h, const oxo_aux *aux, const uint8_t *data, size_t len, uint8_t *out,
                                 size_t *out_len, oxo_ihdr *oh, oxo_aux *oaux); /* color.rs:100-137 */
int oxo_indexed_to_channels(const oxo_ihdr *h, const oxo_aux *aux, const uint8_t *data, size_t len,
                            int allow_grayscale, int optimize_alpha, uint8_t *out, size_t *out_len, oxo_ihdr *oh,
                            oxo_aux *oaux); /* color.rs:141-206 */
int oxo_reduced_palette(const oxo_ihdr *h, const oxo_aux *aux, const uint8_t *data, size_t len, int optimize_alpha,
                        uint8_t *out, oxo_aux *oaux); /* palette.rs:12-73 */
int oxo_sorted_palette(const oxo_ihdr *h, const oxo_aux *aux, const uint8_t *data, size_t len, uint8_t *out,
                       oxo_aux *oaux); /* palette.rs:77-130,197-217 */
/* most_popular_edge_color: returns index or -1 for None (palette.rs:197-217) */
int oxo_most_popular_edge_color(const oxo_ihdr *h, size_t num_colors, const uint8_t *data, size_t len);
/* CoOccurrenceMatrix::build (palette.rs:556-589): matrix is n*n u32 */
int oxo_cooccurrence_build(const oxo_ihdr *h, uint32_t num_colors, const uint8_t *data, size_t len, uint32_t *matrix);
/* apply_palette_reorder (palette.rs:165-194): remapping[i] = old index placed at new index i */
int oxo_apply_palette_reorder(const oxo_ihdr *h, const oxo_aux *aux, const uint32_t *remapping, const uint8_t *data,
                              size_t len, uint8_t *out, oxo_aux *oaux);

#ifdef __cplusplus
}
#endif
#endif
