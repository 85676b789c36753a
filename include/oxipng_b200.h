/*
 * oxipng_b200.h -- C ABI of the B200 (sm_100a) filter-search + reduction-scan path of oxipng.
 *
 * This is the drop-in boundary: plain pointers and sizes, no CUDA or torch types.  The reference
 * (oxipng 10.1.1, Rust) has no FFI layer of its own; each entry point below replaces the body of one
 * Rust function on the hot path and cites it (file:line relative to the reference tree).  The Rust
 * `-sys` declarations and the patch to the reference are shown in INTEGRATION.md.
 *
 * Conventions
 *   - All buffers are caller-allocated and caller-owned; sizes are explicit.
 *   - Return value: OXI_OK (0) on success; OXI_NONE (1) from a reduction means the reference function
 *     would have returned `None`; negative = error (oxi_last_error() gives a thread-local message).
 *   - There is NO CPU fallback: if no CUDA device is usable every compute entry point fails with
 *     OXI_E_CUDA.  The geometry helpers (oxi_raw_data_size, ...) are pure host arithmetic.
 *   - Re-entrant: may be called concurrently from many host threads (the reference calls
 *     filter_image from rayon workers, src/evaluate.rs:141-153).
 *   - "_device" variants take device pointers (inputs already resident in HBM) and a cudaStream_t
 *     passed as void*; they enqueue work and return without synchronising.
 */
#ifndef OXIPNG_B200_H
#define OXIPNG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OXI_OK 0
#define OXI_NONE 1
#define OXI_E_ARG (-1)
#define OXI_E_CUDA (-2)
#define OXI_E_NOMEM (-3)
#define OXI_E_UNSUPPORTED (-4)
#define OXI_E_DATA (-5) /* PngError::InvalidData */

/* IhdrData, src/headers.rs:15-26; colour-type codes src/colors.rs:48-56; BitDepth src/colors.rs:96-107 */
typedef struct oxi_ihdr {
    uint32_t width, height;
    uint8_t color_type; /* 0 gray, 2 rgb, 3 indexed, 4 gray+alpha, 6 rgba */
    uint8_t bit_depth;  /* 1, 2, 4, 8, 16 */
    uint8_t interlaced; /* 0 / 1 */
} oxi_ihdr;

/* FilterStrategy, src/filters/strategies.rs:9-29 (Brute stays on the host and is not representable) */
enum oxi_strategy_kind {
    OXI_BASIC = 0,     /* Basic(RowFilter): basic_filter = 0..4 (RowFilter, src/filters/mod.rs:7-15) */
    OXI_MINSUM = 1,    /* strategies.rs:76-101  */
    OXI_ENTROPY = 2,   /* strategies.rs:105-149 */
    OXI_BIGRAMS = 3,   /* strategies.rs:153-192 */
    OXI_BIGENT = 4,    /* strategies.rs:195-227 */
    OXI_PREDEFINED = 5 /* Predefined(Vec<RowFilter>): rows beyond n_predefined use None (png/mod.rs:389) */
};
typedef struct oxi_strategy {
    uint8_t kind;
    uint8_t basic_filter;
    const uint8_t *predefined; /* host pointer, n_predefined filter ids */
    size_t n_predefined;
} oxi_strategy;

/* The data-carrying parts of ColorType (src/colors.rs:9-29) that reductions read or produce */
typedef struct oxi_color_aux {
    uint32_t n_palette;      /* Indexed { palette }.len() */
    uint8_t palette[256][4]; /* RGBA8 */
    int32_t has_trns;        /* transparent_shade / transparent_color is Some */
    uint16_t trns[3];        /* gray: [0]; rgb: r,g,b */
} oxi_color_aux;

/* flags for the filter entry points */
#define OXI_FLAG_FORCE_GENERIC 1u /* testing: route through the size-generic kernels */
#define OXI_FLAG_FORCE_FAST 2u    /* testing: fail with OXI_E_UNSUPPORTED instead of using the generic kernels */

/* ---- runtime ------------------------------------------------------------------------------ */
const char *oxi_last_error(void);
const char *oxi_version(void);
int oxi_device_count(void);
int oxi_set_device(int device); /* device used by the host-buffer entry points of the calling thread */
/* number of kernels launched by this library in this process since load (bench.py's gpu_launches) */
uint64_t oxi_kernel_launches(void);
/* name of the tier the dispatcher would pick: "fast" or "generic" (diagnostics / tests) */
const char *oxi_filter_tier(const oxi_ihdr *, size_t data_len, const oxi_strategy *, int optimize_alpha);

/* ---- geometry (host arithmetic) --------------------------------------------------------------- */
size_t oxi_bits_per_pixel(const oxi_ihdr *);                /* IhdrData::bpp, src/headers.rs:32 */
size_t oxi_raw_data_size(const oxi_ihdr *);                 /* IhdrData::raw_data_size, src/headers.rs:38-64 */
size_t oxi_filter_bpp(const oxi_ihdr *);                    /* src/png/mod.rs:289-302,361 */
size_t oxi_num_scanlines(const oxi_ihdr *, size_t data_len); /* count of PngImage::scan_lines(false), src/png/scan_lines.rs:82-167 */
/* row table of scan_lines(false): len[i] bytes, pass[i] (0 = None, else 1..7), pixels[i]; returns count */
size_t oxi_scan_lines(const oxi_ihdr *, size_t data_len, uint32_t *len, uint8_t *pass, uint32_t *pixels, size_t cap);

/* ---- filter search: PngImage::filter_image, src/png/mod.rs:355-431 --------------------------- */
/* out: data_len + num_scanlines bytes (= raw_data_size for consistent images);
 * filters_used: num_scanlines bytes (the filter applied to each row, also for Basic/Predefined);
 * *used_is_predefined: 1 when the reference returns Predefined(filters_used) (png/mod.rs:426-430). */
int oxi_filter_image(const oxi_ihdr *, const uint8_t *data, size_t data_len, const oxi_strategy *, int optimize_alpha,
                     uint8_t *out, uint8_t *filters_used, int *used_is_predefined);

/* The evaluator fan-out, src/evaluate.rs:141-153 (`filters.par_iter()` -> filter_image): k strategies on
 * one image in one call; the image is uploaded once. outs[j] / filters_used[j] as above, per strategy. */
int oxi_filter_image_multi(const oxi_ihdr *, const uint8_t *data, size_t data_len, const oxi_strategy *strategies,
                           size_t k, int optimize_alpha, uint8_t *const *outs, uint8_t *const *filters_used);

/* n_images images of identical IHDR, contiguous (image i at data + i*data_len); outs[j] receives
 * n_images * (data_len + rows) bytes, filters_used[j] n_images * rows bytes.  Host buffers: pinned memory is used in
 * place; ordinary pageable memory is staged through the library's own pinned buffers by several copy threads.  Uploads,
 * kernels and downloads are pipelined over chunks of the batch.  For n_images > 1, data_len must be exactly the
 * scanlines of one image (OXI_E_ARG otherwise). */
int oxi_filter_batch(const oxi_ihdr *, const uint8_t *data, size_t data_len, size_t n_images,
                     const oxi_strategy *strategies, size_t k, int optimize_alpha, uint8_t *const *outs,
                     uint8_t *const *filters_used, unsigned flags);

/* Device-resident variant: d_data, d_outs[j], d_filters_used[j] are device pointers on `device`;
 * strategies[].predefined stays a host pointer (copied). stream = cudaStream_t (may be NULL). Asynchronous. */
int oxi_filter_batch_device(int device, void *stream, const oxi_ihdr *, const uint8_t *d_data, size_t data_len,
                            size_t n_images, const oxi_strategy *strategies, size_t k, int optimize_alpha,
                            uint8_t *const *d_outs, uint8_t *const *d_filters_used, unsigned flags);

/* One large image split into row bands with a one-row halo, the bands filtered concurrently on `devices`
 * (png/mod.rs:375-378,422: with optimize_alpha off a row depends only on the previous raw row of its pass).  Host
 * buffers; one worker thread and stream per device; about bands_per_device bands each.  optimize_alpha with an alpha
 * channel makes rows serial: OXI_E_UNSUPPORTED. */
int oxi_filter_image_sharded(const oxi_ihdr *, const uint8_t *data, size_t data_len, const oxi_strategy *strategies, size_t k,
                             int optimize_alpha, const int *devices, int n_devices, int bands_per_device,
                             uint8_t *const *outs, uint8_t *const *filters_used, unsigned flags);

/* ---- reduction scans, src/reduction/{alpha,bit_depth,color,palette}.rs ----------------------- */
/* Every function: host buffers in, host buffers out; `out` must hold the stated worst case (the handle -> handle
 * variants further down allocate their result themselves).
 * out_ihdr / out_aux receive the IhdrData / ColorType payload of the returned image. */

/* cleaned_alpha_channel, alpha.rs:11-32; out: data_len */
int oxi_cleaned_alpha_channel(const oxi_ihdr *, const uint8_t *data, size_t data_len, uint8_t *out);
/* reduced_alpha_channel, alpha.rs:35-108; out: data_len */
int oxi_reduced_alpha_channel(const oxi_ihdr *, const uint8_t *data, size_t data_len, int optimize_alpha, uint8_t *out,
                              size_t *out_len, oxi_ihdr *out_ihdr, oxi_color_aux *out_aux);
/* reduced_bit_depth_16_to_8 / scaled_bit_depth_16_to_8, bit_depth.rs:9-64; out: data_len/2 */
int oxi_reduced_bit_depth_16_to_8(const oxi_ihdr *, const uint8_t *data, size_t data_len, int force_scale,
                                  uint8_t *out, size_t *out_len, oxi_ihdr *out_ihdr);
/* reduced_bit_depth_8_or_less, bit_depth.rs:68-166; out: data_len */
int oxi_reduced_bit_depth_8_or_less(const oxi_ihdr *, const oxi_color_aux *, const uint8_t *data, size_t data_len,
                                    uint8_t *out, size_t *out_len, oxi_ihdr *out_ihdr, oxi_color_aux *out_aux);
/* expanded_bit_depth_to_8, bit_depth.rs:170-230; out: width*height */
int oxi_expanded_bit_depth_to_8(const oxi_ihdr *, const oxi_color_aux *, const uint8_t *data, size_t data_len,
                                uint8_t *out, size_t *out_len, oxi_ihdr *out_ihdr, oxi_color_aux *out_aux);
/* reduced_to_indexed (+ build_palette), color.rs:18-97; out: data_len/channels */
int oxi_reduced_to_indexed(const oxi_ihdr *, const oxi_color_aux *, const uint8_t *data, size_t data_len,
                           int allow_grayscale, uint8_t *out, size_t *out_len, oxi_ihdr *out_ihdr,
                           oxi_color_aux *out_aux);
/* reduced_rgb_to_grayscale, color.rs:100-137; out: data_len */
int oxi_reduced_rgb_to_grayscale(const oxi_ihdr *, const oxi_color_aux *, const uint8_t *data, size_t data_len,
                                 uint8_t *out, size_t *out_len, oxi_ihdr *out_ihdr, oxi_color_aux *out_aux);
/* reduced_palette, palette.rs:12-73; out: data_len */
int oxi_reduced_palette(const oxi_ihdr *, const oxi_color_aux *, const uint8_t *data, size_t data_len,
                        int optimize_alpha, uint8_t *out, oxi_color_aux *out_aux);
/* sorted_palette (+ most_popular_edge_color), palette.rs:77-130,197-217; out: data_len */
int oxi_sorted_palette(const oxi_ihdr *, const oxi_color_aux *, const uint8_t *data, size_t data_len, uint8_t *out,
                       oxi_color_aux *out_aux);
/* the pixel scans of palette.rs exposed on their own */
int oxi_used_histogram256(const uint8_t *data, size_t data_len, uint64_t hist[256]);      /* palette.rs:20-23 (counts; used = count>0) */
int oxi_lut_remap(const uint8_t *data, size_t data_len, const uint8_t lut[256], uint8_t *out); /* palette.rs:43,121,183 */
int oxi_edge_color_counts(const oxi_ihdr *, const uint8_t *data, size_t data_len, uint32_t counts[256]); /* palette.rs:198-204 */
/* CoOccurrenceMatrix::build, palette.rs:556-589; matrix: num_colors*num_colors u32, symmetrised */
int oxi_cooccurrence_build(const oxi_ihdr *, uint32_t num_colors, const uint8_t *data, size_t data_len,
                           uint32_t *matrix);
/* apply_palette_reorder, palette.rs:165-194; remapping[i] = old index placed at new index i */
int oxi_apply_palette_reorder(const oxi_ihdr *, const oxi_color_aux *, const uint32_t *remapping, const uint8_t *data,
                              size_t data_len, uint8_t *out, oxi_color_aux *out_aux);

/* ---- "next" rows of the scope table: the steps either side of the hot path ---------------------- */
/* PngImage::unfilter_image, src/png/mod.rs:335-351 (+ RowFilter::unfilter_line, src/filters/mod.rs:188-239): the
 * inflated IDAT stream (one filter-type byte in front of every scanline) -> PngImage.data. out: len bytes suffice.
 * Returns OXI_E_DATA for a filter type > 4 (the reference's PngError::InvalidData). */
int oxi_unfilter_image(const oxi_ihdr *, const uint8_t *filtered, size_t len, uint8_t *out, size_t *out_len);
/* n_images inflated streams of identical IHDR and length `len`, back to back (what a decoder front end hands over for a
 * directory of equally sized images) -> n_images x *out_len bytes of pixels, back to back.  One launch: groups of CTAs
 * walk the images, so the device is filled even though one image's rows form a dependency chain. */
int oxi_unfilter_batch(const oxi_ihdr *, const uint8_t *filtered, size_t len, size_t n_images, uint8_t *out, size_t *out_len);
/* the same with the inflated stream already in HBM: d_filtered (len bytes) and d_out are device pointers on `device`,
 * the work is enqueued on `stream`; only the few-byte "invalid filter type" flag is read back (synchronises).
 * Both pointers must be 16-byte aligned (OXI_E_ARG otherwise); the stream is read in aligned 4-byte words, i.e. up to
 * 3 bytes past `len` inside the word that holds the last byte.  Scanlines of 16 MB and more: OXI_E_UNSUPPORTED. */
int oxi_unfilter_image_device(int device, void *stream, const oxi_ihdr *, const uint8_t *d_filtered, size_t len,
                              uint8_t *d_out, size_t *out_len);
int oxi_unfilter_batch_device(int device, void *stream, const oxi_ihdr *, const uint8_t *d_filtered, size_t len, size_t n_images,
                              uint8_t *d_out, size_t *out_len);
/* interlace_image, src/interlace.rs:6-60: progressive layout -> Adam7 layout (ihdr describes the image; its
 * `interlaced` flag is ignored). out: oxi_raw_data_size(interlaced ihdr) - rows bytes at most (<= data_len + 7*height). */
int oxi_interlace_image(const oxi_ihdr *, const uint8_t *data, size_t data_len, uint8_t *out, size_t *out_len);
/* deinterlace_image, src/interlace.rs:62-150: Adam7 layout -> progressive layout. out: ceil(width*bpp/8)*height */
int oxi_deinterlace_image(const oxi_ihdr *, const uint8_t *data, size_t data_len, uint8_t *out, size_t *out_len);

/* ---- device-resident images (oxi_image_upload / oxi_image_free; 8f row 3) -------------------------
 * An oxi_image owns PngImage.data in HBM with its IhdrData and palette / tRNS.  Every call below runs on the image's
 * stream (cudaStream_t as void*, may be NULL) and derived images inherit it, so perform_reductions
 * (src/reduction/mod.rs:14-179) and the evaluator (src/evaluate.rs:126-201) chain on one stream: pixels cross PCIe
 * once in (upload) and only the filtered streams come back.  A reduction returns OXI_NONE exactly where the reference
 * function returns None (*out is then NULL); the scan results that decide this (a few bytes) are read back
 * synchronously, the images themselves never leave the device.  Free every image with oxi_image_free. */
typedef struct oxi_image oxi_image;
int oxi_image_upload(int device, void *stream, const oxi_ihdr *, const oxi_color_aux * /* may be NULL */, const uint8_t *data,
                     size_t len, oxi_image **out);
/* the same from bytes already on the device (copied, stream ordered) */
int oxi_image_wrap_device(int device, void *stream, const oxi_ihdr *, const oxi_color_aux *, const uint8_t *d_data, size_t len,
                          oxi_image **out);
/* PngImage::new -> unfilter_image (png/mod.rs:335-351): upload the inflated IDAT stream and unfilter it on the device */
int oxi_image_from_filtered(int device, void *stream, const oxi_ihdr *, const oxi_color_aux *, const uint8_t *filtered,
                            size_t len, oxi_image **out);
void oxi_image_free(oxi_image *);
int oxi_image_info(const oxi_image *, oxi_ihdr *, oxi_color_aux *, size_t *len);
const uint8_t *oxi_image_device_ptr(const oxi_image *);
int oxi_image_download(const oxi_image *, uint8_t *out, size_t cap, size_t *len); /* synchronises */
int oxi_image_sync(const oxi_image *);
/* reductions, handle -> handle: same semantics as the host-buffer entry points above */
int oxi_cleaned_alpha_channel_device(const oxi_image *, oxi_image **out);
int oxi_reduced_alpha_channel_device(const oxi_image *, int optimize_alpha, oxi_image **out);
int oxi_reduced_bit_depth_16_to_8_device(const oxi_image *, int force_scale, oxi_image **out);
int oxi_reduced_bit_depth_8_or_less_device(const oxi_image *, oxi_image **out);
int oxi_expanded_bit_depth_to_8_device(const oxi_image *, oxi_image **out);
int oxi_reduced_to_indexed_device(const oxi_image *, int allow_grayscale, oxi_image **out);
int oxi_reduced_rgb_to_grayscale_device(const oxi_image *, oxi_image **out);
int oxi_reduced_palette_device(const oxi_image *, int optimize_alpha, oxi_image **out);
int oxi_sorted_palette_device(const oxi_image *, oxi_image **out);
int oxi_apply_palette_reorder_device(const oxi_image *, const uint32_t *remapping, oxi_image **out);
int oxi_cooccurrence_build_device(const oxi_image *, uint32_t num_colors, uint32_t *matrix /* host */);
int oxi_used_histogram256_device(const oxi_image *, uint64_t hist[256] /* host */);
int oxi_interlace_image_device(const oxi_image *, oxi_image **out);
int oxi_deinterlace_image_device(const oxi_image *, oxi_image **out);
/* evaluator candidate (evaluate.rs:141-153): k strategies on the resident image; only the filtered streams and the
 * filters used come back to host memory (outs[j]: len + rows bytes, filters_used[j]: rows bytes) */
int oxi_image_filter(const oxi_image *, const oxi_strategy *strategies, size_t k, int optimize_alpha, uint8_t *const *outs,
                     uint8_t *const *filters_used, unsigned flags);
/* the same into device buffers, asynchronous */
int oxi_image_filter_device(const oxi_image *, const oxi_strategy *strategies, size_t k, int optimize_alpha,
                            uint8_t *const *d_outs, uint8_t *const *d_filters_used, unsigned flags);

#ifdef __cplusplus
}
#endif
#endif /* OXIPNG_B200_H */
