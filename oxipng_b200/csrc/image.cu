// image.cu -- device-resident images (SURVEY.md 8b sketch, 8f row 3): an opaque handle owns the PngImage.data bytes in
// HBM together with its IhdrData / palette / tRNS, and every reduction, interlace / deinterlace, unfilter and the filter
// search run handle -> handle on the image's stream, so that perform_reductions (src/reduction/mod.rs:14-179) and the
// evaluator (src/evaluate.rs:126-201) move the pixels over PCIe once in and the chosen filtered streams once out.
// The handle entry points are thin: they allocate the result in stream order and run the very same code as the
// host-buffer entry points in device-resident mode (host.h DevIO), so there is one implementation of every reduction.
#include <cstring>

#include "common.cuh"
#include "host.h"

struct oxi_image {
    int device;
    cudaStream_t stream;
    oxi_ihdr ihdr;
    oxi_color_aux aux;
    uint8_t *d;
    size_t len, cap;
};

namespace oxi {

namespace {

int bind(int device, DeviceCtx **c) {
    int err = 0;
    *c = get_ctx(device, &err);
    if (!*c) return err;
    cudaError_t e = cudaSetDevice(device);
    return e == cudaSuccess ? OXI_OK : fail_cuda(e, "cudaSetDevice");
}

// stream-ordered allocation from the device's default pool (its release threshold is raised once per device, so the
// buffers of a reduction chain are recycled without returning to the driver)
int new_image(int device, cudaStream_t stream, size_t cap, oxi_image **out) {
    oxi_image *im = new oxi_image;
    memset(im, 0, sizeof *im);
    im->device = device;
    im->stream = stream;
    im->cap = cap ? cap : 1;
    cudaError_t e = cudaMallocAsync((void **)&im->d, im->cap + 64, stream);
    if (e != cudaSuccess) { delete im; return fail_cuda(e, "cudaMallocAsync"); }
    *out = im;
    return OXI_OK;
}
void drop(oxi_image *im) {
    if (!im) return;
    cudaSetDevice(im->device);
    cudaFreeAsync(im->d, im->stream);
    delete im;
}

// Runs `body(in->d, in->len, out->d)` in device-resident mode with a result buffer of `cap` bytes; on OXI_OK the result
// handle (header / aux / length filled by `body` through the references) is returned, otherwise it is released.
template <typename F>
int derive(const oxi_image *in, size_t cap, oxi_image **out, F body) {
    if (!in || !out) return fail(OXI_E_ARG, "null image");
    *out = nullptr;
    DeviceCtx *c;
    if (int rc = bind(in->device, &c)) return rc;
    oxi_image *res;
    if (int rc = new_image(in->device, in->stream, cap, &res)) return rc;
    res->ihdr = in->ihdr;
    res->aux = in->aux;
    res->len = in->len;
    const DevIO io{in->device, in->stream};
    int rc;
    {
        DevIOScope scope(&io);
        rc = body(res);
    }
    if (rc != OXI_OK) { drop(res); return rc; }
    *out = res;
    return OXI_OK;
}

} // namespace
} // namespace oxi

using namespace oxi;

extern "C" {

int oxi_image_upload(int device, void *stream, const oxi_ihdr *h, const oxi_color_aux *aux, const uint8_t *data, size_t len,
                     oxi_image **out) {
    if (!valid_ihdr(h) || (!data && len) || !out) return fail(OXI_E_ARG, "bad arguments");
    DeviceCtx *c;
    if (int rc = bind(device, &c)) return rc;
    oxi_image *im;
    if (int rc = new_image(device, (cudaStream_t)stream, len, &im)) return rc;
    im->ihdr = *h;
    if (aux) im->aux = *aux;
    im->len = len;
    cudaError_t e = len ? cudaMemcpyAsync(im->d, data, len, cudaMemcpyHostToDevice, im->stream) : cudaSuccess;
    if (e != cudaSuccess) { drop(im); return fail_cuda(e, "cudaMemcpyAsync(H2D)"); }
    *out = im;
    return OXI_OK;
}

int oxi_image_wrap_device(int device, void *stream, const oxi_ihdr *h, const oxi_color_aux *aux, const uint8_t *d_data,
                          size_t len, oxi_image **out) {
    if (!valid_ihdr(h) || (!d_data && len) || !out) return fail(OXI_E_ARG, "bad arguments");
    DeviceCtx *c;
    if (int rc = bind(device, &c)) return rc;
    oxi_image *im;
    if (int rc = new_image(device, (cudaStream_t)stream, len, &im)) return rc;
    im->ihdr = *h;
    if (aux) im->aux = *aux;
    im->len = len;
    cudaError_t e = len ? cudaMemcpyAsync(im->d, d_data, len, cudaMemcpyDeviceToDevice, im->stream) : cudaSuccess;
    if (e != cudaSuccess) { drop(im); return fail_cuda(e, "cudaMemcpyAsync(D2D)"); }
    *out = im;
    return OXI_OK;
}

void oxi_image_free(oxi_image *im) { drop(im); }

int oxi_image_info(const oxi_image *im, oxi_ihdr *h, oxi_color_aux *aux, size_t *len) {
    if (!im) return fail(OXI_E_ARG, "null image");
    if (h) *h = im->ihdr;
    if (aux) *aux = im->aux;
    if (len) *len = im->len;
    return OXI_OK;
}

const uint8_t *oxi_image_device_ptr(const oxi_image *im) { return im ? im->d : nullptr; }

int oxi_image_download(const oxi_image *im, uint8_t *out, size_t cap, size_t *len) {
    if (!im || (!out && im->len)) return fail(OXI_E_ARG, "bad arguments");
    if (cap < im->len) return fail(OXI_E_ARG, "output buffer too small");
    DeviceCtx *c;
    if (int rc = bind(im->device, &c)) return rc;
    cudaError_t e = im->len ? cudaMemcpyAsync(out, im->d, im->len, cudaMemcpyDeviceToHost, im->stream) : cudaSuccess;
    if (e == cudaSuccess) e = cudaStreamSynchronize(im->stream);
    if (e != cudaSuccess) return fail_cuda(e, "download");
    if (len) *len = im->len;
    return OXI_OK;
}

int oxi_image_sync(const oxi_image *im) {
    if (!im) return fail(OXI_E_ARG, "null image");
    DeviceCtx *c;
    if (int rc = bind(im->device, &c)) return rc;
    cudaError_t e = cudaStreamSynchronize(im->stream);
    return e == cudaSuccess ? OXI_OK : fail_cuda(e, "cudaStreamSynchronize");
}

// ---- reductions, handle -> handle ---------------------------------------------------------------------------------

int oxi_cleaned_alpha_channel_device(const oxi_image *in, oxi_image **out) {
    return derive(in, in ? in->len : 0, out, [&](oxi_image *r) {
        const size_t bpp = oxi_filter_bpp(&in->ihdr);
        r->len = bpp ? in->len / bpp * bpp : 0; // chunks_exact (alpha.rs:17)
        return oxi_cleaned_alpha_channel(&in->ihdr, in->d, in->len, r->d);
    });
}
int oxi_reduced_alpha_channel_device(const oxi_image *in, int optimize_alpha, oxi_image **out) {
    return derive(in, in ? in->len : 0, out, [&](oxi_image *r) {
        return oxi_reduced_alpha_channel(&in->ihdr, in->d, in->len, optimize_alpha, r->d, &r->len, &r->ihdr, &r->aux);
    });
}
int oxi_reduced_bit_depth_16_to_8_device(const oxi_image *in, int force_scale, oxi_image **out) {
    return derive(in, in ? in->len / 2 + 1 : 0, out, [&](oxi_image *r) {
        // the colour type (with its tRNS value) is cloned unchanged, bit_depth.rs:24-31,56-63
        return oxi_reduced_bit_depth_16_to_8(&in->ihdr, in->d, in->len, force_scale, r->d, &r->len, &r->ihdr);
    });
}
int oxi_reduced_bit_depth_8_or_less_device(const oxi_image *in, oxi_image **out) {
    return derive(in, in ? in->len : 0, out, [&](oxi_image *r) {
        return oxi_reduced_bit_depth_8_or_less(&in->ihdr, &in->aux, in->d, in->len, r->d, &r->len, &r->ihdr, &r->aux);
    });
}
int oxi_expanded_bit_depth_to_8_device(const oxi_image *in, oxi_image **out) {
    const size_t cap = in ? (size_t)in->ihdr.width * in->ihdr.height + in->len * 8 : 0;
    return derive(in, cap, out, [&](oxi_image *r) {
        return oxi_expanded_bit_depth_to_8(&in->ihdr, &in->aux, in->d, in->len, r->d, &r->len, &r->ihdr, &r->aux);
    });
}
int oxi_reduced_to_indexed_device(const oxi_image *in, int allow_grayscale, oxi_image **out) {
    return derive(in, in ? in->len : 0, out, [&](oxi_image *r) {
        return oxi_reduced_to_indexed(&in->ihdr, &in->aux, in->d, in->len, allow_grayscale, r->d, &r->len, &r->ihdr, &r->aux);
    });
}
int oxi_reduced_rgb_to_grayscale_device(const oxi_image *in, oxi_image **out) {
    return derive(in, in ? in->len : 0, out, [&](oxi_image *r) {
        return oxi_reduced_rgb_to_grayscale(&in->ihdr, &in->aux, in->d, in->len, r->d, &r->len, &r->ihdr, &r->aux);
    });
}
int oxi_reduced_palette_device(const oxi_image *in, int optimize_alpha, oxi_image **out) {
    return derive(in, in ? in->len : 0, out, [&](oxi_image *r) {
        return oxi_reduced_palette(&in->ihdr, &in->aux, in->d, in->len, optimize_alpha, r->d, &r->aux);
    });
}
int oxi_sorted_palette_device(const oxi_image *in, oxi_image **out) {
    return derive(in, in ? in->len : 0, out, [&](oxi_image *r) {
        return oxi_sorted_palette(&in->ihdr, &in->aux, in->d, in->len, r->d, &r->aux);
    });
}
int oxi_apply_palette_reorder_device(const oxi_image *in, const uint32_t *remapping, oxi_image **out) {
    return derive(in, in ? in->len : 0, out, [&](oxi_image *r) {
        return oxi_apply_palette_reorder(&in->ihdr, &in->aux, remapping, in->d, in->len, r->d, &r->aux);
    });
}
int oxi_cooccurrence_build_device(const oxi_image *in, uint32_t num_colors, uint32_t *matrix) {
    if (!in) return fail(OXI_E_ARG, "null image");
    const DevIO io{in->device, in->stream};
    DevIOScope scope(&io);
    return oxi_cooccurrence_build(&in->ihdr, num_colors, in->d, in->len, matrix);
}
int oxi_used_histogram256_device(const oxi_image *in, uint64_t hist[256]) {
    if (!in) return fail(OXI_E_ARG, "null image");
    const DevIO io{in->device, in->stream};
    DevIOScope scope(&io);
    return oxi_used_histogram256(in->d, in->len, hist);
}
int oxi_interlace_image_device(const oxi_image *in, oxi_image **out) {
    const size_t cap = in ? in->len + 8 * (size_t)in->ihdr.height + 64 : 0;
    return derive(in, cap, out, [&](oxi_image *r) {
        r->ihdr.interlaced = 1;
        return oxi_interlace_image(&in->ihdr, in->d, in->len, r->d, &r->len);
    });
}
int oxi_deinterlace_image_device(const oxi_image *in, oxi_image **out) {
    const size_t cap = in ? ((oxi_bits_per_pixel(&in->ihdr) * in->ihdr.width + 7) / 8) * in->ihdr.height : 0;
    return derive(in, cap, out, [&](oxi_image *r) {
        r->ihdr.interlaced = 0;
        return oxi_deinterlace_image(&in->ihdr, in->d, in->len, r->d, &r->len);
    });
}

// PngImage::unfilter_image (png/mod.rs:335-351) straight into a resident image: the inflated IDAT stream is uploaded
// and unfiltered on the device; this is how an image enters the resident pipeline after the host's inflate.
int oxi_image_from_filtered(int device, void *stream, const oxi_ihdr *h, const oxi_color_aux *aux, const uint8_t *filtered,
                            size_t len, oxi_image **out) {
    if (!out) return fail(OXI_E_ARG, "bad arguments");
    oxi_image *up = nullptr;
    if (int rc = oxi_image_upload(device, stream, h, aux, filtered, len, &up)) return rc;
    const int rc = derive(up, len, out, [&](oxi_image *r) { return oxi_unfilter_image(h, up->d, len, r->d, &r->len); });
    drop(up);
    return rc;
}

int oxi_unfilter_image_device(int device, void *stream, const oxi_ihdr *h, const uint8_t *d_filtered, size_t len, uint8_t *d_out,
                              size_t *out_len) {
    if (!d_filtered || !d_out) return fail(OXI_E_ARG, "bad arguments");
    DeviceCtx *c;
    if (int rc = bind(device, &c)) return rc;
    const DevIO io{device, (cudaStream_t)stream};
    DevIOScope scope(&io);
    return oxi_unfilter_image(h, d_filtered, len, d_out, out_len);
}

int oxi_unfilter_batch_device(int device, void *stream, const oxi_ihdr *h, const uint8_t *d_filtered, size_t len, size_t n_images,
                              uint8_t *d_out, size_t *out_len) {
    if (!d_filtered || !d_out) return fail(OXI_E_ARG, "bad arguments");
    DeviceCtx *c;
    if (int rc = bind(device, &c)) return rc;
    const DevIO io{device, (cudaStream_t)stream};
    DevIOScope scope(&io);
    return oxi_unfilter_batch(h, d_filtered, len, n_images, d_out, out_len);
}

// ---- the filter search on a resident image ---------------------------------------------------------------------------

int oxi_image_filter_device(const oxi_image *im, const oxi_strategy *strategies, size_t k, int optimize_alpha,
                            uint8_t *const *d_outs, uint8_t *const *d_filters_used, unsigned flags) {
    if (!im) return fail(OXI_E_ARG, "null image");
    return filter_resident(im->device, im->stream, &im->ihdr, im->d, im->len, 1, strategies, k, optimize_alpha, d_outs,
                           d_filters_used, flags);
}

// Evaluator candidate (evaluate.rs:141-153): k strategies on the resident image, only the filtered streams (and the
// filters used) cross PCIe.  outs[j]: len + rows bytes of host memory, filters_used[j]: rows bytes (either may be null).
int oxi_image_filter(const oxi_image *im, const oxi_strategy *strategies, size_t k, int optimize_alpha, uint8_t *const *outs,
                     uint8_t *const *filters_used, unsigned flags) {
    if (!im || !strategies || k == 0 || k > (size_t)MAX_STRATEGIES) return fail(OXI_E_ARG, "bad arguments");
    DeviceCtx *c;
    if (int rc = bind(im->device, &c)) return rc;
    Geom g;
    if (!build_geom(&im->ihdr, im->len, optimize_alpha, &g)) return fail(OXI_E_ARG, "invalid IHDR");
    const size_t out_stride = ((size_t)g.out_len + 255) & ~(size_t)255, used_stride = ((size_t)g.total_rows + 255) & ~(size_t)255;
    uint8_t *buf = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&buf, k * (out_stride + used_stride) + 256, im->stream);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMallocAsync");
    uint8_t *d_outs[MAX_STRATEGIES], *d_used[MAX_STRATEGIES];
    for (size_t j = 0; j < k; j++) {
        d_outs[j] = buf + j * (out_stride + used_stride);
        d_used[j] = d_outs[j] + out_stride;
    }
    int rc = filter_resident(im->device, im->stream, &im->ihdr, im->d, g.data_len, 1, strategies, k, optimize_alpha, d_outs,
                             d_used, flags);
    for (size_t j = 0; j < k && rc == OXI_OK; j++) {
        if (outs && outs[j]) e = cudaMemcpyAsync(outs[j], d_outs[j], g.out_len, cudaMemcpyDeviceToHost, im->stream);
        if (e == cudaSuccess && filters_used && filters_used[j])
            e = cudaMemcpyAsync(filters_used[j], d_used[j], g.total_rows, cudaMemcpyDeviceToHost, im->stream);
        if (e != cudaSuccess) rc = fail_cuda(e, "cudaMemcpyAsync(D2H)");
    }
    cudaFreeAsync(buf, im->stream);
    e = cudaStreamSynchronize(im->stream);
    if (e != cudaSuccess && rc == OXI_OK) rc = fail_cuda(e, "cudaStreamSynchronize");
    return rc;
}

} // extern "C"
