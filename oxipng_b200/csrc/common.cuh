// common.cuh -- shared host/device definitions of the B200 filter-search path.
// Domain vocabulary follows the reference: scanline (row), pass (Adam7), filter (RowFilter 0..4), strategy.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/oxipng_b200.h"

namespace oxi {

// One Adam7 pass (or the whole image when not interlaced): `nrows` scanlines of `len` bytes each, stored
// back to back from byte `in_base` of PngImage::data.  Row r of the image (global scanline index) in this
// pass has in_off = in_base + (r-row0)*len and out_off = in_off + r (one filter byte per preceding row),
// which is how filter_image lays out its output (src/png/mod.rs:374-424).
struct PassDesc {
    uint32_t row0;
    uint32_t nrows;
    uint32_t len;
    uint32_t pixels; // pixels per scanline of this pass (ScanLine::num_pixels)
    uint64_t in_base;
};

// Geometry of one image as the ScanLines iterator (src/png/scan_lines.rs:82-167) walks it.
struct Geom {
    uint32_t n_pass;      // 1 when not interlaced, else number of non-empty passes (<= 7)
    uint32_t bpp;         // filter bytes per pixel (src/png/mod.rs:361)
    uint32_t total_rows;  // S
    uint32_t max_len;     // longest scanline in bytes
    uint32_t alpha_bytes; // src/png/mod.rs:363-367 (0 when optimize_alpha is off or no alpha channel)
    uint32_t color_bytes; // bpp - alpha_bytes
    uint64_t data_len;    // N (bytes consumed by the iterator; trailing bytes are ignored like the reference)
    uint64_t out_len;     // data_len + total_rows
    PassDesc pass[7];
};

// A strategy as the kernels see it.
struct DevStrategy {
    int kind;                  // oxi_strategy_kind
    int basic;                 // Basic(filter)
    const uint8_t *predefined; // device pointer (n_predefined entries) or nullptr
    uint32_t n_predefined;
    uint8_t *out;          // n_images * out_len
    uint8_t *filters_used; // n_images * total_rows
};

constexpr int MAX_STRATEGIES = 12; // -o6 evaluates 9 in-scope strategies at once (options.rs:250-273)

struct StrategySet {
    int k;
    DevStrategy s[MAX_STRATEGIES];
};

#ifdef __CUDACC__

__host__ __device__ __forceinline__ int find_pass(const Geom &g, uint32_t row) {
    int p = 0;
#pragma unroll
    for (int i = 1; i < 7; i++)
        if (i < (int)g.n_pass && row >= g.pass[i].row0) p = i;
    return p;
}

// paeth_predictor, src/filters/mod.rs:242-254
__device__ __forceinline__ uint32_t paeth(uint32_t a, uint32_t b, uint32_t c) {
    int p = (int)a + (int)b - (int)c;
    int pa = abs(p - (int)a), pb = abs(p - (int)b), pc = abs(p - (int)c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// ---- byte-SIMD Paeth (shared by the filter search and unfilter_image) ------------------------------------------
__device__ __forceinline__ uint32_t signmask4(uint32_t x) { // 0xFF in every byte whose MSB is set
    uint32_t d;                                             // (prmt's sign-replicate selectors; __byte_perm masks them off)
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(0u), "r"(0xba98u));
    return d;
}
__device__ __forceinline__ uint32_t sel4(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }

// paeth_predictor (src/filters/mod.rs:242-254) on four byte lanes, without widening.
// a=left, b=up, c=upleft.  pa=|b-c|, pb=|a-c|.  pc=|a+b-2c| equals |pa-pb| when c lies between a and b and pa+pb
// (>= both) otherwise; "otherwise" is detected as |pa-pb| == |a-b| and pc saturated to 255, which keeps both
// comparisons against pc true.  Then the winner is the closer of a/b (ties -> a) if its distance <= pc, else c.
// Checked exhaustively over all 2^24 (a,b,c) in every byte lane on the device (tests/cuda/paeth4_exhaustive.cu).
__device__ __forceinline__ uint32_t paeth4(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t pa = __vabsdiffu4(b, c), pb = __vabsdiffu4(a, c);
    const uint32_t pd = __vabsdiffu4(pa, pb);
    const uint32_t pc = pd | __vcmpeq4(pd, __vabsdiffu4(a, b));
    const uint32_t le = __vcmpleu4(pa, pb);
    const uint32_t w = sel4(le, a, b), mn = sel4(le, pa, pb);
    // mn <= pc: mn < 128 whenever pc < 255, so (pc|0x80)-mn never borrows and its MSB decides for pc < 128
    const uint32_t t = (pc | 0x80808080u) - mn;
    return sel4(signmask4(t | pc), w, c);
}

// ilog2i, src/filters/strategies.rs:269-272 (i >= 1)
__device__ __forceinline__ uint32_t ilog2i(uint32_t i) {
    // i * lg + ((i - 2^lg) << 1), written as one multiply-add: i * (lg + 2) - 2^(lg + 1)
    const uint32_t lg = 31u - (uint32_t)__clz((int)i);
    return i * (lg + 2u) - (2u << lg);
}

// residual of one byte under filter f (src/filters/mod.rs:63-110). left/upleft are 0 for the first bpp bytes,
// which reproduces the reference's special cases (Sub: raw, Average: up>>1, Paeth: up).
__device__ __forceinline__ uint32_t residual_byte(int f, uint32_t cur, uint32_t left, uint32_t up, uint32_t upleft) {
    uint32_t pred;
    switch (f) {
    case 0: pred = 0; break;
    case 1: pred = left; break;
    case 2: pred = up; break;
    case 3: pred = (left + up) >> 1; break;
    default: pred = paeth(left, up, upleft); break;
    }
    return (cur - pred) & 0xFFu;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); } // REDUX.SUM
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif // __CUDACC__

// ---- host-side interfaces between translation units -------------------------------------------------

struct LaunchCtx {
    int device;
    cudaStream_t stream;
    int num_sms;
    uint32_t smem_window; // DeviceCtx::smem_window
    // scratch provided by the caller (capi.cu) for the generic tier
    uint8_t *scratch;
    size_t scratch_bytes;
};

// generic tier (filter_generic.cu): any geometry, any strategy, alpha optimisation included.
size_t generic_scratch_bytes(const Geom &g, size_t n_images, int num_sms, int *grid_out);
cudaError_t launch_generic(const LaunchCtx &ctx, const Geom &g, const uint8_t *d_data, size_t n_images,
                           const StrategySet &set);

// fast tier (filter_fast.cu): alpha_bytes == 0, rows that fit the shared-memory staging ring.
bool fast_supported(const Geom &g, const StrategySet &set);
// once per device: raises the dynamic shared-memory limit of every k_fast instantiation to the full 227 KB (so no
// launch ever changes a function attribute) and probes the shared-window address of dynamic shared memory
cudaError_t fast_tier_init(int device, uint32_t *smem_window);
bool fast_tier_usable(uint32_t smem_window);
cudaError_t launch_fast(const LaunchCtx &ctx, const Geom &g, const uint8_t *d_data, size_t n_images,
                        const StrategySet &set);

// optimize_alpha tier (filter_fast.cu, k_alpha): rows serial per (image, strategy) chain, everything staged in shared memory
bool alpha_supported(const Geom &g, const StrategySet &set);
cudaError_t launch_alpha(const LaunchCtx &ctx, const Geom &g, const uint8_t *d_data, size_t n_images,
                         const StrategySet &set);

void count_launch(uint64_t n = 1);

} // namespace oxi
