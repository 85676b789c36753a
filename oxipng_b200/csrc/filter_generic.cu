// filter_generic.cu -- size-generic tier of PngImage::filter_image (src/png/mod.rs:355-431).
//
// Handles every geometry (any row length, packed sub-byte depths, 16-bit, Adam7 passes), every in-scope
// strategy (Basic / Predefined / MinSum / Entropy / Bigrams / BigEnt) and the optimize_alpha pre-pass
// (src/filters/mod.rs:114-186).  It trades speed for generality: rows are read straight from global memory,
// each trial's filtered line is materialised in a per-CTA scratch area and scored from there, and the bigram
// scorers use a 65536-entry u32 table in global memory (L2 resident).  The fast tier (filter_fast.cu) covers the
// benchmark configurations; this tier is the catch-all and the only one that runs optimize_alpha, whose rows are
// serially dependent (row r+1 filters against the rewritten row r chosen by the winning trial, png/mod.rs:412-423).
#include "common.cuh"

namespace oxi {

namespace {

constexpr int GT = 256; // threads per CTA

struct CtaScratch {
    uint32_t *table;    // 65536 u32 bigram counters (kept all-zero between trials)
    uint8_t *trial;     // max_len + 1  (filter byte + residuals)
    uint8_t *best_line; // max_len + 1  (alpha only)
    uint8_t *line;      // max_len      (alpha only: mutable copy of the current row)
    uint8_t *prev;      // max_len      (alpha only: previous row as left by the winning trial)
    uint8_t *best_raw;  // max_len      (alpha only)
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ __device__ inline size_t scratch_per_cta(const Geom &g) {
    return 65536 * sizeof(uint32_t) + 2 * align256((size_t)g.max_len + 1) + 3 * align256(g.max_len);
}

__device__ inline CtaScratch carve(uint8_t *base, const Geom &g) {
    CtaScratch s;
    s.table = reinterpret_cast<uint32_t *>(base);
    base += 65536 * sizeof(uint32_t);
    s.trial = base;
    base += align256((size_t)g.max_len + 1);
    s.best_line = base;
    base += align256((size_t)g.max_len + 1);
    s.line = base;
    base += align256(g.max_len);
    s.prev = base;
    base += align256(g.max_len);
    s.best_raw = base;
    return s;
}

__device__ unsigned long long block_sum64(unsigned long long v, unsigned long long *sh) {
    v = warp_sum64(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
    for (int w = 0; w < GT / 32; w++) t += sh[w];
    return t;
}

// Score of one materialised filtered line t[0..n) under the evaluator `kind`; all threads get the result.
// MinSum/Bigrams: lower is better (usize); Entropy/BigEnt: higher is better (u32 fold cast to i32).
__device__ unsigned long long score_line(int kind, const uint8_t *t, uint32_t n, uint32_t *table, uint32_t *hist,
                                         unsigned long long *sh) {
    const int tid = threadIdx.x;
    unsigned long long acc = 0;
    if (kind == OXI_MINSUM) { // strategies.rs:90-94
        for (uint32_t i = tid; i < n; i += GT) {
            int v = (int8_t)t[i];
            acc += (unsigned)(v < 0 ? -v : v);
        }
        return block_sum64(acc, sh);
    }
    if (kind == OXI_ENTROPY) { // strategies.rs:119-142
        hist[tid] = 0;
        __syncthreads();
        for (uint32_t i = tid; i < n; i += GT) atomicAdd(&hist[t[i]], 1u);
        __syncthreads();
        uint32_t c = hist[tid];
        acc = c ? ilog2i(c) : 0;
        return (uint32_t)block_sum64(acc, sh); // wrapping u32 fold
    }
    // Bigrams (strategies.rs:178-188) and BigEnt (:211-220): count per distinct bigram over windows(2)
    for (uint32_t i = tid; i + 1 < n; i += GT) atomicAdd(&table[((uint32_t)t[i] << 8) | t[i + 1]], 1u);
    __syncthreads();
    for (uint32_t i = tid; i + 1 < n; i += GT) {
        uint32_t c = atomicExch(&table[((uint32_t)t[i] << 8) | t[i + 1]], 0u); // first taker owns the count, clears
        if (c) acc += kind == OXI_BIGRAMS ? 1u : ilog2i(c);
    }
    unsigned long long r = block_sum64(acc, sh);
    return kind == OXI_BIGRAMS ? r : (unsigned long long)(uint32_t)r;
}

__device__ inline bool better(int kind, unsigned long long s, unsigned long long best, bool have_best) {
    if (!have_best) return true; // best = usize::MAX / i32::MIN: the first trial (None) always wins
    if (kind == OXI_MINSUM || kind == OXI_BIGRAMS) return s < best;
    return (int32_t)(uint32_t)s > (int32_t)(uint32_t)best;
}

__device__ inline bool px_transparent(const uint8_t *px, uint32_t cb, uint32_t bpp) {
    for (uint32_t k = cb; k < bpp; k++)
        if (px[k]) return false;
    return true;
}

// RowFilter::optimize_alpha, src/filters/mod.rs:114-186, for one row, cooperatively.
// Alpha bytes are never written, so "transparent" is a fixed property of each pixel; every maximal run of
// transparent pixels is rewritten left to right by the thread that finds its first pixel (runs are disjoint,
// and a run only reads the opaque pixel before it, the previous row, and its own earlier writes).
__device__ void optimize_alpha_row(int f, uint8_t *line, const uint8_t *prev, uint32_t len, uint32_t bpp, uint32_t cb,
                                   uint32_t *sh_first) {
    if (f == 0) return;
    const uint32_t npix = len / bpp;
    const int tid = threadIdx.x;
    if (f == 2) { // Up :157-159
        for (uint32_t i = tid; i < npix; i += GT)
            if (px_transparent(line + (size_t)i * bpp, cb, bpp))
                for (uint32_t j = 0; j < cb; j++) line[(size_t)i * bpp + j] = prev[(size_t)i * bpp + j];
        __syncthreads();
        return;
    }
    // first non-transparent pixel (only matters when pixel 0 is transparent) :126-131
    if (tid == 0) *sh_first = 0xFFFFFFFFu;
    __syncthreads();
    if (npix && px_transparent(line, cb, bpp)) {
        uint32_t mine = 0xFFFFFFFFu;
        for (uint32_t i = tid; i < npix; i += GT)
            if (!px_transparent(line + (size_t)i * bpp, cb, bpp)) { mine = i; break; }
        if (mine != 0xFFFFFFFFu) atomicMin(sh_first, mine);
    }
    __syncthreads();
    const uint32_t first_opaque = *sh_first; // 0xFFFFFFFF: none -> prev = i (= 0)
    for (uint32_t s = tid; s < npix; s += GT) {
        uint8_t *px = line + (size_t)s * bpp;
        if (!px_transparent(px, cb, bpp)) continue;
        if (s > 0 && px_transparent(px - bpp, cb, bpp)) continue; // not a run start
        for (uint32_t i = s; i < npix; i++) {
            px = line + (size_t)i * bpp;
            if (i > s && !px_transparent(px, cb, bpp)) break;
            const uint8_t *up = prev + (size_t)i * bpp;
            const uint32_t pv = i == 0 ? (first_opaque == 0xFFFFFFFFu ? 0u : first_opaque) : i - 1;
            for (uint32_t j = 0; j < cb; j++) {
                uint32_t v;
                if (f == 1) { // Sub :140-156
                    v = line[(size_t)pv * bpp + j];
                } else if (f == 3) { // Average :160-170
                    v = i == 0 ? (uint32_t)(up[j] >> 1) : (((uint32_t)px[(int)j - (int)bpp] + up[j]) >> 1);
                } else { // Paeth :171-181
                    if (i == 0) {
                        uint32_t a = line[(size_t)pv * bpp + j];
                        v = a < up[j] ? a : up[j];
                    } else {
                        v = paeth(px[(int)j - (int)bpp], up[j], up[(int)j - (int)bpp]);
                    }
                }
                px[j] = (uint8_t)v;
            }
        }
    }
    __syncthreads();
}

// filter_line body (src/filters/mod.rs:61-110): t[0] = f, t[1+i] = residual.  prev == nullptr means a zero row.
__device__ void filter_row(int f, const uint8_t *line, const uint8_t *prev, uint32_t len, uint32_t bpp, uint8_t *t) {
    const int tid = threadIdx.x;
    if (tid == 0) t[0] = (uint8_t)f;
    for (uint32_t i = tid; i < len; i += GT) {
        uint32_t cur = line[i];
        uint32_t left = i >= bpp ? line[i - bpp] : 0u;
        uint32_t up = prev ? prev[i] : 0u;
        uint32_t ul = (prev && i >= bpp) ? prev[i - bpp] : 0u;
        t[1 + (size_t)i] = (uint8_t)residual_byte(f, cur, left, up, ul);
    }
}

__device__ void copy_bytes(uint8_t *dst, const uint8_t *src, size_t n) {
    for (size_t i = threadIdx.x; i < n; i += GT) dst[i] = src[i];
}

// One scanline through filter_image's loop body (png/mod.rs:374-424).
// ALPHA: `line` is the mutable scratch copy and sc.prev the carried previous row.
template <bool ALPHA>
__device__ void process_row(const Geom &g, const DevStrategy &st, const CtaScratch &sc, const uint8_t *row_in,
                            const uint8_t *prev_in, uint32_t len, uint32_t row_index, uint8_t *out_row,
                            uint8_t *used_slot, uint32_t *hist, unsigned long long *sh, uint32_t *sh_misc) {
    const int tid = threadIdx.x;
    const uint32_t bpp = g.bpp;
    const uint8_t *line = row_in;
    const uint8_t *prev = prev_in;
    if (ALPHA) {
        copy_bytes(sc.line, row_in, len); // line.data.to_vec() :380
        __syncthreads();
        line = sc.line;
        prev = sc.prev;
    }
    if (st.kind == OXI_BASIC || st.kind == OXI_PREDEFINED) { // :382-394
        int f = st.kind == OXI_BASIC ? st.basic : (row_index < st.n_predefined ? st.predefined[row_index] : 0);
        if (ALPHA) optimize_alpha_row(f, sc.line, prev, len, bpp, g.color_bytes, sh_misc);
        filter_row(f, line, prev, len, bpp, out_row);
        if (tid == 0) *used_slot = (uint8_t)f;
        if (ALPHA) {
            __syncthreads();
            copy_bytes(sc.prev, sc.line, len);
            __syncthreads();
        }
        return;
    }
    // all-zero shortcut :398-404
    unsigned long long nz = 0;
    for (uint32_t i = tid; i < len; i += GT) nz |= line[i];
    nz = block_sum64(nz != 0, sh);
    if (nz == 0) {
        filter_row(0, line, prev, len, bpp, out_row);
        if (tid == 0) *used_slot = 0;
        if (ALPHA) {
            __syncthreads();
            copy_bytes(sc.prev, sc.line, len);
            __syncthreads();
        }
        return;
    }
    int best_f = 0;
    unsigned long long best = 0;
    bool have = false;
    for (int f = 0; f < 5; f++) { // RowFilter::ALL order :412-420
        if (ALPHA) optimize_alpha_row(f, sc.line, prev, len, bpp, g.color_bytes, sh_misc);
        filter_row(f, line, prev, len, bpp, sc.trial);
        __syncthreads();
        unsigned long long s = score_line(st.kind, sc.trial, len + 1, sc.table, hist, sh);
        if (better(st.kind, s, best, have)) {
            best = s;
            have = true;
            best_f = f;
            if (ALPHA) {
                copy_bytes(sc.best_line, sc.trial, (size_t)len + 1);
                copy_bytes(sc.best_raw, sc.line, len);
            }
        }
        __syncthreads();
    }
    if (ALPHA) {
        copy_bytes(out_row, sc.best_line, (size_t)len + 1); // :421
        copy_bytes(sc.prev, sc.best_raw, len);              // :422
        __syncthreads();
    } else {
        filter_row(best_f, line, prev, len, bpp, out_row);
    }
    if (tid == 0) *used_slot = (uint8_t)best_f;
}

template <bool ALPHA>
__global__ void __launch_bounds__(GT) k_filter_generic(Geom g, const uint8_t *__restrict__ data, uint64_t n_images,
                                                       StrategySet set, uint8_t *scratch) {
    __shared__ uint32_t hist[256];
    __shared__ unsigned long long sh[GT / 32];
    __shared__ uint32_t sh_misc;
    const CtaScratch sc = carve(scratch + (size_t)blockIdx.x * scratch_per_cta(g), g);
    // the bigram table must start (and is always left) all-zero
    for (uint32_t i = threadIdx.x; i < 65536; i += GT) sc.table[i] = 0;
    __syncthreads();

    if (ALPHA) {
        // work item = (strategy, image); rows in order
        const uint64_t items = (uint64_t)set.k * n_images;
        for (uint64_t it = blockIdx.x; it < items; it += gridDim.x) {
            const DevStrategy &st = set.s[it / n_images];
            const uint64_t img = it % n_images;
            const uint8_t *in = data + img * g.data_len;
            for (uint32_t p = 0; p < g.n_pass; p++) {
                const PassDesc pd = g.pass[p];
                for (uint32_t i = threadIdx.x; i < pd.len; i += GT) sc.prev[i] = 0; // vec![0; len] :375-378
                __syncthreads();
                for (uint32_t k = 0; k < pd.nrows; k++) {
                    const uint32_t r = pd.row0 + k;
                    const uint64_t in_off = pd.in_base + (uint64_t)k * pd.len;
                    process_row<true>(g, st, sc, in + in_off, nullptr, pd.len, r, st.out + img * g.out_len + in_off + r,
                                      st.filters_used + img * g.total_rows + r, hist, sh, &sh_misc);
                    __syncthreads();
                }
            }
        }
    } else {
        // work item = (strategy, image, row); rows are independent when the previous raw row is the predictor
        const uint64_t rows = (uint64_t)n_images * g.total_rows;
        const uint64_t items = (uint64_t)set.k * rows;
        for (uint64_t it = blockIdx.x; it < items; it += gridDim.x) {
            const DevStrategy &st = set.s[it / rows];
            const uint64_t rr = it % rows;
            const uint64_t img = rr / g.total_rows;
            const uint32_t r = (uint32_t)(rr % g.total_rows);
            const PassDesc pd = g.pass[find_pass(g, r)];
            const uint32_t k = r - pd.row0;
            const uint64_t in_off = pd.in_base + (uint64_t)k * pd.len;
            const uint8_t *row_in = data + img * g.data_len + in_off;
            process_row<false>(g, st, sc, row_in, k == 0 ? nullptr : row_in - pd.len, pd.len, r,
                               st.out + img * g.out_len + in_off + r, st.filters_used + img * g.total_rows + r, hist, sh,
                               &sh_misc);
            __syncthreads();
        }
    }
}

} // namespace

size_t generic_scratch_bytes(const Geom &g, size_t n_images, int num_sms, int *grid_out) {
    uint64_t items = g.alpha_bytes ? (uint64_t)n_images : (uint64_t)n_images * g.total_rows;
    items *= MAX_STRATEGIES; // upper bound on k
    uint64_t grid = (uint64_t)num_sms * 4;
    if (grid > items) grid = items;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = (int)grid;
    return (size_t)grid * scratch_per_cta(g);
}

cudaError_t launch_generic(const LaunchCtx &ctx, const Geom &g, const uint8_t *d_data, size_t n_images,
                           const StrategySet &set) {
    uint64_t items = g.alpha_bytes ? (uint64_t)n_images : (uint64_t)n_images * g.total_rows;
    items *= (uint64_t)set.k;
    if (items == 0) return cudaSuccess;
    uint64_t grid = (uint64_t)ctx.num_sms * 4;
    if (grid > items) grid = items;
    if ((size_t)grid * scratch_per_cta(g) > ctx.scratch_bytes) return cudaErrorMemoryAllocation;
    if (g.alpha_bytes)
        k_filter_generic<true><<<(unsigned)grid, GT, 0, ctx.stream>>>(g, d_data, n_images, set, ctx.scratch);
    else
        k_filter_generic<false><<<(unsigned)grid, GT, 0, ctx.stream>>>(g, d_data, n_images, set, ctx.scratch);
    count_launch();
    return cudaGetLastError();
}

} // namespace oxi
