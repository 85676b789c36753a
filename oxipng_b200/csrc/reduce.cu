// reduce.cu -- the pixel-scan reductions of src/reduction/{alpha,bit_depth,color,palette}.rs as grid-wide CUDA scans,
// with the <=256-entry palette bookkeeping (dedupe sets, luma sort, tRNS choice) kept on the host exactly where the
// reference keeps it.  Every entry point takes host buffers (include/oxipng_b200.h), uploads the image once, runs the
// scan/transform kernels and downloads the result.  Return OXI_NONE where the reference returns `None`.
//
// Kernels are HBM-bound byte maps / reductions: 128-bit accesses where the pixel size divides 16, early-out flags in
// global memory, per-CTA shared-memory partials (flags, 256-bin histograms, distinct-colour hash sets) merged with a
// handful of global atomics per CTA.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "host.h"

namespace oxi {
namespace {

constexpr int RT = 256; // threads per CTA for the scans

inline unsigned grid_for(uint64_t work_items, int num_sms, int per_sm = 8) {
    uint64_t g = (work_items + RT - 1) / RT;
    const uint64_t cap = (uint64_t)num_sms * per_sm;
    if (g > cap) g = cap;
    return g ? (unsigned)g : 1u;
}

// Device-side scratch shared by the scans of one call
struct ScanState {
    uint32_t fail;             // "not reducible" / invalid
    uint32_t has_transparency; // alpha.rs:50-52
    uint32_t used[8];          // 256-bit set (alpha.rs:57-59 used_colors)
    uint32_t max_bits;         // bit_depth.rs:82-112
    uint32_t n_distinct;       // color.rs:18-35
    uint32_t overflow;         // 257th distinct colour
    uint32_t pad_[3];
    unsigned long long hist[256];
    uint32_t counts[256];
};

__device__ __forceinline__ bool alpha_zero(const uint8_t *px, uint32_t cb, uint32_t bpp) {
    for (uint32_t k = cb; k < bpp; k++)
        if (px[k]) return false;
    return true;
}

// ---- alpha.rs -----------------------------------------------------------------------------------------------

// cleaned_alpha_channel (alpha.rs:11-32): pixels whose alpha bytes are all zero become all zero.
// One thread per 16 bytes (whole pixels: bpp in {2,4,8}); `amask` marks the alpha bytes inside one pixel-sized lane.
__global__ void __launch_bounds__(RT) k_clean_alpha(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint64_t n16,
                                                     const uint8_t *__restrict__ in_tail, uint8_t *__restrict__ out_tail,
                                                     uint32_t tail_bytes, uint32_t bpp, uint32_t cb) {
    const uint64_t stride = (uint64_t)gridDim.x * RT;
    for (uint64_t i = (uint64_t)blockIdx.x * RT + threadIdx.x; i < n16; i += stride) {
        uint4 v = in[i];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        if (bpp == 2) { // GA8: alpha = byte 1 of every u16
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t x = w[k];
                if ((x & 0x0000FF00u) == 0) x &= 0xFFFF0000u;
                if ((x & 0xFF000000u) == 0) x &= 0x0000FFFFu;
                w[k] = x;
            }
        } else if (bpp == 4) { // GA16: bytes 2,3; RGBA8: byte 3
            const uint32_t am = cb == 2 ? 0xFFFF0000u : 0xFF000000u;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if ((w[k] & am) == 0) w[k] = 0;
        } else { // RGBA16: bytes 6,7 of every 8
            if ((w[1] & 0xFFFF0000u) == 0) w[0] = w[1] = 0;
            if ((w[3] & 0xFFFF0000u) == 0) w[2] = w[3] = 0;
        }
        out[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { // < 16 trailing bytes
        for (uint32_t p = 0; p + bpp <= tail_bytes; p += bpp) {
            const bool z = alpha_zero(in_tail + p, cb, bpp);
            for (uint32_t k = 0; k < bpp; k++) out_tail[p + k] = z ? 0 : in_tail[p + k];
        }
    }
}

// reduced_alpha_channel scan (alpha.rs:49-60)
__global__ void __launch_bounds__(RT) k_alpha_scan(const uint8_t *__restrict__ in, uint64_t npix, uint32_t bpp, uint32_t cb,
                                                    int optimize_alpha, ScanState *st) {
    __shared__ uint32_t s_used[8];
    __shared__ uint32_t s_flags[2];
    if (threadIdx.x < 8) s_used[threadIdx.x] = 0;
    if (threadIdx.x < 2) s_flags[threadIdx.x] = 0;
    __syncthreads();
    uint32_t fail = 0, transp = 0;
    const uint64_t stride = (uint64_t)gridDim.x * RT;
    for (uint64_t i = (uint64_t)blockIdx.x * RT + threadIdx.x; i < npix; i += stride) {
        const uint8_t *px = in + i * bpp;
        if (optimize_alpha && alpha_zero(px, cb, bpp)) {
            transp = 1;
            continue;
        }
        bool partial = false;
        for (uint32_t k = cb; k < bpp; k++) partial |= px[k] != 255;
        if (partial) {
            fail = 1;
        } else if (optimize_alpha) {
            bool gray = true;
            for (uint32_t k = 1; k < cb; k++) gray &= px[k] == px[0];
            if (gray) atomicOr(&s_used[px[0] >> 5], 1u << (px[0] & 31));
        }
    }
    if (fail) s_flags[0] = 1;
    if (transp) s_flags[1] = 1;
    __syncthreads();
    if (threadIdx.x < 8 && s_used[threadIdx.x]) atomicOr(&st->used[threadIdx.x], s_used[threadIdx.x]);
    if (threadIdx.x == 0) {
        if (s_flags[0]) st->fail = 1;
        if (s_flags[1]) st->has_transparency = 1;
    }
}

// reduced_alpha_channel transform (alpha.rs:77-85): drop the alpha bytes; transparent pixels become `trns` repeated
__global__ void __launch_bounds__(RT) k_alpha_strip(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                                                     uint64_t npix, uint32_t bpp, uint32_t cb, int trns) {
    const uint64_t stride = (uint64_t)gridDim.x * RT;
    for (uint64_t i = (uint64_t)blockIdx.x * RT + threadIdx.x; i < npix; i += stride) {
        const uint8_t *px = in + i * bpp;
        uint8_t *o = out + i * cb;
        const bool t = trns >= 0 && alpha_zero(px, cb, bpp);
        for (uint32_t k = 0; k < cb; k++) o[k] = t ? (uint8_t)trns : px[k];
    }
}

// ---- bit_depth.rs ---------------------------------------------------------------------------------------------

// reduced_bit_depth_16_to_8 / scaled_bit_depth_16_to_8 (bit_depth.rs:9-64): 8 samples (16 B in, 8 B out) per thread.
// The f32 expression round(v * (255/65535)) equals (510 v + 65535) / 131070 for all 65536 inputs (tests pin this).
__device__ __forceinline__ uint32_t scale16(uint32_t hi, uint32_t lo) {
    if (hi == lo) return hi;
    const uint32_t v = (hi << 8) | lo;
    return (510u * v + 65535u) / 131070u;
}
__global__ void __launch_bounds__(RT) k_16to8(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, uint64_t nsamples,
                                               int scale, ScanState *st) {
    const uint64_t n8 = nsamples / 8, stride = (uint64_t)gridDim.x * RT;
    uint32_t bad = 0;
    const uint4 *in4 = reinterpret_cast<const uint4 *>(in);
    uint2 *out2 = reinterpret_cast<uint2 *>(out);
    for (uint64_t i = (uint64_t)blockIdx.x * RT + threadIdx.x; i < n8; i += stride) {
        const uint4 v = in4[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t h0 = w[k] & 255u, l0 = (w[k] >> 8) & 255u, h1 = (w[k] >> 16) & 255u, l1 = w[k] >> 24;
            bad |= (h0 ^ l0) | (h1 ^ l1);
            const uint32_t a = scale ? scale16(h0, l0) : h0, b = scale ? scale16(h1, l1) : h1;
            o[k >> 1] |= (a | (b << 8)) << (16 * (k & 1));
        }
        out2[i] = make_uint2(o[0], o[1]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (uint64_t s = n8 * 8; s < nsamples; s++) {
            const uint32_t h = in[2 * s], l = in[2 * s + 1];
            bad |= h ^ l;
            out[s] = (uint8_t)(scale ? scale16(h, l) : h);
        }
    if (!scale && __syncthreads_or(bad != 0) && threadIdx.x == 0) st->fail = 1;
}

// per-byte minimal bit period of a gray sample (bit_depth.rs:82-112 is the running maximum of this over all bytes)
__device__ __forceinline__ uint32_t min_bits_of(uint32_t b) {
    if (b == 0 || b == 255) return 1;
    if (b == (b & 3u) * 0x55u) return 2;
    if (b == (b & 15u) * 0x11u) return 4;
    return 8;
}
// Word-parallel form: a byte has period d iff bit i equals bit i+d for every i, i.e. ((b >> d) ^ b) has no bit set
// among its low 8-d bits; period 1 implies 2 implies 4, so three OR-accumulators give the maximum over all bytes.
__global__ void __launch_bounds__(RT) k_min_bits(const uint8_t *__restrict__ in, uint64_t n, ScanState *st) {
    uint32_t n8 = 0, n4 = 0, n2 = 0; // some byte is not 4- / 2- / 1-periodic
    const uint64_t n16 = n / 16, stride = (uint64_t)gridDim.x * RT;
    const uint4 *in4 = reinterpret_cast<const uint4 *>(in);
    for (uint64_t i = (uint64_t)blockIdx.x * RT + threadIdx.x; i < n16; i += stride) {
        const uint4 v = in4[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            n8 |= ((w[k] >> 4) ^ w[k]) & 0x0F0F0F0Fu;
            n4 |= ((w[k] >> 2) ^ w[k]) & 0x3F3F3F3Fu;
            n2 |= ((w[k] >> 1) ^ w[k]) & 0x7F7F7F7Fu;
        }
    }
    uint32_t m = n8 ? 8u : n4 ? 4u : n2 ? 2u : 1u;
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (uint64_t i = n16 * 16; i < n; i++) m = max(m, min_bits_of(in[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 1) atomicMax(&st->max_bits, m);
}

struct RowMap { // per pass: rows of `in_len` bytes -> rows of `out_len` bytes
    uint32_t n_pass;
    uint32_t in_len[7], out_len[7], nrows[7], pixels[7];
    uint64_t in_base[7], out_base[7], out_items[8]; // out_items: prefix sum of nrows*out_len
};

__device__ __forceinline__ int pass_of_item(const RowMap &m, uint64_t item) {
    int p = 0;
#pragma unroll
    for (int i = 1; i < 7; i++)
        if (i < (int)m.n_pass && item >= m.out_items[i]) p = i;
    return p;
}

// reduced_bit_depth_8_or_less packing (bit_depth.rs:117-131): one thread per output byte
__global__ void __launch_bounds__(RT) k_pack_bits(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, RowMap m,
                                                   uint32_t bits) {
    const uint32_t per = 8 / bits, mask = (1u << bits) - 1u;
    const uint64_t total = m.out_items[m.n_pass], stride = (uint64_t)gridDim.x * RT;
    for (uint64_t it = (uint64_t)blockIdx.x * RT + threadIdx.x; it < total; it += stride) {
        const int p = pass_of_item(m, it);
        const uint64_t local = it - m.out_items[p];
        const uint64_t row = local / m.out_len[p];
        const uint32_t ob = (uint32_t)(local % m.out_len[p]);
        const uint8_t *src = in + m.in_base[p] + row * m.in_len[p];
        uint32_t nb = 0, shift = 8;
        for (uint32_t k = ob * per; k < ob * per + per && k < m.in_len[p]; k++) {
            shift -= bits;
            nb |= (src[k] & mask) << shift;
        }
        out[m.out_base[p] + local] = (uint8_t)nb;
    }
}

// expanded_bit_depth_to_8 (bit_depth.rs:170-230): one thread per output pixel
__global__ void __launch_bounds__(RT) k_expand_bits(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, RowMap m,
                                                     uint32_t depth, int is_gray) {
    const uint32_t mask = (1u << depth) - 1u;
    const uint32_t rep = depth == 1 ? 0xFFu : depth == 2 ? 0x55u : 0x11u;
    const uint64_t total = m.out_items[m.n_pass], stride = (uint64_t)gridDim.x * RT;
    for (uint64_t it = (uint64_t)blockIdx.x * RT + threadIdx.x; it < total; it += stride) {
        const int p = pass_of_item(m, it);
        const uint64_t local = it - m.out_items[p];
        const uint64_t row = local / m.out_len[p];
        const uint32_t x = (uint32_t)(local % m.out_len[p]);
        const uint32_t bit = x * depth;
        const uint32_t byte = in[m.in_base[p] + row * m.in_len[p] + (bit >> 3)];
        uint32_t val = (byte >> (8 - depth - (bit & 7u))) & mask;
        if (is_gray) val *= rep;
        out[m.out_base[p] + local] = (uint8_t)val;
    }
}

// ---- color.rs ---------------------------------------------------------------------------------------------------

// reduced_rgb_to_grayscale (color.rs:100-137): check R==G==B per pixel, keep the last colour sample (+ alpha)
__global__ void __launch_bounds__(RT) k_rgb_to_gray(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, uint64_t npix,
                                                     uint32_t bpp, uint32_t bd, ScanState *st) {
    const uint64_t stride = (uint64_t)gridDim.x * RT;
    const uint32_t last = 2 * bd, keep = bpp - last;
    uint32_t bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * RT + threadIdx.x; i < npix; i += stride) {
        const uint8_t *px = in + i * bpp;
        if (bd == 1) bad |= (px[0] ^ px[1]) | (px[1] ^ px[2]);
        else bad |= (px[0] ^ px[2]) | (px[1] ^ px[3]) | (px[2] ^ px[4]) | (px[3] ^ px[5]);
        for (uint32_t k = 0; k < keep; k++) out[i * keep + k] = px[last + k];
    }
    if (__syncthreads_or(bad != 0) && threadIdx.x == 0) st->fail = 1;
}

// Distinct-colour set with first-occurrence index (build_palette, color.rs:18-35).
// Global table: 4096 slots of (occupied|key) u64 + first pixel index u64.  Each CTA dedupes its contiguous pixel range
// in a 1024-slot shared-memory table (smem atomics only), then merges its <=256 entries.
constexpr uint32_t GSLOTS = 4096, LSLOTS = 1024;
struct ColorTable {
    unsigned long long key[GSLOTS];   // 0 = empty, else (1<<32) | key
    unsigned long long first[GSLOTS]; // smallest pixel index holding this colour
    uint32_t index[GSLOTS];           // palette index (filled by the host after sorting by `first`)
};
__device__ __forceinline__ uint32_t hash_color(uint32_t key) { return (key * 2654435761u) ^ (key >> 15); }
__device__ __forceinline__ uint32_t load_key(const uint8_t *px, uint32_t ch) {
    if (ch == 4 && (reinterpret_cast<uintptr_t>(px) & 3u) == 0) return *reinterpret_cast<const uint32_t *>(px); // RGBA8
    if (ch == 2 && (reinterpret_cast<uintptr_t>(px) & 1u) == 0) return *reinterpret_cast<const uint16_t *>(px); // GA8
    uint32_t k = px[0];
    if (ch > 1) k |= (uint32_t)px[1] << 8;
    if (ch > 2) k |= (uint32_t)px[2] << 16;
    if (ch > 3) k |= (uint32_t)px[3] << 24;
    return k;
}

__global__ void __launch_bounds__(RT) k_collect_colors(const uint8_t *__restrict__ in, uint64_t npix, uint32_t ch,
                                                        ColorTable *tab, ScanState *st) {
    __shared__ unsigned long long l_key[LSLOTS];
    __shared__ unsigned long long l_first[LSLOTS];
    __shared__ uint32_t l_count;
    for (uint32_t i = threadIdx.x; i < LSLOTS; i += RT) { l_key[i] = 0; l_first[i] = ~0ull; }
    if (threadIdx.x == 0) l_count = 0;
    __syncthreads();
    const uint64_t per = (npix + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = (uint64_t)blockIdx.x * per, hi = min(lo + per, npix);
    uint32_t last_key = 0;
    bool have_last = false;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += RT) {
        if (l_count > 256 || *(volatile uint32_t *)&st->overflow) break; // 257th distinct colour: give up (color.rs:29-31)
        const uint32_t key = load_key(in + i * ch, ch);
        if (have_last && key == last_key) continue; // this thread already recorded an earlier occurrence
        last_key = key;
        have_last = true;
        const unsigned long long tagged = (1ull << 32) | key;
        uint32_t s = hash_color(key) & (LSLOTS - 1);
        for (;;) {
            unsigned long long cur = l_key[s];
            if (cur == 0) {
                cur = atomicCAS(&l_key[s], 0ull, tagged);
                if (cur == 0) { atomicAdd(&l_count, 1u); cur = tagged; }
            }
            if (cur == tagged) {
                // first[] only ever decreases and a thread's pixels come in increasing order: after the first occurrence
                // the (64-bit, CAS-emulated) shared atomic is almost never needed
                if ((unsigned long long)i < *(volatile unsigned long long *)&l_first[s]) atomicMin(&l_first[s], (unsigned long long)i);
                break;
            }
            s = (s + 1) & (LSLOTS - 1);
            if (l_count > 256) break;
        }
    }
    __syncthreads();
    if (l_count > 256) {
        if (threadIdx.x == 0) st->overflow = 1;
        return;
    }
    for (uint32_t i = threadIdx.x; i < LSLOTS; i += RT) {
        const unsigned long long tagged = l_key[i];
        if (!tagged) continue;
        uint32_t s = hash_color((uint32_t)tagged) & (GSLOTS - 1);
        for (uint32_t probes = 0; probes < GSLOTS; probes++) {
            unsigned long long cur = tab->key[s];
            if (cur == 0) {
                cur = atomicCAS(&tab->key[s], 0ull, tagged);
                if (cur == 0) {
                    if (atomicAdd(&st->n_distinct, 1u) >= 256) st->overflow = 1;
                    cur = tagged;
                }
            }
            if (cur == tagged) { atomicMin(&tab->first[s], l_first[i]); break; }
            s = (s + 1) & (GSLOTS - 1);
            if (*(volatile uint32_t *)&st->overflow) break;
        }
    }
}

// pixel -> palette index through the (now read-only) colour table, staged in shared memory
__global__ void __launch_bounds__(RT) k_index_colors(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, uint64_t npix,
                                                      uint32_t ch, const ColorTable *__restrict__ tab) {
    extern __shared__ unsigned long long s_tab[]; // GSLOTS keys, then GSLOTS u8 indices
    uint8_t *s_idx = reinterpret_cast<uint8_t *>(s_tab + GSLOTS);
    for (uint32_t i = threadIdx.x; i < GSLOTS; i += RT) {
        s_tab[i] = tab->key[i];
        s_idx[i] = (uint8_t)tab->index[i];
    }
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * RT;
    for (uint64_t i = (uint64_t)blockIdx.x * RT + threadIdx.x; i < npix; i += stride) {
        const uint32_t key = load_key(in + i * ch, ch);
        const unsigned long long tagged = (1ull << 32) | key;
        uint32_t s = hash_color(key) & (GSLOTS - 1);
        while (s_tab[s] != tagged) s = (s + 1) & (GSLOTS - 1);
        out[i] = s_idx[s];
    }
}

// ---- palette.rs ---------------------------------------------------------------------------------------------------

// 256-bin histogram of the index bytes (palette.rs:20-23 needs presence; counts are exposed too)
__global__ void __launch_bounds__(RT) k_hist256(const uint8_t *__restrict__ in, uint64_t n, ScanState *st) {
    __shared__ uint32_t h[4][256];
    for (uint32_t i = threadIdx.x; i < 1024; i += RT) (&h[0][0])[i] = 0;
    __syncthreads();
    uint32_t *mine = h[threadIdx.x & 3];
    const uint64_t n16 = n / 16, stride = (uint64_t)gridDim.x * RT;
    const uint4 *in4 = reinterpret_cast<const uint4 *>(in);
    for (uint64_t i = (uint64_t)blockIdx.x * RT + threadIdx.x; i < n16; i += stride) {
        const uint4 v = in4[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // runs of equal indices are common in palette images: merge a word of four equal bytes into one add
            const uint32_t b0 = w[k] & 255u;
            if (w[k] == b0 * 0x01010101u) {
                atomicAdd(&mine[b0], 4u);
            } else {
                atomicAdd(&mine[b0], 1u);
                atomicAdd(&mine[(w[k] >> 8) & 255u], 1u);
                atomicAdd(&mine[(w[k] >> 16) & 255u], 1u);
                atomicAdd(&mine[w[k] >> 24], 1u);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (uint64_t i = n16 * 16; i < n; i++) atomicAdd(&mine[in[i]], 1u);
    __syncthreads();
    const uint32_t c = h[0][threadIdx.x] + h[1][threadIdx.x] + h[2][threadIdx.x] + h[3][threadIdx.x];
    if (c) atomicAdd(&st->hist[threadIdx.x], (unsigned long long)c);
}

// byte -> byte look-up-table remap (palette.rs:43,121,183), 16 bytes per thread
__global__ void __launch_bounds__(RT) k_lut(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, uint64_t n,
                                             const uint8_t *__restrict__ lut_g) {
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = lut_g[threadIdx.x];
    __syncthreads();
    const uint64_t n16 = n / 16, stride = (uint64_t)gridDim.x * RT;
    const uint4 *in4 = reinterpret_cast<const uint4 *>(in);
    uint4 *out4 = reinterpret_cast<uint4 *>(out);
    for (uint64_t i = (uint64_t)blockIdx.x * RT + threadIdx.x; i < n16; i += stride) {
        const uint4 v = in4[i];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++)
            w[k] = lut[w[k] & 255u] | (lut[(w[k] >> 8) & 255u] << 8) | (lut[(w[k] >> 16) & 255u] << 16) |
                   (lut[w[k] >> 24] << 24);
        out4[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (uint64_t i = n16 * 16; i < n; i++) out[i] = lut[in[i]];
}

// most_popular_edge_color counts (palette.rs:198-204): first and last byte of every scanline of >= 2 bytes
__global__ void __launch_bounds__(RT) k_edge_counts(const uint8_t *__restrict__ in, Geom g, ScanState *st) {
    const uint32_t stride = gridDim.x * RT;
    for (uint32_t r = blockIdx.x * RT + threadIdx.x; r < g.total_rows; r += stride) {
        const PassDesc pd = g.pass[find_pass(g, r)];
        if (pd.len < 2) continue;
        const uint8_t *row = in + pd.in_base + (uint64_t)(r - pd.row0) * pd.len;
        atomicAdd(&st->counts[row[0]], 1u);
        atomicAdd(&st->counts[row[pd.len - 1]], 1u);
    }
}

// CoOccurrenceMatrix::build counting (palette.rs:556-577).  One thread per 16-byte segment of a scanline; equal
// consecutive (from,to) pairs are merged before they hit the matrix (flat regions are the common case).
// `m` is n*n u32 in global memory (L2 resident, <= 256 KB).
__global__ void __launch_bounds__(RT) k_cooccurrence(const uint8_t *__restrict__ in, Geom g, uint32_t n, uint32_t *m,
                                                      ScanState *st) {
    // work item = (row, segment); rows of different passes have different segment counts, so walk pass by pass
    uint64_t base_item = 0;
    const uint64_t tid = (uint64_t)blockIdx.x * RT + threadIdx.x, stride = (uint64_t)gridDim.x * RT;
    for (uint32_t p = 0; p < g.n_pass; p++) {
        const PassDesc pd = g.pass[p];
        const uint32_t segs = (pd.len + 15) / 16;
        const uint64_t items = (uint64_t)pd.nrows * segs;
        // first item of this pass owned by this thread
        uint64_t it = tid >= base_item % stride ? tid - base_item % stride : tid + stride - base_item % stride;
        for (; it < items; it += stride) {
            const uint32_t k = (uint32_t)(it / segs), seg = (uint32_t)(it % segs);
            const uint8_t *row = in + pd.in_base + (uint64_t)k * pd.len;
            // previous scanline in stream order, also across passes (palette.rs:566-575)
            const uint8_t *prow = nullptr;
            uint32_t plen = 0;
            if (k > 0) { prow = row - pd.len; plen = pd.len; }
            else if (p > 0) {
                const PassDesc pp = g.pass[p - 1];
                prow = in + pp.in_base + (uint64_t)(pp.nrows - 1) * pp.len;
                plen = pp.len;
            }
            const uint32_t i0 = seg * 16, i1 = min(i0 + 16, pd.len);
            uint32_t hk = 0xFFFFFFFFu, hc = 0, vk = 0xFFFFFFFFu, vc = 0;
            uint32_t left = i0 ? row[i0 - 1] : 0xFFFFFFFFu;
            for (uint32_t i = i0; i < i1; i++) {
                const uint32_t val = row[i];
                if (val >= n) { st->fail = 1; left = val; continue; }
                if (left != 0xFFFFFFFFu && left < n) {
                    const uint32_t key = left * n + val;
                    if (key == hk) hc++;
                    else { if (hc) atomicAdd(&m[hk], hc); hk = key; hc = 1; }
                }
                if (prow && i < plen) {
                    const uint32_t pvv = prow[i];
                    if (pvv < n) {
                        const uint32_t key = pvv * n + val;
                        if (key == vk) vc++;
                        else { if (vc) atomicAdd(&m[vk], vc); vk = key; vc = 1; }
                    } else st->fail = 1;
                }
                left = val;
            }
            if (hc) atomicAdd(&m[hk], hc);
            if (vc) atomicAdd(&m[vk], vc);
        }
        base_item += items;
    }
}
// symmetrise (palette.rs:580-587): M[i][j] = M[j][i] = M[i][j] + M[j][i]; the diagonal doubles
__global__ void k_symmetrise(uint32_t *m, uint32_t n) {
    const uint32_t i = blockIdx.x, j = threadIdx.x;
    if (i >= n || j > i) return;
    const uint32_t a = m[i * n + j], b = m[j * n + i];
    const uint32_t s = (i == j) ? a + a : a + b;
    m[i * n + j] = s;
    m[j * n + i] = s;
}


// ---- interlace.rs: Adam7 as a pure gather ----------------------------------------------------------
struct Adam7 {
    uint32_t width, height, bits; // bits per pixel
    uint32_t pw[7], ph[7];        // pixels per row / rows of each pass (0 = empty pass)
    uint32_t stride[7];           // bytes per pass row
    uint64_t base[8];             // byte offset of each pass in the interlaced layout; base[7] = total
    uint32_t full_stride;         // bytes per row of the progressive layout
};
__constant__ uint8_t kA7x0[7] = {0, 4, 0, 2, 0, 1, 0}, kA7y0[7] = {0, 0, 4, 0, 2, 0, 1};
__constant__ uint8_t kA7dx[7] = {8, 8, 4, 4, 2, 2, 1}, kA7dy[7] = {8, 8, 8, 4, 4, 2, 2};

__device__ __forceinline__ uint32_t get_bits(const uint8_t *row, uint32_t bit, uint32_t nbits) { // nbits in {1,2,4}
    return (row[bit >> 3] >> (8 - nbits - (bit & 7u))) & ((1u << nbits) - 1u);
}

// interlace_image (interlace.rs:6-60): one thread per byte of the interlaced layout
__global__ void __launch_bounds__(RT) k_interlace(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, Adam7 a) {
    const uint64_t total = a.base[7], step = (uint64_t)gridDim.x * RT;
    for (uint64_t o = (uint64_t)blockIdx.x * RT + threadIdx.x; o < total; o += step) {
        int p = 0;
#pragma unroll
        for (int i = 1; i < 7; i++)
            if (o >= a.base[i]) p = i;
        const uint64_t local = o - a.base[p];
        const uint32_t r = (uint32_t)(local / a.stride[p]), ob = (uint32_t)(local % a.stride[p]);
        const uint8_t *src = in + (uint64_t)(kA7y0[p] + r * kA7dy[p]) * a.full_stride;
        uint32_t v = 0;
        if (a.bits >= 8) {
            const uint32_t B = a.bits >> 3, c = ob / B, b = ob % B;
            v = src[(uint64_t)(kA7x0[p] + c * kA7dx[p]) * B + b];
        } else {
            const uint32_t ppb = 8 / a.bits;
            for (uint32_t q = 0; q < ppb; q++) {
                const uint32_t c = ob * ppb + q;
                if (c < a.pw[p]) v |= get_bits(src, (kA7x0[p] + c * kA7dx[p]) * a.bits, a.bits) << (8 - a.bits * (q + 1));
            }
        }
        out[o] = (uint8_t)v;
    }
}

// deinterlace_image (interlace.rs:62-150): one thread per byte of the progressive layout
__global__ void __launch_bounds__(RT) k_deinterlace(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, Adam7 a) {
    // pass of pixel (x & 7, y & 7)
    const uint8_t pass_of[64] = {0, 5, 3, 5, 1, 5, 3, 5, 6, 6, 6, 6, 6, 6, 6, 6, 4, 5, 4, 5, 4, 5, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6,
                                 2, 5, 3, 5, 2, 5, 3, 5, 6, 6, 6, 6, 6, 6, 6, 6, 4, 5, 4, 5, 4, 5, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6};
    const uint64_t total = (uint64_t)a.full_stride * a.height, step = (uint64_t)gridDim.x * RT;
    for (uint64_t o = (uint64_t)blockIdx.x * RT + threadIdx.x; o < total; o += step) {
        const uint32_t y = (uint32_t)(o / a.full_stride), ob = (uint32_t)(o % a.full_stride);
        uint32_t v = 0;
        if (a.bits >= 8) {
            const uint32_t B = a.bits >> 3, x = ob / B, b = ob % B;
            const int p = pass_of[(y & 7u) * 8 + (x & 7u)];
            const uint32_t c = (x - kA7x0[p]) / kA7dx[p], r = (y - kA7y0[p]) / kA7dy[p];
            v = in[a.base[p] + (uint64_t)r * a.stride[p] + (uint64_t)c * B + b];
        } else {
            const uint32_t ppb = 8 / a.bits;
            for (uint32_t q = 0; q < ppb; q++) {
                const uint32_t x = ob * ppb + q;
                if (x >= a.width) break;
                const int p = pass_of[(y & 7u) * 8 + (x & 7u)];
                const uint32_t c = (x - kA7x0[p]) / kA7dx[p], r = (y - kA7y0[p]) / kA7dy[p];
                v |= get_bits(in + a.base[p] + (uint64_t)r * a.stride[p], c * a.bits, a.bits) << (8 - a.bits * (q + 1));
            }
        }
        out[o] = (uint8_t)v;
    }
}


// ---- unfilter_image (png/mod.rs:335-351, filters/mod.rs:188-239), SURVEY 8f row 1 ----------------------------------
// Reconstruction is a recurrence along the row (Sub/Average/Paeth use the rebuilt left pixel) and across rows
// (Up/Average/Paeth use the rebuilt previous row): the longest dependency chain of an image is width + height pixels.
// A warp takes a *tile* of 32 consecutive rows, lane = row, and walks it in skewed lockstep: at step s lane l rebuilds
// pixel s - delta(l) of its row.  delta is 0 for a row that does not look at the row above it (None, Sub, the first row
// of a tile) and delta(l-1) + 1 otherwise, so the pixel above was finished by lane l-1 in the step before and arrives
// by one shuffle, "upleft" is the value shuffled in one step earlier and "left" never leaves the lane's registers.
// The rows of a tile use different filters: predictors are selected by per-lane masks instead of a divergent switch,
// and the step body is instantiated three times (MODE) so that a tile without Paeth (or Average) rows does not pay
// for that predictor.
//
// Tiles form a pipeline across warps, CTAs and SMs: the first row of tile t needs the last row of tile t-1 -- unless it
// is a None/Sub row, then the tile is independent.  The producer publishes every pixel of its last row as one 64-bit
// word (tile tag << 32 | pixel) in a per-tile boundary array; the consumer reads eight of them per sub-chunk, one
// sub-chunk ahead of its use, and re-polls only the words whose tag is still missing.  Data and flag travel in the same
// naturally-atomic word: no fence, no acquire/release pair, one L2 round trip off the dependency chain.  The consumer
// trails its producer by delta(last row) + ~16 steps.  All CTAs of a launch are co-resident (cooperative launch,
// grid <= #SMs) and a warp takes its tiles in increasing order, so waiting on tile t-1 cannot deadlock.
//
// Memory traffic stays off the step-to-step chain and off the LSU: a lane's row is a different cache line from its
// neighbour's, so a per-lane load or store costs 32 tag cycles and four warps per SM would spend their time there
// (measured: 320 cycles per step).  Instead the warp moves one row's piece at a time, lanes = consecutive words: the
// filtered bytes of every row's next 32 pixels (16 for 6/8-byte pixels) arrive as aligned words by cp.async in a per-row
// window of shared memory (33-word stride: conflict-free for the per-lane reads), the step extracts its pixel with two
// LDS and one funnel shift and parks the result in a [row][pixel] array, and every row's 32 pixels leave with one
// coalesced store.  Both transfers are software-pipelined INTO the steps: step j of chunk c also fetches row j of chunk
// c+1 and stores row j of chunk c-1 (double-buffered windows and result arrays, branch-free: rows and pixels that do
// not exist are clipped by one unsigned compare on tile-relative 32-bit offsets), so a lone warp per scheduler has
// independent work to issue while the dependency chain of the step waits.
constexpr int UF_WARPS = 4, UF_RAW = 33, UF_RING = 64;
// per-warp shared memory, in words: two windows x 32 rows, two result buffers of one or two [32][CH+1] arrays
constexpr int uf_warp_words(int bpp) { return 2 * 32 * UF_RAW + 2 * (bpp > 4 ? 2 * 32 * 17 : 32 * 33); }
constexpr uint32_t UF_MAX_ROW = 1u << 24; // tile-relative offsets are 32-bit: rows of 16 MB and more are refused

// Shared-memory mailbox between the CTA's four pipeline warps and its poller warp.  Polling the boundary words from a
// pipeline warp costs it one exposed L2 round trip per sub-chunk (its loads queue behind its own stores and copies:
// measured 224 instead of 121 cycles per step, whatever the prefetch distance); the poller warp does nothing else, keeps
// 32 words per pipeline warp in flight and hands the pixels over through a small ring, so a pipeline warp only ever
// waits on shared memory.
struct UfSlot {
    unsigned long long ent;   // request: boundary words of the tile above (second array: + ent_words)
    uint32_t npix, tag;       // request: pixels per row, expected tag
    uint32_t gen;             // request number, bumped last (0 = none yet)
    uint32_t avail;           // poller -> pipeline warp: pixels [0, avail) of the boundary row are in the ring
    uint32_t cons;            // pipeline warp -> poller: pixels [0, cons) have been taken out of the ring
    uint32_t pad_;
    uint32_t ring0[UF_RING], ring1[UF_RING];
};
struct UfCtl {
    UfSlot slot[UF_WARPS];
    uint32_t done; // pipeline warps that have finished all their items
};

struct UfPlan {
    uint64_t ent_base[7]; // first boundary word of each pass (tile t of pass p: ent_base[p] + t * pixels)
    uint64_t ent_words;   // words per boundary array of one image
    uint32_t ent_arrays;  // 1, or 2 for 6/8-byte pixels (second pixel word)
    uint32_t ctas_per_item;
    uint32_t backoff_ns;  // poller sleep after a round without progress (batches: several CTAs share an SM's issue slots)
};

__device__ __forceinline__ uint32_t vadd4_keep(uint32_t d, uint32_t p, uint32_t keep) { // (d + p) per byte, & keep
    const uint32_t s = (d & 0x7f7f7f7fu) + (p & 0x7f7f7f7fu);
    return (s ^ ((d ^ p) & 0x80808080u)) & keep;
}
__device__ __forceinline__ unsigned long long ld_entry(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Stores of the step loop are written as global-space asm WITHOUT a memory clobber: nothing in this thread reads them
// back, and a generic store (or a clobber) would pin every shared-memory load of the unrolled steps behind it.
__device__ __forceinline__ void st_entry(unsigned long long *p, unsigned long long v) {
    asm volatile("st.global.u64 [%0], %1;" ::"l"(p), "l"(v));
}
__device__ __forceinline__ void stg8(uint8_t *p, uint32_t v) { asm volatile("st.global.u8 [%0], %1;" ::"l"(p), "r"(v)); }
__device__ __forceinline__ void stg16(uint8_t *p, uint32_t v) { asm volatile("st.global.u16 [%0], %1;" ::"l"(p), "h"((unsigned short)v)); }
__device__ __forceinline__ void stg32(uint8_t *p, uint32_t v) { asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v)); }
// one rebuilt pixel to global memory; p is aligned to the pixel's natural granule (2 for 2/6-byte pixels, 4 for 4, 8 for 8)
template <int BPP>
__device__ __forceinline__ void st_pixel(uint8_t *p, uint32_t r0, uint32_t r1) {
    if constexpr (BPP == 1) {
        stg8(p, r0);
    } else if constexpr (BPP == 2) {
        stg16(p, r0);
    } else if constexpr (BPP == 3) {
        stg8(p, r0); stg8(p + 1, r0 >> 8); stg8(p + 2, r0 >> 16);
    } else if constexpr (BPP == 4) {
        stg32(p, r0);
    } else if constexpr (BPP == 6) {
        stg16(p, r0); stg16(p + 2, r0 >> 16); stg16(p + 4, r1);
    } else {
        asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(r0), "r"(r1));
    }
}

struct UfTile {
    const uint8_t *src0;        // filtered row of the tile's first row (filter byte first); row l: + l * (len + 1)
    uint8_t *dst0;              // rebuilt row of the tile's first row; row l: + l * len
    const uint8_t *in_lo, *in_hi; // bounds of the filtered buffer (aligned word loads are clipped to them)
    const unsigned long long *ent_in; // boundary words of tile t-1 (read when dep)
    unsigned long long *ent_out;      // boundary words of this tile (written when pub)
    uint64_t ent_words;
    uint32_t len, npix, delta, maxd, lastlane, tag;
    uint32_t m1, m2, m3, m4;
    bool valid, dep, pub;
    UfSlot *slot;  // this warp's mailbox
    uint32_t gen;  // request number of this tile (dep tiles)
};

// MODE 0: None/Sub/Up rows only; 1: + Average; 2: + Paeth.  SUB: steps per sub-chunk (8 or 16) = the unrolled body and the
// granularity of the boundary hand-over: 16 gives a lone warp more independent work per basic block (one image: -8..12 %),
// 8 keeps the body small when three CTAs share an SM's instruction cache (batches) and the registers of 6/8-byte pixels
template <int BPP, int MODE, int SUB>
__device__ __noinline__ void uf_run(const UfTile T, uint32_t *sm, const uint32_t lane) {
    constexpr int N1 = BPP > 4 ? BPP - 4 : 0;
    constexpr int CH = BPP > 4 ? 16 : 32;          // pixels (= steps) per chunk
    constexpr int RAWW = CH * BPP / 4 + 1;         // aligned words that hold a chunk at any misalignment (<= 33)
    constexpr int RAWL = RAWW < 32 ? RAWW : 32;    // of them, moved by the row-wise copy (word 32 by the row's own lane)
    constexpr int OS = CH + 1;                     // stride of the result arrays
    constexpr int OBUF = (N1 ? 2 : 1) * 32 * OS;   // words per result buffer
    constexpr int NOROW = -(1 << 30);              // offset of a row that does not exist: fails every range check
    uint32_t *raw = sm, *O = sm + 2 * 32 * UF_RAW;
    const uint32_t raw_s = (uint32_t)__cvta_generic_to_shared(raw);
    const uint32_t nchunks = (T.npix + T.maxd + CH - 1) / CH;
    // this lane's row, relative to the tile: first byte of chunk 0's window (pixel -delta), aligned down to a word
    const int p0 = (int)(lane * (T.len + 1) + 1) - (int)(T.delta * BPP);
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(T.src0 + p0) & 3u);
    const int srcoff = T.valid ? p0 - (int)mis : NOROW;
    const int dstoff = T.valid ? (int)(lane * T.len) - (int)(T.delta * BPP) : NOROW;
    const ptrdiff_t lo_d = T.in_lo - T.src0, hi_d = T.in_hi - T.src0;
    const int lo_rel = lo_d < -(ptrdiff_t)(1 << 29) ? -(1 << 29) : (int)lo_d;
    const int hi_rel = hi_d > (ptrdiff_t)0x7fffffff ? 0x7fffffff : (int)hi_d;
    const bool publane = T.pub && lane == T.lastlane;

    // row r's words of a chunk -> its window; lanes = consecutive words of that row.  win_s: shared address of this
    // lane's word of row 0 in the destination window, add: the chunk's byte offset + 4 * lane
    auto fetch_row = [&](uint32_t win_s, int add, int r) {
        const int pos = __shfl_sync(0xffffffffu, srcoff, r) + add;
        if (lane < RAWL && pos >= lo_rel && pos < hi_rel)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(win_s + r * (UF_RAW * 4)), "l"(T.src0 + pos));
    };
    auto fetch_own_tail = [&](uint32_t cc) { // word 32 of this lane's own window
        if (RAWW > 32) {
            const int pos = srcoff + (int)(cc * (CH * BPP)) + 128;
            if (pos >= lo_rel && pos < hi_rel)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(raw_s + (((cc & 1u) * 32u + lane) * UF_RAW + 32u) * 4u), "l"(T.src0 + pos));
        }
    };
    const int st_row = CH == 16 ? (int)(lane & 16u) : 0, st_px = CH == 16 ? (int)(lane & 15u) : (int)lane;

    uint32_t l0 = 0, l1 = 0, ul0 = 0, ul1 = 0; // rebuilt left pixel, upleft pixel
    // boundary pixels of the current sub-chunk (lanes 0..SUB-1), taken from the mailbox ring
    uint32_t c0 = 0, c1 = 0;
    volatile UfSlot *slot = T.slot;
    if (T.dep) { // post the request; the poller starts on it when it sees the new request number
        if (lane == 0) {
            slot->ent = (unsigned long long)reinterpret_cast<uintptr_t>(T.ent_in);
            slot->npix = T.npix;
            slot->tag = T.tag;
            slot->avail = 0;
            slot->cons = 0;
            __threadfence_block();
            slot->gen = T.gen;
        }
    }
    __syncwarp();
#pragma unroll 4
    for (int r = 0; r < 32; r++) fetch_row(raw_s + 4u * lane, 4 * (int)lane, r);
    fetch_own_tail(0);
    // one extra iteration: its steps are idle (every pixel is past the row's end), its stores drain the last chunk
    for (uint32_t c = 0; c <= nchunks; c++) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp(); // windows of chunk c complete, results of chunk c-1 visible, stores of chunk c-2 have read their buffer
        fetch_own_tail(c + 1);
        const uint32_t *lb = raw + ((c & 1u) * 32u + lane) * UF_RAW;
        uint32_t *ob0 = O + (c & 1u) * OBUF + lane * OS, *ob1 = ob0 + 32 * OS;
        const uint32_t *sb = O + ((c + 1u) & 1u) * OBUF + st_row * OS + st_px; // results of chunk c-1, this lane's share
        const uint32_t next_s = raw_s + (((c + 1u) & 1u) * 32u * UF_RAW + lane) * 4u;
        const int next_add = (int)((c + 1u) * (CH * BPP)) + 4 * (int)lane;
        const int st_add = ((int)(c * CH) - CH + st_px) * BPP; // chunk c-1 ("-1": nothing passes the range check)
        const int x0 = (int)(c * CH) - (int)T.delta; // this lane's pixel at the chunk's first step
#pragma unroll 1
        for (int q = 0; q < CH / SUB; q++) {
            if (T.dep) {
                const uint32_t xb = c * CH + q * SUB; // lane 0's row is at delta 0: step == pixel
                if (xb < T.npix) {
                    const uint32_t need = min(xb + SUB, T.npix);
                    while (slot->avail < need) {}
                    // (no fence: the ring words were written before `avail` by the same warp, and a warp's volatile
                    // shared-memory accesses are performed in program order)
                    if (lane < SUB) {
                        c0 = slot->ring0[(xb + lane) & (UF_RING - 1)];
                        if (N1) c1 = slot->ring1[(xb + lane) & (UF_RING - 1)];
                    }
                    __syncwarp();
                    if (lane == 0) slot->cons = need;
                }
            }
            // every shared-memory read of the sub-chunk up front (ptxas keeps LDS behind an earlier LDGSTS, so a load
            // inside the steps would sit right in front of its use): the filtered pixels of the sub-chunk and the result
            // words this lane stores for chunk c-1
            uint32_t d0[SUB], d1[SUB], s0[SUB], s1[SUB];
#pragma unroll
            for (int jj = 0; jj < SUB; jj++) {
                const int j = q * SUB + jj;
                const uint32_t o = mis + j * BPP; // bytes [o, o + BPP) of the window
                const uint32_t *wp = lb + (o >> 2);
                const uint32_t w0 = wp[0], w1 = wp[1];
                d0[jj] = __funnelshift_r(w0, w1, 8u * o);
                d1[jj] = N1 ? __funnelshift_r(w1, wp[2], 8u * o) : 0u;
                s0[jj] = sb[j * OS];
                s1[jj] = N1 ? sb[j * OS + 32 * OS] : 0u;
            }
#pragma unroll
            for (int jj = 0; jj < SUB; jj++) {
                const int j = q * SUB + jj;
                // pixel above: lane l-1's result of the previous step; lane 0: the boundary word
                uint32_t a0 = __shfl_up_sync(0xffffffffu, l0, 1);
                uint32_t a1 = N1 ? __shfl_up_sync(0xffffffffu, l1, 1) : 0u;
                const uint32_t b0 = __shfl_sync(0xffffffffu, c0, jj);
                const uint32_t b1 = N1 ? __shfl_sync(0xffffffffu, c1, jj) : 0u;
                if (lane == 0) { a0 = b0; a1 = b1; }
                const int x = x0 + j;
                const bool act = T.valid && (uint32_t)x < T.npix;
                const uint32_t keep = act ? ~0u : 0u;
                uint32_t p0 = (l0 & T.m1) | (a0 & T.m2), p1 = 0;
                if (N1) p1 = (l1 & T.m1) | (a1 & T.m2);
                if (MODE >= 1) {
                    p0 |= __vhaddu4(l0, a0) & T.m3;
                    if (N1) p1 |= __vhaddu4(l1, a1) & T.m3;
                }
                if (MODE >= 2) {
                    p0 |= paeth4(l0, a0, ul0) & T.m4;
                    if (N1) p1 |= paeth4(l1, a1, ul1) & T.m4;
                }
                l0 = vadd4_keep(d0[jj], p0, keep);
                if (N1) l1 = vadd4_keep(d1[jj], p1, keep);
                ul0 = a0; ul1 = a1;
                ob0[j] = l0;
                if (N1) ob1[j] = l1;
                if (publane && act) {
                    st_entry(T.ent_out + x, ((unsigned long long)(T.tag + 1) << 32) | l0);
                    if (N1) st_entry(T.ent_out + T.ent_words + x, ((unsigned long long)(T.tag + 1) << 32) | l1);
                }
                // off the chain: next chunk's row(s) in, previous chunk's row(s) out
                fetch_row(next_s, next_add, j);
                if (CH == 16) fetch_row(next_s, next_add, j + 16);
                const int rr = j + st_row;
                const int pos = __shfl_sync(0xffffffffu, dstoff, rr) + st_add;
                if ((uint32_t)(pos - rr * (int)T.len) < T.len) st_pixel<BPP>(T.dst0 + pos, s0[jj], s1[jj]);
            }
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
}

// the CTA's poller warp: see UfSlot
template <int BPP>
__device__ __noinline__ void uf_poller(volatile UfCtl *ctl, const uint64_t ent_words, const uint32_t backoff_ns, const uint32_t lane) {
    constexpr bool N1 = BPP > 4;
    uint32_t mygen[UF_WARPS] = {}, myavail[UF_WARPS] = {}, npix[UF_WARPS] = {}, tag[UF_WARPS] = {};
    const unsigned long long *ptr[UF_WARPS] = {};
    for (;;) {
        if (ctl->done >= (uint32_t)UF_WARPS) break;
        unsigned long long e0[UF_WARPS], e1[UF_WARPS];
        bool want[UF_WARPS], progress = false;
#pragma unroll
        for (int w = 0; w < UF_WARPS; w++) { // all requests first: the loads of the four mailboxes overlap
            volatile UfSlot *sl = &ctl->slot[w];
            const uint32_t g = sl->gen;
            if (g != mygen[w]) {
                __threadfence_block();
                ptr[w] = reinterpret_cast<const unsigned long long *>((uintptr_t)sl->ent);
                npix[w] = sl->npix;
                tag[w] = sl->tag;
                mygen[w] = g;
                myavail[w] = 0;
            }
            const uint32_t limit = min(npix[w], sl->cons + (uint32_t)UF_RING);
            const uint32_t x = myavail[w] + lane;
            want[w] = mygen[w] != 0 && x < limit;
            e0[w] = 0; e1[w] = 0;
            if (want[w]) {
                e0[w] = ld_entry(ptr[w] + x);
                if (N1) e1[w] = ld_entry(ptr[w] + ent_words + x);
            }
        }
#pragma unroll
        for (int w = 0; w < UF_WARPS; w++) {
            volatile UfSlot *sl = &ctl->slot[w];
            const bool ok = want[w] && (uint32_t)(e0[w] >> 32) == tag[w] && (!N1 || (uint32_t)(e1[w] >> 32) == tag[w]);
            const uint32_t nb = __ballot_sync(0xffffffffu, ok);
            const uint32_t n = nb == 0xffffffffu ? 32u : (uint32_t)__ffs((int)~nb) - 1u; // leading run of valid words
            if (n == 0) continue;
            const uint32_t x = myavail[w] + lane;
            if (lane < n) {
                sl->ring0[x & (UF_RING - 1)] = (uint32_t)e0[w];
                if (N1) sl->ring1[x & (UF_RING - 1)] = (uint32_t)e1[w];
            }
            __syncwarp();
            __threadfence_block();
            myavail[w] += n;
            if (lane == 0) sl->avail = myavail[w];
            progress = true;
        }
        if (backoff_ns && !progress) __nanosleep(backoff_ns);
    }
}

template <int BPP, int SUB>
__global__ void __launch_bounds__((UF_WARPS + 1) * 32) k_unfilter(Geom g, const uint8_t *__restrict__ in, uint8_t *out,
                                                            uint64_t n_images, uint64_t in_stride, ScanState *st,
                                                            unsigned long long *ent, UfPlan plan) {
    extern __shared__ __align__(16) uint32_t uf_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // a group of ctas_per_item CTAs walks the (image, pass) items group, group + #groups, ...: every warp of the group takes
    // the items in the same order and an item's tiles wait only on tiles of the same item, so a batch needs no more
    // co-resident CTAs than fit on the device
    const uint32_t group = blockIdx.x / plan.ctas_per_item, ngroups = gridDim.x / plan.ctas_per_item;
    const uint32_t gw = (blockIdx.x % plan.ctas_per_item) * UF_WARPS + warp, NW = plan.ctas_per_item * UF_WARPS;
    if (group >= ngroups) return;
    UfCtl *ctl = reinterpret_cast<UfCtl *>(uf_smem + UF_WARPS * uf_warp_words(BPP));
    for (uint32_t i = threadIdx.x; i < sizeof(UfCtl) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(ctl)[i] = 0;
    __syncthreads();
    if (warp == UF_WARPS) { // the poller warp
        uf_poller<BPP>(ctl, plan.ent_words, plan.backoff_ns, lane);
        return;
    }
    uint32_t mygen = 0;
    for (uint64_t item = group; item < n_images * g.n_pass; item += ngroups) {
    const uint64_t img = item / g.n_pass;
    const uint32_t pass = (uint32_t)(item % g.n_pass);
    const PassDesc pd = g.pass[pass];
    const uint32_t len = pd.len, npix = len / BPP;
    // filtered stream: row r of the image starts at in_base + r (one filter byte per earlier row), then 1 + len bytes
    const uint8_t *src = in + img * in_stride + pd.in_base + pd.row0;
    uint8_t *dst = out + img * g.data_len + pd.in_base;
    const uint32_t ntiles = (pd.nrows + 31) / 32;
    unsigned long long *ebase = ent + img * plan.ent_words * plan.ent_arrays + plan.ent_base[pass];
    UfTile T;
    T.in_lo = in;
    T.in_hi = in + n_images * in_stride;
    T.ent_words = plan.ent_words;
    T.npix = npix;
    T.len = len;
    for (uint32_t t = gw; t < ntiles; t += NW) {
        const uint32_t k = t * 32 + lane;
        T.valid = k < pd.nrows;
        T.src0 = src + (uint64_t)t * 32 * (len + 1);
        T.dst0 = dst + (uint64_t)t * 32 * len;
        uint32_t f = T.valid ? T.src0[(uint64_t)lane * (len + 1)] : 0;
        if (f > 4) { st->fail = 1; f = 0; } // RowFilter::try_from fails -> PngError::InvalidData
        T.m1 = f == 1 ? ~0u : 0u; T.m2 = f == 2 ? ~0u : 0u; T.m3 = f == 3 ? ~0u : 0u; T.m4 = f == 4 ? ~0u : 0u;
        // delta: rows since the last row that does not depend on the row above it
        const uint32_t heads = __ballot_sync(0xffffffffu, !T.valid || f <= 1 || lane == 0);
        T.delta = lane - (31u - (uint32_t)__clz((int)(heads & (0xffffffffu >> (31u - lane)))));
        T.maxd = __reduce_max_sync(0xffffffffu, T.valid ? T.delta : 0u);
        const uint32_t f_first = __shfl_sync(0xffffffffu, f, 0);
        T.dep = t > 0 && f_first >= 2; // the tile's first row needs tile t-1's last row (row 0 of a pass: zeros)
        T.pub = t + 1 < ntiles;
        T.lastlane = min(31u, pd.nrows - 1u - t * 32u);
        T.tag = t;
        T.slot = &ctl->slot[warp];
        if (T.dep) mygen++;
        T.gen = mygen;
        T.ent_out = ebase + (uint64_t)t * npix;
        T.ent_in = ebase + (uint64_t)(t ? t - 1 : 0) * npix;
        const bool any_avg = __any_sync(0xffffffffu, f == 3), any_paeth = __any_sync(0xffffffffu, f == 4);
        uint32_t *sm = uf_smem + warp * uf_warp_words(BPP);
#ifdef UF_TRACE
        unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#endif
        if (any_paeth) uf_run<BPP, 2, SUB>(T, sm, lane);
        else if (any_avg) uf_run<BPP, 1, SUB>(T, sm, lane);
        else uf_run<BPP, 0, SUB>(T, sm, lane);
#ifdef UF_TRACE
        unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (lane == 0 && t < 128) { st->hist[t] = t0; st->hist[128 + t] = t1; }
#endif
    }
    }
    __syncwarp();
    if (lane == 0) atomicAdd(&ctl->done, 1u);
}

// ---- host-side call scaffolding -------------------------------------------------------------------------------------------

// pageable host buffers of 24 MB and more are staged: the driver copies smaller ones at ~17 GB/s itself (c4: 16 MB images),
// larger ones at ~6 GB/s (measured on 64 MB: 23 ms per call against 9 ms staged)
constexpr size_t STAGE_MIN = (size_t)24 << 20;

struct Call {
    Trace trace{"oxi reduction scan"};
    DeviceCtx *c = nullptr;
    Slot *s = nullptr;
    cudaStream_t stream = nullptr; // the slot's own stream, or the caller's in device-resident mode
    bool resident = false;         // device-resident mode (host.h DevIO): data / out are device pointers
    int err = 0;
    ScanState *st = nullptr; // device
    uint8_t *d_in = nullptr, *d_out = nullptr;
    uint8_t *extra = nullptr; // scratch after ScanState (tables, LUTs)

    // `out` = the caller's output buffer (host, or device in device-resident mode; may be null when the call has no
    // image output)
    Call(const uint8_t *data, size_t len, uint8_t *out, size_t out_cap, size_t extra_bytes = 0) {
        resident = t_devio != nullptr;
        if (resident) {
            c = get_ctx(t_devio->device, &err);
            if (c && cudaSetDevice(c->device) != cudaSuccess) { err = fail(OXI_E_CUDA, "cudaSetDevice"); c = nullptr; }
        } else {
            c = current_ctx(&err);
        }
        if (!c) return;
        s = acquire_slot(c);
        if (!s) { err = fail(OXI_E_CUDA, "cannot create stream"); return; }
        stream = resident ? t_devio->stream : s->stream;
        cudaError_t e = slot_begin(s, stream);
        if (e == cudaSuccess && !resident) {
            e = grow(&s->d_in, &s->in_cap, len + 64);
            if (e == cudaSuccess) e = grow(&s->d_out, &s->out_cap, out_cap + 64);
        }
        if (e == cudaSuccess) e = grow(&s->scratch, &s->scratch_bytes, sizeof(ScanState) + extra_bytes + 256);
        if (e != cudaSuccess) { err = fail_cuda(e, "cudaMalloc"); return; }
        d_in = resident ? const_cast<uint8_t *>(data) : s->d_in;
        d_out = resident ? out : s->d_out;
        st = reinterpret_cast<ScanState *>(s->scratch);
        extra = s->scratch + ((sizeof(ScanState) + 255) & ~(size_t)255);
        e = cudaMemsetAsync(st, 0, sizeof(ScanState), stream);
        if (e == cudaSuccess && len && !resident) {
            // a large pageable buffer (what `&[u8]` gives the Rust side) goes through the slot's pinned buffer: the driver's
            // own staging of pageable copies runs at ~5 GB/s
            const uint8_t *src = data;
            if (len >= STAGE_MIN && is_pageable(data)) {
                e = grow_pinned(&s->h_in, &s->h_in_cap, len + 64);
                if (e == cudaSuccess) { par_memcpy(s->h_in, data, len); src = s->h_in; }
            }
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, src, len, cudaMemcpyHostToDevice, stream);
        }
        if (e != cudaSuccess) err = fail_cuda(e, "upload");
    }
    ~Call() {
        if (!s) return;
        if (resident) {
            release_slot_async(c, s, stream); // nothing synchronises: the scratch is reused in stream order
        } else {
            cudaStreamSynchronize(stream);
            release_slot(c, s);
        }
    }
    bool ok() const { return err == 0; }
    int state(ScanState *h) {
        cudaError_t e = cudaMemcpyAsync(h, st, sizeof(ScanState), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        return e == cudaSuccess ? 0 : fail_cuda(e, "download scan state");
    }
    // the image result: already in place in device-resident mode
    int download(uint8_t *out, size_t n) {
        if (resident) return 0;
        cudaError_t e = cudaSuccess;
        if (n >= STAGE_MIN && is_pageable(out)) {
            e = grow_pinned(&s->h_out, &s->h_out_cap, n + 64);
            if (e == cudaSuccess) e = cudaMemcpyAsync(s->h_out, d_out, n, cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
            if (e == cudaSuccess) par_memcpy(out, s->h_out, n);
            return e == cudaSuccess ? 0 : fail_cuda(e, "download");
        }
        if (n) e = cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        return e == cudaSuccess ? 0 : fail_cuda(e, "download");
    }
    // result = the input bytes unchanged
    int passthrough(uint8_t *out, const uint8_t *data, size_t n) {
        if (!resident) { memcpy(out, data, n); return 0; }
        cudaError_t e = cudaMemcpyAsync(out, data, n, cudaMemcpyDeviceToDevice, stream);
        return e == cudaSuccess ? 0 : fail_cuda(e, "device copy");
    }
    // small results (histograms, matrices) always go to host memory
    int to_host(void *dst, const void *d_src, size_t n) {
        cudaError_t e = cudaMemcpyAsync(dst, d_src, n, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        return e == cudaSuccess ? 0 : fail_cuda(e, "download");
    }
    int check(const char *what) {
        cudaError_t e = cudaGetLastError();
        count_launch();
        return e == cudaSuccess ? 0 : fail_cuda(e, what);
    }
};

bool has_alpha(uint8_t ct) { return ct == 4 || ct == 6; }
size_t byte_depth(const oxi_ihdr *h) { return h->bit_depth == 16 ? 2 : 1; }

// rows of the image with a per-pass output row length chosen by `out_len_of(pass)`
template <typename F>
bool make_row_map(const Geom &g, F out_len_of, RowMap *m) {
    memset(m, 0, sizeof *m);
    m->n_pass = g.n_pass;
    uint64_t ob = 0;
    for (uint32_t p = 0; p < g.n_pass; p++) {
        m->in_len[p] = g.pass[p].len;
        m->nrows[p] = g.pass[p].nrows;
        m->pixels[p] = g.pass[p].pixels;
        m->in_base[p] = g.pass[p].in_base;
        m->out_len[p] = out_len_of(g.pass[p]);
        m->out_base[p] = ob;
        m->out_items[p] = ob;
        ob += (uint64_t)m->nrows[p] * m->out_len[p];
    }
    for (uint32_t p = g.n_pass; p < 8; p++) m->out_items[p] = ob;
    return true;
}

void copy_aux(oxi_color_aux *dst, const oxi_color_aux *src) {
    if (!dst) return;
    if (src) *dst = *src;
    else memset(dst, 0, sizeof *dst);
}

// LUT remap of a device-resident image already uploaded by `call`
int run_lut(Call &call, size_t len, const uint8_t lut[256]) {
    cudaError_t e = cudaMemcpyAsync(call.extra, lut, 256, cudaMemcpyHostToDevice, call.stream);
    if (e != cudaSuccess) return fail_cuda(e, "upload LUT");
    k_lut<<<grid_for((len + 15) / 16, call.c->num_sms), RT, 0, call.stream>>>(call.d_in, call.d_out, len, call.extra);
    return call.check("k_lut");
}

int32_t luma_key(const uint8_t *c) { // palette.rs:92-105
    const int32_t a = c[3];
    return ((a & 0xFE) << 18) + (a & 0x01) - (int32_t)c[0] * 299 - (int32_t)c[1] * 587 - (int32_t)c[2] * 114;
}

} // namespace
} // namespace oxi

using namespace oxi;

extern "C" {

int oxi_cleaned_alpha_channel(const oxi_ihdr *h, const uint8_t *data, size_t len, uint8_t *out) {
    if (!valid_ihdr(h) || !data || !out) return fail(OXI_E_ARG, "bad arguments");
    if (!has_alpha(h->color_type)) return OXI_NONE; // alpha.rs:12-14
    const uint32_t bd = (uint32_t)byte_depth(h), bpp = (uint32_t)channels(h->color_type) * bd, cb = bpp - bd;
    const size_t used = len / bpp * bpp; // chunks_exact
    Call call(data, len, out, len);
    if (!call.ok()) return call.err;
    const uint64_t n16 = used / 16;
    k_clean_alpha<<<grid_for(n16 ? n16 : 1, call.c->num_sms), RT, 0, call.stream>>>(
        reinterpret_cast<const uint4 *>(call.d_in), reinterpret_cast<uint4 *>(call.d_out), n16, call.d_in + n16 * 16,
        call.d_out + n16 * 16, (uint32_t)(used - n16 * 16), bpp, cb);
    if (int rc = call.check("k_clean_alpha")) return rc;
    return call.download(out, used);
}

int oxi_reduced_alpha_channel(const oxi_ihdr *h, const uint8_t *data, size_t len, int optimize_alpha, uint8_t *out,
                              size_t *out_len, oxi_ihdr *oh, oxi_color_aux *oaux) {
    if (!valid_ihdr(h) || !data || !out) return fail(OXI_E_ARG, "bad arguments");
    if (!has_alpha(h->color_type)) return OXI_NONE;
    const uint32_t bd = (uint32_t)byte_depth(h), bpp = (uint32_t)channels(h->color_type) * bd, cb = bpp - bd;
    const uint64_t npix = len / bpp;
    Call call(data, len, out, npix * cb);
    if (!call.ok()) return call.err;
    k_alpha_scan<<<grid_for(npix, call.c->num_sms), RT, 0, call.stream>>>(call.d_in, npix, bpp, cb, optimize_alpha, call.st);
    if (int rc = call.check("k_alpha_scan")) return rc;
    ScanState st;
    if (int rc = call.state(&st)) return rc;
    if (st.fail) return OXI_NONE; // partially transparent pixel (alpha.rs:53-55)
    int trns = -1;
    if (st.has_transparency) { // alpha.rs:62-75
        auto used = [&](int v) { return (st.used[v >> 5] >> (v & 31)) & 1u; };
        if (h->color_type == 4) {
            const int tries[4] = {0x00, 0xFF, 0x55, 0xAA};
            for (int k = 0; k < 4 && trns < 0; k++)
                if (!used(tries[k])) trns = tries[k];
        }
        for (int v = 0; v < 256 && trns < 0; v++)
            if (!used(v)) trns = v;
        if (trns < 0) return OXI_NONE;
    }
    k_alpha_strip<<<grid_for(npix, call.c->num_sms), RT, 0, call.stream>>>(call.d_in, call.d_out, npix, bpp, cb, trns);
    if (int rc = call.check("k_alpha_strip")) return rc;
    if (int rc = call.download(out, npix * cb)) return rc;
    if (out_len) *out_len = npix * cb;
    if (oh) { *oh = *h; oh->color_type = h->color_type == 4 ? 0 : 2; }
    if (oaux) {
        memset(oaux, 0, sizeof *oaux);
        if (trns >= 0) { // alpha.rs:88-91
            const uint16_t t = h->bit_depth == 16 ? (uint16_t)((trns << 8) | trns) : (uint16_t)trns;
            oaux->has_trns = 1;
            oaux->trns[0] = oaux->trns[1] = oaux->trns[2] = t;
        }
    }
    return OXI_OK;
}

int oxi_reduced_bit_depth_16_to_8(const oxi_ihdr *h, const uint8_t *data, size_t len, int force_scale, uint8_t *out,
                                  size_t *out_len, oxi_ihdr *oh) {
    if (!valid_ihdr(h) || !data || !out) return fail(OXI_E_ARG, "bad arguments");
    if (h->bit_depth != 16) return OXI_NONE;
    const uint64_t ns = len / 2;
    Call call(data, len, out, ns);
    if (!call.ok()) return call.err;
    k_16to8<<<grid_for((ns + 7) / 8, call.c->num_sms), RT, 0, call.stream>>>(call.d_in, call.d_out, ns, force_scale ? 1 : 0,
                                                                                 call.st);
    if (int rc = call.check("k_16to8")) return rc;
    if (!force_scale) {
        ScanState st;
        if (int rc = call.state(&st)) return rc;
        if (st.fail) return OXI_NONE; // some sample has hi != lo (bit_depth.rs:19-22)
    }
    if (int rc = call.download(out, ns)) return rc;
    if (out_len) *out_len = ns;
    if (oh) { *oh = *h; oh->bit_depth = 8; }
    return OXI_OK;
}

int oxi_reduced_bit_depth_8_or_less(const oxi_ihdr *h, const oxi_color_aux *aux, const uint8_t *data, size_t len,
                                    uint8_t *out, size_t *out_len, oxi_ihdr *oh, oxi_color_aux *oaux) {
    if (!valid_ihdr(h) || !data || !out) return fail(OXI_E_ARG, "bad arguments");
    if (h->bit_depth != 8 || channels(h->color_type) != 1) return OXI_NONE;
    Geom g;
    if (!build_geom(h, len, 0, &g)) return fail(OXI_E_ARG, "invalid IHDR");
    Call call(data, len, out, len);
    if (!call.ok()) return call.err;
    uint32_t bits;
    if (h->color_type == 3) { // bit_depth.rs:73-80
        const uint32_t n = aux ? aux->n_palette : 0;
        if (n <= 2) bits = 1;
        else if (n <= 4) bits = 2;
        else if (n <= 16) bits = 4;
        else return OXI_NONE;
    } else {
        k_min_bits<<<grid_for((len + 15) / 16, call.c->num_sms), RT, 0, call.stream>>>(call.d_in, len, call.st);
        if (int rc = call.check("k_min_bits")) return rc;
        ScanState st;
        if (int rc = call.state(&st)) return rc;
        bits = st.max_bits ? st.max_bits : 1;
        if (bits >= 8) return OXI_NONE;
    }
    const uint32_t per = 8 / bits;
    RowMap m;
    make_row_map(g, [per](const PassDesc &pd) { return (pd.len + per - 1) / per; }, &m);
    const uint64_t total = m.out_items[g.n_pass];
    k_pack_bits<<<grid_for(total, call.c->num_sms), RT, 0, call.stream>>>(call.d_in, call.d_out, m, bits);
    if (int rc = call.check("k_pack_bits")) return rc;
    if (int rc = call.download(out, total)) return rc;
    if (out_len) *out_len = total;
    if (oh) { *oh = *h; oh->bit_depth = (uint8_t)bits; }
    if (oaux) {
        copy_aux(oaux, aux);
        if (h->color_type == 0 && aux && aux->has_trns) { // bit_depth.rs:134-153
            const uint16_t trans = aux->trns[0];
            const uint16_t reduced = (uint16_t)((trans & 0xFF) >> (8 - bits));
            uint16_t check = reduced;
            for (uint32_t b = bits; b < 8; b <<= 1) check = (uint16_t)((check << b) | check);
            oaux->has_trns = trans == check;
            oaux->trns[0] = trans == check ? reduced : 0;
        }
    }
    return OXI_OK;
}

int oxi_expanded_bit_depth_to_8(const oxi_ihdr *h, const oxi_color_aux *aux, const uint8_t *data, size_t len, uint8_t *out,
                                size_t *out_len, oxi_ihdr *oh, oxi_color_aux *oaux) {
    if (!valid_ihdr(h) || !data || !out) return fail(OXI_E_ARG, "bad arguments");
    const uint32_t depth = h->bit_depth;
    if (depth >= 8) return OXI_NONE;
    Geom g;
    if (!build_geom(h, len, 0, &g)) return fail(OXI_E_ARG, "invalid IHDR");
    RowMap m;
    make_row_map(g, [](const PassDesc &pd) { return pd.pixels; }, &m);
    const uint64_t total = m.out_items[g.n_pass];
    Call call(data, len, out, total);
    if (!call.ok()) return call.err;
    k_expand_bits<<<grid_for(total, call.c->num_sms), RT, 0, call.stream>>>(call.d_in, call.d_out, m, depth,
                                                                               h->color_type == 0);
    if (int rc = call.check("k_expand_bits")) return rc;
    if (int rc = call.download(out, total)) return rc;
    if (out_len) *out_len = total;
    if (oh) { *oh = *h; oh->bit_depth = 8; }
    if (oaux) {
        copy_aux(oaux, aux);
        if (h->color_type == 0 && aux && aux->has_trns) { // bit_depth.rs:206-220
            uint16_t trans = aux->trns[0];
            for (uint32_t b = depth; b < 8; b <<= 1) trans = (uint16_t)((trans << b) | trans);
            oaux->trns[0] = trans;
        }
    }
    return OXI_OK;
}

int oxi_reduced_rgb_to_grayscale(const oxi_ihdr *h, const oxi_color_aux *aux, const uint8_t *data, size_t len, uint8_t *out,
                                 size_t *out_len, oxi_ihdr *oh, oxi_color_aux *oaux) {
    if (!valid_ihdr(h) || !data || !out) return fail(OXI_E_ARG, "bad arguments");
    if (h->color_type != 2 && h->color_type != 6) return OXI_NONE;
    const uint32_t bd = (uint32_t)byte_depth(h), bpp = (uint32_t)channels(h->color_type) * bd, keep = bpp - 2 * bd;
    const uint64_t npix = len / bpp;
    Call call(data, len, out, npix * keep);
    if (!call.ok()) return call.err;
    k_rgb_to_gray<<<grid_for(npix, call.c->num_sms), RT, 0, call.stream>>>(call.d_in, call.d_out, npix, bpp, bd, call.st);
    if (int rc = call.check("k_rgb_to_gray")) return rc;
    ScanState st;
    if (int rc = call.state(&st)) return rc;
    if (st.fail) return OXI_NONE;
    if (int rc = call.download(out, npix * keep)) return rc;
    if (out_len) *out_len = npix * keep;
    if (oh) { *oh = *h; oh->color_type = h->color_type == 2 ? 0 : 4; }
    if (oaux) {
        memset(oaux, 0, sizeof *oaux);
        if (h->color_type == 2 && aux && aux->has_trns && aux->trns[0] == aux->trns[1] && aux->trns[1] == aux->trns[2]) {
            oaux->has_trns = 1; // color.rs:120-128
            oaux->trns[0] = aux->trns[0];
        }
    }
    return OXI_OK;
}

int oxi_reduced_to_indexed(const oxi_ihdr *h, const oxi_color_aux *aux, const uint8_t *data, size_t len, int allow_grayscale,
                           uint8_t *out, size_t *out_len, oxi_ihdr *oh, oxi_color_aux *oaux) {
    if (!valid_ihdr(h) || !data || !out) return fail(OXI_E_ARG, "bad arguments");
    if (h->bit_depth != 8 || h->color_type == 3) return OXI_NONE; // color.rs:39-44
    const bool gray = h->color_type == 0 || h->color_type == 4;
    if (!allow_grayscale && gray) return OXI_NONE;
    const uint32_t ch = (uint32_t)channels(h->color_type);
    const uint64_t npix = len / ch;
    Call call(data, len, out, npix, sizeof(ColorTable));
    if (!call.ok()) return call.err;
    ColorTable *tab = reinterpret_cast<ColorTable *>(call.extra);
    cudaError_t e = cudaMemsetAsync(tab->key, 0, sizeof tab->key, call.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(tab->first, 0xFF, sizeof tab->first, call.stream);
    if (e != cudaSuccess) return fail_cuda(e, "memset colour table");
    if (npix) {
        k_collect_colors<<<grid_for(npix, call.c->num_sms, 4), RT, 0, call.stream>>>(call.d_in, npix, ch, tab, call.st);
        if (int rc = call.check("k_collect_colors")) return rc;
    }
    ScanState st;
    if (int rc = call.state(&st)) return rc;
    if (st.overflow || st.n_distinct > 256) return OXI_NONE; // 257th distinct colour (color.rs:29-31)
    std::vector<unsigned long long> keys(GSLOTS), first(GSLOTS);
    e = cudaMemcpyAsync(keys.data(), tab->key, sizeof tab->key, cudaMemcpyDeviceToHost, call.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(first.data(), tab->first, sizeof tab->first, cudaMemcpyDeviceToHost, call.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(call.stream);
    if (e != cudaSuccess) return fail_cuda(e, "download colour table");
    std::vector<uint32_t> slots;
    for (uint32_t s = 0; s < GSLOTS; s++)
        if (keys[s]) slots.push_back(s);
    std::sort(slots.begin(), slots.end(), [&](uint32_t a, uint32_t b) { return first[a] < first[b]; }); // insertion order
    std::vector<uint32_t> index(GSLOTS, 0);
    for (size_t i = 0; i < slots.size(); i++) index[slots[i]] = (uint32_t)i;
    e = cudaMemcpyAsync(tab->index, index.data(), sizeof tab->index, cudaMemcpyHostToDevice, call.stream);
    if (e != cudaSuccess) return fail_cuda(e, "upload palette indices");
    if (npix) {
        const size_t smem = GSLOTS * 8 + GSLOTS;
        cudaFuncSetAttribute((const void *)k_index_colors, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_index_colors<<<grid_for(npix, call.c->num_sms, 4), RT, smem, call.stream>>>(call.d_in, call.d_out, npix, ch, tab);
        if (int rc = call.check("k_index_colors")) return rc;
    }
    if (int rc = call.download(out, npix)) return rc;
    if (out_len) *out_len = npix;
    if (oh) { *oh = *h; oh->color_type = 3; }
    if (oaux) {
        memset(oaux, 0, sizeof *oaux);
        oaux->n_palette = (uint32_t)slots.size();
        for (size_t i = 0; i < slots.size(); i++) {
            const uint32_t key = (uint32_t)keys[slots[i]];
            const uint8_t c0 = key & 255, c1 = (key >> 8) & 255, c2 = (key >> 16) & 255, c3 = key >> 24;
            uint8_t *p = oaux->palette[i];
            switch (h->color_type) {
            case 0: { // color.rs:51-62
                const bool t = aux && aux->has_trns && (uint8_t)aux->trns[0] == c0;
                p[0] = p[1] = p[2] = c0; p[3] = t ? 0 : 255;
                break;
            }
            case 2: { // color.rs:64-77
                const bool t = aux && aux->has_trns && (uint8_t)aux->trns[0] == c0 && (uint8_t)aux->trns[1] == c1 &&
                               (uint8_t)aux->trns[2] == c2;
                p[0] = c0; p[1] = c1; p[2] = c2; p[3] = t ? 0 : 255;
                break;
            }
            case 4: p[0] = p[1] = p[2] = c0; p[3] = c1; break;
            default: p[0] = c0; p[1] = c1; p[2] = c2; p[3] = c3; break;
            }
        }
    }
    return OXI_OK;
}

int oxi_used_histogram256(const uint8_t *data, size_t len, uint64_t hist[256]) {
    if (!data || !hist) return fail(OXI_E_ARG, "bad arguments");
    Call call(data, len, nullptr, 0);
    if (!call.ok()) return call.err;
    k_hist256<<<grid_for((len + 15) / 16, call.c->num_sms, 2), RT, 0, call.stream>>>(call.d_in, len, call.st);
    if (int rc = call.check("k_hist256")) return rc;
    ScanState st;
    if (int rc = call.state(&st)) return rc;
    for (int i = 0; i < 256; i++) hist[i] = st.hist[i];
    return OXI_OK;
}

int oxi_lut_remap(const uint8_t *data, size_t len, const uint8_t lut[256], uint8_t *out) {
    if (!data || !out || !lut) return fail(OXI_E_ARG, "bad arguments");
    Call call(data, len, out, len, 256);
    if (!call.ok()) return call.err;
    if (int rc = run_lut(call, len, lut)) return rc;
    return call.download(out, len);
}

int oxi_edge_color_counts(const oxi_ihdr *h, const uint8_t *data, size_t len, uint32_t counts[256]) {
    Geom g;
    if (!data || !counts || !build_geom(h, len, 0, &g)) return fail(OXI_E_ARG, "bad arguments");
    Call call(data, len, nullptr, 0);
    if (!call.ok()) return call.err;
    k_edge_counts<<<grid_for(g.total_rows, call.c->num_sms), RT, 0, call.stream>>>(call.d_in, g, call.st);
    if (int rc = call.check("k_edge_counts")) return rc;
    ScanState st;
    if (int rc = call.state(&st)) return rc;
    memcpy(counts, st.counts, sizeof st.counts);
    return OXI_OK;
}

int oxi_reduced_palette(const oxi_ihdr *h, const oxi_color_aux *aux, const uint8_t *data, size_t len, int optimize_alpha,
                        uint8_t *out, oxi_color_aux *oaux) {
    if (!valid_ihdr(h) || !data || !out) return fail(OXI_E_ARG, "bad arguments");
    if (h->bit_depth != 8 || h->color_type != 3 || !aux) return OXI_NONE; // palette.rs:13-18
    Call call(data, len, out, len, 256);
    if (!call.ok()) return call.err;
    k_hist256<<<grid_for((len + 15) / 16, call.c->num_sms, 2), RT, 0, call.stream>>>(call.d_in, len, call.st);
    if (int rc = call.check("k_hist256")) return rc;
    ScanState st;
    if (int rc = call.state(&st)) return rc;
    uint8_t condensed[256][4], byte_map[256];
    uint32_t nc = 0;
    bool did_change = false;
    memset(byte_map, 0, sizeof byte_map);
    for (int i = 0; i < 256; i++) { // palette.rs:28-39
        if (!st.hist[i]) continue;
        uint8_t c[4] = {0, 0, 0, 255}; // indices beyond the palette are opaque black
        if ((uint32_t)i < aux->n_palette) memcpy(c, aux->palette[i], 4);
        if (optimize_alpha && c[3] == 0) c[0] = c[1] = c[2] = 0; // add_color_to_set, palette.rs:64-73
        uint32_t idx = nc;
        for (uint32_t k = 0; k < nc; k++)
            if (!memcmp(condensed[k], c, 4)) { idx = k; break; }
        if (idx == nc) memcpy(condensed[nc++], c, 4);
        byte_map[i] = (uint8_t)idx;
        did_change |= idx != (uint32_t)i;
    }
    if (did_change) { // palette.rs:41-43
        if (int rc = run_lut(call, len, byte_map)) return rc;
        if (int rc = call.download(out, len)) return rc;
    } else if (nc != aux->n_palette) { // palette.rs:44-47: data unchanged, palette resized
        if (int rc = call.passthrough(out, data, len)) return rc;
    } else {
        return OXI_NONE;
    }
    if (oaux) {
        memset(oaux, 0, sizeof *oaux);
        oaux->n_palette = nc;
        memcpy(oaux->palette, condensed, (size_t)nc * 4);
    }
    return OXI_OK;
}

int oxi_sorted_palette(const oxi_ihdr *h, const oxi_color_aux *aux, const uint8_t *data, size_t len, uint8_t *out,
                       oxi_color_aux *oaux) {
    if (!valid_ihdr(h) || !data || !out) return fail(OXI_E_ARG, "bad arguments");
    if (h->bit_depth != 8 || h->color_type != 3 || !aux || aux->n_palette <= 1) return OXI_NONE; // palette.rs:78-84
    Geom g;
    if (!build_geom(h, len, 0, &g)) return fail(OXI_E_ARG, "invalid IHDR");
    const uint32_t n = aux->n_palette > 256 ? 256 : aux->n_palette;
    Call call(data, len, out, len, 256);
    if (!call.ok()) return call.err;
    k_edge_counts<<<grid_for(g.total_rows, call.c->num_sms), RT, 0, call.stream>>>(call.d_in, g, call.st);
    if (int rc = call.check("k_edge_counts")) return rc;
    ScanState st;
    if (int rc = call.state(&st)) return rc;
    // most_popular_edge_color, palette.rs:205-216: last maximum among the first n bins, None when not unique over all 256
    uint32_t best = 0;
    for (uint32_t i = 0; i < n; i++)
        if (st.counts[i] >= st.counts[best]) best = i;
    int equal = 0;
    for (int i = 0; i < 256; i++) equal += st.counts[i] == st.counts[best];
    const int keep_first = equal > 1 ? -1 : (int)best;
    std::vector<uint32_t> order;
    for (uint32_t i = 0; i < n; i++)
        if ((int)i != keep_first) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return luma_key(aux->palette[a]) < luma_key(aux->palette[b]);
    });
    if (keep_first >= 0) order.insert(order.begin(), (uint32_t)keep_first);
    bool identity = true;
    for (uint32_t i = 0; i < n; i++) identity &= order[i] == i;
    if (identity) return OXI_NONE;
    uint8_t byte_map[256];
    memset(byte_map, 0, sizeof byte_map);
    for (uint32_t i = 0; i < n; i++) byte_map[order[i]] = (uint8_t)i;
    if (int rc = run_lut(call, len, byte_map)) return rc;
    if (int rc = call.download(out, len)) return rc;
    if (oaux) {
        memset(oaux, 0, sizeof *oaux);
        oaux->n_palette = n;
        for (uint32_t i = 0; i < n; i++) memcpy(oaux->palette[i], aux->palette[order[i]], 4);
    }
    return OXI_OK;
}

int oxi_apply_palette_reorder(const oxi_ihdr *h, const oxi_color_aux *aux, const uint32_t *remapping, const uint8_t *data,
                              size_t len, uint8_t *out, oxi_color_aux *oaux) {
    if (!valid_ihdr(h) || !data || !out || !remapping) return fail(OXI_E_ARG, "bad arguments");
    if (h->color_type != 3 || !aux) return OXI_NONE;
    const uint32_t n = aux->n_palette > 256 ? 256 : aux->n_palette;
    bool identity = true;
    for (uint32_t i = 0; i < n; i++) identity &= remapping[i] == i;
    if (identity) return OXI_NONE; // palette.rs:171-175
    uint8_t byte_map[256];
    memset(byte_map, 0, sizeof byte_map);
    oxi_color_aux na;
    memset(&na, 0, sizeof na);
    na.n_palette = n;
    for (uint32_t i = 0; i < n; i++) {
        if (remapping[i] >= n) return fail(OXI_E_ARG, "remapping entry beyond the palette");
        memcpy(na.palette[i], aux->palette[remapping[i]], 4);
        byte_map[remapping[i]] = (uint8_t)i;
    }
    Call call(data, len, out, len, 256);
    if (!call.ok()) return call.err;
    if (int rc = run_lut(call, len, byte_map)) return rc;
    if (int rc = call.download(out, len)) return rc;
    if (oaux) *oaux = na;
    return OXI_OK;
}


static bool make_adam7(const oxi_ihdr *h, Adam7 *a) {
    static const uint32_t x0[7] = {0, 4, 0, 2, 0, 1, 0}, y0[7] = {0, 0, 4, 0, 2, 0, 1};
    static const uint32_t dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
    memset(a, 0, sizeof *a);
    a->width = h->width; a->height = h->height;
    a->bits = (uint32_t)(h->bit_depth * channels(h->color_type));
    a->full_stride = (uint32_t)(((uint64_t)h->width * a->bits + 7) / 8);
    uint64_t off = 0;
    for (int p = 0; p < 7; p++) {
        a->base[p] = off;
        if (h->width > x0[p] && h->height > y0[p]) {
            a->pw[p] = (h->width - x0[p] + dx[p] - 1) / dx[p];
            a->ph[p] = (h->height - y0[p] + dy[p] - 1) / dy[p];
            a->stride[p] = (uint32_t)(((uint64_t)a->pw[p] * a->bits + 7) / 8);
            off += (uint64_t)a->ph[p] * a->stride[p];
        } else {
            a->stride[p] = 1; // never addressed
        }
    }
    a->base[7] = off;
    return true;
}

int oxi_interlace_image(const oxi_ihdr *h, const uint8_t *data, size_t len, uint8_t *out, size_t *out_len) {
    if (!valid_ihdr(h) || !data || !out) return fail(OXI_E_ARG, "bad arguments");
    Adam7 a;
    make_adam7(h, &a);
    if (len < (size_t)a.full_stride * h->height) return fail(OXI_E_ARG, "data shorter than the progressive image");
    Call call(data, len, out, a.base[7]);
    if (!call.ok()) return call.err;
    k_interlace<<<grid_for(a.base[7], call.c->num_sms), RT, 0, call.stream>>>(call.d_in, call.d_out, a);
    if (int rc = call.check("k_interlace")) return rc;
    if (int rc = call.download(out, a.base[7])) return rc;
    if (out_len) *out_len = a.base[7];
    return OXI_OK;
}

int oxi_deinterlace_image(const oxi_ihdr *h, const uint8_t *data, size_t len, uint8_t *out, size_t *out_len) {
    if (!valid_ihdr(h) || !data || !out) return fail(OXI_E_ARG, "bad arguments");
    Adam7 a;
    make_adam7(h, &a);
    if (len < a.base[7]) return fail(OXI_E_ARG, "data shorter than the interlaced image");
    const size_t total = (size_t)a.full_stride * h->height;
    Call call(data, len, out, total);
    if (!call.ok()) return call.err;
    k_deinterlace<<<grid_for(total, call.c->num_sms), RT, 0, call.stream>>>(call.d_in, call.d_out, a);
    if (int rc = call.check("k_deinterlace")) return rc;
    if (int rc = call.download(out, total)) return rc;
    if (out_len) *out_len = total;
    return OXI_OK;
}


// n_images filtered streams of `len` bytes each, back to back -> n_images x data_len bytes of pixels, back to back
static int unfilter_streams(const oxi_ihdr *h, const uint8_t *filtered, size_t len, size_t n_images, uint8_t *out, size_t *out_len) {
    if (!valid_ihdr(h) || !filtered || !out || n_images == 0) return fail(OXI_E_ARG, "bad arguments");
    // geometry of the unfiltered data a stream holds: every scanline carries one extra byte
    // (ScanLines with has_filter = true, png/scan_lines.rs:159-163)
    Geom g;
    size_t lo = 0, hi = len; // find data_len such that data_len + rows(data_len) is the largest value <= len
    if (!build_geom(h, len, 0, &g)) return fail(OXI_E_ARG, "invalid IHDR");
    while (lo < hi) {
        const size_t mid = lo + (hi - lo + 1) / 2;
        build_geom(h, mid, 0, &g);
        if (g.out_len <= len) lo = mid; else hi = mid - 1;
    }
    build_geom(h, lo, 0, &g);
    if (g.total_rows == 0) { if (out_len) *out_len = 0; return OXI_OK; }
    if (g.max_len >= UF_MAX_ROW) return fail(OXI_E_UNSUPPORTED, "scanlines of 16 MB and more are not supported");
    if (t_devio && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(filtered)) & 15u))
        return fail(OXI_E_ARG, "device buffers must be 16-byte aligned");
    // boundary words: one per pixel and tile (6/8-byte pixels: two arrays), zeroed so that no tag matches yet
    UfPlan plan{};
    uint32_t max_tiles = 1;
    for (uint32_t p = 0; p < g.n_pass; p++) {
        const uint32_t ntiles = (g.pass[p].nrows + 31) / 32;
        plan.ent_base[p] = plan.ent_words;
        plan.ent_words += (uint64_t)ntiles * (g.pass[p].len / g.bpp);
        max_tiles = std::max(max_tiles, ntiles);
    }
    plan.ent_arrays = g.bpp > 4 ? 2 : 1;
    const size_t ent_bytes = (size_t)plan.ent_words * plan.ent_arrays * n_images * sizeof(unsigned long long);
    Call call(filtered, len * n_images, out, (size_t)g.data_len * n_images, ent_bytes);
    if (!call.ok()) return call.err;
    unsigned long long *ent = reinterpret_cast<unsigned long long *>(call.extra);
    if (cudaMemsetAsync(ent, 0, ent_bytes, call.stream) != cudaSuccess) return fail(OXI_E_CUDA, "memset(boundary words)");
    {
        const void *fn = nullptr;
        const bool wide_body = n_images == 1; // one image: a lone warp per scheduler (see uf_run, SUB)
        switch (g.bpp) {
        case 1: fn = wide_body ? (const void *)k_unfilter<1, 16> : (const void *)k_unfilter<1, 8>; break;
        case 2: fn = wide_body ? (const void *)k_unfilter<2, 16> : (const void *)k_unfilter<2, 8>; break;
        case 3: fn = wide_body ? (const void *)k_unfilter<3, 16> : (const void *)k_unfilter<3, 8>; break;
        case 4: fn = wide_body ? (const void *)k_unfilter<4, 16> : (const void *)k_unfilter<4, 8>; break;
        case 6: fn = (const void *)k_unfilter<6, 8>; break;
        case 8: fn = (const void *)k_unfilter<8, 8>; break;
        default: return fail(OXI_E_ARG, "unsupported filter bpp");
        }
        // always the same value for a given instantiation: concurrent callers cannot shrink each other's limit
        const size_t smem = (size_t)UF_WARPS * uf_warp_words((int)g.bpp) * sizeof(uint32_t) + sizeof(UfCtl);
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int per_sm = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, (UF_WARPS + 1) * 32, smem);
        if (e != cudaSuccess || per_sm < 1) return fail_cuda(e, "k_unfilter occupancy");
        // The tiles of an (image, pass) item wait on each other across CTAs: a cooperative launch guarantees that the
        // whole grid is resident even when other streams keep the device busy.  One image: one CTA per SM (one pipeline
        // warp per scheduler).  A batch: as many CTA groups as fit, each walking its share of the items.
        const uint32_t capacity = (uint32_t)per_sm * (uint32_t)call.c->num_sms;
        const uint64_t n_items = (uint64_t)n_images * g.n_pass;
        const uint32_t want = (max_tiles + UF_WARPS - 1) / UF_WARPS;
        const uint32_t room = n_items == 1 ? (uint32_t)call.c->num_sms : capacity;
        plan.ctas_per_item = std::max(1u, std::min(want, n_items <= g.n_pass ? room / g.n_pass : room));
        const uint32_t groups = (uint32_t)std::min<uint64_t>(n_items, std::max(1u, capacity / plan.ctas_per_item));
        plan.backoff_ns = groups > g.n_pass ? 500u : 0u;
        uint64_t n = n_images, stride = len;
        const uint8_t *din = call.d_in;
        uint8_t *dout = call.d_out;
        ScanState *sst = call.st;
        void *args[] = {&g, &din, &dout, &n, &stride, &sst, &ent, &plan};
        e = cudaLaunchCooperativeKernel(fn, dim3(groups * plan.ctas_per_item), dim3((UF_WARPS + 1) * 32), args, smem, call.stream);
        if (e != cudaSuccess) return fail_cuda(e, "cudaLaunchCooperativeKernel(k_unfilter)");
    }
    if (int rc = call.check("k_unfilter")) return rc;
    ScanState st;
    if (int rc = call.state(&st)) return rc;
#ifdef UF_TRACE
    if (getenv("OXI_UF_TRACE")) {
        for (int t = 0; t < 128; t++)
            fprintf(stderr, "tile %d start %.2f end %.2f us\n", t, (double)(st.hist[t] - st.hist[0]) / 1e3, (double)(st.hist[128 + t] - st.hist[0]) / 1e3);
    }
#endif
    if (st.fail) return fail(OXI_E_DATA, "invalid filter type byte (PngError::InvalidData)");
    if (int rc = call.download(out, (size_t)g.data_len * n_images)) return rc;
    if (out_len) *out_len = g.data_len;
    return OXI_OK;
}

int oxi_unfilter_image(const oxi_ihdr *h, const uint8_t *filtered, size_t len, uint8_t *out, size_t *out_len) {
    return unfilter_streams(h, filtered, len, 1, out, out_len);
}
int oxi_unfilter_batch(const oxi_ihdr *h, const uint8_t *filtered, size_t len, size_t n_images, uint8_t *out, size_t *out_len) {
    return unfilter_streams(h, filtered, len, n_images, out, out_len);
}

int oxi_cooccurrence_build(const oxi_ihdr *h, uint32_t n, const uint8_t *data, size_t len, uint32_t *matrix) {
    Geom g;
    if (!data || !matrix || n == 0 || n > 256 || !build_geom(h, len, 0, &g)) return fail(OXI_E_ARG, "bad arguments");
    if (h->bit_depth != 8) return fail(OXI_E_ARG, "co-occurrence needs 8-bit indices");
    const size_t mbytes = (size_t)n * n * 4;
    Call call(data, len, nullptr, 0, mbytes);
    if (!call.ok()) return call.err;
    uint32_t *m = reinterpret_cast<uint32_t *>(call.extra);
    cudaError_t e = cudaMemsetAsync(m, 0, mbytes, call.stream);
    if (e != cudaSuccess) return fail_cuda(e, "memset matrix");
    const uint64_t items = ((uint64_t)g.data_len + 15) / 16 + g.total_rows;
    k_cooccurrence<<<grid_for(items, call.c->num_sms), RT, 0, call.stream>>>(call.d_in, g, n, m, call.st);
    if (int rc = call.check("k_cooccurrence")) return rc;
    k_symmetrise<<<n, 256, 0, call.stream>>>(m, n);
    if (int rc = call.check("k_symmetrise")) return rc;
    ScanState st;
    if (int rc = call.state(&st)) return rc;
    if (st.fail) return fail(OXI_E_ARG, "pixel index beyond the palette (run reduced_palette first, palette.rs:533-536)");
    return call.to_host(matrix, m, mbytes);
}

} // extern "C"
