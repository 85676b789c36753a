// filter_fast.cu -- fast tier of PngImage::filter_image (src/png/mod.rs:355-431) for sm_100a.
//
// One kernel family, `k_fast<D, SC>`.  A CTA takes a *band* of consecutive scanlines of one pass of one image and
// walks it row by row; rows are staged in a 3- or 4-slot shared-memory ring by 1-D TMA bulk copies (cp.async.bulk +
// mbarrier) so the next row streams in while the current one is scored, and the previous row stays resident
// (each row is read from HBM once, plus one halo row per band).  Per row:
//   1. all five RowFilters are applied with byte-SIMD arithmetic (src/filters/mod.rs:63-110; Paeth :242-254 in a
//      branch-free byte form, common.cuh) and scored:
//        MinSum  (strategies.rs:76-101)   VABSDIFF4 + accumulate: |r as i8| = 128 - ||cur-pred| - 128|
//        Entropy (strategies.rs:105-149)  five 256-bin shared-memory histograms x 4 lane replicas; one prmt + one
//                                         red.shared per residual byte
//        Bigrams (strategies.rs:153-192)  } rows <= 16 KB: dense 32x32 tables of u32 counters per trial for windows whose
//        BigEnt  (strategies.rs:195-227)  } two residuals lie in [-16,15] (one prmt + one red.shared per window), small
//                                           hashes for the outliers, the None trial on coarse keys whose bounds usually
//                                           rule it out (else an exact hash pass), branch-free table sweeps -- see
//                                           bigram_row_small.  Longer rows: one 65536 x u16 table, trial by trial.
//   2. the winner per strategy is picked with the reference's strict comparisons in RowFilter::ALL order
//      (lowest filter id wins ties; all-zero rows short-circuit to None, png/mod.rs:398-404),
//   3. only the chosen filtered row is written: recomputed in the *output-aligned* frame so that every store is a
//      16-byte st.global.v4 (rows of the filtered stream start at r*(len+1), never aligned).
// Requirements: alpha_bytes == 0 (rows independent given the previous raw row) and 3 staged rows (+ tables) must
// fit in 227 KB of shared memory; everything else goes to the generic tier.
#include <map>
#include <mutex>

#include "common.cuh"

namespace oxi {

namespace {

constexpr int SC_MINSUM = 1, SC_ENTROPY = 2, SC_BIGRAM = 4;
constexpr int SC_HALF = 8; // with SC_BIGRAM alone, rows <= 4096 bytes: 256-thread CTAs, four resident per SM
constexpr int HALF_NT = 256, HALF_OCC = 4;
// with SC_BIGRAM | SC_HALF: dense tables of 64 x 64 counters per trial (residuals in [-32, 31]) for images whose filtered rows
// are spread wider than +-16 (film grain, dithering): 512-thread CTAs, two resident per SM
constexpr int SC_WIDE = 16;
__host__ __device__ constexpr int rb_of(int sc) { return (sc & SC_WIDE) ? 6 : 5; } // bits of a range code
constexpr int HIST_REP = 1;      // copies of each Entropy histogram: one (lanes on the same bin are merged by ATOMS.POPC.INC)
// an Entropy histogram has 512 counters: counter 256 + d for the signed byte difference d = cur - pred in [-255, 255];
// the residual value v = d mod 256 is the sum of counters v and v + 256 (entropy_score)
constexpr uint32_t HIST_BINS = 512, HIST_BYTES = 5 * HIST_BINS * 4;
constexpr uint32_t PAD = 32;     // zero bytes in front of every staged row (left/upleft of the first pixel)
constexpr uint32_t SMEM_HDR = 1024;
constexpr uint32_t SCORE_WORDS = 40;   // s_min/s_ent/s_bgc/s_bge/s_out, 8 words each; three sets rotate row by row
constexpr uint32_t HSLOTS = 256;     // slots of a trial's outlier hash table (key << 16 | count)
constexpr int HPROBE = 8;            // probes before an insert gives up and sends the trial's outliers to the big table
constexpr uint32_t SMEM_LIMIT = 232448; // 227 KB

struct FastParams {
    Geom g;
    StrategySet set;
    const uint8_t *data;
    uint64_t n_images;
    uint32_t band_start[8]; // prefix sum of bands per pass
    uint32_t items_per_image;
    uint32_t rb;    // bytes of one ring slot
    uint32_t shift; // 8 * (bpp % 4): funnel-shift amount that turns a row word into its "left" word
    uint32_t nslot;     // ring slots (4 when shared memory allows: removes the end-of-row barrier; else 3)
    uint32_t h0_slots;  // bigram scorer with small_tab: slots of the trial-0 hash table (power of two, > 2 * row bytes)
    const uint8_t *img_class; // optional per-image class written by k_spread_probe (0: residuals within +-16, 1: wider)
    uint32_t want_class;      // the class this launch takes
    uint32_t four;      // the constant 4, opaque to the compiler (scale4)
    uint32_t two31;     // the constant 2^31, opaque to the compiler (havg4)
    uint32_t small_tab; // bigram scorer: 1 = the dense-table layout (rows <= 16384 bytes), 2 = the split-table layout (<= 32768)
};

__host__ __device__ constexpr int threads_for(int sc) { return (sc & SC_WIDE) ? 512 : (sc & SC_HALF) ? HALF_NT : (sc & SC_BIGRAM) ? 1024 : sc == SC_ENTROPY ? 512 : 256; }

// ---- PTX wrappers: mbarrier + 1-D bulk tensor-memory-accelerator copy -----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\t"
                 "bra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- byte-SIMD helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { // full prmt (sign-replicate nibbles kept)
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ void smem_inc(uint32_t addr) { // no-return add of 1 on a shared-window byte address
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(1u) : "memory");
}
// The same with the table's shared-window address as an immediate of the instruction (ATOMS [R + imm]).  The dynamic
// shared memory of a kernel without static shared memory that is not launched in a cluster starts at shared-window
// address SMEM_WIN on sm_100 (the first KB of the window is the system's); k_fast checks it at entry and traps otherwise,
// so a different driver layout fails loudly instead of counting in the wrong place.
constexpr uint32_t SMEM_WIN = 0x400;
template <uint32_t IMM>
__device__ __forceinline__ void smem_inc_at(uint32_t off) {
    asm volatile("red.shared.add.u32 [%0+%1], %2;" ::"r"(off), "n"(IMM), "r"(1u) : "memory");
}
template <int F>
__device__ __forceinline__ uint32_t pred4(uint32_t left, uint32_t up, uint32_t ul) {
    if (F == 0) return 0u;
    if (F == 1) return left;
    if (F == 2) return up;
    if (F == 3) return __vhaddu4(left, up); // floor((left+up)/2), filters/mod.rs:93
    return paeth4(left, up, ul);
}
__device__ __forceinline__ uint32_t pred4_rt(int f, uint32_t left, uint32_t up, uint32_t ul) {
    switch (f) {
    case 0: return 0u;
    case 1: return left;
    case 2: return up;
    case 3: return __vhaddu4(left, up);
    default: return paeth4(left, up, ul);
    }
}

// A 16-byte chunk of a staged row plus the 8 bytes before it: w[j] is row word (4*chunk + j - 2).
struct Chunk {
    uint32_t w[6];
};
__device__ __forceinline__ Chunk load_chunk(const uint8_t *row, uint32_t c) {
    Chunk k;
    const uint2 lo = *reinterpret_cast<const uint2 *>(row + 16 * (size_t)c - 8);
    const uint4 hi = *reinterpret_cast<const uint4 *>(row + 16 * (size_t)c);
    k.w[0] = lo.x; k.w[1] = lo.y; k.w[2] = hi.x; k.w[3] = hi.y; k.w[4] = hi.z; k.w[5] = hi.w;
    return k;
}
// word k (0..3) of the chunk shifted right by bpp bytes: the "left" (or "upleft") neighbours of its four bytes.
// D = bpp / 4, s = 8 * (bpp % 4).
template <int D>
__device__ __forceinline__ uint32_t left_word(const Chunk &c, int k, uint32_t s) {
    if (D == 2) return c.w[k];
    if (D == 1) return __funnelshift_l(c.w[k], c.w[k + 1], s);
    return __funnelshift_l(c.w[k + 1], c.w[k + 2], s);
}
__device__ __forceinline__ uint32_t tail_mask(int valid_bytes) { // bytes of a word that lie inside the row
    return valid_bytes >= 4 ? 0xFFFFFFFFu : (valid_bytes <= 0 ? 0u : ((1u << (8 * valid_bytes)) - 1u));
}

// 16 bytes starting at an arbitrary byte offset of a staged row (shared memory): the two aligned 16-byte granules that
// contain them (conflict-free 128-bit loads; a slot has >= 32 bytes of padding on both sides) and funnel shifts.  The
// misalignment is the same for every thread of a write pass, so the word select is a uniform branch.
__device__ __forceinline__ uint4 ld_unaligned16_wide(const uint8_t *p) {
    const uint32_t a = smem_u32(p);
    const uint32_t s = (a & 3u) * 8u;
    const uint4 *q = reinterpret_cast<const uint4 *>(p - (a & 15u));
    const uint4 A = q[0], B = q[1];
    uint4 r;
    switch ((a >> 2) & 3u) {
    case 0:
        r.x = __funnelshift_r(A.x, A.y, s); r.y = __funnelshift_r(A.y, A.z, s);
        r.z = __funnelshift_r(A.z, A.w, s); r.w = __funnelshift_r(A.w, B.x, s);
        break;
    case 1:
        r.x = __funnelshift_r(A.y, A.z, s); r.y = __funnelshift_r(A.z, A.w, s);
        r.z = __funnelshift_r(A.w, B.x, s); r.w = __funnelshift_r(B.x, B.y, s);
        break;
    case 2:
        r.x = __funnelshift_r(A.z, A.w, s); r.y = __funnelshift_r(A.w, B.x, s);
        r.z = __funnelshift_r(B.x, B.y, s); r.w = __funnelshift_r(B.y, B.z, s);
        break;
    default:
        r.x = __funnelshift_r(A.w, B.x, s); r.y = __funnelshift_r(B.x, B.y, s);
        r.z = __funnelshift_r(B.y, B.z, s); r.w = __funnelshift_r(B.z, B.w, s);
        break;
    }
    return r;
}

// The same via five 32-bit loads (4-way bank conflicts, but no branch): cheaper where shared-memory bandwidth is not
// the limiter.  WIDE is chosen per kernel (the Entropy scorer is bound by shared-memory wavefronts).
template <bool WIDE>
__device__ __forceinline__ uint4 ld_unaligned16(const uint8_t *p) {
    if (WIDE) return ld_unaligned16_wide(p);
    const uint32_t a = smem_u32(p);
    const uint32_t s = (a & 3u) * 8u;
    const uint32_t *q = reinterpret_cast<const uint32_t *>(p - (a & 3u));
    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3], w4 = q[4];
    uint4 r;
    r.x = __funnelshift_r(w0, w1, s);
    r.y = __funnelshift_r(w1, w2, s);
    r.z = __funnelshift_r(w2, w3, s);
    r.w = __funnelshift_r(w3, w4, s);
    return r;
}

// scalar residual of row byte i (i >= 0) under filter f, reading the staged rows (left pads are zero)
__device__ __forceinline__ uint32_t residual_at(int f, const uint8_t *cur, const uint8_t *up, int i, int bpp) {
    return residual_byte(f, cur[i], cur[i - bpp], up[i], up[i - bpp]);
}

// Bigram tables are indexed by (first byte, second byte) with two u16 counters per 32-bit word.  The word index is
// XOR-swizzled with the first byte so that the bank depends on both bytes: runs like "(x, 0)" (every pixel's alpha
// residual) would otherwise put a whole warp on one bank.
__device__ __forceinline__ uint32_t big_word(uint32_t bg) { return (bg >> 1) ^ ((bg >> 8) & 31u); }
// dense tables: 32x32 keys as u32 counters, 2 lane-selected replicas each.  Trials 1..4: both bytes in [-16,15];
// trial 0: the low five bits of both raw bytes (a coarsening that bounds the trial's scores, see bigram_row_small)
// Geometry of the dense-table region for range codes of RB bits (code = residual + 2^(RB-1)): the four trial tables, the
// coarse table of trial 0 (8 KB, every other 128-byte line used), then the outlier hashes.  With RB = 5 the region is kept
// at 36 KB so that the hash of whole rows (h0, 32 KB for 4 KB rows) can live in it while the tables are idle.
template <int RB>
struct Dense {
    static constexpr uint32_t HALF = 1u << (RB - 1);
    static constexpr uint32_t OUT = RB == 5 ? 0xE0u : 0xC0u;       // bits that are set in a code outside the range
    static constexpr uint32_t OUT4 = OUT * 0x01010101u;
    static constexpr uint32_t TABLES = 16u << (2 * RB);            // bytes of the four trial tables: 16 KB / 64 KB
    static constexpr uint32_t COARSE = TABLES;                     // byte offset of the coarse table
    static constexpr uint32_t REGION = RB == 5 ? 36864u : TABLES + 8192u;
    static constexpr uint32_t SMALL = REGION + 4 * HSLOTS * 4;     // ... + outlier hashes
};

// ---- shared-memory carve-up --------------------------------------------------------------------------------------
struct Smem {
    uint64_t *bar;     // 3 mbarriers
    uint32_t *sc;      // score set of the current row: s_min[8] s_ent[8] s_bgc[8] s_bge[8] s_out[8]
    __device__ __forceinline__ uint32_t *s_min_() const { return sc; }      // [5] MinSum sad accumulators
    __device__ __forceinline__ uint32_t *s_ent_() const { return sc + 8; }  // [5]
    __device__ __forceinline__ uint32_t *s_bgc_() const { return sc + 16; } // [5] distinct bigrams
    __device__ __forceinline__ uint32_t *s_bge_() const { return sc + 24; } // [5] bigram entropy
    __device__ __forceinline__ uint32_t *s_out_() const { return sc + 32; } // [5] trial flagged: its outliers overflowed the hash table
    uint8_t *rows;     // 3 slots of rb bytes
    uint32_t *hist;    // [5][HIST_BINS]
    uint32_t *table;   // 32768 words = 65536 u16 counters, or (small_tab) the h0_slots-word hash table of whole rows
    uint32_t *small;   // DENSE_WORDS counters of the in-range bigrams of trials 1..4 (layout: dense_off), optional
    uint32_t *ohash;   // [4][HSLOTS] outlier bigrams of trials 1..4: open-addressing tables of (key << 16 | count)
};
// h0_slots != 0 selects the small-table layout: header | histograms | dense tables | outlier hashes | h0 | row ring.
// The tables come first so that their shared-memory offsets are compile-time constants (immediates of the atomics).
// Otherwise: header | histograms | row ring | 65536-entry bigram table.
constexpr uint32_t SPLIT_SET = 65536;    // split layout: bytes of one table set -- A (32 x 256 u32), then B (256 x 32 u32)
constexpr uint32_t SPLIT_OSLOTS = 128;   // slots of a set's outlier hash
constexpr uint32_t SPLIT_BYTES = 2 * SPLIT_SET + 2 * SPLIT_OSLOTS * 4;
template <int SC>
__device__ __forceinline__ Smem carve(uint8_t *base, uint32_t rb, uint32_t nslot, uint32_t h0_slots, bool split = false) {
    Smem s;
    if (split) { // header | two table sets (= the 65536 x u16 table of the exact trials) | two outlier hashes | row ring
        s.bar = reinterpret_cast<uint64_t *>(base);
        s.sc = reinterpret_cast<uint32_t *>(base + 64);
        s.hist = s.small = s.table = reinterpret_cast<uint32_t *>(base + SMEM_HDR);
        s.ohash = s.table + 2 * SPLIT_SET / 4;
        s.rows = base + SMEM_HDR + SPLIT_BYTES;
        return s;
    }
    s.bar = reinterpret_cast<uint64_t *>(base);
    uint32_t *u = reinterpret_cast<uint32_t *>(base + 64);
    s.sc = u; // set 0; sets 1 and 2 follow at SCORE_WORDS strides
    uint8_t *p = base + SMEM_HDR;
    s.hist = reinterpret_cast<uint32_t *>(p); // histograms first: fixed offset, immediates of the atomics
    if (SC & SC_ENTROPY) p += HIST_BYTES;
    s.small = reinterpret_cast<uint32_t *>(p);
    s.ohash = s.small + Dense<rb_of(SC)>::REGION / 4;
    s.table = s.ohash + 4 * HSLOTS;
    if ((SC & SC_BIGRAM) && h0_slots) p += Dense<rb_of(SC)>::SMALL + ((SC & SC_HALF) ? 0u : h0_slots * 4);
    // SC_HALF: h0 (<= 8192 slots) is only used while the dense tables are swept clean and idle (exact pass of trial 0,
    // flagged outliers) and leaves them clean: it lives in their memory
    if (SC & SC_HALF) s.table = s.small;
    s.rows = p;
    p += nslot * (size_t)rb;
    if (!((SC & SC_BIGRAM) && h0_slots)) s.table = reinterpret_cast<uint32_t *>(p);
    return s;
}
// h0_slots != 0: the small-table layout (hash of whole rows + dense tables + outlier hashes); 0: the 65536-entry table
__host__ inline uint32_t smem_bytes_for(int sc, uint32_t rb, uint32_t h0_slots, uint32_t nslot = 3, bool split = false) {
    if (split) return SMEM_HDR + SPLIT_BYTES + nslot * rb;
    uint32_t b = SMEM_HDR + nslot * rb;
    if (sc & SC_ENTROPY) b += HIST_BYTES;
    if (sc & SC_BIGRAM)
        b += h0_slots ? ((sc & SC_HALF) ? 0u : h0_slots * 4) + ((sc & SC_WIDE) ? Dense<6>::SMALL : Dense<5>::SMALL) : 65536 * 2;
    return b;
}

// ---- scoring passes ------------------------------------------------------------------------------------------------

// MinSum for all five filters in one sweep. Accumulates S_f = sum over bytes of ||cur-pred_f| - 128|; the caller turns
// it into the reference's score 128*bytes - S_f + f (the filter-type byte f counts |f| = f, png/mod.rs:409-414).
// S_0 also answers the all-zero-row test: ||x| - 128| is 128 only for x = 0, so S_0 = 128 * bytes iff the row is zero.
// S0: bpp = 4 (the funnel shift that forms "left" is by zero bits: a plain register pick).
// floor((a + b) / 2) per byte = (a & b) + (((a ^ b) & 0xFE..) >> 1); the shift as a multiply-high by 2^31 -- handed in as a
// kernel parameter so that it stays an IMAD.HI on the FMA pipe -- and the mask folded into the XOR's LOP3: two ALU
// operations instead of the four of __vhaddu4 on the pipe that bounds the MinSum and Entropy scorers
__device__ __forceinline__ uint32_t havg4(uint32_t a, uint32_t b, uint32_t two31) {
#ifdef OXI_HAVG_OLD
    return __vhaddu4(a, b);
#else
    return (a & b) + __umulhi((a ^ b) & 0xFEFEFEFEu, two31);
#endif
}
__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t acc) { // acc + sum_i |a_i - b_i| (bytes)
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}
template <int D, bool TAIL, bool S0>
__device__ __forceinline__ void minsum_chunk(const uint8_t *cur, const uint8_t *up, uint32_t c, int valid, uint32_t shift,
                                             uint32_t (&a)[5], uint32_t two31) {
    const Chunk cw = load_chunk(cur, c), uw = load_chunk(up, c);
    const uint32_t H = 0x80808080u;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t x = cw.w[k + 2], u = uw.w[k + 2];
        uint32_t l = (S0 && D == 1) ? cw.w[k + 1] : left_word<D>(cw, k, shift);
        uint32_t ul = (S0 && D == 1) ? uw.w[k + 1] : left_word<D>(uw, k, shift);
        if (TAIL) { // last, partial chunk: bytes past the row contribute |0-0| -> cancel exactly
            const uint32_t m = tail_mask(valid - 4 * k);
            x &= m; u &= m; l &= m; ul &= m;
        }
#ifdef OXI_MINSUM_OLD
        a[0] = __vsadu4(x, H) + a[0];
        a[1] = __vsadu4(__vabsdiffu4(x, l), H) + a[1];
        a[2] = __vsadu4(__vabsdiffu4(x, u), H) + a[2];
        a[3] = __vsadu4(__vabsdiffu4(x, __vhaddu4(l, u)), H) + a[3];
        a[4] = __vsadu4(__vabsdiffu4(x, paeth4(l, u, ul)), H) + a[4];
#else
        // sum of absolute differences straight into the accumulator (VABSDIFF4.ACC with the accumulator as addend): the
        // compiler's form sums into zero and needs an IADD3 per pair of words on the ALU pipe that bounds this scorer
        a[0] = sad4_acc(x, H, a[0]);
        a[1] = sad4_acc(__vabsdiffu4(x, l), H, a[1]);
        a[2] = sad4_acc(__vabsdiffu4(x, u), H, a[2]);
        a[3] = sad4_acc(__vabsdiffu4(x, havg4(l, u, two31)), H, a[3]);
        a[4] = sad4_acc(__vabsdiffu4(x, paeth4(l, u, ul)), H, a[4]);
#endif
    }
}
template <int D, int NT>
__device__ __forceinline__ void minsum_pass(const uint8_t *cur, const uint8_t *up, uint32_t len, uint32_t shift,
                                            uint32_t *s_min, uint32_t two31) {
    const uint32_t nfull = len >> 4; // complete 16-byte chunks: no masking in the main loop
    uint32_t a[5] = {0, 0, 0, 0, 0};
    if (D == 1 && shift == 0) {
        for (uint32_t c = threadIdx.x; c < nfull; c += NT) minsum_chunk<D, false, true>(cur, up, c, 16, shift, a, two31);
    } else {
        for (uint32_t c = threadIdx.x; c < nfull; c += NT) minsum_chunk<D, false, false>(cur, up, c, 16, shift, a, two31);
    }
    if ((len & 15u) && threadIdx.x == (nfull % NT)) minsum_chunk<D, true, false>(cur, up, nfull, (int)(len & 15u), shift, a, two31);
    a[0] = warp_sum(a[0]); a[1] = warp_sum(a[1]); a[2] = warp_sum(a[2]); a[3] = warp_sum(a[3]); a[4] = warp_sum(a[4]);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_min[0], a[0]); atomicAdd(&s_min[1], a[1]); atomicAdd(&s_min[2], a[2]);
        atomicAdd(&s_min[3], a[3]); atomicAdd(&s_min[4], a[4]);
    }
}

// Entropy: histogram the residual bytes of all five filters (filter-type byte added at scoring time).  The counter of
// byte value v of filter f sits at byte offset f*1024 + v*4: per byte one prmt (v, zero-extended) and one multiply-add
// (v * 4 + the histogram's shared-window address + f*1024; see scale4), then the atomic.
// One copy per histogram: round 1 kept four lane-selected, bank-rotated replicas (2.2 wavefronts per ATOMS measured);
// replaying the lane -> byte map on the c2 image gives 2.14 for that layout and 1.2 for a single copy, because the lanes
// of one instruction that hold the same residual value then hit the same counter and are merged by the hardware.
constexpr uint32_t HIST_ADDR = SMEM_WIN + SMEM_HDR; // the histograms are the first thing behind the header in every layout
// v * four with four == 4 handed in as a kernel parameter: a literal 4 becomes a shift or LEA (ALU pipe), an opaque
// multiplier stays an IMAD (FMA pipe)
__device__ __forceinline__ uint32_t scale4(uint32_t v, uint32_t four) {
    uint32_t d;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(d) : "r"(v), "r"(four));
    return d;
}
// The counters of one row word for all five filters.  Per byte: the value x is extracted once (PRMT) and scaled
// (IMAD: x*4), per filter the predictor byte p is extracted (PRMT) and the counter address formed by one multiply-add
// ((256 + x)*4 - p*4) on the FMA pipe; the histogram base is the atomic's immediate.  No bytewise
// subtraction: `__vsub4` costs three ALU operations per word and filter on the pipe that bounds this scorer (82 %).
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) { // a * b + c, b opaque: stays an IMAD
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
template <int K>
__device__ __forceinline__ void hist_byte(uint32_t x, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t p4, uint32_t four, uint32_t neg4) {
    const uint32_t xq = mad_lo(prmt(x, 0u, 0x4440u + K), four, 1024u); // (256 + x) * 4: the register part stays positive
    smem_inc_at<HIST_ADDR + 0 * 2048u>(xq);
    smem_inc_at<HIST_ADDR + 1 * 2048u>(mad_lo(prmt(p1, 0u, 0x4440u + K), neg4, xq));
    smem_inc_at<HIST_ADDR + 2 * 2048u>(mad_lo(prmt(p2, 0u, 0x4440u + K), neg4, xq));
    smem_inc_at<HIST_ADDR + 3 * 2048u>(mad_lo(prmt(p3, 0u, 0x4440u + K), neg4, xq));
    smem_inc_at<HIST_ADDR + 4 * 2048u>(mad_lo(prmt(p4, 0u, 0x4440u + K), neg4, xq));
}
template <int D, bool TAIL>
__device__ __forceinline__ void entropy_chunk(const uint8_t *cur, const uint8_t *up, uint32_t c, int valid, uint32_t shift,
                                              uint32_t four) {
    const Chunk cw = load_chunk(cur, c), uw = load_chunk(up, c);
    const uint32_t neg4 = 0u - four;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t x = cw.w[k + 2], u = uw.w[k + 2], l = left_word<D>(cw, k, shift), ul = left_word<D>(uw, k, shift);
        const uint32_t pa = havg4(l, u, neg4 << 29), pp = paeth4(l, u, ul); // (-4) << 29 = 2^31, opaque like `four`
        if (!TAIL) {
            hist_byte<0>(x, l, u, pa, pp, four, neg4);
            hist_byte<1>(x, l, u, pa, pp, four, neg4);
            hist_byte<2>(x, l, u, pa, pp, four, neg4);
            hist_byte<3>(x, l, u, pa, pp, four, neg4);
        } else {
            const int nv = min(valid - 4 * k, 4);
            if (nv <= 0) break;
            for (int j = 0; j < nv; j++) {
                const uint32_t xb = (x >> (8 * j)) & 255u;
                smem_inc(HIST_ADDR + 0 * 2048u + 1024u + (xb << 2));
                smem_inc(HIST_ADDR + 1 * 2048u + 1024u + ((xb - ((l >> (8 * j)) & 255u)) << 2));
                smem_inc(HIST_ADDR + 2 * 2048u + 1024u + ((xb - ((u >> (8 * j)) & 255u)) << 2));
                smem_inc(HIST_ADDR + 3 * 2048u + 1024u + ((xb - ((pa >> (8 * j)) & 255u)) << 2));
                smem_inc(HIST_ADDR + 4 * 2048u + 1024u + ((xb - ((pp >> (8 * j)) & 255u)) << 2));
            }
        }
    }
}
template <int D, int NT>
__device__ __forceinline__ void entropy_pass(const uint8_t *cur, const uint8_t *up, uint32_t len, uint32_t shift,
                                             uint32_t *hist, uint32_t four) {
    static_assert(HIST_REP == 1, "the histogram addressing is written for one copy");
    const uint32_t nfull = len >> 4;
    for (uint32_t c = threadIdx.x; c < nfull; c += NT) entropy_chunk<D, false>(cur, up, c, 16, shift, four);
    if ((len & 15u) && threadIdx.x == (nfull % NT)) entropy_chunk<D, true>(cur, up, nfull, (int)(len & 15u), shift, four);
}
// sum over present values of ilog2i(count) (strategies.rs:136-142); clears the histograms for the next row.
template <int NT>
__device__ __forceinline__ void entropy_score(uint32_t *hist, uint32_t *s_ent) {
    for (uint32_t idx = threadIdx.x; idx < 5 * 256; idx += NT) {
        const uint32_t f = idx >> 8, v = idx & 255u;
        uint32_t *h = hist + f * HIST_BINS + v; // residual v: differences v - 256 (counter v) and v (counter 256 + v)
        const uint32_t cnt = h[0] + h[256] + ((v == f) ? 1u : 0u); // + the filter-type byte that leads the filtered line
        h[0] = 0;
        h[256] = 0;
        uint32_t s = warp_sum(ilog2i(cnt | (cnt == 0)) * (cnt != 0));
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_ent[f], s);
    }
}

// One Bigrams/BigEnt trial for filter F over the stream t[0]=F, t[1+i]=residual[i]: windows (t[p-1], t[p]), p=1..len.
// Work is split at word granularity so that every thread of the CTA has windows even for short rows: thread t owns row
// words t, t+NT, ...; the residual words stay in registers for both passes.
// Pass 1 counts into the u16 table (the adder that sees 0 owns the bigram); pass 2 lets owners read the final count,
// contribute 1 to the distinct count and ilog2i(count) to the entropy, and clear the entry for the next trial.
template <int D, int NT, int F, int WPT>
__device__ __forceinline__ void bigram_trial(const uint8_t *cur, const uint8_t *up, uint32_t len, uint32_t shift, int bpp,
                                             uint32_t *table, uint32_t *s_bgc, uint32_t *s_bge) {
    const uint32_t nwords = (len + 3) >> 2;
    const uint32_t *cw = reinterpret_cast<const uint32_t *>(cur), *uw = reinterpret_cast<const uint32_t *>(up);
    uint16_t *t16 = reinterpret_cast<uint16_t *>(table);
    uint32_t r[WPT], pv[WPT];
#pragma unroll
    for (int i = 0; i < WPT; i++) {
        const uint32_t k = threadIdx.x + i * NT;
        uint32_t x = 0;
        if (k < nwords) {
            // row words k-1-D .. k are addressable: the slot has 32 zero bytes in front of the row
            const int kk = (int)k;
            uint32_t l = 0, u = 0, ul = 0;
            x = cw[kk];
            if (F == 1 || F == 3 || F == 4)
                l = D == 2 ? cw[kk - 2] : D == 1 ? __funnelshift_l(cw[kk - 2], cw[kk - 1], shift)
                                                 : __funnelshift_l(cw[kk - 1], cw[kk], shift);
            if (F >= 2) u = uw[kk];
            if (F == 4)
                ul = D == 2 ? uw[kk - 2] : D == 1 ? __funnelshift_l(uw[kk - 2], uw[kk - 1], shift)
                                                  : __funnelshift_l(uw[kk - 1], uw[kk], shift);
            x = __vsub4(x, pred4<F>(l, u, ul));
        }
        r[i] = x;
        // stream byte just before this word: the last residual byte of the previous word (neighbouring lane), the
        // filter id in front of word 0
        uint32_t p = __shfl_up_sync(0xffffffffu, x >> 24, 1);
        if ((threadIdx.x & 31) == 0) p = k == 0 ? (uint32_t)F : (k < nwords ? residual_at(F, cur, up, 4 * (int)k - 1, bpp) : 0u);
        pv[i] = p;
    }
    unsigned long long own = 0;
#pragma unroll
    for (int i = 0; i < WPT; i++) {
        const uint32_t k = threadIdx.x + i * NT;
        const int valid = (int)len - 4 * (int)k;
        uint32_t prev = pv[i];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t b = (r[i] >> (8 * j)) & 255u;
            const uint32_t bg = (prev << 8) | b;
            prev = b;
            if (j < valid) {
                const uint32_t old = atomicAdd(&table[big_word(bg)], (bg & 1u) ? 0x10000u : 1u);
                const uint32_t cnt = (bg & 1u) ? (old >> 16) : (old & 0xFFFFu);
                if (cnt == 0) own |= 1ull << (4 * i + j);
            }
        }
    }
    __syncthreads();
    uint32_t distinct = 0, ent = 0;
    if (own) {
#pragma unroll
        for (int i = 0; i < WPT; i++) {
            uint32_t prev = pv[i];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t b = (r[i] >> (8 * j)) & 255u;
                const uint32_t bg = (prev << 8) | b;
                prev = b;
                if ((own >> (4 * i + j)) & 1ull) {
                    const uint32_t cnt = t16[(big_word(bg) << 1) | (bg & 1u)];
                    t16[(big_word(bg) << 1) | (bg & 1u)] = 0;
                    distinct += 1;
                    ent += ilog2i(cnt);
                }
            }
        }
    }
    distinct = warp_sum(distinct);
    ent = warp_sum(ent);
    if ((threadIdx.x & 31) == 0 && distinct) {
        atomicAdd(&s_bgc[F], distinct);
        atomicAdd(&s_bge[F], ent);
    }
    __syncthreads();
}

template <int D, int NT, int WPT>
__device__ __forceinline__ void bigram_row(const uint8_t *cur, const uint8_t *up, uint32_t len, uint32_t shift, int bpp,
                                           uint32_t *table, uint32_t *s_bgc, uint32_t *s_bge) {
    bigram_trial<D, NT, 0, WPT>(cur, up, len, shift, bpp, table, s_bgc, s_bge);
    bigram_trial<D, NT, 1, WPT>(cur, up, len, shift, bpp, table, s_bgc, s_bge);
    bigram_trial<D, NT, 2, WPT>(cur, up, len, shift, bpp, table, s_bgc, s_bge);
    bigram_trial<D, NT, 3, WPT>(cur, up, len, shift, bpp, table, s_bgc, s_bge);
    bigram_trial<D, NT, 4, WPT>(cur, up, len, shift, bpp, table, s_bgc, s_bge);
}

// ---- Bigrams/BigEnt, all five trials at once (rows <= 16384 bytes) --------------------------------------------------
// Filtered rows of natural images are sharply peaked around 0, which makes atomics on one shared table serialise
// (ncu: 10-32 wavefronts per ATOMS; same-address lanes serialise with or without a return value), and a directly
// indexed table of all 65536 bigrams (128 KB) limits the SM to one CTA.  Instead, per row:
//  * trials 1..4: a window whose two stream bytes both lie in [-16, 15] (as i8) is counted in a dense, bank-swizzled
//    32x32 u16 table private to its trial and to the lane's replica (lane & 3); the remaining (outlier) windows go to a
//    small open-addressing hash table of (key, count) per trial;
//  * trial 0 (None: the raw bytes, rarely near zero) goes to one hash table `h0` with more than twice as many slots as
//    the row has windows, so an insert always finds a slot;
//  * after one barrier every table is swept (and cleared): non-zero counters are the distinct bigrams, their counts
//    give the entropy.  No ownership bookkeeping, nothing kept in registers across the barrier;
//  * a trial whose outliers crowd its small hash in some row (noise-like data) is flagged; its outliers are then
//    re-derived from the staged row and counted in h0, trial by trial, after h0 has been swept.
// Exact for any data: every window is counted in exactly one structure, and their key sets are disjoint.

// h0 entries are ((key + 1) << 15) | count: key + 1 is never 0 (and at most 2^16), a count is at most 16384 (row bytes).
__device__ __forceinline__ void h0_collide(uint32_t *h, uint32_t s, uint32_t tag, uint32_t old, uint32_t mask) {
    for (;;) {
        if ((old & ~0x7FFFu) == tag) {
            atomicAdd(&h[s], 1u);
            return;
        }
        s = (s + 1) & mask;
        old = atomicCAS(&h[s], 0u, tag | 1u);
        if (old == 0) return;
    }
}
__device__ __forceinline__ uint32_t h0_slot(uint32_t key, uint32_t hshift) { return (key * 0x9E3779B1u) >> hshift; }
__device__ __forceinline__ void h0_add(uint32_t *h, uint32_t key, uint32_t hshift, uint32_t mask) {
    const uint32_t tag = (key + 1u) << 15, s = h0_slot(key, hshift);
    const uint32_t old = atomicCAS(&h[s], 0u, tag | 1u);
    if (old != 0) h0_collide(h, s, tag, old, mask);
}
// sweep (and clear) h0: distinct keys and sum of ilog2i(count) over this thread's slots
template <int NT>
__device__ __forceinline__ void h0_sweep(uint32_t *h, uint32_t slots, uint32_t &d, uint32_t &e) {
    uint4 *h4 = reinterpret_cast<uint4 *>(h);
    for (uint32_t q = threadIdx.x; q < slots / 4; q += NT) {
        const uint4 v = h4[q];
        if (v.x | v.y | v.z | v.w) {
            h4[q] = make_uint4(0, 0, 0, 0);
            const uint32_t c0 = v.x & 0x7FFFu, c1 = v.y & 0x7FFFu, c2 = v.z & 0x7FFFu, c3 = v.w & 0x7FFFu;
            d += (c0 != 0) + (c1 != 0) + (c2 != 0) + (c3 != 0);
            if (c0 > 1) e += ilog2i(c0);
            if (c1 > 1) e += ilog2i(c1);
            if (c2 > 1) e += ilog2i(c2);
            if (c3 > 1) e += ilog2i(c3);
        }
    }
}

// Outlier keys of trials 1..4 are never 0 (one of the two bytes lies outside [-16, 15]), so an empty slot is 0; a count
// is at most the number of windows of a row (<= 16384), so it cannot reach the key bits.  Returns false when no slot was
// found within HPROBE probes (table crowded): the caller then flags the trial.
__device__ __forceinline__ bool ohash_add(uint32_t *h, uint32_t key) {
    uint32_t s = ((key * 40503u) >> 7) & (HSLOTS - 1);
    for (int probe = 0; probe < HPROBE; probe++) {
        uint32_t cur = *(volatile uint32_t *)&h[s];
        if (cur == 0) {
            cur = atomicCAS(&h[s], 0u, (key << 16) | 1u);
            if (cur == 0) return true;
        }
        if ((cur >> 16) == key) {
            atomicAdd(&h[s], 1u);
            return true;
        }
        s = (s + 1) & (HSLOTS - 1);
    }
    return false;
}
// the next row of the band, streamed into its ring slot (by TMA, or cooperatively for rows TMA cannot take)
struct Prefetch {
    bool on, tma;
    uint8_t *dst;
    const uint8_t *src;
    uint32_t len;
    uint64_t *bar;
};
template <int NT>
__device__ __noinline__ void coop_load_row(uint8_t *dst, const uint8_t *src, uint32_t len);
template <int NT>
__device__ __forceinline__ void issue_prefetch(const Prefetch &pf) {
    if (!pf.on) return;
    if (pf.tma) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(pf.bar, pf.len);
            tma_load_1d(pf.dst, pf.src, pf.len, pf.bar);
        }
    } else {
        coop_load_row<NT>(pf.dst, pf.src, pf.len);
    }
}

// ilog2i that is also valid for i == 0 (returns 0): no branch in the table sweeps.  PTX shl clamps the shift amount,
// so for i == 0 (lg = -1) the subtrahend is 0 and the product is 0 * 1.
__device__ __forceinline__ uint32_t ilog2i_z(uint32_t i) {
    uint32_t lg, sh; // bfind: position of the most significant set bit, 0xFFFFFFFF for 0 (SASS FLO)
    asm("bfind.u32 %0, %1;" : "=r"(lg) : "r"(i));
    asm("shl.b32 %0, %1, %2;" : "=r"(sh) : "r"(2u), "r"(lg));
    return i * (lg + 2u) - sh;
}
// Dense tables, trials 1..4: u32 counters, 16 KB in all.  Byte offset of the counter of window (first, second) -- both as
// range codes 0..31 -- of trial f:
//     first << 9 | (f-1) << 7 | ((second + 5 * first) & 31) << 2
// i.e. the high byte of the offset is (first << 1 | (f-1) >> 1) and the low byte ((f-1) & 1) << 7 | bank << 2: one prmt
// per window picks the two bytes out of two pre-scaled words, one red.shared per window counts it (the hardware
// merges the lanes of one instruction that hit the same counter: SASS ATOMS.POPC.INC).  The bank is second + 5 * first:
// the keys of a filtered photographic row form a cluster of a few residuals around (16, 16), and a 5 x 5 (6 x 5) block
// of neighbouring keys lands on distinct banks, where first ^ second folds all "equal residual" keys onto one bank.
// Lane-selected replicas and a per-lane rotation of the windows were measured in round 1 at 3.1 wavefronts per ATOMS;
// replaying the kernel's lane -> window map on the bench images (profiles/r02_bank_sim.py) gives 3.06 for that layout
// and 1.85 for this one, because the lanes of one instruction then count the same channel pair and mostly coincide.
// RB = 6 (64 x 64 counters per trial): offset first << 10 | (f-1) << 8 | ((second + 3 * first) & 63) << 2; its high byte
// (first << 2 | f-1) uses all eight bits, so it is stored with bit 7 inverted and the instruction's immediate carries
// + 0x8000: prmt's sign replication then yields offset - 0x8000.
template <int RB>
__device__ __forceinline__ uint32_t dense_off(uint32_t f, uint32_t first, uint32_t second) {
    if (RB == 5) return (first << 9) | ((f - 1u) << 7) | (((second + 5u * first) & 31u) << 2);
    return (first << 10) | ((f - 1u) << 8) | (((second + 3u * first) & 63u) << 2);
}

// One complete row word (four windows) of trial F.  t = range codes of the word's residual bytes, tpw = codes of the
// stream bytes in front of them; all eight codes in range.  DB = shared-window address of the dense tables.
template <int F, uint32_t DB, int RB>
__device__ __forceinline__ void dense_word(uint32_t t, uint32_t tpw) {
    if (RB == 5) {
        const uint32_t hi = tpw * 2u + (uint32_t)((F - 1) >> 1) * 0x01010101u;
        const uint32_t s = tpw * 5u + t; // bytes <= 31 + 155: no carry between bytes; << 2 moves only bits the mask drops
        const uint32_t lo = ((s << 2) & 0x7C7C7C7Cu) | ((uint32_t)((F - 1) & 1) * 0x80808080u);
        smem_inc_at<DB>(prmt(lo, hi, 0xCC40u));
        smem_inc_at<DB>(prmt(lo, hi, 0xDD51u));
        smem_inc_at<DB>(prmt(lo, hi, 0xEE62u));
        smem_inc_at<DB>(prmt(lo, hi, 0xFF73u));
    } else {
        const uint32_t hi = (tpw * 4u + (uint32_t)(F - 1) * 0x01010101u) ^ 0x80808080u; // bytes <= 252 + 3
        const uint32_t s = tpw * 3u + t;                                                 // bytes <= 189 + 63
        const uint32_t lo = (s << 2) & 0xFCFCFCFCu;
        smem_inc_at<DB + 0x8000u>(prmt(lo, hi, 0xCC40u));
        smem_inc_at<DB + 0x8000u>(prmt(lo, hi, 0xDD51u));
        smem_inc_at<DB + 0x8000u>(prmt(lo, hi, 0xEE62u));
        smem_inc_at<DB + 0x8000u>(prmt(lo, hi, 0xFF73u));
    }
}
// the same word window by window: tail words and words with an out-of-range byte
// (one copy for all trials, F at run time: this path is off the straight line and should stay small in the i-cache;
// it takes shared-window addresses, not pointers, so that the callers form no generic pointers on the straight line)
template <uint32_t DB, int RB>
__device__ __noinline__ void dense_word_slow(uint32_t F, uint32_t t, uint32_t tprev, int valid, uint32_t s_out_addr) {
    uint32_t *ohash = reinterpret_cast<uint32_t *>(__cvta_shared_to_generic(DB + Dense<RB>::REGION));
    volatile uint32_t *s_out = reinterpret_cast<volatile uint32_t *>(__cvta_shared_to_generic(s_out_addr));
    uint32_t tp = tprev >> 24;
    for (int j = 0; j < 4; j++) {
        const uint32_t tj = (t >> (8 * j)) & 255u;
        if (j < valid) {
            if (((tp | tj) & Dense<RB>::OUT) == 0) {
                smem_inc(DB + dense_off<RB>(F, tp, tj));
            } else if (s_out[F] == 0) {
                // outliers go to the trial's hash table until an insert finds it crowded; from then on the
                // trial is flagged and all its outliers are counted in h0 after the barrier
                // (the key is the pair of residual bytes; never 0, since one of them lies outside the range)
                if (!ohash_add(ohash + (F - 1) * HSLOTS, (((tp - Dense<RB>::HALF) & 255u) << 8) | ((tj - Dense<RB>::HALF) & 255u)))
                    s_out[F] = 1u;
            }
        }
        tp = tj;
    }
}

// trial 0 on coarse keys (low five bits of both raw bytes): 32x32 counters in the even 128-byte lines of the 8 KB at
// COARSE_OFF; offset first << 8 | ((second ^ first) & 31) << 2, i.e. high byte = first, low byte = bank << 2
template <uint32_t DB, int RB>
__device__ __forceinline__ void dense_word0(uint32_t x, uint32_t xpw) {
    const uint32_t hi = xpw & 0x1F1F1F1Fu;
    const uint32_t lo = ((x ^ xpw) & 0x1F1F1F1Fu) << 2;
    smem_inc_at<DB + Dense<RB>::COARSE>(prmt(lo, hi, 0xCC40u));
    smem_inc_at<DB + Dense<RB>::COARSE>(prmt(lo, hi, 0xDD51u));
    smem_inc_at<DB + Dense<RB>::COARSE>(prmt(lo, hi, 0xEE62u));
    smem_inc_at<DB + Dense<RB>::COARSE>(prmt(lo, hi, 0xFF73u));
}
template <uint32_t DB, int RB>
__device__ __noinline__ void dense_word0_slow(uint32_t x, uint32_t xprev, int valid) {
    uint32_t p = (xprev >> 24) & 31u;
    for (int j = 0; j < valid && j < 4; j++) {
        const uint32_t b = (x >> (8 * j)) & 31u;
        smem_inc(DB + Dense<RB>::COARSE + ((p << 8) | ((b ^ p) << 2)));
        p = b;
    }
}
// the last, partial pair of a row (fewer than eight valid bytes): everything window by window
template <uint32_t DB, int RB>
__device__ __noinline__ void dense_pair_tail(uint32_t x0, uint32_t x1, uint32_t xprev, uint32_t t10, uint32_t t11, uint32_t t20,
                                             uint32_t t21, uint32_t t30, uint32_t t31, uint32_t t40, uint32_t t41, uint32_t tp1,
                                             uint32_t tp2, uint32_t tp3, uint32_t tp4, int valid0, uint32_t s_out_addr) {
    dense_word0_slow<DB, RB>(x0, xprev, valid0);
    dense_word_slow<DB, RB>(1u, t10, tp1, valid0, s_out_addr);
    dense_word_slow<DB, RB>(2u, t20, tp2, valid0, s_out_addr);
    dense_word_slow<DB, RB>(3u, t30, tp3, valid0, s_out_addr);
    dense_word_slow<DB, RB>(4u, t40, tp4, valid0, s_out_addr);
    if (valid0 > 4) {
        dense_word0_slow<DB, RB>(x1, x0, valid0 - 4);
        dense_word_slow<DB, RB>(1u, t11, t10, valid0 - 4, s_out_addr);
        dense_word_slow<DB, RB>(2u, t21, t20, valid0 - 4, s_out_addr);
        dense_word_slow<DB, RB>(3u, t31, t30, valid0 - 4, s_out_addr);
        dense_word_slow<DB, RB>(4u, t41, t40, valid0 - 4, s_out_addr);
    }
}

// Exact Bigrams/BigEnt scores of trial 0 (the raw row), for the rows whose coarse bounds could not rule it out: every
// window goes to the hash table h0 (clean on entry and on exit); distinct keys and entropy are accumulated in
// s_bgc[5] / s_bge[5] (zero on entry).  Two CTA-wide barriers.
template <int NT>
__device__ __noinline__ void bigram_exact0(const uint8_t *cur, uint32_t len, uint32_t *h0, uint32_t h0_slots, uint32_t *s_bgc,
                                           uint32_t *s_bge) {
    const uint32_t hmask = h0_slots - 1, hshift = (uint32_t)__clz((int)h0_slots) + 1;
    const uint32_t *cw = reinterpret_cast<const uint32_t *>(cur);
    for (uint32_t k = threadIdx.x; 4 * k < len; k += NT) {
        const uint32_t x = cw[k];
        uint32_t p = cw[(int)k - 1] >> 24; // word -1 is the zero pad: the filter-type byte 0 in front of the row
        const int valid = (int)len - 4 * (int)k;
        for (int j = 0; j < 4 && j < valid; j++) {
            const uint32_t b = (x >> (8 * j)) & 255u;
            h0_add(h0, (p << 8) | b, hshift, hmask);
            p = b;
        }
    }
    __syncthreads();
    uint32_t d = 0, e = 0;
    h0_sweep<NT>(h0, h0_slots, d, e);
    d = warp_sum(d);
    e = warp_sum(e);
    if ((threadIdx.x & 31) == 0 && d) {
        atomicAdd(&s_bgc[5], d);
        atomicAdd(&s_bge[5], e);
    }
    __syncthreads();
}

// Per row (rows <= 16384 bytes):
//  * trials 1..4: a window whose two stream bytes both lie in [-16, 15] (as i8) is counted in a dense 32x32 table of
//    its trial (two lane-selected replicas, bank = first ^ second); the remaining (outlier) windows go to a small
//    open-addressing hash table of (key, count) per trial;
//  * trial 0 (None: the raw bytes, spread over the whole byte range) is counted on COARSE keys -- the low five bits of
//    both bytes -- in a fifth dense table.  Merging keys can only lower the number of distinct keys and raise
//    sum ilog2i(count) (ilog2i is convex with ilog2i(0) = 0, hence superadditive), so the coarse scores are a lower
//    bound of the trial's Bigrams score and an upper bound of its BigEnt score.  None is first in RowFilter::ALL and
//    wins ties, so it is chosen iff d0 <= min d_f resp. e0 >= max e_f; when the bounds already say it loses (the
//    usual case for photographic rows) its exact scores are never needed.  Otherwise the caller runs bigram_exact0.
//  * after one barrier every table is swept (and cleared): non-zero counters are the distinct bigrams, their counts
//    give the entropy.  No ownership bookkeeping, nothing kept in registers across the barrier;
//  * a trial whose outliers crowd its small hash in some row (noise-like data) is flagged; its outliers are then
//    re-derived from the staged row and counted in h0, trial by trial.
// Exact for any data: every window is counted in exactly one structure, and their key sets are disjoint.
// Flagged trials (their outlier hash crowded in this row): the outlier windows, re-derived from the staged row, are
// counted in h0, one trial at a time.  Off the straight line: kept out of line for the instruction cache.
template <int D, int NT, int RB>
__device__ __noinline__ void bigram_flagged(const uint8_t *cur, const uint8_t *up, uint32_t len, uint32_t shift, int bpp,
                                            const Smem &sm, uint32_t h0_slots, uint32_t flagged) {
    const uint32_t nwords = (len + 3) >> 2, lane = threadIdx.x & 31u;
    uint32_t *h0 = sm.table;
    const uint32_t hmask = h0_slots - 1, hshift = (uint32_t)__clz((int)h0_slots) + 1;
    __syncthreads(); // h0 is clean
    for (int f = 1; f < 5; f++) {
        if (!((flagged >> (f - 1)) & 1u)) continue;
        // word-granular: thread t rebuilds the residual words t, t+NT, ... byte-SIMD; the byte in front of a word comes
        // from the neighbouring lane (lane 0: one scalar pixel, or the filter id in front of word 0)
        const uint32_t *cw = reinterpret_cast<const uint32_t *>(cur), *uw = reinterpret_cast<const uint32_t *>(up);
        for (uint32_t k = threadIdx.x; k < ((nwords + 31u) & ~31u); k += NT) {
            const int kk = (int)k;
            uint32_t r = 0;
            if (k < nwords) {
                const uint32_t x = cw[kk], u = uw[kk];
                const uint32_t l = D == 2 ? cw[kk - 2] : D == 1 ? __funnelshift_l(cw[kk - 2], cw[kk - 1], shift)
                                                                 : __funnelshift_l(cw[kk - 1], cw[kk], shift);
                const uint32_t ul = D == 2 ? uw[kk - 2] : D == 1 ? __funnelshift_l(uw[kk - 2], uw[kk - 1], shift)
                                                                  : __funnelshift_l(uw[kk - 1], uw[kk], shift);
                r = __vsub4(x, pred4_rt(f, l, u, ul));
            }
            uint32_t pr = __shfl_up_sync(0xffffffffu, r >> 24, 1);
            if (lane == 0) pr = k == 0 ? (uint32_t)f : (k < nwords ? residual_at(f, cur, up, 4 * kk - 1, bpp) : 0u);
            const int valid = (int)len - 4 * kk;
            for (int j = 0; j < 4 && j < valid; j++) {
                const uint32_t b = (r >> (8 * j)) & 255u;
                if ((((pr + Dense<RB>::HALF) | (b + Dense<RB>::HALF)) & Dense<RB>::OUT) != 0) h0_add(h0, (pr << 8) | b, hshift, hmask);
                pr = b;
            }
        }
        __syncthreads();
        uint32_t d = 0, e = 0;
        h0_sweep<NT>(h0, h0_slots, d, e);
        d = warp_sum(d);
        e = warp_sum(e);
        if ((threadIdx.x & 31) == 0 && d) {
            atomicAdd(&sm.s_bgc_()[f], d);
            atomicAdd(&sm.s_bge_()[f], e);
        }
        __syncthreads();
    }
}

template <int D, int NT, uint32_t DB, int RB>
__device__ __forceinline__ uint32_t bigram_row_small(const uint8_t *cur, const uint8_t *up, uint32_t len, uint32_t shift, int bpp,
                                                     const Smem &sm, uint32_t h0_slots, const Prefetch &pf) {
    static_assert(NT % 128 == 0, "the sweep gives every trial NT / 128 whole warps");
    const uint32_t nwords = (len + 3) >> 2, npairs = (nwords + 1) >> 1;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t s_out_addr = smem_u32(sm.s_out_());
    uint32_t nz = 0; // OR of this thread's row bytes (all-zero rows short-circuit to None, png/mod.rs:398-404)
    // thread t owns the row words 2t and 2t+1 (one 8-byte load per row and neighbour)
    for (uint32_t pr = threadIdx.x; pr < ((npairs + 31u) & ~31u); pr += NT) {
        uint32_t x[2] = {0, 0}, u[2] = {0, 0}, l[2] = {0, 0}, ul[2] = {0, 0};
        const int k = 2 * (int)pr;
        if (pr < npairs) {
            // words k-2 .. k+1 are addressable: a slot has 32 zero bytes in front of the row and 16+ behind it
            const uint2 cx = *reinterpret_cast<const uint2 *>(cur + 4 * k), cp = *reinterpret_cast<const uint2 *>(cur + 4 * k - 8);
            const uint2 ux = *reinterpret_cast<const uint2 *>(up + 4 * k), upv = *reinterpret_cast<const uint2 *>(up + 4 * k - 8);
            x[0] = cx.x; x[1] = cx.y; u[0] = ux.x; u[1] = ux.y;
            if (D == 2) { l[0] = cp.x; l[1] = cp.y; ul[0] = upv.x; ul[1] = upv.y; }
            else if (D == 1) {
                l[0] = __funnelshift_l(cp.x, cp.y, shift); l[1] = __funnelshift_l(cp.y, cx.x, shift);
                ul[0] = __funnelshift_l(upv.x, upv.y, shift); ul[1] = __funnelshift_l(upv.y, ux.x, shift);
            } else {
                l[0] = __funnelshift_l(cp.y, cx.x, shift); l[1] = __funnelshift_l(cx.x, cx.y, shift);
                ul[0] = __funnelshift_l(upv.y, ux.x, shift); ul[1] = __funnelshift_l(ux.x, ux.y, shift);
            }
        }
        // range codes (residual + 16 per byte) of the four filtered trials
        uint32_t t[4][2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const uint32_t xc = __vadd4(x[i], Dense<RB>::HALF * 0x01010101u);
            t[0][i] = __vsub4(xc, l[i]);
            t[1][i] = __vsub4(xc, u[i]);
            t[2][i] = __vsub4(xc, __vhaddu4(l[i], u[i]));
            t[3][i] = __vsub4(xc, paeth4(l[i], u[i], ul[i]));
        }
        // stream byte in front of the pair: the last byte of the previous pair (neighbouring lane); lane 0 has no
        // neighbour: the filter id in front of word 0, else one scalar pixel
        uint32_t xprev = __shfl_up_sync(0xffffffffu, x[1], 1), tprev[4];
#pragma unroll
        for (int f = 0; f < 4; f++) tprev[f] = __shfl_up_sync(0xffffffffu, t[f][1], 1);
        if (lane == 0) {
            if (k == 0) {
                xprev = 0;
#pragma unroll
                for (int f = 0; f < 4; f++) tprev[f] = ((uint32_t)(f + 1) + Dense<RB>::HALF) << 24;
            } else if (pr < npairs) {
                const int q = 4 * k - 1;
                const uint32_t cx = cur[q], cl = cur[q - bpp], cu = up[q], cul = up[q - bpp];
                xprev = cx << 24;
                tprev[0] = (cx - cl + Dense<RB>::HALF) << 24;
                tprev[1] = (cx - cu + Dense<RB>::HALF) << 24;
                tprev[2] = (cx - ((cl + cu) >> 1) + Dense<RB>::HALF) << 24;
                tprev[3] = (cx - paeth(cl, cu, cul) + Dense<RB>::HALF) << 24;
            }
        }
        if (pr < npairs) {
            const int valid0 = (int)len - 4 * k;
            if (valid0 >= 8) { // both words complete (every pair of a row whose length is a multiple of 8)
                nz |= x[0] | x[1];
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    // "bytes in front" word: byte m is m ? word.byte[m-1] : previous word.byte[3]
                    dense_word0<DB, RB>(x[i], prmt(x[i], i ? x[0] : xprev, 0x2107u));
#pragma unroll
                    for (int f = 0; f < 4; f++) {
                        const uint32_t tp = i ? t[f][0] : tprev[f];
                        const uint32_t tpw = prmt(t[f][i], tp, 0x2107u);
                        if (((t[f][i] | tpw) & Dense<RB>::OUT4) == 0) { // the common case: all four windows in range
                            if (f == 0) dense_word<1, DB, RB>(t[f][i], tpw);
                            if (f == 1) dense_word<2, DB, RB>(t[f][i], tpw);
                            if (f == 2) dense_word<3, DB, RB>(t[f][i], tpw);
                            if (f == 3) dense_word<4, DB, RB>(t[f][i], tpw);
                        } else {
                            dense_word_slow<DB, RB>((uint32_t)f + 1u, t[f][i], tp, 4, s_out_addr);
                        }
                    }
                }
            } else {
                nz |= (x[0] & tail_mask(valid0)) | (x[1] & tail_mask(valid0 - 4));
                dense_pair_tail<DB, RB>(x[0], x[1], xprev, t[0][0], t[0][1], t[1][0], t[1][1], t[2][0], t[2][1], t[3][0], t[3][1],
                                    tprev[0], tprev[1], tprev[2], tprev[3], valid0, s_out_addr);
            }
        }
    }
    __syncthreads();
    // every warp is past the previous row's write pass: with a 3-slot ring the oldest slot may be refilled from here on
    issue_prefetch<NT>(pf);
    // sweep (and clear) the dense tables, four keys at a time, and the outlier hashes.  Trial f belongs to NT / 128 whole
    // warps, so its scores are summed in registers over the loop and added to the row's score set once per warp.
    {
        constexpr uint32_t WPT = NT / 128; // warps per trial
        uint4 *dn = reinterpret_cast<uint4 *>(sm.small);
        const uint32_t warp = threadIdx.x >> 5, f = 1u + warp / WPT;
        uint32_t d = 0, e = 0;
        // c = first : RB | group of four seconds : RB - 2
        constexpr uint32_t G = RB - 2;
        for (uint32_t c = (warp % WPT) * 32u + lane; c < (1u << (2 * RB - 2)); c += WPT * 32u) {
            uint4 *cell = dn + (c >> G) * (4u << G) + (f - 1u) * (1u << G) + (c & ((1u << G) - 1u));
            const uint4 a = *cell;
            *cell = make_uint4(0, 0, 0, 0);
            d += (a.x != 0) + (a.y != 0) + (a.z != 0) + (a.w != 0);
            e += ilog2i_z(a.x) + ilog2i_z(a.y) + ilog2i_z(a.z) + ilog2i_z(a.w);
            // the trial's outlier hash table (HSLOTS / 4 cells): counted unless the trial is flagged (then only cleared)
            if (c < HSLOTS / 4) {
                uint4 *oc = reinterpret_cast<uint4 *>(sm.ohash + (f - 1) * HSLOTS) + c;
                const uint4 v = *oc;
                if (v.x | v.y | v.z | v.w) {
                    *oc = make_uint4(0, 0, 0, 0);
                    if (sm.s_out_()[f] == 0) {
                        if (v.x) { d += 1; e += ilog2i(v.x & 0xFFFFu); }
                        if (v.y) { d += 1; e += ilog2i(v.y & 0xFFFFu); }
                        if (v.z) { d += 1; e += ilog2i(v.z & 0xFFFFu); }
                        if (v.w) { d += 1; e += ilog2i(v.w & 0xFFFFu); }
                    }
                }
            }
        }
        d = warp_sum(d);
        e = warp_sum(e);
        if (lane == 0 && d) {
            atomicAdd(&sm.s_bgc_()[f], d);
            atomicAdd(&sm.s_bge_()[f], e);
        }
        // trial 0 (coarse keys): 1024 counters in 32 lines of 128 bytes (every other line of the 8 KB at COARSE_OFF)
        {
            uint4 *d4 = dn + Dense<RB>::COARSE / 16u;
            uint32_t d0 = 0, e0 = 0;
            for (uint32_t q = threadIdx.x; q < 256u; q += NT) {
                uint4 *cell = d4 + (q >> 3) * 16u + (q & 7u);
                const uint4 a = *cell;
                *cell = make_uint4(0, 0, 0, 0);
                d0 += (a.x != 0) + (a.y != 0) + (a.z != 0) + (a.w != 0);
                e0 += ilog2i_z(a.x) + ilog2i_z(a.y) + ilog2i_z(a.z) + ilog2i_z(a.w);
            }
            if (threadIdx.x < 256u) { // whole warps
                d0 = warp_sum(d0);
                e0 = warp_sum(e0);
                if (lane == 0 && d0) {
                    atomicAdd(&sm.s_bgc_()[0], d0);
                    atomicAdd(&sm.s_bge_()[0], e0);
                }
            }
        }
    }
    // flagged trials: their outlier windows, re-derived from the staged row, are counted in h0, one trial at a time
    const uint32_t flagged = sm.s_out_()[1] | sm.s_out_()[2] << 1 | sm.s_out_()[3] << 2 | sm.s_out_()[4] << 3; // uniform
    if (flagged == 0) return nz;
    bigram_flagged<D, NT, RB>(cur, up, len, shift, bpp, sm, h0_slots, flagged);
    return nz;
}

// ---- Bigrams/BigEnt for rows of 16 KB .. 32 KB (split tables) ---------------------------------------------------------
// Rows this long are mostly 16-bit images, whose low bytes are noise: almost no window has BOTH residuals near zero, so
// the 32 x 32 tables above would count nothing, and the 65536-entry table of `bigram_trial` (u16 counters, returning
// atomics, two passes per trial) ran at 2.7 % of the roofline.  Here a window (a, b) of trial F -- as range codes
// residual + 16 -- is counted
//     in table A[a][b ^ a]        (32 x 256 u32)  when a < 32 (first residual in [-16, 15]), else
//     in table B[a][(b ^ a) & 31] (256 x 32 u32)  when b < 32,
//     else in a small hash (both far from zero: the first pixels of a row, edges).
// Every key has exactly one home, so sweeping A, B and the hash gives the exact distinct count and sum of ilog2i(count).
// Trial 0 (raw bytes) uses a = byte & 31 for every window: keys merge, which bounds its scores (see bigram_row_small);
// if the bounds leave open that None wins, or a trial's hash overflows, that trial is recounted exactly by
// `bigram_trial` in the same memory (the table sets are clean between trials).  The trials run one after the other
// on two alternating table sets: while the CTA counts trial F+1 into one set it sweeps trial F out of the other, with
// one CTA-wide barrier per trial.
__device__ __noinline__ void split_outlier(uint32_t key, uint32_t F, uint32_t ohash_addr, uint32_t s_out_addr) {
    volatile uint32_t *s_out = reinterpret_cast<volatile uint32_t *>(__cvta_shared_to_generic(s_out_addr));
    if (s_out[F]) return;
    uint32_t *h = reinterpret_cast<uint32_t *>(__cvta_shared_to_generic(ohash_addr));
    uint32_t s = ((key * 40503u) >> 6) & (SPLIT_OSLOTS - 1);
    for (int probe = 0; probe < HPROBE; probe++) {
        uint32_t cur = *(volatile uint32_t *)&h[s];
        if (cur == 0) {
            cur = atomicCAS(&h[s], 0u, (key << 16) | 1u);
            if (cur == 0) return;
        }
        if ((cur >> 16) == key) {
            atomicAdd(&h[s], 1u);
            return;
        }
        s = (s + 1) & (SPLIT_OSLOTS - 1);
    }
    s_out[F] = 1u; // crowded: the trial is recounted exactly
}

// Counts trial F of the row into the table set at shared-window address TB (its outlier hash at ohash_addr).
// Returns the OR of this thread's row bytes.
template <int D, int NT, int F, uint32_t TB>
__device__ __forceinline__ uint32_t split_count(const uint8_t *cur, const uint8_t *up, uint32_t len, uint32_t shift, int bpp,
                                                uint32_t ohash_addr, uint32_t s_out_addr) {
    const uint32_t nwords = (len + 3) >> 2, lane = threadIdx.x & 31u;
    const uint32_t *cw = reinterpret_cast<const uint32_t *>(cur), *uw = reinterpret_cast<const uint32_t *>(up);
    uint32_t nz = 0;
    for (uint32_t k = threadIdx.x; k < ((nwords + 31u) & ~31u); k += NT) {
        const int kk = (int)k;
        uint32_t t = 0, x = 0;
        if (k < nwords) {
            // row words k-1-D .. k are addressable: the slot has 32 zero bytes in front of the row
            uint32_t l = 0, u = 0, ul = 0;
            x = cw[kk];
            if (F == 1 || F == 3 || F == 4)
                l = D == 2 ? cw[kk - 2] : D == 1 ? __funnelshift_l(cw[kk - 2], cw[kk - 1], shift)
                                                 : __funnelshift_l(cw[kk - 1], cw[kk], shift);
            if (F >= 2) u = uw[kk];
            if (F == 4)
                ul = D == 2 ? uw[kk - 2] : D == 1 ? __funnelshift_l(uw[kk - 2], uw[kk - 1], shift)
                                                  : __funnelshift_l(uw[kk - 1], uw[kk], shift);
            t = F == 0 ? x : __vsub4(__vadd4(x, 0x10101010u), pred4<F>(l, u, ul)); // range codes (raw bytes for trial 0)
        }
        // code of the stream byte in front of this word: the neighbouring lane's last byte; lane 0: the filter id in front
        // of word 0, else one scalar pixel
        uint32_t p = __shfl_up_sync(0xffffffffu, t >> 24, 1);
        if (lane == 0)
            p = k == 0 ? (F == 0 ? 0u : (uint32_t)F + 16u)
                       : (k < nwords ? ((residual_at(F, cur, up, 4 * kk - 1, bpp) + (F == 0 ? 0u : 16u)) & 255u) : 0u);
        if (k >= nwords) continue;
        const int valid = (int)len - 4 * kk; // >= 1
        if (F == 0) nz |= valid >= 4 ? x : (x & tail_mask(valid));
        const uint32_t cb = t;
        const uint32_t ca = prmt(t, p, 0x2104u); // byte m: m ? t.byte[m-1] : p
        const uint32_t caA = ca & 0x1F1F1F1Fu;   // first code as a table-A index (trial 0: the coarse first byte itself)
        const uint32_t cbx = cb ^ caA;           // only the low five bits of b change: a bijection per a
        // A: offset a << 10 | (b ^ a) << 2       -> high byte a << 2 | (b ^ a) >> 6, low byte ((b ^ a) & 63) << 2
        // B: offset a << 7 | ((b ^ a) & 31) << 2 -> high byte a >> 1,                low byte (a & 1) << 7 | ((b ^ a) & 31) << 2
        // Both sets of address bytes are formed for the whole word and merged bytewise by "a < 32".  The instruction's
        // immediate is the set's address + 0x8000 and the high byte carries bit 7 inverted: prmt's sign replication then
        // yields (offset - 0x8000) for A and (offset + 0x8000 - 0x8000) + 0x8000 - 0x8000 ... i.e. A lands in the first
        // 32 KB of the set and B in the second.
        const uint32_t sh2 = cbx << 2;
        const uint32_t hiA = caA * 4u + (((cbx >> 6) & 0x03030303u) | 0x80808080u), loA = sh2 & 0xFCFCFCFCu;
        const uint32_t hiB = (ca >> 1) & 0x7F7F7F7Fu, loB = ((ca & 0x01010101u) << 7) | (sh2 & 0x7C7C7C7Cu);
        uint32_t hi = hiA, lo = loA, outl = 0;
        if (F != 0) {
            const uint32_t zA = ca & 0xE0E0E0E0u, zB = cb & 0xE0E0E0E0u;
            const uint32_t nzA = (zA | (zA << 1) | (zA << 2)) & 0x80808080u; // 0x80 in every byte with a >= 32
            const uint32_t nzB = (zB | (zB << 1) | (zB << 2)) & 0x80808080u;
            const uint32_t mB = signmask4(nzA);                              // 0xFF where the window goes to B (or nowhere)
            hi = (hiA & ~mB) | (hiB & mB);
            lo = (loA & ~mB) | (loB & mB);
            outl = nzA & nzB; // windows with both residuals far from zero
        }
        if (valid >= 4 && outl == 0) {
            smem_inc_at<TB + 0x8000u>(prmt(lo, hi, 0xCC40u));
            smem_inc_at<TB + 0x8000u>(prmt(lo, hi, 0xDD51u));
            smem_inc_at<TB + 0x8000u>(prmt(lo, hi, 0xEE62u));
            smem_inc_at<TB + 0x8000u>(prmt(lo, hi, 0xFF73u));
        } else {
            for (int j = 0; j < 4 && j < valid; j++) {
                const uint32_t hb = (hi >> (8 * j)) & 255u, lb = (lo >> (8 * j)) & 255u;
                if ((outl >> (8 * j)) & 0x80u)
                    split_outlier((((ca >> (8 * j)) & 255u) << 8) | ((cb >> (8 * j)) & 255u), (uint32_t)F, ohash_addr, s_out_addr);
                else
                    smem_inc(TB + 0x8000u + (uint32_t)(((int32_t)(int8_t)hb) * 256) + lb);
            }
        }
    }
    return nz;
}

// Sweeps (and clears) one table set and its outlier hash: the scores of trial f are added to s_bgc[f] / s_bge[f] -- unless
// the trial is flagged (its hash overflowed): then the set is only cleared and the trial is recounted exactly later.
template <int NT>
__device__ __forceinline__ void split_sweep(uint32_t *set, uint32_t *oh, uint32_t f, const Smem &sr) {
    uint4 *s4 = reinterpret_cast<uint4 *>(set);
    uint32_t d = 0, e = 0;
    for (uint32_t q = threadIdx.x; q < SPLIT_SET / 16; q += NT) {
        const uint4 a = s4[q];
        if (a.x | a.y | a.z | a.w) {
            s4[q] = make_uint4(0, 0, 0, 0);
            d += (a.x != 0) + (a.y != 0) + (a.z != 0) + (a.w != 0);
            e += ilog2i_z(a.x) + ilog2i_z(a.y) + ilog2i_z(a.z) + ilog2i_z(a.w);
        }
    }
    if (threadIdx.x < SPLIT_OSLOTS) {
        const uint32_t v = oh[threadIdx.x];
        if (v) {
            oh[threadIdx.x] = 0;
            d += 1;
            e += ilog2i(v & 0xFFFFu);
        }
    }
    d = warp_sum(d);
    e = warp_sum(e);
    if ((threadIdx.x & 31) == 0 && d && sr.s_out_()[f] == 0) {
        atomicAdd(&sr.s_bgc_()[f], d);
        atomicAdd(&sr.s_bge_()[f], e);
    }
}

// One row.  Returns "some byte of the row is non-zero"; on return the scores of all five trials are final in
// s_bgc / s_bge (exact, trial 0 included if the strategies need it) and every thread is behind a barrier.
template <int D, int NT>
__device__ __forceinline__ bool bigram_row_split(const uint8_t *cur, const uint8_t *up, uint32_t len, uint32_t shift, int bpp,
                                                 const Smem &sr, bool want_bg, bool want_be) {
    constexpr uint32_t T0 = SMEM_WIN + SMEM_HDR, T1 = T0 + SPLIT_SET, O0 = T0 + 2 * SPLIT_SET, O1 = O0 + SPLIT_OSLOTS * 4;
    uint32_t *set0 = sr.table, *set1 = sr.table + SPLIT_SET / 4, *oh0 = sr.ohash, *oh1 = sr.ohash + SPLIT_OSLOTS;
    const uint32_t so = smem_u32(sr.s_out_());
    const uint32_t nz = split_count<D, NT, 0, T0>(cur, up, len, shift, bpp, O0, so);
    const bool hz = __syncthreads_or(nz != 0) != 0;
    split_count<D, NT, 1, T1>(cur, up, len, shift, bpp, O1, so);
    split_sweep<NT>(set0, oh0, 0, sr);
    __syncthreads();
    split_count<D, NT, 2, T0>(cur, up, len, shift, bpp, O0, so);
    split_sweep<NT>(set1, oh1, 1, sr);
    __syncthreads();
    split_count<D, NT, 3, T1>(cur, up, len, shift, bpp, O1, so);
    split_sweep<NT>(set0, oh0, 2, sr);
    __syncthreads();
    split_count<D, NT, 4, T0>(cur, up, len, shift, bpp, O0, so);
    split_sweep<NT>(set1, oh1, 3, sr);
    __syncthreads();
    split_sweep<NT>(set0, oh0, 4, sr);
    __syncthreads();
    if (!hz) return false; // all-zero row: None, whatever the scores (png/mod.rs:398-404)
    // exact recounts, in the memory of the (clean) table sets: flagged trials, and trial 0 when its bounds do not rule it out
    uint32_t *bc = sr.s_bgc_(), *be = sr.s_bge_();
    const uint32_t wpt = (((len + 3) >> 2) + NT - 1) / NT;
    for (int f = 1; f < 5; f++) {
        if (sr.s_out_()[f] == 0) continue; // uniform
        // (the sweeps added nothing for a flagged trial)
        if (f == 1) { if (wpt <= 4) bigram_trial<D, NT, 1, 4>(cur, up, len, shift, bpp, sr.table, bc, be); else bigram_trial<D, NT, 1, 8>(cur, up, len, shift, bpp, sr.table, bc, be); }
        if (f == 2) { if (wpt <= 4) bigram_trial<D, NT, 2, 4>(cur, up, len, shift, bpp, sr.table, bc, be); else bigram_trial<D, NT, 2, 8>(cur, up, len, shift, bpp, sr.table, bc, be); }
        if (f == 3) { if (wpt <= 4) bigram_trial<D, NT, 3, 4>(cur, up, len, shift, bpp, sr.table, bc, be); else bigram_trial<D, NT, 3, 8>(cur, up, len, shift, bpp, sr.table, bc, be); }
        if (f == 4) { if (wpt <= 4) bigram_trial<D, NT, 4, 4>(cur, up, len, shift, bpp, sr.table, bc, be); else bigram_trial<D, NT, 4, 8>(cur, up, len, shift, bpp, sr.table, bc, be); }
    }
    const uint32_t dmin = min(min(bc[1], bc[2]), min(bc[3], bc[4]));
    const int32_t emax = max(max((int32_t)be[1], (int32_t)be[2]), max((int32_t)be[3], (int32_t)be[4]));
    const bool n0 = (want_bg && bc[0] <= dmin) || (want_be && (int32_t)be[0] >= emax); // uniform
    if (n0) {
        __syncthreads(); // everybody has read the bounds
        if (threadIdx.x == 0) { bc[0] = 0; be[0] = 0; }
        __syncthreads();
        if (wpt <= 4) bigram_trial<D, NT, 0, 4>(cur, up, len, shift, bpp, sr.table, bc, be);
        else bigram_trial<D, NT, 0, 8>(cur, up, len, shift, bpp, sr.table, bc, be);
    }
    return true;
}

// ---- writing the chosen row --------------------------------------------------------------------------------------
// out_row points at the filter-type byte of this row in the filtered stream.  Interior 16-byte granules of the
// output are produced in the output's alignment frame; the (at most 31) edge bytes are written one by one.
// `rot` rotates the thread->granule assignment so that the write passes of several strategies of one row land on
// different warps (a 4 KB row has 256 granules: 8 of the 32 warps of a 1024-thread CTA).
template <int NT, bool WIDE>
__device__ __forceinline__ void write_row(int f, const uint8_t *cur, const uint8_t *up, uint32_t len, int bpp,
                                          uint8_t *out_row, int rot = 0) {
    const int tid_r = ((int)threadIdx.x - rot) & (NT - 1);
    const int phase = (int)(reinterpret_cast<uintptr_t>(out_row) & 15u);
    const int total = (int)len + 1;                 // stream positions p = 0..len
    const int q_full_end = (total + phase) >> 4;    // granules [1, q_full_end) are complete and hold residuals only
    uint8_t *base = out_row - phase;
    for (int q = 1 + tid_r; q < q_full_end; q += NT) {
        const int i0 = 16 * q - phase - 1; // row byte of the granule's first position (>= 0)
        const uint4 x = ld_unaligned16<WIDE>(cur + i0);
        uint4 r;
        if (f == 0) {
            r = x;
        } else if (f == 1) {
            const uint4 l = ld_unaligned16<WIDE>(cur + i0 - bpp);
            r.x = __vsub4(x.x, l.x); r.y = __vsub4(x.y, l.y); r.z = __vsub4(x.z, l.z); r.w = __vsub4(x.w, l.w);
        } else if (f == 2) {
            const uint4 u = ld_unaligned16<WIDE>(up + i0);
            r.x = __vsub4(x.x, u.x); r.y = __vsub4(x.y, u.y); r.z = __vsub4(x.z, u.z); r.w = __vsub4(x.w, u.w);
        } else if (f == 3) {
            const uint4 l = ld_unaligned16<WIDE>(cur + i0 - bpp), u = ld_unaligned16<WIDE>(up + i0);
            r.x = __vsub4(x.x, __vhaddu4(l.x, u.x)); r.y = __vsub4(x.y, __vhaddu4(l.y, u.y));
            r.z = __vsub4(x.z, __vhaddu4(l.z, u.z)); r.w = __vsub4(x.w, __vhaddu4(l.w, u.w));
        } else {
            const uint4 l = ld_unaligned16<WIDE>(cur + i0 - bpp), u = ld_unaligned16<WIDE>(up + i0), ul = ld_unaligned16<WIDE>(up + i0 - bpp);
            r.x = __vsub4(x.x, paeth4(l.x, u.x, ul.x)); r.y = __vsub4(x.y, paeth4(l.y, u.y, ul.y));
            r.z = __vsub4(x.z, paeth4(l.z, u.z, ul.z)); r.w = __vsub4(x.w, paeth4(l.w, u.w, ul.w));
        }
        __stcs(reinterpret_cast<uint4 *>(base + 16 * (size_t)q), r);
    }
    // edges: positions [0, head_end) of granule 0 and [tail_begin, total) of the last, partial granule
    const int head_end = min(16 - phase, total);
    const int tail_begin = max(16 * q_full_end - phase, head_end);
    const int n_edge = head_end + (total - tail_begin);
    for (int e = tid_r; e < n_edge; e += NT) {
        const int p = e < head_end ? e : tail_begin + (e - head_end);
        out_row[p] = p == 0 ? (uint8_t)f : (uint8_t)residual_at(f, cur, up, p - 1, bpp);
    }
}

// cooperative, alignment-agnostic row load (rows that TMA cannot take: not 16-byte aligned / sized)
template <int NT>
__device__ __noinline__ void coop_load_row(uint8_t *dst, const uint8_t *src, uint32_t len) {
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    const uint32_t *g = reinterpret_cast<const uint32_t *>(src - mis);
    const uint32_t nsrc = (len + mis + 3) >> 2, ndst = (len + 3) >> 2, s = 8 * mis;
    uint32_t *d = reinterpret_cast<uint32_t *>(dst);
    for (uint32_t k = threadIdx.x; k < ndst; k += NT) {
        const uint32_t lo = __ldg(g + k);
        const uint32_t hi = (mis && k + 1 < nsrc) ? __ldg(g + k + 1) : 0u;
        d[k] = __funnelshift_r(lo, hi, s);
    }
}

// The filter a strategy uses for one row (png/mod.rs:385-421): fixed for Basic / Predefined; for the heuristics the
// winner of the five trials under the reference's strict comparisons in RowFilter::ALL order (lowest id wins ties),
// None for all-zero rows.  `exact0`: the exact Bigrams/BigEnt scores of trial 0 are in slot 5 (bigram_exact0).
__device__ __noinline__ int decide_row(const DevStrategy &st, uint32_t r, uint32_t len, const Smem &sr, bool heur, bool exact0) {
    int f = 0;
    if (st.kind == OXI_BASIC) {
        f = st.basic;
    } else if (st.kind == OXI_PREDEFINED) {
        f = r < st.n_predefined ? st.predefined[r] : 0; // png/mod.rs:389
    } else if (heur) {
        if (st.kind == OXI_MINSUM) { // smaller is better, strict '<' (strategies.rs:95)
            uint32_t best = 0xFFFFFFFFu;
            const uint32_t nbytes = 16 * ((len + 15) >> 4);
#pragma unroll
            for (int q = 0; q < 5; q++) {
                const uint32_t s = 128u * nbytes - sr.s_min_()[q] + (uint32_t)q;
                if (s < best || q == 0) { best = s; f = q; }
            }
        } else if (st.kind == OXI_BIGRAMS) { // fewer distinct bigrams, strict '<' (strategies.rs:182-189)
            uint32_t best = 0;
#pragma unroll
            for (int q = 0; q < 5; q++) {
                const uint32_t s = sr.s_bgc_()[q == 0 && exact0 ? 5 : q];
                if (q == 0 || s < best) { best = s; f = q; }
            }
        } else { // Entropy / BigEnt: larger is better, strict '>' on i32 (strategies.rs:143,221)
            const uint32_t *sc = st.kind == OXI_ENTROPY ? sr.s_ent_() : sr.s_bge_();
            int32_t best = 0;
#pragma unroll
            for (int q = 0; q < 5; q++) {
                const int32_t s = (int32_t)sc[q == 0 && exact0 && st.kind == OXI_BIGENT ? 5 : q];
                if (q == 0 || s > best) { best = s; f = q; }
            }
        }
    }
    return f;
}

// ---- the kernel ------------------------------------------------------------------------------------------------------
template <int D, int SC>
__global__ void __launch_bounds__(threads_for(SC), (SC & SC_WIDE) ? 2 : (SC & SC_HALF) ? ((SC & SC_ENTROPY) ? HALF_OCC - 1 : HALF_OCC) : (SC & SC_BIGRAM) ? 1 : (SC == SC_MINSUM || SC == 0) ? 3 : 2) k_fast(const __grid_constant__ FastParams P) {
    constexpr int NT = threads_for(SC);
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t NS = P.nslot;
    const bool split = SC == SC_BIGRAM && P.small_tab == 2u;
    const Smem sm = carve<SC>(smem_raw, P.rb, NS, P.small_tab == 1u ? P.h0_slots : 0u, split);
    const int tid = threadIdx.x;
    const int bpp = (int)P.g.bpp;
    const uint32_t shift = P.shift;
    if (smem_u32(smem_raw) != SMEM_WIN) __trap(); // the table atomics carry their shared-window address as an immediate

    if (tid == 0) {
        for (uint32_t i = 0; i < NS; i++) mbar_init(&sm.bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 3 * (int)SCORE_WORDS) sm.sc[tid] = 0; // the three rotating score sets
    for (uint32_t i = tid; i < NS * PAD / 4; i += NT) // zero the left pad of each slot (never written again)
        reinterpret_cast<uint32_t *>(sm.rows + (i / (PAD / 4)) * (size_t)P.rb)[i % (PAD / 4)] = 0;
    if (SC & SC_ENTROPY)
        for (uint32_t i = tid; i < 5 * HIST_BINS; i += NT) sm.hist[i] = 0;
    if (SC & SC_BIGRAM) {
        for (uint32_t i = tid; i < (P.small_tab == 1u ? P.h0_slots : split ? SPLIT_BYTES / 4 : 32768u); i += NT) sm.table[i] = 0;
        if (P.small_tab == 1u)
            for (uint32_t i = tid; i < Dense<rb_of(SC)>::SMALL / 4; i += NT) sm.small[i] = 0;
    }
    __syncthreads();

    uint32_t parity = 0;  // bit s = phase to wait for on slot s
    uint32_t set_cur = 0; // score set of the next scored row (three sets rotate)
    const uint64_t total_items = P.n_images * (uint64_t)P.items_per_image;
    bool any_heur = false, want_bg = false, want_be = false;
    for (int j = 0; j < P.set.k; j++) {
        any_heur |= (P.set.s[j].kind != OXI_BASIC && P.set.s[j].kind != OXI_PREDEFINED);
        want_bg |= P.set.s[j].kind == OXI_BIGRAMS;
        want_be |= P.set.s[j].kind == OXI_BIGENT;
    }
    // With a 4-slot ring a slot is refilled two iterations after its last reader, and every scored row contains at
    // least one CTA-wide barrier, so the end-of-row barrier is only needed for the 3-slot ring / unscored rows.
    const bool defer_pf = (SC & SC_BIGRAM) && ((SC & SC_HALF) || P.small_tab == 1u) && any_heur && NS < 4;
    const bool row_end_barrier = (NS < 4 && !defer_pf) || SC == 0 || !any_heur;

    for (uint64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
        const uint64_t img = item / P.items_per_image;
        const uint32_t it = (uint32_t)(item % P.items_per_image);
        if (P.img_class && P.img_class[img] != P.want_class) continue; // another launch (other table range) takes this image
        int p = 0;
#pragma unroll
        for (int i = 1; i < 7; i++)
            if (i < (int)P.g.n_pass && it >= P.band_start[i]) p = i;
        const PassDesc pd = P.g.pass[p];
        const uint32_t nb = P.band_start[p + 1] - P.band_start[p], b = it - P.band_start[p];
        const uint32_t k0 = (uint32_t)(((uint64_t)b * pd.nrows) / nb), k1 = (uint32_t)(((uint64_t)(b + 1) * pd.nrows) / nb);
        if (k0 >= k1) continue;
        const uint32_t len = pd.len;
        const uint8_t *src0 = P.data + img * P.g.data_len + pd.in_base; // first row of the pass
        // row k of the pass starts at band_off + k * (len + 1) of a strategy's output (in_base + k*len bytes of earlier
        // rows of the image plus one filter byte per earlier row) and is row used_off + k of its filters_used
        const uint64_t band_off = img * P.g.out_len + pd.in_base + pd.row0;
        const uint64_t used_off = img * P.g.total_rows + pd.row0;
        const uint32_t g32 = (((len + 16) >> 4) + 31) & ~31u; // 16-byte granules of an output row, in whole warps
        const bool tma = ((reinterpret_cast<uintptr_t>(src0) | len) & 15u) == 0;

        // prologue: previous row (or zeros at a pass start, png/mod.rs:375-378) into slot 0, first row into slot 1
        if (k0 == 0) {
            uint4 *z = reinterpret_cast<uint4 *>(sm.rows + PAD);
            for (uint32_t i = tid; i < (len + 15) / 16; i += NT) z[i] = make_uint4(0, 0, 0, 0);
            fence_proxy_async();
        }
        if (tma) {
            if (tid == 0) {
                if (k0 > 0) {
                    mbar_expect_tx(&sm.bar[0], len);
                    tma_load_1d(sm.rows + PAD, src0 + (size_t)(k0 - 1) * len, len, &sm.bar[0]);
                }
                mbar_expect_tx(&sm.bar[1], len);
                tma_load_1d(sm.rows + P.rb + PAD, src0 + (size_t)k0 * len, len, &sm.bar[1]);
            }
            if (k0 > 0) {
                mbar_wait(&sm.bar[0], parity & 1u);
                parity ^= 1u;
            }
        } else {
            if (k0 > 0) coop_load_row<NT>(sm.rows + PAD, src0 + (size_t)(k0 - 1) * len, len);
            coop_load_row<NT>(sm.rows + P.rb + PAD, src0 + (size_t)k0 * len, len);
        }
        __syncthreads();

        uint32_t sp = 0, sc_ = 1, sn = 2; // ring slots of the previous, current and next row
        for (uint32_t k = k0; k < k1; k++) {
            const uint8_t *cur = sm.rows + sc_ * (size_t)P.rb + PAD;
            const uint8_t *up = sm.rows + sp * (size_t)P.rb + PAD;
            // stream the next row in behind the current one: here, or (3-slot ring, small-table bigram scorer) behind the
            // scorer's first barrier, which makes the end-of-row barrier unnecessary
            Prefetch pf;
            pf.on = k + 1 < k1;
            pf.tma = tma;
            pf.dst = sm.rows + sn * (size_t)P.rb + PAD;
            pf.src = src0 + (size_t)(k + 1) * len;
            pf.len = len;
            pf.bar = &sm.bar[sn];
            if (!defer_pf) {
                issue_prefetch<NT>(pf);
                pf.on = false;
            }
            if (tma) {
                mbar_wait(&sm.bar[sc_], (parity >> sc_) & 1u);
                parity ^= 1u << sc_;
            }

            const uint32_t r = pd.row0 + k;
            const uint64_t row_off = band_off + (uint64_t)k * (len + 1u); // of this row in a strategy's filtered stream
            bool heur = false, need0 = false, decided = false;
            // score set of this row; the set of the row after next is cleared here, before this row's barriers:
            // its last readers (decisions two rows ago) are behind the previous row's barriers
            Smem sr = sm;
            sr.sc = sm.sc + set_cur * SCORE_WORDS;
            const uint32_t set_next = set_cur == 2u ? 0u : set_cur + 1u;
            uint8_t *fsel = reinterpret_cast<uint8_t *>(sr.s_min_() + 5); // 12 bytes: the filter chosen per strategy
            if (SC != 0 && any_heur) {
                if (tid < (int)SCORE_WORDS) sm.sc[set_next * SCORE_WORDS + tid] = 0;
                uint32_t nz = 0;
                bool small_flow = false, split_done = false;
                if (SC & SC_MINSUM) minsum_pass<D, NT>(cur, up, len, shift, sr.s_min_(), P.two31);
                if (SC & SC_ENTROPY) {
                    entropy_pass<D, NT>(cur, up, len, shift, sm.hist, P.four);
                    __syncthreads();
                    entropy_score<NT>(sm.hist, sr.s_ent_());
                }
                if (SC & SC_BIGRAM) {
                    const uint32_t wpt = (((len + 3) >> 2) + NT - 1) / NT; // row words per thread
                    if (SC == SC_BIGRAM && split) {
                        heur = bigram_row_split<D, NT>(cur, up, len, shift, bpp, sr, want_bg, want_be);
                        split_done = true;
                    }
                    else if ((SC & SC_HALF) || P.small_tab == 1u) {
                        nz |= bigram_row_small<D, NT, SMEM_WIN + SMEM_HDR + ((SC & SC_ENTROPY) ? HIST_BYTES : 0u), rb_of(SC)>(cur, up, len, shift, bpp, sr, P.h0_slots, pf);
                        small_flow = true;
                    }
                    else if (wpt <= 1) bigram_row<D, NT, 1>(cur, up, len, shift, bpp, sm.table, sr.s_bgc_(), sr.s_bge_());
                    else if (wpt <= 2) bigram_row<D, NT, 2>(cur, up, len, shift, bpp, sm.table, sr.s_bgc_(), sr.s_bge_());
                    else if (wpt <= 4) bigram_row<D, NT, 4>(cur, up, len, shift, bpp, sm.table, sr.s_bgc_(), sr.s_bge_());
                    else bigram_row<D, NT, 8>(cur, up, len, shift, bpp, sm.table, sr.s_bgc_(), sr.s_bge_());
                }
                if ((SC & SC_BIGRAM) && small_flow) {
                    // Every warp has added its scores; the warp that arrives last decides for all strategies before
                    // the CTA-wide barrier, the others only read its choices behind the barrier.  The arrival counter
                    // (s_out[0]) carries "some byte of the row is non-zero" in its upper half.
                    const bool wnz = __any_sync(0xffffffffu, nz != 0);
                    uint32_t arrived = 0;
                    if ((tid & 31) == 0) {
                        __threadfence_block();
                        arrived = atomicAdd(&sr.s_out_()[0], 1u | (wnz ? 0x10000u : 0u)) + (1u | (wnz ? 0x10000u : 0u));
                    }
                    arrived = __shfl_sync(0xffffffffu, arrived, 0);
                    if ((arrived & 0xFFFFu) == NT / 32) {
                        __threadfence_block();
                        const bool hz = (arrived >> 16) != 0;
                        bool n0 = false;
                        if (hz) {
                            // scores 0 are the coarse bounds of trial 0 (bigram_row_small); the exact pass runs only
                            // when they leave open that None wins (it wins ties: d0 <= min d_f, e0 >= max e_f)
                            const volatile uint32_t *bc = sr.s_bgc_(), *be = sr.s_bge_();
                            const uint32_t dmin = min(min(bc[1], bc[2]), min(bc[3], bc[4]));
                            const int32_t emax = max(max((int32_t)be[1], (int32_t)be[2]), max((int32_t)be[3], (int32_t)be[4]));
                            n0 = (want_bg && bc[0] <= dmin) || (want_be && (int32_t)be[0] >= emax);
                        }
                        const int j = tid & 31;
                        if (j < P.set.k && !n0) fsel[j] = (uint8_t)decide_row(P.set.s[j], r, len, sr, hz, false);
                        if (j == 0) sr.s_out_()[5] = (hz ? 1u : 0u) | (n0 ? 2u : 0u);
                    }
                    __syncthreads();
                    const uint32_t fl = sr.s_out_()[5];
                    heur = (fl & 1u) != 0;
                    need0 = (fl & 2u) != 0;
                    if (need0) {
                        bigram_exact0<NT>(cur, len, sm.table, P.h0_slots, sr.s_bgc_(), sr.s_bge_());
                        if (tid < P.set.k) fsel[tid] = (uint8_t)decide_row(P.set.s[tid], r, len, sr, true, true);
                        __syncthreads();
                    }
                    decided = true;
                } else if (!split_done) {
                    // all-zero rows short-circuit to None (png/mod.rs:398-404); decided at the scorers' last barrier
                    if (SC & SC_MINSUM) { // S_0 = 128 per byte iff every byte is zero (see minsum_chunk)
                        __syncthreads();
                        heur = sr.s_min_()[0] != 128u * 16u * ((len + 15u) >> 4);
                    } else {
                        const uint4 *c4 = reinterpret_cast<const uint4 *>(cur);
                        const uint32_t nchunks = (len + 15) >> 4;
                        for (uint32_t c = tid; c < nchunks; c += NT) {
                            uint4 v = c4[c];
                            const int valid = (int)len - (int)(16 * c);
                            if (valid < 16) {
                                v.x &= tail_mask(valid); v.y &= tail_mask(valid - 4); v.z &= tail_mask(valid - 8);
                                v.w &= tail_mask(valid - 12);
                            }
                            nz |= v.x | v.y | v.z | v.w;
                        }
                        heur = __syncthreads_or(nz != 0) != 0;
                    }
                }
                set_cur = set_next;
            }

            for (int j2 = 0; j2 < P.set.k; j2++) {
                const DevStrategy &st = P.set.s[j2];
                // the write pass of strategy j2 runs on the warps starting at `rot`; the others have nothing to do for
                // it (a row has (len+16)/16 granules plus <= 31 edge bytes) and skip the decision as well
                const int rot = (int)((j2 * g32) & (NT - 1));
                if ((int)(((int)threadIdx.x - rot) & (NT - 1)) >= (int)g32 + 32) continue;
                const int f = decided ? (int)fsel[j2] : decide_row(st, r, len, sr, heur, false);
                write_row<NT, (SC & SC_ENTROPY) != 0>(f, cur, up, len, bpp, st.out + row_off, rot);
                if (tid == rot) st.filters_used[used_off + k] = (uint8_t)f;
            }
            if (row_end_barrier) __syncthreads(); // slot `sp` may be refilled by the next iteration's prefetch
            sp = sc_;
            sc_ = sn;
            sn = sn + 1 == NS ? 0 : sn + 1;
        }
        // rows loaded cooperatively were written through the generic proxy; the next band may refill the same slots by TMA
        // (async proxy): order the two before the barrier that ends the band
        if (!tma) fence_proxy_async();
        if (!row_end_barrier) __syncthreads(); // the next band's prologue rewrites slots 0 and 1
    }
}

// ---- optimize_alpha tier (filters/mod.rs:114-186, png/mod.rs:363-367,412-423) -----------------------------------------------
// With alpha optimisation every trial rewrites the colour bytes of fully transparent pixels so that they filter to zero,
// the five trials mutate ONE line cumulatively, and the next row is filtered against the winning trial's rewritten line:
// rows are serial, so the parallelism is across (image, strategy) chains and inside a row.  One 256-thread CTA walks a
// chain with everything in shared memory: two ring slots that receive the rows by TMA and serve as the mutable line, two
// buffers that alternate as "previous row as left by the winner" and "best line so far", and the scorer's table (one
// 256-bin histogram x 4 replicas, or one hash table of whole-row bigrams).  Trials are scored on the fly from (line,
// previous) with the byte-SIMD filters of the fast tier; the chosen row is written by the fast tier's write_row.
constexpr int ALPHA_NT = 256;
constexpr uint32_t ALPHA_HDR = 512;
enum { AK_FIXED = 0, AK_MINSUM = 1, AK_ENTROPY = 2, AK_BIGRAM = 3 };

struct AlphaParams {
    Geom g;
    StrategySet set;
    const uint8_t *data;
    uint64_t n_images;
    uint32_t rb;       // bytes of one row buffer (PAD + row + PAD)
    uint32_t hslots;   // AK_BIGRAM: slots of the hash table (power of two, > 2 * row bytes)
    uint32_t n_sel;    // strategies handled by this launch
    uint8_t sel[MAX_STRATEGIES];
};

__device__ __forceinline__ bool px_clear(const uint8_t *px, uint32_t cb, uint32_t bpp) { // alpha bytes all zero
    for (uint32_t k = cb; k < bpp; k++)
        if (px[k]) return false;
    return true;
}

// RowFilter::optimize_alpha for one row held in shared memory, any pixel format.  Alpha bytes are never written, so
// "transparent" is a fixed property of each pixel; every maximal run of transparent pixels is rewritten left to right by
// the thread that finds its first pixel (runs are disjoint; a run reads only the opaque pixel before it, the previous row
// and its own earlier writes).  RGBA8 (bpp 4, one pixel = one word) walks the runs with byte-SIMD word operations.
__device__ __noinline__ void alpha_rewrite_row(int f, uint8_t *line, const uint8_t *prev, uint32_t len, uint32_t bpp, uint32_t cb,
                                               uint32_t *sh_first) {
    if (f == 0) return; // "Assume transparent pixels already set to 0"
    const uint32_t npix = len / bpp;
    const int tid = threadIdx.x;
    const bool rgba8 = bpp == 4 && cb == 3;
    uint32_t *lw = reinterpret_cast<uint32_t *>(line);
    const uint32_t *pw = reinterpret_cast<const uint32_t *>(prev);
    if (f == 2) { // Up: copy the colour bytes of the pixel above
        if (rgba8) {
            for (uint32_t i = tid; i < npix; i += ALPHA_NT)
                if ((lw[i] >> 24) == 0) lw[i] = pw[i] & 0x00FFFFFFu;
        } else {
            for (uint32_t i = tid; i < npix; i += ALPHA_NT)
                if (px_clear(line + (size_t)i * bpp, cb, bpp))
                    for (uint32_t j = 0; j < cb; j++) line[(size_t)i * bpp + j] = prev[(size_t)i * bpp + j];
        }
        __syncthreads();
        return;
    }
    // first non-transparent pixel (only matters when pixel 0 is transparent -- a uniform condition), mod.rs:126-131
    uint32_t pv0 = 0;
    if (npix && px_clear(line, cb, bpp)) {
        if (tid == 0) *sh_first = 0xFFFFFFFFu;
        __syncthreads();
        uint32_t mine = 0xFFFFFFFFu;
        for (uint32_t i = tid; i < npix; i += ALPHA_NT)
            if (!px_clear(line + (size_t)i * bpp, cb, bpp)) { mine = i; break; }
        if (mine != 0xFFFFFFFFu) atomicMin(sh_first, mine);
        __syncthreads();
        const uint32_t first_opaque = *sh_first; // 0xFFFFFFFF: none -> prev = i (= 0)
        pv0 = first_opaque == 0xFFFFFFFFu ? 0u : first_opaque;
        __syncthreads(); // (sh_first is rewritten by the next trial)
    }
    for (uint32_t s0 = tid; s0 < npix; s0 += ALPHA_NT) {
        if (rgba8) {
            if ((lw[s0] >> 24) != 0) continue;
            if (s0 > 0 && (lw[s0 - 1] >> 24) == 0) continue; // not a run start
            // end of the run first (independent loads, four at a time), then the recurrence with the pixels above
            // fetched four ahead: the step-to-step chain is arithmetic only
            uint32_t e = s0 + 1;
            while (e + 4 <= npix) {
                const uint32_t a0 = lw[e] >> 24, a1 = lw[e + 1] >> 24, a2 = lw[e + 2] >> 24, a3 = lw[e + 3] >> 24;
                if (a0 | a1 | a2 | a3) break;
                e += 4;
            }
            while (e < npix && (lw[e] >> 24) == 0) e++;
            uint32_t left = s0 ? lw[s0 - 1] : 0u, ul = s0 ? pw[s0 - 1] : 0u;
            uint32_t i = s0;
            if (i == 0) { // first pixel of the row: mod.rs:126-131,162,173
                const uint32_t up = pw[0];
                uint32_t v = f == 1 ? lw[pv0] : f == 3 ? ((up >> 1) & 0x7F7F7F7Fu) : __vminu4(lw[pv0], up);
                v &= 0x00FFFFFFu;
                lw[0] = v;
                left = v;
                ul = up;
                i = 1;
            }
            if (f == 1) { // Sub: the colour of the pixel before the run, all the way
                const uint32_t v = left & 0x00FFFFFFu;
                for (; i < e; i++) lw[i] = v;
            } else {
                for (; i + 4 <= e; i += 4) {
                    const uint32_t u0 = pw[i], u1 = pw[i + 1], u2 = pw[i + 2], u3 = pw[i + 3];
                    const uint32_t v0 = (f == 3 ? __vhaddu4(left, u0) : paeth4(left, u0, ul)) & 0x00FFFFFFu;
                    const uint32_t v1 = (f == 3 ? __vhaddu4(v0, u1) : paeth4(v0, u1, u0)) & 0x00FFFFFFu;
                    const uint32_t v2 = (f == 3 ? __vhaddu4(v1, u2) : paeth4(v1, u2, u1)) & 0x00FFFFFFu;
                    const uint32_t v3 = (f == 3 ? __vhaddu4(v2, u3) : paeth4(v2, u3, u2)) & 0x00FFFFFFu;
                    lw[i] = v0; lw[i + 1] = v1; lw[i + 2] = v2; lw[i + 3] = v3;
                    left = v3;
                    ul = u3;
                }
                for (; i < e; i++) {
                    const uint32_t up = pw[i];
                    const uint32_t v = (f == 3 ? __vhaddu4(left, up) : paeth4(left, up, ul)) & 0x00FFFFFFu;
                    lw[i] = v;
                    left = v;
                    ul = up;
                }
            }
        } else {
            uint8_t *px = line + (size_t)s0 * bpp;
            if (!px_clear(px, cb, bpp)) continue;
            if (s0 > 0 && px_clear(px - bpp, cb, bpp)) continue; // not a run start
            for (uint32_t i = s0; i < npix; i++) {
                px = line + (size_t)i * bpp;
                if (i > s0 && !px_clear(px, cb, bpp)) break;
                const uint8_t *up = prev + (size_t)i * bpp;
                const uint32_t pv = i == 0 ? pv0 : i - 1;
                for (uint32_t j = 0; j < cb; j++) {
                    uint32_t v;
                    if (f == 1) {
                        v = line[(size_t)pv * bpp + j];
                    } else if (f == 3) {
                        v = i == 0 ? (uint32_t)(up[j] >> 1) : (((uint32_t)px[(int)j - (int)bpp] + up[j]) >> 1);
                    } else if (i == 0) {
                        const uint32_t a = line[(size_t)pv * bpp + j];
                        v = a < up[j] ? a : up[j];
                    } else {
                        v = paeth(px[(int)j - (int)bpp], up[j], up[(int)j - (int)bpp]);
                    }
                    px[j] = (uint8_t)v;
                }
            }
        }
    }
    __syncthreads();
}

// residual words of one 16-byte chunk of the row under filter f (run time), bytes past the row masked to zero
template <int D>
__device__ __forceinline__ void alpha_residual_chunk(int f, const uint8_t *cur, const uint8_t *up, uint32_t c, int valid,
                                                     uint32_t shift, uint32_t (&r)[4]) {
    const Chunk cw = load_chunk(cur, c), uw = load_chunk(up, c);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t x = cw.w[k + 2];
        uint32_t pr = 0;
        if (f == 1) pr = left_word<D>(cw, k, shift);
        else if (f == 2) pr = uw.w[k + 2];
        else if (f == 3) pr = __vhaddu4(left_word<D>(cw, k, shift), uw.w[k + 2]);
        else if (f == 4) pr = paeth4(left_word<D>(cw, k, shift), uw.w[k + 2], left_word<D>(uw, k, shift));
        r[k] = __vsub4(x, pr) & tail_mask(valid - 4 * k);
    }
}

// Score of trial f of the line (cur, up) under scorer KIND; every thread returns the same value.
// MinSum / Bigrams: lower is better; Entropy / BigEnt: higher is better as i32 (strategies.rs:95,143,189,221).
template <int D, int KIND>
__device__ __forceinline__ uint32_t alpha_score(int f, int kind, const uint8_t *cur, const uint8_t *up, uint32_t len, uint32_t shift,
                                                uint32_t *tab, uint32_t hslots, uint32_t *acc /* [2], this call's own, zero on entry */,
                                                bool coarse = false /* AK_BIGRAM: keys merged to their low five bits */) {
    const uint32_t nchunks = (len + 15) >> 4, lane = threadIdx.x & 31u;
    if (KIND == AK_MINSUM) {
        uint32_t a = 0; // sum over bytes of ||r| - 128| (see minsum_chunk); bytes past the row give 128 each
        for (uint32_t c = threadIdx.x; c < nchunks; c += ALPHA_NT) {
            uint32_t r[4];
            alpha_residual_chunk<D>(f, cur, up, c, (int)len - 16 * (int)c, shift, r);
#pragma unroll
            for (int k = 0; k < 4; k++) a += __vsadu4(r[k], 0x80808080u); // |r as i8| = 128 - |r - 128|
        }
        a = warp_sum(a);
        if (lane == 0 && a) atomicAdd(&acc[0], a);
        __syncthreads();
        return 128u * 16u * nchunks - acc[0] + (uint32_t)f; // + |f| of the filter-type byte
    }
    if (KIND == AK_ENTROPY) {
        const uint32_t rep = threadIdx.x & 3u;
        for (uint32_t c = threadIdx.x; c < nchunks; c += ALPHA_NT) {
            uint32_t r[4];
            const int valid = (int)len - 16 * (int)c;
            alpha_residual_chunk<D>(f, cur, up, c, valid, shift, r);
#pragma unroll
            for (int k = 0; k < 4; k++)
                for (int j = 0; j < 4; j++)
                    if (4 * k + j < valid) atomicAdd(&tab[rep * 256u + ((r[k] >> (8 * j)) & 255u)], 1u);
        }
        __syncthreads();
        uint32_t e = 0;
        {
            const uint32_t v = threadIdx.x; // one bin per thread (ALPHA_NT == 256)
            const uint32_t cnt = tab[v] + tab[256 + v] + tab[512 + v] + tab[768 + v] + (v == (uint32_t)f ? 1u : 0u);
            tab[v] = tab[256 + v] = tab[512 + v] = tab[768 + v] = 0;
            e = cnt ? ilog2i(cnt) : 0u;
        }
        e = warp_sum(e);
        if (lane == 0 && e) atomicAdd(&acc[0], e);
        __syncthreads();
        return acc[0];
    }
    // AK_BIGRAM: every window (t[p-1], t[p]), p = 1..len, t[0] = f.  Windows whose two residuals lie in [-16, 15] are
    // counted in a dense 32 x 32 table (the first 4 KB of `tab`), the others in the hash table behind it (CAS inserts are
    // an order of magnitude dearer): disjoint key sets, so distinct keys / sum ilog2i(count) add up exactly.
    uint32_t *hash = tab + 1024;
    const uint32_t hmask = hslots - 1, hshift = (uint32_t)__clz((int)hslots) + 1;
    bool any_out = false;
    for (uint32_t c = threadIdx.x; c < ((nchunks + 31u) & ~31u); c += ALPHA_NT) {
        uint32_t r[4] = {0, 0, 0, 0};
        const int valid = (int)len - 16 * (int)c;
        if (c < nchunks) alpha_residual_chunk<D>(f, cur, up, c, valid, shift, r);
        uint32_t p = __shfl_up_sync(0xffffffffu, r[3] >> 24, 1);
        if (lane == 0) p = c == 0 ? (uint32_t)f : (c < nchunks ? residual_at(f, cur, up, 16 * (int)c - 1, (int)(D * 4 + shift / 8)) : 0u);
        if (c >= nchunks) continue;
#pragma unroll
        for (int k = 0; k < 4; k++)
            for (int j = 0; j < 4; j++) {
                const uint32_t b = (r[k] >> (8 * j)) & 255u;
                if (4 * k + j < valid) {
                    const uint32_t pa = coarse ? (p & 31u) : ((p + 16u) & 255u), pb = coarse ? (b & 31u) : ((b + 16u) & 255u);
                    if (((pa | pb) & 0xE0u) == 0) {
                        atomicAdd(&tab[pa * 32u + ((pb + 5u * pa) & 31u)], 1u);
                    } else {
                        h0_add(hash, (p << 8) | b, hshift, hmask);
                        any_out = true;
                    }
                }
                p = b;
            }
    }
    const bool sweep_hash = __syncthreads_or(any_out) != 0;
    uint32_t d = 0, e = 0;
    {
        uint4 *t4 = reinterpret_cast<uint4 *>(tab) + threadIdx.x; // 1024 counters: one uint4 per thread
        const uint4 a = *t4;
        *t4 = make_uint4(0, 0, 0, 0);
        d += (a.x != 0) + (a.y != 0) + (a.z != 0) + (a.w != 0);
        e += ilog2i_z(a.x) + ilog2i_z(a.y) + ilog2i_z(a.z) + ilog2i_z(a.w);
    }
    if (sweep_hash) h0_sweep<ALPHA_NT>(hash, hslots, d, e);
    d = warp_sum(d);
    e = warp_sum(e);
    if (lane == 0 && d) {
        atomicAdd(&acc[0], d);
        atomicAdd(&acc[1], e);
    }
    __syncthreads();
    return kind == OXI_BIGRAMS ? acc[0] : acc[1];
}

template <int D, int KIND>
__global__ void __launch_bounds__(ALPHA_NT, 4) k_alpha(const __grid_constant__ AlphaParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);                 // 2 mbarriers
    uint32_t *acc = reinterpret_cast<uint32_t *>(smem_raw + 64);            // [6][2] score accumulators: one pair per scoring call of a row
    uint32_t *sh_first = reinterpret_cast<uint32_t *>(smem_raw + 128);
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem_raw + ALPHA_HDR);
    const uint32_t tab_bytes = KIND == AK_ENTROPY ? 4096u : KIND == AK_BIGRAM ? 4096u + P.hslots * 4u : 0u;
    uint8_t *bufs = smem_raw + ALPHA_HDR + tab_bytes; // ring slot 0, ring slot 1, keep A, keep B
    const int tid = threadIdx.x;
    const uint32_t bpp = P.g.bpp, cb = P.g.color_bytes, shift = 8 * (bpp % 4);
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 12) acc[tid] = 0;
    for (uint32_t i = tid; i < (tab_bytes + 4u * P.rb) / 4; i += ALPHA_NT) reinterpret_cast<uint32_t *>(smem_raw + ALPHA_HDR)[i] = 0;
    __syncthreads();
    uint32_t parity = 0;
    const uint64_t items = (uint64_t)P.n_sel * P.n_images;
    for (uint64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const DevStrategy &st = P.set.s[P.sel[it / P.n_images]];
        const uint64_t img = it % P.n_images;
        const uint8_t *in = P.data + img * P.g.data_len;
        for (uint32_t p = 0; p < P.g.n_pass; p++) {
            const PassDesc pd = P.g.pass[p];
            const uint32_t len = pd.len;
            const uint8_t *src0 = in + pd.in_base;
            const bool tma = ((reinterpret_cast<uintptr_t>(src0) | len) & 15u) == 0;
            uint8_t *keep_prev = bufs + 2 * (size_t)P.rb + PAD, *keep_best = bufs + 3 * (size_t)P.rb + PAD;
            // previous row of a pass's first row: zeros (png/mod.rs:375-378); also re-zero the tails of the ring slots
            // (a longer row of an earlier pass may have left bytes behind this pass's row length)
            for (uint32_t b = 0; b < 4; b++) {
                uint4 *z = reinterpret_cast<uint4 *>(bufs + b * (size_t)P.rb + PAD);
                for (uint32_t i = tid; i < (P.rb - 2 * PAD) / 16; i += ALPHA_NT) z[i] = make_uint4(0, 0, 0, 0);
            }
            fence_proxy_async();
            __syncthreads();
            // first row of the pass into ring slot 0
            if (tma) {
                if (tid == 0) {
                    mbar_expect_tx(&bar[0], len);
                    tma_load_1d(bufs + PAD, src0, len, &bar[0]);
                }
            } else {
                coop_load_row<ALPHA_NT>(bufs + PAD, src0, len);
            }
            for (uint32_t k = 0; k < pd.nrows; k++) {
                const uint32_t slot = k & 1u;
                uint8_t *w = bufs + slot * (size_t)P.rb + PAD;
                // the next row streams into the other slot (its last reader was row k-1, behind that row's final barrier)
                if (k + 1 < pd.nrows) {
                    uint8_t *nxt = bufs + (slot ^ 1u) * (size_t)P.rb + PAD;
                    if (tma) {
                        if (tid == 0) {
                            mbar_expect_tx(&bar[slot ^ 1u], len);
                            tma_load_1d(nxt, src0 + (size_t)(k + 1) * len, len, &bar[slot ^ 1u]);
                        }
                    } else {
                        coop_load_row<ALPHA_NT>(nxt, src0 + (size_t)(k + 1) * len, len);
                    }
                }
                if (tma) {
                    mbar_wait(&bar[slot], (parity >> slot) & 1u);
                    parity ^= 1u << slot;
                } else {
                    __syncthreads();
                }
                const uint32_t r = pd.row0 + k;
                uint8_t *out_row = st.out + img * P.g.out_len + pd.in_base + r + (uint64_t)k * len;
                int best_f = 0;
                const uint8_t *final_line = w;
                if (KIND == AK_FIXED) { // Basic / Predefined (png/mod.rs:382-394)
                    best_f = st.kind == OXI_BASIC ? st.basic : (r < st.n_predefined ? st.predefined[r] : 0);
                    alpha_rewrite_row(best_f, w, keep_prev, len, bpp, cb, sh_first);
                } else {
                    // all-zero rows short-circuit to None (png/mod.rs:398-404)
                    uint32_t nz = 0;
                    const uint4 *c4 = reinterpret_cast<const uint4 *>(w);
                    for (uint32_t c = tid; c < (len + 15) / 16; c += ALPHA_NT) { const uint4 v = c4[c]; nz |= v.x | v.y | v.z | v.w; }
                    if (tid < 12) acc[tid] = 0; // (the previous row's last reader is behind that row's final barrier)
                    if (__syncthreads_or(nz != 0)) {
                        uint32_t best = 0;
                        // Bigrams / BigEnt: trial 0 (the raw bytes: every window would take the hash path) is first scored
                        // on coarse keys, which bounds its score from the losing side (see bigram_row_small); its exact
                        // score is computed -- from the copy kept in keep_best -- only if a later trial does not beat the
                        // bound while None is still the best
                        bool best_is_bound = false;
                        const bool lower = st.kind == OXI_MINSUM || st.kind == OXI_BIGRAMS;
                        for (int f = 0; f < 5; f++) { // RowFilter::ALL order, cumulative rewriting of the one line
                            alpha_rewrite_row(f, w, keep_prev, len, bpp, cb, sh_first);
                            const bool coarse0 = KIND == AK_BIGRAM && f == 0;
                            const uint32_t sres = alpha_score<D, KIND>(f, st.kind, w, keep_prev, len, shift, tab, P.hslots, acc + 2 * f, coarse0);
                            bool wins = f == 0 || (lower ? sres < best : (int32_t)sres > (int32_t)best);
                            if (KIND == AK_BIGRAM && f > 0 && !wins && best_is_bound) {
                                best = alpha_score<D, KIND>(0, st.kind, keep_best, keep_prev, len, shift, tab, P.hslots, acc + 10, false);
                                best_is_bound = false;
                                wins = lower ? sres < best : (int32_t)sres > (int32_t)best;
                            }
                            if (wins) {
                                best = sres;
                                best_f = f;
                                best_is_bound = coarse0;
                                if (f < 4) { // keep the line as this trial left it (the last trial's line is `w` itself)
                                    const uint4 *s4 = reinterpret_cast<const uint4 *>(w);
                                    uint4 *d4 = reinterpret_cast<uint4 *>(keep_best);
                                    for (uint32_t c = tid; c < (len + 15) / 16; c += ALPHA_NT) d4[c] = s4[c];
                                }
                            }
                        }
                        __syncthreads();
                        final_line = best_f == 4 ? w : keep_best;
                    }
                }
                // write the chosen row from (its rewritten line, previous line); then it becomes the previous line
                write_row<ALPHA_NT, false>(best_f, final_line, keep_prev, len, (int)bpp, out_row);
                if (tid == 0) st.filters_used[img * P.g.total_rows + r] = (uint8_t)best_f;
                __syncthreads();
                if (final_line == keep_best) { // swap the keep buffers
                    uint8_t *t = keep_prev; keep_prev = keep_best; keep_best = t;
                } else { // the line lives in a ring slot that the next-but-one row will overwrite: copy it
                    const uint4 *s4 = reinterpret_cast<const uint4 *>(final_line);
                    uint4 *d4 = reinterpret_cast<uint4 *>(keep_prev);
                    for (uint32_t c = tid; c < (len + 15) / 16; c += ALPHA_NT) d4[c] = s4[c];
                    __syncthreads();
                }
            }
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------

int scorers_needed(const StrategySet &set) {
    int sc = 0;
    for (int j = 0; j < set.k; j++) {
        switch (set.s[j].kind) {
        case OXI_MINSUM: sc |= SC_MINSUM; break;
        case OXI_ENTROPY: sc |= SC_ENTROPY; break;
        case OXI_BIGRAMS: case OXI_BIGENT: sc |= SC_BIGRAM; break;
        default: break;
        }
    }
    return sc;
}
// Kernels are instantiated for the scorer sets that occur in practice; any other mixture runs the full kernel.
// (Entropy + bigrams is the reference's default trial set {None, Sub, Entropy, Bigrams}, options.rs:282-287)
int kernel_sc(int sc) {
    if (sc == 0 || sc == SC_MINSUM || sc == SC_ENTROPY || sc == SC_BIGRAM || sc == (SC_ENTROPY | SC_BIGRAM)) return sc;
    return SC_MINSUM | SC_ENTROPY | SC_BIGRAM;
}
uint32_t slot_bytes(const Geom &g) { return PAD + ((g.max_len + 15u) & ~15u) + PAD; }

using KernelFn = void (*)(const FastParams);
template <int D>
KernelFn pick_sc(int sc) {
    switch (sc) {
    case 0: return k_fast<D, 0>;
    case SC_MINSUM: return k_fast<D, SC_MINSUM>;
    case SC_ENTROPY: return k_fast<D, SC_ENTROPY>;
    case SC_BIGRAM: return k_fast<D, SC_BIGRAM>;
    case SC_BIGRAM | SC_HALF: return k_fast<D, SC_BIGRAM | SC_HALF>;
    case SC_BIGRAM | SC_HALF | SC_WIDE: return k_fast<D, SC_BIGRAM | SC_HALF | SC_WIDE>;
    case SC_ENTROPY | SC_BIGRAM: return k_fast<D, SC_ENTROPY | SC_BIGRAM>;
    case SC_ENTROPY | SC_BIGRAM | SC_HALF: return k_fast<D, SC_ENTROPY | SC_BIGRAM | SC_HALF>;
    default: return k_fast<D, SC_MINSUM | SC_ENTROPY | SC_BIGRAM>;
    }
}
KernelFn pick_kernel(int d, int sc) { return d == 0 ? pick_sc<0>(sc) : d == 1 ? pick_sc<1>(sc) : pick_sc<2>(sc); }


using AlphaFn = void (*)(const AlphaParams);
template <int D>
AlphaFn pick_alpha_kind(int kind) {
    switch (kind) {
    case AK_MINSUM: return k_alpha<D, AK_MINSUM>;
    case AK_ENTROPY: return k_alpha<D, AK_ENTROPY>;
    case AK_BIGRAM: return k_alpha<D, AK_BIGRAM>;
    default: return k_alpha<D, AK_FIXED>;
    }
}
AlphaFn pick_alpha(int d, int kind) { return d == 0 ? pick_alpha_kind<0>(kind) : d == 1 ? pick_alpha_kind<1>(kind) : pick_alpha_kind<2>(kind); }
int alpha_kind_of(int strategy_kind) {
    switch (strategy_kind) {
    case OXI_MINSUM: return AK_MINSUM;
    case OXI_ENTROPY: return AK_ENTROPY;
    case OXI_BIGRAMS: case OXI_BIGENT: return AK_BIGRAM;
    default: return AK_FIXED;
    }
}
uint32_t alpha_hslots(const Geom &g) {
    uint32_t h = 1024;
    while (h < 2 * g.max_len) h *= 2;
    return h;
}
uint32_t alpha_smem(const Geom &g, int kind) {
    return ALPHA_HDR + (kind == AK_ENTROPY ? 4096u : kind == AK_BIGRAM ? 4096u + alpha_hslots(g) * 4u : 0u) + 4u * slot_bytes(g);
}

std::mutex g_occ_mu;
std::map<std::pair<const void *, uint32_t>, int> g_occ; // (kernel, smem) -> resident CTAs per SM

__global__ void k_probe_smem_window(uint32_t *out) {
    extern __shared__ __align__(128) uint8_t probe_raw[];
    *out = smem_u32(probe_raw);
}

} // namespace

cudaError_t fast_tier_init(int device, uint32_t *smem_window) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    const int scs[9] = {0, SC_MINSUM, SC_ENTROPY, SC_BIGRAM, SC_BIGRAM | SC_HALF, SC_BIGRAM | SC_HALF | SC_WIDE,
                        SC_ENTROPY | SC_BIGRAM, SC_ENTROPY | SC_BIGRAM | SC_HALF, SC_MINSUM | SC_ENTROPY | SC_BIGRAM};
    for (int d = 0; d < 3; d++)
        for (int i = 0; i < 9; i++) {
            e = cudaFuncSetAttribute((const void *)pick_kernel(d, scs[i]), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)SMEM_LIMIT);
            if (e != cudaSuccess) return e;
        }
    for (int d = 0; d < 3; d++)
        for (int kind = AK_FIXED; kind <= AK_BIGRAM; kind++) {
            e = cudaFuncSetAttribute((const void *)pick_alpha(d, kind), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
            if (e != cudaSuccess) return e;
        }
    uint32_t *d_w = nullptr, w = 0;
    e = cudaMalloc((void **)&d_w, sizeof w);
    if (e != cudaSuccess) return e;
    k_probe_smem_window<<<1, 32, 1024>>>(d_w);
    e = cudaMemcpy(&w, d_w, sizeof w, cudaMemcpyDeviceToHost);
    cudaFree(d_w);
    if (e != cudaSuccess) return e;
    *smem_window = w;
    return cudaSuccess;
}
// optimize_alpha tier: images with an alpha channel whose rows (four staged copies + the scorer's table) fit shared memory
bool alpha_supported(const Geom &g, const StrategySet &set) {
    if (g.alpha_bytes == 0 || g.total_rows == 0) return false;
    if (g.bpp != 2 && g.bpp != 4 && g.bpp != 8) return false; // GA8, GA16 / RGBA8, RGBA16
    for (int j = 0; j < set.k; j++) {
        const int kind = alpha_kind_of(set.s[j].kind);
        if (alpha_smem(g, kind) > SMEM_LIMIT) return false;
        if (kind == AK_BIGRAM && g.max_len > 16384u) return false; // 15-bit counts in the hash entries
    }
    return true;
}

cudaError_t launch_alpha(const LaunchCtx &ctx, const Geom &g, const uint8_t *d_data, size_t n_images, const StrategySet &set) {
    for (int kind = AK_FIXED; kind <= AK_BIGRAM; kind++) {
        AlphaParams P;
        P.g = g;
        P.set = set;
        P.data = d_data;
        P.n_images = n_images;
        P.rb = slot_bytes(g);
        P.hslots = alpha_hslots(g);
        P.n_sel = 0;
        for (int j = 0; j < set.k; j++)
            if (alpha_kind_of(set.s[j].kind) == kind) P.sel[P.n_sel++] = (uint8_t)j;
        if (P.n_sel == 0) continue;
        const uint32_t smem = alpha_smem(g, kind);
        AlphaFn fn = pick_alpha((int)(g.bpp / 4), kind);
        int occ = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, ALPHA_NT, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
        const uint64_t items = (uint64_t)P.n_sel * n_images;
        uint64_t grid = (uint64_t)ctx.num_sms * occ;
        if (grid > items) grid = items;
        fn<<<(unsigned)grid, ALPHA_NT, smem, ctx.stream>>>(P);
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// the bigram and Entropy scorers address their tables as [register + immediate] from SMEM_WIN
bool fast_tier_usable(uint32_t smem_window) { return smem_window == SMEM_WIN; }

bool fast_supported(const Geom &g, const StrategySet &set) {
    if (g.alpha_bytes != 0 || g.total_rows == 0) return false;
    if (g.bpp != 1 && g.bpp != 2 && g.bpp != 3 && g.bpp != 4 && g.bpp != 6 && g.bpp != 8) return false;
    const int sc = kernel_sc(scorers_needed(set));
    if (smem_bytes_for(sc, slot_bytes(g), 0) > SMEM_LIMIT) return false;
    if ((sc & SC_BIGRAM) && g.max_len > 32768u) return false; // <= 8 row words per thread; u16 counters
    return true;
}

// One launch of the fast tier for scorer set `sc` (SC_HALF / SC_WIDE chosen by the caller).  img_class != nullptr: only the
// images of class want_class are processed (the other class belongs to a second launch).
static cudaError_t launch_fast_variant(const LaunchCtx &ctx, const Geom &g, const uint8_t *d_data, size_t n_images,
                                       const StrategySet &set, int sc, bool small_tab, bool split, uint32_t h0,
                                       const uint8_t *img_class, uint32_t want_class) {
    FastParams P;
    P.g = g;
    P.set = set;
    P.data = d_data;
    P.n_images = n_images;
    P.rb = slot_bytes(g);
    P.shift = 8 * (g.bpp % 4);
    P.four = 4;
    P.two31 = 0x80000000u;
    P.small_tab = split ? 2 : small_tab ? 1 : 0;
    P.h0_slots = small_tab ? h0 : 0;
    P.img_class = img_class;
    P.want_class = want_class;
    const int nt = threads_for(sc);
    const uint32_t half_occ = (sc & SC_WIDE) ? 2 : (sc & SC_ENTROPY) ? HALF_OCC - 1 : HALF_OCC;
    const uint32_t per_cta = (233472u / half_occ - 1024u) & ~127u; // shared memory of one of half_occ resident CTAs
    // a fourth ring slot removes the end-of-row barrier; take it when it does not cost occupancy (rows <= 16 KB)
    const uint32_t budget = (sc & SC_HALF) ? per_cta : SMEM_LIMIT;
    P.nslot = (sc != 0 && !split && P.rb <= 16448 && smem_bytes_for(sc, P.rb, P.h0_slots, 4) <= budget) ? 4 : 3;
    const uint32_t smem = smem_bytes_for(sc, P.rb, P.h0_slots, P.nslot, split);
    KernelFn fn = pick_kernel((int)(g.bpp / 4), sc);
    cudaError_t e; // (the dynamic shared-memory limit was raised once per device by fast_tier_init)
    int occ = 0;
    {
        std::lock_guard<std::mutex> lk(g_occ_mu);
        auto key = std::make_pair((const void *)fn, smem + (uint32_t)ctx.device * 0x1000000u);
        auto it = g_occ.find(key);
        if (it == g_occ.end()) {
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, nt, smem);
            if (e != cudaSuccess) return e;
            if (occ < 1) occ = 1;
            g_occ[key] = occ;
        } else {
            occ = it->second;
        }
    }
    // Bands never cross a pass (png/mod.rs:375-378) and cost one halo row each.  Memory-bound scorers take ~32-row
    // bands (3% extra reads); the shared-memory-bound ones (Entropy, bigrams) take ~8-row bands so that the static
    // grid-stride split stays balanced.  When the whole job is only a few waves, the band count is rounded to a
    // multiple of the resident CTA slots instead (every CTA gets the same number of near-equal bands).
    const uint64_t slots = (uint64_t)ctx.num_sms * occ;
    const uint64_t rows_total = (uint64_t)n_images * g.total_rows;
    uint32_t R = (sc & (SC_BIGRAM | SC_ENTROPY)) ? 8 : 32;
    if (rows_total / 32 >= 24 * slots) R = 32; // plenty of bands either way: fewer band prologues
    uint64_t want_items = (rows_total + R - 1) / R;
    if (want_items < 6 * slots) {
        uint64_t waves = want_items / slots;
        if (waves < 1) waves = 1;
        want_items = waves * slots;
        if (want_items * 2 > rows_total) want_items = rows_total / 2 ? rows_total / 2 : 1; // keep >= 2 rows per band
    }
    uint64_t per_image = (want_items + n_images - 1) / n_images;
    if (per_image < 1) per_image = 1;
    uint32_t acc = 0;
    for (uint32_t p = 0; p < 7; p++) {
        P.band_start[p] = acc;
        if (p < g.n_pass) {
            const uint64_t bytes = (uint64_t)g.pass[p].nrows * g.pass[p].len;
            uint64_t nb = (per_image * bytes + g.data_len - 1) / (g.data_len ? g.data_len : 1);
            const uint64_t max_nb = (g.pass[p].nrows + 1) / 2;
            if (nb > max_nb) nb = max_nb;
            if (nb < 1) nb = 1;
            acc += (uint32_t)nb;
        }
    }
    P.band_start[7] = acc;
    for (uint32_t p = g.n_pass; p < 7; p++) P.band_start[p] = acc;
    P.items_per_image = acc;
    const uint64_t items = (uint64_t)acc * n_images;
    uint64_t grid = items < slots ? items : slots;
    if (grid < 1) grid = 1;
    fn<<<(unsigned)grid, nt, smem, ctx.stream>>>(P);
    count_launch();
    return cudaGetLastError();
}

// Spread of an image's filtered rows, sampled: class 1 when in eight rows of the first pass more than SPREAD_PCT percent of
// the bytes have their Sub residual outside [-16, 15] and as many their Up residual -- such rows put most windows on the
// slow outlier path of the 32 x 32 tables (measured 6x slower at +-12 of noise); the 64 x 64 tables take them instead.
constexpr uint32_t SPREAD_PCT = 4;
__global__ void __launch_bounds__(256) k_spread_probe(Geom g, const uint8_t *__restrict__ data, uint64_t n_images,
                                                      uint8_t *__restrict__ img_class) {
    __shared__ uint32_t s_cnt[2];
    for (uint64_t img = blockIdx.x; img < n_images; img += gridDim.x) {
        if (threadIdx.x == 0) s_cnt[0] = s_cnt[1] = 0;
        __syncthreads();
        const PassDesc pd = g.pass[0];
        const uint8_t *base = data + img * g.data_len + pd.in_base;
        uint32_t cs = 0, cu = 0, total = 0;
        if (pd.nrows >= 2) {
            for (uint32_t sidx = 0; sidx < 8; sidx++) {
                uint32_t r = (uint32_t)(((uint64_t)pd.nrows * (2 * sidx + 1)) / 16);
                if (r == 0) r = 1;
                const uint8_t *row = base + (size_t)r * pd.len, *up = row - pd.len;
                for (uint32_t i = g.bpp + threadIdx.x; i < pd.len; i += 256) {
                    const uint32_t x = row[i];
                    const uint32_t rs = (x - row[i - g.bpp] + 16u) & 255u, ru = (x - up[i] + 16u) & 255u;
                    cs += rs >= 32u;
                    cu += ru >= 32u;
                }
                total += pd.len > g.bpp ? pd.len - g.bpp : 0;
            }
        }
        cs = warp_sum(cs);
        cu = warp_sum(cu);
        if ((threadIdx.x & 31) == 0 && (cs | cu)) {
            atomicAdd(&s_cnt[0], cs);
            atomicAdd(&s_cnt[1], cu);
        }
        __syncthreads();
        if (threadIdx.x == 0)
            img_class[img] = (uint64_t)min(s_cnt[0], s_cnt[1]) * 100u > (uint64_t)SPREAD_PCT * total ? 1 : 0;
        __syncthreads();
    }
}

cudaError_t launch_fast(const LaunchCtx &ctx, const Geom &g, const uint8_t *d_data, size_t n_images,
                        const StrategySet &set) {
    int sc = kernel_sc(scorers_needed(set));
    const uint32_t rb = slot_bytes(g);
    // bigram scorer, rows <= 16384 bytes: dense tables + outlier hashes + hash of whole rows (>= 2 slots per window)
    uint32_t h0 = 1024;
    while (h0 < 2 * g.max_len) h0 *= 2;
    const bool small_tab = (sc & SC_BIGRAM) && g.max_len <= 16384 && smem_bytes_for(sc, rb, h0) <= SMEM_LIMIT;
    // longer rows (<= 32 KB; mostly 16-bit images): the split-table layout, see bigram_row_split
    const bool split = sc == SC_BIGRAM && !small_tab && g.max_len <= 32768 && smem_bytes_for(sc, rb, 0, 3, true) <= SMEM_LIMIT;
    // ... and with rows <= 4096 bytes several small CTAs fit an SM (rows in flight side by side: one CTA's barriers and
    // table sweeps overlap the others' counting): four of 256 threads, three with the Entropy histograms resident as well
    const uint32_t half_occ = (sc & SC_ENTROPY) ? HALF_OCC - 1 : HALF_OCC;
    const uint32_t per_cta = (233472u / half_occ - 1024u) & ~127u;
    if ((sc == SC_BIGRAM || sc == (SC_ENTROPY | SC_BIGRAM)) && small_tab && g.max_len <= 4096 &&
        smem_bytes_for(sc | SC_HALF, rb, h0, 3) <= per_cta)
        sc |= SC_HALF;
    // Bigrams/BigEnt on short rows: the images are classed by the spread of their residuals on the device (no host round
    // trip), and two launches share the batch -- 32 x 32 tables for class 0, 64 x 64 tables for class 1
    if (sc == (SC_BIGRAM | SC_HALF) && ctx.scratch && ctx.scratch_bytes >= n_images &&
        smem_bytes_for(sc | SC_WIDE, rb, h0, 3) <= ((233472u / 2 - 1024u) & ~127u)) {
        uint64_t pg = n_images < (uint64_t)ctx.num_sms * 8 ? n_images : (uint64_t)ctx.num_sms * 8;
        k_spread_probe<<<(unsigned)pg, 256, 0, ctx.stream>>>(g, d_data, n_images, ctx.scratch);
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        e = launch_fast_variant(ctx, g, d_data, n_images, set, sc, small_tab, false, h0, ctx.scratch, 0);
        if (e != cudaSuccess) return e;
        return launch_fast_variant(ctx, g, d_data, n_images, set, sc | SC_WIDE, small_tab, false, h0, ctx.scratch, 1);
    }
    return launch_fast_variant(ctx, g, d_data, n_images, set, sc, small_tab, split, h0, nullptr, 0);
}

} // namespace oxi
