// capi.cu -- the C ABI (include/oxipng_b200.h): host-side geometry, per-device contexts, tier dispatch and
// host<->device staging for the filter-search entry points.  No CPU compute path exists here: when CUDA is
// unavailable every compute entry point returns OXI_E_CUDA.
#include <atomic>
#include <condition_variable>
#include <deque>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "host.h"

namespace oxi {

static std::atomic<uint64_t> g_launches{0};
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

thread_local std::string t_err;
thread_local int t_device = 0;
thread_local const DevIO *t_devio = nullptr;

int fail(int code, const char *msg) {
    t_err = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char *where) {
    t_err = std::string(where) + ": " + cudaGetErrorString(e);
    return OXI_E_CUDA;
}

// ---- geometry -------------------------------------------------------------------------------------------

size_t channels(uint8_t ct) {
    switch (ct) { // src/colors.rs:59-66
    case 0: case 3: return 1;
    case 4: return 2;
    case 2: return 3;
    case 6: return 4;
    default: return 0;
    }
}

bool valid_ihdr(const oxi_ihdr *h) {
    if (!h || h->width == 0 || h->height == 0) return false;
    if (channels(h->color_type) == 0) return false;
    switch (h->bit_depth) {
    case 1: case 2: case 4: return h->color_type == 0 || h->color_type == 3;
    case 8: return true;
    case 16: return h->color_type != 3;
    default: return false;
    }
}

// Adam7 pass origin/stride (src/interlace.rs:186-232)
static const uint32_t kX0[7] = {0, 4, 0, 2, 0, 1, 0}, kY0[7] = {0, 0, 4, 0, 2, 0, 1};
static const uint32_t kDX[7] = {8, 8, 4, 4, 2, 2, 1}, kDY[7] = {8, 8, 8, 4, 4, 2, 2};

// The rows ScanLines (src/png/scan_lines.rs:82-167) yields for `data_len` bytes, grouped by pass.  The iterator's
// per-pass pixel counts and its skipping of empty passes equal ceil((W-x0)/dx) x ceil((H-y0)/dy) with empty passes
// dropped (checked exhaustively against the oracle's state machine in tests/test_geometry.py); it stops as soon as
// the remaining bytes cannot hold another row, and a non-interlaced image keeps yielding rows while bytes remain.
bool build_geom(const oxi_ihdr *h, size_t data_len, int optimize_alpha, Geom *g) {
    memset(g, 0, sizeof *g);
    if (!valid_ihdr(h)) return false;
    const size_t bits = (size_t)h->bit_depth * channels(h->color_type);
    const size_t bpc = h->bit_depth == 16 ? 2 : 1;
    g->bpp = (uint32_t)(bpc * channels(h->color_type));
    const bool alpha = h->color_type == 4 || h->color_type == 6;
    g->alpha_bytes = (optimize_alpha && alpha) ? (uint32_t)bpc : 0; // src/png/mod.rs:363-367
    g->color_bytes = g->bpp - g->alpha_bytes;
    size_t left = data_len, off = 0;
    uint32_t row = 0;
    if (!h->interlaced) {
        const size_t len = ((size_t)h->width * bits + 7) / 8;
        if (len > 0xFFFFFFFFu) return false;
        const size_t n = left / len;
        if (n > 0xFFFFFFFFu) return false;
        if (n) {
            g->pass[0] = PassDesc{0, (uint32_t)n, (uint32_t)len, h->width, 0};
            g->n_pass = 1;
            g->max_len = (uint32_t)len;
            row = (uint32_t)n;
            off = n * len;
        }
    } else {
        for (int p = 0; p < 7; p++) {
            if (h->width <= kX0[p] || h->height <= kY0[p]) continue;
            const size_t pw = ((size_t)h->width - kX0[p] + kDX[p] - 1) / kDX[p];
            const size_t ph = ((size_t)h->height - kY0[p] + kDY[p] - 1) / kDY[p];
            const size_t len = (pw * bits + 7) / 8;
            if (len > 0xFFFFFFFFu) return false;
            size_t n = left / len;
            const bool exhausted = n < ph;
            if (n > ph) n = ph;
            if (n) {
                g->pass[g->n_pass++] = PassDesc{row, (uint32_t)n, (uint32_t)len, (uint32_t)pw, (uint64_t)off};
                if (len > g->max_len) g->max_len = (uint32_t)len;
                row += (uint32_t)n;
                off += n * len;
                left -= n * len;
            }
            if (exhausted) break;
        }
    }
    g->total_rows = row;
    g->data_len = off;
    g->out_len = off + row;
    return true;
}

// ---- per-device context -----------------------------------------------------------------------------------

static std::mutex g_ctx_mu;
static std::map<int, DeviceCtx *> g_ctx;

DeviceCtx *get_ctx(int device, int *err) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    auto it = g_ctx.find(device);
    if (it != g_ctx.end()) return it->second;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *err = fail_cuda(e, "cudaGetDeviceCount"); return nullptr; }
    if (device < 0 || device >= n) { *err = fail(OXI_E_CUDA, "no such CUDA device"); return nullptr; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { *err = fail_cuda(e, "cudaGetDeviceProperties"); return nullptr; }
    // the library holds sm_100a code only (arch-specific targets embed no PTX and are not forward compatible)
    if (prop.major != 10 || prop.minor != 0) {
        *err = fail(OXI_E_UNSUPPORTED, "device is not sm_100 (B200): this build has no kernel image for it");
        return nullptr;
    }
    DeviceCtx *c = new DeviceCtx;
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    {   // keep the stream-ordered pool's memory across synchronisations (oxi_image_* allocate from it per reduction)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    e = fast_tier_init(device, &c->smem_window);
    if (e != cudaSuccess) { delete c; *err = fail_cuda(e, "fast_tier_init"); return nullptr; }
    g_ctx[device] = c;
    return c;
}

cudaError_t grow(uint8_t **p, size_t *cap, size_t need) {
    if (need <= *cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    size_t want = need + need / 4 + 4096;
    cudaError_t e = cudaMalloc((void **)p, want);
    if (e == cudaSuccess) *cap = want;
    return e;
}

DeviceCtx *current_ctx(int *err) {
    DeviceCtx *c = get_ctx(t_device, err);
    if (!c) return nullptr;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) { *err = fail_cuda(e, "cudaSetDevice"); return nullptr; }
    return c;
}

Slot *acquire_slot(DeviceCtx *c) {
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (!c->free_slots.empty()) {
            Slot *s = c->free_slots.back();
            c->free_slots.pop_back();
            return s;
        }
    }
    Slot *s = new Slot;
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) { delete s; return nullptr; }
    if (cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming) != cudaSuccess) {
        cudaStreamDestroy(s->stream);
        delete s;
        return nullptr;
    }
    s->own_stream = true;
    return s;
}
// Slots are pooled, so their number is bounded by the peak number of concurrent calls on the device; beyond
// MAX_POOLED idle slots a returned slot is destroyed instead (its buffers are freed).
constexpr size_t MAX_POOLED = 64;
static void destroy_slot(Slot *s) {
    cudaStreamSynchronize(s->stream);
    if (s->ev_pending) cudaEventSynchronize(s->ev);
    cudaFree(s->scratch); cudaFree(s->d_in); cudaFree(s->d_out); cudaFree(s->d_pre);
    cudaFreeHost(s->h_in); cudaFreeHost(s->h_out);
    cudaEventDestroy(s->ev);
    cudaStreamDestroy(s->stream);
    delete s;
}
void release_slot(DeviceCtx *c, Slot *s) {
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (c->free_slots.size() < MAX_POOLED) { c->free_slots.push_back(s); return; }
    }
    destroy_slot(s);
}
cudaError_t slot_begin(Slot *s, cudaStream_t st) {
    if (!s->ev_pending) return cudaSuccess;
    s->ev_pending = false;
    return cudaStreamWaitEvent(st, s->ev, 0);
}
void release_slot_async(DeviceCtx *c, Slot *s, cudaStream_t st) {
    if (cudaEventRecord(s->ev, st) == cudaSuccess) s->ev_pending = true;
    else cudaStreamSynchronize(st);
    release_slot(c, s);
}

// ---- pageable host memory ---------------------------------------------------------------------------------------
// cudaMemcpyAsync on pageable memory is staged by the driver through small pinned buffers on the calling thread (measured:
// 3.7 GB/s raw pixels end to end against 16 GB/s with pinned buffers).  The host-buffer entry points therefore stage such
// buffers themselves: several threads copy a chunk into the slot's pinned buffer while the previous chunk is on the GPU,
// and copy finished results out of it, so the DMA engines always see pinned memory.
bool is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}
// A small persistent pool of copy threads (created on first use, shared by all calling threads): spawning threads per
// copy cost ~13 % of the pageable end-to-end time (128 copies of 32-96 MB per c5 step).
namespace {
struct CopyPool {
    struct Job { uint8_t *dst; const uint8_t *src; size_t n; std::atomic<int> *left; };
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job> q;
    std::vector<std::thread> workers;
    bool stop = false;
    explicit CopyPool(size_t n) {
        for (size_t i = 0; i < n; i++)
            workers.emplace_back([this] {
                for (;;) {
                    Job j;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [this] { return stop || !q.empty(); });
                        if (q.empty()) return; // stop
                        j = q.front();
                        q.pop_front();
                    }
                    memcpy(j.dst, j.src, j.n);
                    j.left->fetch_sub(1, std::memory_order_release);
                }
            });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto &w : workers) w.join();
    }
};
size_t copy_threads() {
    const unsigned hw = std::thread::hardware_concurrency();
    return hw >= 32 ? 12 : hw >= 16 ? 8 : hw >= 8 ? 4 : hw >= 4 ? 2 : 1;
}
CopyPool &copy_pool() {
    static CopyPool pool(copy_threads() - 1); // the calling thread copies a piece itself
    return pool;
}
} // namespace
void par_memcpy(uint8_t *dst, const uint8_t *src, size_t n) {
    const size_t piece = (size_t)4 << 20;
    size_t t = copy_threads();
    if (n < 2 * piece || t == 1) { memcpy(dst, src, n); return; }
    if (t > n / piece) t = n / piece;
    const size_t per = ((n + t - 1) / t + 4095) & ~(size_t)4095;
    CopyPool &pool = copy_pool();
    std::atomic<int> left{0};
    {
        std::lock_guard<std::mutex> lk(pool.mu);
        for (size_t i = 1; i < t; i++) {
            const size_t off = i * per;
            if (off >= n) break;
            left.fetch_add(1, std::memory_order_relaxed);
            pool.q.push_back({dst + off, src + off, off + per <= n ? per : n - off, &left});
        }
    }
    pool.cv.notify_all();
    memcpy(dst, src, per <= n ? per : n);
    // wait for the other pieces, helping with whatever is still queued (also keeps a forked child, which inherits the
    // pool object but not its threads, from waiting forever)
    while (left.load(std::memory_order_acquire) != 0) {
        CopyPool::Job j{nullptr, nullptr, 0, nullptr};
        {
            std::lock_guard<std::mutex> lk(pool.mu);
            if (!pool.q.empty()) { j = pool.q.front(); pool.q.pop_front(); }
        }
        if (j.left) {
            memcpy(j.dst, j.src, j.n);
            j.left->fetch_sub(1, std::memory_order_release);
        } else {
            std::this_thread::yield();
        }
    }
}
cudaError_t grow_pinned(uint8_t **p, size_t *cap, size_t need) {
    if (need <= *cap) return cudaSuccess;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaHostAlloc((void **)p, need, cudaHostAllocDefault);
    if (e == cudaSuccess) *cap = need;
    return e;
}
// wait for the slot's stream, then hand the staged results to the caller
cudaError_t flush_slot(Slot *s) {
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (e == cudaSuccess)
        for (const Slot::Pending &q : s->pending) par_memcpy(q.dst, q.src, q.n);
    s->pending.clear();
    return e;
}

// ---- dispatch -----------------------------------------------------------------------------------------------

bool valid_strategy(const oxi_strategy *s) {
    if (!s) return false;
    if (s->kind > OXI_PREDEFINED) return false;
    if (s->kind == OXI_BASIC && s->basic_filter > 4) return false;
    if (s->kind == OXI_PREDEFINED) {
        if (s->n_predefined && !s->predefined) return false;
        for (size_t i = 0; i < s->n_predefined; i++)
            if (s->predefined[i] > 4) return false;
    }
    return true;
}

// Upload Predefined lists (stream ordered) and fill the kernel-side strategy set.
int make_set(Slot *slot, cudaStream_t stream, const oxi_strategy *st, size_t k, uint8_t *const *d_outs,
             uint8_t *const *d_used, StrategySet *set) {
    if (k == 0 || k > (size_t)MAX_STRATEGIES) return fail(OXI_E_ARG, "strategy count must be 1..12");
    size_t pre_total = 0;
    for (size_t j = 0; j < k; j++) {
        if (!valid_strategy(&st[j])) return fail(OXI_E_ARG, "invalid strategy");
        if (st[j].kind == OXI_PREDEFINED) pre_total += st[j].n_predefined;
    }
    if (pre_total) {
        cudaError_t e = grow(&slot->d_pre, &slot->pre_cap, pre_total);
        if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(predefined)");
    }
    set->k = (int)k;
    size_t off = 0;
    for (size_t j = 0; j < k; j++) {
        DevStrategy &d = set->s[j];
        d.kind = st[j].kind;
        d.basic = st[j].basic_filter;
        d.predefined = nullptr;
        d.n_predefined = 0;
        if (st[j].kind == OXI_PREDEFINED && st[j].n_predefined) {
            size_t n = st[j].n_predefined > 0xFFFFFFFFu ? 0xFFFFFFFFu : st[j].n_predefined;
            cudaError_t e = cudaMemcpyAsync(slot->d_pre + off, st[j].predefined, n, cudaMemcpyHostToDevice, stream);
            if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpyAsync(predefined)");
            d.predefined = slot->d_pre + off;
            d.n_predefined = (uint32_t)n;
            off += n;
        }
        d.out = d_outs[j];
        d.filters_used = d_used[j];
    }
    return OXI_OK;
}

bool want_fast(const DeviceCtx *c, const Geom &g, const StrategySet &set, unsigned flags) {
    if (flags & OXI_FLAG_FORCE_GENERIC) return false;
    return fast_tier_usable(c->smem_window) && fast_supported(g, set);
}

int run_filters(DeviceCtx *c, Slot *slot, cudaStream_t stream, const Geom &g, const uint8_t *d_data, size_t n_images,
                const StrategySet &set, unsigned flags) {
    if (g.total_rows == 0 || n_images == 0) return OXI_OK;
    LaunchCtx lc;
    lc.device = c->device;
    lc.stream = stream;
    lc.smem_window = c->smem_window;
    lc.num_sms = c->num_sms;
    cudaError_t e;
    if (want_fast(c, g, set, flags)) {
        // one class byte per image for the bigram kernels' table-range choice (k_spread_probe)
        e = grow(&slot->scratch, &slot->scratch_bytes, n_images + 256);
        if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(scratch)");
        lc.scratch = slot->scratch;
        lc.scratch_bytes = slot->scratch_bytes;
        e = launch_fast(lc, g, d_data, n_images, set);
        if (e != cudaSuccess) return fail_cuda(e, "launch_fast");
        return OXI_OK;
    }
    if (g.alpha_bytes && !(flags & OXI_FLAG_FORCE_GENERIC) && alpha_supported(g, set)) {
        lc.scratch = nullptr;
        lc.scratch_bytes = 0;
        e = launch_alpha(lc, g, d_data, n_images, set);
        if (e != cudaSuccess) return fail_cuda(e, "launch_alpha");
        return OXI_OK;
    }
    if (flags & OXI_FLAG_FORCE_FAST) return fail(OXI_E_UNSUPPORTED, "fast tier does not cover this call");
    size_t need = generic_scratch_bytes(g, n_images, c->num_sms, nullptr);
    e = grow(&slot->scratch, &slot->scratch_bytes, need);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(scratch)");
    lc.scratch = slot->scratch;
    lc.scratch_bytes = slot->scratch_bytes;
    e = launch_generic(lc, g, d_data, n_images, set);
    if (e != cudaSuccess) return fail_cuda(e, "launch_generic");
    return OXI_OK;
}

// Host-buffer path for a contiguous batch chunk: upload, filter with all k strategies, download.
int filter_host_chunk(DeviceCtx *c, Slot *slot, const Geom &g, const uint8_t *data, size_t stride_in, size_t n_images,
                      const oxi_strategy *st, size_t k, uint8_t *const *outs, uint8_t *const *used, unsigned flags,
                      bool sync, bool stage_in, bool stage_out) {
    cudaError_t e = grow(&slot->d_in, &slot->in_cap, n_images * (size_t)g.data_len + 64);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(input)");
    const size_t per_out = n_images * (size_t)g.out_len, per_used = n_images * (size_t)g.total_rows;
    const size_t out_stride = (per_out + 255) & ~(size_t)255, used_stride = (per_used + 255) & ~(size_t)255;
    e = grow(&slot->d_out, &slot->out_cap, k * (out_stride + used_stride) + 64);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(output)");
    uint8_t *d_outs[MAX_STRATEGIES], *d_used[MAX_STRATEGIES];
    for (size_t j = 0; j < k && j < (size_t)MAX_STRATEGIES; j++) {
        d_outs[j] = slot->d_out + j * (out_stride + used_stride);
        d_used[j] = d_outs[j] + out_stride;
    }
    StrategySet set;
    int rc = make_set(slot, slot->stream, st, k, d_outs, d_used, &set);
    if (rc) return rc;
    const uint8_t *src = data;
    size_t src_stride = stride_in;
    if (stage_in) { // pageable input: pack the chunk into the slot's pinned buffer with several threads
        e = grow_pinned(&slot->h_in, &slot->h_in_cap, n_images * (size_t)g.data_len + 64);
        if (e != cudaSuccess) return fail_cuda(e, "cudaHostAlloc(input staging)");
        if (stride_in == g.data_len) par_memcpy(slot->h_in, data, n_images * (size_t)g.data_len);
        else
            for (size_t i = 0; i < n_images; i++) memcpy(slot->h_in + i * (size_t)g.data_len, data + i * stride_in, (size_t)g.data_len);
        src = slot->h_in;
        src_stride = g.data_len;
    }
    if (src_stride == g.data_len) {
        e = cudaMemcpyAsync(slot->d_in, src, n_images * (size_t)g.data_len, cudaMemcpyHostToDevice, slot->stream);
    } else {
        e = cudaMemcpy2DAsync(slot->d_in, g.data_len, src, src_stride, g.data_len, n_images, cudaMemcpyHostToDevice,
                              slot->stream);
    }
    if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpyAsync(H2D)");
    rc = run_filters(c, slot, slot->stream, g, slot->d_in, n_images, set, flags);
    if (rc) return rc;
    if (stage_out) {
        e = grow_pinned(&slot->h_out, &slot->h_out_cap, k * (out_stride + used_stride) + 64);
        if (e != cudaSuccess) return fail_cuda(e, "cudaHostAlloc(output staging)");
    }
    for (size_t j = 0; j < k; j++) {
        uint8_t *ho = stage_out ? slot->h_out + j * (out_stride + used_stride) : nullptr;
        if (outs && outs[j]) {
            e = cudaMemcpyAsync(stage_out ? ho : outs[j], d_outs[j], per_out, cudaMemcpyDeviceToHost, slot->stream);
            if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpyAsync(D2H out)");
            if (stage_out) slot->pending.push_back({outs[j], ho, per_out});
        }
        if (used && used[j]) {
            e = cudaMemcpyAsync(stage_out ? ho + out_stride : used[j], d_used[j], per_used, cudaMemcpyDeviceToHost, slot->stream);
            if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpyAsync(D2H filters)");
            if (stage_out) slot->pending.push_back({used[j], ho + out_stride, per_used});
        }
    }
    if (sync) {
        e = flush_slot(slot);
        if (e != cudaSuccess) return fail_cuda(e, "cudaStreamSynchronize");
    }
    return OXI_OK;
}

// Filter search on device-resident images (oxi_filter_batch_device, oxi_image_filter*): enqueues on `stream`, returns
// without synchronising.
int filter_resident(int device, cudaStream_t stream, const oxi_ihdr *h, const uint8_t *d_data, size_t data_len, size_t n_images,
                    const oxi_strategy *strategies, size_t k, int optimize_alpha, uint8_t *const *d_outs,
                    uint8_t *const *d_filters_used, unsigned flags) {
    Trace trace("oxi_filter (resident)");
    Geom g;
    if (!build_geom(h, data_len, optimize_alpha, &g)) return fail(OXI_E_ARG, "invalid IHDR");
    if (g.data_len != data_len) return fail(OXI_E_ARG, "data_len is not a whole number of scanlines for this IHDR");
    if (!d_data || !d_outs || !d_filters_used || !strategies) return fail(OXI_E_ARG, "null pointer");
    int err = 0;
    DeviceCtx *c = get_ctx(device, &err);
    if (!c) return err;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail_cuda(e, "cudaSetDevice");
    // scratch (Predefined lists, generic-tier rows) comes from a pooled slot; the work runs on the caller's stream and
    // the slot goes back with an event, so nothing is shared between concurrent callers and nothing synchronises
    Slot *slot = acquire_slot(c);
    if (!slot) return fail(OXI_E_CUDA, "cannot create stream");
    StrategySet set;
    int rc = OXI_OK;
    e = slot_begin(slot, stream);
    if (e != cudaSuccess) rc = fail_cuda(e, "cudaStreamWaitEvent");
    if (rc == OXI_OK) rc = make_set(slot, stream, strategies, k, d_outs, d_filters_used, &set);
    if (rc == OXI_OK) rc = run_filters(c, slot, stream, g, d_data, n_images, set, flags);
    release_slot_async(c, slot, stream);
    return rc;
}

// ---- one large image split into row bands across devices (png/mod.rs:375-378,422) ---------------------------
// With optimize_alpha off a filtered row depends on its own bytes and on the previous RAW row of the same pass only, so a
// band of rows needs a one-row halo and nothing else.  A band is run as a small non-interlaced image whose first row is
// the halo (its output row is dropped); the bands' outputs are disjoint, contiguous ranges of the filtered stream.
struct RowBand {
    uint32_t pixels, len;   // of one scanline of the band's pass
    uint32_t first_row, n_rows;
    bool halo;
    uint64_t in_begin, in_end;   // bytes of PngImage.data to ship (halo row included)
    uint64_t out_begin;          // where the band's filtered rows start in the filtered stream
};

void split_bands(const Geom &g, size_t n_parts, std::vector<RowBand> *bands) {
    for (uint32_t p = 0; p < g.n_pass; p++) {
        const PassDesc &pd = g.pass[p];
        const double frac = g.data_len ? (double)pd.nrows * pd.len / (double)g.data_len : 1.0;
        size_t share = (size_t)(n_parts * frac + 0.5);
        if (share < 1) share = 1;
        if (share > pd.nrows) share = pd.nrows;
        for (size_t b = 0; b < share; b++) {
            const uint32_t k0 = (uint32_t)(b * pd.nrows / share), k1 = (uint32_t)((b + 1) * pd.nrows / share);
            if (k1 <= k0) continue;
            RowBand rb;
            rb.pixels = pd.pixels;
            rb.len = pd.len;
            rb.first_row = pd.row0 + k0;
            rb.n_rows = k1 - k0;
            rb.halo = k0 > 0;
            rb.in_begin = pd.in_base + (uint64_t)(k0 - (rb.halo ? 1 : 0)) * pd.len;
            rb.in_end = pd.in_base + (uint64_t)k1 * pd.len;
            rb.out_begin = pd.in_base + (uint64_t)k0 * pd.len + pd.row0 + k0;
            bands->push_back(rb);
        }
    }
}

// all bands of one device, in order, on one pooled slot (called on that device's worker thread)
int run_bands_on_device(int device, const oxi_ihdr *h, const uint8_t *data, const std::vector<RowBand> &bands, size_t b0,
                        size_t b1, const oxi_strategy *st, size_t k, uint8_t *const *outs, uint8_t *const *used, unsigned flags,
                        std::string *msg) {
    int err = 0, rc = OXI_OK;
    DeviceCtx *c = get_ctx(device, &err);
    Slot *slot = nullptr;
    cudaError_t e = cudaSuccess;
    if (!c) rc = err;
    if (rc == OXI_OK && (e = cudaSetDevice(device)) != cudaSuccess) rc = fail_cuda(e, "cudaSetDevice");
    if (rc == OXI_OK && !(slot = acquire_slot(c))) rc = fail(OXI_E_CUDA, "cannot create stream");
    if (rc == OXI_OK && (e = slot_begin(slot, slot->stream)) != cudaSuccess) rc = fail_cuda(e, "cudaStreamWaitEvent");
    std::vector<std::vector<uint8_t>> pre(k); // Predefined lists re-based to the band (row i of the band = global row first_row + i)
    for (size_t bi = b0; bi < b1 && rc == OXI_OK; bi++) {
        const RowBand &b = bands[bi];
        const uint32_t skip = b.halo ? 1 : 0, rows = b.n_rows + skip;
        oxi_ihdr sub = *h;
        sub.width = b.pixels;
        sub.height = rows;
        sub.interlaced = 0;
        Geom g;
        if (!build_geom(&sub, (size_t)(b.in_end - b.in_begin), 0, &g) || g.total_rows != rows || g.pass[0].len != b.len) {
            rc = fail(OXI_E_ARG, "band geometry");
            break;
        }
        std::vector<oxi_strategy> bst(st, st + k);
        for (size_t j = 0; j < k; j++) {
            if (st[j].kind != OXI_PREDEFINED) continue;
            pre[j].assign(rows, 0);
            for (uint32_t i = 0; i < b.n_rows; i++)
                if ((size_t)b.first_row + i < st[j].n_predefined) pre[j][skip + i] = st[j].predefined[b.first_row + i];
            bst[j].predefined = pre[j].data();
            bst[j].n_predefined = rows;
        }
        if (bi > b0) { // the slot's buffers (and `pre`) are still in use by the previous band
            if ((e = cudaStreamSynchronize(slot->stream)) != cudaSuccess) { rc = fail_cuda(e, "cudaStreamSynchronize"); break; }
        }
        if ((e = grow(&slot->d_in, &slot->in_cap, (size_t)g.data_len + 64)) != cudaSuccess) { rc = fail_cuda(e, "cudaMalloc(input)"); break; }
        const size_t out_stride = ((size_t)g.out_len + 255) & ~(size_t)255, used_stride = ((size_t)rows + 255) & ~(size_t)255;
        if ((e = grow(&slot->d_out, &slot->out_cap, k * (out_stride + used_stride) + 64)) != cudaSuccess) { rc = fail_cuda(e, "cudaMalloc(output)"); break; }
        uint8_t *d_outs[MAX_STRATEGIES], *d_used[MAX_STRATEGIES];
        for (size_t j = 0; j < k; j++) {
            d_outs[j] = slot->d_out + j * (out_stride + used_stride);
            d_used[j] = d_outs[j] + out_stride;
        }
        StrategySet set;
        if ((rc = make_set(slot, slot->stream, bst.data(), k, d_outs, d_used, &set)) != OXI_OK) break;
        if ((e = cudaMemcpyAsync(slot->d_in, data + b.in_begin, (size_t)g.data_len, cudaMemcpyHostToDevice, slot->stream)) != cudaSuccess) {
            rc = fail_cuda(e, "cudaMemcpyAsync(H2D band)");
            break;
        }
        if ((rc = run_filters(c, slot, slot->stream, g, slot->d_in, 1, set, flags)) != OXI_OK) break;
        const size_t band_out = (size_t)b.n_rows * (b.len + 1u);
        for (size_t j = 0; j < k && e == cudaSuccess; j++) {
            if (outs && outs[j])
                e = cudaMemcpyAsync(outs[j] + b.out_begin, d_outs[j] + (size_t)skip * (b.len + 1u), band_out, cudaMemcpyDeviceToHost, slot->stream);
            if (e == cudaSuccess && used && used[j])
                e = cudaMemcpyAsync(used[j] + b.first_row, d_used[j] + skip, b.n_rows, cudaMemcpyDeviceToHost, slot->stream);
        }
        if (e != cudaSuccess) rc = fail_cuda(e, "cudaMemcpyAsync(D2H band)");
    }
    if (slot) {
        e = cudaStreamSynchronize(slot->stream);
        if (e != cudaSuccess && rc == OXI_OK) rc = fail_cuda(e, "cudaStreamSynchronize");
        release_slot(c, slot);
    }
    if (rc != OXI_OK && msg) *msg = t_err;
    return rc;
}

} // namespace oxi

using namespace oxi;

extern "C" {

const char *oxi_last_error(void) { return t_err.c_str(); }
const char *oxi_version(void) { return "oxipng-b200 0.3.0 (sm_100a)"; }
int oxi_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int oxi_set_device(int device) {
    int err = 0;
    if (!get_ctx(device, &err)) return err;
    t_device = device;
    return OXI_OK;
}
uint64_t oxi_kernel_launches(void) { return g_launches.load(); }

size_t oxi_bits_per_pixel(const oxi_ihdr *h) { return h ? (size_t)h->bit_depth * channels(h->color_type) : 0; }
size_t oxi_filter_bpp(const oxi_ihdr *h) { return h ? (h->bit_depth == 16 ? 2 : 1) * channels(h->color_type) : 0; }

size_t oxi_raw_data_size(const oxi_ihdr *h) { // src/headers.rs:38-64
    if (!h) return 0;
    const size_t w = h->width, ht = h->height, bpp = oxi_bits_per_pixel(h);
    auto bitmap = [bpp](size_t pw, size_t ph) { return ((pw * bpp + 7) / 8) * ph; };
    if (!h->interlaced) return bitmap(w, ht) + ht;
    size_t size = bitmap((w + 7) >> 3, (ht + 7) >> 3) + ((ht + 7) >> 3);
    if (w > 4) size += bitmap((w + 3) >> 3, (ht + 7) >> 3) + ((ht + 7) >> 3);
    size += bitmap((w + 3) >> 2, (ht + 3) >> 3) + ((ht + 3) >> 3);
    if (w > 2) size += bitmap((w + 1) >> 2, (ht + 3) >> 2) + ((ht + 3) >> 2);
    size += bitmap((w + 1) >> 1, (ht + 1) >> 2) + ((ht + 1) >> 2);
    if (w > 1) size += bitmap(w >> 1, (ht + 1) >> 1) + ((ht + 1) >> 1);
    return size + bitmap(w, ht >> 1) + (ht >> 1);
}

size_t oxi_num_scanlines(const oxi_ihdr *h, size_t data_len) {
    Geom g;
    if (!build_geom(h, data_len, 0, &g)) return 0;
    return g.total_rows;
}

size_t oxi_scan_lines(const oxi_ihdr *h, size_t data_len, uint32_t *len, uint8_t *pass, uint32_t *pixels, size_t cap) {
    Geom g;
    if (!build_geom(h, data_len, 0, &g)) return 0;
    const size_t bits = oxi_bits_per_pixel(h);
    size_t n = 0;
    int pid = 0;
    for (uint32_t p = 0; p < g.n_pass; p++) {
        uint8_t pass_id = 0;
        uint32_t px = h->width;
        if (h->interlaced) {
            // recover the pass number: passes appear in order, empty ones skipped
            while (pid < 7 && (h->width <= kX0[pid] || h->height <= kY0[pid])) pid++;
            pass_id = (uint8_t)(pid + 1);
            px = (h->width - kX0[pid] + kDX[pid] - 1) / kDX[pid];
            pid++;
        }
        (void)bits;
        for (uint32_t k = 0; k < g.pass[p].nrows; k++, n++) {
            if (n < cap) {
                if (len) len[n] = g.pass[p].len;
                if (pass) pass[n] = pass_id;
                if (pixels) pixels[n] = px;
            }
        }
    }
    return n;
}

const char *oxi_filter_tier(const oxi_ihdr *h, size_t data_len, const oxi_strategy *st, int optimize_alpha) {
    Geom g;
    if (!build_geom(h, data_len, optimize_alpha, &g) || !valid_strategy(st)) return "invalid";
    StrategySet set;
    memset(&set, 0, sizeof set);
    set.k = 1;
    set.s[0].kind = st->kind;
    set.s[0].basic = st->basic_filter;
    if (fast_supported(g, set)) return "fast"; // (a device whose shared-memory window differs runs generic)
    return alpha_supported(g, set) ? "alpha" : "generic";
}

int oxi_filter_batch_device(int device, void *stream, const oxi_ihdr *h, const uint8_t *d_data, size_t data_len,
                            size_t n_images, const oxi_strategy *strategies, size_t k, int optimize_alpha,
                            uint8_t *const *d_outs, uint8_t *const *d_filters_used, unsigned flags) {
    return filter_resident(device, (cudaStream_t)stream, h, d_data, data_len, n_images, strategies, k, optimize_alpha, d_outs,
                           d_filters_used, flags);
}

int oxi_filter_batch(const oxi_ihdr *h, const uint8_t *data, size_t data_len, size_t n_images,
                     const oxi_strategy *strategies, size_t k, int optimize_alpha, uint8_t *const *outs,
                     uint8_t *const *filters_used, unsigned flags) {
    Trace trace("oxi_filter_batch (host buffers)");
    Geom g;
    if (!build_geom(h, data_len, optimize_alpha, &g)) return fail(OXI_E_ARG, "invalid IHDR");
    if (!data || !strategies || k == 0 || k > (size_t)MAX_STRATEGIES) return fail(OXI_E_ARG, "bad arguments");
    // a single image may carry trailing bytes that hold no further scanline (the reference's iterator ignores them);
    // in a batch the images sit data_len apart, so data_len must be exactly the scanlines of one image
    if (n_images > 1 && g.data_len != data_len)
        return fail(OXI_E_ARG, "data_len is not a whole number of scanlines for this IHDR");
    int err = 0;
    DeviceCtx *c = get_ctx(t_device, &err);
    if (!c) return err;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail_cuda(e, "cudaSetDevice");
    if (n_images == 0) return OXI_OK;
    // ordinary (pageable) caller memory is staged through pinned buffers by this library (see par_memcpy)
    const bool stage_in = is_pageable(data);
    bool stage_out = false;
    for (size_t j = 0; j < k && !stage_out; j++)
        stage_out = (outs && outs[j] && is_pageable(outs[j])) || (filters_used && filters_used[j] && is_pageable(filters_used[j]));
    // chunk the batch so uploads, kernels and downloads of neighbouring chunks overlap (3 slots in flight)
    const size_t target = (stage_in || stage_out) ? (size_t)32 << 20 : (size_t)96 << 20;
    size_t per = g.data_len ? target / (size_t)g.data_len : n_images;
    if (per < 1) per = 1;
    if (per > n_images) per = n_images;
    const size_t n_chunks = (n_images + per - 1) / per;
    const int NS = n_chunks >= 3 ? 3 : (int)n_chunks;
    Slot *slots[3] = {nullptr, nullptr, nullptr};
    int rc = OXI_OK;
    for (int i = 0; i < NS; i++) {
        slots[i] = acquire_slot(c);
        if (!slots[i]) { rc = fail(OXI_E_CUDA, "cannot create stream"); continue; }
        e = slot_begin(slots[i], slots[i]->stream);
        if (e != cudaSuccess) rc = fail_cuda(e, "cudaStreamWaitEvent");
    }
    for (size_t ci = 0; ci < n_chunks && rc == OXI_OK; ci++) {
        Slot *s = slots[ci % NS];
        if (ci >= (size_t)NS) { // buffers of this slot are still in use by chunk ci-NS
            e = flush_slot(s);
            if (e != cudaSuccess) { rc = fail_cuda(e, "cudaStreamSynchronize"); break; }
        }
        const size_t i0 = ci * per, n = (i0 + per <= n_images) ? per : n_images - i0;
        uint8_t *o[MAX_STRATEGIES], *u[MAX_STRATEGIES];
        for (size_t j = 0; j < k; j++) {
            o[j] = (outs && outs[j]) ? outs[j] + i0 * (size_t)g.out_len : nullptr;
            u[j] = (filters_used && filters_used[j]) ? filters_used[j] + i0 * (size_t)g.total_rows : nullptr;
        }
        rc = filter_host_chunk(c, s, g, data + i0 * data_len, data_len, n, strategies, k, o, u, flags, false, stage_in, stage_out);
    }
    for (int i = 0; i < NS; i++) {
        if (!slots[i]) continue;
        e = flush_slot(slots[i]);
        if (e != cudaSuccess && rc == OXI_OK) rc = fail_cuda(e, "cudaStreamSynchronize");
        release_slot(c, slots[i]);
    }
    return rc;
}

int oxi_filter_image_sharded(const oxi_ihdr *h, const uint8_t *data, size_t data_len, const oxi_strategy *strategies, size_t k,
                             int optimize_alpha, const int *devices, int n_devices, int bands_per_device, uint8_t *const *outs,
                             uint8_t *const *filters_used, unsigned flags) {
    Trace trace("oxi_filter_image_sharded");
    Geom g;
    if (!build_geom(h, data_len, optimize_alpha, &g)) return fail(OXI_E_ARG, "invalid IHDR");
    if (!data || !strategies || k == 0 || k > (size_t)MAX_STRATEGIES || !devices || n_devices < 1 || n_devices > 64)
        return fail(OXI_E_ARG, "bad arguments");
    for (size_t j = 0; j < k; j++)
        if (!valid_strategy(&strategies[j])) return fail(OXI_E_ARG, "invalid strategy");
    // with optimize_alpha the next row is filtered against the winning trial's rewritten row (png/mod.rs:412-423):
    // rows are serial and the image cannot be split
    if (g.alpha_bytes != 0) return fail(OXI_E_UNSUPPORTED, "optimize_alpha makes rows serially dependent: no row-band sharding");
    if (g.total_rows == 0) return OXI_OK;
    if (bands_per_device < 1) bands_per_device = 1;
    std::vector<RowBand> bands;
    split_bands(g, (size_t)n_devices * bands_per_device, &bands);
    // contiguous runs of bands per device, one worker thread each (pageable host memory makes cudaMemcpyAsync
    // synchronous for the issuing thread, so the devices are driven from separate threads)
    const size_t nb = bands.size();
    std::vector<int> rcs(n_devices, OXI_OK);
    std::vector<std::string> msgs(n_devices);
    std::vector<std::thread> workers;
    for (int d = 0; d < n_devices; d++) {
        const size_t b0 = nb * d / n_devices, b1 = nb * (d + 1) / n_devices;
        if (b0 == b1) continue;
        workers.emplace_back([&, d, b0, b1] {
            rcs[d] = run_bands_on_device(devices[d], h, data, bands, b0, b1, strategies, k, outs, filters_used, flags, &msgs[d]);
        });
    }
    for (auto &w : workers) w.join();
    for (int d = 0; d < n_devices; d++)
        if (rcs[d] != OXI_OK) return fail(rcs[d], msgs[d].c_str());
    return OXI_OK;
}

int oxi_filter_image_multi(const oxi_ihdr *h, const uint8_t *data, size_t data_len, const oxi_strategy *strategies,
                           size_t k, int optimize_alpha, uint8_t *const *outs, uint8_t *const *filters_used) {
    return oxi_filter_batch(h, data, data_len, 1, strategies, k, optimize_alpha, outs, filters_used, 0);
}

int oxi_filter_image(const oxi_ihdr *h, const uint8_t *data, size_t data_len, const oxi_strategy *strategy,
                     int optimize_alpha, uint8_t *out, uint8_t *filters_used, int *used_is_predefined) {
    if (!strategy) return fail(OXI_E_ARG, "null strategy");
    uint8_t *outs[1] = {out}, *used[1] = {filters_used};
    int rc = oxi_filter_batch(h, data, data_len, 1, strategy, 1, optimize_alpha, outs, used, 0);
    if (rc == OXI_OK && used_is_predefined) { // src/png/mod.rs:426-430
        const bool heuristic = strategy->kind != OXI_BASIC && strategy->kind != OXI_PREDEFINED;
        *used_is_predefined = heuristic && oxi_num_scanlines(h, data_len) > 0;
    }
    return rc;
}

} // extern "C"
