// host.h -- host-side services shared by the C-ABI translation units (defined in capi.cu).
#pragma once
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h> // header-only NVTX v3: ranges show up in Nsight Systems / ncu timelines, no-ops otherwise

#include "common.cuh"

namespace oxi {

// NVTX range around one C-ABI call (tracing)
struct Trace {
    explicit Trace(const char *name) { nvtxRangePushA(name); }
    ~Trace() { nvtxRangePop(); }
};

int fail(int code, const char *msg);
int fail_cuda(cudaError_t e, const char *where);

size_t channels(uint8_t color_type);
bool valid_ihdr(const oxi_ihdr *h);
bool build_geom(const oxi_ihdr *h, size_t data_len, int optimize_alpha, Geom *g);

// A stream plus grow-only device buffers; one slot is owned by one host call at a time.  A call that ran on a caller's
// stream without synchronising records `ev` when it gives the slot back; the next owner makes its stream wait for it,
// so the buffers are reused in stream order (no per-stream slot table, nothing shared between concurrent calls).
struct Slot {
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev = nullptr;
    bool ev_pending = false;
    uint8_t *scratch = nullptr;
    size_t scratch_bytes = 0;
    uint8_t *d_in = nullptr;
    size_t in_cap = 0;
    uint8_t *d_out = nullptr;
    size_t out_cap = 0;
    uint8_t *d_pre = nullptr;
    size_t pre_cap = 0;
    // pinned staging for callers with ordinary (pageable) host memory -- what `&[u8]` / `Vec<u8>` give the Rust side
    uint8_t *h_in = nullptr, *h_out = nullptr;
    size_t h_in_cap = 0, h_out_cap = 0;
    struct Pending { uint8_t *dst; const uint8_t *src; size_t n; };
    std::vector<Pending> pending; // staged results to copy out to the caller once the stream has finished
};

struct DeviceCtx {
    int device = -1;
    int num_sms = 0;
    std::mutex mu;
    std::vector<Slot *> free_slots;
    uint32_t smem_window = 0; // shared-window address of dynamic shared memory (probed once; the fast tier needs 0x400)
};

// Device-resident mode of the calling thread: while set, the reduction / unfilter / interlace entry points treat their
// `data` and `out` arguments as device pointers on `device` and enqueue on `stream` (no upload, no download; the
// few-byte scan results that decide OXI_NONE still come back synchronously).  Set by the oxi_image_* calls.
struct DevIO {
    int device;
    cudaStream_t stream;
};
extern thread_local const DevIO *t_devio;
struct DevIOScope {
    const DevIO *prev;
    explicit DevIOScope(const DevIO *d) : prev(t_devio) { t_devio = d; }
    ~DevIOScope() { t_devio = prev; }
};

DeviceCtx *get_ctx(int device, int *err);
int filter_resident(int device, cudaStream_t stream, const oxi_ihdr *h, const uint8_t *d_data, size_t data_len, size_t n_images,
                    const oxi_strategy *strategies, size_t k, int optimize_alpha, uint8_t *const *d_outs,
                    uint8_t *const *d_filters_used, unsigned flags);
DeviceCtx *current_ctx(int *err); // context of the calling thread's device (oxi_set_device), cudaSetDevice done
Slot *acquire_slot(DeviceCtx *c);
void release_slot(DeviceCtx *c, Slot *s);
// the slot's buffers may still be in use by work the previous owner left on another stream: make `st` wait for it
cudaError_t slot_begin(Slot *s, cudaStream_t st);
// give the slot back while work using its buffers is still queued on `st` (device-resident calls)
void release_slot_async(DeviceCtx *c, Slot *s, cudaStream_t st);
cudaError_t grow(uint8_t **p, size_t *cap, size_t need);
// ordinary (pageable) caller memory is staged through the slot's pinned buffers by several copy threads (capi.cu)
bool is_pageable(const void *p);
void par_memcpy(uint8_t *dst, const uint8_t *src, size_t n);
cudaError_t grow_pinned(uint8_t **p, size_t *cap, size_t need);

} // namespace oxi
