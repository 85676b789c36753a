// host.h -- host-side services shared by the C-ABI translation units (defined in capi.cu).
#pragma once
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace oxi {

int fail(int code, const char *msg);
int fail_cuda(cudaError_t e, const char *where);

size_t channels(uint8_t color_type);
bool valid_ihdr(const oxi_ihdr *h);
bool build_geom(const oxi_ihdr *h, size_t data_len, int optimize_alpha, Geom *g);

// A stream plus grow-only device buffers; one slot is owned by one host call at a time.
struct Slot {
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint8_t *scratch = nullptr;
    size_t scratch_bytes = 0;
    uint8_t *d_in = nullptr;
    size_t in_cap = 0;
    uint8_t *d_out = nullptr;
    size_t out_cap = 0;
    uint8_t *d_pre = nullptr;
    size_t pre_cap = 0;
};

struct DeviceCtx {
    int device = -1;
    int num_sms = 0;
    std::mutex mu;
    std::vector<Slot *> free_slots;
    std::map<cudaStream_t, Slot *> stream_slots;
};

DeviceCtx *get_ctx(int device, int *err);
DeviceCtx *current_ctx(int *err); // context of the calling thread's device (oxi_set_device), cudaSetDevice done
Slot *acquire_slot(DeviceCtx *c);
void release_slot(DeviceCtx *c, Slot *s);
cudaError_t grow(uint8_t **p, size_t *cap, size_t need);

} // namespace oxi
