// Exhaustive device check of the byte-SIMD Paeth predictor used by the fast tier against the scalar definition
// (src/filters/mod.rs:242-254): all 2^24 (a,b,c), each placed in every byte lane with different neighbours.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o paeth4_exhaustive paeth4_exhaustive.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../oxipng_b200/csrc/common.cuh" // the product's own paeth4
using oxi::paeth4;

__device__ uint32_t paeth1(uint32_t a, uint32_t b, uint32_t c) {
    int p = (int)a + (int)b - (int)c;
    int pa = abs(p - (int)a), pb = abs(p - (int)b), pc = abs(p - (int)c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
__global__ void check(unsigned long long *bad, uint32_t *first) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; // 2^24 threads
    uint32_t a = idx & 255, b = (idx >> 8) & 255, c = idx >> 16;
    // neighbours in the other lanes: rotate values so that lanes differ
    uint32_t A = a | ((a * 7 + 13) & 255) << 8 | ((255 - a) << 16) | ((a ^ 0x5A) << 24);
    uint32_t B = b | ((b * 3 + 200) & 255) << 8 | ((b ^ 0xFF) << 16) | (((b + 128) & 255) << 24);
    uint32_t Cc = c | ((c * 5 + 77) & 255) << 8 | (((c >> 1) | 0x80) << 16) | ((c ^ 0x33) << 24);
    uint32_t got = paeth4(A, B, Cc);
    for (int l = 0; l < 4; l++) {
        uint32_t want = paeth1((A >> (8 * l)) & 255, (B >> (8 * l)) & 255, (Cc >> (8 * l)) & 255);
        if (((got >> (8 * l)) & 255) != want) {
            if (atomicAdd(bad, 1ull) == 0) { first[0] = A; first[1] = B; first[2] = Cc; first[3] = got; first[4] = l; }
        }
    }
}
int main() {
    unsigned long long *bad; uint32_t *first;
    cudaMallocManaged(&bad, 8); cudaMallocManaged(&first, 32);
    *bad = 0;
    check<<<65536, 256>>>(bad, first);
    cudaError_t e = cudaDeviceSynchronize();
    printf("paeth4 exhaustive: %s, mismatches=%llu\n", cudaGetErrorString(e), *bad);
    if (*bad) printf("first: A=%08x B=%08x C=%08x got=%08x lane=%u\n", first[0], first[1], first[2], first[3], first[4]);
    return *bad != 0;
}
