// Micro-benchmark: shared-memory update primitives that could carry the Entropy / Bigrams / BigEnt scorers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_ops smem_ops.cu ; run on a B200.
// Reports lane-ops per clock per SM for each primitive (cycles from clock64 inside the kernel).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// idx generator: mode 0 = uniform over `range`, 1 = peaked (16 hot values), 2 = all lanes same
__device__ __forceinline__ uint32_t gen(uint32_t &s, int mode, uint32_t range) {
    uint32_t r = lcg(s);
    if (mode == 0) return r % range;
    if (mode == 1) return ((r & 15u) * 2654435761u >> 8) % range;
    return 12345u % range;
}

extern __shared__ uint32_t smem[];

template <int OP>
__global__ void bench(int iters, int mode, uint32_t *sink, long long *cycles) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t s = 1234567u * (blockIdx.x * blockDim.x + tid + 1);
    if (mode == 2) s = 99;
    // zero 128 KB
    for (int i = tid; i < 32768; i += blockDim.x) smem[i] = 0;
    __syncthreads();
    uint32_t acc = 0;
    long long t0 = clock64();
    if (OP == 0) {  // ATOMS.ADD u32, no return, 32K words
        for (int i = 0; i < iters; i++) atomicAdd(&smem[gen(s, mode, 32768)], 1u);
    } else if (OP == 1) {  // ATOMS.ADD with return
        for (int i = 0; i < iters; i++) acc += atomicAdd(&smem[gen(s, mode, 32768)], 1u);
    } else if (OP == 2) {  // u16-packed counters: atomicAdd word with 1 or 65536 (65536 counters)
        for (int i = 0; i < iters; i++) { uint32_t k = gen(s, mode, 65536); atomicAdd(&smem[k >> 1], (k & 1) ? 65536u : 1u); }
    } else if (OP == 3) {  // lane-private u8 RMW, conflict-free layout, 8 KB per warp
        uint8_t *h = reinterpret_cast<uint8_t *>(smem) + warp * 8192;
        for (int i = 0; i < iters; i++) { uint32_t b = gen(s, mode, 256); uint32_t a = (((b >> 2) * 32 + lane) << 2) + (b & 3); h[a] = h[a] + 1; }
    } else if (OP == 4) {  // STS.U8 scatter into 64 KB + later LDS.U8 verify (write-verify distinct count)
        uint8_t *t = reinterpret_cast<uint8_t *>(smem);
        uint32_t s2 = s;
        for (int i = 0; i < iters; i++) t[gen(s, mode, 65536)] = (uint8_t)tid;
        __syncthreads();
        for (int i = 0; i < iters; i++) acc += (t[gen(s2, mode, 65536)] == (uint8_t)tid);
    } else if (OP == 5) {  // STS.U16 scatter into 128 KB + verify
        uint16_t *t = reinterpret_cast<uint16_t *>(smem);
        uint32_t s2 = s;
        for (int i = 0; i < iters; i++) t[gen(s, mode, 65536)] = (uint16_t)tid;
        __syncthreads();
        for (int i = 0; i < iters; i++) acc += (t[gen(s2, mode, 65536)] == (uint16_t)tid);
    } else if (OP == 6) {  // match.any on a 16-bit key
        for (int i = 0; i < iters; i++) acc += __popc(__match_any_sync(0xffffffffu, gen(s, mode, 65536)));
    } else if (OP == 7) {  // ATOMS.OR bitmap 8 KB (65536 bits) with return
        for (int i = 0; i < iters; i++) { uint32_t k = gen(s, mode, 65536); acc += (atomicOr(&smem[k >> 5], 1u << (k & 31)) >> (k & 31)) & 1; }
    } else if (OP == 8) {  // generator only (ALU baseline)
        for (int i = 0; i < iters; i++) acc += gen(s, mode, 65536);
    } else if (OP == 9) {  // warp-aggregated: match.any + leader ATOMS.ADD(popc)
        for (int i = 0; i < iters; i++) {
            uint32_t k = gen(s, mode, 32768);
            uint32_t m = __match_any_sync(0xffffffffu, k);
            if ((__ffs(m) - 1) == lane) atomicAdd(&smem[k], (uint32_t)__popc(m));
        }
    } else if (OP == 10) {  // plain LDS.32 random (bank-conflict baseline)
        for (int i = 0; i < iters; i++) acc += smem[gen(s, mode, 32768)];
    } else if (OP == 11) {  // lane-private u16 RMW (16 KB per warp)
        uint16_t *h = reinterpret_cast<uint16_t *>(smem) + warp * 8192;
        for (int i = 0; i < iters; i++) { uint32_t b = gen(s, mode, 256); uint32_t a = (((b >> 1) * 32 + lane) << 1) + (b & 1); h[a] = h[a] + 1; }
    }
    long long t1 = clock64();
    __syncthreads();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0xdeadbeef) sink[0] = acc + smem[tid];
}

template <int OP>
int run(const char *name, int threads, int mode, int smem_bytes, int opsPerIter) {
    int iters = 4096;
    int nsm = 148;
    long long *cyc;
    uint32_t *sink;
    CK(cudaMalloc(&cyc, nsm * sizeof(long long)));
    CK(cudaMalloc(&sink, 4096));
    CK(cudaFuncSetAttribute(bench<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    bench<OP><<<nsm, threads, smem_bytes>>>(iters, mode, sink, cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<nsm, threads, smem_bytes>>>(iters, mode, sink, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    CK(cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < nsm; i++) avg += h[i]; avg /= nsm;
    double ops = (double)iters * threads * opsPerIter;
    printf("%-34s thr=%4d mode=%d  %8.0f cyc  %6.2f lane-ops/clk/SM  (%.3f ms)\n", name, threads, mode, avg, ops / avg, ms);
    cudaFree(cyc); cudaFree(sink);
    return 0;
}

int main() {
    const int S = 131072 + 1024;
    for (int mode = 0; mode < 3; mode++) {
        for (int thr : {256, 1024}) {
            run<8>("alu baseline (gen only)", thr, mode, S, 1);
            run<10>("LDS.32 random", thr, mode, S, 1);
            run<0>("ATOMS.ADD u32 noret", thr, mode, S, 1);
            run<1>("ATOMS.ADD u32 ret", thr, mode, S, 1);
            run<2>("ATOMS.ADD u16-packed", thr, mode, S, 1);
            run<7>("ATOMS.OR bitmap ret", thr, mode, S, 1);
            run<9>("match.any + leader ATOMS", thr, mode, S, 1);
            run<6>("match.any only", thr, mode, S, 1);
            run<4>("STS.U8 scatter + LDS.U8 verify", thr, mode, S, 2);
            run<5>("STS.U16 scatter + LDS.U16 verify", thr, mode, S, 2);
        }
        run<3>("lane-private u8 RMW (8KB/warp)", 512, mode, S, 1);
        run<3>("lane-private u8 RMW (8KB/warp)", 256, mode, S, 1);
        run<11>("lane-private u16 RMW (16KB/warp)", 256, mode, S, 1);
    }
    return 0;
}
