// Microbenchmark: shared-memory atomic / load / store throughput at random addresses (design input for K6).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(1024) k(uint32_t* out, int iters, int nbins) {
    extern __shared__ uint32_t s[];
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) s[i] = 0;
    __syncthreads();
    uint32_t x = hash32(blockIdx.x * 1024 + threadIdx.x + 1), acc = 0;
    const uint32_t mask = nbins - 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x = x * 1664525u + 1013904223u;
            const uint32_t b = (x >> 10) & mask;
            if (MODE == 0) atomicAdd(&s[b], 1u);                       // no return (RED)
            else if (MODE == 1) acc += atomicAdd(&s[b], 1u);           // with return
            else if (MODE == 2) acc += s[b];                           // random LDS
            else if (MODE == 3) s[b] = x;                              // random STS
            else if (MODE == 4) atomicAdd(&s[b >> 1], 1u << ((b & 1) * 16));   // packed 16-bit
            else if (MODE == 5) acc += __match_any_sync(0xffffffffu, b & 255u);
            else if (MODE == 6) acc += atomicExch(&s[b], x);
            else if (MODE == 7) { const uint32_t o = atomicAdd(&s[b], x | 1u); if (o > ~(x | 1u)) atomicAdd(&s[(b + 1) & mask], 1u); }
        }
    }
    __syncthreads();
    if (acc == 0x12345678u || MODE != 99) out[blockIdx.x * blockDim.x + threadIdx.x] = acc + s[threadIdx.x & mask];
}
template <int MODE>
void run(const char* name, int threads, int nbins) {
    uint32_t* out; cudaMalloc(&out, 148 * 1024 * 4);
    const int iters = 2000;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, nbins * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, threads, nbins * 4>>>(out, 10, nbins);
    cudaEventRecord(e0);
    k<MODE><<<148, threads, nbins * 4>>>(out, iters, nbins);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 148.0 * threads * iters * 8.0;
    printf("%-28s threads=%4d bins=%6d : %8.3f ms  %7.1f Gop/s  (%.3f op/clk/SM @1.9GHz)  err=%d\n", name, threads, nbins, ms,
           ops / ms * 1e-6, ops / ms * 1e-6 / 148 / 1.9, (int)cudaGetLastError());
    cudaFree(out);
}
int main() {
    for (int threads : {256, 1024}) {
        for (int nbins : {256, 4096, 32768}) {
            run<0>("atomicAdd no return", threads, nbins);
            run<1>("atomicAdd with return", threads, nbins);
            run<4>("packed 16-bit add no ret", threads, nbins);
            run<6>("atomicExch with return", threads, nbins);
            run<7>("acc lo/hi carry", threads, nbins);
            run<2>("random LDS", threads, nbins);
            run<3>("random STS", threads, nbins);
        }
        run<5>("match.any 8-bit", threads, 256);
    }
    return 0;
}
