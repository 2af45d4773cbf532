// Shared helpers for the pymn_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pymn_b200.h"

namespace mn {

// Fixed-point scale of the rank-standardised network weight: W_fix = W * 2^MN_W_SHIFT, W in [0, 1].
// 20 bits: the resolution (9.5e-7 per pair; a vote is a sum over thousands of pairs, so the relative
// rounding noise of a vote is ~1e-8, below float32) leaves room to pack a LUT entry -- W at the lower
// bin edge (21 bits) and the rise across the bin (11 bits) -- into ONE 32-bit word, which halves the
// shared-memory lookups of pass 2 (its bottleneck: random-bank LDS).
constexpr int MN_W_SHIFT = 20;
constexpr int MN_LUT_DELTA_BITS = 11;
constexpr uint32_t MN_LUT_DELTA_MAX = (1u << MN_LUT_DELTA_BITS) - 1u;
constexpr unsigned long long MN_W_ONE = 1ull << MN_W_SHIFT;
// fixed-point scale of a centroid term (centroids.cu, centroid_fixed_kernel): 2^38
constexpr double MN_CQ_ONE = 274877906944.0;

// optional CSR view of the expression matrix (scipy.sparse.csr_matrix layout); indptr == nullptr = dense input
struct CsrIn {
    const int64_t* indptr;    // [n_rows + 1]
    const int32_t* indices;   // [nnz] gene index of each stored value
    const float* data;        // [nnz]
};

// optional int8 limb output of K1 (the centroid paths' operand format, see limbs.cu): d = 128 a1 + a0 with
// a0 in [-64, 63] (G <= 127: a1 = d, a0 = 0), row pitch ldk (multiple of 128, padding zero), and the cell's scale
// 1 / ||d|| in double (0 = dropped cell).  a1 == nullptr: not wanted.
struct LimbOut {
    int8_t* a1;
    int8_t* a0;
    int64_t ldk;
    double* inv64;
};

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// rank_warp.cu
int launch_rank_warp(const float* X, int64_t ldx, const int32_t* row_index, const int32_t* col_index, int64_t M, int G,
                     __half* d_hi, __half* d_lo, int64_t ldd, float* inv_norm, uint8_t* flags, float* xnorm,
                     int32_t* rank2, int drop_mode, cudaStream_t st, bool* may_overflow, int only_mask, CsrIn csr,
                     LimbOut limb);
// auroc_count.cu
int64_t auroc_count_workspace_bytes(int ncols, int64_t M, int n_seg, int n_labels);
int launch_auroc_count(const float* votes, int64_t ldv, int ncols, const float* denom, const int32_t* col_denom,
                       const int32_t* seg_off, int n_seg, const int32_t* cell_label, int n_labels, int64_t M,
                       double* auroc, int32_t* n_pos, void* workspace, cudaStream_t st);
// rank_count.cu
int launch_rank_count(const float* X, int64_t ldx, const int32_t* row_index, const int32_t* col_index, int64_t M, int G,
                      __half* d_hi, __half* d_lo, int64_t ldd, float* inv_norm, uint8_t* flags, float* xnorm,
                      int32_t* rank2, int drop_mode, cudaStream_t st, CsrIn csr, LimbOut limb);

#define MN_CHECK_ARG(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            mn::set_error(__VA_ARGS__);      \
            return 1;                        \
        }                                    \
    } while (0)

#define MN_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            mn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)

#define MN_LAUNCH_CHECK()                                                                 \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            mn::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return 3;                                                                     \
        }                                                                                 \
        mn::count_launch();                                                               \
    } while (0)

// order-preserving map float -> uint32 (ascending floats -> ascending keys)
__device__ __forceinline__ uint32_t float_key(float v) {
    uint32_t u = __float_as_uint(v);
    if (u == 0x80000000u) u = 0u;      // -0.0 ties with +0.0
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// four centred ranks -> one word of each limb plane (byte u = gene g0 + u)
__device__ __forceinline__ void limb_pack4(const int d[4], bool one_limb, uint32_t& w1, uint32_t& w0) {
    w1 = 0u;
    w0 = 0u;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int hi = one_limb ? d[u] : (d[u] + 64) >> 7;      // floor((d + 64) / 128)
        const int lo = one_limb ? 0 : d[u] - (hi << 7);         // [-64, 63]
        w1 |= ((uint32_t)hi & 0xffu) << (8 * u);
        w0 |= ((uint32_t)lo & 0xffu) << (8 * u);
    }
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; scratch must hold >= 32 elements of T; all threads get the result.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    T t = (threadIdx.x < nw) ? scratch[threadIdx.x] : T(0);
    if (warp == 0) {
        t = warp_sum(t);
        if (lane == 0) scratch[0] = t;
    }
    __syncthreads();
    t = scratch[0];
    return t;
}

static inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace mn
