// K1: per-cell average-tie ranking + centring + L2 normalisation.
//
// Replaces normalize_cells (reference pymn/utils.py:122-148) and the ranking
// step of create_nw_spearman (utils.py:69).  One CTA per cell:
//   1. stream the cell's G expression values in (coalesced), compact the
//      non-zeros into shared memory as (order-preserving key, gene index);
//      the exact zeros -- 70-95 % of a scRNA-seq row -- form ONE tie block
//      whose rank is closed-form, so they are never sorted;
//   2. bitonic-sort the non-zeros in shared memory;
//   3. tie runs -> doubled average rank 2r = first + last + 2 (0-based sorted
//      positions), scattered to a gene-indexed staging row in shared memory;
//   4. centre: d = 2r - (G+1) (the mean rank is always (G+1)/2); the squared
//      norm sum d^2 is an exact integer; write d as fp16 plane(s) (exact:
//      |d| < G), 1/||d||, flags, and optionally the reference's float matrix
//      and the integer 2*rank for bit-exact parity checks.
// HBM traffic: 4 B read + 2 (or 4) B written per element.
#include "common.cuh"

namespace mn {

__device__ __forceinline__ void bitonic_sort_smem(uint64_t* a, int P) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool up = ((i & k) == 0);
                const uint64_t x = a[i], y = a[l];
                if ((x > y) == up) {
                    a[i] = y;
                    a[l] = x;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void rank_normalize_kernel(const float* __restrict__ X, int64_t ldx,
                                      const int32_t* __restrict__ row_index,
                                      const int32_t* __restrict__ col_index, int64_t M, int G, int Pmax,
                                      __half* __restrict__ d_hi, __half* __restrict__ d_lo, int64_t ldd,
                                      float* __restrict__ inv_norm, uint8_t* __restrict__ flags,
                                      float* __restrict__ xnorm, int32_t* __restrict__ rank2, int drop_mode,
                                      int only_overflow, CsrIn csr, LimbOut limb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* skey = reinterpret_cast<uint64_t*>(smem_raw);   // [Pmax]
    int* r2 = reinterpret_cast<int*>(skey + Pmax);            // [G]
    __shared__ int s_cnt, s_nneg;
    __shared__ double s_red_d[32];
    __shared__ unsigned long long s_red_u[32];

    const int tid = threadIdx.x, lane = tid & 31, T = blockDim.x;
    const unsigned lt = lanemask_lt();

    for (int64_t row = blockIdx.x; row < M; row += gridDim.x) {
        // second stage of the warp-per-cell kernel: only rows it flagged as too dense (bit 3)
        if (only_overflow && !(flags[row] & 8)) continue;
        const int64_t src = row_index ? (int64_t)row_index[row] : row;
        const float* x = X + src * ldx;
        if (tid == 0) {
            s_cnt = 0;
            s_nneg = 0;
        }
        __syncthreads();

        // ---- 1. load + compact non-zeros -----------------------------------
        double psum = 0.0;
        const int64_t e0 = csr.indptr ? csr.indptr[src] : 0;
        const int n_in = csr.indptr ? (int)(csr.indptr[src + 1] - e0) : G;      // stored values (CSR) or genes (dense)
        for (int base = 0; base < n_in; base += T) {
            const int q = base + tid;
            bool in = q < n_in;
            float v = 0.f;
            int g = q;
            if (in) {
                if (csr.indptr) {
                    v = __ldg(csr.data + e0 + q);
                    g = __ldg(csr.indices + e0 + q);
                    in = (unsigned)g < (unsigned)G;
                    if (!in) v = 0.f;
                } else {
                    v = __ldg(x + (col_index ? col_index[g] : g));
                }
            }
            psum += (double)v;
            const bool nz = in && (v != 0.f);
            const bool ng = in && (v < 0.f);
            const unsigned bal = __ballot_sync(0xffffffffu, nz);
            const unsigned baln = __ballot_sync(0xffffffffu, ng);
            int wbase = 0;
            if (lane == 0) {
                if (bal) wbase = atomicAdd(&s_cnt, __popc(bal));
                if (baln) atomicAdd(&s_nneg, __popc(baln));
            }
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (nz) skey[wbase + __popc(bal & lt)] = ((uint64_t)float_key(v) << 32) | (uint32_t)g;
        }
        __syncthreads();
        const int nnz = s_cnt, nneg = s_nneg, nzero = G - nnz;
        int P = 2;
        while (P < nnz) P <<= 1;
        for (int i = nnz + tid; i < P; i += T) skey[i] = ~0ull;
        const int zero_r2 = 2 * nneg + nzero + 1;
        for (int g = tid; g < G; g += T) r2[g] = zero_r2;
        __syncthreads();

        // ---- 2. sort ---------------------------------------------------------
        if (nnz > 1) bitonic_sort_smem(skey, P);

        // ---- 3. tie runs -> 2*rank ------------------------------------------
        for (int p = tid; p < nnz; p += T) {
            const uint64_t e = skey[p];
            const uint32_t key = (uint32_t)(e >> 32);
            const int g = (int)(uint32_t)e;
            int first = p, last = p;
            if (p > 0 && (uint32_t)(skey[p - 1] >> 32) == key) {
                int lo = 0, hi = p - 1;                 // first position whose key >= key
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if ((uint32_t)(skey[mid] >> 32) < key) lo = mid + 1; else hi = mid;
                }
                first = lo;
            }
            if (p + 1 < nnz && (uint32_t)(skey[p + 1] >> 32) == key) {
                int lo = p + 1, hi = nnz - 1;           // last position whose key <= key
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if ((uint32_t)(skey[mid] >> 32) > key) hi = mid - 1; else lo = mid;
                }
                last = lo;
            }
            const int off = (key & 0x80000000u) ? nzero : 0;   // positives sit above the zero block
            r2[g] = first + last + 2 + 2 * off;
        }
        __syncthreads();

        // ---- 4. centre, norm, write ------------------------------------------
        const int c = G + 1;
        unsigned long long ss = 0;
        for (int g = tid; g < G; g += T) {
            const long long d = (long long)r2[g] - c;
            ss += (unsigned long long)(d * d);
        }
        ss = block_sum<unsigned long long>(ss, s_red_u);
        const double tot = block_sum<double>(psum, s_red_d);
        const bool zerovar = (ss == 0ull);
        const bool nonpos = !(tot > 0.0);
        const bool drop = zerovar || (drop_mode == 1 && nonpos);
        const double inv_true = zerovar ? 0.0 : 1.0 / sqrt((double)ss);
        if (tid == 0) {
            inv_norm[row] = drop ? 0.f : (float)inv_true;
            flags[row] = (uint8_t)((zerovar ? 1 : 0) | (nonpos ? 2 : 0) | (drop ? 4 : 0));
        }
        if (limb.a1) {
            uint32_t* p1 = reinterpret_cast<uint32_t*>(limb.a1 + row * limb.ldk);
            uint32_t* p0 = reinterpret_cast<uint32_t*>(limb.a0 + row * limb.ldk);
            const bool one_limb = G <= 127;
            for (int g4 = tid; g4 < (int)(limb.ldk >> 2); g4 += T) {
                int d[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) d[u] = (!drop && 4 * g4 + u < G) ? r2[4 * g4 + u] - c : 0;
                uint32_t w1, w0;
                limb_pack4(d, one_limb, w1, w0);
                p1[g4] = w1;
                p0[g4] = w0;
            }
            if (tid == 0) limb.inv64[row] = drop ? 0.0 : inv_true;
        }
        __half2* ph = reinterpret_cast<__half2*>(d_hi + row * ldd);
        __half2* pl = d_lo ? reinterpret_cast<__half2*>(d_lo + row * ldd) : nullptr;
        for (int g2 = tid; d_hi && g2 < (int)(ldd >> 1); g2 += T) {
            const int g = 2 * g2;
            const int d0 = (!drop && g < G) ? r2[g] - c : 0;
            const int d1 = (!drop && g + 1 < G) ? r2[g + 1] - c : 0;
            const __half h0 = __int2half_rn(d0), h1 = __int2half_rn(d1);
            ph[g2] = __halves2half2(h0, h1);
            if (pl) pl[g2] = __halves2half2(__int2half_rn(d0 - __half2int_rn(h0)),
                                           __int2half_rn(d1 - __half2int_rn(h1)));
        }
        if (xnorm) {
            float* xo = xnorm + row * (int64_t)G;
            const float qnan = __int_as_float(0x7fc00000);
            for (int g = tid; g < G; g += T)
                xo[g] = zerovar ? qnan : (float)((double)(r2[g] - c) * inv_true);
        }
        if (rank2) {
            int32_t* ro = rank2 + row * (int64_t)G;
            for (int g = tid; g < G; g += T) ro[g] = r2[g];
        }
        __syncthreads();
    }
}

}  // namespace mn

static int rank_dispatch(const float* X, int64_t ldx, mn::CsrIn csr, const int32_t* row_index, const int32_t* col_index,
                         int64_t M, int32_t G, mn_half_t* d_hi, mn_half_t* d_lo, int64_t ldd, float* inv_norm,
                         uint8_t* flags, float* xnorm, int32_t* rank2, int32_t drop_mode, mn_stream_t stream,
                         mn::LimbOut limb = mn::LimbOut{nullptr, nullptr, 0, nullptr}) {
    MN_CHECK_ARG((X || csr.indptr) && (d_hi || limb.a1) && inv_norm && flags, "mn_rank_normalize: null pointer");
    MN_CHECK_ARG(G >= 1 && G <= 16384, "mn_rank_normalize: G=%d outside [1, 16384]", G);
    if (d_hi) {
        MN_CHECK_ARG(ldd >= G && (ldd % 8) == 0, "mn_rank_normalize: ldd=%lld must be >= G and a multiple of 8", (long long)ldd);
        MN_CHECK_ARG(G <= 2048 || d_lo, "mn_rank_normalize: G=%d > 2048 needs the d_lo plane", G);
    }
    if (limb.a1)
        MN_CHECK_ARG(limb.a0 && limb.inv64 && limb.ldk >= G && (limb.ldk % 128) == 0,
                     "mn_rank_normalize_limbs: a0 / inv_norm64 missing or ldk=%lld not a multiple of 128 >= G", (long long)limb.ldk);
    if (M <= 0) return 0;
    // fast path: one warp per cell, register bitonic sort (rank_warp.cu); rows with more than 2048
    // non-zeros (only possible for G > 2048) are flagged and redone below by the CTA-per-cell kernel
    // Three stages, each a persistent grid; a later stage only touches the rows the earlier one flagged:
    //   1. rank_count.cu   rows of small non-negative integers (counts): rank by counting, HBM-bound
    //   2. rank_warp.cu    any other row (flag bit 4): one warp per cell, register bitonic sort
    //   3. below           rows with more than 1024 non-zeros (flag bit 3): CTA per cell, shared-memory sort
    int only_overflow = 0;
    if (!getenv("MN_RANK_CTA")) {
        __half* hi = reinterpret_cast<__half*>(d_hi);
        __half* lo = reinterpret_cast<__half*>(d_lo);
        cudaStream_t st = (cudaStream_t)stream;
        int only_mask = 0;
        if (!getenv("MN_RANK_NOCOUNT")) {
            const int rc = mn::launch_rank_count(X, ldx, row_index, col_index, M, G, hi, lo, ldd, inv_norm, flags, xnorm,
                                                 rank2, drop_mode, st, csr, limb);
            if (rc) return rc;
            only_mask = 16;
        }
        bool may_overflow = false;
        const int rc = mn::launch_rank_warp(X, ldx, row_index, col_index, M, G, hi, lo, ldd, inv_norm, flags, xnorm, rank2,
                                            drop_mode, st, &may_overflow, only_mask, csr, limb);
        if (rc) return rc;
        if (!may_overflow) return 0;
        only_overflow = 1;
    }
    const int Pmax = mn::next_pow2(G < 2 ? 2 : G);
    const size_t smem = (size_t)Pmax * 8 + (size_t)G * 4;
    int T = 256;
    if (G <= 64) T = 32; else if (G <= 256) T = 64; else if (G <= 1024) T = 128;
    // (set on every launch: the attribute is per device and per context, and the call costs about a microsecond)
    if (smem > 40 * 1024)
        MN_CUDA(cudaFuncSetAttribute(mn::rank_normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = M < (int64_t)(1 << 30) ? M : (int64_t)(1 << 30);
    if (only_overflow && grid > 1184) grid = 1184;      // a small persistent grid scans the flags for the few dense rows
    mn::rank_normalize_kernel<<<(unsigned)grid, T, smem, (cudaStream_t)stream>>>(
        X, ldx, row_index, col_index, M, G, Pmax, reinterpret_cast<__half*>(d_hi), reinterpret_cast<__half*>(d_lo),
        ldd, inv_norm, flags, xnorm, rank2, drop_mode, only_overflow, csr, limb);
    MN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mn_rank_normalize(const float* X, int64_t ldx, const int32_t* row_index, const int32_t* col_index,
                                 int64_t M, int32_t G, mn_half_t* d_hi, mn_half_t* d_lo, int64_t ldd,
                                 float* inv_norm, uint8_t* flags, float* xnorm, int32_t* rank2,
                                 int32_t drop_mode, mn_stream_t stream) {
    MN_CHECK_ARG(X != nullptr, "mn_rank_normalize: null matrix");
    mn::CsrIn none = {nullptr, nullptr, nullptr};
    return rank_dispatch(X, ldx, none, row_index, col_index, M, G, d_hi, d_lo, ldd, inv_norm, flags, xnorm, rank2,
                         drop_mode, stream);
}

extern "C" int mn_rank_normalize_csr(const int64_t* indptr, const int32_t* indices, const float* data,
                                     const int32_t* row_index, int64_t M, int32_t G, mn_half_t* d_hi, mn_half_t* d_lo,
                                     int64_t ldd, float* inv_norm, uint8_t* flags, float* xnorm, int32_t* rank2,
                                     int32_t drop_mode, mn_stream_t stream) {
    MN_CHECK_ARG(indptr && indices && data, "mn_rank_normalize_csr: null CSR arrays");
    mn::CsrIn csr = {indptr, indices, data};
    return rank_dispatch(nullptr, 0, csr, row_index, nullptr, M, G, d_hi, d_lo, ldd, inv_norm, flags, xnorm, rank2,
                         drop_mode, stream);
}

// K1 straight into the operand format of the centroid paths (int8 limb planes + double scale, see limbs.cu):
// no fp16 planes are written and no separate split pass runs.
extern "C" int mn_rank_normalize_limbs(const float* X, int64_t ldx, const int32_t* row_index, const int32_t* col_index,
                                       int64_t M, int32_t G, int8_t* a1, int8_t* a0, int64_t ldk, double* inv_norm64,
                                       float* inv_norm, uint8_t* flags, int32_t drop_mode, mn_stream_t stream) {
    MN_CHECK_ARG(X != nullptr && a1 != nullptr, "mn_rank_normalize_limbs: null pointer");
    mn::CsrIn none = {nullptr, nullptr, nullptr};
    return rank_dispatch(X, ldx, none, row_index, col_index, M, G, nullptr, nullptr, 0, inv_norm, flags, nullptr, nullptr,
                         drop_mode, stream, mn::LimbOut{a1, a0, ldk, inv_norm64});
}

extern "C" int mn_rank_normalize_csr_limbs(const int64_t* indptr, const int32_t* indices, const float* data,
                                           const int32_t* row_index, int64_t M, int32_t G, int8_t* a1, int8_t* a0,
                                           int64_t ldk, double* inv_norm64, float* inv_norm, uint8_t* flags,
                                           int32_t drop_mode, mn_stream_t stream) {
    MN_CHECK_ARG(indptr && indices && data && a1, "mn_rank_normalize_csr_limbs: null pointer");
    mn::CsrIn csr = {indptr, indices, data};
    return rank_dispatch(nullptr, 0, csr, row_index, nullptr, M, G, nullptr, nullptr, 0, inv_norm, flags, nullptr, nullptr,
                         drop_mode, stream, mn::LimbOut{a1, a0, ldk, inv_norm64});
}
