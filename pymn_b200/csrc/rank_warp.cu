// K1 (fast path): per-cell average-tie ranking, ONE WARP PER CELL, no block-wide barriers.
//
// Same contract as rank_normalize_kernel (rank_normalize.cu; replaces reference
// pymn/utils.py:122-148), but the sort lives in registers:
//   1. the warp streams the cell's G values in (128-bit coalesced loads, four in flight per lane),
//      compacts the non-zeros into a warp-private shared-memory buffer as
//      (order-preserving key << 32 | gene) -- the zeros are one tie block with a closed-form rank
//      and are never sorted;
//   2. the non-zeros are loaded into registers (E per lane, 32*E >= nnz, E in {2, 8, 16, 32})
//      and sorted by a bitonic network: compare-exchanges between registers of a lane for the
//      small strides, warp shuffles for the large ones;
//   3. tie runs: forward max-scan of run starts / backward min-scan of run ends (in-lane walk +
//      one warp scan each) -> 2*rank = first + last + 2, scattered by gene into a warp-private
//      uint16 staging row;
//   4. centre (d = 2r - (G+1)), exact integer sum of squares, fp16 plane(s), 1/||d||, flags.
// A row with more than 1024 non-zeros (dense input; a scRNA-seq row of 2-3k HVGs has a few hundred)
// is flagged (bit 3) and redone by the CTA-per-cell kernel.
#include "common.cuh"

namespace mn {

constexpr int RW_WARPS = 4;                 // warps per CTA
constexpr int RW_CAP = 1024;                // non-zeros a warp can sort (32 lanes x 32 registers)

template <int E>
__device__ __forceinline__ void warp_bitonic_sort(uint64_t (&x)[E], int lane) {
    // element position i = lane * E + e (blocked); ascending over i
    constexpr int CAP = 32 * E;
#pragma unroll 1
    for (int k = 2; k <= CAP; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= E) {
                // partner in lane ^ (j / E), same register
                const int lm = j / E;
                const bool up = (k >= CAP) || ((lane & (k / E)) == 0);
                const bool keep_min = ((lane & lm) == 0) == up;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const uint64_t y = __shfl_xor_sync(0xffffffffu, x[e], lm);
                    const bool less = x[e] < y;
                    x[e] = (less == keep_min) ? x[e] : y;
                }
            } else {
                // partner in the same lane: registers e and e | j   (j < E is a power of two)
                const bool lane_up = (k >= CAP) || (k < E) || ((lane & (k / E)) == 0);
#pragma unroll
                for (int jj = 1; jj < E; jj <<= 1) {
                    if (jj == j) {
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            if ((e & jj) == 0) {
                                const int f = e | jj;
                                const bool up = (k < E) ? ((e & k) == 0) : lane_up;
                                const uint64_t a = x[e], b = x[f];
                                const bool sw = (a > b) == up;
                                x[e] = sw ? b : a;
                                x[f] = sw ? a : b;
                            }
                        }
                    }
                }
            }
        }
    }
}

__device__ __forceinline__ int warp_excl_max_scan(int v, int lane) {       // exclusive max-scan, identity -1
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, t);
    }
    const int ex = __shfl_up_sync(0xffffffffu, inc, 1);
    return lane == 0 ? -1 : ex;
}

__device__ __forceinline__ int warp_excl_min_scan_rev(int v, int lane) {   // exclusive min-scan from lane 31 down
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_down_sync(0xffffffffu, inc, o);
        if (lane + o < 32) inc = min(inc, t);
    }
    const int ex = __shfl_down_sync(0xffffffffu, inc, 1);
    return lane == 31 ? INT_MAX : ex;
}

// Sort the warp's nnz compacted entries and write 2*rank into r2s[gene].  An entry is
// key(32) | run-start scratch(16) | gene(16); positions are < 2048 and genes < 16384.
template <int E>
__device__ __forceinline__ void warp_rank_nonzeros(const uint64_t* buf, int nnz, int nzero, uint16_t* r2s, int lane) {
    uint64_t x[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int s = e * 32 + lane;                        // conflict-free read; the initial order is irrelevant
        x[e] = s < nnz ? buf[s] : ~0ull;
    }
    warp_bitonic_sort<E>(x, lane);
    const int p0 = lane * E;
    // keys of the neighbours across lane boundaries
    const uint32_t prev_last = (uint32_t)(__shfl_up_sync(0xffffffffu, x[E - 1], 1) >> 32);
    const uint32_t next_first = (uint32_t)(__shfl_down_sync(0xffffffffu, x[0], 1) >> 32);
    // ---- first position of each tie run: in-lane walk, warp scan of the lane results, second walk
    int cur = -1;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int p = p0 + e;
        const uint32_t key = (uint32_t)(x[e] >> 32);
        const uint32_t pk = e == 0 ? prev_last : (uint32_t)(x[e > 0 ? e - 1 : 0] >> 32);
        if (p < nnz && (p == 0 || key != pk)) cur = p;
    }
    cur = warp_excl_max_scan(cur, lane);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int p = p0 + e;
        const uint32_t key = (uint32_t)(x[e] >> 32);
        const uint32_t pk = e == 0 ? prev_last : (uint32_t)(x[e > 0 ? e - 1 : 0] >> 32);
        if (p < nnz && (p == 0 || key != pk)) cur = p;
        // stash the run start in bits 16..31 (it is consumed below; keys of neighbours stay intact)
        x[e] = (x[e] & 0xffffffff0000ffffull) | ((uint64_t)(uint32_t)(cur < 0 ? 0 : cur) << 16);
    }
    // ---- last position of each tie run, then 2*rank = first + last + 2 ----------------------
    int curE = INT_MAX;
#pragma unroll
    for (int e = E - 1; e >= 0; --e) {
        const int p = p0 + e;
        const uint32_t key = (uint32_t)(x[e] >> 32);
        const uint32_t nk = e == E - 1 ? next_first : (uint32_t)(x[e < E - 1 ? e + 1 : e] >> 32);
        if (p < nnz && (p == nnz - 1 || key != nk)) curE = p;
    }
    curE = warp_excl_min_scan_rev(curE, lane);
#pragma unroll
    for (int e = E - 1; e >= 0; --e) {
        const int p = p0 + e;
        const uint32_t key = (uint32_t)(x[e] >> 32);
        const uint32_t nk = e == E - 1 ? next_first : (uint32_t)(x[e < E - 1 ? e + 1 : e] >> 32);
        if (p < nnz && (p == nnz - 1 || key != nk)) curE = p;
        if (p < nnz) {
            const uint32_t low = (uint32_t)x[e];
            const int first = (int)(low >> 16);
            const int off = (key & 0x80000000u) ? nzero : 0;      // positives sit above the zero block
            r2s[low & 0xffffu] = (uint16_t)(first + curE + 2 + 2 * off);
        }
    }
}

// one compaction step: every lane offers one value
__device__ __forceinline__ void compact_one(float v, bool in, int g, uint64_t* buf, int cap, int& nnz, int& nneg,
                                            unsigned lt) {
    const bool nz = in && v != 0.f;
    const unsigned bal = __ballot_sync(0xffffffffu, nz);
    nneg += __popc(__ballot_sync(0xffffffffu, in && v < 0.f));
    const int pos = nnz + __popc(bal & lt);
    if (nz && pos < cap) buf[pos] = ((uint64_t)float_key(v) << 32) | (uint32_t)g;
    nnz += __popc(bal);
}

__global__ void __launch_bounds__(RW_WARPS * 32)
rank_warp_kernel(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ row_index,
                 const int32_t* __restrict__ col_index, int64_t M, int G, int cap, int g_pad,
                 __half* __restrict__ d_hi, __half* __restrict__ d_lo, int64_t ldd, float* __restrict__ inv_norm,
                 uint8_t* __restrict__ flags, float* __restrict__ xnorm, int32_t* __restrict__ rank2, int drop_mode,
                 int only_mask, CsrIn csr, LimbOut limb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    const size_t per_warp = (size_t)cap * 8 + (size_t)g_pad * 2;
    uint64_t* buf = reinterpret_cast<uint64_t*>(smem_raw + wib * per_warp);
    uint16_t* r2s = reinterpret_cast<uint16_t*>(buf + cap);
    const int64_t warps_total = (int64_t)gridDim.x * RW_WARPS;
    const bool vec4 = (col_index == nullptr) && ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);

    for (int64_t row = (int64_t)blockIdx.x * RW_WARPS + wib; row < M; row += warps_total) {
        // second stage of the counting kernel: only the rows it flagged as "not small integers"
        if (only_mask && !(flags[row] & only_mask)) continue;
        const int64_t src = row_index ? (int64_t)row_index[row] : row;
        const float* x = X + src * ldx;
        // ---- 1. load + compact -------------------------------------------------------------
        int nnz = 0, nneg = 0;
        float ps = 0.f;
        if (csr.indptr) {
            const int64_t e0 = csr.indptr[src], e1 = csr.indptr[src + 1];
            for (int64_t eb = e0; eb < e1; eb += 128) {
                float f[4];
                int jj[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t e = eb + u * 32 + lane;
                    f[u] = e < e1 ? __ldg(csr.data + e) : 0.f;
                    jj[u] = e < e1 ? __ldg(csr.indices + e) : 0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool in = (eb + u * 32 + lane < e1) && ((unsigned)jj[u] < (unsigned)G);
                    ps += in ? f[u] : 0.f;
                    compact_one(f[u], in, jj[u], buf, cap, nnz, nneg, lt);
                }
            }
        } else if (vec4) {
            const float4* x4 = reinterpret_cast<const float4*>(x);
            const int G4 = G >> 2;
            for (int base = 0; base < G4; base += 128) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {                     // four 128-bit loads in flight per lane
                    const int q = base + u * 32 + lane;
                    v[u] = q < G4 ? __ldg(x4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = base + u * 32 + lane;
                    const bool in = q < G4;
                    ps += (v[u].x + v[u].y) + (v[u].z + v[u].w);
                    compact_one(v[u].x, in, 4 * q + 0, buf, cap, nnz, nneg, lt);
                    compact_one(v[u].y, in, 4 * q + 1, buf, cap, nnz, nneg, lt);
                    compact_one(v[u].z, in, 4 * q + 2, buf, cap, nnz, nneg, lt);
                    compact_one(v[u].w, in, 4 * q + 3, buf, cap, nnz, nneg, lt);
                }
            }
            if (G & 3) {                                           // up to three trailing values
                const int g = (G4 << 2) + lane;
                const bool in = g < G;
                const float v = in ? __ldg(x + g) : 0.f;
                ps += v;
                compact_one(v, in, g, buf, cap, nnz, nneg, lt);
            }
        } else {
            for (int base = 0; base < G; base += 32) {
                const int g = base + lane;
                const bool in = g < G;
                const float v = in ? __ldg(x + (col_index ? col_index[g] : g)) : 0.f;
                ps += v;
                compact_one(v, in, g, buf, cap, nnz, nneg, lt);
            }
        }
        const double psum = warp_sum((double)ps);
        if (nnz > cap) {                        // leave this row to the CTA-per-cell kernel
            if (lane == 0) flags[row] = 8;
            continue;
        }
        const int nzero = G - nnz;
        const uint16_t zero_r2 = (uint16_t)(2 * nneg + nzero + 1);
        for (int g = lane; g < g_pad; g += 32) r2s[g] = zero_r2;
        __syncwarp();

        // ---- 2./3. sort + tie runs -------------------------------------------------------------
        if (nnz > 512) warp_rank_nonzeros<32>(buf, nnz, nzero, r2s, lane);
        else if (nnz > 256) warp_rank_nonzeros<16>(buf, nnz, nzero, r2s, lane);
        else if (nnz > 64) warp_rank_nonzeros<8>(buf, nnz, nzero, r2s, lane);
        else if (nnz > 0) warp_rank_nonzeros<2>(buf, nnz, nzero, r2s, lane);
        __syncwarp();

        // ---- 4. centre, norm, write -------------------------------------------------------------
        const int c = G + 1;
        unsigned long long ss = 0;
        for (int g = lane; g < G; g += 32) {
            const long long d = (long long)r2s[g] - c;
            ss += (unsigned long long)(d * d);
        }
        ss = warp_sum(ss);
        const bool zerovar = (ss == 0ull);
        const bool nonpos = !(psum > 0.0);
        const bool drop = zerovar || (drop_mode == 1 && nonpos);
        const double inv_true = zerovar ? 0.0 : 1.0 / sqrt((double)ss);
        if (lane == 0) {
            inv_norm[row] = drop ? 0.f : (float)inv_true;
            flags[row] = (uint8_t)((zerovar ? 1 : 0) | (nonpos ? 2 : 0) | (drop ? 4 : 0));
        }
        if (limb.a1) {
            uint32_t* p1 = reinterpret_cast<uint32_t*>(limb.a1 + row * limb.ldk);
            uint32_t* p0 = reinterpret_cast<uint32_t*>(limb.a0 + row * limb.ldk);
            const bool one_limb = G <= 127;
            for (int g4 = lane; g4 < (int)(limb.ldk >> 2); g4 += 32) {
                int d[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) d[u] = (!drop && 4 * g4 + u < G) ? (int)r2s[4 * g4 + u] - c : 0;
                uint32_t w1, w0;
                limb_pack4(d, one_limb, w1, w0);
                p1[g4] = w1;
                p0[g4] = w0;
            }
            if (lane == 0) limb.inv64[row] = drop ? 0.0 : inv_true;
        }
        __half2* ph = reinterpret_cast<__half2*>(d_hi + row * ldd);
        __half2* pl = d_lo ? reinterpret_cast<__half2*>(d_lo + row * ldd) : nullptr;
        for (int g2 = lane; d_hi && g2 < (int)(ldd >> 1); g2 += 32) {
            const int g = 2 * g2;
            const int d0 = (!drop && g < G) ? (int)r2s[g] - c : 0;
            const int d1 = (!drop && g + 1 < G) ? (int)r2s[g + 1] - c : 0;
            const __half h0 = __int2half_rn(d0), h1 = __int2half_rn(d1);
            ph[g2] = __halves2half2(h0, h1);
            if (pl) pl[g2] = __halves2half2(__int2half_rn(d0 - __half2int_rn(h0)), __int2half_rn(d1 - __half2int_rn(h1)));
        }
        if (xnorm) {
            float* xo = xnorm + row * (int64_t)G;
            const float qnan = __int_as_float(0x7fc00000);
            for (int g = lane; g < G; g += 32) xo[g] = zerovar ? qnan : (float)((double)((int)r2s[g] - c) * inv_true);
        }
        if (rank2) {
            int32_t* ro = rank2 + row * (int64_t)G;
            for (int g = lane; g < G; g += 32) ro[g] = (int)r2s[g];
        }
        __syncwarp();
    }
}

// returns 0 on success; *may_overflow tells the caller to run the CTA kernel on flagged rows
int launch_rank_warp(const float* X, int64_t ldx, const int32_t* row_index, const int32_t* col_index, int64_t M, int G,
                     __half* d_hi, __half* d_lo, int64_t ldd, float* inv_norm, uint8_t* flags, float* xnorm,
                     int32_t* rank2, int drop_mode, cudaStream_t st, bool* may_overflow, int only_mask, CsrIn csr,
                     LimbOut limb) {
    int cap = 64;
    while (cap < G && cap < RW_CAP) cap <<= 1;          // never more than the row can hold
    const int g_pad = (G + 7) & ~7;
    const size_t per_warp = (size_t)cap * 8 + (size_t)g_pad * 2;
    const size_t smem = per_warp * RW_WARPS;
    *may_overflow = G > cap;
    // (set on every launch: the attribute is per device and per context, and the call costs about a microsecond)
    if (smem > 40 * 1024)
        MN_CUDA(cudaFuncSetAttribute(rank_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t blocks = (M + RW_WARPS - 1) / RW_WARPS;
    const int64_t max_blocks = (int64_t)sms * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    rank_warp_kernel<<<(unsigned)blocks, RW_WARPS * 32, smem, st>>>(X, ldx, row_index, col_index, M, G, cap, g_pad, d_hi,
                                                                  d_lo, ldd, inv_norm, flags, xnorm, rank2, drop_mode,
                                                                  only_mask, csr, limb);
    MN_LAUNCH_CHECK();
    return 0;
}

}  // namespace mn
