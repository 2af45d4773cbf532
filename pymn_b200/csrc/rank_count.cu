// K1 (count path): per-cell average-tie ranking by COUNTING, one warp per cell.
//
// scRNA-seq expression rows are counts: non-negative integers, mostly 0, rarely above a few
// hundred.  For such a row the rank of an entry depends only on its value, so no sort is needed:
//   1. stream the row in (128-bit coalesced loads, four in flight per lane); stash each value as
//      uint16 in warp-private shared memory and histogram the non-zero values (shared atomics,
//      2048 bins); zeros are only counted;
//   2. an exclusive scan over the occupied bins turns the histogram into a table
//      2*rank(v) = 2*(n_zero + #{non-zeros < v}) + cnt(v) + 1   (average-tie rank, doubled);
//   3. second sweep over the stash: look up 2*rank, centre (d = 2r - (G+1)), accumulate the exact
//      integer sum of squares, write the fp16 plane(s); then 1/||d|| and the flags.
// About 1,000 warp instructions per 2,000-gene cell -- the kernel runs at HBM speed (4 B read +
// 2 B written per element).  A row that is not made of integers in [0, 2048) (log-normalised or
// scaled input) is flagged (bit 4) and ranked by the register-sort kernel (rank_warp.cu) instead;
// the two produce identical results (same tie-averaged ranks).
// Replaces reference pymn/utils.py:141-147 (bottleneck.rankdata(axis=1), centre, L2-normalise).
#include "common.cuh"

namespace mn {

constexpr int RC_WARPS = 4;
constexpr int RC_HB = 2048;          // histogram bins = largest count + 1 this path handles

__global__ void __launch_bounds__(RC_WARPS * 32)
rank_count_kernel(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ row_index,
                  const int32_t* __restrict__ col_index, int64_t M, int G, int g_pad, __half* __restrict__ d_hi,
                  __half* __restrict__ d_lo, int64_t ldd, float* __restrict__ inv_norm, uint8_t* __restrict__ flags,
                  float* __restrict__ xnorm, int32_t* __restrict__ rank2, int drop_mode, CsrIn csr, LimbOut limb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // 16-bit bins (counts <= G and 2*rank <= 2G + 1 fit for G < 32,768; the launcher checks): two per word,
    // incremented with a 32-bit shared atomic on the right half.  8 KB per warp instead of 12 -> 28 warps/SM.
    const size_t per_warp = (size_t)RC_HB * 2 + (size_t)g_pad * 2;
    uint32_t* tab32 = reinterpret_cast<uint32_t*>(smem_raw + wib * per_warp);   // histogram, then 2*rank table
    uint16_t* tab = reinterpret_cast<uint16_t*>(tab32);
    uint16_t* vals = tab + RC_HB;                                                // the row as uint16
    const int64_t warps_total = (int64_t)gridDim.x * RC_WARPS;
    const bool vec4 = (col_index == nullptr) && ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);

    for (int b = lane; b < RC_HB / 2; b += 32) tab32[b] = 0u;
    __syncwarp();

    for (int64_t row = (int64_t)blockIdx.x * RC_WARPS + wib; row < M; row += warps_total) {
        const int64_t src = row_index ? (int64_t)row_index[row] : row;
        const float* x = X + src * ldx;
        // ---- 1. load, stash, histogram -------------------------------------------------------
        bool ok = true;
        int nnz = 0, maxv = 0;
        float ps = 0.f;
        if (csr.indptr) {
            // CSR row: only the stored values travel (8 B per non-zero instead of 4 B per gene)
            for (int g2 = lane; g2 < (g_pad >> 1); g2 += 32) reinterpret_cast<uint32_t*>(vals)[g2] = 0u;
            __syncwarp();
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
                    if (eb + u * 32 + lane >= e1) continue;
                    int iv = __float2int_rz(f[u]);
                    ok = ok && ((unsigned)iv < (unsigned)RC_HB) && ((float)iv == f[u]) && ((unsigned)jj[u] < (unsigned)G);
                    iv = min(max(iv, 0), RC_HB - 1);
                    if (iv > 0) {
                        atomicAdd(&tab32[iv >> 1], 1u << ((iv & 1) << 4));
                        ++nnz;
                        maxv = max(maxv, iv);
                        vals[min(max(jj[u], 0), G - 1)] = (uint16_t)iv;
                    }
                    ps += f[u];
                }
            }
        } else if (vec4) {
            const float4* x4 = reinterpret_cast<const float4*>(x);
            const int G4 = G >> 2;
            for (int base = 0; base < G4; base += 128) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = base + u * 32 + lane;
                    v[u] = q < G4 ? __ldg(x4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = base + u * 32 + lane;
                    if (q < G4) {
                        const float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                        int iv[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            // integers in [0, RC_HB): the unsigned compare also rejects negatives and NaN
                            iv[t] = __float2int_rz(f[t]);
                            ok = ok && ((unsigned)iv[t] < (unsigned)RC_HB) && ((float)iv[t] == f[t]);
                            iv[t] = min(max(iv[t], 0), RC_HB - 1);
                            if (iv[t] > 0) {
                                atomicAdd(&tab32[iv[t] >> 1], 1u << ((iv[t] & 1) << 4));
                                ++nnz;
                                maxv = max(maxv, iv[t]);
                            }
                        }
                        ps += (f[0] + f[1]) + (f[2] + f[3]);
                        *reinterpret_cast<uint2*>(vals + 4 * q) =
                            make_uint2((uint32_t)iv[0] | ((uint32_t)iv[1] << 16), (uint32_t)iv[2] | ((uint32_t)iv[3] << 16));
                    }
                }
            }
            for (int g = (G4 << 2) + lane; g < G; g += 32) {
                const float f = __ldg(x + g);
                int iv = __float2int_rz(f);
                ok = ok && ((unsigned)iv < (unsigned)RC_HB) && ((float)iv == f);
                iv = min(max(iv, 0), RC_HB - 1);
                if (iv > 0) {
                    atomicAdd(&tab32[iv >> 1], 1u << ((iv & 1) << 4));
                    ++nnz;
                    maxv = max(maxv, iv);
                }
                ps += f;
                vals[g] = (uint16_t)iv;
            }
        } else {
            for (int g = lane; g < G; g += 32) {
                const float f = __ldg(x + (col_index ? col_index[g] : g));
                int iv = __float2int_rz(f);
                ok = ok && ((unsigned)iv < (unsigned)RC_HB) && ((float)iv == f);
                iv = min(max(iv, 0), RC_HB - 1);
                if (iv > 0) {
                    atomicAdd(&tab32[iv >> 1], 1u << ((iv & 1) << 4));
                    ++nnz;
                    maxv = max(maxv, iv);
                }
                ps += f;
                vals[g] = (uint16_t)iv;
            }
        }
        ok = __all_sync(0xffffffffu, ok);
        nnz = warp_sum(nnz);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxv = max(maxv, __shfl_xor_sync(0xffffffffu, maxv, o));
        const double psum = warp_sum((double)ps);
        __syncwarp();
        if (!ok) {
            // not a row of small non-negative integers: hand it to the sorting kernel
            for (int b = lane; b <= (maxv >> 1); b += 32) tab32[b] = 0u;
            if (lane == 0) flags[row] = 16;
            __syncwarp();
            continue;
        }
        // ---- 2. histogram -> 2*rank table (bins 1..maxv) -------------------------------------------
        const int nzero = G - nnz;
        const uint32_t zero_r2 = (uint32_t)(nzero + 1);                 // no negatives on this path
        const int c = G + 1;
        unsigned long long ss = 0;
        int below = 0;                                                  // non-zeros smaller than the chunk
        for (int b0 = 1; b0 <= maxv; b0 += 32) {
            const int b = b0 + lane;
            const int cnt = (b <= maxv) ? (int)tab[b] : 0;
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (b <= maxv) {
                const int r2 = 2 * (nzero + below + inc - cnt) + cnt + 1;
                tab[b] = (uint16_t)r2;
                const long long d = (long long)r2 - c;             // all cnt entries of this value share d
                ss += (unsigned long long)(d * d) * (unsigned long long)cnt;
            }
            below += __shfl_sync(0xffffffffu, inc, 31);
        }
        __syncwarp();
        // ---- 3. look up, centre, write ------------------------------------------------------------------
        // the sum of squares comes from the histogram: sum_v cnt(v) * d(v)^2 (+ the zero block), no extra sweep
        // (an all-zero or constant row has a single tie block: d == 0 everywhere, detected through ss == 0)
        if (lane == 0) {
            const long long d = (long long)zero_r2 - c;
            ss += (unsigned long long)(d * d) * (unsigned long long)nzero;
        }
        ss = warp_sum(ss);
        const bool zv = (ss == 0ull);
        const bool nonpos = !(psum > 0.0);
        const bool drop = zv || (drop_mode == 1 && nonpos);
        const double inv_true = zv ? 0.0 : 1.0 / sqrt((double)ss);
        if (lane == 0) {
            inv_norm[row] = drop ? 0.f : (float)inv_true;
            flags[row] = (uint8_t)((zv ? 1 : 0) | (nonpos ? 2 : 0) | (drop ? 4 : 0));
        }
        if (limb.a1) {
            // int8 limb planes (the centroid paths): four genes per lane and step, one 32-bit store per plane
            uint32_t* p1 = reinterpret_cast<uint32_t*>(limb.a1 + row * limb.ldk);
            uint32_t* p0 = reinterpret_cast<uint32_t*>(limb.a0 + row * limb.ldk);
            const bool one_limb = G <= 127;
            for (int g4 = lane; g4 < (int)(limb.ldk >> 2); g4 += 32) {
                const int g = 4 * g4;
                int d[4] = {0, 0, 0, 0};
                if (!drop && g < G) {
                    const uint2 pr = *reinterpret_cast<const uint2*>(vals + g);       // g_pad is a multiple of 8
                    const int v[4] = {(int)(pr.x & 0xffffu), (int)(pr.x >> 16), (int)(pr.y & 0xffffu), (int)(pr.y >> 16)};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (g + u < G) d[u] = (int)(v[u] ? tab[v[u]] : zero_r2) - c;
                }
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
            int d0 = 0, d1 = 0;
            if (!drop && g < G) {
                const uint32_t pr = *reinterpret_cast<const uint32_t*>(vals + g);
                const int v0 = pr & 0xffffu, v1 = pr >> 16;
                d0 = (int)(v0 ? tab[v0] : zero_r2) - c;
                if (g + 1 < G) d1 = (int)(v1 ? tab[v1] : zero_r2) - c;
            }
            const __half h0 = __int2half_rn(d0), h1 = __int2half_rn(d1);
            ph[g2] = __halves2half2(h0, h1);
            if (pl) pl[g2] = __halves2half2(__int2half_rn(d0 - __half2int_rn(h0)), __int2half_rn(d1 - __half2int_rn(h1)));
        }
        if (xnorm) {
            float* xo = xnorm + row * (int64_t)G;
            const float qnan = __int_as_float(0x7fc00000);
            for (int g = lane; g < G; g += 32) {
                const int v = vals[g];
                xo[g] = zv ? qnan : (float)((double)((int)(v ? tab[v] : zero_r2) - c) * inv_true);
            }
        }
        if (rank2) {
            int32_t* ro = rank2 + row * (int64_t)G;
            for (int g = lane; g < G; g += 32) {
                const int v = vals[g];
                ro[g] = (int)(v ? tab[v] : zero_r2);
            }
        }
        __syncwarp();
        for (int b = lane; b <= (maxv >> 1); b += 32) tab32[b] = 0u;   // leave the histogram clean for the next row
        __syncwarp();
    }
}

int launch_rank_count(const float* X, int64_t ldx, const int32_t* row_index, const int32_t* col_index, int64_t M, int G,
                      __half* d_hi, __half* d_lo, int64_t ldd, float* inv_norm, uint8_t* flags, float* xnorm,
                      int32_t* rank2, int drop_mode, cudaStream_t st, CsrIn csr, LimbOut limb) {
    const int g_pad = (G + 7) & ~7;
    if (G >= 32768) {
        set_error("rank_count: G=%d exceeds the 16-bit rank table (max 32767)", G);
        return 1;
    }
    const size_t per_warp = (size_t)RC_HB * 2 + (size_t)g_pad * 2;
    const size_t smem = per_warp * RC_WARPS;
    // (set on every launch: the attribute is per device and per context, and the call costs about a microsecond)
    if (smem > 40 * 1024)
        MN_CUDA(cudaFuncSetAttribute(rank_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t blocks = (M + RC_WARPS - 1) / RC_WARPS;
    const int64_t max_blocks = (int64_t)sms * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    rank_count_kernel<<<(unsigned)blocks, RC_WARPS * 32, smem, st>>>(X, ldx, row_index, col_index, M, G, g_pad, d_hi, d_lo,
                                                                   ldd, inv_norm, flags, xnorm, rank2, drop_mode, csr, limb);
    MN_LAUNCH_CHECK();
    return 0;
}

}  // namespace mn
