// K1 (count path): per-cell average-tie ranking by COUNTING, one warp per cell.
//
// scRNA-seq expression rows are counts: non-negative integers, mostly 0, rarely above a few
// hundred.  For such a row the rank of an entry depends only on its value, so no sort is needed:
//   1. stream the row in (128-bit coalesced loads, four in flight per lane) and histogram the non-zero
//      values (2048 32-bit bins in warp-private shared memory, one predicated shared-memory reduction per
//      element); zeros are not touched at all -- their count is G minus the rest;
//   2. a scan over the occupied bins turns the histogram into a table of OUTPUT-READY entries: for every value
//      the centred doubled rank d = 2*rank - (G+1), already encoded the way it is stored (fp16 bits of the
//      plane(s), or the two int8 limbs of the centroid paths); entry 0 is the zero block.  The sum of squares
//      comes from the histogram too (sum_v cnt(v) d(v)^2);
//   3. second sweep: value -> table entry -> packed stores.  A dense row-major matrix is re-read from L2 (the row
//      was streamed a few microseconds earlier; 8 KB per warp would otherwise go to a shared-memory stash and
//      cost a third of the resident warps); CSR input and column gathers keep a uint16 stash of the row.
// About 15 instructions per element (the first version needed ~45).
// A row that is not made of integers in [0, 2048) (log-normalised or scaled input) is flagged (bit 4) and ranked
// by the register-sort kernel (rank_warp.cu) instead; the two produce identical results (same tie-averaged ranks).
// Replaces reference pymn/utils.py:141-147 (bottleneck.rankdata(axis=1), centre, L2-normalise).
#include "common.cuh"

namespace mn {

constexpr int RC_WARPS = 14;         // warps per CTA without a stash: 14 x 8 KB of bins = 113 KB, two CTAs = 28 warps per SM
constexpr int RC_WARPS_STASH = 4;    // with the uint16 stash of the row (CSR input, column gathers): small CTAs pack better
constexpr int RC_HB = 2048;          // histogram bins = largest count + 1 this path handles
constexpr int RC_INFLIGHT = 8;       // 128-bit loads in flight per lane (4 KB per warp)

enum { RC_MODE_R2 = 0, RC_MODE_LIMB = 1, RC_MODE_PLANE1 = 2, RC_MODE_PLANE2 = 3 };

// one predicated shared-memory reduction: hist[v] += 1 when v != 0 (no branch around it)
__device__ __forceinline__ void hist_inc_nonzero(uint32_t base_addr, int v) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.s32 p, %1, 0;\n\t"
        "@p red.shared.add.u32 [%0], 1;\n\t}"
        :: "r"(base_addr + ((uint32_t)v << 2)), "r"(v) : "memory");
}

// validity and histogram of one value: integers in [0, RC_HB) only; anything else clears `ok`
__device__ __forceinline__ int count_value(float f, uint32_t hist_addr, bool& ok, int& orbits, int& maxv) {
    const int iv = __float2int_rz(f);          // NaN -> 0 (fails the equality), +-inf saturate (fail it too)
    ok = ok && ((float)iv == f);
    orbits |= iv;                              // negative or >= RC_HB: caught once per row through the high bits
    const int v = iv & (RC_HB - 1);
    maxv = max(maxv, v);
    hist_inc_nonzero(hist_addr, v);
    return v;
}

// table entry of a value whose doubled rank is r2 (centred: d = r2 - c)
template <int MODE>
__device__ __forceinline__ uint32_t encode_entry(int r2, int c, bool one_limb) {
    const int d = r2 - c;
    if (MODE == RC_MODE_R2) return (uint32_t)r2;
    if (MODE == RC_MODE_LIMB) {
        const int hi = one_limb ? d : (d + 64) >> 7;
        const int lo = one_limb ? 0 : d - (hi << 7);
        return (((uint32_t)hi & 0xffu) << 8) | ((uint32_t)lo & 0xffu);
    }
    const __half h = __int2half_rn(d);
    if (MODE == RC_MODE_PLANE1) return (uint32_t)__half_as_ushort(h);
    return (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(__int2half_rn(d - __half2int_rn(h))) << 16);
}

template <int MODE, bool STASH>
__global__ void __launch_bounds__(RC_WARPS * 32)
rank_count_kernel(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ row_index,
                  const int32_t* __restrict__ col_index, int64_t M, int G, int g_pad, __half* __restrict__ d_hi,
                  __half* __restrict__ d_lo, int64_t ldd, float* __restrict__ inv_norm, uint8_t* __restrict__ flags,
                  float* __restrict__ xnorm, int32_t* __restrict__ rank2, int drop_mode, CsrIn csr, LimbOut limb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t per_warp = (size_t)RC_HB * 4 + (STASH ? (size_t)g_pad * 2 : 0);
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem_raw + wib * per_warp);      // histogram, then entry table
    uint16_t* vals = reinterpret_cast<uint16_t*>(tab + RC_HB);                    // the row as uint16 (STASH only)
    const uint32_t tab_addr = (uint32_t)__cvta_generic_to_shared(tab);
    const int cta_warps = blockDim.x >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * cta_warps;
    const int G4 = G >> 2;
    const int c = G + 1;
    const bool one_limb = G <= 127;

    for (int b = lane; b < RC_HB; b += 32) tab[b] = 0u;
    __syncwarp();

    for (int64_t row = (int64_t)blockIdx.x * cta_warps + wib; row < M; row += warps_total) {
        const int64_t src = row_index ? (int64_t)row_index[row] : row;
        const float* x = X + src * ldx;
        // ---- 1. load, histogram (and stash) -------------------------------------------------------
        bool ok = true;
        int orbits = 0, maxv = 0;
        if (STASH && csr.indptr) {
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
                    ok = ok && ((unsigned)jj[u] < (unsigned)G);
                    const int v = count_value(f[u], tab_addr, ok, orbits, maxv);
                    if (v) vals[min(max(jj[u], 0), G - 1)] = (uint16_t)v;
                }
            }
        } else if (STASH) {
            for (int g = lane; g < G; g += 32) {
                const float f = __ldg(x + (col_index ? col_index[g] : g));
                vals[g] = (uint16_t)count_value(f, tab_addr, ok, orbits, maxv);
            }
        } else {
            const float4* x4 = reinterpret_cast<const float4*>(x);
            for (int base = 0; base < G4; base += 32 * RC_INFLIGHT) {
                float4 v[RC_INFLIGHT];
#pragma unroll
                for (int u = 0; u < RC_INFLIGHT; ++u) {
                    const int q = base + u * 32 + lane;
                    v[u] = q < G4 ? __ldg(x4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < RC_INFLIGHT; ++u) {         // (the zero fill of lanes past the end counts nothing)
                    count_value(v[u].x, tab_addr, ok, orbits, maxv);
                    count_value(v[u].y, tab_addr, ok, orbits, maxv);
                    count_value(v[u].z, tab_addr, ok, orbits, maxv);
                    count_value(v[u].w, tab_addr, ok, orbits, maxv);
                }
            }
            for (int g = (G4 << 2) + lane; g < G; g += 32) count_value(__ldg(x + g), tab_addr, ok, orbits, maxv);
        }
        ok = ok && ((orbits & ~(RC_HB - 1)) == 0);
        ok = __all_sync(0xffffffffu, ok);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxv = max(maxv, __shfl_xor_sync(0xffffffffu, maxv, o));
        __syncwarp();
        if (!ok) {
            // not a row of small non-negative integers: hand it to the sorting kernel
            for (int b = lane; b <= maxv; b += 32) tab[b] = 0u;
            if (lane == 0) flags[row] = 16;
            __syncwarp();
            continue;
        }
        // ---- 2. histogram -> entry table (bins 1..maxv); nnz and the sum of squares on the way ------------------
        int nnz = 0;                                   // the zero block's size fixes every rank: count first
        for (int b0 = 1; b0 <= maxv; b0 += 32) {
            const int b = b0 + lane;
            nnz += (b <= maxv) ? (int)tab[b] : 0;
        }
        nnz = warp_sum(nnz);
        const int nzero = G - nnz;
        const uint32_t zero_r2 = (uint32_t)(nzero + 1);                 // no negatives on this path
        unsigned long long ss = 0;
        int below = 0;                                                  // non-zeros smaller than the chunk
        for (int b0 = 1; b0 <= maxv; b0 += 32) {
            const int b = b0 + lane;
            const int cnt = (b <= maxv) ? (int)tab[b] : 0;
            if (!__any_sync(0xffffffffu, cnt != 0)) continue;           // (counts are sparse above a few dozen)
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (cnt) {
                const int r2 = 2 * (nzero + below + inc - cnt) + cnt + 1;
                tab[b] = encode_entry<MODE>(r2, c, one_limb);
                const long long d = (long long)r2 - c;                 // all cnt entries of this value share d
                ss += (unsigned long long)(d * d) * (unsigned long long)cnt;
            }
            below += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) {
            const long long d = (long long)zero_r2 - c;
            ss += (unsigned long long)(d * d) * (unsigned long long)nzero;
            tab[0] = encode_entry<MODE>((int)zero_r2, c, one_limb);
        }
        ss = warp_sum(ss);
        __syncwarp();
        // ---- 3. look up, write ------------------------------------------------------------------
        // (an all-zero or constant row has a single tie block: d == 0 everywhere, detected through ss == 0)
        const bool zv = (ss == 0ull);
        const bool nonpos = (nnz == 0);            // non-negative integers: the sum is positive unless all are zero
        const bool drop = zv || (drop_mode == 1 && nonpos);
        const double inv_true = zv ? 0.0 : 1.0 / sqrt((double)ss);
        if (lane == 0) {
            inv_norm[row] = drop ? 0.f : (float)inv_true;
            flags[row] = (uint8_t)((zv ? 1 : 0) | (nonpos ? 2 : 0) | (drop ? 4 : 0));
            if (MODE == RC_MODE_LIMB) limb.inv64[row] = drop ? 0.0 : inv_true;
        }
        // four values of the row, genes g .. g+3 (g a multiple of 4, g + 3 < G)
        auto quad = [&](int g, int v[4]) {
            if (STASH) {
                const uint2 pr = *reinterpret_cast<const uint2*>(vals + g);       // g_pad is a multiple of 8
                v[0] = pr.x & 0xffffu; v[1] = pr.x >> 16; v[2] = pr.y & 0xffffu; v[3] = pr.y >> 16;
            } else {
                const float4 f = __ldg(reinterpret_cast<const float4*>(x) + (g >> 2));
                v[0] = __float2int_rz(f.x); v[1] = __float2int_rz(f.y); v[2] = __float2int_rz(f.z); v[3] = __float2int_rz(f.w);
            }
        };
        auto single = [&](int g) -> int { return STASH ? (int)vals[g] : __float2int_rz(__ldg(x + g)); };
        if (MODE == RC_MODE_LIMB) {
            uint32_t* p1 = reinterpret_cast<uint32_t*>(limb.a1 + row * limb.ldk);
            uint32_t* p0 = reinterpret_cast<uint32_t*>(limb.a0 + row * limb.ldk);
            const int n4 = (int)(limb.ldk >> 2);
            for (int gb = 0; gb < n4; gb += 128) {
                int v[4][4];
                // four independent quads per lane: their loads (the L2 re-read of the row) are all in flight together
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int g = 4 * (gb + w * 32 + lane);
                    if (!drop && g + 3 < G) quad(g, v[w]);
                }
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int g4 = gb + w * 32 + lane;
                    const int g = 4 * g4;
                    if (g4 >= n4) continue;
                    uint32_t e[4] = {0u, 0u, 0u, 0u};
                    if (!drop && g + 3 < G) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) e[u] = tab[v[w][u]];
                    } else if (!drop) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (g + u < G) e[u] = tab[single(g + u)];
                    }
                    const uint32_t r01 = e[0] | (e[1] << 16), r23 = e[2] | (e[3] << 16);
                    p1[g4] = __byte_perm(r01, r23, 0x7531);
                    p0[g4] = __byte_perm(r01, r23, 0x6420);
                }
            }
        } else if (MODE == RC_MODE_PLANE1 || MODE == RC_MODE_PLANE2) {
            uint2* ph = reinterpret_cast<uint2*>(d_hi + row * ldd);
            uint2* pl = reinterpret_cast<uint2*>(d_lo + row * ldd);
            const int n4 = (int)(ldd >> 2);                                       // ldd is a multiple of 8
            for (int gb = 0; gb < n4; gb += 128) {
                int v[4][4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {                                     // four quads in flight per lane
                    const int g = 4 * (gb + w * 32 + lane);
                    if (!drop && g + 3 < G) quad(g, v[w]);
                }
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int g4 = gb + w * 32 + lane;
                    const int g = 4 * g4;
                    if (g4 >= n4) continue;
                    uint32_t e[4] = {0u, 0u, 0u, 0u};
                    if (!drop && g + 3 < G) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) e[u] = tab[v[w][u]];
                    } else if (!drop) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (g + u < G) e[u] = tab[single(g + u)];
                    }
                    if (MODE == RC_MODE_PLANE1) {
                        ph[g4] = make_uint2(e[0] | (e[1] << 16), e[2] | (e[3] << 16));
                    } else {
                        ph[g4] = make_uint2(__byte_perm(e[0], e[1], 0x5410), __byte_perm(e[2], e[3], 0x5410));
                        pl[g4] = make_uint2(__byte_perm(e[0], e[1], 0x7632), __byte_perm(e[2], e[3], 0x7632));
                    }
                }
            }
        } else {
            // general form (test outputs xnorm / rank2 next to the planes): the table holds 2*rank
            __half2* ph = reinterpret_cast<__half2*>(d_hi + row * ldd);
            __half2* pl = d_lo ? reinterpret_cast<__half2*>(d_lo + row * ldd) : nullptr;
            for (int g2 = lane; d_hi && g2 < (int)(ldd >> 1); g2 += 32) {
                const int g = 2 * g2;
                int d0 = 0, d1 = 0;
                if (!drop && g < G) {
                    d0 = (int)tab[single(g)] - c;
                    if (g + 1 < G) d1 = (int)tab[single(g + 1)] - c;
                }
                const __half h0 = __int2half_rn(d0), h1 = __int2half_rn(d1);
                ph[g2] = __halves2half2(h0, h1);
                if (pl) pl[g2] = __halves2half2(__int2half_rn(d0 - __half2int_rn(h0)), __int2half_rn(d1 - __half2int_rn(h1)));
            }
            if (xnorm) {
                float* xo = xnorm + row * (int64_t)G;
                const float qnan = __int_as_float(0x7fc00000);
                for (int g = lane; g < G; g += 32)
                    xo[g] = zv ? qnan : (float)((double)((int)tab[single(g)] - c) * inv_true);
            }
            if (rank2) {
                int32_t* ro = rank2 + row * (int64_t)G;
                for (int g = lane; g < G; g += 32) ro[g] = (int)tab[single(g)];
            }
        }
        __syncwarp();
        for (int b = lane; b <= maxv; b += 32) tab[b] = 0u;   // leave the histogram clean for the next row
        __syncwarp();
    }
}

template <int MODE, bool STASH>
static int launch_mode(const float* X, int64_t ldx, const int32_t* row_index, const int32_t* col_index, int64_t M, int G,
                       __half* d_hi, __half* d_lo, int64_t ldd, float* inv_norm, uint8_t* flags, float* xnorm,
                       int32_t* rank2, int drop_mode, cudaStream_t st, CsrIn csr, LimbOut limb) {
    const int g_pad = (G + 7) & ~7;
    const int cta_warps = STASH ? RC_WARPS_STASH : RC_WARPS;
    const size_t per_warp = (size_t)RC_HB * 4 + (STASH ? (size_t)g_pad * 2 : 0);
    const size_t smem = per_warp * cta_warps;
    // (set on every launch: the attribute is per device and per context, and the call costs about a microsecond)
    if (smem > 40 * 1024)
        MN_CUDA(cudaFuncSetAttribute(rank_count_kernel<MODE, STASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // a persistent grid of exactly the resident CTAs: more would run as a second, partly filled wave
    int per_sm = 1;
    MN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rank_count_kernel<MODE, STASH>, cta_warps * 32, smem));
    int64_t blocks = (M + cta_warps - 1) / cta_warps;
    const int64_t max_blocks = (int64_t)sms * (per_sm > 0 ? per_sm : 1);
    if (blocks > max_blocks) blocks = max_blocks;
    rank_count_kernel<MODE, STASH><<<(unsigned)blocks, cta_warps * 32, smem, st>>>(
        X, ldx, row_index, col_index, M, G, g_pad, d_hi, d_lo, ldd, inv_norm, flags, xnorm, rank2, drop_mode, csr, limb);
    MN_LAUNCH_CHECK();
    return 0;
}

int launch_rank_count(const float* X, int64_t ldx, const int32_t* row_index, const int32_t* col_index, int64_t M, int G,
                      __half* d_hi, __half* d_lo, int64_t ldd, float* inv_norm, uint8_t* flags, float* xnorm,
                      int32_t* rank2, int drop_mode, cudaStream_t st, CsrIn csr, LimbOut limb) {
    // the row is re-read from L2 in the write sweep unless it has to be assembled first (CSR, column gather) or
    // cannot be read with aligned 128-bit loads
    const bool stash = csr.indptr || col_index || (ldx & 3) || (reinterpret_cast<uintptr_t>(X) & 15);
#define MN_RC_LAUNCH(MODE)                                                                                            \
    return stash ? launch_mode<MODE, true>(X, ldx, row_index, col_index, M, G, d_hi, d_lo, ldd, inv_norm, flags, xnorm, \
                                           rank2, drop_mode, st, csr, limb)                                          \
                 : launch_mode<MODE, false>(X, ldx, row_index, col_index, M, G, d_hi, d_lo, ldd, inv_norm, flags, xnorm, \
                                            rank2, drop_mode, st, csr, limb)
    if (limb.a1 && (d_hi || xnorm || rank2)) {
        set_error("rank_count: limb output cannot be combined with planes / xnorm / rank2");
        return 1;
    }
    if (xnorm || rank2) { MN_RC_LAUNCH(RC_MODE_R2); }
    if (limb.a1) { MN_RC_LAUNCH(RC_MODE_LIMB); }
    if (d_lo) { MN_RC_LAUNCH(RC_MODE_PLANE2); }
    MN_RC_LAUNCH(RC_MODE_PLANE1);
#undef MN_RC_LAUNCH
}

}  // namespace mn
