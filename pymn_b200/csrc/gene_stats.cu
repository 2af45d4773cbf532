// Per-study, per-gene median and variance of the expression matrix: the statistics behind
// variableGenes (reference pymn/variableGenes.py:22-27 -- bottleneck.median(X, axis=0) and
// np.var(X, axis=0) over the cells of one study; SURVEY.md section 8f-2, the step right before the hot path).
//
// X is cells x genes, row-major, so a gene is a strided column: one lane owns two neighbouring genes and a
// warp reads 64 genes of a row as 256 contiguous bytes.  A CTA owns (64 genes, one study) and finds
// the medians by a 4-pass most-significant-digit radix SELECT over order-preserving uint32 keys:
// every pass sweeps the study's rows once and histograms the next 8-bit digit of the elements that
// still share the prefix of the wanted order statistic -- into a [256 bins][64 genes] shared-memory
// table whose columns are private to a lane (two-way bank conflicts at most; only warps collide on an address).
// Two order statistics are tracked (ranks (n-1)/2 and n/2; their mean is the median), each with its own
// table.  Expression columns are runs of equal values (mostly zeros), so every lane keeps a (digit, count)
// run in registers and only touches the table when the digit changes.  The first sweep also
// accumulates sum and sum of squares in float64 for the variance (population variance, ddof = 0, as
// np.var).  HBM-bound by design: 4 sweeps x 4 B per element.
#include "common.cuh"

namespace mn {

constexpr int GS_WARPS = 32;
constexpr int GS_UNROLL = 4;          // row loads (of 8 bytes) in flight per lane
constexpr int GS_THREADS = GS_WARPS * 32;
constexpr int GS_GENES = 64;          // genes per CTA: two per lane, 256 contiguous bytes per row and warp

__device__ __forceinline__ float key_to_float(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// run-length cache in front of a shared-memory histogram column
struct RunCache {
    uint32_t d, c;
    __device__ __forceinline__ void init() { d = 0xffffffffu; c = 0u; }
    __device__ __forceinline__ void add(uint32_t* col, uint32_t digit) {
        if (digit == d) {
            ++c;
        } else {
            if (c) atomicAdd(col + d * GS_GENES, c);
            d = digit;
            c = 1u;
        }
    }
    __device__ __forceinline__ void flush(uint32_t* col) {
        if (c) atomicAdd(col + d * GS_GENES, c);
    }
};

__global__ void __launch_bounds__(GS_THREADS, 1)
gene_stats_kernel(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ row_index,
                  const int32_t* __restrict__ seg_off, int G, double* __restrict__ median,
                  double* __restrict__ variance) {
    extern __shared__ __align__(16) unsigned char gs_raw[];
    uint32_t* hist = reinterpret_cast<uint32_t*>(gs_raw);                          // [2 tracks][256 bins][64 genes]
    double* s_sum = reinterpret_cast<double*>(hist + 2 * 256 * GS_GENES);           // [GS_WARPS][64]
    double* s_sq = s_sum + GS_WARPS * GS_GENES;                                     // [GS_WARPS][64]
    uint32_t* s_prefix = reinterpret_cast<uint32_t*>(s_sq + GS_WARPS * GS_GENES);   // [2][64] selected high digits
    uint32_t* s_rank = s_prefix + 2 * GS_GENES;                                     // [2][64] rank inside the prefix

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = 2 * lane;                                   // this lane's two genes inside the CTA's 64
    const int g0 = blockIdx.x * GS_GENES + gl;
    const int seg = blockIdx.y;
    const int r0 = seg_off[seg], r1 = seg_off[seg + 1];
    const int n = r1 - r0;
    const bool live0 = g0 < G, live1 = g0 + 1 < G;
    // 8-byte loads need an even row pitch and an 8-byte aligned base
    const bool vec2 = live1 && ((ldx & 1) == 0) && ((reinterpret_cast<uintptr_t>(X) & 7) == 0);
    if (n <= 0) {
        if (warp == 0) {
            const double qnan = __longlong_as_double(0x7ff8000000000000ll);
            if (live0) median[(int64_t)seg * G + g0] = variance[(int64_t)seg * G + g0] = qnan;
            if (live1) median[(int64_t)seg * G + g0 + 1] = variance[(int64_t)seg * G + g0 + 1] = qnan;
        }
        return;
    }
    if (threadIdx.x < 2 * GS_GENES) {
        s_prefix[threadIdx.x] = 0u;
        s_rank[threadIdx.x] = (threadIdx.x < GS_GENES) ? (uint32_t)((n - 1) / 2) : (uint32_t)(n / 2);
    }
    double sum0 = 0.0, sq0 = 0.0, sum1 = 0.0, sq1 = 0.0;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = threadIdx.x; i < 2 * 256 * GS_GENES; i += GS_THREADS) hist[i] = 0u;
        __syncthreads();
        // [track][gene of the lane]
        const uint32_t preA0 = s_prefix[gl], preA1 = s_prefix[gl + 1];
        const uint32_t preB0 = s_prefix[GS_GENES + gl], preB1 = s_prefix[GS_GENES + gl + 1];
        const bool same0 = (preA0 == preB0), same1 = (preA1 == preB1);
        uint32_t* hA = hist + gl;                              // track A column of gene 0 (gene 1: + 1)
        uint32_t* hB = hist + 256 * GS_GENES + gl;
        RunCache cA0, cA1, cB0, cB1;
        cA0.init(); cA1.init(); cB0.init(); cB1.init();
        for (int rb = r0 + warp; rb < r1; rb += GS_WARPS * GS_UNROLL) {
            float2 v[GS_UNROLL];
            bool ok[GS_UNROLL];
#pragma unroll
            for (int u = 0; u < GS_UNROLL; ++u) {
                const int r = rb + u * GS_WARPS;
                ok[u] = live0 && r < r1;
                v[u] = make_float2(0.f, 0.f);
                if (ok[u]) {
                    const int64_t src = row_index ? (int64_t)row_index[r] : (int64_t)r;
                    const float* px = X + src * ldx + g0;
                    if (vec2) {
                        v[u] = __ldg(reinterpret_cast<const float2*>(px));
                    } else {
                        v[u].x = __ldg(px);
                        if (live1) v[u].y = __ldg(px + 1);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < GS_UNROLL; ++u) {
                if (!ok[u]) continue;
                const uint32_t k0 = float_key(v[u].x), k1 = float_key(v[u].y);
                if (pass == 0) {
                    sum0 += (double)v[u].x;
                    sq0 += (double)v[u].x * (double)v[u].x;
                    sum1 += (double)v[u].y;
                    sq1 += (double)v[u].y * (double)v[u].y;
                }
                const uint32_t hi0 = (pass == 0) ? 0u : (k0 >> (shift + 8)), d0 = (k0 >> shift) & 0xffu;
                const uint32_t hi1 = (pass == 0) ? 0u : (k1 >> (shift + 8)), d1 = (k1 >> shift) & 0xffu;
                if (hi0 == preA0) cA0.add(hA, d0);
                if (!same0 && hi0 == preB0) cB0.add(hB, d0);
                if (live1) {
                    if (hi1 == preA1) cA1.add(hA + 1, d1);
                    if (!same1 && hi1 == preB1) cB1.add(hB + 1, d1);
                }
            }
        }
        cA0.flush(hA); cB0.flush(hB); cA1.flush(hA + 1); cB1.flush(hB + 1);
        __syncthreads();
        // per gene and tracked order statistic: the bin that holds the wanted rank (128 threads: 2 tracks x 64 genes)
        if (threadIdx.x < 2 * GS_GENES) {
            const int t = threadIdx.x / GS_GENES, gg = threadIdx.x % GS_GENES;
            const bool same = s_prefix[gg] == s_prefix[GS_GENES + gg];
            const uint32_t* h = hist + ((t == 1 && pass > 0 && !same) ? 256 * GS_GENES : 0) + gg;
            const uint32_t want = s_rank[t * GS_GENES + gg];
            uint32_t cum = 0u;
            int bin = 255;
            for (int b = 0; b < 256; ++b) {
                const uint32_t c = h[b * GS_GENES];
                if (want < cum + c) {
                    bin = b;
                    break;
                }
                cum += c;
            }
            // (all 128 threads read the old prefixes above before any of them writes: same warp pair -> barrier)
            __syncwarp();
            asm volatile("bar.sync 2, 128;" ::: "memory");
            s_rank[t * GS_GENES + gg] = want - cum;
            s_prefix[t * GS_GENES + gg] = (s_prefix[t * GS_GENES + gg] << 8) | (uint32_t)bin;
        }
        __syncthreads();
    }
    s_sum[warp * GS_GENES + gl] = sum0;
    s_sum[warp * GS_GENES + gl + 1] = sum1;
    s_sq[warp * GS_GENES + gl] = sq0;
    s_sq[warp * GS_GENES + gl + 1] = sq1;
    __syncthreads();
    if (threadIdx.x < GS_GENES) {
        const int gg = threadIdx.x, g = blockIdx.x * GS_GENES + gg;
        if (g < G) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < GS_WARPS; ++w) {
                a += s_sum[w * GS_GENES + gg];
                b += s_sq[w * GS_GENES + gg];
            }
            const double mean = a / (double)n;
            double var = b / (double)n - mean * mean;
            if (var < 0.0) var = 0.0;
            const double lo = (double)key_to_float(s_prefix[gg]), hi = (double)key_to_float(s_prefix[GS_GENES + gg]);
            median[(int64_t)seg * G + g] = 0.5 * (lo + hi);
            variance[(int64_t)seg * G + g] = var;
        }
    }
}

}  // namespace mn

extern "C" int mn_gene_stats(const float* X, int64_t ldx, const int32_t* row_index, const int32_t* seg_off,
                             int32_t n_seg, int32_t G, double* median, double* variance, mn_stream_t stream) {
    using namespace mn;
    MN_CHECK_ARG(X && seg_off && median && variance && n_seg >= 1 && G >= 1 && ldx >= G, "mn_gene_stats: bad arguments");
    const size_t smem = (size_t)2 * 256 * GS_GENES * 4 + (size_t)2 * GS_WARPS * GS_GENES * 8 + 4 * GS_GENES * 4;
    MN_CUDA(cudaFuncSetAttribute(gene_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((G + GS_GENES - 1) / GS_GENES), (unsigned)n_seg);
    gene_stats_kernel<<<grid, GS_THREADS, smem, (cudaStream_t)stream>>>(X, ldx, row_index, seg_off, G, median, variance);
    MN_LAUNCH_CHECK();
    return 0;
}
