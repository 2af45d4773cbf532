// Per-study, per-gene median and variance of the expression matrix: the statistics behind
// variableGenes (reference pymn/variableGenes.py:22-27 -- bottleneck.median(X, axis=0) and
// np.var(X, axis=0) over the cells of one study; SURVEY.md section 8f-2, the step right before the hot path).
//
// X is cells x genes, row-major, so a gene is a strided column: one lane owns one gene and a warp
// reads 32 neighbouring genes of a row as one 128-byte line.  A CTA owns (32 genes, one study) and finds
// the medians by a 4-pass most-significant-digit radix SELECT over order-preserving uint32 keys:
// every pass sweeps the study's rows once and histograms the next 8-bit digit of the elements that
// still share the prefix of the wanted order statistic -- into a [256 bins][32 genes] shared-memory
// table, so lane g only ever touches bank g: no bank conflicts, and only warps collide on an address.
// Two order statistics are tracked (ranks (n-1)/2 and n/2; their mean is the median), each with its own
// table.  Expression columns are runs of equal values (mostly zeros), so every lane keeps a (digit, count)
// run in registers and only touches the table when the digit changes.  The first sweep also
// accumulates sum and sum of squares in float64 for the variance (population variance, ddof = 0, as
// np.var).  HBM-bound by design: 4 sweeps x 4 B per element.
#include "common.cuh"

namespace mn {

constexpr int GS_WARPS = 16;
constexpr int GS_UNROLL = 8;          // row loads in flight per lane
constexpr int GS_THREADS = GS_WARPS * 32;

__device__ __forceinline__ float key_to_float(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

__global__ void __launch_bounds__(GS_THREADS)
gene_stats_kernel(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ row_index,
                  const int32_t* __restrict__ seg_off, int G, double* __restrict__ median,
                  double* __restrict__ variance) {
    extern __shared__ __align__(16) unsigned char gs_raw[];
    uint32_t* hist = reinterpret_cast<uint32_t*>(gs_raw);                     // [2][256][32]
    double* s_sum = reinterpret_cast<double*>(hist + 2 * 256 * 32);            // [GS_WARPS][32]
    double* s_sq = s_sum + GS_WARPS * 32;                                      // [GS_WARPS][32]
    uint32_t* s_prefix = reinterpret_cast<uint32_t*>(s_sq + GS_WARPS * 32);    // [2][32] selected high digits
    uint32_t* s_rank = s_prefix + 64;                                          // [2][32] rank inside the prefix

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * 32 + lane;
    const int seg = blockIdx.y;
    const int r0 = seg_off[seg], r1 = seg_off[seg + 1];
    const int n = r1 - r0;
    const bool live = g < G;
    if (n <= 0) {
        if (warp == 0 && live) {
            const double qnan = __longlong_as_double(0x7ff8000000000000ll);
            median[(int64_t)seg * G + g] = qnan;
            variance[(int64_t)seg * G + g] = qnan;
        }
        return;
    }
    if (threadIdx.x < 64) {
        s_prefix[threadIdx.x] = 0u;
        s_rank[threadIdx.x] = (threadIdx.x < 32) ? (uint32_t)((n - 1) / 2) : (uint32_t)(n / 2);
    }
    double sum = 0.0, sq = 0.0;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = threadIdx.x; i < 2 * 256 * 32; i += GS_THREADS) hist[i] = 0u;
        __syncthreads();
        const uint32_t pre0 = s_prefix[lane], pre1 = s_prefix[32 + lane];
        const bool same = (pre0 == pre1);
        // rows of the study, GS_WARPS apart, GS_UNROLL loads in flight per lane
        uint32_t run_d0 = 0xffffffffu, run_c0 = 0u, run_d1 = 0xffffffffu, run_c1 = 0u;
        for (int rb = r0 + warp; rb < r1; rb += GS_WARPS * GS_UNROLL) {
            float v[GS_UNROLL];
            bool ok[GS_UNROLL];
#pragma unroll
            for (int u = 0; u < GS_UNROLL; ++u) {
                const int r = rb + u * GS_WARPS;
                ok[u] = live && r < r1;
                v[u] = 0.f;
                if (ok[u]) {
                    const int64_t src = row_index ? (int64_t)row_index[r] : (int64_t)r;
                    v[u] = __ldg(X + src * ldx + g);
                }
            }
#pragma unroll
            for (int u = 0; u < GS_UNROLL; ++u) {
                if (!ok[u]) continue;
                const uint32_t k = float_key(v[u]);
                if (pass == 0) {
                    sum += (double)v[u];
                    sq += (double)v[u] * (double)v[u];
                }
                const uint32_t hi = (pass == 0) ? 0u : (k >> (shift + 8));
                const uint32_t d = (k >> shift) & 0xffu;
                if (hi == pre0) {
                    if (d == run_d0) {
                        ++run_c0;
                    } else {
                        if (run_c0) atomicAdd(&hist[run_d0 * 32 + lane], run_c0);
                        run_d0 = d;
                        run_c0 = 1u;
                    }
                }
                if (!same && hi == pre1) {
                    if (d == run_d1) {
                        ++run_c1;
                    } else {
                        if (run_c1) atomicAdd(&hist[(256 + run_d1) * 32 + lane], run_c1);
                        run_d1 = d;
                        run_c1 = 1u;
                    }
                }
            }
        }
        if (run_c0) atomicAdd(&hist[run_d0 * 32 + lane], run_c0);
        if (run_c1) atomicAdd(&hist[(256 + run_d1) * 32 + lane], run_c1);
        __syncthreads();
        // per gene and tracked order statistic: the bin that holds the wanted rank
        if (warp < 2) {
            const int t = warp;                                    // warp 0: lower statistic, warp 1: upper
            const uint32_t* h = hist + ((t == 1 && pass > 0 && !same) ? 256 * 32 : 0);
            uint32_t want = s_rank[t * 32 + lane], cum = 0u;
            int bin = 255;
            for (int b = 0; b < 256; ++b) {
                const uint32_t c = h[b * 32 + lane];
                if (want < cum + c) {
                    bin = b;
                    break;
                }
                cum += c;
            }
            __syncwarp();
            s_rank[t * 32 + lane] = want - cum;
            s_prefix[t * 32 + lane] = (s_prefix[t * 32 + lane] << 8) | (uint32_t)bin;
        }
        __syncthreads();
    }
    s_sum[warp * 32 + lane] = sum;
    s_sq[warp * 32 + lane] = sq;
    __syncthreads();
    if (warp == 0 && live) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int w = 0; w < GS_WARPS; ++w) {
            a += s_sum[w * 32 + lane];
            b += s_sq[w * 32 + lane];
        }
        const double mean = a / (double)n;
        double var = b / (double)n - mean * mean;
        if (var < 0.0) var = 0.0;
        const double lo = (double)key_to_float(s_prefix[lane]), hi = (double)key_to_float(s_prefix[32 + lane]);
        median[(int64_t)seg * G + g] = 0.5 * (lo + hi);
        variance[(int64_t)seg * G + g] = var;
    }
}

}  // namespace mn

extern "C" int mn_gene_stats(const float* X, int64_t ldx, const int32_t* row_index, const int32_t* seg_off,
                             int32_t n_seg, int32_t G, double* median, double* variance, mn_stream_t stream) {
    using namespace mn;
    MN_CHECK_ARG(X && seg_off && median && variance && n_seg >= 1 && G >= 1 && ldx >= G, "mn_gene_stats: bad arguments");
    const size_t smem = (size_t)2 * 256 * 32 * 4 + (size_t)2 * GS_WARPS * 32 * 8 + 2 * 64 * 4;
    static bool configured = false;
    if (!configured) {
        MN_CUDA(cudaFuncSetAttribute(gene_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((unsigned)((G + 31) / 32), (unsigned)n_seg);
    gene_stats_kernel<<<grid, GS_THREADS, smem, (cudaStream_t)stream>>>(X, ldx, row_index, seg_off, G, median, variance);
    MN_LAUNCH_CHECK();
    return 0;
}
