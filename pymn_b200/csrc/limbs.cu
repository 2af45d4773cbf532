// Integer limb planes for the exact (int8 tensor-core) form of the centroid votes, mn_votes_fast_i8.
//
// votes = X_test.T @ centroids (reference pymn/MetaNeighborUS.py:361, MetaNeighbor.py:170-174) with
// X_test[:, i] = d_i / ||d_i||, d = centred doubled ranks (integers, |d| < G).  The FP32 accumulator of the
// fp16 tensor-core form truncates ~2e-6 of every dot product -- 3-4 rank flips per AUROC cell at 5,000 cells
// per study.  Splitting both operands into signed base-128 limbs makes every limb product an int8 MMA with an
// int32 accumulator: the dot product of the integer ranks with the centroid, quantised to 21 bits of its
// largest entry, is EXACT, independent of tile shape, k order and GPU count.
//   cell      d = 128 a1 + a0,              a0 in [-64, 63], |a1| <= 127   (G <= 16,000);  G <= 127: a1 = d, a0 = 0
//   centroid  b = rint(c * 2^s) = 128^3 b3 + 128^2 b2 + 128 b1 + b0,  b0 in [-64, 63], b1 .. b3 in [-64, 64]
// 28 bits of the centroid's largest entry: measured on the config-4-shaped fixture (7 x 5,000 cells, 700 clusters)
// 21 bits cost up to 3 rank flips per AUROC cell, 24 bits 2, 28 bits 1 (and 0.002 on average).
#include "common.cuh"

namespace mn {

constexpr int LS_WARPS = 8;

__global__ void __launch_bounds__(LS_WARPS * 32)
limb_split_cells_kernel(const __half* __restrict__ d_hi, const __half* __restrict__ d_lo, int64_t ldd,
                        const float* __restrict__ inv_norm, int64_t M, int G, int8_t* __restrict__ a1,
                        int8_t* __restrict__ a0, int64_t ldk, double* __restrict__ inv_norm64) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * LS_WARPS + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * LS_WARPS;
    const bool one_limb = G <= 127;          // |d| < G fits one digit: the products with a0 vanish
    for (int64_t row = warp0; row < M; row += n_warps) {
        const __half* ph = d_hi + row * ldd;
        const __half* pl = d_lo ? d_lo + row * ldd : nullptr;
        unsigned long long ss = 0ull;
        for (int g0 = lane * 4; g0 < (int)ldk; g0 += 128) {
            int d[4] = {0, 0, 0, 0};
            if (g0 < G) {                    // G <= ldd, both multiples of 4 apart from the tail handled below
                const uint2 h = *reinterpret_cast<const uint2*>(ph + g0);
                const __half2 h01 = *reinterpret_cast<const __half2*>(&h.x), h23 = *reinterpret_cast<const __half2*>(&h.y);
                d[0] = __half2int_rn(__low2half(h01));
                d[1] = __half2int_rn(__high2half(h01));
                d[2] = __half2int_rn(__low2half(h23));
                d[3] = __half2int_rn(__high2half(h23));
                if (pl) {
                    const uint2 l = *reinterpret_cast<const uint2*>(pl + g0);
                    const __half2 l01 = *reinterpret_cast<const __half2*>(&l.x), l23 = *reinterpret_cast<const __half2*>(&l.y);
                    d[0] += __half2int_rn(__low2half(l01));
                    d[1] += __half2int_rn(__high2half(l01));
                    d[2] += __half2int_rn(__low2half(l23));
                    d[3] += __half2int_rn(__high2half(l23));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (g0 + u >= G) d[u] = 0;
            }
            uint32_t w1 = 0u, w0 = 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int hi = one_limb ? d[u] : (d[u] + 64) >> 7;      // floor((d + 64) / 128)
                const int lo = one_limb ? 0 : d[u] - (hi << 7);         // [-64, 63]
                w1 |= ((uint32_t)hi & 0xffu) << (8 * u);
                w0 |= ((uint32_t)lo & 0xffu) << (8 * u);
                ss += (unsigned long long)((long long)d[u] * d[u]);
            }
            *reinterpret_cast<uint32_t*>(a1 + row * ldk + g0) = w1;
            *reinterpret_cast<uint32_t*>(a0 + row * ldk + g0) = w0;
        }
        ss = warp_sum(ss);
        if (lane == 0) inv_norm64[row] = (ss > 0ull && inv_norm[row] > 0.f) ? 1.0 / sqrt((double)ss) : 0.0;
    }
}

// one CTA per centroid row: power-of-two scale so that max|v| * scale is in [2^26, 2^27), then four limbs
__global__ void limb_split_centroids_kernel(const float* __restrict__ in, int64_t ld, int G, int8_t* __restrict__ limbs,
                                            int64_t ldk, int64_t rows_pad, float* __restrict__ inv_scale,
                                            int32_t* __restrict__ q32) {
    __shared__ float s_red[32];
    const int r = blockIdx.x;
    const float* x = in + (int64_t)r * ld;
    float m = 0.f;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const float a = fabsf(x[g]);
        if (a == a && a < __int_as_float(0x7f800000)) m = fmaxf(m, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = (threadIdx.x < ((blockDim.x + 31) >> 5)) ? s_red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (threadIdx.x == 0) s_red[0] = t;
    }
    __syncthreads();
    m = s_red[0];
    float sc = 1.f;
    if (m > 0.f) {
        int e;
        frexpf(m, &e);                 // m = f * 2^e, f in [0.5, 1)
        sc = ldexpf(1.f, 27 - e);      // m * sc in [2^26, 2^27)
    }
    if (threadIdx.x == 0) inv_scale[r] = 1.f / sc;   // exact: sc is a power of two
    int8_t* p3 = limbs ? limbs + ((int64_t)0 * rows_pad + r) * ldk : nullptr;
    int8_t* p2 = limbs ? limbs + ((int64_t)1 * rows_pad + r) * ldk : nullptr;
    int8_t* p1 = limbs ? limbs + ((int64_t)2 * rows_pad + r) * ldk : nullptr;
    int8_t* p0 = limbs ? limbs + ((int64_t)3 * rows_pad + r) * ldk : nullptr;
    for (int g = threadIdx.x; g < (int)ldk; g += blockDim.x) {
        int b = 0;
        if (g < G) {
            const float v = x[g] * sc;                        // exact (power-of-two scale); |v| < 2^27, 24-bit mantissa
            b = (v == v) ? __float2int_rn(fminf(fmaxf(v, -134217728.f), 134217728.f)) : 0;
        }
        const int b3 = (b + (1 << 20)) >> 21;                 // [-64, 64]
        const int r3 = b - (b3 << 21);                        // [-2^20, 2^20)
        const int b2 = (r3 + 8192) >> 14;                     // [-64, 64]
        const int rem = r3 - (b2 << 14);                      // [-8192, 8191]
        const int b1 = (rem + 64) >> 7;                       // [-64, 64]
        const int b0 = rem - (b1 << 7);                       // [-64, 63]
        if (q32) q32[(int64_t)r * ldk + g] = b;
        if (limbs) {
            p3[g] = (int8_t)b3;
            p2[g] = (int8_t)b2;
            p1[g] = (int8_t)b1;
            p0[g] = (int8_t)b0;
        }
    }
}

// Node degree in double: den[s, m] = inv_norm64[m] * inv_scale[s] * sum_g d[m, g] * q[s, g] + addend[s]
// (MetaNeighborUS.py:363-366), q = the study centroid quantised to 28 bits exactly as the limb planes are
// (limb_split_centroids_kernel's q32 plane).  One warp per cell; every lane owns 4 genes per step and SB running
// 64-bit integer sums (one IMAD.WIDE per product: exact, and no int->double conversions in the loop).
// S is a handful of studies: 2 S G integer FMAs per cell, hidden under the 2 G bytes of limbs the cell streams in.
template <int SB>
__global__ void __launch_bounds__(LS_WARPS * 32)
votes_den64_kernel(const int8_t* __restrict__ a1, const int8_t* __restrict__ a0, int64_t ldk,
                   const double* __restrict__ inv_norm64, int64_t M, int G, const int32_t* __restrict__ q32,
                   const float* __restrict__ inv_scale, const float* __restrict__ addend, int s0, int S,
                   double* __restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * LS_WARPS + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * LS_WARPS;
    const int ns = (S - s0) < SB ? (S - s0) : SB;
    const int w1mul = G <= 127 ? 1 : 128;
    for (int64_t row = warp0; row < M; row += n_warps) {
        long long acc[SB];
#pragma unroll
        for (int q = 0; q < SB; ++q) acc[q] = 0;
        for (int g0 = lane * 4; g0 < G; g0 += 128) {
            const uint32_t w1 = *reinterpret_cast<const uint32_t*>(a1 + row * ldk + g0);
            const uint32_t w0 = *reinterpret_cast<const uint32_t*>(a0 + row * ldk + g0);
            int d[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                d[u] = w1mul * (int)(int8_t)(w1 >> (8 * u)) + (int)(int8_t)(w0 >> (8 * u));   // 0 beyond G (padding)
#pragma unroll
            for (int q = 0; q < SB; ++q) {
                if (q < ns) {
                    const int4 c = __ldg(reinterpret_cast<const int4*>(q32 + (int64_t)(s0 + q) * ldk + g0));
                    acc[q] += (long long)d[0] * c.x;
                    acc[q] += (long long)d[1] * c.y;
                    acc[q] += (long long)d[2] * c.z;
                    acc[q] += (long long)d[3] * c.w;
                }
            }
        }
        const double w = inv_norm64[row];
#pragma unroll
        for (int q = 0; q < SB; ++q) {
            const long long t = warp_sum(acc[q]);
            if (lane == 0 && q < ns)
                out[(int64_t)(s0 + q) * ldo + row] =
                    fma((double)t * w, (double)inv_scale[s0 + q], addend ? (double)addend[s0 + q] : 0.0);
        }
    }
}

}  // namespace mn

extern "C" {

int mn_votes_den64(const int8_t* a1, const int8_t* a0, int64_t ldk, const double* inv_norm64, int64_t M, int32_t G,
                   const int32_t* q32, const float* inv_scale, const float* addend, int32_t S, double* out, int64_t ldo,
                   mn_stream_t stream) {
    MN_CHECK_ARG(a1 && a0 && inv_norm64 && q32 && inv_scale && out, "mn_votes_den64: null pointer");
    MN_CHECK_ARG(M >= 1 && G >= 1 && S >= 1 && ldo >= M && ldk % 128 == 0 && ldk >= G, "mn_votes_den64: bad shape");
    MN_CHECK_ARG((reinterpret_cast<uintptr_t>(q32) & 15) == 0, "mn_votes_den64: q32 must be 16-byte aligned");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t blocks = (M + mn::LS_WARPS - 1) / mn::LS_WARPS;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    cudaStream_t st = (cudaStream_t)stream;
    for (int s0 = 0; s0 < S; s0 += 8) {
        if (S - s0 <= 4)
            mn::votes_den64_kernel<4><<<(unsigned)blocks, mn::LS_WARPS * 32, 0, st>>>(a1, a0, ldk, inv_norm64, M, G, q32,
                                                                                 inv_scale, addend, s0, S, out, ldo);
        else
            mn::votes_den64_kernel<8><<<(unsigned)blocks, mn::LS_WARPS * 32, 0, st>>>(a1, a0, ldk, inv_norm64, M, G, q32,
                                                                                 inv_scale, addend, s0, S, out, ldo);
        MN_LAUNCH_CHECK();
    }
    return 0;
}

int mn_limb_split_cells(const mn_half_t* d_hi, const mn_half_t* d_lo, int64_t ldd, const float* inv_norm, int64_t M,
                        int32_t G, int8_t* a1, int8_t* a0, int64_t ldk, double* inv_norm64, mn_stream_t stream) {
    MN_CHECK_ARG(d_hi && inv_norm && a1 && a0 && inv_norm64, "mn_limb_split_cells: null pointer");
    MN_CHECK_ARG(M >= 1 && G >= 1 && G <= 16000 && ldd >= G && ldd % 4 == 0, "mn_limb_split_cells: bad shape (G <= 16000)");
    MN_CHECK_ARG(G <= 2048 || d_lo, "mn_limb_split_cells: G=%d > 2048 needs the d_lo plane", G);
    MN_CHECK_ARG(ldk % 128 == 0 && ldk >= G, "mn_limb_split_cells: ldk must be a multiple of 128 >= G");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t blocks = (M + mn::LS_WARPS - 1) / mn::LS_WARPS;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    mn::limb_split_cells_kernel<<<(unsigned)blocks, mn::LS_WARPS * 32, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __half*>(d_hi), reinterpret_cast<const __half*>(d_lo), ldd, inv_norm, M, G, a1, a0, ldk,
        inv_norm64);
    MN_LAUNCH_CHECK();
    return 0;
}

int mn_limb_split_centroids(const float* in, int64_t ld, int32_t R, int32_t G, int8_t* limbs, int64_t ldk,
                            int64_t rows_pad, float* inv_scale, int32_t* q32, mn_stream_t stream) {
    MN_CHECK_ARG(in && (limbs || q32) && inv_scale && R >= 1 && G >= 1 && ld >= G, "mn_limb_split_centroids: bad arguments");
    MN_CHECK_ARG(ldk % 128 == 0 && ldk >= G && rows_pad >= R, "mn_limb_split_centroids: bad plane shape");
    mn::limb_split_centroids_kernel<<<(unsigned)R, 256, 0, (cudaStream_t)stream>>>(in, ld, G, limbs, ldk, rows_pad, inv_scale, q32);
    MN_LAUNCH_CHECK();
    return 0;
}

}
