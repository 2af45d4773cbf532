// K2: cluster centroids, study centroids, fp16 operand split, vote finalisation.
//
// Replaces (reference pymn/):  X_norm @ onehot  (MetaNeighborUS.py:284,
// trainModel.py:43, MetaNeighbor.py:170), cluster_centroids @ study_matrix
// (MetaNeighborUS.py:349-350), the node-degree division of the full-network
// path (MetaNeighborUS.py:167-169) and the leave-one-study-out label matrix of
// the supervised paths (MetaNeighbor.py:198-217, :134-147).
// All reductions run in a fixed order (no float atomics): results are
// deterministic and independent of the launch geometry.
#include "common.cuh"

namespace mn {

// grid (ceil(ldc/2/T), L); each thread owns two adjacent genes, walks the
// label's rows in order, accumulates in double.
__global__ void centroid_kernel(const __half* __restrict__ d_hi, const __half* __restrict__ d_lo, int64_t ldd,
                                const float* __restrict__ inv_norm, const double* __restrict__ inv_norm64,
                                const int32_t* __restrict__ row_off, int G,
                                float* __restrict__ out, int64_t ldc, float* __restrict__ n_valid) {
    const int l = blockIdx.y;
    const int g2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = 2 * g2;
    const int r0 = row_off[l], r1 = row_off[l + 1];
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        int cnt = 0;
        for (int r = r0 + (int)threadIdx.x; r < r1; r += 32) cnt += (inv_norm[r] > 0.f) ? 1 : 0;
        cnt = warp_sum(cnt);
        if (threadIdx.x == 0) n_valid[l] = (float)cnt;
    }
    if (g >= G) return;
    double a0 = 0.0, a1 = 0.0;
    const __half2* ph = reinterpret_cast<const __half2*>(d_hi) + g2;
    const __half2* pl = d_lo ? reinterpret_cast<const __half2*>(d_lo) + g2 : nullptr;
    const int64_t stride2 = ldd >> 1;
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
        const float w = __ldg(inv_norm + r);
        float2 v = __half22float2(ph[(int64_t)r * stride2]);
        if (pl) {
            const float2 u = __half22float2(pl[(int64_t)r * stride2]);
            v.x += u.x;   // exact: |d| < 2^15
            v.y += u.y;
        }
        if (inv_norm64) {          // exact integers times the double-precision scale: only the final store rounds
            const double w64 = inv_norm64[r];
            a0 = fma((double)v.x, w64, a0);
            a1 = fma((double)v.y, w64, a1);
        } else {
            a0 += (double)(v.x * w);   // one fp32 rounding per term
            a1 += (double)(v.y * w);
        }
    }
    out[(int64_t)l * ldc + g] = (float)a0;
    if (g + 1 < G) out[(int64_t)l * ldc + g + 1] = (float)a1;
}

// ---- fixed-point centroids: exact, order-independent sums ----------------------------------------------
// Every term x*_ig = d_ig / ||d_i|| (|x*| <= 1) is rounded ONCE to a multiple of 2^-38 and the terms are summed
// as 64-bit integers: the sum does not depend on the order of the cells, on the launch geometry, or on how
// the cells are dealt to GPUs (partial sums are added with an integer all-reduce), so every world size sees
// bit-identical centroids.  2^-38 per term is ~1e-11 of a centroid entry, four orders below the float32
// rounding of the reference's own product; a label may hold up to 2^24 cells before the sum could overflow.
// grid (ceil(ldc / 2 / T), L, Z): CTA (x, l, z) walks rows r0 + z, r0 + z + Z, ... of label l.
__global__ void centroid_fixed_kernel(const __half* __restrict__ d_hi, const __half* __restrict__ d_lo, int64_t ldd,
                                      const double* __restrict__ inv_norm64, const int32_t* __restrict__ row_off,
                                      int G, unsigned long long* __restrict__ out, int64_t ldc,
                                      int32_t* __restrict__ n_valid) {
    const int l = blockIdx.y;
    const int g2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = 2 * g2;
    const int Z = gridDim.z;
    const int r0 = row_off[l] + (int)blockIdx.z, r1 = row_off[l + 1];
    if (r0 >= r1) return;
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        int cnt = 0;
        for (int r = r0 + (int)threadIdx.x * Z; r < r1; r += 32 * Z) cnt += (inv_norm64[r] > 0.0) ? 1 : 0;
        cnt = warp_sum(cnt);
        if (threadIdx.x == 0 && cnt) atomicAdd(n_valid + l, cnt);
    }
    if (g >= G) return;
    long long a0 = 0, a1 = 0;
    const __half2* ph = reinterpret_cast<const __half2*>(d_hi) + g2;
    const __half2* pl = d_lo ? reinterpret_cast<const __half2*>(d_lo) + g2 : nullptr;
    const int64_t stride2 = ldd >> 1;
#pragma unroll 4
    for (int r = r0; r < r1; r += Z) {
        const double w = __ldg(inv_norm64 + r) * MN_CQ_ONE;       // exact: a power-of-two scale
        float2 v = __half22float2(ph[(int64_t)r * stride2]);
        if (pl) {
            const float2 u = __half22float2(pl[(int64_t)r * stride2]);
            v.x += u.x;   // exact: |d| < 2^15
            v.y += u.y;
        }
        a0 += __double2ll_rn((double)v.x * w);
        a1 += __double2ll_rn((double)v.y * w);
    }
    if (a0) atomicAdd(out + (int64_t)l * ldc + g, (unsigned long long)a0);
    if (g + 1 < G && a1) atomicAdd(out + (int64_t)l * ldc + g + 1, (unsigned long long)a1);
}

// the same sums from the int8 limb planes (d = 128 a1 + a0; G <= 127: d = a1): four genes per thread
__global__ void centroid_fixed_i8_kernel(const int8_t* __restrict__ a1, const int8_t* __restrict__ a0, int64_t ldk,
                                         const double* __restrict__ inv_norm64, const int32_t* __restrict__ row_off,
                                         int G, unsigned long long* __restrict__ out, int64_t ldc,
                                         int32_t* __restrict__ n_valid) {
    const int l = blockIdx.y;
    const int g4 = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = 4 * g4;
    const int Z = gridDim.z;
    const int r0 = row_off[l] + (int)blockIdx.z, r1 = row_off[l + 1];
    if (r0 >= r1) return;
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        int cnt = 0;
        for (int r = r0 + (int)threadIdx.x * Z; r < r1; r += 32 * Z) cnt += (inv_norm64[r] > 0.0) ? 1 : 0;
        cnt = warp_sum(cnt);
        if (threadIdx.x == 0 && cnt) atomicAdd(n_valid + l, cnt);
    }
    if (g >= G) return;
    long long acc[4] = {0, 0, 0, 0};
    const int hs = G <= 127 ? 1 : 128;          // one digit per rank up to 127 genes: a1 = d, a0 = 0
    const char4* p1 = reinterpret_cast<const char4*>(a1) + g4;
    const char4* p0 = reinterpret_cast<const char4*>(a0) + g4;
    const int64_t stride4 = ldk >> 2;
#pragma unroll 4
    for (int r = r0; r < r1; r += Z) {
        const double w = __ldg(inv_norm64 + r) * MN_CQ_ONE;       // exact: a power-of-two scale
        const char4 h = p1[(int64_t)r * stride4], q = p0[(int64_t)r * stride4];
        acc[0] += __double2ll_rn((double)(hs * (int)h.x + (int)q.x) * w);
        acc[1] += __double2ll_rn((double)(hs * (int)h.y + (int)q.y) * w);
        acc[2] += __double2ll_rn((double)(hs * (int)h.z + (int)q.z) * w);
        acc[3] += __double2ll_rn((double)(hs * (int)h.w + (int)q.w) * w);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (g + u < G && acc[u]) atomicAdd(out + (int64_t)l * ldc + g + u, (unsigned long long)acc[u]);
}

__global__ void group_sum_fixed_kernel(const long long* __restrict__ in, int64_t ld, int R, int width,
                                       const int32_t* __restrict__ group, long long* __restrict__ out,
                                       const int32_t* __restrict__ n_in, int32_t* __restrict__ n_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int gq = blockIdx.y;
    if (c == 0 && n_in) {
        int t = 0;
        for (int r = 0; r < R; ++r)
            if (group[r] == gq) t += n_in[r];
        n_out[gq] = t;
    }
    if (c >= width) return;
    long long acc = 0;
    for (int r = 0; r < R; ++r)
        if (group[r] == gq) acc += in[(int64_t)r * ld + c];
    out[(int64_t)gq * ld + c] = acc;
}

__global__ void centroid_finish_kernel(const long long* __restrict__ cq, const int32_t* __restrict__ nq, int64_t n_elem,
                                       int R, float* __restrict__ C, float* __restrict__ n_valid) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_elem) C[i] = (float)((double)cq[i] * (1.0 / MN_CQ_ONE));
    if (i < R) n_valid[i] = (float)nq[i];
}

__global__ void group_sum_kernel(const float* __restrict__ in, int64_t ld, int R, int width,
                                 const int32_t* __restrict__ group, int n_groups, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int gq = blockIdx.y;
    if (c >= width) return;
    double acc = 0.0;
    for (int r = 0; r < R; ++r)
        if (group[r] == gq) acc += (double)in[(int64_t)r * ld + c];
    out[(int64_t)gq * ld + c] = (float)acc;
}

// one CTA per row: power-of-two scale so that max|v|*scale is in [2^13, 2^14), then hi/lo fp16.
__global__ void split_f16_kernel(const float* __restrict__ in, int64_t ld, int G, __half* __restrict__ hi,
                                 __half* __restrict__ lo, int64_t ldh, float* __restrict__ scale) {
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
        sc = ldexpf(1.f, 14 - e);      // m*sc in [2^13, 2^14)
    }
    if (threadIdx.x == 0) scale[r] = 1.f / sc;   // exact: sc is a power of two
    __half* ph = hi + (int64_t)r * ldh;
    __half* pl = lo + (int64_t)r * ldh;
    for (int g = threadIdx.x; g < (int)ldh; g += blockDim.x) {
        float v = (g < G) ? x[g] * sc : 0.f;
        const __half h = __float2half_rn(v);
        ph[g] = h;
        pl[g] = __float2half_rn(v - __half2float(h));
    }
}

// votes u64 [N, L] (+ degree) -> float column-major [L, ldo]
__global__ void votes_finalize_kernel(const unsigned long long* __restrict__ votes,
                                      const unsigned long long* __restrict__ degree,
                                      const int32_t* __restrict__ label, int64_t N, int L, int ndn,
                                      float* __restrict__ out, int64_t ldo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const unsigned long long one = MN_W_ONE;
    const int li = label[i];
    unsigned long long dsum = 0ull;
    if (degree) {
        dsum = degree[i];
    } else {                                  // node degree = sum of the cell's votes over all labels
        for (int l = 0; l < L; ++l) dsum += votes[i * L + l];
    }
    const double deg = (double)(dsum + one);
    const double inv_one = 1.0 / (double)MN_W_ONE;
    for (int l = 0; l < L; ++l) {
        unsigned long long v = votes[i * L + l];
        if (l == li) v += one;
        const double x = ndn ? (double)v / deg : (double)v * inv_one;
        out[(int64_t)l * ldo + i] = (float)x;
    }
}

// test labels of the centroid paths: dropped cells (inv_norm == 0) leave the ranking; *poison is raised when
// some cell is constant but NON-ZERO (flag bit 0 without bit 1), see include/pymn_b200.h
__global__ void mask_dropped_kernel(const int32_t* __restrict__ cell_label, const float* __restrict__ inv_norm,
                                    const uint8_t* __restrict__ flags, int64_t N, int32_t* __restrict__ out_label,
                                    uint8_t* __restrict__ poison) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    out_label[i] = inv_norm[i] > 0.f ? cell_label[i] : -1;
    if (poison && flags && (flags[i] & 3) == 1) *poison = 1;
}

// leave-one-study-out combination (see include/pymn_b200.h)
__global__ void loso_finalize_kernel(const void* __restrict__ in, int64_t ldi, int is_fixed,
                                     const unsigned long long* __restrict__ degree,
                                     const float* __restrict__ n_joint, const int32_t* __restrict__ cell_study,
                                     int64_t N, int S, int K, int ndn, float* __restrict__ out, int64_t ldo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int si = cell_study[i];
    if (is_fixed) {
        const unsigned long long* v = reinterpret_cast<const unsigned long long*>(in) + i * (int64_t)(S * K);
        const double deg = (double)(degree[i] + MN_W_ONE);
        const double inv_one = 1.0 / (double)MN_W_ONE;
        for (int k = 0; k < K; ++k) {
            unsigned long long acc = 0;
            for (int s = 0; s < S; ++s)
                if (s != si) acc += v[s * K + k];
            const double x = ndn ? (double)acc / deg : (double)acc * inv_one;
            out[(int64_t)k * ldo + i] = (float)x;
        }
    } else {
        const float* v = reinterpret_cast<const float*>(in);
        double tot = 0.0, ntot = 0.0;
        for (int k = 0; k < K; ++k) {
            double acc = 0.0, nk = 0.0;
            for (int s = 0; s < S; ++s)
                if (s != si) {
                    acc += (double)v[(int64_t)(s * K + k) * ldi + i];
                    nk += (double)n_joint[s * K + k];
                }
            tot += acc;
            ntot += nk;
            out[(int64_t)k * ldo + i] = (float)(ndn ? acc + nk : acc);
        }
        if (ndn) {
            const double norm = tot + ntot;
            for (int k = 0; k < K; ++k)
                out[(int64_t)k * ldo + i] = (float)((double)out[(int64_t)k * ldo + i] / norm);
        }
    }
}

}  // namespace mn

extern "C" {

int mn_centroids(const mn_half_t* d_hi, const mn_half_t* d_lo, int64_t ldd, const float* inv_norm,
                 const double* inv_norm64, const int32_t* row_off, int32_t L, int32_t G, float* centroids, int64_t ldc, float* n_valid,
                 mn_stream_t stream) {
    MN_CHECK_ARG(d_hi && inv_norm && row_off && centroids && n_valid, "mn_centroids: null pointer");
    MN_CHECK_ARG(L >= 1 && G >= 1 && ldc >= G && (ldd % 2) == 0, "mn_centroids: bad shape");
    const int T = 128;
    dim3 grid((unsigned)((G + 2 * T - 1) / (2 * T)), (unsigned)L);
    mn::centroid_kernel<<<grid, T, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __half*>(d_hi), reinterpret_cast<const __half*>(d_lo), ldd, inv_norm, inv_norm64, row_off,
        G, centroids, ldc, n_valid);
    MN_LAUNCH_CHECK();
    return 0;
}

int mn_centroids_fixed(const mn_half_t* d_hi, const mn_half_t* d_lo, int64_t ldd, const double* inv_norm64,
                       const int32_t* row_off, int32_t L, int32_t G, int64_t max_label_rows, int64_t* cq, int64_t ldc,
                       int32_t* nq, mn_stream_t stream) {
    MN_CHECK_ARG(d_hi && inv_norm64 && row_off && cq && nq, "mn_centroids_fixed: null pointer");
    MN_CHECK_ARG(L >= 1 && G >= 1 && ldc >= G && (ldd % 2) == 0, "mn_centroids_fixed: bad shape");
    const int T = 128;
    const unsigned gx = (unsigned)((G + 2 * T - 1) / (2 * T));
    // row slices per label: enough CTAs to fill the machine, at least ~32 rows each
    int64_t Z = (148 * 16 + (int64_t)gx * L - 1) / ((int64_t)gx * L);
    const int64_t zmax = max_label_rows > 0 ? (max_label_rows + 31) / 32 : 1;
    if (Z > zmax) Z = zmax;
    if (Z > 1024) Z = 1024;
    if (Z < 1) Z = 1;
    dim3 grid(gx, (unsigned)L, (unsigned)Z);
    mn::centroid_fixed_kernel<<<grid, T, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __half*>(d_hi), reinterpret_cast<const __half*>(d_lo), ldd, inv_norm64, row_off, G,
        reinterpret_cast<unsigned long long*>(cq), ldc, nq);
    MN_LAUNCH_CHECK();
    return 0;
}

int mn_centroids_fixed_i8(const int8_t* a1, const int8_t* a0, int64_t ldk, const double* inv_norm64,
                          const int32_t* row_off, int32_t L, int32_t G, int64_t max_label_rows, int64_t* cq, int64_t ldc,
                          int32_t* nq, mn_stream_t stream) {
    MN_CHECK_ARG(a1 && a0 && inv_norm64 && row_off && cq && nq, "mn_centroids_fixed_i8: null pointer");
    MN_CHECK_ARG(L >= 1 && G >= 1 && ldc >= G && ldk >= G && (ldk % 4) == 0, "mn_centroids_fixed_i8: bad shape");
    const int T = 128;
    const unsigned gx = (unsigned)((G + 4 * T - 1) / (4 * T));
    int64_t Z = (148 * 16 + (int64_t)gx * L - 1) / ((int64_t)gx * L);
    const int64_t zmax = max_label_rows > 0 ? (max_label_rows + 31) / 32 : 1;
    if (Z > zmax) Z = zmax;
    if (Z > 1024) Z = 1024;
    if (Z < 1) Z = 1;
    dim3 grid(gx, (unsigned)L, (unsigned)Z);
    mn::centroid_fixed_i8_kernel<<<grid, T, 0, (cudaStream_t)stream>>>(a1, a0, ldk, inv_norm64, row_off, G,
                                                                      reinterpret_cast<unsigned long long*>(cq), ldc, nq);
    MN_LAUNCH_CHECK();
    return 0;
}

int mn_group_sum_fixed(const int64_t* in, int64_t ld, int32_t R, int32_t width, const int32_t* group, int32_t n_groups,
                       int64_t* out, const int32_t* n_in, int32_t* n_out, mn_stream_t stream) {
    MN_CHECK_ARG(in && group && out && R >= 1 && n_groups >= 1 && width >= 1 && (!n_in || n_out),
                 "mn_group_sum_fixed: bad arguments");
    dim3 grid((unsigned)((width + 127) / 128), (unsigned)n_groups);
    mn::group_sum_fixed_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const long long*>(in), ld, R, width, group, reinterpret_cast<long long*>(out), n_in, n_out);
    MN_LAUNCH_CHECK();
    return 0;
}

int mn_centroids_finish(const int64_t* cq, const int32_t* nq, int32_t R, int64_t ld, float* centroids, float* n_valid,
                        mn_stream_t stream) {
    MN_CHECK_ARG(cq && nq && centroids && n_valid && R >= 1 && ld >= 1, "mn_centroids_finish: bad arguments");
    const int64_t n = (int64_t)R * ld;
    mn::centroid_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const long long*>(cq), nq, n, R, centroids, n_valid);
    MN_LAUNCH_CHECK();
    return 0;
}

int mn_group_sum(const float* in, int64_t ld, int32_t R, int32_t width, const int32_t* group, int32_t n_groups,
                 float* out, mn_stream_t stream) {
    MN_CHECK_ARG(in && group && out && R >= 1 && n_groups >= 1 && width >= 1, "mn_group_sum: bad arguments");
    dim3 grid((unsigned)((width + 127) / 128), (unsigned)n_groups);
    mn::group_sum_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(in, ld, R, width, group, n_groups, out);
    MN_LAUNCH_CHECK();
    return 0;
}

int mn_split_f16(const float* in, int64_t ld, int32_t R, int32_t G, mn_half_t* hi, mn_half_t* lo, int64_t ldh,
                 float* inv_scale, mn_stream_t stream) {
    MN_CHECK_ARG(in && hi && lo && inv_scale && R >= 1 && G >= 1 && ldh >= G, "mn_split_f16: bad arguments");
    mn::split_f16_kernel<<<(unsigned)R, 256, 0, (cudaStream_t)stream>>>(
        in, ld, G, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), ldh, inv_scale);
    MN_LAUNCH_CHECK();
    return 0;
}

int mn_votes_finalize(const uint64_t* votes, const uint64_t* degree, const int32_t* label, int64_t N, int32_t L,
                      int32_t ndn, float* out, int64_t ldo, mn_stream_t stream) {
    MN_CHECK_ARG(votes && label && out && N >= 1 && L >= 1 && ldo >= N, "mn_votes_finalize: bad arguments");
    mn::votes_finalize_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const unsigned long long*>(votes), reinterpret_cast<const unsigned long long*>(degree),
        label, N, L, ndn, out, ldo);
    MN_LAUNCH_CHECK();
    return 0;
}

int mn_mask_dropped(const int32_t* cell_label, const float* inv_norm, const uint8_t* flags, int64_t N,
                    int32_t* out_label, uint8_t* poison, mn_stream_t stream) {
    MN_CHECK_ARG(cell_label && inv_norm && out_label && N >= 1, "mn_mask_dropped: bad arguments");
    mn::mask_dropped_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cell_label, inv_norm, flags, N,
                                                                                          out_label, poison);
    MN_LAUNCH_CHECK();
    return 0;
}

int mn_loso_finalize(const void* in, int64_t ldi, int32_t is_fixed, const uint64_t* degree, const float* n_joint,
                     const int32_t* cell_study, int64_t N, int32_t S, int32_t K, int32_t ndn, float* out,
                     int64_t ldo, mn_stream_t stream) {
    MN_CHECK_ARG(in && cell_study && out && N >= 1 && S >= 2 && K >= 1 && ldo >= N, "mn_loso_finalize: bad arguments");
    MN_CHECK_ARG(is_fixed ? (degree != nullptr) : (n_joint != nullptr), "mn_loso_finalize: missing degree / n_joint");
    mn::loso_finalize_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        in, ldi, is_fixed, reinterpret_cast<const unsigned long long*>(degree), n_joint, cell_study, N, S, K, ndn,
        out, ldo);
    MN_LAUNCH_CHECK();
    return 0;
}

}
