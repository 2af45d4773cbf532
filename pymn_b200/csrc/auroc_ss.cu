// K6 (fast path): rank-sum AUROC by SAMPLE SORT, one CTA per (train column, test study).
//
// Same contract as auroc_kernel (auroc.cu; replaces reference pymn/utils.py:151-178,
// MetaNeighborUS.py:194-213, MetaNeighbor.py:232-238).  A full radix sort of the column costs
// ~500 instructions per vote; the rank sums only need every element's count of smaller / equal
// votes, which a bucket decomposition delivers with a handful of block barriers:
//   1. sample the column (up to 8192 votes, strided), sort the sample in shared memory and keep
//      every second value as a splitter;
//   2. classify every vote by binary search into the class "strictly between two splitters"
//      (even id) or "equal to a splitter value" (odd id) -- heavy ties therefore get a class of
//      their own; count the classes, remember (key, class id);
//   3. exclusive scan of the class counts = number of votes below each class;
//   4. scatter (key, label) class by class (shared-memory cursors; order inside a class is free);
//   5. one warp per class: an "equal" class is one tie run (closed-form rank); a "between" class
//      (about 16 votes on average) is ranked by all-pairs comparison over warp shuffles:
//      2*rank = 2*(below + less) + equal + 1, summed per test label with shared atomics.
// Exact average-tie ranks of the float32 votes, integer rank sums -> bit-identical to the oracle.
#include "common.cuh"

namespace mn {

constexpr int ST = 512;                  // threads per CTA
constexpr int SNW = ST / 32;
constexpr int SS_MAX_SPLIT = 4096;       // splitters
constexpr int SS_MAX_SAMPLE = 8192;      // sample size (2 per splitter)

__device__ __forceinline__ void ss_acc_add(unsigned int* lo, unsigned int* hi, int idx, unsigned int x) {
    const unsigned int old = atomicAdd(&lo[idx], x);
    if (old > 0xffffffffu - x) atomicAdd(&hi[idx], 1u);
}

__device__ __forceinline__ float vote_value_ss(const float* __restrict__ v, const float* __restrict__ dn, int i) {
    const float x = v[i];
    return dn ? __fdiv_rn(x, dn[i]) : x;
}

__device__ __forceinline__ void bitonic_sort_u32(uint32_t* a, int P) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool up = ((i & k) == 0);
                const uint32_t x = a[i], y = a[l];
                if ((x > y) == up) {
                    a[i] = y;
                    a[l] = x;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(ST)
auroc_ss_kernel(const float* __restrict__ votes, int64_t ldv, int ncols, const float* __restrict__ denom,
                const int32_t* __restrict__ col_denom, const int32_t* __restrict__ seg_off,
                const int32_t* __restrict__ cell_label, int n_labels, int64_t M, double* __restrict__ auroc,
                int32_t* __restrict__ n_pos_out, uint32_t* keys0, uint32_t* keys1, uint16_t* bids, uint16_t* labs1) {
    extern __shared__ __align__(16) unsigned char dyn[];
    uint32_t* spl = reinterpret_cast<uint32_t*>(dyn);                     // [SS_MAX_SPLIT] splitters
    uint32_t* offs = spl + SS_MAX_SPLIT;                                    // [2*SS_MAX_SPLIT + 2] class offsets
    uint32_t* curs = offs + 2 * SS_MAX_SPLIT + 2;                           // [2*SS_MAX_SPLIT + 2] sample buffer, then cursors
    unsigned int* acc_lo = curs + 2 * SS_MAX_SPLIT + 2;                     // [n_labels]
    unsigned int* acc_hi = acc_lo + n_labels;
    unsigned int* cnt = acc_hi + n_labels;
    __shared__ unsigned int s_scan[SNW];
    __shared__ unsigned int s_total;

    const int col = blockIdx.x, seg = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = seg_off[seg], r1 = seg_off[seg + 1];
    const int n = r1 - r0;
    if (n <= 0) return;
    const float* v = votes + (int64_t)col * ldv + r0;
    const float* dn = nullptr;
    if (denom && col_denom && col_denom[col] >= 0) dn = denom + (int64_t)col_denom[col] * ldv + r0;
    const int32_t* lab_in = cell_label + r0;
    uint32_t* k0 = keys0 + (int64_t)col * M + r0;       // keys in cell order
    uint32_t* k1 = keys1 + (int64_t)col * M + r0;       // keys grouped by class
    uint16_t* bid = bids + (int64_t)col * M + r0;       // class id in cell order
    uint16_t* l1 = labs1 + (int64_t)col * M + r0;       // labels grouped by class

    // number of splitters: classes of ~16 votes, at most SS_MAX_SPLIT
    int ns = 32;
    while (ns < SS_MAX_SPLIT && ns * 16 < n) ns <<= 1;
    const int n_samp = 2 * ns;                           // power of two <= SS_MAX_SAMPLE
    const int n_class = 2 * ns + 1;

    for (int i = tid; i < n_labels; i += ST) {
        acc_lo[i] = 0u;
        acc_hi[i] = 0u;
        cnt[i] = 0u;
    }
    // ---- 1. sample + sort + splitters ----------------------------------------------------------
    for (int k = tid; k < n_samp; k += ST) {
        uint32_t key = 0xffffffffu;                      // padding / excluded cells sort last
        const int i = (n_samp <= n) ? (int)(((long long)k * n) / n_samp) : k;     // strided sample, or everything
        if (i < n && lab_in[i] >= 0) key = float_key(vote_value_ss(v, dn, i));
        curs[k] = key;
    }
    __syncthreads();
    bitonic_sort_u32(curs, n_samp);
    for (int j = tid; j < ns; j += ST) spl[j] = curs[2 * j + 1];
    __syncthreads();
    for (int b = tid; b < n_class + 1; b += ST) offs[b] = 0u;
    __syncthreads();

    // ---- 2. classify, count ---------------------------------------------------------------------
    for (int i0 = 0; i0 < n; i0 += 4 * ST) {
        float fv[4], fd[4];
        int fl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * ST + tid;
            fl[u] = i < n ? lab_in[i] : -1;
            fv[u] = i < n ? v[i] : 0.f;
            fd[u] = (dn && i < n) ? dn[i] : 1.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * ST + tid;
            if (i >= n) continue;
            uint32_t key = 0xffffffffu;
            if (fl[u] >= 0) {
                key = float_key(dn ? __fdiv_rn(fv[u], fd[u]) : fv[u]);
                atomicAdd(&cnt[fl[u]], 1u);
            }
            int lo = 0, hi = ns;                          // number of splitters < key
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (spl[mid] < key) lo = mid + 1; else hi = mid;
            }
            const int id = 2 * lo + ((lo < ns && spl[lo] == key) ? 1 : 0);
            atomicAdd(&offs[id], 1u);
            k0[i] = key;
            bid[i] = (uint16_t)id;
        }
    }
    __syncthreads();

    // ---- 3. exclusive scan of the class counts ----------------------------------------------------------
    {
        const int per = (n_class + ST - 1) / ST;
        const int b0 = tid * per, b1 = min(n_class, b0 + per);
        unsigned int local = 0;
        for (int b = b0; b < b1; ++b) local += offs[b];
        unsigned int inc = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned int w = lane < SNW ? s_scan[lane] : 0u;
            unsigned int winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            if (lane < SNW) s_scan[lane] = winc - w;     // exclusive warp offsets
            if (lane == SNW - 1) s_total = winc;
        }
        __syncthreads();
        unsigned int run = s_scan[warp] + inc - local;
        for (int b = b0; b < b1; ++b) {
            const unsigned int c = offs[b];
            offs[b] = run;
            curs[b] = run;
            run += c;
        }
        if (tid == 0) offs[n_class] = (unsigned int)n;
    }
    __syncthreads();

    // ---- 4. scatter by class ---------------------------------------------------------------------
    for (int i0 = 0; i0 < n; i0 += 4 * ST) {
        uint32_t kk[4];
        uint16_t bb[4];
        int fl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * ST + tid;
            kk[u] = i < n ? k0[i] : 0u;
            bb[u] = i < n ? bid[i] : (uint16_t)0;
            fl[u] = i < n ? lab_in[i] : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * ST + tid;
            if (i >= n) continue;
            const unsigned int pos = atomicAdd(&curs[bb[u]], 1u);
            k1[pos] = kk[u];
            l1[pos] = (uint16_t)(fl[u] >= 0 ? fl[u] : 0xffff);
        }
    }
    __syncthreads();

    // ---- 5. rank inside the classes, one warp per class ---------------------------------------------
    for (int b = warp; b < n_class; b += SNW) {
        const int o0 = (int)offs[b], o1 = (int)offs[b + 1];
        const int m = o1 - o0;
        if (m <= 0) continue;
        if (b & 1) {
            // all votes of the class equal one splitter value: one tie run
            const unsigned int r2 = (unsigned int)(o0 + o1 - 1 + 2);
            for (int e = o0 + lane; e < o1; e += 32) {
                const uint16_t lb = l1[e];
                if (lb != 0xffff) ss_acc_add(acc_lo, acc_hi, lb, r2);
            }
            continue;
        }
        for (int ib = o0; ib < o1; ib += 32) {
            const int e = ib + lane;
            const bool in_i = e < o1;
            const uint32_t ki = in_i ? k1[e] : 0u;
            const uint16_t li = in_i ? l1[e] : (uint16_t)0xffff;
            int less = 0, eq = 0;
            for (int jb = o0; jb < o1; jb += 32) {
                const int f = jb + lane;
                const uint32_t kj = f < o1 ? k1[f] : 0xffffffffu;
                const int cntj = min(32, o1 - jb);
                for (int t = 0; t < cntj; ++t) {
                    const uint32_t kt = __shfl_sync(0xffffffffu, kj, t);
                    less += (kt < ki) ? 1 : 0;
                    eq += (kt == ki) ? 1 : 0;
                }
            }
            if (in_i && li != 0xffff) ss_acc_add(acc_lo, acc_hi, li, (unsigned int)(2 * (o0 + less) + eq + 1));
        }
    }
    __syncthreads();

    // ---- AUROC per test label (same operation order as the reference, utils.py:175-178) ------------------
    unsigned nv_part = 0;
    for (int i = tid; i < n_labels; i += ST) nv_part += cnt[i];
    __shared__ unsigned int s_red[32];
    const int n_valid = (int)block_sum<unsigned int>(nv_part, s_red);
    for (int c = tid; c < n_labels; c += ST) {
        const unsigned np = cnt[c];
        if (np == 0) continue;
        const double n_p = (double)np, n_n = (double)(n_valid - (int)np);
        const unsigned long long sum2 = ((unsigned long long)acc_hi[c] << 32) | acc_lo[c];
        double roc = ((double)sum2 * 0.5) / n_p;
        roc -= (n_p + 1.0) / 2.0;
        roc /= n_n;
        auroc[(int64_t)c * ncols + col] = roc;
        if (col == 0 && n_pos_out) n_pos_out[c] = (int)np;
    }
}

size_t auroc_ss_smem_bytes(int n_labels) {
    return (size_t)(SS_MAX_SPLIT + 2 * (2 * SS_MAX_SPLIT + 2)) * 4 + (size_t)n_labels * 12 + 16;
}

int launch_auroc_ss(const float* votes, int64_t ldv, int ncols, const float* denom, const int32_t* col_denom,
                    const int32_t* seg_off, int n_seg, const int32_t* cell_label, int n_labels, int64_t per,
                    double* auroc, int32_t* n_pos, uint32_t* keys0, uint32_t* keys1, uint16_t* labs0, uint16_t* labs1,
                    cudaStream_t st) {
    const size_t dyn = auroc_ss_smem_bytes(n_labels);
    static size_t configured = 0;
    if (dyn > configured) {
        MN_CUDA(cudaFuncSetAttribute(auroc_ss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        configured = dyn;
    }
    dim3 grid((unsigned)ncols, (unsigned)n_seg);
    auroc_ss_kernel<<<grid, ST, dyn, st>>>(votes, ldv, ncols, denom, col_denom, seg_off, cell_label, n_labels, per, auroc,
                                          n_pos, keys0, keys1, labs0, labs1);
    MN_LAUNCH_CHECK();
    return 0;
}

}  // namespace mn
