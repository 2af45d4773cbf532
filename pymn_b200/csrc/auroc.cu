// K6 (sort form, used beyond 2,048 test labels; the product kernel is auroc_count.cu): rank-sum AUROC,
// K7: one-vs-best.
//
// K6 replaces compute_aurocs (reference pymn/utils.py:151-178: rankdata over
// the test cells of one study per train column, positives.T @ ranks),
// compute_aurocs_default (MetaNeighborUS.py:194-213) and the rank-sum block of
// score_default (MetaNeighbor.py:232-238).  One CTA per (train column, test
// study): the column's votes are read once from HBM, turned into
// order-preserving 32-bit keys (node-degree division fused into the load),
// sorted with a stable LSD radix sort whose ping-pong buffers stay in L2, and
// the exact average-tie ranks (integer 2*rank = first + last + 2) are summed
// per test label with integer shared-memory atomics.  One sort serves all
// test labels of the study (the reference re-sorts per label).
//
// K7 replaces compute_1v1_aurocs / find_top_candidates (utils.py:194-248).
#include "common.cuh"

namespace mn {

// The sort kernel is instantiated for 512 and 256 threads per CTA (8 keys per thread per tile): many
// (column, study) segments (config 4: 7,350 of 71k cells) run faster with more, smaller CTAs per SM, few
// long segments (config 5: 150 of 1M cells; config 2: 160 of 50k) with more threads per segment
// (measured on one box: c4 47.9 vs 40.9 ms, c5 28.7 vs 42.9 ms, c1 0.18 vs 0.21 ms for 512 vs 256).

__device__ __forceinline__ float vote_value(const float* __restrict__ v, const float* __restrict__ dn, int64_t i) {
    const float x = v[i];
    return dn ? __fdiv_rn(x, dn[i]) : x;
}

// inclusive max-scan over the block; returns the scanned value, *total = block max.
template <int ANW>
__device__ __forceinline__ int block_scan_max(int v, int* s_warp, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = max(v, t);
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < ANW) ? s_warp[lane] : INT_MIN;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w = max(w, t);
        }
        if (lane < ANW) s_warp[lane] = w;
    }
    __syncthreads();
    if (warp > 0) v = max(v, s_warp[warp - 1]);
    *total = s_warp[ANW - 1];
    return v;
}

__device__ __forceinline__ void acc_add(unsigned int* lo, unsigned int* hi, int idx, unsigned int x) {
    const unsigned int old = atomicAdd(&lo[idx], x);
    if (old > 0xffffffffu - x) atomicAdd(&hi[idx], 1u);
}

__global__ void fill_f64_kernel(double* p, int64_t n, double val) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = val;
}

template <int AT, int AITEMS>
__global__ void __launch_bounds__(AT)
auroc_kernel(const float* __restrict__ votes, int64_t ldv, int ncols, const float* __restrict__ denom,
             const int32_t* __restrict__ col_denom, const int32_t* __restrict__ seg_off,
             const int32_t* __restrict__ cell_label, int n_labels, int64_t M, double* __restrict__ auroc,
             int32_t* __restrict__ n_pos_out, uint32_t* keys0, uint32_t* keys1, uint16_t* labs0, uint16_t* labs1) {
    constexpr int ANW = AT / 32;            // warps
    constexpr int ATILE = AT * AITEMS;      // keys per tile
    static_assert(AT >= 256, "256 threads own the 256 digits");
    extern __shared__ __align__(16) unsigned char dyn[];
    // per-label sum of 2*rank as a (lo, hi) pair of 32-bit words: native 32-bit shared
    // atomics with an explicit carry instead of a 64-bit CAS loop
    unsigned int* acc_lo = reinterpret_cast<unsigned int*>(dyn);               // [n_labels]
    unsigned int* acc_hi = acc_lo + n_labels;                                   // [n_labels]
    unsigned int* cnt = acc_hi + n_labels;                                      // [n_labels]
    __shared__ unsigned int hist[4][256];
    __shared__ unsigned int wcnt[ANW][256];
    __shared__ unsigned int base[256];
    __shared__ unsigned int tcount[256], tstart[256];      // this tile's keys per digit / their start inside the tile
    __shared__ unsigned int s_scan[8];
    __shared__ uint32_t s_key[ATILE];                      // the tile in digit order, staged for coalesced write-out
    __shared__ uint16_t s_lab[ATILE];
    __shared__ int s_warp[32];
    __shared__ int s_flag[4];

    const int col = blockIdx.x, seg = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = lanemask_lt();
    const int r0 = seg_off[seg], r1 = seg_off[seg + 1];
    const int n = r1 - r0;
    if (n <= 0) return;
    const float* v = votes + (int64_t)col * ldv + r0;
    const float* dn = nullptr;
    if (denom && col_denom && col_denom[col] >= 0) dn = denom + (int64_t)col_denom[col] * ldv + r0;
    const int32_t* lab_in = cell_label + r0;
    uint32_t* kb[2] = {keys0 + (int64_t)col * M + r0, keys1 + (int64_t)col * M + r0};
    uint16_t* lb[2] = {labs0 + (int64_t)col * M + r0, labs1 + (int64_t)col * M + r0};

    for (int i = tid; i < n_labels; i += AT) {
        acc_lo[i] = 0u;
        acc_hi[i] = 0u;
        cnt[i] = 0u;
    }
    for (int i = tid; i < 4 * 256; i += AT) (&hist[0][0])[i] = 0u;
    __syncthreads();

    // ---- sweep A: range of the kept keys.  Sorting (key - kmin) instead of key makes the high
    // digits constant whenever the votes span less than 2^24 / 2^16 key values (they usually sit in
    // one or two binades), and constant digits are skipped below: 3 radix passes instead of 4.
    __shared__ uint32_t s_kmn[32], s_kmx[32];
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    for (int i0 = 0; i0 < n; i0 += 4 * AT) {
        float fv[4], fd[4];
        int fl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                      // four independent loads in flight per thread
            const int i = i0 + u * AT + tid;
            fl[u] = i < n ? lab_in[i] : -1;
            fv[u] = i < n ? v[i] : 0.f;
            fd[u] = (dn && i < n) ? dn[i] : 1.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (fl[u] >= 0) {
                const uint32_t k = float_key(dn ? __fdiv_rn(fv[u], fd[u]) : fv[u]);
                kmin = min(kmin, k);
                kmax = max(kmax, k);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if (lane == 0) {
        s_kmn[warp] = kmin;
        s_kmx[warp] = kmax;
    }
    __syncthreads();
    if (warp == 0) {
        kmin = (lane < ANW) ? s_kmn[lane] : 0xffffffffu;
        kmax = (lane < ANW) ? s_kmx[lane] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        }
        if (lane == 0) {
            s_kmn[0] = kmin;
            s_kmx[0] = kmax;
        }
    }
    __syncthreads();
    kmin = s_kmn[0];
    kmax = s_kmx[0];
    const uint32_t krange = (kmax >= kmin) ? (kmax - kmin) : 0u;
    const uint32_t sentinel = (krange == 0xffffffffu) ? krange : krange + 1u;   // excluded cells sort last
    if (kmax < kmin) kmin = 0u;

    // ---- sweep B: digit histograms of all four bytes + label counts -------
    for (int i0 = 0; i0 < n; i0 += 4 * AT) {
        float fv[4], fd[4];
        int fl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * AT + tid;
            fl[u] = i < n ? lab_in[i] : -1;
            fv[u] = i < n ? v[i] : 0.f;
            fd[u] = (dn && i < n) ? dn[i] : 1.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u * AT + tid >= n) continue;          // past the end of the segment
            uint32_t key = sentinel;
            if (fl[u] >= 0) {
                key = float_key(dn ? __fdiv_rn(fv[u], fd[u]) : fv[u]) - kmin;
                atomicAdd(&cnt[fl[u]], 1u);
            }
            atomicAdd(&hist[0][key & 255u], 1u);
            atomicAdd(&hist[1][(key >> 8) & 255u], 1u);
            atomicAdd(&hist[2][(key >> 16) & 255u], 1u);
            atomicAdd(&hist[3][key >> 24], 1u);
        }
    }
    __syncthreads();
    if (tid < 4) s_flag[tid] = 1;     // 1 = pass needed
    __syncthreads();
    for (int i = tid; i < 4 * 256; i += AT)
        if ((&hist[0][0])[i] == (unsigned)n) s_flag[i >> 8] = 0;   // every key shares this digit
    __syncthreads();

    // ---- stable LSD radix passes ----------------------------------------------
    int cur = -1;                       // -1: data still lives in votes / cell_label
    for (int pass = 0; pass < 4; ++pass) {
        if (!s_flag[pass]) continue;    // uniform
        const int shift = 8 * pass;
        if (tid == 0) {
            unsigned run = 0;
            for (int d = 0; d < 256; ++d) {
                base[d] = run;
                run += hist[pass][d];
            }
        }
        const int nxt = (cur == 0) ? 1 : 0;
        const uint32_t* ksrc = (cur >= 0) ? kb[cur] : nullptr;
        const uint16_t* lsrc = (cur >= 0) ? lb[cur] : nullptr;
        uint32_t* kdst = kb[nxt];
        uint16_t* ldst = lb[nxt];
        for (int t0 = 0; t0 < n; t0 += ATILE) {
            for (int x = tid; x < ANW * 256; x += AT) (&wcnt[0][0])[x] = 0u;
            __syncthreads();
            uint32_t key[AITEMS];
            uint16_t lab[AITEMS];
            int off[AITEMS];
            unsigned int* wc = wcnt[warp];
            // all of the tile's loads first (independent, AITEMS in flight per thread) ...
            if (ksrc) {
#pragma unroll
                for (int j = 0; j < AITEMS; ++j) {
                    const int idx = t0 + warp * (32 * AITEMS) + j * 32 + lane;
                    key[j] = idx < n ? ksrc[idx] : 0u;
                    lab[j] = idx < n ? lsrc[idx] : (uint16_t)0;
                }
            } else {
                float fv[AITEMS], fd[AITEMS];
                int fl[AITEMS];
#pragma unroll
                for (int j = 0; j < AITEMS; ++j) {
                    const int idx = t0 + warp * (32 * AITEMS) + j * 32 + lane;
                    fv[j] = idx < n ? v[idx] : 0.f;
                    fd[j] = (dn && idx < n) ? dn[idx] : 1.f;
                    fl[j] = idx < n ? lab_in[idx] : -1;
                }
#pragma unroll
                for (int j = 0; j < AITEMS; ++j) {
                    const float val = dn ? __fdiv_rn(fv[j], fd[j]) : fv[j];
                    key[j] = (fl[j] >= 0) ? float_key(val) - kmin : sentinel;
                    lab[j] = (uint16_t)(fl[j] >= 0 ? fl[j] : 0xffff);
                }
            }
            // ... then the stable in-warp ranking of their digits
#pragma unroll
            for (int j = 0; j < AITEMS; ++j) {
                const int idx = t0 + warp * (32 * AITEMS) + j * 32 + lane;
                const bool in = idx < n;
                off[j] = -1;
                const unsigned active = __ballot_sync(0xffffffffu, in);
                unsigned peers = 0, prior = 0;
                const unsigned dig = (key[j] >> shift) & 255u;
                if (in) {
                    peers = __match_any_sync(active, dig);
                    prior = wc[dig];
                    off[j] = (int)(prior + __popc(peers & lt));
                }
                __syncwarp();
                if (in && (peers & lt) == 0) wc[dig] = prior + __popc(peers);
                __syncwarp();
            }
            __syncthreads();
            // Scatter through shared memory: a key's 4-byte store to its bucket is a whole 32-byte sector write
            // for L2, and with 256 buckets almost every store of a tile hits a different sector.  Instead the
            // tile is first put into digit order in shared memory; consecutive threads then write consecutive
            // elements of a digit's run to consecutive global addresses.
            unsigned incl = 0;
            if (tid < 256) {
                unsigned run = 0;
#pragma unroll
                for (int w = 0; w < ANW; ++w) {
                    const unsigned c = wcnt[w][tid];
                    wcnt[w][tid] = run;                    // offset of warp w's keys inside digit tid's run of this tile
                    run += c;
                }
                tcount[tid] = run;
                incl = run;                                 // inclusive scan over the 256 digits -> start of each run
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (lane == 31) s_scan[warp] = incl;
            }
            __syncthreads();
            if (tid < 256) {
                unsigned before = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w)
                    if (w < warp) before += s_scan[w];
                tstart[tid] = before + incl - tcount[tid];
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < AITEMS; ++j) {
                if (off[j] >= 0) {
                    const unsigned dig = (key[j] >> shift) & 255u;
                    const unsigned lpos = tstart[dig] + wc[dig] + (unsigned)off[j];
                    s_key[lpos] = key[j];
                    s_lab[lpos] = lab[j];
                }
            }
            __syncthreads();
            const int tile_n = min(ATILE, n - t0);
            for (int i = tid; i < tile_n; i += AT) {
                const uint32_t k = s_key[i];
                const unsigned dig = (k >> shift) & 255u;
                const unsigned pos = base[dig] + ((unsigned)i - tstart[dig]);
                kdst[pos] = k;
                ldst[pos] = s_lab[i];
            }
            __syncthreads();
            if (tid < 256) base[tid] += tcount[tid];
        }
        cur = nxt;
        __threadfence_block();
        __syncthreads();
    }

    // number of kept cells
    unsigned nv_part = 0;
    for (int i = tid; i < n_labels; i += AT) nv_part += cnt[i];
    __shared__ unsigned int s_red[32];
    const int n_valid = (int)block_sum<unsigned int>(nv_part, s_red);

    if (cur < 0) {
        // every key identical: either one big tie of kept cells, or nothing kept
        for (int i = tid; i < n_labels; i += AT) {
            const unsigned long long t = (unsigned long long)cnt[i] * (unsigned long long)(n_valid + 1);
            acc_lo[i] = (unsigned int)t;
            acc_hi[i] = (unsigned int)(t >> 32);
        }
        __syncthreads();
    } else if (n_valid > 0) {
        const uint32_t* ks = kb[cur];
        const uint16_t* ls = lb[cur];
        // Sorted keys -> 2*rank = first + last + 2 of the element's tie run.  Ties are rare among
        // float votes, so every element just compares with its two neighbours; only a tied element
        // binary-searches the ends of its run.  No block-wide communication: eight independent
        // elements per thread keep the L2 loads in flight.
        for (int t0 = 0; t0 < n_valid; t0 += 8 * AT) {
            uint32_t k[8], kp[8], kn[8];
            uint16_t lab8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int p = t0 + u * AT + tid;
                const bool in = p < n_valid;
                k[u] = in ? ks[p] : 0u;
                kp[u] = (in && p > 0) ? ks[p - 1] : ~k[u];
                kn[u] = (in && p + 1 < n_valid) ? ks[p + 1] : ~k[u];
                lab8[u] = in ? ls[p] : (uint16_t)0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int p = t0 + u * AT + tid;
                if (p >= n_valid) continue;
                int first = p, last = p;
                if (kp[u] == k[u]) {                     // first position whose key >= k
                    int lo = 0, hi = p - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (ks[mid] < k[u]) lo = mid + 1; else hi = mid;
                    }
                    first = lo;
                }
                if (kn[u] == k[u]) {                     // last position whose key <= k
                    int lo = p + 1, hi = n_valid - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (ks[mid] > k[u]) hi = mid - 1; else lo = mid;
                    }
                    last = lo;
                }
                acc_add(acc_lo, acc_hi, lab8[u], (unsigned int)(first + last + 2));
            }
        }
        __syncthreads();
    }

    for (int c = tid; c < n_labels; c += AT) {
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

// ---------------------------------------------------------------------------
// K7 one-vs-best: one CTA per (train column, test study)
// ---------------------------------------------------------------------------
constexpr int OT = 256;
constexpr int OCHUNK = 2048;

__global__ void __launch_bounds__(OT)
one_vs_best_kernel(const float* __restrict__ votes, int64_t ldv, int ncols, const float* __restrict__ denom,
                   const int32_t* __restrict__ col_denom, const int32_t* __restrict__ seg_lab_off,
                   const int32_t* __restrict__ lab_row_off, const int32_t* __restrict__ cell_label,
                   const double* __restrict__ auroc, double* __restrict__ out) {
    __shared__ float s_vals[OCHUNK];
    __shared__ int s_cand[5];
    __shared__ int s_ncand;
    __shared__ int s_best, s_second;
    __shared__ double s_score;
    __shared__ unsigned long long s_red[32];

    const int col = blockIdx.x, seg = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    const int lb0 = seg_lab_off[seg], lb1 = seg_lab_off[seg + 1];
    const int nl = lb1 - lb0;
    if (nl < 2) return;
    const float* v = votes + (int64_t)col * ldv;
    const float* dn = nullptr;
    if (denom && col_denom && col_denom[col] >= 0) dn = denom + (int64_t)col_denom[col] * ldv;

    // ---- top-5 labels by AUROC: descending, NaN never selected, ties -> lower index
    if (tid < 32) {
        int chosen[5];
        int nc = 0;
        for (int r = 0; r < 5; ++r) {
            double bv = 0.0;
            int bi = -1;
            for (int c = lane; c < nl; c += 32) {
                bool used = false;
                for (int q = 0; q < r; ++q) used |= (q < nc && chosen[q] == c);
                if (used) continue;
                const double a = auroc[(int64_t)(lb0 + c) * ncols + col];
                if (a != a) continue;
                if (bi < 0 || a > bv) {
                    bv = a;
                    bi = c;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) {
                    bv = ov;
                    bi = oi;
                }
            }
            if (bi < 0) break;        // uniform across the warp after the butterfly
            chosen[nc++] = bi;
        }
        if (lane == 0) {
            s_ncand = nc;
            for (int q = 0; q < nc; ++q) s_cand[q] = chosen[q];
            s_best = nc > 0 ? chosen[0] : -1;
            s_second = nc > 1 ? chosen[1] : -1;
            s_score = 1.0;
        }
    }
    __syncthreads();
    const int nc = s_ncand;
    if (nc < 2) return;

    for (int ci = 1; ci < nc; ++ci) {
        const int best = s_best, cont = s_cand[ci];
        const int b0 = lab_row_off[lb0 + best], b1 = lab_row_off[lb0 + best + 1];
        const int c0 = lab_row_off[lb0 + cont], c1 = lab_row_off[lb0 + cont + 1];
        unsigned long long u2 = 0, nb = 0, ncn = 0;
        const float pinf = __int_as_float(0x7f800000);
        for (int cc = c0; cc < c1; cc += OCHUNK) {
            const int m = min(OCHUNK, c1 - cc);
            __syncthreads();
            for (int j = tid; j < m; j += OT) {
                const bool ok = cell_label[cc + j] >= 0;
                s_vals[j] = ok ? vote_value(v, dn, cc + j) : pinf;
                ncn += ok ? 1 : 0;
            }
            // sort the contender's chunk (bitonic, +inf = excluded cell / padding), then every cell of the best label
            // counts the smaller and the equal votes with two binary searches: the same integers as comparing
            // all pairs, at ~1/6 of the instructions for ~500-cell clusters
            int P = 32;
            while (P < m) P <<= 1;
            for (int j = m + tid; j < P; j += OT) s_vals[j] = __int_as_float(0x7f800000);
            __syncthreads();
            for (int k = 2; k <= P; k <<= 1) {
                for (int jj = k >> 1; jj > 0; jj >>= 1) {
                    for (int t = tid; t < P; t += OT) {
                        const int ixj = t ^ jj;
                        if (ixj > t) {
                            const float a = s_vals[t], b = s_vals[ixj];
                            const bool up = (t & k) == 0;
                            if ((a > b) == up) {
                                s_vals[t] = b;
                                s_vals[ixj] = a;
                            }
                        }
                    }
                    __syncthreads();
                }
            }
            for (int i = b0 + tid; i < b1; i += OT) {
                if (cell_label[i] < 0) continue;
                const float vi = vote_value(v, dn, i);
                int lo = 0, hi = P;                          // first position with s_vals[pos] >= vi
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (s_vals[mid] < vi) lo = mid + 1; else hi = mid;
                }
                const int less = lo;
                hi = P;                                      // first position with s_vals[pos] > vi
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (s_vals[mid] <= vi) lo = mid + 1; else hi = mid;
                }
                u2 += 2u * (unsigned)less + (unsigned)(lo - less);
            }
        }
        for (int i = b0 + tid; i < b1; i += OT) nb += (cell_label[i] >= 0) ? 1 : 0;
        u2 = block_sum<unsigned long long>(u2, s_red);
        nb = block_sum<unsigned long long>(nb, s_red);
        ncn = block_sum<unsigned long long>(ncn, s_red);
        if (tid == 0) {
            const double n_b = (double)nb, n_c = (double)ncn;
            // rank-sum of the best's cells inside best U contender, same op order as utils.py:175-178
            double a = ((double)(nb * (nb + 1ull) + u2) * 0.5) / n_b;
            a -= (n_b + 1.0) / 2.0;
            a /= n_c;
            if (a < 0.5) {
                s_second = best;
                s_best = cont;
                s_score = 1.0 - a;
            } else if (a < s_score) {
                s_score = a;
                s_second = cont;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        out[(int64_t)(lb0 + s_best) * ncols + col] = s_score;
        out[(int64_t)(lb0 + s_second) * ncols + col] = 1.0 - s_score;
    }
}

}  // namespace mn

extern "C" {

static int64_t sort_workspace_bytes(int32_t ncols, int64_t M) {
    const int64_t per = ((M + 63) / 64) * 64;
    return (int64_t)ncols * per * (4 + 4 + 2 + 2) + 1024;
}

// The counting kernel (auroc_count.cu) packs (key remainder | segment-local label) into 32 bits: up to 2,048
// labels.  Beyond that (or with MN_AUROC_SORT=1, for A/B measurements) the radix-sort kernel above is used.
static bool use_sort_kernel(int32_t n_labels) {
    if (n_labels > 2048) return true;
    const char* e = getenv("MN_AUROC_SORT");
    return e && atoi(e) != 0;
}

int64_t mn_auroc_workspace_bytes(int32_t ncols, int64_t M, int32_t n_seg, int32_t n_labels) {
    if (use_sort_kernel(n_labels)) return sort_workspace_bytes(ncols, M);
    return mn::auroc_count_workspace_bytes(ncols, M, n_seg, n_labels);
}

int mn_auroc(const float* votes, int64_t ldv, int32_t ncols, const float* denom, const int32_t* col_denom,
             const int32_t* seg_off, int32_t n_seg, const int32_t* cell_label, int32_t n_labels, int64_t M,
             double* auroc, int32_t* n_pos, void* workspace, int64_t workspace_bytes, mn_stream_t stream) {
    MN_CHECK_ARG(votes && seg_off && cell_label && auroc && workspace, "mn_auroc: null pointer");
    MN_CHECK_ARG(ncols >= 1 && n_seg >= 1 && n_labels >= 1 && M >= 1 && ldv >= M, "mn_auroc: bad shape");
    MN_CHECK_ARG(n_labels <= 16000, "mn_auroc: n_labels=%d > 16000 unsupported", n_labels);
    MN_CHECK_ARG(M < (1ll << 30), "mn_auroc: M too large");
    MN_CHECK_ARG(workspace_bytes >= mn_auroc_workspace_bytes(ncols, M, n_seg, n_labels), "mn_auroc: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (!use_sort_kernel(n_labels))
        return mn::launch_auroc_count(votes, ldv, ncols, denom, col_denom, seg_off, n_seg, cell_label, n_labels, M, auroc,
                                      n_pos, workspace, st);
    const int64_t per = ((M + 63) / 64) * 64;
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    uint32_t* keys0 = reinterpret_cast<uint32_t*>(w);
    uint32_t* keys1 = keys0 + (int64_t)ncols * per;
    uint16_t* labs0 = reinterpret_cast<uint16_t*>(keys1 + (int64_t)ncols * per);
    uint16_t* labs1 = labs0 + (int64_t)ncols * per;
    const int64_t tot = (int64_t)n_labels * ncols;
    mn::fill_f64_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(auroc, tot, nan(""));
    MN_LAUNCH_CHECK();
    if (n_pos) MN_CUDA(cudaMemsetAsync(n_pos, 0, sizeof(int32_t) * n_labels, st));
    const size_t dyn = (size_t)n_labels * 12 + 16;
    dim3 grid((unsigned)ncols, (unsigned)n_seg);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // many segments -> 256-thread CTAs (more of them per SM); few -> 512 threads per segment
    const bool small_cta = (int64_t)ncols * n_seg >= 8ll * sms;
    // (the kernels' static shared memory is ~28-47 KB: any dynamic part needs the opt-in; set on every call --
    // the attribute is per device and a process may drive several)
    if (small_cta) {
        MN_CUDA(cudaFuncSetAttribute(mn::auroc_kernel<256, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        mn::auroc_kernel<256, 8><<<grid, 256, dyn, st>>>(votes, ldv, ncols, denom, col_denom, seg_off, cell_label, n_labels,
                                                        per, auroc, n_pos, keys0, keys1, labs0, labs1);
    } else {
        MN_CUDA(cudaFuncSetAttribute(mn::auroc_kernel<512, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        mn::auroc_kernel<512, 8><<<grid, 512, dyn, st>>>(votes, ldv, ncols, denom, col_denom, seg_off, cell_label, n_labels,
                                                        per, auroc, n_pos, keys0, keys1, labs0, labs1);
    }
    MN_LAUNCH_CHECK();
    return 0;
}

int mn_one_vs_best(const float* votes, int64_t ldv, int32_t ncols, const float* denom, const int32_t* col_denom,
                   const int32_t* seg_lab_off, int32_t n_seg, const int32_t* lab_row_off,
                   const int32_t* cell_label, int32_t n_labels, const double* auroc, double* out,
                   mn_stream_t stream) {
    MN_CHECK_ARG(votes && seg_lab_off && lab_row_off && cell_label && auroc && out, "mn_one_vs_best: null pointer");
    MN_CHECK_ARG(ncols >= 1 && n_seg >= 1 && n_labels >= 1, "mn_one_vs_best: bad shape");
    dim3 grid((unsigned)ncols, (unsigned)n_seg);
    mn::one_vs_best_kernel<<<grid, mn::OT, 0, (cudaStream_t)stream>>>(votes, ldv, ncols, denom, col_denom, seg_lab_off,
                                                                     lab_row_off, cell_label, auroc, out);
    MN_LAUNCH_CHECK();
    return 0;
}

}
