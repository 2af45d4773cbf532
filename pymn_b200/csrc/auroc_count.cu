// K6 (counting form): exact rank-sum AUROC without sorting.
//
// Replaces compute_aurocs (reference pymn/utils.py:151-178: rankdata over the test cells of one study per
// train column, positives.T @ ranks), compute_aurocs_default (MetaNeighborUS.py:194-213) and the rank-sum
// block of score_default (MetaNeighbor.py:232-238).
//
// Only the per-label SUMS of the average-tie ranks are needed, so nothing is sorted.  A persistent CTA owns
// one (train column, test study) segment at a time:
//   P0  votes (/ node degree) -> order-preserving 32-bit keys, written once to the CTA's key slot (global
//       memory, re-read from L2 by the later sweeps), min / max of the kept keys
//   P1  exact coarse histogram (2,048 bins over [min, max]) -> cumulative counts
//   more keys than one round holds (cap ~ 43k = what fits in shared memory): one PARTITION sweep deals keys and
//       labels to the rounds' contiguous ranges of a second slot (round = the coarse bins whose cumulative
//       count starts in [r capq, (r + 1) capq)), so that every round streams only its own keys
//   per ROUND:
//     LUT  every occupied coarse bin gets fine bins in proportion to its count, so the 16,384 fine bins are
//          equi-populated (~2.6 keys each) whatever the shape of the vote distribution
//     P2   fine histogram (packed 16-bit shared-memory atomics), scan, bitmap of the bin starts
//     P3   counting-sort scatter of (key remainder | label) into shared memory -- bin order, arbitrary inside
//     P4   one thread per slot: both ends of its bin from a 64-bit window of the bitmap (no branch), its bin's few
//          entries compared pairwise,
//          2*rank = 2*(keys below the bin) + 2*(smaller in the bin) + (equal in the bin, self included) + 1,
//          added to its label's integer accumulator (32-bit shared atomics with carry)
//   A coarse bin holding more than half a round (heavy ties / extreme skew) switches the task to greedy rounds
//   with a range filter; a bin that alone exceeds cap is refined recursively (task stack); a range of one key
//   value is a tie block and is closed in one sweep.  Fine bins of more than 32 entries (tie blocks) are handled
//   by the whole CTA: all-equal -> closed form, otherwise cooperative pairwise counting.
// Random-bank shared-memory atomics run at ~9 lanes/clk/SM on B200 (tools/micro/smem_atomics.cu), the same as
// random loads, which is why counting beats the radix sort it replaces (`match.any` ranking: 0.55 lanes/clk/SM).
// Rank sums are integers: given the float32 votes the AUROCs are bit-identical to the oracle's.
#include "common.cuh"

namespace mn {
namespace {

constexpr int QT = 1024;                 // threads per CTA (one CTA per SM)
constexpr int QCB_LOG = 11;
constexpr int QCB = 1 << QCB_LOG;        // coarse bins per task
constexpr int QF = 16 * QT;              // fine bins per round: thread t scans bins [16t, 16t+16)
constexpr uint32_t QSENT = 0xffffffffu;  // key of an excluded cell
constexpr int QBIG = 64;                 // bins above this are resolved by the whole CTA (tie blocks)
constexpr int QSMALL = 8;                // bins above this go to the CTA-cooperative path (every slot of a bin of
                                         // <= 8 entries finds both ends inside one 32-bit window of the bitmap)
constexpr int QFIXED = (QCB + QCB + QF / 2 + (QT + 4) + 36) * 4;   // fixed part of the dynamic shared memory
constexpr int QSMEM = 232448 - 1024;     // dynamic shared memory per CTA (227 KB minus the static part)

struct QTask {
    uint32_t lo, range, below, count;    // keys in [lo, lo + range], `below` kept keys are smaller, `count` inside
};

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_w, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_w[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        s_w[lane] = wi - w;
        if (lane == 31) s_w[32] = wi;
    }
    __syncthreads();
    *total = s_w[32];
    return inc - v + s_w[warp];
}

__device__ __forceinline__ void acc_add32(unsigned int* lo, unsigned int* hi, uint32_t idx, uint32_t x) {
    const unsigned int old = atomicAdd(&lo[idx], x);
    if (old > 0xffffffffu - x) atomicAdd(&hi[idx], 1u);
}

// NaN-fill of the table, zero of the counters
__global__ void auroc_init_kernel(double* auroc, int64_t n_auroc, int32_t* n_pos, int n_labels, int32_t* seg_cnt,
                                  int64_t n_seg_cnt, int32_t* seg_info, int n_seg, unsigned int* counter) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    for (int64_t k = i; k < n_auroc; k += stride) auroc[k] = qnan;
    for (int64_t k = i; k < n_seg_cnt; k += stride) seg_cnt[k] = 0;
    if (n_pos)
        for (int64_t k = i; k < n_labels; k += stride) n_pos[k] = 0;
    for (int64_t k = i; k < n_seg; k += stride) {
        seg_info[4 * k + 0] = 0;            // kept cells
        seg_info[4 * k + 1] = INT_MAX;      // lowest label id
        seg_info[4 * k + 2] = -1;           // highest label id
        seg_info[4 * k + 3] = 0;
    }
    if (i == 0) *counter = 0u;
}

// kept cells per (segment, label), label range and kept cells per segment (column-independent)
__global__ void auroc_label_count_kernel(const int32_t* __restrict__ seg_off, int n_seg,
                                         const int32_t* __restrict__ cell_label, int n_labels, int64_t M,
                                         int32_t* seg_cnt, int32_t* seg_info) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int lab = -1, seg = -1;
    if (i < M) {
        lab = cell_label[i];
        if (lab >= n_labels) lab = -1;
        int lo = 0, hi = n_seg;             // last segment with seg_off[seg] <= i
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (seg_off[mid] <= i) lo = mid; else hi = mid;
        }
        seg = lo;
        if (i < seg_off[0] || i >= seg_off[n_seg]) lab = -1;
    }
    const unsigned live = __ballot_sync(0xffffffffu, lab >= 0);
    if (lab < 0) return;
    const unsigned peers = __match_any_sync(live, ((unsigned long long)seg << 32) | (unsigned)lab);
    if (lane == __ffs(peers) - 1) {
        const int c = __popc(peers);
        atomicAdd(&seg_cnt[(int64_t)seg * n_labels + lab], c);
        atomicAdd(&seg_info[4 * seg + 0], c);
        atomicMin(&seg_info[4 * seg + 1], lab);
        atomicMax(&seg_info[4 * seg + 2], lab);
    }
}

__device__ __forceinline__ uint32_t block_max_u32(uint32_t v, uint32_t* s_w) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) s_w[warp] = v;
    __syncthreads();
    v = s_w[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(QT, 1)
auroc_count_kernel(const float* __restrict__ votes, int64_t ldv, int ncols, const float* __restrict__ denom,
                   const int32_t* __restrict__ col_denom, const int32_t* __restrict__ seg_off, int n_seg,
                   const int32_t* __restrict__ cell_label, int n_labels, const int32_t* __restrict__ seg_cnt,
                   const int32_t* __restrict__ seg_info, double* __restrict__ auroc, int32_t* __restrict__ n_pos_out,
                   uint32_t* slots, uint32_t* slots2, uint16_t* labs2, int64_t slot_stride, QTask* stacks, int stack_cap,
                   unsigned int* counter) {
    extern __shared__ __align__(16) uint32_t dyn[];
    uint32_t* coarse = dyn;                  // [QCB] counts -> inclusive cumulative counts of the task
    uint32_t* lut = coarse + QCB;            // [QCB] (first fine bin << 15) | number of fine bins, per coarse bin of the round
    uint32_t* fine = lut + QCB;              // [QF/2] two 16-bit counters per word
    uint32_t* gbase = fine + QF / 2;         // [QT + 4] slot of the first entry of each 16-bin group; big-bin list; round cursors
    uint32_t* s_w = gbase + QT + 4;          // [36] scan scratch
    uint32_t* tail = s_w + 36;               // accumulators | bitmap | sorted entries (sized per segment)
    __shared__ int s_item, s_ntask, s_nbig, s_e2;
    __shared__ uint32_t s_mn[32], s_mx[32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* slot = slots + (int64_t)blockIdx.x * slot_stride;
    uint32_t* slot2 = slots2 + (int64_t)blockIdx.x * slot_stride;
    uint16_t* lab2 = labs2 + (int64_t)blockIdx.x * slot_stride;
    QTask* stack = stacks + (int64_t)blockIdx.x * stack_cap;
    const int n_items = ncols * n_seg;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = (int)atomicAdd(counter, 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= n_items) break;
        const int col = item / n_seg, seg = item - col * n_seg;
        const int r0 = seg_off[seg], n = seg_off[seg + 1] - r0;
        const int nv = seg_info[4 * seg + 0];
        if (n <= 0 || nv <= 0) continue;
        const int lab_lo = seg_info[4 * seg + 1];
        const int nl = seg_info[4 * seg + 2] - lab_lo + 1;
        const int LB = 32 - __clz(nl - 1 > 0 ? nl - 1 : 0);       // bits of a segment-local label (0 when nl == 1)
        const uint32_t lmask = (1u << LB) - 1u;
        const int acc_words = (2 * nl + 3) & ~3;
        unsigned int* acc_lo = tail;
        unsigned int* acc_hi = tail + nl;
        // capacity of a round: 4 B per entry + 1 bit of the bin-start bitmap
        int cap = (int)(((long long)(QSMEM - QFIXED - 4 * acc_words - 64) * 8) / 33) & ~31;   // (8 spare entries after the last)
        if (cap > 65504) cap = 65504;        // 16-bit in-group prefixes
        uint32_t* bm = tail + acc_words;     // [(cap >> 5) + 4], bit of slot p at position p + 32
        uint32_t* sorted = bm + (cap >> 5) + 4;

        const float* v = votes + (int64_t)col * ldv + r0;
        const float* dn = nullptr;
        if (denom && col_denom && col_denom[col] >= 0) dn = denom + (int64_t)col_denom[col] * ldv + r0;
        const int32_t* lab_in = cell_label + r0;

        for (int i = tid; i < 2 * nl; i += QT) tail[i] = 0u;

        // ---- P0: keys, once; range of the kept keys ---------------------------------------------------
        uint32_t kmin = 0xffffffffu, kmax = 0u;
        for (int i0 = 0; i0 < n; i0 += 4 * QT) {
            float fv[4], fd[4];
            int fl[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * QT + tid;
                fl[u] = i < n ? lab_in[i] : -1;
                fv[u] = i < n ? v[i] : 0.f;
                fd[u] = (dn && i < n) ? dn[i] : 1.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * QT + tid;
                if (i >= n) continue;
                uint32_t k = QSENT;
                if (fl[u] >= 0 && fl[u] < n_labels) {
                    k = min(float_key(dn ? __fdiv_rn(fv[u], fd[u]) : fv[u]), 0xfffffffeu);
                    kmin = min(kmin, k);
                    kmax = max(kmax, k);
                }
                slot[i] = k;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        }
        if (lane == 0) {
            s_mn[warp] = kmin;
            s_mx[warp] = kmax;
        }
        __syncthreads();
        kmin = s_mn[lane];
        kmax = s_mx[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        }
        if (tid == 0) {
            stack[0] = QTask{kmin, kmax - kmin, 0u, (uint32_t)nv};
            s_ntask = 1;
        }
        bool top = true;                     // the first task is the whole segment

        // ---- tasks ---------------------------------------------------------------------------------------
        for (;;) {
            __syncthreads();                 // stack / s_ntask of thread 0, the keys in the slot
            const int nt = s_ntask;
            if (nt == 0) break;
            const QTask t = stack[nt - 1];
            __syncthreads();
            if (tid == 0) s_ntask = nt - 1;
            const bool is_top = top;
            top = false;
            if (t.range == 0u) {
                // one key value: a tie block of t.count keys above t.below smaller ones
                const uint32_t r2 = 2u * t.below + t.count + 1u;
                for (int i = tid; i < n; i += QT)
                    if (slot[i] == t.lo) acc_add32(acc_lo, acc_hi, (uint32_t)(lab_in[i] - lab_lo), r2);
                continue;
            }
            const int bits = 32 - __clz(t.range);
            const int cshift = bits > QCB_LOG ? bits - QCB_LOG : 0;
            const uint32_t cmask = (1u << cshift) - 1u;
            const int ncb = (int)(t.range >> cshift) + 1;          // <= QCB
            for (int i = tid; i < QCB; i += QT) coarse[i] = 0u;
            __syncthreads();
            // ---- P1: exact coarse histogram of the task's keys
            for (int i0 = 0; i0 < n; i0 += 8 * QT) {
                uint32_t k[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * QT + tid;
                    k[u] = i < n ? slot[i] : QSENT;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t c = k[u] - t.lo;
                    if (k[u] != QSENT && c <= t.range) atomicAdd(&coarse[c >> cshift], 1u);
                }
            }
            __syncthreads();
            uint32_t maxc;
            {
                const uint32_t a = coarse[2 * tid], b = coarse[2 * tid + 1];
                maxc = block_max_u32(max(a, b), s_w);
                uint32_t tot;
                const uint32_t ex = block_excl_scan(a + b, s_w, &tot);
                coarse[2 * tid] = ex + a;
                coarse[2 * tid + 1] = ex + a + b;
            }
            __syncthreads();

            // One ROUND: the keys of coarse bins [cb0, cb1) (m of them, rb kept keys below) taken from `src`
            // (ns keys; labels from lab16 when given, else from the segment's own label array).
            auto run_round = [&](const uint32_t* src, const uint16_t* lab16, int ns, int cb0, int cb1, uint32_t m, uint32_t rb) {
                const int ncbr = cb1 - cb0;
                // ---- LUT: fine bins per coarse bin in proportion to its count (>= 1, <= key values in the bin)
                {
                    const unsigned long long alpha = ((unsigned long long)(QF - ncbr) << 16) / m;
                    uint32_t nf[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cb = cb0 + 2 * tid + e;
                        nf[e] = 0u;
                        if (cb < cb1) {
                            const uint32_t cnt = coarse[cb] - (cb > 0 ? coarse[cb - 1] : 0u);
                            if (cnt) {
                                uint32_t x = (uint32_t)((cnt * alpha) >> 16);
                                x = x > 0u ? x : 1u;
                                if (cshift < 15 && x > (1u << cshift)) x = 1u << cshift;
                                nf[e] = x;
                            }
                        }
                    }
                    uint32_t tot;
                    const uint32_t ex = block_excl_scan(nf[0] + nf[1], s_w, &tot);
                    if (2 * tid < ncbr) lut[2 * tid] = (ex << 15) | nf[0];
                    if (2 * tid + 1 < ncbr) lut[2 * tid + 1] = ((ex + nf[0]) << 15) | nf[1];
                }
                {
                    uint4* f4 = reinterpret_cast<uint4*>(fine);
                    f4[2 * tid] = make_uint4(0u, 0u, 0u, 0u);
                    f4[2 * tid + 1] = make_uint4(0u, 0u, 0u, 0u);
                    for (int i = tid; i < (int)(m >> 5) + 4; i += QT) bm[i] = 0u;
                }
                __syncthreads();
                const int nshift = 32 - cshift;           // remainder -> 0.32 fixed point (cshift == 0: remainder is 0)
                // ---- P2: fine histogram
                for (int i0 = 0; i0 < ns; i0 += 8 * QT) {
                    uint32_t k[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + u * QT + tid;
                        k[u] = i < ns ? src[i] : QSENT;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t c = k[u] - t.lo;
                        const uint32_t cb = (c >> cshift) - (uint32_t)cb0;
                        if (k[u] != QSENT && c <= t.range && cb < (uint32_t)ncbr) {
                            const uint32_t l = lut[cb];
                            const uint32_t rem = c & cmask;
                            const uint32_t fb = (l >> 15) + (cshift ? __umulhi(rem << nshift, l & 0x7fffu) : 0u);
                            atomicAdd(&fine[fb >> 1], 1u << ((fb & 1u) << 4));
                        }
                    }
                }
                __syncthreads();
                // ---- scan: thread t owns fine bins [16t, 16t + 16); counts -> exclusive in-group prefixes
                {
                    uint4* f4 = reinterpret_cast<uint4*>(fine);
                    uint4 wa = f4[2 * tid], wb = f4[2 * tid + 1];
                    uint32_t w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                    uint32_t cnt[16], pre[16];
                    uint32_t run = 0u;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        cnt[2 * q] = w[q] & 0xffffu;
                        cnt[2 * q + 1] = w[q] >> 16;
                    }
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        pre[q] = run;
                        run += cnt[q];
                    }
                    uint32_t tot;
                    const uint32_t base = block_excl_scan(run, s_w, &tot);
                    gbase[tid] = base;
#pragma unroll
                    for (int q = 0; q < 8; ++q) w[q] = pre[2 * q] | (pre[2 * q + 1] << 16);
                    f4[2 * tid] = make_uint4(w[0], w[1], w[2], w[3]);
                    f4[2 * tid + 1] = make_uint4(w[4], w[5], w[6], w[7]);
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        if (cnt[q]) {
                            const uint32_t pos = base + pre[q] + 32u;       // (the bitmap is stored one word up)
                            atomicOr(&bm[pos >> 5], 1u << (pos & 31u));
                        }
                    }
                    if (tid == 0) {
                        atomicOr(&bm[(m >> 5) + 1], 1u << (m & 31u));   // end marker
                        s_nbig = 0;
                    }
                }
                __syncthreads();
                // ---- P3: scatter (key remainder | label) into bin order
                for (int i0 = 0; i0 < ns; i0 += 8 * QT) {
                    uint32_t k[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + u * QT + tid;
                        k[u] = i < ns ? src[i] : QSENT;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t c = k[u] - t.lo;
                        const uint32_t cb = (c >> cshift) - (uint32_t)cb0;
                        if (k[u] != QSENT && c <= t.range && cb < (uint32_t)ncbr) {
                            const uint32_t l = lut[cb];
                            const uint32_t rem = c & cmask;
                            const uint32_t fb = (l >> 15) + (cshift ? __umulhi(rem << nshift, l & 0x7fffu) : 0u);
                            const uint32_t sh = (fb & 1u) << 4;
                            const uint32_t old = atomicAdd(&fine[fb >> 1], 1u << sh);
                            const uint32_t pos = gbase[fb >> 4] + ((old >> sh) & 0xffffu);
                            const int i = i0 + u * QT + tid;
                            const uint32_t lab = lab16 ? (uint32_t)lab16[i] : (uint32_t)(lab_in[i] - lab_lo);
                            sorted[pos] = (rem << LB) | lab;
                        }
                    }
                }
                __syncthreads();
                // ---- P4: one thread per slot.  Bins hold ~2.6 entries: a bin of <= 8 has both ends inside the 32-bit
                // window of the bin-start bitmap that begins 8 slots below the slot (one funnel shift of two bitmap
                // words; the bitmap is stored one word up so that the window never starts below word 0), and its
                // entries are compared in two fixed, branch-free groups of four.  Longer bins (tie blocks) go to the
                // CTA-cooperative path below.
                for (int j = tid; j < (int)m; j += QT) {
                    const uint32_t e = sorted[j];
                    const int p0 = j + 24;                         // window = bitmap positions [j - 8, j + 24), shifted by 32
                    const uint32_t win = __funnelshift_r(bm[p0 >> 5], bm[(p0 >> 5) + 1], p0 & 31);
                    const uint32_t lo9 = win & 0x1ffu;             // starts at slots j - 8 .. j   (bit 8 = slot j)
                    const uint32_t hi9 = win & 0x3fe00u;           // starts at slots j + 1 .. j + 9
                    const int s = j - 8 + (31 - __clz((int)lo9));  // lo9 == 0: clz = 32 -> s = j - 9 (rejected below)
                    const int e2 = j - 8 + (__ffs((int)hi9) - 1);  // hi9 == 0: ffs = 0 -> e2 = j - 9 (rejected below)
                    const int cnt = e2 - s;
                    if (lo9 != 0u && hi9 != 0u && cnt <= QSMALL) {
                        const uint32_t kj = e >> LB;
                        uint32_t less = 0u, eq = 0u;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t kq = sorted[s + q] >> LB;       // (reads past the bin stay inside `sorted`)
                            const bool in = q < cnt;
                            less += (in && kq < kj) ? 1u : 0u;
                            eq += (in && kq == kj) ? 1u : 0u;
                        }
                        if (cnt > 4) {
#pragma unroll
                            for (int q = 4; q < 8; ++q) {
                                const uint32_t kq = sorted[s + q] >> LB;
                                const bool in = q < cnt;
                                less += (in && kq < kj) ? 1u : 0u;
                                eq += (in && kq == kj) ? 1u : 0u;
                            }
                        }
                        acc_add32(acc_lo, acc_hi, e & lmask, 2u * (rb + (uint32_t)s) + 2u * less + eq + 1u);
                    } else {
                        // a longer bin (ties): look up to two bitmap words either side -- enough for <= 64 entries
                        const int w0 = (j >> 5) + 1, b = j & 31;
                        const uint32_t x = bm[w0];
                        const uint32_t lowmask = 0xffffffffu >> (31 - b);
                        int s2 = -1, e3 = -1;
                        {
                            uint32_t xs = x & lowmask;
                            int ws = w0;
                            if (!xs && ws > 1) xs = bm[--ws];
                            if (!xs && ws > 1) xs = bm[--ws];
                            if (xs) s2 = ((ws - 1) << 5) + 31 - __clz(xs);
                        }
                        {
                            uint32_t xe = x & ~lowmask;
                            int we = w0;
                            const int wlast = (int)(m >> 5) + 1;
                            if (!xe && we < wlast) xe = bm[++we];
                            if (!xe && we < wlast) xe = bm[++we];
                            if (xe) e3 = ((we - 1) << 5) + __ffs(xe) - 1;
                        }
                        if (s2 >= 0 && e3 >= 0 && e3 - s2 <= QBIG) {
                            const uint32_t kj = e >> LB;
                            uint32_t less = 0u, eq = 0u;
                            for (int q = s2; q < e3; ++q) {
                                const uint32_t kq = sorted[q] >> LB;
                                less += kq < kj ? 1u : 0u;
                                eq += kq == kj ? 1u : 0u;
                            }
                            acc_add32(acc_lo, acc_hi, e & lmask, 2u * (rb + (uint32_t)s2) + 2u * less + eq + 1u);
                        } else if ((win >> 8) & 1u) {
                            lut[atomicAdd(&s_nbig, 1)] = (uint32_t)j;      // first slot of a big bin (> 64 entries: <= m/65
                                                                           // < QCB of them; the LUT is free once P3 is done)
                        }
                    }
                }
                __syncthreads();
                // ---- big bins (heavy ties): the whole CTA per bin
                const int nbig = s_nbig;
                for (int bi = 0; bi < nbig; ++bi) {
                    const int s = (int)lut[bi];
                    if (tid == 0) {
                        int we = (s >> 5) + 1;
                        uint32_t xe = bm[we] & ~(0xffffffffu >> (31 - (s & 31)));
                        while (!xe) xe = bm[++we];
                        s_e2 = ((we - 1) << 5) + __ffs(xe) - 1;
                    }
                    __syncthreads();
                    const int e2 = s_e2;
                    const uint32_t k0 = sorted[s] >> LB;
                    int differs = 0;
                    for (int j = s + tid; j < e2; j += QT) differs |= (sorted[j] >> LB) != k0;
                    differs = __syncthreads_or(differs);
                    if (!differs) {
                        const uint32_t r2 = 2u * (rb + (uint32_t)s) + (uint32_t)(e2 - s) + 1u;
                        for (int j = s + tid; j < e2; j += QT) acc_add32(acc_lo, acc_hi, sorted[j] & lmask, r2);
                    } else {
                        for (int j = s + tid; j < e2; j += QT) {
                            const uint32_t e = sorted[j];
                            const uint32_t kj = e >> LB;
                            uint32_t less = 0u, eq = 0u;
                            for (int q = s; q < e2; ++q) {
                                const uint32_t kq = sorted[q] >> LB;
                                less += kq < kj ? 1u : 0u;
                                eq += kq == kj ? 1u : 0u;
                            }
                            acc_add32(acc_lo, acc_hi, e & lmask, 2u * (rb + (uint32_t)s) + 2u * less + eq + 1u);
                        }
                    }
                }
                __syncthreads();
            };

            if (t.count <= (uint32_t)cap) {
                // everything fits one round
                run_round(slot, nullptr, n, 0, ncb, t.count, t.below);
                continue;
            }
            const uint32_t capq = (uint32_t)cap - maxc;
            // (two or three rounds re-read the slot with a range filter instead: the second slot would double the
            // CTA's footprint in L2 -- measured at config 4, 71k keys in 2 rounds: 9.4 ms partitioned, 8.1 filtered)
            if (is_top && t.count > 3u * (uint32_t)cap && 2u * maxc <= (uint32_t)cap &&
                (t.count + capq - 1u) / capq + 1u <= (uint32_t)QT) {
                // ---- PARTITION: round r = the coarse bins whose cumulative count starts in [r capq, (r + 1) capq):
                // at most capq + maxc <= cap keys.  One sweep deals the keys (and their labels) to the rounds'
                // contiguous ranges of the second slot; every round then streams only its own keys.
                uint16_t* rid16 = reinterpret_cast<uint16_t*>(fine);      // [QCB] round of each coarse bin
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int cb = 2 * tid + e;
                    const uint32_t ce = cb > 0 ? coarse[cb - 1] : 0u;
                    rid16[cb] = (uint16_t)(cb < ncb ? ce / capq : 0u);
                }
                const int n_rounds = (int)((coarse[ncb - 1] - 1u) / capq) + 1;      // (an empty top bin cannot exist: kmax is in it)
                __syncthreads();
                // first coarse bin of round r: smallest cb with cum_excl(cb) >= r * capq
                auto round_begin = [&](int r) {
                    if (r >= n_rounds) return ncb;
                    const uint32_t want = (uint32_t)r * capq;
                    int lo_ = 0, hi_ = ncb - 1;
                    while (lo_ < hi_) {
                        const int mid = (lo_ + hi_) >> 1;
                        if ((mid > 0 ? coarse[mid - 1] : 0u) >= want) hi_ = mid; else lo_ = mid + 1;
                    }
                    return lo_;
                };
                if (tid < n_rounds) {
                    const int cbA = round_begin(tid);
                    gbase[tid] = cbA > 0 ? coarse[cbA - 1] : 0u;          // cursor = offset of the round's range
                }
                __syncthreads();
                for (int i0 = 0; i0 < n; i0 += 8 * QT) {
                    uint32_t k[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + u * QT + tid;
                        k[u] = i < n ? slot[i] : QSENT;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (k[u] != QSENT) {
                            const uint32_t r = rid16[(k[u] - t.lo) >> cshift];
                            const uint32_t pos = atomicAdd(&gbase[r], 1u);
                            const int i = i0 + u * QT + tid;
                            slot2[pos] = k[u];
                            lab2[pos] = (uint16_t)(lab_in[i] - lab_lo);
                        }
                    }
                }
                __syncthreads();
                for (int r = 0; r < n_rounds; ++r) {
                    const int cbA = round_begin(r), cbB = round_begin(r + 1);
                    const uint32_t off = cbA > 0 ? coarse[cbA - 1] : 0u;
                    const uint32_t m = coarse[cbB - 1] - off;
                    if (m == 0u) continue;
                    run_round(slot2 + off, lab2 + off, (int)m, cbA, cbB, m, t.below + off);
                }
                continue;
            }

            // ---- greedy rounds over the slot with a range filter; a coarse bin above the capacity becomes a task
            int cb0 = 0;
            uint32_t before = 0u;
            while (cb0 < ncb) {
                int lo_ = cb0, hi_ = ncb;    // largest cb1 with cum(cb1) - before <= cap
                while (lo_ < hi_) {
                    const int mid = (lo_ + hi_ + 1) >> 1;
                    if (coarse[mid - 1] - before <= (uint32_t)cap) lo_ = mid; else hi_ = mid - 1;
                }
                const int cb1 = lo_;
                if (cb1 == cb0) {
                    const uint32_t cnt = coarse[cb0] - before;
                    const uint32_t off = (uint32_t)cb0 << cshift;
                    if (tid == 0) {
                        stack[s_ntask] = QTask{t.lo + off, min(cmask, t.range - off), t.below + before, cnt};
                        s_ntask = s_ntask + 1;
                    }
                    before = coarse[cb0];
                    ++cb0;
                    continue;
                }
                const uint32_t m = coarse[cb1 - 1] - before;
                if (m > 0u) run_round(slot, nullptr, n, cb0, cb1, m, t.below + before);
                before = coarse[cb1 - 1];
                cb0 = cb1;
            }
        }

        // ---- AUROC of every test label of the segment, in the reference's operation order (utils.py:175-178)
        for (int c = tid; c < nl; c += QT) {
            const int np = seg_cnt[(int64_t)seg * n_labels + lab_lo + c];
            if (np == 0) continue;
            const double n_p = (double)np, n_n = (double)(nv - np);
            const unsigned long long sum2 = ((unsigned long long)acc_hi[c] << 32) | acc_lo[c];
            double roc = ((double)sum2 * 0.5) / n_p;
            roc -= (n_p + 1.0) / 2.0;
            roc /= n_n;
            auroc[(int64_t)(lab_lo + c) * ncols + col] = roc;
            if (col == 0 && n_pos_out) n_pos_out[lab_lo + c] = np;
        }
    }
}

struct CountLayout {
    int64_t slot_stride, stack_cap;
    int64_t off_counter, off_seg_info, off_seg_cnt, off_stacks, off_slots, off_slots2, off_labs2, total;
    int grid;
};

CountLayout count_layout(int ncols, int64_t M, int n_seg, int n_labels, int sms) {
    CountLayout L;
    const int64_t items = (int64_t)ncols * n_seg;
    L.grid = (int)(items < sms ? items : sms);
    L.slot_stride = ((M + 63) / 64) * 64;
    L.stack_cap = M / 32768 + 8;             // a pending task holds more than one round (> 39k) of distinct keys
    int64_t o = 0;
    L.off_counter = o;
    o += 256;
    L.off_seg_info = o;
    o += (((int64_t)n_seg * 16 + 255) / 256) * 256;
    L.off_seg_cnt = o;
    o += (((int64_t)n_seg * n_labels * 4 + 255) / 256) * 256;
    L.off_stacks = o;
    o += (((int64_t)sms * L.stack_cap * (int64_t)sizeof(QTask) + 255) / 256) * 256;
    L.off_slots = o;
    o += (int64_t)sms * L.slot_stride * 4;
    L.off_slots2 = o;
    o += (int64_t)sms * L.slot_stride * 4;
    L.off_labs2 = o;
    o += (int64_t)sms * L.slot_stride * 2;
    L.total = o;
    return L;
}

int device_sms() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace

int64_t auroc_count_workspace_bytes(int ncols, int64_t M, int n_seg, int n_labels) {
    return count_layout(ncols, M, n_seg, n_labels, device_sms()).total;
}

int launch_auroc_count(const float* votes, int64_t ldv, int ncols, const float* denom, const int32_t* col_denom,
                       const int32_t* seg_off, int n_seg, const int32_t* cell_label, int n_labels, int64_t M,
                       double* auroc, int32_t* n_pos, void* workspace, cudaStream_t st) {
    const CountLayout L = count_layout(ncols, M, n_seg, n_labels, device_sms());
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    unsigned int* counter = reinterpret_cast<unsigned int*>(w + L.off_counter);
    int32_t* seg_info = reinterpret_cast<int32_t*>(w + L.off_seg_info);
    int32_t* seg_cnt = reinterpret_cast<int32_t*>(w + L.off_seg_cnt);
    QTask* stacks = reinterpret_cast<QTask*>(w + L.off_stacks);
    uint32_t* slots = reinterpret_cast<uint32_t*>(w + L.off_slots);
    uint32_t* slots2 = reinterpret_cast<uint32_t*>(w + L.off_slots2);
    uint16_t* labs2 = reinterpret_cast<uint16_t*>(w + L.off_labs2);
    const int64_t n_auroc = (int64_t)n_labels * ncols, n_seg_cnt = (int64_t)n_seg * n_labels;
    const int64_t most = n_auroc > n_seg_cnt ? n_auroc : n_seg_cnt;
    int init_blocks = (int)((most + 255) / 256);
    if (init_blocks > 1184) init_blocks = 1184;
    if (init_blocks < 1) init_blocks = 1;
    auroc_init_kernel<<<init_blocks, 256, 0, st>>>(auroc, n_auroc, n_pos, n_labels, seg_cnt, n_seg_cnt, seg_info, n_seg,
                                                  counter);
    MN_LAUNCH_CHECK();
    auroc_label_count_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(seg_off, n_seg, cell_label, n_labels, M, seg_cnt,
                                                                         seg_info);
    MN_LAUNCH_CHECK();
    // (set on every call: a process may drive several devices, and the attribute is per device)
    MN_CUDA(cudaFuncSetAttribute(auroc_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, QSMEM));
    auroc_count_kernel<<<L.grid, QT, QSMEM, st>>>(votes, ldv, ncols, denom, col_denom, seg_off, n_seg, cell_label, n_labels,
                                                 seg_cnt, seg_info, auroc, n_pos, slots, slots2, labs2, L.slot_stride, stacks,
                                                 (int)L.stack_cap, counter);
    MN_LAUNCH_CHECK();
    return 0;
}

}  // namespace mn
