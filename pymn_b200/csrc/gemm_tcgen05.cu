// K3 / K4 / K5: tcgen05 tensor-core contraction with fused epilogues (sm_100a).
//
//   D[m, n] = sum_g A[m, g] * B[n, g]        A, B: K-major fp16 planes, D: FP32 in TMEM
//
// Operands are the integer-valued centred doubled ranks d = 2*rank - (G+1)
// written by K1 as fp16 planes (exact: one plane for G <= 2048, hi+lo beyond),
// so the tensor cores multiply EXACT integers and only the FP32 accumulation
// rounds; the 1/||d|| scaling happens in the epilogue.  (The reference computes
// the same quantity as np.corrcoef of the ranks, utils.py:71, or as dot
// products of normalised cells, MetaNeighborUS.py:361.)  Multi-plane products
// (hi*hi + hi*lo + lo*hi) are extra accumulation passes into the same TMEM
// accumulator -- the fp16 analogue of a 3xTF32 split, but with exact limbs.
//
// Kernel anatomy (persistent, one CTA per SM, 320 threads; the two network passes run as CTA PAIRS --
// clusters of two CTAs, tcgen05 cta_group::2, M = 256 -- see gemm_kernel):
//   warp 0      TMA producer: the (cluster) leader claims tiles from a global atomic counter (dynamic
//               scheduler) and publishes them to every role; cp.async.bulk.tensor 2D tiles
//               (128B swizzle) -> smem ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, M = 128 or 256, N = BN,
//               K = 16), tcgen05.commit (multicast to both CTAs of a pair) -> mbarriers
//   warps 2..9  epilogue: tcgen05.ld (32 lanes x 32 columns per warp; two warps per TMEM lane
//               quarter, each half of the columns) from one of two TMEM accumulator stages while
//               the next tile's MMAs run
// Epilogues:
//   EPI_STORE  votes_fast: scale + addend, store column-major FP32   (MetaNeighborUS.py:361-366)
//   EPI_HIST   pass 1 of the global rank: rho -> shared-memory histogram (utils.py:44-45), and the
//              16.16 fixed-point bin position of every pair -> the rho spill in HBM
//   EPI_VOTE   pass 2 for tiles that were not spilled: rho -> W via CDF LUT -> integer per-label vote
//              sums + node degree (utils.py:46-51 + MetaNeighborUS.py:165-169)
// vote_spill_kernel is pass 2 for the spilled tiles: it streams the bin positions back from HBM.
// The float64 N x N network of the reference never exists.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace mn {

constexpr int BM = 128;            // rows per tile = TMEM lanes
constexpr int BK = 64;             // fp16 elements per k-block = 128 bytes = one swizzle atom
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;             // two warps per TMEM lane quarter, each takes half of the tile columns
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int GEMM_THREADS = 64 + EPI_THREADS;
constexpr int MAX_STAGES = 6;
constexpr int MAX_TERMS = 8;
constexpr int TQ = 4;                     // depth of the tile-index queue of the dynamic scheduler

enum { EPI_STORE = 0, EPI_HIST = 1, EPI_VOTE = 2 };

struct GemmParams {
    int64_t M;        // rows of A
    int64_t N;        // rows of B
    int k_blocks;     // ceil(G / 64)
    int k_last;       // MMA steps (of 16 genes) the last k-block needs: ceil((G - 64 (k_blocks - 1)) / 16)
    int n_terms;      // plane products accumulated into one tile
    int a_sel[MAX_TERMS];   // 0 = hi plane, 1 = lo plane          (int8 form: 0 = limb a1, 1 = limb a0)
    int b_sel[MAX_TERMS];   //                                     (int8 form: 0 .. 3 = limb b3 .. b0)
    // int8 form (exact integer dot products, mn_votes_fast_i8): every limb product has a weight 128^cls and is
    // accumulated into the int32 accumulator of its class (the "chain" slots of TMEM); the limb planes of B are
    // stacked in ONE tensor map, limb l at rows [l * b_limb_rows, ...)
    int cls[MAX_TERMS];
    int cls_first[MAX_TERMS];     // 1 = first term of its class: its first MMA overwrites the accumulator
    int b_limb_rows;
    int cls_mul;                  // weight of accumulator slot 0 (128 with two limbs of A, 1 with one)
    int n_cls;                    // accumulator slots in use (4; 1 for the lone a0 * b0 product of mn_votes_den_i8)
    int acc64;                    // 1: out64 += val instead of out64 = val
    // peer mode (multi-GPU centroid paths, mn_votes_fast_i8_peer): column n belongs to rank n / peer_cols and is
    // stored straight into THAT rank's receive buffer over NVLink, at [peer_src][n % peer_cols][m] (pitch peer_per)
    float* peer_out[8];
    int peer_world;               // 0 = off
    int peer_src;
    int peer_cols;
    int64_t peer_per;
    const double* inv_a64;        // [M] exact-ish row scale 1/||d|| (double)
    // ... its epilogue, all in double:  val = dot * inv_a64[m] * inv_b[n] + addend[n];
    //   den != null:  val = val / den[col_den[n] * ld_den + m] * col_scale[n] - 1     (node-degree ratio, centred)
    //   else       :  val = val - col_scale[n]                                         (col_scale may be null = 0)
    //   out64 != null: val is stored as double there, else rounded once to float32 in out
    double* out64;
    const double* den;
    int64_t ld_den;
    const int32_t* col_den;
    const double* col_scale;
    int tri;          // 1: only tiles that contain some column index > row index
    int mt0, mt1;     // row-tile range handled by this launch
    int mt_step;      // ... of which every mt_step-th, starting at mt0 (row tiles dealt round-robin to GPUs)
    int n_tiles;      // number of column tiles
    int band;         // row tiles per L2 band of the tile schedule
    unsigned int* tile_counter;   // dynamic tile scheduler: next linear tile index (zeroed before the launch)
    int total_tiles;
    int stages;
    // epilogue
    const float* inv_a;      // [M] row scale (1/||d||), 0 = dropped cell
    const float* inv_b;      // [N] column scale
    const float* addend;     // [N] or null            (EPI_STORE)
    float* out;              // column-major [N, ldo]  (EPI_STORE)
    int64_t ldo;
    int nbins;               // power of two           (EPI_HIST / EPI_VOTE)
    unsigned long long* hist;      // [nbins]          (EPI_HIST)
    const uint32_t* lut;     // [nbins + 1]            (EPI_VOTE)
    const int32_t* label;    // [N]                    (EPI_VOTE)
    int L;
    unsigned long long* votes;     // [M, L]
    unsigned long long* degree;    // [M]
    // rho spill (EPI_HIST writes, vote_spill_kernel reads): tile t < spill_tiles of the linear tile
    // order keeps its BM x BN fixed-point bin positions in HBM so that pass 2 need not recompute them
    uint32_t* spill;
    int spill_tiles;
    int tile_base;           // first linear tile index of this launch (pass 2 over the tiles that were not spilled)
    int ctas;                // 1, or 2 = CTA pairs (cta_group::2): a tile is 256 rows, mt counts 256-row tiles,
                             //    CTA r of the pair owns rows [128 r, 128 r + 128) of it
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok != 0;
}
// Spin until the barrier phase with the given parity completes.  A wait of more than 2^30 polls (each
// try_wait suspends the thread for a hardware-bounded time: tens of seconds in all) can only be a protocol bug:
// trap instead of hanging the GPU.  Polls are counted, not clock cycles, so that time spent preempted
// (time-slicing, a debugger) never trips it.
constexpr uint32_t MBAR_SPIN_LIMIT = 1u << 30;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    if (mbar_try_wait(addr, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try_wait(addr, parity)) {
        if (++spins > MBAR_SPIN_LIMIT) {
            printf("pymn_b200 gemm: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// ---- CTA-pair (cta_group::2) variants: the two CTAs of a cluster sit on the two SMs of one TPC
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_remote_u32(uint32_t cluster_addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok != 0;
}
// wait on a barrier that a thread of the peer CTA arrives on after writing data into this CTA's shared memory
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    if (mbar_try_wait_cluster(addr, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(addr, parity)) {
        if (++spins > MBAR_SPIN_LIMIT) {
            printf("pymn_b200 gemm: cluster mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}
// TMA load issued by either CTA of a pair; the transaction bytes are counted on the LEADER CTA's barrier
// (same shared-memory offset, peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
// ... and the same tile delivered to several CTAs of the cluster (bit r of cta_mask = CTA rank r): one L2 read,
// the bytes are counted on the barrier at this offset in the pair leader of every destination
__device__ __forceinline__ void tma_load_2d_pair_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                    uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "h"(cta_mask),
          "r"(c0), "r"(c1)
        : "memory");
}
// completion of all prior MMAs of the pair -> arrive on the barrier at this offset in the CTAs of cta_mask
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma_i8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 columns of this thread's row: sum of the NCHAIN partial accumulators (chain c lives BN columns further)
template <int BN, int NCHAIN>
__device__ __forceinline__ void tmem_ld_chains(uint32_t taddr, int n_chains, uint32_t (&r)[32]) {
    tmem_ld32(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int c = 1; c < NCHAIN; ++c) {
        if (c < n_chains) {
            uint32_t t[32];
            tmem_ld32(taddr + (uint32_t)(c * BN), t);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(t[j]));
        }
    }
}
// 64-bit integer reduction to global memory whose L2 line is marked evict-first: the column-side vote
// sums stream through the whole votes array once per band and must not push the operand tiles out of L2
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void red_add_u64_hint(unsigned long long* addr, unsigned long long val, uint64_t pol) {
    asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(addr), "l"(val), "l"(pol) : "memory");
}
template <int THREADS = EPI_THREADS>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }

// K-major, 128B-swizzled operand tile: 8-row groups of 1024 bytes (SBO), LBO unused.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);   // start address, 16-byte units
    d |= (uint64_t)1 << 16;                        // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
// kind::f16, A=B=F16, D=F32, both K-major
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// kind::i8, A=B=signed 8 bit, D=S32, both K-major (K = 32 per instruction)
__device__ __forceinline__ uint32_t make_idesc_i8(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Position of rho on the bin axis; pass 1 and pass 2 must agree bit for bit.
//   pos = rho * bin_scale + bin_offset,  bin_scale = nbins/2 - 0.81,  bin_offset = nbins/2 + 0.043
// so rho in [-1, 1] maps into (0.85, nbins - 0.76).  The odd constants keep "nice" correlations off the
// bin edges: with pos = (rho + 1) * nbins / 2 every rho = k / 2^m sat exactly ON an edge, and the FP32
// rounding of inv_i * inv_j dealt the members of one tie block (tiny gene sets: thousands of cell pairs
// share rho = 1/2) to the two bins around it -- half of the block then ranked a whole bin too low.
__host__ __device__ __forceinline__ float bin_scale(int nbins) { return 0.5f * (float)nbins - 0.81f; }
__host__ __device__ __forceinline__ float bin_offset(int nbins) { return 0.5f * (float)nbins + 0.043f; }
__device__ __forceinline__ float rho_binpos(float acc, float inv_i_s, float inv_j, float offset) {
    return fmaf(acc * inv_i_s, inv_j, offset);
}
// ... as 16.16 fixed point: bin = q >> 16, position inside the bin = q & 0xffff.  NaN (a dropped
// cell's scale) and negative values convert to 0.  This integer is what pass 1 bins, what the rho
// spill stores and what pass 2 maps to a weight, so the three can never disagree.
constexpr int Q_FRAC_BITS = 16;
__device__ __forceinline__ uint32_t rho_q(float acc, float inv_i_s, float inv_j, float offset, uint32_t q_top) {
    return min(__float2uint_rd(rho_binpos(acc, inv_i_s, inv_j, offset) * (float)(1 << Q_FRAC_BITS)), q_top);
}
// The LUT (corr_lut_kernel) holds one packed word per bin: (W0 << 11) | rise.
//   rise > 0: W0 = weight at the bin's lower edge, the weight rises linearly by `rise` across the bin;
//   rise = 0: the bin is ONE TIE BLOCK (it holds >= 0.2 % of all pairs -- only heavily tied correlations of
//             tiny gene sets do that) and W0 is the block's mid-rank weight, as bottleneck.rankdata gives a
//             tie block (utils.py:45).
// One shared-memory load per pair either way.
__device__ __forceinline__ uint32_t w_from_q(uint32_t q, const uint32_t* tab) {
    const uint32_t e = tab[q >> Q_FRAC_BITS];
    return (e >> MN_LUT_DELTA_BITS) + (((e & MN_LUT_DELTA_MAX) * (q & ((1u << Q_FRAC_BITS) - 1u))) >> Q_FRAC_BITS);
}

// ------------------------------------------------------------------ pass-2 vote accumulation
// One thread = one row (cell i) of a tile; it walks the tile's columns in chunks of 32 fixed-point bin
// positions q_ij (from the tensor core or from the rho spill).
//   row side:    votes[i, label(j)] += W_ij as a per-thread running sum over the label runs of the
//                columns (cells are sorted by label, so a tile holds 1-3 runs), one 64-bit atomic per run;
//   column side: votes[j, label(i)] += W_ij through a 32 x 32 shuffle transpose-reduce per warp (rows
//                of a warp almost always share a label); per-warp column sums land in shared memory
//                (colw) and are merged across the row-quarter warps by vote_merge_columns.
struct VoteRows {
    unsigned long long run_acc;
    unsigned long long* vrow;
    int cur_lab, lab_i;
    unsigned gmask;
    bool warp_uniform, grp_leader;

    __device__ __forceinline__ void begin(int lab_i_, unsigned long long* vrow_, int lane) {
        run_acc = 0ull;
        cur_lab = 0;
        lab_i = lab_i_;
        vrow = vrow_;
        gmask = __match_any_sync(0xffffffffu, lab_i);
        warp_uniform = (gmask == 0xffffffffu);
        grp_leader = (gmask & lanemask_lt()) == 0;
    }
    __device__ __forceinline__ void flush() {
        if (run_acc) {
            atomicAdd(vrow + cur_lab, run_acc);
            run_acc = 0ull;
        }
    }
    // r: 32 bin positions in, weights out.  brk: bit c set = a label run starts at column c of the chunk;
    // labs.at(c) = label of column c of the chunk.  vcol = &votes[(first column of the chunk) * L].
    // Returns in lane l the sum of column l over the warp's 32 rows when the rows share one label
    // (otherwise 0: the mixed-label columns have been flushed to global memory already).
    template <typename LabSrc>
    __device__ __forceinline__ uint32_t chunk(uint32_t (&r)[32], unsigned brk, const LabSrc labs,
                                              const uint32_t* s_table, bool sym,
                                              unsigned long long* vcol, int L, uint64_t pol_ef, int lane) {
        uint32_t mycol = 0u;
        if (brk & 1u) {                                  // warp-uniform: a label run starts with this chunk
            flush();
            cur_lab = labs.at(0);
        }
        if ((brk >> 1) == 0u) {
            // fast path (almost every chunk): one label run -> branch-free
            uint32_t acc32 = 0u;                         // 32 weights <= 2^20 each fit 32 bits
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const uint32_t w = w_from_q(r[c], s_table);
                acc32 += w;
                r[c] = w;
            }
            run_acc += acc32;
        } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                if (c > 0 && (brk & (1u << c))) {
                    flush();
                    cur_lab = labs.at(c);
                }
                const uint32_t w = w_from_q(r[c], s_table);
                run_acc += w;
                r[c] = w;
            }
        }
        if (sym) {
            if (warp_uniform) {
                // 32 x 32 transpose-reduce: 31 shuffles leave in lane l the sum over the warp's
                // 32 rows of column l of the chunk  (<= 32 * 2^20, fits 32 bits).  (One REDUX.SUM per
                // column instead measured slower: pass 2 at config 2 28.6 vs 26.4 ms.)
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const bool up = (lane & off) != 0;
#pragma unroll
                    for (int c = 0; c < off; ++c) {
                        const uint32_t send = up ? r[c] : r[c + off];
                        const uint32_t keep = up ? r[c + off] : r[c];
                        r[c] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                mycol = r[0];
            } else {                                     // rows of several labels in one warp (rare)
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const uint32_t s = __reduce_add_sync(gmask, r[c]);
                    if (grp_leader && s) red_add_u64_hint(vcol + (int64_t)c * L + lab_i, (unsigned long long)s, pol_ef);
                }
            }
        }
        return mycol;
    }
    __device__ __forceinline__ void finish() { flush(); }
};
// column labels of a chunk: staged in shared memory / one per lane in a register
struct SmemLabels {
    const int32_t* p;
    __device__ __forceinline__ int at(int c) const { return p[c]; }
};
struct LaneLabels {
    int v;
    __device__ __forceinline__ int at(int c) const { return __shfl_sync(0xffffffffu, v, c); }
};

// Column side of one tile: s_colw [EPI_WARPS][BN/2] per-warp column sums (warps 0-3 = first half of the
// columns, 4-7 = second half; s_wlab[w] = the row label of warp w, or -2 if its rows were mixed and
// already flushed).  Call after a barrier over the 8 warps.
template <int BN>
__device__ __forceinline__ void vote_merge_columns(const uint32_t* s_colw, const int32_t* s_wlab,
                                                   unsigned long long* votes, int64_t n0, int L, uint64_t pol_ef,
                                                   int ew, int lane) {
    constexpr int HALF_COLS = BN / 2;
    const int et = ew * 32 + lane;
    const int lab0 = s_wlab[0];
    bool tile_uniform = lab0 >= 0;
#pragma unroll
    for (int w8 = 1; w8 < EPI_WARPS; ++w8) tile_uniform = tile_uniform && (s_wlab[w8] == lab0);
    if (tile_uniform) {
        for (int c = et; c < BN; c += EPI_THREADS) {
            const uint32_t* src = s_colw + (c / HALF_COLS) * 4 * HALF_COLS + (c % HALF_COLS);
            const unsigned long long sum = (unsigned long long)src[0] + src[HALF_COLS] +
                                           src[2 * HALF_COLS] + src[3 * HALF_COLS];
            if (sum) red_add_u64_hint(votes + (n0 + c) * (int64_t)L + lab0, sum, pol_ef);
        }
    } else {
        const int my_lab = s_wlab[ew];
        if (my_lab >= 0) {
            const uint32_t* src = s_colw + ew * HALF_COLS;
            const int cbeg = (ew >> 2) * HALF_COLS;
            for (int k = lane; k < HALF_COLS; k += 32) {
                const uint32_t v = src[k];
                if (v) red_add_u64_hint(votes + (n0 + cbeg + k) * (int64_t)L + my_lab, (unsigned long long)v, pol_ef);
            }
        }
    }
}

// ------------------------------------------------------------------ tile schedule
// Row tiles owned by this launch: mt(k) = mt0 + k * mt_step, k in [0, n_my).  They are processed in
// BANDS of `band` row tiles; inside a band the order is column tile outer, row tile inner.  The CTAs
// of the persistent grid therefore work on a few neighbouring column tiles of one band at a time:
// the band's A tiles (band * 128 rows) stay in L2 and every B tile is fetched from HBM once per band
// instead of once per row tile.  In `tri` mode only tiles holding some column index > some row index
// exist:  mt * BM < (nt + 1) * bn.
struct TileIter {
    int mt0, mt_step, n_my, band, n_tiles, tri, bn, bm;
    int k0, nt_, r, mt;
    bool ok;
    __device__ __forceinline__ int mt_of(int k) const { return mt0 + k * mt_step; }
    __device__ __forceinline__ int rows_in_band() const {
        const int v = n_my - k0;
        return v < band ? v : band;
    }
    __device__ __forceinline__ int first_nt_of_band() const { return tri ? (mt_of(k0) * bm) / bn : 0; }
    __device__ __forceinline__ int col_len() const {
        const int rows = rows_in_band();
        if (!tri) return rows;
        const int max_mt = ((nt_ + 1) * bn - 1) / bm;
        if (max_mt < mt0) return 0;
        const int cnt = (max_mt - mt0) / mt_step + 1 - k0;
        return cnt < 0 ? 0 : (cnt < rows ? cnt : rows);
    }
    __device__ __forceinline__ void init(const GemmParams& p, int bn_, int start) {
        mt0 = p.mt0; mt_step = p.mt_step; band = p.band; n_tiles = p.n_tiles; tri = p.tri; bn = bn_;
        bm = BM * p.ctas;
        n_my = p.mt1 > p.mt0 ? (p.mt1 - p.mt0 + p.mt_step - 1) / p.mt_step : 0;
        k0 = 0; r = 0; mt = mt0;
        ok = n_my > 0 && n_tiles > 0;
        if (ok) {
            nt_ = first_nt_of_band();
            advance(start);
        }
    }
    __device__ __forceinline__ void advance(int step) {
        if (!ok) return;
        int rem = r + step;
        for (;;) {
            const int len = col_len();
            if (rem < len) {
                r = rem;
                mt = mt_of(k0 + r);
                return;
            }
            rem -= len;
            if (++nt_ >= n_tiles) {
                k0 += band;
                if (k0 >= n_my) {
                    ok = false;
                    return;
                }
                nt_ = first_nt_of_band();
            }
        }
    }
    __device__ __forceinline__ bool valid() const { return ok; }
    __device__ __forceinline__ int nt() const { return nt_; }
};

// ------------------------------------------------------------------ the kernel
// CTAS = 2: CTA pairs (clusters of two CTAs on the two SMs of a TPC) work on 256 x BN tiles with
// tcgen05.mma.cta_group::2 (M = 256).  CTA r loads its own 128 rows of A and only HALF of the B tile
// (BN / 2 rows); the tensor cores of both SMs read both halves.  Per SM and k-block that is 32 KB
// written + 32 KB read of shared memory instead of 48 + 48 KB -- the single-CTA form needs 96 B/clk of
// the 128 B/clk shared-memory bandwidth for the MMA reads alone and tops out near 2/3 of the tensor
// peak once TMA writes and the epilogue's lookups share the same pipe.  Only the leader (rank 0)
// claims tiles and issues MMAs; completion is multicast to the barriers of both CTAs.
// EW = epilogue warps: 8 (two per TMEM lane quarter, each half of the tile's columns), or 16 for the pass-1
// epilogue (four per quarter, a quarter of the columns each; opt-in MN_HIST_EW16).  Measured at config 2:
// 16 warps are no faster (73.0 vs 70.8 ms), and dropping the histogram atomics altogether changes nothing
// (72.8 vs 72.9 ms): the epilogue is not what paces pass 1 -- the power cap is.
template <int BN, int EPI, int NCHAIN, int CTAS, int EW = EPI_WARPS, int I8 = 0>
__global__ void __launch_bounds__(64 + EW * 32, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
            const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
            const GemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // CTAS = 4 (opt-in, MN_GEMM_CTAS=4): two such pairs stacked in M form one cluster and work on a 512 x BN
    // tile.  Both pairs need the same B tile: pair 0 loads it ONCE with TMA multicast into the shared memory of
    // both pairs (CTA r of pair 0 -> CTAs r and r + 2), 768 instead of 1,024 operand rows per 512 x 256 outputs.
    // Correct (same tests) but measured slower than pairs: see corr_ctas().
    constexpr bool PAIR = (CTAS >= 2);
    constexpr bool QUAD = (CTAS == 4);
    // elements of one k-block along the TMA's K coordinate: 128 bytes either way (64 fp16 / 128 int8)
    constexpr int BKE = I8 ? 128 : BK;
    static_assert(!I8 || (EPI == EPI_STORE && NCHAIN == 4), "the int8 form is wired for the vote store epilogue");
    static_assert(!I8 || (EW * (BN / 2) * 4 >= BN * 16), "int8 epilogue keeps a double and a pointer per column in s_colw");
    static_assert(CTAS == 1 || CTAS == 2 || CTAS == 4, "1 CTA, a pair, or two pairs");
    constexpr int B_ROWS = BN / (PAIR ? 2 : 1);    // rows of the B tile this CTA stages
    constexpr uint32_t A_BYTES = BM * BK * 2;
    constexpr uint32_t B_BYTES = B_ROWS * BK * 2;
    // int8 form: one stage holds BOTH limb tiles of A and ALL FOUR digit tiles of B for one k-block; the (up to) seven
    // digit products are issued from that one stage, so every operand byte is staged once per k-block (the
    // term-major order of the fp16 form would stage 7 x (A + B) = 2.3x as much -- and the kernel is bound by
    // shared-memory bandwidth: TMA writes + MMA operand reads)
    constexpr uint32_t STAGE_BYTES = I8 ? (2 * A_BYTES + 4 * B_BYTES) : (A_BYTES + B_BYTES);
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0u;                     // cluster leader: claims tiles
    const uint32_t pair_rank = cta_rank & 1u;               // rank inside the CTA pair
    const uint32_t pair_lead = cta_rank & ~1u;              // cluster rank of this pair's leader (issues the MMAs)
    const bool pair_leader = pair_rank == 0u;
    // NCHAIN independent accumulators per tile (k-blocks dealt round-robin, summed in FP32 by the
    // epilogue): the tensor core truncates when it adds into a long FP32 chain, so shorter chains
    // cut the accumulation bias NCHAIN-fold.  Two tiles in flight when TMEM (512 columns) allows.
    constexpr int ACC_COLS = NCHAIN * BN;
    constexpr int ACC_STAGES = (2 * ACC_COLS <= 512) ? 2 : 1;
    static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
    constexpr int WANT_COLS = ACC_COLS * ACC_STAGES;
    constexpr int TMEM_COLS = (WANT_COLS <= 32) ? 32 : (WANT_COLS <= 64) ? 64 : (WANT_COLS <= 128) ? 128 : (WANT_COLS <= 256) ? 256 : 512;

    // carve-up: [stages * STAGE_BYTES] operand ring (1024-aligned) | epilogue table | barriers etc.
    unsigned char* ring = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);   // swizzle atoms need 1024-byte alignment
    unsigned char* tail = ring + (size_t)p.stages * STAGE_BYTES;
    // hist: [nbins] counts + slot nbins (rho rounding above 1, folded into the top bin at the end)
    //       + slot nbins+1 (sink for dropped cells);  lut: [nbins+1] edges + one pad entry
    uint32_t* s_table = reinterpret_cast<uint32_t*>(tail);
    const int table_words = (EPI == EPI_STORE) ? 0 : (p.nbins + 2 + 3) & ~3;
    float* s_invb = reinterpret_cast<float*>(s_table + table_words);       // [BN]
    int32_t* s_lab = reinterpret_cast<int32_t*>(s_invb + BN);              // [BN]
    float* s_add = reinterpret_cast<float*>(s_lab + BN);                   // [BN] (EPI_STORE) / flush masks (EPI_VOTE)
    uint32_t* s_flush = reinterpret_cast<uint32_t*>(s_add);
    static_assert(EW == 8 || (EW == 16 && EPI == EPI_HIST), "16 epilogue warps are wired for the pass-1 epilogue only");
    constexpr int EPI_T = EW * 32;                                         // epilogue threads of this instantiation
    constexpr int ALL_T = 64 + EPI_T;
    uint32_t* s_colw = reinterpret_cast<uint32_t*>(s_add + BN);            // [EW][BN/2] column sums (EPI_VOTE)
    int32_t* s_wlab = reinterpret_cast<int32_t*>(s_colw + EW * (BN / 2));  // [EW] row label of a uniform warp
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_wlab + EW);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + MAX_STAGES;
    uint64_t* tfull_bar = bars + 2 * MAX_STAGES;
    uint64_t* tempty_bar = bars + 2 * MAX_STAGES + 2;
    uint64_t* qfull_bar = bars + 2 * MAX_STAGES + 4;                       // tile-index queue (dynamic scheduler)
    uint64_t* qempty_bar = bars + 2 * MAX_STAGES + 4 + TQ;
    int32_t* s_tileq = reinterpret_cast<int32_t*>(bars + 2 * MAX_STAGES + 4 + 2 * TQ);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_tileq + TQ);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_iters = p.n_terms * p.k_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi);
        tma_prefetch_desc(&tm_b_hi);
        if (p.n_terms > 1) {
            tma_prefetch_desc(&tm_a_lo);
            tma_prefetch_desc(&tm_b_lo);
        }
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], QUAD ? 2 : 1);        // quad: both pairs must have consumed the shared B slot
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], EW * (PAIR ? 2 : 1));    // a pair leader's MMA thread waits for the epilogues of both its CTAs
        }
        for (int qi = 0; qi < TQ; ++qi) {
            mbar_init(&qfull_bar[qi], 1);
            // cluster leader: every consumer of a queue slot in the cluster -- per CTA one lane of every epilogue
            // warp, the producers of the non-leader CTAs, the MMA thread of every pair leader
            mbar_init(&qempty_bar[qi], CTAS * EW + (CTAS - 1) + (QUAD ? 2 : 1));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    if (EPI != EPI_STORE) {
        // hist starts at zero; lut is copied in
        for (int i = threadIdx.x; i < p.nbins + 2; i += ALL_T)
            s_table[i] = (EPI == EPI_VOTE) ? p.lut[i] : 0u;
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();        // the peer's barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    // where the consumers of this CTA hand queue slots back (the cluster leader's barriers) and accumulator
    // stages back (their pair leader's)
    auto arrive_leader = [&](uint64_t* bar) {
        if constexpr (PAIR) mbar_arrive_remote(mapa_u32(smem_u32(bar), 0u));
        else mbar_arrive(bar);
    };
    auto arrive_pair_leader = [&](uint64_t* bar) {
        if constexpr (PAIR) mbar_arrive_remote(mapa_u32(smem_u32(bar), pair_lead));
        else mbar_arrive(bar);
    };
    auto wait_queue = [&](uint64_t* bar, uint32_t parity) {
        if constexpr (PAIR) mbar_wait_cluster(bar, parity);
        else mbar_wait(bar, parity);
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            // Dynamic tile scheduler: the producer claims the next linear tile index with one atomicAdd and
            // publishes it to the MMA and epilogue roles through a small shared-memory queue.  All CTAs
            // therefore work inside one narrow window of the (band, column, row) order even when their
            // speeds differ -- which keeps the band's operands in L2 (a static round-robin lets slow CTAs
            // fall behind until the working set of in-flight column tiles exceeds L2).
            TileIter it;
            it.init(p, BN, 0);
            int cur_tile = 0;
            int s = 0, qs = 0;
            uint32_t ph = 0, qph = 0;
            for (;;) {
                int tile;
                if (leader) {
                    mbar_wait(&qempty_bar[qs], qph ^ 1u);
                    const unsigned int t = atomicAdd(p.tile_counter, 1u) + (unsigned int)p.tile_base;
                    tile = (t < (unsigned int)p.total_tiles) ? (int)t : -1;
                    s_tileq[qs] = tile;
                    mbar_arrive(&qfull_bar[qs]);
                    if constexpr (PAIR) {          // ... and into the other CTAs' copies of the queue
                        for (uint32_t r = 1; r < (uint32_t)CTAS; ++r) {
                            st_remote_u32(mapa_u32(smem_u32(&s_tileq[qs]), r), (uint32_t)tile);
                            mbar_arrive_remote(mapa_u32(smem_u32(&qfull_bar[qs]), r));
                        }
                    }
                } else {
                    wait_queue(&qfull_bar[qs], qph);
                    tile = s_tileq[qs];
                    arrive_leader(&qempty_bar[qs]);
                }
                if (++qs == TQ) {
                    qs = 0;
                    qph ^= 1u;
                }
                if (tile < 0) break;
                it.advance(tile - cur_tile);
                cur_tile = tile;
                const int m0 = it.mt * (BM * CTAS) + (int)cta_rank * BM, n0 = it.nt() * BN + (int)pair_rank * B_ROWS;
                if constexpr (I8) {
                    for (int kb = 0; kb < p.k_blocks; ++kb) {
                        mbar_wait(&empty_bar[s], ph ^ 1u);
                        unsigned char* sa = ring + (size_t)s * STAGE_BYTES;
                        if constexpr (PAIR) {
                            if (pair_leader) mbar_expect_tx(&full_bar[s], STAGE_BYTES * 2);
                            tma_load_2d_pair(sa, &tm_a_hi, &full_bar[s], kb * BKE, m0);
                            tma_load_2d_pair(sa + A_BYTES, &tm_a_lo, &full_bar[s], kb * BKE, m0);
#pragma unroll
                            for (int d = 0; d < 4; ++d)
                                tma_load_2d_pair(sa + 2 * A_BYTES + d * B_BYTES, &tm_b_hi, &full_bar[s], kb * BKE,
                                                 n0 + d * p.b_limb_rows);
                        } else {
                            mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                            tma_load_2d(sa, &tm_a_hi, &full_bar[s], kb * BKE, m0);
                            tma_load_2d(sa + A_BYTES, &tm_a_lo, &full_bar[s], kb * BKE, m0);
#pragma unroll
                            for (int d = 0; d < 4; ++d)
                                tma_load_2d(sa + 2 * A_BYTES + d * B_BYTES, &tm_b_hi, &full_bar[s], kb * BKE,
                                            n0 + d * p.b_limb_rows);
                        }
                        if (++s == p.stages) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
                    continue;
                }
                for (int term = 0; term < p.n_terms; ++term) {
                    const CUtensorMap* ma = p.a_sel[term] ? &tm_a_lo : &tm_a_hi;
                    const CUtensorMap* mb = (!I8 && p.b_sel[term]) ? &tm_b_lo : &tm_b_hi;
                    const int nb0 = I8 ? n0 + p.b_sel[term] * p.b_limb_rows : n0;
                    for (int kb = 0; kb < p.k_blocks; ++kb) {
                        mbar_wait(&empty_bar[s], ph ^ 1u);
                        unsigned char* sa = ring + (size_t)s * STAGE_BYTES;
                        if constexpr (PAIR) {
                            if (pair_leader) mbar_expect_tx(&full_bar[s], STAGE_BYTES * 2);
                            tma_load_2d_pair(sa, ma, &full_bar[s], kb * BKE, m0);
                            if constexpr (QUAD) {      // pair 0 fetches the B half for its own CTA and the one below it
                                if (cta_rank < 2u)
                                    tma_load_2d_pair_mc(sa + A_BYTES, mb, &full_bar[s], kb * BKE, nb0, (uint16_t)(5u << pair_rank));
                            } else {
                                tma_load_2d_pair(sa + A_BYTES, mb, &full_bar[s], kb * BKE, nb0);
                            }
                        } else {
                            mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                            tma_load_2d(sa, ma, &full_bar[s], kb * BKE, m0);
                            tma_load_2d(sa + A_BYTES, mb, &full_bar[s], kb * BKE, nb0);
                        }
                        if (++s == p.stages) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0 && pair_leader) {
            const uint32_t idesc = I8 ? make_idesc_i8(BM * (PAIR ? 2 : 1), BN) : make_idesc(BM * (PAIR ? 2 : 1), BN);
            const uint16_t pair_mask = (uint16_t)(3u << pair_lead);            // the two CTAs of this pair
            const uint16_t slot_mask = QUAD ? (uint16_t)0xF : pair_mask;       // everyone whose producer reuses the slot
            int s = 0, a = 0, qs = 0;
            uint32_t ph = 0, aph = 0, qph = 0;
            for (;;) {
                wait_queue(&qfull_bar[qs], qph);
                const int tile = s_tileq[qs];
                arrive_leader(&qempty_bar[qs]);
                if (++qs == TQ) {
                    qs = 0;
                    qph ^= 1u;
                }
                if (tile < 0) break;
                mbar_wait(&tempty_bar[a], aph ^ 1u);
                tc_fence_after();
                if constexpr (I8) {
                    // k-block-major: all digit products of a k-block from one stage; accumulator = weight class
                    for (int kb = 0; kb < p.k_blocks; ++kb) {
                        mbar_wait(&full_bar[s], ph);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(ring + (size_t)s * STAGE_BYTES);
                        const int k_steps = (kb == p.k_blocks - 1) ? p.k_last : BK / UMMA_K;
                        for (int term = 0; term < p.n_terms; ++term) {
                            const uint32_t d_tmem = tmem_base + (uint32_t)(a * ACC_COLS + p.cls[term] * BN);
                            const bool fresh = (kb == 0 && p.cls_first[term] != 0);
                            const uint64_t adesc = make_smem_desc(sa + (uint32_t)p.a_sel[term] * A_BYTES);
                            const uint64_t bdesc = make_smem_desc(sa + 2 * A_BYTES + (uint32_t)p.b_sel[term] * B_BYTES);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k) {
                                if (k >= k_steps) break;
                                const uint32_t accum = (!fresh || k > 0) ? 1u : 0u;
                                if constexpr (PAIR) tc_mma_i8_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accum);
                                else tc_mma_i8(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accum);
                            }
                        }
                        if constexpr (PAIR) tc_commit_pair(&empty_bar[s], slot_mask);
                        else tc_commit(&empty_bar[s]);
                        if (++s == p.stages) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
                }
                int kb = 0, term = 0;
                for (int i = 0; i < (I8 ? 0 : n_iters); ++i) {
                    // fp16 form: k-blocks dealt round-robin to the NCHAIN accumulators; int8 form: accumulator = weight class
                    const int chain = I8 ? p.cls[term] : (i % NCHAIN);
                    const bool fresh = I8 ? (kb == 0 && p.cls_first[term] != 0) : (i < NCHAIN);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(a * ACC_COLS + chain * BN);
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(ring + (size_t)s * STAGE_BYTES);
                    const uint64_t adesc = make_smem_desc(sa);
                    const uint64_t bdesc = make_smem_desc(sa + A_BYTES);
                    // the last k-block of a plane product holds G mod 64 genes (the rest is zero padding):
                    // only the 16-gene steps that carry data are issued (G = 2000: 1 of 4)
                    const int k_steps = (kb == p.k_blocks - 1) ? p.k_last : BK / UMMA_K;
                    if (++kb == p.k_blocks) {
                        kb = 0;
                        ++term;
                    }
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if (k >= k_steps) break;
                        // advance the start address by k * 16 elements * 2 B (32 int8) = 32 B = 2 (16-byte units)
                        const uint32_t accum = (!fresh || k > 0) ? 1u : 0u;
                        if constexpr (I8) {
                            if constexpr (PAIR) tc_mma_i8_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accum);
                            else tc_mma_i8(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accum);
                        } else {
                            if constexpr (PAIR) tc_mma_f16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accum);
                            else tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accum);
                        }
                    }
                    // frees the smem slot (of both CTAs) once these MMAs retire
                    if constexpr (PAIR) tc_commit_pair(&empty_bar[s], slot_mask);
                    else tc_commit(&empty_bar[s]);
                    if (++s == p.stages) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                // accumulator complete -> epilogue (of both CTAs)
                if constexpr (PAIR) tc_commit_pair(&tfull_bar[a], pair_mask);
                else tc_commit(&tfull_bar[a]);
                if (++a == ACC_STAGES) {
                    a = 0;
                    aph ^= 1u;
                }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int ew = warp - 2;                      // 0..7
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int half = ew >> 2;                     // which half of the tile's columns
        const int et = ew * 32 + lane;                // 0..255 among epilogue threads
        const int row_in_tile = q * 32 + lane;
        constexpr int HALF_COLS = BN / (EW / 4);      // columns per warp: half of the tile (EW = 8) or a quarter (16)
        const int cbeg = half * HALF_COLS;            // (half = ew >> 2 is the warp's column part, 0 .. EW/4 - 1)
        TileIter it;
        it.init(p, BN, 0);
        int cur_tile = 0;
        int a = 0, qs = 0;
        uint32_t aph = 0, qph = 0;
        const float bscale = bin_scale(p.nbins), boff = bin_offset(p.nbins);
        const uint64_t pol_ef = l2_evict_first_policy();
        for (;;) {
            wait_queue(&qfull_bar[qs], qph);
            const int tile = s_tileq[qs];
            __syncwarp();
            if (lane == 0) arrive_leader(&qempty_bar[qs]);
            if (++qs == TQ) {
                qs = 0;
                qph ^= 1u;
            }
            if (tile < 0) break;
            it.advance(tile - cur_tile);
            cur_tile = tile;
            const int64_t m0 = (int64_t)it.mt * (BM * CTAS) + (int64_t)cta_rank * BM, n0 = (int64_t)it.nt() * BN;
            const int64_t m = m0 + row_in_tile;
            // stage per-column metadata of this tile
            epi_bar_sync<EPI_T>();                           // previous tile's readers are done
            for (int c = et; c < BN; c += EPI_T) {
                const int64_t n = n0 + c;
                const bool ok = n < p.N;
                float ib = ok ? p.inv_b[n] : 0.f;
                if (EPI == EPI_VOTE && !(ib > 0.f)) ib = __int_as_float(0x7fc00000);   // dropped / absent column: NaN
                s_invb[c] = ib;
                if (EPI == EPI_STORE) s_add[c] = (ok && p.addend) ? p.addend[n] : 0.f;
                if (EPI == EPI_STORE && I8) {
                    s_lab[c] = (ok && p.den) ? p.col_den[n] : 0;
                    reinterpret_cast<double*>(s_colw)[c] = (ok && p.col_scale) ? p.col_scale[n] : (p.den ? 1.0 : 0.0);
                    // where this column's votes go: the local column-major matrix, or (peer mode) the receive buffer of
                    // the rank that owns the column -- a peer pointer, written over NVLink by the stores below
                    float* dst = nullptr;
                    if (ok && p.out) {
                        dst = p.out + n * p.ldo;
                        if (p.peer_world > 0) {
                            const int d = (int)(n / p.peer_cols);
                            dst = p.peer_out[d] + ((int64_t)p.peer_src * p.peer_cols + (n - (int64_t)d * p.peer_cols)) * p.peer_per;
                        }
                    }
                    reinterpret_cast<float**>(reinterpret_cast<double*>(s_colw) + BN)[c] = dst;
                }
                if (EPI == EPI_VOTE) {
                    const int lab = ok ? p.label[n] : 0;
                    s_lab[c] = lab;
                    // a label run starts here (each half of the tile starts its own run)
                    const bool brk = (c == 0) || (c == HALF_COLS) || !ok || (n >= 1 && p.label[n - 1] != lab);
                    const unsigned msk = __ballot_sync(0xffffffffu, brk);
                    if (lane == 0) s_flush[c >> 5] = msk;
                }
            }
            epi_bar_sync<EPI_T>();
            const bool row_ok = m < p.M;
            const float inv_i = row_ok ? p.inv_a[m] : 0.f;

            mbar_wait(&tfull_bar[a], aph);
            tc_fence_after();
            const uint32_t t_base = tmem_base + (uint32_t)(a * ACC_COLS) + ((uint32_t)(q * 32) << 16);
            const int n_chains = n_iters < NCHAIN ? n_iters : NCHAIN;
            // some column index <= some row index: the tile touches the diagonal or (the lower CTAs of a
            // quad's first column tile) lies entirely below it -- either way the per-element test decides
            const bool diag_tile = (n0 <= m0 + BM - 1);

            if (EPI == EPI_STORE && I8) {
                // exact integer dot product: sum over the weight classes, class c scaled by 128^(c+1) (Horner, int64)
                const double inv_i64 = row_ok ? p.inv_a64[m] : 0.0;
#pragma unroll 1
                for (int c0 = cbeg; c0 < cbeg + HALF_COLS; c0 += 32) {
                    long long v[32];
                    uint32_t r[32];
                    tmem_ld32(t_base + (uint32_t)(c0 + (p.n_cls - 1) * BN), r);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = (long long)(int)r[c];
#pragma unroll
                    for (int cl = 2; cl >= 0; --cl) {
                        if (cl >= p.n_cls - 1) continue;
                        tmem_ld32(t_base + (uint32_t)(c0 + cl * BN), r);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] = v[c] * 128 + (long long)(int)r[c];
                    }
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] *= p.cls_mul;
                    if (row_ok) {
                        const double* s_scale64 = reinterpret_cast<const double*>(s_colw);
                        float* const* s_dst = reinterpret_cast<float* const*>(s_scale64 + BN);
                        int last_sd = -1;
                        double rden = 1.0;
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const int64_t n = n0 + c0 + c;
                            if (n < p.N) {
                                double val = fma((double)v[c] * inv_i64, (double)s_invb[c0 + c], (double)s_add[c0 + c]);
                                if (p.den) {
                                    // the denominator depends on (cell, train study) only: one reciprocal per run of
                                    // columns of one study (warp-uniform branch: all lanes work on the same column)
                                    const int sd = s_lab[c0 + c];
                                    if (sd != last_sd) {
                                        last_sd = sd;
                                        rden = 1.0 / p.den[(int64_t)sd * p.ld_den + m];
                                    }
                                    val = fma(val * rden, s_scale64[c0 + c], -1.0);
                                } else {
                                    val -= s_scale64[c0 + c];
                                }
                                if (p.out64) p.out64[n * p.ldo + m] = p.acc64 ? p.out64[n * p.ldo + m] + val : val;
                                else s_dst[c0 + c][m] = (float)val;
                            }
                        }
                    }
                }
            } else if (EPI == EPI_STORE) {
#pragma unroll 1
                for (int c0 = cbeg; c0 < cbeg + HALF_COLS; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_chains<BN, NCHAIN>(t_base + (uint32_t)c0, n_chains, r);
                    if (row_ok) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const int64_t n = n0 + c0 + c;
                            if (n < p.N)
                                p.out[n * p.ldo + m] = fmaf(__uint_as_float(r[c]) * inv_i, s_invb[c0 + c], s_add[c0 + c]);
                        }
                    }
                }
            } else if (EPI == EPI_HIST) {
                const float inv_i_half = inv_i * bscale;
                const unsigned sink = (unsigned)p.nbins + 1u;
                const uint32_t q_top = ((uint32_t)p.nbins << Q_FRAC_BITS) - 1u;
                const bool row_live = inv_i > 0.f;
                // rho spill: this warp's 32 x 32 chunks in tile-native order, [ew][chunk][k][lane] x 16 B
                // (coalesced 512-byte stores); masked entries are stored as 0, which maps to W = lut[0] = 0
                // (a pair's 256-row tile is stored as two 128-row tiles: 2 * tile + rank)
                const int sp_tile = tile * CTAS + (int)cta_rank;
                const bool do_spill = p.spill != nullptr && sp_tile < p.spill_tiles;
                // the spill keeps the 8-warp layout [role][chunk]: role = (lane quarter, column half), BN/64 chunks
                const int sp_role = ((q + 2) & 3) + 4 * (half / (EW / 8));
                const int sp_chunk0 = (half % (EW / 8)) * (HALF_COLS / 32);
                uint4* sp = reinterpret_cast<uint4*>(p.spill + (size_t)sp_tile * (size_t)(BM * BN)) +
                            (size_t)(sp_role * (BN / 64) + sp_chunk0) * 256 + lane;
#pragma unroll 1
                for (int c0 = cbeg; c0 < cbeg + HALF_COLS; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_chains<BN, NCHAIN>(t_base + (uint32_t)c0, n_chains, r);
                    // column scales four at a time (one broadcast LDS.128 instead of four LDS: the epilogue's
                    // shared-memory wavefronts compete with the TMA writes and the MMA operand reads)
                    const float4* invb4 = reinterpret_cast<const float4*>(s_invb + c0);
                    if (row_live) {
                        if (!diag_tile) {
#pragma unroll
                            for (int c4 = 0; c4 < 8; ++c4) {
                                const float4 ib = invb4[c4];
                                const float ibv[4] = {ib.x, ib.y, ib.z, ib.w};
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const int c = 4 * c4 + u;
                                    const uint32_t qv = rho_q(__uint_as_float(r[c]), inv_i_half, ibv[u], boff, q_top);
                                    const bool use = ibv[u] > 0.f;
                                    atomicAdd(&s_table[use ? (qv >> Q_FRAC_BITS) : sink], 1u);
                                    r[c] = use ? qv : 0u;
                                }
                            }
                        } else {
#pragma unroll
                            for (int c4 = 0; c4 < 8; ++c4) {
                                const float4 ib = invb4[c4];
                                const float ibv[4] = {ib.x, ib.y, ib.z, ib.w};
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const int c = 4 * c4 + u;
                                    const uint32_t qv = rho_q(__uint_as_float(r[c]), inv_i_half, ibv[u], boff, q_top);
                                    const bool use = ibv[u] > 0.f && (n0 + c0 + c > m);
                                    atomicAdd(&s_table[use ? (qv >> Q_FRAC_BITS) : sink], 1u);
                                    r[c] = use ? qv : 0u;
                                }
                            }
                        }
                    } else if (do_spill) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) r[c] = 0u;
                    }
                    if (do_spill) {
                        // streaming stores (plain stores and an evict-first L2 hint measured the same)
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            __stcs(sp + k * 32, make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]));
                        sp += 256;
                    }
                }
            } else {   // EPI_VOTE
                // Every weight W_ij (j > i in `tri` mode: the network is symmetric, only the upper
                // triangle is computed) is added to votes[i, label(j)] (row side) and, in `tri` mode, to
                // votes[j, label(i)] (column side); see VoteRows.  Dropped cells carry NaN scales here:
                // NaN propagates through the bin position, the float->uint conversion turns it into
                // q = 0 and lut[0] == 0, so their weights vanish without a predicate in the inner loop.
                const float qnan = __int_as_float(0x7fc00000);
                const float inv_i_half = (inv_i > 0.f) ? inv_i * bscale : qnan;
                const bool sym = p.tri != 0;
                const uint32_t q_top = ((uint32_t)p.nbins << Q_FRAC_BITS) - 1u;
                VoteRows vr;
                vr.begin(p.label[row_ok ? m : p.M - 1], p.votes + (row_ok ? m : 0) * (int64_t)p.L, lane);
                uint32_t* colw = s_colw + ew * HALF_COLS;
#pragma unroll 1
                for (int c0 = cbeg; c0 < cbeg + HALF_COLS; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_chains<BN, NCHAIN>(t_base + (uint32_t)c0, n_chains, r);
                    const float4* invb4 = reinterpret_cast<const float4*>(s_invb + c0);
                    if (!diag_tile) {
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4) {
                            const float4 ib = invb4[c4];
                            r[4 * c4] = rho_q(__uint_as_float(r[4 * c4]), inv_i_half, ib.x, boff, q_top);
                            r[4 * c4 + 1] = rho_q(__uint_as_float(r[4 * c4 + 1]), inv_i_half, ib.y, boff, q_top);
                            r[4 * c4 + 2] = rho_q(__uint_as_float(r[4 * c4 + 2]), inv_i_half, ib.z, boff, q_top);
                            r[4 * c4 + 3] = rho_q(__uint_as_float(r[4 * c4 + 3]), inv_i_half, ib.w, boff, q_top);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const uint32_t qv = rho_q(__uint_as_float(r[c]), inv_i_half, s_invb[c0 + c], boff, q_top);
                            const bool use = sym ? (n0 + c0 + c > m) : (n0 + c0 + c != m);
                            r[c] = use ? qv : 0u;
                        }
                    }
                    const uint32_t mycol = vr.chunk(r, s_flush[c0 >> 5], SmemLabels{s_lab + c0}, s_table,
                                                    sym, p.votes + (n0 + c0) * (int64_t)p.L, p.L, pol_ef, lane);
                    if (sym) colw[c0 - cbeg + lane] = mycol;
                }
                vr.finish();
                if (sym && lane == 0) s_wlab[ew] = vr.warp_uniform ? vr.lab_i : -2;
            }

            // the accumulator stage is drained: hand it back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_pair_leader(&tempty_bar[a]);
            if (++a == ACC_STAGES) {
                a = 0;
                aph ^= 1u;
            }

            if (EPI == EPI_VOTE && p.tri != 0) {
                epi_bar_sync<EPI_T>();
                vote_merge_columns<BN>(s_colw, s_wlab, p.votes, n0, p.L, pol_ef, ew, lane);
            }
        }
        if (EPI == EPI_HIST) {
            epi_bar_sync<EPI_T>();
            for (int b = et; b <= p.nbins; b += EPI_T) {
                const uint32_t v = s_table[b];
                if (v) atomicAdd(p.hist + (b < p.nbins ? b : p.nbins - 1), (unsigned long long)v);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();        // nobody leaves (or frees TMEM) while the peer may still signal it
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ pass 2 from the rho spill
// Streams the bin positions pass 1 left in HBM (tile-native layout, see EPI_HIST) instead of
// recomputing the tiles on the tensor cores: 4 B per cell pair.  Same VoteRows arithmetic as the GEMM
// epilogue, hence bit-identical vote sums.
// Every warp works on its own: it owns one 32-row quarter of `tb` consecutive stored tiles (same column
// tile, consecutive row tiles inside a band) and walks all 256 columns of each tile, so that
//   - a row's running sum is flushed once per label run of the whole tile row (one 64-bit atomic per row
//     and tile),
//   - the column sums stay in registers across the run of tiles and are flushed when the column tile
//     or the rows' label changes (one 64-bit reduction per column and ~tb tiles),
//   - no CTA barrier is needed.
// Data path: each warp owns a 4 KB shared-memory slot; one lane fetches the warp's next chunk with a
// bulk async copy (cp.async.bulk, mbarrier completion) as soon as the current one has been moved to
// registers, so the HBM latency hides behind the ~600 instructions a chunk costs and no registers are
// spent on prefetching: 20 warps per SM at 96 registers (the column sums live in shared memory too).  The kernel is bound by shared-memory lookups (one
// random-bank LDS per pair, ~3.5-way conflicts) and latency, not by HBM.
constexpr int SPILL_WARPS = 24;          // most the slots allow; the launcher's default is 20 (96 registers, measured best)
__device__ __forceinline__ void bulk_load_4k(void* smem_dst, const void* gsrc, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4096;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 4096, [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(smem_u32(bar)) : "memory");
}
template <int BN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) vote_spill_kernel(const GemmParams p, int tb) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    constexpr int HALF_COLS = BN / 2;
    constexpr int NCH = HALF_COLS / 32;         // chunks per column half
    constexpr int NCK = 2 * NCH;                // chunks per tile row
    const int n_warps = (int)blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4* s_slot = reinterpret_cast<uint4*>(s_raw) + (size_t)warp * 256;                    // this warp's 4 KB slot
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_raw + (size_t)n_warps * 4096) + warp;
    // column sums carried across a run of tiles: [chunk][lane] per warp, in shared memory -- registers are
    // what limits the overlap of the table lookups (at 80 registers the compiler funnelled all 32 lookups of a
    // chunk through one temporary, exposing the full LDS latency 32 times)
    unsigned long long* s_col = reinterpret_cast<unsigned long long*>(s_raw + (size_t)n_warps * (4096 + 8)) +
                                (size_t)warp * (NCK * 32) + lane;
    uint32_t* s_table = reinterpret_cast<uint32_t*>(s_raw + (size_t)n_warps * (4096 + 8 + NCK * 32 * 8));   // [nbins + 2]
    for (int i = threadIdx.x; i < p.nbins + 2; i += blockDim.x) s_table[i] = p.lut[i];
    if (lane == 0) {
        mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int gw = blockIdx.x * n_warps + warp;
    const int quarter = gw & 3;                                  // rows quarter * 32 .. + 31 of every tile
    const int role0 = (quarter + 2) & 3;                         // GEMM epilogue warp (ew) that wrote them, column half 0
    const int n_blocks = (p.spill_tiles + tb - 1) / tb;
    const int blk_step = ((int)gridDim.x * n_warps) >> 2;
    const uint64_t pol_ef = l2_evict_first_policy();
    constexpr size_t TILE_V4 = (size_t)BM * BN / 4;
    // (tile, chunk k of the tile row): role = role0 + 4 * (k / NCH), chunk-in-role = k % NCH; 4 KB contiguous
    const uint4* base = reinterpret_cast<const uint4*>(p.spill);
    auto chunk_src = [&](int tile, int k) {
        return base + (size_t)tile * TILE_V4 + (size_t)(((role0 + 4 * (k / NCH)) * NCH + (k % NCH)) * 256);
    };

    TileIter it;
    it.init(p, BN, 0);
    int cur_tile = 0;
    uint32_t phase = 0;
    int blk = gw >> 2;
    if (blk < n_blocks && lane == 0) bulk_load_4k(s_slot, chunk_src(blk * tb, 0), s_bar);
    // column side carried across tiles
#pragma unroll
    for (int k = 0; k < NCK; ++k) s_col[k * 32] = 0ull;
    int64_t col_n0 = -1;
    int col_lab = 0;

    auto flush_cols = [&]() {
        if (col_n0 >= 0) {
#pragma unroll 1
            for (int k = 0; k < NCK; ++k) {
                const unsigned long long v = s_col[k * 32];
                if (v) red_add_u64_hint(p.votes + (col_n0 + 32 * k + lane) * (int64_t)p.L + col_lab, v, pol_ef);
                s_col[k * 32] = 0ull;
            }
        }
        col_n0 = -1;
    };

    for (; blk < n_blocks; blk += blk_step) {
        const int t_end = min(p.spill_tiles, (blk + 1) * tb);
        for (int tile = blk * tb; tile < t_end; ++tile) {
            // stored tile -> (tile of the GEMM's schedule, CTA rank inside a pair)
            const int gt = tile / p.ctas, gr = tile - gt * p.ctas;
            it.advance(gt - cur_tile);
            cur_tile = gt;
            const int64_t m0 = (int64_t)it.mt * (BM * p.ctas) + (int64_t)gr * BM, n0 = (int64_t)it.nt() * BN;
            const int64_t m = m0 + quarter * 32 + lane;
            const bool row_ok = m < p.M;
            VoteRows vr;
            vr.begin(p.label[row_ok ? m : p.M - 1], p.votes + (row_ok ? m : 0) * (int64_t)p.L, lane);
            if (col_n0 >= 0 && (col_n0 != n0 || col_lab != vr.lab_i || !vr.warp_uniform)) flush_cols();
            if (vr.warp_uniform) {
                col_n0 = n0;
                col_lab = vr.lab_i;
            }
            // NOT unrolled: eight copies of the chunk body (~6,000 instructions each) overflow the instruction
            // cache -- ncu showed "no instruction" as the top stall of the unrolled version.
#pragma unroll 1
            for (int k = 0; k < NCK; ++k) {
                uint32_t r[32];
                // labels of this chunk's columns and of the column before each (L1-resident; issued before the wait)
                const int64_t n = n0 + 32 * k + lane;
                const int lab_k = n < p.N ? __ldg(p.label + n) : -1 - k;      // past the end: a run of its own (zero weights)
                const int lab_prev = (n >= 1 && n - 1 < p.N) ? __ldg(p.label + n - 1) : -1 - k;
                mbar_wait(s_bar, phase);
                phase ^= 1u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 v = s_slot[j * 32 + lane];
                    r[4 * j] = v.x;
                    r[4 * j + 1] = v.y;
                    r[4 * j + 2] = v.z;
                    r[4 * j + 3] = v.w;
                }
                __syncwarp();                     // every lane has taken its part of the slot
                // refill the slot with this warp's next chunk: same tile, next tile of the run, or the first
                // tile of its next run
                int nt_tile = tile, nk = k + 1;
                if (nk == NCK) {
                    nk = 0;
                    nt_tile = tile + 1;
                    if (nt_tile == t_end) nt_tile = (blk + blk_step < n_blocks) ? (blk + blk_step) * tb : -1;
                }
                if (nt_tile >= 0 && lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads before the async write
                    bulk_load_4k(s_slot, chunk_src(nt_tile, nk), s_bar);
                }
                // label runs of this chunk's columns: a run starts where the label differs from the previous column's
                const unsigned brk = __ballot_sync(0xffffffffu, (k == 0 && lane == 0) || lab_prev != lab_k);
                const int lab_use = lab_k < 0 ? 0 : lab_k;
                const uint32_t mycol = vr.chunk(r, brk, LaneLabels{lab_use}, s_table, true,
                                                p.votes + (n0 + 32 * k) * (int64_t)p.L, p.L, pol_ef, lane);
                if (mycol) s_col[k * 32] += mycol;
            }
            vr.finish();
        }
    }
    flush_cols();
}

// ------------------------------------------------------------------ CDF LUT
// hist: counts of rho over pairs i<j.  Ranking covers both triangles plus the
// n_diag diagonal entries, which tie at the top (utils.py:72).  For a value at
// fraction t of bin b:  rank - 1 ~= 2*(cum[b] + t*cnt[b]);  W = (rank-1)/top with
// top = F - (n_diag-1)/2 - 1, F = 2*total + n_diag (utils.py:47-49), in 2^MN_W_SHIFT fixed point.
// Output: one packed word per bin, (W0 << 11) | rise (see w_from_q).  A bin whose weight would rise by
// 2047 units or more holds >= 0.2 % of all pairs: that is a tie block (a few genes, heavily tied ranks),
// and bottleneck.rankdata (utils.py:45) gives every member of a tie block its MID-rank -- so the bin gets
// W0 = W(2 cum + cnt - 1/2) and rise 0 instead of an interpolation by the position inside the bin, which
// would shift the whole block by up to half its size.  lut[nbins] = top edge (rise 0),
// lut[nbins + 1] = number of such bins.
__global__ void corr_lut_kernel(const unsigned long long* __restrict__ hist, int nbins, long long n_diag,
                                uint32_t* __restrict__ lut) {
    __shared__ unsigned long long s_part[1024];
    __shared__ int s_heavy;
    const int T = blockDim.x, tid = threadIdx.x;
    const int per = (nbins + T - 1) / T;
    const int b0 = tid * per, b1 = min(nbins, b0 + per);
    unsigned long long local = 0;
    for (int b = b0; b < b1; ++b) local += hist[b];
    s_part[tid] = local;
    if (tid == 0) s_heavy = 0;
    __syncthreads();
    if (tid == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < T; ++i) {
            const unsigned long long v = s_part[i];
            s_part[i] = run;
            run += v;
        }
        s_part[T] = run;      // T < 1024 is guaranteed by the launcher
    }
    __syncthreads();
    const double total = (double)s_part[T];
    const double F = 2.0 * total + (double)n_diag;
    double top = F - ((double)n_diag - 1.0) * 0.5 - 1.0;
    if (!(top > 0.0)) top = 1.0;
    const double one = (double)MN_W_ONE;
    const double scale = one / top;
    unsigned long long run = s_part[tid];
    int heavy = 0;
    for (int b = b0; b < b1; ++b) {
        const unsigned long long cnt = hist[b];
        uint32_t w0 = (uint32_t)fmin(one, rint(2.0 * (double)run * scale));
        const uint32_t w1 = (uint32_t)fmin(one, rint(2.0 * (double)(run + cnt) * scale));
        uint32_t rise = w1 - w0;
        if (rise >= MN_LUT_DELTA_MAX) {
            w0 = (uint32_t)fmin(one, rint((2.0 * (double)run + (double)cnt - 0.5) * scale));
            rise = 0u;
            ++heavy;
        }
        lut[b] = (w0 << MN_LUT_DELTA_BITS) | rise;
        run += cnt;
    }
    if (heavy) atomicAdd(&s_heavy, heavy);
    __syncthreads();
    if (tid == 0) {
        lut[nbins] = (uint32_t)fmin(one, rint(2.0 * total * scale)) << MN_LUT_DELTA_BITS;
        lut[nbins + 1] = (uint32_t)s_heavy;
    }
}

__global__ void vote_rowsum_kernel(const unsigned long long* __restrict__ votes, int64_t N, int L,
                                   unsigned long long* __restrict__ degree) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    unsigned long long s = 0ull;
    for (int l = 0; l < L; ++l) s += votes[i * L + l];
    degree[i] = s;
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// rows x cols fp16, row pitch ld elements; box = 64 columns x box_rows rows, 128B swizzle, zero OOB fill
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    bool i8 = false) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return 4;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * (i8 ? 1 : 2)};
    cuuint32_t box[2] = {(cuuint32_t)(i8 ? 128 : BK), (cuuint32_t)box_rows};      // 128 bytes of K either way
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, i8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed: CUresult %d (rows=%lld cols=%lld ld=%lld box_rows=%d base=%p)", (int)r,
                  (long long)rows, (long long)cols, (long long)ld, box_rows, base);
        return 5;
    }
    return 0;
}

static int sm_count() {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

// row tiles per L2 band of the tile schedule: 64 x 128 rows x (G * 2 B) of A = 32 MB at G = 2000, L2-resident
static int gemm_band() {
    if (const char* e = getenv("MN_GEMM_BAND")) {      // tuning knob
        const int v = atoi(e);
        if (v > 0) return v;
    }
    return 64;
}

// number of tiles of the triangular passes for one shard (the linear tile order of TileIter)
// (bm = rows per tile: 128, or 256 for CTA pairs)
static int64_t corr_tile_count(int64_t N, int64_t row0, int64_t row1, int shard_rank, int shard_world, int bn, int bm) {
    const int n_tiles = (int)((N + bn - 1) / bn);
    const int mt1 = (int)((row1 + bm - 1) / bm);
    int64_t tiles = 0;
    for (int mt = (int)(row0 / bm) + shard_rank; mt < mt1; mt += shard_world) {
        const int first = (mt * bm) / bn;
        if (first < n_tiles) tiles += n_tiles - first;
    }
    return tiles;
}

// CTAs per cluster for the two network passes: 2 (one cta_group::2 pair).  MN_GEMM_CTAS = 4 selects two pairs
// sharing the B tile through TMA multicast, = 1 (or MN_GEMM_1CTA) single CTAs.  Measured at config 2 with the
// board pinned at its 1,000 W cap in every variant (pass 1 is POWER-bound): pair 74.2 ms @ 1.44 GHz, pair
// without spill 66.5 ms @ 1.53 GHz, single CTAs without spill 68.1 ms, quad 79.1 ms @ 1.55 GHz -- the quad
// moves 25 % fewer operand bytes from L2 but its two pairs stall on each other's shared B slot.
static int corr_ctas(bool wide) {
    if (!wide || getenv("MN_GEMM_1CTA")) return 1;
    if (const char* e = getenv("MN_GEMM_CTAS")) {
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4) return v;
    }
    return 2;
}

// Tile counters of the dynamic scheduler: a small rotating pool PER DEVICE (a process may drive several GPUs,
// and launches on different streams may overlap), handed out under a lock.
static int next_tile_counter(unsigned int** out) {
    constexpr int N_COUNTERS = 64, MAX_DEV = 64;
    static std::mutex mu;
    static unsigned int* pools[MAX_DEV] = {nullptr};
    static unsigned int next[MAX_DEV] = {0};
    int dev = 0;
    MN_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEV) {
        set_error("gemm: device ordinal %d out of range", dev);
        return 6;
    }
    std::lock_guard<std::mutex> lock(mu);
    if (!pools[dev]) MN_CUDA(cudaMalloc(&pools[dev], N_COUNTERS * sizeof(unsigned int)));
    *out = pools[dev] + (next[dev]++ % N_COUNTERS);
    return 0;
}

template <int BN>
static int launch_vote_spill(const GemmParams& p, cudaStream_t st) {
    int warps = 20;
    if (const char* e = getenv("MN_SPILL_WARPS")) {    // tuning knob (multiple of 4): 16 / 20 / 24 measured 22.4 / 21.4 / 22.6 ms at config 2
        const int v = atoi(e);
        if (v >= 4 && v <= SPILL_WARPS && (v & 3) == 0) warps = v;
    }
    // per warp a 4 KB slot, its mbarrier and its column sums (8 chunks x 32 lanes x 8 B), then the packed LUT
    const size_t per_warp = 4096 + 8 + (size_t)(BN / 32) * 32 * 8, table = (size_t)((p.nbins + 2 + 3) & ~3) * 4;
    while (warps > 4 && warps * per_warp + table > 232448) warps -= 4;        // 32,768 bins leave room for 16 warps
    const size_t smem = (size_t)warps * per_warp + table;
    // (the register budget per thread follows the warp count: 24 warps -> 80, 20 -> 96, 16 -> 128)
    auto kern = warps > 20 ? vote_spill_kernel<BN, 24> : (warps > 16 ? vote_spill_kernel<BN, 20> : vote_spill_kernel<BN, 16>);
    MN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // tiles per run: the column sums are flushed about once per run, so long runs are cheaper (64 = one
    // column segment of a band), but every warp should still get >= 8 runs to balance the grid
    int tb = (int)(((long long)p.spill_tiles * 4) / (8ll * sm_count() * warps));
    tb = tb < 1 ? 1 : (tb > 64 ? 64 : tb);
    if (const char* e = getenv("MN_SPILL_TB")) {       // tuning knob
        const int v = atoi(e);
        if (v > 0) tb = v;
    }
    const int n_blocks = (p.spill_tiles + tb - 1) / tb;
    int grid = sm_count();
    const int want = (n_blocks * 4 + warps - 1) / warps;     // one warp per (run, row quarter)
    if (want < grid) grid = want;
    kern<<<grid, warps * 32, smem, st>>>(p, tb);
    MN_LAUNCH_CHECK();
    return 0;
}

// I8 = 1: a_hi / a_lo are the int8 limb planes a1 / a0 of A (row pitch lda BYTES), b_hi the stacked limb planes of B
// (4 * p.b_limb_rows rows, row pitch ldb bytes), b_lo unused.
template <int BN, int EPI, int NCHAIN, int CTAS = 1, int EW = EPI_WARPS, int I8 = 0>
static int launch_gemm(const void* a_hi, const void* a_lo, int64_t lda, int64_t M, const void* b_hi,
                       const void* b_lo, int64_t ldb, int64_t N, int G, GemmParams p, cudaStream_t st) {
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    const int64_t cols_a = lda, cols_b = ldb;     // padding columns are zero; beyond ld the TMA zero-fills
    int rc;
    if ((rc = make_map(&ta_hi, a_hi, M, cols_a, lda, BM, I8))) return rc;
    if ((rc = make_map(&ta_lo, a_lo ? a_lo : a_hi, M, cols_a, lda, BM, I8))) return rc;
    constexpr int B_BOX = BN / (CTAS >= 2 ? 2 : 1);      // rows of B one CTA stages (half a tile in the pair forms)
    const int64_t b_rows = I8 ? 4 * (int64_t)p.b_limb_rows : N;
    if ((rc = make_map(&tb_hi, b_hi, b_rows, cols_b, ldb, B_BOX, I8))) return rc;
    if ((rc = make_map(&tb_lo, b_lo ? b_lo : b_hi, b_rows, cols_b, ldb, B_BOX, I8))) return rc;
    p.M = M;
    p.N = N;
    p.ctas = CTAS;
    constexpr int BKE = I8 ? 128 : BK, KSTEP = I8 ? 32 : UMMA_K;
    p.k_blocks = (G + BKE - 1) / BKE;
    p.k_last = (G - (p.k_blocks - 1) * BKE + KSTEP - 1) / KSTEP;
    p.n_tiles = (int)((N + BN - 1) / BN);
    if (p.band <= 0) p.band = gemm_band() / CTAS;
    const size_t stage_bytes = I8 ? (size_t)(2 * BM + 4 * B_BOX) * BK * 2 : (size_t)(BM + B_BOX) * BK * 2;
    const size_t table = (EPI == EPI_STORE) ? 0 : (size_t)((p.nbins + 2 + 3) & ~3) * 4;
    const size_t fixed = table + (size_t)BN * 12 + (size_t)EW * (BN / 2) * 4 + EW * 4 +
                         (2 * MAX_STAGES + 4 + 2 * TQ) * 8 + TQ * 4 + 16;
    const size_t budget = 232448 - 1024;          // 227 KB minus alignment slack
    int stages = (int)((budget - fixed) / stage_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (const char* e = getenv("MN_GEMM_STAGES")) {    // tuning knob (can only lower the depth)
        const int v = atoi(e);
        if (v >= 2 && v < stages) stages = v;
    }
    if (stages < 2) {
        set_error("gemm: not enough shared memory for 2 pipeline stages (nbins=%d)", p.nbins);
        return 6;
    }
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + fixed + 1024;
    auto kern = gemm_kernel<BN, EPI, NCHAIN, CTAS, EW, I8>;
    MN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // number of tiles (of BM * CTAS rows)
    long long tiles = 0;
    for (int mt = p.mt0; mt < p.mt1; mt += p.mt_step) tiles += p.n_tiles - (p.tri ? (mt * BM * CTAS) / BN : 0);
    if (tiles <= p.tile_base) return 0;
    int grid = sm_count() / CTAS;                  // one CTA (or CTA pair) per SM (TPC)
    if (tiles - p.tile_base < grid) grid = (int)(tiles - p.tile_base);
    grid *= CTAS;
    // dynamic scheduler: one counter per launch out of a small rotating pool (launches on different
    // streams may overlap), zeroed on the launch stream
    if ((rc = next_tile_counter(&p.tile_counter))) return rc;
    p.total_tiles = (int)tiles;
    MN_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(unsigned int), st));
    if (CTAS == 1) {
        kern<<<grid, 64 + EW * 32, smem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
    } else {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(64 + EW * 32);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CTAS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MN_CUDA(cudaLaunchKernelEx(&cfg, kern, ta_hi, ta_lo, tb_hi, tb_lo, p));
    }
    MN_LAUNCH_CHECK();
    return 0;
}

static int check_planes(const char* who, const void* hi, int64_t ld, int G) {
    MN_CHECK_ARG(hi != nullptr, "%s: null plane", who);
    MN_CHECK_ARG((reinterpret_cast<uintptr_t>(hi) & 15) == 0, "%s: plane must be 16-byte aligned", who);
    MN_CHECK_ARG(ld >= G && (ld % 8) == 0, "%s: ld=%lld must be >= G=%d and a multiple of 8", who, (long long)ld, G);
    return 0;
}

}  // namespace mn

extern "C" {

int mn_votes_fast(const mn_half_t* a_hi, const mn_half_t* a_lo, int64_t lda, const float* inv_norm, int64_t M,
                  const mn_half_t* b_hi, const mn_half_t* b_lo, int64_t ldb, const float* inv_scale,
                  const float* addend, int32_t Nn, int32_t G, float* out, int64_t ldo, mn_stream_t stream) {
    using namespace mn;
    if (check_planes("mn_votes_fast(a)", a_hi, lda, G) || check_planes("mn_votes_fast(b)", b_hi, ldb, G)) return 1;
    MN_CHECK_ARG(inv_norm && inv_scale && out && M >= 1 && Nn >= 1 && ldo >= M, "mn_votes_fast: bad arguments");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    int t = 0;
    p.a_sel[t] = 0; p.b_sel[t] = 0; ++t;
    if (b_lo) { p.a_sel[t] = 0; p.b_sel[t] = 1; ++t; }
    if (a_lo) { p.a_sel[t] = 1; p.b_sel[t] = 0; ++t; }
    p.n_terms = t;
    p.tri = 0;
    p.mt0 = 0;
    p.mt1 = (int)((M + BM - 1) / BM);
    p.mt_step = 1;
    p.inv_a = inv_norm;
    p.inv_b = inv_scale;
    p.addend = addend;
    p.out = out;
    p.ldo = ldo;
    const __half* ah = reinterpret_cast<const __half*>(a_hi);
    const __half* al = reinterpret_cast<const __half*>(a_lo);
    const __half* bh = reinterpret_cast<const __half*>(b_hi);
    const __half* bl = reinterpret_cast<const __half*>(b_lo);
    cudaStream_t st = (cudaStream_t)stream;
    // BN <= 128 leaves room for 4 accumulator chains in TMEM (accuracy over tile width)
    if (Nn <= 64) return launch_gemm<64, EPI_STORE, 4>(ah, al, lda, M, bh, bl, ldb, Nn, G, p, st);
    return launch_gemm<128, EPI_STORE, 4>(ah, al, lda, M, bh, bl, ldb, Nn, G, p, st);
}

struct PeerOut {
    float* const* ptrs;     // host array of `world` device pointers (receive buffers), nullptr = off
    int world, src, cols;
    int64_t per;
};

static int votes_i8(int low_only, PeerOut peer, const int8_t* a1, const int8_t* a0, int64_t lda, const double* inv_norm64, int64_t M,
                     const int8_t* b_limbs, int64_t ldb, int32_t b_limb_rows, const float* inv_scale,
                     const float* addend, int32_t Nn, int32_t G, float* out, double* out64, int64_t ldo,
                     const double* den, int64_t ld_den, const int32_t* col_den, const double* col_scale,
                     mn_stream_t stream) {
    using namespace mn;
    MN_CHECK_ARG(a1 && a0 && b_limbs && inv_norm64 && inv_scale && (out || out64), "mn_votes_fast_i8: null pointer");
    MN_CHECK_ARG(!den || (col_den && ld_den >= M), "mn_votes_fast_i8: den needs col_den and ld_den >= M");
    MN_CHECK_ARG(M >= 1 && Nn >= 1 && ldo >= M && G >= 1 && G <= 16000, "mn_votes_fast_i8: bad shape (G <= 16000)");
    MN_CHECK_ARG(lda % 128 == 0 && ldb % 128 == 0 && lda >= G && ldb >= G, "mn_votes_fast_i8: row pitches must be multiples of 128 bytes");
    MN_CHECK_ARG(b_limb_rows >= Nn && b_limb_rows % 128 == 0, "mn_votes_fast_i8: b_limb_rows must be a multiple of 128 >= Nn");
    MN_CHECK_ARG(((reinterpret_cast<uintptr_t>(a1) | reinterpret_cast<uintptr_t>(a0) | reinterpret_cast<uintptr_t>(b_limbs)) & 15) == 0,
                 "mn_votes_fast_i8: planes must be 16-byte aligned");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    bool seen[4] = {false, false, false, false};
    p.n_cls = 4;
    if (low_only) {
        // the weight-1 product a0 * b0 alone, added to what out64 already holds (mn_votes_den_i8)
        p.a_sel[0] = 1;
        p.b_sel[0] = 3;
        p.cls[0] = 0;
        p.cls_first[0] = 1;
        p.n_terms = 1;
        p.cls_mul = 1;
        p.n_cls = 1;
        p.acc64 = 1;
    } else if (G <= 127) {
        // one limb of A (a1 = d itself, weight 1): the four products a1 * b3 .. b0 are the whole dot product
        for (int t = 0; t < 4; ++t) {
            p.a_sel[t] = 0;
            p.b_sel[t] = t;
            p.cls[t] = 3 - t;
            p.cls_first[t] = 1;
        }
        p.n_terms = 4;
        p.cls_mul = 1;
    } else {
        // limb weights: a1 = 128, a0 = 1; b3 .. b0 = 128^3 .. 1 -> class = exponent of 128 of the product, 1 .. 4
        // (accumulator slot = class - 1).  The weight-1 product a0 * b0 is NOT formed: TMEM holds four 128-column
        // accumulators.  sum_g a0 b0 is a zero-mean sum of G products of balanced 7-bit digits, ~1,400-2,400 sqrt(G),
        // against a dot product of ~2^25 G^2: 3e-7 of it at G = 128, 3e-9 at G = 3,000 (include/pymn_b200.h).
        static const int A_SEL[7] = {0, 0, 1, 0, 1, 0, 1}, B_SEL[7] = {0, 1, 0, 2, 1, 3, 2};
        for (int t = 0; t < 7; ++t) {
            p.a_sel[t] = A_SEL[t];
            p.b_sel[t] = B_SEL[t];
            p.cls[t] = (1 - A_SEL[t]) + (3 - B_SEL[t]) - 1;
            p.cls_first[t] = seen[p.cls[t]] ? 0 : 1;
            seen[p.cls[t]] = true;
        }
        p.n_terms = 7;
        p.cls_mul = 128;
    }
    // CTA pairs (cta_group::2, 256 x 128 tiles: each CTA stages its own 128 rows of A and HALF of every B digit
    // tile) when there are enough row tiles to keep every pair busy; MN_I8_CTAS=1 / 2 forces a form
    // a handful of columns (the node-degree rows, the supervised path's joint labels): 128 x 64 tiles of single CTAs --
    // half the MMA work of a 128-wide tile, and TMEM then holds two accumulator stages
    const bool narrow = (Nn <= 64 && !peer.ptrs && !getenv("MN_I8_WIDE"));
    int ctas = (!narrow && M >= (int64_t)2 * BM * sm_count()) ? 2 : 1;
    if (const char* e = getenv("MN_I8_CTAS")) {
        const int v = atoi(e);
        if (!narrow && (v == 1 || v == 2)) ctas = v;
    }
    if (peer.ptrs) {
        MN_CHECK_ARG(peer.world >= 2 && peer.world <= 8 && peer.src >= 0 && peer.src < peer.world && peer.cols >= 1 &&
                         peer.per >= M && Nn <= peer.world * peer.cols && out && !out64,
                     "mn_votes_fast_i8_peer: bad peer layout");
        for (int r = 0; r < peer.world; ++r) {
            MN_CHECK_ARG(peer.ptrs[r] != nullptr, "mn_votes_fast_i8_peer: null receive buffer of rank %d", r);
            p.peer_out[r] = peer.ptrs[r];
        }
        p.peer_world = peer.world;
        p.peer_src = peer.src;
        p.peer_cols = peer.cols;
        p.peer_per = peer.per;
    }
    p.b_limb_rows = b_limb_rows;
    p.tri = 0;
    p.mt0 = 0;
    p.mt1 = (int)((M + BM * ctas - 1) / (BM * ctas));
    p.mt_step = 1;
    p.inv_a64 = inv_norm64;
    p.inv_b = inv_scale;
    p.addend = addend;
    p.out = out;
    p.out64 = out64;
    p.den = den;
    p.ld_den = ld_den;
    p.col_den = col_den;
    p.col_scale = col_scale;
    p.ldo = ldo;
    if (narrow)
        return launch_gemm<64, EPI_STORE, 4, 1, EPI_WARPS, 1>(a1, a0, lda, M, b_limbs, nullptr, ldb, Nn, G, p, (cudaStream_t)stream);
    if (ctas == 2)
        return launch_gemm<128, EPI_STORE, 4, 2, EPI_WARPS, 1>(a1, a0, lda, M, b_limbs, nullptr, ldb, Nn, G, p, (cudaStream_t)stream);
    return launch_gemm<128, EPI_STORE, 4, 1, EPI_WARPS, 1>(a1, a0, lda, M, b_limbs, nullptr, ldb, Nn, G, p, (cudaStream_t)stream);
}

int mn_votes_fast_i8(const int8_t* a1, const int8_t* a0, int64_t lda, const double* inv_norm64, int64_t M,
                     const int8_t* b_limbs, int64_t ldb, int32_t b_limb_rows, const float* inv_scale,
                     const float* addend, int32_t Nn, int32_t G, float* out, double* out64, int64_t ldo,
                     const double* den, int64_t ld_den, const int32_t* col_den, const double* col_scale,
                     mn_stream_t stream) {
    return votes_i8(0, PeerOut{nullptr, 0, 0, 0, 0}, a1, a0, lda, inv_norm64, M, b_limbs, ldb, b_limb_rows, inv_scale, addend, Nn,
                    G, out, out64, ldo, den, ld_den, col_den, col_scale, stream);
}

int mn_votes_fast_i8_peer(const int8_t* a1, const int8_t* a0, int64_t lda, const double* inv_norm64, int64_t M,
                          const int8_t* b_limbs, int64_t ldb, int32_t b_limb_rows, const float* inv_scale,
                          const float* addend, int32_t Nn, int32_t G, float* const* peer_out_host, int32_t world,
                          int32_t src_rank, int32_t cols_per_rank, int64_t per, const double* den, int64_t ld_den,
                          const int32_t* col_den, const double* col_scale, mn_stream_t stream) {
    MN_CHECK_ARG(peer_out_host != nullptr, "mn_votes_fast_i8_peer: null pointer table");
    return votes_i8(0, PeerOut{peer_out_host, world, src_rank, cols_per_rank, per}, a1, a0, lda, inv_norm64, M, b_limbs, ldb,
                    b_limb_rows, inv_scale, addend, Nn, G, peer_out_host[src_rank], nullptr, per, den, ld_den, col_den,
                    col_scale, stream);
}

int mn_votes_den_i8(const int8_t* a1, const int8_t* a0, int64_t lda, const double* inv_norm64, int64_t M,
                    const int8_t* b_limbs, int64_t ldb, int32_t b_limb_rows, const float* inv_scale, const float* addend,
                    int32_t S, int32_t G, double* out, int64_t ldo, mn_stream_t stream) {
    MN_CHECK_ARG(out != nullptr, "mn_votes_den_i8: null output");
    const PeerOut none{nullptr, 0, 0, 0, 0};
    int rc = votes_i8(0, none, a1, a0, lda, inv_norm64, M, b_limbs, ldb, b_limb_rows, inv_scale, addend, S, G, nullptr, out, ldo,
                      nullptr, 0, nullptr, nullptr, stream);
    if (rc || G <= 127) return rc;          // one digit per rank: the four products above are the whole dot product
    return votes_i8(1, none, a1, a0, lda, inv_norm64, M, b_limbs, ldb, b_limb_rows, inv_scale, nullptr, S, G, nullptr, out, ldo,
                    nullptr, 0, nullptr, nullptr, stream);
}

static int corr_terms(mn::GemmParams& p, bool has_lo) {
    int t = 0;
    p.a_sel[t] = 0; p.b_sel[t] = 0; ++t;
    if (has_lo) {
        p.a_sel[t] = 0; p.b_sel[t] = 1; ++t;
        p.a_sel[t] = 1; p.b_sel[t] = 0; ++t;
        p.a_sel[t] = 1; p.b_sel[t] = 1; ++t;      // lo*lo too: the integer product stays exact
    }
    return t;
}

int mn_corr_hist(const mn_half_t* d_hi, const mn_half_t* d_lo, int64_t ldd, const float* inv_norm, int64_t N,
                 int64_t row0, int64_t row1, int32_t shard_rank, int32_t shard_world, int32_t G, int32_t nbins,
                 uint64_t* hist, uint32_t* spill, int64_t spill_tiles, mn_stream_t stream) {
    using namespace mn;
    if (check_planes("mn_corr_hist", d_hi, ldd, G)) return 1;
    MN_CHECK_ARG(inv_norm && hist && N >= 1 && row0 >= 0 && row1 <= N && row0 < row1, "mn_corr_hist: bad arguments");
    MN_CHECK_ARG(nbins >= 256 && nbins <= 32768 && (nbins & (nbins - 1)) == 0, "mn_corr_hist: nbins must be a power of two in [256, 32768]");
    MN_CHECK_ARG(shard_world >= 1 && shard_rank >= 0 && shard_rank < shard_world, "mn_corr_hist: bad shard");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.n_terms = corr_terms(p, d_lo != nullptr);
    p.tri = 1;
    const bool wide = nbins <= 16384 && !getenv("MN_GEMM_BN128");
    const int ctas = corr_ctas(wide), bm = BM * ctas;
    MN_CHECK_ARG((row0 % bm) == 0, "mn_corr_hist: row0 must be a multiple of %d", bm);
    p.mt0 = (int)(row0 / bm) + shard_rank;
    p.mt1 = (int)((row1 + bm - 1) / bm);
    p.mt_step = shard_world;
    p.inv_a = inv_norm;
    p.inv_b = inv_norm;
    p.nbins = nbins;
    p.hist = reinterpret_cast<unsigned long long*>(hist);
    MN_CHECK_ARG(spill_tiles >= 0 && (spill_tiles == 0 || spill != nullptr), "mn_corr_hist: spill_tiles without a spill buffer");
    MN_CHECK_ARG((reinterpret_cast<uintptr_t>(spill) & 15) == 0, "mn_corr_hist: spill must be 16-byte aligned");
    p.spill_tiles = (int)(spill_tiles > INT_MAX ? INT_MAX : spill_tiles) & ~(ctas - 1);    // whole pairs only
    p.spill = p.spill_tiles > 0 ? spill : nullptr;
    const __half* h = reinterpret_cast<const __half*>(d_hi);
    const __half* l = reinterpret_cast<const __half*>(d_lo);
    // rows beyond row1 inside the last tile belong to the next shard: mask them via M
    // 128 x 256 tiles move 25 % fewer operand bytes per flop through L2 (the measured bound); the table
    // must then leave room for >= 2 pipeline stages of 48 KB, i.e. nbins <= 16384
    if (wide && ctas == 4)
        return launch_gemm<256, EPI_HIST, 1, 4>(h, l, ldd, row1, h, l, ldd, N, G, p, (cudaStream_t)stream);
    if (wide && ctas == 2 && getenv("MN_HIST_EW16"))     // 16 epilogue warps: measured no faster (73.0 vs 70.8 ms)
        return launch_gemm<256, EPI_HIST, 1, 2, 16>(h, l, ldd, row1, h, l, ldd, N, G, p, (cudaStream_t)stream);
    if (wide && ctas == 2)
        return launch_gemm<256, EPI_HIST, 1, 2>(h, l, ldd, row1, h, l, ldd, N, G, p, (cudaStream_t)stream);
    if (wide)
        return launch_gemm<256, EPI_HIST, 1>(h, l, ldd, row1, h, l, ldd, N, G, p, (cudaStream_t)stream);
    return launch_gemm<128, EPI_HIST, 2>(h, l, ldd, row1, h, l, ldd, N, G, p, (cudaStream_t)stream);
}

int mn_corr_spill_layout(int64_t N, int64_t row0, int64_t row1, int32_t shard_rank, int32_t shard_world,
                         int32_t nbins, int64_t* n_tiles, int64_t* tile_bytes) {
    using namespace mn;
    MN_CHECK_ARG(n_tiles && tile_bytes && N >= 1 && row0 >= 0 && row1 <= N && row0 < row1 && (row0 % BM) == 0 &&
                     shard_world >= 1 && shard_rank >= 0 && shard_rank < shard_world,
                 "mn_corr_spill_layout: bad arguments");
    const bool wide = nbins <= 16384 && !getenv("MN_GEMM_BN128");
    const int bn = wide ? 256 : 128, ctas = corr_ctas(wide);
    // stored tiles are 128 x bn; a CTA pair's 256-row tile counts as two
    *n_tiles = corr_tile_count(N, row0, row1, shard_rank, shard_world, bn, BM * ctas) * ctas;
    *tile_bytes = (int64_t)BM * bn * 4;
    return 0;
}

int mn_corr_lut(const uint64_t* hist, int32_t nbins, int64_t n_diag, uint32_t* lut, mn_stream_t stream) {
    MN_CHECK_ARG(hist && lut && nbins >= 256 && nbins <= 32768, "mn_corr_lut: bad arguments");
    mn::corr_lut_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(hist), nbins,
                                                            (long long)n_diag, lut);
    MN_LAUNCH_CHECK();
    return 0;
}

int mn_corr_vote(const mn_half_t* d_hi, const mn_half_t* d_lo, int64_t ldd, const float* inv_norm, int64_t N,
                 int64_t row0, int64_t row1, int32_t shard_rank, int32_t shard_world, int32_t G, int32_t nbins,
                 const uint32_t* lut, const int32_t* label, int32_t L, uint64_t* votes, uint64_t* degree,
                 const uint32_t* spill, int64_t spill_tiles, mn_stream_t stream) {
    using namespace mn;
    if (check_planes("mn_corr_vote", d_hi, ldd, G)) return 1;
    MN_CHECK_ARG(inv_norm && lut && label && votes && degree && N >= 1 && L >= 1 && row0 >= 0 && row1 <= N && row0 < row1,
                 "mn_corr_vote: bad arguments");
    MN_CHECK_ARG(nbins >= 256 && nbins <= 32768 && (nbins & (nbins - 1)) == 0, "mn_corr_vote: nbins must be a power of two in [256, 32768]");
    MN_CHECK_ARG(shard_world >= 1 && shard_rank >= 0 && shard_rank < shard_world, "mn_corr_vote: bad shard");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.n_terms = corr_terms(p, d_lo != nullptr);
    // symmetric network: compute the upper triangle only and add every weight to both cells
    // (MN_VOTE_FULL=1 computes all N^2 tiles instead; debugging aid)
    p.tri = getenv("MN_VOTE_FULL") ? 0 : 1;
    const bool wide = nbins <= 16384 && !getenv("MN_GEMM_BN128");
    const int ctas = p.tri ? corr_ctas(wide) : 1, bm = BM * ctas;
    MN_CHECK_ARG((row0 % bm) == 0, "mn_corr_vote: row0 must be a multiple of %d", bm);
    p.mt0 = (int)(row0 / bm) + shard_rank;
    p.mt1 = (int)((row1 + bm - 1) / bm);
    p.mt_step = shard_world;
    p.inv_a = inv_norm;
    p.inv_b = inv_norm;
    p.nbins = nbins;
    p.lut = lut;
    p.label = label;
    p.L = L;
    p.votes = reinterpret_cast<unsigned long long*>(votes);
    p.degree = reinterpret_cast<unsigned long long*>(degree);
    const __half* h = reinterpret_cast<const __half*>(d_hi);
    const __half* l = reinterpret_cast<const __half*>(d_lo);
    MN_CHECK_ARG(spill_tiles >= 0 && (spill_tiles == 0 || spill != nullptr), "mn_corr_vote: spill_tiles without a spill buffer");
    int rc = 0;
    // tiles [0, n_sp) of the linear tile order were spilled by pass 1: stream them from HBM;
    // the tensor cores recompute only the rest
    const int64_t total = corr_tile_count(N, row0, row1, shard_rank, shard_world, wide ? 256 : 128, bm) * ctas;
    int64_t n_sp = (p.tri && spill) ? (spill_tiles < total ? spill_tiles : total) : 0;
    n_sp &= ~(int64_t)(ctas - 1);                  // whole pairs only, as in pass 1
    if (n_sp > 0) {
        p.spill = const_cast<uint32_t*>(spill);
        p.spill_tiles = (int)n_sp;
        p.M = row1;
        p.N = N;
        p.ctas = ctas;
        p.n_tiles = (int)((N + (wide ? 256 : 128) - 1) / (wide ? 256 : 128));
        p.band = gemm_band() / ctas;
        rc = wide ? launch_vote_spill<256>(p, (cudaStream_t)stream) : launch_vote_spill<128>(p, (cudaStream_t)stream);
        if (rc) return rc;
    }
    p.tile_base = (int)(n_sp / ctas);
    if (wide && ctas == 4)
        rc = launch_gemm<256, EPI_VOTE, 1, 4>(h, l, ldd, row1, h, l, ldd, N, G, p, (cudaStream_t)stream);
    else if (wide && ctas == 2)
        rc = launch_gemm<256, EPI_VOTE, 1, 2>(h, l, ldd, row1, h, l, ldd, N, G, p, (cudaStream_t)stream);
    else if (wide)
        rc = launch_gemm<256, EPI_VOTE, 1>(h, l, ldd, row1, h, l, ldd, N, G, p, (cudaStream_t)stream);
    else
        rc = launch_gemm<128, EPI_VOTE, 2>(h, l, ldd, row1, h, l, ldd, N, G, p, (cudaStream_t)stream);
    if (rc) return rc;
    // node degree = sum of a cell's votes over all labels (every voter carries a label)
    vote_rowsum_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p.votes, N, L, p.degree);
    MN_LAUNCH_CHECK();
    return 0;
}

}
