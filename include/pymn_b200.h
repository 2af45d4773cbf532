/*
 * pymn_b200 -- C ABI of the B200 (sm_100a) MetaNeighbor scoring hot path.
 *
 * Drop-in boundary.  The reference (gillislab/pyMN) is
 * pure Python: its "FFI" for this path is the set of NumPy / bottleneck calls
 * inside pymn/utils.py, pymn/MetaNeighborUS.py, pymn/MetaNeighbor.py and
 * pymn/trainModel.py.  Each entry point below replaces one of those call sites
 * (cited per function, paths relative to the reference's pymn/ directory) and
 * is what a ctypes stub inside those files would bind.
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer unless it says "host";
 *  - the caller owns all buffers (inputs, outputs, workspaces); nothing is
 *    allocated inside and nothing outlives a call except the enqueued work
 *    (one exception: a 4 KB pool of tile counters per device for the dynamic
 *    tile scheduler of the GEMM kernels, allocated on first use);
 *  - every call enqueues on the given stream (cudaStream_t passed as void*)
 *    and returns without synchronising; thread-compatible (one stream per call);
 *  - return value 0 = success, non-zero = error; mn_last_error() returns the
 *    message of the last failure on the calling thread;
 *  - "cells" are rows.  Device-side cell order is the caller's (the Python host
 *    sorts cells by (test study, label) so that studies and labels are
 *    contiguous row ranges);
 *  - fp16 "planes": centred doubled ranks d = 2*rank - (G+1) are integers with
 *    |d| < G.  d_hi = fp16(d) is exact for G <= 2048; otherwise d = d_hi + d_lo
 *    exactly with both fp16.  These feed the tcgen05 kind::f16 tensor-core GEMMs
 *    with FP32 accumulation in TMEM (integer-exact operands instead of a
 *    3xTF32 split of normalised floats);
 *  - int8 "limbs" (the centroid paths): the same integers as d = 128 a1 + a0,
 *    a0 in [-64, 63]; they feed the tcgen05 kind::i8 GEMM with int32
 *    accumulators, whose dot products are exact.
 */
#ifndef PYMN_B200_H
#define PYMN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mn_stream_t;   /* cudaStream_t */
typedef uint16_t mn_half_t;  /* IEEE binary16 bit pattern (__half) */

/* ---- library ------------------------------------------------------------- */
const char* mn_last_error(void);
int mn_abi_version(void);              /* 3 = round 2 (limb / fixed-point / peer / batch entry points added) */
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t mn_launch_count(void);

/* ---- K1: per-cell rank + normalise --------------------------------------
 * Replaces normalize_cells (utils.py:122-148: bottleneck.rankdata(axis=1),
 * centre, L2-normalise) and the ranking half of create_nw_spearman (utils.py:69).
 *   X          [*, ldx] float32 row-major cells x genes (any finite floats)
 *   row_index  int32[M] or NULL: output row i is input row row_index[i]
 *   col_index  int32[G] or NULL: gene g is input column col_index[g]
 *                (per-gene-set sub-setting, MetaNeighbor.py:85)
 *   d_hi,d_lo  fp16 [M, ldd] planes (d_lo may be NULL when G <= 2048);
 *                columns G..ldd-1 are written as zero
 *   inv_norm   float32[M]: 1/||d||_2, or 0 for a dropped cell
 *   flags      uint8[M]: bit0 = zero variance (NaN row in the reference,
 *                utils.py:146-147), bit1 = row sum <= 0 (trainModel.py:36,
 *                MetaNeighbor.py:121), bit2 = dropped (d = 0, inv_norm = 0)
 *   xnorm      float32 [M, G] or NULL: the reference's normalised matrix
 *   rank2      int32 [M, G] or NULL: 2*rank (bit-exact integer encoding)
 *   drop_mode  0: drop zero-variance cells; 1: drop cells with row sum <= 0
 *                (zero-variance cells are always zeroed; flags tell them apart)
 */
int mn_rank_normalize(const float* X, int64_t ldx, const int32_t* row_index, const int32_t* col_index,
                      int64_t M, int32_t G, mn_half_t* d_hi, mn_half_t* d_lo, int64_t ldd,
                      float* inv_norm, uint8_t* flags, float* xnorm, int32_t* rank2,
                      int32_t drop_mode, mn_stream_t stream);

/* Same, reading the expression matrix in CSR form (scipy.sparse.csr_matrix: indptr int64 [n_rows+1],
 * indices int32 [nnz], data float32 [nnz]) -- the layout AnnData keeps X in.  Replaces the
 * X.toarray() + rankdata of normalize_cells (utils.py:136-141): only the stored values cross
 * PCIe / HBM (8 B per non-zero instead of 4 B per gene).  Implicit entries are zeros; explicit
 * zeros are allowed; indices must be unique within a row.  No column sub-setting. */
int mn_rank_normalize_csr(const int64_t* indptr, const int32_t* indices, const float* data,
                          const int32_t* row_index, int64_t M, int32_t G, mn_half_t* d_hi, mn_half_t* d_lo,
                          int64_t ldd, float* inv_norm, uint8_t* flags, float* xnorm, int32_t* rank2,
                          int32_t drop_mode, mn_stream_t stream);

/* K1 straight into the operand format of the centroid paths: int8 limb planes a1, a0 [M, ldk] (d = 128 a1 + a0,
 * a0 in [-64, 63]; G <= 127: a1 = d, a0 = 0; ldk a multiple of 128, padding zero) and the cell's scale in double,
 * inv_norm64[m] = 1 / ||d_m|| (0 = dropped cell) -- exactly what mn_limb_split_cells derives from the fp16 planes,
 * without writing the planes or running the split pass.  inv_norm / flags as in mn_rank_normalize. */
int mn_rank_normalize_limbs(const float* X, int64_t ldx, const int32_t* row_index, const int32_t* col_index,
                            int64_t M, int32_t G, int8_t* a1, int8_t* a0, int64_t ldk, double* inv_norm64,
                            float* inv_norm, uint8_t* flags, int32_t drop_mode, mn_stream_t stream);
int mn_rank_normalize_csr_limbs(const int64_t* indptr, const int32_t* indices, const float* data,
                                const int32_t* row_index, int64_t M, int32_t G, int8_t* a1, int8_t* a0, int64_t ldk,
                                double* inv_norm64, float* inv_norm, uint8_t* flags, int32_t drop_mode,
                                mn_stream_t stream);

/* ---- gene statistics for variableGenes (the step before the hot path) -------------
 * Replaces bottleneck.median(adata.X, axis=0) and np.var(adata.X, axis=0) over the cells of each
 * study (variableGenes.py:22-27, called per study at :69-72).
 *   X          [*, ldx] float32 row-major cells x genes (finite values)
 *   row_index  int32[M] or NULL: position i of the study-sorted order is input row row_index[i]
 *   seg_off    int32[n_seg+1]: study t owns positions [seg_off[t], seg_off[t+1])
 *   median     float64 [n_seg, G]: per-gene median (mean of the two middle order statistics; exact)
 *   variance   float64 [n_seg, G]: per-gene population variance (ddof = 0), accumulated in float64
 * An empty study yields NaN.  4-pass radix select: HBM-bound, 4 x 4 B per element. */
int mn_gene_stats(const float* X, int64_t ldx, const int32_t* row_index, const int32_t* seg_off,
                  int32_t n_seg, int32_t G, double* median, double* variance, mn_stream_t stream);

/* Columns of a sparse matrix in CSC form (indptr int64 [G + 1], indices int32 = row ids, data float32) expanded
 * into a dense row-major tile: out[r, j] = X[r, cols[j]] (cols == NULL: columns col0 .. col0 + n_cols - 1), zero
 * elsewhere; out float32 [N, ld], ld >= n_cols.  Feeds the gene-set slices of the supervised path
 * (MetaNeighbor.py:84-93) and the per-gene statistics of variableGenes (variableGenes.py:109-135) from a sparse
 * AnnData.X without densifying the matrix. */
int mn_csc_gather_dense(const int64_t* indptr, const int32_t* indices, const float* data, const int32_t* cols,
                        int32_t col0, int32_t n_cols, int64_t N, float* out, int64_t ld, mn_stream_t stream);

/* ---- K2: cluster centroids ----------------------------------------------
 * Replaces X_norm @ onehot (MetaNeighborUS.py:284, trainModel.py:43,
 * MetaNeighbor.py:170 inner product) and n_cells_per_cluster (:288).
 * Cells are sorted by label: label l owns rows [row_off[l], row_off[l+1]).
 *   centroids  float32 [L, ldc]  (row l = sum over the label's kept cells of d*inv_norm)
 *   n_valid    float32 [L]       number of kept cells (inv_norm > 0) per label
 *   inv_norm64 double [M] or NULL: the cells' scale in double (mn_limb_split_cells); with it every term is the
 *              exact integer rank times the double scale and only the final float32 store rounds
 */
int mn_centroids(const mn_half_t* d_hi, const mn_half_t* d_lo, int64_t ldd, const float* inv_norm,
                 const double* inv_norm64, const int32_t* row_off, int32_t L, int32_t G, float* centroids, int64_t ldc,
                 float* n_valid, mn_stream_t stream);

/* Fixed-point form of mn_centroids (the product path): every term d * inv_norm64 is rounded once to a multiple
 * of 2^-38 and summed as a 64-bit integer, so the sums are exact, independent of the order of the cells, of the
 * launch geometry and of how the cells are dealt to GPUs (partial sums of different ranks add with an integer
 * all-reduce): every world size sees bit-identical centroids.
 *   cq [L, ldc] int64, nq [L] int32: ACCUMULATED into (the caller zero-fills them once)
 *   max_label_rows: the largest label's row count (sizes the row slices per label; 0 = one slice)
 * mn_group_sum_fixed: the study centroids / study sizes as integer sums of label rows (n_in / n_out may be NULL)
 * mn_centroids_finish: float32 centroids = cq * 2^-38, n_valid = (float)nq, for rows [0, R) */
int mn_centroids_fixed(const mn_half_t* d_hi, const mn_half_t* d_lo, int64_t ldd, const double* inv_norm64,
                       const int32_t* row_off, int32_t L, int32_t G, int64_t max_label_rows, int64_t* cq, int64_t ldc,
                       int32_t* nq, mn_stream_t stream);
/* the same from the int8 limb planes of mn_rank_normalize_limbs / mn_limb_split_cells (d = 128 a1 + a0) */
int mn_centroids_fixed_i8(const int8_t* a1, const int8_t* a0, int64_t ldk, const double* inv_norm64,
                          const int32_t* row_off, int32_t L, int32_t G, int64_t max_label_rows, int64_t* cq, int64_t ldc,
                          int32_t* nq, mn_stream_t stream);
int mn_group_sum_fixed(const int64_t* in, int64_t ld, int32_t R, int32_t width, const int32_t* group, int32_t n_groups,
                       int64_t* out, const int32_t* n_in, int32_t* n_out, mn_stream_t stream);
int mn_centroids_finish(const int64_t* cq, const int32_t* nq, int32_t R, int64_t ld, float* centroids, float* n_valid,
                        mn_stream_t stream);

/* out[g_out, :] = sum over rows r with group[r] == g_out of in[r, :]   (rows with group < 0 skipped).
 * Replaces cluster_centroids @ study_matrix and n_cells_per_cluster @ study_matrix
 * (MetaNeighborUS.py:349-350).  in/out float32 [R, ld] / [n_groups, ld]; width = columns used. */
int mn_group_sum(const float* in, int64_t ld, int32_t R, int32_t width, const int32_t* group,
                 int32_t n_groups, float* out, mn_stream_t stream);

/* Split float32 rows into two fp16 planes with a per-row power-of-two scale:
 *   in[r, g] ~= (hi[r, g] + lo[r, g]) * inv_scale[r]   (22+ significant bits).
 * Prepares the centroid operand of mn_votes_fast.  Columns G..ldh-1 zeroed. */
int mn_split_f16(const float* in, int64_t ld, int32_t R, int32_t G, mn_half_t* hi, mn_half_t* lo,
                 int64_t ldh, float* inv_scale, mn_stream_t stream);

/* ---- K3: centroid votes (tcgen05 GEMM) ------------------------------------
 * Replaces votes = X_test.T @ centroids (+ n_cells) and node_degree =
 * X_test.T @ study_centroids (+ n_cells_per_study) (MetaNeighborUS.py:361-366;
 * MetaNeighbor.py:170-174).
 *   out[n, m] = inv_norm[m] * inv_scale[n] * sum_g (a_hi+a_lo)[m, g] * (b_hi+b_lo)[n, g] + addend[n]
 *   a_*: [M, lda] fp16 cell planes (a_lo may be NULL); b_*: [Nn, ldb] fp16 centroid planes
 *   out: float32 column-major [Nn, ldo] (ldo >= M): column n is contiguous over cells
 */
int mn_votes_fast(const mn_half_t* a_hi, const mn_half_t* a_lo, int64_t lda, const float* inv_norm, int64_t M,
                  const mn_half_t* b_hi, const mn_half_t* b_lo, int64_t ldb, const float* inv_scale,
                  const float* addend, int32_t Nn, int32_t G, float* out, int64_t ldo, mn_stream_t stream);

/* ---- K3, exact form: centroid votes on the int8 tensor cores -----------------
 * Same product as mn_votes_fast (MetaNeighborUS.py:361-366; MetaNeighbor.py:170-174), computed as integer limb
 * products with int32 accumulators: the dot product of the cell's integer ranks with the centroid (quantised
 * to 21 bits of its largest entry) is exact, hence independent of tiling, k order and GPU count.
 *   mn_limb_split_cells      fp16 planes of K1 -> int8 limb planes a1, a0 [M, ldk] (d = 128 a1 + a0; G <= 127:
 *                            a1 = d, a0 = 0) and the
 *                            double-precision scale inv_norm64[m] = 1 / ||d_m|| (0 = dropped cell)
 *   mn_limb_split_centroids  float32 rows [R, ld] -> three stacked int8 limb planes
 *                            limbs[(l * rows_pad + r) * ldk + g], l = 0 .. 3 (weights 128^3 .. 1), with a per-row
 *                            power-of-two scale: in[r, g] ~= (128^3 b3 + 128^2 b2 + 128 b1 + b0) * inv_scale[r]
 *                            (28 bits of the row's largest entry); the caller zero-fills rows R..rows_pad-1
 *                            q32 (optional): the same 28-bit integers as one int32 plane [R, ldk]
 *                            (limbs may be NULL when only q32 is wanted)
 *   mn_votes_den64           den[s, m] = inv_norm64[m] * inv_scale[s] * sum_g d[m, g] * q32[s, g] + addend[s] for the
 *                            few study-centroid rows (node degree, MetaNeighborUS.py:363-366): exact 64-bit
 *                            integer dot products on the CUDA cores, scaled in double
 *   mn_votes_fast_i8         val = dot(m, n) * inv_norm64[m] * inv_scale[n] + addend[n], evaluated in double;
 *                            G <= 127 (one digit per rank): dot = the exact integer product; beyond, dot = the
 *                            exact product minus the weight-1 digit cross term sum_g a0 b0 (four accumulators
 *                            fit TMEM): a zero-mean sum, ~3e-7 of the product at G = 128, ~3e-9 at G = 3,000;
 *                            den != NULL: val = val / den[col_den[n] * ld_den + m] * col_scale[n] - 1 -- the
 *                              node-degree ratio (MetaNeighborUS.py:365-371) taken in double and CENTRED on a
 *                              per-column reference ratio 1 / col_scale[n] (any positive constant: a monotone
 *                              map per column, to which the AUROC is invariant), so that the single float32
 *                              rounding of the result resolves ~1e-9 of the ratio where the cells crowd;
 *                            den == NULL: val = val - col_scale[n] (NULL = 0);
 *                            out64 != NULL: val stored as double [Nn, ldo] (the denominators themselves),
 *                            else rounded once into out, float32 column-major [Nn, ldo]
 * ldk, ldb: row pitch in bytes, multiples of 128 (padding zero); b_limb_rows (= rows_pad): multiple of 128. */
int mn_limb_split_cells(const mn_half_t* d_hi, const mn_half_t* d_lo, int64_t ldd, const float* inv_norm, int64_t M,
                        int32_t G, int8_t* a1, int8_t* a0, int64_t ldk, double* inv_norm64, mn_stream_t stream);
int mn_limb_split_centroids(const float* in, int64_t ld, int32_t R, int32_t G, int8_t* limbs, int64_t ldk,
                            int64_t rows_pad, float* inv_scale, int32_t* q32, mn_stream_t stream);
int mn_votes_den64(const int8_t* a1, const int8_t* a0, int64_t ldk, const double* inv_norm64, int64_t M, int32_t G,
                   const int32_t* q32, const float* inv_scale, const float* addend, int32_t S, double* out,
                   int64_t ldo, mn_stream_t stream);
int mn_votes_fast_i8(const int8_t* a1, const int8_t* a0, int64_t lda, const double* inv_norm64, int64_t M,
                     const int8_t* b_limbs, int64_t ldb, int32_t b_limb_rows, const float* inv_scale,
                     const float* addend, int32_t Nn, int32_t G, float* out, double* out64, int64_t ldo,
                     const double* den, int64_t ld_den, const int32_t* col_den, const double* col_scale,
                     mn_stream_t stream);
/* mn_votes_fast_i8 fused with the votes exchange of the multi-GPU centroid paths: the Nn = world *
 * cols_per_rank train columns are arranged by owner (rank d owns columns [d cols_per_rank, (d + 1) cols_per_rank)),
 * and the epilogue stores the votes of this rank's M cells for column n straight into the OWNER's receive buffer --
 * a peer pointer, written over NVLink tile by tile while the tensor cores work on the next tiles -- at
 * peer_out[d][(src_rank * cols_per_rank + n % cols_per_rank) * per + m].  peer_out_host: HOST array of `world`
 * device pointers (float32 [world * cols_per_rank, per] each, mapped into this process: CUDA IPC / peer access are
 * the caller's business); per >= M.  No NCCL collective moves the votes; the caller fences the ranks before the
 * launch (the receive buffers are free) and after it (all writes have landed). */
int mn_votes_fast_i8_peer(const int8_t* a1, const int8_t* a0, int64_t lda, const double* inv_norm64, int64_t M,
                          const int8_t* b_limbs, int64_t ldb, int32_t b_limb_rows, const float* inv_scale,
                          const float* addend, int32_t Nn, int32_t G, float* const* peer_out_host, int32_t world,
                          int32_t src_rank, int32_t cols_per_rank, int64_t per, const double* den, int64_t ld_den,
                          const int32_t* col_den, const double* col_scale, mn_stream_t stream);
/* Node degree on the int8 tensor cores, FULLY exact: den[s, m] = inv_norm64[m] * inv_scale[s] * sum_g d[m, g] q[s, g]
 * + addend[s] for the S study-centroid rows whose limbs are b_limbs (mn_limb_split_centroids): the seven digit
 * products of mn_votes_fast_i8 plus, in a second launch that adds into `out`, the weight-1 product a0 * b0 that the
 * vote GEMM leaves out (a denominator shared by every train column of a study is worth the exact value).  Same result
 * as mn_votes_den64 up to the last bit of a double; out: double [S, ldo]. */
int mn_votes_den_i8(const int8_t* a1, const int8_t* a0, int64_t lda, const double* inv_norm64, int64_t M,
                    const int8_t* b_limbs, int64_t ldb, int32_t b_limb_rows, const float* inv_scale, const float* addend,
                    int32_t S, int32_t G, double* out, int64_t ldo, mn_stream_t stream);

/* ---- K4/K5: full cell x cell network, never materialised -----------------
 * Replaces create_nw_spearman + rank (utils.py:35-75: np.corrcoef, one ranking
 * of all N^2 entries) and the vote products that consume it
 * (MetaNeighborUS.py:165-169; MetaNeighbor.py:195-217).
 *
 * Pass 1: mn_corr_hist   -- rho tiles on tensor cores -> exact histogram of rho
 *                           over pairs i<j, i in [row0,row1), both cells kept
 * LUT   : mn_corr_lut    -- histogram -> fixed-point CDF table W = (rank-1)/max
 * Pass 2: mn_corr_vote   -- rho tiles again (or from the rho spill) -> W via LUT -> integer vote sums
 *   votes  uint64 [N, L]  sum_j W_fix(rho_ij) [label(j)==l]   (W_fix = W * 2^20; j != i)
 *   degree uint64 [N]     sum_j W_fix(rho_ij)                 (j != i)
 * Both accumulate with integer atomics (exact, order-independent, so the
 * result is bit-identical for any tile schedule or GPU count).  The caller
 * zeroes hist / votes.  Every cell must carry a label in [0, L).  Pass 2 computes
 * the upper triangle only (the network is symmetric) and adds each weight to both
 * cells; `degree` is overwritten with the row sums of `votes`.
 * Multi-GPU: of the 128-row tiles of [row0, row1) a launch handles those with
 * (tile index) % shard_world == shard_rank, so row tiles are dealt round-robin
 * (balances the triangular pass 1); single GPU = (0, 1).
 */
int mn_corr_hist(const mn_half_t* d_hi, const mn_half_t* d_lo, int64_t ldd, const float* inv_norm,
                 int64_t N, int64_t row0, int64_t row1, int32_t shard_rank, int32_t shard_world,
                 int32_t G, int32_t nbins, uint64_t* hist, uint32_t* spill, int64_t spill_tiles,
                 mn_stream_t stream);
/* The rho spill (optional; spill = NULL, spill_tiles = 0 turns it off).  The 180 GB of HBM are large
 * enough to keep what pass 1 computes: the first `spill_tiles` tiles (128 x 256 pairs each, in the
 * kernel's own tile order) leave their bin positions -- 16.16 fixed point, 4 B per cell pair, 80 GB for
 * 200,000 cells -- in `spill`, and pass 2 streams them back (an HBM-bound kernel) instead of
 * recomputing those tiles on the tensor cores; tiles beyond `spill_tiles` are recomputed.  Both
 * routes feed the same integers into the same weight map, so the vote sums are bit-identical for
 * any `spill_tiles`.  mn_corr_spill_layout reports the tile count and size of one shard: a buffer of
 * n_tiles * tile_bytes holds everything.  The same (N, row0, row1, shard, nbins) must be passed to
 * the layout query, mn_corr_hist and mn_corr_vote; row0 must be a multiple of 512 (the tallest tile of
 * the GEMM forms).  spill_tiles is rounded down to whole cluster tiles (2 stored tiles per CTA pair). */
int mn_corr_spill_layout(int64_t N, int64_t row0, int64_t row1, int32_t shard_rank, int32_t shard_world,
                         int32_t nbins, int64_t* n_tiles /* host */, int64_t* tile_bytes /* host */);
/* hist -> packed CDF table.  n_diag = number of diagonal entries (all cells; they tie at rho = 1, utils.py:72).
 * lut uint32 [nbins + 2]: lut[b] = (W0 << 11) | rise for bin b, W in 2^20 fixed point:
 *   rise > 0: W0 = weight at the lower edge of the bin, rising linearly by `rise` (< 2047) across it;
 *   rise = 0: the bin holds >= 0.2 % of all pairs -- one tie block of heavily tied correlations (gene sets of a
 *             handful of genes) -- and W0 is the block's MID-rank weight, the rank bottleneck.rankdata gives
 *             every member of a tie block (utils.py:45).
 * lut[nbins] = weight at the top edge (rise 0); lut[nbins + 1] = number of tie-block bins (diagnostic).
 * Bin of a correlation: floor(rho * (nbins/2 - 0.81) + nbins/2 + 0.043) -- the odd constants keep simple
 * fractions (rho = 1/2, 1/4 ...) away from the bin edges, where FP32 rounding would split their tie block. */
int mn_corr_lut(const uint64_t* hist, int32_t nbins, int64_t n_diag, uint32_t* lut /* [nbins+2] */,
                mn_stream_t stream);
int mn_corr_vote(const mn_half_t* d_hi, const mn_half_t* d_lo, int64_t ldd, const float* inv_norm,
                 int64_t N, int64_t row0, int64_t row1, int32_t shard_rank, int32_t shard_world,
                 int32_t G, int32_t nbins, const uint32_t* lut, const int32_t* label, int32_t L,
                 uint64_t* votes, uint64_t* degree, const uint32_t* spill, int64_t spill_tiles,
                 mn_stream_t stream);

/* votes uint64 [N, L] + degree uint64 [N] -> float32 column-major [L, ldo] vote matrix.
 * Adds the diagonal (W_ii = 1, utils.py:74) to votes[i, label[i]] and degree[i];
 * ndn != 0 divides by the node degree (MetaNeighborUS.py:167-169).  degree == NULL: the node degree is taken as
 * the row sum of votes (every voter carries a label) -- after a multi-GPU sum of the votes no second
 * collective is needed for the degree. */
int mn_votes_finalize(const uint64_t* votes, const uint64_t* degree, const int32_t* label, int64_t N,
                      int32_t L, int32_t ndn, float* out, int64_t ldo, mn_stream_t stream);

/* Test labels of the centroid paths: out_label[i] = cell_label[i], or -1 for a dropped cell (inv_norm[i] == 0:
 * zero-variance cells leave the ranking, MetaNeighborUS.py:273-281; score_low_mem drops cells whose sum is <= 0,
 * MetaNeighbor.py:121).  poison (optional, one byte, set to 1, never cleared): some cell is constant but NON-ZERO
 * (K1 flag bit 0 without bit 1) -- in score_low_mem such a cell normalises to a NaN row that turns the gene
 * set's whole result into NaN (MetaNeighbor.py:121-130, utils.py:146-147). */
int mn_mask_dropped(const int32_t* cell_label, const float* inv_norm, const uint8_t* flags, int64_t N,
                    int32_t* out_label, uint8_t* poison, mn_stream_t stream);

/* Leave-one-study-out combination for the supervised paths (MetaNeighbor.py:198-217
 * and :134-147,170-175).  Input: per-cell sums over joint (study, type) labels,
 * joint id = study * K + type, either uint64 fixed-point [N, S*K] (full network,
 * is_fixed = 1; `degree` uint64 [N] is the node degree incl. the diagonal = W_ii)
 * or float32 column-major [S*K, ldi] (centroid votes, is_fixed = 0; n_joint
 * float32 [S*K] = kept cells per joint label).  Output float32 column-major [K, ldo]:
 *   out[k, i] = sum_{s' != study(i)} in[i, (s', k)]  (+ n_k of other studies)  / degree
 */
int mn_loso_finalize(const void* in, int64_t ldi, int32_t is_fixed, const uint64_t* degree,
                     const float* n_joint, const int32_t* cell_study, int64_t N, int32_t S, int32_t K,
                     int32_t ndn, float* out, int64_t ldo, mn_stream_t stream);

/* ---- K6: rank-sum AUROC --------------------------------------------------
 * Replaces compute_aurocs (utils.py:151-178), compute_aurocs_default
 * (MetaNeighborUS.py:194-213) and the rank-sum block of score_default
 * (MetaNeighbor.py:232-238).
 *   votes     float32 column-major [ncols, ldv]; column l = votes of all cells for train label l
 *   denom     float32 column-major [n_denom, ldv] or NULL; col_denom int32[ncols] (or NULL):
 *               key = votes[l, i] / denom[col_denom[l], i]   (node degree per train study,
 *               MetaNeighborUS.py:365-371)
 *   seg_off   int32[n_seg+1]: test study t owns cells [seg_off[t], seg_off[t+1])
 *   cell_label int32[M]: test-label id of each cell (0..n_labels-1), or < 0 = cell excluded
 *   auroc     float64 [n_labels, ncols]: rows = test labels, cols = train labels
 *               (a label's row is filled from the segment its cells are in)
 *   n_pos     int32 [n_labels] or NULL: kept cells per test label
 *   workspace from mn_auroc_workspace_bytes(ncols, M, n_seg, n_labels)
 * Ranks are exact average-tie ranks of the float32 keys (integer 2*rank sums).
 */
int64_t mn_auroc_workspace_bytes(int32_t ncols, int64_t M, int32_t n_seg, int32_t n_labels);
int mn_auroc(const float* votes, int64_t ldv, int32_t ncols, const float* denom, const int32_t* col_denom,
             const int32_t* seg_off, int32_t n_seg, const int32_t* cell_label, int32_t n_labels, int64_t M,
             double* auroc, int32_t* n_pos, void* workspace, int64_t workspace_bytes, mn_stream_t stream);

/* ---- K7: one-vs-best ------------------------------------------------------
 * Replaces compute_1v1_aurocs + find_top_candidates (utils.py:194-248).
 * Cells must be sorted by label inside each segment: label c owns rows
 * [lab_row_off[c], lab_row_off[c+1]) (excluded cells have inv. keys skipped via
 * cell_label < 0).  seg_lab_off int32[n_seg+1]: segment t owns labels
 * [seg_lab_off[t], seg_lab_off[t+1]).  out float64 [n_labels, ncols] must be
 * pre-filled with NaN by the caller; two cells per (segment, column) are written.
 */
int mn_one_vs_best(const float* votes, int64_t ldv, int32_t ncols, const float* denom,
                   const int32_t* col_denom, const int32_t* seg_lab_off, int32_t n_seg,
                   const int32_t* lab_row_off, const int32_t* cell_label, int32_t n_labels,
                   const double* auroc, double* out, mn_stream_t stream);

/* Multi-GPU centroid paths: after the all-to-all of the votes, rank r holds its train columns for the
 * cells of all ranks as [chunk][source rank][column][cell in chunk]; this puts them into [column, global device row]
 * order for K6 / K7: out[c, p] = recv[cell_off[p] + c * col_stride] (cell_off: int64 [N], built by the host from
 * the cell shard bookkeeping). */
int mn_votes_unpack(const float* recv, const int64_t* cell_off, int64_t N, int32_t n_cols, int64_t col_stride,
                    float* out, int64_t ldo, mn_stream_t stream);

/* ---- supervised MetaNeighbor, fast_version: all gene sets of one rank in one call --------------------------
 * Replaces the per-gene-set loop of MetaNeighbor.py:84-93 over score_low_mem (:106-148) / compute_votes (:151-176):
 * for each set j (gene columns all_cols[set_off[j] .. set_off[j+1]) of X, at most g_max of them) the host loop --
 * in C, no interpreter or FFI cost per set -- enqueues mn_rank_normalize_limbs (drop_mode 1, cells gathered through
 * perm into joint (study, type) order) -> mn_centroids_fixed_i8 / _finish over the S*K joint labels ->
 * mn_limb_split_centroids + mn_votes_fast_i8 -> mn_loso_finalize -> mn_mask_dropped -> mn_auroc.
 *   X float32 [*, ldx] device; perm, lab_row_off [S*K+1], cell_label, cell_study, seg_off [S+1]: device int32
 *   all_cols: device int32; set_off_host: HOST int64 [n_sets + 1]
 *   tables: device float64 [n_sets, S*K, K] (rows = joint test labels); poison: device uint8 [n_sets]
 *   workspace: device, mn_supervised_fast_workspace_bytes(N, g_max, S, K) bytes.  Nothing synchronises. */
int64_t mn_supervised_fast_workspace_bytes(int64_t N, int32_t g_max, int32_t S, int32_t K);
int mn_supervised_fast_sets(const float* X, int64_t ldx, const int32_t* perm, int64_t N, const int32_t* all_cols,
                            const int64_t* set_off_host, int32_t n_sets, int32_t g_max, const int32_t* lab_row_off,
                            const int32_t* cell_label, const int32_t* cell_study, const int32_t* seg_off, int32_t S,
                            int32_t K, int32_t ndn, int64_t max_label_rows, void* workspace, int64_t workspace_bytes,
                            double* tables, uint8_t* poison, mn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PYMN_B200_H */
