// Column access to a sparse expression matrix (scipy.sparse.csc_matrix layout) on the device.
//
// The supervised path ranks every cell over one gene set at a time (reference pymn/MetaNeighbor.py:84-93 slices
// `expression[:, gene_set]`), and variableGenes needs per-gene statistics over the cells of a study
// (variableGenes.py:22-27, with a CSC median at :109-135): both want a few COLUMNS of all cells.  In CSC form those
// are directly addressable, so a gene set's columns are expanded into a small dense [N, g] tile (zero fill +
// one scattered store per stored value) that the dense kernels then consume -- the matrix itself never
// densifies, on the host or on the device.
#include "common.cuh"

namespace mn {

// one warp per selected column: out[row, j] = value for every stored entry of column cols[j]
__global__ void csc_gather_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                  const float* __restrict__ data, const int32_t* __restrict__ cols, int n_cols,
                                  int64_t N, float* __restrict__ out, int64_t ld) {
    const int lane = threadIdx.x & 31;
    const int j = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    if (j >= n_cols) return;
    const int64_t c = cols ? (int64_t)cols[j] : (int64_t)j;
    const int64_t e0 = indptr[c], e1 = indptr[c + 1];
    for (int64_t e = e0 + lane; e < e1; e += 32) {
        const int64_t r = indices[e];
        if (r >= 0 && r < N) out[r * ld + j] = data[e];
    }
}

// out[c, p] = recv[cell_off[p] + c * col_stride]: the received vote chunks of the cell-sharded centroid path
// ([chunk][source rank][column][cell in chunk]) -> [column, global device row] in one pass
__global__ void votes_unpack_kernel(const float* __restrict__ recv, const int64_t* __restrict__ cell_off, int64_t N,
                                    int64_t col_stride, float* __restrict__ out, int64_t ldo) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int c = blockIdx.y;
    out[(int64_t)c * ldo + p] = recv[cell_off[p] + (int64_t)c * col_stride];
}

}  // namespace mn

extern "C" int mn_votes_unpack(const float* recv, const int64_t* cell_off, int64_t N, int32_t n_cols, int64_t col_stride,
                               float* out, int64_t ldo, mn_stream_t stream) {
    MN_CHECK_ARG(recv && cell_off && out && N >= 1 && n_cols >= 1 && n_cols <= 65535 && ldo >= N,
                 "mn_votes_unpack: bad arguments");
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)n_cols);
    mn::votes_unpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(recv, cell_off, N, col_stride, out, ldo);
    MN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mn_csc_gather_dense(const int64_t* indptr, const int32_t* indices, const float* data,
                                   const int32_t* cols, int32_t col0, int32_t n_cols, int64_t N, float* out, int64_t ld,
                                   mn_stream_t stream) {
    MN_CHECK_ARG(indptr && indices && data && out && n_cols >= 1 && N >= 1 && ld >= n_cols,
                 "mn_csc_gather_dense: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    MN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N * (size_t)ld, st));
    const int warps_per_block = 8;
    const unsigned blocks = (unsigned)((n_cols + warps_per_block - 1) / warps_per_block);
    // cols == NULL: the contiguous column range [col0, col0 + n_cols)
    mn::csc_gather_kernel<<<blocks, warps_per_block * 32, 0, st>>>(cols ? indptr : indptr + col0, indices, data, cols,
                                                                   n_cols, N, out, ld);
    MN_LAUNCH_CHECK();
    return 0;
}
