// Supervised MetaNeighbor, fast_version (score_low_mem): all gene sets of one rank in ONE call.
//
// Replaces the per-gene-set loop of reference pymn/MetaNeighbor.py:84-93 calling score_low_mem (:106-148) and
// compute_votes (:151-176).  A gene set is ~0.1 ms of device work; driven from Python, the dozen launches of a
// set cost twice that in interpreter and FFI overhead.  Here the host loop is C: per set it enqueues exactly the
// entry points the Python pipeline would call (K1 into int8 limbs over the set's gene columns -> fixed-point joint
// (study, type) centroids -> exact int8 vote GEMM -> leave-one-study-out combination -> test labels / NaN-poison
// flag -> rank-sum AUROC), on caller-provided workspace, without synchronising.
#include "common.cuh"

namespace {

struct SupLayout {
    int64_t ldk, ldd, rows_pad;
    int64_t off_a1, off_a0, off_inv64, off_inv, off_flags, off_cq, off_nq, off_c, off_nv, off_limbs, off_scale, off_votes,
        off_vmat, off_lab, off_ws, ws_bytes, total;
};

int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

SupLayout sup_layout(int64_t N, int32_t g_max, int32_t S, int32_t K) {
    SupLayout L;
    const int64_t SK = (int64_t)S * K;
    L.ldk = ((int64_t)g_max + 127) / 128 * 128;
    L.ldd = ((int64_t)g_max + 63) / 64 * 64;
    L.rows_pad = (SK + 127) / 128 * 128;
    int64_t o = 0;
    auto take = [&](int64_t bytes) { const int64_t at = o; o += align256(bytes); return at; };
    L.off_a1 = take(N * L.ldk);
    L.off_a0 = take(N * L.ldk);
    L.off_inv64 = take(N * 8);
    L.off_inv = take(N * 4);
    L.off_flags = take(N);
    L.off_cq = take(SK * L.ldd * 8);
    L.off_nq = take(SK * 4);
    L.off_c = take(SK * L.ldd * 4);
    L.off_nv = take(SK * 4);
    L.off_limbs = take(4 * L.rows_pad * L.ldk);
    L.off_scale = take(SK * 4);
    L.off_votes = take(SK * N * 4);
    L.off_vmat = take((int64_t)K * N * 4);
    L.off_lab = take(N * 4);
    L.ws_bytes = mn_auroc_workspace_bytes(K, N, S, (int32_t)SK);
    L.off_ws = take(L.ws_bytes);
    L.total = o;
    return L;
}

}  // namespace

extern "C" int64_t mn_supervised_fast_workspace_bytes(int64_t N, int32_t g_max, int32_t S, int32_t K) {
    if (N < 1 || g_max < 1 || S < 2 || K < 1) return -1;
    return sup_layout(N, g_max, S, K).total;
}

extern "C" int mn_supervised_fast_sets(const float* X, int64_t ldx, const int32_t* perm, int64_t N,
                                       const int32_t* all_cols, const int64_t* set_off_host, int32_t n_sets,
                                       int32_t g_max, const int32_t* lab_row_off, const int32_t* cell_label,
                                       const int32_t* cell_study, const int32_t* seg_off, int32_t S, int32_t K,
                                       int32_t ndn, int64_t max_label_rows, void* workspace, int64_t workspace_bytes,
                                       double* tables, uint8_t* poison, mn_stream_t stream) {
    MN_CHECK_ARG(X && perm && all_cols && set_off_host && lab_row_off && cell_label && cell_study && seg_off &&
                     workspace && tables && poison,
                 "mn_supervised_fast_sets: null pointer");
    MN_CHECK_ARG(N >= 1 && n_sets >= 0 && S >= 2 && K >= 1 && g_max >= 1, "mn_supervised_fast_sets: bad shape");
    const SupLayout L = sup_layout(N, g_max, S, K);
    MN_CHECK_ARG(workspace_bytes >= L.total, "mn_supervised_fast_sets: workspace too small (%lld < %lld)",
                 (long long)workspace_bytes, (long long)L.total);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    int8_t* a1 = reinterpret_cast<int8_t*>(w + L.off_a1);
    int8_t* a0 = reinterpret_cast<int8_t*>(w + L.off_a0);
    double* inv64 = reinterpret_cast<double*>(w + L.off_inv64);
    float* inv = reinterpret_cast<float*>(w + L.off_inv);
    uint8_t* flags = w + L.off_flags;
    int64_t* cq = reinterpret_cast<int64_t*>(w + L.off_cq);
    int32_t* nq = reinterpret_cast<int32_t*>(w + L.off_nq);
    float* C = reinterpret_cast<float*>(w + L.off_c);
    float* nv = reinterpret_cast<float*>(w + L.off_nv);
    int8_t* limbs = reinterpret_cast<int8_t*>(w + L.off_limbs);
    float* scale = reinterpret_cast<float*>(w + L.off_scale);
    float* votes = reinterpret_cast<float*>(w + L.off_votes);
    float* vmat = reinterpret_cast<float*>(w + L.off_vmat);
    int32_t* lab = reinterpret_cast<int32_t*>(w + L.off_lab);
    void* aws = w + L.off_ws;
    const int32_t SK = S * K;
    MN_CUDA(cudaMemsetAsync(poison, 0, (size_t)(n_sets > 0 ? n_sets : 1), st));
    for (int32_t j = 0; j < n_sets; ++j) {
        const int64_t c0 = set_off_host[j], c1 = set_off_host[j + 1];
        const int32_t G = (int32_t)(c1 - c0);
        MN_CHECK_ARG(G >= 1 && G <= g_max, "mn_supervised_fast_sets: gene set %d has %d genes (g_max = %d)", j, G, g_max);
        // every set uses its own row pitches (the kernels' padding rules), inside the buffers sized for g_max
        const int64_t ldk = ((int64_t)G + 127) / 128 * 128, ldd = ((int64_t)G + 63) / 64 * 64;
        int rc = mn_rank_normalize_limbs(X, ldx, perm, all_cols + c0, N, G, a1, a0, ldk, inv64, inv, flags, 1, stream);
        if (rc) return rc;
        MN_CUDA(cudaMemsetAsync(cq, 0, (size_t)SK * ldd * 8, st));
        MN_CUDA(cudaMemsetAsync(nq, 0, (size_t)SK * 4, st));
        rc = mn_centroids_fixed_i8(a1, a0, ldk, inv64, lab_row_off, SK, G, max_label_rows, cq, ldd, nq, stream);
        if (rc) return rc;
        rc = mn_centroids_finish(cq, nq, SK, ldd, C, nv, stream);
        if (rc) return rc;
        MN_CUDA(cudaMemsetAsync(limbs, 0, (size_t)(4 * L.rows_pad * ldk), st));
        rc = mn_limb_split_centroids(C, ldd, SK, G, limbs, ldk, L.rows_pad, scale, nullptr, stream);
        if (rc) return rc;
        rc = mn_votes_fast_i8(a1, a0, ldk, inv64, N, limbs, ldk, (int32_t)L.rows_pad, scale, nullptr, SK, G, votes, nullptr,
                              N, nullptr, 0, nullptr, nullptr, stream);
        if (rc) return rc;
        rc = mn_loso_finalize(votes, N, 0, nullptr, nv, cell_study, N, S, K, ndn, vmat, N, stream);
        if (rc) return rc;
        rc = mn_mask_dropped(cell_label, inv, flags, N, lab, poison + j, stream);
        if (rc) return rc;
        rc = mn_auroc(vmat, N, K, nullptr, nullptr, seg_off, S, lab, SK, N, tables + (int64_t)j * SK * K, nullptr, aws,
                      L.ws_bytes, stream);
        if (rc) return rc;
    }
    return 0;
}
