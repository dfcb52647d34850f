// Shared internals of libsimspread_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/simspread_b200.h"

namespace ss {

void set_error(const char* fmt, ...);

#define SS_CHECK_CUDA(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ss::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                          __LINE__);                                                          \
            return (_e == cudaErrorMemoryAllocation) ? SS_ERR_OOM : SS_ERR_CUDA;              \
        }                                                                                     \
    } while (0)

#define SS_REQUIRE(cond, ...)             \
    do {                                  \
        if (!(cond)) {                    \
            ss::set_error(__VA_ARGS__);   \
            return SS_ERR_INVALID;        \
        }                                 \
    } while (0)

#define SS_TRY(expr)                 \
    do {                             \
        int32_t _s = (expr);         \
        if (_s != SS_OK) return _s;  \
    } while (0)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// grow-only device scratch buffer
struct Scratch {
    void* p = nullptr;
    size_t bytes = 0;
};

}  // namespace ss

struct ss_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;   // compute
    cudaStream_t copy_in = nullptr;  // H2D pipeline
    cudaStream_t copy_out = nullptr; // D2H pipeline
    int64_t launches = 0;
    // int8-sliced products since context creation: {products, products re-run on the DMMA path, entries that
    // failed the certificate in the last product}
    int64_t int8_stats[3] = {0, 0, 0};
    int int8_last_pairs = 0;  // slice pairs of the last int8 product (planes of zeros are skipped)
    // workspaces reused across predict calls (never shrink)
    ss::Scratch ws[28];
    int32_t* tile_counter = nullptr;
    void* stage[3] = {nullptr, nullptr, nullptr};  // pinned staging buffers of the pageable-memory copies
    cudaEvent_t stage_ev[3] = {nullptr, nullptr, nullptr};
    bool gemm_attr_set = false;
    bool tsp_attr_set = false;
    // optional per-GEMM timing (ss_ctx_profile)
    bool profile = false;
    struct ProfRec {
        cudaEvent_t start, stop;
        double flops;
    };
    std::vector<ProfRec> prof;
};

struct ss_mat {
    ss_ctx* ctx = nullptr;
    double* d = nullptr;
    int64_t rows = 0, cols = 0, ld = 0;
    bool owned = false;
    bool pooled = false;  // allocated with cudaMallocAsync on the context stream
};

struct ss_ivec {
    ss_ctx* ctx = nullptr;
    int32_t* d = nullptr;
    int64_t n = 0;
    bool owned = false;
};

struct ss_csr {
    ss_ctx* ctx = nullptr;
    int64_t rows = 0, cols = 0, nnz = 0;
    int32_t* row_ptr = nullptr;  // rows + 1
    int32_t* col_idx = nullptr;  // nnz
    double* values = nullptr;    // nnz or null
    bool owned = true;           // false: wraps caller-managed device arrays (ss_csr_wrap)
};

namespace ss {

int32_t scratch_get(ss_ctx* ctx, int slot, size_t bytes, void** out);

// ---- kernels (launchers; all asynchronous on ctx->stream) --------------------------------------
int32_t launch_featurize(ss_ctx* ctx, const double* S, int64_t rows, int64_t cols, int64_t lds,
                         double alpha, bool weighted, double* X, int64_t ldx);
int32_t launch_gather(ss_ctx* ctx, const double* src, int64_t lds, const int32_t* ridx,
                      const int32_t* cidx, double* dst, int64_t rows, int64_t cols, int64_t ldd);
int32_t launch_transpose(ss_ctx* ctx, const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows,
                         int64_t cols);
int32_t launch_degrees(ss_ctx* ctx, const double* M, int64_t rows, int64_t cols, int64_t ld,
                       int32_t* row_deg, int32_t* col_deg);  // accumulates into row_deg (+=)
int32_t launch_spread_rows(ss_ctx* ctx, const double* G, int64_t rows, int64_t cols, int64_t ldg,
                           const int32_t* k, double* W, int64_t ldw);
int32_t launch_clean(ss_ctx* ctx, double* R, int64_t rows, int64_t cols, int64_t ld,
                     const int32_t* kt);
int32_t launch_gemm_f64(ss_ctx* ctx, int opA, const double* A, int64_t lda, const double* B,
                        int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                        const int32_t* row_div, const int32_t* col_flag, bool accumulate, int nmirror = 0,
                        double* const* mirrors = nullptr);
int32_t launch_gemm_tf32(ss_ctx* ctx, int opA, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                         int64_t ldc, int64_t M, int64_t N, int64_t K, const int32_t* row_div, const int32_t* col_flag,
                         bool split);
int32_t launch_gemm_i8(ss_ctx* ctx, int opA, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                       int64_t ldc, int64_t M, int64_t N, int64_t K, const int32_t* row_div, const int32_t* col_flag,
                       int S, double cert_tol, int64_t* uncertified);
// T = (Xs' * (Y ./ ks)) ./ kf from the edges of a sparse label matrix (csrc/ss_tsparse.cu); *used = false: declined,
// the caller runs spread + the dense GEMM
int32_t t_from_sparse_labels(ss_ctx* ctx, const double* Xs, int64_t ldxs, const double* Y, int64_t ldy, int64_t ns, int64_t nf,
                             int64_t nt, const int32_t* ks, const int32_t* kf, const int32_t* kt, double* T, int64_t ldt,
                             int nmirror, double* const* mirrors, bool* used);
int32_t featurize_csr(ss_ctx* ctx, const ss_mat* S, double alpha, bool weighted, ss_csr** out);
int32_t featurize_csc(ss_ctx* ctx, const ss_mat* S, double alpha, bool weighted, ss_csr** out);
int32_t predict_query_csr(ss_ctx* ctx, const ss_csr* Xq, const ss_csr* XsT, const ss_mat* Y, ss_mat* R, uint32_t flags,
                          int32_t* kt_out);
int32_t launch_topl(ss_ctx* ctx, const double* R, int64_t rows, int64_t cols, int64_t ld, int L,
                    int32_t* idx_out, double* val_out, int64_t ldv);
int32_t atl(ss_ctx* ctx, const ss_mat* Y, const ss_mat* R, int L, double* out2);
int32_t auroc_auprc(ss_ctx* ctx, const uint8_t* labels, const double* scores, int64_t M,
                    double* out2);
int32_t auroc_auprc_mat(ss_ctx* ctx, const ss_mat* Y, const ss_mat* R, double* out2);
int32_t recommend_topl(ss_ctx* ctx, const ss_csr* Y, const ss_csr* YT, int L, int64_t s_begin, int64_t s_end,
                       int32_t* idx_out, double* val_out, int64_t ldv);
int32_t recommend_topl_stream(ss_ctx* ctx, const ss_csr* Y, const ss_csr* YT, int L, int64_t s_begin, int64_t s_end,
                              int32_t* idx_out, double* val_out, int64_t ldv, bool* declined);
int32_t predict_query_folds_batched(ss_ctx* ctx, const ss_mat* X, const ss_mat* Y, int nfolds, const int32_t* q_ptr,
                                    const int32_t* s_ptr, const int32_t* f_ptr, const int32_t* dq, const int32_t* dsx,
                                    const int32_t* dsy, const int32_t* df, ss_mat* R, bool clean);
int32_t jaccard_featurize(ss_ctx* ctx, const ss_mat* A, const ss_mat* B, double alpha, bool weighted, ss_mat* X);
int32_t tanimoto_bits_featurize(ss_ctx* ctx, const uint64_t* FA, int64_t na, const uint64_t* FB, int64_t nb, int64_t words,
                                double alpha, bool weighted, ss_mat* X);
int32_t auc_sort(ss_ctx* ctx, const uint8_t* labels, const double* scores, const uint64_t* keys_in, int64_t M,
                 uint64_t** keys_out, uint8_t** labels_out);
int32_t auc_lower_bound(ss_ctx* ctx, const uint64_t* keys_sorted, int64_t M, const uint64_t* query_host, int nq,
                        int64_t* pos_host);
int32_t auc_segment_summary(ss_ctx* ctx, const uint64_t* keys, const uint8_t* lab, int64_t M, int64_t* summary3);
int32_t auc_segment_integrate(ss_ctx* ctx, const uint64_t* keys, const uint8_t* lab, int64_t M, const int64_t* global6,
                              double* out2);
int32_t threshold_sweep(ss_ctx* ctx, const ss_mat* Y, const ss_mat* R, int metric, double* out4);
int32_t bedroc(ss_ctx* ctx, const ss_mat* Y, const ss_mat* R, int rev, double alpha, double* out);

}  // namespace ss
