/*
 * simspread_b200.h -- C ABI of libsimspread_b200.so
 *
 * B200 (sm_100a) implementation of the resource-spreading hot path of SimSpread.jl:
 *   featurize/cutoff -> construct (+ degrees k) -> spread -> predict -> clean! -> ranking metrics.
 *
 * The reference (cvigilv/SimSpread.jl) is pure Julia and has no FFI of its own; the boundary this
 * header replaces is the set of exported Julia generics (src/SimSpread.jl:21-56).  Every entry
 * point below names the reference function (file:line, relative to the reference checkout) whose
 * array work it performs.  The Julia wrappers (`simspread.jl_b200/julia/`) and the Python ctypes
 * mirror (`simspread.jl_b200/host.py`) bind exactly these symbols; see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns an int32 status (SS_OK == 0); the message of the last failure on the
 *     calling thread is available from ss_last_error().  No C++ exception crosses the ABI.
 *   - host buffers are caller-owned, COLUMN-MAJOR float64 with an explicit leading dimension
 *     (Julia `Matrix{Float64}` layout), valid for the duration of the call only.
 *   - device memory is library-owned behind opaque handles (ss_mat, ss_ivec, ss_csr) unless it was
 *     wrapped with ss_mat_wrap()/ss_ivec_wrap(); handles are destroyed explicitly.
 *   - indices inside the ABI are 0-based int32 (the Julia wrapper converts from 1-based).
 *   - one ss_ctx per GPU; calls on a context are issued on its stream and are synchronous on
 *     return unless the name ends in _async.  There is NO CPU fallback: without a CUDA device
 *     ss_ctx_create() fails with SS_ERR_NO_DEVICE.
 */
#ifndef SIMSPREAD_B200_H
#define SIMSPREAD_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define SS_API __attribute__((visibility("default")))
#else
#define SS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SS_VERSION 100 /* 0.1.0 */

/* status codes */
#define SS_OK 0
#define SS_ERR_INVALID 1     /* bad argument (shape, null pointer, alignment) */
#define SS_ERR_CUDA 2        /* a CUDA runtime/driver call failed */
#define SS_ERR_NO_DEVICE 3   /* no usable sm_100 device: there is no CPU fallback */
#define SS_ERR_ASSERT 4      /* a reference @assert fired; ss_last_error() holds its message */
#define SS_ERR_OOM 5         /* device or pinned-host allocation failed */
#define SS_ERR_UNSUPPORTED 6 /* valid request that this build does not implement */

/* ss_predict_* flags */
#define SS_PREDICT_CLEAN 1u /* fuse clean! (src/core.jl:478-484): columns with kt == 0 -> -99 */

/* precision of the two chain products, OR-ed into the ss_predict_* flags (default FP64 on the DMMA
 * pipe).  SS_PRECISION_TF32 is the opt-in fast mode, the analogue of the reference's reduced
 * precision GPU=true path (Float32 cuBLAS SGEMM, src/core.jl:404): operands rounded to TF32,
 * tcgen05.mma kind::tf32, FP32 accumulation in TMEM; measured max relative error 2e-4 at K = 20000. */
#define SS_PRECISION_F64 0u
#define SS_PRECISION_TF32 (1u << 4)
/* FP64-grade results from exact INT8 tensor-core products (Ozaki-style 8-bit slicing of non-negative
 * operands, INT32 accumulation in TMEM, FP64 recombination; csrc/ss_umma.cu).  The a-priori error bound is
 * normwise (K * (S + 1.5) * 2^(-8S) * rowmax * colmax), so every entry is CERTIFIED a posteriori in the epilogue
 * (bound <= 4e-13 * entry, or the entry is an exact 0); a product with an uncertified entry is re-run on the FP64
 * DMMA path (ss_ctx_int8_stats).  SS_INT8_SLICES=4..8 changes the slice count (default 6), SS_INT8_TOL the
 * certificate, SS_INT8_CERTIFY=0 disables it. */
#define SS_PRECISION_F64_INT8 (3u << 4)
#define SS_PRECISION_MASK (15u << 4)

/* ss_gemm_f64 operand form of A */
#define SS_OP_N 0 /* A is M x K column-major (m contiguous)            : C = A  * B */
#define SS_OP_T 1 /* A is stored K x M column-major (k contiguous)     : C = A' * B */

typedef struct ss_ctx ss_ctx;   /* one GPU + one stream + workspaces */
typedef struct ss_mat ss_mat;   /* dense float64 column-major device matrix */
typedef struct ss_ivec ss_ivec; /* int32 device vector (degrees, index lists) */
typedef struct ss_csr ss_csr;   /* CSR device matrix (int32 row_ptr/col_idx, optional f64 values) */
typedef struct ss_transfer ss_transfer; /* item x item block of W*W of a 2-layer graph, split by column tile */
typedef struct ss_comm ss_comm;       /* one rank of a single-node NCCL communicator (one process per GPU) */
typedef struct ss_sharded ss_sharded; /* state of the sharded predict: T (IPC-shared), degrees, peer mappings */

/* ---- library / context ------------------------------------------------------------------- */
SS_API int32_t ss_version(void);
SS_API const char* ss_last_error(void);
SS_API int32_t ss_device_count(int32_t* count);
SS_API int32_t ss_ctx_create(int32_t device, ss_ctx** out);
SS_API int32_t ss_ctx_destroy(ss_ctx* ctx);
SS_API int32_t ss_ctx_sync(ss_ctx* ctx);
/* the CUDA stream the context launches on (a cudaStream_t), for event timing by the caller */
SS_API int32_t ss_ctx_stream(ss_ctx* ctx, void** stream_out);
/* number of kernels launched on this context since creation (bench.py's gpu_launches) */
SS_API int32_t ss_ctx_launch_count(ss_ctx* ctx, int64_t* count);
/* SS_PRECISION_F64_INT8 bookkeeping: stats4 = {products computed on the INT8 tensor pipe, products whose a-posteriori
 * certificate failed and that were re-run on the FP64 DMMA path, entries that failed in the last product, slice pairs
 * of the last product (planes that hold only zeros are skipped: a 0/1 matrix is one plane)}. */
SS_API int32_t ss_ctx_int8_stats(ss_ctx* ctx, int64_t* stats4);
/* Per-kernel device timing of the chain-product GEMMs (CUDA events on the context stream, recorded
 * around each launch while enabled).  ss_ctx_profile_read() synchronises, returns up to `cap`
 * (milliseconds, algorithmic flops = 2*M*N*K) pairs in launch order and clears the list. */
SS_API int32_t ss_ctx_profile(ss_ctx* ctx, int32_t enable);
SS_API int32_t ss_ctx_profile_read(ss_ctx* ctx, double* ms_out, double* flops_out, int32_t cap, int32_t* n_out);
SS_API int32_t ss_host_alloc(int64_t bytes, void** out); /* pinned host memory */
SS_API int32_t ss_host_free(void* p);

/* ---- device containers ------------------------------------------------------------------- */
/* zero-initialised rows x cols matrix; ld is padded to a multiple of 16 elements (TMA alignment) */
SS_API int32_t ss_mat_create(ss_ctx* ctx, int64_t rows, int64_t cols, ss_mat** out);
/* same, allocated with cudaMalloc so that ss_mat_ipc_handle() can export it to peer processes */
SS_API int32_t ss_mat_create_ipc(ss_ctx* ctx, int64_t rows, int64_t cols, ss_mat** out);
/* non-owning view of caller-managed device memory (ld >= rows).  GEMM inputs additionally need a
 * 16-byte aligned base and an even ld (TMA); a GEMM output may be a row-offset view (8-byte aligned). */
SS_API int32_t ss_mat_wrap(ss_ctx* ctx, void* devptr, int64_t rows, int64_t cols, int64_t ld, ss_mat** out);
SS_API int32_t ss_mat_destroy(ss_mat* m);
SS_API int32_t ss_mat_info(const ss_mat* m, int64_t* rows, int64_t* cols, int64_t* ld, void** devptr);
SS_API int32_t ss_mat_upload(ss_ctx* ctx, ss_mat* m, const double* host, int64_t ld_host);
/* row-major host array (row pitch ld_host >= cols, NumPy's default order): uploaded as it lies, transposed on the device */
SS_API int32_t ss_mat_upload_rowmajor(ss_ctx* ctx, ss_mat* m, const double* host, int64_t ld_host);
SS_API int32_t ss_mat_download(ss_ctx* ctx, const ss_mat* m, double* host, int64_t ld_host);
/* column range [col0, col0+ncols) only; asynchronous on the context stream (pinned host memory) */
SS_API int32_t ss_mat_upload_cols_async(ss_ctx* ctx, ss_mat* m, int64_t col0, int64_t ncols,
                                 const double* host, int64_t ld_host);
SS_API int32_t ss_mat_download_cols_async(ss_ctx* ctx, const ss_mat* m, int64_t col0, int64_t ncols,
                                   double* host, int64_t ld_host);
SS_API int32_t ss_ivec_create(ss_ctx* ctx, int64_t n, ss_ivec** out);
SS_API int32_t ss_ivec_wrap(ss_ctx* ctx, void* devptr, int64_t n, ss_ivec** out);
SS_API int32_t ss_ivec_destroy(ss_ivec* v);
SS_API int32_t ss_ivec_info(const ss_ivec* v, int64_t* n, void** devptr);
SS_API int32_t ss_ivec_upload(ss_ctx* ctx, ss_ivec* v, const int32_t* host);
SS_API int32_t ss_ivec_download(ss_ctx* ctx, const ss_ivec* v, int32_t* host);

/* ---- host I/O: read_namedmatrix fast path ---------------------------------------------------------- */
/* read_namedmatrix [src/utils.jl:50-53: readdlm(filepath, delimiter, String); :30-32 parse.(Float64, block)]
 * for large matrices: all host cores parse the value block of a delimited text file into a caller-owned
 * column-major float64 buffer.  Every occurrence of `delimiter` separates two fields (readdlm with an explicit
 * delimiter), '\n' (optionally preceded by '\r') ends a line.  The host layer reads the names (first line /
 * first field of each line) itself and applies the reference's sort by name (:38). */
SS_API int32_t ss_text_matrix_dims(const char* path, int32_t delimiter, int64_t* lines_out, int64_t* fields_out);
SS_API int32_t ss_text_matrix_read(const char* path, int32_t delimiter, int32_t skip_lines, int32_t skip_fields,
                                   double* values, int64_t rows, int64_t cols, int64_t ld);
/* `save(filepath, yhat, y; delimiter)` / `save(filepath, fidx, yhat, y; delimiter)` [src/core.jl:503-522, 542-561]:
 * one line `fold <d> "query" <d> "target" <d> score <d> label` per pair, numbers as Julia's string(x) (shortest
 * round-trip digits; an integer-typed matrix prints integers), byte-exact against test/data/save1..4.  fold < 0: the
 * 1-based index of the query [src/core.jl:512].  yhat / y: host, column-major nq x nt in the row order of y; formatted
 * by all host cores.  ss_save_rows_mat: the same from device-resident blocks (downloaded through pinned memory). */
SS_API int32_t ss_save_rows(const char* path, int32_t append, int64_t fold, int64_t nq, int64_t nt, const char* const* qnames,
                            const char* const* tnames, const double* yhat, int64_t ld_yhat, int32_t yhat_is_int,
                            const double* y, int64_t ld_y, int32_t y_is_int, int32_t delimiter, int64_t* bytes_written);
SS_API int32_t ss_save_rows_mat(ss_ctx* ctx, const char* path, int32_t append, int64_t fold, const char* const* qnames,
                                const char* const* tnames, const ss_mat* yhat, const ss_mat* y, int32_t y_is_int,
                                int32_t delimiter, int64_t* bytes_written);

/* ---- (1) featurization ------------------------------------------------------------------- */
/* cutoff.(S, alpha, weighted)  [src/core.jl:37-43 scalar rule, :55-60 array, :106-112 featurize,
 * :129-132 featurize!]: X[i,j] = S[i,j] >= alpha ? (weighted ? S[i,j] : 1.0) : 0.0 (NaN -> 0).
 * X may be S itself (featurize!).  One pass, 128-bit loads/stores. */
SS_API int32_t ss_featurize(ss_ctx* ctx, const ss_mat* S, double alpha, int32_t weighted, ss_mat* X);
/* Same threshold, output compacted to CSR of S (row = source, ascending column index) by a
 * warp-ballot kernel; values are stored when weighted != 0.  An entry is kept iff the
 * thresholded value is an edge for src/graphs.jl:10 (non-zero). */
SS_API int32_t ss_featurize_csr(ss_ctx* ctx, const ss_mat* S, double alpha, int32_t weighted, ss_csr** out);
/* Same, but the CSR of S' (one CSR row per COLUMN of S, ascending row index) -- the natural
 * direction for a column-major S and the form the sparse chain needs for Xs (features x sources). */
SS_API int32_t ss_featurize_csc(ss_ctx* ctx, const ss_mat* S, double alpha, int32_t weighted, ss_csr** out);
/* Upstream similarity fused with the threshold: the tutorial's
 * `S = 1 .- pairwise(Jaccard(), X, dims=1)` (docs/src/tutorial/fishers-flowers.jl:66, Distances.jl) followed by
 * `featurize(S[rows, cols], alpha, weighted)` (src/core.jl:106-112) in one kernel; S is never materialised.
 * DA: na x d and DB: nb x d descriptor matrices (entities in rows); X: na x nb,
 * X[i,j] = cutoff(1 - (1 - a1/a2), alpha, weighted) with Distances.jl's accumulation a1 = sum_k |a+b| - |a-b|,
 * a2 = sum_k |a+b| + |a-b| (k ascending); 0/0 counts as distance 0.  Bit-exact against the shipped iris.simmat. */
SS_API int32_t ss_jaccard_featurize(ss_ctx* ctx, const ss_mat* DA, const ss_mat* DB, double alpha, int32_t weighted,
                                    ss_mat* X);
/* Same for bit-packed fingerprints (the Jaccard index of 0/1 descriptors = Tanimoto coefficient):
 * fa_dev / fb_dev are device arrays of na x words / nb x words uint64, row-major, one fingerprint per row;
 * X[i,j] = cutoff(|a & b| / |a | b|, alpha, weighted). */
SS_API int32_t ss_tanimoto_featurize_bits(ss_ctx* ctx, const void* fa_dev, int64_t na, const void* fb_dev, int64_t nb,
                                          int64_t words, double alpha, int32_t weighted, ss_mat* X);
SS_API int32_t ss_csr_info(const ss_csr* c, int64_t* rows, int64_t* cols, int64_t* nnz, int32_t* has_values);
SS_API int32_t ss_csr_download(ss_ctx* ctx, const ss_csr* c, int32_t* row_ptr, int32_t* col_idx, double* values);
SS_API int32_t ss_csr_destroy(ss_csr* c);
/* non-owning CSR over caller-managed device arrays (int32 row_ptr[rows+1], col_idx[nnz] ascending
 * within a row, optional float64 values[nnz]; values == NULL means every stored entry is 1.0) */
SS_API int32_t ss_csr_wrap(ss_ctx* ctx, int64_t rows, int64_t cols, int64_t nnz, void* row_ptr_dev, void* col_idx_dev,
                           void* values_dev, ss_csr** out);

/* ---- (2) graph construction + degrees ------------------------------------------------------ */
/* Block extraction of construct() [src/core.jl:167,171-172: X[queries,features], X[sources,features],
 * y[sources,targets]]: dst[i,j] = src[row_idx[i], col_idx[j]]; a NULL index list means identity. */
SS_API int32_t ss_gather(ss_ctx* ctx, const ss_mat* src, const ss_ivec* row_idx, const ss_ivec* col_idx, ss_mat* dst);
/* Degrees of the masked graph B [src/graphs.jl:9-11 applied to the B of src/core.jl:196-198]:
 * ks[s] = nnz(Xs[s,:]) + nnz(Y[s,:]), kf[f] = nnz(Xs[:,f]), kt[t] = nnz(Y[:,t]).
 * Non-zero test is !iszero (NaN counts, -0.0 does not).  Xs may be NULL (2-layer NBI); any output
 * may be NULL. */
SS_API int32_t ss_degrees(ss_ctx* ctx, const ss_mat* Xs, const ss_mat* Y, ss_ivec* ks, ss_ivec* kf, ss_ivec* kt);
/* k(G) = mapslices(k, G; dims=2) [src/graphs.jl:11]: row non-zero counts of any dense matrix. */
SS_API int32_t ss_k_rows(ss_ctx* ctx, const ss_mat* G, ss_ivec* k);

/* ---- (3) spread + predict ------------------------------------------------------------------ */
/* spread [src/core.jl:365-371]: W[i,j] = G[i,j] / k[i] (true division), Inf -> 0, NaN -> 0.
 * k == NULL computes k(G) first (the literal reference call); W may alias G. */
SS_API int32_t ss_spread_rows(ss_ctx* ctx, const ss_mat* G, const ss_ivec* k, ss_mat* W);
/* One step of the chain product: C = op(A) * B on the FP64 tensor pipe (TMA-staged, DMMA), with the
 * degree normalisation fused in the epilogue:
 *   row_div  != NULL : C[m,n] = acc / row_div[m]   (0 when row_div[m] == 0)   -- the 1/kf of W[f,:]
 *   col_flag != NULL : C[m,n] = -99 where col_flag[n] == 0                     -- clean!
 * B is K x N column-major. */
SS_API int32_t ss_gemm_f64(ss_ctx* ctx, int32_t opA, const ss_mat* A, const ss_mat* B, ss_mat* C,
                    const ss_ivec* row_div, const ss_ivec* col_flag);
/* Same contract as ss_gemm_f64 (FP64 operands in, FP64 C out) with the product computed by
 * tcgen05.mma (accumulators in TMEM): precision = SS_PRECISION_TF32 or SS_PRECISION_F64_INT8. */
SS_API int32_t ss_gemm_lowp(ss_ctx* ctx, int32_t opA, const ss_mat* A, const ss_mat* B, ss_mat* C,
                            const ss_ivec* row_div, const ss_ivec* col_flag, uint32_t precision);
/* Fused GEMM + all-gather for the multi-GPU chain: as ss_gemm_f64, and every element of C is also
 * stored, from the epilogue, to `n_mirrors` (<= 7) peer-GPU matrices with the same leading dimension
 * (`mirrors[i]` = device address, mapped into this process, of the peer's element C(0,0)) over
 * NVLink P2P.  The caller synchronises the ranks afterwards (stream sync + barrier). */
SS_API int32_t ss_gemm_f64_mirrored(ss_ctx* ctx, int32_t opA, const ss_mat* A, const ss_mat* B, ss_mat* C,
                                    const ss_ivec* row_div, const ss_ivec* col_flag, int32_t n_mirrors,
                                    void* const* mirrors);
/* CUDA IPC plumbing for the mirrors: 64-byte handle of a library-owned matrix, and mapping of a
 * peer's handle into this process (peer access is enabled lazily). */
SS_API int32_t ss_mat_ipc_handle(ss_ctx* ctx, const ss_mat* m, void* handle64_out);
SS_API int32_t ss_ipc_open(ss_ctx* ctx, const void* handle64, void** devptr_out);
SS_API int32_t ss_ipc_close(ss_ctx* ctx, void* devptr);
/* predict((A,B), ytest) for query rows [src/core.jl:402-423 with the blocks of :165-198]:
 *   R = Xq * T,  T = (Xs' ./ kf) * (Y ./ ks)      
 * Xq: Nq x Nf, Xs: Ns x Nf, Y: Ns x Nt, R: Nq x Nt.  Runs degrees -> spread -> T -> R on the
 * context stream.  kt_out (optional) receives the target degrees used by clean!.
 * T is a dense FP64 tensor-core product, or -- for large problems whose label matrix Y is at most 10 % dense and whose
 * feature weights are finite -- a sum over the edges of Y in ascending source order (same value to ~1e-15, independent
 * of tiles and shards; environment SS_T_FORM=dense|sparse|auto overrides the choice). */
SS_API int32_t ss_predict_query(ss_ctx* ctx, const ss_mat* Xq, const ss_mat* Xs, const ss_mat* Y, ss_mat* R,
                         uint32_t flags, ss_ivec* kt_out);
/* The same chain with the result delivered to host memory [the `F[names(ytest,1), names(ytest,2)]` that
 * predict returns, src/core.jl:423]: the second product runs in column blocks and each finished block is copied to
 * `host` (column-major, leading dimension ld_host; staged by several host threads when the memory is pageable) while
 * the next block is computed.  Bit-identical to ss_predict_query + ss_mat_download; R keeps the device copy. */
SS_API int32_t ss_predict_query_fetch(ss_ctx* ctx, const ss_mat* Xq, const ss_mat* Xs, const ss_mat* Y, ss_mat* R,
                                      uint32_t flags, double* host, int64_t ld_host);
/* k-fold cross-validation in one call (the user-side loop of docs/src/api.md:17-21 over construct / predict /
 * clean!).  X: featurized N x N similarity matrix, Y: N' x Nt labels (both resident).  For fold f the
 * queries are rows q_idx[q_ptr[f] .. q_ptr[f+1]) of X, the sources rows s_idx[s_ptr[f] ..) of X with their label rows
 * ys_idx[s_ptr[f] ..) of Y, the features columns f_idx[f_ptr[f] ..) of X (host int32 arrays, 0-based: exactly the
 * name filtering of src/core.jl:152-154 done by the host layer).  Rows [q_ptr[f] - q_ptr[0], ...) of R receive the
 * predictions of fold f.  All folds are queued on the stream, one synchronisation at the end. */
SS_API int32_t ss_predict_query_folds(ss_ctx* ctx, const ss_mat* X, const ss_mat* Y, int32_t nfolds, const int32_t* q_ptr,
                                      const int32_t* q_idx, const int32_t* s_ptr, const int32_t* s_idx,
                                      const int32_t* ys_idx, const int32_t* f_ptr, const int32_t* f_idx, ss_mat* R,
                                      uint32_t flags);
/* Sparse form of ss_predict_query for high-alpha (few-percent dense) feature blocks: Xq is the CSR
 * of the Nq x Nf query block, XsT the CSR of Xs' (Nf x Ns, from ss_featurize_csc), Y the dense
 * Ns x Nt label block.  Both products run as row-split SpMMs; R is dense column-major as before. */
SS_API int32_t ss_predict_query_csr(ss_ctx* ctx, const ss_csr* Xq, const ss_csr* XsT, const ss_mat* Y, ss_mat* R,
                                    uint32_t flags, ss_ivec* kt_out);
/* predict(A, ytrain) / source rows [src/core.jl:446-466]: R = Xs*T + Y*U, U = (Y' ./ kt)*(Y ./ ks).
 * Xs may be NULL (classical 2-layer NBI: R = Y*U). */
SS_API int32_t ss_predict_source(ss_ctx* ctx, const ss_mat* Xs, const ss_mat* Y, ss_mat* R, uint32_t flags);
/* Sparse 2-layer NBI with fused top-L (BASELINE config 5; predict(A, ytrain) of src/core.jl:446-466
 * on the graph [0 Y; Y' 0], reduced to `sortperm(rev=true)[1:L]` per source, src/performance.jl:315):
 * Y = CSR of the source x target graph, YT = CSR of its transpose.  F is never materialised.  Sources
 * [s_begin, s_end) are processed (shard the range across GPUs); idx_out is L x sources (0-based target
 * indices, -1 padding), val_out (optional) the matching scores, an L x sources matrix.  L <= 32. */
SS_API int32_t ss_recommend_topl(ss_ctx* ctx, const ss_csr* Y, const ss_csr* YT, int32_t L, int64_t s_begin,
                                 int64_t s_end, ss_ivec* idx_out, ss_mat* val_out);
/* The same computation in two steps, for callers that rank several source ranges of one graph (multi-GPU
 * sharding, repeated calls): ss_transfer_build materialises U = (W*W)[targets, targets] =
 * (Y' ./ kt) * (Y ./ ks) of the reference's `Aarr * Warr^2` [src/core.jl:456] once -- rows sorted by column,
 * split by column tile, every entry summed in ascending source order -- and ss_recommend_topl_transfer streams
 * it: F[s,:] = sum_{t' in Y[s,:], ascending} Y[s,t'] * U[t',:] accumulated in shared memory (no atomics: scores
 * and the order of tied scores are reproducible bit for bit), reduced to the top L.  ss_recommend_topl does both
 * and falls back to column-tile chunks when U does not fit in the free device memory.
 * ss_transfer_info: info4 = {entries of U, device bytes, tile width, column tiles}.
 * ss_transfer_download (tests): U as a host CSR (row_ptr: targets + 1 int64, global int32 columns, values). */
SS_API int32_t ss_transfer_build(ss_ctx* ctx, const ss_csr* Y, const ss_csr* YT, ss_transfer** out);
SS_API int32_t ss_transfer_info(const ss_transfer* U, int64_t* info4);
SS_API int32_t ss_transfer_download(const ss_transfer* U, int64_t* row_ptr_host, int32_t* col_host, double* val_host);
SS_API int32_t ss_transfer_destroy(ss_transfer* U);
SS_API int32_t ss_recommend_topl_transfer(ss_ctx* ctx, const ss_csr* Y, const ss_transfer* U, int32_t L, int64_t s_begin,
                                          int64_t s_end, ss_ivec* idx_out, ss_mat* val_out);
/* clean! [src/core.jl:478-484]: R[:,t] = -99 for every t with kt[t] == 0. */
SS_API int32_t ss_clean(ss_ctx* ctx, ss_mat* R, const ss_ivec* kt);

/* Reference-facing one-call form with HOST buffers (what the Julia `predict` wrapper ccalls):
 * uploads Xq/Xs/Y, runs ss_predict_query, downloads R.  Xq slabs / R slabs are pipelined against
 * the R GEMM on separate streams when the host buffers are pinned. */
SS_API int32_t ss_predict_query_host(ss_ctx* ctx, const double* Xq, int64_t ldxq, const double* Xs, int64_t ldxs,
                              const double* Y, int64_t ldy, int64_t nq, int64_t ns, int64_t nf, int64_t nt,
                              uint32_t flags, double* R, int64_t ldr);

/* Second half of the same call for callers that already hold T on the device (multi-GPU: T comes
 * from the sharded chain): R_host (nq x N) = Xq_host (nq x K) * T (K x N) with the clean! flag of
 * col_flag (optional), query-row slabs streamed H2D / GEMM / D2H on three streams. */
SS_API int32_t ss_stream_product_host(ss_ctx* ctx, const double* Xq, int64_t ldxq, int64_t nq, const ss_mat* T,
                                      const ss_ivec* col_flag, double* R, int64_t ldr);

/* ---- (3b) multi-GPU: the exchange steps of the sharded predict (SURVEY 8b `ss_comm_init`, 8e) ------------------
 * One process per GPU of one node.  NCCL (dlopen of libnccl.so.2; override with SS_NCCL_LIBRARY) carries the
 * all-reduce of the source degrees and the all-gather of the target degrees; the all-gather of the T tiles is fused
 * into the T-GEMM epilogue over CUDA-IPC peer mappings (NVLink P2P stores), with an NCCL all-gather as the fallback
 * (SS_FUSED_ALLGATHER=0 or no peer access).  The reference has no multi-GPU path: this is the north star's
 * "sharded by query rows ... NCCL only to all-gather the target-side degree vector and W tiles".
 * ss_comm_unique_id: 128 bytes created by ONE rank and passed to every rank through the host's own channel;
 * ss_comm_init_file: the same through a file (rank 0 writes `path`, the others poll up to timeout_s; `path` must be
 * unique to the job).  All calls are collective over the ranks of the communicator and synchronous on return. */
SS_API int32_t ss_comm_unique_id(void* id128_out);
SS_API int32_t ss_comm_init(ss_ctx* ctx, int32_t rank, int32_t world, const void* id128, ss_comm** out);
SS_API int32_t ss_comm_init_file(ss_ctx* ctx, int32_t rank, int32_t world, const char* path, double timeout_s, ss_comm** out);
SS_API int32_t ss_comm_destroy(ss_comm* comm);
SS_API int32_t ss_comm_info(const ss_comm* comm, int32_t* rank, int32_t* world, int32_t* nccl_version);
SS_API int32_t ss_comm_barrier(ss_comm* comm);
SS_API int32_t ss_comm_allreduce_i32(ss_comm* comm, ss_ivec* v);                          /* in place, sum */
SS_API int32_t ss_comm_allgather_i32(ss_comm* comm, const ss_ivec* send, ss_ivec* recv);  /* recv: world x send */
SS_API int32_t ss_comm_allgather_host(ss_comm* comm, const void* send, void* recv, int64_t bytes); /* small host blobs */
/* a replicated operand from sharded uploads: rank r filled column block r of M (cols divisible by world) */
SS_API int32_t ss_comm_allgather_cols(ss_comm* comm, ss_mat* M);
SS_API int32_t ss_comm_allreduce_host_f64(ss_comm* comm, double* x, int32_t n, int32_t op); /* op 0 = sum, 1 = max */
/* Sharded predict for query rows: rank r owns Xq[r] / R[r] and the target-column block Y[:, r] (ns x nt_blk,
 * nt_blk = ceil(nt / world), columns beyond nt zero); Xs (ns x nf) is replicated.  ss_sharded_create allocates T
 * (nf x nt_blk*world per rank) and maps the peers' copies; ss_sharded_front = degrees -> all-reduce ks, all-gather kt
 * -> Wst = Yblk ./ ks -> T[:, r] = (Xs' * Wst) ./ kf stored into every rank's T -> barrier; ss_predict_query_sharded
 * = front + R slab = Xq slab * T with clean! fused (Xq / R may be NULL on a rank without query rows);
 * ss_sharded_views: non-owning views of T (nf x nt) and kt (nt) for callers that stream their query rows
 * (ss_stream_product_host). */
SS_API int32_t ss_sharded_create(ss_comm* comm, int64_t ns, int64_t nf, int64_t nt, ss_sharded** out);
SS_API int32_t ss_sharded_destroy(ss_sharded* plan);
SS_API int32_t ss_sharded_info(const ss_sharded* plan, int64_t* nt_blk, int32_t* fused_allgather);
SS_API int32_t ss_sharded_front(ss_sharded* plan, const ss_mat* Xs, const ss_mat* Yblk);
SS_API int32_t ss_sharded_views(ss_sharded* plan, ss_mat** T, ss_ivec** kt);
SS_API int32_t ss_predict_query_sharded(ss_sharded* plan, const ss_mat* Xq, const ss_mat* Xs, const ss_mat* Yblk, ss_mat* R,
                                        uint32_t flags);

/* ---- (4) ranking + metrics ----------------------------------------------------------------- */
/* Per-row top-L of R under `sortperm(row; rev=true)` order [src/performance.jl:315,377]:
 * descending by isless, ties by ascending column.  idx_out: L x rows int32 (column-major, 0-based
 * columns), val_out (optional): L x rows. */
SS_API int32_t ss_topl_rows(ss_ctx* ctx, const ss_mat* R, int32_t L, ss_ivec* idx_out, ss_mat* val_out);
/* mean recall@L / precision@L over the rows (groups) of R [src/performance.jl:341-357, 398-414]:
 * out[0] = mean recall@L (NaN if any row of Ytrue has no positive, as in the reference),
 * out[1] = mean precision@L.  Requires cols > L (the reference's strict assert, :311-312). */
SS_API int32_t ss_atl(ss_ctx* ctx, const ss_mat* Ytrue, const ss_mat* R, int32_t L, double* out2);
/* AuROC / AuPRC over M (label, score) pairs [src/performance.jl:49-63, 74-89; MLBase.roc,
 * Trapz.trapz semantics]: device radix sort + scan.  labels: uint8 (non-zero = positive),
 * scores: float64, both device pointers.  out[0] = AuROC, out[1] = AuPRC. */
SS_API int32_t ss_auroc_auprc(ss_ctx* ctx, const void* labels_u8_dev, const void* scores_f64_dev, int64_t M,
                       double* out2);
/* AuROC / AuPRC of a list spread over several GPUs (one ss_ctx per rank; the exchange itself is done by the host
 * layer with NCCL, simspread.jl_b200/sharded.py).  Keys are the order-preserving uint64 image of a score under
 * Julia's isless.  Returned device pointers live in the context's sort buffers until the next metric call.
 *   ss_auc_sort            (labels, scores) or (labels, keys_in != NULL) -> ascending (key, label) arrays
 *   ss_auc_lower_bound     pos[q] = #keys < query[q]  (splitter positions; host arrays in / out)
 *   ss_auc_segment_summary {positives, index of the last run start or -1, positives before it} of a sorted array
 *   ss_auc_segment_integrate  signed trapezoid sums (ROC, PR) of the thresholds inside this key range, including
 *       the joint to the range below; global6 = {P, M_total, pairs below, positives below, global index of the last
 *       run start below (-1: none), positives before it}.  |sum over ranks| = AuROC, AuPRC of the whole list. */
SS_API int32_t ss_auc_sort(ss_ctx* ctx, const void* labels_u8_dev, const void* scores_f64_dev, const void* keys_u64_dev,
                           int64_t M, void** keys_sorted_out, void** labels_sorted_out);
SS_API int32_t ss_auc_lower_bound(ss_ctx* ctx, const void* keys_sorted_dev, int64_t M, const uint64_t* query, int32_t nq,
                                  int64_t* pos_out);
SS_API int32_t ss_auc_segment_summary(ss_ctx* ctx, const void* keys_sorted_dev, const void* labels_sorted_dev, int64_t M,
                                      int64_t* summary3);
SS_API int32_t ss_auc_segment_integrate(ss_ctx* ctx, const void* keys_sorted_dev, const void* labels_sorted_dev, int64_t M,
                                        const int64_t* global6, double* out2);
/* Same for the entries of two device matrices (Ytrue != 0 is the label), column-major order. */
SS_API int32_t ss_auroc_auprc_mat(ss_ctx* ctx, const ss_mat* Ytrue, const ss_mat* R, double* out2);

/* BEDROC [src/performance.jl:22-38] over the entries of two device matrices (label = Ytrue == 1):
 * stable descending radix sort (ascending when rev == 0), sum of exp(-alpha*rank/N) over the
 * positives, closed-form normalisation. */
SS_API int32_t ss_bedroc(ss_ctx* ctx, const ss_mat* Ytrue, const ss_mat* R, int32_t rev, double alpha, double* out);
/* maxperformance / meanperformance / meanstdperformance [src/performance.jl:425-531] of one
 * confusion-matrix metric over all unique-score thresholds (MLBase.roc semantics):
 * metric: 0 f1score, 1 mcc, 2 accuracy, 3 balancedaccuracy, 4 recall, 5 precision.
 * out[0] = maximum, out[1] = mean, out[2] = corrected sample std, out[3] = number of thresholds. */
SS_API int32_t ss_threshold_sweep(ss_ctx* ctx, const ss_mat* Ytrue, const ss_mat* R, int32_t metric, double* out4);

#ifdef __cplusplus
}
#endif
#endif /* SIMSPREAD_B200_H */
