// Upstream similarity fused with the featurization threshold.
//
// The reference leaves the similarity matrix to the user: the tutorial builds it as
// `S = 1 .- pairwise(Jaccard(), X, dims=1)` (docs/src/tutorial/fishers-flowers.jl:66, Distances.jl) and
// then calls `featurize(S[rows, cols], alpha, weighted)` (src/core.jl:106-112, rule :37-43).  Here both
// steps are one kernel: the N x N matrix S is never written, only the featurized block comes out.
//
//   jaccard_featurize_kernel : real-valued descriptors (entities x d, column-major FP64)
//        Distances.jl's accumulation, k ascending:  a1 += |a+b| - |a-b| (= 2 min),  a2 += |a+b| + |a-b| (= 2 max),
//        distance = 1 - a1/a2 (NaN, i.e. 0/0, -> 0),  S(i,j) = 1 - distance  [the tutorial's `1 .- `];
//        this order of operations reproduces the shipped docs/src/tutorial/data/iris.simmat bit for bit
//   tanimoto_bits_kernel     : bit-packed fingerprints (entities x words, row-major uint64)
//        S(i,j) = |a & b| / (|a| + |b| - |a & b|)  (the same Jaccard index on 0/1 descriptors)
// followed by X(i,j) = S >= alpha ? (weighted ? S : 1) : 0.
//
// Both are ALU-bound (d, or words, operations per output element against 8 B written): 64 x 64 output
// tile per block, descriptor slabs staged in shared memory (conflict-free: a half-warp reads 16
// consecutive entities, the other operand is a broadcast), 4 x 4 outputs per thread in registers,
// 128-byte store runs down the (column-major) output columns.
#include "ss_common.cuh"

namespace {

constexpr int ST = 64;    // output tile edge
constexpr int SK = 16;    // descriptor slab (FP64 kernel)
constexpr int SW = 8;     // word slab (bit kernel)
constexpr int STPB = 256;

__device__ __forceinline__ double sim_cutoff(double x, double alpha, bool weighted) {
    return x >= alpha ? (weighted ? x : 1.0) : 0.0;  // src/core.jl:37-43 (NaN >= alpha is false -> 0.0)
}

// A: na x d (lda), B: nb x d (ldb), X: na x nb (ldx); all column-major
__global__ void __launch_bounds__(STPB)
    jaccard_featurize_kernel(const double* __restrict__ A, int64_t na, int64_t lda, const double* __restrict__ B, int64_t nb,
                             int64_t ldb, int64_t d, double alpha, int weighted, double* __restrict__ X, int64_t ldx) {
    __shared__ double sa[SK][ST];
    __shared__ double sb[SK][ST + 1];
    const int tid = threadIdx.x;
    const int ti = tid & 15, tj = tid >> 4;  // thread owns rows i0 + ti + 16*r, columns j0 + tj + 16*c
    const int64_t i0 = int64_t(blockIdx.x) * ST, j0 = int64_t(blockIdx.y) * ST;
    double mn[4][4], mx[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) mn[r][c] = mx[r][c] = 0.0;
    for (int64_t k0 = 0; k0 < d; k0 += SK) {
        for (int e = tid; e < SK * ST; e += STPB) {  // entity index fastest: coalesced along the columns of A / B
            const int k = e / ST, x = e % ST;
            const bool kin = k0 + k < d;
            sa[k][x] = (kin && i0 + x < na) ? A[(k0 + k) * lda + i0 + x] : 0.0;
            sb[k][x] = (kin && j0 + x < nb) ? B[(k0 + k) * ldb + j0 + x] : 0.0;
        }
        __syncthreads();
        const int kk = (d - k0 < SK) ? int(d - k0) : SK;
        for (int k = 0; k < kk; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = sa[k][ti + 16 * r];
#pragma unroll
            for (int c = 0; c < 4; ++c) b[c] = sb[k][tj + 16 * c];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double pl = fabs(a[r] + b[c]), mi = fabs(a[r] - b[c]);
                    mn[r][c] += pl - mi;
                    mx[r][c] += pl + mi;
                }
        }
        __syncthreads();
    }
    const bool w = weighted != 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int64_t j = j0 + tj + 16 * c;
        if (j >= nb) continue;
        double o[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const double dq = 1.0 - mn[r][c] / mx[r][c];
            const double dist = (dq != dq) ? 0.0 : dq;  // Distances.jl: a NaN distance (0/0) is 0
            o[r] = sim_cutoff(1.0 - dist, alpha, w);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {  // a half-warp writes 16 consecutive rows of one output column (128 B)
            const int64_t i = i0 + ti + 16 * r;
            if (i < na) X[j * ldx + i] = o[r];
        }
    }
}

// |row| of bit-packed fingerprints: one warp per entity
__global__ void __launch_bounds__(256) bits_popcount_kernel(const uint64_t* __restrict__ F, int64_t n, int64_t words,
                                                            int32_t* __restrict__ cnt) {
    const int64_t row = (int64_t(blockIdx.x) * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    int c = 0;
    for (int64_t w = lane; w < words; w += 32) c += __popcll(F[row * words + w]);
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) cnt[row] = c;
}

// FA: na x words, FB: nb x words (row-major uint64); X: na x nb column-major
__global__ void __launch_bounds__(STPB)
    tanimoto_bits_kernel(const uint64_t* __restrict__ FA, const int32_t* __restrict__ ca, int64_t na,
                         const uint64_t* __restrict__ FB, const int32_t* __restrict__ cb, int64_t nb, int64_t words, double alpha,
                         int weighted, double* __restrict__ X, int64_t ldx) {
    __shared__ uint64_t sa[SW][ST];
    __shared__ uint64_t sb[SW][ST + 1];
    const int tid = threadIdx.x;
    const int ti = tid & 15, tj = tid >> 4;
    const int64_t i0 = int64_t(blockIdx.x) * ST, j0 = int64_t(blockIdx.y) * ST;
    int both[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) both[r][c] = 0;
    for (int64_t w0 = 0; w0 < words; w0 += SW) {
        for (int e = tid; e < SW * ST; e += STPB) {  // word index fastest: 64-byte runs of one fingerprint
            const int x = e / SW, k = e % SW;
            const bool kin = w0 + k < words;
            sa[k][x] = (kin && i0 + x < na) ? FA[(i0 + x) * words + w0 + k] : 0ull;
            sb[k][x] = (kin && j0 + x < nb) ? FB[(j0 + x) * words + w0 + k] : 0ull;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SW; ++k) {
            uint64_t a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = sa[k][ti + 16 * r];
#pragma unroll
            for (int c = 0; c < 4; ++c) b[c] = sb[k][tj + 16 * c];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) both[r][c] += __popcll(a[r] & b[c]);
        }
        __syncthreads();
    }
    const bool w = weighted != 0;
    int na_bits[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) na_bits[r] = (i0 + ti + 16 * r < na) ? ca[i0 + ti + 16 * r] : 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int64_t j = j0 + tj + 16 * c;
        if (j >= nb) continue;
        const int nbb = cb[j];
        double o[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int uni = na_bits[r] + nbb - both[r][c];
            const double s = uni == 0 ? 1.0 : double(both[r][c]) / double(uni);  // 0/0: distance 0
            o[r] = sim_cutoff(s, alpha, w);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {  // a half-warp writes 16 consecutive rows of one output column (128 B)
            const int64_t i = i0 + ti + 16 * r;
            if (i < na) X[j * ldx + i] = o[r];
        }
    }
}

}  // namespace

namespace ss {

int32_t jaccard_featurize(ss_ctx* ctx, const ss_mat* A, const ss_mat* B, double alpha, bool weighted, ss_mat* X) {
    const int64_t na = A->rows, nb = B->rows, d = A->cols;
    if (na == 0 || nb == 0) return SS_OK;
    const int64_t gy = ceil_div(nb, ST);
    SS_REQUIRE(gy <= 65535, "jaccard_featurize: too many columns (%lld)", (long long)nb);
    dim3 grid(unsigned(ceil_div(na, ST)), unsigned(gy));
    jaccard_featurize_kernel<<<grid, STPB, 0, ctx->stream>>>(A->d, na, A->ld, B->d, nb, B->ld, d, alpha, weighted ? 1 : 0, X->d,
                                                           X->ld);
    SS_CHECK_CUDA(cudaGetLastError());
    ctx->launches++;
    return SS_OK;
}

int32_t tanimoto_bits_featurize(ss_ctx* ctx, const uint64_t* FA, int64_t na, const uint64_t* FB, int64_t nb, int64_t words,
                                double alpha, bool weighted, ss_mat* X) {
    if (na == 0 || nb == 0) return SS_OK;
    const int64_t gy = ceil_div(nb, ST);
    SS_REQUIRE(gy <= 65535, "tanimoto_bits_featurize: too many columns (%lld)", (long long)nb);
    void* w;
    SS_TRY(scratch_get(ctx, 15, size_t(na + nb) * 4 + 256, &w));
    int32_t* ca = static_cast<int32_t*>(w);
    int32_t* cb = ca + na;
    bits_popcount_kernel<<<unsigned(ceil_div(na * 32, 256)), 256, 0, ctx->stream>>>(FA, na, words, ca);
    bits_popcount_kernel<<<unsigned(ceil_div(nb * 32, 256)), 256, 0, ctx->stream>>>(FB, nb, words, cb);
    dim3 grid(unsigned(ceil_div(na, ST)), unsigned(gy));
    tanimoto_bits_kernel<<<grid, STPB, 0, ctx->stream>>>(FA, ca, na, FB, cb, nb, words, alpha, weighted ? 1 : 0, X->d, X->ld);
    SS_CHECK_CUDA(cudaGetLastError());
    ctx->launches += 3;
    return SS_OK;
}

}  // namespace ss
