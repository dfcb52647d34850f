// All folds of a cross-validation in three launches (SURVEY.md 7 "batch all 10 folds", 8f-2; the loop of the
// reference's docs/src/api.md:17-21: per fold construct -> predict -> clean!).
//
// At Enzyme size (445 sources x 664 targets, 10 folds) one fold is a 400 x 400 x 664 and a 45 x 400 x 664 product:
// 24 tiles of the persistent DMMA GEMM, i.e. 24 of 148 SMs, and 3 gathers + degrees + spread + 2 GEMMs per fold in
// sequence.  Here the fold is a grid dimension and the blocks of construct() are never extracted: every operand is
// read through the fold's index lists (the name filtering of src/core.jl:152-154 done by the host layer),
//
//   folds_degrees_kernel : ks / kf / kt of every fold                      [src/graphs.jl:9-11 on the B of :196-198]
//   folds_gemm_kernel<0> : T_f = (Xs_f' * (Y_f ./ ks_f)) ./ kf_f           [W = spread(B), (W*W)[f,t]; src/core.jl:366, 413]
//   folds_gemm_kernel<1> : R_f = Xq_f * T_f, clean! fused                  [src/core.jl:413, 421, 478-484]
//
// with Xs_f[s,i] = X[sx[s], fx[i]], Y_f[s,t] = Y[sy[s], t], Xq_f[q,i] = X[qx[q], fx[i]].  The division Y ./ ks is the
// IEEE division of the reference, done at operand load.  FP64 FMA on 64 x 64 x 16 tiles, 4 x 4 outputs per thread:
// the whole CV is 2.4 GFLOP, so the tensor pipe would not be visible next to the launches it saves.
#include <algorithm>

#include "ss_common.cuh"

namespace {

struct FoldsArgs {
    const double* X;
    int64_t ldx;
    const double* Y;
    int64_t ldy;
    int64_t nt;
    int nfolds;
    const int32_t *q_ptr, *s_ptr, *f_ptr;         // device copies of the fold pointers (nfolds + 1 each)
    const int32_t *q_idx, *s_idx, *ys_idx, *f_idx; // device index lists (rebased to the first fold)
    int32_t *ks, *kf, *kt;                         // ks[s_ptr..], kf[f_ptr..], kt[fold * nt ..]
    double* T;                                     // [fold][nt][ldt]
    int64_t ldt;
    double* R;                                     // rows q_ptr[f] - q_ptr[0] .. of the result
    int64_t ldr;
    int clean;
};

// blockIdx.z: 0 = ks (one thread per source), 1 = kf (per feature), 2 = kt (per target); blockIdx.y = fold
__global__ void __launch_bounds__(256) folds_degrees_kernel(const FoldsArgs a) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int32_t s0 = a.s_ptr[f] - a.s_ptr[0], ns = a.s_ptr[f + 1] - a.s_ptr[f];
    const int32_t f0 = a.f_ptr[f] - a.f_ptr[0], nf = a.f_ptr[f + 1] - a.f_ptr[f];
    if (blockIdx.z == 0) {
        if (i >= ns) return;
        const int64_t row = a.s_idx[s0 + i], yrow = a.ys_idx[s0 + i];
        int k = 0;
        for (int j = 0; j < nf; ++j) k += a.X[int64_t(a.f_idx[f0 + j]) * a.ldx + row] != 0.0;  // !iszero: NaN counts
        for (int64_t t = 0; t < a.nt; ++t) k += a.Y[t * a.ldy + yrow] != 0.0;
        a.ks[s0 + i] = k;
    } else if (blockIdx.z == 1) {
        if (i >= nf) return;
        const double* col = a.X + int64_t(a.f_idx[f0 + i]) * a.ldx;
        int k = 0;
        for (int s = 0; s < ns; ++s) k += col[a.s_idx[s0 + s]] != 0.0;
        a.kf[f0 + i] = k;
    } else {
        if (i >= a.nt) return;
        const double* col = a.Y + int64_t(i) * a.ldy;
        int k = 0;
        for (int s = 0; s < ns; ++s) k += col[a.ys_idx[s0 + s]] != 0.0;
        a.kt[int64_t(f) * a.nt + i] = k;
    }
}

constexpr int FT = 64;  // tile edge
constexpr int FK = 16;  // k slab

// MODE 0: C = T_f (M = features, N = targets, K = sources); MODE 1: C = R_f (M = queries, N = targets, K = features)
template <int MODE>
__global__ void __launch_bounds__(256) folds_gemm_kernel(const FoldsArgs a) {
    __shared__ double As[FK][FT + 2];
    __shared__ double Bs[FK][FT + 2];
    const int f = blockIdx.y;
    const int32_t q0 = a.q_ptr[f] - a.q_ptr[0], nq = a.q_ptr[f + 1] - a.q_ptr[f];
    const int32_t s0 = a.s_ptr[f] - a.s_ptr[0], ns = a.s_ptr[f + 1] - a.s_ptr[f];
    const int32_t f0 = a.f_ptr[f] - a.f_ptr[0], nf = a.f_ptr[f + 1] - a.f_ptr[f];
    const int M = MODE == 0 ? nf : nq, K = MODE == 0 ? ns : nf;
    const int64_t N = a.nt;
    const int tiles_n = int((N + FT - 1) / FT);
    const int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
    const int m0 = tm * FT;
    const int64_t n0 = int64_t(tn) * FT;
    if (m0 >= M) return;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    double* Tf = a.T + int64_t(f) * a.nt * a.ldt;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += FK) {
        // stage the slabs: 16 x 64 elements each, 4 per thread; gathers through the fold's index lists
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + e * 256;
            if (MODE == 0) {
                const int k = idx & 15, m = idx >> 4;  // consecutive threads walk the sources (rows of X / Y)
                double va = 0.0, vb = 0.0;
                if (k0 + k < K) {
                    const int32_t srow = a.s_idx[s0 + k0 + k];
                    if (m0 + m < M) va = a.X[int64_t(a.f_idx[f0 + m0 + m]) * a.ldx + srow];
                    if (n0 + m < N) {
                        const double y = a.Y[(n0 + m) * a.ldy + a.ys_idx[s0 + k0 + k]];
                        const int kk = a.ks[s0 + k0 + k];
                        const double w = y / double(kk);     // W = G ./ k, then Inf -> 0, NaN -> 0 (src/core.jl:366-368)
                        vb = (w - w == 0.0) ? w : 0.0;       // finite
                    }
                }
                As[k][m] = va;
                Bs[k][m] = vb;
            } else {
                const int m = idx & 63, k = idx >> 6;  // A: consecutive threads walk the queries
                double va = 0.0;
                if (k0 + k < K && m0 + m < M) va = a.X[int64_t(a.f_idx[f0 + k0 + k]) * a.ldx + a.q_idx[q0 + m0 + m]];
                As[k][m] = va;
                const int kb = idx & 15, nb = idx >> 4;  // B: consecutive threads walk the rows of T_f
                double vb = 0.0;
                if (k0 + kb < K && n0 + nb < N) vb = Tf[(n0 + nb) * a.ldt + k0 + kb];
                Bs[kb][nb] = vb;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < FK; ++k) {
            double av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t n = n0 + tx * 4 + j;
            if (n >= N) continue;
            if (MODE == 0) {
                const int kk = a.kf[f0 + m];
                Tf[n * a.ldt + m] = kk > 0 ? acc[i][j] / double(kk) : 0.0;
            } else {
                const bool flag = a.clean && a.kt[int64_t(f) * a.nt + n] == 0;  // clean!: target without edges -> -99
                a.R[n * a.ldr + q0 + m] = flag ? -99.0 : acc[i][j];
            }
        }
    }
}

}  // namespace

namespace ss {

// device index lists already uploaded by the caller (dq, dsx, dsy, df rebased to the first fold); host pointers for sizes
int32_t predict_query_folds_batched(ss_ctx* ctx, const ss_mat* X, const ss_mat* Y, int nfolds, const int32_t* q_ptr,
                                    const int32_t* s_ptr, const int32_t* f_ptr, const int32_t* dq, const int32_t* dsx,
                                    const int32_t* dsy, const int32_t* df, ss_mat* R, bool clean) {
    const int64_t nt = Y->cols;
    int64_t mq = 0, ms = 0, mf = 0;
    for (int f = 0; f < nfolds; ++f) {
        mq = std::max<int64_t>(mq, q_ptr[f + 1] - q_ptr[f]);
        ms = std::max<int64_t>(ms, s_ptr[f + 1] - s_ptr[f]);
        mf = std::max<int64_t>(mf, f_ptr[f + 1] - f_ptr[f]);
    }
    const int64_t nsi = s_ptr[nfolds] - s_ptr[0], nfi = f_ptr[nfolds] - f_ptr[0];
    const int64_t ldt = round_up(std::max<int64_t>(mf, 1), 16);
    void* p;
    const size_t ptr_ints = size_t(3) * (nfolds + 1);
    const size_t k_ints = size_t(nsi + nfi) + size_t(nfolds) * nt;
    SS_TRY(scratch_get(ctx, 20, (ptr_ints + k_ints + 16) * 4, &p));
    int32_t* dptr = static_cast<int32_t*>(p);
    std::vector<int32_t> hp(ptr_ints);
    for (int f = 0; f <= nfolds; ++f) {
        hp[f] = q_ptr[f];
        hp[nfolds + 1 + f] = s_ptr[f];
        hp[2 * (nfolds + 1) + f] = f_ptr[f];
    }
    // pageable source: the copy is staged by the driver before the call returns, so `hp` may go out of scope
    SS_CHECK_CUDA(cudaMemcpyAsync(dptr, hp.data(), ptr_ints * 4, cudaMemcpyHostToDevice, ctx->stream));
    FoldsArgs a{};
    a.X = X->d;
    a.ldx = X->ld;
    a.Y = Y->d;
    a.ldy = Y->ld;
    a.nt = nt;
    a.nfolds = nfolds;
    a.q_ptr = dptr;
    a.s_ptr = dptr + (nfolds + 1);
    a.f_ptr = dptr + 2 * (nfolds + 1);
    a.q_idx = dq;
    a.s_idx = dsx;
    a.ys_idx = dsy;
    a.f_idx = df;
    a.ks = dptr + ptr_ints;
    a.kf = a.ks + nsi;
    a.kt = a.kf + nfi;
    SS_TRY(scratch_get(ctx, 21, size_t(nfolds) * size_t(nt) * size_t(ldt) * 8, &p));
    a.T = static_cast<double*>(p);
    a.ldt = ldt;
    a.R = R->d;
    a.ldr = R->ld;
    a.clean = clean ? 1 : 0;
    const int64_t longest = std::max({ms, mf, nt, int64_t(1)});
    folds_degrees_kernel<<<dim3(unsigned(ceil_div(longest, 256)), unsigned(nfolds), 3), 256, 0, ctx->stream>>>(a);
    const int64_t tiles_n = ceil_div(nt, FT);
    if (mf > 0 && nt > 0)
        folds_gemm_kernel<0><<<dim3(unsigned(ceil_div(mf, FT) * tiles_n), unsigned(nfolds)), 256, 0, ctx->stream>>>(a);
    if (mq > 0 && nt > 0)
        folds_gemm_kernel<1><<<dim3(unsigned(ceil_div(mq, FT) * tiles_n), unsigned(nfolds)), 256, 0, ctx->stream>>>(a);
    ctx->launches += 3;
    SS_CHECK_CUDA(cudaGetLastError());
    return SS_OK;
}

}  // namespace ss
