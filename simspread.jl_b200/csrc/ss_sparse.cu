// Sparse form of the predict chain (north-star subsystem 3, "when A is sparse"): at high alpha the
// feature blocks Xq / Xs keep only a few percent of their entries, and the two products of
// `A * W^2` (reference src/core.jl:413) are cheaper as row-split SpMMs than as dense DMMA GEMMs:
//
//     T[f,:] = (sum_{s in nz(Xs[:,f])} Xs[s,f] * Wst[s,:]) / kf[f]      (CSR of Xs' x dense Wst)
//     R[q,:] =  sum_{f in nz(Xq[q,:])} Xq[q,f] * T[f,:]                 (CSR of Xq  x dense T)
//
// The dense right-hand sides are kept ROW-major (one target row contiguous), so a block stages the
// (column, value) list of a sparse row in shared memory and every thread streams its own target
// columns of the referenced rows with 128-bit loads.  Work unit = one partial product (a_ij * B[j,c]);
// algorithmic bytes = 8 B of B per partial product (+ 12 B per non-zero, amortised over the tile).
// Sums run in ascending column order of the CSR row: deterministic.
#include "ss_common.cuh"

namespace {

constexpr int SP_TPB = 128;         // threads; each owns two adjacent target columns (double2)
constexpr int SP_COLS = SP_TPB * 2; // 256 target columns per block
constexpr int SP_ROWS = 32;         // sparse rows per block (transposed-output tile height)
constexpr int SP_CHUNK = 128;       // non-zeros staged per pass

// CSR of S' (one "row" per COLUMN of the column-major S): lanes run along the contiguous column,
// kept entries are compacted with __ballot_sync + popc prefix, warps are ordered by a block scan.
template <bool COUNT_ONLY>
__global__ void __launch_bounds__(256)
    csc_kernel(const double* __restrict__ S, int64_t rows, int64_t cols, int64_t ld, double alpha, int weighted,
               int32_t* __restrict__ col_count, const int32_t* __restrict__ col_ptr, int32_t* __restrict__ row_idx,
               double* __restrict__ values) {
    __shared__ int wcount[8];
    __shared__ int base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool w = weighted != 0;
    for (int64_t c = blockIdx.x; c < cols; c += gridDim.x) {
        if (threadIdx.x == 0) base_s = COUNT_ONLY ? 0 : col_ptr[c];
        __syncthreads();
        const double* col = S + c * ld;
        for (int64_t r0 = 0; r0 < rows; r0 += 256) {
            const int64_t r = r0 + threadIdx.x;
            const double x = (r < rows) ? __ldg(col + r) : __longlong_as_double(0x7ff8000000000000ll);
            const bool keep = (x >= alpha) && (!w || x != 0.0);
            const unsigned ballot = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) wcount[warp] = __popc(ballot);
            __syncthreads();
            int before = 0, total = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int v = wcount[i];
                if (i < warp) before += v;
                total += v;
            }
            const int base = base_s;
            if (!COUNT_ONLY && keep) {
                const int pos = base + before + __popc(ballot & ((1u << lane) - 1u));
                row_idx[pos] = int32_t(r);
                if (values) values[pos] = x;
            }
            __syncthreads();
            if (threadIdx.x == 0) base_s = base + total;
        }
        __syncthreads();
        if (COUNT_ONLY && threadIdx.x == 0) col_count[c] = base_s;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) scan_counts2_kernel(const int32_t* __restrict__ in, int64_t n,
                                                            int32_t* __restrict__ out, int32_t* overflow) {
    __shared__ long long part[1024];
    const int t = threadIdx.x;
    const int64_t chunk = (n + 1023) / 1024;
    const int64_t b = t * chunk, e = min(n, b + chunk);
    long long s = 0;
    for (int64_t i = b; i < e; ++i) s += in[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        long long run = 0;
        for (int i = 0; i < 1024; ++i) {
            const long long v = part[i];
            part[i] = run;
            run += v;
        }
        out[n] = int32_t(run);
        *overflow = run > 2147483647ll;
    }
    __syncthreads();
    long long run = part[t];
    for (int64_t i = b; i < e; ++i) {
        out[i] = int32_t(run);
        run += in[i];
    }
}

// column histogram of a CSR (degree of the nodes on the column side): integer atomics, exact
__global__ void __launch_bounds__(256)
    csr_col_hist_kernel(const int32_t* __restrict__ col_idx, int64_t nnz, int32_t* __restrict__ hist) {
    for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < nnz; i += int64_t(gridDim.x) * 256)
        atomicAdd(hist + col_idx[i], 1);
}

// Wt[s][t] (row-major, ldw) = spread(Y[s,t], ks[s]) from the column-major Y: 32x32 smem transpose
__global__ void __launch_bounds__(256)
    spread_transpose_kernel(const double* __restrict__ Y, int64_t rows, int64_t cols, int64_t ldy,
                            const int32_t* __restrict__ k, double* __restrict__ Wt, int64_t ldw) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t r0 = int64_t(blockIdx.x) * 32, c0 = int64_t(blockIdx.y) * 32;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int64_t r = r0 + tx, c = c0 + j;
        double v = 0.0;
        if (r < rows && c < cols) {
            const double x = Y[c * ldy + r];
            const double kk = double(k[r]);
            double q;
            if (kk != 0.0 && x == 0.0) q = x;
            else q = x / kk;  // true division, reference src/core.jl:366
            v = (q != q || q == __longlong_as_double(0x7ff0000000000000ll)) ? 0.0 : q;
        }
        tile[j][tx] = v;  // tile[c][r]
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int64_t r = r0 + j, c = c0 + tx;
        if (r < rows && c < cols) Wt[r * ldw + c] = tile[tx][j];
    }
}

struct SpmmParams {
    const int32_t* row_ptr;
    const int32_t* col_idx;
    const double* values;  // null: every stored entry is 1.0 (binary features)
    int64_t rows;          // sparse rows
    const double* B;       // dense, row-major: B[j * ldb + c]
    int64_t ldb;
    int64_t ncols;         // target columns
    double* out;
    int64_t ldo;
    int div_by_rowlen;        // out[i,:] /= nnz(row i)   (the 1/kf of W[f,:]; 0 rows -> 0)
    const int32_t* col_flag;  // optional: out[:,c] = -99 where col_flag[c] == 0 (clean!)
};

// TRANSPOSED_OUT == false: out is row-major  (out[i * ldo + c])   -- used for T
// TRANSPOSED_OUT == true : out is column-major (out[c * ldo + i]) -- used for R (Julia layout)
template <bool TRANSPOSED_OUT>
__global__ void __launch_bounds__(SP_TPB) spmm_kernel(const SpmmParams p) {
    extern __shared__ double sp_smem[];
    __shared__ int32_t sj[SP_CHUNK];
    __shared__ double sa[SP_CHUNK];
    double(*tile)[SP_COLS + 1] = reinterpret_cast<double(*)[SP_COLS + 1]>(sp_smem);  // [SP_ROWS][SP_COLS+1]
    const int64_t row0 = int64_t(blockIdx.x) * SP_ROWS;
    const int64_t c0 = int64_t(blockIdx.y) * SP_COLS;
    const int64_t c = c0 + 2 * threadIdx.x;
    const bool in0 = c < p.ncols, in1 = c + 1 < p.ncols;
    const bool vec = in1 && ((p.ldb & 1) == 0);
    for (int r = 0; r < SP_ROWS; ++r) {
        const int64_t row = row0 + r;
        if (row >= p.rows) break;  // uniform across the block
        const int32_t beg = p.row_ptr[row], end = p.row_ptr[row + 1];
        double a0 = 0.0, a1 = 0.0;
        for (int32_t e0 = beg; e0 < end; e0 += SP_CHUNK) {
            const int n = min(SP_CHUNK, end - e0);
            __syncthreads();
            if (threadIdx.x < n) {
                sj[threadIdx.x] = p.col_idx[e0 + threadIdx.x];
                sa[threadIdx.x] = p.values ? p.values[e0 + threadIdx.x] : 1.0;
            }
            __syncthreads();
            if (in0) {
#pragma unroll 4
                for (int e = 0; e < n; ++e) {
                    const double* b = p.B + int64_t(sj[e]) * p.ldb + c;
                    const double a = sa[e];
                    if (vec) {
                        const double2 v = *reinterpret_cast<const double2*>(b);
                        a0 = fma(a, v.x, a0);
                        a1 = fma(a, v.y, a1);
                    } else {
                        a0 = fma(a, b[0], a0);
                        if (in1) a1 = fma(a, b[1], a1);
                    }
                }
            }
        }
        if (p.div_by_rowlen) {
            const int len = end - beg;
            a0 = len ? a0 / double(len) : 0.0;
            a1 = len ? a1 / double(len) : 0.0;
        }
        if (TRANSPOSED_OUT) {
            tile[r][2 * threadIdx.x] = a0;
            tile[r][2 * threadIdx.x + 1] = a1;
        } else {
            if (in0) p.out[row * p.ldo + c] = a0;
            if (in1) p.out[row * p.ldo + c + 1] = a1;
        }
    }
    if (TRANSPOSED_OUT) {
        __syncthreads();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int64_t row = row0 + lane;
        for (int cc = warp; cc < SP_COLS; cc += SP_TPB / 32) {
            const int64_t col = c0 + cc;
            if (col < p.ncols && row < p.rows) {
                double v = tile[lane][cc];
                if (p.col_flag && __ldg(p.col_flag + col) == 0) v = -99.0;
                p.out[col * p.ldo + row] = v;
            }
        }
    }
}

int32_t launch_spmm(ss_ctx* ctx, const SpmmParams& p, bool transposed_out) {
    if (p.rows == 0 || p.ncols == 0) return SS_OK;
    const int64_t gx = ss::ceil_div(p.rows, SP_ROWS), gy = ss::ceil_div(p.ncols, SP_COLS);
    SS_REQUIRE(gy <= 65535, "spmm: too many target columns");
    dim3 grid{unsigned(gx), unsigned(gy)};
    if (transposed_out) {
        const size_t smem = size_t(SP_ROWS) * (SP_COLS + 1) * 8;
        SS_CHECK_CUDA(cudaFuncSetAttribute(spmm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        spmm_kernel<true><<<grid, SP_TPB, smem, ctx->stream>>>(p);
    } else {
        spmm_kernel<false><<<grid, SP_TPB, 0, ctx->stream>>>(p);
    }
    SS_CHECK_CUDA(cudaGetLastError());
    ctx->launches++;
    return SS_OK;
}

}  // namespace

namespace ss {

int32_t featurize_csc(ss_ctx* ctx, const ss_mat* S, double alpha, bool weighted, ss_csr** out) {
    *out = nullptr;
    const int64_t rows = S->rows, cols = S->cols;
    SS_REQUIRE(rows < (1ll << 31) && cols < (1ll << 31) - 64, "featurize_csc: matrix too large for int32 indices");
    ss_csr* c = new ss_csr();
    c->ctx = ctx;
    c->rows = cols;  // CSR of S'
    c->cols = rows;
    auto fail = [&](int32_t s) {
        ss_csr_destroy(c);
        return s;
    };
    if (cudaMallocAsync(reinterpret_cast<void**>(&c->row_ptr), size_t(cols + 2) * 4, ctx->stream) != cudaSuccess) {
        cudaGetLastError();
        set_error("featurize_csc: out of device memory");
        return fail(SS_ERR_OOM);
    }
    void* p;
    int32_t st;
    if ((st = scratch_get(ctx, 8, size_t(cols + 2) * 4, &p)) != SS_OK) return fail(st);
    int32_t* counts = static_cast<int32_t*>(p);
    int32_t* overflow = counts + cols + 1;
    int grid = int(cols < int64_t(ctx->sm_count) * 8 ? (cols > 0 ? cols : 1) : int64_t(ctx->sm_count) * 8);
    if (rows > 0 && cols > 0) {
        csc_kernel<true><<<grid, 256, 0, ctx->stream>>>(S->d, rows, cols, S->ld, alpha, weighted ? 1 : 0, counts, nullptr,
                                                         nullptr, nullptr);
        ctx->launches++;
    } else {
        cudaMemsetAsync(counts, 0, size_t(cols + 1) * 4, ctx->stream);
    }
    scan_counts2_kernel<<<1, 1024, 0, ctx->stream>>>(counts, cols, c->row_ptr, overflow);
    ctx->launches++;
    int32_t h[2] = {0, 0};
    cudaMemcpyAsync(&h[0], c->row_ptr + cols, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&h[1], overflow, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        set_error("featurize_csc: %s", cudaGetErrorString(e));
        return fail(SS_ERR_CUDA);
    }
    if (h[1]) {
        set_error("featurize_csc: more than 2^31-1 edges; int32 CSR cannot hold them");
        return fail(SS_ERR_UNSUPPORTED);
    }
    c->nnz = h[0];
    const size_t n1 = size_t(c->nnz > 0 ? c->nnz : 1);
    if (cudaMallocAsync(reinterpret_cast<void**>(&c->col_idx), n1 * 4, ctx->stream) != cudaSuccess ||
        (weighted && cudaMallocAsync(reinterpret_cast<void**>(&c->values), n1 * 8, ctx->stream) != cudaSuccess)) {
        cudaGetLastError();
        set_error("featurize_csc: out of device memory for %lld edges", (long long)c->nnz);
        return fail(SS_ERR_OOM);
    }
    if (c->nnz > 0) {
        csc_kernel<false><<<grid, 256, 0, ctx->stream>>>(S->d, rows, cols, S->ld, alpha, weighted ? 1 : 0, nullptr,
                                                          c->row_ptr, c->col_idx, c->values);
        ctx->launches++;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("featurize_csc: %s", cudaGetErrorString(e));
        return fail(SS_ERR_CUDA);
    }
    *out = c;
    return SS_OK;
}

// R = Xq * T, T = (Xs' * (Y ./ ks)) ./ kf with CSR feature blocks (Xq: Nq x Nf, XsT: Nf x Ns = CSR of
// Xs') and a dense column-major Y (Ns x Nt).  R dense column-major (Nq x Nt).
int32_t predict_query_csr(ss_ctx* ctx, const ss_csr* Xq, const ss_csr* XsT, const ss_mat* Y, ss_mat* R, uint32_t flags,
                          int32_t* kt_out) {
    const int64_t nq = Xq->rows, nf = Xq->cols, ns = Y->rows, nt = Y->cols;
    if (nq == 0 || nt == 0) return SS_OK;
    void* p;
    const size_t kbytes = size_t(round_up(ns, 64) + round_up(nt, 64)) * 4;
    SS_TRY(scratch_get(ctx, 0, kbytes, &p));
    int32_t* ks = static_cast<int32_t*>(p);
    int32_t* kt = ks + round_up(ns, 64);
    SS_CHECK_CUDA(cudaMemsetAsync(ks, 0, kbytes, ctx->stream));
    // ks = nnz_row(Xs) + nnz_row(Y): column histogram of XsT + dense row counts of Y; kt = nnz_col(Y)
    if (XsT->nnz > 0) {
        int grid = int(ceil_div(XsT->nnz, 256 * 8));
        if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
        csr_col_hist_kernel<<<grid, 256, 0, ctx->stream>>>(XsT->col_idx, XsT->nnz, ks);
        ctx->launches++;
    }
    SS_TRY(launch_degrees(ctx, Y->d, ns, nt, Y->ld, ks, kt));
    // Wt (row-major Ns x Nt)
    const int64_t ldw = round_up(nt, 16);
    SS_TRY(scratch_get(ctx, 1, size_t(ldw) * size_t(ns > 0 ? ns : 1) * 8, &p));
    double* Wt = static_cast<double*>(p);
    if (ns > 0) {
        dim3 g{unsigned(ceil_div(ns, 32)), unsigned(ceil_div(nt, 32))};
        SS_REQUIRE(g.y <= 65535, "predict_query_csr: too many targets");
        spread_transpose_kernel<<<g, 256, 0, ctx->stream>>>(Y->d, ns, nt, Y->ld, ks, Wt, ldw);
        ctx->launches++;
    }
    // T (row-major Nf x Nt) = (XsT * Wt) ./ kf,  kf[f] = nnz of row f of XsT
    SS_TRY(scratch_get(ctx, 2, size_t(ldw) * size_t(nf > 0 ? nf : 1) * 8, &p));
    double* T = static_cast<double*>(p);
    SpmmParams a{};
    a.row_ptr = XsT->row_ptr;
    a.col_idx = XsT->col_idx;
    a.values = XsT->values;
    a.rows = nf;
    a.B = Wt;
    a.ldb = ldw;
    a.ncols = nt;
    a.out = T;
    a.ldo = ldw;
    a.div_by_rowlen = 1;
    a.col_flag = nullptr;
    SS_TRY(launch_spmm(ctx, a, false));
    SpmmParams b{};
    b.row_ptr = Xq->row_ptr;
    b.col_idx = Xq->col_idx;
    b.values = Xq->values;
    b.rows = nq;
    b.B = T;
    b.ldb = ldw;
    b.ncols = nt;
    b.out = R->d;
    b.ldo = R->ld;
    b.div_by_rowlen = 0;
    b.col_flag = (flags & SS_PREDICT_CLEAN) ? kt : nullptr;
    SS_TRY(launch_spmm(ctx, b, true));
    if (kt_out) SS_CHECK_CUDA(cudaMemcpyAsync(kt_out, kt, size_t(nt) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    return SS_OK;
}

}  // namespace ss
