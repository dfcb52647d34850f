// HBM-bound passes of the hot path: featurize (cutoff), block gather (construct), degrees (k),
// spread (G ./ k) and clean!.  All matrices are column-major float64; a thread owns two adjacent
// rows (one 128-bit load per column) and walks over columns, so every warp access is a 512-byte
// contiguous run of one column.  Grids are sized in multiples of the SM count.
#include "ss_common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int COLS_PER_BLOCK = 64;  // degrees: columns handled by one block (smem column counters)

__device__ __forceinline__ double2 ld2(const double* p) {
    return *reinterpret_cast<const double2*>(p);
}
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

// reference src/core.jl:37-43: x >= alpha ? (weighted ? x : 1.0) : 0.0   (NaN >= alpha is false)
__device__ __forceinline__ double cutoff1(double x, double alpha, bool weighted) {
    return (x >= alpha) ? (weighted ? x : 1.0) : 0.0;
}

__global__ void __launch_bounds__(TPB)
    featurize_kernel(const double* S, int64_t rows, int64_t cols, int64_t lds, double alpha,
                     int weighted, double* X, int64_t ldx) {  // X may alias S (featurize!)
    const int64_t r = (int64_t(blockIdx.x) * TPB + threadIdx.x) * 2;
    if (r >= rows) return;
    const bool pair = (r + 1 < rows);
    const bool w = weighted != 0;
    int64_t c = blockIdx.y;
    // 4 independent 128-bit loads in flight per thread
    for (; c + 3 * int64_t(gridDim.y) < cols; c += 4 * int64_t(gridDim.y)) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double* p = S + (c + u * int64_t(gridDim.y)) * lds + r;
            if (pair) v[u] = ld2(p);
            else { v[u].x = *p; v[u].y = 0.0; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double* q = X + (c + u * int64_t(gridDim.y)) * ldx + r;
            double2 o;
            o.x = cutoff1(v[u].x, alpha, w);
            o.y = cutoff1(v[u].y, alpha, w);
            if (pair) st2(q, o);
            else *q = o.x;
        }
    }
    for (; c < cols; c += gridDim.y) {
        const double* p = S + c * lds + r;
        double* q = X + c * ldx + r;
        if (pair) {
            double2 v = ld2(p), o;
            o.x = cutoff1(v.x, alpha, w);
            o.y = cutoff1(v.y, alpha, w);
            st2(q, o);
        } else {
            *q = cutoff1(*p, alpha, w);
        }
    }
}

// reference src/core.jl:167,171-172: X[queries, features] etc.  dst[i,j] = src[ridx[i], cidx[j]]
__global__ void __launch_bounds__(TPB)
    gather_kernel(const double* __restrict__ src, int64_t lds, const int32_t* __restrict__ ridx,
                  const int32_t* __restrict__ cidx, double* __restrict__ dst, int64_t rows, int64_t cols,
                  int64_t ldd) {
    const int64_t i = int64_t(blockIdx.x) * TPB + threadIdx.x;
    if (i >= rows) return;
    const int64_t sr = ridx ? ridx[i] : i;
    for (int64_t j = blockIdx.y; j < cols; j += gridDim.y) {
        const int64_t sc = cidx ? cidx[j] : j;
        dst[j * ldd + i] = __ldg(src + sc * lds + sr);
    }
}

// reference src/graphs.jl:9-11 : k = count(!iszero, row).  One pass produces row counts (added
// into row_deg) and column counts (added into col_deg): warp ballot + popc per column, shared
// memory column counters per block, one global atomic per (block, column) and per (thread, row).
__global__ void __launch_bounds__(TPB)
    degrees_kernel(const double* __restrict__ M, int64_t rows, int64_t cols, int64_t ld,
                   int32_t* __restrict__ row_deg, int32_t* __restrict__ col_deg) {
    __shared__ int32_t ccount[COLS_PER_BLOCK];
    const int64_t r = (int64_t(blockIdx.x) * TPB + threadIdx.x) * 2;
    const int64_t c0 = int64_t(blockIdx.y) * COLS_PER_BLOCK;
    const int ncol = int(min(int64_t(COLS_PER_BLOCK), cols - c0));
    if (threadIdx.x < COLS_PER_BLOCK) ccount[threadIdx.x] = 0;
    __syncthreads();
    const bool in0 = r < rows, in1 = r + 1 < rows;
    int cnt0 = 0, cnt1 = 0;
    const int lane = threadIdx.x & 31;
#pragma unroll 4
    for (int j = 0; j < ncol; ++j) {
        double2 v = make_double2(0.0, 0.0);
        const double* p = M + (c0 + j) * ld + r;
        if (in1) v = ld2(p);
        else if (in0) v.x = *p;
        const int nz0 = (v.x != 0.0), nz1 = (v.y != 0.0);  // true for NaN, false for -0.0
        cnt0 += nz0;
        cnt1 += nz1;
        if (col_deg) {
            const unsigned b0 = __ballot_sync(0xffffffffu, nz0);
            const unsigned b1 = __ballot_sync(0xffffffffu, nz1);
            if (lane == 0) {
                const int s = __popc(b0) + __popc(b1);
                if (s) atomicAdd(&ccount[j], s);
            }
        }
    }
    if (row_deg) {
        if (in0 && cnt0) atomicAdd(row_deg + r, cnt0);
        if (in1 && cnt1) atomicAdd(row_deg + r + 1, cnt1);
    }
    if (col_deg) {
        __syncthreads();
        if (threadIdx.x < ncol && ccount[threadIdx.x]) atomicAdd(col_deg + c0 + threadIdx.x, ccount[threadIdx.x]);
    }
}

// reference src/core.jl:365-371 : W = G ./ k(G); replace!(W, Inf => 0.0); replace!(W, NaN => 0.0)
// FP64 division is a ~15-instruction sequence on the (slow) FP64 pipe and would make this pass
// compute-bound.  Bipartite label / binary feature blocks are almost entirely 0.0 or 1.0, and for
// those the IEEE quotient is known without dividing: (+-0)/k = +-0 and 1/k = RN(1/k) (computed once
// per row).  Every other value takes the true division, so the result is bit-identical to `x / k`.
__device__ __forceinline__ double spread1(double x, double kk, double rk) {
    double q;
    if (kk != 0.0 && x == 0.0) q = x;
    else if (x == 1.0) q = rk;
    else q = x / kk;  // IEEE division, as Julia's `/`
    return (q != q || q == __longlong_as_double(0x7ff0000000000000ll)) ? 0.0 : q;
}

__global__ void __launch_bounds__(TPB)
    spread_kernel(const double* G, int64_t rows, int64_t cols, int64_t ldg,
                  const int32_t* __restrict__ k, double* W, int64_t ldw) {  // W may alias G
    const int64_t r = (int64_t(blockIdx.x) * TPB + threadIdx.x) * 2;
    if (r >= rows) return;
    const bool pair = (r + 1 < rows);
    const double k0 = double(k[r]);
    const double k1 = pair ? double(k[r + 1]) : 1.0;
    const double r0 = 1.0 / k0, r1 = 1.0 / k1;
    int64_t c = blockIdx.y;
    for (; c + 3 * int64_t(gridDim.y) < cols; c += 4 * int64_t(gridDim.y)) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double* p = G + (c + u * int64_t(gridDim.y)) * ldg + r;
            if (pair) v[u] = ld2(p);
            else { v[u].x = *p; v[u].y = 0.0; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double* q = W + (c + u * int64_t(gridDim.y)) * ldw + r;
            double2 o;
            o.x = spread1(v[u].x, k0, r0);
            o.y = spread1(v[u].y, k1, r1);
            if (pair) st2(q, o);
            else *q = o.x;
        }
    }
    for (; c < cols; c += gridDim.y) {
        const double* p = G + c * ldg + r;
        double* q = W + c * ldw + r;
        if (pair) {
            double2 v = ld2(p), o;
            o.x = spread1(v.x, k0, r0);
            o.y = spread1(v.y, k1, r1);
            st2(q, o);
        } else {
            *q = spread1(*p, k0, r0);
        }
    }
}

// reference src/core.jl:478-484 : yhat[:, t] .= -99 where k(A[t,:]) == 0
__global__ void __launch_bounds__(TPB)
    clean_kernel(double* __restrict__ R, int64_t rows, int64_t cols, int64_t ld, const int32_t* __restrict__ kt) {
    const int64_t r = int64_t(blockIdx.x) * TPB + threadIdx.x;
    if (r >= rows) return;
    for (int64_t c = blockIdx.y; c < cols; c += gridDim.y)
        if (__ldg(kt + c) == 0) R[c * ld + r] = -99.0;
}

// grid.y so that grid.x * grid.y is about `waves` full waves of the SM count (and <= cols)
inline unsigned pick_grid_y(const ss_ctx* ctx, int64_t gx, int64_t cols, int waves_x_resident) {
    int64_t want = ss::ceil_div(int64_t(ctx->sm_count) * waves_x_resident, gx);
    if (want < 1) want = 1;
    if (want > cols) want = cols;
    if (want > 65535) want = 65535;
    return unsigned(want);
}

}  // namespace

namespace {

// dst (rows x cols, column-major, ld ldd) = transpose of src (cols x rows, column-major, ld lds): the device half of
// ss_mat_upload_rowmajor (a row-major host array is the column-major image of its transpose).  32 x 32 tiles through
// padded shared memory, both sides coalesced.
__global__ void __launch_bounds__(256)
    transpose_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int64_t ldd, int64_t rows,
                     int64_t cols) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows of the tile per pass
    const int64_t c0 = int64_t(blockIdx.x) * 32, r0 = int64_t(blockIdx.y) * 32;
    // src element (c, r) at src[r * lds + c]: lanes along c
    for (int j = ty; j < 32; j += 8) {
        const int64_t r = r0 + j, c = c0 + tx;
        tile[j][tx] = (r < rows && c < cols) ? src[r * lds + c] : 0.0;
    }
    __syncthreads();
    // dst element (r, c) at dst[c * ldd + r]: lanes along r
    for (int j = ty; j < 32; j += 8) {
        const int64_t c = c0 + j, r = r0 + tx;
        if (r < rows && c < cols) dst[c * ldd + r] = tile[tx][j];
    }
}

}  // namespace

namespace ss {

int32_t launch_featurize(ss_ctx* ctx, const double* S, int64_t rows, int64_t cols, int64_t lds,
                         double alpha, bool weighted, double* X, int64_t ldx) {
    if (rows == 0 || cols == 0) return SS_OK;
    SS_REQUIRE((lds % 2) == 0 && (ldx % 2) == 0, "featurize: leading dimensions must be even");
    SS_REQUIRE(((reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(X)) & 15) == 0,
               "featurize: matrices must be 16-byte aligned");
    const int64_t gx = ceil_div(ceil_div(rows, 2), TPB);
    dim3 grid(unsigned(gx), pick_grid_y(ctx, gx, cols, 16));
    featurize_kernel<<<grid, TPB, 0, ctx->stream>>>(S, rows, cols, lds, alpha, weighted ? 1 : 0, X, ldx);
    SS_CHECK_CUDA(cudaGetLastError());
    ctx->launches++;
    return SS_OK;
}

int32_t launch_transpose(ss_ctx* ctx, const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows,
                         int64_t cols) {
    if (rows == 0 || cols == 0) return SS_OK;
    SS_REQUIRE(ceil_div(rows, 32) <= 65535, "transpose: more than 2 M rows are not supported");
    const dim3 grid(unsigned(ceil_div(cols, 32)), unsigned(ceil_div(rows, 32)));
    transpose_kernel<<<grid, 256, 0, ctx->stream>>>(src, lds, dst, ldd, rows, cols);
    SS_CHECK_CUDA(cudaGetLastError());
    ctx->launches++;
    return SS_OK;
}

int32_t launch_gather(ss_ctx* ctx, const double* src, int64_t lds, const int32_t* ridx,
                      const int32_t* cidx, double* dst, int64_t rows, int64_t cols, int64_t ldd) {
    if (rows == 0 || cols == 0) return SS_OK;
    const int64_t gx = ceil_div(rows, TPB);
    dim3 grid(unsigned(gx), pick_grid_y(ctx, gx, cols, 16));
    gather_kernel<<<grid, TPB, 0, ctx->stream>>>(src, lds, ridx, cidx, dst, rows, cols, ldd);
    SS_CHECK_CUDA(cudaGetLastError());
    ctx->launches++;
    return SS_OK;
}

int32_t launch_degrees(ss_ctx* ctx, const double* M, int64_t rows, int64_t cols, int64_t ld,
                       int32_t* row_deg, int32_t* col_deg) {
    if (rows == 0 || cols == 0) return SS_OK;
    SS_REQUIRE((ld % 2) == 0 && (reinterpret_cast<uintptr_t>(M) & 15) == 0,
               "degrees: matrix must be 16-byte aligned with an even leading dimension");
    const int64_t gx = ceil_div(ceil_div(rows, 2), TPB);
    const int64_t gy = ceil_div(cols, COLS_PER_BLOCK);
    SS_REQUIRE(gy <= 65535, "degrees: too many columns (%lld)", (long long)cols);
    dim3 grid{unsigned(gx), unsigned(gy)};
    degrees_kernel<<<grid, TPB, 0, ctx->stream>>>(M, rows, cols, ld, row_deg, col_deg);
    SS_CHECK_CUDA(cudaGetLastError());
    ctx->launches++;
    return SS_OK;
}

int32_t launch_spread_rows(ss_ctx* ctx, const double* G, int64_t rows, int64_t cols, int64_t ldg,
                           const int32_t* k, double* W, int64_t ldw) {
    if (rows == 0 || cols == 0) return SS_OK;
    SS_REQUIRE((ldg % 2) == 0 && (ldw % 2) == 0, "spread: leading dimensions must be even");
    SS_REQUIRE(((reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(W)) & 15) == 0,
               "spread: matrices must be 16-byte aligned");
    const int64_t gx = ceil_div(ceil_div(rows, 2), TPB);
    dim3 grid(unsigned(gx), pick_grid_y(ctx, gx, cols, 16));
    spread_kernel<<<grid, TPB, 0, ctx->stream>>>(G, rows, cols, ldg, k, W, ldw);
    SS_CHECK_CUDA(cudaGetLastError());
    ctx->launches++;
    return SS_OK;
}

int32_t launch_clean(ss_ctx* ctx, double* R, int64_t rows, int64_t cols, int64_t ld,
                     const int32_t* kt) {
    if (rows == 0 || cols == 0) return SS_OK;
    const int64_t gx = ceil_div(rows, TPB);
    dim3 grid(unsigned(gx), pick_grid_y(ctx, gx, cols, 16));
    clean_kernel<<<grid, TPB, 0, ctx->stream>>>(R, rows, cols, ld, kt);
    SS_CHECK_CUDA(cudaGetLastError());
    ctx->launches++;
    return SS_OK;
}

}  // namespace ss
