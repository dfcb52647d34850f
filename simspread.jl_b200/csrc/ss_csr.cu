// Fused alpha-threshold -> CSR featurization (north-star subsystem 1).
//
// Reference semantics: cutoff.(S, alpha, weighted) (src/core.jl:37-43, :106-112); an entry becomes
// an edge iff the thresholded value is non-zero (src/graphs.jl:10).  Output is the CSR of S: row =
// source node, column indices ascending -- bit-exact with a row-wise scan of the dense result.
//
// S is column-major, so rows are strided.  A block owns 32 consecutive rows and sweeps all
// columns in tiles of 64: the tile is loaded with lanes along rows (256-byte coalesced runs of one
// column), transposed through padded shared memory, and compacted with lanes along columns:
// __ballot_sync + popc prefix give every kept entry its slot, so the col_idx / value stores of one
// row are contiguous.  The same kernel run in COUNT mode produces the row counts for the scan -- and the keep-mask
// (one bit per cell), which the fill pass of a SPARSE result replays instead of reading S a second time
// (csr_fill_mask_kernel): S is then read once, plus the sectors of the kept entries.
#include <stdlib.h>
#include <string.h>

#include "ss_common.cuh"

namespace {

constexpr int ROWS_PER_BLOCK = 32;
constexpr int TILE_COLS = 64;
constexpr int CSR_TPB = 256;  // 8 warps

__device__ __forceinline__ bool keep_edge(double x, double alpha, bool weighted) {
    // thresholded value != 0: binary -> 1.0 when x >= alpha; weighted -> x itself (x == 0 is no edge)
    return (x >= alpha) && (!weighted || x != 0.0);
}

template <bool COUNT_ONLY>
__global__ void __launch_bounds__(CSR_TPB)
    csr_kernel(const double* __restrict__ S, int64_t rows, int64_t cols, int64_t ld, double alpha, int weighted,
               int32_t* __restrict__ row_count, const int32_t* __restrict__ row_ptr, int32_t* __restrict__ col_idx,
               double* __restrict__ values, uint32_t* __restrict__ keep_mask, int64_t mask_words) {
    __shared__ double tile[TILE_COLS][ROWS_PER_BLOCK + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = int64_t(blockIdx.x) * ROWS_PER_BLOCK;
    const bool w = weighted != 0;
    // warp `warp` compacts rows 4*warp .. 4*warp+3 of the block
    int32_t cursor[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t r = row0 + 4 * warp + q;
        cursor[q] = (COUNT_ONLY || r >= rows) ? 0 : row_ptr[r];
    }
    for (int64_t c0 = 0; c0 < cols; c0 += TILE_COLS) {
        // load: warp handles columns warp, warp+8, ...; lane = row inside the block
#pragma unroll
        for (int cc = warp; cc < TILE_COLS; cc += CSR_TPB / 32) {
            const int64_t c = c0 + cc, r = row0 + lane;
            // NaN fails x >= alpha, so out-of-range cells are never edges
            tile[cc][lane] = (c < cols && r < rows) ? __ldg(S + c * ld + r) : __longlong_as_double(0x7ff8000000000000ll);
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int rr = 4 * warp + q;
#pragma unroll
            for (int h = 0; h < TILE_COLS / 32; ++h) {
                const double x = tile[32 * h + lane][rr];
                const bool keep = keep_edge(x, alpha, w);
                const unsigned ballot = __ballot_sync(0xffffffffu, keep);
                if (COUNT_ONLY && keep_mask && lane == 0 && row0 + rr < rows && c0 + 32 * h < cols)
                    keep_mask[(row0 + rr) * mask_words + (c0 >> 5) + h] = ballot;  // replayed by csr_fill_mask_kernel
                if (!COUNT_ONLY && keep) {
                    const int32_t pos = cursor[q] + __popc(ballot & ((1u << lane) - 1u));
                    col_idx[pos] = int32_t(c0 + 32 * h + lane);
                    if (values) values[pos] = x;
                }
                cursor[q] += __popc(ballot);
            }
        }
        __syncthreads();
    }
    if (COUNT_ONLY && lane == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t r = row0 + 4 * warp + q;
            if (r < rows) row_count[r] = cursor[q];
        }
    }
}

// Fill pass for sparse results: instead of reading S a second time, replay the keep-mask written by the count pass
// (1/64 of the bytes of S) and touch only the sectors of the kept entries.  One warp per row; lane l takes mask word
// l, l + 32, ...; a warp prefix over the popcounts gives every kept entry its slot, so col_idx / values of a row are
// written in ascending column order, bit-identical to the tiled fill.
__global__ void __launch_bounds__(256)
    csr_fill_mask_kernel(const double* __restrict__ S, int64_t rows, int64_t ld, const uint32_t* __restrict__ keep_mask,
                         int64_t mask_words, const int32_t* __restrict__ row_ptr, int32_t* __restrict__ col_idx,
                         double* __restrict__ values) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const uint32_t* m = keep_mask + row * mask_words;
    int32_t base = row_ptr[row];
    for (int64_t w0 = 0; w0 < mask_words; w0 += 32) {
        const int64_t w = w0 + lane;
        uint32_t bits = w < mask_words ? __ldg(m + w) : 0u;
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        int32_t pos = base + incl - cnt;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int64_t c = (w << 5) + b;
            col_idx[pos] = int32_t(c);
            if (values) values[pos] = __ldg(S + c * ld + row);  // weighted: the kept value is the similarity itself
            ++pos;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// single-block exclusive scan of n int32 counts into out[0..n]; overflow (> INT32_MAX) reported
// through *overflow.
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int32_t* __restrict__ in, int64_t n,
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

}  // namespace

namespace ss {

int32_t featurize_csr(ss_ctx* ctx, const ss_mat* S, double alpha, bool weighted, ss_csr** out) {
    *out = nullptr;
    const int64_t rows = S->rows, cols = S->cols;
    SS_REQUIRE(rows < (1ll << 31) - 64 && cols < (1ll << 31), "featurize_csr: matrix too large for int32 indices");
    ss_csr* c = new ss_csr();
    c->ctx = ctx;
    c->rows = rows;
    c->cols = cols;
    int32_t status = SS_OK;
    auto fail = [&](int32_t s) {
        ss_csr_destroy(c);
        return s;
    };
    if (cudaMallocAsync(reinterpret_cast<void**>(&c->row_ptr), size_t(rows + 2) * 4, ctx->stream) != cudaSuccess) {
        cudaGetLastError();
        set_error("featurize_csr: out of device memory");
        return fail(SS_ERR_OOM);
    }
    void* p;
    if ((status = scratch_get(ctx, 8, size_t(rows + 2) * 4, &p)) != SS_OK) return fail(status);
    int32_t* counts = static_cast<int32_t*>(p);
    int32_t* overflow = counts + rows + 1;
    const unsigned grid = unsigned(ceil_div(rows > 0 ? rows : 1, ROWS_PER_BLOCK));
    // keep-mask of the count pass (one bit per cell): lets the fill pass of a sparse result skip the second read of S
    const int64_t mask_words = ceil_div(cols, 32);
    uint32_t* keep_mask = nullptr;
    {
        const char* e = getenv("SS_CSR_FILL");  // "tiled": always re-read S through the transposing tiles (A/B runs)
        if (!(e && !strcmp(e, "tiled")) && rows > 0 && cols > 0) {
            void* mp;
            if ((status = scratch_get(ctx, 22, size_t(rows) * size_t(mask_words) * 4, &mp)) != SS_OK) return fail(status);
            keep_mask = static_cast<uint32_t*>(mp);
        }
    }
    if (rows > 0 && cols > 0) {
        csr_kernel<true><<<grid, CSR_TPB, 0, ctx->stream>>>(S->d, rows, cols, S->ld, alpha, weighted ? 1 : 0, counts,
                                                             nullptr, nullptr, nullptr, keep_mask, mask_words);
        ctx->launches++;
    } else {
        cudaMemsetAsync(counts, 0, size_t(rows + 1) * 4, ctx->stream);
    }
    scan_counts_kernel<<<1, 1024, 0, ctx->stream>>>(counts, rows, c->row_ptr, overflow);
    ctx->launches++;
    int32_t h[2] = {0, 0};
    cudaMemcpyAsync(&h[0], c->row_ptr + rows, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&h[1], overflow, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        set_error("featurize_csr: %s", cudaGetErrorString(e));
        return fail(SS_ERR_CUDA);
    }
    if (h[1]) {
        set_error("featurize_csr: more than 2^31-1 edges; int32 CSR cannot hold them");
        return fail(SS_ERR_UNSUPPORTED);
    }
    c->nnz = h[0];
    const size_t n1 = size_t(c->nnz > 0 ? c->nnz : 1);
    if (cudaMallocAsync(reinterpret_cast<void**>(&c->col_idx), n1 * 4, ctx->stream) != cudaSuccess ||
        (weighted && cudaMallocAsync(reinterpret_cast<void**>(&c->values), n1 * 8, ctx->stream) != cudaSuccess)) {
        cudaGetLastError();
        set_error("featurize_csr: out of device memory for %lld edges", (long long)c->nnz);
        return fail(SS_ERR_OOM);
    }
    if (c->nnz > 0) {
        // below ~10 % density the mask replay moves fewer bytes than a second sweep of S (a kept value costs one
        // 32-byte sector); above it the transposing tiles are the better access pattern
        const bool replay = keep_mask && double(c->nnz) < 0.10 * double(rows) * double(cols);
        if (replay) {
            csr_fill_mask_kernel<<<unsigned(ceil_div(rows * 32, 256)), 256, 0, ctx->stream>>>(S->d, rows, S->ld, keep_mask, mask_words,
                                                                                          c->row_ptr, c->col_idx, c->values);
        } else {
            csr_kernel<false><<<grid, CSR_TPB, 0, ctx->stream>>>(S->d, rows, cols, S->ld, alpha, weighted ? 1 : 0, nullptr,
                                                                  c->row_ptr, c->col_idx, c->values, nullptr, 0);
        }
        ctx->launches++;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("featurize_csr: %s", cudaGetErrorString(e));
        return fail(SS_ERR_CUDA);
    }
    *out = c;
    return SS_OK;
}

}  // namespace ss
