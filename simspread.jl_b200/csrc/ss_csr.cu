// Fused alpha-threshold -> CSR featurization (north-star subsystem 1).
//
// Reference semantics: cutoff.(S, alpha, weighted) (src/core.jl:37-43, :106-112); an entry becomes
// an edge iff the thresholded value is non-zero (src/graphs.jl:10).  Output is the CSR of S: row =
// source node, column indices ascending -- bit-exact with a row-wise scan of the dense result.
//
// S is column-major, so the rows of the CSR are strided in memory.  Three steps, S is read ONCE:
//   1. csr_count_kernel: lanes along ROWS (every load is a 256-byte run of one column, as in the degree kernel), a
//      warp owns 32 rows x one segment of columns.  A lane thresholds 32 columns of its row into one word of the
//      keep-mask (bit j = column 32 w + j kept) straight from registers -- no shared-memory transpose, no barriers --
//      and counts its row's edges of the segment.  The words of a warp are transposed through shared memory and the
//      mask is stored row-major.
//   2. scan_*: device-wide exclusive scan of the (row, segment) counts in row-major order, which IS the CSR order:
//      at segment 0 of every row it yields row_ptr.
//   3. csr_fill_kernel (results below 10 % density): one warp per row replays the mask -- 1/64 of the bytes of S --
//      and writes col_idx / values of the row with contiguous stores; for weighted graphs only the 32-byte sectors of
//      the kept similarities are read again, up to four per lane in flight.  Denser results keep the transposing tiled fill (csr_tiled_fill_kernel): lanes along rows for
//      the loads, a padded shared-memory tile, lanes along columns with __ballot_sync + popc for the compaction.
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

// ---- 1. count + keep-mask ----------------------------------------------------------------------------------------
// grid.x covers the 32-row blocks (8 per thread block, one per warp), grid.y the column segments of seg_words words
// (4, 8, 16 or 32).  The words of a warp (32 rows x seg_words) are transposed through shared memory so that the mask
// is stored row-major (mask[row][word], seg_words * 4 contiguous bytes per row): the fill walks a row with lanes along
// its words.
constexpr int CNT_BATCH = 16;  // independent 8-byte loads per lane in flight (two batches make one mask word)

//
// Measured and dropped: queueing the kept similarities of a weighted graph from the registers of this kernel (shared-
// memory queues copied to staging slots, read back by the fill as contiguous runs instead of one 32-byte sector of S
// per edge).  The fill drops from 2.0 to 0.6 ms at 4 % density, but the 32 conditional queue stores per word cost the
// count pass 1.2 ms at any density (and placed next to the compares they make ptxas sink every load to its use: 8.9
// ms), so it only pays above ~3.5 % density -- where predict switches to the dense chain anyway.
__global__ void __launch_bounds__(CSR_TPB, 3)
    csr_count_kernel(const double* __restrict__ S, int64_t rows, int64_t cols, int64_t ld, double alpha, int weighted,
                     int64_t mask_words, int seg_words, int nseg, uint32_t* __restrict__ keep_mask,
                     int32_t* __restrict__ seg_count) {
    extern __shared__ uint32_t csr_tr[];  // mask words of the block: [warp][lane][seg_words + 1]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* tr = csr_tr + size_t(warp) * 32 * (seg_words + 1);
    const int64_t rb = int64_t(blockIdx.x) * (CSR_TPB / 32) + warp;
    const int64_t row = rb * 32 + lane;
    if (rb * 32 >= rows) return;  // warp-uniform
    const bool w = weighted != 0;
    const bool live = row < rows;
    const double* base = S + (live ? row : rows - 1);
    const int seg = blockIdx.y;
    const int64_t w0 = int64_t(seg) * seg_words, w1 = min(mask_words, w0 + seg_words);
    int32_t cnt = 0;
    for (int64_t wd = w0; wd < w1; ++wd) {
        const int64_t c0 = wd << 5;
        uint32_t bits = 0;
        if (c0 + 32 <= cols) {
            const double* pc = base + c0 * ld;
#pragma unroll
            for (int h = 0; h < 32 / CNT_BATCH; ++h) {
                double v[CNT_BATCH];
#pragma unroll
                for (int j = 0; j < CNT_BATCH; ++j) v[j] = __ldg(pc + int64_t(h * CNT_BATCH + j) * ld);
#pragma unroll
                for (int j = 0; j < CNT_BATCH; ++j) bits |= uint32_t(keep_edge(v[j], alpha, w)) << (h * CNT_BATCH + j);
            }
        } else {  // last word of a row: clamp the address, drop the bit
#pragma unroll 4
            for (int j = 0; j < 32; ++j) {
                const int64_t c = c0 + j;
                const double x = __ldg(base + min(c, cols - 1) * ld);
                bits |= uint32_t(c < cols && keep_edge(x, alpha, w)) << j;
            }
        }
        if (!live) bits = 0;
        tr[lane * (seg_words + 1) + int(wd - w0)] = bits;
        cnt += __popc(bits);
    }
    if (live) seg_count[row * nseg + seg] = cnt;
    __syncwarp();
    // row-major store of the mask: one instruction covers 32 / seg_words rows x seg_words words
    const int nw = int(w1 - w0);
    const int j = lane % seg_words, r_in = lane / seg_words, r_step = 32 / seg_words;
    for (int rr = r_in; rr < 32; rr += r_step) {
        const int64_t r = rb * 32 + rr;
        if (r < rows && j < nw) keep_mask[r * mask_words + w0 + j] = tr[rr * (seg_words + 1) + j];
    }
}

// ---- 2. device-wide exclusive scan of int32 counts (int64 running sums; overflow of the int32 result reported) ------
constexpr int SCAN_TPB = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_TPB * SCAN_ITEMS;  // 4096 counts per block, thread t owns items [16 t, 16 t + 16)

__global__ void __launch_bounds__(SCAN_TPB)
    scan_block_sums_kernel(const int32_t* __restrict__ in, int64_t n, long long* __restrict__ bsum) {
    __shared__ long long red[SCAN_TPB / 32];
    const int64_t b0 = int64_t(blockIdx.x) * SCAN_TILE;
    long long s = 0;
    for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_TPB) {
        const int64_t k = b0 + i;
        if (k < n) s += in[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int i = 0; i < SCAN_TPB / 32; ++i) t += red[i];
        bsum[blockIdx.x] = t;
    }
}

// single block: exclusive scan of the block sums in place; total and overflow flag
__global__ void __launch_bounds__(1024)
    scan_block_prefix_kernel(long long* __restrict__ bsum, int64_t nb, int32_t* __restrict__ total, int32_t* __restrict__ overflow) {
    __shared__ long long part[1024];
    const int t = threadIdx.x;
    const int64_t chunk = (nb + 1023) / 1024;
    const int64_t b = t * chunk, e = min(nb, b + chunk);
    long long s = 0;
    for (int64_t i = b; i < e; ++i) s += bsum[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        long long run = 0;
        for (int i = 0; i < 1024; ++i) {
            const long long v = part[i];
            part[i] = run;
            run += v;
        }
        *total = int32_t(run);
        *overflow = run > 2147483647ll;
    }
    __syncthreads();
    long long run = part[t];
    for (int64_t i = b; i < e; ++i) {
        const long long v = bsum[i];
        bsum[i] = run;
        run += v;
    }
}

// per-tile exclusive scan + block prefix -> offs[0..n); row_ptr[r] = offs[r * nseg] (the start of row r)
__global__ void __launch_bounds__(SCAN_TPB)
    scan_apply_kernel(const int32_t* __restrict__ in, int64_t n, const long long* __restrict__ bpre, int nseg,
                      int32_t* __restrict__ offs, int32_t* __restrict__ row_ptr) {
    __shared__ long long wsum[SCAN_TPB / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t k0 = int64_t(blockIdx.x) * SCAN_TILE + int64_t(threadIdx.x) * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (k0 + i < n) ? in[k0 + i] : 0;
        s += v[i];
    }
    long long incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    long long run = bpre[blockIdx.x] + incl - s;
    for (int i = 0; i < warp; ++i) run += wsum[i];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const int64_t k = k0 + i;
        if (k < n) {
            if (offs) offs[k] = int32_t(run);
            if (k % nseg == 0) row_ptr[k / nseg] = int32_t(run);
        }
        run += v[i];
    }
}

// ---- 3a. fill by mask replay (sparse results) -------------------------------------------------------------------
// One warp per row; lane l takes mask word l, l + 32, ...; a warp prefix over the popcounts gives every kept entry its
// slot, so col_idx / values of a row are written in ascending column order and the stores of a warp are contiguous.
// Up to four set bits of a word are taken per round so that their gathers are in flight together (the loads are
// unconditional on a valid dummy column: a predicated load would make every load blocking).
constexpr int FILL_TAKE = 4;

__global__ void __launch_bounds__(CSR_TPB)
    csr_fill_kernel(const double* __restrict__ S, int64_t rows, int64_t ld, const uint32_t* __restrict__ keep_mask,
                    int64_t mask_words, const int32_t* __restrict__ row_ptr, int32_t* __restrict__ col_idx,
                    double* __restrict__ values) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const uint32_t* m = keep_mask + row * mask_words;
    const double* base = S + row;
    int32_t run = row_ptr[row];
    for (int64_t wb = 0; wb < mask_words; wb += 32) {
        const int64_t wd = wb + lane;
        uint32_t bits = wd < mask_words ? __ldg(m + wd) : 0u;
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        int32_t pos = run + incl - cnt;
        const int32_t c0 = int32_t(wd << 5);
        while (bits) {
            int32_t c[FILL_TAKE];
            bool ok[FILL_TAKE];
#pragma unroll
            for (int u = 0; u < FILL_TAKE; ++u) {
                ok[u] = bits != 0;
                c[u] = c0 + (ok[u] ? __ffs(bits) - 1 : 0);  // column c0 exists whenever the word has a bit
                bits &= bits - 1;                            // 0 stays 0
            }
            if (values) {
                double v[FILL_TAKE];
#pragma unroll
                for (int u = 0; u < FILL_TAKE; ++u) v[u] = __ldg(base + int64_t(c[u]) * ld);  // weighted: the similarity itself
#pragma unroll
                for (int u = 0; u < FILL_TAKE; ++u)
                    if (ok[u]) values[pos + u] = v[u];
            }
#pragma unroll
            for (int u = 0; u < FILL_TAKE; ++u)
                if (ok[u]) col_idx[pos + u] = c[u];
            pos += int(ok[0]) + int(ok[1]) + int(ok[2]) + int(ok[3]);
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// ---- 3b. transposing tiled fill (dense results) ------------------------------------------------------------------
// A block owns 32 consecutive rows and sweeps all columns in tiles of 64: the tile is loaded with lanes along rows,
// transposed through padded shared memory, and compacted with lanes along columns: __ballot_sync + popc prefix give
// every kept entry its slot, so the col_idx / value stores of one row are contiguous.
__global__ void __launch_bounds__(CSR_TPB)
    csr_tiled_fill_kernel(const double* __restrict__ S, int64_t rows, int64_t cols, int64_t ld, double alpha, int weighted,
                          const int32_t* __restrict__ row_ptr, int32_t* __restrict__ col_idx, double* __restrict__ values) {
    __shared__ double tile[TILE_COLS][ROWS_PER_BLOCK + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = int64_t(blockIdx.x) * ROWS_PER_BLOCK;
    const bool w = weighted != 0;
    // warp `warp` compacts rows 4*warp .. 4*warp+3 of the block
    int32_t cursor[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t r = row0 + 4 * warp + q;
        cursor[q] = (r >= rows) ? 0 : row_ptr[r];
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
                if (keep) {
                    const int32_t pos = cursor[q] + __popc(ballot & ((1u << lane) - 1u));
                    col_idx[pos] = int32_t(c0 + 32 * h + lane);
                    if (values) values[pos] = x;
                }
                cursor[q] += __popc(ballot);
            }
        }
        __syncthreads();
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
    if (rows == 0 || cols == 0) {  // no cells: every row is empty
        cudaMemsetAsync(c->row_ptr, 0, size_t(rows + 1) * 4, ctx->stream);
        SS_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->col_idx), 4, ctx->stream));
        if (weighted) SS_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->values), 8, ctx->stream));
        c->nnz = 0;
        *out = c;
        return SS_OK;
    }
    // decomposition: 32-row blocks x column segments of seg_words mask words; short segments for small matrices so
    // that the grid still covers the GPU, at most 32 words (1024 columns) per segment
    const int64_t mask_words = ceil_div(cols, 32);
    const int64_t row_blocks = ceil_div(rows, 32);
    int seg_words = 32;
    while (seg_words > 4 && row_blocks * ceil_div(mask_words, seg_words) < 8 * 64 * int64_t(ctx->sm_count) &&
           ceil_div(mask_words, seg_words / 2) <= 65535)
        seg_words >>= 1;
    SS_REQUIRE(ceil_div(mask_words, seg_words) <= 65535, "featurize_csr: more than 67 M columns are not supported");
    const int nseg = int(ceil_div(mask_words, seg_words));
    const int64_t n = rows * nseg;  // (row, segment) counts, row-major = CSR order
    const int64_t nb = ceil_div(n, SCAN_TILE);
    void* p;
    if ((status = scratch_get(ctx, 8, size_t(n) * 4 + size_t(nb) * 8 + 64, &p)) != SS_OK) return fail(status);
    int32_t* seg_count = static_cast<int32_t*>(p);
    long long* bsum = reinterpret_cast<long long*>(reinterpret_cast<char*>(p) + ((size_t(n) * 4 + 15) & ~size_t(15)));
    int32_t* overflow = reinterpret_cast<int32_t*>(bsum + nb);
    void* mp;
    if ((status = scratch_get(ctx, 22, size_t(rows) * size_t(mask_words) * 4, &mp)) != SS_OK) return fail(status);
    uint32_t* keep_mask = static_cast<uint32_t*>(mp);
    const dim3 grid(unsigned(ceil_div(row_blocks, CSR_TPB / 32)), unsigned(nseg));
    const size_t tr_bytes = size_t(CSR_TPB / 32) * 32 * (seg_words + 1) * 4;
    csr_count_kernel<<<grid, CSR_TPB, tr_bytes, ctx->stream>>>(S->d, rows, cols, S->ld, alpha, weighted ? 1 : 0, mask_words, seg_words,
                                                                nseg, keep_mask, seg_count);
    scan_block_sums_kernel<<<unsigned(nb), SCAN_TPB, 0, ctx->stream>>>(seg_count, n, bsum);
    scan_block_prefix_kernel<<<1, 1024, 0, ctx->stream>>>(bsum, nb, c->row_ptr + rows, overflow);
    scan_apply_kernel<<<unsigned(nb), SCAN_TPB, 0, ctx->stream>>>(seg_count, n, bsum, nseg, nullptr, c->row_ptr);
    ctx->launches += 4;
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
        // 32-byte sector); above it the transposing tiles are the better access pattern.  SS_CSR_FILL=tiled / replay
        // force one form (A/B runs, tests)
        bool replay = double(c->nnz) < 0.10 * double(rows) * double(cols);
        if (const char* env = getenv("SS_CSR_FILL")) {
            if (!strcmp(env, "tiled")) replay = false;
            if (!strcmp(env, "replay")) replay = true;
        }
        if (replay) {
            csr_fill_kernel<<<unsigned(ceil_div(rows * 32, CSR_TPB)), CSR_TPB, 0, ctx->stream>>>(S->d, rows, S->ld, keep_mask, mask_words,
                                                                                             c->row_ptr, c->col_idx, c->values);
        } else {
            csr_tiled_fill_kernel<<<unsigned(row_blocks), CSR_TPB, 0, ctx->stream>>>(S->d, rows, cols, S->ld, alpha, weighted ? 1 : 0,
                                                                                     c->row_ptr, c->col_idx, c->values);
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
