// Deterministic, atomic-free form of the sparse 2-layer NBI with fused top-L (BASELINE config 5;
// reference `predict(A, ytrain)`, src/core.jl:446-466, on the graph [0 Y; Y' 0], ranked as
// `sortperm(rev=true)[1:L]`, src/performance.jl:315).
//
// The reference computes F = A * (W * W): first the square of the transfer matrix, then its product
// with the adjacency rows.  On the source rows that is
//
//     U[t',t] = sum_{s' asc} fl(Y[s',t'] / kt[t']) * fl(Y[s',t] / ks[s'])      (item x item block of W*W)
//     F[s,t]  = sum_{t' asc} Y[s,t'] * U[t',t]
//
// and this file evaluates exactly that association, every sum in ascending index order with separately
// rounded multiply and add (what a CSR x CSR Gustavson product with sorted rows does, e.g.
// scipy.sparse), so scores -- and therefore the order of tied scores -- are reproducible bit for bit:
//
//   tr_fill_kernel  : one block per item row t': the two-hop expansion t' -> co-raters s' -> their items t
//                     marks a shared-memory bitmap (a second bitmap holds the columns reached more than
//                     once); a popcount prefix gives every column its rank, so U[t',:] is written SORTED by
//                     column, split by column tile (16-bit in-tile column + FP64 value): a column reached
//                     through one co-rater is written directly, the others are collected in shared memory
//                     and summed in ascending co-rater order
//   tr_stream_kernel: one WARP owns a source s and walks the column tiles (TW columns: TW x 8 B of FP64
//                     accumulators in SHARED memory per warp): the tile segments of the rows U[t',:],
//                     t' in Y[s,:] ascending, are streamed from HBM (contiguous runs, register-prefetched;
//                     consecutive tiles continue the same row streams) and added with plain LDS/DADD/STS --
//                     columns inside a segment are distinct and segments follow one another in program
//                     order, so there are no atomics, no block barriers, and the order of additions is
//                     fixed; each tile is then scanned once against the running L-th best score
//                     (warp-distributed sorted list in registers) and cleared
//   tr_merge_kernel : few sources are split over ranges of tiles to fill the GPU; their lists (and the
//                     running result of earlier tile chunks) are merged under (score desc, column asc)
//
// Work unit: one partial product Y[s,t'] * U[t',t]; algorithmic bytes per partial product: 10 B (2 B column
// + 8 B value of U, read once from HBM).  U is 5e9 entries (50 GB) at 2M x 500k x 1e-4.  Rows of U are
// allocated by an upper bound (the number of two-hop paths) when that fits the free device memory; otherwise
// tr_count_kernel sizes every (row, tile) exactly and the column tiles are processed in chunks (U is built for
// one chunk of tiles at a time and the running top-L is merged across chunks).
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "ss_common.cuh"

struct ss_transfer {
    ss_ctx* ctx = nullptr;
    int64_t ns = 0, nt = 0;
    int tw = 0;          // tile width (columns), multiple of 64, <= 65536
    int ntiles = 0;      // ceil(nt / tw)
    int range_tiles = 0; // tiles covered by one bitmap pass of the build kernels
    bool weighted = false;
    bool counted = false;                      // `pre` / tile totals come from tr_count_kernel
    uint32_t* pre = nullptr;                   // [nt][ntiles + 1]: distinct columns of row t' before tile c
    unsigned long long* tile_total = nullptr;  // [ntiles + 1] device ([ntiles]: entries written by the fill)
    std::vector<unsigned long long> tile_total_host;
    // materialised chunk of tiles
    int tile_begin = 0, tile_end = 0;
    int64_t* rowbase = nullptr;  // [nt + 1]: entries of (t', tile c) start at rowbase[t'] + pre[t'][c]
    uint16_t* col = nullptr;
    double* val = nullptr;
    int64_t nnz = 0;  // entries of the materialised tiles
    int64_t cap = 0;  // allocated entries (>= nnz: rows are sized by an upper bound in the one-chunk build)
};

namespace sstr {

__device__ __forceinline__ uint64_t tr_key(double v) {  // order-preserving image of a double under isless
    if (v != v) return 0xFFFFFFFFFFFFFFFFull;
    const uint64_t b = uint64_t(__double_as_longlong(v));
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double tr_value(uint64_t k) {
    if (k == 0xFFFFFFFFFFFFFFFFull) return __longlong_as_double(0x7ff8000000000000ll);
    const uint64_t b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

struct TrGraph {
    const int32_t* y_ptr;   // CSR of Y (sources x targets), rows sorted by column
    const int32_t* y_idx;
    const double* y_val;    // null: binary
    const int32_t* yt_ptr;  // CSR of Y' (targets x sources), rows sorted by source
    const int32_t* yt_idx;
    const double* yt_val;
    int64_t ns, nt;
};

struct TrBuild {
    TrGraph g;
    int tw, ntiles, range_tiles;
    uint32_t* pre;
    unsigned long long* tile_total;
    int tile_begin, tile_end;  // fill: tiles to materialise
    const int64_t* rowbase;
    uint16_t* col;
    double* val;
    int* row_counter;
};

constexpr int TB_THREADS = 1024;
constexpr int TB_MAX_RANGE_TILES = 2048;

// f(t, u, e, ks): every product of row t' = [u0, u1) of Y' whose column t lies in [c_lo, c_hi): u = position of the
// co-rater s' in Y'[t',:], e = position of t in Y[s',:], ks = degree of s'.  A warp takes 32 co-raters at a time
// (their row extents are fetched by the 32 lanes in parallel), then the lanes run over each co-rater's targets.
template <class F>
__device__ __forceinline__ void tr_products(const TrGraph& g, int32_t u0, int32_t u1, int32_t c_lo, int32_t c_hi, F&& f) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int32_t ub = u0 + warp * 32; ub < u1; ub += nwarps * 32) {
        const int32_t u = ub + lane;
        int32_t r0 = 0, r1 = 0;
        if (u < u1) {
            const int32_t sp = __ldg(g.yt_idx + u);
            r0 = __ldg(g.y_ptr + sp);
            r1 = __ldg(g.y_ptr + sp + 1);
        }
        const int nb = min(32, u1 - ub);
        for (int l = 0; l < nb; l += 2) {
            const int32_t a0 = __shfl_sync(0xffffffffu, r0, l), a1 = __shfl_sync(0xffffffffu, r1, l);
            int32_t b0 = __shfl_sync(0xffffffffu, r0, (l + 1) & 31), b1 = __shfl_sync(0xffffffffu, r1, (l + 1) & 31);
            if (l + 1 >= nb) b1 = b0;
            const int32_t ea = a0 + lane, eb = b0 + lane;
            const int32_t ta = ea < a1 ? __ldg(g.y_idx + ea) : -1;  // both rows' first 32 targets in flight
            const int32_t tb = eb < b1 ? __ldg(g.y_idx + eb) : -1;
            if (ta >= c_lo && ta < c_hi) f(ta, ub + l, ea, a1 - a0);
            for (int32_t e = ea + 32; e < a1; e += 32) {
                const int32_t t = __ldg(g.y_idx + e);
                if (t >= c_lo && t < c_hi) f(t, ub + l, e, a1 - a0);
            }
            if (tb >= c_lo && tb < c_hi) f(tb, ub + l + 1, eb, b1 - b0);
            for (int32_t e = eb + 32; e < b1; e += 32) {
                const int32_t t = __ldg(g.y_idx + e);
                if (t >= c_lo && t < c_hi) f(t, ub + l + 1, e, b1 - b0);
            }
        }
    }
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// distinct columns of U[t',:] per column tile -> pre[t'][c + 1] (inclusive prefix over tiles), tile totals
__global__ void __launch_bounds__(TB_THREADS) tr_count_kernel(const TrBuild p) {
    extern __shared__ __align__(16) uint32_t tb_sm[];
    const int wpt = p.tw >> 5;  // bitmap words per tile
    uint32_t* A = tb_sm;
    unsigned long long* s_tot = reinterpret_cast<unsigned long long*>(A + size_t(p.range_tiles) * wpt);
    __shared__ uint32_t s_tcnt[TB_MAX_RANGE_TILES];
    __shared__ int s_row;
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (int t = tid; t < p.ntiles; t += blockDim.x) s_tot[t] = 0;
    const int64_t stride = int64_t(p.ntiles) + 1;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_row = atomicAdd(p.row_counter, 1);
        __syncthreads();
        const int row = s_row;
        if (row >= p.g.nt) break;
        const int32_t u0 = __ldg(p.g.yt_ptr + row), u1 = __ldg(p.g.yt_ptr + row + 1);
        uint32_t carry = 0;
        if (tid == 0) p.pre[row * stride] = 0;
        for (int tb = 0; tb < p.ntiles; tb += p.range_tiles) {
            const int ntr = min(p.range_tiles, p.ntiles - tb);
            const int32_t c_lo = tb * p.tw;
            const int32_t c_hi = int32_t(min(p.g.nt, int64_t(tb + ntr) * p.tw));
            const int nw = ntr * wpt;
            for (int w = tid; w < nw; w += blockDim.x) A[w] = 0;
            __syncthreads();
            tr_products(p.g, u0, u1, c_lo, c_hi, [&](int32_t t, int32_t, int32_t, int32_t) {
                const int32_t rel = t - c_lo;
                atomicOr(&A[rel >> 5], 1u << (rel & 31));
            });
            __syncthreads();
            for (int ti = warp; ti < ntr; ti += nwarps) {
                uint32_t sum = 0;
                for (int w = lane; w < wpt; w += 32) sum += __popc(A[ti * wpt + w]);
                sum = warp_sum(sum);
                if (lane == 0) s_tcnt[ti] = sum;
            }
            __syncthreads();
            if (warp == 0) {
                uint32_t run = carry;
                for (int b = 0; b < ntr; b += 32) {
                    const uint32_t v = (b + lane < ntr) ? s_tcnt[b + lane] : 0;
                    uint32_t incl = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += x;
                    }
                    if (b + lane < ntr) {
                        p.pre[row * stride + tb + b + lane + 1] = run + incl;
                        s_tot[tb + b + lane] += v;
                    }
                    run += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane == 0) s_carry = run;
            }
            __syncthreads();
            carry = s_carry;
        }
    }
    __syncthreads();
    for (int t = tid; t < p.ntiles; t += blockDim.x)
        if (s_tot[t]) atomicAdd(p.tile_total + t, s_tot[t]);
}

// sum over sources of ks^2 = two-hop paths source -> item -> ... = upper bound of the entries of U (picks the tile width)
__global__ void __launch_bounds__(256) tr_paths_kernel(const int32_t* __restrict__ y_ptr, int64_t ns, unsigned long long* __restrict__ out) {
    unsigned long long sum = 0;
    for (int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; s < ns; s += int64_t(gridDim.x) * blockDim.x) {
        const unsigned long long k = (unsigned long long)(__ldg(y_ptr + s + 1) - __ldg(y_ptr + s));
        sum += k * k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(out, sum);
}

// len[t'] for the row allocation: bound = number of two-hop paths through t' (capped by the columns of the chunk),
// or, after the count pass, the exact entries of tiles [tb, te).  One warp per row.
__global__ void __launch_bounds__(256) tr_rowlen_kernel(const TrGraph g, const uint32_t* __restrict__ pre, int ntiles, int tb,
                                                        int te, int64_t chunk_cols, int64_t* __restrict__ len) {
    const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= g.nt) return;
    if (pre) {
        if (lane == 0) len[row] = int64_t(pre[row * (int64_t(ntiles) + 1) + te] - pre[row * (int64_t(ntiles) + 1) + tb]);
        return;
    }
    const int32_t u0 = __ldg(g.yt_ptr + row), u1 = __ldg(g.yt_ptr + row + 1);
    long long sum = 0;
    for (int32_t u = u0 + lane; u < u1; u += 32) {
        const int32_t sp = __ldg(g.yt_idx + u);
        sum += __ldg(g.y_ptr + sp + 1) - __ldg(g.y_ptr + sp);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) len[row] = min((long long)chunk_cols, sum);
}

// rowbase[r] = (exclusive scan of len) - pre[r][tb] (when `pre` is given); rowbase[n] = total.  One block; in place.
__global__ void __launch_bounds__(1024) tr_rowscan_kernel(int64_t* __restrict__ len_base, int64_t n, const uint32_t* __restrict__ pre,
                                                         int ntiles, int tb) {
    __shared__ long long s_sum[1024];
    const int tid = threadIdx.x;
    const int64_t per = (n + 1023) / 1024;
    const int64_t r0 = min(n, tid * per), r1 = min(n, r0 + per);
    long long sum = 0;
    for (int64_t r = r0; r < r1; ++r) sum += len_base[r];
    s_sum[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
        const long long v = tid >= o ? s_sum[tid - o] : 0;
        __syncthreads();
        s_sum[tid] += v;
        __syncthreads();
    }
    long long run = tid ? s_sum[tid - 1] : 0;
    for (int64_t r = r0; r < r1; ++r) {
        const long long l = len_base[r];
        len_base[r] = run - (pre ? (long long)pre[r * (int64_t(ntiles) + 1) + tb] : 0);
        run += l;
    }
    if (tid == 1023) len_base[n] = s_sum[1023];
}

// ordered sum of the products through the common raters of items t' (row [u0,u1) of Y') and t (row [v0,v1)):
// sum_{s' asc} fl(Y[s',t']/kt') * fl(Y[s',t]/ks[s']); one warp, every lane returns the sum
template <bool WEIGHTED>
__device__ double tr_dup_sum(const TrGraph& g, int32_t u0, int32_t u1, int32_t v0, int32_t v1) {
    const int lane = threadIdx.x & 31;
    const double ktp = double(u1 - u0);
    const bool iter_u = (u1 - u0) <= (v1 - v0);  // iterate the shorter list, search the longer (both ascending)
    const int32_t i0 = iter_u ? u0 : v0, i1 = iter_u ? u1 : v1;
    const int32_t j0 = iter_u ? v0 : u0, j1 = iter_u ? v1 : u1;
    double sum = 0.0;
    for (int32_t ib = i0; ib < i1; ib += 32) {
        const int32_t i = ib + lane;
        bool found = false;
        double pr = 0.0;
        if (i < i1) {
            const int32_t sp = __ldg(g.yt_idx + i);
            int32_t lo = j0, hi = j1;
            while (lo < hi) {
                const int32_t mid = (lo + hi) >> 1;
                if (__ldg(g.yt_idx + mid) < sp) lo = mid + 1; else hi = mid;
            }
            if (lo < j1 && __ldg(g.yt_idx + lo) == sp) {
                found = true;
                const int32_t up = iter_u ? i : lo, vp = iter_u ? lo : i;
                const double ks = double(__ldg(g.y_ptr + sp + 1) - __ldg(g.y_ptr + sp));
                const double w1 = (WEIGHTED ? __ldg(g.yt_val + up) : 1.0) / ktp;
                const double w2 = (WEIGHTED ? __ldg(g.yt_val + vp) : 1.0) / ks;
                pr = __dmul_rn(w1, w2);
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, found);
        while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            sum = __dadd_rn(sum, __shfl_sync(0xffffffffu, pr, l));
        }
    }
    return sum;
}

constexpr int TB_NDUP = 2048;  // products on multiply-reached columns kept in shared memory per bitmap range

// U[t', tiles [tile_begin, tile_end)]: sorted by column, split by tile.  WRITE_PRE: the one-chunk build -- the row's
// `pre` entries come out of the same bitmap (no count pass) and the entries written are added to tile_total[ntiles].
template <bool WEIGHTED, bool WRITE_PRE>
__global__ void __launch_bounds__(TB_THREADS) tr_fill_kernel(const TrBuild p) {
    extern __shared__ __align__(16) uint32_t tb_sm[];
    const int wpt = p.tw >> 5;
    const int maxw = p.range_tiles * wpt;
    uint32_t* A = tb_sm;          // column reached at least once
    uint32_t* B = A + maxw;       // column reached more than once
    uint32_t* P = B + maxw;       // exclusive popcount prefix of A
    double* d_val = reinterpret_cast<double*>(P + maxw);              // products on the columns of B
    uint64_t* d_key = reinterpret_cast<uint64_t*>(d_val + TB_NDUP);   // (column relative to the range, co-rater ordinal)
    __shared__ uint32_t s_part[TB_THREADS / 32];
    __shared__ int s_row, s_ndup;
    __shared__ uint32_t s_total;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const int64_t stride = int64_t(p.ntiles) + 1;
    unsigned long long written = 0;  // thread 0
    for (;;) {
        __syncthreads();
        if (tid == 0) s_row = atomicAdd(p.row_counter, 1);
        __syncthreads();
        const int row = s_row;
        if (row >= p.g.nt) break;
        const int32_t u0 = __ldg(p.g.yt_ptr + row), u1 = __ldg(p.g.yt_ptr + row + 1);
        if (u1 == u0) {
            if (WRITE_PRE)
                for (int c = tid; c <= p.ntiles; c += blockDim.x) p.pre[row * stride + c] = 0;
            continue;
        }
        const double ktp = double(u1 - u0);
        const int64_t rbase = p.rowbase[row];
        uint32_t carry = 0;  // WRITE_PRE: entries of the row before this range
        if (WRITE_PRE && tid == 0) p.pre[row * stride] = 0;
        for (int tb = p.tile_begin; tb < p.tile_end; tb += p.range_tiles) {
            const int ntr = min(p.range_tiles, p.tile_end - tb);
            if (!WRITE_PRE && p.pre[row * stride + tb + ntr] == p.pre[row * stride + tb]) continue;  // nothing in this range
            const int32_t c_lo = tb * p.tw;
            const int32_t c_hi = int32_t(min(p.g.nt, int64_t(tb + ntr) * p.tw));
            const int nw = ntr * wpt;
            __syncthreads();
            for (int w = tid; w < nw; w += blockDim.x) {
                A[w] = 0;
                B[w] = 0;
            }
            if (tid == 0) s_ndup = 0;
            __syncthreads();
            tr_products(p.g, u0, u1, c_lo, c_hi, [&](int32_t t, int32_t, int32_t, int32_t) {
                const int32_t rel = t - c_lo;
                const uint32_t bit = 1u << (rel & 31);
                const uint32_t old = atomicOr(&A[rel >> 5], bit);
                if (old & bit) atomicOr(&B[rel >> 5], bit);
            });
            __syncthreads();
            {   // block-wide exclusive scan of popc(A[w]): contiguous word runs per thread
                const int per = (nw + TB_THREADS - 1) / TB_THREADS;
                const int w0 = min(nw, tid * per), w1 = min(nw, w0 + per);
                uint32_t sum = 0;
                for (int w = w0; w < w1; ++w) sum += __popc(A[w]);
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += x;
                }
                if (lane == 31) s_part[warp] = incl;
                __syncthreads();
                if (warp == 0) {
                    uint32_t v = lane < nwarps ? s_part[lane] : 0;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t x = __shfl_up_sync(0xffffffffu, v, o);
                        if (lane >= o) v += x;
                    }
                    if (lane < nwarps) s_part[lane] = v;
                    if (lane == 31) s_total = v;
                }
                __syncthreads();
                uint32_t run = (warp ? s_part[warp - 1] : 0) + incl - sum;
                for (int w = w0; w < w1; ++w) {
                    P[w] = run;
                    run += __popc(A[w]);
                }
            }
            __syncthreads();
            if (WRITE_PRE) {
                for (int c = tid + 1; c < ntr; c += blockDim.x) p.pre[row * stride + tb + c] = carry + P[c * wpt];
                if (tid == 0) p.pre[row * stride + tb + ntr] = carry + s_total;
            }
            const int64_t base = rbase + int64_t(WRITE_PRE ? carry : p.pre[row * stride + tb]);
            tr_products(p.g, u0, u1, c_lo, c_hi, [&](int32_t t, int32_t u, int32_t e, int32_t ks) {
                const int32_t rel = t - c_lo;
                const int w = rel >> 5;
                const uint32_t bit = 1u << (rel & 31);
                const double w1 = (WEIGHTED ? __ldg(p.g.yt_val + u) : 1.0) / ktp;          // W[t',s'] (true division)
                const double w2 = (WEIGHTED ? __ldg(p.g.y_val + e) : 1.0) / double(ks);    // W[s',t]
                const double pr = __dmul_rn(w1, w2);
                if (B[w] & bit) {  // several co-raters reach t: summed in source order below
                    if (t == row) return;  // U[t',t']: every co-rater contributes, summed by warp 0 below
                    const int slot = atomicAdd(&s_ndup, 1);
                    if (slot < TB_NDUP) {
                        d_val[slot] = pr;
                        d_key[slot] = (uint64_t(uint32_t(rel)) << 32) | uint32_t(u);
                    }
                    return;
                }
                const int64_t pos = base + P[w] + __popc(A[w] & (bit - 1));
                p.col[pos] = uint16_t(t % p.tw);
                p.val[pos] = pr;
            });
            if (warp == 0 && row >= c_lo && row < c_hi) {  // the diagonal entry, when more than one source has item t'
                const int32_t rel = row - c_lo;
                const int w = rel >> 5;
                const uint32_t bit = 1u << (rel & 31);
                if (B[w] & bit) {
                    double sum = 0.0;
                    for (int32_t ub = u0; ub < u1; ub += 32) {
                        const int32_t u = ub + lane;
                        double pr = 0.0;
                        if (u < u1) {
                            const int32_t sp = __ldg(p.g.yt_idx + u);
                            const double ks = double(__ldg(p.g.y_ptr + sp + 1) - __ldg(p.g.y_ptr + sp));
                            const double yv = WEIGHTED ? __ldg(p.g.yt_val + u) : 1.0;
                            pr = __dmul_rn(yv / ktp, yv / ks);
                        }
                        const int nb = min(32, u1 - ub);
                        for (int l = 0; l < nb; ++l) sum = __dadd_rn(sum, __shfl_sync(0xffffffffu, pr, l));
                    }
                    if (lane == 0) {
                        const int64_t pos = base + P[w] + __popc(A[w] & (bit - 1));
                        p.col[pos] = uint16_t(row % p.tw);
                        p.val[pos] = sum;
                    }
                }
            }
            __syncthreads();
            const int nd = s_ndup;
            if (nd <= TB_NDUP) {
                // sort the products by (column, co-rater ordinal): rank by counting, permuted in place through registers
                uint64_t mk[TB_NDUP / TB_THREADS];
                double mv[TB_NDUP / TB_THREADS];
                int mr[TB_NDUP / TB_THREADS];
#pragma unroll
                for (int q = 0; q < TB_NDUP / TB_THREADS; ++q) {
                    const int i = tid + q * TB_THREADS;
                    mr[q] = -1;
                    if (i < nd) {
                        mk[q] = d_key[i];
                        mv[q] = d_val[i];
                        int rank = 0;
                        for (int j = 0; j < nd; ++j) rank += d_key[j] < mk[q] ? 1 : 0;
                        mr[q] = rank;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < TB_NDUP / TB_THREADS; ++q)
                    if (mr[q] >= 0) {
                        d_key[mr[q]] = mk[q];
                        d_val[mr[q]] = mv[q];
                    }
                __syncthreads();
                for (int q = tid; q < nd; q += blockDim.x) {  // the first product of a column adds up the column in order
                    const uint32_t rel = uint32_t(d_key[q] >> 32);
                    if (q > 0 && uint32_t(d_key[q - 1] >> 32) == rel) continue;
                    double sum = d_val[q];  // 0 + x = x
                    for (int e = q + 1; e < nd && uint32_t(d_key[e] >> 32) == rel; ++e) sum = __dadd_rn(sum, d_val[e]);
                    const int w = int(rel >> 5);
                    const uint32_t bit = 1u << (rel & 31);
                    const int64_t pos = base + P[w] + __popc(A[w] & (bit - 1));
                    p.col[pos] = uint16_t((c_lo + int32_t(rel)) % p.tw);
                    p.val[pos] = sum;
                }
            } else {
                // more multiply-reached columns than the list holds (very popular items): every such column is summed
                // from the intersection of the two rater lists instead
                for (int w = warp; w < nw; w += nwarps) {
                    uint32_t bits = B[w];
                    while (bits) {
                        const int b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const int32_t t = c_lo + w * 32 + b;
                        if (t == row) continue;
                        const double sum = tr_dup_sum<WEIGHTED>(p.g, u0, u1, __ldg(p.g.yt_ptr + t), __ldg(p.g.yt_ptr + t + 1));
                        if (lane == 0) {
                            const int64_t pos = base + P[w] + __popc(A[w] & ((1u << b) - 1));
                            p.col[pos] = uint16_t(t % p.tw);
                            p.val[pos] = sum;
                        }
                    }
                }
            }
            if (WRITE_PRE) {
                carry += s_total;
                if (tid == 0) written += s_total;
            }
        }
    }
    if (WRITE_PRE && tid == 0 && written) atomicAdd(p.tile_total + p.ntiles, written);
}

// ------------------------------------------------------------------------------------------------
// streaming accumulation + running top-L: one warp owns (source s, a range of column tiles)
// ------------------------------------------------------------------------------------------------
struct TrStream {
    const int32_t* y_ptr;
    const int32_t* y_idx;
    const double* y_val;
    const uint32_t* pre;
    const int64_t* rowbase;
    const uint16_t* col;
    const double* val;
    int64_t nt;
    int ntiles, tile_begin, ntl;  // all tiles, first tile of the chunk, tiles in the chunk
    int64_t s0;                   // first source
    int parts;                    // tile ranges per source (work item = (source, part))
    int nwork;                    // sources * parts
    int L;
    uint64_t* cand_key;           // [nwork][L]
    int32_t* cand_col;
    int* counter;
};

constexpr int TS_D = 4;  // rows of U per register bank (two banks: one being consumed, one in flight)

// insert (ck, ci) into the warp-distributed list sorted by (key desc, column asc); lane l holds rank l
__device__ __forceinline__ void tr_list_insert(uint64_t& lkey, int32_t& lidx, int& cnt, int L, uint64_t ck, int32_t ci, int lane) {
    const bool before = (lane < cnt) && (lkey > ck || (lkey == ck && lidx < ci));
    const int pos = __popc(__ballot_sync(0xffffffffu, before));
    if (pos >= L) return;
    const uint64_t upk = __shfl_up_sync(0xffffffffu, lkey, 1);
    const int32_t upi = __shfl_up_sync(0xffffffffu, lidx, 1);
    if (lane == pos) {
        lkey = ck;
        lidx = ci;
    } else if (lane > pos) {
        lkey = upk;
        lidx = upi;
    }
    if (cnt < L) ++cnt;
}

// high word of tr_key(v) (not valid for NaN, which the caller tests separately)
__device__ __forceinline__ uint32_t tr_key_hi(double v) {
    const int hi = __double2hiint(v);
    return uint32_t(hi ^ ((hi >> 31) | int(0x80000000)));
}

// one bank of TS_D rows: lane l holds entries l, l + 32, ... (EPL slots) of each row
template <int EPL>
struct TrBank {
    uint32_t c[TS_D][EPL];
    double v[TS_D][EPL];
    int n[TS_D];
};

__device__ __forceinline__ const void* tr_shfl_ptr(const void* p, int src) {
    return reinterpret_cast<const void*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(p), src));
}

// pc / pv: lane r holds the pointers to the first column / value of the segment of row r, cn its length.
// The loads are UNCONDITIONAL (lanes beyond the segment re-read its last entry; an empty row reads the entry its pointer
// designates, which exists: the arrays carry slack).  A predicated load leaves the old value of its destination
// register live, and ptxas then merges old and new through predicated moves that wait for the load: every load
// became blocking (long-scoreboard stalls on IMAD.MOV in the round-2 profiles).
// a 16-bit column as an opaque 32-bit value: when the compiler knows the upper half is zero it packs two columns
// into one register with a PRMT placed right behind the load, which turns the load into a blocking one
__device__ __forceinline__ uint32_t tr_ldg_col(const uint16_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <int EPL>
__device__ __forceinline__ void tr_bank_load(TrBank<EPL>& b, const uint16_t* pc, const double* pv, int cn, int rb, int nrows,
                                             int lane) {
#pragma unroll
    for (int k = 0; k < TS_D; ++k) {
        const int r = rb + k;
        const uint16_t* pcr = static_cast<const uint16_t*>(tr_shfl_ptr(pc, r & 31));
        const double* pvr = static_cast<const double*>(tr_shfl_ptr(pv, r & 31));
        int n = __shfl_sync(0xffffffffu, cn, r & 31);
        if (r >= nrows) n = 0;
        b.n[k] = n;
        const int nm1 = max(n, 1) - 1;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const int i = min(lane + 32 * e, nm1);
            b.c[k][e] = tr_ldg_col(pcr + i);
            b.v[k][e] = __ldg(pvr + i);
        }
    }
}

template <int EPL, bool WEIGHTED>
__device__ __forceinline__ void tr_bank_consume(const TrBank<EPL>& b, double* __restrict__ acc, const uint16_t* pc, const double* pv,
                                                double cf, int rb, int nrows, int lane) {
#pragma unroll
    for (int k = 0; k < TS_D; ++k) {
        const int r = rb + k;
        if (r >= nrows) break;
        const int n = b.n[k];
        double cfr = 1.0;
        if (WEIGHTED) cfr = __shfl_sync(0xffffffffu, cf, r);
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            if (lane < n - 32 * e) {
                double* a = acc + b.c[k][e];
                *a = __dadd_rn(*a, WEIGHTED ? __dmul_rn(cfr, b.v[k][e]) : b.v[k][e]);
            }
        }
        if (n > 32 * EPL) {  // segment longer than the register slots (rare by the choice of the tile width)
            const uint16_t* pcr = static_cast<const uint16_t*>(tr_shfl_ptr(pc, r));
            const double* pvr = static_cast<const double*>(tr_shfl_ptr(pv, r));
            for (int i = lane + 32 * EPL; i < n; i += 32) {
                double* a = acc + pcr[i];
                const double v = pvr[i];
                *a = __dadd_rn(*a, WEIGHTED ? __dmul_rn(cfr, v) : v);
            }
        }
        __syncwarp();  // the next row may touch the same columns
    }
}

__device__ __forceinline__ bool tr_either_nan(double x, double y) {
    int r;
    asm("{ .reg .pred p; setp.nan.f64 p, %1, %2; selp.s32 %0, 1, 0, p; }" : "=r"(r) : "d"(x), "d"(y));
    return r != 0;
}

// Per warp: for every tile of its range, the segments U[t', tile] of the source's items t' (ascending) are added into
// the warp's shared-memory accumulators -- lane l takes entries l, l + 32, ... of a segment (columns inside a segment
// are distinct, segments follow one another in program order: no atomics, fixed order of additions) -- then the tile
// is scanned once: entries that beat the running L-th best enter the warp-distributed top-L list, and the tile is
// cleared.  TW = 1024 * EPL: a segment holds about 20 * EPL entries at the density of config 5.
// Latency: the segment table of a unit (tile, 32 items) is read two units ahead and the rows of the current unit go
// through three register banks of TS_D rows (two in flight while the third is consumed).  Measured and dropped
// (DESIGN.md 4.4): pulling the next unit into L2 with prefetch.global.L2 (no gain, +45 % DRAM reads: 128-byte lines
// against 32-byte sectors) and staging the segments in shared memory with cp.async.bulk + mbarrier (1.8 x slower: the
// per-row copies serialise on the uniform datapath and the 34 KB per warp leave 6 warps per SM).
constexpr int ts_warps(int epl) { return epl == 1 ? 16 : epl == 2 ? 12 : epl == 3 ? 8 : 6; }  // register-bound (three banks)

template <int EPL, bool WEIGHTED>
__global__ void __launch_bounds__(ts_warps(EPL) * 32, 1) tr_stream_kernel(const TrStream p) {
    constexpr int TW = 1024 * EPL;
    extern __shared__ __align__(16) double ts_acc[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* acc = ts_acc + warp * TW;
    const int L = p.L;
    for (int j = 0; j < TW / 64; ++j) reinterpret_cast<double2*>(acc)[j * 32 + lane] = make_double2(0.0, 0.0);
    __syncwarp();
    const int64_t pstride = int64_t(p.ntiles) + 1;

    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(p.counter, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= p.nwork) break;
        const int64_t s = p.s0 + w / p.parts;
        const int part = w % p.parts;
        const int tb = p.tile_begin + int(int64_t(part) * p.ntl / p.parts);
        int te = p.tile_begin + int(int64_t(part + 1) * p.ntl / p.parts);
        const int32_t b0 = __ldg(p.y_ptr + s), b1 = __ldg(p.y_ptr + s + 1);
        const int R = b1 - b0;
        if (R == 0 && te > tb) te = tb + 1;  // no items: every score is 0, the first tile of the range decides
        const int nchunks = (R + 31) >> 5;

        uint64_t lkey = 0;  // running top-L of this work item
        int32_t lidx = -1;
        int cnt = 0;

        // metadata of a unit (tile, chunk of 32 items): lane r holds the segment of row r.  The table entries of a unit
        // are loaded two units ahead as raw values (m2), turned into pointers one unit ahead (m1: nothing waits on a
        // load that was just issued) and used in the current unit (pc, pv, cn, cf).
        struct Raw {
            long long rb;
            uint32_t p0, p1;
            double cf;
        };
        Raw m2{0, 0, 0, 1.0};
        const uint16_t *pc = p.col, *pc_1 = p.col;
        const double *pv = p.val, *pv_1 = p.val;
        int cn = 0, cn_1 = 0;
        double cf = 1.0, cf_1 = 1.0;
        int mt = tb, mc = 0;  // unit whose table entries are loaded next
        auto load_raw = [&](Raw& o) {
            const int row = (mc << 5) + lane;
            o.rb = 0;
            o.p0 = o.p1 = 0;
            o.cf = 1.0;
            if (row < R && mt < te) {
                const int32_t tp = __ldg(p.y_idx + b0 + row);
                const uint32_t* pr = p.pre + int64_t(tp) * pstride + mt;
                o.p0 = __ldg(pr);
                o.p1 = __ldg(pr + 1);
                o.rb = __ldg(p.rowbase + tp);
                if (WEIGHTED) o.cf = __ldg(p.y_val + b0 + row);
            }
            if (++mc >= nchunks) {
                mc = 0;
                ++mt;
            }
        };
        auto to_pointers = [&](const Raw& m, const uint16_t*& o_pc, const double*& o_pv, int& o_cn, double& o_cf) {
            const long long st = m.rb + (long long)m.p0;
            o_pc = p.col + st;
            o_pv = p.val + st;
            o_cn = int(m.p1 - m.p0);
            o_cf = m.cf;
        };
        if (nchunks > 0) {
            load_raw(m2);
            to_pointers(m2, pc_1, pv_1, cn_1, cf_1);  // the first unit waits for its table entries
            load_raw(m2);
        }

        for (int tile = tb; tile < te; ++tile) {
            for (int chunk = 0; chunk < nchunks; ++chunk) {
                pc = pc_1, pv = pv_1, cn = cn_1, cf = cf_1;
                to_pointers(m2, pc_1, pv_1, cn_1, cf_1);  // loaded during the previous unit
                load_raw(m2);                              // two units ahead
                const int nrows = min(32, R - (chunk << 5));
                TrBank<EPL> ba, bb, bc;  // two banks in flight while the third is consumed
                tr_bank_load<EPL>(ba, pc, pv, cn, 0, nrows, lane);
                tr_bank_load<EPL>(bb, pc, pv, cn, TS_D, nrows, lane);
                for (int rb = 0; rb < nrows; rb += 3 * TS_D) {
                    tr_bank_load<EPL>(bc, pc, pv, cn, rb + 2 * TS_D, nrows, lane);
                    tr_bank_consume<EPL, WEIGHTED>(ba, acc, pc, pv, cf, rb, nrows, lane);
                    tr_bank_load<EPL>(ba, pc, pv, cn, rb + 3 * TS_D, nrows, lane);
                    tr_bank_consume<EPL, WEIGHTED>(bb, acc, pc, pv, cf, rb + TS_D, nrows, lane);
                    tr_bank_load<EPL>(bb, pc, pv, cn, rb + 4 * TS_D, nrows, lane);
                    tr_bank_consume<EPL, WEIGHTED>(bc, acc, pc, pv, cf, rb + 2 * TS_D, nrows, lane);
                }
            }
            // ---- scan + clear the tile: entries above the running L-th best enter the list.  Columns are visited in
            // ascending order, so an entry that only ties the L-th best never displaces it.
            const int tile_cols = int(min(int64_t(TW), p.nt - int64_t(tile) * TW));
            const int32_t colbase = tile * TW;
            uint64_t thr = cnt >= L ? __shfl_sync(0xffffffffu, lkey, L - 1) : 0;
            uint32_t thr_hi = uint32_t(thr >> 32);
            auto scan = [&](auto full) {
                constexpr bool FULL = decltype(full)::value;
#pragma unroll 4
                for (int j = 0; j < TW / 64; ++j) {
                    const int cc = j * 64 + 2 * lane;
                    double2* ap = reinterpret_cast<double2*>(acc + cc);
                    const double2 v = *ap;
                    *ap = make_double2(0.0, 0.0);
                    bool pass = max(tr_key_hi(v.x), tr_key_hi(v.y)) >= thr_hi || tr_either_nan(v.x, v.y);
                    if (!FULL) pass = pass && cc < tile_cols;
                    unsigned any = __ballot_sync(0xffffffffu, pass);
                    if (any) {
                        const uint64_t kx = tr_key(v.x), ky = tr_key(v.y);
                        while (any) {
                            const int src = __ffs(any) - 1;
                            any &= any - 1;
                            const uint64_t ckx = __shfl_sync(0xffffffffu, kx, src), cky = __shfl_sync(0xffffffffu, ky, src);
                            const int32_t c0 = j * 64 + 2 * src;
                            if (cnt < L || ckx > thr) {
                                tr_list_insert(lkey, lidx, cnt, L, ckx, colbase + c0, lane);
                                if (cnt >= L) thr = __shfl_sync(0xffffffffu, lkey, L - 1);
                            }
                            if ((FULL || c0 + 1 < tile_cols) && (cnt < L || cky > thr)) {
                                tr_list_insert(lkey, lidx, cnt, L, cky, colbase + c0 + 1, lane);
                                if (cnt >= L) thr = __shfl_sync(0xffffffffu, lkey, L - 1);
                            }
                        }
                        thr_hi = uint32_t(thr >> 32);
                    }
                }
            };
            if (tile_cols == TW) scan(std::true_type{}); else scan(std::false_type{});
            __syncwarp();
        }
        if (lane < L) {
            p.cand_key[int64_t(w) * L + lane] = lane < cnt ? lkey : 0;
            p.cand_col[int64_t(w) * L + lane] = lane < cnt ? lidx : -1;
        }
    }
}

// one warp per source: tile candidates (+ the running top-L of earlier tile chunks) -> top-L
__global__ void __launch_bounds__(256) tr_merge_kernel(const uint64_t* __restrict__ cand_key, const int32_t* __restrict__ cand_col,
                                                       int ntl, int L, int64_t s0, int nusers, bool have_prev,
                                                       uint64_t* __restrict__ run_key, int32_t* __restrict__ idx_out,
                                                       double* __restrict__ val_out, int64_t ldv) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ul = blockIdx.x * 8 + warp;
    if (ul >= nusers) return;
    const int64_t s = s0 + ul;
    uint64_t lkey = 0;
    int32_t lidx = -1;
    int cnt = 0;
    if (have_prev) {
        if (lane < L) {
            lkey = run_key[s * L + lane];
            lidx = idx_out[s * L + lane];
        }
        cnt = __popc(__ballot_sync(0xffffffffu, lane < L && lidx >= 0));
    }
    for (int tile = 0; tile < ntl; ++tile) {
        const int64_t o = (int64_t(ul) * ntl + tile) * L + lane;
        const uint64_t key = lane < L ? cand_key[o] : 0;
        const int32_t colv = lane < L ? cand_col[o] : -1;
        const uint64_t thr = __shfl_sync(0xffffffffu, lkey, L - 1);
        const int32_t thc = __shfl_sync(0xffffffffu, lidx, L - 1);
        unsigned cd = __ballot_sync(0xffffffffu, colv >= 0 && (cnt < L || key > thr || (key == thr && colv < thc)));
        while (cd) {
            const int src = __ffs(cd) - 1;
            cd &= cd - 1;
            tr_list_insert(lkey, lidx, cnt, L, __shfl_sync(0xffffffffu, key, src), __shfl_sync(0xffffffffu, colv, src), lane);
        }
    }
    if (lane < L) {
        idx_out[s * L + lane] = (lane < cnt) ? lidx : -1;
        if (val_out) val_out[s * ldv + lane] = (lane < cnt) ? tr_value(lkey) : 0.0;
        if (run_key) run_key[s * L + lane] = (lane < cnt) ? lkey : 0;
    }
}

int tr_env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

TrGraph make_graph(const ss_csr* Y, const ss_csr* YT) {
    TrGraph g{};
    g.y_ptr = Y->row_ptr;
    g.y_idx = Y->col_idx;
    g.y_val = Y->values;
    g.yt_ptr = YT->row_ptr;
    g.yt_idx = YT->col_idx;
    g.yt_val = YT->values;
    g.ns = Y->rows;
    g.nt = Y->cols;
    return g;
}

constexpr size_t TB_BITMAP_SMEM = 188 * 1024;                 // two bitmaps + prefix of the fill kernel
constexpr size_t TB_DUP_SMEM = size_t(TB_NDUP) * 16;          // products on multiply-reached columns

}  // namespace sstr

using namespace sstr;

namespace ss {

void transfer_free_chunk(ss_transfer* T) {
    if (T->rowbase) cudaFree(T->rowbase);
    if (T->col) cudaFree(T->col);
    if (T->val) cudaFree(T->val);
    T->rowbase = nullptr;
    T->col = nullptr;
    T->val = nullptr;
    T->nnz = T->cap = 0;
    T->tile_begin = T->tile_end = 0;
}

void transfer_free(ss_transfer* T) {
    if (!T) return;
    cudaSetDevice(T->ctx->device);
    cudaStreamSynchronize(T->ctx->stream);
    transfer_free_chunk(T);
    if (T->pre) cudaFree(T->pre);
    if (T->tile_total) cudaFree(T->tile_total);
    delete T;
}

// Tile width of the streaming kernel: SS_RECSYS_TILE, else the width at which a row of U holds about 40 entries per
// tile (two register slots per lane are then filled 63 %, and a segment rarely exceeds them): the entries of a row
// of U are bounded by the two-hop paths, sum_s ks^2 / targets.
int32_t transfer_pick_tile(ss_ctx* ctx, const ss_csr* Y, int* tw, unsigned long long* paths_out) {
    const int t = tr_env_int("SS_RECSYS_TILE", 0);
    void* w;
    SS_TRY(scratch_get(ctx, 16, 256, &w));
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(static_cast<char*>(w) + 128);
    SS_CHECK_CUDA(cudaMemsetAsync(acc, 0, 8, ctx->stream));
    tr_paths_kernel<<<unsigned(std::min<int64_t>(1024, ceil_div(std::max<int64_t>(Y->rows, 1), 256))), 256, 0, ctx->stream>>>(Y->row_ptr, Y->rows, acc);
    ctx->launches += 1;
    unsigned long long paths = 0;
    SS_CHECK_CUDA(cudaMemcpyAsync(&paths, acc, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (paths_out) *paths_out = paths;
    const double per_row = double(paths) / double(std::max<int64_t>(Y->cols, 1));
    const double want = per_row > 0 ? 40.0 * double(Y->cols) / per_row : 4096.0;  // columns that hold ~40 entries of a row
    *tw = want < 1536 ? 1024 : want < 2560 ? 2048 : want < 3584 ? 3072 : 4096;
    if (t == 1024 || t == 2048 || t == 3072 || t == 4096) *tw = t;
    while (*tw < 4096 && ceil_div(Y->cols, *tw) > 8192) *tw += 1024;
    return SS_OK;
}

// the handle and its tile table (not yet filled)
int32_t transfer_create(ss_ctx* ctx, const ss_csr* Y, int tw, ss_transfer** out) {
    *out = nullptr;
    SS_REQUIRE(tw >= 64 && tw <= 65536 && tw % 64 == 0, "transfer matrix: tile width %d must be a multiple of 64, <= 65536", tw);
    const int64_t ntiles = ceil_div(Y->cols, tw);
    SS_REQUIRE(ntiles <= 8192, "transfer matrix: %lld column tiles (more than 8192): raise the tile width", (long long)ntiles);
    ss_transfer* T = new ss_transfer();
    T->ctx = ctx;
    T->ns = Y->rows;
    T->nt = Y->cols;
    T->tw = tw;
    T->ntiles = int(ntiles);
    T->weighted = Y->values != nullptr;
    const int max_words = int(TB_BITMAP_SMEM / 12);
    T->range_tiles = std::max(1, std::min({T->ntiles, max_words * 32 / tw, TB_MAX_RANGE_TILES}));
    cudaError_t e = cudaMalloc(&T->pre, size_t(T->nt) * (T->ntiles + 1) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&T->tile_total, size_t(T->ntiles + 1) * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemsetAsync(T->tile_total, 0, size_t(T->ntiles + 1) * sizeof(unsigned long long), ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("transfer matrix: cannot allocate the tile table (%s)", cudaGetErrorString(e));
        transfer_free(T);
        return SS_ERR_OOM;
    }
    *out = T;
    return SS_OK;
}

static int32_t row_counter(ss_ctx* ctx, int** out) {
    void* cnt;
    SS_TRY(scratch_get(ctx, 16, 256, &cnt));
    SS_CHECK_CUDA(cudaMemsetAsync(cnt, 0, 256, ctx->stream));
    *out = static_cast<int*>(cnt);
    return SS_OK;
}

// count pass (only when U has to be cut into chunks of tiles): the `pre` table and the entries of U per column tile
int32_t transfer_count(ss_ctx* ctx, const ss_csr* Y, const ss_csr* YT, ss_transfer* T) {
    TrBuild b{};
    b.g = make_graph(Y, YT);
    b.tw = T->tw;
    b.ntiles = T->ntiles;
    b.range_tiles = T->range_tiles;
    b.pre = T->pre;
    b.tile_total = T->tile_total;
    SS_TRY(row_counter(ctx, &b.row_counter));
    const size_t smem = size_t(T->range_tiles) * (T->tw / 32) * 4 + size_t(T->ntiles) * 8;
    SS_CHECK_CUDA(cudaFuncSetAttribute(tr_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const int per_sm = smem * 2 + 20 * 1024 <= 227 * 1024 ? 2 : 1;
    const int grid = int(std::min<int64_t>(T->nt, int64_t(ctx->sm_count) * per_sm));
    tr_count_kernel<<<grid, TB_THREADS, smem, ctx->stream>>>(b);
    ctx->launches += 1;
    T->tile_total_host.resize(T->ntiles);
    SS_CHECK_CUDA(cudaMemcpyAsync(T->tile_total_host.data(), T->tile_total, size_t(T->ntiles) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    T->counted = true;
    return SS_OK;
}

// row offsets of the tiles [tile_begin, tile_end): exact after the count pass, else by the two-hop path bound;
// *entries = entries to allocate
int32_t transfer_rows(ss_ctx* ctx, const ss_csr* Y, const ss_csr* YT, ss_transfer* T, int tile_begin, int tile_end, int64_t* entries) {
    transfer_free_chunk(T);
    SS_REQUIRE(tile_begin >= 0 && tile_begin < tile_end && tile_end <= T->ntiles, "transfer matrix: bad tile range");
    SS_CHECK_CUDA(cudaMalloc(&T->rowbase, size_t(T->nt + 1) * sizeof(int64_t)));
    const int64_t chunk_cols = std::min<int64_t>(T->nt, int64_t(tile_end) * T->tw) - int64_t(tile_begin) * T->tw;
    tr_rowlen_kernel<<<unsigned(ceil_div(T->nt * 32, 256)), 256, 0, ctx->stream>>>(make_graph(Y, YT), T->counted ? T->pre : nullptr,
                                                                                T->ntiles, tile_begin, tile_end, chunk_cols, T->rowbase);
    tr_rowscan_kernel<<<1, 1024, 0, ctx->stream>>>(T->rowbase, T->nt, T->counted ? T->pre : nullptr, T->ntiles, tile_begin);
    ctx->launches += 2;
    SS_CHECK_CUDA(cudaMemcpyAsync(entries, T->rowbase + T->nt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    T->tile_begin = tile_begin;
    T->tile_end = tile_end;
    return SS_OK;
}

// materialise the tiles chosen by transfer_rows
int32_t transfer_fill(ss_ctx* ctx, const ss_csr* Y, const ss_csr* YT, ss_transfer* T, int64_t entries) {
    T->cap = entries;
    T->nnz = entries;
    SS_CHECK_CUDA(cudaMalloc(&T->col, size_t(entries + 64) * sizeof(uint16_t)));  // slack: see tr_bank_load
    SS_CHECK_CUDA(cudaMalloc(&T->val, size_t(entries + 64) * sizeof(double)));
    if (entries == 0) {
        if (!T->counted) SS_CHECK_CUDA(cudaMemsetAsync(T->pre, 0, size_t(T->nt) * (T->ntiles + 1) * 4, ctx->stream));
        return SS_OK;
    }
    TrBuild b{};
    b.g = make_graph(Y, YT);
    b.tw = T->tw;
    b.ntiles = T->ntiles;
    b.range_tiles = T->range_tiles;
    b.pre = T->pre;
    b.tile_total = T->tile_total;
    b.tile_begin = T->tile_begin;
    b.tile_end = T->tile_end;
    b.rowbase = T->rowbase;
    b.col = T->col;
    b.val = T->val;
    SS_TRY(row_counter(ctx, &b.row_counter));
    const size_t smem = size_t(T->range_tiles) * (T->tw / 32) * 12 + TB_DUP_SMEM;
    const int grid = int(std::min<int64_t>(T->nt, ctx->sm_count));
#define SS_FILL(W_, P_)                                                                                                       \
    do {                                                                                                                      \
        SS_CHECK_CUDA(cudaFuncSetAttribute(tr_fill_kernel<W_, P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));  \
        tr_fill_kernel<W_, P_><<<grid, TB_THREADS, smem, ctx->stream>>>(b);                                                   \
    } while (0)
    if (T->counted) {
        if (T->weighted) SS_FILL(true, false); else SS_FILL(false, false);
    } else {
        SS_REQUIRE(T->tile_begin == 0 && T->tile_end == T->ntiles, "transfer matrix: a chunk of tiles needs the count pass");
        if (T->weighted) SS_FILL(true, true); else SS_FILL(false, true);
        unsigned long long written = 0;
        SS_CHECK_CUDA(cudaMemcpyAsync(&written, T->tile_total + T->ntiles, 8, cudaMemcpyDeviceToHost, ctx->stream));
        SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        T->nnz = int64_t(written);
    }
#undef SS_FILL
    ctx->launches += 1;
    SS_CHECK_CUDA(cudaGetLastError());
    return SS_OK;
}

template <int EPL>
static int32_t launch_stream(ss_ctx* ctx, TrStream& p, bool weighted, int ntl, int64_t nsrc, int L, void** scratch) {
    constexpr int TW = 1024 * EPL;
    constexpr int warps_per_cta = ts_warps(EPL);
    const size_t smem = size_t(warps_per_cta) * TW * sizeof(double);
    if (weighted) {
        SS_CHECK_CUDA(cudaFuncSetAttribute(tr_stream_kernel<EPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    } else {
        SS_CHECK_CUDA(cudaFuncSetAttribute(tr_stream_kernel<EPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    }
    const int64_t warps = int64_t(ctx->sm_count) * warps_per_cta;
    // work item = (source, range of tiles): few sources are split over tile ranges so that every warp has work
    int parts = int(std::min<int64_t>(ntl, std::max<int64_t>(1, ceil_div(4 * warps, nsrc))));
    const int forced = tr_env_int("SS_RECSYS_PARTS", 0);
    if (forced > 0) parts = std::min(forced, ntl);
    const int64_t nwork = nsrc * parts;
    SS_REQUIRE(nwork < (int64_t(1) << 31), "recommend: too many work items");
    const size_t key_bytes = size_t(nwork) * L * 8, col_bytes = round_up(size_t(nwork) * L * 4, 256);
    SS_TRY(scratch_get(ctx, 15, key_bytes + col_bytes + 256, scratch));
    p.parts = parts;
    p.nwork = int(nwork);
    p.cand_key = static_cast<uint64_t*>(*scratch);
    p.cand_col = reinterpret_cast<int32_t*>(static_cast<char*>(*scratch) + key_bytes);
    p.counter = reinterpret_cast<int*>(static_cast<char*>(*scratch) + key_bytes + col_bytes);
    SS_CHECK_CUDA(cudaMemsetAsync(p.counter, 0, 256, ctx->stream));
    const int grid = int(std::min<int64_t>(ceil_div(nwork, warps_per_cta), ctx->sm_count));
    if (weighted) tr_stream_kernel<EPL, true><<<grid, warps_per_cta * 32, smem, ctx->stream>>>(p);
    else tr_stream_kernel<EPL, false><<<grid, warps_per_cta * 32, smem, ctx->stream>>>(p);
    return SS_OK;
}

// top-L of the sources [s_begin, s_end) over the materialised tiles of T, merged with the running result of
// earlier chunks when `have_prev`
int32_t transfer_stream(ss_ctx* ctx, const ss_csr* Y, const ss_transfer* T, int L, int64_t s_begin, int64_t s_end,
                        bool have_prev, uint64_t* run_key, int32_t* idx_out, double* val_out, int64_t ldv) {
    const int ntl = T->tile_end - T->tile_begin;
    SS_REQUIRE(ntl > 0, "recommend: no materialised tiles");
    const int64_t nsrc = s_end - s_begin;
    if (nsrc <= 0) return SS_OK;
    SS_REQUIRE(nsrc < (int64_t(1) << 30), "recommend: more than 2^30 sources in one call");
    TrStream p{};
    p.y_ptr = Y->row_ptr;
    p.y_idx = Y->col_idx;
    p.y_val = Y->values;
    p.pre = T->pre;
    p.rowbase = T->rowbase;
    p.col = T->col;
    p.val = T->val;
    p.nt = T->nt;
    p.ntiles = T->ntiles;
    p.tile_begin = T->tile_begin;
    p.ntl = ntl;
    p.s0 = s_begin;
    p.L = L;
    void* w = nullptr;
    int32_t st;
    switch (T->tw) {
        case 1024: st = launch_stream<1>(ctx, p, T->weighted, ntl, nsrc, L, &w); break;
        case 2048: st = launch_stream<2>(ctx, p, T->weighted, ntl, nsrc, L, &w); break;
        case 3072: st = launch_stream<3>(ctx, p, T->weighted, ntl, nsrc, L, &w); break;
        case 4096: st = launch_stream<4>(ctx, p, T->weighted, ntl, nsrc, L, &w); break;
        default:
            set_error("recommend: no streaming kernel for tile width %d", T->tw);
            st = SS_ERR_INVALID;
    }
    SS_TRY(st);
    tr_merge_kernel<<<unsigned(ceil_div(nsrc, 8)), 256, 0, ctx->stream>>>(p.cand_key, p.cand_col, p.parts, L, s_begin, int(nsrc),
                                                                        have_prev, run_key, idx_out, val_out, ldv);
    ctx->launches += 2;
    SS_CHECK_CUDA(cudaGetLastError());
    return SS_OK;
}

static int64_t transfer_budget() {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    int64_t budget = int64_t(free_b) - (int64_t(3) << 30);  // candidate scratch, head-room
    const int cap_mb = tr_env_int("SS_RECSYS_U_MB", 0);      // test hook: force several chunks
    if (cap_mb > 0) budget = std::min<int64_t>(budget, int64_t(cap_mb) << 20);
    return std::max<int64_t>(budget, int64_t(64) << 20);
}

// U in one piece (rows sized by the path bound); SS_ERR_OOM when it does not fit the free device memory
int32_t transfer_build_whole(ss_ctx* ctx, const ss_csr* Y, const ss_csr* YT, ss_transfer* T, bool* fits) {
    int64_t entries = 0;
    SS_TRY(transfer_rows(ctx, Y, YT, T, 0, T->ntiles, &entries));
    *fits = entries * 10 <= transfer_budget();
    if (!*fits) return SS_OK;
    return transfer_fill(ctx, Y, YT, T, entries);
}

// the whole call: U in one piece when it fits, else count + chunks of column tiles with a running top-L
// *declined: the transfer matrix does not fit the free device memory in one piece (heavy-tailed graphs: U approaches
// a dense items x items matrix and nearly every entry sums many paths) -- nothing was computed, the caller takes the
// two-hop expansion kernel of ss_recsys.cu instead.
int32_t recommend_topl_stream(ss_ctx* ctx, const ss_csr* Y, const ss_csr* YT, int L, int64_t s_begin, int64_t s_end,
                              int32_t* idx_out, double* val_out, int64_t ldv, bool* declined) {
    *declined = false;
    ss_transfer* T = nullptr;
    int tw = 0;
    unsigned long long paths = 0;
    SS_TRY(transfer_pick_tile(ctx, Y, &tw, &paths));
    // chunks of column tiles are exact but slow on such graphs (count pass + one build per chunk, most entries of U
    // summing many paths): taken only on request (SS_RECSYS_MODE=chunks) or under the test hook SS_RECSYS_U_MB
    const char* mode = getenv("SS_RECSYS_MODE");
    const bool chunks_ok = tr_env_int("SS_RECSYS_U_MB", 0) > 0 || (mode && !strcmp(mode, "chunks"));
    if (!chunks_ok && double(paths) * 10.0 > 2.0 * double(transfer_budget())) {
        *declined = true;
        return SS_OK;
    }
    SS_TRY(transfer_create(ctx, Y, tw, &T));
    bool fits = false;
    int32_t st = transfer_build_whole(ctx, Y, YT, T, &fits);
    if (st == SS_OK && !fits && !chunks_ok) {
        transfer_free(T);
        *declined = true;
        return SS_OK;
    }
    if (st == SS_OK && fits) {
        st = transfer_stream(ctx, Y, T, L, s_begin, s_end, false, nullptr, idx_out, val_out, ldv);
        if (st == SS_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            set_error("recommend: %s", cudaGetErrorString(cudaGetLastError()));
            st = SS_ERR_CUDA;
        }
        transfer_free(T);
        return st;
    }
    if (st == SS_OK) st = transfer_count(ctx, Y, YT, T);
    std::vector<int> cuts{0};
    if (st == SS_OK) {
        const int64_t budget = transfer_budget();
        int64_t acc = 0;
        for (int t = 0; t < T->ntiles; ++t) {
            const int64_t bytes = int64_t(T->tile_total_host[t]) * 10;
            if (acc > 0 && acc + bytes > budget) {
                cuts.push_back(t);
                acc = 0;
            }
            acc += bytes;
        }
        cuts.push_back(T->ntiles);
    }
    uint64_t* run_key = nullptr;
    if (st == SS_OK) {
        void* rk;
        st = scratch_get(ctx, 17, size_t(Y->rows) * L * 8, &rk);
        run_key = static_cast<uint64_t*>(rk);
    }
    for (size_t c = 0; st == SS_OK && c + 1 < cuts.size(); ++c) {
        int64_t entries = 0;
        st = transfer_rows(ctx, Y, YT, T, cuts[c], cuts[c + 1], &entries);
        if (st == SS_OK) st = transfer_fill(ctx, Y, YT, T, entries);
        if (st == SS_OK) st = transfer_stream(ctx, Y, T, L, s_begin, s_end, c > 0, run_key, idx_out, val_out, ldv);
        if (st == SS_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            set_error("recommend: %s", cudaGetErrorString(cudaGetLastError()));
            st = SS_ERR_CUDA;
        }
    }
    transfer_free(T);
    return st;
}

}  // namespace ss

using namespace ss;

extern "C" {

int32_t ss_transfer_build(ss_ctx* ctx, const ss_csr* Y, const ss_csr* YT, ss_transfer** out) {
    SS_REQUIRE(ctx && Y && YT && out, "ss_transfer_build: null argument");
    SS_CHECK_CUDA(cudaSetDevice(ctx->device));
    SS_REQUIRE(YT->rows == Y->cols && YT->cols == Y->rows && YT->nnz == Y->nnz,
               "ss_transfer_build: YT must be the CSR of the transpose of Y");
    SS_REQUIRE((Y->values == nullptr) == (YT->values == nullptr), "ss_transfer_build: Y and YT must both be binary or both weighted");
    *out = nullptr;
    ss_transfer* T = nullptr;
    int tw = 0;
    SS_TRY(transfer_pick_tile(ctx, Y, &tw, nullptr));
    SS_TRY(transfer_create(ctx, Y, tw, &T));
    bool fits = false;
    int32_t st = transfer_build_whole(ctx, Y, YT, T, &fits);
    if (st == SS_OK && !fits) {
        set_error("ss_transfer_build: the transfer matrix does not fit the free device memory (ss_recommend_topl processes it in chunks)");
        st = SS_ERR_OOM;
    }
    if (st == SS_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        set_error("ss_transfer_build: %s", cudaGetErrorString(cudaGetLastError()));
        st = SS_ERR_CUDA;
    }
    if (st != SS_OK) {
        transfer_free(T);
        return st;
    }
    *out = T;
    return SS_OK;
}

int32_t ss_transfer_info(const ss_transfer* T, int64_t* info4) {
    SS_REQUIRE(T && info4, "ss_transfer_info: null argument");
    info4[0] = T->nnz;
    info4[1] = T->cap * 10 + int64_t(T->nt) * (T->ntiles + 1) * 4 + (T->nt + 1) * 8;
    info4[2] = T->tw;
    info4[3] = T->ntiles;
    return SS_OK;
}

int32_t ss_transfer_destroy(ss_transfer* T) {
    transfer_free(T);
    return SS_OK;
}

/* debugging / tests: the materialised U as host arrays (row offsets per item, global column, value) */
int32_t ss_transfer_download(const ss_transfer* T, int64_t* row_ptr_host, int32_t* col_host, double* val_host) {
    SS_REQUIRE(T && row_ptr_host, "ss_transfer_download: null argument");
    SS_REQUIRE(T->tile_begin == 0 && T->tile_end == T->ntiles, "ss_transfer_download: U is not fully materialised");
    SS_CHECK_CUDA(cudaSetDevice(T->ctx->device));
    SS_CHECK_CUDA(cudaStreamSynchronize(T->ctx->stream));
    const int64_t stride = int64_t(T->ntiles) + 1;
    std::vector<uint32_t> pre(size_t(T->nt) * stride);
    std::vector<int64_t> rb(T->nt + 1);
    SS_CHECK_CUDA(cudaMemcpy(pre.data(), T->pre, pre.size() * 4, cudaMemcpyDeviceToHost));
    SS_CHECK_CUDA(cudaMemcpy(rb.data(), T->rowbase, rb.size() * 8, cudaMemcpyDeviceToHost));
    row_ptr_host[0] = 0;
    for (int64_t r = 0; r < T->nt; ++r) row_ptr_host[r + 1] = row_ptr_host[r] + pre[r * stride + T->ntiles];
    if (!col_host || !val_host || T->cap == 0) return SS_OK;
    std::vector<uint16_t> c16(T->cap);
    std::vector<double> v64(T->cap);
    SS_CHECK_CUDA(cudaMemcpy(c16.data(), T->col, size_t(T->cap) * 2, cudaMemcpyDeviceToHost));
    SS_CHECK_CUDA(cudaMemcpy(v64.data(), T->val, size_t(T->cap) * 8, cudaMemcpyDeviceToHost));
    for (int64_t r = 0; r < T->nt; ++r) {  // rows may be padded on the device (sized by a bound): compact them
        int64_t o = row_ptr_host[r];
        for (int c = 0; c < T->ntiles; ++c) {
            const int64_t a = rb[r] + pre[r * stride + c], b = rb[r] + pre[r * stride + c + 1];
            for (int64_t e = a; e < b; ++e, ++o) {
                col_host[o] = int32_t(c) * T->tw + c16[e];
                val_host[o] = v64[e];
            }
        }
    }
    return SS_OK;
}

int32_t ss_recommend_topl_transfer(ss_ctx* ctx, const ss_csr* Y, const ss_transfer* T, int32_t L, int64_t s_begin,
                                   int64_t s_end, ss_ivec* idx_out, ss_mat* val_out) {
    SS_REQUIRE(ctx && Y && T && idx_out, "ss_recommend_topl_transfer: null argument");
    SS_CHECK_CUDA(cudaSetDevice(ctx->device));
    SS_REQUIRE(T->ns == Y->rows && T->nt == Y->cols && T->weighted == (Y->values != nullptr),
               "ss_recommend_topl_transfer: the transfer matrix was built for another graph");
    SS_REQUIRE(T->tile_begin == 0 && T->tile_end == T->ntiles, "ss_recommend_topl_transfer: U is not fully materialised");
    SS_REQUIRE(L >= 1 && L <= 32 && L <= Y->cols, "ss_recommend_topl_transfer: L must be in 1..min(32, targets)");
    SS_REQUIRE(s_begin >= 0 && s_end <= Y->rows && s_begin <= s_end, "ss_recommend_topl_transfer: bad source range");
    SS_REQUIRE(idx_out->n == int64_t(L) * Y->rows, "ss_recommend_topl_transfer: idx_out must hold L x sources entries");
    SS_REQUIRE(!val_out || (val_out->rows == L && val_out->cols == Y->rows),
               "ss_recommend_topl_transfer: val_out must be an L x sources matrix");
    SS_TRY(transfer_stream(ctx, Y, T, L, s_begin, s_end, false, nullptr, idx_out->d, val_out ? val_out->d : nullptr,
                           val_out ? val_out->ld : 0));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

}  // extern "C"
