// Ranking and metric kernels (north-star subsystem 4).
//
//  * per-row top-L in `sortperm(row; rev=true)` order (reference src/performance.jl:315,377:
//    descending under isless, ties by ascending index) and recall@L / precision@L
//    (src/performance.jl:308-328, 341-357, 370-385, 398-414);
//  * AuROC / AuPRC (src/performance.jl:49-63, 74-89) = MLBase.roc over the ascending unique scores
//    + Trapz.trapz: a stable LSD radix sort of (isless-key, label) pairs, then one scan whose state
//    carries (positives so far, latest run start, positives before that run start); every run
//    boundary contributes one trapezoid.
//
// Scores are mapped to uint64 keys that are monotone under Julia's isless
// (-Inf < ... < -0.0 < 0.0 < ... < Inf < NaN, all NaNs equal).
#include <algorithm>

#include "ss_common.cuh"

namespace {

__device__ __forceinline__ uint64_t isless_key(double v) {
    if (v != v) return 0xFFFFFFFFFFFFFFFFull;
    const uint64_t b = uint64_t(__double_as_longlong(v));
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// ------------------------------------------------------------------------------------------------
// top-L per row: one thread per row (lanes along rows -> coalesced column-major reads), private
// sorted list in shared memory (list[l][thread]); a candidate is inserted only when it beats the
// current L-th entry, which becomes rare after the first few hundred columns.
// ------------------------------------------------------------------------------------------------
constexpr int TOPL_TPB = 64;
constexpr int TOPL_UNROLL = 8;

__global__ void __launch_bounds__(TOPL_TPB)
    topl_kernel(const double* __restrict__ R, int64_t rows, int64_t cols, int64_t ld, int L,
                int32_t* __restrict__ idx_out, double* __restrict__ val_out, int64_t ldv) {
    extern __shared__ uint8_t topl_smem[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(topl_smem);                   // [L][TOPL_TPB]
    double* vals = reinterpret_cast<double*>(keys + size_t(L) * TOPL_TPB);      // [L][TOPL_TPB]
    int32_t* idxs = reinterpret_cast<int32_t*>(vals + size_t(L) * TOPL_TPB);    // [L][TOPL_TPB]
    const int t = threadIdx.x;
    const int64_t row = int64_t(blockIdx.x) * TOPL_TPB + t;
    if (row >= rows) return;
    int cnt = 0;
    uint64_t worst = 0;  // key of the current L-th entry once the list is full
    auto insert = [&](uint64_t key, double v, int32_t c) {
        // position after every entry with key >= new key (earlier columns win ties)
        int p = (cnt < L) ? cnt : L - 1;
        while (p > 0 && keys[(p - 1) * TOPL_TPB + t] < key) {
            keys[p * TOPL_TPB + t] = keys[(p - 1) * TOPL_TPB + t];
            vals[p * TOPL_TPB + t] = vals[(p - 1) * TOPL_TPB + t];
            idxs[p * TOPL_TPB + t] = idxs[(p - 1) * TOPL_TPB + t];
            --p;
        }
        keys[p * TOPL_TPB + t] = key;
        vals[p * TOPL_TPB + t] = v;
        idxs[p * TOPL_TPB + t] = c;
        if (cnt < L) ++cnt;
        if (cnt == L) worst = keys[(L - 1) * TOPL_TPB + t];
    };
    int64_t c = 0;
    for (; c + TOPL_UNROLL <= cols; c += TOPL_UNROLL) {
        double v[TOPL_UNROLL];
#pragma unroll
        for (int u = 0; u < TOPL_UNROLL; ++u) v[u] = __ldg(R + (c + u) * ld + row);
#pragma unroll
        for (int u = 0; u < TOPL_UNROLL; ++u) {
            const uint64_t key = isless_key(v[u]);
            if (cnt < L || key > worst) insert(key, v[u], int32_t(c + u));
        }
    }
    for (; c < cols; ++c) {
        const double v = __ldg(R + c * ld + row);
        const uint64_t key = isless_key(v);
        if (cnt < L || key > worst) insert(key, v, int32_t(c));
    }
    for (int l = 0; l < L; ++l) {
        idx_out[row * L + l] = (l < cnt) ? idxs[l * TOPL_TPB + t] : -1;
        if (val_out) val_out[row * ldv + l] = (l < cnt) ? vals[l * TOPL_TPB + t] : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// Warp-level top-L (L <= 32).  A block owns 32 consecutive rows: column chunks are loaded with lanes
// along the rows (coalesced for the column-major R), transposed through shared memory, and each warp
// then scans ITS rows with lanes along the columns.  The sorted list of a row lives across the lanes
// of the warp (lane i = rank i); a chunk is filtered with one compare + ballot against the current
// L-th key, and the rare survivors are inserted with a popc rank and one shuffle.  Candidates are
// taken in ascending column order and equal keys never displace an earlier column, which is exactly
// the stable descending order of `sortperm(row; rev=true)`.
// ------------------------------------------------------------------------------------------------
constexpr int WT_ROWS = 32;
constexpr int WT_TPB = 256;  // 8 warps, 4 rows each

__device__ __forceinline__ double key_to_value(uint64_t k) {
    if (k == 0xFFFFFFFFFFFFFFFFull) return __longlong_as_double(0x7ff8000000000000ll);
    const uint64_t b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

template <int WT_COLS, int MIN_BLOCKS>
__global__ void __launch_bounds__(WT_TPB, MIN_BLOCKS)
    topl_warp_kernel(const double* __restrict__ R, int64_t rows, int64_t cols, int64_t ld, int L,
                     int32_t* __restrict__ idx_out, double* __restrict__ val_out, int64_t ldv) {
    extern __shared__ double wt_tile[];  // [WT_ROWS][WT_COLS + 1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = int64_t(blockIdx.x) * WT_ROWS;
    // per-warp state of its 4 rows: lane i holds the i-th best (key, column); cnt = entries so far
    uint64_t lkey[4];
    int32_t lidx[4];
    int cnt[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        lkey[q] = 0;
        lidx[q] = -1;
        cnt[q] = 0;
    }
    // The next tile is loaded into registers while the current one is scanned (the loads used to sit between two
    // barriers with nothing to overlap them).  Warp w takes columns w, w+8, ...; lane = row (256 B contiguous per
    // column); out-of-range cells read a clamped, valid address and are ignored by the scan.
    constexpr int WT_PER = WT_COLS / (WT_TPB / 32);
    double pre[WT_PER];
    const int64_t rl = min(row0 + lane, rows - 1);
    auto prefetch = [&](int64_t c0) {
        if (c0 + WT_COLS <= cols) {  // whole tile in range: one pointer, constant stride
            const double* pc = R + (c0 + warp) * ld + rl;
            const int64_t step = int64_t(WT_TPB / 32) * ld;
#pragma unroll
            for (int j = 0; j < WT_PER; ++j) pre[j] = __ldg(pc + j * step);
        } else {
#pragma unroll
            for (int j = 0; j < WT_PER; ++j) {
                const int64_t c = min(c0 + warp + j * (WT_TPB / 32), cols - 1);
                pre[j] = __ldg(R + c * ld + rl);
            }
        }
    };
    prefetch(0);
    // Filter state of the four rows of this warp.  Once a list is full a cell qualifies iff its key beats the key of
    // rank L-1.  For non-NaN values the key order is the numeric order with -0.0 < +0.0, so the filter is ONE FP64
    // compare per cell, !(v <= tv) (true for v > tv and for NaN, whose key is the largest), instead of the 64-bit key
    // transform + integer compare; the two thresholds where that differs from the key order (tv = -0.0: a +0.0 cell
    // beats it; tv = NaN: nothing beats it) switch the row to the exact key compare.  Survivors are re-checked against
    // the exact keys when they are ranked, so the filter only has to let every qualifying cell through.
    constexpr uint64_t KEY_NEGZERO = 0x7FFFFFFFFFFFFFFFull, KEY_NAN = 0xFFFFFFFFFFFFFFFFull;
    double tv[4];
    bool full[4], live[4], exact[4];
    auto refresh = [&](int q) {
        const uint64_t thr = __shfl_sync(0xffffffffu, lkey[q], L - 1);  // key of rank L-1 once the list is full
        full[q] = cnt[q] >= L;
        tv[q] = key_to_value(thr);
        exact[q] = full[q] && (thr == KEY_NEGZERO || thr == KEY_NAN);
    };
#pragma unroll
    for (int q = 0; q < 4; ++q) live[q] = row0 + 4 * warp + q < rows;  // warp-uniform
    for (int64_t c0 = 0; c0 < cols; c0 += WT_COLS) {
        const int nc = int(min(int64_t(WT_COLS), cols - c0));
        __syncthreads();
#pragma unroll
        for (int j = 0; j < WT_PER; ++j) wt_tile[lane * (WT_COLS + 1) + warp + j * (WT_TPB / 32)] = pre[j];
        __syncthreads();
        if (c0 + WT_COLS < cols) prefetch(c0 + WT_COLS);
        // the four rows of this warp are independent chains: issue their loads / ballots together
#pragma unroll
        for (int q = 0; q < 4; ++q) refresh(q);
        for (int cb = 0; cb < nc; cb += 32) {
            const int c = cb + lane;
            const bool inb = c < nc;
            double v[4];
            unsigned cand[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = wt_tile[(4 * warp + q) * (WT_COLS + 1) + c];  // c < WT_COLS always
            if (exact[0] | exact[1] | exact[2] | exact[3]) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint64_t thr = __shfl_sync(0xffffffffu, lkey[q], L - 1);
                    cand[q] = __ballot_sync(0xffffffffu, inb && live[q] && (!full[q] || isless_key(v[q]) > thr));
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    cand[q] = __ballot_sync(0xffffffffu, inb && live[q] && (!full[q] || !(v[q] <= tv[q])));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (!cand[q]) continue;  // the common case after the first few hundred columns
                unsigned cd = cand[q];
                while (cd) {
                    const int src = __ffs(cd) - 1;
                    cd &= cd - 1;
                    const uint64_t ck = isless_key(__shfl_sync(0xffffffffu, v[q], src));
                    const int32_t ci = int32_t(c0 + cb + src);
                    // rank = number of list entries with key >= ck (earlier columns win ties)
                    const bool ge = (lane < cnt[q]) && (lkey[q] >= ck);
                    const int pos = __popc(__ballot_sync(0xffffffffu, ge));
                    if (pos >= L) continue;  // no longer qualifies after earlier inserts of this chunk
                    const uint64_t upk = __shfl_up_sync(0xffffffffu, lkey[q], 1);
                    const int32_t upi = __shfl_up_sync(0xffffffffu, lidx[q], 1);
                    if (lane == pos) {
                        lkey[q] = ck;
                        lidx[q] = ci;
                    } else if (lane > pos) {
                        lkey[q] = upk;
                        lidx[q] = upi;
                    }
                    if (cnt[q] < L) ++cnt[q];
                }
                refresh(q);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t row = row0 + 4 * warp + q;
        if (row < rows && lane < L) {
            const bool have = lane < cnt[q];
            idx_out[row * L + lane] = have ? lidx[q] : -1;
            if (val_out) val_out[row * ldv + lane] = have ? key_to_value(lkey[q]) : 0.0;
        }
    }
}

// recall@L / precision@L per row from the top-L index lists (src/performance.jl:315-327,377-384):
// Xi = sum(y), Xi_L = sum(first(y[order], L)).
constexpr int ATL_TPB = 128;
__global__ void __launch_bounds__(ATL_TPB)
    atl_rows_kernel(const double* __restrict__ Y, int64_t rows, int64_t cols, int64_t ldy, int L,
                    const int32_t* __restrict__ idx, double* __restrict__ rec, double* __restrict__ prec) {
    const int64_t row = int64_t(blockIdx.x) * ATL_TPB + threadIdx.x;
    if (row >= rows) return;
    double xi = 0.0;
    for (int64_t c = 0; c < cols; ++c) xi += __ldg(Y + c * ldy + row);
    double xil = 0.0;
    for (int l = 0; l < L; ++l) xil += __ldg(Y + int64_t(idx[row * L + l]) * ldy + row);
    rec[row] = (xi > 0.0) ? xil / xi : __longlong_as_double(0x7ff8000000000000ll);
    prec[row] = xil / double(L);
}

// deterministic single-block sum of up to two arrays; out[j] = sum(a_j) / n
__global__ void __launch_bounds__(1024) mean2_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                     int64_t n, double* __restrict__ out) {
    __shared__ double sa[1024], sb[1024];
    const int t = threadIdx.x;
    double xa = 0.0, xb = 0.0;
    for (int64_t i = t; i < n; i += 1024) {
        xa += a[i];
        xb += b[i];
    }
    sa[t] = xa;
    sb[t] = xb;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (t < s) {
            sa[t] += sa[t + s];
            sb[t] += sb[t + s];
        }
        __syncthreads();
    }
    if (t == 0) {
        out[0] = sa[0] / double(n);
        out[1] = sb[0] / double(n);
    }
}

// ------------------------------------------------------------------------------------------------
// radix sort (LSD, 8-bit digits) of (uint64 key, uint8 label) pairs, stable.
// ------------------------------------------------------------------------------------------------
constexpr int RS_TPB = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_TPB * RS_ITEMS;  // 4096 keys per block

__global__ void __launch_bounds__(256)
    make_keys_kernel(const double* __restrict__ scores, const uint8_t* __restrict__ labels, int64_t M,
                     uint64_t* __restrict__ keys, uint8_t* __restrict__ lab) {
    for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < M; i += int64_t(gridDim.x) * 256) {
        keys[i] = isless_key(scores[i]);
        lab[i] = labels[i] ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256)
    make_keys_mat_kernel(const double* __restrict__ R, int64_t ldr, const double* __restrict__ Y, int64_t ldy,
                         int64_t rows, int64_t cols, uint64_t* __restrict__ keys, uint8_t* __restrict__ lab) {
    const int64_t r = int64_t(blockIdx.x) * 256 + threadIdx.x;
    if (r >= rows) return;
    for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) {
        keys[c * rows + r] = isless_key(R[c * ldr + r]);
        lab[c * rows + r] = (Y[c * ldy + r] != 0.0) ? 1 : 0;
    }
}

// histogram of all 8 digits in one pass (decides which passes are trivial)
__global__ void __launch_bounds__(256)
    digit_hist_kernel(const uint64_t* __restrict__ keys, int64_t M, unsigned long long* __restrict__ ghist) {
    __shared__ unsigned int h[8 * 256];
    for (int i = threadIdx.x; i < 8 * 256; i += 256) h[i] = 0;
    __syncthreads();
    // a block sees at most 2^32-1 keys only if M / gridDim.x < 2^32: guaranteed by the launcher
    for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < M; i += int64_t(gridDim.x) * 256) {
        const uint64_t k = keys[i];
#pragma unroll
        for (int d = 0; d < 8; ++d) atomicAdd(&h[d * 256 + ((k >> (8 * d)) & 255)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 256; i += 256)
        if (h[i]) atomicAdd(&ghist[i], (unsigned long long)h[i]);
}

// per-block digit counts for one pass: hist[digit * nblocks + block]
__global__ void __launch_bounds__(RS_TPB)
    rs_upsweep_kernel(const uint64_t* __restrict__ keys, int64_t M, int shift, uint32_t* __restrict__ hist,
                      int64_t nblocks) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = int64_t(blockIdx.x) * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int64_t idx = base + i * RS_TPB + threadIdx.x;
        if (idx < M) atomicAdd(&h[(keys[idx] >> shift) & 255], 1u);
    }
    __syncthreads();
    hist[int64_t(threadIdx.x) * nblocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of every digit row of hist[digit][block] (one block per digit, coalesced, running
// carry) -> per-(digit, block) offset relative to the digit's start + digit totals
__global__ void __launch_bounds__(1024)
    rs_scan_rows_kernel(const uint32_t* __restrict__ hist, int64_t nblocks, uint32_t* __restrict__ rel,
                        unsigned long long* __restrict__ totals) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    const int d = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t* in = hist + int64_t(d) * nblocks;
    uint32_t* out = rel + int64_t(d) * nblocks;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int64_t i0 = 0; i0 < nblocks; i0 += 1024) {
        const int64_t i = i0 + t;
        const unsigned long long v = (i < nblocks) ? in[i] : 0;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        unsigned long long wpre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            if (w == warp) wpre = tot;
            tot += wsum[w];
        }
        const unsigned long long carry = carry_s;
        if (i < nblocks) out[i] = uint32_t(carry + wpre + inc - v);
        __syncthreads();
        if (t == 0) carry_s = carry + tot;
    }
    __syncthreads();
    if (t == 0) totals[d] = carry_s;
}

// exclusive scan of the 256 digit totals -> global start of every digit
__global__ void __launch_bounds__(256)
    rs_scan_totals_kernel(const unsigned long long* __restrict__ totals, unsigned long long* __restrict__ base) {
    __shared__ unsigned long long s[256];
    s[threadIdx.x] = totals[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 256; ++i) {
            const unsigned long long v = s[i];
            s[i] = run;
            run += v;
        }
    }
    __syncthreads();
    base[threadIdx.x] = s[threadIdx.x];
}

// stable scatter: rank inside the warp by __match_any_sync, across warps by a per-digit prefix; the
// tile is first ordered by digit in shared memory so that the global stores of one digit run are
// contiguous (coalesced) instead of one scattered 8-byte store per key.
__global__ void __launch_bounds__(RS_TPB)
    rs_downsweep_kernel(const uint64_t* __restrict__ keys_in, const uint8_t* __restrict__ lab_in, int64_t M, int shift,
                        const uint32_t* __restrict__ rel, const unsigned long long* __restrict__ base, int64_t nblocks,
                        uint64_t* __restrict__ keys_out, uint8_t* __restrict__ lab_out) {
    extern __shared__ uint8_t rs_smem[];
    uint64_t* skey = reinterpret_cast<uint64_t*>(rs_smem);                      // [RS_TILE]
    unsigned int(*wcount)[256] = reinterpret_cast<unsigned int(*)[256]>(skey + RS_TILE);  // [8][256]
    unsigned long long* gbase = reinterpret_cast<unsigned long long*>(wcount + RS_TPB / 32);  // [256]
    unsigned int* dstart = reinterpret_cast<unsigned int*>(gbase + 256);        // [256]
    unsigned int* wtot = dstart + 256;                                          // [8]
    uint8_t* slab = reinterpret_cast<uint8_t*>(wtot + 8);                       // [RS_TILE]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (RS_TPB / 32) * 256; i += RS_TPB) (&wcount[0][0])[i] = 0;
    __syncthreads();
    // order inside the tile: (warp, item, lane) == ascending index
    const int64_t tbase = int64_t(blockIdx.x) * RS_TILE;
    const int64_t wbase = tbase + int64_t(warp) * (32 * RS_ITEMS);
    uint64_t key[RS_ITEMS];
    uint8_t lab[RS_ITEMS];
    unsigned int rank[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int64_t idx = wbase + i * 32 + lane;
        const bool valid = idx < M;
        key[i] = valid ? keys_in[idx] : 0xFFFFFFFFFFFFFFFFull;
        lab[i] = valid ? lab_in[idx] : 0;
    }
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const unsigned int d = (unsigned int)((key[i] >> shift) & 255);
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        unsigned int old = 0;
        if (lane == leader) {
            old = wcount[warp][d];
            wcount[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[i] = old + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {
        const int d = threadIdx.x;  // 256 threads <-> 256 digits
        unsigned int run = 0;
#pragma unroll
        for (int w = 0; w < RS_TPB / 32; ++w) {
            const unsigned int v = wcount[w][d];
            wcount[w][d] = run;
            run += v;
        }
        gbase[d] = base[d] + rel[int64_t(d) * nblocks + blockIdx.x];
        // exclusive scan of the 256 digit totals of this tile
        unsigned int inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        unsigned int wpre = 0;
#pragma unroll
        for (int w = 0; w < RS_TPB / 32; ++w)
            if (w < warp) wpre += wtot[w];
        dstart[d] = wpre + inc - run;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const unsigned int d = (unsigned int)((key[i] >> shift) & 255);
        const unsigned int pos = dstart[d] + wcount[warp][d] + rank[i];
        skey[pos] = key[i];
        slab[pos] = lab[i];
    }
    __syncthreads();
    const int64_t rem = M - tbase;
    const int count = int(rem < RS_TILE ? rem : RS_TILE);
    for (int i = threadIdx.x; i < count; i += RS_TPB) {
        const uint64_t k = skey[i];
        const unsigned int d = (unsigned int)((k >> shift) & 255);
        const unsigned long long g = gbase[d] + (unsigned int)(i - dstart[d]);
        keys_out[g] = k;
        lab_out[g] = slab[i];
    }
}

// ------------------------------------------------------------------------------------------------
// ROC / PR curve integration over the sorted pairs
// ------------------------------------------------------------------------------------------------
struct CurveState {
    unsigned long long pos;       // positives in the covered range
    long long start;              // index of the latest run start in the range (-1: none)
    unsigned long long startpos;  // positives in the range strictly before that run start
};
__device__ __forceinline__ CurveState cs_identity() { return {0ull, -1ll, 0ull}; }
__device__ __forceinline__ CurveState cs_combine(const CurveState& l, const CurveState& r) {
    CurveState o;
    o.pos = l.pos + r.pos;
    if (r.start >= 0) {
        o.start = r.start;
        o.startpos = l.pos + r.startpos;
    } else {
        o.start = l.start;
        o.startpos = l.startpos;
    }
    return o;
}
__device__ __forceinline__ CurveState cs_shfl_up(const CurveState& s, int delta) {
    CurveState o;
    o.pos = __shfl_up_sync(0xffffffffu, s.pos, delta);
    o.start = __shfl_up_sync(0xffffffffu, s.start, delta);
    o.startpos = __shfl_up_sync(0xffffffffu, s.startpos, delta);
    return o;
}

constexpr int CV_TPB = 256;
constexpr int CV_ITEMS = 8;
constexpr int CV_TILE = CV_TPB * CV_ITEMS;

// block-wide exclusive scan of per-thread aggregates; returns the exclusive prefix of this thread
// and (in *block_total) the aggregate of the block.
__device__ CurveState cs_block_exclusive(CurveState agg, CurveState* block_total) {
    __shared__ CurveState warp_tot[CV_TPB / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    CurveState inc = agg;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        CurveState o = cs_shfl_up(inc, d);
        if (lane >= d) inc = cs_combine(o, inc);
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    CurveState wprefix = cs_identity();
    CurveState total = cs_identity();
#pragma unroll
    for (int w = 0; w < CV_TPB / 32; ++w) {
        if (w == warp) wprefix = total;
        total = cs_combine(total, warp_tot[w]);
    }
    CurveState excl = cs_shfl_up(inc, 1);
    if (lane == 0) excl = cs_identity();
    *block_total = total;
    __syncthreads();
    return cs_combine(wprefix, excl);
}

// What a sorted array is a SEGMENT of (multi-GPU AuROC: every rank holds one key range of the global order).
// P / Mtot: positives / pairs of the whole list; idx0: pairs in the lower segments; init: scan state of everything
// below (its `start` is a GLOBAL index).  A stand-alone array is the segment {P = its positives, Mtot = M, 0, identity}.
struct CurveGlobal {
    unsigned long long P, Mtot;
    long long idx0;
    CurveState init;
};

// EMIT == false: write the block aggregate.  EMIT == true: use the scanned block prefixes and add
// one trapezoid per run boundary into partial[block] (roc) / partial[nblocks + block] (pr).
template <bool EMIT>
__global__ void __launch_bounds__(CV_TPB)
    curve_kernel(const uint64_t* __restrict__ keys, const uint8_t* __restrict__ lab, int64_t M,
                 CurveState* __restrict__ block_state, const CurveGlobal* __restrict__ glob,
                 double* __restrict__ partial, int64_t nblocks) {
    const int64_t base = int64_t(blockIdx.x) * CV_TILE + int64_t(threadIdx.x) * CV_ITEMS;
    uint64_t k[CV_ITEMS + 1];
    uint8_t l[CV_ITEMS];
    k[0] = (base > 0 && base - 1 < M) ? keys[base - 1] : 0;
#pragma unroll
    for (int i = 0; i < CV_ITEMS; ++i) {
        const int64_t idx = base + i;
        k[i + 1] = (idx < M) ? keys[idx] : 0;
        l[i] = (idx < M) ? lab[idx] : 0;
    }
    CurveState agg = cs_identity();
#pragma unroll
    for (int i = 0; i < CV_ITEMS; ++i) {
        const int64_t idx = base + i;
        if (idx < M) {
            const bool is_start = (idx == 0) || (k[i + 1] != k[i]);
            CurveState e = {(unsigned long long)l[i], is_start ? (long long)idx : -1ll, 0ull};
            agg = cs_combine(agg, e);
        }
    }
    CurveState total;
    CurveState excl = cs_block_exclusive(agg, &total);
    if (!EMIT) {
        if (threadIdx.x == 0) block_state[blockIdx.x] = total;
        return;
    }
    const CurveGlobal g = *glob;
    // exclusive prefix over the whole (global) list; local run starts are shifted to global indices
    CurveState loc = cs_combine(block_state[blockIdx.x], excl);
    if (loc.start >= 0) loc.start += g.idx0;
    CurveState run = cs_combine(g.init, loc);
    const double P = double(g.P);
    const double N = double(g.Mtot - g.P);
    double roc = 0.0, pr = 0.0;
#pragma unroll
    for (int i = 0; i < CV_ITEMS; ++i) {
        const int64_t idx = base + i;
        if (idx < M) {
            const bool is_start = (idx == 0) || (k[i + 1] != k[i]);  // keys never straddle two segments
            const long long gidx = (long long)idx + g.idx0;
            if (is_start && run.start >= 0) {
                // lower threshold = previous run (start p), higher threshold = this run (start gidx)
                const double p = double(run.start);
                const double tp1 = P - double(run.startpos), fp1 = N - (p - double(run.startpos));
                const double tp2 = P - double(run.pos), fp2 = N - (double(gidx) - double(run.pos));
                roc += (fp2 / N - fp1 / N) * (tp1 / P + tp2 / P) * 0.5;
                pr += (tp2 / P - tp1 / P) * (tp1 / (tp1 + fp1) + tp2 / (tp2 + fp2)) * 0.5;
            }
            CurveState e = {(unsigned long long)l[i], is_start ? gidx : -1ll, 0ull};
            run = cs_combine(run, e);
        }
    }
    // block reduction (deterministic order)
    __shared__ double sroc[CV_TPB], spr[CV_TPB];
    sroc[threadIdx.x] = roc;
    spr[threadIdx.x] = pr;
    __syncthreads();
    for (int s = CV_TPB / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            sroc[threadIdx.x] += sroc[threadIdx.x + s];
            spr[threadIdx.x] += spr[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = sroc[0];
        partial[nblocks + blockIdx.x] = spr[0];
    }
}

// single-block exclusive scan of the block aggregates (in place) + grand total of positives
__global__ void __launch_bounds__(1024)
    curve_scan_kernel(CurveState* __restrict__ st, int64_t n, int64_t M, CurveGlobal* __restrict__ glob,
                      CurveState* __restrict__ summary) {
    __shared__ CurveState part[1024];
    const int t = threadIdx.x;
    const int64_t chunk = (n + 1023) / 1024;
    const int64_t b = t * chunk, e = min(n, b + chunk);
    CurveState s = cs_identity();
    for (int64_t i = b; i < e; ++i) s = cs_combine(s, st[i]);
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        CurveState run = cs_identity();
        for (int i = 0; i < 1024; ++i) {
            const CurveState v = part[i];
            part[i] = run;
            run = cs_combine(run, v);
        }
        glob->P = run.pos;  // stand-alone list; a segment caller overwrites *glob with the global picture
        glob->Mtot = (unsigned long long)M;
        glob->idx0 = 0;
        glob->init = cs_identity();
        *summary = run;     // (positives, last run start, positives before it) of this array
    }
    __syncthreads();
    CurveState run = part[t];
    for (int64_t i = b; i < e; ++i) {
        const CurveState v = st[i];
        st[i] = run;
        run = cs_combine(run, v);
    }
}

// out[0] = |sum roc partials|, out[1] = |sum pr partials|
__global__ void __launch_bounds__(1024)
    curve_final_kernel(const double* __restrict__ partial, int64_t nblocks, double* __restrict__ out, int keep_sign) {
    __shared__ double sa[1024], sb[1024];
    const int t = threadIdx.x;
    double a = 0.0, b = 0.0;
    for (int64_t i = t; i < nblocks; i += 1024) {
        a += partial[i];
        b += partial[nblocks + i];
    }
    sa[t] = a;
    sb[t] = b;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (t < s) {
            sa[t] += sa[t + s];
            sb[t] += sb[t + s];
        }
        __syncthreads();
    }
    if (t == 0) {
        out[0] = keep_sign ? sa[0] : fabs(sa[0]);
        out[1] = keep_sign ? sb[0] : fabs(sb[0]);
    }
}

// ---- confusion-matrix metrics over every unique threshold (reference src/performance.jl:102-296,
// ---- 425-531: maxperformance / meanperformance / meanstdperformance) ----------------------------
enum { MET_F1 = 0, MET_MCC = 1, MET_ACC = 2, MET_BACC = 3, MET_RECALL = 4, MET_PRECISION = 5 };

__device__ __forceinline__ double mcc_eps(double a, double b) {  // src/performance.jl:150-152
    const double e = 2.2250738585072014e-308;                    // floatmin(Float64)
    return (a * e - b * e) / sqrt((a + b) * (a + e) * (b + e) * (e + e));
}

__device__ double confusion_metric(int metric, long long tn, long long fp, long long fn, long long tp) {
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    switch (metric) {
        case MET_F1: {
            const double den = double(tp) + 0.5 * double(fp + fn);
            return den == 0.0 ? nan : double(tp) / den;
        }
        case MET_MCC: {
            const long long p_pred = tp + fp, n_pred = fn + tn, p_act = tp + fn, n_act = fp + tn;
            if (p_pred == 0) return mcc_eps(double(tn), double(fn));
            if (n_pred == 0) return mcc_eps(double(tp), double(fp));
            if (p_act == 0) return mcc_eps(double(tn), double(fp));
            if (n_act == 0) return mcc_eps(double(tp), double(fn));
            const __int128 num = (__int128)tp * tn - (__int128)fp * fn;
            // the reference multiplies the four Int64 counts (and overflows beyond ~3e18); here the
            // product is formed in floating point
            return double(num) / sqrt(double(p_pred) * double(n_pred) * double(p_act) * double(n_act));
        }
        case MET_ACC: {
            const long long den = (tp + tn) + (fp + fn);
            return den == 0 ? nan : double(tp + tn) / double(den);
        }
        case MET_BACC:
            return (double(tp) / double(tp + fn) + double(tn) / double(tn + fp)) / 2.0;
        case MET_RECALL:
            return (tp + fn) == 0 ? nan : double(tp) / double(tp + fn);
        default:
            return (tp + fp) == 0 ? nan : double(tp) / double(tp + fp);
    }
}

// one value per unique threshold (run start); PASS 0: sum / max / NaN flag / count, PASS 1: sum of
// squared deviations from `mean`.  partial: [4][nblocks]
template <int PASS>
__global__ void __launch_bounds__(CV_TPB)
    sweep_kernel(const uint64_t* __restrict__ keys, const uint8_t* __restrict__ lab, int64_t M,
                 const CurveState* __restrict__ block_state, const unsigned long long* __restrict__ totals, int metric,
                 double mean, double* __restrict__ partial, int64_t nblocks) {
    const int64_t base = int64_t(blockIdx.x) * CV_TILE + int64_t(threadIdx.x) * CV_ITEMS;
    uint64_t k[CV_ITEMS + 1];
    uint8_t l[CV_ITEMS];
    k[0] = (base > 0 && base - 1 < M) ? keys[base - 1] : 0;
#pragma unroll
    for (int i = 0; i < CV_ITEMS; ++i) {
        const int64_t idx = base + i;
        k[i + 1] = (idx < M) ? keys[idx] : 0;
        l[i] = (idx < M) ? lab[idx] : 0;
    }
    CurveState agg = cs_identity();
#pragma unroll
    for (int i = 0; i < CV_ITEMS; ++i) {
        const int64_t idx = base + i;
        if (idx < M) {
            const bool is_start = (idx == 0) || (k[i + 1] != k[i]);
            CurveState e = {(unsigned long long)l[i], is_start ? (long long)idx : -1ll, 0ull};
            agg = cs_combine(agg, e);
        }
    }
    CurveState total;
    CurveState excl = cs_block_exclusive(agg, &total);
    CurveState run = cs_combine(block_state[blockIdx.x], excl);
    const long long P = (long long)totals[0], N = (long long)M - P;
    double sum = 0.0, mx = -1.0 / 0.0, cnt = 0.0, nanflag = 0.0;
#pragma unroll
    for (int i = 0; i < CV_ITEMS; ++i) {
        const int64_t idx = base + i;
        if (idx < M) {
            const bool is_start = (idx == 0) || (k[i + 1] != k[i]);
            if (is_start) {  // threshold = this score: predicted positive <=> score >= threshold
                const long long below_pos = (long long)run.pos, below_all = idx;
                const long long fn = below_pos, tn = below_all - below_pos;
                const double v = confusion_metric(metric, tn, N - tn, fn, P - fn);
                if (PASS == 0) {
                    if (v != v) nanflag = 1.0;
                    else { sum += v; mx = fmax(mx, v); }
                    cnt += 1.0;
                } else {
                    const double dlt = v - mean;
                    sum += dlt * dlt;
                }
            }
            CurveState e = {(unsigned long long)l[i], is_start ? (long long)idx : -1ll, 0ull};
            run = cs_combine(run, e);
        }
    }
    __shared__ double s0[CV_TPB], s1[CV_TPB], s2[CV_TPB], s3[CV_TPB];
    s0[threadIdx.x] = sum;
    s1[threadIdx.x] = mx;
    s2[threadIdx.x] = cnt;
    s3[threadIdx.x] = nanflag;
    __syncthreads();
    for (int st = CV_TPB / 2; st > 0; st >>= 1) {
        if (threadIdx.x < st) {
            s0[threadIdx.x] += s0[threadIdx.x + st];
            s1[threadIdx.x] = fmax(s1[threadIdx.x], s1[threadIdx.x + st]);
            s2[threadIdx.x] += s2[threadIdx.x + st];
            s3[threadIdx.x] += s3[threadIdx.x + st];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = s0[0];
        partial[nblocks + blockIdx.x] = s1[0];
        partial[2 * nblocks + blockIdx.x] = s2[0];
        partial[3 * nblocks + blockIdx.x] = s3[0];
    }
}

// out[0] = sum, out[1] = max, out[2] = count, out[3] = NaN count
__global__ void __launch_bounds__(1024)
    sweep_final_kernel(const double* __restrict__ partial, int64_t nblocks, double* __restrict__ out) {
    __shared__ double s0[1024], s1[1024], s2[1024], s3[1024];
    const int t = threadIdx.x;
    double a = 0.0, b = -1.0 / 0.0, c = 0.0, d = 0.0;
    for (int64_t i = t; i < nblocks; i += 1024) {
        a += partial[i];
        b = fmax(b, partial[nblocks + i]);
        c += partial[2 * nblocks + i];
        d += partial[3 * nblocks + i];
    }
    s0[t] = a; s1[t] = b; s2[t] = c; s3[t] = d;
    __syncthreads();
    for (int st = 512; st > 0; st >>= 1) {
        if (t < st) {
            s0[t] += s0[t + st];
            s1[t] = fmax(s1[t], s1[t + st]);
            s2[t] += s2[t + st];
            s3[t] += s3[t + st];
        }
        __syncthreads();
    }
    if (t == 0) { out[0] = s0[0]; out[1] = s1[0]; out[2] = s2[0]; out[3] = s3[0]; }
}

// BEDROC (reference src/performance.jl:22-38): sum over the positives of exp(-alpha * rank / N),
// ranks 1-based in the stable descending order (= ascending order of the inverted keys).
__global__ void __launch_bounds__(256)
    bedroc_sum_kernel(const uint8_t* __restrict__ lab, int64_t M, double alpha, double* __restrict__ partial,
                      unsigned long long* __restrict__ npos_partial) {
    __shared__ double s[256];
    __shared__ unsigned long long c[256];
    double acc = 0.0;
    unsigned long long n = 0;
    for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < M; i += int64_t(gridDim.x) * 256)
        if (lab[i]) {
            acc += exp(-alpha * double(i + 1) / double(M));
            ++n;
        }
    s[threadIdx.x] = acc;
    c[threadIdx.x] = n;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) {
            s[threadIdx.x] += s[threadIdx.x + st];
            c[threadIdx.x] += c[threadIdx.x + st];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = s[0];
        npos_partial[blockIdx.x] = c[0];
    }
}

__global__ void __launch_bounds__(256)
    make_keys_mat_inv_kernel(const double* __restrict__ R, int64_t ldr, const double* __restrict__ Y, int64_t ldy,
                             int64_t rows, int64_t cols, int invert, uint64_t* __restrict__ keys,
                             uint8_t* __restrict__ lab) {
    const int64_t r = int64_t(blockIdx.x) * 256 + threadIdx.x;
    if (r >= rows) return;
    for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) {
        const uint64_t k = isless_key(R[c * ldr + r]);
        keys[c * rows + r] = invert ? ~k : k;
        lab[c * rows + r] = (Y[c * ldy + r] == 1.0) ? 1 : 0;  // BEDROC counts `y .== 1`
    }
}

// pos[q] = number of keys < query[q] in the ascending array (one thread per query)
__global__ void __launch_bounds__(64)
    lower_bound_kernel(const uint64_t* __restrict__ keys, int64_t M, const uint64_t* __restrict__ query, int nq,
                       long long* __restrict__ pos) {
    const int q = blockIdx.x * 64 + threadIdx.x;
    if (q >= nq) return;
    const uint64_t v = query[q];
    int64_t lo = 0, hi = M;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    pos[q] = lo;
}

// stable ascending LSD radix sort of the (key, label) pairs; *kout / *lout point at the sorted arrays
int32_t sort_pairs(ss_ctx* ctx, uint64_t* keysA, uint8_t* labA, uint64_t* keysB, uint8_t* labB, int64_t M,
                   uint64_t** kout_p, uint8_t** lout_p) {
    using namespace ss;
    void* p;
    const int64_t nblocks = ceil_div(M, RS_TILE);
    // scratch: global digit histogram (8*256 u64) | digit totals, digit bases (256 u64 each) |
    // per-pass hist (256*nblocks u32) | per-(digit, block) relative offsets (u32)
    SS_TRY(scratch_get(ctx, 11, size_t(8 * 256 + 512) * 8 + size_t(256) * nblocks * 4 * 2 + 64, &p));
    unsigned long long* ghist = static_cast<unsigned long long*>(p);
    unsigned long long* dtotals = ghist + 8 * 256;
    unsigned long long* dbase = dtotals + 256;
    uint32_t* hist = reinterpret_cast<uint32_t*>(dbase + 256);
    uint32_t* rel = hist + 256 * nblocks;
    const size_t ds_smem = size_t(RS_TILE) * 8 + (RS_TPB / 32) * 256 * 4 + 256 * 8 + 256 * 4 + 8 * 4 + RS_TILE;
    SS_CHECK_CUDA(cudaFuncSetAttribute(rs_downsweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ds_smem)));
    SS_CHECK_CUDA(cudaMemsetAsync(ghist, 0, 8 * 256 * 8, ctx->stream));
    int hgrid = int(ceil_div(M, 256 * 64));
    if (hgrid > ctx->sm_count * 8) hgrid = ctx->sm_count * 8;
    if (hgrid < 1) hgrid = 1;
    digit_hist_kernel<<<hgrid, 256, 0, ctx->stream>>>(keysA, M, ghist);
    ctx->launches++;
    unsigned long long hh[8 * 256];
    SS_CHECK_CUDA(cudaMemcpyAsync(hh, ghist, sizeof(hh), cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t* kin = keysA;
    uint64_t* kout = keysB;
    uint8_t* lin = labA;
    uint8_t* lout = labB;
    for (int d = 0; d < 8; ++d) {
        bool trivial = false;
        for (int b = 0; b < 256; ++b)
            if (hh[d * 256 + b] == (unsigned long long)M) trivial = true;
        if (trivial) continue;  // every key has the same digit: the pass is the identity
        rs_upsweep_kernel<<<unsigned(nblocks), RS_TPB, 0, ctx->stream>>>(kin, M, 8 * d, hist, nblocks);
        rs_scan_rows_kernel<<<256, 1024, 0, ctx->stream>>>(hist, nblocks, rel, dtotals);
        rs_scan_totals_kernel<<<1, 256, 0, ctx->stream>>>(dtotals, dbase);
        rs_downsweep_kernel<<<unsigned(nblocks), RS_TPB, ds_smem, ctx->stream>>>(kin, lin, M, 8 * d, rel, dbase, nblocks,
                                                                                  kout, lout);
        ctx->launches += 4;
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint8_t* tl = lin; lin = lout; lout = tl;
    }
    SS_CHECK_CUDA(cudaGetLastError());
    *kout_p = kin;
    *lout_p = lin;
    return SS_OK;
}

// scratch of the curve integration over M sorted pairs
struct CurveBufs {
    CurveState* bstate;
    double* partial;
    CurveGlobal* glob;
    CurveState* summary;
    double* dout;
    int64_t cblocks;
};

int32_t curve_bufs(ss_ctx* ctx, int64_t M, CurveBufs* b) {
    using namespace ss;
    void* p;
    b->cblocks = ceil_div(std::max<int64_t>(M, 1), CV_TILE);
    SS_TRY(scratch_get(ctx, 12, size_t(b->cblocks) * sizeof(CurveState) + size_t(2 * b->cblocks) * 8 + 256, &p));
    b->bstate = static_cast<CurveState*>(p);
    b->partial = reinterpret_cast<double*>(b->bstate + b->cblocks);
    b->glob = reinterpret_cast<CurveGlobal*>(b->partial + 2 * b->cblocks);
    b->summary = reinterpret_cast<CurveState*>(b->glob + 1);
    b->dout = reinterpret_cast<double*>(b->summary + 1);
    return SS_OK;
}

int32_t sort_and_integrate(ss_ctx* ctx, uint64_t* keysA, uint8_t* labA, uint64_t* keysB, uint8_t* labB, int64_t M,
                           double* out2) {
    using namespace ss;
    uint64_t* kin;
    uint8_t* lin;
    SS_TRY(sort_pairs(ctx, keysA, labA, keysB, labB, M, &kin, &lin));
    // curve integration over (kin, lin)
    CurveBufs b;
    SS_TRY(curve_bufs(ctx, M, &b));
    curve_kernel<false><<<unsigned(b.cblocks), CV_TPB, 0, ctx->stream>>>(kin, lin, M, b.bstate, nullptr, nullptr, b.cblocks);
    curve_scan_kernel<<<1, 1024, 0, ctx->stream>>>(b.bstate, b.cblocks, M, b.glob, b.summary);
    curve_kernel<true><<<unsigned(b.cblocks), CV_TPB, 0, ctx->stream>>>(kin, lin, M, b.bstate, b.glob, b.partial, b.cblocks);
    curve_final_kernel<<<1, 1024, 0, ctx->stream>>>(b.partial, b.cblocks, b.dout, 0);
    ctx->launches += 4;
    SS_CHECK_CUDA(cudaGetLastError());
    SS_CHECK_CUDA(cudaMemcpyAsync(out2, b.dout, 16, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t alloc_sort_buffers(ss_ctx* ctx, int64_t M, uint64_t** kA, uint64_t** kB, uint8_t** lA, uint8_t** lB) {
    using namespace ss;
    void* p;
    const size_t mk = size_t(round_up(M, 64));
    SS_TRY(scratch_get(ctx, 9, mk * 8 * 2, &p));
    *kA = static_cast<uint64_t*>(p);
    *kB = *kA + mk;
    SS_TRY(scratch_get(ctx, 10, mk * 2, &p));
    *lA = static_cast<uint8_t*>(p);
    *lB = *lA + mk;
    return SS_OK;
}

}  // namespace

namespace ss {

int32_t launch_topl(ss_ctx* ctx, const double* R, int64_t rows, int64_t cols, int64_t ld, int L, int32_t* idx_out,
                    double* val_out, int64_t ldv) {
    if (rows == 0) return SS_OK;
    if (L <= 32) {
        // 32 x 256 tiles, two blocks per SM (measured: 128- and 64-column tiles with 3 / 4 blocks are 12 % / 25 % slower)
        constexpr int WT_COLS = 256;
        const size_t wsm = size_t(WT_ROWS) * (WT_COLS + 1) * 8;
        SS_CHECK_CUDA(cudaFuncSetAttribute(topl_warp_kernel<WT_COLS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(wsm)));
        topl_warp_kernel<WT_COLS, 2><<<unsigned(ceil_div(rows, WT_ROWS)), WT_TPB, wsm, ctx->stream>>>(R, rows, cols, ld, L, idx_out,
                                                                                                   val_out, ldv);
        SS_CHECK_CUDA(cudaGetLastError());
        ctx->launches++;
        return SS_OK;
    }
    const size_t smem = size_t(L) * TOPL_TPB * (8 + 8 + 4);
    SS_REQUIRE(smem <= 200 * 1024, "top-L: L = %d is too large (max %d)", L, int(200 * 1024 / (TOPL_TPB * 20)));
    if (smem > 48 * 1024)
        SS_CHECK_CUDA(cudaFuncSetAttribute(topl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    topl_kernel<<<unsigned(ceil_div(rows, TOPL_TPB)), TOPL_TPB, smem, ctx->stream>>>(R, rows, cols, ld, L, idx_out,
                                                                                     val_out, ldv);
    SS_CHECK_CUDA(cudaGetLastError());
    ctx->launches++;
    return SS_OK;
}

int32_t atl(ss_ctx* ctx, const ss_mat* Y, const ss_mat* R, int L, double* out2) {
    const int64_t rows = R->rows;
    if (rows == 0) {
        out2[0] = out2[1] = __builtin_nan("");
        return SS_OK;
    }
    void* p;
    SS_TRY(scratch_get(ctx, 13, size_t(rows) * L * 4 + size_t(rows) * 16 + 64, &p));
    double* rec = static_cast<double*>(p);
    double* prec = rec + rows;
    double* dout = prec + rows;
    int32_t* idx = reinterpret_cast<int32_t*>(dout + 2);
    SS_TRY(launch_topl(ctx, R->d, rows, R->cols, R->ld, L, idx, nullptr, 0));
    atl_rows_kernel<<<unsigned(ceil_div(rows, ATL_TPB)), ATL_TPB, 0, ctx->stream>>>(Y->d, rows, Y->cols, Y->ld, L, idx,
                                                                                    rec, prec);
    mean2_kernel<<<1, 1024, 0, ctx->stream>>>(rec, prec, rows, dout);
    ctx->launches += 2;
    SS_CHECK_CUDA(cudaGetLastError());
    SS_CHECK_CUDA(cudaMemcpyAsync(out2, dout, 16, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t auroc_auprc(ss_ctx* ctx, const uint8_t* labels, const double* scores, int64_t M, double* out2) {
    if (M <= 1) {  // trapz of a curve with <= 1 point
        out2[0] = out2[1] = 0.0;
        return SS_OK;
    }
    uint64_t *kA, *kB;
    uint8_t *lA, *lB;
    SS_TRY(alloc_sort_buffers(ctx, M, &kA, &kB, &lA, &lB));
    int grid = int(ceil_div(M, 256 * 8));
    if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    make_keys_kernel<<<grid, 256, 0, ctx->stream>>>(scores, labels, M, kA, lA);
    ctx->launches++;
    return sort_and_integrate(ctx, kA, lA, kB, lB, M, out2);
}

// max / mean / std (corrected) of a confusion-matrix metric over all unique thresholds
int32_t threshold_sweep(ss_ctx* ctx, const ss_mat* Y, const ss_mat* R, int metric, double* out4) {
    const int64_t M = R->rows * R->cols;
    const double nan = __builtin_nan("");
    if (M == 0) {
        out4[0] = out4[1] = out4[2] = nan;
        out4[3] = 0;
        return SS_OK;
    }
    uint64_t *kA, *kB, *kin;
    uint8_t *lA, *lB, *lin;
    SS_TRY(alloc_sort_buffers(ctx, M, &kA, &kB, &lA, &lB));
    const int64_t gx = ceil_div(R->rows, 256);
    int64_t gy = ceil_div(int64_t(ctx->sm_count) * 16, gx);
    if (gy > R->cols) gy = R->cols;
    if (gy > 65535) gy = 65535;
    if (gy < 1) gy = 1;
    dim3 grid{unsigned(gx), unsigned(gy)};
    make_keys_mat_kernel<<<grid, 256, 0, ctx->stream>>>(R->d, R->ld, Y->d, Y->ld, R->rows, R->cols, kA, lA);
    ctx->launches++;
    SS_TRY(sort_pairs(ctx, kA, lA, kB, lB, M, &kin, &lin));
    const int64_t cblocks = ceil_div(M, CV_TILE);
    void* p;
    SS_TRY(scratch_get(ctx, 12, size_t(cblocks) * sizeof(CurveState) + size_t(4 * cblocks) * 8 + 256, &p));
    CurveState* bstate = static_cast<CurveState*>(p);
    double* partial = reinterpret_cast<double*>(bstate + cblocks);
    CurveGlobal* glob = reinterpret_cast<CurveGlobal*>(partial + 4 * cblocks);
    CurveState* summary = reinterpret_cast<CurveState*>(glob + 1);
    const unsigned long long* totals = &glob->P;  // positives of the whole list
    double* dout = reinterpret_cast<double*>(summary + 1);
    curve_kernel<false><<<unsigned(cblocks), CV_TPB, 0, ctx->stream>>>(kin, lin, M, bstate, nullptr, nullptr, cblocks);
    curve_scan_kernel<<<1, 1024, 0, ctx->stream>>>(bstate, cblocks, M, glob, summary);
    sweep_kernel<0><<<unsigned(cblocks), CV_TPB, 0, ctx->stream>>>(kin, lin, M, bstate, totals, metric, 0.0, partial, cblocks);
    sweep_final_kernel<<<1, 1024, 0, ctx->stream>>>(partial, cblocks, dout);
    ctx->launches += 4;
    double h[4];
    SS_CHECK_CUDA(cudaMemcpyAsync(h, dout, 32, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    const double n = h[2];
    const bool has_nan = h[3] > 0;
    const double mean = has_nan ? nan : h[0] / n;
    out4[0] = has_nan ? nan : h[1];  // Julia's maximum / mean propagate NaN
    out4[1] = mean;
    out4[3] = n;
    if (has_nan || n < 2) {
        out4[2] = nan;
        return SS_OK;
    }
    sweep_kernel<1><<<unsigned(cblocks), CV_TPB, 0, ctx->stream>>>(kin, lin, M, bstate, totals, metric, mean, partial, cblocks);
    sweep_final_kernel<<<1, 1024, 0, ctx->stream>>>(partial, cblocks, dout);
    ctx->launches += 2;
    SS_CHECK_CUDA(cudaMemcpyAsync(h, dout, 32, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    out4[2] = sqrt(h[0] / (n - 1.0));
    return SS_OK;
}

int32_t bedroc(ss_ctx* ctx, const ss_mat* Y, const ss_mat* R, int rev, double alpha, double* out) {
    const int64_t M = R->rows * R->cols;
    if (M == 0) {
        *out = __builtin_nan("");
        return SS_OK;
    }
    uint64_t *kA, *kB, *kin;
    uint8_t *lA, *lB, *lin;
    SS_TRY(alloc_sort_buffers(ctx, M, &kA, &kB, &lA, &lB));
    const int64_t gx = ceil_div(R->rows, 256);
    int64_t gy = ceil_div(int64_t(ctx->sm_count) * 16, gx);
    if (gy > R->cols) gy = R->cols;
    if (gy > 65535) gy = 65535;
    if (gy < 1) gy = 1;
    dim3 grid{unsigned(gx), unsigned(gy)};
    make_keys_mat_inv_kernel<<<grid, 256, 0, ctx->stream>>>(R->d, R->ld, Y->d, Y->ld, R->rows, R->cols, rev ? 1 : 0, kA, lA);
    ctx->launches++;
    if (M > 1) SS_TRY(sort_pairs(ctx, kA, lA, kB, lB, M, &kin, &lin));
    else { kin = kA; lin = lA; }
    int nb = int(ceil_div(M, 256 * 16));
    if (nb > 1024) nb = 1024;
    void* p;
    SS_TRY(scratch_get(ctx, 12, size_t(nb) * 16 + 64, &p));
    double* partial = static_cast<double*>(p);
    unsigned long long* npart = reinterpret_cast<unsigned long long*>(partial + nb);
    bedroc_sum_kernel<<<nb, 256, 0, ctx->stream>>>(lin, M, alpha, partial, npart);
    ctx->launches++;
    std::vector<double> hs(nb);
    std::vector<unsigned long long> hn(nb);
    SS_CHECK_CUDA(cudaMemcpyAsync(hs.data(), partial, size_t(nb) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaMemcpyAsync(hn.data(), npart, size_t(nb) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    double ssum = 0.0;
    unsigned long long n = 0;
    for (int i = 0; i < nb; ++i) {
        ssum += hs[i];
        n += hn[i];
    }
    // closed-form normalisation, reference src/performance.jl:32-37
    const double Nn = double(M), Ra = double(n) / Nn;
    const double rand_sum = Ra * (1.0 - exp(-alpha)) / (exp(alpha / Nn) - 1.0);
    const double fac = Ra * sinh(alpha / 2.0) / (cosh(alpha / 2.0) - cosh(alpha / 2.0 - alpha * Ra));
    const double cte = 1.0 / (1.0 - exp(alpha * (1.0 - Ra)));
    *out = ssum * fac / rand_sum + cte;
    return SS_OK;
}

// ---- AuROC / AuPRC of a list that is spread over several GPUs -----------------------------------
// Every rank sorts what it holds, the ranks agree on key splitters and exchange the pairs so that rank r owns one
// contiguous key range of the global order (host side: simspread.jl_b200/sharded.py), then each rank integrates
// the trapezoids of ITS range given what lies below it; the signed partial areas add up to the global ones.

// (scores, labels) -> ascending (key, label) arrays in the context's sort buffers (valid until the next metric call)
int32_t auc_sort(ss_ctx* ctx, const uint8_t* labels, const double* scores, const uint64_t* keys_in, int64_t M,
                 uint64_t** keys_out, uint8_t** labels_out) {
    uint64_t *kA, *kB;
    uint8_t *lA, *lB;
    SS_TRY(alloc_sort_buffers(ctx, std::max<int64_t>(M, 1), &kA, &kB, &lA, &lB));
    *keys_out = kA;
    *labels_out = lA;
    if (M == 0) return SS_OK;
    if (keys_in) {  // pairs that already carry keys (received from the peers)
        SS_CHECK_CUDA(cudaMemcpyAsync(kA, keys_in, size_t(M) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        SS_CHECK_CUDA(cudaMemcpyAsync(lA, labels, size_t(M), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        int grid = int(ceil_div(M, 256 * 8));
        if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
        make_keys_kernel<<<grid, 256, 0, ctx->stream>>>(scores, labels, M, kA, lA);
        ctx->launches++;
    }
    SS_TRY(sort_pairs(ctx, kA, lA, kB, lB, M, keys_out, labels_out));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t auc_lower_bound(ss_ctx* ctx, const uint64_t* keys_sorted, int64_t M, const uint64_t* query_host, int nq,
                        int64_t* pos_host) {
    if (nq == 0) return SS_OK;
    void* p;
    SS_TRY(scratch_get(ctx, 12, size_t(nq) * 16 + 64, &p));
    uint64_t* dq = static_cast<uint64_t*>(p);
    long long* dp = reinterpret_cast<long long*>(dq + nq);
    SS_CHECK_CUDA(cudaMemcpyAsync(dq, query_host, size_t(nq) * 8, cudaMemcpyHostToDevice, ctx->stream));
    lower_bound_kernel<<<unsigned(ceil_div(nq, 64)), 64, 0, ctx->stream>>>(keys_sorted, M, dq, nq, dp);
    ctx->launches++;
    SS_CHECK_CUDA(cudaGetLastError());
    SS_CHECK_CUDA(cudaMemcpyAsync(pos_host, dp, size_t(nq) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

// summary3 = {positives, index of the last run start (-1: empty), positives before that run start} of a sorted array
int32_t auc_segment_summary(ss_ctx* ctx, const uint64_t* keys, const uint8_t* lab, int64_t M, int64_t* summary3) {
    if (M == 0) {
        summary3[0] = 0;
        summary3[1] = -1;
        summary3[2] = 0;
        return SS_OK;
    }
    CurveBufs b;
    SS_TRY(curve_bufs(ctx, M, &b));
    curve_kernel<false><<<unsigned(b.cblocks), CV_TPB, 0, ctx->stream>>>(keys, lab, M, b.bstate, nullptr, nullptr, b.cblocks);
    curve_scan_kernel<<<1, 1024, 0, ctx->stream>>>(b.bstate, b.cblocks, M, b.glob, b.summary);
    ctx->launches += 2;
    SS_CHECK_CUDA(cudaGetLastError());
    CurveState h;
    SS_CHECK_CUDA(cudaMemcpyAsync(&h, b.summary, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    summary3[0] = int64_t(h.pos);
    summary3[1] = h.start;
    summary3[2] = int64_t(h.startpos);
    return SS_OK;
}

// signed trapezoid sums of the run boundaries inside this segment, including the joint to the segment below.
// global6 = {P, Mtot, idx0, init.pos, init.start (global index, -1: nothing below), init.startpos}; must follow
// auc_segment_summary on the same arrays (re-uses its block scan).
int32_t auc_segment_integrate(ss_ctx* ctx, const uint64_t* keys, const uint8_t* lab, int64_t M, const int64_t* global6,
                              double* out2) {
    out2[0] = out2[1] = 0.0;
    if (M == 0) return SS_OK;
    CurveBufs b;
    SS_TRY(curve_bufs(ctx, M, &b));
    CurveGlobal g;
    g.P = (unsigned long long)global6[0];
    g.Mtot = (unsigned long long)global6[1];
    g.idx0 = global6[2];
    g.init.pos = (unsigned long long)global6[3];
    g.init.start = global6[4];
    g.init.startpos = (unsigned long long)global6[5];
    SS_CHECK_CUDA(cudaMemcpyAsync(b.glob, &g, sizeof(g), cudaMemcpyHostToDevice, ctx->stream));
    curve_kernel<true><<<unsigned(b.cblocks), CV_TPB, 0, ctx->stream>>>(keys, lab, M, b.bstate, b.glob, b.partial, b.cblocks);
    curve_final_kernel<<<1, 1024, 0, ctx->stream>>>(b.partial, b.cblocks, b.dout, 1);
    ctx->launches += 2;
    SS_CHECK_CUDA(cudaGetLastError());
    SS_CHECK_CUDA(cudaMemcpyAsync(out2, b.dout, 16, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t auroc_auprc_mat(ss_ctx* ctx, const ss_mat* Y, const ss_mat* R, double* out2) {
    const int64_t M = R->rows * R->cols;
    if (M <= 1) {
        out2[0] = out2[1] = 0.0;
        return SS_OK;
    }
    uint64_t *kA, *kB;
    uint8_t *lA, *lB;
    SS_TRY(alloc_sort_buffers(ctx, M, &kA, &kB, &lA, &lB));
    const int64_t gx = ceil_div(R->rows, 256);
    int64_t gy = ceil_div(int64_t(ctx->sm_count) * 16, gx);
    if (gy > R->cols) gy = R->cols;
    if (gy > 65535) gy = 65535;
    if (gy < 1) gy = 1;
    dim3 grid{unsigned(gx), unsigned(gy)};
    make_keys_mat_kernel<<<grid, 256, 0, ctx->stream>>>(R->d, R->ld, Y->d, Y->ld, R->rows, R->cols, kA, lA);
    ctx->launches++;
    return sort_and_integrate(ctx, kA, lA, kB, lB, M, out2);
}

}  // namespace ss
