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

// single-block exclusive scan of n uint32 counts (n = 256 * nblocks), 64-bit running sum stored as
// uint64 offsets
__global__ void __launch_bounds__(1024)
    rs_scan_kernel(const uint32_t* __restrict__ in, int64_t n, uint64_t* __restrict__ out) {
    __shared__ unsigned long long part[1024];
    const int t = threadIdx.x;
    const int64_t chunk = (n + 1023) / 1024;
    const int64_t b = t * chunk, e = min(n, b + chunk);
    unsigned long long s = 0;
    for (int64_t i = b; i < e; ++i) s += in[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; ++i) {
            const unsigned long long v = part[i];
            part[i] = run;
            run += v;
        }
    }
    __syncthreads();
    unsigned long long run = part[t];
    for (int64_t i = b; i < e; ++i) {
        out[i] = run;
        run += in[i];
    }
}

// stable scatter: rank inside the warp by __match_any_sync, across warps by a per-digit prefix
__global__ void __launch_bounds__(RS_TPB)
    rs_downsweep_kernel(const uint64_t* __restrict__ keys_in, const uint8_t* __restrict__ lab_in, int64_t M, int shift,
                        const uint64_t* __restrict__ offsets, int64_t nblocks, uint64_t* __restrict__ keys_out,
                        uint8_t* __restrict__ lab_out) {
    __shared__ unsigned int wcount[RS_TPB / 32][256];
    __shared__ unsigned long long dbase[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (RS_TPB / 32) * 256; i += RS_TPB) (&wcount[0][0])[i] = 0;
    __syncthreads();
    // order inside the tile: (warp, item, lane) == ascending index
    const int64_t wbase = int64_t(blockIdx.x) * RS_TILE + int64_t(warp) * (32 * RS_ITEMS);
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
        dbase[d] = offsets[int64_t(d) * nblocks + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int64_t idx = wbase + i * 32 + lane;
        if (idx < M) {
            const unsigned int d = (unsigned int)((key[i] >> shift) & 255);
            const unsigned long long pos = dbase[d] + wcount[warp][d] + rank[i];
            keys_out[pos] = key[i];
            lab_out[pos] = lab[i];
        }
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

// EMIT == false: write the block aggregate.  EMIT == true: use the scanned block prefixes and add
// one trapezoid per run boundary into partial[block] (roc) / partial[nblocks + block] (pr).
template <bool EMIT>
__global__ void __launch_bounds__(CV_TPB)
    curve_kernel(const uint64_t* __restrict__ keys, const uint8_t* __restrict__ lab, int64_t M,
                 CurveState* __restrict__ block_state, const unsigned long long* __restrict__ totals,
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
    CurveState run = cs_combine(block_state[blockIdx.x], excl);  // exclusive prefix over the whole array
    const double P = double(totals[0]);
    const double N = double((unsigned long long)M - totals[0]);
    double roc = 0.0, pr = 0.0;
#pragma unroll
    for (int i = 0; i < CV_ITEMS; ++i) {
        const int64_t idx = base + i;
        if (idx < M) {
            const bool is_start = (idx == 0) || (k[i + 1] != k[i]);
            if (is_start && idx > 0) {
                // lower threshold = previous run (start p), higher threshold = this run (start idx)
                const double p = double(run.start);
                const double tp1 = P - double(run.startpos), fp1 = N - (p - double(run.startpos));
                const double tp2 = P - double(run.pos), fp2 = N - (double(idx) - double(run.pos));
                roc += (fp2 / N - fp1 / N) * (tp1 / P + tp2 / P) * 0.5;
                pr += (tp2 / P - tp1 / P) * (tp1 / (tp1 + fp1) + tp2 / (tp2 + fp2)) * 0.5;
            }
            CurveState e = {(unsigned long long)l[i], is_start ? (long long)idx : -1ll, 0ull};
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
    curve_scan_kernel(CurveState* __restrict__ st, int64_t n, unsigned long long* __restrict__ totals) {
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
        totals[0] = run.pos;
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
    curve_final_kernel(const double* __restrict__ partial, int64_t nblocks, double* __restrict__ out) {
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
        out[0] = fabs(sa[0]);
        out[1] = fabs(sb[0]);
    }
}

int32_t sort_and_integrate(ss_ctx* ctx, uint64_t* keysA, uint8_t* labA, uint64_t* keysB, uint8_t* labB, int64_t M,
                           double* out2) {
    using namespace ss;
    void* p;
    const int64_t nblocks = ceil_div(M, RS_TILE);
    // scratch: global digit histogram (8*256 u64) | per-pass hist (256*nblocks u32) | offsets (u64)
    SS_TRY(scratch_get(ctx, 11, size_t(8 * 256) * 8 + size_t(256) * nblocks * 4 + size_t(256) * nblocks * 8 + 64, &p));
    unsigned long long* ghist = static_cast<unsigned long long*>(p);
    uint64_t* offsets = reinterpret_cast<uint64_t*>(ghist + 8 * 256);
    uint32_t* hist = reinterpret_cast<uint32_t*>(offsets + 256 * nblocks);
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
        rs_scan_kernel<<<1, 1024, 0, ctx->stream>>>(hist, 256 * nblocks, offsets);
        rs_downsweep_kernel<<<unsigned(nblocks), RS_TPB, 0, ctx->stream>>>(kin, lin, M, 8 * d, offsets, nblocks, kout,
                                                                            lout);
        ctx->launches += 3;
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint8_t* tl = lin; lin = lout; lout = tl;
    }
    SS_CHECK_CUDA(cudaGetLastError());
    // curve integration over (kin, lin)
    const int64_t cblocks = ceil_div(M, CV_TILE);
    SS_TRY(scratch_get(ctx, 12, size_t(cblocks) * sizeof(CurveState) + size_t(2 * cblocks) * 8 + 64, &p));
    CurveState* bstate = static_cast<CurveState*>(p);
    double* partial = reinterpret_cast<double*>(bstate + cblocks);
    unsigned long long* totals = reinterpret_cast<unsigned long long*>(partial + 2 * cblocks);
    double* dout = reinterpret_cast<double*>(totals + 2);
    curve_kernel<false><<<unsigned(cblocks), CV_TPB, 0, ctx->stream>>>(kin, lin, M, bstate, nullptr, nullptr, cblocks);
    curve_scan_kernel<<<1, 1024, 0, ctx->stream>>>(bstate, cblocks, totals);
    curve_kernel<true><<<unsigned(cblocks), CV_TPB, 0, ctx->stream>>>(kin, lin, M, bstate, totals, partial, cblocks);
    curve_final_kernel<<<1, 1024, 0, ctx->stream>>>(partial, cblocks, dout);
    ctx->launches += 4;
    SS_CHECK_CUDA(cudaGetLastError());
    SS_CHECK_CUDA(cudaMemcpyAsync(out2, dout, 16, cudaMemcpyDeviceToHost, ctx->stream));
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
