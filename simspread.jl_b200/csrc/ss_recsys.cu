// Sparse 2-layer NBI with fused top-L (BASELINE config 5, SURVEY.md 8a row a10 without a feature
// layer): for every source (user) s
//
//     F[s,t] = sum_{t'} A[s,t'] * (W^2)[t',t],  (W^2)[t',t] = sum_{s'} (Y[s',t']/kt[t']) * (Y[s',t]/ks[s'])
//
// (reference `predict(A, ytrain)`, src/core.jl:446-466, on the graph [0 Y; Y' 0]), reduced on the fly to
// the L best targets of each source in `sortperm(rev=true)` order (src/performance.jl:315).  Neither F
// (10^12 scores at 2M x 500k) nor the item x item transfer matrix is ever materialised: the product is
// expanded as two hops over the CSR of Y and the CSR of Y' (Gustavson, row-split by source).
//
//   expand : a group of G sources owns G dense FP64 accumulator rows that stay L2-resident
//            (G * Nt * 8 B <= 64 MB); one block per (source, item t'); a warp fetches 32 co-raters s'
//            at once, then lanes run over the items of each s' -> red.global.add.f64 into acc[g][t]
//   extract: every warp scans a column chunk of one accumulator row with a warp-distributed sorted
//            list (ballot filter + shuffle insert), emits its L candidates and zeroes the chunk
//   merge  : one block per source merges the chunk candidates under the composite order
//            (score descending, column ascending) into the final top-L
// Work unit: one partial product; algorithmic bytes: 4 B (column index) per partial product, or
// 12 B for a weighted graph.
#include "ss_common.cuh"

namespace {

__device__ __forceinline__ uint64_t rk_isless_key(double v) {
    if (v != v) return 0xFFFFFFFFFFFFFFFFull;
    const uint64_t b = uint64_t(__double_as_longlong(v));
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double rk_key_to_value(uint64_t k) {
    if (k == 0xFFFFFFFFFFFFFFFFull) return __longlong_as_double(0x7ff8000000000000ll);
    const uint64_t b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

struct RecParams {
    const int32_t* y_ptr;   // CSR of Y (sources x targets)
    const int32_t* y_idx;
    const double* y_val;    // null: binary
    const int32_t* yt_ptr;  // CSR of Y' (targets x sources)
    const int32_t* yt_idx;
    const double* yt_val;
    int64_t ns, nt;
    int64_t ldacc;          // accumulator row pitch (>= nt)
    double* acc;            // [G][ldacc], zero on entry and on exit
    int64_t s0;             // first source of the group
    int G;
};

constexpr int EX_TPB = 256;
constexpr int EX_SLICES = 64;  // blocks per source: about one item t' each at ~50 items per source

__global__ void __launch_bounds__(EX_TPB) rec_expand_kernel(const RecParams p) {
    const int g = blockIdx.x;
    const int64_t s = p.s0 + g;
    if (s >= p.ns) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* acc = p.acc + int64_t(g) * p.ldacc;
    const int32_t b0 = p.y_ptr[s], b1 = p.y_ptr[s + 1];
    for (int32_t j = b0 + blockIdx.y; j < b1; j += gridDim.y) {
        const int32_t tp = p.y_idx[j];
        const double a = p.y_val ? p.y_val[j] : 1.0;  // A[s,t']
        const int32_t u0 = p.yt_ptr[tp], u1 = p.yt_ptr[tp + 1];
        const double kt = double(u1 - u0);  // degree of t' = non-zeros of its row in Y'
        // each warp takes 32 co-raters s' at a time: their ids / row extents are fetched by the 32 lanes
        // in parallel (one dependent-load chain per batch instead of one per co-rater)
        for (int32_t ub = u0 + warp * 32; ub < u1; ub += (EX_TPB / 32) * 32) {
            const int32_t u = ub + lane;
            int32_t r0 = 0, r1 = 0;
            double c = 0.0;
            if (u < u1) {
                const int32_t sp = p.yt_idx[u];
                r0 = p.y_ptr[sp];
                r1 = p.y_ptr[sp + 1];
                c = a * ((p.yt_val ? p.yt_val[u] : 1.0) / kt);  // A[s,t'] * W[t',s'] (true division)
            }
            const int nb = min(32, u1 - ub);
            for (int l = 0; l < nb; ++l) {
                const int32_t q0 = __shfl_sync(0xffffffffu, r0, l), q1 = __shfl_sync(0xffffffffu, r1, l);
                const double cl = __shfl_sync(0xffffffffu, c, l);
                const double ks = double(q1 - q0);
                for (int32_t e = q0 + lane; e < q1; e += 32) {
                    const double w_st = (p.y_val ? p.y_val[e] : 1.0) / ks;  // W[s',t]
                    atomicAdd(acc + p.y_idx[e], cl * w_st);
                }
            }
        }
    }
}

// candidates: cand_key / cand_idx [G][nchunk_warps][32] (lane l = rank l; unused ranks: idx = -1)
__global__ void __launch_bounds__(256)
    rec_extract_kernel(const RecParams p, int L, int64_t cols_per_warp, int nwarp_chunks, uint64_t* __restrict__ cand_key,
                       int32_t* __restrict__ cand_idx) {
    const int g = blockIdx.y;
    if (p.s0 + g >= p.ns) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wc = blockIdx.x * 8 + warp;  // warp chunk index
    if (wc >= nwarp_chunks) return;
    double* acc = p.acc + int64_t(g) * p.ldacc;
    const int64_t c0 = int64_t(wc) * cols_per_warp;
    const int64_t c1 = min(p.nt, c0 + cols_per_warp);
    uint64_t lkey = 0;
    int32_t lidx = -1;
    int cnt = 0;
    uint64_t thr = 0;
    constexpr int UNR = 4;  // 4 independent 256-byte loads in flight per warp
    for (int64_t cb0 = c0; cb0 < c1; cb0 += 32 * UNR) {
        double v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t c = cb0 + 32 * u + lane;
            v[u] = (c < c1) ? acc[c] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t c = cb0 + 32 * u + lane;
            if (c < c1) acc[c] = 0.0;  // leave the accumulator clean for the next group
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t cb = cb0 + 32 * u;
            const bool inb = cb + lane < c1;
            const uint64_t key = inb ? rk_isless_key(v[u]) : 0;
            unsigned cd = __ballot_sync(0xffffffffu, inb && (cnt < L || key > thr));
            if (!cd) continue;
            while (cd) {
                const int src = __ffs(cd) - 1;
                cd &= cd - 1;
                const uint64_t ck = __shfl_sync(0xffffffffu, key, src);
                const int32_t ci = int32_t(cb + src);
                const bool ge = (lane < cnt) && (lkey >= ck);
                const int pos = __popc(__ballot_sync(0xffffffffu, ge));
                if (pos >= L) continue;
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
            thr = __shfl_sync(0xffffffffu, lkey, L - 1);
        }
    }
    const int64_t o = (int64_t(g) * nwarp_chunks + wc) * 32 + lane;
    cand_key[o] = (lane < cnt) ? lkey : 0;
    cand_idx[o] = (lane < cnt) ? lidx : -1;
}

// ties must resolve to the lower column: a candidate wins over a list entry with the same key iff its
// column is lower.  Chunks are visited out of order here (8 warps, strided), so the rank uses the
// composite (key desc, column asc) order explicitly.
__device__ __forceinline__ void list_insert_ordered(uint64_t& lkey, int32_t& lidx, int& cnt, int L, uint64_t ck,
                                                    int32_t ci, int lane) {
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

__global__ void __launch_bounds__(256)
    rec_merge_kernel(const RecParams p, int L, int nwarp_chunks, const uint64_t* __restrict__ cand_key,
                     const int32_t* __restrict__ cand_idx, int32_t* __restrict__ idx_out, double* __restrict__ val_out) {
    __shared__ uint64_t skey[8][32];
    __shared__ int32_t sidx[8][32];
    const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t s = p.s0 + g;
    if (s >= p.ns) return;
    uint64_t lkey = 0;
    int32_t lidx = -1;
    int cnt = 0;
    for (int wc = warp; wc < nwarp_chunks; wc += 8) {
        const int64_t o = (int64_t(g) * nwarp_chunks + wc) * 32 + lane;
        const uint64_t key = cand_key[o];
        const int32_t idx = cand_idx[o];
        unsigned cd = __ballot_sync(0xffffffffu, idx >= 0);
        while (cd) {
            const int src = __ffs(cd) - 1;
            cd &= cd - 1;
            list_insert_ordered(lkey, lidx, cnt, L, __shfl_sync(0xffffffffu, key, src),
                                __shfl_sync(0xffffffffu, idx, src), lane);
        }
    }
    skey[warp][lane] = (lane < cnt) ? lkey : 0;
    sidx[warp][lane] = (lane < cnt) ? lidx : -1;
    __syncthreads();
    if (warp != 0) return;
    for (int w = 1; w < 8; ++w) {
        const uint64_t key = skey[w][lane];
        const int32_t idx = sidx[w][lane];
        unsigned cd = __ballot_sync(0xffffffffu, idx >= 0);
        while (cd) {
            const int src = __ffs(cd) - 1;
            cd &= cd - 1;
            list_insert_ordered(lkey, lidx, cnt, L, __shfl_sync(0xffffffffu, key, src),
                                __shfl_sync(0xffffffffu, idx, src), lane);
        }
    }
    if (lane < L) {
        idx_out[s * L + lane] = (lane < cnt) ? lidx : -1;
        if (val_out) val_out[s * L + lane] = (lane < cnt) ? rk_key_to_value(lkey) : 0.0;
    }
}

}  // namespace

namespace ss {

// top-L targets of every source of the 2-layer graph Y (CSR) / Y' (CSR), sources [s_begin, s_end)
int32_t recommend_topl(ss_ctx* ctx, const ss_csr* Y, const ss_csr* YT, int L, int64_t s_begin, int64_t s_end,
                       int32_t* idx_out, double* val_out) {
    const int64_t ns = Y->rows, nt = Y->cols;
    if (s_end <= s_begin || nt == 0) return SS_OK;
    const int64_t ldacc = round_up(nt, 32);
    int64_t G = (int64_t(32) << 20) / (ldacc * 8);  // 3 concurrent groups x 32 MB of accumulators stay in L2
    if (G < 1) G = 1;
    if (G > 64) G = 64;
    if (G > s_end - s_begin) G = s_end - s_begin;
    // warp chunks: enough warps to fill the GPU for one group, at least 1024 columns each
    int64_t nwc = ceil_div(int64_t(ctx->sm_count) * 8 * 2, G);
    if (nwc > ceil_div(nt, 1024)) nwc = ceil_div(nt, 1024);
    if (nwc < 1) nwc = 1;
    const int64_t cpw = round_up(ceil_div(nt, nwc), 32);
    nwc = ceil_div(nt, cpw);
    // The three phases of a group are latency-bound, so consecutive groups run on three streams with
    // their own accumulators / candidate buffers and overlap each other.
    constexpr int NS = 3;
    cudaStream_t streams[NS] = {ctx->stream, ctx->copy_in, ctx->copy_out};
    const size_t per = size_t(G) * ldacc * 8 + size_t(G) * nwc * 32 * 12 + 256;
    void* p;
    SS_TRY(scratch_get(ctx, 13, per * NS, &p));
    cudaEvent_t ready, done[NS];
    SS_CHECK_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    SS_CHECK_CUDA(cudaEventRecord(ready, ctx->stream));
    RecParams q{};
    q.y_ptr = Y->row_ptr;
    q.y_idx = Y->col_idx;
    q.y_val = Y->values;
    q.yt_ptr = YT->row_ptr;
    q.yt_idx = YT->col_idx;
    q.yt_val = YT->values;
    q.ns = s_end;  // sources beyond the requested range are skipped by the kernels
    q.nt = nt;
    q.ldacc = ldacc;
    double* accs[NS];
    uint64_t* ckeys[NS];
    int32_t* cidxs[NS];
    for (int i = 0; i < NS; ++i) {
        accs[i] = reinterpret_cast<double*>(static_cast<char*>(p) + per * i);
        ckeys[i] = reinterpret_cast<uint64_t*>(accs[i] + G * ldacc);
        cidxs[i] = reinterpret_cast<int32_t*>(ckeys[i] + G * nwc * 32);
        if (i) SS_CHECK_CUDA(cudaStreamWaitEvent(streams[i], ready, 0));
        SS_CHECK_CUDA(cudaMemsetAsync(accs[i], 0, size_t(G) * ldacc * 8, streams[i]));
    }
    int64_t gi = 0;
    for (int64_t s0 = s_begin; s0 < s_end; s0 += G, ++gi) {
        const int b = int(gi % NS);
        q.s0 = s0;
        q.G = int(s_end - s0 < G ? s_end - s0 : G);
        q.acc = accs[b];
        rec_expand_kernel<<<dim3(unsigned(q.G), EX_SLICES), EX_TPB, 0, streams[b]>>>(q);
        rec_extract_kernel<<<dim3(unsigned(ceil_div(nwc, 8)), unsigned(q.G)), 256, 0, streams[b]>>>(q, L, cpw, int(nwc),
                                                                                                  ckeys[b], cidxs[b]);
        rec_merge_kernel<<<unsigned(q.G), 256, 0, streams[b]>>>(q, L, int(nwc), ckeys[b], cidxs[b], idx_out, val_out);
        ctx->launches += 3;
    }
    for (int i = 1; i < NS; ++i) {  // join the helper streams back into the context stream
        SS_CHECK_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        SS_CHECK_CUDA(cudaEventRecord(done[i], streams[i]));
        SS_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, done[i], 0));
        SS_CHECK_CUDA(cudaEventDestroy(done[i]));
    }
    SS_CHECK_CUDA(cudaEventDestroy(ready));
    SS_CHECK_CUDA(cudaGetLastError());
    return SS_OK;
}

}  // namespace ss
