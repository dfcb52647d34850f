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
#include <cooperative_groups.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "ss_common.cuh"

namespace cg = cooperative_groups;

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
    int64_t s0;             // first source of the group (or first position in `slist`)
    int G;
    const int32_t* slist;   // optional: the group kernels process sources slist[s0 + g] (ns = list length)
};

__device__ __forceinline__ int64_t rec_source(const RecParams& p, int g) { return p.slist ? int64_t(p.slist[p.s0 + g]) : p.s0 + g; }

constexpr int EX_TPB = 256;
constexpr int EX_SLICES = 64;  // blocks per source: about one item t' each at ~50 items per source

__global__ void __launch_bounds__(EX_TPB) rec_expand_kernel(const RecParams p) {
    const int g = blockIdx.x;
    if (p.s0 + g >= p.ns) return;
    const int64_t s = rec_source(p, g);
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
                     const int32_t* __restrict__ cand_idx, int32_t* __restrict__ idx_out, double* __restrict__ val_out,
                     int64_t ldv) {
    __shared__ uint64_t skey[8][32];
    __shared__ int32_t sidx[8][32];
    const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (p.s0 + g >= p.ns) return;
    const int64_t s = rec_source(p, g);
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
        if (val_out) val_out[s * ldv + lane] = (lane < cnt) ? rk_key_to_value(lkey) : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// Fused persistent form: a thread-block cluster of FU_C CTAs owns ONE source at a time and one
// L2-resident accumulator row (clusters x Nt x 8 B <= ~64 MB).  Per source:
//   expand : work units = (item t', batch of 32 co-raters), dealt round-robin to the CTAs of the
//            cluster (the bound is the per-SM RED issue rate, so the split across SMs is what counts);
//            the index loads of 4 co-raters (8 x 128 B) are in flight before their REDs are issued
//   select : every thread finds the maximum of its (interleaved) columns, the L-th largest of the
//            FU_NW warp maxima is a lower bound tau of the L-th best score; only threads whose maximum
//            reaches tau re-read their columns and append the entries >= tau to the candidate list in
//            CTA 0's shared memory (DSMEM atomics); CTA 0 ranks the candidates under (score desc,
//            column asc) and writes the top-L while the other CTAs already expand the next source
//   zero   : every thread clears its own columns (no second kernel, no memset)
// A source whose candidate list overflows (massive ties, e.g. an all-zero row) is put on a redo list
// and goes through the group kernels above.
constexpr int FU_CAP = 2048;            // candidate capacity

struct FusedParams {
    RecParams r;          // r.acc: [clusters][ldacc]
    int64_t s_begin, s_end;
    int L;
    int32_t* idx_out;
    double* val_out;
    int64_t ldv;          // leading dimension of val_out (>= L)
    int32_t* redo_cnt;
    int32_t* redo_list;
};

// order-preserving image of the high word of a double under Julia's isless (sign-magnitude -> two's complement;
// -0.0 < +0.0, +NaN above +Inf)
__device__ __forceinline__ int hi_image(double v) {
    const int hi = __double2hiint(v);
    return hi ^ ((hi >> 31) & 0x7fffffff);
}


// FU_C CTAs per cluster (launch attribute), FU_TPB threads per CTA, 1024 / FU_TPB CTAs per SM: with two
// CTAs of different clusters on one SM the select / zero phases of one source overlap the RED stream of another.
template <bool WEIGHTED, int FU_C, int FU_TPB, int FU_UNIT>  // FU_UNIT: co-raters per work unit
__global__ void __launch_bounds__(FU_TPB, 1024 / FU_TPB) rec_fused_kernel(const FusedParams p) {
    constexpr int FU_WARPS = FU_TPB / 32;
    constexpr int FU_NW = FU_C * FU_WARPS;  // warp maxima per source
    constexpr int FU_MAXI = FU_TPB;         // items of the source staged per pass
    static_assert(FU_NW <= FU_TPB && FU_NW >= 32 && FU_C <= 32, "tau is ranked by the first FU_NW threads");
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ int32_t s_pref[FU_MAXI + 1];
    __shared__ int32_t s_u0[FU_MAXI];
    __shared__ int32_t s_u1[FU_MAXI];
    __shared__ double s_a[FU_MAXI];
    __shared__ int32_t s_wsum[32];
    __shared__ int s_wmax[FU_NW];
    __shared__ int s_tau;
    __shared__ uint64_t s_ckey[FU_CAP];  // candidate list: used in CTA 0 only, written by the whole cluster
    __shared__ int32_t s_ccol[FU_CAP];
    __shared__ int s_ccnt;
    __shared__ int s_next;  // next work unit of this CTA (warps draw units dynamically)

    const RecParams& r = p.r;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = int(cluster.block_rank());
    const int ncl = gridDim.x / FU_C, cl = blockIdx.x / FU_C;
    double* acc = r.acc + int64_t(cl) * r.ldacc;
    const int64_t npairs = r.ldacc >> 1;
    const int L = p.L;
    if (tid == 0) s_ccnt = 0;
    int* ccnt0 = cluster.map_shared_rank(&s_ccnt, 0);
    uint64_t* ckey0 = cluster.map_shared_rank(s_ckey, 0);
    int32_t* ccol0 = cluster.map_shared_rank(s_ccol, 0);
    cluster.sync();

    for (int64_t s = p.s_begin + cl; s < p.s_end; s += ncl) {
        // ---------------- expand ----------------
        const int32_t b0 = __ldg(r.y_ptr + s), b1 = __ldg(r.y_ptr + s + 1);
        for (int32_t chunk = b0; chunk < b1; chunk += FU_MAXI) {
            const int nit = min(FU_MAXI, b1 - chunk);
            int nb = 0;
            if (tid < nit) {
                const int32_t tp = __ldg(r.y_idx + chunk + tid);
                const int32_t u0 = __ldg(r.yt_ptr + tp), u1 = __ldg(r.yt_ptr + tp + 1);
                s_u0[tid] = u0;
                s_u1[tid] = u1;
                s_a[tid] = WEIGHTED ? __ldg(r.y_val + chunk + tid) : 1.0;  // A[s,t']
                nb = (u1 - u0 + FU_UNIT - 1) / FU_UNIT;
            }
            int incl = nb;  // block-wide exclusive scan of the unit counts
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                int w = lane < FU_WARPS ? s_wsum[lane] : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, w, o);
                    if (lane >= o) w += v;
                }
                if (lane < FU_WARPS) s_wsum[lane] = w;  // inclusive over warps
            }
            __syncthreads();
            const int wbase = warp ? s_wsum[warp - 1] : 0;
            if (tid < nit) s_pref[tid] = wbase + incl - nb;
            const int total = s_wsum[FU_WARPS - 1];
            if (tid == 0) {
                s_pref[nit] = total;
                s_next = 0;
            }
            __syncthreads();
            for (;;) {
                int unit = 0;
                if (lane == 0) unit = atomicAdd(&s_next, 1) * FU_C + cta;  // units are dealt round-robin to the CTAs
                unit = __shfl_sync(0xffffffffu, unit, 0);
                if (unit >= total) break;
                int lo = 0, hi = nit - 1;  // largest j with s_pref[j] <= unit
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_pref[mid] <= unit) lo = mid; else hi = mid - 1;
                }
                const int32_t u1 = s_u1[lo];
                const int32_t u0 = s_u0[lo];
                const double a = s_a[lo];
                const double kt = double(u1 - u0);  // degree of t' = non-zeros of its row in Y'
                const int32_t ub = u0 + (unit - s_pref[lo]) * FU_UNIT;
                const int32_t u = ub + lane;
                int32_t r0 = 0, r1 = 0;
                double cw = 0.0;
                if (lane < FU_UNIT && u < u1) {
                    const int32_t sp = __ldcs(r.yt_idx + u);
                    r0 = __ldg(r.y_ptr + sp);
                    r1 = __ldg(r.y_ptr + sp + 1);
                    const double c = a * ((WEIGHTED ? __ldg(r.yt_val + u) : 1.0) / kt);  // A[s,t'] * W[t',s']
                    cw = WEIGHTED ? c : c * (1.0 / double(r1 - r0));                     // binary: * W[s',t]
                }
                const int nbr = min(FU_UNIT, u1 - ub);
                for (int l = 0; l < nbr; l += 4) {
                    int32_t q0[4], q1[4], ia[4], ib[4];
                    double cl4[4], va[4], vb[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int src = (l + k) & 31;
                        q0[k] = __shfl_sync(0xffffffffu, r0, src);
                        q1[k] = __shfl_sync(0xffffffffu, r1, src);
                        cl4[k] = __shfl_sync(0xffffffffu, cw, src);
                        if (l + k >= nbr) q1[k] = q0[k];
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int32_t e = q0[k] + lane;
                        ia[k] = e < q1[k] ? __ldcs(r.y_idx + e) : -1;
                        ib[k] = e + 32 < q1[k] ? __ldcs(r.y_idx + e + 32) : -1;
                        if (WEIGHTED) {
                            va[k] = e < q1[k] ? __ldcs(r.y_val + e) : 0.0;
                            vb[k] = e + 32 < q1[k] ? __ldcs(r.y_val + e + 32) : 0.0;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double ks = double(q1[k] - q0[k]);
                        if (ia[k] >= 0) atomicAdd(acc + ia[k], WEIGHTED ? cl4[k] * (va[k] / ks) : cl4[k]);
                        if (ib[k] >= 0) atomicAdd(acc + ib[k], WEIGHTED ? cl4[k] * (vb[k] / ks) : cl4[k]);
                        for (int32_t e = q0[k] + 64 + lane; e < q1[k]; e += 32)  // sources with more than 64 targets
                            atomicAdd(acc + __ldg(r.y_idx + e), WEIGHTED ? cl4[k] * (__ldg(r.y_val + e) / ks) : cl4[k]);
                    }
                }
            }
            __syncthreads();
        }
        cluster.sync();  // all partial products of this source are in the accumulator row

        // ---------------- select ----------------
        // The threshold works on the order-preserving image of the HIGH word of a score (sign, exponent, 20
        // mantissa bits; 3 integer instructions per column): tau = L-th largest warp maximum is a lower bound of
        // the L-th best score, and every entry whose image reaches tau is a candidate (ranked below by its full key).
        const double2* accp = reinterpret_cast<const double2*>(acc);
        const int64_t stride = int64_t(FU_C) * FU_TPB;
        const int64_t nfull = r.nt >> 1;  // column pairs entirely inside the row
        int hmax = INT_MIN;
        {
            int64_t pi = int64_t(cta) * FU_TPB + tid;
            for (; pi + 3 * stride < nfull; pi += 4 * stride) {  // 4 independent 128-bit loads in flight per thread
                const double2 v0 = __ldcg(accp + pi), v1 = __ldcg(accp + pi + stride);
                const double2 v2 = __ldcg(accp + pi + 2 * stride), v3 = __ldcg(accp + pi + 3 * stride);
                hmax = max(hmax, max(max(hi_image(v0.x), hi_image(v0.y)), max(hi_image(v1.x), hi_image(v1.y))));
                hmax = max(hmax, max(max(hi_image(v2.x), hi_image(v2.y)), max(hi_image(v3.x), hi_image(v3.y))));
            }
            for (; pi < nfull; pi += stride) {
                const double2 v = __ldcg(accp + pi);
                hmax = max(hmax, max(hi_image(v.x), hi_image(v.y)));
            }
            if ((r.nt & 1) && pi == nfull) hmax = max(hmax, hi_image(__ldcg(acc + r.nt - 1)));  // last, odd column
        }
        const int wm = __reduce_max_sync(0xffffffffu, hmax);
        if (lane < FU_C) *cluster.map_shared_rank(&s_wmax[cta * FU_WARPS + warp], lane) = wm;
        cluster.sync();
        if (tid < FU_NW) {  // tau = L-th largest warp maximum (rank by counting; ties broken by position)
            const int my = s_wmax[tid];
            int rank = 0;
#pragma unroll 8
            for (int j = 0; j < FU_NW; ++j) {
                const int o = s_wmax[j];
                rank += (o > my || (o == my && j < tid)) ? 1 : 0;
            }
            if (rank == L - 1) s_tau = my;
        }
        __syncthreads();
        const int tau = s_tau;
        // rare: a thread whose maximum reaches tau holds candidates; the warp re-reads that thread's columns
        // together (one load round per such thread) and appends every entry >= tau to CTA 0's list
        unsigned hot = __ballot_sync(0xffffffffu, hmax != INT_MIN && hmax >= tau);
        while (hot) {
            const int src = __ffs(hot) - 1;
            hot &= hot - 1;
            const int64_t first = int64_t(cta) * FU_TPB + (warp << 5) + src;
            for (int64_t pi = first + int64_t(lane) * stride; pi <= nfull; pi += 32 * stride) {
                const int64_t c = pi << 1;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (c + h >= r.nt) continue;
                    const double v = __ldcg(acc + c + h);
                    if (hi_image(v) >= tau) {
                        const int pos = atomicAdd(ccnt0, 1);
                        if (pos < FU_CAP) {
                            ckey0[pos] = rk_isless_key(v);
                            ccol0[pos] = int32_t(c + h);
                        }
                    }
                }
            }
        }
        // ---------------- zero the row (every thread its own columns) ----------------
        for (int64_t pi = int64_t(cta) * FU_TPB + tid; pi < npairs; pi += int64_t(FU_C) * FU_TPB)
            __stcg(reinterpret_cast<double2*>(acc) + pi, make_double2(0.0, 0.0));
        cluster.sync();  // candidates complete; the row is clean for the next source
        if (cta == 0) {
            const int n = s_ccnt;
            if (n > FU_CAP) {
                if (tid == 0) p.redo_list[atomicAdd(p.redo_cnt, 1)] = int32_t(s);
            } else {
                for (int i = tid; i < n; i += FU_TPB) {
                    const uint64_t my = s_ckey[i];
                    const int32_t mc = s_ccol[i];
                    int rank = 0;
                    for (int j = 0; j < n; ++j) {
                        const uint64_t o = s_ckey[j];
                        rank += (o > my || (o == my && s_ccol[j] < mc)) ? 1 : 0;
                    }
                    if (rank < L) {
                        p.idx_out[s * L + rank] = mc;
                        if (p.val_out) p.val_out[s * p.ldv + rank] = rk_key_to_value(my);
                    }
                }
                if (tid >= n && tid < L) {  // fewer entries than L (cannot happen for L <= targets)
                    p.idx_out[s * L + tid] = -1;
                    if (p.val_out) p.val_out[s * p.ldv + tid] = 0.0;
                }
            }
            __syncthreads();
            if (tid == 0) s_ccnt = 0;
        }
    }
    cluster.sync();  // no CTA exits while its shared memory may still be addressed by a peer
}

}  // namespace

namespace ss {

namespace {

// group kernels (expand / extract / merge) over the sources slist[0..count) or [s_begin, s_end)
int32_t recommend_groups(ss_ctx* ctx, const ss_csr* Y, const ss_csr* YT, int L, int64_t s_begin, int64_t s_end,
                         const int32_t* slist, int32_t* idx_out, double* val_out, int64_t ldv) {
    const int64_t nt = Y->cols;
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
    q.ns = s_end;  // sources (or list positions) beyond the requested range are skipped by the kernels
    q.nt = nt;
    q.ldacc = ldacc;
    q.slist = slist;
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
        rec_merge_kernel<<<unsigned(q.G), 256, 0, streams[b]>>>(q, L, int(nwc), ckeys[b], cidxs[b], idx_out, val_out, ldv);
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

}  // namespace

namespace {

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

template <bool WEIGHTED, int FU_C, int FU_TPB, int FU_UNIT>
int32_t launch_fused(ss_ctx* ctx, FusedParams& fp, int64_t ldacc, int64_t nsrc, int* ncl_out) {
    auto kern = rec_fused_kernel<WEIGHTED, FU_C, FU_TPB, FU_UNIT>;
    if (FU_C > 8) SS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = FU_C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(FU_TPB);
    cfg.gridDim = dim3(unsigned(FU_C));
    cfg.stream = ctx->stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int ncl = 0;  // clusters that can be co-resident
    SS_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
    SS_REQUIRE(ncl >= 1, "ss_recommend_topl: no thread-block cluster of %d CTAs fits on this device", FU_C);
    // ... capped so that the accumulator rows (one per cluster) stay L2-resident
    const int64_t l2_rows = std::max<int64_t>(4, (int64_t(env_int("SS_RECSYS_L2MB", 72)) << 20) / (ldacc * 8));
    if (ncl > l2_rows) ncl = int(l2_rows);
    const int cap = env_int("SS_RECSYS_CLUSTERS", 0);
    if (cap > 0 && ncl > cap) ncl = cap;
    if (ncl > nsrc) ncl = int(nsrc);
    const size_t acc_bytes = size_t(ncl) * ldacc * 8;
    void* w;
    SS_TRY(scratch_get(ctx, 14, acc_bytes + 256 + size_t(nsrc) * 4, &w));
    fp.r.acc = static_cast<double*>(w);
    fp.redo_cnt = reinterpret_cast<int32_t*>(static_cast<char*>(w) + acc_bytes);
    fp.redo_list = fp.redo_cnt + 64;
    SS_CHECK_CUDA(cudaMemsetAsync(w, 0, acc_bytes + 256, ctx->stream));
    cfg.gridDim = dim3(unsigned(ncl * FU_C));
    // SS_RECSYS_L2PERSIST=1: accumulator rows in the persisting part of L2, everything else streaming
    // (measured: no gain on B200, the rows stay resident anyway; kept as an option)
    int max_persist = 0, max_window = 0;
    SS_CHECK_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device));
    SS_CHECK_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device));
    const bool persist = env_int("SS_RECSYS_L2PERSIST", 0) != 0 && max_persist > 0 && max_window > 0;
    if (persist) {
        const size_t win = std::min(acc_bytes, size_t(max_window));
        SS_CHECK_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min(win, size_t(max_persist))));
        attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[1].val.accessPolicyWindow.base_ptr = w;
        attr[1].val.accessPolicyWindow.num_bytes = win;
        attr[1].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.numAttrs = 2;
    }
    SS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, fp));
    ctx->launches += 1;
    if (persist) {
        SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        SS_CHECK_CUDA(cudaCtxResetPersistingL2Cache());
    }
    *ncl_out = ncl;
    return SS_OK;
}

}  // namespace

// top-L targets of every source of the 2-layer graph Y (CSR) / Y' (CSR), sources [s_begin, s_end)
int32_t recommend_topl(ss_ctx* ctx, const ss_csr* Y, const ss_csr* YT, int L, int64_t s_begin, int64_t s_end,
                       int32_t* idx_out, double* val_out, int64_t ldv) {
    const int64_t nt = Y->cols;
    if (s_end <= s_begin || nt == 0) return SS_OK;
    // default: the deterministic shared-memory form over the materialised transfer matrix (ss_transfer.cu).
    // SS_RECSYS_MODE = "atomic": the round-1 cluster kernel below (two-hop expansion, red.global.add.f64 into
    // L2-resident rows; sums in no fixed order); "groups": its three-kernel predecessor.  Both kept for A/B runs.
    // A graph whose transfer matrix cannot be materialised (heavy-tailed degrees: it approaches items x items) is
    // declined by that form and also takes the cluster kernel.
    const char* mode = getenv("SS_RECSYS_MODE");
    if (!mode || (strcmp(mode, "atomic") && strcmp(mode, "groups"))) {
        bool declined = false;
        SS_TRY(recommend_topl_stream(ctx, Y, YT, L, s_begin, s_end, idx_out, val_out, ldv, &declined));
        if (!declined) return SS_OK;
    } else if (!strcmp(mode, "groups")) {
        return recommend_groups(ctx, Y, YT, L, s_begin, s_end, nullptr, idx_out, val_out, ldv);
    }
    const bool weighted = Y->values != nullptr;
    const int64_t ldacc = round_up(nt, 32);
    FusedParams fp{};
    fp.r.y_ptr = Y->row_ptr;
    fp.r.y_idx = Y->col_idx;
    fp.r.y_val = Y->values;
    fp.r.yt_ptr = YT->row_ptr;
    fp.r.yt_idx = YT->col_idx;
    fp.r.yt_val = YT->values;
    fp.r.ns = Y->rows;
    fp.r.nt = nt;
    fp.r.ldacc = ldacc;
    fp.s_begin = s_begin;
    fp.s_end = s_end;
    fp.L = L;
    fp.idx_out = idx_out;
    fp.val_out = val_out;
    fp.ldv = ldv;
    // cluster shape: "8x1024" (default) = 8 CTAs of 1024 threads, one per SM (portable cluster size; 15 clusters =
    // 120 SMs co-resident on B200); "16x512" = 16 CTAs of 512 threads, two CTAs per SM (non-portable size; 14
    // clusters on B200, measured 3 % slower: the kernel is bound by the RED issue rate of the SMs it covers)
    const char* shape = getenv("SS_RECSYS_SHAPE");
    const int unit = env_int("SS_RECSYS_UNIT", 16);
    const int64_t nsrc = s_end - s_begin;
    int ncl = 0;
    int32_t st;
#define SS_FUSED(C_, T_, U_) (weighted ? launch_fused<true, C_, T_, U_>(ctx, fp, ldacc, nsrc, &ncl) : launch_fused<false, C_, T_, U_>(ctx, fp, ldacc, nsrc, &ncl))
    if (shape && !strcmp(shape, "16x512")) st = SS_FUSED(16, 512, 16);
    else if (shape && !strcmp(shape, "8x512")) st = SS_FUSED(8, 512, 16);
    else if (shape && !strcmp(shape, "6x1024")) st = SS_FUSED(6, 1024, 16);
    else if (shape && !strcmp(shape, "7x1024")) st = SS_FUSED(7, 1024, 16);
    else if (unit == 8) st = SS_FUSED(8, 1024, 8);
    else if (unit == 32) st = SS_FUSED(8, 1024, 32);
    else st = SS_FUSED(8, 1024, 16);
#undef SS_FUSED
    const bool big = shape && strcmp(shape, "8x1024");
    SS_TRY(st);
    if (env_int("SS_RECSYS_VERBOSE", 0)) fprintf(stderr, "ss_recommend_topl: %d clusters (%s, unit %d)\n", ncl, big ? shape : "8x1024", unit);
    int32_t redo = 0;  // sources whose candidate list overflowed take the three-kernel form
    SS_CHECK_CUDA(cudaMemcpyAsync(&redo, fp.redo_cnt, 4, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (redo > 0) SS_TRY(recommend_groups(ctx, Y, YT, L, 0, redo, fp.redo_list, idx_out, val_out, ldv));
    return SS_OK;
}

}  // namespace ss
