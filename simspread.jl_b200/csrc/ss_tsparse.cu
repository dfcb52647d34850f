// First product of the chain when the LABEL matrix is sparse:  T = (Xs' * (Y ./ ks)) ./ kf  with Y a few percent dense.
//
// Reference: W = spread(B) = G ./ k(G) (src/core.jl:365-371) and the first factor of `Aarr * Warr^2` (src/core.jl:413)
// after block reduction: T[f,t] = (sum_s Xs[s,f] * W[s,t]) / kf[f], W[s,t] = Y[s,t] / ks[s].
//
// A drug-target / user-item label matrix is sparse (C4: 5 %, Yamanishi: 1-3 %), so 95 % of the 2 Ns Nf Nt flop of the
// dense DMMA product multiply by zero -- and the DMMA rate of B200 equals its DFMA rate, there is no tensor-core bonus
// to lose.  Here W is compacted by target column (CSC: sources ascending, weight = the IEEE quotient spread() computes)
// and a CTA computes a 128-feature x 128-target tile of T:
//   * slabs of 64 sources x 128 features of Xs' (a transposed copy, features contiguous) are staged in shared memory
//     with cp.async (16-byte chunks, two buffers);
//   * a warp owns 8 targets; per slab it loads a 32-edge window of each target's list (lane l: edge cursor + l), counts
//     the edges that fall into the slab with a ballot and broadcasts them with shuffles; for every edge (s, t) it reads
//     the 128 features of source s with two conflict-free LDS.128 (lane l: features 2l, 2l+1, 64+2l, 65+2l) and issues
//     4 DFMA per lane into accumulators that are indexed statically (the loop over the 8 targets is unrolled);
//   * epilogue: `/ kf[f]` (true division, kf == 0 -> 0), 16-byte stores, optionally replicated into peer-GPU copies of T
//     (the fused all-gather of the sharded chain).
// Every T[f,t] is the sum over its edges in ascending source order, so the result does not depend on tiles, shards or
// the launch geometry.  The kernel is bound by shared-memory bandwidth (8 bytes of Xs' per FMA): ~25 % of the FP64 peak
// on 5 % of the flop.
//
// Declined (the caller falls back to spread + dense GEMM): small products, labels denser than 10 %, non-finite feature
// weights (0 * Inf = NaN in the dense form, skipped here), more than 2^31-1 edges.  SS_T_FORM=dense / sparse forces a form.
#include <stdlib.h>
#include <string.h>

#include "ss_common.cuh"

namespace {

constexpr int TS_FB = 128;      // features per CTA tile
constexpr int TS_TB = 128;      // targets per CTA tile
constexpr int TS_SB = 64;       // sources per shared-memory slab
constexpr int TS_WARPS = 16;
constexpr int TS_TPW = TS_TB / TS_WARPS;  // targets per warp
constexpr int TS_THREADS = TS_WARPS * 32;
constexpr int TS_SLAB_DOUBLES = TS_SB * TS_FB;
constexpr size_t TS_SMEM = size_t(2) * TS_SLAB_DOUBLES * 8;  // 128 KB
constexpr const char* kDefaultForm = "auto";

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- Xs' with a finiteness check (dst: nf x ns column-major, features contiguous) ----------------------------------
__global__ void __launch_bounds__(256)
    transpose_check_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int64_t ldd, int64_t rows,
                           int64_t cols, int32_t* __restrict__ nonfinite) {
    // src: rows x cols (column-major), dst: cols x rows
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = int64_t(blockIdx.x) * 32, c0 = int64_t(blockIdx.y) * 32;
    bool bad = false;
    for (int j = ty; j < 32; j += 8) {
        const int64_t c = c0 + j, r = r0 + tx;
        double v = 0.0;
        if (r < rows && c < cols) {
            v = src[c * lds + r];
            bad |= !isfinite(v);
        }
        tile[j][tx] = v;
    }
    if (bad) *nonfinite = 1;
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int64_t r = r0 + j, c = c0 + tx;
        if (r < rows && c < cols) dst[r * ldd + c] = tile[tx][j];
    }
}

// ---- col_ptr = exclusive scan of the column degrees of Y (single block; nt is at most a few million) --------------
__global__ void __launch_bounds__(1024) ts_scan_kernel(const int32_t* __restrict__ kt, int64_t nt, int32_t* __restrict__ col_ptr,
                                                       long long* __restrict__ total) {
    __shared__ long long part[1024];
    const int t = threadIdx.x;
    const int64_t chunk = (nt + 1023) / 1024;
    const int64_t b = t * chunk, e = min(nt, b + chunk);
    long long s = 0;
    for (int64_t i = b; i < e; ++i) s += kt[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        long long run = 0;
        for (int i = 0; i < 1024; ++i) {
            const long long v = part[i];
            part[i] = run;
            run += v;
        }
        *total = run;
        col_ptr[nt] = int32_t(run > 2147483647ll ? 2147483647ll : run);
    }
    __syncthreads();
    long long run = part[t];
    for (int64_t i = b; i < e; ++i) {
        col_ptr[i] = int32_t(run > 2147483647ll ? 2147483647ll : run);
        run += kt[i];
    }
}

// ---- W by target column: one warp per column of Y, ballot compaction, weight = Y[s,t] / ks[s] as spread() rounds it ----
__global__ void __launch_bounds__(256)
    wcsc_fill_kernel(const double* __restrict__ Y, int64_t ldy, int64_t ns, int64_t nt, const int32_t* __restrict__ ks,
                     const int32_t* __restrict__ col_ptr, int32_t* __restrict__ row_idx, double* __restrict__ val) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (t >= nt) return;
    const double* y = Y + t * ldy;
    int32_t base = col_ptr[t];
    for (int64_t s0 = 0; s0 < ns; s0 += 32) {
        const int64_t s = s0 + lane;
        const double v = s < ns ? __ldg(y + s) : 0.0;
        const bool keep = v != 0.0;  // count(!iszero): NaN is an edge (src/graphs.jl:10), as in the degree kernel
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int32_t pos = base + __popc(ballot & ((1u << lane) - 1u));
            double w = v / double(__ldg(ks + s));  // ks[s] >= 1 here
            if (!isfinite(w)) w = 0.0;              // spread(): Inf -> 0, NaN -> 0 (src/core.jl:368-369)
            row_idx[pos] = int32_t(s);
            val[pos] = w;
        }
        base += __popc(ballot);
    }
}

struct TsParams {
    const double* XsT;  // nf x ns, features contiguous
    int64_t ldf;
    int64_t ns, nf, nt;
    const int32_t* col_ptr;
    const int32_t* row_idx;
    const double* val;
    const int32_t* kf;
    double* T;
    int64_t ldt;
    int nmirror;
    double* mirror[7];
};

__global__ void __launch_bounds__(TS_THREADS, 1) tsp_kernel(const TsParams p) {
    extern __shared__ double ts_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t t0 = int64_t(blockIdx.x) * TS_TB, f0 = int64_t(blockIdx.y) * TS_FB;

    // per target of this warp: cursor into its edge list (warp-uniform)
    int32_t cur[TS_TPW], end[TS_TPW];
    double acc[TS_TPW][4];
    constexpr int32_t kNone = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < TS_TPW; ++j) {
        const int64_t t = t0 + warp * TS_TPW + j;
        cur[j] = t < p.nt ? __ldg(p.col_ptr + t) : 0;
        end[j] = t < p.nt ? __ldg(p.col_ptr + t + 1) : 0;
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0;
    }

    const int64_t nslab = (p.ns + TS_SB - 1) / TS_SB;
    auto load_slab = [&](int64_t slab, int buf) {
        const uint32_t base = smem_addr(ts_smem + size_t(buf) * TS_SLAB_DOUBLES);
#pragma unroll
        for (int i = 0; i < (TS_SB * TS_FB / 2) / TS_THREADS; ++i) {  // 16-byte chunks: 64 per source row
            const int chunk = threadIdx.x + i * TS_THREADS;
            const int sl = chunk >> 6, ch = chunk & 63;
            const int64_t s = slab * TS_SB + sl, f = f0 + 2 * ch;
            const bool ok = s < p.ns && f < p.ldf;  // the padding rows [nf, ldf) are read but never stored
            const double* src = ok ? p.XsT + s * p.ldf + f : p.XsT;
            cp_async16(base + uint32_t(sl * TS_FB + 2 * ch) * 8u, src, ok ? 16 : 0);
        }
        cp_async_commit();
    };

    load_slab(0, 0);
    for (int64_t slab = 0; slab < nslab; ++slab) {
        const int buf = int(slab & 1);
        if (slab + 1 < nslab) {
            load_slab(slab + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const double* A = ts_smem + size_t(buf) * TS_SLAB_DOUBLES;
        const int32_t s_lo = int32_t(slab * TS_SB), s_hi = s_lo + TS_SB;
        // Edge windows: lane l loads edge cur[j] + l of target j (one coalesced load per list, all eight lists in flight
        // together); the edges of this slab are a prefix of the window (sources ascend), counted with a ballot and
        // broadcast one by one with shuffles -- no load sits in the dependent chain of the FMAs.
        int32_t wi[TS_TPW];
        double wv[TS_TPW];
#pragma unroll
        for (int j = 0; j < TS_TPW; ++j) {
            const int32_t c = cur[j] + lane;
            const bool ok = c < end[j];
            wi[j] = ok ? __ldg(p.row_idx + c) : kNone;
            wv[j] = ok ? __ldg(p.val + c) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < TS_TPW; ++j) {
            int n = __popc(__ballot_sync(0xffffffffu, wi[j] < s_hi));
            for (;;) {
                for (int k = 0; k < n; ++k) {
                    const int32_t sj = __shfl_sync(0xffffffffu, wi[j], k);
                    const double w = __shfl_sync(0xffffffffu, wv[j], k);
                    const double2* a = reinterpret_cast<const double2*>(A + size_t(sj - s_lo) * TS_FB);
                    const double2 a0 = a[lane], a1 = a[lane + 32];
                    acc[j][0] = fma(w, a0.x, acc[j][0]);
                    acc[j][1] = fma(w, a0.y, acc[j][1]);
                    acc[j][2] = fma(w, a1.x, acc[j][2]);
                    acc[j][3] = fma(w, a1.y, acc[j][3]);
                }
                cur[j] += n;
                if (n < 32) break;
                // a full window inside one slab (a target with more than 32 of the slab's 64 sources): next window
                const int32_t c = cur[j] + lane;
                const bool ok = c < end[j];
                wi[j] = ok ? __ldg(p.row_idx + c) : kNone;
                wv[j] = ok ? __ldg(p.val + c) : 0.0;
                n = __popc(__ballot_sync(0xffffffffu, wi[j] < s_hi));
            }
        }
        __syncthreads();  // the other buffer is refilled at the top of the next round
    }

    // epilogue: / kf (true division as in W = G ./ k(G); kf == 0 -> 0), T and its peer copies
    const int64_t fa = f0 + 2 * lane, fb = f0 + 64 + 2 * lane;
    double kd_f[4];
    int32_t kd[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t f = (q < 2 ? fa : fb) + (q & 1);
        kd[q] = f < p.nf ? __ldg(p.kf + f) : 0;
        kd_f[q] = double(kd[q]);
    }
#pragma unroll
    for (int j = 0; j < TS_TPW; ++j) {
        const int64_t t = t0 + warp * TS_TPW + j;
        if (t >= p.nt) continue;
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = kd[q] ? acc[j][q] / kd_f[q] : 0.0;
        const int64_t off = t * p.ldt;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t f = h ? fb : fa;
            if (f + 1 < p.nf) {
                const double2 o = make_double2(v[2 * h], v[2 * h + 1]);
                *reinterpret_cast<double2*>(p.T + off + f) = o;
                for (int m = 0; m < p.nmirror; ++m) *reinterpret_cast<double2*>(p.mirror[m] + off + f) = o;
            } else if (f < p.nf) {
                p.T[off + f] = v[2 * h];
                for (int m = 0; m < p.nmirror; ++m) p.mirror[m][off + f] = v[2 * h];
            }
        }
    }
}

}  // namespace

namespace ss {

int32_t t_from_sparse_labels(ss_ctx* ctx, const double* Xs, int64_t ldxs, const double* Y, int64_t ldy, int64_t ns, int64_t nf,
                             int64_t nt, const int32_t* ks, const int32_t* kf, const int32_t* kt, double* T, int64_t ldt,
                             int nmirror, double* const* mirrors, bool* used) {
    *used = false;
    // SS_T_FORM = dense | sparse (forced) | auto (sparse for large products with labels up to 10 % dense)
    const char* form = getenv("SS_T_FORM");
    if (!form) form = kDefaultForm;
    const bool force = !strcmp(form, "sparse");
    if (!force && strcmp(form, "auto")) return SS_OK;
    if (ns <= 0 || nf <= 0 || nt <= 0 || nmirror < 0 || nmirror > 7) return SS_OK;
    if (ns >= (1ll << 31) || nf >= (1ll << 31) || nt >= (1ll << 31) || ceil_div(nf, 32) > 65535) return SS_OK;  // grid.y limits
    // the dense DMMA product of a small problem is a few milliseconds: not worth a CSC build and a host round trip
    if (!force && 2.0 * double(ns) * double(nf) * double(nt) < 2e11) return SS_OK;
    if ((ldt & 1) || (reinterpret_cast<uintptr_t>(T) & 15)) return SS_OK;
    for (int i = 0; i < nmirror; ++i)
        if (reinterpret_cast<uintptr_t>(mirrors[i]) & 15) return SS_OK;

    void* p;
    SS_TRY(scratch_get(ctx, 23, size_t(nt + 1) * 4 + 64, &p));
    int32_t* col_ptr = static_cast<int32_t*>(p);
    long long* total = reinterpret_cast<long long*>(reinterpret_cast<char*>(p) + ((size_t(nt + 1) * 4 + 15) & ~size_t(15)));
    int32_t* nonfinite = reinterpret_cast<int32_t*>(total + 1);
    const int64_t ldf = round_up(nf, 16);
    SS_TRY(scratch_get(ctx, 25, size_t(ldf) * size_t(ns) * 8, &p));
    double* XsT = static_cast<double*>(p);
    SS_CHECK_CUDA(cudaMemsetAsync(nonfinite, 0, 4, ctx->stream));
    ts_scan_kernel<<<1, 1024, 0, ctx->stream>>>(kt, nt, col_ptr, total);
    {
        const dim3 grid(unsigned(ceil_div(ns, 32)), unsigned(ceil_div(nf, 32)));
        transpose_check_kernel<<<grid, 256, 0, ctx->stream>>>(Xs, ldxs, XsT, ldf, ns, nf, nonfinite);
    }
    ctx->launches += 2;
    long long h_total = 0;
    int32_t h_bad = 0;
    SS_CHECK_CUDA(cudaMemcpyAsync(&h_total, total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaMemcpyAsync(&h_bad, nonfinite, 4, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h_bad || h_total > 2147483647ll) return SS_OK;
    if (!force && double(h_total) > 0.10 * double(ns) * double(nt)) return SS_OK;
    const size_t nnz = size_t(h_total > 0 ? h_total : 1);
    SS_TRY(scratch_get(ctx, 24, nnz * 12 + 64, &p));
    double* val = static_cast<double*>(p);
    int32_t* row_idx = reinterpret_cast<int32_t*>(val + nnz);
    wcsc_fill_kernel<<<unsigned(ceil_div(nt * 32, 256)), 256, 0, ctx->stream>>>(Y, ldy, ns, nt, ks, col_ptr, row_idx, val);
    ctx->launches++;

    TsParams q;
    q.XsT = XsT;
    q.ldf = ldf;
    q.ns = ns;
    q.nf = nf;
    q.nt = nt;
    q.col_ptr = col_ptr;
    q.row_idx = row_idx;
    q.val = val;
    q.kf = kf;
    q.T = T;
    q.ldt = ldt;
    q.nmirror = nmirror;
    for (int i = 0; i < 7; ++i) q.mirror[i] = i < nmirror ? mirrors[i] : nullptr;
    if (!ctx->tsp_attr_set) {
        SS_CHECK_CUDA(cudaFuncSetAttribute(tsp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TS_SMEM)));
        ctx->tsp_attr_set = true;
    }
    ss_ctx::ProfRec rec{nullptr, nullptr, 2.0 * double(h_total) * double(nf)};
    if (ctx->profile) {
        SS_CHECK_CUDA(cudaEventCreate(&rec.start));
        SS_CHECK_CUDA(cudaEventCreate(&rec.stop));
        SS_CHECK_CUDA(cudaEventRecord(rec.start, ctx->stream));
    }
    // target tiles vary fastest: the CTAs in flight share one 128-feature panel of Xs' through L2
    const dim3 grid(unsigned(ceil_div(nt, TS_TB)), unsigned(ceil_div(nf, TS_FB)));
    tsp_kernel<<<grid, TS_THREADS, TS_SMEM, ctx->stream>>>(q);
    SS_CHECK_CUDA(cudaGetLastError());
    if (ctx->profile) {
        SS_CHECK_CUDA(cudaEventRecord(rec.stop, ctx->stream));
        ctx->prof.push_back(rec);
    }
    ctx->launches++;
    *used = true;
    return SS_OK;
}

}  // namespace ss
