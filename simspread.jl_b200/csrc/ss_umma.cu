// tcgen05 (5th-generation tensor core) form of the chain-product GEMM: the opt-in reduced-precision
// modes of predict.  The reference's only GPU mode (`GPU=true`, src/core.jl:404,411-413) silently
// computes `A * W^2` in Float32 through cuBLAS SGEMM; here the same products run on tcgen05.mma
// kind::tf32 with accumulators in TMEM:
//
//   SS_PRECISION_TF32 : one pass, operands rounded to TF32 (10-bit mantissa), FP32 accumulation.
//   (A 3xTF32 split -- hi*hi + hi*lo + lo*hi, the `split` argument below -- was measured too: the
//   in-TMEM FP32 accumulation truncates, so at K = 20000 the split result (3.3e-4) is no better than
//   the single pass (1.9e-4) at 3x the cost; it is therefore not exposed through the ABI.)
//
// Structure (one CTA per SM, persistent over 128x256 C tiles):
//   warp 0  TMA producer : cp.async.bulk.tensor of K-major, SWIZZLE_128B operand slabs (A 128 rows,
//                          B 256 rows, 128 bytes of K per slab) into a 4-stage smem ring
//   warp 1  MMA issuer   : one lane issues tcgen05.mma (M=128, N=256, 32 bytes of K per instruction)
//                          from smem descriptors; tcgen05.commit releases the smem stage / publishes
//                          the accumulator
//   warp 2  TMEM owner   : tcgen05.alloc of all 512 columns = two 128x256 FP32 accumulators, so the
//                          epilogue of tile i overlaps the MMAs of tile i+1
//   warps 4-7 epilogue   : tcgen05.ld (32 lanes x 32 columns per instruction) -> FP64 -> fused
//                          normalisation (/kf, clean! flag) -> column-major C
// Operand pairs (A_p, B_p) are concatenated along K inside one accumulator ("group"); a tile may run
// several groups, each with its own TMEM accumulator pass and epilogue scale.
//
//   SS_PRECISION_F64_INT8 : FP64-grade products from INTEGER tensor-core math (Ozaki-style slicing).
//   Each non-negative FP64 operand row is scaled by a power of two and cut into S unsigned 8-bit
//   slices, x = 2^e * sum_i q_i 2^(-8i); slice products q_i^A * q_j^B are accumulated EXACTLY in
//   INT32 by tcgen05.mma kind::i8 (groups are sized so that the unsigned 32-bit sum cannot wrap), and
//   the epilogue adds 2^(ea+eb-8(i+j)) * G_ij into the FP64 result (every term exact, ~S roundings in
//   the final sum).  Pairs with i+j > S+1 are dropped: the error is bounded by
//   ~(S+1)*K*2^(-8S) * rowmax(A)*colmax(B) -- a NORMWISE bound (1e-13 of a typical entry for S = 6,
//   K = 20000), not the element-wise 1e-12 of the default DMMA path, which is why this mode is opt-in.
#include <cuda.h>
#include <stdlib.h>

#include "ss_common.cuh"

namespace {

constexpr int UM = 128, UN = 256;
constexpr int U_STAGES = 4;
constexpr int A_ST_BYTES = UM * 128;  // 16 KB
constexpr int B_ST_BYTES = UN * 128;  // 32 KB
constexpr int ST_BYTES = A_ST_BYTES + B_ST_BYTES;
constexpr int U_THREADS = 256;
constexpr int MAXS = 8;    // operand slices (tensor maps) per side
constexpr int MAXPAIR = 36; // slice pairs per tile
constexpr int MAXGRP = 36;  // accumulator groups per tile
constexpr int U_GROUP_M = 8;
constexpr int U_SYNC_CHUNK = 32;  // 128-byte slabs between lockstep checkpoints
constexpr size_t U_SMEM = size_t(U_STAGES) * ST_BYTES + 1024 + 256;

struct UmmaMaps {
    CUtensorMap a[MAXS];
    CUtensorMap b[MAXS];
};

struct UmmaParams {
    int M, N;
    int kslabs;  // 128-byte K slabs per operand pair
    int bke;     // elements per slab
    int tiles_m, tiles_n;
    uint32_t idesc;
    int ngroups;                 // accumulator passes per tile
    int gstart[MAXGRP + 1];      // pairs [gstart[g], gstart[g+1]) share accumulator g
    double gscale[MAXGRP];       // epilogue scale of group g (2^(-8(i+j)) for integer slices)
    signed char pa[MAXPAIR], pb[MAXPAIR];  // slice index of A / B for every pair
    double* C;
    int64_t ldc;
    const int32_t* row_div;
    const int32_t* col_flag;
    const double* rscale;  // integer mode: 2^ea[m]
    const double* cscale;  // integer mode: 2^eb[n]
    int* sync_prog;        // loose lockstep of the producers (see ss_gemm.cu), optional
    // integer mode, a-posteriori certificate: an entry passes if acc >= bound(m, n) / tol (see launch_gemm_i8) or if
    // it is exactly 0 and no operand entry was truncated to 0
    double certA, certB, certD;     // (truncation of A, truncation of B, dropped slice pairs * K) / tol
    const double* rsum;             // sum_k |A[m,k]|
    const double* csum;             // sum_k |B[k,n]|
    int zero_ok;
    unsigned long long* cert_fail;  // number of entries that did not pass (null: no check)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "UWAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni UWAIT_DONE;\n"
        "bra.uni UWAIT_LOOP;\n"
        "UWAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm100):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major) | SBO>>4 [32,46) = 8 rows * 128 B |
// version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
    return uint64_t((smem_addr >> 4) & 0x3FFF) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
           (uint64_t(2) << 61);
}

template <int KIND>  // 0: kind::tf32 (FP32 accumulate), 1: kind::i8 (INT32 accumulate)
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (KIND == 0) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct UTile {
    int tm, tn;
};
__device__ __forceinline__ UTile utile(int tile, int tiles_m, int tiles_n) {
    const int group_size = U_GROUP_M * tiles_n;
    const int gid = tile / group_size;
    const int first_m = gid * U_GROUP_M;
    const int gm = min(tiles_m - first_m, U_GROUP_M);
    const int r = tile - gid * group_size;
    return {first_m + r % gm, r / gm};
}

template <int KIND>
__global__ void __launch_bounds__(U_THREADS, 1)
    ss_umma_kernel(const __grid_constant__ UmmaMaps maps, const __grid_constant__ UmmaParams p) {
    extern __shared__ uint8_t usmem_raw[];
    __shared__ uint32_t tmem_base_slot;
    const uint32_t smem_base = (smem_u32(usmem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + U_STAGES * ST_BYTES;
    // barriers: full[U_STAGES], empty[U_STAGES], tmem_full[2], tmem_empty[2]
    const uint32_t bar_full = bar_base, bar_empty = bar_base + 8 * U_STAGES;
    const uint32_t bar_tfull = bar_base + 16 * U_STAGES, bar_tempty = bar_tfull + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.tiles_m * p.tiles_n;

    if (threadIdx.x == 0) {
        for (int s = 0; s < U_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, 4);  // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int checkpoint = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const UTile tc = utile(tile, p.tiles_m, p.tiles_n);
                const int m0 = tc.tm * UM, n0 = tc.tn * UN;
                const int npairs = p.gstart[p.ngroups];
                for (int pr = 0; pr < npairs; ++pr)
                    for (int kb = 0; kb < p.kslabs; ++kb) {
                        if (p.sync_prog && (kb % U_SYNC_CHUNK) == 0) {
                            // loose lockstep (same scheme and rationale as ss_gemm.cu): concurrent tiles
                            // share operand panels through L2 only while they stream K at nearby positions
                            ++checkpoint;
                            *reinterpret_cast<volatile int*>(p.sync_prog + blockIdx.x) = checkpoint;
                            const long long t0 = clock64();
                            for (;;) {
                                int mn = 0x7fffffff;
                                for (int i = 0; i < int(gridDim.x); ++i)
                                    mn = min(mn, *reinterpret_cast<volatile int*>(p.sync_prog + i));
                                if (mn >= checkpoint - 1 || clock64() - t0 > 400000ll) break;
                                __nanosleep(256);
                            }
                        }
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        mbar_expect_tx(bar_full + 8 * stage, ST_BYTES);
                        const uint32_t sA = smem_base + stage * ST_BYTES;
                        tma_load_2d(sA, &maps.a[p.pa[pr]], kb * p.bke, m0, bar_full + 8 * stage);
                        tma_load_2d(sA + A_ST_BYTES, &maps.b[p.pb[pr]], kb * p.bke, n0, bar_full + 8 * stage);
                        if (++stage == U_STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
            }
            if (p.sync_prog) *reinterpret_cast<volatile int*>(p.sync_prog + blockIdx.x) = 0x7fffffff;
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x)
            for (int g = 0; g < p.ngroups; ++g, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(bar_tempty + 8 * as, aphase ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * UN;
                uint32_t accumulate = 0;
                const int nslab = (p.gstart[g + 1] - p.gstart[g]) * p.kslabs;
                for (int s = 0; s < nslab; ++s) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sA = smem_base + stage * ST_BYTES;
                    const uint64_t adesc = make_kmajor_desc(sA);
                    const uint64_t bdesc = make_kmajor_desc(sA + A_ST_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // 4 x 32 bytes of K per 128-byte slab
                        umma<KIND>(tmem_d, adesc + 2 * k, bdesc + 2 * k, p.idesc, accumulate);
                        accumulate = 1;
                    }
                    umma_commit(bar_empty + 8 * stage);  // smem stage reusable once these MMAs retire
                    if (++stage == U_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(bar_tfull + 8 * as);  // accumulator complete
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int ew = warp - 4;  // TMEM lanes [32*ew, 32*ew+32)
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const UTile tc = utile(tile, p.tiles_m, p.tiles_n);
            const int m0 = tc.tm * UM, n0 = tc.tn * UN;
            const int row = m0 + 32 * ew + lane;
            double inv = 1.0, rs = 1.0, rsm = 0.0;
            bool zero_row = false;
            if (KIND == 1 && p.cert_fail && row < p.M) rsm = __ldg(p.rsum + row);
            if (row < p.M) {
                if (p.row_div) {
                    const int d = __ldg(p.row_div + row);
                    zero_row = (d == 0);
                    inv = double(d);
                }
                if (KIND == 1) rs = __ldg(p.rscale + row);
            }
            unsigned nfail = 0;
            for (int g = 0; g < p.ngroups; ++g, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(bar_tfull + 8 * as, aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (uint32_t(32 * ew) << 16) + as * UN;
                const bool first = (g == 0), last = (g == p.ngroups - 1);
                const double gs = (KIND == 1) ? rs * p.gscale[g] : 1.0;
#pragma unroll 1
                for (int c = 0; c < UN / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c * 32, r);
                    if (row < p.M) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = n0 + c * 32 + j;
                            if (col < p.N) {
                                double* dst = p.C + int64_t(col) * p.ldc + row;
                                double v;
                                if (KIND == 0) {
                                    v = double(__uint_as_float(r[j]));
                                } else {
                                    // exact: (uint32 < 2^32) * 2^k; the running FP64 sum lives in C (L2-hot)
                                    v = double(r[j]) * (gs * __ldg(p.cscale + col));
                                    if (!first) v += *dst;
                                }
                                if (last) {
                                    if (KIND == 1 && p.cert_fail) {
                                        const double cs_ = __ldg(p.cscale + col);
                                        const double need = p.certA * rs * __ldg(p.csum + col) + p.certB * cs_ * rsm + p.certD * rs * cs_;
                                        // 1 % slack covers the FP64 rounding of the recombination (<= ngroups * 2^-53 relative)
                                        const bool pass = (v * 0.99 >= need) || (v == 0.0 && p.zero_ok);
                                        nfail += pass ? 0u : 1u;
                                    }
                                    if (p.row_div) v = zero_row ? 0.0 : v / inv;
                                    if (p.col_flag && __ldg(p.col_flag + col) == 0) v = -99.0;
                                }
                                *dst = v;
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
            }
            if (KIND == 1 && p.cert_fail) {
                nfail = __reduce_add_sync(0xffffffffu, nfail);
                if (lane == 0 && nfail) atomicAdd(p.cert_fail, (unsigned long long)nfail);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- FP64 -> TF32 operand preparation ---------------------------------------------------------
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t o;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o) : "f"(x));
    return __uint_as_float(o);
}

// src: column-major FP64 with K contiguous (element (k, j) at src[j*ld + k]) -> dst[j][k] float,
// row pitch kp.  hi = tf32(x); lo (optional) = tf32(float(x) - hi).
__global__ void __launch_bounds__(256)
    cvt_kmajor_kernel(const double* __restrict__ src, int64_t K, int64_t J, int64_t ld, float* __restrict__ hi,
                      float* __restrict__ lo, int64_t kp) {
    const int64_t k = int64_t(blockIdx.x) * 256 + threadIdx.x;
    if (k >= kp) return;
    for (int64_t j = blockIdx.y; j < J; j += gridDim.y) {
        const float x = (k < K) ? float(src[j * ld + k]) : 0.f;
        const float h = to_tf32(x);
        hi[j * kp + k] = h;
        if (lo) lo[j * kp + k] = to_tf32(x - h);
    }
}

// src: column-major FP64 with M contiguous (element (m, k) at src[k*ld + m]) -> dst[m][k] float
__global__ void __launch_bounds__(256)
    cvt_transpose_kernel(const double* __restrict__ src, int64_t M, int64_t K, int64_t ld, float* __restrict__ hi,
                         float* __restrict__ lo, int64_t kp) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t m0 = int64_t(blockIdx.x) * 32, k0 = int64_t(blockIdx.y) * 32;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int64_t m = m0 + tx, k = k0 + j;
        tile[j][tx] = (m < M && k < K) ? float(src[k * ld + m]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int64_t m = m0 + j, k = k0 + tx;
        if (m < M && k < kp) {
            const float x = tile[tx][j];
            const float h = to_tf32(x);
            hi[m * kp + k] = h;
            if (lo) lo[m * kp + k] = to_tf32(x - h);
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// K-major operand: `rows` x K elements of `esize` bytes, row pitch `pitch_bytes`; box = 128 B of K x box_rows
int32_t make_map_kmajor(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int esize, int64_t K, int64_t rows,
                        int64_t pitch_bytes, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        ss::set_error("cuTensorMapEncodeTiled is not available from the driver");
        return SS_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {cuuint64_t(K), cuuint64_t(rows)};
    cuuint64_t gstride[1] = {cuuint64_t(pitch_bytes)};
    cuuint32_t box[2] = {cuuint32_t(128 / esize), cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ss::set_error("cuTensorMapEncodeTiled (k-major, %d-byte elements) failed with CUresult %d", esize, int(r));
        return SS_ERR_CUDA;
    }
    return SS_OK;
}

// ---- FP64 -> unsigned 8-bit slices (integer mode) -----------------------------------------------
// per-row scale 2^e with max|x| < 2^e, and flags: bit0 = negative entry, bit1 = NaN / Inf entry
__device__ __forceinline__ double pow2_above(double mx) {
    if (mx == 0.0) return 1.0;
    int e;
    frexp(mx, &e);  // mx = f * 2^e, f in [0.5, 1)
    return ldexp(1.0, e);
}

// k-contiguous source: element (k, j) at src[j*ld + k]; one block per row j
__global__ void __launch_bounds__(256)
    rowscale_kmajor_kernel(const double* __restrict__ src, int64_t K, int64_t J, int64_t ld, double* __restrict__ scale,
                           double* __restrict__ sums, int* __restrict__ flags) {
    __shared__ double smx[256], ssm[256];
    for (int64_t j = blockIdx.x; j < J; j += gridDim.x) {
        double mx = 0.0, sm = 0.0;
        int bad = 0;
        for (int64_t k = threadIdx.x; k < K; k += 256) {
            const double x = src[j * ld + k];
            if (x < 0.0) bad |= 1;
            if (!(fabs(x) <= 1.7976931348623157e308)) bad |= 2;
            mx = fmax(mx, fabs(x));
            sm += fabs(x);
        }
        if (bad) atomicOr(flags, bad);
        smx[threadIdx.x] = mx;
        ssm[threadIdx.x] = sm;
        __syncthreads();
        for (int st = 128; st > 0; st >>= 1) {
            if (threadIdx.x < st) {
                smx[threadIdx.x] = fmax(smx[threadIdx.x], smx[threadIdx.x + st]);
                ssm[threadIdx.x] += ssm[threadIdx.x + st];
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            scale[j] = pow2_above(smx[0]);
            sums[j] = ssm[0] * (1.0 + 1e-12);  // an upper bound of the exact sum (rounding of the reduction)
        }
        __syncthreads();
    }
}

// m-contiguous source: element (m, k) at src[k*ld + m]; one thread per row m (coalesced over m)
__global__ void __launch_bounds__(256)
    rowscale_mmajor_kernel(const double* __restrict__ src, int64_t M, int64_t K, int64_t ld, double* __restrict__ scale,
                           double* __restrict__ sums, int* __restrict__ flags) {
    const int64_t m = int64_t(blockIdx.x) * 256 + threadIdx.x;
    if (m >= M) return;
    double mx = 0.0, sm = 0.0;
    int bad = 0;
    for (int64_t k = 0; k < K; ++k) {
        const double x = __ldg(src + k * ld + m);
        if (x < 0.0) bad |= 1;
        if (!(fabs(x) <= 1.7976931348623157e308)) bad |= 2;
        mx = fmax(mx, fabs(x));
        sm += fabs(x);
    }
    if (bad) atomicOr(flags, bad);
    scale[m] = pow2_above(mx);
    sums[m] = sm * (1.0 + 1e-12);
}

// r in [0,1) -> S unsigned 8-bit digits, most significant first (all operations exact)
// returns the mask of planes that received a non-zero digit
template <int DUMMY = 0>
// bit 31 of the result: the entry is not exactly representable in S digits (it was truncated)
__device__ __forceinline__ unsigned slice_digits(double r, int S, uint8_t* out, int64_t plane, int* flags) {
    if (r != 0.0 && r < ldexp(1.0, -8 * S)) atomicOr(flags, 4);  // a non-zero entry whose digits are all 0
    unsigned used = 0;
    for (int i = 0; i < S; ++i) {
        r *= 256.0;
        const double q = floor(r);
        r -= q;
        out[int64_t(i) * plane] = uint8_t(int(q));
        used |= (q != 0.0 ? 1u : 0u) << i;
    }
    if (r != 0.0) used |= 0x80000000u;
    return used;
}

// planes that hold only zeros need no tensor-core pass (a 0/1 label matrix is ONE plane): OR the per-thread masks
__device__ __forceinline__ void publish_planes(unsigned used, int* plane_mask) {
    used = __reduce_or_sync(0xffffffffu, used);
    if ((threadIdx.x & 31) == 0 && used) atomicOr(plane_mask, int(used));
}

// k-contiguous source -> planes[i][j*kp + k]
__global__ void __launch_bounds__(256)
    slice_kmajor_kernel(const double* __restrict__ src, int64_t K, int64_t J, int64_t ld, const double* __restrict__ scale,
                        int S, uint8_t* __restrict__ planes, int64_t kp, int* __restrict__ flags,
                        int* __restrict__ plane_mask) {
    const int64_t k = int64_t(blockIdx.x) * 256 + threadIdx.x;
    const int64_t plane = J * kp;
    unsigned used = 0;
    if (k < kp)
        for (int64_t j = blockIdx.y; j < J; j += gridDim.y) {
            const double x = (k < K) ? src[j * ld + k] : 0.0;
            used |= slice_digits(fabs(x) / scale[j], S, planes + j * kp + k, plane, flags);  // scale is a power of two: exact
        }
    publish_planes(used, plane_mask);
}

// m-contiguous source -> planes[i][m*kp + k] (32x32 transpose through shared memory)
__global__ void __launch_bounds__(256)
    slice_mmajor_kernel(const double* __restrict__ src, int64_t M, int64_t K, int64_t ld, const double* __restrict__ scale,
                        int S, uint8_t* __restrict__ planes, int64_t kp, int* __restrict__ flags,
                        int* __restrict__ plane_mask) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t m0 = int64_t(blockIdx.x) * 32, k0 = int64_t(blockIdx.y) * 32;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int64_t m = m0 + tx, k = k0 + j;
        tile[j][tx] = (m < M && k < K) ? src[k * ld + m] : 0.0;
    }
    __syncthreads();
    const int64_t plane = M * kp;
    unsigned used = 0;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int64_t m = m0 + j, k = k0 + tx;
        if (m < M && k < kp) used |= slice_digits(fabs(tile[tx][j]) / scale[m], S, planes + m * kp + k, plane, flags);
    }
    publish_planes(used, plane_mask);
}

inline unsigned grid_y_for(const ss_ctx* ctx, int64_t gx, int64_t n) {
    int64_t want = ss::ceil_div(int64_t(ctx->sm_count) * 16, gx);
    if (want < 1) want = 1;
    if (want > n) want = n;
    if (want > 65535) want = 65535;
    return unsigned(want);
}

template <int KIND>
int32_t launch_umma(ss_ctx* ctx, const UmmaMaps& maps, UmmaParams q, double flops) {
    const int64_t total = int64_t(q.tiles_m) * q.tiles_n;
    const int grid = int(total < ctx->sm_count ? total : ctx->sm_count);
    q.sync_prog = nullptr;
    // measured at C4: the lockstep turns the 21-pair integer kernel from HBM/power-bound (0.83 GHz,
    // 3.55 s) into 1.70 GHz / 2.16 s; the short single-pass TF32 tiles lose 15 % to it, so it stays off
    if (KIND == 1 && total > grid && !getenv("SS_NO_LOCKSTEP")) {
        if (!ctx->tile_counter) SS_CHECK_CUDA(cudaMalloc(&ctx->tile_counter, 4096));
        SS_CHECK_CUDA(cudaMemsetAsync(ctx->tile_counter, 0, size_t(grid) * 4, ctx->stream));
        q.sync_prog = ctx->tile_counter;
    }
    SS_CHECK_CUDA(cudaFuncSetAttribute(ss_umma_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(U_SMEM)));
    ss_ctx::ProfRec rec{nullptr, nullptr, flops};
    if (ctx->profile) {
        SS_CHECK_CUDA(cudaEventCreate(&rec.start));
        SS_CHECK_CUDA(cudaEventCreate(&rec.stop));
        SS_CHECK_CUDA(cudaEventRecord(rec.start, ctx->stream));
    }
    ss_umma_kernel<KIND><<<grid, U_THREADS, U_SMEM, ctx->stream>>>(maps, q);
    SS_CHECK_CUDA(cudaGetLastError());
    if (ctx->profile) {
        SS_CHECK_CUDA(cudaEventRecord(rec.stop, ctx->stream));
        ctx->prof.push_back(rec);
    }
    ctx->launches++;
    return SS_OK;
}

}  // namespace

namespace ss {

// C (FP64, column-major, M x N) = op(A) * B with TF32 operands / FP32 accumulation on tcgen05.
// Scratch slots 14 (A) and 15 (B) hold the converted operands.
int32_t launch_gemm_tf32(ss_ctx* ctx, int opA, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                         int64_t ldc, int64_t M, int64_t N, int64_t K, const int32_t* row_div, const int32_t* col_flag,
                         bool /*split*/) {
    SS_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_tf32: empty problem");
    SS_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm_tf32: dimension too large");
    const int64_t kp = round_up(K, 32);
    void* p;
    SS_TRY(scratch_get(ctx, 14, size_t(M) * kp * 4, &p));
    float* Ahi = static_cast<float*>(p);
    SS_TRY(scratch_get(ctx, 15, size_t(N) * kp * 4, &p));
    float* Bhi = static_cast<float*>(p);
    if (opA == SS_OP_N) {
        dim3 g{unsigned(ceil_div(M, 32)), unsigned(ceil_div(kp, 32))};
        SS_REQUIRE(g.y <= 65535, "gemm_tf32: K too large for the transpose grid");
        cvt_transpose_kernel<<<g, 256, 0, ctx->stream>>>(A, M, K, lda, Ahi, nullptr, kp);
    } else {
        const int64_t gx = ceil_div(kp, 256);
        dim3 g{unsigned(gx), grid_y_for(ctx, gx, M)};
        cvt_kmajor_kernel<<<g, 256, 0, ctx->stream>>>(A, K, M, lda, Ahi, nullptr, kp);
    }
    {
        const int64_t gx = ceil_div(kp, 256);
        dim3 g{unsigned(gx), grid_y_for(ctx, gx, N)};
        cvt_kmajor_kernel<<<g, 256, 0, ctx->stream>>>(B, K, N, ldb, Bhi, nullptr, kp);
    }
    ctx->launches += 2;
    SS_CHECK_CUDA(cudaGetLastError());
    UmmaMaps maps;
    UmmaParams q{};
    SS_TRY(make_map_kmajor(&maps.a[0], Ahi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, K, M, kp * 4, UM));
    SS_TRY(make_map_kmajor(&maps.b[0], Bhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, K, N, kp * 4, UN));
    for (int i = 1; i < MAXS; ++i) {
        maps.a[i] = maps.a[0];
        maps.b[i] = maps.b[0];
    }
    q.M = int(M);
    q.N = int(N);
    q.kslabs = int(kp / 32);
    q.bke = 32;
    q.tiles_m = int(ceil_div(M, UM));
    q.tiles_n = int(ceil_div(N, UN));
    // cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6) | a_format TF32 (2) [7,10) | b_format TF32 (2) [10,13) |
    // a_major K (0) [15] | b_major K (0) [16] | N>>3 [17,23) | M>>4 [24,29)
    q.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(UN >> 3) << 17) | (uint32_t(UM >> 4) << 24);
    q.ngroups = 1;
    q.gstart[0] = 0;
    q.gstart[1] = 1;
    q.gscale[0] = 1.0;
    q.pa[0] = q.pb[0] = 0;
    q.C = C;
    q.ldc = ldc;
    q.row_div = row_div;
    q.col_flag = col_flag;
    return launch_umma<0>(ctx, maps, q, 2.0 * double(M) * double(N) * double(K));
}

// C = op(A) * B for NON-NEGATIVE FP64 operands from exact integer slice products (see file header).
// S = number of 8-bit slices per operand (4..8).
int32_t launch_gemm_i8(ss_ctx* ctx, int opA, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                       int64_t ldc, int64_t M, int64_t N, int64_t K, const int32_t* row_div, const int32_t* col_flag,
                       int S, double cert_tol, int64_t* uncertified) {
    SS_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_i8: empty problem");
    SS_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm_i8: dimension too large");
    SS_REQUIRE(S >= 2 && S <= MAXS, "gemm_i8: 2..%d slices", MAXS);
    const int64_t kp = round_up(K, 128);
    // unsigned 32-bit accumulation of `g` pairs stays exact while g * K * 255^2 < 2^32
    const int64_t per_group = int64_t(4294967295.0 / (double(kp) * 65025.0));
    SS_REQUIRE(per_group >= 1, "gemm_i8: K = %lld is too long for exact INT32 accumulation (max 66048)", (long long)K);
    void* p;
    SS_TRY(scratch_get(ctx, 14, size_t(S) * M * kp + size_t(M) * 16 + 64, &p));
    uint8_t* Ap = static_cast<uint8_t*>(p);
    double* rs = reinterpret_cast<double*>(Ap + size_t(S) * M * kp);
    double* rsum = rs + M;
    SS_TRY(scratch_get(ctx, 15, size_t(S) * N * kp + size_t(N) * 16 + 64, &p));
    uint8_t* Bp = static_cast<uint8_t*>(p);
    double* cs = reinterpret_cast<double*>(Bp + size_t(S) * N * kp);
    double* csum = cs + N;
    int* flags = reinterpret_cast<int*>(csum + N);
    // flags[0]: bad-entry bits, flags[1] / flags[2]: planes of A / B that hold a non-zero digit
    unsigned long long* cert_fail = reinterpret_cast<unsigned long long*>(flags + 4);
    SS_CHECK_CUDA(cudaMemsetAsync(flags, 0, 24, ctx->stream));
    const int rgrid = ctx->sm_count * 8;
    if (opA == SS_OP_N) {
        rowscale_mmajor_kernel<<<unsigned(ceil_div(M, 256)), 256, 0, ctx->stream>>>(A, M, K, lda, rs, rsum, flags);
        dim3 g{unsigned(ceil_div(M, 32)), unsigned(ceil_div(kp, 32))};
        SS_REQUIRE(g.y <= 65535, "gemm_i8: K too large for the transpose grid");
        slice_mmajor_kernel<<<g, 256, 0, ctx->stream>>>(A, M, K, lda, rs, S, Ap, kp, flags, flags + 1);
    } else {
        rowscale_kmajor_kernel<<<unsigned(M < rgrid ? M : rgrid), 256, 0, ctx->stream>>>(A, K, M, lda, rs, rsum, flags);
        const int64_t gx = ceil_div(kp, 256);
        dim3 g{unsigned(gx), grid_y_for(ctx, gx, M)};
        slice_kmajor_kernel<<<g, 256, 0, ctx->stream>>>(A, K, M, lda, rs, S, Ap, kp, flags, flags + 1);
    }
    {
        rowscale_kmajor_kernel<<<unsigned(N < rgrid ? N : rgrid), 256, 0, ctx->stream>>>(B, K, N, ldb, cs, csum, flags);
        const int64_t gx = ceil_div(kp, 256);
        dim3 g{unsigned(gx), grid_y_for(ctx, gx, N)};
        slice_kmajor_kernel<<<g, 256, 0, ctx->stream>>>(B, K, N, ldb, cs, S, Bp, kp, flags, flags + 2);
    }
    ctx->launches += 4;
    SS_CHECK_CUDA(cudaGetLastError());
    int hf[3] = {0, 0, 0};
    SS_CHECK_CUDA(cudaMemcpyAsync(hf, flags, 12, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    const int hflags = hf[0];
    const unsigned planesA = unsigned(hf[1]), planesB = unsigned(hf[2]);
    if (hflags & 3) {
        set_error("the int8-sliced precision mode needs finite, non-negative operands (%s entry found); "
                  "use the default FP64 mode", (hflags & 2) ? "NaN/Inf" : "negative");
        return SS_ERR_UNSUPPORTED;
    }
    UmmaMaps maps;
    UmmaParams q{};
    for (int i = 0; i < MAXS; ++i) {
        const int s = i < S ? i : 0;
        SS_TRY(make_map_kmajor(&maps.a[i], Ap + size_t(s) * M * kp, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, kp, M, kp, UM));
        SS_TRY(make_map_kmajor(&maps.b[i], Bp + size_t(s) * N * kp, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, kp, N, kp, UN));
    }
    q.M = int(M);
    q.N = int(N);
    q.kslabs = int(kp / 128);
    q.bke = 128;
    q.tiles_m = int(ceil_div(M, UM));
    q.tiles_n = int(ceil_div(N, UN));
    // c_format S32 (2) | a_format / b_format unsigned 8-bit (0) | K-major | N>>3 | M>>4
    q.idesc = (2u << 4) | (uint32_t(UN >> 3) << 17) | (uint32_t(UM >> 4) << 24);
    // slice i (1-based weight 2^(-8i)) x slice j: keep i + j <= S + 1; smallest terms first
    double dropped = 0.0;  // sum over the dropped pairs of non-zero planes of 2^(16 - 8(i+j)) >= q_i q_j 2^(-8(i+j))
    for (int i = 1; i <= S; ++i)
        for (int j = 1; j <= S; ++j)
            if (i + j > S + 1 && ((planesA >> (i - 1)) & 1) && ((planesB >> (j - 1)) & 1)) dropped += ldexp(1.0, 16 - 8 * (i + j));
    int np = 0, ng = 0;
    q.gstart[0] = 0;
    for (int d = S + 1; d >= 2; --d) {
        int in_group = 0;
        for (int i = 1; i <= S; ++i) {
            const int j = d - i;
            if (j < 1 || j > S) continue;
            // a plane of zeros contributes nothing (keep the leading pair so that C is always written)
            if ((!((planesA >> (i - 1)) & 1) || !((planesB >> (j - 1)) & 1)) && !(i == 1 && j == 1)) continue;
            if (in_group == per_group) {
                q.gscale[ng] = ldexp(1.0, -8 * d);
                q.gstart[++ng] = np;
                in_group = 0;
            }
            q.pa[np] = (signed char)(i - 1);
            q.pb[np] = (signed char)(j - 1);
            ++np;
            ++in_group;
        }
        if (in_group) {
            q.gscale[ng] = ldexp(1.0, -8 * d);
            q.gstart[++ng] = np;
        }
    }
    SS_REQUIRE(np >= 1 && np <= MAXPAIR && ng <= MAXGRP, "gemm_i8: bad number of slice pairs");
    ctx->int8_last_pairs = np;
    q.ngroups = ng;
    q.C = C;
    q.ldc = ldc;
    q.row_div = row_div;
    q.col_flag = col_flag;
    q.rscale = rs;
    q.cscale = cs;
    // A-posteriori certificate.  With x = 2^e (sum_i q_i 2^(-8i) + d), 0 <= d < 2^(-8S) only if x is not exactly
    // representable (flag from the slicing), the absolute error of entry (m, n) is below
    //   tA * 2^ea[m] * sum_k |B[k,n]|  +  tB * 2^eb[n] * sum_k |A[m,k]|  +  dropped * K * 2^ea[m] * 2^eb[n]
    // (tA = 2^(-8S) if A was truncated, else 0; dropped = the slice pairs left out, zero planes not counted).
    // An entry passes when this is <= tol * entry.  A 0/1 operand is exact and costs nothing here.
    const bool certify = uncertified != nullptr && cert_tol > 0.0;
    const double trunc = ldexp(1.0, -8 * S) * 1.0000001;
    q.certA = certify && (planesA >> 31) ? trunc / cert_tol : 0.0;
    q.certB = certify && (planesB >> 31) ? trunc / cert_tol : 0.0;
    q.certD = certify ? dropped * double(K) / cert_tol : 0.0;
    q.rsum = rsum;
    q.csum = csum;
    q.zero_ok = (hflags & 4) ? 0 : 1;
    q.cert_fail = certify ? cert_fail : nullptr;
    SS_TRY(launch_umma<1>(ctx, maps, q, 2.0 * double(M) * double(N) * double(K)));
    if (uncertified) *uncertified = 0;
    if (certify) {
        unsigned long long h = 0;
        SS_CHECK_CUDA(cudaMemcpyAsync(&h, cert_fail, 8, cudaMemcpyDeviceToHost, ctx->stream));
        SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        *uncertified = int64_t(h);
    }
    return SS_OK;
}

}  // namespace ss
