// FP64 chain-product GEMM for sm_100a: C = op(A) * B with the degree normalisation fused into the
// epilogue.  This is the kernel behind both products of predict (`Aarr * Warr^2`,
// reference src/core.jl:413 and :456) after block reduction:
//     T = (Xs' * Wst) ./ kf        op(A) = Xs'  (k-major A),  B = Wst = Y ./ ks
//     R =  Xq  * T   (+ clean!)    op(A) = Xq   (m-major A),  B = T
//
// Design (B200):
//   * persistent CTAs (one per SM), 128x128 C tile, K consumed in slabs of 16;
//   * a dedicated producer warp stages A/B slabs with TMA (cp.async.bulk.tensor, SWIZZLE_128B)
//     into a STAGES-deep shared-memory ring guarded by full/empty mbarriers;
//   * 8 consumer warps (64x32 warp tile) feed the FP64 tensor pipe with mma.sync m8n8k4 (SASS
//     DMMA.8x8x4 -- tcgen05 has no f64 kind); fragments are read with conflict-free LDS.128 thanks
//     to a k / row permutation that is consistent between A and B (see frag_* below and
//     tests/test_gemm_layout_sim.py which replays this index math on the CPU);
//   * tiles are rasterised in groups of 16 row-tiles so that the 148 concurrently processed
//     tiles share A/B slabs through L2.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "ss_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 6;
constexpr int CONSUMER_WARPS = 8;
constexpr int WARPS_N = 4;            // warp grid 2 (m) x 4 (n)
constexpr int WM = 64, WN = 32;       // warp tile
constexpr int NT = WN / 8;
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int A_BYTES = BM * BK * 8;  // 16 KB
constexpr int B_BYTES = BN * BK * 8;  // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int GROUP_M = 16;
constexpr int SYNC_CHUNK = 64;  // slabs between lockstep checkpoints
constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
        "{%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}

// rho: row permutation inside an 8-row group of a k-major (128 B per row) swizzled tile that
// makes quarter-warp LDS.128 conflict-free: lanes g=2q,2q+1 must differ in bit 2 of the row.
__device__ __forceinline__ int rho(int g) { return (g >> 1) | ((g & 1) << 2); }

struct TileCoord {
    int tm, tn;
};
__device__ __forceinline__ TileCoord tile_coord(int tile, int tiles_m, int tiles_n) {
    const int group_size = GROUP_M * tiles_n;
    const int gid = tile / group_size;
    const int first_m = gid * GROUP_M;
    const int gm = min(tiles_m - first_m, GROUP_M);
    const int r = tile - gid * group_size;
    return {first_m + r % gm, r / gm};
}

// Work units.  Whole waves are 128 x 128 tiles; the tiles of the last, partial wave are cut into `tail_split` row
// bands of 128 / tail_split rows so that the SMs a partial wave would leave idle get work too (C3 shape: 640 tiles
// = 4 waves + 48 tiles -> 96 half tiles, 4.5 instead of 5 tile times).  A band is computed by the same 2 x 4 warp
// grid with a shorter warp tile; every output element still accumulates k = 0 .. K-1 in the same DMMA sequence, so
// the result does not depend on the split (bit-identical to whole tiles).
struct Unit {
    int tm, tn, band, split;
};

struct GemmParams {
    double* C;
    int64_t ldc;
    int M, N, K;
    int tiles_m, tiles_n;
    int full_tiles;           // units [0, full_tiles) are whole 128 x 128 tiles (whole waves of them)
    int tail_split;           // the remaining tiles are cut into 1 / 2 / 4 row bands, one band per work unit
    int reg_rows;             // row tiles enumerated by tile_coord: tiles_m, or tiles_m - 1 when the last row tile is
    int reg_tiles;            //   short (few valid rows); reg_tiles = reg_rows * tiles_n
    int short_bands;          // bands that cover the valid rows of a short last row tile (0: no such tiles); those
    int total_units;          //   tiles come last in the unit list, short_bands units each
    const int32_t* row_div;   // optional: C[m,:] = acc / row_div[m] (0 when row_div[m] == 0)
    const int32_t* col_flag;  // optional: C[:,n] = -99 when col_flag[n] == 0
    int accumulate;           // C += result
    int cvec;                 // C (and ldc) allow 16-byte vector stores
    int* sync_prog;           // optional: per-CTA checkpoint counters for the loose lockstep (see producer)
    int nmirror;              // fused all-gather: every C element is also stored to these peer-GPU
    double* mirror[7];        // copies of C (same ld), over NVLink P2P, straight from the epilogue
};

// Unit list: whole tiles [0, full_tiles), then the bands of the remaining regular tiles, then the bands of the short
// tiles of the last row tile (only the bands that hold valid rows).
__device__ __forceinline__ Unit unit_of(int u, const GemmParams& p) {
    if (u < p.full_tiles) {
        const TileCoord tc = tile_coord(u, p.reg_rows, p.tiles_n);
        return {tc.tm, tc.tn, 0, 1};
    }
    const int v = u - p.full_tiles;
    const int nreg = (p.reg_tiles - p.full_tiles) * p.tail_split;
    if (v < nreg) {
        const TileCoord tc = tile_coord(p.full_tiles + v / p.tail_split, p.reg_rows, p.tiles_n);
        return {tc.tm, tc.tn, v % p.tail_split, p.tail_split};
    }
    const int w = v - nreg;
    return {p.tiles_m - 1, w / p.short_bands, w % p.short_bands, p.tail_split};
}

__device__ __forceinline__ void store1(const GemmParams& p, int64_t off, double v) {
    p.C[off] = v;
    for (int i = 0; i < p.nmirror; ++i) p.mirror[i][off] = v;
}
__device__ __forceinline__ void store2(const GemmParams& p, int64_t off, double2 v) {
    *reinterpret_cast<double2*>(p.C + off) = v;
    for (int i = 0; i < p.nmirror; ++i) *reinterpret_cast<double2*>(p.mirror[i] + off) = v;
}

__device__ __forceinline__ double finish(double acc, int row, int col, const GemmParams& p,
                                         const double* cptr) {
    double v = acc;
    if (p.row_div) {
        const int d = __ldg(p.row_div + row);
        v = d ? v / double(d) : 0.0;  // true division as in W = G ./ k(G); k == 0 -> 0
    }
    if (p.accumulate) v += *cptr;
    if (p.col_flag && __ldg(p.col_flag + col) == 0) v = -99.0;
    return v;
}

// ---- whole-tile kernel ---------------------------------------------------------------------------------------------
// Launches that need no row bands (every large shape: C4 is 2 066 whole waves) run this kernel, which is the band
// kernel below reduced to whole 128 x 128 tiles with the per-thread offsets hoisted out of the tile loop.  Kept as its
// own kernel because ptxas schedules its K loop slightly better: 36.4 against 35.9 TFLOP/s on the C4 R product.
constexpr int MT_WHOLE = WM / 8;

// A_MMAJOR: A is M x K column-major (m contiguous) staged as 8 boxes [16 k][16 m] per slab.
// !A_MMAJOR: A is stored K x M column-major (k contiguous) staged as one box [128 m][16 k].
template <bool A_MMAJOR>
__global__ void __launch_bounds__(THREADS, 1)
    ss_dgemm_whole_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;  // full[STAGES], empty[STAGES]
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int kblocks = (p.K + BK - 1) / BK;
    const int total_tiles = p.tiles_m * p.tiles_n;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_base + 8 * s, 1);                          // full: producer's expect_tx
            mbar_init(bar_base + 8 * (STAGES + s), CONSUMER_WARPS);  // empty: one arrive per warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CONSUMER_WARPS) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int checkpoint = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const TileCoord tc = tile_coord(tile, p.tiles_m, p.tiles_n);
                const int m0 = tc.tm * BM, n0 = tc.tn * BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    if (p.sync_prog && (kb % SYNC_CHUNK) == 0) {
                        // Loose lockstep.  The 148 CTAs of a wave share ~25 A/B panels through L2
                        // only while they stream K at nearby positions; left alone they drift
                        // apart, every slab is re-fetched from HBM (measured at C4: 10.6 TB
                        // instead of ~1 TB) and the high fill rate shortens L2 residency further.
                        // Each producer publishes a checkpoint count every SYNC_CHUNK slabs and
                        // may run at most one checkpoint ahead of the slowest CTA, so jitter
                        // averages out instead of adding up as with a hard per-tile barrier.
                        // The wait is bounded: a CTA that is not co-resident cannot dead-lock us.
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
                    const uint32_t full = bar_base + 8 * stage;
                    const uint32_t empty = bar_base + 8 * (STAGES + stage);
                    mbar_wait(empty, phase ^ 1);
                    mbar_expect_tx(full, STAGE_BYTES);
                    const uint32_t sA = smem_base + stage * STAGE_BYTES;
                    const uint32_t sB = sA + A_BYTES;
                    if (A_MMAJOR) {
#pragma unroll
                        for (int b = 0; b < BM / 16; ++b)
                            tma_load_2d(sA + b * 2048, &mapA, m0 + b * 16, kb * BK, full);
                    } else {
                        tma_load_2d(sA, &mapA, kb * BK, m0, full);
                    }
                    tma_load_2d(sB, &mapB, kb * BK, n0, full);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            if (p.sync_prog) *reinterpret_cast<volatile int*>(p.sync_prog + blockIdx.x) = 0x7fffffff;
        }
        return;
    }

    // ================= DMMA consumers =================
    const int g = lane >> 2, t = lane & 3;
    const int rg = rho(g);
    const int m_warp = (warp / WARPS_N) * WM;
    const int n_warp = (warp % WARPS_N) * WN;

    // per-thread shared-memory offsets (bytes, relative to the stage's A / B base)
    // B (k-major): row n = n_warp + 8j + rho(g); chunk (t + 4h) ^ rho(g)
    uint32_t offB[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) offB[h] = (n_warp + rg) * 128 + (((t + 4 * h) ^ rg) << 4);
    // A k-major: row m = m_warp + 8i + rho(g), same chunk rule
    uint32_t offAk[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) offAk[h] = (m_warp + rg) * 128 + (((t + 4 * h) ^ rg) << 4);
    // A m-major: block (m_warp/16 + b), row k = 2t + (s&1) + 8(s>>1), chunk g ^ (k & 7)
    uint32_t offAm[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int k = 2 * t + (s & 1) + 8 * (s >> 1);
        offAm[s] = (m_warp >> 4) * 2048 + k * 128 + ((g ^ (k & 7)) << 4);
    }

    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = tile_coord(tile, p.tiles_m, p.tiles_n);
        const int m0 = tc.tm * BM, n0 = tc.tn * BN;

        double acc[MT_WHOLE][NT][2];
#pragma unroll
        for (int i = 0; i < MT_WHOLE; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(bar_base + 8 * stage, phase);
            const uint32_t sA = smem_base + stage * STAGE_BYTES;
            const uint32_t sB = sA + A_BYTES;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double2 bf[NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) bf[j] = lds128(sB + offB[h] + j * 8 * 128);
                if (A_MMAJOR) {
#pragma unroll
                    for (int ss2 = 0; ss2 < 2; ++ss2) {
                        const int s = 2 * h + ss2;
                        double2 af[MT_WHOLE / 2];
#pragma unroll
                        for (int b = 0; b < MT_WHOLE / 2; ++b) af[b] = lds128(sA + offAm[s] + b * 2048);
#pragma unroll
                        for (int b = 0; b < MT_WHOLE / 2; ++b)
#pragma unroll
                            for (int j = 0; j < NT; ++j) {
                                const double bv = ss2 ? bf[j].y : bf[j].x;
                                dmma(acc[2 * b][j], af[b].x, bv);
                                dmma(acc[2 * b + 1][j], af[b].y, bv);
                            }
                    }
                } else {
                    double2 af[MT_WHOLE];
#pragma unroll
                    for (int i = 0; i < MT_WHOLE; ++i) af[i] = lds128(sA + offAk[h] + i * 8 * 128);
#pragma unroll
                    for (int ss2 = 0; ss2 < 2; ++ss2)
#pragma unroll
                        for (int i = 0; i < MT_WHOLE; ++i)
#pragma unroll
                            for (int j = 0; j < NT; ++j)
                                dmma(acc[i][j], ss2 ? af[i].y : af[i].x, ss2 ? bf[j].y : bf[j].x);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_base + 8 * (STAGES + stage));
            if (++stage == STAGES) {
                stage = 0;
                phase ^= 1;
            }
        }

        // ---- epilogue: registers -> global (column-major C), fused normalisation / clean! ----
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = n0 + n_warp + 8 * j + t + 4 * e;  // rho(2t+e) = t + 4e
                if (col >= p.N) continue;
                const int64_t coff = int64_t(col) * p.ldc;
                const double* ccol = p.C + coff;
                if (A_MMAJOR) {
#pragma unroll
                    for (int b = 0; b < MT_WHOLE / 2; ++b) {
                        const int row = m0 + m_warp + 16 * b + 2 * g;  // rows (row, row+1)
                        if (row + 1 < p.M && p.cvec) {
                            double2 v;
                            v.x = finish(acc[2 * b][j][e], row, col, p, ccol + row);
                            v.y = finish(acc[2 * b + 1][j][e], row + 1, col, p, ccol + row + 1);
                            store2(p, coff + row, v);
                        } else {
                            if (row < p.M) store1(p, coff + row, finish(acc[2 * b][j][e], row, col, p, ccol + row));
                            if (row + 1 < p.M)
                                store1(p, coff + row + 1,
                                       finish(acc[2 * b + 1][j][e], row + 1, col, p, ccol + row + 1));
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < MT_WHOLE; ++i) {
                        const int row = m0 + m_warp + 8 * i + rg;
                        if (row < p.M) store1(p, coff + row, finish(acc[i][j][e], row, col, p, ccol + row));
                    }
                }
            }
    }
}

// One work unit of the consumer warps: the (MT_ * 16) x 128 row band `band` of tile (m0, n0), warp tile (MT_ * 8) x 32.
// MT_ = 8 is the whole tile.
template <bool A_MMAJOR, int MT_>
__device__ __forceinline__ void consume_unit(const GemmParams& p, uint32_t smem_base, uint32_t bar_base, int m0, int n0,
                                             int band, int kblocks, int warp, int lane, int& stage, uint32_t& phase) {
    const int g = lane >> 2, t = lane & 3;
    const int rg = rho(g);
    const int m_warp = band * (2 * MT_ * 8) + (warp / WARPS_N) * (MT_ * 8);
    const int n_warp = (warp % WARPS_N) * WN;

    // per-thread shared-memory offsets (bytes, relative to the stage's A / B base)
    // B (k-major): row n = n_warp + 8j + rho(g); chunk (t + 4h) ^ rho(g)
    uint32_t offB[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) offB[h] = (n_warp + rg) * 128 + (((t + 4 * h) ^ rg) << 4);
    // A k-major: row m = m_warp + 8i + rho(g), same chunk rule
    uint32_t offAk[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) offAk[h] = (m_warp + rg) * 128 + (((t + 4 * h) ^ rg) << 4);
    // A m-major: block (m_warp/16 + b), row k = 2t + (s&1) + 8(s>>1), chunk g ^ (k & 7)
    uint32_t offAm[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int k = 2 * t + (s & 1) + 8 * (s >> 1);
        offAm[s] = (m_warp >> 4) * 2048 + k * 128 + ((g ^ (k & 7)) << 4);
    }

    double acc[MT_][NT][2];
#pragma unroll
    for (int i = 0; i < MT_; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(bar_base + 8 * stage, phase);
        const uint32_t sA = smem_base + stage * STAGE_BYTES;
        const uint32_t sB = sA + A_BYTES;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double2 bf[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[j] = lds128(sB + offB[h] + j * 8 * 128);
            if (A_MMAJOR) {
#pragma unroll
                for (int ss2 = 0; ss2 < 2; ++ss2) {
                    const int s = 2 * h + ss2;
                    double2 af[MT_ / 2];
#pragma unroll
                    for (int b = 0; b < MT_ / 2; ++b) af[b] = lds128(sA + offAm[s] + b * 2048);
#pragma unroll
                    for (int b = 0; b < MT_ / 2; ++b)
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            const double bv = ss2 ? bf[j].y : bf[j].x;
                            dmma(acc[2 * b][j], af[b].x, bv);
                            dmma(acc[2 * b + 1][j], af[b].y, bv);
                        }
                }
            } else {
                double2 af[MT_];
#pragma unroll
                for (int i = 0; i < MT_; ++i) af[i] = lds128(sA + offAk[h] + i * 8 * 128);
#pragma unroll
                for (int ss2 = 0; ss2 < 2; ++ss2)
#pragma unroll
                    for (int i = 0; i < MT_; ++i)
#pragma unroll
                        for (int j = 0; j < NT; ++j)
                            dmma(acc[i][j], ss2 ? af[i].y : af[i].x, ss2 ? bf[j].y : bf[j].x);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_base + 8 * (STAGES + stage));
        if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
        }
    }

    // ---- epilogue: registers -> global (column-major C), fused normalisation / clean! ----
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int col = n0 + n_warp + 8 * j + t + 4 * e;  // rho(2t+e) = t + 4e
            if (col >= p.N) continue;
            const int64_t coff = int64_t(col) * p.ldc;
            const double* ccol = p.C + coff;
            if (A_MMAJOR) {
#pragma unroll
                for (int b = 0; b < MT_ / 2; ++b) {
                    const int row = m0 + m_warp + 16 * b + 2 * g;  // rows (row, row+1)
                    if (row + 1 < p.M && p.cvec) {
                        double2 v;
                        v.x = finish(acc[2 * b][j][e], row, col, p, ccol + row);
                        v.y = finish(acc[2 * b + 1][j][e], row + 1, col, p, ccol + row + 1);
                        store2(p, coff + row, v);
                    } else {
                        if (row < p.M) store1(p, coff + row, finish(acc[2 * b][j][e], row, col, p, ccol + row));
                        if (row + 1 < p.M)
                            store1(p, coff + row + 1,
                                   finish(acc[2 * b + 1][j][e], row + 1, col, p, ccol + row + 1));
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < MT_; ++i) {
                    const int row = m0 + m_warp + 8 * i + rg;
                    if (row < p.M) store1(p, coff + row, finish(acc[i][j][e], row, col, p, ccol + row));
                }
            }
        }
}

// A_MMAJOR: A is M x K column-major (m contiguous) staged as 8 boxes [16 k][16 m] per slab.
// !A_MMAJOR: A is stored K x M column-major (k contiguous) staged as one box [128 m][16 k].
template <bool A_MMAJOR>
__global__ void __launch_bounds__(THREADS, 1)
    ss_dgemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;  // full[STAGES], empty[STAGES]
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int kblocks = (p.K + BK - 1) / BK;
    const int total_units = p.total_units;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_base + 8 * s, 1);                          // full: producer's expect_tx
            mbar_init(bar_base + 8 * (STAGES + s), CONSUMER_WARPS);  // empty: one arrive per warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CONSUMER_WARPS) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int checkpoint = 0;
            for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
                const Unit un = unit_of(u, p);
                const Unit& tc = un;
                const int n0 = tc.tn * BN;
                // rows of A this unit needs: the whole tile, or one band of it (multiples of 16 rows)
                const int band_rows = BM / un.split;
                const int m0 = tc.tm * BM + un.band * band_rows;
                const uint32_t a_off = uint32_t(un.band * band_rows) * (BK * 8);  // both layouts: 128 bytes per row
                for (int kb = 0; kb < kblocks; ++kb) {
                    if (p.sync_prog && (kb % SYNC_CHUNK) == 0) {
                        // Loose lockstep.  The 148 CTAs of a wave share ~25 A/B panels through L2
                        // only while they stream K at nearby positions; left alone they drift
                        // apart, every slab is re-fetched from HBM (measured at C4: 10.6 TB
                        // instead of ~1 TB) and the high fill rate shortens L2 residency further.
                        // Each producer publishes a checkpoint count every SYNC_CHUNK slabs and
                        // may run at most one checkpoint ahead of the slowest CTA, so jitter
                        // averages out instead of adding up as with a hard per-tile barrier.
                        // The wait is bounded: a CTA that is not co-resident cannot dead-lock us.
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
                    const uint32_t full = bar_base + 8 * stage;
                    const uint32_t empty = bar_base + 8 * (STAGES + stage);
                    mbar_wait(empty, phase ^ 1);
                    const uint32_t sA = smem_base + stage * STAGE_BYTES;
                    const uint32_t sB = sA + A_BYTES;
                    if (un.split == 1) {
                        mbar_expect_tx(full, STAGE_BYTES);
                        if (A_MMAJOR) {
#pragma unroll
                            for (int b = 0; b < BM / 16; ++b)
                                tma_load_2d(sA + b * 2048, &mapA, m0 + b * 16, kb * BK, full);
                        } else {
                            tma_load_2d(sA, &mapA, kb * BK, m0, full);
                        }
                    } else if (A_MMAJOR) {
                        // a band is band_rows / 16 of the 16-row boxes, placed where the whole tile would have them
                        mbar_expect_tx(full, B_BYTES + band_rows * BK * 8);
                        for (int b = 0; b < band_rows / 16; ++b)
                            tma_load_2d(sA + a_off + b * 2048, &mapA, m0 + b * 16, kb * BK, full);
                    } else {
                        // k-major A has one 128-row box: load it whole, the consumers read their band of it
                        mbar_expect_tx(full, STAGE_BYTES);
                        tma_load_2d(sA, &mapA, kb * BK, tc.tm * BM, full);
                    }
                    tma_load_2d(sB, &mapB, kb * BK, n0, full);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            if (p.sync_prog) *reinterpret_cast<volatile int*>(p.sync_prog + blockIdx.x) = 0x7fffffff;
        }
        return;
    }

    // ================= DMMA consumers =================
    int stage = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const Unit un = unit_of(u, p);
        const int m0 = un.tm * BM, n0 = un.tn * BN;
        if (un.split == 1)
            consume_unit<A_MMAJOR, 8>(p, smem_base, bar_base, m0, n0, 0, kblocks, warp, lane, stage, phase);
        else if (un.split == 2)
            consume_unit<A_MMAJOR, 4>(p, smem_base, bar_base, m0, n0, un.band, kblocks, warp, lane, stage, phase);
        else
            consume_unit<A_MMAJOR, 2>(p, smem_base, bar_base, m0, n0, un.band, kblocks, warp, lane, stage, phase);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D FP64 tensor map: dim0 (contiguous) x dim1 with row pitch ld elements, SWIZZLE_128B boxes.
int32_t make_map(CUtensorMap* map, const double* base, int64_t dim0, int64_t dim1, int64_t ld,
                 int box0, int box1) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        ss::set_error("cuTensorMapEncodeTiled is not available from the driver");
        return SS_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {cuuint64_t(dim0), cuuint64_t(dim1)};
    cuuint64_t gstride[1] = {cuuint64_t(ld) * 8};
    cuuint32_t box[2] = {cuuint32_t(box0), cuuint32_t(box1)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim,
                     gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ss::set_error("cuTensorMapEncodeTiled failed with CUresult %d (dims %lld x %lld, ld %lld)",
                      int(r), (long long)dim0, (long long)dim1, (long long)ld);
        return SS_ERR_CUDA;
    }
    return SS_OK;
}

}  // namespace

namespace ss {

static int32_t launch_gemm_f64_one(ss_ctx* ctx, int opA, const double* A, int64_t lda, const double* B,
                                   int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                                   const int32_t* row_div, const int32_t* col_flag, bool accumulate, int nmirror,
                                   double* const* mirrors) {
    SS_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem (M=%lld N=%lld K=%lld)", (long long)M,
               (long long)N, (long long)K);
    SS_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm: dimension too large");
    SS_REQUIRE((lda % 2) == 0 && (ldb % 2) == 0, "gemm: lda / ldb must be even (16-byte TMA alignment)");
    SS_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
               "gemm: A and B must be 16-byte aligned");
    SS_REQUIRE((reinterpret_cast<uintptr_t>(C) & 7) == 0, "gemm: C must be 8-byte aligned");
    CUtensorMap mapA, mapB;
    if (opA == SS_OP_N) {
        SS_TRY(make_map(&mapA, A, M, K, lda, 16, 16));
    } else {
        SS_TRY(make_map(&mapA, A, K, M, lda, 16, BM));
    }
    SS_TRY(make_map(&mapB, B, K, N, ldb, 16, BN));
    GemmParams p;
    p.C = C;
    p.ldc = ldc;
    p.M = int(M);
    p.N = int(N);
    p.K = int(K);
    p.tiles_m = int(ceil_div(M, BM));
    p.tiles_n = int(ceil_div(N, BN));
    p.row_div = row_div;
    p.col_flag = col_flag;
    p.accumulate = accumulate ? 1 : 0;
    p.cvec = ((reinterpret_cast<uintptr_t>(C) & 15) == 0 && (ldc % 2) == 0) ? 1 : 0;
    p.sync_prog = nullptr;
    SS_REQUIRE(nmirror >= 0 && nmirror <= 7 && !(nmirror && accumulate), "gemm: bad mirror request");
    p.nmirror = nmirror;
    for (int i = 0; i < 7; ++i) p.mirror[i] = (i < nmirror) ? mirrors[i] : nullptr;
    for (int i = 0; i < nmirror; ++i)
        if (reinterpret_cast<uintptr_t>(mirrors[i]) & 15) p.cvec = 0;
    const int64_t total = int64_t(p.tiles_m) * p.tiles_n;
    SS_REQUIRE(total < (1ll << 31), "gemm: too many tiles");
    // Partial last wave: cut the trailing tiles into 2 or 4 row bands when that shortens the launch (see Unit).  The
    // candidates are the tiles of the partial wave alone, or together with the last whole wave; a band costs a little
    // more per row than a whole tile (the B slab is re-read per band), hence the 4 % / 8 % handicap.  With bands a
    // last row tile that holds few valid rows (M = 5000: 8 of 128) only gets the bands that hold rows.
    // SS_GEMM_TAIL_SPLIT=0 keeps whole tiles, =2 allows halves only (A/B measurements, tests).
    const int64_t sms = ctx->sm_count;
    int split = 1, short_bands = 0;
    int64_t full = total, reg_rows = p.tiles_m;
    {
        const char* e = getenv("SS_GEMM_TAIL_SPLIT");
        const int allow = e ? atoi(e) : 4;
        double best = double(ceil_div(total, sms));
        const int64_t valid_last = M - int64_t(p.tiles_m - 1) * BM;  // rows of the last row tile
        if (K >= 64) {
            for (int s = 2; s <= allow && s <= 4; s *= 2) {
                const int64_t band_rows = BM / s;
                const int64_t sb = ceil_div(valid_last, band_rows) < s ? ceil_div(valid_last, band_rows) : 0;  // short?
                const int64_t rr = sb ? p.tiles_m - 1 : p.tiles_m;
                const int64_t reg = rr * p.tiles_n, shorts = sb ? p.tiles_n : 0;
                const int64_t rem = reg % sms;
                for (int j = 0; j <= 1; ++j) {
                    const int64_t f = reg - rem - j * sms;
                    if (f < 0 || (reg - f) * s + shorts * sb == 0) continue;
                    const double cost = double(f / sms) +
                                        double(ceil_div((reg - f) * s + shorts * sb, sms)) / s * (s == 2 ? 1.04 : 1.08);
                    if (cost < best * 0.98) {
                        best = cost;
                        split = s;
                        full = f;
                        short_bands = int(sb);
                        reg_rows = rr;
                    }
                }
            }
        }
    }
    p.tail_split = split;
    p.full_tiles = int(full);
    p.reg_rows = int(reg_rows);
    p.reg_tiles = int(reg_rows * p.tiles_n);
    p.short_bands = short_bands;
    const int64_t units = full + (p.reg_tiles - full) * split + (short_bands ? int64_t(p.tiles_n) * short_bands : 0);
    SS_REQUIRE(units < (1ll << 31), "gemm: too many work units");
    p.total_units = int(units);
    const int grid = int(units < sms ? units : sms);
    if (units > grid) {  // more than one wave: keep the waves in lockstep for L2 reuse
        if (!ctx->tile_counter) SS_CHECK_CUDA(cudaMalloc(&ctx->tile_counter, 4096));
        SS_CHECK_CUDA(cudaMemsetAsync(ctx->tile_counter, 0, size_t(grid) * 4, ctx->stream));
        p.sync_prog = ctx->tile_counter;
    }
    if (!ctx->gemm_attr_set) {
        SS_CHECK_CUDA(cudaFuncSetAttribute(ss_dgemm_kernel<true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
        SS_CHECK_CUDA(cudaFuncSetAttribute(ss_dgemm_kernel<false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
        SS_CHECK_CUDA(cudaFuncSetAttribute(ss_dgemm_whole_kernel<true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
        SS_CHECK_CUDA(cudaFuncSetAttribute(ss_dgemm_whole_kernel<false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
        ctx->gemm_attr_set = true;
    }
    ss_ctx::ProfRec rec{nullptr, nullptr, 2.0 * double(M) * double(N) * double(K)};
    if (ctx->profile) {
        SS_CHECK_CUDA(cudaEventCreate(&rec.start));
        SS_CHECK_CUDA(cudaEventCreate(&rec.stop));
        SS_CHECK_CUDA(cudaEventRecord(rec.start, ctx->stream));
    }
    const char* force = getenv("SS_GEMM_KERNEL");  // "bands": the band kernel for whole tiles too (A/B measurements)
    if (split == 1 && !(force && !strcmp(force, "bands"))) {
        if (opA == SS_OP_N)
            ss_dgemm_whole_kernel<true><<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(mapA, mapB, p);
        else
            ss_dgemm_whole_kernel<false><<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(mapA, mapB, p);
    } else if (opA == SS_OP_N) {
        ss_dgemm_kernel<true><<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(mapA, mapB, p);
    } else {
        ss_dgemm_kernel<false><<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(mapA, mapB, p);
    }
    SS_CHECK_CUDA(cudaGetLastError());
    if (ctx->profile) {
        SS_CHECK_CUDA(cudaEventRecord(rec.stop, ctx->stream));
        ctx->prof.push_back(rec);
    }
    ctx->launches++;
    return SS_OK;
}

// K-blocking (off by default, SS_GEMM_KBLOCK=<k>): the product is run as ceil(K / k) launches that accumulate into C, so
// that the A panels of one wave (16 row tiles x 128 rows x k x 8 B) stay in L2 between the column tiles that re-read
// them.  The split depends on K alone, so every caller (whole matrix, slab, shard) rounds the same way.  Not used with
// the row division (it applies to the complete sum) nor with mirrors.
int32_t launch_gemm_f64(ss_ctx* ctx, int opA, const double* A, int64_t lda, const double* B,
                        int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                        const int32_t* row_div, const int32_t* col_flag, bool accumulate, int nmirror,
                        double* const* mirrors) {
    int64_t kb = 0;
    if (const char* env = getenv("SS_GEMM_KBLOCK")) kb = atoll(env) & ~int64_t(15);
    if (kb < 256 || K < 2 * kb || row_div || nmirror)
        return launch_gemm_f64_one(ctx, opA, A, lda, B, ldb, C, ldc, M, N, K, row_div, col_flag, accumulate, nmirror,
                                   mirrors);
    const int64_t nblk = (K + kb - 1) / kb;
    const int64_t step = (((K + nblk - 1) / nblk) + 15) & ~int64_t(15);  // even blocks, 16-aligned starts
    for (int64_t k0 = 0; k0 < K; k0 += step) {
        const int64_t kk = (K - k0 < step) ? (K - k0) : step;
        const bool last = (k0 + kk >= K);
        const double* Ab = (opA == SS_OP_N) ? A + k0 * lda : A + k0;
        SS_TRY(launch_gemm_f64_one(ctx, opA, Ab, lda, B + k0, ldb, C, ldc, M, N, kk, nullptr, last ? col_flag : nullptr,
                                   accumulate || k0 > 0, 0, nullptr));
    }
    return SS_OK;
}

}  // namespace ss
