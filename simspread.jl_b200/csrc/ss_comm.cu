// Multi-GPU exchange steps behind the C ABI (SURVEY.md 8b `ss_comm_init`, 8e): one process per GPU, NCCL over
// NVLink for the two small collectives of the sharded predict (all-reduce of the source degrees, all-gather of the
// target degrees) and a fused GEMM + all-gather for the T tiles: every rank's T lives in cudaMalloc memory whose CUDA
// IPC handle is mapped by the peers, and the T-GEMM epilogue stores each tile into the local T and the same offsets
// of all peers over NVLink P2P (ss_gemm.cu, `mirrors`).  A host language needs nothing but this library: NCCL is
// dlopen()ed (libnccl.so.2 -- the copy already loaded by the process if there is one), the unique id travels through
// whatever channel the host has (ss_comm_unique_id + ss_comm_init) or through a file (ss_comm_init_file).
//
// Sharding of R = Xq * T, T = (Xs' * (Y ./ ks)) ./ kf [reference src/core.jl:402-423 on the blocks of App. B]:
//   rank r owns the query rows Xq[r] / R[r] (no exchange) and the target-column block Y[:, r]; Xs is replicated;
//   ks = nnz_row(Xs) + sum_r nnz_row(Y[:, r])  -> all-reduce (rank 0 adds the Xs term);
//   kt[r] = nnz_col(Y[:, r])                    -> all-gather (clean! flag of the R epilogue);
//   T[:, r] = (Xs' * (Y[:, r] ./ ks)) ./ kf     -> written into every rank's T by the GEMM epilogue, then a barrier.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <mutex>

#include "ss_common.cuh"

namespace {

// the few NCCL entry points used, declared here so that the build does not depend on nccl.h
typedef void* nccl_comm_t;
struct nccl_unique_id {
    char internal[128];
};
enum { NCCL_INT32 = 2, NCCL_INT64 = 4, NCCL_FLOAT64 = 8, NCCL_UINT8 = 1 };
enum { NCCL_SUM = 0, NCCL_MAX = 2 };
struct NcclApi {
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    void* handle = nullptr;
    bool ok = false;
    char why[256] = "";
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* override_path = getenv("SS_NCCL_LIBRARY");
        const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            snprintf(api.why, sizeof(api.why), "cannot load libnccl.so.2 (%s); set SS_NCCL_LIBRARY", dlerror());
            return;
        }
#define SS_NCCL_SYM(field, name)                                                      \
    *reinterpret_cast<void**>(&api.field) = dlsym(api.handle, name);                  \
    if (!api.field) {                                                                 \
        snprintf(api.why, sizeof(api.why), "libnccl has no symbol %s", name);         \
        return;                                                                       \
    }
        SS_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        SS_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        SS_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        SS_NCCL_SYM(GetErrorString, "ncclGetErrorString")
        SS_NCCL_SYM(AllReduce, "ncclAllReduce")
        SS_NCCL_SYM(AllGather, "ncclAllGather")
        SS_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef SS_NCCL_SYM
        api.ok = true;
    });
    return api;
}

#define SS_CHECK_NCCL(expr)                                                                               \
    do {                                                                                                  \
        int _r = (expr);                                                                                  \
        if (_r != 0) {                                                                                    \
            ss::set_error("%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(_r), __FILE__, __LINE__); \
            return SS_ERR_CUDA;                                                                           \
        }                                                                                                 \
    } while (0)

}  // namespace

struct ss_comm {
    ss_ctx* ctx = nullptr;
    int rank = 0, world = 1;
    nccl_comm_t nc = nullptr;
    int32_t* scratch = nullptr;  // device: barrier token + staging of small host blobs (64 KiB)
};

struct ss_sharded {
    ss_comm* comm = nullptr;
    int64_t ns = 0, nf = 0, nt = 0, nt_blk = 0, nt_padded = 0, ldt = 0, ldw = 0;
    double* T = nullptr;      // nf x nt_padded, cudaMalloc (IPC-exported)
    double* Wst = nullptr;    // ns x nt_blk
    int32_t *ks = nullptr, *kf = nullptr, *kt_blk = nullptr, *kt = nullptr;
    void* peer_base[8] = {nullptr};  // mapped T of the other ranks (index = rank; own entry null)
    int n_mirrors = 0;
    double* mirrors[7] = {nullptr};  // my column block inside every peer's T
    bool fused = false;
    ss_mat T_view;            // nf x nt
    ss_ivec kt_view;          // nt
};

using namespace ss;

namespace {

int32_t comm_barrier(ss_comm* c) {
    if (c->world > 1) SS_CHECK_NCCL(nccl().AllReduce(c->scratch, c->scratch, 1, NCCL_INT32, NCCL_SUM, c->nc, c->ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(c->ctx->stream));
    return SS_OK;
}

// small host blobs: send (bytes) from every rank -> recv (bytes * world), staged through the device scratch
int32_t comm_allgather_host(ss_comm* c, const void* send, void* recv, int64_t bytes) {
    SS_REQUIRE(bytes > 0 && bytes * (c->world + 1) <= 60 * 1024, "ss_comm_allgather_host: blob too large");
    char* base = reinterpret_cast<char*>(c->scratch) + 1024;
    char* dsend = base;
    char* drecv = base + ((bytes + 255) & ~int64_t(255));
    SS_CHECK_CUDA(cudaMemcpyAsync(dsend, send, size_t(bytes), cudaMemcpyHostToDevice, c->ctx->stream));
    if (c->world > 1) {
        SS_CHECK_NCCL(nccl().AllGather(dsend, drecv, size_t(bytes), NCCL_UINT8, c->nc, c->ctx->stream));
    } else {
        SS_CHECK_CUDA(cudaMemcpyAsync(drecv, dsend, size_t(bytes), cudaMemcpyDeviceToDevice, c->ctx->stream));
    }
    SS_CHECK_CUDA(cudaMemcpyAsync(recv, drecv, size_t(bytes) * c->world, cudaMemcpyDeviceToHost, c->ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(c->ctx->stream));
    return SS_OK;
}

void sharded_free(ss_sharded* p) {
    if (!p) return;
    ss_ctx* ctx = p->comm->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < 8; ++r)
        if (p->peer_base[r]) cudaIpcCloseMemHandle(p->peer_base[r]);
    if (p->T) cudaFree(p->T);
    if (p->Wst) cudaFree(p->Wst);
    if (p->ks) cudaFree(p->ks);
    delete p;
}

}  // namespace

extern "C" {

int32_t ss_comm_unique_id(void* id128_out) {
    SS_REQUIRE(id128_out, "ss_comm_unique_id: null output");
    SS_REQUIRE(nccl().ok, "ss_comm_unique_id: %s", nccl().why);
    nccl_unique_id id;
    SS_CHECK_NCCL(nccl().GetUniqueId(&id));
    memcpy(id128_out, &id, 128);
    return SS_OK;
}

int32_t ss_comm_init(ss_ctx* ctx, int32_t rank, int32_t world, const void* id128, ss_comm** out) {
    SS_REQUIRE(ctx && out, "ss_comm_init: null argument");
    *out = nullptr;
    SS_CHECK_CUDA(cudaSetDevice(ctx->device));
    SS_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "ss_comm_init: rank %d of %d (1..8 ranks of one node)", rank, world);
    SS_REQUIRE(world == 1 || id128, "ss_comm_init: the NCCL unique id is required for more than one rank");
    ss_comm* c = new ss_comm();
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    if (cudaMalloc(&c->scratch, 64 * 1024) != cudaSuccess) {
        cudaGetLastError();
        delete c;
        set_error("ss_comm_init: out of device memory");
        return SS_ERR_OOM;
    }
    cudaMemsetAsync(c->scratch, 0, 64 * 1024, ctx->stream);
    if (world > 1) {
        if (!nccl().ok) {
            set_error("ss_comm_init: %s", nccl().why);
            cudaFree(c->scratch);
            delete c;
            return SS_ERR_UNSUPPORTED;
        }
        nccl_unique_id id;
        memcpy(&id, id128, 128);
        const int r = nccl().CommInitRank(&c->nc, world, id, rank);
        if (r != 0) {
            set_error("ncclCommInitRank failed: %s", nccl().GetErrorString(r));
            cudaFree(c->scratch);
            delete c;
            return SS_ERR_CUDA;
        }
    }
    *out = c;
    const int32_t st = comm_barrier(c);
    if (st != SS_OK) return st;
    return SS_OK;
}

/* Rendezvous through a file for hosts without their own channel: rank 0 writes the unique id to `path`
 * (temporary file + rename), the other ranks poll for it.  `path` must be unique to the job. */
int32_t ss_comm_init_file(ss_ctx* ctx, int32_t rank, int32_t world, const char* path, double timeout_s, ss_comm** out) {
    SS_REQUIRE(ctx && out && (world == 1 || path), "ss_comm_init_file: null argument");
    char id[128] = {0};
    if (world > 1 && rank == 0) {
        SS_TRY(ss_comm_unique_id(id));
        char tmp[4096];
        snprintf(tmp, sizeof(tmp), "%s.tmp.%d", path, int(getpid()));
        FILE* f = fopen(tmp, "wb");
        SS_REQUIRE(f, "ss_comm_init_file: cannot write %s", tmp);
        const size_t w = fwrite(id, 1, 128, f);
        fclose(f);
        SS_REQUIRE(w == 128 && rename(tmp, path) == 0, "ss_comm_init_file: cannot publish %s", path);
    } else if (world > 1) {
        const double t0 = double(time(nullptr));
        for (;;) {
            struct stat sb;
            if (stat(path, &sb) == 0 && sb.st_size == 128) {
                FILE* f = fopen(path, "rb");
                if (f) {
                    const size_t r = fread(id, 1, 128, f);
                    fclose(f);
                    if (r == 128) break;
                }
            }
            SS_REQUIRE(double(time(nullptr)) - t0 <= timeout_s, "ss_comm_init_file: no unique id at %s after %.0f s", path, timeout_s);
            usleep(20000);
        }
    }
    const int32_t st = ss_comm_init(ctx, rank, world, id, out);
    if (st == SS_OK && world > 1 && rank == 0) unlink(path);  // every rank has joined: the id is spent
    return st;
}

int32_t ss_comm_destroy(ss_comm* c) {
    if (!c) return SS_OK;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->nc) nccl().CommDestroy(c->nc);
    if (c->scratch) cudaFree(c->scratch);
    delete c;
    return SS_OK;
}

int32_t ss_comm_info(const ss_comm* c, int32_t* rank, int32_t* world, int32_t* nccl_version) {
    SS_REQUIRE(c, "ss_comm_info: null communicator");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (nccl_version) {
        int v = 0;
        if (nccl().ok) nccl().GetVersion(&v);
        *nccl_version = v;
    }
    return SS_OK;
}

int32_t ss_comm_barrier(ss_comm* c) {
    SS_REQUIRE(c, "ss_comm_barrier: null communicator");
    SS_CHECK_CUDA(cudaSetDevice(c->ctx->device));
    return comm_barrier(c);
}

int32_t ss_comm_allreduce_i32(ss_comm* c, ss_ivec* v) {
    SS_REQUIRE(c && v, "ss_comm_allreduce_i32: null argument");
    SS_CHECK_CUDA(cudaSetDevice(c->ctx->device));
    if (c->world > 1 && v->n) SS_CHECK_NCCL(nccl().AllReduce(v->d, v->d, size_t(v->n), NCCL_INT32, NCCL_SUM, c->nc, c->ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(c->ctx->stream));
    return SS_OK;
}

int32_t ss_comm_allgather_i32(ss_comm* c, const ss_ivec* send, ss_ivec* recv) {
    SS_REQUIRE(c && send && recv && recv->n == send->n * c->world, "ss_comm_allgather_i32: recv must hold world x send entries");
    SS_CHECK_CUDA(cudaSetDevice(c->ctx->device));
    if (c->world > 1) {
        if (send->n) SS_CHECK_NCCL(nccl().AllGather(send->d, recv->d, size_t(send->n), NCCL_INT32, c->nc, c->ctx->stream));
    } else if (send->n) {
        SS_CHECK_CUDA(cudaMemcpyAsync(recv->d, send->d, size_t(send->n) * 4, cudaMemcpyDeviceToDevice, c->ctx->stream));
    }
    SS_CHECK_CUDA(cudaStreamSynchronize(c->ctx->stream));
    return SS_OK;
}

int32_t ss_comm_allgather_host(ss_comm* c, const void* send, void* recv, int64_t bytes) {
    SS_REQUIRE(c && send && recv, "ss_comm_allgather_host: null argument");
    SS_CHECK_CUDA(cudaSetDevice(c->ctx->device));
    return comm_allgather_host(c, send, recv, bytes);
}

/* Replicated operand from sharded uploads: rank r has filled the column block [r * cols/world, (r+1) * cols/world) of
 * M (same shape and ld on every rank, cols divisible by world); afterwards every rank holds all of M.  One NCCL
 * all-gather over NVLink instead of `world` host -> device copies of the whole matrix. */
int32_t ss_comm_allgather_cols(ss_comm* c, ss_mat* M) {
    SS_REQUIRE(c && M, "ss_comm_allgather_cols: null argument");
    SS_CHECK_CUDA(cudaSetDevice(c->ctx->device));
    SS_REQUIRE(M->cols % c->world == 0, "ss_comm_allgather_cols: %lld columns do not divide over %d ranks", (long long)M->cols, c->world);
    if (c->world > 1 && M->cols > 0) {
        const size_t per = size_t(M->cols / c->world) * size_t(M->ld);
        SS_CHECK_NCCL(nccl().AllGather(M->d + per * c->rank, M->d, per, NCCL_FLOAT64, c->nc, c->ctx->stream));
    }
    SS_CHECK_CUDA(cudaStreamSynchronize(c->ctx->stream));
    return SS_OK;
}

/* x[i] <- max / sum over the ranks (host doubles; op 0 = sum, 1 = max): timings, checksums */
int32_t ss_comm_allreduce_host_f64(ss_comm* c, double* x, int32_t n, int32_t op) {
    SS_REQUIRE(c && x && n >= 1 && n <= 1024 && (op == 0 || op == 1), "ss_comm_allreduce_host_f64: bad argument");
    SS_CHECK_CUDA(cudaSetDevice(c->ctx->device));
    double* d = reinterpret_cast<double*>(reinterpret_cast<char*>(c->scratch) + 1024);
    SS_CHECK_CUDA(cudaMemcpyAsync(d, x, size_t(n) * 8, cudaMemcpyHostToDevice, c->ctx->stream));
    if (c->world > 1) SS_CHECK_NCCL(nccl().AllReduce(d, d, size_t(n), NCCL_FLOAT64, op ? NCCL_MAX : NCCL_SUM, c->nc, c->ctx->stream));
    SS_CHECK_CUDA(cudaMemcpyAsync(x, d, size_t(n) * 8, cudaMemcpyDeviceToHost, c->ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(c->ctx->stream));
    return SS_OK;
}

/* ---- sharded predict ------------------------------------------------------------------------------------------ */

int32_t ss_sharded_create(ss_comm* c, int64_t ns, int64_t nf, int64_t nt, ss_sharded** out) {
    SS_REQUIRE(c && out && ns >= 1 && nf >= 1 && nt >= 1, "ss_sharded_create: bad argument");
    *out = nullptr;
    ss_ctx* ctx = c->ctx;
    SS_CHECK_CUDA(cudaSetDevice(ctx->device));
    ss_sharded* p = new ss_sharded();
    p->comm = c;
    p->ns = ns;
    p->nf = nf;
    p->nt = nt;
    p->nt_blk = ceil_div(nt, c->world);
    p->nt_padded = p->nt_blk * c->world;
    p->ldt = round_up(nf, 16);
    p->ldw = round_up(ns, 16);
    const size_t t_bytes = size_t(p->ldt) * p->nt_padded * 8, w_bytes = size_t(p->ldw) * p->nt_blk * 8;
    const size_t k_ints = size_t(round_up(ns, 64) + round_up(nf, 64) + round_up(p->nt_blk, 64) + round_up(p->nt_padded, 64));
    cudaError_t e = cudaMalloc(&p->T, t_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&p->Wst, w_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&p->ks, k_ints * 4);
    if (e != cudaSuccess) {
        cudaGetLastError();
        sharded_free(p);
        set_error("ss_sharded_create: out of device memory (T: %zu MB per rank)", t_bytes >> 20);
        return SS_ERR_OOM;
    }
    p->kf = p->ks + round_up(ns, 64);
    p->kt_blk = p->kf + round_up(nf, 64);
    p->kt = p->kt_blk + round_up(p->nt_blk, 64);
    cudaMemsetAsync(p->T, 0, t_bytes, ctx->stream);
    cudaMemsetAsync(p->Wst, 0, w_bytes, ctx->stream);
    cudaMemsetAsync(p->ks, 0, k_ints * 4, ctx->stream);
    p->T_view.ctx = ctx;
    p->T_view.d = p->T;
    p->T_view.rows = nf;
    p->T_view.cols = nt;
    p->T_view.ld = p->ldt;
    p->kt_view.ctx = ctx;
    p->kt_view.d = p->kt;
    p->kt_view.n = nt;
    // map every peer's T (CUDA IPC over NVLink P2P) unless SS_FUSED_ALLGATHER=0; all ranks must agree
    const char* fa = getenv("SS_FUSED_ALLGATHER");
    int ok = (c->world > 1 && !(fa && !strcmp(fa, "0"))) ? 1 : 0;
    if (c->world > 1) {
        cudaIpcMemHandle_t mine;
        if (cudaIpcGetMemHandle(&mine, p->T) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
            memset(&mine, 0, sizeof(mine));
        }
        char all[8 * 64];
        int32_t st = comm_allgather_host(c, &mine, all, 64);
        if (st != SS_OK) {
            sharded_free(p);
            return st;
        }
        for (int r = 0; ok && r < c->world; ++r) {
            if (r == c->rank) continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, all + 64 * r, 64);
            if (cudaIpcOpenMemHandle(&p->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                p->peer_base[r] = nullptr;
                ok = 0;
            }
        }
        double v = double(ok);
        v = -v;  // min over ranks = -max(-x)
        st = ss_comm_allreduce_host_f64(c, &v, 1, 1);
        if (st != SS_OK) {
            sharded_free(p);
            return st;
        }
        ok = (-v) > 0.5 ? 1 : 0;
    }
    p->fused = ok != 0;
    if (p->fused) {
        const size_t blk_off = size_t(c->rank) * p->nt_blk * p->ldt;
        for (int r = 0; r < c->world; ++r)
            if (r != c->rank) p->mirrors[p->n_mirrors++] = static_cast<double*>(p->peer_base[r]) + blk_off;
    }
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = p;
    return SS_OK;
}

int32_t ss_sharded_destroy(ss_sharded* p) {
    if (p && p->comm->world > 1) comm_barrier(p->comm);  // no rank unmaps while a peer may still store into its T
    sharded_free(p);
    return SS_OK;
}

int32_t ss_sharded_info(const ss_sharded* p, int64_t* nt_blk, int32_t* fused_allgather) {
    SS_REQUIRE(p, "ss_sharded_info: null plan");
    if (nt_blk) *nt_blk = p->nt_blk;
    if (fused_allgather) *fused_allgather = p->fused ? 1 : 0;
    return SS_OK;
}

/* everything up to the assembled T and kt on every rank.  Xs: ns x nf (replicated); Yblk: ns x nt_blk, this rank's
 * target columns [rank * nt_blk, ...) (columns beyond nt must be zero). */
int32_t ss_sharded_front(ss_sharded* p, const ss_mat* Xs, const ss_mat* Yblk) {
    SS_REQUIRE(p && Xs && Yblk, "ss_sharded_front: null argument");
    ss_comm* c = p->comm;
    ss_ctx* ctx = c->ctx;
    SS_CHECK_CUDA(cudaSetDevice(ctx->device));
    SS_REQUIRE(Xs->rows == p->ns && Xs->cols == p->nf && Yblk->rows == p->ns && Yblk->cols == p->nt_blk,
               "ss_sharded_front: Xs must be %lld x %lld and the Y block %lld x %lld", (long long)p->ns, (long long)p->nf,
               (long long)p->ns, (long long)p->nt_blk);
    // degrees: kf from the replicated Xs; ks partial = nnz_row(Y blk) (+ nnz_row(Xs) on rank 0); kt of my block
    SS_CHECK_CUDA(cudaMemsetAsync(p->ks, 0, size_t(p->ns) * 4, ctx->stream));
    SS_CHECK_CUDA(cudaMemsetAsync(p->kf, 0, size_t(p->nf) * 4, ctx->stream));
    SS_CHECK_CUDA(cudaMemsetAsync(p->kt_blk, 0, size_t(p->nt_blk) * 4, ctx->stream));
    SS_TRY(launch_degrees(ctx, Xs->d, p->ns, p->nf, Xs->ld, c->rank == 0 ? p->ks : nullptr, p->kf));
    SS_TRY(launch_degrees(ctx, Yblk->d, p->ns, p->nt_blk, Yblk->ld, p->ks, p->kt_blk));
    if (c->world > 1) {
        SS_CHECK_NCCL(nccl().AllReduce(p->ks, p->ks, size_t(p->ns), NCCL_INT32, NCCL_SUM, c->nc, ctx->stream));
        SS_CHECK_NCCL(nccl().AllGather(p->kt_blk, p->kt, size_t(p->nt_blk), NCCL_INT32, c->nc, ctx->stream));
    } else {
        SS_CHECK_CUDA(cudaMemcpyAsync(p->kt, p->kt_blk, size_t(p->nt_blk) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    double* Tblk = p->T + size_t(c->rank) * p->nt_blk * p->ldt;
    // sparse labels: T block from the edge list of the Y block (csrc/ss_tsparse.cu), stored into every peer's T as well
    bool used = false;
    SS_TRY(t_from_sparse_labels(ctx, Xs->d, Xs->ld, Yblk->d, Yblk->ld, p->ns, p->nf, p->nt_blk, p->ks, p->kf, p->kt_blk, Tblk, p->ldt,
                                p->fused ? p->n_mirrors : 0, p->fused ? p->mirrors : nullptr, &used));
    if (!used) {
        SS_TRY(launch_spread_rows(ctx, Yblk->d, p->ns, p->nt_blk, Yblk->ld, p->ks, p->Wst, p->ldw));
        SS_TRY(launch_gemm_f64(ctx, SS_OP_T, Xs->d, Xs->ld, p->Wst, p->ldw, Tblk, p->ldt, p->nf, p->nt_blk, p->ns, p->kf, nullptr, false,
                               p->fused ? p->n_mirrors : 0, p->fused ? p->mirrors : nullptr));
    }
    if (c->world > 1) {
        if (p->fused) {
            SS_TRY(comm_barrier(c));  // every rank's epilogue has stored its block into every T
        } else {
            SS_CHECK_NCCL(nccl().AllGather(Tblk, p->T, size_t(p->nt_blk) * p->ldt, NCCL_FLOAT64, c->nc, ctx->stream));
        }
    }
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

/* non-owning views of the assembled T (nf x nt) and kt (nt), valid until ss_sharded_destroy */
int32_t ss_sharded_views(ss_sharded* p, ss_mat** T, ss_ivec** kt) {
    SS_REQUIRE(p, "ss_sharded_views: null plan");
    if (T) *T = &p->T_view;
    if (kt) *kt = &p->kt_view;
    return SS_OK;
}

/* one sharded spread + predict step: front, then R slab = Xq slab * T with clean! fused (flags & SS_PREDICT_CLEAN).
 * Xq: nq_local x nf, R: nq_local x nt (a rank may own zero query rows: pass NULL for both). */
int32_t ss_predict_query_sharded(ss_sharded* p, const ss_mat* Xq, const ss_mat* Xs, const ss_mat* Yblk, ss_mat* R, uint32_t flags) {
    SS_REQUIRE(p && Xs && Yblk, "ss_predict_query_sharded: null argument");
    SS_REQUIRE((Xq == nullptr) == (R == nullptr), "ss_predict_query_sharded: Xq and R go together");
    SS_REQUIRE((flags & SS_PRECISION_MASK) == SS_PRECISION_F64, "ss_predict_query_sharded: the sharded chain is FP64 (DMMA) only");
    if (Xq)
        SS_REQUIRE(Xq->cols == p->nf && R->rows == Xq->rows && R->cols == p->nt, "ss_predict_query_sharded: Xq must be nq x %lld and R nq x %lld",
                   (long long)p->nf, (long long)p->nt);
    SS_TRY(ss_sharded_front(p, Xs, Yblk));
    ss_ctx* ctx = p->comm->ctx;
    if (Xq && Xq->rows > 0) {
        SS_TRY(launch_gemm_f64(ctx, SS_OP_N, Xq->d, Xq->ld, p->T, p->ldt, R->d, R->ld, Xq->rows, p->nt, p->nf, nullptr,
                               (flags & SS_PREDICT_CLEAN) ? p->kt : nullptr, false));
        SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return SS_OK;
}

}  // extern "C"
