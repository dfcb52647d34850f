// C ABI of libsimspread_b200.so (see include/simspread_b200.h): contexts, device containers, and
// the orchestration of the predict chain.  Kernels live in ss_elementwise.cu / ss_gemm.cu /
// ss_rank.cu / ss_csr.cu.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "ss_common.cuh"

namespace ss {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int32_t scratch_get(ss_ctx* ctx, int slot, size_t bytes, void** out) {
    Scratch& s = ctx->ws[slot];
    if (s.bytes < bytes) {
        if (s.p) {
            SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
            SS_CHECK_CUDA(cudaFree(s.p));
            s.p = nullptr;
            s.bytes = 0;
        }
        size_t want = (bytes + 255) & ~size_t(255);
        SS_CHECK_CUDA(cudaMalloc(&s.p, want));
        s.bytes = want;
    }
    *out = s.p;
    return SS_OK;
}

}  // namespace ss

using namespace ss;

#define SS_ENTER(ctx)                                                   \
    SS_REQUIRE((ctx) != nullptr, "%s: null context", __func__);         \
    SS_CHECK_CUDA(cudaSetDevice((ctx)->device))

extern "C" {

int32_t ss_version(void) { return SS_VERSION; }
const char* ss_last_error(void) { return g_err; }

int32_t ss_device_count(int32_t* count) {
    SS_REQUIRE(count, "ss_device_count: null output");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return SS_OK;
}

int32_t ss_ctx_create(int32_t device, ss_ctx** out) {
    SS_REQUIRE(out, "ss_ctx_create: null output");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device is visible: libsimspread_b200 has no CPU fallback");
        return SS_ERR_NO_DEVICE;
    }
    SS_REQUIRE(device >= 0 && device < n, "ss_ctx_create: device %d out of range (0..%d)", device, n - 1);
    cudaDeviceProp prop;
    SS_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                  prop.major, prop.minor);
        return SS_ERR_NO_DEVICE;
    }
    SS_CHECK_CUDA(cudaSetDevice(device));
    ss_ctx* c = new ss_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    SS_CHECK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SS_CHECK_CUDA(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
    SS_CHECK_CUDA(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
    {   // keep freed CSR buffers in the default pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    *out = c;
    return SS_OK;
}

int32_t ss_ctx_destroy(ss_ctx* ctx) {
    if (!ctx) return SS_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& s : ctx->ws)
        if (s.p) cudaFree(s.p);
    if (ctx->tile_counter) cudaFree(ctx->tile_counter);
    for (int i = 0; i < 3; ++i)
        if (ctx->stage[i]) {
            cudaFreeHost(ctx->stage[i]);
            cudaEventDestroy(ctx->stage_ev[i]);
        }
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->copy_in);
    cudaStreamDestroy(ctx->copy_out);
    delete ctx;
    return SS_OK;
}

int32_t ss_ctx_sync(ss_ctx* ctx) {
    SS_ENTER(ctx);
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->copy_in));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->copy_out));
    return SS_OK;
}

int32_t ss_ctx_stream(ss_ctx* ctx, void** stream_out) {
    SS_REQUIRE(ctx && stream_out, "ss_ctx_stream: null argument");
    *stream_out = reinterpret_cast<void*>(ctx->stream);
    return SS_OK;
}

int32_t ss_ctx_int8_stats(ss_ctx* ctx, int64_t* stats3) {  // 4 values, see the header
    SS_REQUIRE(ctx && stats3, "ss_ctx_int8_stats: null argument");
    for (int i = 0; i < 3; ++i) stats3[i] = ctx->int8_stats[i];
    stats3[3] = ctx->int8_last_pairs;
    return SS_OK;
}

int32_t ss_ctx_launch_count(ss_ctx* ctx, int64_t* count) {
    SS_REQUIRE(ctx && count, "ss_ctx_launch_count: null argument");
    *count = ctx->launches;
    return SS_OK;
}

int32_t ss_ctx_profile(ss_ctx* ctx, int32_t enable) {
    SS_ENTER(ctx);
    ctx->profile = enable != 0;
    return SS_OK;
}

int32_t ss_ctx_profile_read(ss_ctx* ctx, double* ms_out, double* flops_out, int32_t cap, int32_t* n_out) {
    SS_ENTER(ctx);
    SS_REQUIRE(n_out && cap >= 0, "ss_ctx_profile_read: bad argument");
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    int32_t n = 0;
    for (auto& r : ctx->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.start, r.stop);
        if (n < cap) {
            if (ms_out) ms_out[n] = ms;
            if (flops_out) flops_out[n] = r.flops;
            ++n;
        }
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    ctx->prof.clear();
    *n_out = n;
    return SS_OK;
}

int32_t ss_host_alloc(int64_t bytes, void** out) {
    SS_REQUIRE(out && bytes >= 0, "ss_host_alloc: bad argument");
    *out = nullptr;
    if (bytes == 0) return SS_OK;
    SS_CHECK_CUDA(cudaHostAlloc(out, size_t(bytes), cudaHostAllocDefault));
    return SS_OK;
}

int32_t ss_host_free(void* p) {
    if (p) SS_CHECK_CUDA(cudaFreeHost(p));
    return SS_OK;
}

// ---- containers ------------------------------------------------------------------------------

static int32_t mat_create(ss_ctx* ctx, int64_t rows, int64_t cols, ss_mat** out, bool ipc) {
    SS_ENTER(ctx);
    SS_REQUIRE(out && rows >= 0 && cols >= 0, "ss_mat_create: bad shape %lld x %lld", (long long)rows,
               (long long)cols);
    ss_mat* m = new ss_mat();
    m->ctx = ctx;
    m->rows = rows;
    m->cols = cols;
    // columns start on 128-byte boundaries (TMA, 128-bit loads); a skinny matrix (a vector uploaded as 1 x n for the
    // ungrouped @L metrics or k(vector)) keeps 16-byte column alignment only: padding 1 row to 16 is a 16 x blow-up
    m->ld = rows >= 16 ? round_up(rows, 16) : round_up(rows > 0 ? rows : 1, 2);
    m->owned = true;
    const size_t bytes = size_t(m->ld) * size_t(cols > 0 ? cols : 1) * 8 + 256;
    // stream-ordered pool allocation: no device-wide synchronisation per matrix (CV loops create
    // and drop many small blocks); ss_mat_create_ipc() is the cudaMalloc variant for IPC sharing
    cudaError_t e = ipc ? cudaMalloc(&m->d, bytes)
                        : cudaMallocAsync(reinterpret_cast<void**>(&m->d), bytes, ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        delete m;
        set_error("ss_mat_create: allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return SS_ERR_OOM;
    }
    m->pooled = !ipc;
    SS_CHECK_CUDA(cudaMemsetAsync(m->d, 0, bytes, ctx->stream));
    if (ipc) SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = m;
    return SS_OK;
}

int32_t ss_mat_create(ss_ctx* ctx, int64_t rows, int64_t cols, ss_mat** out) {
    return mat_create(ctx, rows, cols, out, false);
}

int32_t ss_mat_create_ipc(ss_ctx* ctx, int64_t rows, int64_t cols, ss_mat** out) {
    return mat_create(ctx, rows, cols, out, true);
}

int32_t ss_mat_wrap(ss_ctx* ctx, void* devptr, int64_t rows, int64_t cols, int64_t ld, ss_mat** out) {
    SS_ENTER(ctx);
    SS_REQUIRE(out && devptr && rows >= 0 && cols >= 0, "ss_mat_wrap: bad argument");
    SS_REQUIRE(ld >= rows, "ss_mat_wrap: ld (%lld) must be >= rows (%lld)", (long long)ld, (long long)rows);
    SS_REQUIRE((reinterpret_cast<uintptr_t>(devptr) & 7) == 0, "ss_mat_wrap: pointer must be 8-byte aligned");
    ss_mat* m = new ss_mat();
    m->ctx = ctx;
    m->d = static_cast<double*>(devptr);
    m->rows = rows;
    m->cols = cols;
    m->ld = ld;
    m->owned = false;
    *out = m;
    return SS_OK;
}

int32_t ss_mat_destroy(ss_mat* m) {
    if (!m) return SS_OK;
    if (m->owned && m->d) {
        cudaSetDevice(m->ctx->device);
        if (m->pooled) cudaFreeAsync(m->d, m->ctx->stream);
        else cudaFree(m->d);
    }
    delete m;
    return SS_OK;
}

int32_t ss_mat_info(const ss_mat* m, int64_t* rows, int64_t* cols, int64_t* ld, void** devptr) {
    SS_REQUIRE(m, "ss_mat_info: null matrix");
    if (rows) *rows = m->rows;
    if (cols) *cols = m->cols;
    if (ld) *ld = m->ld;
    if (devptr) *devptr = m->d;
    return SS_OK;
}

static int32_t copy2d(ss_ctx* ctx, cudaStream_t st, void* dst, int64_t dpitch, const void* src, int64_t spitch,
                      int64_t rows, int64_t cols, cudaMemcpyKind kind) {
    if (rows == 0 || cols == 0) return SS_OK;
    SS_CHECK_CUDA(cudaMemcpy2DAsync(dst, size_t(dpitch) * 8, src, size_t(spitch) * 8, size_t(rows) * 8,
                                    size_t(cols), kind, st));
    return SS_OK;
}

// ---- large copies from / to PAGEABLE host memory (what a NumPy / Julia array is) --------------------------------
// cudaMemcpy on pageable memory stages through the driver's bounce buffer on one host thread (~10 GB/s).  Here the
// columns go through three pinned staging buffers owned by the context: several host threads memcpy a chunk while the
// DMA engine moves the previous one, which approaches the PCIe rate.  Pinned / registered memory takes the direct path.
static bool host_is_pageable(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// Persistent host workers for the staging memcpys: spawning 16 threads per 128 MB chunk cost ~50 ms per 13 GB upload.
class CopyPool {
public:
    explicit CopyPool(int n) : n_(n) {
        for (int t = 0; t < n_; ++t) th_.emplace_back([this, t] { run(t); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            ++epoch_;
        }
        cv_.notify_all();
        for (auto& x : th_) x.join();
    }
    int size() const { return n_; }
    // fn(worker index) on every worker; returns when all are done
    void run_all(const std::function<void(int)>& fn) {
        std::unique_lock<std::mutex> g(m_);
        fn_ = &fn;
        pending_ = n_;
        ++epoch_;
        cv_.notify_all();
        done_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void run(int t) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* fn;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(t);
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* fn_ = nullptr;
    int pending_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

static CopyPool& copy_pool() {
    static CopyPool pool(int(std::max(1u, std::min(16u, std::thread::hardware_concurrency()))));
    return pool;
}
static std::mutex g_copy_pool_mutex;  // one staged copy at a time drives the pool (contexts may live on several host threads)

static void parallel_copy_cols(char* dst, size_t dpitch, const char* src, size_t spitch, size_t row_bytes, int64_t cols,
                               int nthreads) {
    auto work = [=](int64_t c0, int64_t c1) {
        if (dpitch == row_bytes && spitch == row_bytes) {
            memcpy(dst + size_t(c0) * dpitch, src + size_t(c0) * spitch, size_t(c1 - c0) * row_bytes);
        } else {
            for (int64_t c = c0; c < c1; ++c) memcpy(dst + size_t(c) * dpitch, src + size_t(c) * spitch, row_bytes);
        }
    };
    if (nthreads <= 1 || cols < 2 * nthreads) {
        work(0, cols);
        return;
    }
    CopyPool& pool = copy_pool();
    const int n = pool.size();
    const int64_t per = (cols + n - 1) / n;
    pool.run_all([&](int t) {
        const int64_t c0 = std::min<int64_t>(cols, int64_t(t) * per), c1 = std::min<int64_t>(cols, c0 + per);
        if (c0 < c1) work(c0, c1);
    });
}

// `ready` / `ready_cols` (download only): columns [i * ready_cols, (i + 1) * ready_cols) of the device matrix are complete
// once event ready[i] has fired -- the copy stream waits for it instead of draining the compute stream first, so the
// download of finished column blocks overlaps the kernels that produce the next ones.  `cchunk_cols` overrides the
// number of columns per staging chunk (it must divide ready_cols).
static constexpr size_t kStage = size_t(128) << 20;
static int32_t ensure_stage(ss_ctx* ctx) {
    if (!ctx->stage[0]) {
        for (int i = 0; i < 3; ++i) {
            SS_CHECK_CUDA(cudaHostAlloc(&ctx->stage[i], kStage, cudaHostAllocDefault));
            SS_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
        }
    }
    return SS_OK;
}

// Medium-sized pageable copies (256 KB .. 32 MB: the matrices of a C2-sized cross-validation): one pinned staging
// buffer, packed / unpacked by the calling thread.  cudaMemcpy on pageable memory took 0.77 ms for a 2.4 MB download;
// DMA into pinned memory + one memcpy is ~3 x faster.
static int32_t small_staged_copy2d(ss_ctx* ctx, double* dev, int64_t ldd, double* host, int64_t ldh, int64_t rows, int64_t cols,
                                   bool upload) {
    SS_TRY(ensure_stage(ctx));
    std::lock_guard<std::mutex> pool_guard(g_copy_pool_mutex);
    const size_t row_bytes = size_t(rows) * 8;
    char* st = static_cast<char*>(ctx->stage[0]);
    SS_CHECK_CUDA(cudaEventSynchronize(ctx->stage_ev[0]));  // a DMA that last used this buffer is done
    if (upload) {
        if (ldh == rows) {
            memcpy(st, host, row_bytes * size_t(cols));
        } else {
            for (int64_t c = 0; c < cols; ++c) memcpy(st + size_t(c) * row_bytes, host + c * ldh, row_bytes);
        }
        SS_CHECK_CUDA(cudaMemcpy2DAsync(dev, size_t(ldd) * 8, st, row_bytes, row_bytes, size_t(cols), cudaMemcpyHostToDevice,
                                        ctx->stream));
        SS_CHECK_CUDA(cudaEventRecord(ctx->stage_ev[0], ctx->stream));
        SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        SS_CHECK_CUDA(cudaMemcpy2DAsync(st, row_bytes, dev, size_t(ldd) * 8, row_bytes, size_t(cols), cudaMemcpyDeviceToHost,
                                        ctx->stream));
        SS_CHECK_CUDA(cudaEventRecord(ctx->stage_ev[0], ctx->stream));
        SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ldh == rows) {
            memcpy(host, st, row_bytes * size_t(cols));
        } else {
            for (int64_t c = 0; c < cols; ++c) memcpy(host + c * ldh, st + size_t(c) * row_bytes, row_bytes);
        }
    }
    return SS_OK;
}

static int32_t staged_copy2d(ss_ctx* ctx, double* dev, int64_t ldd, double* host, int64_t ldh, int64_t rows, int64_t cols,
                             bool upload, const cudaEvent_t* ready = nullptr, int64_t ready_cols = 0,
                             int64_t cchunk_cols = 0) {
    constexpr int kBufs = 3;
    const size_t row_bytes = size_t(rows) * 8;
    if (row_bytes > kStage || rows == 0 || cols == 0) {  // a single column does not fit a staging buffer: direct copy
        SS_TRY(copy2d(ctx, ctx->stream, upload ? (void*)dev : (void*)host, upload ? ldd : ldh, upload ? (const void*)host : (const void*)dev,
                      upload ? ldh : ldd, rows, cols, upload ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost));
        SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        return SS_OK;
    }
    SS_TRY(ensure_stage(ctx));
    std::lock_guard<std::mutex> pool_guard(g_copy_pool_mutex);
    const int nthreads = copy_pool().size();
    const int64_t cchunk = cchunk_cols > 0 ? cchunk_cols : std::max<int64_t>(1, int64_t(kStage / row_bytes));
    cudaStream_t st = upload ? ctx->copy_in : ctx->copy_out;
    if (!ready) SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));  // order against work already queued on the compute stream
    if (upload) {
        int b = 0;
        for (int64_t c0 = 0; c0 < cols; c0 += cchunk, b = (b + 1) % kBufs) {
            const int64_t nc = std::min(cchunk, cols - c0);
            SS_CHECK_CUDA(cudaEventSynchronize(ctx->stage_ev[b]));  // the DMA that last read this buffer is done
            parallel_copy_cols(static_cast<char*>(ctx->stage[b]), row_bytes, reinterpret_cast<const char*>(host + c0 * ldh),
                               size_t(ldh) * 8, row_bytes, nc, nthreads);
            SS_CHECK_CUDA(cudaMemcpy2DAsync(dev + c0 * ldd, size_t(ldd) * 8, ctx->stage[b], row_bytes, row_bytes, size_t(nc),
                                            cudaMemcpyHostToDevice, st));
            SS_CHECK_CUDA(cudaEventRecord(ctx->stage_ev[b], st));
        }
        SS_CHECK_CUDA(cudaStreamSynchronize(st));
    } else {
        // DMA of chunk i + 1 and i + 2 in flight while chunk i is copied out of its staging buffer
        const int64_t nchunks = ceil_div(cols, cchunk);
        auto issue = [&](int64_t i) -> cudaError_t {
            const int b = int(i % kBufs);
            const int64_t c0 = i * cchunk, nc = std::min(cchunk, cols - c0);
            cudaError_t e = ready ? cudaStreamWaitEvent(st, ready[(c0 + nc - 1) / ready_cols], 0) : cudaSuccess;
            if (e == cudaSuccess)
                e = cudaMemcpy2DAsync(ctx->stage[b], row_bytes, dev + c0 * ldd, size_t(ldd) * 8, row_bytes, size_t(nc),
                                      cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaEventRecord(ctx->stage_ev[b], st);
            return e;
        };
        for (int64_t i = 0; i < std::min<int64_t>(kBufs - 1, nchunks); ++i) SS_CHECK_CUDA(issue(i));
        for (int64_t i = 0; i < nchunks; ++i) {
            const int b = int(i % kBufs);
            const int64_t c0 = i * cchunk, nc = std::min(cchunk, cols - c0);
            SS_CHECK_CUDA(cudaEventSynchronize(ctx->stage_ev[b]));
            if (i + kBufs - 1 < nchunks) SS_CHECK_CUDA(issue(i + kBufs - 1));  // its buffer was emptied in the previous round
            parallel_copy_cols(reinterpret_cast<char*>(host + c0 * ldh), size_t(ldh) * 8, static_cast<const char*>(ctx->stage[b]),
                               row_bytes, row_bytes, nc, nthreads);
        }
        SS_CHECK_CUDA(cudaStreamSynchronize(st));
    }
    return SS_OK;
}

static constexpr int64_t kStagedCopyMinBytes = int64_t(32) << 20;
static constexpr int64_t kSmallStageMinBytes = int64_t(256) << 10;

int32_t ss_mat_upload(ss_ctx* ctx, ss_mat* m, const double* host, int64_t ld_host) {
    SS_ENTER(ctx);
    SS_REQUIRE(m && (host || m->rows * m->cols == 0), "ss_mat_upload: null argument");
    SS_REQUIRE(ld_host >= m->rows, "ss_mat_upload: ld_host (%lld) < rows (%lld)", (long long)ld_host,
               (long long)m->rows);
    if (m->rows * m->cols * 8 >= kStagedCopyMinBytes && host_is_pageable(host))
        return staged_copy2d(ctx, m->d, m->ld, const_cast<double*>(host), ld_host, m->rows, m->cols, true);
    if (m->rows * m->cols * 8 >= kSmallStageMinBytes && host_is_pageable(host))
        return small_staged_copy2d(ctx, m->d, m->ld, const_cast<double*>(host), ld_host, m->rows, m->cols, true);
    SS_TRY(copy2d(ctx, ctx->stream, m->d, m->ld, host, ld_host, m->rows, m->cols, cudaMemcpyHostToDevice));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

// Row-major host array (rows x cols, row pitch ld_host >= cols; NumPy's default order) into a column-major device
// matrix: the array is the column-major image of its transpose, so it is uploaded as it lies (staged copy for pageable
// memory) and transposed on the device -- instead of a strided transposing copy on the host (0.3 s for 800 MB in NumPy).
int32_t ss_mat_upload_rowmajor(ss_ctx* ctx, ss_mat* m, const double* host, int64_t ld_host) {
    SS_ENTER(ctx);
    SS_REQUIRE(m && (host || m->rows * m->cols == 0), "ss_mat_upload_rowmajor: null argument");
    SS_REQUIRE(ld_host >= m->cols, "ss_mat_upload_rowmajor: ld_host (%lld) < cols (%lld)", (long long)ld_host,
               (long long)m->cols);
    if (m->rows == 0 || m->cols == 0) return SS_OK;
    const int64_t ldt = round_up(m->cols, 16);
    double* tmp = nullptr;
    if (cudaMallocAsync(reinterpret_cast<void**>(&tmp), size_t(ldt) * size_t(m->rows) * 8, ctx->stream) != cudaSuccess) {
        cudaGetLastError();
        set_error("ss_mat_upload_rowmajor: out of device memory for the %lld x %lld staging copy", (long long)m->cols,
                  (long long)m->rows);
        return SS_ERR_OOM;
    }
    int32_t st;
    if (m->rows * m->cols * 8 >= kStagedCopyMinBytes && host_is_pageable(host)) {
        st = staged_copy2d(ctx, tmp, ldt, const_cast<double*>(host), ld_host, m->cols, m->rows, true);
    } else {
        st = copy2d(ctx, ctx->stream, tmp, ldt, host, ld_host, m->cols, m->rows, cudaMemcpyHostToDevice);
    }
    if (st == SS_OK) st = launch_transpose(ctx, tmp, ldt, m->d, m->ld, m->rows, m->cols);
    cudaFreeAsync(tmp, ctx->stream);
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return st;
}

int32_t ss_mat_download(ss_ctx* ctx, const ss_mat* m, double* host, int64_t ld_host) {
    SS_ENTER(ctx);
    SS_REQUIRE(m && (host || m->rows * m->cols == 0), "ss_mat_download: null argument");
    SS_REQUIRE(ld_host >= m->rows, "ss_mat_download: ld_host (%lld) < rows (%lld)", (long long)ld_host,
               (long long)m->rows);
    if (m->rows * m->cols * 8 >= kStagedCopyMinBytes && host_is_pageable(host))
        return staged_copy2d(ctx, m->d, m->ld, host, ld_host, m->rows, m->cols, false);
    if (m->rows * m->cols * 8 >= kSmallStageMinBytes && host_is_pageable(host))
        return small_staged_copy2d(ctx, m->d, m->ld, host, ld_host, m->rows, m->cols, false);
    SS_TRY(copy2d(ctx, ctx->stream, host, ld_host, m->d, m->ld, m->rows, m->cols, cudaMemcpyDeviceToHost));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_mat_upload_cols_async(ss_ctx* ctx, ss_mat* m, int64_t col0, int64_t ncols, const double* host,
                                 int64_t ld_host) {
    SS_ENTER(ctx);
    SS_REQUIRE(m && host && col0 >= 0 && ncols >= 0 && col0 + ncols <= m->cols && ld_host >= m->rows,
               "ss_mat_upload_cols_async: bad argument");
    return copy2d(ctx, ctx->stream, m->d + col0 * m->ld, m->ld, host, ld_host, m->rows, ncols,
                  cudaMemcpyHostToDevice);
}

int32_t ss_mat_download_cols_async(ss_ctx* ctx, const ss_mat* m, int64_t col0, int64_t ncols, double* host,
                                   int64_t ld_host) {
    SS_ENTER(ctx);
    SS_REQUIRE(m && host && col0 >= 0 && ncols >= 0 && col0 + ncols <= m->cols && ld_host >= m->rows,
               "ss_mat_download_cols_async: bad argument");
    return copy2d(ctx, ctx->stream, host, ld_host, m->d + col0 * m->ld, m->ld, m->rows, ncols,
                  cudaMemcpyDeviceToHost);
}

int32_t ss_ivec_create(ss_ctx* ctx, int64_t n, ss_ivec** out) {
    SS_ENTER(ctx);
    SS_REQUIRE(out && n >= 0, "ss_ivec_create: bad argument");
    ss_ivec* v = new ss_ivec();
    v->ctx = ctx;
    v->n = n;
    v->owned = true;
    const size_t bytes = size_t(n > 0 ? n : 1) * 4 + 64;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&v->d), bytes, ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        delete v;
        set_error("ss_ivec_create: allocation of %zu bytes failed", bytes);
        return SS_ERR_OOM;
    }
    SS_CHECK_CUDA(cudaMemsetAsync(v->d, 0, bytes, ctx->stream));
    *out = v;
    return SS_OK;
}

int32_t ss_ivec_wrap(ss_ctx* ctx, void* devptr, int64_t n, ss_ivec** out) {
    SS_ENTER(ctx);
    SS_REQUIRE(out && devptr && n >= 0, "ss_ivec_wrap: bad argument");
    ss_ivec* v = new ss_ivec();
    v->ctx = ctx;
    v->d = static_cast<int32_t*>(devptr);
    v->n = n;
    v->owned = false;
    *out = v;
    return SS_OK;
}

int32_t ss_ivec_destroy(ss_ivec* v) {
    if (!v) return SS_OK;
    if (v->owned && v->d) {
        cudaSetDevice(v->ctx->device);
        cudaFreeAsync(v->d, v->ctx->stream);
    }
    delete v;
    return SS_OK;
}

int32_t ss_ivec_info(const ss_ivec* v, int64_t* n, void** devptr) {
    SS_REQUIRE(v, "ss_ivec_info: null vector");
    if (n) *n = v->n;
    if (devptr) *devptr = v->d;
    return SS_OK;
}

int32_t ss_ivec_upload(ss_ctx* ctx, ss_ivec* v, const int32_t* host) {
    SS_ENTER(ctx);
    SS_REQUIRE(v && (host || v->n == 0), "ss_ivec_upload: null argument");
    if (v->n) SS_CHECK_CUDA(cudaMemcpyAsync(v->d, host, size_t(v->n) * 4, cudaMemcpyHostToDevice, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_ivec_download(ss_ctx* ctx, const ss_ivec* v, int32_t* host) {
    SS_ENTER(ctx);
    SS_REQUIRE(v && (host || v->n == 0), "ss_ivec_download: null argument");
    if (v->n) SS_CHECK_CUDA(cudaMemcpyAsync(host, v->d, size_t(v->n) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

// ---- (1) featurize -----------------------------------------------------------------------------

int32_t ss_featurize(ss_ctx* ctx, const ss_mat* S, double alpha, int32_t weighted, ss_mat* X) {
    SS_ENTER(ctx);
    SS_REQUIRE(S && X, "ss_featurize: null matrix");
    SS_REQUIRE(S->rows == X->rows && S->cols == X->cols, "ss_featurize: shape mismatch");
    SS_TRY(launch_featurize(ctx, S->d, S->rows, S->cols, S->ld, alpha, weighted != 0, X->d, X->ld));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_featurize_csr(ss_ctx* ctx, const ss_mat* S, double alpha, int32_t weighted, ss_csr** out) {
    SS_ENTER(ctx);
    SS_REQUIRE(S && out, "ss_featurize_csr: null argument");
    SS_TRY(featurize_csr(ctx, S, alpha, weighted != 0, out));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_featurize_csc(ss_ctx* ctx, const ss_mat* S, double alpha, int32_t weighted, ss_csr** out) {
    SS_ENTER(ctx);
    SS_REQUIRE(S && out, "ss_featurize_csc: null argument");
    SS_TRY(featurize_csc(ctx, S, alpha, weighted != 0, out));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_jaccard_featurize(ss_ctx* ctx, const ss_mat* DA, const ss_mat* DB, double alpha, int32_t weighted, ss_mat* X) {
    SS_ENTER(ctx);
    SS_REQUIRE(DA && DB && X, "ss_jaccard_featurize: null matrix");
    SS_REQUIRE(DA->cols == DB->cols, "ss_jaccard_featurize: descriptor counts differ (%lld vs %lld)", (long long)DA->cols,
               (long long)DB->cols);
    SS_REQUIRE(X->rows == DA->rows && X->cols == DB->rows, "ss_jaccard_featurize: X must be rows(DA) x rows(DB)");
    SS_TRY(jaccard_featurize(ctx, DA, DB, alpha, weighted != 0, X));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_tanimoto_featurize_bits(ss_ctx* ctx, const void* fa_dev, int64_t na, const void* fb_dev, int64_t nb, int64_t words,
                                   double alpha, int32_t weighted, ss_mat* X) {
    SS_ENTER(ctx);
    SS_REQUIRE(X && na >= 0 && nb >= 0 && words >= 1, "ss_tanimoto_featurize_bits: bad argument");
    SS_REQUIRE((fa_dev || na == 0) && (fb_dev || nb == 0), "ss_tanimoto_featurize_bits: null fingerprint array");
    SS_REQUIRE(((reinterpret_cast<uintptr_t>(fa_dev) | reinterpret_cast<uintptr_t>(fb_dev)) & 7) == 0,
               "ss_tanimoto_featurize_bits: fingerprints must be 8-byte aligned");
    SS_REQUIRE(X->rows == na && X->cols == nb, "ss_tanimoto_featurize_bits: X must be na x nb");
    SS_TRY(tanimoto_bits_featurize(ctx, static_cast<const uint64_t*>(fa_dev), na, static_cast<const uint64_t*>(fb_dev), nb, words,
                                   alpha, weighted != 0, X));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_csr_info(const ss_csr* c, int64_t* rows, int64_t* cols, int64_t* nnz, int32_t* has_values) {
    SS_REQUIRE(c, "ss_csr_info: null csr");
    if (rows) *rows = c->rows;
    if (cols) *cols = c->cols;
    if (nnz) *nnz = c->nnz;
    if (has_values) *has_values = c->values ? 1 : 0;
    return SS_OK;
}

int32_t ss_csr_download(ss_ctx* ctx, const ss_csr* c, int32_t* row_ptr, int32_t* col_idx, double* values) {
    SS_ENTER(ctx);
    SS_REQUIRE(c, "ss_csr_download: null csr");
    if (row_ptr)
        SS_CHECK_CUDA(cudaMemcpyAsync(row_ptr, c->row_ptr, size_t(c->rows + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (col_idx && c->nnz)
        SS_CHECK_CUDA(cudaMemcpyAsync(col_idx, c->col_idx, size_t(c->nnz) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (values && c->values && c->nnz)
        SS_CHECK_CUDA(cudaMemcpyAsync(values, c->values, size_t(c->nnz) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_csr_destroy(ss_csr* c) {
    if (!c) return SS_OK;
    cudaSetDevice(c->ctx->device);
    if (c->owned) {  // stream-ordered pool allocations (cudaMallocAsync): no device-wide synchronisation per CSR
        if (c->row_ptr) cudaFreeAsync(c->row_ptr, c->ctx->stream);
        if (c->col_idx) cudaFreeAsync(c->col_idx, c->ctx->stream);
        if (c->values) cudaFreeAsync(c->values, c->ctx->stream);
    }
    delete c;
    return SS_OK;
}

int32_t ss_csr_wrap(ss_ctx* ctx, int64_t rows, int64_t cols, int64_t nnz, void* row_ptr_dev, void* col_idx_dev,
                    void* values_dev, ss_csr** out) {
    SS_ENTER(ctx);
    SS_REQUIRE(out && rows >= 0 && cols >= 0 && nnz >= 0 && row_ptr_dev && (nnz == 0 || col_idx_dev),
               "ss_csr_wrap: bad argument");
    SS_REQUIRE(nnz < (1ll << 31) && cols < (1ll << 31), "ss_csr_wrap: int32 CSR limits exceeded");
    ss_csr* c = new ss_csr();
    c->ctx = ctx;
    c->rows = rows;
    c->cols = cols;
    c->nnz = nnz;
    c->row_ptr = static_cast<int32_t*>(row_ptr_dev);
    c->col_idx = static_cast<int32_t*>(col_idx_dev);
    c->values = static_cast<double*>(values_dev);
    c->owned = false;
    *out = c;
    return SS_OK;
}

int32_t ss_recommend_topl(ss_ctx* ctx, const ss_csr* Y, const ss_csr* YT, int32_t L, int64_t s_begin, int64_t s_end,
                          ss_ivec* idx_out, ss_mat* val_out) {
    SS_ENTER(ctx);
    SS_REQUIRE(Y && YT && idx_out, "ss_recommend_topl: null argument");
    SS_REQUIRE(YT->rows == Y->cols && YT->cols == Y->rows && YT->nnz == Y->nnz,
               "ss_recommend_topl: YT must be the CSR of the transpose of Y");
    SS_REQUIRE((Y->values == nullptr) == (YT->values == nullptr), "ss_recommend_topl: Y and YT must both be binary or both weighted");
    SS_REQUIRE(L >= 1 && L <= 32 && L <= Y->cols, "ss_recommend_topl: L must be in 1..min(32, targets)");
    SS_REQUIRE(s_begin >= 0 && s_end <= Y->rows && s_begin <= s_end, "ss_recommend_topl: bad source range");
    SS_REQUIRE(idx_out->n == int64_t(L) * Y->rows, "ss_recommend_topl: idx_out must hold L x sources entries");
    SS_REQUIRE(!val_out || (val_out->rows == L && val_out->cols == Y->rows), "ss_recommend_topl: val_out must be an L x sources matrix");
    SS_TRY(recommend_topl(ctx, Y, YT, L, s_begin, s_end, idx_out->d, val_out ? val_out->d : nullptr, val_out ? val_out->ld : 0));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

// ---- (2) construct / degrees -------------------------------------------------------------------

int32_t ss_gather(ss_ctx* ctx, const ss_mat* src, const ss_ivec* row_idx, const ss_ivec* col_idx, ss_mat* dst) {
    SS_ENTER(ctx);
    SS_REQUIRE(src && dst, "ss_gather: null matrix");
    SS_REQUIRE((row_idx ? row_idx->n : src->rows) == dst->rows && (col_idx ? col_idx->n : src->cols) == dst->cols,
               "ss_gather: destination shape does not match the index lists");
    SS_TRY(launch_gather(ctx, src->d, src->ld, row_idx ? row_idx->d : nullptr, col_idx ? col_idx->d : nullptr,
                         dst->d, dst->rows, dst->cols, dst->ld));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

static int32_t degrees_async(ss_ctx* ctx, const ss_mat* Xs, const ss_mat* Y, int32_t* ks, int32_t* kf, int32_t* kt) {
    const int64_t ns = Y->rows;
    if (ks) SS_CHECK_CUDA(cudaMemsetAsync(ks, 0, size_t(ns) * 4, ctx->stream));
    if (kf && Xs) SS_CHECK_CUDA(cudaMemsetAsync(kf, 0, size_t(Xs->cols) * 4, ctx->stream));
    if (kt) SS_CHECK_CUDA(cudaMemsetAsync(kt, 0, size_t(Y->cols) * 4, ctx->stream));
    if (Xs && (ks || kf)) SS_TRY(launch_degrees(ctx, Xs->d, Xs->rows, Xs->cols, Xs->ld, ks, kf));
    if (ks || kt) SS_TRY(launch_degrees(ctx, Y->d, Y->rows, Y->cols, Y->ld, ks, kt));
    return SS_OK;
}

int32_t ss_degrees(ss_ctx* ctx, const ss_mat* Xs, const ss_mat* Y, ss_ivec* ks, ss_ivec* kf, ss_ivec* kt) {
    SS_ENTER(ctx);
    SS_REQUIRE(Y, "ss_degrees: Y is required");
    SS_REQUIRE(!Xs || Xs->rows == Y->rows, "ss_degrees: Xs and Y must have the same number of source rows");
    SS_REQUIRE(!ks || ks->n == Y->rows, "ss_degrees: ks has wrong length");
    SS_REQUIRE(!kf || (Xs && kf->n == Xs->cols), "ss_degrees: kf has wrong length");
    SS_REQUIRE(!kt || kt->n == Y->cols, "ss_degrees: kt has wrong length");
    SS_TRY(degrees_async(ctx, Xs, Y, ks ? ks->d : nullptr, kf ? kf->d : nullptr, kt ? kt->d : nullptr));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_k_rows(ss_ctx* ctx, const ss_mat* G, ss_ivec* k) {
    SS_ENTER(ctx);
    SS_REQUIRE(G && k && k->n == G->rows, "ss_k_rows: bad argument");
    if (k->n) SS_CHECK_CUDA(cudaMemsetAsync(k->d, 0, size_t(k->n) * 4, ctx->stream));
    SS_TRY(launch_degrees(ctx, G->d, G->rows, G->cols, G->ld, k->d, nullptr));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

// ---- (3) spread / predict ----------------------------------------------------------------------

int32_t ss_spread_rows(ss_ctx* ctx, const ss_mat* G, const ss_ivec* k, ss_mat* W) {
    SS_ENTER(ctx);
    SS_REQUIRE(G && W && G->rows == W->rows && G->cols == W->cols, "ss_spread_rows: shape mismatch");
    const int32_t* kd = nullptr;
    if (k) {
        SS_REQUIRE(k->n == G->rows, "ss_spread_rows: k has wrong length");
        kd = k->d;
    } else {
        void* p;
        SS_TRY(scratch_get(ctx, 0, size_t(G->rows + 1) * 4, &p));
        SS_CHECK_CUDA(cudaMemsetAsync(p, 0, size_t(G->rows + 1) * 4, ctx->stream));
        SS_TRY(launch_degrees(ctx, G->d, G->rows, G->cols, G->ld, static_cast<int32_t*>(p), nullptr));
        kd = static_cast<int32_t*>(p);
    }
    SS_TRY(launch_spread_rows(ctx, G->d, G->rows, G->cols, G->ld, kd, W->d, W->ld));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_gemm_f64(ss_ctx* ctx, int32_t opA, const ss_mat* A, const ss_mat* B, ss_mat* C, const ss_ivec* row_div,
                    const ss_ivec* col_flag) {
    SS_ENTER(ctx);
    SS_REQUIRE(A && B && C, "ss_gemm_f64: null matrix");
    SS_REQUIRE(opA == SS_OP_N || opA == SS_OP_T, "ss_gemm_f64: bad opA");
    const int64_t M = (opA == SS_OP_N) ? A->rows : A->cols;
    const int64_t K = (opA == SS_OP_N) ? A->cols : A->rows;
    SS_REQUIRE(B->rows == K && C->rows == M && C->cols == B->cols, "ss_gemm_f64: shape mismatch");
    SS_REQUIRE(!row_div || row_div->n == M, "ss_gemm_f64: row_div has wrong length");
    SS_REQUIRE(!col_flag || col_flag->n == B->cols, "ss_gemm_f64: col_flag has wrong length");
    SS_TRY(launch_gemm_f64(ctx, opA, A->d, A->ld, B->d, B->ld, C->d, C->ld, M, B->cols, K,
                           row_div ? row_div->d : nullptr, col_flag ? col_flag->d : nullptr, false));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

// dispatch of one chain product on the requested precision
static int32_t chain_gemm(ss_ctx* ctx, uint32_t precision, int opA, const double* A, int64_t lda, const double* B,
                          int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t N, int64_t K, const int32_t* row_div,
                          const int32_t* col_flag) {
    if (precision == SS_PRECISION_F64)
        return launch_gemm_f64(ctx, opA, A, lda, B, ldb, C, ldc, M, N, K, row_div, col_flag, false);
    if (precision == SS_PRECISION_TF32)
        return launch_gemm_tf32(ctx, opA, A, lda, B, ldb, C, ldc, M, N, K, row_div, col_flag, false);
    if (precision == SS_PRECISION_F64_INT8) {
        int S = 6;
        if (const char* env = getenv("SS_INT8_SLICES")) {
            const int v = atoi(env);
            if (v >= 2 && v <= 8) S = v;
        }
        // every entry is certified a posteriori (error bound <= tol * entry, default 4e-13 so that the two
        // products of a chain stay below 1e-12); a product with uncertified entries is re-run on the FP64 DMMA
        // path.  SS_INT8_CERTIFY=0 skips the check, SS_INT8_TOL overrides the tolerance.
        double tol = 4e-13;
        if (const char* env = getenv("SS_INT8_TOL")) {
            const double v = atof(env);
            if (v > 0.0) tol = v;
        }
        const char* ce = getenv("SS_INT8_CERTIFY");
        const bool certify = !(ce && ce[0] == '0');
        int64_t bad = 0;
        const int32_t st8 = launch_gemm_i8(ctx, opA, A, lda, B, ldb, C, ldc, M, N, K, row_div, col_flag, S, certify ? tol : 0.0,
                                           certify ? &bad : nullptr);
        if (st8 == SS_ERR_INVALID) {
            // a shape the sliced mode cannot take (K too long for exact INT32 accumulation): the documented behaviour of
            // this mode is the FP64 DMMA path, counted as a re-run.  (Negative / non-finite operands stay an error:
            // SS_ERR_UNSUPPORTED, the caller asked for a mode that is not defined for them.)
            ctx->int8_stats[1] += 1;
            return launch_gemm_f64(ctx, opA, A, lda, B, ldb, C, ldc, M, N, K, row_div, col_flag, false);
        }
        SS_TRY(st8);
        ctx->int8_stats[0] += 1;
        ctx->int8_stats[2] = bad;
        if (bad > 0) {
            ctx->int8_stats[1] += 1;
            return launch_gemm_f64(ctx, opA, A, lda, B, ldb, C, ldc, M, N, K, row_div, col_flag, false);
        }
        return SS_OK;
    }
    set_error("unknown precision flag 0x%x", precision);
    return SS_ERR_INVALID;
}

int32_t ss_gemm_lowp(ss_ctx* ctx, int32_t opA, const ss_mat* A, const ss_mat* B, ss_mat* C, const ss_ivec* row_div,
                     const ss_ivec* col_flag, uint32_t precision) {
    SS_ENTER(ctx);
    SS_REQUIRE(A && B && C, "ss_gemm_lowp: null matrix");
    SS_REQUIRE(opA == SS_OP_N || opA == SS_OP_T, "ss_gemm_lowp: bad opA");
    const int64_t M = (opA == SS_OP_N) ? A->rows : A->cols;
    const int64_t K = (opA == SS_OP_N) ? A->cols : A->rows;
    SS_REQUIRE(B->rows == K && C->rows == M && C->cols == B->cols, "ss_gemm_lowp: shape mismatch");
    SS_REQUIRE(!row_div || row_div->n == M, "ss_gemm_lowp: row_div has wrong length");
    SS_REQUIRE(!col_flag || col_flag->n == B->cols, "ss_gemm_lowp: col_flag has wrong length");
    SS_TRY(chain_gemm(ctx, precision & SS_PRECISION_MASK, opA, A->d, A->ld, B->d, B->ld, C->d, C->ld, M, B->cols, K,
                      row_div ? row_div->d : nullptr, col_flag ? col_flag->d : nullptr));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_gemm_f64_mirrored(ss_ctx* ctx, int32_t opA, const ss_mat* A, const ss_mat* B, ss_mat* C,
                             const ss_ivec* row_div, const ss_ivec* col_flag, int32_t n_mirrors, void* const* mirrors) {
    SS_ENTER(ctx);
    SS_REQUIRE(A && B && C, "ss_gemm_f64_mirrored: null matrix");
    SS_REQUIRE(opA == SS_OP_N || opA == SS_OP_T, "ss_gemm_f64_mirrored: bad opA");
    SS_REQUIRE(n_mirrors >= 0 && n_mirrors <= 7 && (n_mirrors == 0 || mirrors), "ss_gemm_f64_mirrored: 0..7 mirrors");
    const int64_t M = (opA == SS_OP_N) ? A->rows : A->cols;
    const int64_t K = (opA == SS_OP_N) ? A->cols : A->rows;
    SS_REQUIRE(B->rows == K && C->rows == M && C->cols == B->cols, "ss_gemm_f64_mirrored: shape mismatch");
    SS_REQUIRE(!row_div || row_div->n == M, "ss_gemm_f64_mirrored: row_div has wrong length");
    SS_REQUIRE(!col_flag || col_flag->n == B->cols, "ss_gemm_f64_mirrored: col_flag has wrong length");
    double* mp[7] = {nullptr};
    for (int i = 0; i < n_mirrors; ++i) {
        SS_REQUIRE(mirrors[i] && (reinterpret_cast<uintptr_t>(mirrors[i]) & 7) == 0, "ss_gemm_f64_mirrored: bad mirror pointer");
        mp[i] = static_cast<double*>(mirrors[i]);
    }
    SS_TRY(launch_gemm_f64(ctx, opA, A->d, A->ld, B->d, B->ld, C->d, C->ld, M, B->cols, K,
                           row_div ? row_div->d : nullptr, col_flag ? col_flag->d : nullptr, false, n_mirrors, mp));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_mat_ipc_handle(ss_ctx* ctx, const ss_mat* m, void* handle64_out) {
    SS_ENTER(ctx);
    SS_REQUIRE(m && handle64_out, "ss_mat_ipc_handle: null argument");
    SS_REQUIRE(m->owned && !m->pooled, "ss_mat_ipc_handle: only matrices from ss_mat_create_ipc can be shared");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    SS_CHECK_CUDA(cudaIpcGetMemHandle(&h, m->d));
    memcpy(handle64_out, &h, 64);
    return SS_OK;
}

int32_t ss_ipc_open(ss_ctx* ctx, const void* handle64, void** devptr_out) {
    SS_ENTER(ctx);
    SS_REQUIRE(handle64 && devptr_out, "ss_ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SS_CHECK_CUDA(cudaIpcOpenMemHandle(devptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return SS_OK;
}

int32_t ss_ipc_close(ss_ctx* ctx, void* devptr) {
    SS_ENTER(ctx);
    if (devptr) SS_CHECK_CUDA(cudaIpcCloseMemHandle(devptr));
    return SS_OK;
}

// Shared front half of both predict forms: degrees -> Wst = Y ./ ks -> T = (Xs' * Wst) ./ kf.
// Returns device pointers into the context workspaces.
struct ChainWs {
    int32_t *ks = nullptr, *kf = nullptr, *kt = nullptr;
    double* Wst = nullptr;
    int64_t ldw = 0;
    double* T = nullptr;
    int64_t ldt = 0;
};

// need_wst: the caller also reads Wst (predict_source); otherwise a sparse label matrix takes the edge-list form of the
// first product (csrc/ss_tsparse.cu) and Wst is never materialised.
static int32_t chain_front(ss_ctx* ctx, const ss_mat* Xs, const ss_mat* Y, ChainWs* w,
                           uint32_t precision = SS_PRECISION_F64, bool need_wst = false) {
    const int64_t ns = Y->rows, nt = Y->cols, nf = Xs ? Xs->cols : 0;
    void* p;
    const size_t kbytes = size_t(round_up(ns, 64) + round_up(nf, 64) + round_up(nt, 64)) * 4;
    SS_TRY(scratch_get(ctx, 0, kbytes, &p));
    w->ks = static_cast<int32_t*>(p);
    w->kf = w->ks + round_up(ns, 64);
    w->kt = w->kf + round_up(nf, 64);
    SS_TRY(degrees_async(ctx, Xs, Y, w->ks, Xs ? w->kf : nullptr, w->kt));
    if (Xs && precision == SS_PRECISION_F64_INT8) {
        // INT8-sliced mode: fold the 1/ks of Wst = Y ./ ks into the rows of Xs, T = ((Xs ./ ks)' * Y) ./ kf
        // (x/k * y instead of x * (y/k): one rounding apart).  Y itself is then an operand, and a 0/1 label matrix
        // is a single 8-bit plane: 6 slice pairs instead of 21.
        const int64_t ldx = round_up(ns, 16);
        SS_TRY(scratch_get(ctx, 1, size_t(ldx) * size_t(nf) * 8, &p));
        double* Xk = static_cast<double*>(p);
        SS_TRY(launch_spread_rows(ctx, Xs->d, ns, nf, Xs->ld, w->ks, Xk, ldx));
        w->ldt = round_up(nf, 16);
        SS_TRY(scratch_get(ctx, 2, size_t(w->ldt) * size_t(nt) * 8, &p));
        w->T = static_cast<double*>(p);
        SS_TRY(chain_gemm(ctx, precision, SS_OP_T, Xk, ldx, Y->d, Y->ld, w->T, w->ldt, nf, nt, ns, w->kf, nullptr));
        return SS_OK;
    }
    if (Xs && !need_wst && precision == SS_PRECISION_F64) {
        w->ldt = round_up(nf, 16);
        SS_TRY(scratch_get(ctx, 2, size_t(w->ldt) * size_t(nt) * 8, &p));
        w->T = static_cast<double*>(p);
        bool used = false;
        SS_TRY(t_from_sparse_labels(ctx, Xs->d, Xs->ld, Y->d, Y->ld, ns, nf, nt, w->ks, w->kf, w->kt, w->T, w->ldt, 0, nullptr,
                                    &used));
        if (used) return SS_OK;
    }
    w->ldw = round_up(ns, 16);
    SS_TRY(scratch_get(ctx, 1, size_t(w->ldw) * size_t(nt) * 8, &p));
    w->Wst = static_cast<double*>(p);
    SS_TRY(launch_spread_rows(ctx, Y->d, ns, nt, Y->ld, w->ks, w->Wst, w->ldw));
    if (Xs) {
        w->ldt = round_up(nf, 16);
        SS_TRY(scratch_get(ctx, 2, size_t(w->ldt) * size_t(nt) * 8, &p));
        w->T = static_cast<double*>(p);
        // T[f,t] = (sum_s Xs[s,f] * Wst[s,t]) / kf[f]
        SS_TRY(chain_gemm(ctx, precision, SS_OP_T, Xs->d, Xs->ld, w->Wst, w->ldw, w->T, w->ldt, nf, nt, ns, w->kf,
                          nullptr));
    }
    return SS_OK;
}

int32_t ss_predict_query(ss_ctx* ctx, const ss_mat* Xq, const ss_mat* Xs, const ss_mat* Y, ss_mat* R, uint32_t flags,
                         ss_ivec* kt_out) {
    SS_ENTER(ctx);
    SS_REQUIRE(Xq && Xs && Y && R, "ss_predict_query: null matrix");
    SS_REQUIRE(Xq->cols == Xs->cols, "Number of features between test and training sets doesn't match");
    SS_REQUIRE(Xs->rows == Y->rows, "Labels and features have different number of source nodes");
    SS_REQUIRE(R->rows == Xq->rows && R->cols == Y->cols, "ss_predict_query: R must be Nq x Nt");
    SS_REQUIRE(!kt_out || kt_out->n == Y->cols, "ss_predict_query: kt_out has wrong length");
    if (R->rows == 0 || R->cols == 0) return SS_OK;
    if (Xs->rows == 0 || Xs->cols == 0) {  // no sources / no features: every product is empty -> zeros
        SS_CHECK_CUDA(cudaMemset2DAsync(R->d, size_t(R->ld) * 8, 0, size_t(R->rows) * 8, size_t(R->cols), ctx->stream));
        SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        return SS_OK;
    }
    ChainWs w;
    const uint32_t prec = flags & SS_PRECISION_MASK;
    SS_TRY(chain_front(ctx, Xs, Y, &w, prec));
    SS_TRY(chain_gemm(ctx, prec, SS_OP_N, Xq->d, Xq->ld, w.T, w.ldt, R->d, R->ld, Xq->rows, Y->cols, Xq->cols,
                      nullptr, (flags & SS_PREDICT_CLEAN) ? w.kt : nullptr));
    if (kt_out)
        SS_CHECK_CUDA(cudaMemcpyAsync(kt_out->d, w.kt, size_t(Y->cols) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

// predict + clean! with the result delivered to HOST memory: the second product runs in column blocks and every
// finished block is copied out (through the staging buffers when `host` is pageable) while the next one is computed.
// Bit-identical to ss_predict_query + ss_mat_download (the K loop of an entry does not depend on its tile); R keeps the
// device copy.  This is the call behind host.predict() / SimSpreadB200.predict when the whole query block is asked for.
int32_t ss_predict_query_fetch(ss_ctx* ctx, const ss_mat* Xq, const ss_mat* Xs, const ss_mat* Y, ss_mat* R, uint32_t flags,
                               double* host, int64_t ld_host) {
    SS_ENTER(ctx);
    SS_REQUIRE(Xq && Xs && Y && R && host, "ss_predict_query_fetch: null argument");
    SS_REQUIRE(Xq->cols == Xs->cols, "Number of features between test and training sets doesn't match");
    SS_REQUIRE(Xs->rows == Y->rows, "Labels and features have different number of source nodes");
    SS_REQUIRE(R->rows == Xq->rows && R->cols == Y->cols, "ss_predict_query_fetch: R must be Nq x Nt");
    SS_REQUIRE(ld_host >= R->rows, "ss_predict_query_fetch: leading dimension too small");
    const int64_t nq = R->rows, nt = R->cols;
    if (nq == 0 || nt == 0) return SS_OK;
    const uint32_t prec = flags & SS_PRECISION_MASK;
    const size_t row_bytes = size_t(nq) * 8;
    const bool degenerate = Xs->rows == 0 || Xs->cols == 0;
    // columns per staging chunk (a multiple of the GEMM's column tile) and per GEMM launch (>= ~20 waves of tiles)
    const int64_t cchunk = int64_t(kStage / row_bytes) & ~int64_t(127);
    if (degenerate || prec == SS_PRECISION_F64_INT8 || cchunk < 128 || nt < 4 * cchunk ||
        row_bytes * size_t(nt) < size_t(kStagedCopyMinBytes)) {
        SS_TRY(ss_predict_query(ctx, Xq, Xs, Y, R, flags, nullptr));
        return ss_mat_download(ctx, R, host, ld_host);
    }
    const int64_t tiles_chunk = ceil_div(nq, 128) * (cchunk / 128);
    const int64_t per_block = std::max<int64_t>(1, ceil_div(int64_t(20) * ctx->sm_count, tiles_chunk));
    const int64_t bcols = cchunk * per_block;
    const int64_t nblocks = ceil_div(nt, bcols);
    ChainWs w;
    SS_TRY(chain_front(ctx, Xs, Y, &w, prec));
    std::vector<cudaEvent_t> ready(size_t(nblocks), nullptr);
    int32_t status = SS_OK;
    auto run = [&]() -> int32_t {
        for (int64_t b = 0; b < nblocks; ++b) SS_CHECK_CUDA(cudaEventCreateWithFlags(&ready[size_t(b)], cudaEventDisableTiming));
        for (int64_t b = 0; b < nblocks; ++b) {
            const int64_t c0 = b * bcols, nc = std::min(bcols, nt - c0);
            SS_TRY(chain_gemm(ctx, prec, SS_OP_N, Xq->d, Xq->ld, w.T + c0 * w.ldt, w.ldt, R->d + c0 * R->ld, R->ld, nq, nc,
                              Xq->cols, nullptr, (flags & SS_PREDICT_CLEAN) ? w.kt + c0 : nullptr));
            SS_CHECK_CUDA(cudaEventRecord(ready[size_t(b)], ctx->stream));
        }
        if (host_is_pageable(host)) {
            SS_TRY(staged_copy2d(ctx, R->d, R->ld, host, ld_host, nq, nt, false, ready.data(), bcols, cchunk));
        } else {
            for (int64_t b = 0; b < nblocks; ++b) {
                const int64_t c0 = b * bcols, nc = std::min(bcols, nt - c0);
                SS_CHECK_CUDA(cudaStreamWaitEvent(ctx->copy_out, ready[size_t(b)], 0));
                SS_TRY(copy2d(ctx, ctx->copy_out, host + c0 * ld_host, ld_host, R->d + c0 * R->ld, R->ld, nq, nc,
                              cudaMemcpyDeviceToHost));
            }
            SS_CHECK_CUDA(cudaStreamSynchronize(ctx->copy_out));
        }
        SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        return SS_OK;
    };
    status = run();
    if (status != SS_OK) {  // drain before the events go away
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->copy_out);
    }
    for (cudaEvent_t e : ready)
        if (e) cudaEventDestroy(e);
    return status;
}

// k-fold cross-validation in one call (the loop of docs/src/api.md:17-21): per fold the blocks
// X[queries, features], X[sources, features], y[sources, targets] are extracted by index list (construct,
// src/core.jl:167,171-172) and predict + clean! runs on them; everything is queued on the context stream with the
// workspaces re-used fold after fold and ONE synchronisation at the end (a fold of an Enzyme-sized data set is tens of
// microseconds of kernels: the per-fold host round trips were the cost).
int32_t ss_predict_query_folds(ss_ctx* ctx, const ss_mat* X, const ss_mat* Y, int32_t nfolds, const int32_t* q_ptr,
                               const int32_t* q_idx, const int32_t* s_ptr, const int32_t* s_idx, const int32_t* ys_idx,
                               const int32_t* f_ptr, const int32_t* f_idx, ss_mat* R, uint32_t flags) {
    SS_ENTER(ctx);
    SS_REQUIRE(X && Y && R && nfolds >= 0, "ss_predict_query_folds: bad argument");
    if (nfolds == 0) return SS_OK;
    SS_REQUIRE(q_ptr && q_idx && s_ptr && s_idx && ys_idx && f_ptr && f_idx, "ss_predict_query_folds: null index list");
    const int64_t nt = Y->cols;
    SS_REQUIRE(R->cols == nt && R->rows == q_ptr[nfolds] - q_ptr[0], "ss_predict_query_folds: R must be (all queries) x targets");
    const uint32_t prec = flags & SS_PRECISION_MASK;
    int64_t mq = 0, ms = 0, mf = 0;
    for (int f = 0; f < nfolds; ++f) {
        SS_REQUIRE(q_ptr[f + 1] >= q_ptr[f] && s_ptr[f + 1] >= s_ptr[f] && f_ptr[f + 1] >= f_ptr[f],
                   "ss_predict_query_folds: index pointers must be non-decreasing");
        mq = std::max<int64_t>(mq, q_ptr[f + 1] - q_ptr[f]);
        ms = std::max<int64_t>(ms, s_ptr[f + 1] - s_ptr[f]);
        mf = std::max<int64_t>(mf, f_ptr[f + 1] - f_ptr[f]);
    }
    for (int64_t i = q_ptr[0]; i < q_ptr[nfolds]; ++i) SS_REQUIRE(q_idx[i] >= 0 && q_idx[i] < X->rows, "ss_predict_query_folds: query index out of range");
    for (int64_t i = s_ptr[0]; i < s_ptr[nfolds]; ++i)
        SS_REQUIRE(s_idx[i] >= 0 && s_idx[i] < X->rows && ys_idx[i] >= 0 && ys_idx[i] < Y->rows, "ss_predict_query_folds: source index out of range");
    for (int64_t i = f_ptr[0]; i < f_ptr[nfolds]; ++i) SS_REQUIRE(f_idx[i] >= 0 && f_idx[i] < X->cols, "ss_predict_query_folds: feature index out of range");
    // index lists -> device, one copy
    const int64_t nqi = q_ptr[nfolds] - q_ptr[0], nsi = s_ptr[nfolds] - s_ptr[0], nfi = f_ptr[nfolds] - f_ptr[0];
    void* p;
    SS_TRY(scratch_get(ctx, 16, size_t(nqi + 2 * nsi + nfi + 4) * 4, &p));
    int32_t* dq = static_cast<int32_t*>(p);
    int32_t* dsx = dq + nqi;
    int32_t* dsy = dsx + nsi;
    int32_t* df = dsy + nsi;
    SS_CHECK_CUDA(cudaMemcpyAsync(dq, q_idx + q_ptr[0], size_t(nqi) * 4, cudaMemcpyHostToDevice, ctx->stream));
    SS_CHECK_CUDA(cudaMemcpyAsync(dsx, s_idx + s_ptr[0], size_t(nsi) * 4, cudaMemcpyHostToDevice, ctx->stream));
    SS_CHECK_CUDA(cudaMemcpyAsync(dsy, ys_idx + s_ptr[0], size_t(nsi) * 4, cudaMemcpyHostToDevice, ctx->stream));
    SS_CHECK_CUDA(cudaMemcpyAsync(df, f_idx + f_ptr[0], size_t(nfi) * 4, cudaMemcpyHostToDevice, ctx->stream));
    // Small folds (a cross-validation at Enzyme size): the fold is a grid dimension of three launches and the blocks
    // are read through the index lists (ss_folds.cu).  Large folds keep the per-fold chain on the persistent DMMA GEMM.
    {
        bool small = prec == SS_PRECISION_F64 && std::max(ms, mf) <= 2048 && nt >= 1 && !getenv("SS_FOLDS_SERIAL");
        for (int f = 0; small && f < nfolds; ++f)
            small = q_ptr[f + 1] > q_ptr[f] && s_ptr[f + 1] > s_ptr[f] && f_ptr[f + 1] > f_ptr[f];
        if (small) {
            SS_TRY(predict_query_folds_batched(ctx, X, Y, nfolds, q_ptr, s_ptr, f_ptr, dq, dsx, dsy, df, R,
                                               (flags & SS_PREDICT_CLEAN) != 0));
            SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
            return SS_OK;
        }
    }
    // block buffers sized for the largest fold
    ss_mat bXq, bXs, bY, bR;
    bXq.ctx = bXs.ctx = bY.ctx = bR.ctx = ctx;
    SS_TRY(scratch_get(ctx, 17, size_t(round_up(std::max<int64_t>(mq, 1), 16)) * size_t(std::max<int64_t>(mf, 1)) * 8, &p));
    bXq.d = static_cast<double*>(p);
    SS_TRY(scratch_get(ctx, 18, size_t(round_up(std::max<int64_t>(ms, 1), 16)) * size_t(std::max<int64_t>(mf, 1)) * 8, &p));
    bXs.d = static_cast<double*>(p);
    SS_TRY(scratch_get(ctx, 19, size_t(round_up(std::max<int64_t>(ms, 1), 16)) * size_t(std::max<int64_t>(nt, 1)) * 8, &p));
    bY.d = static_cast<double*>(p);
    for (int f = 0; f < nfolds; ++f) {
        const int64_t nq = q_ptr[f + 1] - q_ptr[f], ns = s_ptr[f + 1] - s_ptr[f], nf = f_ptr[f + 1] - f_ptr[f];
        if (nq == 0 || nt == 0) continue;
        const int64_t off = q_ptr[f] - q_ptr[0];
        bR.d = R->d + off;  // rows [off, off + nq) of R
        bR.rows = nq; bR.cols = nt; bR.ld = R->ld;
        if (ns == 0 || nf == 0) {  // no sources / no features: every product is empty -> zeros
            SS_CHECK_CUDA(cudaMemset2DAsync(bR.d, size_t(bR.ld) * 8, 0, size_t(nq) * 8, size_t(nt), ctx->stream));
            continue;
        }
        bXq.rows = nq; bXq.cols = nf; bXq.ld = round_up(nq, 16);
        bXs.rows = ns; bXs.cols = nf; bXs.ld = round_up(ns, 16);
        bY.rows = ns; bY.cols = nt; bY.ld = round_up(ns, 16);
        const int32_t* fq = dq + off;
        const int32_t* fs = dsx + (s_ptr[f] - s_ptr[0]);
        const int32_t* fy = dsy + (s_ptr[f] - s_ptr[0]);
        const int32_t* ff = df + (f_ptr[f] - f_ptr[0]);
        SS_TRY(launch_gather(ctx, X->d, X->ld, fq, ff, bXq.d, nq, nf, bXq.ld));
        SS_TRY(launch_gather(ctx, X->d, X->ld, fs, ff, bXs.d, ns, nf, bXs.ld));
        SS_TRY(launch_gather(ctx, Y->d, Y->ld, fy, nullptr, bY.d, ns, nt, bY.ld));
        ChainWs w;
        SS_TRY(chain_front(ctx, &bXs, &bY, &w, prec));
        SS_TRY(chain_gemm(ctx, prec, SS_OP_N, bXq.d, bXq.ld, w.T, w.ldt, bR.d, bR.ld, nq, nt, nf, nullptr,
                          (flags & SS_PREDICT_CLEAN) ? w.kt : nullptr));
    }
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_predict_query_csr(ss_ctx* ctx, const ss_csr* Xq, const ss_csr* XsT, const ss_mat* Y, ss_mat* R,
                             uint32_t flags, ss_ivec* kt_out) {
    SS_ENTER(ctx);
    SS_REQUIRE(Xq && XsT && Y && R, "ss_predict_query_csr: null argument");
    SS_REQUIRE(Xq->cols == XsT->rows, "Number of features between test and training sets doesn't match");
    SS_REQUIRE(XsT->cols == Y->rows, "Labels and features have different number of source nodes");
    SS_REQUIRE(R->rows == Xq->rows && R->cols == Y->cols, "ss_predict_query_csr: R must be Nq x Nt");
    SS_REQUIRE(!kt_out || kt_out->n == Y->cols, "ss_predict_query_csr: kt_out has wrong length");
    SS_TRY(predict_query_csr(ctx, Xq, XsT, Y, R, flags, kt_out ? kt_out->d : nullptr));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_predict_source(ss_ctx* ctx, const ss_mat* Xs, const ss_mat* Y, ss_mat* R, uint32_t flags) {
    SS_ENTER(ctx);
    SS_REQUIRE(Y && R, "ss_predict_source: null matrix");
    SS_REQUIRE(!Xs || Xs->rows == Y->rows, "Labels and features have different number of source nodes");
    SS_REQUIRE(R->rows == Y->rows && R->cols == Y->cols, "ss_predict_source: R must be Ns x Nt");
    if (R->rows == 0 || R->cols == 0) return SS_OK;
    const ss_mat* Xeff = (Xs && Xs->cols > 0) ? Xs : nullptr;
    ChainWs w;
    SS_TRY(chain_front(ctx, Xeff, Y, &w, SS_PRECISION_F64, true));
    const int64_t ns = Y->rows, nt = Y->cols;
    // U[t',t] = (sum_s Y[s,t'] * Wst[s,t]) / kt[t']
    void* p;
    const int64_t ldu = round_up(nt, 16);
    SS_TRY(scratch_get(ctx, 3, size_t(ldu) * size_t(nt) * 8, &p));
    double* U = static_cast<double*>(p);
    SS_TRY(launch_gemm_f64(ctx, SS_OP_T, Y->d, Y->ld, w.Wst, w.ldw, U, ldu, nt, nt, ns, w.kt, nullptr, false));
    const int32_t* flag = (flags & SS_PREDICT_CLEAN) ? w.kt : nullptr;
    if (Xeff) {
        SS_TRY(launch_gemm_f64(ctx, SS_OP_N, Xeff->d, Xeff->ld, w.T, w.ldt, R->d, R->ld, ns, nt, Xeff->cols, nullptr,
                               nullptr, false));
        SS_TRY(launch_gemm_f64(ctx, SS_OP_N, Y->d, Y->ld, U, ldu, R->d, R->ld, ns, nt, nt, nullptr, flag, true));
    } else {
        SS_TRY(launch_gemm_f64(ctx, SS_OP_N, Y->d, Y->ld, U, ldu, R->d, R->ld, ns, nt, nt, nullptr, flag, false));
    }
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_clean(ss_ctx* ctx, ss_mat* R, const ss_ivec* kt) {
    SS_ENTER(ctx);
    SS_REQUIRE(R && kt && kt->n == R->cols, "ss_clean: bad argument");
    SS_TRY(launch_clean(ctx, R->d, R->rows, R->cols, R->ld, kt->d));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

// R_host (nq x nt) = Xq_host (nq x nf) * T (device, nf x nt) [+ clean! flag], streamed: query-row slabs
// of Xq go up on copy_in while the GEMM of the previous slab runs on `stream` and finished R slabs
// come down on copy_out (double-buffered; overlap needs pinned host memory).
static int32_t stream_product(ss_ctx* ctx, const double* Xq, int64_t ldxq, int64_t nq, int64_t nf, const double* T,
                              int64_t ldt, int64_t nt, const int32_t* col_flag, double* R, int64_t ldr) {
    void* p;
    // slab size: about 1 GiB of Xq+R per buffer, multiple of 128 rows
    int64_t slab = (int64_t(1) << 30) / ((nf + nt) * 8);
    slab = slab / 128 * 128;
    if (slab < 128) slab = 128;
    if (const char* env = getenv("SS_SLAB_ROWS")) {  // test hook: force many small slabs
        const long long v = atoll(env);
        if (v >= 16) slab = round_up(v, 16);
    }
    if (slab > nq) slab = round_up(nq, 16);
    const int64_t nslab = ceil_div(nq, slab);
    double* dXq[2];
    double* dR[2];
    SS_TRY(scratch_get(ctx, 6, size_t(slab) * size_t(nf) * 8 * 2, &p));
    dXq[0] = static_cast<double*>(p);
    dXq[1] = dXq[0] + slab * nf;
    SS_TRY(scratch_get(ctx, 7, size_t(slab) * size_t(nt) * 8 * 2, &p));
    dR[0] = static_cast<double*>(p);
    dR[1] = dR[0] + slab * nt;
    cudaEvent_t in_done[2], mm_done[2], out_done[2];
    for (int i = 0; i < 2; ++i) {
        SS_CHECK_CUDA(cudaEventCreateWithFlags(&in_done[i], cudaEventDisableTiming));
        SS_CHECK_CUDA(cudaEventCreateWithFlags(&mm_done[i], cudaEventDisableTiming));
        SS_CHECK_CUDA(cudaEventCreateWithFlags(&out_done[i], cudaEventDisableTiming));
    }
    int32_t status = SS_OK;
    for (int64_t i = 0; i < nslab && status == SS_OK; ++i) {
        const int b = int(i & 1);
        const int64_t r0 = i * slab;
        const int64_t nr = (nq - r0 < slab) ? nq - r0 : slab;
        // upload slab i (needs the GEMM that last read this buffer to be done)
        if (i >= 2) cudaStreamWaitEvent(ctx->copy_in, mm_done[b], 0);
        status = copy2d(ctx, ctx->copy_in, dXq[b], slab, Xq + r0, ldxq, nr, nf, cudaMemcpyHostToDevice);
        if (status != SS_OK) break;
        cudaEventRecord(in_done[b], ctx->copy_in);
        // GEMM slab i (needs its input and the download that last read this R buffer)
        cudaStreamWaitEvent(ctx->stream, in_done[b], 0);
        if (i >= 2) cudaStreamWaitEvent(ctx->stream, out_done[b], 0);
        status = launch_gemm_f64(ctx, SS_OP_N, dXq[b], slab, T, ldt, dR[b], slab, nr, nt, nf, nullptr,
                                 col_flag, false);
        if (status != SS_OK) break;
        cudaEventRecord(mm_done[b], ctx->stream);
        // download slab i
        cudaStreamWaitEvent(ctx->copy_out, mm_done[b], 0);
        status = copy2d(ctx, ctx->copy_out, R + r0, ldr, dR[b], slab, nr, nt, cudaMemcpyDeviceToHost);
        cudaEventRecord(out_done[b], ctx->copy_out);
    }
    cudaError_t e1 = cudaStreamSynchronize(ctx->copy_in);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaError_t e3 = cudaStreamSynchronize(ctx->copy_out);
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(in_done[i]);
        cudaEventDestroy(mm_done[i]);
        cudaEventDestroy(out_done[i]);
    }
    if (status != SS_OK) return status;
    SS_CHECK_CUDA(e1);
    SS_CHECK_CUDA(e2);
    SS_CHECK_CUDA(e3);
    return SS_OK;
}

// Host-buffer form: Xs / Y are uploaded once, then query-row slabs of Xq stream in on copy_in while
// the R GEMM of the previous slab runs on `stream` and finished R slabs stream out on copy_out.
int32_t ss_predict_query_host(ss_ctx* ctx, const double* Xq, int64_t ldxq, const double* Xs, int64_t ldxs,
                              const double* Y, int64_t ldy, int64_t nq, int64_t ns, int64_t nf, int64_t nt,
                              uint32_t flags, double* R, int64_t ldr) {
    SS_ENTER(ctx);
    SS_REQUIRE(nq >= 0 && ns >= 0 && nf >= 0 && nt >= 0, "ss_predict_query_host: negative dimension");
    SS_REQUIRE(ldxq >= nq && ldxs >= ns && ldy >= ns && ldr >= nq, "ss_predict_query_host: leading dimension too small");
    if (nq == 0 || nt == 0) return SS_OK;
    SS_REQUIRE(Xq && Xs && Y && R, "ss_predict_query_host: null buffer");
    if (ns == 0 || nf == 0) {
        for (int64_t c = 0; c < nt; ++c) memset(R + c * ldr, 0, size_t(nq) * 8);
        return SS_OK;
    }
    // resident operands
    ss_mat mXs, mY;
    void* p;
    mXs.ctx = mY.ctx = ctx;
    mXs.rows = ns; mXs.cols = nf; mXs.ld = round_up(ns, 16);
    mY.rows = ns; mY.cols = nt; mY.ld = round_up(ns, 16);
    SS_TRY(scratch_get(ctx, 4, size_t(mXs.ld) * size_t(nf) * 8, &p));
    mXs.d = static_cast<double*>(p);
    SS_TRY(scratch_get(ctx, 5, size_t(mY.ld) * size_t(nt) * 8, &p));
    mY.d = static_cast<double*>(p);
    SS_TRY(copy2d(ctx, ctx->stream, mXs.d, mXs.ld, Xs, ldxs, ns, nf, cudaMemcpyHostToDevice));
    SS_TRY(copy2d(ctx, ctx->stream, mY.d, mY.ld, Y, ldy, ns, nt, cudaMemcpyHostToDevice));
    ChainWs w;
    SS_TRY(chain_front(ctx, &mXs, &mY, &w));

    return stream_product(ctx, Xq, ldxq, nq, nf, w.T, w.ldt, nt, (flags & SS_PREDICT_CLEAN) ? w.kt : nullptr, R, ldr);
}

int32_t ss_stream_product_host(ss_ctx* ctx, const double* Xq, int64_t ldxq, int64_t nq, const ss_mat* T,
                               const ss_ivec* col_flag, double* R, int64_t ldr) {
    SS_ENTER(ctx);
    SS_REQUIRE(T && nq >= 0 && ldxq >= nq && ldr >= nq, "ss_stream_product_host: bad argument");
    SS_REQUIRE(!col_flag || col_flag->n == T->cols, "ss_stream_product_host: col_flag has wrong length");
    if (nq == 0 || T->cols == 0) return SS_OK;
    SS_REQUIRE(Xq && R, "ss_stream_product_host: null buffer");
    if (T->rows == 0) {
        for (int64_t c = 0; c < T->cols; ++c) memset(R + c * ldr, 0, size_t(nq) * 8);
        return SS_OK;
    }
    return stream_product(ctx, Xq, ldxq, nq, T->rows, T->d, T->ld, T->cols, col_flag ? col_flag->d : nullptr, R, ldr);
}

// ---- (4) ranking / metrics ---------------------------------------------------------------------

int32_t ss_topl_rows(ss_ctx* ctx, const ss_mat* R, int32_t L, ss_ivec* idx_out, ss_mat* val_out) {
    SS_ENTER(ctx);
    SS_REQUIRE(R && idx_out, "ss_topl_rows: null argument");
    SS_REQUIRE(L > 0 && L <= R->cols, "ss_topl_rows: L must be in 1..cols");
    SS_REQUIRE(idx_out->n == int64_t(L) * R->rows, "ss_topl_rows: idx_out must hold L x rows entries");
    SS_REQUIRE(!val_out || (val_out->rows == L && val_out->cols == R->rows), "ss_topl_rows: val_out must be L x rows");
    SS_TRY(launch_topl(ctx, R->d, R->rows, R->cols, R->ld, L, idx_out->d, val_out ? val_out->d : nullptr,
                       val_out ? val_out->ld : 0));
    SS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return SS_OK;
}

int32_t ss_atl(ss_ctx* ctx, const ss_mat* Ytrue, const ss_mat* R, int32_t L, double* out2) {
    SS_ENTER(ctx);
    SS_REQUIRE(Ytrue && R && out2, "ss_atl: null argument");
    SS_REQUIRE(Ytrue->rows == R->rows && Ytrue->cols == R->cols, "Number of predictions must match number of labels");
    if (L <= 0) {
        set_error("Please use a list length greater than 0 (L > 0)");
        return SS_ERR_ASSERT;
    }
    if (R->cols <= L) {
        set_error("Number of labels is less than length (L > y)");
        return SS_ERR_ASSERT;
    }
    return atl(ctx, Ytrue, R, L, out2);
}

int32_t ss_auroc_auprc(ss_ctx* ctx, const void* labels_u8_dev, const void* scores_f64_dev, int64_t M, double* out2) {
    SS_ENTER(ctx);
    SS_REQUIRE(out2 && M >= 0 && (M == 0 || (labels_u8_dev && scores_f64_dev)), "ss_auroc_auprc: bad argument");
    return auroc_auprc(ctx, static_cast<const uint8_t*>(labels_u8_dev), static_cast<const double*>(scores_f64_dev), M,
                       out2);
}

int32_t ss_auc_sort(ss_ctx* ctx, const void* labels_u8_dev, const void* scores_f64_dev, const void* keys_u64_dev, int64_t M,
                    void** keys_sorted_out, void** labels_sorted_out) {
    SS_ENTER(ctx);
    SS_REQUIRE(keys_sorted_out && labels_sorted_out && M >= 0, "ss_auc_sort: bad argument");
    SS_REQUIRE(M == 0 || (labels_u8_dev && (scores_f64_dev || keys_u64_dev)), "ss_auc_sort: null input");
    uint64_t* k = nullptr;
    uint8_t* l = nullptr;
    SS_TRY(auc_sort(ctx, static_cast<const uint8_t*>(labels_u8_dev), static_cast<const double*>(scores_f64_dev),
                    static_cast<const uint64_t*>(keys_u64_dev), M, &k, &l));
    *keys_sorted_out = k;
    *labels_sorted_out = l;
    return SS_OK;
}

int32_t ss_auc_lower_bound(ss_ctx* ctx, const void* keys_sorted_dev, int64_t M, const uint64_t* query, int32_t nq,
                           int64_t* pos_out) {
    SS_ENTER(ctx);
    SS_REQUIRE(M >= 0 && nq >= 0 && (nq == 0 || (query && pos_out)) && (M == 0 || keys_sorted_dev), "ss_auc_lower_bound: bad argument");
    return auc_lower_bound(ctx, static_cast<const uint64_t*>(keys_sorted_dev), M, query, nq, pos_out);
}

int32_t ss_auc_segment_summary(ss_ctx* ctx, const void* keys_sorted_dev, const void* labels_sorted_dev, int64_t M,
                               int64_t* summary3) {
    SS_ENTER(ctx);
    SS_REQUIRE(summary3 && M >= 0 && (M == 0 || (keys_sorted_dev && labels_sorted_dev)), "ss_auc_segment_summary: bad argument");
    return auc_segment_summary(ctx, static_cast<const uint64_t*>(keys_sorted_dev), static_cast<const uint8_t*>(labels_sorted_dev), M,
                               summary3);
}

int32_t ss_auc_segment_integrate(ss_ctx* ctx, const void* keys_sorted_dev, const void* labels_sorted_dev, int64_t M,
                                 const int64_t* global6, double* out2) {
    SS_ENTER(ctx);
    SS_REQUIRE(global6 && out2 && M >= 0 && (M == 0 || (keys_sorted_dev && labels_sorted_dev)), "ss_auc_segment_integrate: bad argument");
    return auc_segment_integrate(ctx, static_cast<const uint64_t*>(keys_sorted_dev), static_cast<const uint8_t*>(labels_sorted_dev), M,
                                 global6, out2);
}

int32_t ss_auroc_auprc_mat(ss_ctx* ctx, const ss_mat* Ytrue, const ss_mat* R, double* out2) {
    SS_ENTER(ctx);
    SS_REQUIRE(Ytrue && R && out2, "ss_auroc_auprc_mat: null argument");
    if (Ytrue->rows != R->rows || Ytrue->cols != R->cols) {
        set_error("The number of scores must be equal to the number of labels");
        return SS_ERR_ASSERT;
    }
    return auroc_auprc_mat(ctx, Ytrue, R, out2);
}

int32_t ss_bedroc(ss_ctx* ctx, const ss_mat* Ytrue, const ss_mat* R, int32_t rev, double alpha, double* out) {
    SS_ENTER(ctx);
    SS_REQUIRE(Ytrue && R && out, "ss_bedroc: null argument");
    if (Ytrue->rows != R->rows || Ytrue->cols != R->cols) {
        set_error("The number of scores must be equal to the number of labels");
        return SS_ERR_ASSERT;
    }
    return bedroc(ctx, Ytrue, R, rev, alpha, out);
}

int32_t ss_threshold_sweep(ss_ctx* ctx, const ss_mat* Ytrue, const ss_mat* R, int32_t metric, double* out4) {
    SS_ENTER(ctx);
    SS_REQUIRE(Ytrue && R && out4, "ss_threshold_sweep: null argument");
    SS_REQUIRE(metric >= 0 && metric <= 5, "ss_threshold_sweep: unknown metric %d", metric);
    if (Ytrue->rows != R->rows || Ytrue->cols != R->cols) {
        set_error("The number of scores must be equal to the number of labels");
        return SS_ERR_ASSERT;
    }
    return threshold_sweep(ctx, Ytrue, R, metric, out4);
}

}  // extern "C"
