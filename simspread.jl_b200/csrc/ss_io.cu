// Host-side fast path of `read_namedmatrix` (reference src/utils.jl:50-53: `readdlm(filepath, delimiter,
// String)` followed by `parse.(Float64, M[r_idx:end, c_idx:end])`, :30-32) for large similarity / label
// matrices.  The value block of a delimited text file is parsed by all host cores
// straight into a column-major Float64 buffer (Julia's `Matrix{Float64}` layout, ready for ss_mat_upload);
// the names (first line / first field of every line) stay with the host layer.
//
// Field rule of the reference call: EVERY occurrence of the delimiter separates two fields (an explicit
// delimiter, even ' ', is not merged by readdlm), lines end with '\n' (a preceding '\r' is dropped), a
// trailing empty line is ignored.  Numbers go through std::from_chars (correctly rounded, like Julia's
// parse(Float64, .)), plus the spellings Julia accepts that from_chars does not: a leading '+' and
// surrounding blanks.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "ss_common.cuh"

namespace {

struct Mapped {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    ~Mapped() {
        if (p && n) munmap(const_cast<char*>(p), n);
        if (fd >= 0) close(fd);
    }
};

int32_t map_file(const char* path, Mapped& m) {
    m.fd = open(path, O_RDONLY);
    SS_REQUIRE(m.fd >= 0, "read_namedmatrix: cannot open %s", path);
    struct stat st;
    SS_REQUIRE(fstat(m.fd, &st) == 0, "read_namedmatrix: cannot stat %s", path);
    m.n = size_t(st.st_size);
    if (m.n == 0) return SS_OK;
    void* p = mmap(nullptr, m.n, PROT_READ, MAP_PRIVATE, m.fd, 0);
    SS_REQUIRE(p != MAP_FAILED, "read_namedmatrix: cannot map %s", path);
    m.p = static_cast<const char*>(p);
    return SS_OK;
}

// line starts of the whole file (a final line without '\n' counts; a trailing empty line does not)
void line_starts(const Mapped& m, std::vector<size_t>& starts) {
    const unsigned nthr = std::max(1u, std::min(std::thread::hardware_concurrency(), unsigned(m.n >> 22) + 1));
    std::vector<std::vector<size_t>> part(nthr);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthr; ++t)
        th.emplace_back([&, t] {
            const size_t b = m.n * t / nthr, e = m.n * (t + 1) / nthr;
            const char* q = m.p + b;
            while (q < m.p + e) {
                const char* nl = static_cast<const char*>(memchr(q, '\n', size_t(m.p + e - q)));
                if (!nl) break;
                part[t].push_back(size_t(nl - m.p) + 1);
                q = nl + 1;
            }
        });
    for (auto& x : th) x.join();
    starts.clear();
    if (m.n) starts.push_back(0);
    for (auto& v : part) starts.insert(starts.end(), v.begin(), v.end());
    if (!starts.empty() && starts.back() >= m.n) starts.pop_back();  // file ends with '\n'
}

inline size_t line_end(const Mapped& m, const std::vector<size_t>& starts, size_t i) {
    size_t e = (i + 1 < starts.size()) ? starts[i + 1] - 1 : m.n;
    if (e > starts[i] && i + 1 >= starts.size() && m.p[e - 1] == '\n') --e;
    if (e > starts[i] && m.p[e - 1] == '\r') --e;
    return e;
}

bool parse_f64(const char* b, const char* e, double& out) {
    while (b < e && (*b == ' ' || *b == '\t')) ++b;
    while (e > b && (e[-1] == ' ' || e[-1] == '\t')) --e;
    if (b < e && *b == '+') ++b;
    if (b >= e) return false;
    auto r = std::from_chars(b, e, out);
    return r.ec == std::errc() && r.ptr == e;
}

}  // namespace

extern "C" {

// number of lines and of delimiter-separated fields of the first line
int32_t ss_text_matrix_dims(const char* path, int32_t delimiter, int64_t* lines_out, int64_t* fields_out) {
    SS_REQUIRE(path && lines_out && fields_out, "ss_text_matrix_dims: null argument");
    Mapped m;
    SS_TRY(map_file(path, m));
    std::vector<size_t> starts;
    line_starts(m, starts);
    *lines_out = int64_t(starts.size());
    int64_t f = 0;
    if (!starts.empty()) {
        const size_t e = line_end(m, starts, 0);
        f = 1;
        for (size_t i = starts[0]; i < e; ++i) f += (m.p[i] == char(delimiter));
    }
    *fields_out = f;
    return SS_OK;
}

// values[i + j*ld] = parse(Float64, field (j + skip_fields) of line (i + skip_lines)); every line must hold
// exactly cols + skip_fields fields
int32_t ss_text_matrix_read(const char* path, int32_t delimiter, int32_t skip_lines, int32_t skip_fields, double* values,
                            int64_t rows, int64_t cols, int64_t ld) {
    SS_REQUIRE(path && (values || rows * cols == 0) && ld >= rows && skip_lines >= 0 && skip_fields >= 0,
               "ss_text_matrix_read: bad argument");
    Mapped m;
    SS_TRY(map_file(path, m));
    std::vector<size_t> starts;
    line_starts(m, starts);
    SS_REQUIRE(int64_t(starts.size()) == rows + skip_lines, "ss_text_matrix_read: %lld lines in %s, expected %lld",
               (long long)starts.size(), path, (long long)(rows + skip_lines));
    const unsigned nthr = unsigned(std::max<int64_t>(1, std::min<int64_t>(std::thread::hardware_concurrency(), rows / 8 + 1)));
    std::vector<int64_t> bad(nthr, -1);
    std::vector<std::thread> th;
    const char d = char(delimiter);
    for (unsigned t = 0; t < nthr; ++t)
        th.emplace_back([&, t] {
            for (int64_t i = rows * t / nthr; i < rows * (t + 1) / nthr; ++i) {
                const size_t li = size_t(i + skip_lines);
                const char* b = m.p + starts[li];
                const char* e = m.p + line_end(m, starts, li);
                int64_t field = 0;
                for (;;) {
                    const char* f = static_cast<const char*>(memchr(b, d, size_t(e - b)));
                    const char* fe = f ? f : e;
                    if (field >= skip_fields) {
                        const int64_t j = field - skip_fields;
                        double v;
                        if (j >= cols || !parse_f64(b, fe, v)) {
                            bad[t] = i;
                            return;
                        }
                        values[i + j * ld] = v;
                    }
                    ++field;
                    if (!f) break;
                    b = f + 1;
                }
                if (field != cols + skip_fields) {
                    bad[t] = i;
                    return;
                }
            }
        });
    for (auto& x : th) x.join();
    for (unsigned t = 0; t < nthr; ++t)
        SS_REQUIRE(bad[t] < 0, "ss_text_matrix_read: line %lld of %s does not hold %lld numeric fields",
                   (long long)(bad[t] + skip_lines + 1), path, (long long)cols);
    return SS_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// `save` (reference src/core.jl:503-522, 542-561): one line `fold, "source", "target", score, label` per
// (query, target) pair, appended to the file; numbers as Julia's `string(x)` prints them (shortest round-trip
// digits, Base.Ryu.writeshortest: fixed notation for decimal exponents -4..5, `d.ddde-7` otherwise; an
// integer-valued matrix prints integers).  All host cores format blocks of query rows; the blocks are written in order.
// ------------------------------------------------------------------------------------------------
namespace {

inline void jl_append_float(std::string& out, double x) {
    if (x != x) {
        out += "NaN";
        return;
    }
    if (x == 0.0) {
        out += std::signbit(x) ? "-0.0" : "0.0";
        return;
    }
    if (std::isinf(x)) {
        out += x > 0 ? "Inf" : "-Inf";
        return;
    }
    char buf[48];
    auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);  // shortest: d[.ddd]e[+-]XX
    const char* p = buf;
    if (*p == '-') {
        out += '-';
        ++p;
    }
    char digits[24];
    int nd = 0;
    for (; p < r.ptr && *p != 'e'; ++p)
        if (*p != '.') digits[nd++] = *p;
    ++p;  // 'e'
    int e10 = 0;
    bool eneg = false;
    if (*p == '-') {
        eneg = true;
        ++p;
    } else if (*p == '+') {
        ++p;
    }
    for (; p < r.ptr; ++p) e10 = e10 * 10 + (*p - '0');
    if (eneg) e10 = -e10;
    while (nd > 1 && digits[nd - 1] == '0') --nd;
    if (e10 > -5 && e10 < 6) {
        if (e10 >= 0) {
            for (int i = 0; i <= e10; ++i) out += i < nd ? digits[i] : '0';
            out += '.';
            if (nd > e10 + 1) out.append(digits + e10 + 1, size_t(nd - e10 - 1));
            else out += '0';
        } else {
            out += "0.";
            out.append(size_t(-e10 - 1), '0');
            out.append(digits, size_t(nd));
        }
    } else {
        out += digits[0];
        out += '.';
        if (nd > 1) out.append(digits + 1, size_t(nd - 1));
        else out += '0';
        out += 'e';
        out += std::to_string(e10);
    }
}

inline void jl_append_value(std::string& out, double x, bool as_int) {
    if (as_int && x == double((long long)x)) out += std::to_string((long long)x);
    else jl_append_float(out, x);
}

}  // namespace

extern "C" {

/* yhat, y: host, column-major nq x nt (rows = the queries of y in its row order).  fold >= 0: that fold id on every
 * line (src/core.jl:542-561); fold < 0: the 1-based index of the query (src/core.jl:512).  *_is_int: the matrix has
 * an integer element type in the caller (prints `1`, not `1.0`). */
int32_t ss_save_rows(const char* path, int32_t append, int64_t fold, int64_t nq, int64_t nt, const char* const* qnames,
                     const char* const* tnames, const double* yhat, int64_t ld_yhat, int32_t yhat_is_int, const double* y,
                     int64_t ld_y, int32_t y_is_int, int32_t delimiter, int64_t* bytes_written) {
    SS_REQUIRE(path && nq >= 0 && nt >= 0 && (nq * nt == 0 || (qnames && tnames && yhat && y)), "ss_save_rows: null argument");
    SS_REQUIRE(ld_yhat >= nq && ld_y >= nq, "ss_save_rows: leading dimension smaller than the number of queries");
    FILE* f = fopen(path, append ? "a+" : "w");
    SS_REQUIRE(f, "ss_save_rows: cannot open %s", path);
    const char d = char(delimiter);
    int64_t total = 0;
    const int nthreads = int(std::max(1u, std::min(16u, std::thread::hardware_concurrency())));
    // target-name fields are the same for every query: quote them once
    std::vector<std::string> tq(static_cast<size_t>(nt));
    for (int64_t t = 0; t < nt; ++t) tq[size_t(t)] = std::string("\"") + tnames[t] + "\"";
    const int64_t rows_per_block = std::max<int64_t>(1, (int64_t(1) << 22) / std::max<int64_t>(nt, 1));  // ~4M lines per round
    // two sets of per-thread buffers: the previous round is written by its own thread while this one is formatted
    std::vector<std::string> sets[2] = {std::vector<std::string>(static_cast<size_t>(nthreads)),
                                        std::vector<std::string>(static_cast<size_t>(nthreads))};
    bool ok = true;
    std::thread writer;
    int cur = 0;
    for (int64_t q0 = 0; q0 < nq; q0 += rows_per_block * nthreads, cur ^= 1) {
        std::vector<std::string>& parts = sets[cur];
        auto work = [&](int t) {
            std::string& out = parts[size_t(t)];
            out.clear();
            const int64_t a = std::min(nq, q0 + int64_t(t) * rows_per_block), b = std::min(nq, a + rows_per_block);
            for (int64_t q = a; q < b; ++q) {
                const std::string head = std::to_string(fold >= 0 ? fold : q + 1) + d + "\"" + qnames[q] + "\"" + d;
                for (int64_t c = 0; c < nt; ++c) {
                    out += head;
                    out += tq[size_t(c)];
                    out += d;
                    jl_append_value(out, yhat[c * ld_yhat + q], yhat_is_int != 0);
                    out += d;
                    jl_append_value(out, y[c * ld_y + q], y_is_int != 0);
                    out += '\n';
                }
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        if (writer.joinable()) writer.join();  // the other set is free again; rounds are written in order
        writer = std::thread([&ok, &total, f, &parts]() {
            for (const std::string& s : parts) {
                if (ok && !s.empty()) ok = fwrite(s.data(), 1, s.size(), f) == s.size();
                total += int64_t(s.size());
            }
        });
    }
    if (writer.joinable()) writer.join();
    const bool closed = fclose(f) == 0;
    SS_REQUIRE(ok && closed, "ss_save_rows: write to %s failed", path);
    if (bytes_written) *bytes_written = total;
    return SS_OK;
}

/* the same from device-resident score / label blocks (nq x nt): downloaded through pinned staging, then written */
int32_t ss_save_rows_mat(ss_ctx* ctx, const char* path, int32_t append, int64_t fold, const char* const* qnames,
                         const char* const* tnames, const ss_mat* yhat, const ss_mat* y, int32_t y_is_int, int32_t delimiter,
                         int64_t* bytes_written) {
    SS_REQUIRE(ctx && yhat && y && yhat->rows == y->rows && yhat->cols == y->cols, "ss_save_rows_mat: yhat and y must have the same shape");
    SS_CHECK_CUDA(cudaSetDevice(ctx->device));
    const int64_t nq = yhat->rows, nt = yhat->cols;
    double *hp = nullptr, *hy = nullptr;
    const size_t bytes = size_t(std::max<int64_t>(nq * nt, 1)) * 8;
    SS_CHECK_CUDA(cudaMallocHost(&hp, bytes));
    if (cudaMallocHost(&hy, bytes) != cudaSuccess) {
        cudaGetLastError();
        cudaFreeHost(hp);
        ss::set_error("ss_save_rows_mat: out of pinned host memory");
        return SS_ERR_OOM;
    }
    int32_t st = SS_OK;
    if (nq * nt > 0) {
        cudaError_t e = cudaMemcpy2DAsync(hp, size_t(nq) * 8, yhat->d, size_t(yhat->ld) * 8, size_t(nq) * 8, size_t(nt), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpy2DAsync(hy, size_t(nq) * 8, y->d, size_t(y->ld) * 8, size_t(nq) * 8, size_t(nt), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            ss::set_error("ss_save_rows_mat: download failed: %s", cudaGetErrorString(e));
            st = SS_ERR_CUDA;
        }
    }
    if (st == SS_OK) st = ss_save_rows(path, append, fold, nq, nt, qnames, tnames, hp, nq, 0, hy, nq, y_is_int, delimiter, bytes_written);
    cudaFreeHost(hp);
    cudaFreeHost(hy);
    return st;
}

}  // extern "C"
