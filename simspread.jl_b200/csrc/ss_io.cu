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
#include <charconv>
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
