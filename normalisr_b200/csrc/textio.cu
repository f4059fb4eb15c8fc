// Text I/O of the command-line layer around the hot path (reference src/normalisr/run.py:10-35):
//   file_read_tsv   numpy.loadtxt(f, delimiter='\t')          -> nsr_tsv_shape + nsr_tsv_read
//   file_write_tsv  numpy.savetxt(f, d, delimiter, '%.8G')    -> nsr_tsv_write
//   file_read_coo   scipy.io.mmread(f) (MatrixMarket coordinate) -> nsr_mtx_shape + nsr_mtx_read_dense
// Host code (no kernels): the files are mapped, cut into line ranges and parsed by a pool of threads
// straight into the caller's buffer - page-locked memory allocated by the Python layer, so the matrix
// goes to the device with one asynchronous copy and never exists as a Python object per value.
// Every function returns 0 on success; the message of a failure is in nsr_last_error().
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "nsr_common.cuh"

namespace {

struct Mapped {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    ~Mapped() {
        if (p && n) munmap((void*)p, n);
        if (fd >= 0) close(fd);
    }
};

int map_file(const char* path, Mapped& m) {
    m.fd = open(path, O_RDONLY);
    NSR_REQUIRE(m.fd >= 0, "cannot open %s: %s", path, strerror(errno));
    struct stat st;
    NSR_REQUIRE(fstat(m.fd, &st) == 0, "cannot stat %s: %s", path, strerror(errno));
    m.n = (size_t)st.st_size;
    if (m.n == 0) return 0;
    void* p = mmap(nullptr, m.n, PROT_READ, MAP_PRIVATE, m.fd, 0);
    NSR_REQUIRE(p != MAP_FAILED, "cannot map %s: %s", path, strerror(errno));
    m.p = (const char*)p;
    madvise(p, m.n, MADV_SEQUENTIAL);
    return 0;
}

inline bool blank_or_comment(const char* b, const char* e, char comment) {
    while (b < e && (*b == ' ' || *b == '\t' || *b == '\r')) ++b;
    return b == e || *b == comment;
}

// offsets of the first byte of every data line (blank lines and '#' comments skipped, like loadtxt)
void data_lines(const Mapped& m, char comment, std::vector<size_t>& starts) {
    size_t i = 0;
    while (i < m.n) {
        const char* nl = (const char*)memchr(m.p + i, '\n', m.n - i);
        const size_t end = nl ? (size_t)(nl - m.p) : m.n;
        if (!blank_or_comment(m.p + i, m.p + end, comment)) starts.push_back(i);
        i = end + 1;
    }
}

inline size_t line_end(const Mapped& m, size_t start) {
    const char* nl = (const char*)memchr(m.p + start, '\n', m.n - start);
    size_t end = nl ? (size_t)(nl - m.p) : m.n;
    while (end > start && (m.p[end - 1] == '\r')) --end;
    return end;
}

int64_t count_fields(const char* b, const char* e, char delim) {
    int64_t c = 1;
    for (const char* p = b; p < e; ++p) c += *p == delim;
    return c;
}

int n_threads(int asked, size_t work) {
    int t = asked > 0 ? asked : (int)std::thread::hardware_concurrency();
    if (t < 1) t = 1;
    if (t > 64) t = 64;
    if ((size_t)t > work) t = work ? (int)work : 1;
    return t;
}

}  // namespace

extern "C" int nsr_tsv_shape(const char* path, char delimiter, int64_t* rows, int64_t* cols) {
    NSR_REQUIRE(path && rows && cols, "nsr_tsv_shape: null argument");
    Mapped m;
    if (map_file(path, m)) return 1;
    std::vector<size_t> starts;
    data_lines(m, '#', starts);
    *rows = (int64_t)starts.size();
    *cols = starts.empty() ? 0 : count_fields(m.p + starts[0], m.p + line_end(m, starts[0]), delimiter);
    return 0;
}

// out[r * ld + c] for the rows x cols table of `path` (shape from nsr_tsv_shape).  Values are parsed with
// strtod (what numpy.loadtxt's float conversion amounts to); a row with another number of fields or an
// unparsable field is an error.
extern "C" int nsr_tsv_read(const char* path, char delimiter, double* out, int64_t rows, int64_t cols, int64_t ld,
                            int threads) {
    NSR_REQUIRE(path && out && rows >= 0 && cols >= 0 && ld >= cols, "nsr_tsv_read: bad arguments");
    Mapped m;
    if (map_file(path, m)) return 1;
    std::vector<size_t> starts;
    data_lines(m, '#', starts);
    NSR_REQUIRE((int64_t)starts.size() == rows, "nsr_tsv_read: %s has %lld data lines, expected %lld", path,
                (long long)starts.size(), (long long)rows);
    const int nt = n_threads(threads, (size_t)rows);
    std::atomic<int64_t> bad_row(-1);
    auto work = [&](int t) {
        const int64_t r0 = rows * t / nt, r1 = rows * (t + 1) / nt;
        std::string field;
        for (int64_t r = r0; r < r1 && bad_row.load(std::memory_order_relaxed) < 0; ++r) {
            const char* p = m.p + starts[(size_t)r];
            const char* e = m.p + line_end(m, starts[(size_t)r]);
            int64_t c = 0;
            while (true) {
                const char* d = (const char*)memchr(p, delimiter, (size_t)(e - p));
                const char* fe = d ? d : e;
                if (c >= cols) { bad_row = r; break; }
                field.assign(p, (size_t)(fe - p));          // strtod needs a terminator; fields are short
                char* endp = nullptr;
                errno = 0;
                const double v = strtod(field.c_str(), &endp);
                while (endp && (*endp == ' ' || *endp == '\r')) ++endp;
                if (endp == field.c_str() || (endp && *endp != '\0')) { bad_row = r; break; }
                out[r * ld + c] = v;
                ++c;
                if (!d) break;
                p = d + 1;
            }
            if (c != cols && bad_row.load() < 0) bad_row = r;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    NSR_REQUIRE(bad_row.load() < 0, "nsr_tsv_read: %s: data line %lld does not hold %lld numeric fields", path,
                (long long)bad_row.load() + 1, (long long)cols);
    return 0;
}

// numpy.savetxt(path, data, delimiter=delimiter, fmt='%.8G'): C printf semantics, '\n' line ends.
// `precision` is the 8 of the reference's fmt_float (run.py:6).  Rows are formatted by a pool of threads
// into per-thread buffers and written in order.
extern "C" int nsr_tsv_write(const char* path, char delimiter, const double* data, int64_t rows, int64_t cols,
                             int64_t ld, int precision, int threads) {
    NSR_REQUIRE(path && (data || rows * cols == 0) && rows >= 0 && cols >= 0 && ld >= cols && precision >= 1 && precision <= 17,
                "nsr_tsv_write: bad arguments");
    FILE* f = fopen(path, "wb");
    NSR_REQUIRE(f != nullptr, "cannot open %s for writing: %s", path, strerror(errno));
    const int64_t rows_per_batch = cols > 0 ? ((int64_t)(8 << 20) / (cols * 12) + 1) : rows + 1;
    char fmt[16];
    snprintf(fmt, sizeof(fmt), "%%.%dG", precision);
    int rc = 0;
    for (int64_t b0 = 0; b0 < rows && rc == 0; b0 += rows_per_batch * 64) {
        const int64_t b1 = b0 + rows_per_batch * 64 < rows ? b0 + rows_per_batch * 64 : rows;
        const int nt = n_threads(threads, (size_t)(b1 - b0));
        std::vector<std::string> bufs((size_t)nt);
        auto work = [&](int t) {
            const int64_t r0 = b0 + (b1 - b0) * t / nt, r1 = b0 + (b1 - b0) * (t + 1) / nt;
            std::string& s = bufs[(size_t)t];
            s.reserve((size_t)((r1 - r0) * cols * 12));
            char tmp[40];
            for (int64_t r = r0; r < r1; ++r) {
                for (int64_t c = 0; c < cols; ++c) {
                    const int len = snprintf(tmp, sizeof(tmp), fmt, data[r * ld + c]);
                    s.append(tmp, (size_t)len);
                    s.push_back(c + 1 < cols ? delimiter : '\n');
                }
                if (cols == 0) s.push_back('\n');
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
        for (auto& s : bufs)
            if (!s.empty() && fwrite(s.data(), 1, s.size(), f) != s.size()) rc = 1;
    }
    if (fclose(f) != 0) rc = 1;
    NSR_REQUIRE(rc == 0, "nsr_tsv_write: write to %s failed: %s", path, strerror(errno));
    return 0;
}

// ---- MatrixMarket coordinate files (scipy.io.mmread, run.py:10-17) --------------------------------
namespace {
struct MtxHeader {
    int64_t rows = 0, cols = 0, nnz = 0;
    int field = 0;        // 0 real, 1 integer, 2 pattern
    int symmetric = 0;    // 0 general, 1 symmetric, 2 skew-symmetric
    size_t data_start = 0;
};

int parse_mtx_header(const Mapped& m, const char* path, MtxHeader& h) {
    NSR_REQUIRE(m.n > 14 && !strncmp(m.p, "%%MatrixMarket", 14), "%s: not a MatrixMarket file", path);
    const size_t e0 = line_end(m, 0);
    std::string banner(m.p, e0);
    for (auto& ch : banner) ch = (char)tolower(ch);
    NSR_REQUIRE(banner.find("matrix") != std::string::npos && banner.find("coordinate") != std::string::npos,
                "%s: only 'matrix coordinate' MatrixMarket files are supported", path);
    NSR_REQUIRE(banner.find("complex") == std::string::npos && banner.find("hermitian") == std::string::npos,
                "%s: complex MatrixMarket files are not supported", path);
    h.field = banner.find("integer") != std::string::npos ? 1 : (banner.find("pattern") != std::string::npos ? 2 : 0);
    h.symmetric = banner.find("skew-symmetric") != std::string::npos ? 2 : (banner.find("symmetric") != std::string::npos ? 1 : 0);
    size_t i = e0 + 1;
    while (i < m.n) {                                       // comments, then the size line
        const size_t e = line_end(m, i);
        if (!blank_or_comment(m.p + i, m.p + e, '%')) {
            std::string s(m.p + i, e - i);
            long long r, c, z;
            NSR_REQUIRE(sscanf(s.c_str(), "%lld %lld %lld", &r, &c, &z) == 3 && r >= 0 && c >= 0 && z >= 0, "%s: bad size line",
                        path);
            h.rows = r; h.cols = c; h.nnz = z;
            const char* nl = (const char*)memchr(m.p + i, '\n', m.n - i);
            h.data_start = nl ? (size_t)(nl - m.p) + 1 : m.n;
            return 0;
        }
        const char* nl = (const char*)memchr(m.p + i, '\n', m.n - i);
        i = nl ? (size_t)(nl - m.p) + 1 : m.n;
    }
    nsr_set_error("%s: no size line", path);
    return 2;
}
}  // namespace

extern "C" int nsr_mtx_shape(const char* path, int64_t* rows, int64_t* cols, int64_t* nnz, int* is_integer) {
    NSR_REQUIRE(path && rows && cols && nnz && is_integer, "nsr_mtx_shape: null argument");
    Mapped m;
    if (map_file(path, m)) return 1;
    MtxHeader h;
    if (int rc = parse_mtx_header(m, path, h)) return rc;
    *rows = h.rows; *cols = h.cols; *nnz = h.nnz; *is_integer = h.field == 1;
    return 0;
}

// Entries of a coordinate file as triplets: row[i], col[i] (0-based) and val[i] (float64; pattern files give
// 1).  Exactly nnz entries as stored (symmetric files are NOT expanded here; *symmetric reports the banner).
// Duplicate entries are kept (scipy sums them when converting; so does the Python layer).
extern "C" int nsr_mtx_read(const char* path, int64_t nnz, int32_t* row, int32_t* col, double* val, int* symmetric,
                            int threads) {
    NSR_REQUIRE(path && row && col && val && nnz >= 0, "nsr_mtx_read: bad arguments");
    Mapped m;
    if (map_file(path, m)) return 1;
    MtxHeader h;
    if (int rc = parse_mtx_header(m, path, h)) return rc;
    NSR_REQUIRE(h.nnz == nnz, "nsr_mtx_read: %s holds %lld entries, expected %lld", path, (long long)h.nnz, (long long)nnz);
    if (symmetric) *symmetric = h.symmetric;
    // cut the data section into line ranges
    std::vector<size_t> starts;
    starts.reserve((size_t)nnz);
    size_t i = h.data_start;
    while (i < m.n) {
        const char* nl = (const char*)memchr(m.p + i, '\n', m.n - i);
        const size_t end = nl ? (size_t)(nl - m.p) : m.n;
        if (!blank_or_comment(m.p + i, m.p + end, '%')) starts.push_back(i);
        i = end + 1;
    }
    NSR_REQUIRE((int64_t)starts.size() == nnz, "nsr_mtx_read: %s: %lld entry lines, header says %lld", path,
                (long long)starts.size(), (long long)nnz);
    const int nt = n_threads(threads, (size_t)nnz);
    std::atomic<int64_t> bad(-1);
    auto work = [&](int t) {
        const int64_t e0 = nnz * t / nt, e1 = nnz * (t + 1) / nt;
        for (int64_t k = e0; k < e1 && bad.load(std::memory_order_relaxed) < 0; ++k) {
            const char* p = m.p + starts[(size_t)k];
            char* q = nullptr;
            const long long r = strtoll(p, &q, 10);
            if (q == p) { bad = k; break; }
            p = q;
            const long long c = strtoll(p, &q, 10);
            if (q == p) { bad = k; break; }
            double v = 1.0;
            if (h.field != 2) {
                p = q;
                v = strtod(p, &q);
                if (q == p) { bad = k; break; }
            }
            if (r < 1 || r > h.rows || c < 1 || c > h.cols) { bad = k; break; }
            row[k] = (int32_t)(r - 1);
            col[k] = (int32_t)(c - 1);
            val[k] = v;
        }
    };
    // strtoll / strtod read up to the next non-numeric byte: the mapping must end with one.  A file whose last
    // byte is a digit is handled by parsing the final line from a terminated copy.
    int64_t tail_fix = -1;
    if (nnz > 0 && m.n > 0 && m.p[m.n - 1] != '\n') tail_fix = nnz - 1;
    std::vector<std::thread> pool;
    const int64_t n_main = tail_fix >= 0 ? nnz - 1 : nnz;
    if (tail_fix < 0) {
        for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
    } else {
        // rare: do everything on one thread with the last line copied
        for (int64_t k = 0; k < n_main; ++k) {
            const char* p = m.p + starts[(size_t)k];
            char* q = nullptr;
            const long long r = strtoll(p, &q, 10);
            const long long c = strtoll(q, &q, 10);
            const double v = h.field != 2 ? strtod(q, &q) : 1.0;
            if (r < 1 || r > h.rows || c < 1 || c > h.cols) { bad = k; break; }
            row[k] = (int32_t)(r - 1); col[k] = (int32_t)(c - 1); val[k] = v;
        }
        std::string last(m.p + starts[(size_t)tail_fix], m.n - starts[(size_t)tail_fix]);
        long long r = 0, c = 0;
        double v = 1.0;
        const int got = h.field != 2 ? sscanf(last.c_str(), "%lld %lld %lf", &r, &c, &v) : sscanf(last.c_str(), "%lld %lld", &r, &c);
        if (got < (h.field != 2 ? 3 : 2) || r < 1 || r > h.rows || c < 1 || c > h.cols) bad = tail_fix;
        else { row[tail_fix] = (int32_t)(r - 1); col[tail_fix] = (int32_t)(c - 1); val[tail_fix] = v; }
    }
    NSR_REQUIRE(bad.load() < 0, "nsr_mtx_read: %s: bad entry line %lld", path, (long long)bad.load() + 1);
    return 0;
}
