// f4: compiled A2M / FASTA ingest (host code, no device work).
//
// Replaces the text -> matrix part of what plmc does before inference (8a row a4) and the reference's in-tree
// readers read_fasta / sequences_to_matrix / map_matrix (evcouplings/align/alignment.py:42-74, 410-443, 479-495:
// a Python generator plus np.vectorize over N*L characters).  The file is mmap'ed once; records may be wrapped
// over several lines; '\r' is ignored.
//   evc_a2m_scan    rows, common row width, bytes needed for the NUL-separated ids
//   evc_a2m_read    raw character matrix (rows x width) + ids
//   evc_msa_encode  raw characters -> model codes through a 256-entry table, row validity (ANY character of the
//                   row outside the table invalidates it, insert columns included), focus-column selection and
//                   compaction to the valid rows -- multi-threaded, one pass over the matrix per step
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/evcplm.h"
#include "common.cuh"

namespace evc {

struct MappedFile {
    const unsigned char *p = nullptr;
    size_t n = 0;
    int fd = -1;
    ~MappedFile()
    {
        if (p && n) munmap(const_cast<unsigned char *>(p), n);
        if (fd >= 0) close(fd);
    }
    int open_path(const char *path)
    {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) { set_error(std::string("cannot open alignment ") + path); return 1; }
        struct stat st;
        if (fstat(fd, &st) != 0) { set_error("fstat failed"); return 1; }
        n = (size_t)st.st_size;
        if (n == 0) { p = nullptr; return 0; }
        void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { p = nullptr; set_error("mmap failed"); return 1; }
        p = static_cast<const unsigned char *>(m);
        madvise(m, n, MADV_SEQUENTIAL);
        return 0;
    }
};

static inline bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\f' || c == '\v'; }

// Walk the records.  on_header(begin, end) for the header text (without '>' and without surrounding whitespace);
// on_chunk(ptr, len) for every sequence line of the current record (already stripped).
template <class H, class C>
static void walk(const unsigned char *p, size_t n, H on_header, C on_chunk)
{
    size_t pos = 0;
    bool in_record = false;
    while (pos < n) {
        const unsigned char *nl = static_cast<const unsigned char *>(memchr(p + pos, '\n', n - pos));
        size_t end = nl ? (size_t)(nl - p) : n;
        size_t b = pos, e = end;
        if (p[pos] == '>') {
            b = pos + 1;
            while (b < e && is_space(p[b])) b++;
            while (e > b && is_space(p[e - 1])) e--;
            on_header(p + b, p + e);
            in_record = true;
        } else if (in_record) {
            while (b < e && is_space(p[b])) b++;
            while (e > b && is_space(p[e - 1])) e--;
            if (e > b) on_chunk(p + b, e - b);
        }
        pos = end + 1;
    }
}

static int n_threads(int64_t rows)
{
    unsigned hw = std::thread::hardware_concurrency();
    int t = (int)std::min<unsigned>(hw ? hw : 1, 32);
    return (int)std::max<int64_t>(1, std::min<int64_t>(t, rows / 4096 + 1));
}

template <class F>
static void parallel_rows(int64_t rows, F fn)
{
    const int T = n_threads(rows);
    if (T == 1) { fn(0, rows); return; }
    std::vector<std::thread> th;
    const int64_t per = (rows + T - 1) / T;
    for (int t = 0; t < T; t++) {
        const int64_t lo = t * per, hi = std::min(rows, lo + per);
        if (lo >= hi) break;
        th.emplace_back([=] { fn(lo, hi); });
    }
    for (auto &x : th) x.join();
}

}  // namespace evc

using namespace evc;

extern "C" {

int evc_a2m_scan(const char *path, int64_t *n_rows, int64_t *width, int64_t *ids_bytes)
{
    if (!path || !n_rows || !width || !ids_bytes) { set_error("evc_a2m_scan: null pointer"); return 1; }
    MappedFile f;
    if (f.open_path(path)) return 1;
    int64_t rows = 0, w0 = -1, cur = 0, idb = 0, bad_row = -1, bad_len = 0;
    auto close_record = [&]() {
        if (rows == 0) return;
        if (w0 < 0) w0 = cur;
        else if (cur != w0 && bad_row < 0) { bad_row = rows - 1; bad_len = cur; }
    };
    walk(f.p, f.n,
         [&](const unsigned char *b, const unsigned char *e) {
             close_record();
             rows++;
             cur = 0;
             idb += (e - b) + 1;
         },
         [&](const unsigned char *, size_t len) { cur += (int64_t)len; });
    close_record();
    if (rows == 0) { set_error(std::string("alignment ") + path + " contains no sequences"); return 2; }
    if (w0 == 0) { set_error(std::string("alignment ") + path + " has zero-length sequences"); return 2; }
    if (bad_row >= 0) {
        set_error("ragged alignment: row " + std::to_string(bad_row) + " has length " + std::to_string(bad_len) +
                  ", expected " + std::to_string(w0));
        return 2;
    }
    *n_rows = rows;
    *width = w0;
    *ids_bytes = idb;
    return 0;
}

int evc_a2m_read(const char *path, int64_t n_rows, int64_t width, uint8_t *raw, char *ids, int64_t ids_bytes)
{
    if (!path || !raw || !ids) { set_error("evc_a2m_read: null pointer"); return 1; }
    MappedFile f;
    if (f.open_path(path)) return 1;
    int64_t row = -1, col = 0, idpos = 0;
    bool overflow = false;
    walk(f.p, f.n,
         [&](const unsigned char *b, const unsigned char *e) {
             row++;
             col = 0;
             const int64_t len = e - b;
             if (idpos + len + 1 > ids_bytes || row >= n_rows) { overflow = true; return; }
             memcpy(ids + idpos, b, (size_t)len);
             ids[idpos + len] = '\0';
             idpos += len + 1;
         },
         [&](const unsigned char *ptr, size_t len) {
             if (overflow || row < 0 || row >= n_rows || col + (int64_t)len > width) { overflow = true; return; }
             memcpy(raw + row * width + col, ptr, len);
             col += (int64_t)len;
         });
    if (overflow || row + 1 != n_rows) { set_error("evc_a2m_read: file changed since evc_a2m_scan"); return 2; }
    return 0;
}

int evc_msa_encode(const uint8_t *raw, int64_t n_rows, int64_t width, const uint8_t *lut /* 256 */,
                   const int64_t *cols, int64_t n_cols, uint8_t *valid_out /* n_rows */,
                   uint8_t *codes_out /* n_valid x n_cols, capacity n_rows x n_cols */, int64_t *n_valid_out)
{
    if (!raw || !lut || !cols || !valid_out || !codes_out || !n_valid_out) { set_error("evc_msa_encode: null pointer"); return 1; }
    for (int64_t k = 0; k < n_cols; k++)
        if (cols[k] < 0 || cols[k] >= width) { set_error("evc_msa_encode: column index out of range"); return 1; }
    // pass 1: a row is valid iff every character (all columns, inserts included) maps to a code
    parallel_rows(n_rows, [&](int64_t lo, int64_t hi) {
        for (int64_t r = lo; r < hi; r++) {
            const uint8_t *row = raw + r * width;
            unsigned bad = 0;
            for (int64_t c = 0; c < width; c++) bad |= (lut[row[c]] == 255u);
            valid_out[r] = bad ? 0 : 1;
        }
    });
    // destination row of every valid row (exclusive prefix sum)
    std::vector<int64_t> dst((size_t)n_rows);
    int64_t nv = 0;
    for (int64_t r = 0; r < n_rows; r++) { dst[(size_t)r] = nv; nv += valid_out[r]; }
    // pass 2: selected columns of the valid rows, coded
    parallel_rows(n_rows, [&](int64_t lo, int64_t hi) {
        for (int64_t r = lo; r < hi; r++) {
            if (!valid_out[r]) continue;
            const uint8_t *row = raw + r * width;
            uint8_t *out = codes_out + dst[(size_t)r] * n_cols;
            for (int64_t k = 0; k < n_cols; k++) out[k] = lut[row[cols[k]]];
        }
    });
    *n_valid_out = nv;
    return 0;
}

}  // extern "C"
