// mm_host.cpp -- Matrix Market reader (.mtx, .gz, .tar.gz, .tgz) for the host side of the C ABI.
//
// Behaviour follows the reference reader (matrix/matrix-market.cpp:346-555, 738-861):
//   * first line "%%MatrixMarket matrix {coordinate|array} {real|complex|integer|pattern}
//     {general|symmetric|skew-symmetric|hermitian}", keywords case-insensitive, banner exact;
//   * then lines starting with '%', then one size line, then num_entries records that are plain
//     whitespace-separated tokens (records may span or share lines);
//   * complex keeps the real part, integer is widened, pattern becomes 1.0; symmetric files are
//     NOT expanded;
//   * a .tar.gz/.tgz archive is searched for the member "<name>/<name>.mtx".
// The implementation is a single pass over an in-memory buffer (the whole file is read, and
// inflated with zlib when compressed) instead of iostream extraction, which is what makes the
// reference's loader take seconds on an 80 MB file; files with one record per line (all of
// SuiteSparse) are parsed by several threads.
#include "mm_host.hpp"

#include "../../include/spmv_b200.h"

#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>

namespace spmvb200 {

int fail(int code, const std::string & msg);

namespace {

struct Cursor {
    const char * p;
    const char * end;
    bool at_end() const { return p >= end; }
};

inline bool is_space(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

inline void skip_space(Cursor & c)
{
    while (c.p < c.end && is_space(*c.p)) ++c.p;
}

Cursor take_line(Cursor & c)
{
    Cursor line{c.p, c.p};
    while (c.p < c.end && *c.p != '\n') ++c.p;
    line.end = c.p;
    if (c.p < c.end) ++c.p;
    return line;
}

std::string take_word(Cursor & c, bool lower)
{
    skip_space(c);
    std::string w;
    while (c.p < c.end && !is_space(*c.p)) {
        char ch = *c.p++;
        if (lower && ch >= 'A' && ch <= 'Z') ch = (char)(ch - 'A' + 'a');
        w.push_back(ch);
    }
    return w;
}

bool take_i32(Cursor & c, int32_t & v)
{
    skip_space(c);
    const char * b = c.p;
    if (b < c.end && *b == '+') ++b;
    auto r = std::from_chars(b, c.end, v, 10);
    if (r.ec != std::errc() || r.ptr == b) return false;
    c.p = r.ptr;
    return true;
}

bool take_f64(Cursor & c, double & v)
{
    skip_space(c);
    const char * b = c.p;
    if (b < c.end && *b == '+') ++b;
    auto r = std::from_chars(b, c.end, v, std::chars_format::general);
    if (r.ec == std::errc::result_out_of_range) {  // denormal / overflow: fall back to strtod semantics
        std::string tok(b, r.ptr);
        v = strtod(tok.c_str(), nullptr);
        c.p = r.ptr;
        return true;
    }
    if (r.ec != std::errc() || r.ptr == b) return false;
    c.p = r.ptr;
    return true;
}

// Parallel fast path for the common layout -- exactly one record per line.  The text after the size line
// is cut into chunks at line boundaries, every chunk is parsed by its own thread into private arrays, and
// the pieces are concatenated.  Any line that is not one complete record (records spanning or sharing
// lines, blank lines, junk) makes the whole fast path stand down, and the sequential tokenizer below --
// which defines the behaviour and the error messages -- runs instead.  Returns true when it filled m.
inline void skip_blank(Cursor & c)  // spaces and tabs, never a line end
{
    while (c.p < c.end && (*c.p == ' ' || *c.p == '\t')) ++c.p;
}

inline bool blank_i32(Cursor & c, int32_t & v)
{
    skip_blank(c);
    auto r = std::from_chars(c.p, c.end, v, 10);
    if (r.ec != std::errc() || r.ptr == c.p) return false;
    c.p = r.ptr;
    return true;
}

inline bool blank_f64(Cursor & c, double & v)
{
    skip_blank(c);
    auto r = std::from_chars(c.p, c.end, v, std::chars_format::general);
    if (r.ec != std::errc() || r.ptr == c.p) return false;  // also "+1.0", overflow, denormals: left to the sequential reader
    c.p = r.ptr;
    return true;
}

// Parses the lines of `c` into i[0..], j[0..], a[0..]; returns the number of records, or -1 when a line is
// not exactly one record or more than `room` records turn up.
int64_t parse_lines(Cursor c, int field, int32_t rows, int32_t columns, int32_t * pi, int32_t * pj, double * pa, int64_t room)
{
    int64_t k = 0;
    while (!c.at_end()) {
        int32_t i = 0, j = 0;
        double a = 1.0;
        bool ok = blank_i32(c, i) && blank_i32(c, j);
        if (ok) {
            switch (field) {
            case 0: ok = blank_f64(c, a); break;
            case 1: { double im; ok = blank_f64(c, a) && blank_f64(c, im); break; }
            case 2: { int32_t v = 0; ok = blank_i32(c, v); a = (double)v; break; }
            default: break;
            }
        }
        if (ok) {
            skip_blank(c);
            if (c.p < c.end && *c.p == '\r') ++c.p;
            ok = c.at_end() || *c.p == '\n';
        }
        if (!ok || k >= room || i < 1 || i > rows || j < 1 || j > columns) return -1;
        if (!c.at_end()) ++c.p;  // the line end
        pi[k] = i; pj[k] = j; pa[k] = a;
        ++k;
    }
    return k;
}

bool parse_entries_parallel(Cursor c, spmvb200_mm_s * m)
{
    const size_t n = (size_t)m->num_entries;
    const size_t bytes = (size_t)(c.end - c.p);
    unsigned T = std::min<unsigned>(std::thread::hardware_concurrency(), 32u);
    if (const char * env = getenv("SPMVB200_PARSE_THREADS")) T = (unsigned)std::max(1, atoi(env));
    T = (unsigned)std::min<size_t>(T, bytes / ((size_t)4 << 20));  // at least 4 MB of text per thread
    if (n < ((size_t)1 << 16) || T < 2) return false;
    std::vector<Cursor> chunk(T);
    const char * b = c.p;
    for (unsigned t = 0; t < T; t++) {
        const char * e = t + 1 == T ? c.end : c.p + bytes / T * (t + 1);
        if (e < b) e = b;
        while (e < c.end && e > c.p && e[-1] != '\n') ++e;  // extend to the end of the line
        chunk[t] = Cursor{b, e};
        b = e;
    }
    // pass 1: lines per chunk (memchr speed), so that every thread knows where its records go
    std::vector<int64_t> lines(T, 0), got(T, 0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; t++)
        th.emplace_back([&, t] {
            int64_t cnt = 0;
            const char * p = chunk[t].p;
            while (p < chunk[t].end) {
                const char * q = (const char *)memchr(p, '\n', (size_t)(chunk[t].end - p));
                ++cnt;
                if (!q) break;
                p = q + 1;
            }
            lines[t] = cnt;
        });
    for (auto & x : th) x.join();
    size_t total = 0;
    for (auto v : lines) total += (size_t)v;
    if (total != n) return false;  // blank lines, too few or too many records: the sequential reader decides
    m->i.resize(n); m->j.resize(n); m->a.resize(n);
    // pass 2: parse straight into place
    th.clear();
    size_t off = 0;
    for (unsigned t = 0; t < T; t++) {
        th.emplace_back([&, t, off] {
            got[t] = parse_lines(chunk[t], m->field, m->rows, m->columns, m->i.data() + off, m->j.data() + off, m->a.data() + off, lines[t]);
        });
        off += (size_t)lines[t];
    }
    for (auto & x : th) x.join();
    for (unsigned t = 0; t < T; t++)
        if (got[t] != lines[t]) return false;
    return true;
}

int parse(Cursor c, spmvb200_mm_s * m)
{
    if (c.at_end()) return fail(SPMVB200_ERR_PARSE, "Failed to parse header: Expected \"%%MatrixMarket\", got \"\"");
    Cursor h = take_line(c);
    std::string w = take_word(h, false);
    if (w != "%%MatrixMarket")
        return fail(SPMVB200_ERR_PARSE, "Failed to parse header: Expected \"%%MatrixMarket\", got \"" + w + "\"");
    w = take_word(h, true);
    if (w != "matrix") return fail(SPMVB200_ERR_PARSE, "Failed to parse header: Expected \"matrix\", got \"" + w + "\"");
    w = take_word(h, true);
    if (w == "coordinate") m->format = 0;
    else if (w == "array") m->format = 1;
    else return fail(SPMVB200_ERR_PARSE, "Expected \"coordinate\" or \"array\", got \"" + w + "\"");
    w = take_word(h, true);
    if (w == "real") m->field = 0;
    else if (w == "complex") m->field = 1;
    else if (w == "integer") m->field = 2;
    else if (w == "pattern") m->field = 3;
    else return fail(SPMVB200_ERR_PARSE, "Expected \"real\", \"complex\", \"integer\", or \"pattern\", got \"" + w + "\"");
    w = take_word(h, true);
    if (w == "general") m->symmetry = 0;
    else if (w == "symmetric") m->symmetry = 1;
    else if (w == "skew-symmetric") m->symmetry = 2;
    else if (w == "hermitian") m->symmetry = 3;
    else return fail(SPMVB200_ERR_PARSE,
                     "Expected \"general\", \"symmetric\", \"skew-symmetric\", or \"hermitian\", got \"" + w + "\"");

    while (!c.at_end() && *c.p == '%') take_line(c);

    if (c.at_end()) return fail(SPMVB200_ERR_PARSE, "Failed to parse size");
    Cursor s = take_line(c);
    if (!take_i32(s, m->rows))
        return fail(SPMVB200_ERR_OVERFLOW, "Failed to parse size: Integer overflow when reading number of rows");
    if (!take_i32(s, m->columns))
        return fail(SPMVB200_ERR_OVERFLOW, "Failed to parse size: Integer overflow when reading number of columns");
    if (m->format == 1) {
        m->num_entries = 0;
        return 0;
    }
    if (!take_i32(s, m->num_entries))
        return fail(SPMVB200_ERR_OVERFLOW, "Failed to parse size: Integer overflow when reading number of non-zeros");
    if (m->rows < 0 || m->columns < 0 || m->num_entries < 0) return fail(SPMVB200_ERR_PARSE, "Failed to parse size");

    const size_t n = (size_t)m->num_entries;
    if (parse_entries_parallel(c, m)) return 0;
    m->i.resize(n); m->j.resize(n); m->a.resize(n);
    for (size_t k = 0; k < n; k++) {
        bool ok = take_i32(c, m->i[k]) && take_i32(c, m->j[k]);
        if (ok) {
            switch (m->field) {
            case 0: ok = take_f64(c, m->a[k]); break;
            case 1: { double im; ok = take_f64(c, m->a[k]) && take_f64(c, im); break; }
            case 2: { int32_t v = 0; ok = take_i32(c, v); m->a[k] = (double)v; break; }
            default: m->a[k] = 1.0;
            }
        }
        if (!ok)
            return fail(SPMVB200_ERR_PARSE, "Failed to parse entries: Expected " + std::to_string(n) +
                                                " entries, got " + std::to_string(k) + " entries.");
        if (m->i[k] < 1 || m->i[k] > m->rows || m->j[k] < 1 || m->j[k] > m->columns)
            return fail(SPMVB200_ERR_PARSE, "Failed to parse entries: index outside the matrix in entry " +
                                                std::to_string(k + 1));
    }
    return 0;
}

bool ends_with(const std::string & s, const std::string & t)
{
    return s.size() > t.size() && s.compare(s.size() - t.size(), t.size(), t) == 0;
}

int read_file(const std::string & path, std::string & out)
{
    FILE * f = fopen(path.c_str(), "rb");
    if (!f) return fail(SPMVB200_ERR_IO, strerror(errno));
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, got);
    const bool bad = ferror(f) != 0;
    fclose(f);
    if (bad) return fail(SPMVB200_ERR_IO, "read error");
    return 0;
}

// gzip or zlib wrapped deflate, auto-detected (windowBits 15+32 like the reference, zlibstream.cpp:67)
int inflate_all(const std::string & in, std::string & out)
{
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    int err = inflateInit2(&zs, 15 + 32);
    if (err != Z_OK) return fail(SPMVB200_ERR_IO, std::string("inflateInit2: ") + zError(err));
    zs.next_in = (Bytef *)in.data();
    zs.avail_in = (uInt)std::min<size_t>(in.size(), 1u << 30);
    size_t consumed = 0;
    std::vector<unsigned char> buf(1 << 20);
    while (true) {
        zs.next_out = buf.data();
        zs.avail_out = (uInt)buf.size();
        err = inflate(&zs, Z_NO_FLUSH);
        if (err != Z_OK && err != Z_STREAM_END && err != Z_BUF_ERROR) {
            inflateEnd(&zs);
            return fail(SPMVB200_ERR_IO, std::string("inflate: ") + zError(err));
        }
        out.append((const char *)buf.data(), buf.size() - zs.avail_out);
        if (err == Z_STREAM_END) break;
        if (zs.avail_in == 0) {
            consumed = (const char *)zs.next_in - in.data();
            if (consumed >= in.size()) break;  // truncated stream: hand over what we have
            zs.avail_in = (uInt)std::min<size_t>(in.size() - consumed, 1u << 30);
        }
    }
    inflateEnd(&zs);
    return 0;
}

uint64_t tar_number(const char * s, size_t n)
{
    if ((unsigned char)s[0] & 0x80) {  // base-256
        uint64_t v = (unsigned char)s[0] & 0x7f;
        for (size_t k = 1; k < n; k++) v = (v << 8) | (unsigned char)s[k];
        return v;
    }
    uint64_t v = 0;
    for (size_t k = 0; k < n && s[k]; k++)
        if (s[k] >= '0' && s[k] <= '7') v = v * 8 + (uint64_t)(s[k] - '0');
    return v;
}

// Locate member `name` in an (uncompressed) tar image: 512-byte headers, name at 0..99, size at 124..135.
int tar_member(const std::string & tar, const std::string & name, Cursor & out)
{
    size_t pos = 0;
    while (pos + 512 <= tar.size()) {
        const char * h = tar.data() + pos;
        const uint64_t size = tar_number(h + 124, 12);
        const size_t len = std::min<size_t>(name.size(), 100);
        if (strncmp(name.c_str(), h, len) == 0) {
            const size_t b = pos + 512;
            const size_t e = std::min<size_t>(tar.size(), b + size);
            out = Cursor{tar.data() + b, tar.data() + e};
            return 0;
        }
        pos += 512 + (size + 511) / 512 * 512;
    }
    return fail(SPMVB200_ERR_PARSE, "Failed to parse header: Expected \"%%MatrixMarket\", got \"\"");
}

}  // namespace

int mm_parse_text(const char * text, size_t len, spmvb200_mm_s ** out)
{
    auto * m = new spmvb200_mm_s();
    int rc = parse(Cursor{text, text + len}, m);
    if (rc) { delete m; return rc; }
    *out = m;
    return 0;
}

static int load_plain(const std::string & path, spmvb200_mm_s ** out)
{
    std::string raw;
    int rc = read_file(path, raw);
    if (rc) return rc;
    std::string ext;
    if (ends_with(path, ".tar.gz")) ext = ".tar.gz";
    else if (ends_with(path, ".tgz")) ext = ".tgz";
    if (!ext.empty()) {
        std::string tar;
        if ((rc = inflate_all(raw, tar))) return rc;
        size_t start = path.find_last_of('/');
        start = start == std::string::npos ? 0 : start + 1;
        const std::string base = path.substr(start, path.size() - ext.size() - start);
        Cursor member{nullptr, nullptr};
        if ((rc = tar_member(tar, base + "/" + base + ".mtx", member))) return rc;
        return mm_parse_text(member.p, (size_t)(member.end - member.p), out);
    }
    if (ends_with(path, ".gz")) {
        std::string text;
        if ((rc = inflate_all(raw, text))) return rc;
        return mm_parse_text(text.data(), text.size(), out);
    }
    return mm_parse_text(raw.data(), raw.size(), out);
}

int mm_load_path(const char * path_c, spmvb200_mm_s ** out)
{
    std::string path(path_c);
    // "__RCM" / "__GP<n>" suffixes select a reordering (matrix-market.cpp:786-802): the suffix is cut off,
    // the file is loaded, then the order is computed and applied (RCM first, then GP).
    bool rcm = false, gp = false;
    int nparts = 0;
    size_t pos = path.rfind("__RCM");
    if (pos != std::string::npos) {
        rcm = true;
        path.erase(pos);
    }
    pos = path.rfind("__GP");
    if (pos != std::string::npos) {
        // without METIS the reference's graph-partitioning order is the identity; with the global option
        // "mm.gp_partitioner" the library's own K-way partitioner stands in for METIS (reorder_host.cpp)
        gp = gp_partitioner() != 0;
        if (pos + 4 < path.size()) nparts = std::atoi(path.c_str() + pos + 4);  // sscanf("%d"), matrix-market.cpp:797-801
        path.erase(pos);
    }
    int rc = load_plain(path, out);
    if (rc || !(rcm || gp)) return rc;
    std::vector<int32_t> order((size_t)(*out)->rows);
    if (rcm) {
        rc = mm_order_rcm(*out, order.data());
        if (rc == 0) rc = mm_permute(*out, order.data());
    }
    if (rc == 0 && gp) {
        rc = mm_order_gp_kway(*out, nparts, order.data());
        if (rc == 0) rc = mm_permute(*out, order.data());
    }
    if (rc) {
        delete *out;
        *out = nullptr;
    }
    return rc;
}

int mm_from_entries(int32_t rows, int32_t columns, int32_t n, const int32_t * i, const int32_t * j,
                    const double * a, spmvb200_mm_s ** out)
{
    for (int32_t k = 0; k < n; k++)
        if (i[k] < 1 || i[k] > rows || j[k] < 1 || j[k] > columns)
            return fail(SPMVB200_ERR_INVALID, "entry index outside the matrix");
    auto * m = new spmvb200_mm_s();
    m->rows = rows; m->columns = columns; m->num_entries = n;
    m->i.assign(i, i + n); m->j.assign(j, j + n); m->a.assign(a, a + n);
    *out = m;
    return 0;
}

int mm_row_lengths(const spmvb200_mm_s * mm, int32_t * lengths)
{
    std::fill(lengths, lengths + mm->rows, 0);
    for (int32_t k = 0; k < mm->num_entries; k++) ++lengths[mm->i[k] - 1];
    return 0;
}

int mm_sort(spmvb200_mm_s * mm, bool row_major)
{
    const size_t n = (size_t)mm->num_entries;
    std::vector<uint32_t> perm(n);
    std::iota(perm.begin(), perm.end(), 0u);
    const auto & p = row_major ? mm->i : mm->j;
    const auto & q = row_major ? mm->j : mm->i;
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t x, uint32_t y) {
        return p[x] != p[y] ? p[x] < p[y] : q[x] < q[y];
    });
    std::vector<int32_t> ni(n), nj(n);
    std::vector<double> na(n);
    for (size_t k = 0; k < n; k++) { ni[k] = mm->i[perm[k]]; nj[k] = mm->j[perm[k]]; na[k] = mm->a[perm[k]]; }
    mm->i.swap(ni); mm->j.swap(nj); mm->a.swap(na);
    return 0;
}

}  // namespace spmvb200
