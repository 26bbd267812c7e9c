// cache_model.cpp -- the reference's cache model applied to the SpMV reference string, on the host.
//
// North-star item (4) / SURVEY 8(f)1: the cache-simulation subsystem of the reference stays a host
// tool; what the GPU path needs from it is the predicted x-gather miss traffic for the partition that
// was actually chosen, to put next to ncu's dram__bytes.  The reference cannot deliver that as it is:
// its reference strings hard-code the equal-rows partition and carry no array attribution
// (matrix/csr-matrix.cpp:97-143), and its LRU scans the whole recency queue on every hit
// (cache-simulation/lru.cpp:38) -- minutes for 1M rows, hopeless for a 126 MB cache.  This file is the
// same model computed differently:
//   * the reference string of {csr,ell,coo}_matrix::Matrix::spmv_memory_reference_string
//     (csr-matrix.cpp:97-143, ell-matrix.cpp:103-143, coo-matrix.cpp:144-185 [atomic form]), generated
//     on the fly per part for ARBITRARY contiguous row (COO: entry) ranges, every reference tagged
//     with the array it touches and, for x and y, with whether the element belongs to the part itself;
//   * a fully associative LRU cache of cache_bytes / line_bytes lines (lru.cpp:31-54), exact, O(1)
//     per reference (hash table + recency list);
//   * one cache shared by all parts with the round-robin interleaving of replacement.cpp:41-95, or one
//     private cache per part (one L2 per GPU);
//   * optional warm-up pass (cache-trace.cpp:128-140).
// With page_bytes > 0 the owner of x_j / y_i is the reference's page rule (thread_of_index,
// util/aligned-allocator.hpp:156-211), which makes the per-thread-per-NUMA-domain numbers of the
// reference reproducible bit for bit (tests/test_cache_model.py pins this against the reference's own
// LRU and against the config-1 known answer of SURVEY section 6).  With page_bytes = 0 the owner is the
// part whose row range holds the index: the row-partitioned multi-GPU mode, where remote x misses are
// exactly the elements a rank must receive.
// "stream_bypass" models the evict-first policy the kernels put on the matrix streams: index and
// value references miss (they are read once) but do not enter the cache.
#include "../../include/spmv_b200.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <exception>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <new>
#include <stdexcept>
#include <vector>

// No exception crosses the C ABI: host-side allocations (std::vector, std::string, new) may throw.
#define SPMV_ABI_CATCH                                                                                  \
    catch (const std::bad_alloc &) { return ::spmvb200::fail(SPMVB200_ERR_NOMEM, "out of host memory"); } \
    catch (const std::exception & e) { return ::spmvb200::fail(SPMVB200_ERR_INVALID, e.what()); }

namespace spmvb200 {
int fail(int code, const std::string & msg);

namespace {

enum Cat { kIndex = 0, kCol, kVal, kXLocal, kXRemote, kYLocal, kYRemote, kCats };

// Exact fully associative LRU over line numbers.
class Lru {
public:
    explicit Lru(int64_t lines) : cap_(std::max<int64_t>(lines, 1))
    {
        size_t b = 16;
        while ((int64_t)b < 2 * cap_) b <<= 1;
        mask_ = b - 1;
        bucket_.assign(b, -1);
        key_.resize((size_t)cap_);
        chain_.resize((size_t)cap_);
        prev_.resize((size_t)cap_);
        next_.resize((size_t)cap_);
    }
    // true = miss (the line is allocated, evicting the least recently used one when full)
    bool access(uint64_t line)
    {
        const size_t h = hash(line);
        for (int32_t n = bucket_[h]; n >= 0; n = chain_[n])
            if (key_[n] == line) {
                touch(n);
                return false;
            }
        int32_t n;
        if (used_ < cap_) {
            n = (int32_t)used_++;
        } else {  // evict the tail
            n = tail_;
            unlink_list(n);
            const size_t hb = hash(key_[n]);
            int32_t * p = &bucket_[hb];
            while (*p != n) p = &chain_[*p];
            *p = chain_[n];
        }
        key_[n] = line;
        chain_[n] = bucket_[h];
        bucket_[h] = n;
        push_front(n);
        return true;
    }

private:
    size_t hash(uint64_t k) const
    {
        k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33;
        return (size_t)k & mask_;
    }
    void unlink_list(int32_t n)
    {
        if (prev_[n] >= 0) next_[prev_[n]] = next_[n]; else head_ = next_[n];
        if (next_[n] >= 0) prev_[next_[n]] = prev_[n]; else tail_ = prev_[n];
    }
    void push_front(int32_t n)
    {
        prev_[n] = -1;
        next_[n] = head_;
        if (head_ >= 0) prev_[head_] = n;
        head_ = n;
        if (tail_ < 0) tail_ = n;
    }
    void touch(int32_t n)
    {
        if (head_ == n) return;
        unlink_list(n);
        push_front(n);
    }
    int64_t cap_, used_ = 0;
    size_t mask_ = 0;
    int32_t head_ = -1, tail_ = -1;
    std::vector<int32_t> bucket_, chain_, prev_, next_;
    std::vector<uint64_t> key_;
};

// Synthetic addresses: every array starts on its own 2^40 boundary, which keeps the reference's
// property that arrays are page aligned (aligned_allocator<T, 4096>) and never share a line.
constexpr uint64_t base_of(int array) { return (uint64_t)(array + 1) << 40; }
enum Array { aIndex = 0, aCol, aVal, aX, aY };

struct Owner {
    int parts = 1;
    int64_t n = 0, page = 0;
    const int64_t * starts = nullptr;  // page == 0: owner = part whose [starts[p], starts[p+1]) holds the index
    int of(int64_t idx) const
    {
        if (page > 0) {  // thread_of_index (aligned-allocator.hpp:156-211) for a page-aligned array of doubles
            const int64_t per = (n + parts - 1) / parts;
            const int64_t page_byte = (idx * 8 / page) * page;
            const int64_t t = per > 0 ? page_byte / (per * 8) : 0;
            return (int)std::min<int64_t>(t, parts - 1);
        }
        if (starts)  // number of part ends <= idx
            return std::min((int)(std::upper_bound(starts + 1, starts + parts + 1, idx) - (starts + 1)), parts - 1);
        const int64_t per = (n + parts - 1) / parts;
        return per > 0 ? (int)std::min<int64_t>(idx / per, parts - 1) : 0;
    }
};

struct Ref {
    uint64_t addr;
    int cat;
};

// One part's reference string, produced on demand.
class Source {
public:
    virtual ~Source() = default;
    virtual bool next(Ref & r) = 0;
    virtual void rewind() = 0;
};

class CsrSource : public Source {
public:
    CsrSource(const int64_t * rp, const int32_t * col, int64_t r0, int64_t r1, int part, const Owner & xo)
        : rp_(rp), col_(col), r0_(r0), r1_(r1), part_(part), xo_(xo) { rewind(); }
    void rewind() override { i_ = r0_; k_ = rp_[r0_]; state_ = 0; first_ = true; }
    bool next(Ref & r) override
    {
        if (first_) {  // &row_ptr[start_row] (csr-matrix.cpp:119-120)
            first_ = false;
            r = {base_of(aIndex) + 4 * (uint64_t)r0_, kIndex};
            return true;
        }
        for (;;) {
            if (i_ >= r1_) return false;
            switch (state_) {
            case 0:  // &row_ptr[i+1]
                state_ = 1;
                k_ = rp_[i_];
                r = {base_of(aIndex) + 4 * (uint64_t)(i_ + 1), kIndex};
                return true;
            case 1:
                if (k_ >= rp_[i_ + 1]) { state_ = 4; continue; }
                state_ = 2;
                r = {base_of(aCol) + 4 * (uint64_t)k_, kCol};
                return true;
            case 2:
                state_ = 3;
                r = {base_of(aVal) + 8 * (uint64_t)k_, kVal};
                return true;
            case 3: {
                const int64_t j = col_[k_++];
                state_ = 1;
                r = {base_of(aX) + 8 * (uint64_t)j, xo_.of(j) == part_ ? kXLocal : kXRemote};
                return true;
            }
            default:  // &y[i]
                state_ = 0;
                r = {base_of(aY) + 8 * (uint64_t)i_++, kYLocal};
                return true;
            }
        }
    }

private:
    const int64_t * rp_;
    const int32_t * col_;
    int64_t r0_, r1_, i_ = 0, k_ = 0;
    int part_, state_ = 0;
    bool first_ = true;
    Owner xo_;
};

// ELLPACK in the reference's ROW-MAJOR layout (ell-matrix.cpp:103-143): k = i*W + l.
class EllSource : public Source {
public:
    EllSource(const int32_t * col, int64_t W, int64_t r0, int64_t r1, int part, const Owner & xo)
        : col_(col), W_(W), r0_(r0), r1_(r1), part_(part), xo_(xo) { rewind(); }
    void rewind() override { i_ = r0_; l_ = 0; state_ = 0; }
    bool next(Ref & r) override
    {
        for (;;) {
            if (i_ >= r1_) return false;
            const int64_t k = i_ * W_ + l_;
            if (l_ >= W_) {
                r = {base_of(aY) + 8 * (uint64_t)i_, kYLocal};
                ++i_; l_ = 0; state_ = 0;
                return true;
            }
            switch (state_) {
            case 0: state_ = 1; r = {base_of(aCol) + 4 * (uint64_t)k, kCol}; return true;
            case 1: state_ = 2; r = {base_of(aVal) + 8 * (uint64_t)k, kVal}; return true;
            default: {
                int64_t j = col_[k];
                if (j == INT32_MAX || j < 0) j = 0;  // skip-padding sentinel: the reference would index x[INT32_MAX]
                state_ = 0; ++l_;
                r = {base_of(aX) + 8 * (uint64_t)j, xo_.of(j) == part_ ? kXLocal : kXRemote};
                return true;
            }
            }
        }
    }

private:
    const int32_t * col_;
    int64_t W_, r0_, r1_, i_ = 0, l_ = 0;
    int part_, state_ = 0;
    Owner xo_;
};

// COO, the atomic form (coo-matrix.cpp:144-185): row, column, value, x[j], y[i] per entry.
class CooSource : public Source {
public:
    CooSource(const int32_t * row, const int32_t * col, int64_t k0, int64_t k1, int part, const Owner & xo, const Owner & yo)
        : row_(row), col_(col), k0_(k0), k1_(k1), part_(part), xo_(xo), yo_(yo) { rewind(); }
    void rewind() override { k_ = k0_; state_ = 0; }
    bool next(Ref & r) override
    {
        if (k_ >= k1_) return false;
        switch (state_) {
        case 0: state_ = 1; r = {base_of(aIndex) + 4 * (uint64_t)k_, kIndex}; return true;
        case 1: state_ = 2; r = {base_of(aCol) + 4 * (uint64_t)k_, kCol}; return true;
        case 2: state_ = 3; r = {base_of(aVal) + 8 * (uint64_t)k_, kVal}; return true;
        case 3: {
            const int64_t j = col_[k_];
            state_ = 4;
            r = {base_of(aX) + 8 * (uint64_t)j, xo_.of(j) == part_ ? kXLocal : kXRemote};
            return true;
        }
        default: {
            const int64_t i = row_[k_++];
            state_ = 0;
            r = {base_of(aY) + 8 * (uint64_t)i, yo_.of(i) == part_ ? kYLocal : kYRemote};
            return true;
        }
        }
    }

private:
    const int32_t * row_;
    const int32_t * col_;
    int64_t k0_, k1_, k_ = 0;
    int part_, state_ = 0;
    Owner xo_, yo_;
};

void count(spmvb200_cache_misses & o, const Ref & r, bool miss)
{
    o.references++;
    if (r.cat == kXLocal || r.cat == kXRemote) o.x_references++;
    if (r.cat == kXRemote) o.x_remote_references++;
    if (!miss) return;
    switch (r.cat) {
    case kIndex: o.misses_index++; break;
    case kCol: o.misses_column_index++; break;
    case kVal: o.misses_value++; break;
    case kXLocal: o.misses_x_local++; break;
    case kXRemote: o.misses_x_remote++; break;
    case kYLocal: o.misses_y_local++; break;
    default: o.misses_y_remote++; break;
    }
}

void simulate(std::vector<std::unique_ptr<Source>> & src, const spmvb200_cache_config & cfg, spmvb200_cache_misses * out)
{
    const int P = (int)src.size();
    const int64_t lines = (cfg.cache_bytes + cfg.line_bytes - 1) / cfg.line_bytes;  // cache-trace.cpp:127
    const uint64_t L = (uint64_t)cfg.line_bytes;
    // stream_bypass: an index/value reference misses once per line (consecutive references to the line a
    // stream is on are served by the load that fetched it) and never enters the cache.
    std::vector<uint64_t> last((size_t)P * 3, ~0ull);
    auto touch = [&](Lru & cache, int p, const Ref & r) -> bool {
        const uint64_t line = r.addr / L;
        if (cfg.stream_bypass && r.cat <= kVal) {
            uint64_t & l = last[(size_t)p * 3 + r.cat];
            const bool miss = l != line;
            l = line;
            return miss;
        }
        return cache.access(line);
    };
    if (cfg.shared) {
        Lru cache(lines);
        for (int pass = cfg.warmup ? 0 : 1; pass < 2; ++pass) {
            for (auto & s : src) s->rewind();
            std::fill(last.begin(), last.end(), ~0ull);
            std::vector<char> live((size_t)P, 1);
            int alive = P;
            while (alive > 0) {  // one reference of every part per step (replacement.cpp:72-82)
                for (int p = 0; p < P; ++p) {
                    if (!live[p]) continue;
                    Ref r;
                    if (!src[p]->next(r)) { live[p] = 0; --alive; continue; }
                    const bool miss = touch(cache, p, r);
                    if (pass == 1) count(out[p], r, miss);
                }
            }
        }
        return;
    }
    // private caches: the parts are independent of each other, one host thread each
    auto run_part = [&](int p) {
        Lru cache(lines);
        for (int pass = cfg.warmup ? 0 : 1; pass < 2; ++pass) {
            src[p]->rewind();
            for (int k = 0; k < 3; ++k) last[(size_t)p * 3 + k] = ~0ull;
            Ref r;
            while (src[p]->next(r)) {
                const bool miss = touch(cache, p, r);
                if (pass == 1) count(out[p], r, miss);
            }
        }
    };
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    if (P == 1 || hw == 1) {
        for (int p = 0; p < P; ++p) run_part(p);
        return;
    }
    std::exception_ptr failure;
    std::mutex failure_mutex;
    for (int first = 0; first < P; first += (int)hw) {  // at most `hw` parts at a time
        std::vector<std::thread> workers;
        for (int p = first; p < std::min<int>(P, first + (int)hw); ++p)
            workers.emplace_back([&, p] {
                try {
                    run_part(p);
                } catch (...) {
                    std::lock_guard<std::mutex> lk(failure_mutex);
                    failure = std::current_exception();
                }
            });
        for (auto & w : workers) w.join();
    }
    if (failure) std::rethrow_exception(failure);
}

int check_config(const spmvb200_cache_config * cfg, const spmvb200_cache_misses * out)
{
    if (!cfg || !out) return fail(SPMVB200_ERR_INVALID, "null argument");
    if (cfg->cache_bytes < 1 || cfg->line_bytes < 1 || cfg->parts < 1 || cfg->page_bytes < 0)
        return fail(SPMVB200_ERR_INVALID, "cache model: cache_bytes, line_bytes and parts must be positive");
    return 0;
}

}  // namespace
}  // namespace spmvb200

using namespace spmvb200;

extern "C" {

int spmvb200_cache_trace_csr(int64_t rows, int64_t columns, const int64_t * row_ptr, const int32_t * column_index,
                             const spmvb200_cache_config * cfg, spmvb200_cache_misses * out)
try {
    if (int rc = check_config(cfg, out)) return rc;
    if (rows < 0 || columns < 0 || !row_ptr || (!column_index && row_ptr[rows] > 0)) return fail(SPMVB200_ERR_INVALID, "bad CSR arrays");
    const int P = cfg->parts;
    std::vector<int64_t> starts((size_t)P + 1);
    for (int p = 0; p <= P; ++p)
        starts[p] = cfg->starts ? cfg->starts[p] : std::min<int64_t>(rows, (int64_t)p * ((rows + P - 1) / P));  // csr-matrix.cpp:106-108
    for (int p = 0; p < P; ++p)
        if (starts[p] < 0 || starts[p] > starts[p + 1] || starts[p + 1] > rows) return fail(SPMVB200_ERR_INVALID, "cache model: bad partition");
    // x_j belongs to the part that owns row j (square matrices: the iterated x <- A x of the row-partitioned
    // mode); for a rectangular matrix the columns are cut by the reference's equal-count rule
    Owner xo{P, columns, cfg->page_bytes, (cfg->page_bytes == 0 && rows == columns) ? starts.data() : nullptr};
    std::memset(out, 0, sizeof(*out) * (size_t)P);
    std::vector<std::unique_ptr<Source>> src;
    for (int p = 0; p < P; ++p) src.emplace_back(new CsrSource(row_ptr, column_index, starts[p], starts[p + 1], p, xo));
    simulate(src, *cfg, out);
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_cache_trace_ell(int64_t rows, int64_t columns, int64_t row_length, const int32_t * column_index_row_major,
                             const spmvb200_cache_config * cfg, spmvb200_cache_misses * out)
try {
    if (int rc = check_config(cfg, out)) return rc;
    if (rows < 0 || columns < 0 || row_length < 0 || (!column_index_row_major && rows * row_length > 0))
        return fail(SPMVB200_ERR_INVALID, "bad ELL arrays");
    const int P = cfg->parts;
    std::vector<int64_t> starts((size_t)P + 1);
    for (int p = 0; p <= P; ++p)
        starts[p] = cfg->starts ? cfg->starts[p] : std::min<int64_t>(rows, (int64_t)p * ((rows + P - 1) / P));  // ell-matrix.cpp:111-113
    for (int p = 0; p < P; ++p)
        if (starts[p] < 0 || starts[p] > starts[p + 1] || starts[p + 1] > rows) return fail(SPMVB200_ERR_INVALID, "cache model: bad partition");
    Owner xo{P, columns, cfg->page_bytes, (cfg->page_bytes == 0 && rows == columns) ? starts.data() : nullptr};
    std::memset(out, 0, sizeof(*out) * (size_t)P);
    std::vector<std::unique_ptr<Source>> src;
    for (int p = 0; p < P; ++p) src.emplace_back(new EllSource(column_index_row_major, row_length, starts[p], starts[p + 1], p, xo));
    simulate(src, *cfg, out);
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_cache_trace_coo(int64_t rows, int64_t columns, int64_t num_entries, const int32_t * row_index,
                             const int32_t * column_index, const spmvb200_cache_config * cfg, spmvb200_cache_misses * out)
try {
    if (int rc = check_config(cfg, out)) return rc;
    if (rows < 0 || columns < 0 || num_entries < 0 || (num_entries > 0 && (!row_index || !column_index)))
        return fail(SPMVB200_ERR_INVALID, "bad COO arrays");
    const int P = cfg->parts;
    std::vector<int64_t> starts((size_t)P + 1);  // ENTRY ranges (coo-matrix.cpp:152-154)
    for (int p = 0; p <= P; ++p)
        starts[p] = cfg->starts ? cfg->starts[p] : std::min<int64_t>(num_entries, (int64_t)p * ((num_entries + P - 1) / P));
    for (int p = 0; p < P; ++p)
        if (starts[p] < 0 || starts[p] > starts[p + 1] || starts[p + 1] > num_entries) return fail(SPMVB200_ERR_INVALID, "cache model: bad partition");
    Owner xo{P, columns, cfg->page_bytes, nullptr}, yo{P, rows, cfg->page_bytes, nullptr};
    std::memset(out, 0, sizeof(*out) * (size_t)P);
    std::vector<std::unique_ptr<Source>> src;
    for (int p = 0; p < P; ++p) src.emplace_back(new CooSource(row_index, column_index, starts[p], starts[p + 1], p, xo, yo));
    simulate(src, *cfg, out);
    return 0;
}
SPMV_ABI_CATCH

}  // extern "C"
