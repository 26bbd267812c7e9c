/*
 * ref_shim.cpp -- C-ABI window onto the UNMODIFIED reference library.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is ours; it is compiled together with
 * the reference's own sources where they lie under /root/reference (see
 * oracle/Makefile, target _ref/libspmvref.so) and only forwards to the
 * reference's public functions:
 *   matrix_market::fromStream / load_matrix / sort_matrix_row_major
 *   {csr,coo,ell,hybrid}_matrix::from_matrix_market, ::spmv, ::spmv_atomic
 *   {csr,ell,coo}_matrix::Matrix::spmv[_atomic]_memory_reference_string,
 *   replacement::LRU, replacement::trace_cache_misses
 * No reference source is copied into this repository.  Used (a) to pin
 * oracle/spmv_oracle.c, (b) to generate tests/golden/ref_vectors.npz, and
 * (c) as bench.py's CPU baseline ("kind": "reference").
 */
#include "matrix/coo-matrix.hpp"
#include "matrix/csr-matrix.hpp"
#include "matrix/ell-matrix.hpp"
#include "matrix/hybrid-matrix.hpp"
#include "matrix/matrix-error.hpp"
#include "matrix/matrix-market.hpp"
#include "matrix/matrix-market-reorder.hpp"
#include "cache-simulation/replacement.hpp"

#include <omp.h>
#include <sched.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

struct RefMatrix {
    std::unique_ptr<matrix_market::Matrix> mm;
    std::string format; /* "", "csr", "coo", "coo-atomic", "ell", "hybrid" */
    csr_matrix::Matrix csr;
    coo_matrix::Matrix coo;
    ell_matrix::Matrix ell;
    hybrid_matrix::Matrix hyb;
    /* Kernel-object state (csr-spmv.hpp:34-38 etc.) */
    std::vector<double, aligned_allocator<double, 4096>> x, y, workspace;
    int workspace_threads = 0;
};

template <typename F>
int guarded(F && f)
{
    try {
        g_err.clear();
        f();
        return 0;
    } catch (std::exception const & e) {
        g_err = e.what();
        return 1;
    }
}

}  // namespace

extern "C" {

const char * ref_last_error() { return g_err.c_str(); }

/* ---- Matrix Market ------------------------------------------------------ */

int ref_mm_from_text(const char * text, size_t len, void ** out)
{
    return guarded([&] {
        std::istringstream s(std::string(text, len));
        auto m = std::make_unique<RefMatrix>();
        m->mm = std::make_unique<matrix_market::Matrix>(matrix_market::fromStream(s));
        *out = m.release();
    });
}

int ref_mm_load(const char * path, void ** out)
{
    return guarded([&] {
        std::ostringstream log;
        auto m = std::make_unique<RefMatrix>();
        m->mm = std::make_unique<matrix_market::Matrix>(matrix_market::load_matrix(path, log, false));
        *out = m.release();
    });
}

/* Public ctor matrix_market::Matrix(Header, Comments, Size, vector<CoordinateEntryReal>)
 * (matrix-market.hpp:81-84); entries 1-based. */
int ref_mm_from_entries(int32_t rows, int32_t cols, int32_t n, const int32_t * i, const int32_t * j,
                        const double * a, void ** out)
{
    return guarded([&] {
        using namespace matrix_market;
        std::vector<CoordinateEntryReal> e((size_t)n);
        for (int32_t k = 0; k < n; k++) e[k] = CoordinateEntryReal{i[k], j[k], a[k]};
        Header h{Object::matrix, Format::coordinate, Field::real, Symmetry::general};
        auto m = std::make_unique<RefMatrix>();
        m->mm = std::make_unique<Matrix>(h, Comments{}, Size{rows, cols, n}, e);
        *out = m.release();
    });
}

void ref_free(void * h) { delete static_cast<RefMatrix *>(h); }

int ref_mm_info(void * h, int32_t * rows, int32_t * cols, int32_t * n, int32_t * field, int32_t * symmetry)
{
    auto * m = static_cast<RefMatrix *>(h);
    *rows = m->mm->rows(); *cols = m->mm->columns(); *n = m->mm->num_entries();
    *field = (int32_t)m->mm->field(); *symmetry = (int32_t)m->mm->symmetry();
    return 0;
}

int ref_mm_entries(void * h, int32_t * i, int32_t * j, double * a)
{
    return guarded([&] {
        auto * m = static_cast<RefMatrix *>(h);
        auto ri = m->mm->row_indices();
        auto ci = m->mm->column_indices();
        auto v = m->mm->values_real();
        std::copy(ri.begin(), ri.end(), i);
        std::copy(ci.begin(), ci.end(), j);
        std::copy(v.begin(), v.end(), a);
    });
}

int32_t ref_mm_max_row_length(void * h) { return static_cast<RefMatrix *>(h)->mm->max_row_length(); }

int ref_mm_sort_row_major(void * h)
{
    return guarded([&] {
        auto * m = static_cast<RefMatrix *>(h);
        m->mm = std::make_unique<matrix_market::Matrix>(matrix_market::sort_matrix_row_major(*m->mm));
    });
}

/* ---- conversions (X_matrix::from_matrix_market) --------------------------- */

int ref_convert(void * h, const char * format, int32_t arg)
{
    return guarded([&] {
        auto * m = static_cast<RefMatrix *>(h);
        std::string f(format);
        std::ostringstream log;
        if (f == "csr") m->csr = csr_matrix::from_matrix_market_row_aligned(*m->mm, arg > 0 ? arg : 1);
        else if (f == "coo" || f == "coo-atomic") m->coo = coo_matrix::from_matrix_market(*m->mm);
        else if (f == "ell") m->ell = ell_matrix::from_matrix_market(*m->mm, arg != 0);
        else if (f == "hybrid") m->hyb = hybrid_matrix::from_matrix_market(*m->mm, arg != 0, log, false);
        else throw matrix::matrix_error("unknown format " + f);
        m->format = f;
    });
}

/* sizes[]: csr {rows, cols, nnz, stored, row_alignment}; coo {rows, cols, nnz};
 * ell {rows, cols, nnz, row_length, skip}; hybrid {rows, cols, nnz, W, n_ell, n_coo, skip}.
 * bytes = Matrix::size() as the reference prints it in "matrix_size". */
int ref_sizes(void * h, int64_t * sizes, int64_t * bytes)
{
    auto * m = static_cast<RefMatrix *>(h);
    if (m->format == "csr") {
        sizes[0] = m->csr.rows; sizes[1] = m->csr.columns; sizes[2] = m->csr.num_entries;
        sizes[3] = (int64_t)m->csr.column_index.size(); sizes[4] = m->csr.row_alignment;
        *bytes = (int64_t)m->csr.size();
    } else if (m->format == "coo" || m->format == "coo-atomic") {
        sizes[0] = m->coo.rows; sizes[1] = m->coo.columns; sizes[2] = m->coo.num_entries;
        *bytes = (int64_t)m->coo.size();
    } else if (m->format == "ell") {
        sizes[0] = m->ell.rows; sizes[1] = m->ell.columns; sizes[2] = m->ell.num_entries;
        sizes[3] = m->ell.row_length; sizes[4] = m->ell.skip_padding;
        *bytes = (int64_t)m->ell.size();
    } else if (m->format == "hybrid") {
        sizes[0] = m->hyb.rows; sizes[1] = m->hyb.columns; sizes[2] = m->hyb.num_entries;
        sizes[3] = m->hyb.ell_row_length; sizes[4] = m->hyb.num_ell_entries;
        sizes[5] = m->hyb.num_coo_entries; sizes[6] = m->hyb.ell_skip_padding;
        *bytes = (int64_t)m->hyb.size();
    } else {
        g_err = "no converted matrix";
        return 1;
    }
    return 0;
}

/* Borrowed pointers to the converted arrays (valid until ref_free). */
int ref_arrays(void * h, const int32_t ** p0, const int32_t ** p1, const double ** v0,
               const int32_t ** p2, const int32_t ** p3, const double ** v1)
{
    auto * m = static_cast<RefMatrix *>(h);
    *p0 = *p1 = *p2 = *p3 = nullptr; *v0 = *v1 = nullptr;
    if (m->format == "csr") {
        *p0 = m->csr.row_ptr.data(); *p1 = m->csr.column_index.data(); *v0 = m->csr.value.data();
    } else if (m->format == "coo" || m->format == "coo-atomic") {
        *p0 = m->coo.row_index.data(); *p1 = m->coo.column_index.data(); *v0 = m->coo.value.data();
    } else if (m->format == "ell") {
        *p1 = m->ell.column_index.data(); *v0 = m->ell.value.data();
    } else if (m->format == "hybrid") {
        *p1 = m->hyb.ell_column_index.data(); *v0 = m->hyb.ell_value.data();
        *p2 = m->hyb.coo_row_index.data(); *p3 = m->hyb.coo_column_index.data();
        *v1 = m->hyb.coo_value.data();
    } else {
        g_err = "no converted matrix";
        return 1;
    }
    return 0;
}

int32_t ref_csr_rows_per_thread(void * h, int t, int T)
{
    return static_cast<RefMatrix *>(h)->csr.spmv_rows_per_thread(t, T);
}
int32_t ref_csr_nonzeros_per_thread(void * h, int t, int T)
{
    return static_cast<RefMatrix *>(h)->csr.spmv_nonzeros_per_thread(t, T);
}

/* ---- SpMV: the Kernel::run bodies (csr-spmv.cpp:64-67, coo-spmv.cpp:76-81,
 * coo-spmv-atomic.cpp:63-68, ell-spmv.cpp:63-66, hybrid-spmv.cpp:78-83),
 * entered by every thread of an omp parallel region (profile-kernel.cpp:227). */

static void run_once(RefMatrix * m, int T)
{
    if (m->format == "csr") csr_matrix::spmv(m->csr, m->x, m->y);
    else if (m->format == "coo") coo_matrix::spmv(T, m->coo, m->x, m->y, m->workspace);
    else if (m->format == "coo-atomic") coo_matrix::spmv_atomic(T, m->coo, m->x, m->y);
    else if (m->format == "ell") ell_matrix::spmv(m->ell, m->x, m->y);
    else if (m->format == "hybrid") hybrid_matrix::spmv(T, m->hyb, m->x, m->y, m->workspace, 0);
}

static void shape(RefMatrix * m, int64_t * rows, int64_t * cols)
{
    if (m->format == "csr") { *rows = m->csr.rows; *cols = m->csr.columns; }
    else if (m->format == "coo" || m->format == "coo-atomic") { *rows = m->coo.rows; *cols = m->coo.columns; }
    else if (m->format == "ell") { *rows = m->ell.rows; *cols = m->ell.columns; }
    else { *rows = m->hyb.rows; *cols = m->hyb.columns; }
}

/* One y += A*x with T threads on zeroed workspace; x has `cols`, y has `rows` entries. */
int ref_spmv(void * h, int T, const double * x, double * y)
{
    return guarded([&] {
        auto * m = static_cast<RefMatrix *>(h);
        int64_t rows, cols;
        shape(m, &rows, &cols);
        m->x.assign(x, x + cols);
        m->y.assign(y, y + rows);
        m->workspace.assign((size_t)T * rows, 0.0);
        omp_set_dynamic(0);
        omp_set_num_threads(T);
        #pragma omp parallel num_threads(T)
        {
            run_once(m, T);
        }
        std::copy(m->y.begin(), m->y.end(), y);
    });
}

/* Timing protocol of profile_kernel_run (profile-kernel.cpp:137-179): inside
 * one parallel region, pin each thread, one warm-up, then `reps` runs each
 * bracketed barrier / steady_clock / barrier.  x = 1, y = 0 as Kernel::init
 * sets them (csr-spmv.cpp:35-36).  ns[] receives the per-run durations. */
int ref_time(void * h, int T, int pin, int reps, double * ns)
{
    return guarded([&] {
        auto * m = static_cast<RefMatrix *>(h);
        int64_t rows, cols;
        shape(m, &rows, &cols);
        m->x.assign((size_t)cols, 1.0);
        m->y.assign((size_t)rows, 0.0);
        if (m->format == "coo" || m->format == "hybrid") m->workspace.assign((size_t)T * rows, 0.0);
        omp_set_dynamic(0);
        omp_set_num_threads(T);
        using clk = std::chrono::steady_clock;
        #pragma omp parallel num_threads(T)
        {
            int t = omp_get_thread_num();
            if (pin) {
                cpu_set_t set;
                CPU_ZERO(&set);
                CPU_SET(t, &set);
                sched_setaffinity(0, sizeof set, &set);
            }
            run_once(m, T); /* warm-up (profile-kernel.cpp:263-264) */
            for (int r = 0; r < reps; r++) {
                clk::time_point t0, t1;
                #pragma omp barrier
                #pragma omp master
                t0 = clk::now();
                #pragma omp barrier
                run_once(m, T);
                #pragma omp barrier
                #pragma omp master
                {
                    t1 = clk::now();
                    ns[r] = (double)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
                }
                #pragma omp barrier
            }
        }
    });
}

/* find_new_order_RCM of the reference (matrix-market-reorder.cpp:60-170); out has `rows` entries.
 * (The reference prints progress lines to stdout.) */
int ref_mm_order_rcm(void * h, int32_t * out)
{
    return guarded([&] {
        auto * m = static_cast<RefMatrix *>(h);
        std::vector<int> order = find_new_order_RCM(*m->mm);
        std::copy(order.begin(), order.end(), out);
    });
}

/* The reference's cache trace for one cache (cache-trace.cpp:92-161): reference strings of all T
 * threads (numa domain of thread t = t), LRU of cache_bytes/line_bytes lines shared by them,
 * optional warm-up pass; misses[t*T + d] = misses of thread t on lines owned by domain d.
 * Formats: csr, ell, coo-atomic. */
int ref_cache_trace(void * h, int T, int64_t cache_bytes, int line_bytes, int warmup, int page_size, uint64_t * misses)
{
    return guarded([&] {
        auto * m = static_cast<RefMatrix *>(h);
        int64_t rows, cols;
        shape(m, &rows, &cols);
        m->x.assign((size_t)cols, 1.0);
        m->y.assign((size_t)rows, 0.0);
        std::vector<int> domains(T);
        for (int t = 0; t < T; t++) domains[t] = t;
        std::vector<replacement::MemoryReferenceString> ws(T);
        for (int t = 0; t < T; t++) {
            if (m->format == "csr") ws[t] = m->csr.spmv_memory_reference_string(m->x, m->y, t, T, domains.data(), page_size);
            else if (m->format == "ell") ws[t] = m->ell.spmv_memory_reference_string(m->x, m->y, t, T, domains.data(), page_size);
            else if (m->format == "coo-atomic") ws[t] = m->coo.spmv_atomic_memory_reference_string(m->x, m->y, t, T, domains.data(), page_size);
            else throw std::runtime_error("ref_cache_trace: format must be csr, ell or coo-atomic");
        }
        int64_t lines = (cache_bytes + line_bytes - 1) / line_bytes;
        replacement::LRU lru(lines, line_bytes);
        if (warmup) replacement::trace_cache_misses(lru, ws, T, false, 0);
        auto r = replacement::trace_cache_misses(lru, ws, T, false, 0);
        for (int t = 0; t < T; t++)
            for (int d = 0; d < T; d++) misses[t * T + d] = r[t][d];
    });
}

}  // extern "C"
