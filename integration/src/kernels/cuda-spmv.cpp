#include "cuda-spmv.hpp"

#include "matrix/matrix-error.hpp"
#include "matrix/matrix-market.hpp"

#ifdef USE_OPENMP
#include <omp.h>
#endif

#include <ostream>
#include <sstream>
#include <string>
#include <system_error>

namespace
{

std::vector<int> numa_domains_of(TraceConfig const & trace_config)
{
    auto const & thread_affinities = trace_config.thread_affinities();
    std::vector<int> numa_domain_affinity(thread_affinities.size(), 0);
    for (size_t i = 0; i < thread_affinities.size(); i++)
        numa_domain_affinity[i] = thread_affinities[i].numa_domain;
    return numa_domain_affinity;
}

int page_size_of()
{
#ifdef HAVE_LIBNUMA
    return numa_pagesize();
#else
    return 4096;
#endif
}

/* The host kernels turn loader and converter failures into kernel_error("<path>: <what>"). */
template <typename F>
void with_kernel_errors(std::string const & matrix_path, F && f)
{
    try {
        f();
    } catch (matrix::matrix_error & e) {
        std::stringstream s;
        s << matrix_path << ": " << e.what();
        throw kernel_error(s.str());
    } catch (std::system_error & e) {
        std::stringstream s;
        s << matrix_path << ": " << e.what();
        throw kernel_error(s.str());
    }
}

std::ostream & print_common(
    std::ostream & o, std::string const & name, std::string const & matrix_path,
    char const * format, long rows, long columns, long nonzeros, std::size_t matrix_size)
{
    return o
        << "{\n"
        << '"' << "name" << '"' << ": " << '"' << name << '"' << ',' << '\n'
        << '"' << "matrix_path" << '"' << ": " << '"' << matrix_path << '"' << ',' << '\n'
        << '"' << "matrix_format" << '"' << ": " << '"' << format << '"' << ',' << '\n'
        << '"' << "rows" << '"' << ": " << rows << ',' << '\n'
        << '"' << "columns" << '"' << ": " << columns << ',' << '\n'
        << '"' << "nonzeros" << '"' << ": " << nonzeros << ',' << '\n'
        << '"' << "matrix_size" << '"' << ": " << matrix_size << ',' << '\n'
        << '"' << "x_size" << '"' << ": " << sizeof(double) * columns << ',' << '\n'
        << '"' << "y_size" << '"' << ": " << sizeof(double) * rows;
}

}

/*
 * cuda_spmv_device
 */

cuda_spmv_device::cuda_spmv_device(std::string const & matrix_path)
    : matrix_path(matrix_path)
    , dA(nullptr)
{
}

cuda_spmv_device::~cuda_spmv_device()
{
    if (dA)
        spmvb200_destroy(dA);
}

void cuda_spmv_device::check(int status) const
{
    if (status != 0) {
        std::stringstream s;
        s << matrix_path << ": " << spmvb200_last_error();
        throw kernel_error(s.str());
    }
}

void cuda_spmv_device::upload_vectors(
    double const * x, std::size_t, double const * y, std::size_t)
{
    check(spmvb200_set_x(dA, x));
    check(spmvb200_set_y(dA, y));
}

/*
 * prepare() and run() are entered by every thread of the OpenMP team of profile_kernel
 * (profile-kernel.cpp:227, 262-264, 159-161).  The GPU is driven by the master; the others wait at the
 * barrier, which also keeps the timed region of profile_kernel_run honest: t1 is taken after the kernel
 * has finished, not after it was launched.  Exceptions must not leave the master construct with the
 * team split, so the status is carried across the barrier.
 */
void cuda_spmv_device::prepare_device()
{
    static int status;
    #pragma omp master
    status = spmvb200_prepare(dA);
    #pragma omp barrier
    check(status);
}

void cuda_spmv_device::run_device()
{
    static int status;
    #pragma omp master
    {
        status = spmvb200_spmv(dA);
        if (status == 0)
            status = spmvb200_sync(dA);
    }
    #pragma omp barrier
    check(status);
}

std::vector<double> cuda_spmv_device::device_y() const
{
    spmvb200_info info;
    check(spmvb200_matrix_info(dA, &info));
    std::vector<double> y(info.rows);
    if (info.rows > 0)
        check(spmvb200_get_y(dA, y.data()));
    return y;
}

std::ostream & cuda_spmv_device::print_device(std::ostream & o) const
{
    spmvb200_info info;
    if (!dA || spmvb200_matrix_info(dA, &info) != 0)
        return o;
    return o << ',' << '\n'
        << '"' << "device_kernel" << '"' << ": " << '"' << spmvb200_kernel_name(dA) << '"' << ',' << '\n'
        << '"' << "device_bytes" << '"' << ": " << info.device_bytes;
}

/*
 * CSR
 */

cuda_csr_spmv_kernel::cuda_csr_spmv_kernel(std::string const & matrix_path)
    : Kernel()
    , cuda_spmv_device(matrix_path)
{
}

void cuda_csr_spmv_kernel::init(
    TraceConfig const & trace_config,
    std::ostream & o,
    bool verbose)
{
    with_kernel_errors(matrix_path, [&] {
        matrix_market::Matrix mm =
            matrix_market::load_matrix(matrix_path, o, verbose);
        A = csr_matrix::from_matrix_market(mm);
        x = csr_matrix::value_array_type(A.columns, 1.0);
        y = csr_matrix::value_array_type(A.rows, 0.0);
    });
    check(spmvb200_csr_create(
              A.rows, A.columns, A.num_entries,
              A.row_ptr.data(), A.column_index.data(), A.value.data(), &dA));
    upload_vectors(x.data(), x.size(), y.data(), y.size());
}

void cuda_csr_spmv_kernel::prepare(TraceConfig const &)
{
    prepare_device();
}

void cuda_csr_spmv_kernel::run(TraceConfig const &)
{
    run_device();
}

replacement::MemoryReferenceString cuda_csr_spmv_kernel::memory_reference_string(
    TraceConfig const & trace_config,
    int thread,
    int num_threads) const
{
    return A.spmv_memory_reference_string(
        x, y, thread, num_threads,
        numa_domains_of(trace_config).data(), page_size_of());
}

std::string cuda_csr_spmv_kernel::name() const
{
    return "cuda-csr-spmv";
}

std::ostream & cuda_csr_spmv_kernel::print(std::ostream & o) const
{
    print_common(o, name(), matrix_path, "csr", A.rows, A.columns, A.num_entries, A.size());
    return print_device(o) << "\n}";
}

/*
 * COO (segmented reduction on row-sorted entries / file order with reductions into y)
 */

cuda_coo_spmv_kernel::cuda_coo_spmv_kernel(std::string const & matrix_path, bool atomic)
    : Kernel()
    , cuda_spmv_device(matrix_path)
    , atomic(atomic)
{
}

void cuda_coo_spmv_kernel::init(
    TraceConfig const & trace_config,
    std::ostream & o,
    bool verbose)
{
    auto const & thread_affinities = trace_config.thread_affinities();
    int num_threads = thread_affinities.size();
    with_kernel_errors(matrix_path, [&] {
        matrix_market::Matrix mm =
            matrix_market::load_matrix(matrix_path, o, verbose);
        A = coo_matrix::from_matrix_market(mm);
        x = coo_matrix::value_array_type(A.columns, 1.0);
        y = coo_matrix::value_array_type(A.rows, 0.0);
        if (!atomic) {
            /* never touched on the device; the cache model of coo-spmv references it */
            size_t workspace_size;
            if (__builtin_mul_overflow(num_threads, A.rows, &workspace_size)) {
                throw matrix::matrix_error(
                    "Failed to compute COO SpMV: "
                    "Integer overflow when computing workspace size");
            }
            workspace = coo_matrix::value_array_type(workspace_size, 0.0);
        }
    });
    check(spmvb200_coo_create(
              A.rows, A.columns, A.num_entries,
              A.row_index.data(), A.column_index.data(), A.value.data(),
              atomic ? SPMVB200_COO_ATOMIC : SPMVB200_COO_SEGMENTED, &dA));
    upload_vectors(x.data(), x.size(), y.data(), y.size());
}

void cuda_coo_spmv_kernel::prepare(TraceConfig const &)
{
    prepare_device();
}

void cuda_coo_spmv_kernel::run(TraceConfig const &)
{
    run_device();
}

replacement::MemoryReferenceString cuda_coo_spmv_kernel::memory_reference_string(
    TraceConfig const & trace_config,
    int thread,
    int num_threads) const
{
    if (atomic) {
        return A.spmv_atomic_memory_reference_string(
            x, y, thread, num_threads,
            numa_domains_of(trace_config).data(), page_size_of());
    }
    return A.spmv_memory_reference_string(
        x, y, workspace, thread, num_threads,
        numa_domains_of(trace_config).data(), page_size_of());
}

std::string cuda_coo_spmv_kernel::name() const
{
    return atomic ? "cuda-coo-spmv-atomic" : "cuda-coo-spmv";
}

std::ostream & cuda_coo_spmv_kernel::print(std::ostream & o) const
{
    print_common(o, name(), matrix_path, "coo", A.rows, A.columns, A.num_entries, A.size());
    return print_device(o) << "\n}";
}

/*
 * ELLPACK
 */

cuda_ell_spmv_kernel::cuda_ell_spmv_kernel(std::string const & matrix_path)
    : Kernel()
    , cuda_spmv_device(matrix_path)
{
}

void cuda_ell_spmv_kernel::init(
    TraceConfig const & trace_config,
    std::ostream & o,
    bool verbose)
{
    with_kernel_errors(matrix_path, [&] {
        matrix_market::Matrix mm =
            matrix_market::load_matrix(matrix_path, o, verbose);
        A = ell_matrix::from_matrix_market(mm);
        x = ell_matrix::value_array_type(A.columns, 1.0);
        y = ell_matrix::value_array_type(A.rows, 0.0);
    });
    check(spmvb200_ell_create(
              A.rows, A.columns, A.num_entries, A.row_length,
              A.column_index.data(), A.value.data(), A.skip_padding ? 1 : 0, &dA));
    upload_vectors(x.data(), x.size(), y.data(), y.size());
}

void cuda_ell_spmv_kernel::prepare(TraceConfig const &)
{
    prepare_device();
}

void cuda_ell_spmv_kernel::run(TraceConfig const &)
{
    run_device();
}

replacement::MemoryReferenceString cuda_ell_spmv_kernel::memory_reference_string(
    TraceConfig const & trace_config,
    int thread,
    int num_threads) const
{
    return A.spmv_memory_reference_string(
        x, y, thread, num_threads,
        numa_domains_of(trace_config).data(), page_size_of());
}

std::string cuda_ell_spmv_kernel::name() const
{
    return "cuda-ell-spmv";
}

std::ostream & cuda_ell_spmv_kernel::print(std::ostream & o) const
{
    print_common(o, name(), matrix_path, "ell", A.rows, A.columns, A.num_entries, A.size());
    return print_device(o) << "\n}";
}

/*
 * Hybrid ELL + COO
 */

cuda_hybrid_spmv_kernel::cuda_hybrid_spmv_kernel(std::string const & matrix_path)
    : Kernel()
    , cuda_spmv_device(matrix_path)
{
}

void cuda_hybrid_spmv_kernel::init(
    TraceConfig const & trace_config,
    std::ostream & o,
    bool verbose)
{
    auto const & thread_affinities = trace_config.thread_affinities();
    int num_threads = thread_affinities.size();
    with_kernel_errors(matrix_path, [&] {
        matrix_market::Matrix mm =
            matrix_market::load_matrix(matrix_path, o, verbose);
        A = hybrid_matrix::from_matrix_market(mm, false, o, verbose);
        x = hybrid_matrix::value_array_type(A.columns, 1.0);
        y = hybrid_matrix::value_array_type(A.rows, 0.0);
        size_t workspace_size;
        if (__builtin_mul_overflow(num_threads, A.rows, &workspace_size)) {
            throw matrix::matrix_error(
                "Failed to compute hybrid SpMV: "
                "Integer overflow when computing workspace size");
        }
        workspace = hybrid_matrix::value_array_type(workspace_size, 0.0);
    });
    check(spmvb200_hyb_create(
              A.rows, A.columns, A.num_entries,
              A.ell_row_length, A.ell_column_index.data(), A.ell_value.data(),
              A.ell_skip_padding ? 1 : 0,
              A.num_coo_entries, A.coo_row_index.data(), A.coo_column_index.data(),
              A.coo_value.data(), &dA));
    upload_vectors(x.data(), x.size(), y.data(), y.size());
}

void cuda_hybrid_spmv_kernel::prepare(TraceConfig const &)
{
    prepare_device();
}

void cuda_hybrid_spmv_kernel::run(TraceConfig const &)
{
    run_device();
}

replacement::MemoryReferenceString cuda_hybrid_spmv_kernel::memory_reference_string(
    TraceConfig const & trace_config,
    int thread,
    int num_threads) const
{
    return A.spmv_memory_reference_string(
        x, y, workspace, thread, num_threads,
        numa_domains_of(trace_config).data(), page_size_of());
}

std::string cuda_hybrid_spmv_kernel::name() const
{
    return "cuda-hybrid-spmv";
}

std::ostream & cuda_hybrid_spmv_kernel::print(std::ostream & o) const
{
    /* (hybrid-spmv.cpp:124 emits a stray ',' line after matrix_size; this is the same object as valid JSON) */
    print_common(o, name(), matrix_path, "hybrid", A.rows, A.columns, A.num_entries, A.size());
    o << ',' << '\n'
      << '"' << "ell_row_length" << '"' << ": " << A.ell_row_length << ',' << '\n'
      << '"' << "num_ell_entries" << '"' << ": " << A.num_ell_entries << ',' << '\n'
      << '"' << "num_coo_entries" << '"' << ": " << A.num_coo_entries;
    return print_device(o) << "\n}";
}

/*
 * Row-partitioned CSR over several GPUs
 */

cuda_csr_dist_spmv_kernel::cuda_csr_dist_spmv_kernel(std::string const & matrix_path)
    : Kernel()
    , matrix_path(matrix_path)
{
}

cuda_csr_dist_spmv_kernel::~cuda_csr_dist_spmv_kernel()
{
    for (auto d : ranks)
        if (d) spmvb200_dist_destroy(d);
    for (auto c : comms)
        if (c) spmvb200_comm_destroy(c);
}

void cuda_csr_dist_spmv_kernel::init(
    TraceConfig const & trace_config,
    std::ostream & o,
    bool verbose)
{
    auto fail = [&](int status) {
        if (status != 0) {
            std::stringstream s;
            s << matrix_path << ": " << spmvb200_last_error();
            throw kernel_error(s.str());
        }
    };
    int num_ranks = trace_config.thread_affinities().size();
    with_kernel_errors(matrix_path, [&] {
        matrix_market::Matrix mm =
            matrix_market::load_matrix(matrix_path, o, verbose);
        A = csr_matrix::from_matrix_market(mm);
        if (A.rows != A.columns)
            throw matrix::matrix_error("the iteration x <- A*x needs a square matrix");
        x = csr_matrix::value_array_type(A.columns, 1.0);
        y = csr_matrix::value_array_type(A.rows, 0.0);
    });

    /* rank t owns the rows of thread t (csr-matrix.cpp:77-83) */
    starts.assign(num_ranks + 1, 0);
    for (int t = 0; t < num_ranks; t++)
        starts[t + 1] = starts[t] + A.spmv_rows_per_thread(t, num_ranks);

    comms.assign(num_ranks, nullptr);
    ranks.assign(num_ranks, nullptr);
    fail(spmvb200_comm_create_local(num_ranks, nullptr, comms.data()));
    for (int t = 0; t < num_ranks; t++) {
        int device = 0;
        fail(spmvb200_comm_rank(comms[t], nullptr, nullptr, &device));
        fail(spmvb200_set_device(device));
        int64_t const b = starts[t], e = starts[t + 1];
        std::vector<int64_t> row_ptr(e - b + 1);
        for (int64_t r = b; r <= e; r++)
            row_ptr[r - b] = A.row_ptr[r] - A.row_ptr[b];
        spmvb200_matrix_t local = nullptr;
        fail(spmvb200_csr_create64(
                 e - b, A.columns, A.row_ptr[e] - A.row_ptr[b], row_ptr.data(),
                 A.column_index.data() + A.row_ptr[b], A.value.data() + A.row_ptr[b], &local));
        fail(spmvb200_dist_create(
                 comms[t], local, starts.data(), SPMVB200_EXCHANGE_AUTO, SPMVB200_CSR,
                 SPMVB200_DIST_CONSUME_LOCAL, &ranks[t]));
        fail(spmvb200_dist_set_x(ranks[t], x.data() + b));
    }
}

void cuda_csr_dist_spmv_kernel::prepare(TraceConfig const &)
{
    /* the executors were prepared when they were created */
    #pragma omp barrier
}

/*
 * Thread t issues step k of rank t; the barrier separates step k of all ranks from step k+1 of any, which is
 * what the in-process exchange asks for, and thread t then waits for its own GPU.
 */
void cuda_csr_dist_spmv_kernel::run(TraceConfig const &)
{
#ifdef USE_OPENMP
    int const thread = omp_get_thread_num();
#else
    int const thread = 0;
#endif
    static int status;
    #pragma omp single
    status = 0;
    int mine = 0;
    if ((size_t) thread < ranks.size())
        mine = spmvb200_dist_step(ranks[thread], 1.0);
    #pragma omp barrier
    if (mine == 0 && (size_t) thread < ranks.size())
        mine = spmvb200_dist_sync(ranks[thread]);
    if (mine != 0) {
        #pragma omp atomic write
        status = mine;
    }
    #pragma omp barrier
    if (status != 0) {
        std::stringstream s;
        s << matrix_path << ": " << spmvb200_last_error();
        throw kernel_error(s.str());
    }
}

replacement::MemoryReferenceString cuda_csr_dist_spmv_kernel::memory_reference_string(
    TraceConfig const & trace_config,
    int thread,
    int num_threads) const
{
    return A.spmv_memory_reference_string(
        x, y, thread, num_threads,
        numa_domains_of(trace_config).data(), page_size_of());
}

std::string cuda_csr_dist_spmv_kernel::name() const
{
    return "cuda-csr-dist-spmv";
}

std::vector<double> cuda_csr_dist_spmv_kernel::gather_x() const
{
    std::vector<double> out(A.rows);
    for (size_t t = 0; t < ranks.size(); t++) {
        if (spmvb200_dist_get_x(ranks[t], out.data() + starts[t]) != 0)
            throw kernel_error(matrix_path + ": " + spmvb200_last_error());
    }
    return out;
}

std::ostream & cuda_csr_dist_spmv_kernel::print(std::ostream & o) const
{
    print_common(o, name(), matrix_path, "csr", A.rows, A.columns, A.num_entries, A.size());
    o << ',' << '\n' << '"' << "ranks" << '"' << ": [";
    for (size_t t = 0; t < ranks.size(); t++) {
        spmvb200_dist_info_t info;
        if (spmvb200_dist_info(ranks[t], &info) != 0)
            continue;
        o << (t ? "," : "") << '\n'
          << "{" << '"' << "rows" << '"' << ": " << info.rows << ", "
          << '"' << "nonzeros" << '"' << ": " << info.num_entries << ", "
          << '"' << "recv_bytes_per_step" << '"' << ": " << info.recv_bytes_per_step << ", "
          << '"' << "interior_rows" << '"' << ": " << info.interior_rows << "}";
    }
    return o << '\n' << "]" << "\n}";
}
