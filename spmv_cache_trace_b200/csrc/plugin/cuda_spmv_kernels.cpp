// cuda_spmv_kernels.cpp -- see cuda_spmv_kernels.hpp.
//
// init()    = load_matrix + from_matrix_market + x = 1, y = 0 (csr-spmv.cpp:26-46), with the
//             conversion done by the device builders; matrix errors become
//             kernel_error("<path>: <what>") exactly like csr-spmv.cpp:37-45.
// prepare() = the reference moves pages next to the threads that use them (distribute_pages,
//             csr-spmv.cpp:48-62); here the data is already in HBM, so only the device is selected.
// run()     = y += A*x.  The reference's kernels are orphaned OpenMP loops entered by every thread;
//             a GPU launch must happen once, so the master thread launches and waits and the team
//             meets at a barrier -- the surrounding barriers of profile_kernel_run
//             (profile-kernel.cpp:159-161) then bracket the device work like they bracket the loops.
#include "cuda_spmv_kernels.hpp"

#include <ostream>
#include <sstream>

static void check(int rc, std::string const & matrix_path)
{
    if (rc != 0) {
        std::stringstream s;
        s << matrix_path << ": " << spmvb200_last_error();
        throw kernel_error(s.str());
    }
}

cuda_spmv_kernel::cuda_spmv_kernel(cuda_format format, std::string const & matrix_path)
    : Kernel(), format(format), matrix_path(matrix_path), A(nullptr)
{
}

cuda_spmv_kernel::~cuda_spmv_kernel()
{
    if (A) spmvb200_destroy(A);
}

void cuda_spmv_kernel::init(TraceConfig const &, std::ostream & o, bool verbose)
{
    if (verbose) o << "Loading matrix from " << matrix_path << '\n';
    spmvb200_mm_t mm = nullptr;
    check(spmvb200_mm_load(matrix_path.c_str(), &mm), matrix_path);
    int rc = 0;
    switch (format) {
    case cuda_format::csr: rc = spmvb200_csr_from_mm(mm, 1, &A); break;
    case cuda_format::coo: rc = spmvb200_coo_from_mm(mm, SPMVB200_COO_SEGMENTED, &A); break;
    case cuda_format::coo_atomic: rc = spmvb200_coo_from_mm(mm, SPMVB200_COO_ATOMIC, &A); break;
    case cuda_format::ell: rc = spmvb200_ell_from_mm(mm, 0, &A); break;
    case cuda_format::hybrid:
        if (verbose) o << "Converting matrix to hybrid format" << std::endl;
        rc = spmvb200_hyb_from_mm(mm, 0, &A);
        break;
    }
    spmvb200_mm_free(mm);
    check(rc, matrix_path);
}

void cuda_spmv_kernel::prepare(TraceConfig const &)
{
    int rc = 0;
#pragma omp master
    {
        // nothing to migrate (the matrix, x and y have been resident in HBM since init()); build the
        // kernel's launch metadata now so the first timed run() does not pay for it
        rc = spmvb200_prepare(A);
    }
#pragma omp barrier
    if (rc != 0) check(rc, matrix_path);
}

void cuda_spmv_kernel::run(TraceConfig const &)
{
    int rc = 0;
#pragma omp master
    {
        rc = spmvb200_spmv(A);
        if (rc == 0) rc = spmvb200_sync(A);
    }
#pragma omp barrier
    if (rc != 0) check(rc, matrix_path);
}

void cuda_spmv_kernel::result(std::vector<double> & y) const
{
    spmvb200_info info;
    check(spmvb200_matrix_info(A, &info), matrix_path);
    y.resize((size_t)info.rows);
    if (info.rows > 0) check(spmvb200_get_y(A, y.data()), matrix_path);
}

std::string cuda_spmv_kernel::name() const
{
    switch (format) {
    case cuda_format::csr: return "cuda-csr-spmv";
    case cuda_format::coo: return "cuda-coo-spmv";
    case cuda_format::coo_atomic: return "cuda-coo-spmv-atomic";
    case cuda_format::ell: return "cuda-ell-spmv";
    default: return "cuda-hybrid-spmv";
    }
}

// Same keys, same order as the reference (csr-spmv.cpp:97-112; hybrid-spmv.cpp:113-131 minus
// its stray comma, which makes the reference's hybrid output invalid JSON).
std::ostream & cuda_spmv_kernel::print(std::ostream & o) const
{
    spmvb200_info info{};
    if (A) spmvb200_matrix_info(A, &info);
    const char * fmt = format == cuda_format::csr ? "csr" : format == cuda_format::ell ? "ell"
                     : format == cuda_format::hybrid ? "hybrid" : "coo";
    o << "{\n"
      << "\"name\": \"" << name() << "\",\n"
      << "\"matrix_path\": \"" << matrix_path << "\",\n"
      << "\"matrix_format\": \"" << fmt << "\",\n"
      << "\"rows\": " << info.rows << ",\n"
      << "\"columns\": " << info.columns << ",\n"
      << "\"nonzeros\": " << info.num_entries << ",\n"
      << "\"matrix_size\": " << info.matrix_size << ",\n"
      << "\"x_size\": " << info.x_size << ",\n"
      << "\"y_size\": " << info.y_size;
    if (format == cuda_format::hybrid)
        o << ",\n\"ell_row_length\": " << info.ell_row_length << ",\n\"num_ell_entries\": " << info.num_ell_entries
          << ",\n\"num_coo_entries\": " << info.num_coo_entries;
    return o << "\n}";
}

std::unique_ptr<Kernel> make_cuda_kernel(std::string const & f, std::string const & path)
{
    if (f == "cuda-csr") return std::make_unique<cuda_csr_spmv_kernel>(path);
    if (f == "cuda-coo") return std::make_unique<cuda_coo_spmv_kernel>(path);
    if (f == "cuda-coo-atomic") return std::make_unique<cuda_coo_spmv_atomic_kernel>(path);
    if (f == "cuda-ell") return std::make_unique<cuda_ell_spmv_kernel>(path);
    if (f == "cuda-hybrid") return std::make_unique<cuda_hybrid_spmv_kernel>(path);
    return nullptr;
}
