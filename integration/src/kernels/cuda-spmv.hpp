#ifndef CUDA_SPMV_HPP
#define CUDA_SPMV_HPP

/*
 * Kernel plugins that run the SpMV of every format on an NVIDIA B200 through libspmvb200.so
 * (C ABI: include/spmv_b200.h of the spmv-cache-trace_b200 repository).
 *
 * Each class is the GPU twin of the host kernel of the same format (csr-spmv.hpp, coo-spmv.hpp,
 * coo-spmv-atomic.hpp, ell-spmv.hpp, hybrid-spmv.hpp): it owns the same host matrix and vectors, so
 * `memory_reference_string` -- and with it the whole cache-trace mode -- behaves exactly like the host
 * kernel's, plus a device copy that `run` multiplies with.  Selected with
 * `--spmv-format cuda-csr | cuda-coo | cuda-coo-atomic | cuda-ell | cuda-hybrid | cuda-csr-dist`.
 */

#include "kernel.hpp"
#include "trace-config.hpp"
#include "cache-simulation/replacement.hpp"
#include "matrix/coo-matrix.hpp"
#include "matrix/csr-matrix.hpp"
#include "matrix/ell-matrix.hpp"
#include "matrix/hybrid-matrix.hpp"

#include <spmv_b200.h>

#include <iosfwd>
#include <string>
#include <vector>

/* What the five plugins share: the device handle and the transfers of x and y. */
class cuda_spmv_device
{
public:
    cuda_spmv_device(std::string const & matrix_path);
    ~cuda_spmv_device();
    cuda_spmv_device(cuda_spmv_device const &) = delete;
    cuda_spmv_device & operator=(cuda_spmv_device const &) = delete;

protected:
    void check(int status) const;            // non-zero status -> kernel_error("<path>: <text>")
    void upload_vectors(double const * x, std::size_t columns, double const * y, std::size_t rows);
    void prepare_device();                   // called by every thread of the team; the master does the work
    void run_device();                       // likewise: one y += A*x on the GPU, then the team meets
    std::ostream & print_device(std::ostream & o) const;

    std::string matrix_path;
    spmvb200_matrix_t dA;

public:
    /* y as the device holds it after the runs so far (the host kernels expose y the same way: a member). */
    std::vector<double> device_y() const;
};

class cuda_csr_spmv_kernel : public Kernel, public cuda_spmv_device
{
public:
    cuda_csr_spmv_kernel(std::string const & matrix_path);
    void init(TraceConfig const & trace_config, std::ostream & o, bool verbose) override;
    void prepare(TraceConfig const & trace_config) override;
    void run(TraceConfig const & trace_config) override;
    replacement::MemoryReferenceString memory_reference_string(
        TraceConfig const & trace_config, int thread, int num_threads) const override;
    std::string name() const override;
    std::ostream & print(std::ostream & o) const override;

    csr_matrix::Matrix A;
    csr_matrix::value_array_type x;
    csr_matrix::value_array_type y;
};

class cuda_coo_spmv_kernel : public Kernel, public cuda_spmv_device
{
public:
    cuda_coo_spmv_kernel(std::string const & matrix_path, bool atomic);
    void init(TraceConfig const & trace_config, std::ostream & o, bool verbose) override;
    void prepare(TraceConfig const & trace_config) override;
    void run(TraceConfig const & trace_config) override;
    replacement::MemoryReferenceString memory_reference_string(
        TraceConfig const & trace_config, int thread, int num_threads) const override;
    std::string name() const override;
    std::ostream & print(std::ostream & o) const override;

    bool atomic;
    coo_matrix::Matrix A;
    coo_matrix::value_array_type x;
    coo_matrix::value_array_type y;
    coo_matrix::value_array_type workspace;  // host only: the cache model of the workspace variant references it
};

class cuda_ell_spmv_kernel : public Kernel, public cuda_spmv_device
{
public:
    cuda_ell_spmv_kernel(std::string const & matrix_path);
    void init(TraceConfig const & trace_config, std::ostream & o, bool verbose) override;
    void prepare(TraceConfig const & trace_config) override;
    void run(TraceConfig const & trace_config) override;
    replacement::MemoryReferenceString memory_reference_string(
        TraceConfig const & trace_config, int thread, int num_threads) const override;
    std::string name() const override;
    std::ostream & print(std::ostream & o) const override;

    ell_matrix::Matrix A;
    ell_matrix::value_array_type x;
    ell_matrix::value_array_type y;
};

class cuda_hybrid_spmv_kernel : public Kernel, public cuda_spmv_device
{
public:
    cuda_hybrid_spmv_kernel(std::string const & matrix_path);
    void init(TraceConfig const & trace_config, std::ostream & o, bool verbose) override;
    void prepare(TraceConfig const & trace_config) override;
    void run(TraceConfig const & trace_config) override;
    replacement::MemoryReferenceString memory_reference_string(
        TraceConfig const & trace_config, int thread, int num_threads) const override;
    std::string name() const override;
    std::ostream & print(std::ostream & o) const override;

    hybrid_matrix::Matrix A;
    hybrid_matrix::value_array_type x;
    hybrid_matrix::value_array_type y;
    hybrid_matrix::value_array_type workspace;
};

/*
 * Row-partitioned CSR on several GPUs: thread t of the team drives rank t, which owns the rows the
 * reference gives thread t (csr_matrix::Matrix::spmv_rows_per_thread, csr-matrix.cpp:77-83) on GPU
 * t mod #GPUs.  One `run` is one step x <- A*x of the iteration, including the exchange of x between
 * the ranks (direct peer copies of exactly the columns each rank's rows reference).
 */
class cuda_csr_dist_spmv_kernel : public Kernel
{
public:
    cuda_csr_dist_spmv_kernel(std::string const & matrix_path);
    ~cuda_csr_dist_spmv_kernel();
    void init(TraceConfig const & trace_config, std::ostream & o, bool verbose) override;
    void prepare(TraceConfig const & trace_config) override;
    void run(TraceConfig const & trace_config) override;
    replacement::MemoryReferenceString memory_reference_string(
        TraceConfig const & trace_config, int thread, int num_threads) const override;
    std::string name() const override;
    std::ostream & print(std::ostream & o) const override;

    /* the current x, gathered from the ranks (for checks) */
    std::vector<double> gather_x() const;

    std::string matrix_path;
    csr_matrix::Matrix A;
    csr_matrix::value_array_type x;
    csr_matrix::value_array_type y;
    std::vector<spmvb200_comm_t> comms;
    std::vector<spmvb200_dist_t> ranks;
    std::vector<int64_t> starts;
};

#endif
