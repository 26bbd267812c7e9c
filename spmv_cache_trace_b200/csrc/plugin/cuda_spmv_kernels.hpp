// cuda_spmv_kernels.hpp -- GPU counterparts of the reference's SpMV kernel objects
// (src/kernels/{csr,coo,coo-atomic,ell,hybrid}-spmv.{hpp,cpp}), on top of the C ABI.
#pragma once

#include "kernel.hpp"

#include "../../../include/spmv_b200.h"

#include <memory>

enum class cuda_format { csr, coo, coo_atomic, ell, hybrid };

class cuda_spmv_kernel : public Kernel
{
public:
    cuda_spmv_kernel(cuda_format format, std::string const & matrix_path);
    ~cuda_spmv_kernel() override;
    cuda_spmv_kernel(cuda_spmv_kernel const &) = delete;
    cuda_spmv_kernel & operator=(cuda_spmv_kernel const &) = delete;

    void init(TraceConfig const & trace_config, std::ostream & o, bool verbose) override;
    void prepare(TraceConfig const & trace_config) override;
    void run(TraceConfig const & trace_config) override;
    std::string name() const override;
    std::ostream & print(std::ostream & o) const override;

    // beyond the reference interface: what the GPU-aware driver needs
    spmvb200_matrix_t handle() const { return A; }
    void result(std::vector<double> & y) const;  // copy y back (the reference never prints it)

private:
    cuda_format format;
    std::string matrix_path;
    spmvb200_matrix_t A;
};

// One class per format, like the reference, so the factory switch in main.cpp reads the same.
struct cuda_csr_spmv_kernel : cuda_spmv_kernel { explicit cuda_csr_spmv_kernel(std::string const & p) : cuda_spmv_kernel(cuda_format::csr, p) {} };
struct cuda_coo_spmv_kernel : cuda_spmv_kernel { explicit cuda_coo_spmv_kernel(std::string const & p) : cuda_spmv_kernel(cuda_format::coo, p) {} };
struct cuda_coo_spmv_atomic_kernel : cuda_spmv_kernel { explicit cuda_coo_spmv_atomic_kernel(std::string const & p) : cuda_spmv_kernel(cuda_format::coo_atomic, p) {} };
struct cuda_ell_spmv_kernel : cuda_spmv_kernel { explicit cuda_ell_spmv_kernel(std::string const & p) : cuda_spmv_kernel(cuda_format::ell, p) {} };
struct cuda_hybrid_spmv_kernel : cuda_spmv_kernel { explicit cuda_hybrid_spmv_kernel(std::string const & p) : cuda_spmv_kernel(cuda_format::hybrid, p) {} };

std::unique_ptr<Kernel> make_cuda_kernel(std::string const & spmv_format, std::string const & matrix_path);
