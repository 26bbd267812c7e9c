// harness.cpp -- drives the CUDA Kernel plugins exactly as profile_kernel does (every thread of an OpenMP team
// enters prepare() and run(); profile-kernel.cpp:227, 262-264, 159-161) and compares what the GPU computed with the
// HOST kernel of the same format, linked from the same (patched) reference tree.
//
//   harness <matrix.mtx> <trace-config.json> <runs>
//
// Prints one JSON object per format: name, runs, max |y_gpu - y_host| / bound, ok.  Exit code 0 iff every format
// (and the row-partitioned iteration) is within the per-row tolerance 1e-12 * sum_j |a_ij x_j| * runs.
#include "kernels.hpp"
#include "matrix/csr-matrix.hpp"
#include "matrix/matrix-market.hpp"
#include "trace-config.hpp"

#include <omp.h>

#include <cmath>
#include <cstdio>
#include <exception>
#include <iostream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

template <typename K>
static void drive(K & kernel, TraceConfig const & cfg, int runs)
{
    int const T = cfg.thread_affinities().size();
    std::exception_ptr error = nullptr;
    std::mutex m;
    omp_set_num_threads(T);
    #pragma omp parallel
    {
        try {
            kernel.prepare(cfg);
            for (int r = 0; r < runs; r++) {
                #pragma omp barrier
                kernel.run(cfg);
                #pragma omp barrier
            }
        } catch (std::exception const &) {
            std::lock_guard<std::mutex> l(m);
            if (!error) error = std::current_exception();
        }
    }
    if (error) std::rethrow_exception(error);
}

struct Verdict { double worst; bool ok; };

static Verdict compare(std::vector<double> const & got, std::vector<double> const & want, std::vector<double> const & bound, double factor)
{
    Verdict v{0.0, got.size() == want.size()};
    for (size_t i = 0; v.ok && i < got.size(); i++) {
        double const err = std::fabs(got[i] - want[i]);
        double const lim = 1e-12 * factor * bound[i];
        if (bound[i] > 0) v.worst = std::max(v.worst, err / bound[i]);
        if (!(err <= lim)) v.ok = false;
    }
    return v;
}

int main(int argc, char ** argv)
{
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s matrix.mtx trace-config.json runs\n", argv[0]);
        return 2;
    }
    std::string const path = argv[1];
    int const runs = std::atoi(argv[3]);
    int failures = 0;
    try {
        TraceConfig cfg = read_trace_config(argv[2]);
        std::ostringstream log;

        // the reference's own CSR product with x = 1 is the yardstick for every format: y = runs * A * 1
        matrix_market::Matrix mm = matrix_market::load_matrix(path, log, false);
        csr_matrix::Matrix A = csr_matrix::from_matrix_market(mm);
        csr_matrix::value_array_type x(A.columns, 1.0), y1(A.rows, 0.0);
        csr_matrix::spmv(A, x, y1);
        std::vector<double> want(A.rows), bound(A.rows, 0.0);
        for (int i = 0; i < A.rows; i++) {
            want[i] = runs * y1[i];
            for (int k = A.row_ptr[i]; k < A.row_ptr[i + 1]; k++) bound[i] += std::fabs(A.value[k]);
        }

        auto report = [&](std::string const & name, Verdict v, std::string const & extra) {
            std::cout << "{\"name\": \"" << name << "\", \"runs\": " << runs << ", \"max_err_over_bound\": " << v.worst
                      << ", \"ok\": " << (v.ok ? "true" : "false") << extra << "}" << std::endl;
            if (!v.ok) failures++;
        };

        {
            cuda_csr_spmv_kernel k(path);
            k.init(cfg, log, false);
            drive(k, cfg, runs);
            std::ostringstream p; k.print(p);
            report(k.name(), compare(k.device_y(), want, bound, runs), ", \"print_bytes\": " + std::to_string(p.str().size()));
        }
        {
            cuda_ell_spmv_kernel k(path);
            k.init(cfg, log, false);
            drive(k, cfg, runs);
            report(k.name(), compare(k.device_y(), want, bound, runs), "");
        }
        for (int atomic = 0; atomic < 2; atomic++) {
            cuda_coo_spmv_kernel k(path, atomic != 0);
            k.init(cfg, log, false);
            drive(k, cfg, runs);
            report(k.name(), compare(k.device_y(), want, bound, runs), "");
        }
        {
            cuda_hybrid_spmv_kernel k(path);
            k.init(cfg, log, false);
            drive(k, cfg, runs);
            report(k.name(), compare(k.device_y(), want, bound, runs),
                   ", \"ell_row_length\": " + std::to_string(k.A.ell_row_length) + ", \"num_coo_entries\": " + std::to_string(k.A.num_coo_entries));
        }
        if (A.rows == A.columns) {
            // x <- A x, `runs` times, one rank per thread of the trace configuration, against the host CSR kernel
            cuda_csr_dist_spmv_kernel k(path);
            k.init(cfg, log, false);
            drive(k, cfg, runs);
            csr_matrix::value_array_type xi(A.columns, 1.0);
            std::vector<double> b(A.rows, 1.0);
            for (int r = 0; r < runs; r++) {
                csr_matrix::value_array_type yi(A.rows, 0.0);
                csr_matrix::spmv(A, xi, yi);
                std::vector<double> nb(A.rows, 0.0);
                for (int i = 0; i < A.rows; i++)
                    for (int kk = A.row_ptr[i]; kk < A.row_ptr[i + 1]; kk++) nb[i] += std::fabs(A.value[kk]) * b[A.column_index[kk]];
                xi = yi;
                b = nb;
            }
            std::vector<double> w(xi.begin(), xi.end());
            std::ostringstream p; k.print(p);
            report(k.name(), compare(k.gather_x(), w, b, runs + 1), ", \"ranks\": " + std::to_string(k.ranks.size()));
        }
    } catch (std::exception const & e) {
        std::cerr << "harness: " << e.what() << std::endl;
        return 1;
    }
    return failures ? 3 : 0;
}
