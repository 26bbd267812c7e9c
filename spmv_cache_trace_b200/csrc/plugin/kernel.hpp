// kernel.hpp -- the plugin interface of the reference, mirrored for the CUDA kernels.
//
// Same virtuals, argument meaning and error behaviour as `class Kernel` in the reference
// (src/kernels/kernel.hpp:18-45) so a GPU variant is registered exactly like a CPU one
// (enum + strcmp chain + factory switch, src/main.cpp:28-37, 139-147, 209-232).  Two things are
// deliberately absent: TraceConfig's cache hierarchy (only the thread affinities reach a kernel,
// e.g. csr-spmv.cpp:51-55) and memory_reference_string(), which belongs to the cache-simulation
// subsystem that stays on the host unchanged -- INTEGRATION.md shows the reference-side adapter
// that keeps it by delegating to the host matrix.
#pragma once

#include <iosfwd>
#include <stdexcept>
#include <string>
#include <vector>

// kernel_error (src/kernels/kernel.hpp:11-16)
class kernel_error : public std::runtime_error
{
public:
    explicit kernel_error(std::string const & message) : std::runtime_error(message) {}
};

// What a kernel reads from the reference's TraceConfig (src/trace-config.hpp): one entry per
// OpenMP thread that will enter prepare() / run().
struct ThreadAffinity {
    int thread;
    int cpu;
    int numa_domain;
};

class TraceConfig
{
public:
    explicit TraceConfig(int num_threads = 1)
    {
        for (int t = 0; t < num_threads; t++) affinities_.push_back(ThreadAffinity{t, t, 0});
    }
    std::vector<ThreadAffinity> const & thread_affinities() const { return affinities_; }

private:
    std::vector<ThreadAffinity> affinities_;
};

class Kernel
{
public:
    virtual ~Kernel() {}

    // main thread only (src/main.cpp:237): load + convert; x = 1, y = 0
    virtual void init(TraceConfig const & trace_config, std::ostream & o, bool verbose) = 0;
    // called by EVERY thread of the enclosing omp parallel region (src/profile-kernel.cpp:227, 262)
    virtual void prepare(TraceConfig const & trace_config) = 0;
    // called by every thread, between barriers (src/profile-kernel.cpp:159-161): y += A*x
    virtual void run(TraceConfig const & trace_config) = 0;

    virtual std::string name() const = 0;
    virtual std::ostream & print(std::ostream & o) const = 0;
};

inline std::ostream & operator<<(std::ostream & o, Kernel const & kernel) { return kernel.print(o); }
