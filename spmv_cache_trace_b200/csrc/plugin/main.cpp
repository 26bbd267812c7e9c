// spmv-b200 -- the profile mode of the reference CLI for the CUDA kernels.
//
//   spmv-b200 --spmv-format cuda-csr|cuda-coo|cuda-coo-atomic|cuda-ell|cuda-hybrid \
//             --matrix PATH [--profile N] [--threads T] [--verbose]
//
// Follows src/main.cpp (option names :166-187, kernel factory :209-232, error mapping :261-270) and
// the protocol of profile_kernel (src/profile-kernel.cpp:197-313): an omp parallel region of T
// threads, prepare, one warm-up run, then N runs each bracketed barrier / clock / barrier -- and,
// because the host clock around an asynchronous launch says little, the same N runs timed again
// with CUDA events.  Output: one JSON document in the reference's shape ("kernel",
// "execution_time" with print_sample's statistics, src/util/sample.hpp:137-165) plus "roofline".
// The cache-trace mode (--profile 0 in the reference) stays with the reference binary; what the GPU
// path adds is --x-gather PARTS: the reference's LRU model (spmvb200_cache_trace) run for the row
// partition the multi-GPU mode would choose (balanced non-zeros for CSR, the reference rule otherwise),
// one private cache of the size of this GPU's L2 per part, misses attributed per array -- the predicted
// x-gather traffic to put next to ncu's dram__bytes ("x_gather" in the output).
#include "cuda_spmv_kernels.hpp"

#include <getopt.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

// Number of entries of "thread_affinities" in a reference trace-config file (src/trace-config.cpp:263-343):
// the only thing a GPU kernel takes from it (one OpenMP thread per entry enters prepare()/run()).
// Returns 0 when the file cannot be read or has no such array.
static int trace_config_threads(std::string const & path)
{
    FILE * f = std::fopen(path.c_str(), "rb");
    if (!f) return 0;
    std::string text;
    char buf[4096];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
    std::fclose(f);
    size_t pos = text.find("\"thread_affinities\"");
    if (pos == std::string::npos) return 0;
    pos = text.find('[', pos);
    if (pos == std::string::npos) return 0;
    int depth = 0, count = 0;
    bool in_string = false;
    for (size_t k = pos; k < text.size(); k++) {
        const char ch = text[k];
        if (in_string) {
            if (ch == '\\') k++;
            else if (ch == '"') in_string = false;
            continue;
        }
        if (ch == '"') in_string = true;
        else if (ch == '[' || ch == '{') {
            if (ch == '{' && depth == 1) count++;
            depth++;
        } else if (ch == ']' || ch == '}') {
            if (--depth == 0) break;
        }
    }
    return count;
}

static void print_sample(std::ostream & o, std::vector<double> v, const char * unit)
{
    const size_t n = v.size();
    double mn = 0, mx = 0, mean = 0, med = 0, var = 0, skew = NAN, kurt = NAN;
    if (n) {
        std::sort(v.begin(), v.end());
        mn = v.front(); mx = v.back(); med = v[n / 2];  // upper median, like sample.hpp:51-53
        for (double x : v) mean += x;
        mean /= (double)n;
        double m2 = 0, m3 = 0, m4 = 0;
        for (double x : v) { const double d = x - mean; m2 += d * d; m3 += d * d * d; m4 += d * d * d * d; }
        m2 /= (double)n; m3 /= (double)n; m4 /= (double)n;
        var = n > 1 ? m2 * (double)n / (double)(n - 1) : 0.0;
        if (m2 > 0) { skew = m3 / std::pow(m2, 1.5); kurt = m4 / (m2 * m2); }
    }
    auto num = [&](double x) { if (std::isnan(x)) o << "\"nan\""; else o << x; };
    o << "{\n\"samples\": " << n << ",\n\"min\": "; num(mn);
    o << ",\n\"max\": "; num(mx);
    o << ",\n\"mean\": "; num(mean);
    o << ",\n\"median\": "; num(med);
    o << ",\n\"variance\": "; num(var);
    o << ",\n\"standard_deviation\": "; num(std::sqrt(var));
    o << ",\n\"skewness\": "; num(skew);
    o << ",\n\"kurtosis\": "; num(kurt);
    o << ",\n\"unit\": \"" << unit << "\"\n}";
}

int main(int argc, char ** argv)
{
    std::string format, matrix_path;
    int profile = 10, threads = 1, gather_parts = 0, line_bytes = 32;
    long long cache_bytes = 0;
    bool verbose = false;
    static option longopts[] = {{"spmv-format", required_argument, nullptr, 'f'}, {"matrix", required_argument, nullptr, 'm'},
                                {"profile", required_argument, nullptr, 'p'}, {"threads", required_argument, nullptr, 't'},
                                {"x-gather", required_argument, nullptr, 'x'}, {"cache-bytes", required_argument, nullptr, 1001},
                                {"line-bytes", required_argument, nullptr, 'l'}, {"trace-config", required_argument, nullptr, 'c'},
                                {"warmup", no_argument, nullptr, 1002}, {"flush-caches", no_argument, nullptr, 1003},
                                {"gp-partitioner", no_argument, nullptr, 1004},
                                {"verbose", no_argument, nullptr, 'v'}, {"help", no_argument, nullptr, 'h'},
                                {nullptr, 0, nullptr, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "f:m:p:t:x:c:l:vh", longopts, nullptr)) != -1) {
        switch (c) {
        case 'f': format = optarg; break;
        case 'm': matrix_path = optarg; break;
        case 'p': profile = std::atoi(optarg); break;
        case 't': threads = std::max(1, std::atoi(optarg)); break;
        case 'x': gather_parts = std::max(1, std::atoi(optarg)); break;
        case 1001: cache_bytes = std::atoll(optarg); break;
        case 'c': {  // the reference's mandatory -c/--trace-config (main.cpp:152-153): only its thread count matters here
            const int t = trace_config_threads(optarg);
            if (t <= 0) {
                std::cerr << "spmv-b200: " << optarg << ": no thread_affinities found\n";
                return EXIT_FAILURE;
            }
            threads = t;
            break;
        }
        case 1002: case 1003: break;  // --warmup / --flush-caches of the reference: a warm-up run is always done, and
                                      // cache flushing is a CPU notion (the L2-cold protocol lives in bench.py)
        case 1004: spmvb200_set_global_option("mm.gp_partitioner", 1); break;
        case 'l': line_bytes = std::max(1, std::atoi(optarg)); break;
        case 'v': verbose = true; break;
        default:
            std::cout << "Usage: spmv-b200 --spmv-format FMT --matrix PATH [--profile N] [--threads T | -c TRACE_CONFIG] [--verbose]\n"
                         "                 [--x-gather PARTS [--cache-bytes B] [--line-bytes L]] [--gp-partitioner]\n"
                         "  FMT: cuda-csr, cuda-coo, cuda-coo-atomic, cuda-ell, cuda-hybrid\n"
                         "  --x-gather PARTS  cache-model prediction of the x-gather misses for a PARTS-way partition\n"
                         "                    (default cache: this GPU's L2; default line: the 32 B DRAM sector)\n"
                         "  --gp-partitioner  a PATH ending in __GP<n> is reordered by the built-in K-way graph partitioner\n"
                         "                    (the reference needs a METIS build for that suffix; default: no reordering)\n";
            return c == 'h' ? EXIT_SUCCESS : EXIT_FAILURE;
        }
    }
    std::unique_ptr<Kernel> kernel = make_cuda_kernel(format, matrix_path);
    if (!kernel) {
        std::cerr << "spmv-b200: invalid argument for --spmv-format\n";
        return EXIT_FAILURE;
    }
    try {
        TraceConfig trace_config(threads);
        kernel->init(trace_config, std::cerr, verbose);
        std::vector<double> host_ns((size_t)std::max(profile, 0));
#ifdef _OPENMP
        omp_set_num_threads(threads);
#endif
        std::string failure;
#pragma omp parallel
        {
            try {
                kernel->prepare(trace_config);
                kernel->run(trace_config);  // warm-up (profile-kernel.cpp:263-264)
                for (int r = 0; r < profile; r++) {
                    std::chrono::steady_clock::time_point t0, t1;
#pragma omp barrier
#pragma omp master
                    t0 = std::chrono::steady_clock::now();
#pragma omp barrier
                    kernel->run(trace_config);
#pragma omp barrier
#pragma omp master
                    {
                        t1 = std::chrono::steady_clock::now();
                        host_ns[(size_t)r] = (double)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
                    }
#pragma omp barrier
                }
            } catch (std::exception const & e) {  // profile-kernel.cpp:299-307
#pragma omp critical
                failure = e.what();
            }
        }
        if (!failure.empty()) throw kernel_error(failure);

        auto * gpu = dynamic_cast<cuda_spmv_kernel *>(kernel.get());
        std::vector<float> ms((size_t)std::max(profile, 0));
        if (profile > 0 && spmvb200_time(gpu->handle(), 0, profile, ms.data()) != 0) throw kernel_error(spmvb200_last_error());
        std::vector<double> dev_ns(ms.begin(), ms.end());
        for (double & v : dev_ns) v *= 1e6;
        spmvb200_info info{};
        spmvb200_matrix_info(gpu->handle(), &info);
        const double bytes = (double)(info.matrix_size + info.x_size + info.y_size);
        double best = 0;
        if (!dev_ns.empty()) best = *std::min_element(dev_ns.begin(), dev_ns.end());

        std::cout << "{\n\"kernel\": " << *kernel << ",\n\"execution_time\": ";
        print_sample(std::cout, dev_ns, "ns");
        std::cout << ",\n\"host_execution_time\": ";
        print_sample(std::cout, host_ns, "ns");
        std::cout << ",\n\"roofline\": {\n\"bytes\": " << (long long)bytes << ",\n\"flops\": " << 2 * info.num_entries
                  << ",\n\"gpu_kernel\": \"" << spmvb200_kernel_name(gpu->handle()) << "\""
                  << ",\n\"best_gbs\": " << (best > 0 ? bytes / best : 0.0)
                  << ",\n\"best_gflops\": " << (best > 0 ? 2.0 * (double)info.num_entries / best : 0.0)
                  << ",\n\"fraction_of_8TBs\": " << (best > 0 ? bytes / best / 8000.0 : 0.0) << "\n}";
        if (gather_parts > 0) {
            int dev = 0, sms = 0, maj = 0, min = 0;
            int64_t l2 = 0, mem = 0;
            char name[128];
            if (cache_bytes <= 0 && spmvb200_device_props(dev, name, sizeof name, &sms, &l2, &mem, &maj, &min) == 0) cache_bytes = l2;
            std::vector<int64_t> starts((size_t)gather_parts + 1);
            const bool by_nnz = info.format == SPMVB200_CSR;
            int rc = by_nnz ? spmvb200_partition_rows_nnz(gpu->handle(), gather_parts, starts.data())
                            : spmvb200_partition_rows_ref(info.format == SPMVB200_COO ? info.num_entries : info.rows, gather_parts, starts.data());
            if (rc != 0) throw kernel_error(spmvb200_last_error());
            spmvb200_cache_config cfg{};
            cfg.cache_bytes = cache_bytes; cfg.line_bytes = line_bytes; cfg.parts = gather_parts;
            cfg.starts = info.format == SPMVB200_HYB ? nullptr : starts.data();
            cfg.shared = 0; cfg.warmup = 0; cfg.page_bytes = 0; cfg.stream_bypass = 1;
            std::vector<spmvb200_cache_misses> miss((size_t)gather_parts);
            if (spmvb200_cache_trace(gpu->handle(), &cfg, miss.data()) != 0) throw kernel_error(spmvb200_last_error());
            std::cout << ",\n\"x_gather\": {\n\"model\": \"fully associative LRU (reference cache-simulation/lru.cpp), one private cache per part, "
                         "matrix streams bypass the cache\",\n\"cache_bytes\": " << cache_bytes << ",\n\"line_bytes\": " << line_bytes
                      << ",\n\"partition\": \"" << (by_nnz ? "balanced non-zeros" : "reference rule ceil(n/parts)") << "\",\n\"parts\": [";
            for (int p = 0; p < gather_parts; p++) {
                const spmvb200_cache_misses & q = miss[(size_t)p];
                std::cout << (p ? ",\n" : "\n") << "{\"first\": " << starts[(size_t)p] << ", \"end\": " << starts[(size_t)p + 1]
                          << ", \"references\": " << q.references << ", \"x_references\": " << q.x_references
                          << ", \"x_remote_references\": " << q.x_remote_references
                          << ", \"misses\": {\"index\": " << q.misses_index << ", \"column_index\": " << q.misses_column_index
                          << ", \"value\": " << q.misses_value << ", \"x_local\": " << q.misses_x_local << ", \"x_remote\": "
                          << q.misses_x_remote << ", \"y_local\": " << q.misses_y_local << ", \"y_remote\": " << q.misses_y_remote
                          << "}, \"x_gather_miss_bytes\": " << (q.misses_x_local + q.misses_x_remote) * (long long)line_bytes
                          << ", \"predicted_dram_bytes\": "
                          << (q.misses_index + q.misses_column_index + q.misses_value + q.misses_x_local + q.misses_x_remote +
                              q.misses_y_local + q.misses_y_remote) * (long long)line_bytes << "}";
            }
            std::cout << "\n]\n}";
        }
        std::cout << "\n}\n";
    } catch (kernel_error const & e) {
        std::cerr << kernel->name() << ": " << e.what() << '\n';
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}
