"""The drop-in boundary, compiled against the reference itself (integration/).

integration/src/kernels/cuda-spmv.{hpp,cpp} derive from the reference's own `Kernel` (src/kernels/kernel.hpp:18-45);
integration/reference.patch registers them in src/main.cpp (enum :28-37, strcmp chain :139-147, factory :209-232),
src/kernels.hpp and the Makefile (USE_CUDA=1), and lets `--profile` start without libpfm (util/perf-events.cpp:35-45).
integration/build.sh applies all of that to a scratch copy of /root/reference, builds the STOCK binary and a harness that
drives the plugins like profile_kernel does; only the two binaries land in integration/_build/.

CPU: the patch applies and everything compiles and links (when /root/reference is present); the patched binary still
runs the host kernels, profiles without libpfm, and the CUDA formats fail loudly without a GPU.
GPU: every CUDA plugin multiplies like the host kernel of the same format (harness), the stock CLI traces and profiles
with them, and the row-partitioned plugin iterates x <- A x with one rank per thread.
"""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INT = os.path.join(ROOT, "integration")
BIN = os.path.join(INT, "_build", "spmv-cache-trace")
HARNESS = os.path.join(INT, "_build", "harness")
CFG = os.path.join(INT, "trace-config-2threads.json")
MTX = os.path.join(ROOT, "tests", "golden", "poisson2D.mtx")
FORMATS = ["cuda-csr", "cuda-ell", "cuda-coo", "cuda-coo-atomic", "cuda-hybrid"]


def run(*cmd, timeout=600):
    return subprocess.run(list(cmd), capture_output=True, text=True, timeout=timeout)


def need_binaries():
    if not (os.path.exists(BIN) and os.path.exists(HARNESS)):
        pytest.skip("integration/_build not built (needs /root/reference: integration/build.sh)")


def test_patch_applies_and_the_reference_builds_with_the_cuda_plugins():
    if not os.path.isdir("/root/reference/src/kernels"):
        pytest.skip("/root/reference not present")
    p = run(os.path.join(INT, "build.sh"), timeout=1200)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert os.path.exists(BIN) and os.path.exists(HARNESS)
    # no reference source was copied into the repository tree
    allowed = {"cuda-spmv.hpp", "cuda-spmv.cpp", "reference.patch", "build.sh", "harness.cpp", "trace-config-2threads.json",
               "spmv-cache-trace", "harness", "README.md"}
    for dirpath, _, files in os.walk(INT):
        for f in files:
            assert f in allowed, os.path.join(dirpath, f)


def test_patched_binary_keeps_the_host_kernels_and_lists_the_cuda_formats():
    need_binaries()
    p = run(BIN, "--help")
    assert p.returncode == 0 and "cuda-hybrid" in p.stdout and "cuda-csr-dist" in p.stdout
    p = run(BIN, "--spmv-format", "csr", "-m", MTX, "-c", CFG)
    assert p.returncode == 0, p.stderr
    doc = json.loads(p.stdout)
    assert doc["kernel"]["name"] == "csr-spmv" and doc["kernel"]["nonzeros"] == 2417
    assert doc["cache_misses"]["L3"] == [[297, 0], [24, 249]]


def test_profile_mode_starts_without_libpfm():
    """the stock --profile threw 'Please re-build with libpfm enabled' before it measured anything."""
    need_binaries()
    p = run(BIN, "--spmv-format", "ell", "-m", MTX, "-c", CFG, "--profile", "4")
    assert p.returncode == 0, p.stderr
    doc = json.loads(p.stdout)
    assert doc["execution_time"]["samples"] == 4 and doc["execution_time"]["unit"] == "ns"
    assert doc["profiling_events"] == []


def test_cuda_formats_fail_loudly_without_a_gpu():
    need_binaries()
    import spmv_cache_trace_b200 as sp
    if sp.device_count() > 0:
        pytest.skip("a GPU is present")
    for fmt in FORMATS + ["cuda-csr-dist"]:
        p = run(BIN, "--spmv-format", fmt, "-m", MTX, "-c", CFG)
        assert p.returncode != 0 and p.stdout == ""
        name = {"cuda-coo-atomic": "cuda-coo-spmv-atomic"}.get(fmt, fmt + "-spmv")
        assert p.stderr.startswith(f"{name}: {MTX}: "), p.stderr  # main.cpp:264-266, csr-spmv.cpp:37-45
        assert "CUDA" in p.stderr or "device" in p.stderr


@pytest.mark.gpu
def test_harness_every_plugin_multiplies_like_the_host_kernel():
    need_binaries()
    p = run(HARNESS, MTX, CFG, "3")
    assert p.returncode == 0, p.stdout + p.stderr
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert [l["name"] for l in lines] == ["cuda-csr-spmv", "cuda-ell-spmv", "cuda-coo-spmv", "cuda-coo-spmv-atomic",
                                          "cuda-hybrid-spmv", "cuda-csr-dist-spmv"]
    assert all(l["ok"] and l["runs"] == 3 and l["max_err_over_bound"] <= 3e-12 for l in lines)
    hyb = lines[4]
    assert (hyb["ell_row_length"], hyb["num_coo_entries"]) == (7, 83)  # SURVEY 8a11 [probed]
    assert lines[5]["ranks"] == 2


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", FORMATS)
def test_stock_cli_traces_and_profiles_with_the_cuda_plugins(fmt):
    need_binaries()
    host = {"cuda-csr": "csr", "cuda-ell": "ell", "cuda-coo": "coo", "cuda-coo-atomic": "coo-atomic", "cuda-hybrid": "hybrid"}[fmt]
    # trace mode: the plugin owns the same host matrix, so the cache model sees what the host kernel's shows
    a = run(BIN, "--spmv-format", fmt, "-m", MTX, "-c", CFG)
    b = run(BIN, "--spmv-format", host, "-m", MTX, "-c", CFG)
    assert a.returncode == 0, a.stderr
    if host == "hybrid":  # hybrid-spmv.cpp:124 prints a stray ',' line: the host kernel's document is not valid JSON
        assert a.stdout[a.stdout.index('"cache_misses"'):] == b.stdout[b.stdout.index('"cache_misses"'):]
    else:
        assert json.loads(a.stdout)["cache_misses"] == json.loads(b.stdout)["cache_misses"]
    assert json.loads(a.stdout)["kernel"]["name"].startswith("cuda-")
    # profile mode (no libpfm here): execution times of the GPU kernel, through the stock profile_kernel
    p = run(BIN, "--spmv-format", fmt, "-m", MTX, "-c", CFG, "--profile", "5")
    assert p.returncode == 0, p.stderr
    doc = json.loads(p.stdout)
    assert doc["execution_time"]["samples"] == 5 and doc["execution_time"]["min"] > 0
    assert doc["kernel"]["rows"] == 367 and "kernel" in doc["kernel"]["device_kernel"]


@pytest.mark.gpu
def test_stock_cli_row_partitioned_plugin():
    need_binaries()
    p = run(BIN, "--spmv-format", "cuda-csr-dist", "-m", MTX, "-c", CFG, "--profile", "3")
    assert p.returncode == 0, p.stderr
    doc = json.loads(p.stdout)
    ranks = doc["kernel"]["ranks"]
    assert len(ranks) == 2 and sum(r["rows"] for r in ranks) == 367 and sum(r["nonzeros"] for r in ranks) == 2417
    assert [r["rows"] for r in ranks] == [184, 183]  # ceil(367 / 2) and the rest (csr-matrix.cpp:77-83)
