"""The C++ host layer above the C ABI: Kernel-plugin mirror + spmv-b200 CLI (csrc/plugin/).

CPU part: the binary exists, parses options like the reference CLI, and maps failures to
"<kernel name>: <path>: <what>" + EXIT_FAILURE (src/main.cpp:261-270, csr-spmv.cpp:37-45).
GPU part: profile mode on the poisson2D fixture for every format, JSON in the reference's shape.
"""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "spmv_cache_trace_b200", "bin", "spmv-b200")
MTX = os.path.join(ROOT, "tests", "golden", "poisson2D.mtx")


def run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=300)


def test_cli_usage_and_errors():
    assert os.path.exists(CLI), "build it with python -m spmv_cache_trace_b200.build"
    p = run("--help")
    assert p.returncode == 0 and "cuda-hybrid" in p.stdout
    p = run("--spmv-format", "csr", "--matrix", MTX)  # CPU formats belong to the reference binary
    assert p.returncode != 0 and "invalid argument" in p.stderr
    p = run("--spmv-format", "cuda-ell", "--matrix", "/no/such/file.mtx")
    assert p.returncode != 0
    assert p.stderr.startswith("cuda-ell-spmv: /no/such/file.mtx: ")


def test_cli_accepts_the_reference_trace_config(tmp_path):
    """-c/--trace-config is mandatory in the reference CLI (main.cpp:152-153); here only its thread count matters."""
    bad = tmp_path / "cfg.json"
    bad.write_text('{"trace-config": {"caches": {}, "numa_domains": []}}')
    p = run("--spmv-format", "cuda-csr", "--matrix", MTX, "-c", str(bad))
    assert p.returncode != 0 and "no thread_affinities" in p.stderr


def test_cli_bad_matrix_market(tmp_path):
    bad = tmp_path / "bad.mtx"
    bad.write_text("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n")
    p = run("--spmv-format", "cuda-csr", "--matrix", str(bad))
    assert p.returncode != 0 and "Failed to parse entries" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,name,mfmt", [("cuda-csr", "cuda-csr-spmv", "csr"), ("cuda-coo", "cuda-coo-spmv", "coo"),
                                           ("cuda-coo-atomic", "cuda-coo-spmv-atomic", "coo"),
                                           ("cuda-ell", "cuda-ell-spmv", "ell"), ("cuda-hybrid", "cuda-hybrid-spmv", "hybrid")])
def test_cli_profile_mode(fmt, name, mfmt):
    p = run("--spmv-format", fmt, "--matrix", MTX, "--profile", "5", "--threads", "2")
    assert p.returncode == 0, p.stderr
    doc = json.loads(p.stdout)
    k = doc["kernel"]
    assert (k["name"], k["matrix_format"], k["rows"], k["columns"], k["nonzeros"]) == (name, mfmt, 367, 367, 2417)
    assert (k["x_size"], k["y_size"]) == (8 * 367, 8 * 367)
    sizes = {"csr": 12 * 2417 + 4 * 368, "coo": 16 * 2417, "ell": 12 * 367 * 9, "hybrid": 12 * 367 * 7 + 16 * 83}
    assert k["matrix_size"] == sizes[mfmt]
    if mfmt == "hybrid":
        assert (k["ell_row_length"], k["num_ell_entries"], k["num_coo_entries"]) == (7, 367 * 7, 83)
    t = doc["execution_time"]
    assert t["samples"] == 5 and t["unit"] == "ns" and 0 < t["min"] <= t["median"] <= t["max"]
    assert doc["host_execution_time"]["samples"] == 5
    assert doc["roofline"]["bytes"] == k["matrix_size"] + k["x_size"] + k["y_size"]
    assert doc["roofline"]["flops"] == 2 * 2417


@pytest.mark.gpu
def test_cli_with_readme_trace_config(tmp_path):
    cfg = tmp_path / "trace-config.json"  # the README's two-thread hierarchy (README.md:53-66)
    cfg.write_text("""{"trace-config": {"caches": {
      "L1-0": {"size": 32768, "line_size": 64, "parent": "L2-0"}, "L1-1": {"size": 32768, "line_size": 64, "parent": "L2-1"},
      "L2-0": {"size": 262144, "line_size": 64, "parent": "L3"}, "L2-1": {"size": 262144, "line_size": 64, "parent": "L3"},
      "L3": {"size": 20971520, "line_size": 64, "parent": null}},
      "numa_domains": ["a [string] with {braces}"],
      "thread_affinities": [{"thread": 0, "cpu": 0, "cache": "L1-0", "numa_domain": 0},
                            {"thread": 1, "cpu": 1, "cache": "L1-1", "numa_domain": 1}]}}""")
    p = run("--spmv-format", "cuda-ell", "-m", MTX, "-c", str(cfg), "-p", "3", "--warmup", "--flush-caches")
    assert p.returncode == 0, p.stderr
    doc = json.loads(p.stdout)
    assert doc["execution_time"]["samples"] == 3 and doc["kernel"]["name"] == "cuda-ell-spmv"


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["cuda-csr", "cuda-ell", "cuda-coo", "cuda-hybrid"])
def test_cli_x_gather_report(fmt):
    """--x-gather: the cache-model prediction for a 2-way partition, tiny cache so that x is re-fetched."""
    p = run("--spmv-format", fmt, "--matrix", MTX, "--profile", "2", "--x-gather", "2", "--cache-bytes", "512",
            "--line-bytes", "32")
    assert p.returncode == 0, p.stderr
    g = json.loads(p.stdout)["x_gather"]
    assert g["cache_bytes"] == 512 and g["line_bytes"] == 32 and len(g["parts"]) == 2
    stored = {"cuda-csr": 2417, "cuda-coo": 2417, "cuda-ell": 367 * 9, "cuda-hybrid": 367 * 7 + 83}[fmt]
    assert sum(q["x_references"] for q in g["parts"]) == stored
    for q in g["parts"]:
        m = q["misses"]
        assert q["x_gather_miss_bytes"] == 32 * (m["x_local"] + m["x_remote"]) > 0
        assert q["predicted_dram_bytes"] == 32 * sum(m.values())
    if fmt == "cuda-csr":  # balanced non-zeros: the two parts differ by less than one row
        a, b = (q["x_references"] for q in g["parts"])
        assert abs(a - b) <= 9
        assert g["partition"] == "balanced non-zeros"
