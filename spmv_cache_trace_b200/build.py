"""Build libspmvb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m spmv_cache_trace_b200.build [--force] [--verbose]

Objects go to spmv_cache_trace_b200/build/, the library to spmv_cache_trace_b200/lib/.
Both are git-ignored; the .so travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# SPMVB200_VARIANT=name builds an experiment variant next to the product (objects in build_<name>/, library
# lib/libspmvb200_<name>.so, flags from SPMVB200_CFLAGS); the Python layer loads it when SPMVB200_LIB names it.
VARIANT = os.environ.get("SPMVB200_VARIANT", "")
OBJ = os.path.join(HERE, "build" + ("_" + VARIANT if VARIANT else ""))
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libspmvb200" + ("_" + VARIANT if VARIANT else "") + ".so")
CLI = os.path.join(HERE, "bin", "spmv-b200")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

CUDA_SOURCES = ["abi.cu", "kernels_csr.cu", "kernels_csr_warp.cu", "kernels_csr_flat.cu", "kernels_csr_sliced.cu", "kernels_ell.cu", "kernels_coo.cu", "builders.cu", "generators.cu", "dist.cu"]
CXX_SOURCES = ["mm_host.cpp", "cache_model.cpp", "reorder_host.cpp"]
HEADERS = ["common.cuh", "ptx.cuh", "launch.cuh", "segreduce.cuh", "mm_host.hpp", os.path.join(INCLUDE, "spmv_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx() -> str:
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), p.stdout))
    if verbose and p.stdout.strip():
        print(p.stdout)
    return p.stdout


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc, cxx = _nvcc(), _host_cxx()
    headers = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    jobs, objs = [], []
    for src in CUDA_SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or not _newer(o, [s] + headers):
            flags = NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + os.environ.get("SPMVB200_CFLAGS", "").split()
            jobs.append([nvcc, "-ccbin", cxx] + flags + ["-I", INCLUDE, "-c", s, "-o", o])
    for src in CXX_SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or not _newer(o, [s] + headers):
            jobs.append([cxx, "-O2", "-std=c++17", "-fPIC", "-pthread", "-Wall", "-I", INCLUDE, "-c", s, "-o", o])
    outputs = []
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            outputs = list(ex.map(lambda c: _run(c, verbose), jobs))
    if jobs or not os.path.exists(LIB):
        _run([nvcc, "-ccbin", cxx, "-shared", "-o", LIB] + objs + ["-lz", "-ldl", "-Xlinker", "--no-undefined"], verbose)
    if ptxas_info:
        print("\n".join(outputs))
    # the C++ host layer above the C ABI: Kernel-plugin mirror + the profile-mode CLI
    plugin = os.path.join(CSRC, "plugin")
    srcs = [os.path.join(plugin, f) for f in ("main.cpp", "cuda_spmv_kernels.cpp")]
    deps = srcs + [os.path.join(plugin, f) for f in ("kernel.hpp", "cuda_spmv_kernels.hpp")] + [LIB]
    if not VARIANT and (force or not _newer(CLI, deps)):
        os.makedirs(os.path.dirname(CLI), exist_ok=True)
        _run([cxx, "-O2", "-std=c++17", "-fopenmp", "-Wall", "-I", INCLUDE, "-o", CLI] + srcs +
             ["-L", LIBDIR, "-lspmvb200", "-Wl,-rpath,$ORIGIN/../lib"], verbose)
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ptxas_info="--ptxas" in sys.argv)
    print(lib)
