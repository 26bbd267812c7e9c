"""The host-side cache model (spmvb200_cache_trace_*) against the reference's own cache simulation.

The model is the reference's (fully associative LRU, cache-simulation/lru.cpp; reference strings of
matrix/{csr,ell,coo}-matrix.cpp; interleaving of replacement.cpp:41-95) computed with an O(1) LRU and
extended with per-array attribution and arbitrary partitions.  Pinned three ways:
  1. live against the UNMODIFIED reference simulator (oracle/_ref), random matrices, 1-3 threads;
  2. against committed golden numbers produced by it (tests/golden/cache_trace.json, made by
     tests/golden/make_cache_golden.py) so the pin also holds where /root/reference is absent;
  3. against the known answer for BASELINE config 1 recorded in SURVEY.md section 6 (CSR, 2 threads,
     README hierarchy: L1 32 KiB, L2 256 KiB, L3 20 MiB, 64 B lines).
No GPU is needed: the model never touches CUDA.
"""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import spmv_cache_trace_b200 as sp
from oracle.generators_ref import stencil_entries
from oracle.oracle import Ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cache_trace.json")


def per_domain(res):
    """[[local, remote]...] summed over arrays, the shape the reference reports per thread."""
    out = []
    for r in res:
        local = r["misses_index"] + r["misses_column_index"] + r["misses_value"] + r["misses_x_local"] + r["misses_y_local"]
        out.append((local, r["misses_x_remote"] + r["misses_y_remote"]))
    return out


def ref_matrix_local_remote(m, T):
    """The reference's [thread][domain] table folded to (own domain, other domains)."""
    return [(int(m[t, t]), int(m[t].sum() - m[t, t])) for t in range(T)]


def random_case(seed, rows=300, cols=300, per_row=6):
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, per_row + 1, rows)
    ii = np.repeat(np.arange(rows), lens)
    jj = np.concatenate([np.sort(rng.choice(cols, l, replace=False)) for l in lens])
    a = rng.uniform(-1, 1, len(ii))
    return (ii + 1).astype(np.int32), (jj + 1).astype(np.int32), a


def csr_of(rows, i, j):
    order = np.lexsort((j, i))
    i0, j0 = i[order] - 1, j[order] - 1
    rp = np.zeros(rows + 1, np.int64)
    np.add.at(rp, i0 + 1, 1)
    return np.cumsum(rp), j0.astype(np.int32), i0.astype(np.int32)


CASES = [(seed, T, cache, line, warm) for seed in (1, 2) for T in (1, 2, 3)
         for cache, line in ((2048, 64), (16384, 64), (4096, 32)) for warm in (False, True)]


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref/libspmvref.so not built")
@pytest.mark.parametrize("seed,T,cache,line,warm", CASES)
def test_csr_ell_coo_match_the_reference_simulator(seed, T, cache, line, warm):
    rows = cols = 300
    i, j, a = random_case(seed, rows, cols)
    R = Ref().from_entries(rows, cols, i, j, a)
    rp, col, row = csr_of(rows, i, j)
    # CSR
    R.convert("csr")
    want = ref_matrix_local_remote(R.cache_trace(T, cache, line, warm), T)
    got = per_domain(sp.cache_model.csr(rows, cols, rp, col, cache, line, parts=T, shared=True, warmup=warm, page_bytes=4096))
    assert got == want
    # ELL (row-major reference layout)
    E = R.convert("ell")
    W = E.row_length
    want = ref_matrix_local_remote(R.cache_trace(T, cache, line, warm), T)
    got = per_domain(sp.cache_model.ell(rows, cols, W, np.asarray(E.column_index), cache, line, parts=T, shared=True,
                                        warmup=warm, page_bytes=4096))
    assert got == want
    # COO, atomic form, file order
    Cm = R.convert("coo-atomic")
    want = ref_matrix_local_remote(R.cache_trace(T, cache, line, warm), T)
    got = per_domain(sp.cache_model.coo(rows, cols, np.asarray(Cm.row_index), np.asarray(Cm.column_index), cache, line,
                                        parts=T, shared=True, warmup=warm, page_bytes=4096))
    assert got == want


def test_golden_numbers_from_the_reference_simulator():
    gold = json.load(open(GOLD))
    for g in gold["cases"]:
        i, j, a = random_case(g["seed"], g["rows"], g["cols"])
        rp, col, row = csr_of(g["rows"], i, j)
        got = per_domain(sp.cache_model.csr(g["rows"], g["cols"], rp, col, g["cache_bytes"], g["line_bytes"], parts=g["threads"],
                                            shared=True, warmup=g["warmup"], page_bytes=4096))
        assert [list(t) for t in got] == g["csr"], g
        got = per_domain(sp.cache_model.coo(g["rows"], g["cols"], i - 1, j - 1, g["cache_bytes"], g["line_bytes"],
                                            parts=g["threads"], shared=True, warmup=g["warmup"], page_bytes=4096))
        assert [list(t) for t in got] == g["coo_atomic"], g


def test_config1_known_answer_from_the_survey():
    """reference binary, config 1 (2D 5-point 1000x1000), CSR, 2 threads:
    L1-0 [[749404,97],[0,0]], L1-1 [[0,0],[181,749320]], L2-0 [[624654,97],..], L2-1 [..,[153,624598]],
    L3 [[624654,97],[153,624598]].  L1-x / L2-x are private to one thread (only that thread is active for
    them); L3 is shared.  The reference runs every cache on the full reference string."""
    n = 1000
    i, j, a = stencil_entries(0, n, n)
    rp, col, _ = csr_of(n * n, i, j)
    N = n * n
    l1 = per_domain(sp.cache_model.csr(N, N, rp, col, 32768, 64, parts=2, shared=False, page_bytes=4096))
    assert l1 == [(749404, 97), (749320, 181)]
    l2 = per_domain(sp.cache_model.csr(N, N, rp, col, 262144, 64, parts=2, shared=False, page_bytes=4096))
    assert l2 == [(624654, 97), (624598, 153)]
    l3 = per_domain(sp.cache_model.csr(N, N, rp, col, 20971520, 64, parts=2, shared=True, page_bytes=4096))
    assert l3 == [(624654, 97), (624598, 153)]
    # compulsory traffic: sum of L3 misses x 64 B ~ matrix + x + y bytes (79.95 MB)
    total = sum(a + b for a, b in l3) * 64
    assert abs(total - 79952004) < 0.001 * 79952004


def test_attribution_partition_and_bypass():
    n = 64
    i, j, a = stencil_entries(1, n, n, n)  # 3D 7-point, 262144 rows
    N = n ** 3
    rp, col, _ = csr_of(N, i, j)
    nnz = int(rp[-1])
    one = sp.cache_model.csr(N, N, rp, col, 1 << 20, 128, parts=1)[0]
    # references: 3 per entry + 2 per row + 1 (csr-matrix.cpp:114)
    assert one["references"] == 3 * nnz + 2 * N + 1 and one["x_references"] == nnz
    # streamed arrays miss exactly once per line
    assert one["misses_column_index"] == (4 * nnz + 127) // 128 and one["misses_value"] == (8 * nnz + 127) // 128
    assert one["misses_y_local"] == N * 8 // 128 and one["misses_x_remote"] == 0
    # a 1 MiB cache holds the three x planes a 64^3 sweep needs (3 * 64*64*8 B = 96 KiB): x misses are compulsory
    assert one["misses_x_local"] == N * 8 // 128
    # Reusing the z-1 / z+1 neighbours needs two planes of x (64 KiB) to survive ~8200 rows of sweep, during
    # which the matrix streams bring in 8200*7*12 B = 690 KiB: a 256 KiB LRU loses the far planes ...
    small = sp.cache_model.csr(N, N, rp, col, 256 * 1024, 128, parts=1)[0]
    assert small["misses_x_local"] > 2 * one["misses_x_local"]
    # ... unless the streams bypass the cache (the kernels' evict-first policy): x misses are compulsory again
    bypass = sp.cache_model.csr(N, N, rp, col, 256 * 1024, 128, parts=1, stream_bypass=True)[0]
    assert bypass["misses_x_local"] == one["misses_x_local"]
    assert bypass["misses_value"] == one["misses_value"]
    # row partition into 4 z-slabs, private caches, ownership by row range: remote x = the halo planes
    starts = np.array([0, N // 4, N // 2, 3 * N // 4, N], np.int64)
    parts = sp.cache_model.csr(N, N, rp, col, 1 << 20, 128, parts=4, starts=starts, shared=False)
    plane_lines = n * n * 8 // 128
    assert [p["misses_x_remote"] for p in parts] == [plane_lines, 2 * plane_lines, 2 * plane_lines, plane_lines]
    assert sum(p["references"] for p in parts) == 3 * nnz + 2 * N + 4
    # an unbalanced partition is accepted, a malformed one is not
    sp.cache_model.csr(N, N, rp, col, 1 << 20, 128, parts=2, starts=[0, 1000, N], shared=False)
    with pytest.raises(sp.matrix_error):
        sp.cache_model.csr(N, N, rp, col, 1 << 20, 128, parts=2, starts=[0, N + 1, N])
    with pytest.raises(sp.matrix_error):
        sp.cache_model.csr(N, N, rp, col, 0, 128)
