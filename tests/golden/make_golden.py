#!/usr/bin/env python
"""Regenerate tests/golden/* from the reference (run in the build container only).

Two sources, both read-only:
  1. /root/reference/test/poisson2D.hpp -- the reference's golden fixture
     (SuiteSparse FEMLAB/poisson2D, 367x367, 2417 nnz, with x and expected y);
     written out as poisson2D.mtx / poisson2D_b.txt / poisson2D_result.txt.
  2. The reference library itself (oracle/_ref/libspmvref.so): conversions and
     y = A*x for seeded matrices, incl. the edge cases the reference tests touch
     (ragged rows, empty rows, unsorted and column-major input, row-aligned CSR,
     skip_padding ELL, multi-thread COO/hybrid workspace reduce) -> ref_vectors.npz.

The small known-answer tests of test/test_{csr,coo,ell,hybrid}-matrix.cpp are
transcribed by hand into kats.json (line numbers cited there).

Nothing in tests/ reads /root/reference at run time; only this script does.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REF = os.environ.get("SPMV_REFERENCE", "/root/reference")


def extract_poisson2d():
    src = open(os.path.join(REF, "test", "poisson2D.hpp")).read()
    m = re.search(r'poisson2D\{R"\((.*?)\)"\}', src, re.S)
    mtx = m.group(1)
    if not mtx.endswith("\n"):
        mtx += "\n"
    open(os.path.join(HERE, "poisson2D.mtx"), "w").write(mtx)
    for name, out in (("poisson2D_b", "poisson2D_b.txt"), ("poisson2D_result", "poisson2D_result.txt")):
        m = re.search(name + r" = std::vector<double>\{\s*\{(.*?)\}\};", src, re.S)
        vals = [t.strip() for t in m.group(1).replace("\n", " ").split(",") if t.strip()]
        assert len(vals) == 367, (name, len(vals))
        # keep the literal text: float(repr) round-trips, and the file shows what the reference holds
        open(os.path.join(HERE, out), "w").write("\n".join(vals) + "\n")
    print("poisson2D: wrote mtx + b + result")


def seeded_cases():
    """(name, rows, cols, i, j, a) with 1-based unsorted entries and no duplicate (i,j)."""
    rng = np.random.default_rng(0x5EED)
    cases = []

    def rand_case(name, rows, cols, density, empty_rows=(), order="random", heavy=()):
        mask = rng.random((rows, cols)) < density
        for r in empty_rows:
            mask[r, :] = False
        for r in heavy:
            mask[r, :] = rng.random(cols) < 0.9
        ii, jj = np.nonzero(mask)
        if order == "random":
            p = rng.permutation(len(ii))
        elif order == "column":
            p = np.lexsort((ii, jj))
        else:
            p = np.arange(len(ii))
        ii, jj = ii[p], jj[p]
        a = rng.uniform(-1.0, 1.0, len(ii))
        cases.append((name, rows, cols, (ii + 1).astype(np.int32), (jj + 1).astype(np.int32), a))

    rand_case("rand_40x50", 40, 50, 0.15)
    rand_case("rand_colmajor_64x64", 64, 64, 0.1, order="column")
    rand_case("empty_rows_33x20", 33, 20, 0.2, empty_rows=(5, 6, 17, 32))   # first row non-empty (ELL pad rule)
    rand_case("ragged_heavy_97x300", 97, 300, 0.02, heavy=(3, 50), empty_rows=(96,))
    rand_case("single_row_1x77", 1, 77, 0.5)
    rand_case("single_col_55x1", 55, 1, 0.6, empty_rows=())
    rand_case("tall_1000x37", 1000, 37, 0.08, empty_rows=(1, 2, 3, 999))
    rand_case("powerlaw_512", 512, 512, 0.01, heavy=(0, 7, 100, 101, 511))
    # make sure the single-col case has a non-empty first row (reference ELL reads column_indices[-1] otherwise)
    name, rows, cols, ii, jj, a = cases[5]
    if 1 not in ii:
        ii = np.concatenate([[1], ii]).astype(np.int32)
        jj = np.concatenate([[1], jj]).astype(np.int32)
        a = np.concatenate([[0.5], a])
        cases[5] = (name, rows, cols, ii, jj, a)
    # 2D 5-point Poisson 30x30, row-major order (the shape of BASELINE config 1)
    n = 30
    ii, jj, aa = [], [], []
    for r in range(n * n):
        gx, gy = r % n, r // n
        for dx, dy, v in ((0, -1, -1.0), (-1, 0, -1.0), (0, 0, 4.0), (1, 0, -1.0), (0, 1, -1.0)):
            x, y = gx + dx, gy + dy
            if 0 <= x < n and 0 <= y < n:
                ii.append(r + 1); jj.append(y * n + x + 1); aa.append(v)
    cases.append(("poisson5_30", n * n, n * n, np.array(ii, np.int32), np.array(jj, np.int32), np.array(aa)))
    return cases


def make_ref_vectors():
    from oracle.oracle import Ref, build
    build(ref=True)
    ref = Ref()
    rng = np.random.default_rng(7)
    out = {}
    names = []
    for name, rows, cols, i, j, a in seeded_cases():
        names.append(name)
        x = rng.uniform(0.5, 1.5, cols)
        y0 = rng.uniform(-1.0, 1.0, rows)
        p = name + "/"
        out[p + "shape"] = np.array([rows, cols, len(i)], np.int64)
        out[p + "i"], out[p + "j"], out[p + "a"], out[p + "x"], out[p + "y0"] = i, j, a, x, y0
        m = ref.from_entries(rows, cols, i, j, a)
        out[p + "max_row_length"] = np.array([m.max_row_length()], np.int64)
        for align in (1, 2, 4):
            A = m.convert("csr", align)
            q = f"{p}csr{align}/"
            out[q + "row_ptr"], out[q + "col"], out[q + "val"] = A.row_ptr, A.column_index, A.value
            out[q + "size"] = np.array([A.size], np.int64)
            out[q + "y_t1"] = m.spmv(x, y0, threads=1)
            out[q + "y_t3"] = m.spmv(x, y0, threads=3)
            if align == 1:
                for T in (1, 2, 3, 8):
                    out[f"{q}part_T{T}"] = np.array(
                        [[m.csr_rows_per_thread(t, T), m.csr_nonzeros_per_thread(t, T)] for t in range(T)], np.int64)
        A = m.convert("coo")
        q = p + "coo/"
        out[q + "row"], out[q + "col"], out[q + "val"] = A.row_index, A.column_index, A.value
        out[q + "size"] = np.array([A.size], np.int64)
        for T in (1, 2, 3):
            out[f"{q}y_t{T}"] = m.spmv(x, y0, threads=T)
        m.convert("coo-atomic")
        out[q + "y_atomic_t1"] = m.spmv(x, y0, threads=1)
        for skip in (0, 1):
            A = m.convert("ell", skip)
            q = f"{p}ell{skip}/"
            out[q + "row_length"] = np.array([A.row_length], np.int64)
            out[q + "col"], out[q + "val"] = A.column_index, A.value
            out[q + "size"] = np.array([A.size], np.int64)
            out[q + "y_t1"] = m.spmv(x, y0, threads=1)
            out[q + "y_t2"] = m.spmv(x, y0, threads=2)
            A = m.convert("hybrid", skip)
            q = f"{p}hyb{skip}/"
            out[q + "dims"] = np.array([A.ell_row_length, A.num_ell_entries, A.num_coo_entries], np.int64)
            out[q + "ell_col"], out[q + "ell_val"] = A.ell_column_index, A.ell_value
            out[q + "coo_row"], out[q + "coo_col"], out[q + "coo_val"] = A.coo_row_index, A.coo_column_index, A.coo_value
            out[q + "size_reference"] = np.array([A.size_reference], np.int64)
            for T in (1, 2, 3):
                out[f"{q}y_t{T}"] = m.spmv(x, y0, threads=T)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("ref_vectors.npz:", len(names), "cases,", len(out), "arrays")


if __name__ == "__main__":
    extract_poisson2d()
    make_ref_vectors()
