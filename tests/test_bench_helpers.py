"""CPU tests of the arithmetic bench.py uses to CHECK results before it times anything (a wrong checker is as bad as a
wrong kernel: round 2 found one that cut a row short when empty rows followed it)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle.generators_ref import rmat_entries, stencil_entries  # noqa: E402


@pytest.mark.parametrize("kind,dims", [(0, (9, 7, 1)), (1, (6, 5, 4)), (2, (5, 6, 7)), (2, (1, 4, 3))])
def test_stencil_closed_form_matches_the_oracle(oracle, kind, dims):
    nx, ny, nz = dims
    N = nx * ny * nz
    i, j, a = stencil_entries(kind, nx, ny, nz)
    O = oracle.csr(N, N, i, j, a)
    x = bench.x_pattern(np.arange(N))
    want = oracle.csr_spmv(O, x)
    rows = np.arange(N)
    assert np.array_equal(bench.stencil_rows_closed_form(kind, nx, ny, nz, rows), want)
    # a rank's slice: parity_stencil samples rows of [row_begin, row_begin + len) and compares exactly
    b, e = N // 3, N - 2
    ok = bench.parity_stencil(kind, nx, ny, nz, b, want[b:e])
    assert ok["ok"] and ok["bad_rows"] == 0 and ok["rows_checked"] == e - b
    broken = want[b:e].copy()
    broken[5] += 0.125
    bad = bench.parity_stencil(kind, nx, ny, nz, b, broken)
    assert not bad["ok"] and bad["bad_rows"] == 1


def test_sample_rows_cover_both_ends():
    rows = bench.sample_rows(1000, 9_000_000, 300, 5000, seed=1)
    assert rows[0] == 1000 and rows[-1] == 8_999_999 and np.all(np.diff(rows) > 0)
    assert np.array_equal(rows[:300], np.arange(1000, 1300)) and np.array_equal(rows[-300:], np.arange(8_999_700, 9_000_000))
    assert 300 * 2 + 4000 < len(rows) <= 300 * 2 + 5000
    assert np.array_equal(bench.sample_rows(10, 50, 30, 100, seed=1), np.arange(10, 50))  # small range: every row


def test_parity_csr_rows_with_empty_rows(oracle):
    """R-MAT row blocks are full of empty rows, also at the end of a sample."""
    scale, ef, seed = 10, 4, 0x5EED0004
    n = 1 << scale
    r, c, v = rmat_entries(scale, ef, seed)
    O = oracle.csr(n, n, r + 1, c + 1, v)
    rp = np.asarray(O.row_ptr, np.int64)
    assert np.any(np.diff(rp) == 0)
    x = bench.x_pattern(np.arange(n))
    y = oracle.csr_spmv(O, x)
    for b, e in ((0, n), (100, 613), (n - 300, n)):
        while rp[e] > rp[e - 1]:  # make the sample END in an empty row
            e -= 1
        cols, vals = np.asarray(O.column_index)[rp[b]:rp[e]], np.asarray(O.value)[rp[b]:rp[e]]
        bad, rel = bench.parity_csr_rows(rp[b:e + 1] - rp[b], cols, vals, y[b:e])
        assert bad == 0 and rel <= 1e-13
        wrong = y[b:e].copy()
        k = int(np.nonzero(np.diff(rp[b:e + 1]) > 0)[0][-1])  # the last non-empty row of the sample
        wrong[k] *= 1.0 + 1e-9
        bad, rel = bench.parity_csr_rows(rp[b:e + 1] - rp[b], cols, vals, wrong)
        assert bad == 1
    assert bench.parity_csr_rows(np.zeros(5, np.int64), np.zeros(0, np.int64), np.zeros(0), np.zeros(4)) == (0, 0.0)


def test_l2_cold_copies_rule():
    class Info:
        x_size, y_size, device_bytes = 8_000_000, 8_000_000, 100_000_000

    class A:
        info = Info()

        @staticmethod
        def algorithmic_bytes():
            return 80_000_000

    l2 = 132_644_864
    copies = bench.l2_cold_copies(A, l2, free_bytes=100 << 30)
    assert copies * 16_000_000 >= 2 * l2 and copies * 80_000_000 >= 3 * l2  # x + y AND the matrix data leave L2
    A.algorithmic_bytes = staticmethod(lambda: 46_000_000_000)
    assert bench.l2_cold_copies(A, l2, free_bytes=100 << 30) == 1
