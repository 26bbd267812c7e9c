"""Live cross-check: oracle (our C restatement) vs the reference library itself.

Runs only where oracle/_ref/libspmvref.so exists (built in the container from
/root/reference by oracle/Makefile; it travels to the GPU box as a built .so).
Fresh random matrices each parametrisation -- beyond the committed golden vectors.
"""
import numpy as np
import pytest


def random_entries(rng, rows, cols, nnz, first_row_nonempty=True):
    flat = rng.choice(rows * cols, size=min(nnz, rows * cols), replace=False)
    i, j = flat // cols, flat % cols
    if first_row_nonempty and 0 not in i:
        i = np.concatenate([[0], i]); j = np.concatenate([[0], j])
    a = rng.uniform(-2.0, 2.0, len(i))
    return (i + 1).astype(np.int32), (j + 1).astype(np.int32), a


@pytest.mark.parametrize("seed,rows,cols,nnz", [(1, 17, 23, 60), (2, 300, 300, 2500), (3, 1, 9, 5),
                                                (4, 2000, 50, 9000), (5, 64, 4096, 5000)])
def test_conversions_and_products_match_reference(oracle, ref, seed, rows, cols, nnz):
    rng = np.random.default_rng(seed)
    i, j, a = random_entries(rng, rows, cols, nnz)
    x = rng.uniform(-1, 1, cols)
    y0 = rng.uniform(-1, 1, rows)
    m = ref.from_entries(rows, cols, i, j, a)

    R = m.convert("csr", 1)
    O = oracle.csr(rows, cols, i, j, a, 1)
    assert np.array_equal(R.row_ptr, O.row_ptr) and np.array_equal(R.column_index, O.column_index)
    assert np.array_equal(R.value, O.value) and R.size == O.size
    assert np.array_equal(m.spmv(x, y0, threads=2), oracle.csr_spmv(O, x, y0))

    R = m.convert("coo")
    O = oracle.coo(rows, cols, i, j, a)
    assert np.array_equal(R.row_index, O.row_index) and np.array_equal(R.value, O.value)
    for T in (1, 2, 4):
        assert np.array_equal(m.spmv(x, y0, threads=T), oracle.coo_spmv(O, x, y0, num_threads=T))

    for skip in (0, 1):
        R = m.convert("ell", skip)
        O = oracle.ell(rows, cols, i, j, a, skip)
        assert R.row_length == O.row_length
        assert np.array_equal(R.column_index, O.column_index) and np.array_equal(R.value, O.value)
        assert np.array_equal(m.spmv(x, y0, threads=3), oracle.ell_spmv(O, x, y0))
        R = m.convert("hybrid", skip)
        O = oracle.hyb(rows, cols, i, j, a, skip)
        assert (R.ell_row_length, R.num_ell_entries, R.num_coo_entries) == (O.ell_row_length, O.num_ell_entries, O.num_coo_entries)
        assert np.array_equal(R.ell_column_index, O.ell_column_index) and np.array_equal(R.ell_value, O.ell_value)
        assert np.array_equal(R.coo_row_index, O.coo_row_index) and np.array_equal(R.coo_value, O.coo_value)
        for T in (1, 2, 5):
            assert np.array_equal(m.spmv(x, y0, threads=T), oracle.hyb_spmv(O, x, y0, num_threads=T))


def test_parser_matches_reference(oracle, ref, poisson2d):
    text, _, _ = poisson2d
    ri, rj, ra = ref.from_text(text).entries()
    mm = oracle.mm_parse(text)
    assert np.array_equal(ri, mm.i) and np.array_equal(rj, mm.j) and np.array_equal(ra, mm.a)


def test_reference_timer_runs(ref):
    """The profile_kernel_run-protocol timer used as bench.py's CPU baseline."""
    n = 64
    r = np.arange(n * n)
    i = np.concatenate([r, r[:-1]]) + 1
    j = np.concatenate([r, r[1:]]) + 1
    m = ref.from_entries(n * n, n * n, i.astype(np.int32), j.astype(np.int32), np.ones(len(i)))
    for fmt in ("csr", "coo", "ell", "hybrid"):
        m.convert(fmt)
        ns = m.time(threads=2, reps=3, pin=False)
        assert ns.shape == (3,) and np.all(ns > 0)
