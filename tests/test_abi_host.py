"""CPU-side checks of the C ABI (no GPU needed, no compute calls).

* libspmvb200.so loads and exports every function include/spmv_b200.h declares.
* The host-side Matrix Market reader behind the ABI agrees with the oracle (and hence
  with the reference) on real/complex/integer/pattern files, .gz, .tar, .tar.gz, and the
  two sort orders -- the cases of the reference's test/test_matrix-market.cpp.
* Without a CUDA device every compute entry point fails loudly (no CPU fallback).
"""
import ctypes as C
import gzip
import io
import os
import re
import tarfile

import numpy as np
import pytest

import spmv_cache_trace_b200 as sp
from spmv_cache_trace_b200 import _abi, matrix_error, matrix_market

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "spmv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spmvb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_abi.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 55
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_abi.SIGNATURES) == names  # the Python binding covers the header exactly


def test_version_and_error_text():
    L = _abi.lib()
    assert L.spmvb200_version() == 210
    assert isinstance(L.spmvb200_last_error(), bytes)


KAT = ("%%MatrixMarket matrix coordinate real general\n% Test matrix\n4 5 7\n"
       "1 1 1.0\n1 2 2.0\n2 2 1.0\n3 3 3.0\n4 1 -1.0\n4 4 2.0\n4 5 1.0\n")


def test_fromStream_real(oracle):
    m = matrix_market.fromStream(io.StringIO(KAT))
    assert (m.rows, m.columns, m.num_entries, m.field, m.symmetry, m.format) == (4, 5, 7, 0, 0, 0)
    o = oracle.mm_parse(KAT)
    assert np.array_equal(m.row_indices(), o.i) and np.array_equal(m.column_indices(), o.j)
    assert np.array_equal(m.values_real(), o.a)
    assert m.max_row_length() == 3
    assert m.row_lengths().tolist() == [2, 1, 1, 3]


@pytest.mark.parametrize("text", [
    "%%MatrixMarket matrix coordinate complex general\n2 2 2\n1 1 1.5 -2.0\n2 2 3.0 4.0\n",
    "%%MatrixMarket MATRIX Coordinate Integer Symmetric\n% c\n%c2\n3 3 2\n2 1 7\n3 3 -2\n",
    "%%MatrixMarket matrix coordinate pattern general\n2 3 3\n1 1\n1 3\n2 2\n",
    "%%MatrixMarket matrix coordinate real general\n3 3 4\n1 1 1e-3 2 2\n-.5\n3 3 +4.25 3 1 1E2\n",
])
def test_fromStream_fields_match_oracle(oracle, text):
    m = matrix_market.fromStream(text)
    o = oracle.mm_parse(text)
    assert (m.rows, m.columns, m.num_entries, m.field, m.symmetry) == (o.rows, o.columns, o.num_entries, o.field, o.symmetry)
    assert np.array_equal(m.row_indices(), o.i) and np.array_equal(m.column_indices(), o.j)
    assert np.array_equal(m.values_real(), o.a)


def test_poisson2D_parse_matches_oracle(oracle, poisson2d):
    text, _, _ = poisson2d
    m = matrix_market.fromStream(text)
    o = oracle.mm_parse(text)
    assert np.array_equal(m.row_indices(), o.i) and np.array_equal(m.column_indices(), o.j)
    assert np.array_equal(m.values_real(), o.a)  # bit-exact text -> double


@pytest.mark.parametrize("bad", [
    "%MatrixMarket matrix coordinate real general\n1 1 0\n",
    "%%MatrixMarket vector coordinate real general\n1 1 0\n",
    "%%MatrixMarket matrix coordinate quaternion general\n1 1 0\n",
    "%%MatrixMarket matrix coordinate real weird\n1 1 0\n",
    "%%MatrixMarket matrix coordinate real general\n99999999999 1 0\n",
    "%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n",
    "",
])
def test_fromStream_errors(bad):
    with pytest.raises(matrix_error):
        matrix_market.fromStream(bad)


def test_sort_orders_match_oracle(oracle, poisson2d):
    text, _, _ = poisson2d
    m = matrix_market.fromStream(text)
    o = oracle.mm_parse(text)
    r = matrix_market.sort_matrix_row_major(m)
    i, j, a = oracle.sort_row_major(o.i, o.j, o.a)
    assert np.array_equal(r.row_indices(), i) and np.array_equal(r.column_indices(), j) and np.array_equal(r.values_real(), a)
    c = matrix_market.sort_matrix_column_major(r)
    i, j, a = oracle.sort_column_major(i, j, a)
    assert np.array_equal(c.row_indices(), i) and np.array_equal(c.column_indices(), j) and np.array_equal(c.values_real(), a)


def test_load_matrix_mtx_gz_tar(tmp_path, oracle):
    o = oracle.mm_parse(KAT)
    p = tmp_path / "kat.mtx"
    p.write_text(KAT)
    gz = tmp_path / "kat.mtx.gz"
    with gzip.open(gz, "wb") as f:
        f.write(KAT.encode())
    d = tmp_path / "kat"
    d.mkdir()
    (d / "kat.mtx").write_text(KAT)
    (d / "other.txt").write_text("x" * 1000)
    for name, mode in (("kat.tar.gz", "w:gz"), ("kat.tgz", "w:gz")):
        with tarfile.open(tmp_path / name, mode) as t:
            t.add(d / "other.txt", arcname="kat/other.txt")
            t.add(d / "kat.mtx", arcname="kat/kat.mtx")
    for path in (p, gz, tmp_path / "kat.tar.gz", tmp_path / "kat.tgz"):
        m = matrix_market.load_matrix(str(path))
        assert (m.rows, m.columns, m.num_entries) == (4, 5, 7), path
        assert np.array_equal(m.values_real(), o.a), path
    with pytest.raises(matrix_error):
        matrix_market.load_matrix(str(tmp_path / "missing.mtx"))
    with pytest.raises(matrix_error):
        matrix_market.load_matrix(str(p) + "__RCM")


def test_partition_rows_ref_matches_oracle(oracle):
    for rows, P in ((10, 4), (3, 8), (1000000, 2), (134217728, 8), (0, 3)):
        assert np.array_equal(sp.partition.rows_ref(rows, P), oracle.partition_rows_ref(rows, P))


def test_compute_fails_loudly_without_gpu():
    if sp.device_count() > 0:
        pytest.skip("a CUDA device is present")
    m = matrix_market.fromStream(KAT)
    with pytest.raises(matrix_error) as e:
        sp.csr_matrix.from_matrix_market(m)
    assert e.value.status == 5  # SPMVB200_ERR_CUDA
    with pytest.raises(matrix_error):
        sp.generators.stencil(sp.STENCIL_2D5, 8, 8)


def _big_text(n, seed=0, fmt="%d %d %.17g"):
    rng = np.random.default_rng(seed)
    rows, cols = 700000, 650000
    i = rng.integers(1, rows + 1, n).astype(np.int32)
    j = rng.integers(1, cols + 1, n).astype(np.int32)
    a = rng.uniform(-1e3, 1e3, n)
    lines = [fmt % (ii, jj, aa) for ii, jj, aa in zip(i.tolist(), j.tolist(), a.tolist())]
    head = "%%MatrixMarket matrix coordinate real general\n% a comment\n" + f"{rows} {cols} {n}\n"
    return head, lines, i, j, a


def test_parallel_parse_matches_the_sequential_reader():
    """Files with one record per line are parsed by several threads (>= 4 MB of text per thread); anything
    irregular falls back to the sequential tokenizer, with the same results and the same errors."""
    import spmv_cache_trace_b200 as sp
    head, lines, i, j, a = _big_text(400000)
    text = head + "\n".join(lines) + "\n"
    assert len(text) > 9 << 20
    mm = sp.matrix_market.fromStream(text)
    assert np.array_equal(mm.row_indices(), i) and np.array_equal(mm.column_indices(), j)
    assert np.array_equal(mm.values_real(), a)  # %.17g round-trips
    os.environ["SPMVB200_PARSE_THREADS"] = "1"  # the sequential reader on the same text
    try:
        m1 = sp.matrix_market.fromStream(text)
    finally:
        del os.environ["SPMVB200_PARSE_THREADS"]
    assert np.array_equal(m1.row_indices(), i) and np.array_equal(m1.values_real(), a)
    # no trailing newline, CRLF line ends, tabs
    m2 = sp.matrix_market.fromStream(head + "\r\n".join(l.replace(" ", "\t", 1) for l in lines))
    assert np.array_equal(m2.column_indices(), j) and np.array_equal(m2.values_real(), a)
    # irregular layouts: two records on one line / a record split over two lines / a blank line
    two = lines[:]
    two[1000] = two[1000] + " " + two.pop(1001)
    m3 = sp.matrix_market.fromStream(head + "\n".join(two) + "\n")
    assert np.array_equal(m3.row_indices(), i) and np.array_equal(m3.values_real(), a)
    split = lines[:]
    first, rest = split[2000].split(" ", 1)
    split[2000] = first + "\n" + rest
    m4 = sp.matrix_market.fromStream(head + "\n".join(split) + "\n")
    assert np.array_equal(m4.column_indices(), j)
    blank = lines[:]
    blank.insert(3000, "")
    m5 = sp.matrix_market.fromStream(head + "\n".join(blank) + "\n")
    assert np.array_equal(m5.row_indices(), i)
    # errors read like the sequential reader's: one record short, an index outside the matrix
    with pytest.raises(sp.matrix_error, match="Expected 400000 entries, got 399999 entries"):
        sp.matrix_market.fromStream(head + "\n".join(lines[:-1]) + "\n")
    bad = lines[:]
    bad[123456] = "700001 1 1.0"
    with pytest.raises(sp.matrix_error, match="index outside the matrix in entry 123457"):
        sp.matrix_market.fromStream(head + "\n".join(bad) + "\n")
