"""ctypes front-end to the CPU oracle (TEST INFRASTRUCTURE ONLY).

``Oracle``  -> oracle/_build/liboracle.so  (our plain-C restatement, spmv_oracle.c)
``Ref``     -> oracle/_ref/libspmvref.so   (the unmodified reference library + ref_shim.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (spmv_cache_trace_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libspmvref.so")

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


def build(ref: bool = True) -> None:
    """Compile the checkers (gcc only; the reference part needs /root/reference)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def _ptr(a, t):
    return a.ctypes.data_as(t)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _take(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class _MM(C.Structure):
    _fields_ = [("rows", C.c_int32), ("columns", C.c_int32), ("num_entries", C.c_int32),
                ("format", C.c_int32), ("field", C.c_int32), ("symmetry", C.c_int32),
                ("num_comments", C.c_int32), ("i", _i32p), ("j", _i32p), ("a", _f64p)]


class _CSR(C.Structure):
    _fields_ = [("rows", C.c_int32), ("columns", C.c_int32), ("num_entries", C.c_int32),
                ("row_alignment", C.c_int32), ("row_ptr", _i32p), ("column_index", _i32p),
                ("value", _f64p)]


class _COO(C.Structure):
    _fields_ = [("rows", C.c_int32), ("columns", C.c_int32), ("num_entries", C.c_int32),
                ("row_index", _i32p), ("column_index", _i32p), ("value", _f64p)]


class _ELL(C.Structure):
    _fields_ = [("rows", C.c_int32), ("columns", C.c_int32), ("num_entries", C.c_int32),
                ("row_length", C.c_int32), ("skip_padding", C.c_int32),
                ("column_index", _i32p), ("value", _f64p)]


class _HYB(C.Structure):
    _fields_ = [("rows", C.c_int32), ("columns", C.c_int32), ("num_entries", C.c_int32),
                ("ell_row_length", C.c_int32), ("num_ell_entries", C.c_int32),
                ("ell_skip_padding", C.c_int32), ("ell_column_index", _i32p), ("ell_value", _f64p),
                ("num_coo_entries", C.c_int32), ("coo_row_index", _i32p),
                ("coo_column_index", _i32p), ("coo_value", _f64p)]


@dataclass
class MM:
    rows: int
    columns: int
    num_entries: int
    field: int
    symmetry: int
    format: int
    i: np.ndarray  # 1-based
    j: np.ndarray
    a: np.ndarray


@dataclass
class Csr:
    rows: int
    columns: int
    num_entries: int
    row_alignment: int
    row_ptr: np.ndarray
    column_index: np.ndarray
    value: np.ndarray
    size: int = 0


@dataclass
class Coo:
    rows: int
    columns: int
    num_entries: int
    row_index: np.ndarray
    column_index: np.ndarray
    value: np.ndarray
    size: int = 0


@dataclass
class Ell:
    rows: int
    columns: int
    num_entries: int
    row_length: int
    skip_padding: int
    column_index: np.ndarray  # row-major rows*row_length
    value: np.ndarray
    size: int = 0
    first_row_empty: int = 0


@dataclass
class Hyb:
    rows: int
    columns: int
    num_entries: int
    ell_row_length: int
    num_ell_entries: int
    ell_skip_padding: int
    ell_column_index: np.ndarray
    ell_value: np.ndarray
    num_coo_entries: int
    coo_row_index: np.ndarray
    coo_column_index: np.ndarray
    coo_value: np.ndarray
    size: int = 0
    size_reference: int = 0


class OracleError(RuntimeError):
    pass


class Oracle:
    """Plain-C restatement (oracle/spmv_oracle.c)."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.orc_last_error.restype = C.c_char_p
        for name in ("orc_csr_size", "orc_coo_size", "orc_ell_size", "orc_hyb_size",
                     "orc_hyb_size_reference"):
            getattr(L, name).restype = C.c_size_t
        L.orc_max_row_length.restype = C.c_int32
        L.orc_hyb_ell_row_length.restype = C.c_int32
        for name in ("orc_csr_rows_per_thread", "orc_csr_start_row", "orc_csr_nonzeros_per_thread"):
            getattr(L, name).restype = C.c_int32

    def _check(self, rc):
        if rc != 0:
            raise OracleError(self.lib.orc_last_error().decode())

    # -- matrix market ---------------------------------------------------
    def mm_parse(self, text) -> MM:
        if isinstance(text, str):
            text = text.encode()
        m = _MM()
        self._check(self.lib.orc_mm_parse(C.c_char_p(text), C.c_size_t(len(text)), C.byref(m)))
        n = m.num_entries
        out = MM(m.rows, m.columns, n, m.field, m.symmetry, m.format,
                 _take(m.i, n, np.int32), _take(m.j, n, np.int32), _take(m.a, n, np.float64))
        self.lib.orc_mm_free(C.byref(m))
        return out

    def sort_row_major(self, i, j, a):
        i, j, a = _i32(i).copy(), _i32(j).copy(), _f64(a).copy()
        self.lib.orc_sort_row_major(C.c_int32(len(i)), _ptr(i, _i32p), _ptr(j, _i32p), _ptr(a, _f64p))
        return i, j, a

    def sort_column_major(self, i, j, a):
        i, j, a = _i32(i).copy(), _i32(j).copy(), _f64(a).copy()
        self.lib.orc_sort_column_major(C.c_int32(len(i)), _ptr(i, _i32p), _ptr(j, _i32p), _ptr(a, _f64p))
        return i, j, a

    def row_lengths(self, rows, i):
        i = _i32(i)
        out = np.zeros(rows, dtype=np.int32)
        self.lib.orc_row_lengths(C.c_int32(rows), C.c_int32(len(i)), _ptr(i, _i32p), _ptr(out, _i32p))
        return out

    def hyb_ell_row_length(self, row_lengths):
        rl = _i32(row_lengths)
        return int(self.lib.orc_hyb_ell_row_length(C.c_int32(len(rl)), _ptr(rl, _i32p)))

    # -- conversions -----------------------------------------------------
    def csr(self, rows, cols, i, j, a, row_alignment=1) -> Csr:
        i, j, a = _i32(i), _i32(j), _f64(a)
        m = _CSR()
        self._check(self.lib.orc_csr_from_entries(C.c_int32(rows), C.c_int32(cols), C.c_int32(len(i)),
                                                  _ptr(i, _i32p), _ptr(j, _i32p), _ptr(a, _f64p),
                                                  C.c_int32(row_alignment), C.byref(m)))
        rp = _take(m.row_ptr, rows + 1, np.int32)
        st = int(rp[rows]) if rows >= 0 else 0
        out = Csr(rows, cols, len(i), row_alignment, rp, _take(m.column_index, st, np.int32),
                  _take(m.value, st, np.float64), int(self.lib.orc_csr_size(C.byref(m))))
        self.lib.orc_csr_free(C.byref(m))
        return out

    def coo(self, rows, cols, i, j, a) -> Coo:
        i, j, a = _i32(i), _i32(j), _f64(a)
        m = _COO()
        n = len(i)
        self._check(self.lib.orc_coo_from_entries(C.c_int32(rows), C.c_int32(cols), C.c_int32(n),
                                                  _ptr(i, _i32p), _ptr(j, _i32p), _ptr(a, _f64p), C.byref(m)))
        out = Coo(rows, cols, n, _take(m.row_index, n, np.int32), _take(m.column_index, n, np.int32),
                  _take(m.value, n, np.float64), int(self.lib.orc_coo_size(C.byref(m))))
        self.lib.orc_coo_free(C.byref(m))
        return out

    def ell(self, rows, cols, i, j, a, skip_padding=False) -> Ell:
        i, j, a = _i32(i), _i32(j), _f64(a)
        m = _ELL()
        fre = C.c_int32(0)
        self._check(self.lib.orc_ell_from_entries(C.c_int32(rows), C.c_int32(cols), C.c_int32(len(i)),
                                                  _ptr(i, _i32p), _ptr(j, _i32p), _ptr(a, _f64p),
                                                  C.c_int32(int(skip_padding)), C.byref(m), C.byref(fre)))
        slots = rows * m.row_length
        out = Ell(rows, cols, len(i), m.row_length, int(skip_padding),
                  _take(m.column_index, slots, np.int32), _take(m.value, slots, np.float64),
                  int(self.lib.orc_ell_size(C.byref(m))), fre.value)
        self.lib.orc_ell_free(C.byref(m))
        return out

    def hyb(self, rows, cols, i, j, a, skip_padding=False) -> Hyb:
        i, j, a = _i32(i), _i32(j), _f64(a)
        m = _HYB()
        self._check(self.lib.orc_hyb_from_entries(C.c_int32(rows), C.c_int32(cols), C.c_int32(len(i)),
                                                  _ptr(i, _i32p), _ptr(j, _i32p), _ptr(a, _f64p),
                                                  C.c_int32(int(skip_padding)), C.byref(m)))
        ne, nc = m.num_ell_entries, m.num_coo_entries
        out = Hyb(rows, cols, len(i), m.ell_row_length, ne, int(skip_padding),
                  _take(m.ell_column_index, ne, np.int32), _take(m.ell_value, ne, np.float64),
                  nc, _take(m.coo_row_index, nc, np.int32), _take(m.coo_column_index, nc, np.int32),
                  _take(m.coo_value, nc, np.float64),
                  int(self.lib.orc_hyb_size(C.byref(m))), int(self.lib.orc_hyb_size_reference(C.byref(m))))
        self.lib.orc_hyb_free(C.byref(m))
        return out

    # -- struct views over numpy arrays (no copies) ------------------------
    @staticmethod
    def _csr_view(A: Csr):
        A.row_ptr, A.column_index, A.value = _i32(A.row_ptr), _i32(A.column_index), _f64(A.value)
        return _CSR(A.rows, A.columns, A.num_entries, A.row_alignment, _ptr(A.row_ptr, _i32p),
                    _ptr(A.column_index, _i32p), _ptr(A.value, _f64p))

    @staticmethod
    def _coo_view(A: Coo):
        A.row_index, A.column_index, A.value = _i32(A.row_index), _i32(A.column_index), _f64(A.value)
        return _COO(A.rows, A.columns, A.num_entries, _ptr(A.row_index, _i32p),
                    _ptr(A.column_index, _i32p), _ptr(A.value, _f64p))

    @staticmethod
    def _ell_view(A: Ell):
        A.column_index, A.value = _i32(A.column_index), _f64(A.value)
        return _ELL(A.rows, A.columns, A.num_entries, A.row_length, A.skip_padding,
                    _ptr(A.column_index, _i32p), _ptr(A.value, _f64p))

    @staticmethod
    def _hyb_view(A: Hyb):
        A.ell_column_index, A.ell_value = _i32(A.ell_column_index), _f64(A.ell_value)
        A.coo_row_index, A.coo_column_index = _i32(A.coo_row_index), _i32(A.coo_column_index)
        A.coo_value = _f64(A.coo_value)
        return _HYB(A.rows, A.columns, A.num_entries, A.ell_row_length, A.num_ell_entries,
                    A.ell_skip_padding, _ptr(A.ell_column_index, _i32p), _ptr(A.ell_value, _f64p),
                    A.num_coo_entries, _ptr(A.coo_row_index, _i32p), _ptr(A.coo_column_index, _i32p),
                    _ptr(A.coo_value, _f64p))

    # -- y += A x ----------------------------------------------------------
    def csr_spmv(self, A: Csr, x, y=None, threads: int = 0):
        x = _f64(x)
        y = np.zeros(A.rows) if y is None else _f64(y).copy()
        v = self._csr_view(A)
        if threads > 0:
            self.lib.orc_csr_spmv_omp(C.byref(v), _ptr(x, _f64p), _ptr(y, _f64p), C.c_int(threads))
        else:
            self.lib.orc_csr_spmv(C.byref(v), _ptr(x, _f64p), _ptr(y, _f64p))
        return y

    def csr_abs_rowsum(self, A: Csr, x):
        x = _f64(x)
        out = np.zeros(A.rows)
        v = self._csr_view(A)
        self.lib.orc_csr_abs_rowsum(C.byref(v), _ptr(x, _f64p), _ptr(out, _f64p))
        return out

    def coo_spmv(self, A: Coo, x, y=None, num_threads: int = 1, workspace=None, chunk_size: int = 0,
                 omp: bool = False):
        x = _f64(x)
        y = np.zeros(A.rows) if y is None else _f64(y).copy()
        ws = np.zeros(max(1, num_threads * A.rows)) if workspace is None else workspace
        v = self._coo_view(A)
        if omp:
            self.lib.orc_coo_spmv_omp(C.byref(v), _ptr(x, _f64p), _ptr(y, _f64p), _ptr(ws, _f64p),
                                      C.c_int(num_threads))
        else:
            self.lib.orc_coo_spmv(C.c_int(num_threads), C.byref(v), _ptr(x, _f64p), _ptr(y, _f64p),
                                  _ptr(ws, _f64p), C.c_int32(chunk_size))
        return y

    def coo_spmv_atomic(self, A: Coo, x, y=None):
        x = _f64(x)
        y = np.zeros(A.rows) if y is None else _f64(y).copy()
        v = self._coo_view(A)
        self.lib.orc_coo_spmv_atomic(C.byref(v), _ptr(x, _f64p), _ptr(y, _f64p))
        return y

    def ell_spmv(self, A: Ell, x, y=None, threads: int = 0):
        x = _f64(x)
        y = np.zeros(A.rows) if y is None else _f64(y).copy()
        v = self._ell_view(A)
        if threads > 0:
            self.lib.orc_ell_spmv_omp(C.byref(v), _ptr(x, _f64p), _ptr(y, _f64p), C.c_int(threads))
        else:
            self.lib.orc_ell_spmv(C.byref(v), _ptr(x, _f64p), _ptr(y, _f64p))
        return y

    def hyb_spmv(self, A: Hyb, x, y=None, num_threads: int = 1, workspace=None, omp: bool = False):
        x = _f64(x)
        y = np.zeros(A.rows) if y is None else _f64(y).copy()
        ws = np.zeros(max(1, num_threads * A.rows)) if workspace is None else workspace
        v = self._hyb_view(A)
        if omp:
            self.lib.orc_hyb_spmv_omp(C.byref(v), _ptr(x, _f64p), _ptr(y, _f64p), _ptr(ws, _f64p),
                                      C.c_int(num_threads))
        else:
            self.lib.orc_hyb_spmv(C.c_int(num_threads), C.byref(v), _ptr(x, _f64p), _ptr(y, _f64p),
                                  _ptr(ws, _f64p))
        return y

    # -- partitions ----------------------------------------------------------
    def csr_rows_per_thread(self, rows, t, T):
        return int(self.lib.orc_csr_rows_per_thread(C.c_int32(rows), C.c_int(t), C.c_int(T)))

    def csr_start_row(self, rows, t, T):
        return int(self.lib.orc_csr_start_row(C.c_int32(rows), C.c_int(t), C.c_int(T)))

    def csr_nonzeros_per_thread(self, row_ptr, rows, t, T):
        rp = _i32(row_ptr)
        return int(self.lib.orc_csr_nonzeros_per_thread(_ptr(rp, _i32p), C.c_int32(rows), C.c_int(t), C.c_int(T)))

    # -- R-MAT at full size (C twin of the device generator; numpy twin: generators_ref.rmat_entries) ---------
    def rmat_csr(self, scale: int, edge_factor: int, seed: int, a=0.57, b=0.19, c=0.19) -> Csr:
        """The R-MAT matrix as the CSR the reference's converter produces from its sorted, de-duplicated entries
        (row_ptr = prefix counts, columns ascending inside a row).  Built directly: the converter's sort would be the
        identity on this input (checked against orc_csr_from_entries at small sizes in tests/test_oracle_golden.py)."""
        m = edge_factor << scale
        n = 1 << scale
        keys = np.empty(m, dtype=np.uint64)
        self.lib.orc_rmat_keys(C.c_int(scale), C.c_uint64(seed), C.c_double(a), C.c_double(b), C.c_double(c),
                               C.c_uint64(0), C.c_uint64(m), keys.ctypes.data_as(C.POINTER(C.c_uint64)))
        keys.sort()
        keep = np.empty(m, dtype=bool)
        keep[0] = True
        np.not_equal(keys[1:], keys[:-1], out=keep[1:])
        keys = keys[keep]
        del keep
        nnz = int(keys.shape[0])
        col = np.empty(nnz, dtype=np.int32)
        val = np.empty(nnz, dtype=np.float64)
        self.lib.orc_rmat_unpack(C.c_int64(nnz), keys.ctypes.data_as(C.POINTER(C.c_uint64)), _ptr(col, _i32p), _ptr(val, _f64p))
        rp = np.searchsorted(keys, np.arange(n + 1, dtype=np.uint64) << np.uint64(32)).astype(np.int32)
        return Csr(n, n, nnz, 1, rp, col, val, 12 * nnz + 4 * (n + 1))

    def partition_rows_ref(self, rows, P):
        out = np.zeros(P + 1, dtype=np.int64)
        self.lib.orc_partition_rows_ref(C.c_int64(rows), C.c_int(P), out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out

    def partition_rows_nnz(self, row_ptr, P):
        rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
        out = np.zeros(P + 1, dtype=np.int64)
        self.lib.orc_partition_rows_nnz(C.c_int64(len(rp) - 1), rp.ctypes.data_as(C.POINTER(C.c_int64)),
                                        C.c_int(P), out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out

    def order_from_parts(self, part, nparts: int):
        part = _i32(part)
        out = np.zeros(max(part.size, 1), dtype=np.int32)
        self._check(self.lib.orc_order_from_parts(C.c_int32(part.size), C.c_int32(nparts), _ptr(part, _i32p), _ptr(out, _i32p)))
        return out[: part.size]

    def partition_rows_weighted(self, row_ptr, P, row_weight: float):
        rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
        out = np.zeros(P + 1, dtype=np.int64)
        self.lib.orc_partition_rows_weighted(C.c_int64(len(rp) - 1), rp.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int(P),
                                             C.c_int64(int(round(1024 * row_weight))), out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out


class RefError(RuntimeError):
    pass


class Ref:
    """The unmodified reference matrix library (oracle/_ref/libspmvref.so)."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def __init__(self, path: str = REF_SO):
        self.lib = L = C.CDLL(path)
        L.ref_last_error.restype = C.c_char_p
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_mm_max_row_length.restype = C.c_int32
        L.ref_mm_max_row_length.argtypes = [C.c_void_p]
        L.ref_csr_rows_per_thread.restype = C.c_int32
        L.ref_csr_nonzeros_per_thread.restype = C.c_int32
        L.ref_csr_rows_per_thread.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_csr_nonzeros_per_thread.argtypes = [C.c_void_p, C.c_int, C.c_int]

    def _check(self, rc):
        if rc != 0:
            raise RefError(self.lib.ref_last_error().decode())

    def from_text(self, text) -> "RefMatrix":
        if isinstance(text, str):
            text = text.encode()
        h = C.c_void_p()
        self._check(self.lib.ref_mm_from_text(C.c_char_p(text), C.c_size_t(len(text)), C.byref(h)))
        return RefMatrix(self, h)

    def load(self, path: str) -> "RefMatrix":
        h = C.c_void_p()
        self._check(self.lib.ref_mm_load(C.c_char_p(path.encode()), C.byref(h)))
        return RefMatrix(self, h)

    def from_entries(self, rows, cols, i, j, a) -> "RefMatrix":
        i, j, a = _i32(i), _i32(j), _f64(a)
        h = C.c_void_p()
        self._check(self.lib.ref_mm_from_entries(C.c_int32(rows), C.c_int32(cols), C.c_int32(len(i)),
                                                 _ptr(i, _i32p), _ptr(j, _i32p), _ptr(a, _f64p), C.byref(h)))
        return RefMatrix(self, h)


class RefMatrix:
    def __init__(self, ref: Ref, handle):
        self.ref, self.h = ref, handle
        self.format = None

    def __del__(self):
        try:
            if self.h:
                self.ref.lib.ref_free(self.h)
                self.h = None
        except Exception:
            pass

    def info(self):
        r, c, n, f, s = (C.c_int32() for _ in range(5))
        self.ref.lib.ref_mm_info(self.h, C.byref(r), C.byref(c), C.byref(n), C.byref(f), C.byref(s))
        return dict(rows=r.value, columns=c.value, num_entries=n.value, field=f.value, symmetry=s.value)

    def entries(self):
        n = self.info()["num_entries"]
        i, j, a = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
        self.ref._check(self.ref.lib.ref_mm_entries(self.h, _ptr(i, _i32p), _ptr(j, _i32p), _ptr(a, _f64p)))
        return i, j, a

    def max_row_length(self):
        return int(self.ref.lib.ref_mm_max_row_length(self.h))

    def order_rcm(self):
        """find_new_order_RCM of the reference itself (new_order[old] = new)."""
        out = np.zeros(max(self.info()["rows"], 1), np.int32)
        self.ref._check(self.ref.lib.ref_mm_order_rcm(self.h, _ptr(out, _i32p)))
        return out[: self.info()["rows"]]

    def sort_row_major(self):
        self.ref._check(self.ref.lib.ref_mm_sort_row_major(self.h))
        return self

    def convert(self, fmt: str, arg: int = 0):
        """fmt in csr|coo|coo-atomic|ell|hybrid; arg = row_alignment (csr) or skip_padding."""
        self.ref._check(self.ref.lib.ref_convert(self.h, C.c_char_p(fmt.encode()), C.c_int32(arg)))
        self.format = fmt
        sizes = (C.c_int64 * 8)()
        nbytes = C.c_int64()
        self.ref._check(self.ref.lib.ref_sizes(self.h, sizes, C.byref(nbytes)))
        p0, p1, p2, p3 = (_i32p() for _ in range(4))
        v0, v1 = _f64p(), _f64p()
        self.ref._check(self.ref.lib.ref_arrays(self.h, C.byref(p0), C.byref(p1), C.byref(v0),
                                                C.byref(p2), C.byref(p3), C.byref(v1)))
        s = [int(v) for v in sizes]
        if fmt == "csr":
            return Csr(s[0], s[1], s[2], s[4], _take(p0, s[0] + 1, np.int32), _take(p1, s[3], np.int32),
                       _take(v0, s[3], np.float64), nbytes.value)
        if fmt in ("coo", "coo-atomic"):
            return Coo(s[0], s[1], s[2], _take(p0, s[2], np.int32), _take(p1, s[2], np.int32),
                       _take(v0, s[2], np.float64), nbytes.value)
        if fmt == "ell":
            slots = s[0] * s[3]
            return Ell(s[0], s[1], s[2], s[3], s[4], _take(p1, slots, np.int32),
                       _take(v0, slots, np.float64), nbytes.value)
        return Hyb(s[0], s[1], s[2], s[3], s[4], s[6], _take(p1, s[4], np.int32), _take(v0, s[4], np.float64),
                   s[5], _take(p2, s[5], np.int32), _take(p3, s[5], np.int32), _take(v1, s[5], np.float64),
                   size_reference=nbytes.value)

    def spmv(self, x, y=None, threads: int = 1):
        """One y += A x through the reference's own kernels, `threads` OpenMP threads."""
        info = self.info()
        x = _f64(x)
        y = np.zeros(info["rows"]) if y is None else _f64(y).copy()
        self.ref._check(self.ref.lib.ref_spmv(self.h, C.c_int(threads), _ptr(x, _f64p), _ptr(y, _f64p)))
        return y

    def time(self, threads: int, reps: int = 10, pin: bool = True):
        """profile_kernel_run protocol; returns per-run nanoseconds."""
        ns = np.zeros(reps)
        self.ref._check(self.ref.lib.ref_time(self.h, C.c_int(threads), C.c_int(int(pin)), C.c_int(reps),
                                              _ptr(ns, _f64p)))
        return ns

    def cache_trace(self, threads: int, cache_bytes: int, line_bytes: int = 64, warmup: bool = False,
                    page_bytes: int = 4096):
        """The reference's own cache trace of one cache shared by all threads (cache-trace.cpp:92-161):
        misses[t, d] = misses of thread t on lines whose page belongs to thread/NUMA domain d."""
        if not hasattr(self.ref.lib, "ref_cache_trace"):
            raise RefError("libspmvref.so was built without the cache-simulation sources")
        out = np.zeros((threads, threads), np.uint64)
        self.ref._check(self.ref.lib.ref_cache_trace(self.h, C.c_int(threads), C.c_int64(cache_bytes), C.c_int(line_bytes),
                                                     C.c_int(int(warmup)), C.c_int(page_bytes),
                                                     out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out.astype(np.int64)

    def csr_rows_per_thread(self, t, T):
        return int(self.ref.lib.ref_csr_rows_per_thread(self.h, t, T))

    def csr_nonzeros_per_thread(self, t, T):
        return int(self.ref.lib.ref_csr_nonzeros_per_thread(self.h, t, T))
