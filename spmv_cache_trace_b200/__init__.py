"""spmv_cache_trace_b200 -- B200-native SpMV engine behind the interface of jamtrott/spmv-cache-trace.

Python face of libspmvb200.so (C ABI in include/spmv_b200.h).  The names mirror the
reference's C++ namespaces so tests read like the reference's own gtest files:

    mm = matrix_market.fromStream(io.StringIO(text))          # matrix/matrix-market.cpp:530
    A  = csr_matrix.from_matrix_market(mm)                    # matrix/csr-matrix.cpp:187 (built on the GPU)
    y  = A * x                                                # operator* (csr-matrix.cpp:245-259), CUDA kernel
    csr_matrix.spmv(A, x, y)                                  # y += A x  (csr-matrix-spmv.cpp:148)

Everything that computes runs on the GPU through the C ABI; this package contains no
numerical fallback.  Errors surface as `matrix_error` (reference: matrix::matrix_error,
matrix/matrix-error.hpp:10) and, from the kernel objects, `kernel_error`
(kernels/kernel.hpp:11).
"""
from __future__ import annotations

import ctypes as C
import io
from typing import Optional

import numpy as np

from . import _abi
from ._abi import Info, f32p, f64p, i32p, i64p

__all__ = [
    "matrix_error", "kernel_error", "matrix_market", "csr_matrix", "coo_matrix", "ell_matrix",
    "hybrid_matrix", "DeviceMatrix", "MatrixMarket", "generators", "partition", "device_count",
    "launch_count", "CSR", "COO", "ELL", "HYB", "COO_SEGMENTED", "COO_ATOMIC",
]

CSR, COO, ELL, HYB = 0, 1, 2, 3
COO_SEGMENTED, COO_ATOMIC = 0, 1
STENCIL_2D5, STENCIL_3D7, STENCIL_3D27 = 0, 1, 2
FORMAT_NAMES = {CSR: "csr", COO: "coo", ELL: "ell", HYB: "hybrid"}


class matrix_error(RuntimeError):
    """matrix::matrix_error of the reference (matrix/matrix-error.hpp:10-15)."""

    def __init__(self, message: str, status: int = 0):
        super().__init__(message)
        self.status = status


class kernel_error(RuntimeError):
    """kernel_error of the reference (kernels/kernel.hpp:11-16)."""


def _check(rc: int) -> None:
    if rc != 0:
        raise matrix_error(_abi.lib().spmvb200_last_error().decode(), rc)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(t)


def device_count() -> int:
    n = C.c_int(0)
    rc = _abi.lib().spmvb200_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def set_device(device: int) -> None:
    _check(_abi.lib().spmvb200_set_device(int(device)))


def launch_count() -> int:
    """Number of kernels this library launched so far in this process."""
    return int(_abi.lib().spmvb200_launch_count())


def set_global_option(key: str, value: int) -> None:
    _check(_abi.lib().spmvb200_set_global_option(key.encode(), int(value)))


def device_props(device: int = 0) -> dict:
    name = C.create_string_buffer(256)
    sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    l2, mem = C.c_int64(), C.c_int64()
    _check(_abi.lib().spmvb200_device_props(device, name, 256, C.byref(sm), C.byref(l2), C.byref(mem),
                                            C.byref(maj), C.byref(mnr)))
    return dict(name=name.value.decode(), sm_count=sm.value, l2_bytes=l2.value, mem_bytes=mem.value,
                cc=(maj.value, mnr.value))


# ---------------------------------------------------------------------------
# Matrix Market (host)
# ---------------------------------------------------------------------------

class MatrixMarket:
    """matrix_market::Matrix (matrix/matrix-market.hpp:78-136): size, header facts and entries."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        try:
            if self._h:
                _abi.lib().spmvb200_mm_free(self._h)
                self._h = None
        except Exception:
            pass

    def _info(self):
        v = [C.c_int32() for _ in range(6)]
        _check(_abi.lib().spmvb200_mm_info(self._h, *[C.byref(t) for t in v]))
        return [t.value for t in v]

    rows = property(lambda self: self._info()[0])
    columns = property(lambda self: self._info()[1])
    num_entries = property(lambda self: self._info()[2])
    field = property(lambda self: self._info()[3])
    symmetry = property(lambda self: self._info()[4])
    format = property(lambda self: self._info()[5])

    def _entries(self):
        n = self.num_entries
        pi, pj, pa = i32p(), i32p(), f64p()
        _check(_abi.lib().spmvb200_mm_entries(self._h, C.byref(pi), C.byref(pj), C.byref(pa)))
        if n == 0:
            return np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0)
        return (np.ctypeslib.as_array(pi, (n,)).copy(), np.ctypeslib.as_array(pj, (n,)).copy(),
                np.ctypeslib.as_array(pa, (n,)).copy())

    def row_indices(self):
        return self._entries()[0]

    def column_indices(self):
        return self._entries()[1]

    def values_real(self):
        return self._entries()[2]

    def max_row_length(self) -> int:
        v = C.c_int32()
        _check(_abi.lib().spmvb200_mm_max_row_length(self._h, C.byref(v)))
        return v.value

    def row_lengths(self):
        out = np.zeros(max(self.rows, 1), np.int32)
        _check(_abi.lib().spmvb200_mm_row_lengths(self._h, _p(out, i32p)))
        return out[: self.rows]

    def permute(self, new_order):
        """Matrix::permute (matrix-market.cpp:309-333), in place: i, j <- new_order[i-1]+1, new_order[j-1]+1."""
        order = _i32(new_order)
        if order.size != self.rows:
            raise matrix_error("The dimension of the matrix doesn't match")
        _check(_abi.lib().spmvb200_mm_permute(self._h, _p(order, i32p)))


class matrix_market:
    """namespace matrix_market"""

    Matrix = MatrixMarket

    @staticmethod
    def fromStream(stream) -> MatrixMarket:
        text = stream.read() if hasattr(stream, "read") else stream
        if isinstance(text, str):
            text = text.encode()
        h = C.c_void_p()
        _check(_abi.lib().spmvb200_mm_parse(text, len(text), C.byref(h)))
        return MatrixMarket(h)

    @staticmethod
    def load_matrix(path: str, o=None, verbose: bool = False) -> MatrixMarket:
        if verbose and o is not None:
            o.write(f"Loading matrix from {path}\n")
        h = C.c_void_p()
        _check(_abi.lib().spmvb200_mm_load(str(path).encode(), C.byref(h)))
        return MatrixMarket(h)

    @staticmethod
    def from_entries(rows, columns, i, j, a) -> MatrixMarket:
        i, j, a = _i32(i), _i32(j), _f64(a)
        h = C.c_void_p()
        _check(_abi.lib().spmvb200_mm_from_entries(rows, columns, len(i), _p(i, i32p), _p(j, i32p), _p(a, f64p),
                                                   C.byref(h)))
        return MatrixMarket(h)

    @staticmethod
    def find_new_order_RCM(m: MatrixMarket):
        """matrix/matrix-market-reorder.cpp:60-170: new_order[old index] = new index."""
        out = np.zeros(max(m.rows, 1), np.int32)
        _check(_abi.lib().spmvb200_mm_order_rcm(m._h, _p(out, i32p)))
        return out[: m.rows]

    @staticmethod
    def find_new_order_GP(m: MatrixMarket, nparts: int, partitioner: bool = None):
        """find_new_order_GP (matrix-market-reorder.cpp:172-278).  partitioner=None: what the global option
        "mm.gp_partitioner" says (default: the identity, the reference without METIS); True: the library's own K-way
        partitioner in METIS's place."""
        out = np.zeros(max(m.rows, 1), np.int32)
        fn = _abi.lib().spmvb200_mm_order_gp_kway if partitioner else _abi.lib().spmvb200_mm_order_gp
        _check(fn(m._h, int(nparts), _p(out, i32p)))
        return out[: m.rows]

    @staticmethod
    def partition_kway(m: MatrixMarket, nparts: int, ub: float = 1.05):
        """What the reference calls METIS_PartGraphKway for (:236-237): (part[rows], edgecut)."""
        part = np.zeros(max(m.rows, 1), np.int32)
        cut = C.c_int64(0)
        _check(_abi.lib().spmvb200_mm_partition_kway(m._h, int(nparts), int(round(ub * 1000)), _p(part, i32p), C.byref(cut)))
        return part[: m.rows], int(cut.value)

    @staticmethod
    def order_from_parts(part, nparts: int):
        """The grouping step of find_new_order_GP (:246-266): new_order[old] = new."""
        part = _i32(part)
        out = np.zeros(max(part.size, 1), np.int32)
        _check(_abi.lib().spmvb200_order_from_parts(int(part.size), int(nparts), _p(part, i32p), _p(out, i32p)))
        return out[: part.size]

    @staticmethod
    def _copy(m: MatrixMarket) -> MatrixMarket:
        i, j, a = m._entries()
        return matrix_market.from_entries(m.rows, m.columns, i, j, a)

    @staticmethod
    def sort_matrix_row_major(m: MatrixMarket) -> MatrixMarket:
        out = matrix_market._copy(m)
        _check(_abi.lib().spmvb200_mm_sort_row_major(out._h))
        return out

    @staticmethod
    def sort_matrix_column_major(m: MatrixMarket) -> MatrixMarket:
        out = matrix_market._copy(m)
        _check(_abi.lib().spmvb200_mm_sort_column_major(out._h))
        return out


# ---------------------------------------------------------------------------
# Device matrices
# ---------------------------------------------------------------------------

class DeviceMatrix:
    """A matrix resident on the GPU together with the kernel object's x and y vectors."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        self.destroy()

    def destroy(self):
        try:
            if getattr(self, "_h", None):
                _abi.lib().spmvb200_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- facts ---------------------------------------------------------------
    @property
    def info(self) -> Info:
        inf = Info()
        _check(_abi.lib().spmvb200_matrix_info(self._h, C.byref(inf)))
        return inf

    rows = property(lambda self: int(self.info.rows))
    columns = property(lambda self: int(self.info.columns))
    num_entries = property(lambda self: int(self.info.num_entries))
    row_length = property(lambda self: int(self.info.ell_row_length))
    ell_row_length = property(lambda self: int(self.info.ell_row_length))
    num_ell_entries = property(lambda self: int(self.info.num_ell_entries))
    num_coo_entries = property(lambda self: int(self.info.num_coo_entries))
    format = property(lambda self: int(self.info.format))

    def size(self) -> int:
        """Matrix::size() as the reference prints it in "matrix_size"."""
        return int(self.info.matrix_size)

    def algorithmic_bytes(self) -> int:
        """matrix_size + x_size + y_size (reference kernels/csr-spmv.cpp:108-110)."""
        inf = self.info
        return int(inf.matrix_size + inf.x_size + inf.y_size)

    @property
    def kernel_name(self) -> str:
        return _abi.lib().spmvb200_kernel_name(self._h).decode()

    # -- vectors ---------------------------------------------------------------
    def set_x(self, x):
        x = _f64(x)
        if x.shape[0] != self.columns:
            raise matrix_error(f"Size mismatch: A.size()={self.rows}x{self.columns}, x.size()={x.shape[0]}")
        _check(_abi.lib().spmvb200_set_x(self._h, _p(x, f64p)))

    def set_y(self, y):
        y = _f64(y)
        if y.shape[0] != self.rows:
            raise matrix_error(f"Size mismatch: A.size()={self.rows}x{self.columns}, y.size()={y.shape[0]}")
        _check(_abi.lib().spmvb200_set_y(self._h, _p(y, f64p)))

    def get_y(self):
        y = np.empty(self.rows)
        if self.rows:
            _check(_abi.lib().spmvb200_get_y(self._h, _p(y, f64p)))
        return y

    def get_x(self):
        x = np.empty(self.columns)
        if self.columns:
            _check(_abi.lib().spmvb200_get_x(self._h, _p(x, f64p)))
        return x

    def fill_x(self, v: float):
        _check(_abi.lib().spmvb200_fill_x(self._h, float(v)))

    def fill_y(self, v: float):
        _check(_abi.lib().spmvb200_fill_y(self._h, float(v)))

    def x_device(self) -> int:
        p = C.c_void_p()
        _check(_abi.lib().spmvb200_x_device(self._h, C.byref(p)))
        return p.value or 0

    def y_device(self) -> int:
        p = C.c_void_p()
        _check(_abi.lib().spmvb200_y_device(self._h, C.byref(p)))
        return p.value or 0

    def bind_x(self, device_ptr: int):
        _check(_abi.lib().spmvb200_bind_x(self._h, C.c_void_p(device_ptr)))

    def bind_y(self, device_ptr: int):
        _check(_abi.lib().spmvb200_bind_y(self._h, C.c_void_p(device_ptr)))

    def set_stream(self, cuda_stream: int):
        _check(_abi.lib().spmvb200_set_stream(self._h, C.c_void_p(cuda_stream)))

    # -- the hot path ------------------------------------------------------------
    def set_alpha(self, alpha: float):
        """y += alpha*A*x from now on (1.0 = the reference's semantics, exact)."""
        _check(_abi.lib().spmvb200_set_alpha(self._h, float(alpha)))

    def prepare(self):
        """Build the selected kernel's launch metadata now (Kernel::prepare, kernels/kernel.hpp:28)."""
        _check(_abi.lib().spmvb200_prepare(self._h))

    def spmv(self):
        """y += A x on the device (asynchronous)."""
        _check(_abi.lib().spmvb200_spmv(self._h))

    def sync(self):
        _check(_abi.lib().spmvb200_sync(self._h))

    def spmv_host(self, x, y):
        """y += A x with host buffers: H2D x and y, kernel, D2H y.  y is updated in place."""
        if not (isinstance(y, np.ndarray) and y.dtype == np.float64 and y.flags.c_contiguous):
            raise matrix_error("y must be a contiguous float64 numpy array")
        x = _f64(x)
        if x.shape[0] != self.columns or y.shape[0] != self.rows:
            raise matrix_error(f"Size mismatch: A.size()={self.rows}x{self.columns}, x.size()={x.shape[0]}")
        _check(_abi.lib().spmvb200_spmv_host(self._h, _p(x, f64p), _p(y, f64p)))
        return y

    def time(self, reps: int = 10, warmup: int = 3):
        """Per-launch milliseconds (CUDA events on the launching stream)."""
        ms = np.zeros(reps, np.float32)
        _check(_abi.lib().spmvb200_time(self._h, warmup, reps, _p(ms, f32p)))
        return ms

    def set_option(self, key: str, value: int):
        _check(_abi.lib().spmvb200_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        v = C.c_int64()
        _check(_abi.lib().spmvb200_get_option(self._h, key.encode(), C.byref(v)))
        return v.value

    def __mul__(self, x):
        """operator*: y = A x on a fresh zero y (csr-matrix.cpp:245-259 and siblings)."""
        self.set_x(x)
        self.fill_y(0.0)
        self.spmv()
        return self.get_y()

    # -- export in the reference layout ---------------------------------------------
    def export(self) -> dict:
        inf = self.info
        L = _abi.lib()
        if inf.format == CSR:
            rp = np.zeros(inf.rows + 1, np.int64)
            col = np.zeros(max(inf.stored_entries, 1), np.int32)
            val = np.zeros(max(inf.stored_entries, 1))
            _check(L.spmvb200_csr_export(self._h, _p(rp, i64p), _p(col, i32p), _p(val, f64p)))
            return dict(row_ptr=rp, column_index=col[: inf.stored_entries], value=val[: inf.stored_entries])
        if inf.format == COO:
            n = inf.stored_entries
            row, col, val = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1))
            _check(L.spmvb200_coo_export(self._h, _p(row, i32p), _p(col, i32p), _p(val, f64p)))
            return dict(row_index=row[:n], column_index=col[:n], value=val[:n])
        if inf.format == ELL:
            n = inf.rows * inf.ell_row_length
            col, val = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1))
            _check(L.spmvb200_ell_export(self._h, _p(col, i32p), _p(val, f64p)))
            return dict(column_index=col[:n], value=val[:n], row_length=int(inf.ell_row_length))
        ne, nc = inf.num_ell_entries, inf.num_coo_entries
        ecol, eval_ = np.zeros(max(ne, 1), np.int32), np.zeros(max(ne, 1))
        crow, ccol, cval = np.zeros(max(nc, 1), np.int32), np.zeros(max(nc, 1), np.int32), np.zeros(max(nc, 1))
        _check(L.spmvb200_hyb_export(self._h, _p(ecol, i32p), _p(eval_, f64p), _p(crow, i32p), _p(ccol, i32p),
                                     _p(cval, f64p)))
        return dict(ell_column_index=ecol[:ne], ell_value=eval_[:ne], coo_row_index=crow[:nc],
                    coo_column_index=ccol[:nc], coo_value=cval[:nc], ell_row_length=int(inf.ell_row_length))

    def convert(self, fmt: int, arg: int = 0) -> "DeviceMatrix":
        h = C.c_void_p()
        _check(_abi.lib().spmvb200_convert(self._h, fmt, arg, C.byref(h)))
        return DeviceMatrix(h)

    def column_split(self, col_begin: int, col_end: int):
        """(inside, outside): the entries with column in [col_begin, col_end) and all the others, as two CSR matrices."""
        a, b = C.c_void_p(), C.c_void_p()
        _check(_abi.lib().spmvb200_csr_column_split(self._h, int(col_begin), int(col_end), C.byref(a), C.byref(b)))
        return DeviceMatrix(a), DeviceMatrix(b)

    def column_span(self, col_begin: int, col_end: int) -> dict:
        """Columns this (CSR) row block references and its rows that need no remote x (see spmv_b200.h)."""
        v = [C.c_int64() for _ in range(4)]
        _check(_abi.lib().spmvb200_csr_column_span(self._h, col_begin, col_end, *[C.byref(t) for t in v]))
        return dict(col_min=v[0].value, col_max=v[1].value, lo_end=v[2].value, hi_begin=v[3].value)

    def row_block(self, row_begin: int, row_end: int) -> "DeviceMatrix":
        h = C.c_void_p()
        _check(_abi.lib().spmvb200_csr_row_block(self._h, row_begin, row_end, C.byref(h)))
        return DeviceMatrix(h)

    # -- Kernel::print (kernels/csr-spmv.cpp:97-112) -----------------------------------
    def describe(self, matrix_path: str = "") -> dict:
        inf = self.info
        name = FORMAT_NAMES[inf.format]
        d = {"name": f"cuda-{name}-spmv", "matrix_path": matrix_path, "matrix_format": name,
             "rows": int(inf.rows), "columns": int(inf.columns), "nonzeros": int(inf.num_entries),
             "matrix_size": int(inf.matrix_size), "x_size": int(inf.x_size), "y_size": int(inf.y_size)}
        if inf.format == HYB:
            d.update(ell_row_length=int(inf.ell_row_length), num_ell_entries=int(inf.num_ell_entries),
                     num_coo_entries=int(inf.num_coo_entries))
        return d


def _from_mm(fn_name: str, mm: MatrixMarket, arg: int) -> DeviceMatrix:
    h = C.c_void_p()
    _check(getattr(_abi.lib(), fn_name)(mm._h, arg, C.byref(h)))
    return DeviceMatrix(h)


def _spmv(A: DeviceMatrix, x, y):
    """y += A x, y a numpy array updated in place (the reference's spmv signature)."""
    A.set_x(x)
    A.set_y(y)
    A.spmv()
    y[:] = A.get_y()
    return y


class csr_matrix:
    """namespace csr_matrix (matrix/csr-matrix.hpp)."""

    @staticmethod
    def from_matrix_market(m: MatrixMarket) -> DeviceMatrix:
        return _from_mm("spmvb200_csr_from_mm", m, 1)

    @staticmethod
    def from_matrix_market_row_aligned(m: MatrixMarket, row_alignment: int) -> DeviceMatrix:
        return _from_mm("spmvb200_csr_from_mm", m, row_alignment)

    @staticmethod
    def Matrix(rows, columns, num_entries, row_alignment, row_ptr, column_index, value) -> DeviceMatrix:
        rp, col, val = np.ascontiguousarray(row_ptr, np.int64), _i32(column_index), _f64(value)
        h = C.c_void_p()
        _check(_abi.lib().spmvb200_csr_create64(rows, columns, num_entries, _p(rp, i64p), _p(col, i32p),
                                                _p(val, f64p), C.byref(h)))
        return DeviceMatrix(h)

    spmv = staticmethod(_spmv)

    @staticmethod
    def spmv_regular_traffic(A: DeviceMatrix, x, y):
        """y_i += sum_k a_k: the matrix values streamed, no gather (csr-matrix.hpp:132, csr-matrix-spmv.cpp:35-47, 119-131)."""
        A.set_option("csr.probe", 1)
        try:
            return _spmv(A, x, y)
        finally:
            A.set_option("csr.probe", 0)

    @staticmethod
    def spmv_irregular_traffic(A: DeviceMatrix, x, y):
        """y_i += sum_k x[j_k]: the gather alone (csr-matrix.hpp:137, csr-matrix-spmv.cpp:49-61, 133-146)."""
        A.set_option("csr.probe", 2)
        try:
            return _spmv(A, x, y)
        finally:
            A.set_option("csr.probe", 0)

    @staticmethod
    def spmv_rows_per_thread(A: DeviceMatrix, thread: int, num_threads: int) -> int:
        s = partition.rows_ref(A.rows, num_threads)
        return int(s[thread + 1] - s[thread])

    @staticmethod
    def spmv_nonzeros_per_thread(A: DeviceMatrix, thread: int, num_threads: int) -> int:
        s = partition.rows_ref(A.rows, num_threads)
        rp = A.export()["row_ptr"]
        return int(rp[s[thread + 1]] - rp[s[thread]])


class coo_matrix:
    """namespace coo_matrix (matrix/coo-matrix.hpp)."""

    @staticmethod
    def from_matrix_market(m: MatrixMarket, mode: int = COO_SEGMENTED) -> DeviceMatrix:
        return _from_mm("spmvb200_coo_from_mm", m, mode)

    @staticmethod
    def Matrix(rows, columns, num_entries, row_index, column_index, value, mode: int = COO_SEGMENTED) -> DeviceMatrix:
        r, c, v = _i32(row_index), _i32(column_index), _f64(value)
        h = C.c_void_p()
        _check(_abi.lib().spmvb200_coo_create(rows, columns, len(r), _p(r, i32p), _p(c, i32p), _p(v, f64p), mode,
                                              C.byref(h)))
        return DeviceMatrix(h)

    @staticmethod
    def spmv(num_threads, A, x, y, workspace=None, chunk_size=0):
        # num_threads / workspace / chunk_size belong to the OpenMP algorithm (coo-matrix.cpp:313-335)
        return _spmv(A, x, y)

    @staticmethod
    def spmv_atomic(num_threads, A, x, y, chunk_size=0):
        return _spmv(A, x, y)


class ell_matrix:
    """namespace ell_matrix (matrix/ell-matrix.hpp)."""

    @staticmethod
    def from_matrix_market(m: MatrixMarket, skip_padding: bool = False) -> DeviceMatrix:
        return _from_mm("spmvb200_ell_from_mm", m, int(skip_padding))

    @staticmethod
    def Matrix(rows, columns, num_entries, row_length, column_index, value, skip_padding=False) -> DeviceMatrix:
        c, v = _i32(column_index), _f64(value)
        h = C.c_void_p()
        _check(_abi.lib().spmvb200_ell_create(rows, columns, num_entries, row_length, _p(c, i32p), _p(v, f64p),
                                              int(skip_padding), C.byref(h)))
        return DeviceMatrix(h)

    spmv = staticmethod(_spmv)


class hybrid_matrix:
    """namespace hybrid_matrix (matrix/hybrid-matrix.hpp)."""

    @staticmethod
    def from_matrix_market(m: MatrixMarket, skip_padding: bool = False, o=None, verbose: bool = False) -> DeviceMatrix:
        if verbose and o is not None:
            o.write("Converting matrix to hybrid format\n")
        return _from_mm("spmvb200_hyb_from_mm", m, int(skip_padding))

    @staticmethod
    def Matrix(rows, columns, num_entries, ell_row_length, num_ell_entries, ell_column_index, ell_value,
               ell_skip_padding, num_coo_entries, coo_row_index, coo_column_index, coo_value) -> DeviceMatrix:
        ec, ev = _i32(ell_column_index), _f64(ell_value)
        cr, cc, cv = _i32(coo_row_index), _i32(coo_column_index), _f64(coo_value)
        h = C.c_void_p()
        _check(_abi.lib().spmvb200_hyb_create(rows, columns, num_entries, ell_row_length, _p(ec, i32p), _p(ev, f64p),
                                              int(ell_skip_padding), num_coo_entries, _p(cr, i32p), _p(cc, i32p),
                                              _p(cv, f64p), C.byref(h)))
        return DeviceMatrix(h)

    @staticmethod
    def spmv(num_threads, A, x, y, workspace=None, chunk_size=0):
        return _spmv(A, x, y)


class generators:
    """Synthetic matrices of BASELINE.json, generated on the device."""

    @staticmethod
    def stencil(kind: int, nx: int, ny: int, nz: int = 1, fmt: int = CSR, row_begin: int = 0,
                row_end: int = 0) -> DeviceMatrix:
        h = C.c_void_p()
        _check(_abi.lib().spmvb200_gen_stencil(kind, nx, ny, nz, row_begin, row_end, fmt, C.byref(h)))
        return DeviceMatrix(h)

    @staticmethod
    def rmat(scale: int, edge_factor: int, seed: int, a=0.57, b=0.19, c=0.19, fmt: int = CSR,
             coo_mode: int = COO_SEGMENTED, row_begin: int = 0, row_end: int = 0) -> DeviceMatrix:
        h = C.c_void_p()
        _check(_abi.lib().spmvb200_gen_rmat(scale, edge_factor, C.c_uint64(seed), a, b, c, row_begin, row_end, fmt,
                                            coo_mode, C.byref(h)))
        return DeviceMatrix(h)


class PinnedBuffer:
    """float64 numpy view over page-locked host memory (spmvb200_host_alloc)."""

    def __init__(self, n: int):
        self._p = C.c_void_p()
        _check(_abi.lib().spmvb200_host_alloc(8 * max(n, 1), C.byref(self._p)))
        self.array = np.ctypeslib.as_array(C.cast(self._p, f64p), (max(n, 1),))[:n]

    def __del__(self):
        try:
            if self._p:
                _abi.lib().spmvb200_host_free(self._p)
                self._p = None
        except Exception:
            pass


def time_rotating(mats, steps: int, warmup: int = 3, per_launch: bool = False):
    """`steps` launches round-robin over copies of one workload (L2-cold); returns (total_ms, per_launch_ms|None)."""
    arr = (C.c_void_p * len(mats))(*[m._h for m in mats])
    total = C.c_float()
    per = np.zeros(steps, np.float32) if per_launch else None
    _check(_abi.lib().spmvb200_time_rotating(arr, len(mats), warmup, steps, C.byref(total),
                                             _p(per, f32p) if per_launch else None))
    return float(total.value), per


def time_copy(nbytes: int, copies: int = 8, reps: int = 200, warmup: int = 20):
    """Per-copy milliseconds of an isolated device-to-device cudaMemcpyAsync of `nbytes` (event pair per copy, L2-cold)."""
    ms = np.zeros(reps, np.float32)
    _check(_abi.lib().spmvb200_time_copy(int(nbytes), int(copies), int(warmup), int(reps), _p(ms, f32p)))
    return ms


def time_host_rotating(mats, xs, ys, steps: int, warmup: int = 1) -> float:
    """End-to-end steps with host buffers (H2D x, y; kernel; D2H y); returns total milliseconds."""
    n = len(mats)
    arr = (C.c_void_p * n)(*[m._h for m in mats])
    xp = (f64p * n)(*[_p(x, f64p) for x in xs])
    yp = (f64p * n)(*[_p(y, f64p) for y in ys])
    total = C.c_float()
    _check(_abi.lib().spmvb200_time_host_rotating(arr, n, xp, yp, warmup, steps, C.byref(total)))
    return float(total.value)


class partition:
    """Row partitions of the multi-GPU mode."""

    @staticmethod
    def rows_ref(rows: int, parts: int):
        """The reference rule (csr-matrix.cpp:77-83)."""
        out = np.zeros(parts + 1, np.int64)
        _check(_abi.lib().spmvb200_partition_rows_ref(rows, parts, _p(out, i64p)))
        return out

    @staticmethod
    def rows_nnz(A: DeviceMatrix, parts: int):
        """Balanced non-zeros: start_p = first row with row_ptr >= floor(p*nnz/P)."""
        out = np.zeros(parts + 1, np.int64)
        _check(_abi.lib().spmvb200_partition_rows_nnz(A._h, parts, _p(out, i64p)))
        return out


    @staticmethod
    def rows_weighted(A: DeviceMatrix, parts: int, row_weight: float):
        """Balanced cost, a row costing its entries + row_weight (in entries): spmvb200_partition_rows_weighted."""
        out = np.zeros(parts + 1, np.int64)
        _check(_abi.lib().spmvb200_partition_rows_weighted(A._h, parts, int(round(1024 * row_weight)), _p(out, i64p)))
        return out


class cache_model:
    """The reference's LRU cache model (cache-simulation/lru.cpp, cache-trace.cpp:92-161) over the SpMV
    reference string, with per-array attribution and an arbitrary partition (include/spmv_b200.h).
    Host side only.  Every function returns one dict per part."""

    FIELDS = [n for n, _ in _abi.CacheMisses._fields_]

    @staticmethod
    def _config(cache_bytes, line_bytes, parts, starts, shared, warmup, page_bytes, stream_bypass):
        keep = None
        cfg = _abi.CacheConfig(int(cache_bytes), int(line_bytes), int(parts), None, int(bool(shared)),
                               int(bool(warmup)), int(page_bytes), int(bool(stream_bypass)))
        if starts is not None:
            keep = np.ascontiguousarray(starts, dtype=np.int64)
            if keep.size != parts + 1:
                raise matrix_error("starts must have parts+1 entries")
            cfg.starts = _p(keep, i64p)
        return cfg, keep

    @staticmethod
    def _result(out):
        return [{n: int(getattr(o, n)) for n in cache_model.FIELDS} for o in out]

    @staticmethod
    def csr(rows, columns, row_ptr, column_index, cache_bytes, line_bytes=64, parts=1, starts=None, shared=True,
            warmup=False, page_bytes=0, stream_bypass=False):
        rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
        col = np.ascontiguousarray(column_index, dtype=np.int32)
        cfg, keep = cache_model._config(cache_bytes, line_bytes, parts, starts, shared, warmup, page_bytes, stream_bypass)
        out = (_abi.CacheMisses * parts)()
        _check(_abi.lib().spmvb200_cache_trace_csr(rows, columns, _p(rp, i64p), _p(col, i32p), C.byref(cfg), out))
        return cache_model._result(out)

    @staticmethod
    def ell(rows, columns, row_length, column_index_row_major, cache_bytes, line_bytes=64, parts=1, starts=None,
            shared=True, warmup=False, page_bytes=0, stream_bypass=False):
        col = np.ascontiguousarray(column_index_row_major, dtype=np.int32)
        cfg, keep = cache_model._config(cache_bytes, line_bytes, parts, starts, shared, warmup, page_bytes, stream_bypass)
        out = (_abi.CacheMisses * parts)()
        _check(_abi.lib().spmvb200_cache_trace_ell(rows, columns, row_length, _p(col, i32p), C.byref(cfg), out))
        return cache_model._result(out)

    @staticmethod
    def coo(rows, columns, row_index, column_index, cache_bytes, line_bytes=64, parts=1, starts=None, shared=True,
            warmup=False, page_bytes=0, stream_bypass=False):
        row = np.ascontiguousarray(row_index, dtype=np.int32)
        col = np.ascontiguousarray(column_index, dtype=np.int32)
        cfg, keep = cache_model._config(cache_bytes, line_bytes, parts, starts, shared, warmup, page_bytes, stream_bypass)
        out = (_abi.CacheMisses * parts)()
        _check(_abi.lib().spmvb200_cache_trace_coo(rows, columns, len(row), _p(row, i32p), _p(col, i32p), C.byref(cfg), out))
        return cache_model._result(out)

    @staticmethod
    def matrix(A: "DeviceMatrix", cache_bytes, line_bytes=64, parts=1, starts=None, shared=True, warmup=False,
               page_bytes=0, stream_bypass=False):
        """The model for a device matrix (its index arrays are copied back to the host)."""
        cfg, keep = cache_model._config(cache_bytes, line_bytes, parts, starts, shared, warmup, page_bytes, stream_bypass)
        out = (_abi.CacheMisses * parts)()
        _check(_abi.lib().spmvb200_cache_trace(A._h, C.byref(cfg), out))
        return cache_model._result(out)
