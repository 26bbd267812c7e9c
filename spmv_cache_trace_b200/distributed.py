"""Row-partitioned multi-GPU SpMV: Python face of spmvb200_comm_* / spmvb200_dist_* (include/spmv_b200.h).

The executor itself lives below the C ABI (csrc/dist.cu) so that a C++ caller -- the reference's profile_kernel,
whose OpenMP threads map one to one onto ranks -- can drive N GPUs; this module is a thin binding plus a plain-Python
restatement of the bookkeeping (partition, exchange plan, row split) that the CPU tests run over gloo and compare
with the library's own arithmetic (spmvb200_exchange_plan).

The reference is a single-process OpenMP code; what it has is the PARTITION (every thread owns
ceil(rows/T) consecutive rows, matrix/csr-matrix.cpp:77-95) and a model of who owns which part of x
(aligned-allocator.hpp:156-211).  The executor turns that into a distributed iteration
x_{k+1} = A x_k on the GPUs of one NVSwitch domain:

  * rank p owns rows [s_p, s_{p+1}) of A (global column indices) and the matching slice of x and y;
    the partition is the reference rule (`partition.rows_ref`) or balanced non-zeros (`rows_nnz`);
  * x lives in two full-length device buffers used in ping-pong: iteration k reads X_k and writes its
    rows of A X_k straight into its slice of X_{k+1}, so there is no y -> x copy;
  * the exchange fills the other ranks' slices of X_k, either with one NCCL all-gather (in place) or,
    when a rank only needs a narrow band of columns (stencils: one grid plane from each neighbour),
    with NCCL send/recv of exactly those columns ("halo");
  * the local rows are split in three CSR blocks -- rows that reference columns below the rank's own
    slice, rows that reference only own columns (interior), rows that reference columns above --
    the interior block runs on the compute stream WHILE the exchange runs on the communication
    stream; the two boundary blocks run after the exchange's event.

  * the boundary blocks run on their own stream as soon as the halo has arrived, CONCURRENTLY with the interior
    block, and the next exchange starts as soon as the boundary rows are done.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _abi
from ._abi import f32p, f64p, i32p, i64p


# --------------------------------------------------------------------------------------------
# pure bookkeeping (no GPU needed)
# --------------------------------------------------------------------------------------------

def partition_rows_ref(rows: int, parts: int):
    """The reference rule, start_p = min(rows, p * ceil(rows/P)) (matrix/csr-matrix.cpp:77-83)."""
    rpt = (rows + parts - 1) // parts
    return np.minimum(rows, np.arange(parts + 1, dtype=np.int64) * rpt)


def owner_of(starts, col: int) -> int:
    """Rank whose slice of x holds column `col`."""
    return int(np.searchsorted(starts, col, side="right") - 1)


@dataclass
class ExchangePlan:
    """Who sends which columns of x to whom.  sends/recvs: lists of (peer, lo, hi), hi exclusive."""
    mode: str
    sends: list = field(default_factory=list)
    recvs: list = field(default_factory=list)
    recv_bytes: int = 0


def make_exchange_plan(starts, need, rank: int, mode: str = "auto", halo_fraction: float = 0.25) -> ExchangePlan:
    """`need[q] = (lo, hi)`: the column range rank q's rows reference (hi exclusive; lo >= hi: nothing).

    mode "allgather": every rank receives every other slice.  "halo": rank q receives
    need[q] minus its own slice, from the owners.  "auto": halo if every rank's remote need is at
    most `halo_fraction` of the vector, else allgather.
    """
    P = len(starts) - 1
    n = int(starts[-1])

    def remote(q):
        lo, hi = need[q]
        out = []
        if hi <= lo:
            return out
        for o in range(P):
            if o == q:
                continue
            a, b = max(lo, int(starts[o])), min(hi, int(starts[o + 1]))
            if b > a:
                out.append((o, a, b))
        return out

    if mode == "auto":
        worst = max((sum(b - a for _, a, b in remote(q)) for q in range(P)), default=0)
        mode = "halo" if worst <= halo_fraction * n else "allgather"
    plan = ExchangePlan(mode)
    if mode == "allgather":
        plan.recv_bytes = 8 * (n - int(starts[rank + 1] - starts[rank]))
        return plan
    plan.recvs = remote(rank)
    for q in range(P):
        if q != rank:
            plan.sends += [(q, a, b) for o, a, b in remote(q) if o == rank]
    plan.recv_bytes = 8 * sum(b - a for _, a, b in plan.recvs)
    return plan


def exchange(dist, x_full, starts, rank: int, plan: ExchangePlan):
    """Fill the remote parts of x_full this rank needs.  Works on CPU (gloo) and CUDA (nccl) tensors."""
    P = len(starts) - 1
    if P == 1:
        return
    if plan.mode == "allgather":
        sizes = np.diff(starts)
        mine = x_full[int(starts[rank]):int(starts[rank + 1])]
        if np.all(sizes == sizes[0]) and hasattr(dist, "all_gather_into_tensor") and x_full.is_cuda:
            dist.all_gather_into_tensor(x_full[: int(starts[-1])], mine)  # in place
        else:  # uneven slices (or gloo): one broadcast per owner
            for q in range(P):
                dist.broadcast(x_full[int(starts[q]):int(starts[q + 1])], src=q)
        return
    ops = []
    for peer, a, b in plan.sends:
        ops.append(dist.P2POp(dist.isend, x_full[a:b], peer))
    for peer, a, b in plan.recvs:
        ops.append(dist.P2POp(dist.irecv, x_full[a:b], peer))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def split_rows(lo_end: int, hi_begin: int, rows: int):
    """Row blocks (begin, end, needs_remote_x) of a rank: [0,lo_end) boundary, [lo_end,hi_begin) interior,
    [hi_begin,rows) boundary.  If the interior is empty the whole rank is one boundary block."""
    if lo_end >= hi_begin:
        return [(0, rows, True)]
    blocks = []
    if lo_end > 0:
        blocks.append((0, lo_end, True))
    blocks.append((lo_end, hi_begin, False))
    if hi_begin < rows:
        blocks.append((hi_begin, rows, True))
    return blocks


# --------------------------------------------------------------------------------------------
# the executor below the C ABI
# --------------------------------------------------------------------------------------------

EXCHANGE = {"auto": 0, "allgather": 1, "halo": 2}
EXCHANGE_NAMES = {v: k for k, v in EXCHANGE.items()}
CONSUME_LOCAL, COLUMN_SPLIT, NO_OVERLAP, PEER_COPY, PEER_PUSH = 1, 2, 4, 8, 16


def _check(rc: int) -> None:
    if rc != 0:
        from . import matrix_error
        raise matrix_error(_abi.lib().spmvb200_last_error().decode(), rc)


def exchange_plan(starts, need, rank: int, mode: str = "auto") -> ExchangePlan:
    """The library's own plan arithmetic (spmvb200_exchange_plan): must equal make_exchange_plan above."""
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    P = len(starts) - 1
    lo = np.ascontiguousarray([a for a, _ in need], dtype=np.int64)
    hi = np.ascontiguousarray([b for _, b in need], dtype=np.int64)
    cap = 4 * P + 4
    sends, recvs = np.zeros(3 * cap, np.int64), np.zeros(3 * cap, np.int64)
    chosen, ns, nr, rb = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
    _check(_abi.lib().spmvb200_exchange_plan(P, starts.ctypes.data_as(i64p), lo.ctypes.data_as(i64p), hi.ctypes.data_as(i64p),
                                             rank, EXCHANGE[mode], cap, C.byref(chosen), C.byref(ns),
                                             sends.ctypes.data_as(i64p), C.byref(nr), recvs.ctypes.data_as(i64p), C.byref(rb)))
    plan = ExchangePlan(EXCHANGE_NAMES[chosen.value])
    if plan.mode != "allgather":
        plan.sends = [tuple(int(v) for v in sends[3 * i:3 * i + 3]) for i in range(ns.value)]
        plan.recvs = [tuple(int(v) for v in recvs[3 * i:3 * i + 3]) for i in range(nr.value)]
    plan.recv_bytes = int(rb.value)
    return plan


class Comm:
    """spmvb200_comm_t: the ranks of a row-partitioned run."""

    def __init__(self, handle, keep=None):
        self._h = handle
        self._keep = keep

    @staticmethod
    def local(nranks: int, devices=None):
        """In-process communicator: returns one Comm per rank (several ranks may share a GPU)."""
        arr = (C.c_void_p * nranks)()
        dev = None if devices is None else (C.c_int * nranks)(*[int(v) for v in devices])
        _check(_abi.lib().spmvb200_comm_create_local(nranks, dev, arr))
        return [Comm(C.c_void_p(arr[r])) for r in range(nranks)]

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(_abi.lib().spmvb200_comm_unique_id(buf))
        return buf.raw

    @staticmethod
    def nccl(uid: bytes, rank: int, nranks: int) -> "Comm":
        h = C.c_void_p()
        buf = C.create_string_buffer(bytes(uid), 128)
        _check(_abi.lib().spmvb200_comm_create_nccl(buf, rank, nranks, C.byref(h)))
        return Comm(h)

    @staticmethod
    def from_torch(dist, torch) -> "Comm":
        """One process per GPU under torchrun: torch.distributed carries rank 0's id to the others (plumbing only)."""
        rank, world = dist.get_rank(), dist.get_world_size()
        uid = Comm.unique_id() if rank == 0 else bytes(128)
        t = torch.frombuffer(bytearray(uid), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        return Comm.nccl(bytes(t.cpu().numpy().tobytes()), rank, world)

    @property
    def rank(self):
        r, n, d = C.c_int(), C.c_int(), C.c_int()
        _check(_abi.lib().spmvb200_comm_rank(self._h, C.byref(r), C.byref(n), C.byref(d)))
        return r.value

    @property
    def nranks(self):
        r, n, d = C.c_int(), C.c_int(), C.c_int()
        _check(_abi.lib().spmvb200_comm_rank(self._h, C.byref(r), C.byref(n), C.byref(d)))
        return n.value

    def barrier(self):
        _check(_abi.lib().spmvb200_comm_barrier(self._h))

    def allreduce(self, value: float, op: str = "max") -> float:
        v = C.c_double(float(value))
        _check(_abi.lib().spmvb200_comm_allreduce(self._h, C.byref(v), {"max": 0, "sum": 1, "min": 2}[op]))
        return v.value

    def destroy(self):
        if getattr(self, "_h", None):
            _abi.lib().spmvb200_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class DistributedSpMV:
    """spmvb200_dist_t: iterated x <- alpha A x on this rank's row block (see include/spmv_b200.h)."""

    def __init__(self, comm: Comm, local, starts, mode: str = "auto", fmt: int = 0, column_split: bool = False,
                 overlap: bool = True, consume_local: bool = False, peer_copy: bool = False, peer_push: bool = False):
        self.comm = comm
        self.starts = np.ascontiguousarray(starts, dtype=np.int64)
        flags = ((CONSUME_LOCAL if consume_local else 0) | (COLUMN_SPLIT if column_split else 0) | (0 if overlap else NO_OVERLAP)
                 | (PEER_COPY if peer_copy else 0) | (PEER_PUSH if peer_push else 0))
        h = C.c_void_p()
        _check(_abi.lib().spmvb200_dist_create(comm._h, local._h, self.starts.ctypes.data_as(i64p), EXCHANGE[mode], int(fmt),
                                               flags, C.byref(h)))
        self._h = h
        self._local = local
        if consume_local:
            local._h = None  # the executor owns (or has destroyed) it
        r = comm.rank
        self.rank = r
        self.s, self.e = int(self.starts[r]), int(self.starts[r + 1])
        self.rows = self.e - self.s

    def destroy(self):
        if getattr(self, "_h", None):
            _abi.lib().spmvb200_dist_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    @property
    def info(self) -> dict:
        inf = _abi.DistInfo()
        _check(_abi.lib().spmvb200_dist_info(self._h, C.byref(inf)))
        d = {n: int(getattr(inf, n)) for n, _ in _abi.DistInfo._fields_}
        d["exchange"] = EXCHANGE_NAMES[d["exchange"]]
        return d

    def blocks(self):
        """[(row_begin, row_end, needs_remote_x, kernel_name)]"""
        from . import DeviceMatrix
        out = []
        for b in range(self.info["n_blocks"]):
            rb, re_, rem, mh = C.c_int64(), C.c_int64(), C.c_int32(), C.c_void_p()
            _check(_abi.lib().spmvb200_dist_block(self._h, b, C.byref(rb), C.byref(re_), C.byref(rem), C.byref(mh)))
            name = _abi.lib().spmvb200_kernel_name(mh).decode()
            out.append((rb.value, re_.value, bool(rem.value), name))
        return out

    def block_matrix(self, b: int):
        """Borrowed DeviceMatrix view of block b (do not destroy)."""
        from . import DeviceMatrix
        mh = C.c_void_p()
        _check(_abi.lib().spmvb200_dist_block(self._h, b, None, None, None, C.byref(mh)))
        m = DeviceMatrix(mh)
        m.destroy = lambda: None  # borrowed
        return m

    def set_x(self, x_local):
        x = np.ascontiguousarray(x_local, dtype=np.float64)
        assert x.shape[0] == self.rows
        _check(_abi.lib().spmvb200_dist_set_x(self._h, x.ctypes.data_as(f64p)))

    def get_x(self):
        x = np.empty(self.rows)
        _check(_abi.lib().spmvb200_dist_get_x(self._h, x.ctypes.data_as(f64p)))
        return x

    def x_device(self) -> int:
        p = C.c_void_p()
        _check(_abi.lib().spmvb200_dist_x_device(self._h, C.byref(p)))
        return p.value or 0

    def step(self, alpha: float = 1.0):
        _check(_abi.lib().spmvb200_dist_step(self._h, float(alpha)))

    def synchronize(self):
        _check(_abi.lib().spmvb200_dist_sync(self._h))

    def time(self, steps: int, warmup: int = 3, alpha: float = 1.0) -> float:
        """Device milliseconds of `steps` steps on this rank (NCCL communicators)."""
        return time_steps([self], steps, warmup, alpha)[0]

    def run_host(self, xs, ys, alpha: float = 1.0) -> float:
        """len(xs) independent products through pinned host slices, pipelined; returns device milliseconds."""
        n = len(xs)
        xp = (f64p * n)(*[x.ctypes.data_as(f64p) for x in xs])
        yp = (f64p * n)(*[y.ctypes.data_as(f64p) for y in ys])
        ms = C.c_float()
        _check(_abi.lib().spmvb200_dist_run_host(self._h, n, xp, yp, float(alpha), C.byref(ms)))
        return float(ms.value)


def time_steps(engines, steps: int, warmup: int = 3, alpha: float = 1.0):
    """spmvb200_dist_time over the executors this thread drives; returns per-rank device milliseconds."""
    n = len(engines)
    arr = (C.c_void_p * n)(*[e._h for e in engines])
    ms = np.zeros(n, np.float32)
    _check(_abi.lib().spmvb200_dist_time(arr, n, warmup, steps, float(alpha), ms.ctypes.data_as(f32p)))
    return [float(v) for v in ms]
