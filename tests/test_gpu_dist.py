"""GPU tests of the row-partitioned executor below the C ABI (spmvb200_comm_*, spmvb200_dist_*).

The in-process communicator lets several ranks share one GPU, so everything but NCCL itself -- the exchange plan, the
interior / boundary row split, the column split, the three-stream step with its event edges, the ping-pong buffers,
the pipelined host path -- runs on the single-GPU box of the driver, against the oracle.  (The NCCL backend differs
only in how a planned range travels; bench.py checks every multi-GPU run's result before it times it.)
"""
import os
import sys
import threading

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import spmv_cache_trace_b200 as sp
from oracle.generators_ref import rmat_entries, stencil_entries
from spmv_cache_trace_b200 import distributed as D

pytestmark = pytest.mark.gpu
TOL = 1e-12


def iterate_oracle(oracle, O, x0, steps, alpha):
    x, bound = x0.copy(), np.abs(x0)
    for _ in range(steps):
        bound = abs(alpha) * oracle.csr_abs_rowsum(O, bound)
        x = alpha * oracle.csr_spmv(O, x)
    return x, bound


def run_iteration(engines, x0, starts, steps, alpha):
    for r, eng in enumerate(engines):
        eng.set_x(x0[starts[r]:starts[r + 1]])
    for _ in range(steps):  # step k of every rank before step k+1 of any (the rule of in-process communicators)
        for eng in engines:
            eng.step(alpha)
    for eng in engines:
        eng.synchronize()
    return np.concatenate([eng.get_x() for eng in engines])


@pytest.mark.parametrize("P,mode", [(1, "auto"), (2, "halo"), (2, "allgather"), (3, "auto"), (4, "halo"), (4, "allgather")])
def test_stencil_iteration_matches_the_oracle(oracle, P, mode):
    nx, ny, nz = 12, 10, 4 * max(P, 2) + 3  # z-slabs of uneven thickness
    N = nx * ny * nz
    i, j, a = stencil_entries(2, nx, ny, nz)
    O = oracle.csr(N, N, i, j, a)
    starts = D.partition_rows_ref(N, P)
    comms = D.Comm.local(P, [0] * P)
    engines = []
    for r in range(P):
        local = sp.generators.stencil(sp.STENCIL_3D27, nx, ny, nz, fmt=sp.CSR, row_begin=int(starts[r]), row_end=int(starts[r + 1]))
        engines.append(D.DistributedSpMV(comms[r], local, starts, mode=mode, consume_local=True))
    x0 = np.random.default_rng(100 + P).uniform(-1, 1, N)
    steps, alpha = 7, 1.0 / 52.0
    got = run_iteration(engines, x0, starts, steps, alpha)
    ref, bound = iterate_oracle(oracle, O, x0, steps, alpha)
    assert np.all(np.abs(got - ref) <= TOL * (steps + 1) * bound)
    rp, cj = np.asarray(O.row_ptr, np.int64), np.asarray(O.column_index, np.int64)
    need = [(int(cj[rp[starts[q]]:rp[starts[q + 1]]].min()), int(cj[rp[starts[q]]:rp[starts[q + 1]]].max()) + 1) for q in range(P)]
    for r, eng in enumerate(engines):
        inf = eng.info
        assert inf["nranks"] == P and inf["rank"] == r and inf["steps_done"] == steps
        if P == 1:
            assert inf["n_blocks"] == 1 and inf["recv_bytes_per_step"] == 0
            continue
        # the plan the executor runs is the documented arithmetic (distributed.make_exchange_plan) on the column ranges
        # the rows really reference
        plan = D.make_exchange_plan(starts, need, r, mode)
        assert D.exchange_plan(starts, need, r, mode) == plan
        assert inf["exchange"] == plan.mode == ("halo" if mode == "auto" else mode)
        assert inf["recv_bytes_per_step"] == plan.recv_bytes
        neighbours = (r > 0) + (r < P - 1)
        if plan.mode == "halo":
            assert (inf["n_recvs"], inf["n_sends"]) == (len(plan.recvs), len(plan.sends)) == (neighbours, neighbours)
            assert 8 * neighbours * nx * ny <= plan.recv_bytes <= 8 * neighbours * (nx * ny + nx + 1)  # one grid plane (+ a line)
        blocks = eng.blocks()
        assert len(blocks) == 1 + neighbours  # interior + one boundary block per neighbour
        assert sum(1 for b in blocks if not b[2]) == 1 and inf["interior_rows"] > 0
    # exact data: the same iteration with x_j = 1 + (j mod 7)/8 and alpha = 1 is bit-identical to the oracle's
    xp = 1.0 + (np.arange(N) % 7) / 8.0
    got = run_iteration(engines, xp, starts, 2, 1.0)
    assert np.array_equal(got, oracle.csr_spmv(O, oracle.csr_spmv(O, xp)))


@pytest.mark.parametrize("P", [2, 3, 4])
def test_fused_halo_push(oracle, P):
    """SPMVB200_DIST_PEER_PUSH: the sliced CSR kernel that computes the rows a neighbour references stores them into the
    neighbour's x buffer as well, so from the second step on the exchange copies nothing.  Same numbers as the copying
    exchange, bit for bit on exact data, over many steps (every step consumes the halo the previous one pushed)."""
    nx, ny, nz = 12, 10, 4 * max(P, 2) + 3
    N = nx * ny * nz
    i, j, a = stencil_entries(2, nx, ny, nz)
    O = oracle.csr(N, N, i, j, a)
    starts = D.partition_rows_ref(N, P)

    def engines(push):
        comms = D.Comm.local(P, [0] * P)
        out = []
        for r in range(P):
            local = sp.generators.stencil(sp.STENCIL_3D27, nx, ny, nz, fmt=sp.CSR, row_begin=int(starts[r]), row_end=int(starts[r + 1]))
            out.append(D.DistributedSpMV(comms[r], local, starts, mode="halo", consume_local=True, peer_push=push))
        return out

    pushers, copiers = engines(True), engines(False)
    xp = 1.0 + (np.arange(N) % 7) / 8.0
    steps = 6
    got = run_iteration(pushers, xp, starts, steps, 1.0 / 4.0)  # powers of two: still exact
    want = run_iteration(copiers, xp, starts, steps, 1.0 / 4.0)
    assert all(e.info["halo_push"] == 1 for e in pushers) and all(e.info["halo_push"] == 0 for e in copiers)
    ref = xp
    for _ in range(steps):
        ref = 0.25 * oracle.csr_spmv(O, ref)
    assert np.array_equal(want, ref) and np.array_equal(got, ref)
    # set_x in between: the next exchange copies again, then the pushes take over
    x0 = np.random.default_rng(3).uniform(-1, 1, N)
    got = run_iteration(pushers, x0, starts, 5, 1.0 / 52.0)
    ref, bound = iterate_oracle(oracle, O, x0, 5, 1.0 / 52.0)
    assert np.all(np.abs(got - ref) <= TOL * 6 * bound)
    # a block that is not CSR cannot push: the flag is ignored, the copies stay
    comms = D.Comm.local(2, [0, 0])
    s2 = D.partition_rows_ref(N, 2)
    ell = [D.DistributedSpMV(comms[r], sp.generators.stencil(sp.STENCIL_3D27, nx, ny, nz, fmt=sp.CSR, row_begin=int(s2[r]), row_end=int(s2[r + 1])),
                             s2, mode="halo", fmt=sp.ELL, consume_local=True, peer_push=True) for r in range(2)]
    got = run_iteration(ell, xp, s2, 3, 0.25)
    assert all(e.info["halo_push"] == 0 for e in ell)
    ref = xp
    for _ in range(3):
        ref = 0.25 * oracle.csr_spmv(O, ref)
    assert np.array_equal(got, ref)


def test_all_ranks_run_one_plan(oracle):
    """A rank whose block allows no column analysis (here: an ELL block) needs the all-gather; the ranks that asked for
    the halo plan must follow, or the exchange would not match."""
    nx, ny, nz = 10, 9, 12
    N = nx * ny * nz
    i, j, a = stencil_entries(2, nx, ny, nz)
    O = oracle.csr(N, N, i, j, a)
    starts = D.partition_rows_ref(N, 3)
    comms = D.Comm.local(3, [0, 0, 0])
    engines = []
    for r in range(3):
        fmt = sp.ELL if r == 1 else sp.CSR
        local = sp.generators.stencil(sp.STENCIL_3D27, nx, ny, nz, fmt=fmt, row_begin=int(starts[r]), row_end=int(starts[r + 1]))
        engines.append(D.DistributedSpMV(comms[r], local, starts, mode="halo", consume_local=True))
    xp = 1.0 + (np.arange(N) % 7) / 8.0
    got = run_iteration(engines, xp, starts, 3, 0.5)
    ref = xp
    for _ in range(3):
        ref = 0.5 * oracle.csr_spmv(O, ref)
    assert np.array_equal(got, ref)
    assert [e.info["exchange"] for e in engines] == ["allgather"] * 3


def test_uneven_partition_and_unstructured_rows(oracle):
    """Balanced-nnz cut of a ragged matrix: no band, so no interior rows; automatic mode picks the all-gather."""
    rng = np.random.default_rng(21)
    n = 3000
    rows = np.repeat(np.arange(n), rng.integers(1, 40, n))
    cols = rng.integers(0, n, rows.size)
    keys = np.unique(rows.astype(np.int64) * n + cols)
    rows, cols = (keys // n).astype(np.int32), (keys % n).astype(np.int32)
    vals = rng.uniform(-0.05, 0.05, keys.size)
    O = oracle.csr(n, n, rows + 1, cols + 1, vals)
    full = sp.csr_matrix.from_matrix_market(sp.matrix_market.from_entries(n, n, rows + 1, cols + 1, vals))
    P = 3
    starts = sp.partition.rows_nnz(full, P)
    assert list(starts) == list(oracle.partition_rows_nnz(np.asarray(O.row_ptr, np.int64), P))
    comms = D.Comm.local(P, [0] * P)
    engines = [D.DistributedSpMV(comms[r], full.row_block(int(starts[r]), int(starts[r + 1])), starts, mode="auto", consume_local=True)
               for r in range(P)]
    x0 = rng.uniform(-1, 1, n)
    got = run_iteration(engines, x0, starts, 5, 1.0)
    ref, bound = iterate_oracle(oracle, O, x0, 5, 1.0)
    assert np.all(np.abs(got - ref) <= TOL * 6 * bound)
    assert all(e.info["exchange"] == "allgather" and e.info["n_blocks"] == 1 for e in engines)


@pytest.mark.parametrize("fmt", [sp.HYB, sp.COO, sp.ELL, sp.CSR])
def test_column_split_pieces_overlap_the_all_gather(oracle, fmt):
    """BASELINE configs[3] in small: R-MAT row blocks of equal non-zeros, each cut by columns into the entries that
    reference the rank's own slice of x (run during the exchange, stored) and the rest (added afterwards)."""
    scale, ef, seed = 13, 16, 0x5EED0004
    n = 1 << scale
    r, c, v = rmat_entries(scale, ef, seed)
    O = oracle.csr(n, n, r + 1, c + 1, v)
    full = sp.generators.rmat(scale, ef, seed)
    P = 2
    starts = sp.partition.rows_nnz(full, P)
    comms = D.Comm.local(P, [0] * P)
    engines = [D.DistributedSpMV(comms[q], full.row_block(int(starts[q]), int(starts[q + 1])), starts, mode="allgather", fmt=fmt,
                                 column_split=True, consume_local=True) for q in range(P)]
    for eng in engines:
        blocks = eng.blocks()
        assert [b[2] for b in blocks] == [False, True] and blocks[0][:2] == blocks[1][:2] == (0, eng.rows)
    x0 = np.random.default_rng(3).uniform(-1, 1, n)
    alpha = 1.0 / 64.0
    got = run_iteration(engines, x0, starts, 4, alpha)
    ref, bound = iterate_oracle(oracle, O, x0, 4, alpha)
    assert np.all(np.abs(got - ref) <= TOL * 10 * bound)
    # without the split: one block per rank after the exchange
    engines2 = [D.DistributedSpMV(comms2, full.row_block(int(starts[q]), int(starts[q + 1])), starts, mode="allgather", fmt=fmt,
                                  overlap=False, consume_local=True) for q, comms2 in enumerate(D.Comm.local(P, [0] * P))]
    got2 = run_iteration(engines2, x0, starts, 4, alpha)
    assert np.all(np.abs(got2 - ref) <= TOL * 10 * bound)


@pytest.mark.parametrize("P,mode", [(1, "auto"), (2, "halo"), (3, "allgather")])
def test_run_host_pipeline(oracle, P, mode):
    """spmvb200_dist_run_host: independent products through pinned host slices, uploads / downloads overlapped with the
    steps; one host thread per rank (they meet once per step)."""
    nx, ny, nz = 16, 12, 15
    N = nx * ny * nz
    i, j, a = stencil_entries(2, nx, ny, nz)
    O = oracle.csr(N, N, i, j, a)
    starts = D.partition_rows_ref(N, P)
    comms = D.Comm.local(P, [0] * P)
    engines = []
    for r in range(P):
        local = sp.generators.stencil(sp.STENCIL_3D27, nx, ny, nz, fmt=sp.CSR, row_begin=int(starts[r]), row_end=int(starts[r + 1]))
        engines.append(D.DistributedSpMV(comms[r], local, starts, mode=mode, consume_local=True))
    steps = 9
    rng = np.random.default_rng(5)
    X = [1.0 + rng.integers(0, 8, N) / 8.0 for _ in range(steps)]  # exact data
    xs = [[sp.PinnedBuffer(int(starts[r + 1] - starts[r])) for _ in range(steps)] for r in range(P)]
    ys = [[sp.PinnedBuffer(int(starts[r + 1] - starts[r])) for _ in range(steps)] for r in range(P)]
    for r in range(P):
        for k in range(steps):
            xs[r][k].array[:] = X[k][starts[r]:starts[r + 1]]
            ys[r][k].array[:] = np.nan
    errors = []

    def work(r):
        try:
            sp.set_device(0)
            engines[r].run_host([b.array for b in xs[r]], [b.array for b in ys[r]], 1.0)
        except Exception as ex:  # pragma: no cover
            errors.append(ex)

    threads = [threading.Thread(target=work, args=(r,)) for r in range(P)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors and not any(t.is_alive() for t in threads)
    for k in range(steps):
        got = np.concatenate([ys[r][k].array for r in range(P)])
        assert np.array_equal(got, oracle.csr_spmv(O, X[k])), f"step {k}"
    # the executor still iterates afterwards
    got = run_iteration(engines, X[0], starts, 2, 1.0)
    assert np.array_equal(got, oracle.csr_spmv(O, oracle.csr_spmv(O, X[0])))


def test_timing_entry_point_and_errors(oracle):
    nx = 24
    N = nx ** 3
    P = 2
    starts = D.partition_rows_ref(N, P)
    comms = D.Comm.local(P, [0] * P)
    locals_ = [sp.generators.stencil(sp.STENCIL_3D27, nx, nx, nx, fmt=sp.CSR, row_begin=int(starts[r]), row_end=int(starts[r + 1]))
               for r in range(P)]
    e0 = D.DistributedSpMV(comms[0], locals_[0], starts, mode="halo")
    with pytest.raises(sp.matrix_error, match="every rank"):
        e0.step(1.0)  # rank 1 has no executor yet: no plan
    with pytest.raises(sp.matrix_error, match="already has an executor"):
        D.DistributedSpMV(comms[0], locals_[0], starts, mode="halo")
    e1 = D.DistributedSpMV(comms[1], locals_[1], starts, mode="halo")
    for e, x in ((e0, 0.5), (e1, 0.25)):
        e.set_x(np.full(e.rows, x))
    before = sp.launch_count()
    ms = D.time_steps([e0, e1], steps=6, warmup=2, alpha=1.0 / 52.0)
    assert len(ms) == 2 and all(m > 0 for m in ms)
    assert sp.launch_count() - before == (6 + 2) * (e0.info["launches_per_step"] + e1.info["launches_per_step"])
    with pytest.raises(sp.matrix_error):
        D.time_steps([e0], steps=1)  # all ranks of an in-process communicator, or none
    wrong = sp.generators.stencil(sp.STENCIL_3D27, nx, nx, nx, fmt=sp.CSR, row_begin=0, row_end=100)
    with pytest.raises(sp.matrix_error, match="rows"):
        D.DistributedSpMV(D.Comm.local(1)[0], wrong, np.array([0, N], dtype=np.int64))
