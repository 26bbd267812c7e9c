"""The row-partitioned mode on CPU: world_size 2 (and 3) over gloo.

What runs here is everything of spmv_cache_trace_b200/distributed.py except the CUDA kernel: the
partition, the exchange plan, the exchange itself (gloo broadcast / send-recv instead of NCCL), the
interior/boundary split and the ping-pong iteration -- with the oracle's CSR loop standing in for
the local kernel.  The result after a few iterations x <- A x must equal the single-process oracle.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.generators_ref import stencil_entries  # noqa: E402
from spmv_cache_trace_b200.distributed import (exchange, exchange_plan, make_exchange_plan, owner_of,  # noqa: E402
                                               partition_rows_ref, split_rows)


def csr_from_entries(n, i, j, a):
    order = np.lexsort((j, i))
    i, j, a = i[order] - 1, j[order] - 1, a[order]
    rp = np.zeros(n + 1, np.int64)
    np.add.at(rp, i + 1, 1)
    return np.cumsum(rp), j.astype(np.int64), a


def local_spmv(rp, col, val, x, b, e):
    y = np.zeros(e - b)
    for r in range(b, e):
        k0, k1 = rp[r], rp[r + 1]
        y[r - b] = np.dot(val[k0:k1], x[col[k0:k1]])
    return y


def column_span(rp, col, b, e, cb, ce):
    lo_end, hi_begin, cmin, cmax = 0, e - b, None, -1
    for r in range(b, e):
        c = col[rp[r]:rp[r + 1]]
        if len(c) == 0:
            continue
        cmin = c.min() if cmin is None else min(cmin, c.min())
        cmax = max(cmax, c.max())
        if c.min() < cb:
            lo_end = max(lo_end, r - b + 1)
        if c.max() >= ce:
            hi_begin = min(hi_begin, r - b)
    return dict(col_min=-1 if cmin is None else int(cmin), col_max=int(cmax), lo_end=lo_end, hi_begin=hi_begin)


def worker(rank, world, port, case, mode, starts, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, i, j, a = case
    rp, col, val = csr_from_entries(n, i, j, a)
    s, e = int(starts[rank]), int(starts[rank + 1])
    span = column_span(rp, col, s, e, s, e)
    mine = torch.tensor([span["col_min"], span["col_max"] + 1], dtype=torch.int64)
    allneed = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allneed, mine)
    need = [tuple(int(v) for v in t.tolist()) for t in allneed]
    plan = make_exchange_plan(starts, need, rank, mode)
    blocks = split_rows(span["lo_end"], span["hi_begin"], e - s)
    X = [torch.zeros(n, dtype=torch.float64) for _ in range(2)]
    x0 = np.random.default_rng(5).uniform(-1, 1, n)
    X[0][s:e] = torch.from_numpy(x0[s:e])
    for k in range(3):
        cur, nxt = X[k % 2], X[(k + 1) % 2]
        # interior rows use only the local slice: compute them BEFORE the exchange to prove it
        y = np.zeros(e - s)
        xin = cur.numpy().copy()
        xin[:s] = np.nan
        xin[e:] = np.nan
        for b, e2, remote in blocks:
            if not remote:
                y[b:e2] = local_spmv(rp, col, val, xin, s + b, s + e2)
        exchange(dist, cur, starts, rank, plan)
        xc = cur.numpy()
        for b, e2, remote in blocks:
            if remote:
                y[b:e2] = local_spmv(rp, col, val, xc, s + b, s + e2)
        assert not np.isnan(y).any()
        nxt[s:e] = torch.from_numpy(y)
    np.save(os.path.join(out_dir, f"y_{rank}.npy"), X[3 % 2][s:e].numpy())
    np.save(os.path.join(out_dir, f"plan_{rank}.npy"), np.array([plan.recv_bytes, len(blocks)]))
    dist.barrier()
    dist.destroy_process_group()


def run_case(tmp_path, world, case, mode, starts, port):
    mp.spawn(worker, args=(world, port, case, mode, starts, str(tmp_path)), nprocs=world, join=True)
    n, i, j, a = case
    rp, col, val = csr_from_entries(n, i, j, a)
    x = np.random.default_rng(5).uniform(-1, 1, n)
    for _ in range(3):
        x = local_spmv(rp, col, val, x, 0, n)
    got = np.concatenate([np.load(tmp_path / f"y_{r}.npy") for r in range(world)])
    np.testing.assert_allclose(got, x, rtol=0, atol=1e-12 * np.abs(x).max())
    return [np.load(tmp_path / f"plan_{r}.npy") for r in range(world)]


def stencil_case(kind, nx, ny, nz):
    i, j, a = stencil_entries(kind, nx, ny, nz)
    return nx * ny * nz, i.astype(np.int64), j.astype(np.int64), a / 32.0


@pytest.mark.parametrize("mode", ["allgather", "halo", "auto"])
def test_two_ranks_stencil(tmp_path, mode):
    case = stencil_case(2, 6, 5, 8)  # 27-point, 240 rows: z-slabs of 4 planes
    starts = partition_rows_ref(case[0], 2)
    plans = run_case(tmp_path, 2, case, mode, starts, 29511 + ["allgather", "halo", "auto"].index(mode))
    if mode != "allgather":  # one 6x5 plane from the neighbour
        assert [int(p[0]) for p in plans] == [8 * 30, 8 * 30]
    assert all(int(p[1]) == 2 for p in plans)  # interior + one boundary block per rank


def test_three_ranks_uneven_partition(tmp_path):
    case = stencil_case(0, 7, 9, 1)  # 2D 5-point, 63 rows, uneven split
    starts = np.array([0, 17, 40, 63], dtype=np.int64)
    plans = run_case(tmp_path, 3, case, "halo", starts, 29521)
    assert int(plans[1][1]) == 3  # middle rank: low boundary, interior, high boundary


def test_two_ranks_unstructured_no_interior(tmp_path):
    rng = np.random.default_rng(11)
    n = 60
    mask = rng.random((n, n)) < 0.15
    mask[np.arange(n), np.arange(n)] = True
    i, j = np.nonzero(mask)
    case = (n, (i + 1).astype(np.int64), (j + 1).astype(np.int64), rng.uniform(-0.1, 0.1, len(i)))
    plans = run_case(tmp_path, 2, case, "auto", partition_rows_ref(n, 2), 29531)
    assert all(int(p[1]) == 1 for p in plans)  # every row needs remote x: one block, no overlap


def column_split_worker(rank, world, port, case, starts, out_dir):
    """The overlap for matrices without a band: a rank's rows are cut by COLUMNS into the entries that reference
    its own slice of x (computed BEFORE the exchange, with every remote x poisoned) and the rest (added after)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, i, j, a = case
    rp, col, val = csr_from_entries(n, i, j, a)
    s, e = int(starts[rank]), int(starts[rank + 1])
    plan = make_exchange_plan(starts, [(0, n)] * world, rank, "allgather")
    own = (col >= s) & (col < e)
    v_in, v_out = np.where(own, val, 0.0), np.where(own, 0.0, val)
    c_in = np.where(own, col, s)  # the inside piece never indexes outside [s, e)
    X = [torch.zeros(n, dtype=torch.float64) for _ in range(2)]
    x0 = np.random.default_rng(5).uniform(-1, 1, n)
    X[0][s:e] = torch.from_numpy(x0[s:e])
    for k in range(3):
        cur, nxt = X[k % 2], X[(k + 1) % 2]
        xin = cur.numpy().copy()
        xin[:s] = np.nan
        xin[e:] = np.nan
        y = local_spmv(rp, c_in, v_in, xin, s, e)  # y = A_inside x, stored
        assert not np.isnan(y).any()
        exchange(dist, cur, starts, rank, plan)
        y += local_spmv(rp, col, v_out, cur.numpy(), s, e)  # y += A_outside x
        nxt[s:e] = torch.from_numpy(y)
    np.save(os.path.join(out_dir, f"y_{rank}.npy"), X[3 % 2][s:e].numpy())
    np.save(os.path.join(out_dir, f"plan_{rank}.npy"), np.array([plan.recv_bytes, int(own[rp[s]:rp[e]].sum())]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_column_split_overlap(tmp_path):
    rng = np.random.default_rng(13)
    n = 70
    mask = rng.random((n, n)) < 0.2
    mask[np.arange(n), np.arange(n)] = True
    i, j = np.nonzero(mask)
    case = (n, (i + 1).astype(np.int64), (j + 1).astype(np.int64), rng.uniform(-0.1, 0.1, len(i)))
    starts = np.array([0, 25, 70], dtype=np.int64)  # uneven, like a balanced-nnz cut of a power-law matrix
    mp.spawn(column_split_worker, args=(2, 29541, case, starts, str(tmp_path)), nprocs=2, join=True)
    rp, col, val = csr_from_entries(n, case[1], case[2], case[3])
    x = np.random.default_rng(5).uniform(-1, 1, n)
    for _ in range(3):
        x = local_spmv(rp, col, val, x, 0, n)
    got = np.concatenate([np.load(tmp_path / f"y_{r}.npy") for r in range(2)])
    np.testing.assert_allclose(got, x, rtol=0, atol=1e-12 * np.abs(x).max())
    plans = [np.load(tmp_path / f"plan_{r}.npy") for r in range(2)]
    assert [int(p[0]) for p in plans] == [8 * 45, 8 * 25]  # all-gather: everything but the own slice
    assert all(int(p[1]) > 0 for p in plans)  # both ranks had work to overlap with the exchange


def test_plan_arithmetic():
    starts = partition_rows_ref(100, 4)
    assert starts.tolist() == [0, 25, 50, 75, 100]
    assert [owner_of(starts, c) for c in (0, 24, 25, 99)] == [0, 0, 1, 3]
    need = [(0, 30), (20, 55), (45, 80), (70, 100)]
    p1 = make_exchange_plan(starts, need, 1, "halo")
    assert sorted(p1.recvs) == [(0, 20, 25), (2, 50, 55)]
    assert sorted(p1.sends) == [(0, 25, 30), (2, 45, 50)]
    assert p1.recv_bytes == 8 * 10
    assert make_exchange_plan(starts, need, 1, "auto").mode == "halo"
    assert make_exchange_plan(starts, [(0, 100)] * 4, 1, "auto").mode == "allgather"
    assert make_exchange_plan(starts, need, 2, "allgather").recv_bytes == 8 * 75
    assert split_rows(3, 20, 25) == [(0, 3, True), (3, 20, False), (20, 25, True)]
    assert split_rows(0, 20, 25) == [(0, 20, False), (20, 25, True)]
    assert split_rows(10, 10, 25) == [(0, 25, True)]


def test_library_plan_equals_the_python_restatement():
    """spmvb200_exchange_plan (the arithmetic the executor below the C ABI runs; host only, no device) against
    make_exchange_plan above, on random partitions and column ranges."""
    rng = np.random.default_rng(17)
    for _ in range(200):
        P = int(rng.integers(1, 9))
        n = int(rng.integers(P, 400))
        cuts = np.sort(rng.integers(0, n + 1, P - 1)) if P > 1 else np.zeros(0, dtype=np.int64)
        starts = np.concatenate([[0], cuts, [n]]).astype(np.int64)
        need = []
        for q in range(P):
            lo = int(rng.integers(0, n + 1))
            hi = int(rng.integers(0, n + 1))
            need.append((lo, hi))  # hi <= lo: the rank references nothing
        for mode in ("auto", "halo", "allgather"):
            for rank in range(P):
                a, b = exchange_plan(starts, need, rank, mode), make_exchange_plan(starts, need, rank, mode)
                assert a.mode == b.mode and a.recv_bytes == b.recv_bytes
                assert sorted(a.sends) == sorted(b.sends) and sorted(a.recvs) == sorted(b.recvs), (starts, need, rank, mode)
