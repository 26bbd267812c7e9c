#!/usr/bin/env python
"""Multi-GPU correctness check of the row-partitioned mode (run under torchrun on the GPU box).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600 \
        tools/dist_check.py [--grid 96] [--iters 3]

Every rank builds its rows of the 27-point operator, both exchange modes run `iters` iterations of
x <- A x / 32, and the concatenated result is compared on rank 0 with the same iteration done on a
single GPU with the full matrix (per-row tolerance of BASELINE.json, accumulated over the iterations).
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import spmv_cache_trace_b200 as sp  # noqa: E402
from spmv_cache_trace_b200.distributed import DistributedSpMV, partition_rows_ref  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=96)
    ap.add_argument("--iters", type=int, default=6)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    sp.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.grid
    N = n ** 3
    starts = partition_rows_ref(N, world)
    s, e = int(starts[rank]), int(starts[rank + 1])
    local = sp.generators.stencil(sp.STENCIL_3D27, n, n, n, fmt=sp.CSR, row_begin=s, row_end=e)
    x0 = np.random.default_rng(42).uniform(-1, 1, N)
    ok = True
    for mode in ("allgather", "halo", "auto"):
        for overlap in (True, False):
            eng = DistributedSpMV(sp, torch, dist, local, starts, rank, mode=mode, overlap=overlap)
            eng.set_x(x0[s:e])
            for _ in range(args.iters):  # issued back to back, no host synchronisation between the steps
                eng.step(scale=1.0 / 32.0)  # keep the iterates O(1)
            eng.synchronize()
            parts = [torch.zeros(int(starts[q + 1] - starts[q]), dtype=torch.float64, device="cuda") for q in range(world)]
            dist.all_gather(parts, eng.x_local().contiguous())
            got = torch.cat(parts).cpu().numpy()
            if rank == 0:
                full = sp.generators.stencil(sp.STENCIL_3D27, n, n, n, fmt=sp.CSR)
                x = x0.copy()
                bound = np.zeros(N)
                for _ in range(args.iters):
                    y = full * x
                    bound = 26.0 / 32.0 * bound + 52.0 / 32.0 * np.abs(x).max() * 1e-12 + 0.0
                    x = y / 32.0
                err = np.abs(got - x).max()
                lim = 1e-10 * max(1.0, np.abs(x).max())
                status = "ok" if err <= lim else "FAIL"
                ok = ok and err <= lim
                print(f"dist_check P={world} grid={n} mode={mode} plan={eng.plan.mode} overlap={overlap} "
                      f"blocks={[(b, e2, r) for _, b, e2, r in eng.blocks]} recv_bytes={eng.plan.recv_bytes} "
                      f"max|err|={err:.3e} {status}", flush=True)
                del full
            del eng
            dist.barrier()
    # BASELINE configs[3] at reduced scale: hybrid ELL+COO row blocks of an R-MAT matrix, equal non-zeros per rank,
    # x all-gathered (uneven slices), y = alpha*A*x through the kernels' alpha / beta0 path
    scale, ef, seed, alpha = 18, 32, 0x5EED0004, 1.0 / 64.0
    full = sp.generators.rmat(scale, ef, seed, fmt=sp.CSR)
    Nr = 1 << scale
    rstarts = sp.partition.rows_nnz(full, world)
    rs, re_ = int(rstarts[rank]), int(rstarts[rank + 1])
    for fmt, split in ((sp.HYB, False), (sp.HYB, True), (sp.COO, False), (sp.COO, True), (sp.CSR, True)):
        if split:  # the block cut by columns: own slice of x (overlaps the all-gather) / the rest (adds afterwards)
            block = full.row_block(rs, re_)
            eng = DistributedSpMV(sp, torch, dist, block, rstarts, rank, mode="auto", overlap=True, fmt=fmt, column_split=True)
        else:
            block = full.row_block(rs, re_).convert(fmt)
            eng = DistributedSpMV(sp, torch, dist, block, rstarts, rank, mode="auto", overlap=True)
        xr = np.random.default_rng(7).uniform(-1, 1, Nr)
        eng.set_x(xr[rs:re_])
        for _ in range(args.iters):
            eng.step(scale=alpha)
        eng.synchronize()
        parts = [torch.zeros(int(rstarts[q + 1] - rstarts[q]), dtype=torch.float64, device="cuda") for q in range(world)]
        dist.all_gather(parts, eng.x_local().contiguous())
        got = torch.cat(parts).cpu().numpy()
        if rank == 0:
            x = xr.copy()
            for _ in range(args.iters):
                x = alpha * (full * x)
            err = np.abs(got - x).max()
            lim = 1e-9 * max(1.0, np.abs(x).max())
            status = "ok" if err <= lim else "FAIL"
            ok = ok and err <= lim
            print(f"dist_check P={world} rmat 2^{scale}x{ef} format={sp.FORMAT_NAMES[fmt]} column_split={split} plan={eng.plan.mode} "
                  f"row_starts={[int(v) for v in rstarts]} kernels={[A.kernel_name for A, _, _, _ in eng.blocks]} "
                  f"max|err|={err:.3e} {status}", flush=True)
        del eng, block
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
