#!/usr/bin/env python
"""Multi-GPU correctness check of the row-partitioned executor (run under torchrun on the GPU box).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29600 \
        tools/dist_check.py [--grid 96] [--iters 6] [--scale 18]

Every plan (halo, all-gather) x every transport (NCCL, copy-engine pulls out of IPC-mapped peer buffers) on the 27-point
operator, and R-MAT row blocks of equal non-zeros as hybrid / COO / CSR with and without the column split: `iters`
steps of x <- alpha A x issued back to back (no host synchronisation), the concatenated result compared on every rank
with the same iteration done on ONE GPU with the full matrix, per-row tolerance of BASELINE.json accumulated over the
iterations.  Exact data (x_j = 1 + (j mod 7)/8, alpha = 1, two steps) must agree bit for bit on the stencil.
bench.py performs the one-step version of this before it times anything; this is the long form (P = 2, 4, 8).
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
from spmv_cache_trace_b200 import distributed as D  # noqa: E402


def gather(comm_t, x_local, starts, rank, world):
    n = int(starts[-1])
    full = torch.zeros(n, dtype=torch.float64, device="cuda")
    full[int(starts[rank]):int(starts[rank + 1])] = torch.from_numpy(x_local).cuda()
    dist.all_reduce(full)
    return full.cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=96)
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--scale", type=int, default=18)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    sp.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    comm = D.Comm.from_torch(dist, torch)
    failures = 0

    def report(what, ok, detail=""):
        nonlocal failures
        failures += 0 if ok else 1
        if rank == 0:
            print(f"{'ok  ' if ok else 'FAIL'} {what} {detail}", flush=True)

    # ---- 27-point stencil: plans x transports -------------------------------------------------------------------------
    n = args.grid
    N = n ** 3
    starts = D.partition_rows_ref(N, world)
    s, e = int(starts[rank]), int(starts[rank + 1])
    full = sp.generators.stencil(sp.STENCIL_3D27, n, n, n)
    full.set_option("beta0", 1)
    alpha = 1.0 / 52.0
    full.set_alpha(alpha)
    x0 = np.random.default_rng(7).uniform(-1, 1, N)
    ref, bound = x0.copy(), np.abs(x0)
    for _ in range(args.iters):
        ref = full * ref  # beta0 + alpha: y = alpha A x
    full.set_alpha(1.0)
    xp = 1.0 + (np.arange(N) % 7) / 8.0
    ref_exact = full * (full * xp)
    absA = 52.0  # row sums of |A| are at most 52
    tol = 1e-12 * (args.iters + 1) * np.abs(x0).max()
    for mode, peer, push in (("halo", False, False), ("halo", True, False), ("halo", True, True), ("allgather", False, False),
                             ("allgather", True, False)):
        local = sp.generators.stencil(sp.STENCIL_3D27, n, n, n, fmt=sp.CSR, row_begin=s, row_end=e)
        eng = D.DistributedSpMV(comm, local, starts, mode=mode, consume_local=True, peer_copy=peer, peer_push=push)
        how = "fused push" if push else "peer copy" if peer else "NCCL"
        if push:
            report(f"27-point {n}^3, halo push is in effect", eng.info["halo_push"] == 1)
        eng.set_x(x0[s:e])
        for _ in range(args.iters):
            eng.step(alpha)
        got = gather(comm, eng.get_x(), starts, rank, world)
        err = float(np.abs(got - ref).max())
        report(f"27-point {n}^3, {mode}, {how}, {args.iters} steps", err <= tol, f"max|err|={err:.3e}")
        eng.set_x(xp[s:e])
        eng.step(1.0)
        eng.step(1.0)
        got = gather(comm, eng.get_x(), starts, rank, world)
        report(f"27-point {n}^3, {mode}, {how}, exact data", bool(np.array_equal(got, ref_exact)))
        eng.destroy()
    del full

    # ---- R-MAT row blocks of equal non-zeros -----------------------------------------------------------------------------
    scale, ef, seed = args.scale, 32, 0x5EED0004
    n2 = 1 << scale
    A = sp.generators.rmat(scale, ef, seed)
    starts2 = sp.partition.rows_nnz(A, world)
    s2, e2 = int(starts2[rank]), int(starts2[rank + 1])
    H = A.convert(sp.HYB)
    H.set_option("beta0", 1)
    a2 = 1.0 / 64.0
    H.set_alpha(a2)
    x1 = np.random.default_rng(9).uniform(-1, 1, n2)
    r2 = x1.copy()
    for _ in range(args.iters):
        r2 = H * r2
    scale_ = max(float(np.abs(r2).max()), 1e-300)
    for fmt, name in ((sp.HYB, "hybrid"), (sp.COO, "coo"), (sp.CSR, "csr")):
        for split in (True, False):
            for peer in (False, True):
                eng = D.DistributedSpMV(comm, A.row_block(s2, e2), starts2, mode="allgather", fmt=fmt, column_split=split,
                                        overlap=False, consume_local=True, peer_copy=peer)
                eng.set_x(x1[s2:e2])
                for _ in range(args.iters):
                    eng.step(a2)
                got = gather(comm, eng.get_x(), starts2, rank, world)
                err = float(np.abs(got - r2).max())
                report(f"R-MAT 2^{scale} x {ef}, {name}, column_split={split}, {'peer copy' if peer else 'NCCL'}",
                       err <= 1e-9 * scale_, f"max|err|={err:.3e} (scale {scale_:.3e})")
                eng.destroy()
    comm.destroy()
    dist.destroy_process_group()
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
