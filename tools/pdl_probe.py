#!/usr/bin/env python
"""Root-cause probe for wrong rows under programmatic dependent launch (PDL).

Round 1 saw wrong rows when the pipelined hybrid step of the row-partitioned mode was launched with the PDL attribute
 and worked around it.  Two explanations were open: (a) the x gathers use ld.global.nc, which PTX
defines only for data nobody writes during the kernel's lifetime -- and under PDL that lifetime overlaps the producer of
x; (b) an ordering hole between PDL launches and event edges to other streams.  This probe runs the two situations with
PDL forced ("pdl" = 2) and with PDL off, on the product library and on the experiment build whose gathers are ordinary
coherent loads (SPMVB200_LIB=.../libspmvb200_xcoherent.so, built with SPMVB200_VARIANT=xcoherent
SPMVB200_CFLAGS=-DSPMVB200_X_COHERENT):

  A  owned stream, x_(k+1) = A x_k / 8 with the two vectors swapped by bind_x / bind_y, no host synchronisation;
  B  the row-partitioned executor (in-process communicator, --devices), R-MAT hybrid pieces cut by columns, steps
     issued back to back.

    python tools/pdl_probe.py --pdl 2 [--devices 0,1]
Prints one JSON line: bad rows per repetition in A (per format) and B.
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import spmv_cache_trace_b200 as sp  # noqa: E402
from spmv_cache_trace_b200 import distributed as D  # noqa: E402


def probe_a(pdl, reps, steps):
    cudart = ctypes.CDLL("libcudart.so")
    out = {}
    n = 1024
    N = n * n
    x0 = np.random.default_rng(7).uniform(-1, 1, N)
    for name, fmt in (("csr", sp.CSR), ("ell", sp.ELL), ("coo", sp.COO)):
        A = sp.generators.stencil(sp.STENCIL_2D5, n, n, 1, fmt=fmt)
        A.prepare()
        bufs = [ctypes.c_void_p(), ctypes.c_void_p()]
        for b in bufs:
            assert cudart.cudaMalloc(ctypes.byref(b), ctypes.c_size_t(8 * (N + 16))) == 0
            assert cudart.cudaMemset(b, 0, ctypes.c_size_t(8 * (N + 16))) == 0
        A.set_alpha(0.125)
        A.set_option("beta0", 1)

        def iterate(sync_every_step, pdl_opt):
            A.set_option("pdl", pdl_opt)
            assert cudart.cudaMemcpy(bufs[0], x0.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(8 * N), 1) == 0
            A.sync()
            for k in range(steps):
                A.bind_x(bufs[k % 2].value)
                A.bind_y(bufs[(k + 1) % 2].value)
                A.spmv()
                if sync_every_step:
                    A.sync()
            A.sync()
            got = np.empty(N)
            assert cudart.cudaMemcpy(got.ctypes.data_as(ctypes.c_void_p), bufs[steps % 2], ctypes.c_size_t(8 * N), 2) == 0
            return got

        ref = iterate(True, 0)
        bad = []
        pdl_seen = 0
        for _ in range(reps):
            got = iterate(False, pdl)
            pdl_seen = max(pdl_seen, A.get_option("last_launch.pdl"))
            bad.append(int(np.count_nonzero(got != ref)))
        out[name] = {"bad_rows": bad, "launched_with_pdl": bool(pdl_seen)}
        del A
        for b in bufs:
            cudart.cudaFree(b)
    return out


def probe_b(pdl, reps, steps, devices, scale, ef):
    P = len(devices)
    n = 1 << scale
    sp.set_device(devices[0])
    full = sp.generators.rmat(scale, ef, 0x5EED0004)
    starts = sp.partition.rows_nnz(full, P)
    H = full.convert(sp.HYB)
    H.set_option("beta0", 1)
    alpha = 1.0 / 64.0
    H.set_alpha(alpha)
    x0 = np.random.default_rng(3).uniform(-1, 1, n)
    x = x0.copy()
    for _ in range(steps):  # reference: one GPU, host round trip per step
        x = H * x
    blocks_host = [full.row_block(int(starts[r]), int(starts[r + 1])).export() for r in range(P)]
    del H, full
    comms = D.Comm.local(P, devices)
    engines = []
    for r in range(P):
        sp.set_device(devices[r])
        b = blocks_host[r]
        local = sp.csr_matrix.Matrix(int(starts[r + 1] - starts[r]), n, len(b["column_index"]), 1, b["row_ptr"], b["column_index"], b["value"])
        engines.append(D.DistributedSpMV(comms[r], local, starts, mode="allgather", fmt=sp.HYB, column_split=True, consume_local=True))
    for eng in engines:
        for b in range(eng.info["n_blocks"]):
            eng.block_matrix(b).set_option("pdl", pdl)
    bad = []
    for _ in range(reps):
        for r, eng in enumerate(engines):
            eng.set_x(x0[starts[r]:starts[r + 1]])
        for _ in range(steps):
            for eng in engines:
                eng.step(alpha)
        got = np.concatenate([eng.get_x() for eng in engines])
        scale_ = np.maximum(np.abs(x), 1e-300)
        bad.append(int(np.count_nonzero(np.abs(got - x) > 1e-9 * np.abs(x).max())))
    launched = [eng.block_matrix(0).get_option("last_launch.pdl") for eng in engines]
    return {"bad_rows": bad, "launched_with_pdl": launched, "ranks": P, "devices": devices}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pdl", type=int, default=2)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--devices", default="0,0")
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--ef", type=int, default=32)
    args = ap.parse_args()
    devices = [int(v) for v in args.devices.split(",")]
    line = {"lib": os.path.basename(os.environ.get("SPMVB200_LIB", "libspmvb200.so")), "pdl": args.pdl,
            "A_owned_stream_ping_pong": probe_a(args.pdl, args.reps, args.steps),
            "B_row_partitioned_hybrid": probe_b(args.pdl, args.reps, args.steps, devices, args.scale, args.ef)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
