#!/usr/bin/env python
"""Run one workload a few times (for ncu / quick A-B timing on the GPU box).

    python tools/run_workload.py c1_csr --steps 20 [--opt csr.tile=1024 --opt csr.stages=4] [--copies 4]

Extra reduced-size workloads for profiling: c5s_csr (27-point 256^3), c3s_coo (R-MAT 2^21 x 16),
c3s_coo_atomic, c4s_hyb (R-MAT 2^22 x 32).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import spmv_cache_trace_b200 as sp  # noqa: E402
from bench import make_workload, measured_peak  # noqa: E402


def factory(name):
    g = sp.generators
    extra = {
        "c5s_csr": lambda: g.stencil(sp.STENCIL_3D27, 256, 256, 256, sp.CSR),
        "c5s_ell": lambda: g.stencil(sp.STENCIL_3D27, 256, 256, 256, sp.ELL),
        "c3s_coo": lambda: g.rmat(21, 16, 0x5EED0003, fmt=sp.COO),
        "c3s_coo_atomic": lambda: g.rmat(21, 16, 0x5EED0003, fmt=sp.COO, coo_mode=sp.COO_ATOMIC),
        "c3s_csr": lambda: g.rmat(21, 16, 0x5EED0003, fmt=sp.CSR),
        "c3_csr": lambda: g.rmat(24, 16, 0x5EED0003, fmt=sp.CSR),
        "c4s_hyb": lambda: g.rmat(22, 32, 0x5EED0004, fmt=sp.HYB),
        # short-row stencils at a size where launch effects no longer matter (flat vs sliced + diagonal slices)
        "big7_csr": lambda: g.stencil(sp.STENCIL_3D7, 400, 400, 400, sp.CSR),
        "big5_csr": lambda: g.stencil(sp.STENCIL_2D5, 8000, 8000, 1, sp.CSR),
        "big7_ell": lambda: g.stencil(sp.STENCIL_3D7, 400, 400, 400, sp.ELL),
    }
    return extra[name] if name in extra else make_workload(sp, name)[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--copies", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--gopt", action="append", default=[], help="global option key=value, set before the matrices are built")
    ap.add_argument("--sweep", action="append", default=[], help="key=v1,v2,... (cartesian product of all sweeps)")
    args = ap.parse_args()
    for kv in args.gopt:
        sp.set_global_option(kv.split("=")[0], int(kv.split("=")[1]))
    make = factory(args.workload)
    A = make()
    B = A.algorithmic_bytes()
    l2 = sp.device_props(0)["l2_bytes"]
    copies = args.copies or (1 if B >= 3 * l2 else min(8, int(np.ceil(3.0 * l2 / B))))
    mats = [A] + [make() for _ in range(copies - 1)]
    peak, _ = measured_peak()
    base = dict(kv.split("=") for kv in args.opt)
    sweeps = [(s.split("=")[0], s.split("=")[1].split(",")) for s in args.sweep]

    def run(opts):
        for m in mats:
            for k, v in opts.items():
                m.set_option(k, int(v))
        try:
            total, per = sp.time_rotating(mats, args.steps, args.warmup, True)
        except sp.matrix_error as e:
            print(f"{args.workload} {opts}: ERROR {e}")
            return
        t = total / args.steps
        print(f"{args.workload} {opts} copies={copies} kernel={A.kernel_name} ms/step={t:.5f} "
              f"GB/s={B / t / 1e6:.0f} frac8T={B / t / 1e6 / 8000:.3f} fracMeas={B / t / 1e6 / peak:.3f} "
              f"kernel_ms(mean/min)={per.mean():.5f}/{per.min():.5f}", flush=True)

    def rec(i, opts):
        if i == len(sweeps):
            run(opts)
            return
        k, vals = sweeps[i]
        for v in vals:
            rec(i + 1, {**opts, k: v})

    rec(0, dict(base))


if __name__ == "__main__":
    main()
